for CS in 0 1 0 1; do MDSEA_FORCE_CS=$CS python profiles/ab_cutoff.py fp64 40 c3 100 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('f64 CS$CS', d['force_us'], d['rebuild_us'], d['integrate_us_per_step'], d['pe_last'])"; done 2>&1 | tee gpurun_out/ab_cs64.log
timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
