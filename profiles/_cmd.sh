MDSEA_NBR_BUILD=seg timeout 400 python -m pytest tests/test_gpu_cutoff.py -m gpu -x -q 2>&1 | tail -3
for B in warp seg warp seg; do MDSEA_NBR_BUILD=$B python profiles/ab_cutoff.py fp32 40 c3 200 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$B', d['rebuild_us'], d['force_us'], d['integrate_us_per_step'], d['rebuilds'], d['pe_last'])"; done 2>&1 | tee gpurun_out/ab_seg.log
