set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['roofline']['ms_per_launch'], d['roofline_hbm']['frac'])"
A="python profiles/ab_cutoff.py fp32 20"
$A > gpurun_out/plain_ab.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/l_fin.csv $A > gpurun_out/ncu_ab.log 2>&1
