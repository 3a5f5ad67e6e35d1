set -x
timeout 900 python -m pytest tests/test_gpu_allpairs.py tests/test_gpu_api.py -x -q 2>&1 | tail -4
python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','timesteps_per_s')}, d['roofline']['ms_per_launch'], d['roofline']['frac'])"
MDSEA_AP_N3=0 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','timesteps_per_s')}, d['roofline']['ms_per_launch'], d['roofline']['frac'])"
