set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','timesteps_per_s','gpu_launches')}, d['roofline']['ms_per_launch'], d['e2e']['ms_per_step'])"
