timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err
python bench.py --steps 10 --warmup 3 --evals 2 --no-cpu-baseline --no-extra > gpurun_out/bench_r01_c3_evals2.json 2>/dev/null
python -c "
import json
for f in ('bench_r01_final','bench_r01_c3_evals2'):
    d=json.loads(open('gpurun_out/'+f+'.json').read().strip().splitlines()[-1]); print(f, d.get('value'), d.get('ms_per_step'), d['e2e']['value'], d['roofline']['frac'], d['roofline_hbm']['frac'], d['gpu_launches'])
"
