set -x
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/run_nccl_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -3
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_p2p.json 2> gpurun_out/bench_2gpu_p2p.err
for f in p2p; do python -c "import json,sys; d=json.loads(open('gpurun_out/bench_2gpu_$f.json').read().strip().splitlines()[-1]); print('$f', {k:d[k] for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'])"; done
MDSEA_TRACE_REBUILD=1 timeout 300 $TR --master-port 29515 profiles/mgpu_breakdown.py > gpurun_out/mgpu2.log 2>&1
grep "rank 0" gpurun_out/mgpu2.log | tail -6; tail -2 gpurun_out/mgpu2.log
