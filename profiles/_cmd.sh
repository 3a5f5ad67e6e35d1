set -x
timeout 600 python -m pytest tests/test_gpu_multirank.py -x -q 2>&1 | tail -3
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 300 $TR --master-port 29511 tests/run_nccl_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -2
MDSEA_HALO=nccl timeout 300 $TR --master-port 29512 tests/run_nccl_check.py 2>&1 | grep -v "^\*\|OMP_NUM\|^$" | tail -2
timeout 300 $TR --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_2gpu_p2p.json 2> gpurun_out/bench_2gpu_p2p.err
python -c "import json,sys; d=json.loads(open('gpurun_out/bench_2gpu_p2p.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step')}, d['e2e']['ms_per_step'])"
