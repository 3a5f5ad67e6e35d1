"""torchrun script: per-family CUDA-event times of the decomposed cut-off path (profiling mode: every kernel family is
bracketed by events and synchronised, so the sum is NOT the step time; it shows where the device time goes)."""
import os, sys, json
import torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import WORKLOADS, PROC_GRIDS, initial_state, make_ctx
from mdsea_b200 import _capi

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
wl = WORKLOADS["c3"]
n, L, r0, v0, dt = initial_state(3, wl["per_row_by_world"][world], wl["vf"])
ctx = make_ctx(wl, n, L, dt, local, "fp32", 20, rank=rank, world=world, proc_grid=PROC_GRIDS[world])
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(_capi.Context.nccl_unique_id()), dtype=torch.uint8))
dist.broadcast(uid, src=0)
ctx.comm_init(bytes(uid.cpu().numpy().tobytes()))
ctx.set_temperature(1.0); ctx.set_state(r0, v0)
for _ in range(3):
    ctx.run(20, kb=1.0, record=False)
torch.cuda.synchronize(); dist.barrier()
import time
t0 = time.perf_counter(); ctx.run(20, kb=1.0, record=False); torch.cuda.synchronize(); t_plain = time.perf_counter() - t0
ctx.set_profiling(True); q0 = ctx.stats()
ctx.run(20, kb=1.0, record=False)
q1 = ctx.stats()
out = {k: round(q1[k] - q0[k], 3) for k in ("ms_force", "ms_integrate", "ms_rebuild", "ms_comm", "list_rebuilds", "kernel_launches")}
out.update(rank=rank, n_local=q1["n_local"], n_ghost=q1["n_ghost"], plain_ms_20_steps=round(1e3 * t_plain, 3))
print(json.dumps(out), flush=True)
ctx.close(); dist.destroy_process_group()
