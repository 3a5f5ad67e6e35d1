"""torchrun script (NOT collected by pytest): the NCCL transport of the domain decomposition on real GPUs.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tests/run_nccl_check.py

Every rank steps its share of a 16^3 LJ liquid; rank 0 compares the global energies and the assembled trajectory
with the C oracle (same contract as tests/test_gpu_multirank.py, which uses the in-process transport).
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from mdsea_b200 import _capi  # noqa: E402
from mdsea_b200.parallel import choose_proc_grid  # noqa: E402
from oracle import oracle_c  # noqa: E402
from test_gpu_cutoff import RC, SKIN, _liquid  # noqa: E402

GRIDS = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}


def main():
    import faulthandler

    faulthandler.dump_traceback_later(240, exit=True)   # a hung collective must not hold the box
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, L, r, v, dt = _liquid(16)
    steps = 40
    # evals = 2: the reference's two force evaluations per step; evals = 1: a(t+dt) reused, the finish kernel also starts
    # the next step (fused drift, tagged skin trigger riding on the halo messages); the last case adds a quench schedule
    # (one all-reduced mean_ke per step) and gravity, checked against the NumPy oracle with a cutoff
    for evals, quench in ((2, False), (1, False), (1, True)):
        ctx = _capi.Context(ndim=3, num_particles=n, boxlen=L, mass=1.0, radius=0.5, dt=dt, pbc=1, algorithm=0, pot_kind=1, precision=0,
                            epsilon=1.0, sigma=1.0, mie_m=12.0, mie_n=6.0, mie_a=0.0, r_cutoff=RC, skin=SKIN,
                            force_evals_per_step=evals, ring_rows=steps // 2, device=local, rank=rank, world=world,
                            proc_grid=GRIDS[world])
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")   # one NCCL communicator per context
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(_capi.Context.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()))
        ctx.set_temperature(1.0)
        ctx.set_state(r, v)
        q = None
        if quench:
            q = np.full((steps, 2), np.nan)
            q[7, 1], q[20, 1], q[33, 1] = 0.6, 0.3, 0.9     # step 20 = first step of the second batch
        rr = np.zeros((steps, 3, n)); owners = np.zeros((steps, n))
        pe, ke, temp = [], [], []
        half = steps // 2
        for b in range(2):   # two batches: the fused drift never crosses a batch boundary
            rows = ctx.run(half, kb=1.0, g=9.80665 if quench else 0.0, quench=None if q is None else q[b * half:(b + 1) * half])
            pe.append(rows["pe"].copy()); ke.append(rows["ke"].copy()); temp.append(rows["temp"].copy())
            for k in range(half):
                m = int(rows["n_owned"][k]); ids = rows["ids"][k, :m]
                rr[b * half + k][:, ids] = rows["r"][k, :, :m]; owners[b * half + k, ids] = 1
        pe, ke, temp = np.concatenate(pe), np.concatenate(ke), np.concatenate(temp)
        # gather the compact rows on rank 0 through torch.distributed (plumbing)
        t = torch.from_numpy(rr).cuda(); o = torch.from_numpy(owners).cuda()
        dist.all_reduce(t); dist.all_reduce(o)
        st = ctx.stats()
        if rank == 0:
            if quench:
                from oracle.oracle_np import MiePotential, OracleSolver

                ref = OracleSolver(r=r, v=v, boxlen=L, dt=dt, pot=MiePotential(), r_cutoff=RC, temp=1.0, kb=1.0, gravity=True,
                                   quench_steps=[7, 20, 33], quench_temps=[0.6, 0.3, 0.9]).run(steps)
                np.testing.assert_allclose(temp, ref["temp"], rtol=1e-10)
            else:
                s = oracle_c.CutoffSystem(r, v, L, RC, SKIN, dt)
                ref = s.run(steps, record=True)
                assert st["list_rebuilds"] == s.counters()["rebuilds"]
            np.testing.assert_allclose(pe, ref["pe"], rtol=1e-10)
            np.testing.assert_allclose(ke, ref["ke"], rtol=1e-10)
            assert bool((o == 1).all()), "every particle must be owned by exactly one rank"
            assert np.abs(t.cpu().numpy() - ref["r"]).max() <= 1e-9 * L
            print(f"NCCL domain decomposition check OK: world={world} evals={evals} quench={quench} rebuilds={st['list_rebuilds']} "
                  f"n_local={st['n_local']} n_ghost={st['n_ghost']}", flush=True)
        ctx.close()
    # fp32 pair math on a system large enough for interior tile blocks (tile lists inside the domain, per-particle lists on
    # its faces): energies to 1e-5, the neighbour lists in force at the end bit for bit, same rebuild schedule
    n, L, r, v, dt = _liquid(32, melt_steps=20)
    steps = 20
    ctx = _capi.Context(ndim=3, num_particles=n, boxlen=L, mass=1.0, radius=0.5, dt=dt, pbc=1, algorithm=0, pot_kind=1, precision=1,
                        epsilon=1.0, sigma=1.0, mie_m=12.0, mie_n=6.0, mie_a=0.0, r_cutoff=RC, skin=SKIN,
                        force_evals_per_step=1, ring_rows=steps, device=local, rank=rank, world=world,
                        proc_grid=choose_proc_grid(world, int(L // (RC + SKIN))))
    uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        uid.copy_(torch.frombuffer(bytearray(_capi.Context.nccl_unique_id()), dtype=torch.uint8))
    dist.broadcast(uid, src=0)
    ctx.comm_init(bytes(uid.cpu().numpy().tobytes()))
    ctx.set_temperature(1.0)
    ctx.set_state(r, v)
    rows = ctx.run(steps, kb=1.0)
    off, idx = ctx.get_neighbors()
    cnt = torch.from_numpy(np.diff(off)).cuda()
    dist.all_reduce(cnt)
    st = ctx.stats()
    s = oracle_c.CutoffSystem(r, v, L, RC, SKIN, dt)
    ref = s.run(steps)
    ref_off, ref_idx = s.neighbor_list()
    cell_of, _, _ = ctx.get_cells()
    mine = np.repeat(cell_of >= 0, np.diff(ref_off))
    np.testing.assert_array_equal(idx, ref_idx[mine])          # this rank's rows, bit for bit
    np.testing.assert_array_equal(cnt.cpu().numpy(), np.diff(ref_off))
    np.testing.assert_allclose(rows["pe"], ref["pe"], rtol=1e-5)
    np.testing.assert_allclose(rows["ke"], ref["ke"], rtol=1e-5)
    assert st["list_rebuilds"] == s.counters()["rebuilds"]
    tl = torch.tensor([st["tile_lists"]], device="cuda"); dist.all_reduce(tl)
    if rank == 0:
        print(f"NCCL domain decomposition check OK: world={world} fp32 tile lists on {int(tl.item())}/{world} ranks "
              f"transport={st['halo_transport']} (1 = peer memory, 2 = ncclSend/ncclRecv) rebuilds={st['list_rebuilds']}", flush=True)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
