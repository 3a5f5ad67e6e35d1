"""CPU, world_size 2 over gloo: the host-side logic of multi-GPU runs (mdsea_b200/parallel.py) -- process-grid choice,
unique-id hand-out, disjoint per-rank writes of compact dataset rows into ONE shared simfile, and the reassembly of
full host arrays from owned pieces.  (The device side of the decomposition is covered by tests/test_gpu_multirank.py.)"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, workdir, q):
    try:
        sys.path.insert(0, ROOT)
        os.chdir(workdir)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
        import torch.distributed as dist

        dist.init_process_group("gloo", rank=rank, world_size=world)
        from mdsea_b200 import parallel
        from mdsea_b200.core import SysManager

        assert parallel.dist_info() == (rank, world)
        uid = parallel.broadcast_unique_id(lambda: bytes(range(128)))
        assert uid == bytes(range(128))

        steps, ndim, n = 6, 3, 40
        kw = dict(simid="shared", ndim=ndim, num_particles=n, steps=steps, vol_fraction=0.3, radius_particle=0.5)
        if rank == 0:
            sm = SysManager.new(**kw)          # creates the tree and the datasets
            sm._open_datafile()
        parallel.barrier()
        if rank != 0:
            sm = SysManager.new(initfilesys=False, **kw)
            sm._open_datafile()
        # ground truth and a time-dependent ownership pattern (particles change owner between rows)
        rng = np.random.default_rng(5)
        R = rng.normal(size=(steps, ndim, n)); V = rng.normal(size=(steps, ndim, n))
        width = n  # rank capacity
        for first, nb in ((0, 4), (4, 2)):
            rows = dict(r=np.zeros((nb, ndim, width)), v=np.zeros((nb, ndim, width)), ids=np.full((nb, width), -1, np.int32),
                        n_owned=np.zeros(nb, np.int64), pe=np.arange(first, first + nb) * 1.0, ke=np.arange(first, first + nb) * 2.0,
                        temp=np.ones(nb))
            for k in range(nb):
                step = first + k
                owner = (np.arange(n) * 7 + step) % world           # deterministic, changes every step
                mine = np.where(owner == rank)[0][::-1]       # slot order is not id order
                m = len(mine)
                rows["ids"][k, :m] = mine
                rows["n_owned"][k] = m
                rows["r"][k, :, :m] = R[step][:, mine]
                rows["v"][k, :, :m] = V[step][:, mine]
            parallel.scatter_rows(sm, rows, nb, first, rank)
        sm.dfile.flush() if hasattr(sm.dfile, "flush") else None
        parallel.barrier()
        sm._close_datafile()
        parallel.barrier()
        sm2 = SysManager.load("shared")
        np.testing.assert_array_equal(sm2.get_ds("r_coords"), R)
        np.testing.assert_array_equal(sm2.get_ds("v_coords"), V)
        np.testing.assert_array_equal(sm2.get_ds("pot_energy"), np.arange(steps) * 1.0)
        np.testing.assert_array_equal(sm2.get_ds("kin_energy"), np.arange(steps) * 2.0)
        # owned pieces -> full arrays
        full = rng.normal(size=(ndim, n))
        box = [full if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        full = box[0]
        piece = np.where((np.arange(n) % world) == rank, full, np.nan)
        np.testing.assert_array_equal(parallel.assemble_owned(piece), full)
        dist.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as e:  # pragma: no cover
        import traceback

        q.put((rank, "FAIL: " + traceback.format_exc()))


def test_two_ranks_share_one_simfile(tmp_path):
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    [p.start() for p in procs]
    res = [q.get(timeout=180) for _ in procs]
    [p.join(timeout=60) for p in procs]
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_choose_proc_grid():
    from mdsea_b200.parallel import choose_proc_grid

    assert choose_proc_grid(1) == (1, 1, 1)
    # x (the contiguous cell index) stays whole; y and z share the ranks as evenly as possible
    assert choose_proc_grid(2) == (1, 1, 2)
    assert choose_proc_grid(4) == (1, 2, 2)
    assert choose_proc_grid(8) == (1, 2, 4)
    assert choose_proc_grid(6) == (1, 2, 3)
    assert choose_proc_grid(16) == (1, 4, 4)
    assert choose_proc_grid(8, ncells=43) == (1, 2, 4)
    assert choose_proc_grid(8, ncells=3) == (2, 2, 2)   # y and z alone cannot hold 8 ranks of >= 1 cell
    with pytest.raises(ValueError):
        choose_proc_grid(7, ncells=3)


def test_numa_binding_helpers_without_a_gpu(tmp_path):
    """bench.py binds every rank to the CPUs next to its GPU before it allocates pinned row buffers: the sysfs parsing and the
    'cannot tell' path (no device, no sysfs entry) must not raise or change the affinity."""
    import os

    from mdsea_b200.parallel import _parse_cpulist, bind_to_gpu_numa

    assert _parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert _parse_cpulist("") == set()
    before = os.sched_getaffinity(0)
    assert bind_to_gpu_numa(0, sysfs=str(tmp_path)) is None    # no CUDA device here / no such PCI function
    assert os.sched_getaffinity(0) == before
