"""TEST INFRASTRUCTURE ONLY -- ctypes loader for oracle/lj_oracle.c (the cut-off / large-N CPU restatement).
Imported by tests/, __graft_entry__.smoke() and bench.py's CPU legs only."""
import ctypes as C
import os
import subprocess
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "liboracle.so")
_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    src = os.path.join(HERE, "lj_oracle.c")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.run(["make", "-s", "-C", HERE], check=True)
    lib = C.CDLL(LIB)
    vp, i64, dbl = C.c_void_p, C.c_int64, C.c_double
    lib.oracle_num_threads.restype = C.c_int
    lib.oracle_cell_grid.restype = C.c_int
    lib.oracle_cell_grid.argtypes = [dbl, dbl]
    lib.oracle_cells.argtypes = [i64, vp, dbl, dbl, vp, vp, vp, C.POINTER(C.c_int)]
    lib.oracle_cells.restype = None
    for f in (lib.oracle_neighbors_brute, lib.oracle_neighbors_cells):
        f.restype = i64
        f.argtypes = [i64, vp, dbl, dbl, vp, vp, i64]
    lib.oracle_sys_create.restype = vp
    lib.oracle_sys_create.argtypes = [i64, dbl, dbl, dbl, dbl, dbl, dbl, dbl, vp, vp]
    lib.oracle_sys_destroy.argtypes = [vp]
    lib.oracle_sys_forces.argtypes = [vp, vp, C.POINTER(dbl)]
    lib.oracle_sys_step.argtypes = [vp, C.POINTER(dbl), C.POINTER(dbl)]
    lib.oracle_sys_get.argtypes = [vp, vp, vp, vp]
    lib.oracle_sys_counters.argtypes = [vp, vp]
    lib.oracle_sys_list.restype = i64
    lib.oracle_sys_list.argtypes = [vp, vp, vp]
    lib.oracle_set_threads.argtypes = [C.c_int]
    lib.oracle_sys_set_walls.argtypes = [vp, dbl, dbl, dbl]
    lib.oracle_sys_set_walls.restype = None
    _lib = lib
    return lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def cells(r, L, rlist):
    """Canonical binning: (cell_of[N], order[N], nc)."""
    lib = load()
    r = np.ascontiguousarray(r, dtype=np.float64)
    n = r.shape[1]
    cell_of = np.empty(n, dtype=np.int32)
    order = np.empty(n, dtype=np.int32)
    nc = C.c_int(0)
    lib.oracle_cells(n, _p(r), L, rlist, _p(cell_of), _p(order), None, C.byref(nc))
    return cell_of, order, nc.value


def neighbors(r, L, rlist, brute=False):
    """Canonical neighbour list as CSR (offsets[N+1] int64, idx int32), rows ascending."""
    lib = load()
    r = np.ascontiguousarray(r, dtype=np.float64)
    n = r.shape[1]
    fn = lib.oracle_neighbors_brute if brute else lib.oracle_neighbors_cells
    off = np.empty(n + 1, dtype=np.int64)
    total = fn(n, _p(r), L, rlist, _p(off), None, 0)
    idx = np.empty(max(total, 1), dtype=np.int32)
    fn(n, _p(r), L, rlist, _p(off), _p(idx), total)
    return off, idx[:total]


class CutoffSystem:
    """Verlet-list LJ system stepped by the C oracle (two force evaluations per step, like the reference)."""

    def __init__(self, r0, v0, L, rc, skin, dt, eps=1.0, sigma=1.0, mass=1.0, walls=None):
        """walls = (radius, restitution): hard walls at [radius, L - radius] (simulator.py:168-179).  Cells, lists and the
        minimum image then use the PERIOD L + 2 (rc + skin) of the padded grid the product embeds the box in (self.period)."""
        self.lib = load()
        r0 = np.ascontiguousarray(r0, dtype=np.float64)
        v0 = np.ascontiguousarray(v0, dtype=np.float64)
        assert r0.shape[0] == 3
        self.n = r0.shape[1]
        self.period = L if walls is None else L + 2.0 * (rc + skin)
        self.h = self.lib.oracle_sys_create(self.n, self.period, rc, skin, eps, sigma, mass, dt, _p(r0), _p(v0))
        if walls is not None:
            self.lib.oracle_sys_set_walls(self.h, L, float(walls[0]), float(walls[1]))

    def close(self):
        if self.h:
            self.lib.oracle_sys_destroy(self.h)
            self.h = None

    def __del__(self):
        self.close()

    def forces(self):
        acc = np.empty((3, self.n))
        pe = C.c_double(0.0)
        self.lib.oracle_sys_forces(self.h, _p(acc), C.byref(pe))
        return acc, pe.value

    def step(self):
        pe, ke = C.c_double(0.0), C.c_double(0.0)
        self.lib.oracle_sys_step(self.h, C.byref(pe), C.byref(ke))
        return pe.value, ke.value

    def run(self, steps, record=False):
        pe, ke, rr, vv = [], [], [], []
        for _ in range(steps):
            p, k = self.step()
            pe.append(p)
            ke.append(k)
            if record:
                r, v, _ = self.state()
                rr.append(r)
                vv.append(v)
        out = dict(pe=np.array(pe), ke=np.array(ke))
        if record:
            out.update(r=np.array(rr), v=np.array(vv))
        return out

    def state(self):
        r, v, a = (np.empty((3, self.n)) for _ in range(3))
        self.lib.oracle_sys_get(self.h, _p(r), _p(v), _p(a))
        return r, v, a

    def counters(self):
        c = np.zeros(5, dtype=np.int64)
        self.lib.oracle_sys_counters(self.h, _p(c))
        return dict(rebuilds=int(c[0]), pairs_unique=int(c[1]), force_evals=int(c[2]), list_checksum=int(c[3]),
                    nbr_entries=int(c[4]))

    def neighbor_list(self):
        off = np.empty(self.n + 1, dtype=np.int64)
        total = self.lib.oracle_sys_list(self.h, _p(off), None)
        idx = np.empty(max(total, 1), dtype=np.int32)
        self.lib.oracle_sys_list(self.h, _p(off), _p(idx))
        return off, idx[:total]


def bench_cutoff(wl, steps, warmup):
    """CPU timing leg for the cut-off workloads: all host threads (pthreads), a bounded number of MD steps."""
    import sys

    sys.path.insert(0, os.path.dirname(HERE))
    from bench import initial_state

    lib = load()
    n, L, r0, v0, dt = initial_state(wl["ndim"], wl["per_row"], wl["vf"])
    s = CutoffSystem(r0, v0, L, wl["r_cutoff"], wl["skin"], dt)
    for _ in range(max(warmup, 1)):
        s.step()
    c0 = s.counters()
    t0 = time.perf_counter()
    for _ in range(steps):
        s.step()
    t = time.perf_counter() - t0
    c1 = s.counters()
    s.close()
    return dict(pairs=c1["pairs_unique"] - c0["pairs_unique"], seconds=t, md_steps=steps, n=n, cores=lib.oracle_num_threads(),
                sample=f"{steps} verlet step(s) = {2 * steps} force evaluations of the same N={n} system, C/pthreads restatement "
                       f"(cell + Verlet list, plain truncation at rc={wl['r_cutoff']}) on {lib.oracle_num_threads()} threads")
