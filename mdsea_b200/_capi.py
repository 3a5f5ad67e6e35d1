"""ctypes binding of libmdsea_b200.so (include/mdsea_b200.h).  No fallback: if the library is missing it is
built in-tree with nvcc; if that fails, or no GPU is present at create() time, errors propagate."""
import ctypes as C
import os

import numpy as np

from mdsea_b200 import _build

ABI_VERSION = 2

STATUS = {0: "MDSEA_OK", -1: "MDSEA_ERR_BAD_ARG", -2: "MDSEA_ERR_CUDA", -3: "MDSEA_ERR_NCCL", -4: "MDSEA_ERR_STATE",
          -5: "MDSEA_ERR_NONFINITE", -6: "MDSEA_ERR_CAPACITY", -7: "MDSEA_ERR_UNSUPPORTED"}


class MdseaError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"{STATUS.get(status, status)}: {message}")
        self.status = status


class Params(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("ndim", C.c_int32), ("num_particles", C.c_int64),
        ("boxlen", C.c_double), ("mass", C.c_double), ("radius", C.c_double), ("dt", C.c_double),
        ("pbc", C.c_int32), ("algorithm", C.c_int32), ("restitution", C.c_double),
        ("pot_kind", C.c_int32), ("precision", C.c_int32),
        ("epsilon", C.c_double), ("sigma", C.c_double),
        ("mie_m", C.c_double), ("mie_n", C.c_double), ("mie_a", C.c_double),
        ("r_cutoff", C.c_double), ("skin", C.c_double),
        ("force_evals_per_step", C.c_int32), ("ring_rows", C.c_int32), ("device", C.c_int32),
        ("max_neighbors", C.c_int32), ("rank", C.c_int32), ("world", C.c_int32),
        ("proc_grid", C.c_int32 * 3), ("reserved", C.c_int32 * 8),
    ]


class RunArgs(C.Structure):
    _fields_ = [("nsteps", C.c_int32), ("record_coords", C.c_int32), ("k_boltzmann", C.c_double),
                ("gravity_acceleration", C.c_double), ("quench_targets", C.POINTER(C.c_double))]


class Rows(C.Structure):
    _fields_ = [("r", C.c_void_p), ("v", C.c_void_p), ("pe", C.c_void_p), ("ke", C.c_void_p), ("temp", C.c_void_p),
                ("ids", C.c_void_p), ("n_owned", C.c_void_p), ("row_stride", C.c_int64)]


class Stats(C.Structure):
    _fields_ = [("steps", C.c_int64), ("force_evals", C.c_int64), ("pair_evals_unique", C.c_int64),
                ("pair_evals_executed", C.c_int64), ("kernel_launches", C.c_int64), ("list_rebuilds", C.c_int64),
                ("n_local", C.c_int64), ("n_ghost", C.c_int64), ("nbr_entries", C.c_int64),
                ("ms_force", C.c_double), ("ms_integrate", C.c_double), ("ms_rebuild", C.c_double), ("ms_comm", C.c_double),
                ("tile_lists", C.c_int64), ("tile_window_max", C.c_int64), ("tile_fallbacks", C.c_int64),
                ("halo_transport", C.c_int64)]

    def as_dict(self):
        return {f: getattr(self, f) for f, _ in self._fields_}


EXPORTS = [
    "mdsea_b200_version", "mdsea_b200_create", "mdsea_b200_destroy", "mdsea_b200_last_error", "mdsea_b200_set_state",
    "mdsea_b200_get_state", "mdsea_b200_run", "mdsea_b200_run_async", "mdsea_b200_wait", "mdsea_b200_update_acc",
    "mdsea_b200_update_dists", "mdsea_b200_get_cells", "mdsea_b200_get_neighbors", "mdsea_b200_set_temperature",
    "mdsea_b200_get_stats", "mdsea_b200_first_nonfinite_step", "mdsea_b200_set_profiling", "mdsea_b200_stream",
    "mdsea_b200_nccl_unique_id", "mdsea_b200_comm_init", "mdsea_b200_comm_init_local", "mdsea_b200_rank_capacity",
    "mdsea_b200_measure_fma_peak", "mdsea_b200_analyse_speeds", "mdsea_b200_generate_state", "mdsea_b200_set_dt",
    "mdsea_b200_get_tile_windows", "mdsea_b200_apply_bc", "mdsea_b200_com_rog",
]

_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Load (building first if needed) the shared library; raises if it cannot be produced."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_build.LIB):
        _build.build()
    lib = C.CDLL(_build.LIB)
    vp, i32, i64p, dp = C.c_void_p, C.c_int32, C.POINTER(C.c_int64), C.POINTER(C.c_double)
    lib.mdsea_b200_version.restype = C.c_char_p
    lib.mdsea_b200_last_error.restype = C.c_char_p
    lib.mdsea_b200_last_error.argtypes = [vp]
    lib.mdsea_b200_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    lib.mdsea_b200_destroy.argtypes = [vp]
    lib.mdsea_b200_destroy.restype = None
    lib.mdsea_b200_set_state.argtypes = [vp, vp, vp]
    lib.mdsea_b200_get_state.argtypes = [vp, vp, vp, vp, vp, i64p]
    lib.mdsea_b200_run.argtypes = [vp, C.POINTER(RunArgs), C.POINTER(Rows)]
    lib.mdsea_b200_run_async.argtypes = [vp, C.POINTER(RunArgs), C.POINTER(Rows)]
    lib.mdsea_b200_wait.argtypes = [vp]
    lib.mdsea_b200_update_acc.argtypes = [vp, vp, dp]
    lib.mdsea_b200_update_dists.argtypes = [vp, vp, vp]
    lib.mdsea_b200_get_cells.argtypes = [vp, vp, vp, C.POINTER(i32 * 3)]
    lib.mdsea_b200_get_neighbors.argtypes = [vp, i64p, vp, vp]
    lib.mdsea_b200_generate_state.argtypes = [vp, vp, C.c_int64, C.c_double, C.c_uint64, i32, dp]
    lib.mdsea_b200_set_dt.argtypes = [vp, C.c_double]
    lib.mdsea_b200_apply_bc.argtypes = [vp]
    lib.mdsea_b200_com_rog.argtypes = [vp, vp, dp]
    lib.mdsea_b200_set_temperature.argtypes = [vp, C.c_double, i32]
    lib.mdsea_b200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.mdsea_b200_first_nonfinite_step.argtypes = [vp, i64p]
    lib.mdsea_b200_set_profiling.argtypes = [vp, i32]
    lib.mdsea_b200_stream.argtypes = [vp]
    lib.mdsea_b200_stream.restype = vp
    lib.mdsea_b200_nccl_unique_id.argtypes = [vp]
    lib.mdsea_b200_comm_init.argtypes = [vp, vp]
    lib.mdsea_b200_comm_init_local.argtypes = [vp, C.c_int64]
    lib.mdsea_b200_rank_capacity.argtypes = [vp]
    lib.mdsea_b200_rank_capacity.restype = C.c_int64
    lib.mdsea_b200_measure_fma_peak.argtypes = [i32, i32, dp]
    lib.mdsea_b200_analyse_speeds.argtypes = [i32, vp, C.c_int64, i32, C.c_int64, vp, dp, dp, dp]
    _lib = lib
    return lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


class Context:
    """Thin OO wrapper over one mdsea_ctx*; raises MdseaError on any non-zero status."""

    def __init__(self, **kw):
        self.lib = load()
        p = Params()
        p.abi_version = ABI_VERSION
        p.world = 1
        p.proc_grid[0] = p.proc_grid[1] = p.proc_grid[2] = 1
        p.force_evals_per_step = 2
        p.ring_rows = 1
        p.restitution = 1.0
        for k, v in kw.items():
            if k == "proc_grid":
                for d in range(3):
                    p.proc_grid[d] = int(v[d])
            else:
                setattr(p, k, v)
        self.params = p
        self.ndim, self.n = int(p.ndim), int(p.num_particles)
        self.world, self.rank = int(p.world), int(p.rank)
        h = C.c_void_p()
        st = self.lib.mdsea_b200_create(C.byref(p), C.byref(h))
        if st != 0:
            raise MdseaError(st, self.lib.mdsea_b200_last_error(None).decode())
        self.h = h

    def _ck(self, st):
        if st != 0:
            raise MdseaError(st, self.lib.mdsea_b200_last_error(self.h).decode())

    def close(self):
        if getattr(self, "h", None):
            self.lib.mdsea_b200_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- multi-rank plumbing ----------------------------------------------------------------------
    @staticmethod
    def nccl_unique_id() -> bytes:
        lib = load()
        buf = C.create_string_buffer(128)
        st = lib.mdsea_b200_nccl_unique_id(buf)
        if st != 0:
            raise MdseaError(st, lib.mdsea_b200_last_error(None).decode())
        return buf.raw

    def comm_init(self, unique_id: bytes):
        buf = C.create_string_buffer(unique_id, 128)
        self._ck(self.lib.mdsea_b200_comm_init(self.h, buf))

    def comm_init_local(self, group_key: int):
        self._ck(self.lib.mdsea_b200_comm_init_local(self.h, int(group_key)))

    def rank_capacity(self) -> int:
        return int(self.lib.mdsea_b200_rank_capacity(self.h))

    def set_state(self, r, v):
        r = np.ascontiguousarray(r, dtype=np.float64)
        v = np.ascontiguousarray(v, dtype=np.float64)
        assert r.shape == (self.ndim, self.n) and v.shape == r.shape
        self._ck(self.lib.mdsea_b200_set_state(self.h, _ptr(r), _ptr(v)))

    def generate_state(self, speed_cdf, speed_step, seed=0, isotropic=False) -> float:
        """Lattice + Maxwell-Boltzmann state generated on the device; returns the mean speed (input of get_dt)."""
        ms = C.c_double(0.0)
        if speed_cdf is None:
            self._ck(self.lib.mdsea_b200_generate_state(self.h, None, 0, 0.0, C.c_uint64(int(seed)), int(bool(isotropic)), C.byref(ms)))
        else:
            cdf = np.ascontiguousarray(speed_cdf, dtype=np.float64)
            self._ck(self.lib.mdsea_b200_generate_state(self.h, _ptr(cdf), cdf.shape[0], float(speed_step), C.c_uint64(int(seed)),
                                                        int(bool(isotropic)), C.byref(ms)))
        return float(ms.value)

    def set_dt(self, dt):
        self._ck(self.lib.mdsea_b200_set_dt(self.h, float(dt)))

    def apply_bc(self):
        self._ck(self.lib.mdsea_b200_apply_bc(self.h))

    def com_rog(self):
        com = np.empty(self.ndim)
        rog = C.c_double(0.0)
        self._ck(self.lib.mdsea_b200_com_rog(self.h, _ptr(com), C.byref(rog)))
        return com, float(rog.value)

    def set_temperature(self, temp, reset_ke=True):
        self._ck(self.lib.mdsea_b200_set_temperature(self.h, float(temp), 1 if reset_ke else 0))

    def get_state(self, want_acc=True):
        fill = np.empty if self.world == 1 else (lambda shp: np.full(shp, np.nan))
        r = fill((self.ndim, self.n)); v = fill((self.ndim, self.n)); a = fill((self.ndim, self.n)) if want_acc else None
        ids = np.empty(self.n, dtype=np.int64); n_owned = C.c_int64(0)
        self._ck(self.lib.mdsea_b200_get_state(self.h, _ptr(r), _ptr(v), _ptr(a) if want_acc else None, _ptr(ids), C.byref(n_owned)))
        return r, v, a, ids[:n_owned.value]

    def update_acc(self, want_pe=True):
        a = np.empty((self.ndim, self.n)) if self.world == 1 else np.full((self.ndim, self.n), np.nan)
        pe = C.c_double(0.0)
        self._ck(self.lib.mdsea_b200_update_acc(self.h, _ptr(a), C.byref(pe) if want_pe else None))
        return a, pe.value

    def update_dists(self):
        P = self.n * (self.n - 1) // 2
        d = np.empty(P); u = np.empty((P, self.ndim))
        self._ck(self.lib.mdsea_b200_update_dists(self.h, _ptr(d), _ptr(u)))
        return d, u

    def run(self, nsteps, kb, g=0.0, quench=None, record=True, rows=None, asynchronous=False):
        """Advance nsteps; returns dict of row arrays (views into `rows` buffers when given)."""
        E = self.ndim * self.n
        if rows is None:
            width = self.n if self.world == 1 else self.rank_capacity()
            rows = dict(r=np.empty((nsteps, self.ndim, width)) if record else None,
                        v=np.empty((nsteps, self.ndim, width)) if record else None,
                        pe=np.empty(nsteps), ke=np.empty(nsteps), temp=np.empty(nsteps))
            if self.world > 1:
                rows["ids"] = np.full((nsteps, width), -1, dtype=np.int32)
                rows["n_owned"] = np.zeros(nsteps, dtype=np.int64)
        a = RunArgs()
        a.nsteps, a.record_coords, a.k_boltzmann, a.gravity_acceleration = nsteps, 1 if record else 0, float(kb), float(g)
        qbuf = None
        if quench is not None:
            qbuf = np.ascontiguousarray(quench, dtype=np.float64)
            assert qbuf.shape == (nsteps, 2)
            a.quench_targets = qbuf.ctypes.data_as(C.POINTER(C.c_double))
        rw = Rows()
        rw.r = rows["r"].ctypes.data if (record and rows.get("r") is not None) else None
        rw.v = rows["v"].ctypes.data if (record and rows.get("v") is not None) else None
        rw.pe, rw.ke, rw.temp = rows["pe"].ctypes.data, rows["ke"].ctypes.data, rows["temp"].ctypes.data
        if self.world > 1 and record and rows.get("ids") is not None:
            rw.ids, rw.n_owned = rows["ids"].ctypes.data, rows["n_owned"].ctypes.data
            rw.row_stride = rows["r"].shape[2]
        fn = self.lib.mdsea_b200_run_async if asynchronous else self.lib.mdsea_b200_run
        self._ck(fn(self.h, C.byref(a), C.byref(rw)))
        # the library copies into `rows` asynchronously: keep the buffers of the batches in flight alive
        self._keep = (getattr(self, "_keep", ()) + ((qbuf, rows),))[-3:]
        del E
        return rows

    def wait(self):
        """Complete the OLDEST batch submitted with run(..., asynchronous=True); its row buffers are valid afterwards.
        Up to two batches can be in flight (the library alternates between two device ring sets)."""
        self._ck(self.lib.mdsea_b200_wait(self.h))

    def stats(self) -> dict:
        s = Stats()
        self._ck(self.lib.mdsea_b200_get_stats(self.h, C.byref(s)))
        return s.as_dict()

    def first_nonfinite_step(self) -> int:
        s = C.c_int64(-1)
        self._ck(self.lib.mdsea_b200_first_nonfinite_step(self.h, C.byref(s)))
        return s.value

    def set_profiling(self, on=True):
        self._ck(self.lib.mdsea_b200_set_profiling(self.h, 1 if on else 0))

    def stream(self) -> int:
        return self.lib.mdsea_b200_stream(self.h)

    def get_cells(self):
        cell_of = np.empty(self.n, dtype=np.int32); order = np.empty(self.n, dtype=np.int32); nc = (C.c_int32 * 3)()
        self._ck(self.lib.mdsea_b200_get_cells(self.h, _ptr(cell_of), _ptr(order), C.byref(nc)))
        return cell_of, order, tuple(nc)

    def get_neighbors(self):
        ne = C.c_int64(0)
        offsets = np.empty(self.n + 1, dtype=np.int64)
        self._ck(self.lib.mdsea_b200_get_neighbors(self.h, C.byref(ne), _ptr(offsets), None))
        idx = np.empty(max(ne.value, 1), dtype=np.int32)
        self._ck(self.lib.mdsea_b200_get_neighbors(self.h, C.byref(ne), _ptr(offsets), _ptr(idx)))
        return offsets, idx[:ne.value]


    def get_tile_windows(self):
        """(window entries, slot ranges) per block of 128 slots of the current tile lists; empty arrays if not in tile format"""
        nb = C.c_int64(0)
        self._ck(self.lib.mdsea_b200_get_tile_windows(self.h, C.byref(nb), None, None))
        nw = np.zeros(nb.value, dtype=np.int32)
        nr = np.zeros(nb.value, dtype=np.int32)
        if nb.value:
            self._ck(self.lib.mdsea_b200_get_tile_windows(self.h, C.byref(nb), _ptr(nw), _ptr(nr)))
        return nw, nr


def measure_fma_peak(device=0, fp64=False) -> float:
    lib = load()
    out = C.c_double(0.0)
    st = lib.mdsea_b200_measure_fma_peak(device, 1 if fp64 else 0, C.byref(out))
    if st != 0:
        raise MdseaError(st, lib.mdsea_b200_last_error(None).decode())
    return out.value


def analyse_speeds(v_coords: np.ndarray, device=0, want_speeds=True):
    """Device version of Analyser's speed statistics (mdsea/analytics.py:53-56): v_coords is the (STEPS, NDIM, N)
    dataset; returns (speeds (STEPS, N) or None, mean, std, max)."""
    lib = load()
    v = np.ascontiguousarray(v_coords, dtype=np.float64)
    steps, ndim, n = v.shape
    speeds = np.empty((steps, n)) if want_speeds else None
    mean, std, mx = C.c_double(0.0), C.c_double(0.0), C.c_double(0.0)
    st = lib.mdsea_b200_analyse_speeds(device, _ptr(v), steps, ndim, n, _ptr(speeds) if want_speeds else None, C.byref(mean),
                                       C.byref(std), C.byref(mx))
    if st != 0:
        raise MdseaError(st, lib.mdsea_b200_last_error(None).decode())
    return speeds, mean.value, std.value, mx.value
