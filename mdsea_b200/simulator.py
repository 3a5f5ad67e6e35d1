"""ContinuousPotentialSolver -- mdsea's solver API over the B200 step loop.

Drop-in for ``mdsea.simulator.ContinuousPotentialSolver`` (mdsea/simulator.py:19-49, 511-589): same
dataclass fields, same public attributes (``r_vec v_vec acc dists drunits mean_ke mean_pe temp step dt``),
same methods (``run_simulation advance update_acc update_dists update_energies update_files``), same
initial conditions (gen.py under the caller's ``np.random.seed``), same dataset rows.  What changed is
where the work happens: the state lives on the GPU, every step of ``advance`` runs in the CUDA kernels
of ``csrc/`` through the C ABI (include/mdsea_b200.h), rows are buffered in a device ring and flushed
to the HDF5 datasets every ``flush_every`` steps.  There is no NumPy fallback for the step loop.

New keyword fields (defaults keep the reference's behaviour):
    precision             "fp64" (default) | "fp32"  -- arithmetic of the pair kernel; state is always float64
    r_cutoff              None => all pairs (the reference's only working mode); a number => plain truncation
                          at dist < r_cutoff through the cell-list / Verlet-list kernels (3-D boxes: periodic with
                          >= 3 cells per side, or hard walls; 2-D boxes on one GPU, run as the 3-D system (x, L/2, y);
                          otherwise the all-pairs kernel applies the cutoff)
    skin                  Verlet skin of the neighbour list (default 0.3 sigma)
    flush_every           rows buffered on the device between HDF5 flushes (default: sized from memory)
    force_evals_per_step  1 (default, periodic boxes): a(t+dt) is reused as the next step's a(t);
                          2: evaluate twice per step exactly like simulator.py:540,544
    device                CUDA device ordinal
"""
from __future__ import annotations

import logging
import math
import os
from dataclasses import dataclass, field
from typing import Optional

import numpy as np

from mdsea_b200 import _capi, parallel, quicker
from mdsea_b200.config import Config
from mdsea_b200.constants import DTYPE
from mdsea_b200.core import SysManager
from mdsea_b200.gen import PosGen, VelGen
from mdsea_b200.helpers import ProgressBar, get_dt
from mdsea_b200.potentials import Potential

log = logging.getLogger(__name__)

_ALGORITHMS = {"verlet": 0, "simple": 1}
_PRECISIONS = {"fp64": 0, "fp32": 1}


def _pinned_empty(shape, keep: list):
    """Page-locked float64 host buffer (torch is used for allocation only); plain numpy if CUDA is absent.  The owning
    tensor goes into `keep` (a list on the solver instance): the pages are unpinned when the solver drops it."""
    try:
        import torch

        if torch.cuda.is_available():
            t = torch.empty(tuple(shape), dtype=torch.float64, pin_memory=True)
            a = t.numpy()
            a.setflags(write=True)
            keep.append(t)
            return a
    except Exception:  # pragma: no cover
        pass
    return np.empty(shape, dtype=np.float64)


@dataclass
class _BaseSimulator:
    sm: SysManager = field()
    dt: Optional[float] = field(default=None)
    pot: Potential = field(default_factory=Potential.ideal)
    restitution_coeff: float = field(default=1.0)
    isothermal: bool = field(default=False)
    gravity: bool = field(default=False)
    r_cutoff: Optional[float] = field(default=None)
    # --- new, B200-side knobs -------------------------------------------------------------------
    precision: str = field(default="fp64")
    skin: Optional[float] = field(default=None)
    flush_every: Optional[int] = field(default=None)
    force_evals_per_step: Optional[int] = field(default=None)
    device: int = field(default=0)
    max_neighbors: int = field(default=0)
    # initial conditions generated on the device (opt-in: same lattice bit for bit and same Maxwell-Boltzmann sampling
    # scheme as gen.py, but a counter-based generator instead of the global NumPy stream -- not for parity runs)
    device_init: bool = field(default=False)
    device_init_seed: int = field(default=0)
    device_init_isotropic: bool = field(default=False)

    def __post_init__(self) -> None:
        sm = self.sm
        if self.device_init:
            self._r = self._v = None   # filled from the device on first access
        else:
            # initial state: same generators, same RNG call order as simulator.py:54-59
            self._r = PosGen(ndim=sm.NDIM, boxlen=sm.LEN_BOX, nparticles=sm.NUM_PARTICLES).simplecubic()
            self._v = VelGen(ndim=sm.NDIM, nparticles=sm.NUM_PARTICLES).mb(sm.MASS, sm.TEMP, Config.k_boltzmann)
        # device_init runs are the large ones (configs[4]: 64 M particles on 8 ranks): no O(N) host mirrors until asked for
        self._acc = None if self.device_init else np.zeros((sm.NDIM, sm.NUM_PARTICLES), dtype=DTYPE)
        self.ndnp_zeroes = None if self.device_init else np.zeros((sm.NDIM, sm.NUM_PARTICLES), dtype=DTYPE)
        self._dists = self._drunits = self.pairs_indexes = None
        self.mean_ke = None
        self.mean_pe = None
        self.temp_init = sm.TEMP
        self.temp = sm.TEMP
        self.step = 0
        if self.dt is None and not self.device_init:  # simulator.py:86-87
            self.dt = get_dt(radius=sm.RADIUS_PARTICLE, mean_speed=float(quicker.norm(self._v, axis=0).mean()))
        self.apply_restcoeff = bool(self.restitution_coeff < 1)
        self.pbarr = ProgressBar("Simulator", sm.STEPS, __name__)
        self._FLIPID = quicker.flipid(sm.NDIM)

        if self.precision not in _PRECISIONS:
            raise ValueError(f"precision '{self.precision}' not in {tuple(_PRECISIONS)}")
        self._ctx = None
        self._embed = None           # 2-D system run as the 3-D system (x, L/2, y) on the cut-off path (_ensure_ctx)
        self._nd_dev = sm.NDIM
        self._pinned = []            # page-locked row buffers of this solver (released by close())
        self._device_fresh = False   # device holds a newer state than the host copies
        self._host_exposed = True    # host arrays may have been edited -> upload before the next device op
        if self.device_init:
            self._generate_on_device()

    def _generate_on_device(self) -> None:
        """PosGen.simplecubic + VelGen.mb + get_dt (simulator.py:54-59,86-87) on the device; needs the GPU at construction."""
        from mdsea_b200.gen import MONTECARLO_SPEEDRANGE, mb_cdf

        sm = self.sm
        dt_given = self.dt
        if self.dt is None:
            self.dt = 0.0   # placeholder until the mean speed is known
        ctx = self._ensure_ctx()
        cdf = None if sm.TEMP == 0 else mb_cdf(sm.MASS, sm.TEMP, Config.k_boltzmann)
        step = float(MONTECARLO_SPEEDRANGE[1] - MONTECARLO_SPEEDRANGE[0])
        mean_speed = ctx.generate_state(cdf, step, seed=self.device_init_seed, isotropic=self.device_init_isotropic)
        if dt_given is None:
            self.dt = get_dt(radius=sm.RADIUS_PARTICLE, mean_speed=mean_speed)
            ctx.set_dt(self.dt)
        self._host_exposed = False
        self._device_fresh = True

    # -- device context --------------------------------------------------------------------------
    def _algorithm_code(self) -> int:
        return 0

    def _ensure_ctx(self):
        if self._ctx is not None:
            return self._ctx
        sm = self.sm
        pp = self.pot.device_params()  # NotImplementedError for foreign callables
        n, nd = sm.NUM_PARTICLES, sm.NDIM
        k = self.flush_every
        if k is None:
            row_bytes = 2 * nd * n * 8
            k = max(1, min(256, (256 << 20) // max(row_bytes, 1)))
        k = max(1, min(int(k), max(1, sm.STEPS)))
        self._k = k
        fe = self.force_evals_per_step
        if fe is None:
            fe = 1 if sm.PBC else 2
        if fe == 1 and not sm.PBC:
            raise ValueError("force_evals_per_step=1 needs periodic boundaries (hard walls move particles between evaluations)")
        sigma = pp["sigma"]
        # multi-GPU: one process per GPU (torchrun); only the cut-off path decomposes, all-pairs runs stay replicas
        rank, world = parallel.dist_info()
        self._rank, self._world = (rank, world) if (world > 1 and self.r_cutoff) else (0, 1)
        # 2-D boxes on the cut-off path: the cell / list kernels are 3-D, so a 2-D system with a cutoff runs as the 3-D system
        # (x, L/2, y): the padding coordinate is the same for every particle, every pair has dy' = 0 exactly, no force ever acts
        # along it and it never changes -- r^2, forces, energies and (with the 2/3 of simulator.py:238) temperatures are those of
        # the 2-D system; gravity acts on the LAST dimension in both (simulator.py:205).  Only the occupied layer of the cubic cell
        # grid holds particles.  Single-rank runs; boxes whose cubic grid would not fit stay on the all-pairs kernels.
        self._embed = None
        if nd == 2 and self.r_cutoff and self._world == 1 and not os.environ.get("MDSEA_NO_2D_EMBED"):
            rl = float(self.r_cutoff) + (float(self.skin) if self.skin is not None else 0.3 * sigma)
            period = float(sm.LEN_BOX) if sm.PBC else float(sm.LEN_BOX) + 2.0 * rl
            nc = int(period // rl)
            if (nc >= 3 if sm.PBC else int(sm.LEN_BOX // rl) >= 1) and nc ** 3 <= 200_000_000:
                self._embed = [0, 2]   # device dimensions that hold x and y
        nd_dev = 3 if self._embed else nd
        self._nd_dev = nd_dev
        max_nbr = int(self.max_neighbors)
        if self._embed and max_nbr == 0:   # the library sizes the lists from N / L^3, which is not this layer's density
            rl = float(self.r_cutoff) + (float(self.skin) if self.skin is not None else 0.3 * sigma)
            max_nbr = int(1.6 * math.pi * rl * rl * n / float(sm.LEN_BOX) ** 2) + 32
        grid = (1, 1, 1)
        if self._world > 1:
            rlist = float(self.r_cutoff) + (float(self.skin) if self.skin is not None else 0.3 * sigma)
            grid = parallel.choose_proc_grid(self._world, int(sm.LEN_BOX // rlist))
        self._ctx = _capi.Context(
            rank=self._rank, world=self._world, proc_grid=grid,
            ndim=nd_dev, num_particles=n, boxlen=float(sm.LEN_BOX), mass=float(sm.MASS), radius=float(sm.RADIUS_PARTICLE),
            dt=float(self.dt), pbc=1 if sm.PBC else 0, algorithm=self._algorithm_code(),
            restitution=float(self.restitution_coeff), pot_kind=pp["kind"], precision=_PRECISIONS[self.precision],
            epsilon=pp["epsilon"], sigma=sigma, mie_m=pp["m"], mie_n=pp["n"], mie_a=pp["a"],
            r_cutoff=float(self.r_cutoff) if self.r_cutoff else 0.0,
            skin=float(self.skin) if self.skin is not None else 0.3 * sigma,
            force_evals_per_step=int(fe), ring_rows=k, device=int(self.device), max_neighbors=max_nbr,
        )
        if self._world > 1:
            self._ctx.comm_init(parallel.broadcast_unique_id(_capi.Context.nccl_unique_id))
            parallel.barrier()
            # rank 0 created the simfile tree; every rank (rank 0 included) now holds the file through the shared
            # memory-mapped backend: concurrent read-write opens are not something libhdf5's file locking allows
            if sm.dfile is not None and getattr(sm.dfile, "id", None) is not None and sm.dfile.id.valid:
                sm._close_datafile()
            sm._open_datafile(shared=True)
        self._ctx.set_temperature(self.temp, reset_ke=self.mean_ke is None)
        width = n if self._world == 1 else self._ctx.rank_capacity()
        self._rows_width = width
        self._rows = self._new_rows()
        self._rows_alt = None   # second host buffer set, created when run_simulation pipelines more than one batch
        return self._ctx

    def close(self) -> None:
        """Release the device context and the page-locked row buffers (the state stays readable on the host)."""
        if self._ctx is not None:
            self._sync_to_host()
            self._ctx.close()
            self._ctx = None
            self._host_exposed = True
        self._rows = self._rows_alt = None
        self._pinned.clear()

    def __del__(self):
        try:
            if getattr(self, "_ctx", None) is not None:
                self._ctx.close()
            getattr(self, "_pinned", []).clear()
        except Exception:  # pragma: no cover - interpreter shutdown
            pass

    def _new_rows(self) -> dict:
        k, nd, width = self._k, self._nd_dev, self._rows_width
        keep = self._pinned
        rows = dict(r=_pinned_empty((k, nd, width), keep), v=_pinned_empty((k, nd, width), keep),
                    pe=_pinned_empty((k,), keep), ke=_pinned_empty((k,), keep), temp=_pinned_empty((k,), keep))
        if self._world > 1:
            rows["ids"] = np.full((k, width), -1, dtype=np.int32)
            rows["n_owned"] = np.zeros(k, dtype=np.int64)
        return rows

    def _sync_to_device(self):
        ctx = self._ensure_ctx()
        if self._host_exposed:
            if self._embed:
                n = self.sm.NUM_PARTICLES
                r3 = np.full((3, n), 0.5 * float(self.sm.LEN_BOX), dtype=DTYPE)
                v3 = np.zeros((3, n), dtype=DTYPE)
                r3[self._embed], v3[self._embed] = self._r, self._v
                ctx.set_state(r3, v3)
            else:
                ctx.set_state(self._r, self._v)
            self._host_exposed = False
            self._device_fresh = False
        return ctx

    def _sync_to_host(self):
        if self._device_fresh and self._ctx is not None:
            self._r, self._v, self._acc, _ = self._ctx.get_state()
            if self._embed:
                self._r, self._v, self._acc = (np.ascontiguousarray(a[self._embed]) for a in (self._r, self._v, self._acc))
            if self._world > 1:  # every rank holds its owned entries only
                self._r, self._v, self._acc = (parallel.assemble_owned(a) for a in (self._r, self._v, self._acc))
            self._device_fresh = False

    # -- state attributes (host views of the device state) -----------------------------------------
    # Reading r_vec / v_vec downloads the state if the device holds a newer one and -- because the reference mutates these
    # arrays in place (simulator.py:165-166, 194-195, 205) -- marks it as possibly edited: the next device call uploads it
    # again.  With world > 1 the download is COLLECTIVE (every rank holds only the particles it owns and the pieces are
    # all-reduced): read these attributes on all ranks or on none.
    @property
    def r_vec(self) -> np.ndarray:
        self._sync_to_host()
        self._host_exposed = True
        return self._r

    @r_vec.setter
    def r_vec(self, value):
        self._sync_to_host()
        self._r = np.array(value, dtype=DTYPE)
        self._host_exposed = True

    @property
    def v_vec(self) -> np.ndarray:
        self._sync_to_host()
        self._host_exposed = True
        return self._v

    @v_vec.setter
    def v_vec(self, value):
        self._sync_to_host()
        self._v = np.array(value, dtype=DTYPE)
        self._host_exposed = True

    @property
    def acc(self) -> np.ndarray:
        self._sync_to_host()
        if self._acc is None:
            self._acc = np.zeros((self.sm.NDIM, self.sm.NUM_PARTICLES), dtype=DTYPE)
        return self._acc

    @acc.setter
    def acc(self, value):
        self._acc = np.array(value, dtype=DTYPE)

    @property
    def dists(self):
        return self._dists

    @property
    def drunits(self):
        return self._drunits

    # -- reference methods ---------------------------------------------------------------------------
    def update_files(self) -> None:
        """Write the current state as row ``self.step`` of the five datasets (simulator.py:143-150)."""
        values = [[self.r_vec], [self.v_vec], [self.mean_pe], [self.mean_ke], [self.temp]]
        for dataset, val in zip(self.sm.all_dsnames, values):
            self.sm.update_ds(dataset, val, self.step)

    def update_dists(self, radius: Optional[float] = None, where: str = "inside") -> np.ndarray:
        """Pair distances / unit vectors in combinations order, computed on the device
        (simulator.py:257-293).  Small N only: the output is O(N^2)."""
        if where not in ("inside", "outside"):
            raise ValueError(f"'{where}' is not a valid value for 'where'. Try 'inside' or 'outside' instead.")
        ctx = self._sync_to_device()
        d, u = ctx.update_dists()
        if self._embed:
            u = np.ascontiguousarray(u[:, self._embed])
        if radius is None:
            self.pairs_indexes = np.repeat(True, d.shape[0])
        else:
            self.pairs_indexes = d < radius if where == "inside" else d > radius
            d, u = d[self.pairs_indexes], u[self.pairs_indexes]
        self._dists, self._drunits = d, u
        return d

    def update_acc(self, radius: float = None) -> np.ndarray:
        """Accelerations for the current positions, evaluated by the force kernels (simulator.py:295-311)."""
        if radius is not None:
            raise NotImplementedError("update_acc(radius=...) is broken in the reference (simulator.py:289-309); "
                                      "use the solver's r_cutoff field instead")
        ctx = self._sync_to_device()
        self._acc, self._last_pe = ctx.update_acc()
        if self._embed:
            self._acc = np.ascontiguousarray(self._acc[self._embed])
        return self._acc

    def update_mean_pe(self):
        ctx = self._sync_to_device()
        _, self.mean_pe = ctx.update_acc()
        return self.mean_pe

    def update_mean_ke(self):
        v = self.v_vec
        self.mean_ke = 0.5 * self.sm.MASS * float(np.einsum("dn,dn->n", v, v).mean())
        return self.mean_ke

    def update_energies(self) -> None:
        self.update_mean_pe()
        self.update_mean_ke()

    def update_temp(self) -> Optional[float]:
        if self.mean_ke is None:
            log.warning("Cannot update the temperature without first evaluating the mean kinetic energy.")
            return None
        self.temp = (2 / 3) * self.mean_ke / Config.k_boltzmann
        return self.temp

    @property
    def com(self):
        """Centre of mass (simulator.py:207-213), reduced on the device; None (with the reference's warning) when periodic."""
        if self.sm.PBC:
            log.warning("COM computation not available for PBC yet!")
            return
        if self._world > 1:
            return np.mean(self.r_vec, axis=1)
        com = self._sync_to_device().com_rog()[0]
        return com[self._embed] if self._embed else com

    @property
    def rog(self) -> float:
        """Radius of gyration: root-mean-square distance to the centre of mass (simulator.py:215-231), on the device."""
        if self.sm.PBC:
            # the reference subtracts `self.com` (None for periodic boxes) from the positions: same TypeError here
            raise TypeError("unsupported operand type(s) for -: 'float' and 'NoneType'")
        if self._world > 1:
            d = np.stack(self.r_vec, axis=-1) - self.com
            return float(np.sqrt(np.mean(quicker.norm(d, axis=1) ** 2)))
        return self._sync_to_device().com_rog()[1]

    # -- boundary conditions, special events, fields as callable steps (simulator.py:156-205) -----------------
    def _apply_bc_device(self) -> None:
        ctx = self._sync_to_device()
        ctx.apply_bc()
        self._device_fresh = True

    def apply_pbc(self) -> None:
        """One-shot strict wrap (simulator.py:156-166).  On the device when the box is periodic."""
        if self.sm.PBC:
            return self._apply_bc_device()
        r, L = self.r_vec, self.sm.LEN_BOX
        r[np.where(r < 0)] += L
        r[np.where(r > L)] -= L

    def apply_hbc(self) -> None:
        """Clamp to the walls and reflect with the restitution coefficient (simulator.py:168-179).  On the device when
        the box has hard walls."""
        if not self.sm.PBC:
            return self._apply_bc_device()
        r, v, L, a = self.r_vec, self.v_vec, self.sm.LEN_BOX, self.sm.RADIUS_PARTICLE
        whr = np.where(r - a < 0)
        r[whr] = a
        v[whr] *= -self.restitution_coeff
        whr = np.where(r + a > L)
        r[whr] = L - a
        v[whr] *= -self.restitution_coeff

    def apply_bc(self) -> None:
        """``apply_pbc if sm.PBC else apply_hbc`` (simulator.py:90)."""
        self._apply_bc_device()

    def apply_special(self) -> None:
        """Isothermal rescale and scheduled quenches of the current step (simulator.py:185-190)."""
        if self.isothermal:
            self._quench(self.temp_init)
        if self.step in self.sm.QUENCH_STEP:
            self._quench(self.sm.QUENCH_T.pop(0))

    def _quench(self, temperature: float) -> None:
        self.update_temp()   # simulator.py:192-195
        if self.temp != 0.0:
            self.v_vec *= (temperature / self.temp) ** 0.5

    def apply_field(self) -> None:
        if self.gravity:     # simulator.py:201-205
            self.v_vec[-1] -= Config.gravity_acceleration * self.dt

    # -- the pair table (simulator.py:122-137): built on first use, O(N^2) like the reference's ------------
    @property
    def all_pairs(self) -> np.ndarray:
        if getattr(self, "_all_pairs", None) is None:
            i, j = np.triu_indices(self.sm.NUM_PARTICLES, k=1)   # itertools.combinations order
            self._all_pairs = np.stack([i, j], axis=1).astype(np.int32)
        return self._all_pairs

    @property
    def pairs(self) -> np.ndarray:
        """Pairs selected by the last update_dists / update_pairs call (simulator.py:313-316)."""
        return self.all_pairs[self.pairs_indexes]

    def update_pairs(self, radius: Optional[float] = None, where: str = "inside"):
        """Pairs within (or outside) a radius with their distances and unit vectors (simulator.py:318-323)."""
        self.update_dists(radius=radius, where=where)
        return zip(self.all_pairs[self.pairs_indexes], self.dists, self.drunits)


@dataclass
class ContinuousPotentialSolver(_BaseSimulator):
    algorithm: str = field(default="verlet")

    def __post_init__(self) -> None:
        super().__post_init__()
        if self.algorithm not in _ALGORITHMS:  # simulator.py:524-528
            raise KeyError(f"Algorithm '{self.algorithm}' not found in {tuple(_ALGORITHMS.keys())}")

    def _algorithm_code(self) -> int:
        return _ALGORITHMS[self.algorithm]

    # -- batches of steps on the device ------------------------------------------------------------
    def _quench_targets(self, first_step: int, nsteps: int):
        """Per-step (isothermal, scheduled) quench targets in the reference's call order (simulator.py:185-190)."""
        if not self.isothermal and not self.sm.QUENCH_STEP:
            return None
        q = np.full((nsteps, 2), np.nan)
        for k in range(nsteps):
            s = first_step + k
            if self.isothermal:
                q[k, 0] = self.temp_init
            if s in self.sm.QUENCH_STEP:
                q[k, 1] = self.sm.QUENCH_T.pop(0)
            if s == 0 and self.mean_ke is None and not np.isnan(q[k]).all():
                log.warning("Cannot update the temperature without first evaluating the mean kinetic energy.")
        return q if not np.isnan(q).all() else None

    def _submit_batch(self, first_step: int, nsteps: int, record: bool, rows: dict):
        """Enqueue `nsteps` steps; the rows land in `rows` once the batch has been completed with _complete_batch."""
        if self.apply_restcoeff:
            raise NotImplementedError  # apply_collision_damping, simulator.py:546-548,565-566
        ctx = self._sync_to_device()
        q = self._quench_targets(first_step, nsteps)
        g = Config.gravity_acceleration if self.gravity else 0.0
        ctx.run(nsteps, kb=Config.k_boltzmann, g=g, quench=q, record=record, rows=rows, asynchronous=True)
        self._device_fresh = True

    def _complete_batch(self, nsteps: int, rows: dict):
        self._ctx.wait()   # oldest batch in flight
        self.mean_pe = float(rows["pe"][nsteps - 1])
        self.mean_ke = float(rows["ke"][nsteps - 1])
        self.temp = float(rows["temp"][nsteps - 1])
        return rows

    def _run_batch(self, first_step: int, nsteps: int, record: bool):
        if self.apply_restcoeff:
            raise NotImplementedError  # before any device work, like the reference (simulator.py:565-566)
        self._ensure_ctx()
        self._submit_batch(first_step, nsteps, record, self._rows)
        return self._complete_batch(nsteps, self._rows)

    def advance(self):
        """Advance the simulation one step (simulator.py:560-568)."""
        self._run_batch(self.step, 1, record=False)

    def run_simulation(self) -> None:
        """Run steps ``self.step .. STEPS-1`` and store one row per step (simulator.py:570-589)."""
        sm = self.sm
        simrange = range(self.step, sm.STEPS)
        if len(simrange) == 0:
            self.update_files()
        self.pbarr.set_start()
        self._ensure_ctx()
        # Two batches in flight: while batch b is being computed, the rows of batch b-1 travel to the host (the
        # library alternates between two device ring sets) and are written into the datasets.
        def store(first: int, n: int, rows: dict) -> None:
            self._complete_batch(n, rows)
            if self._world > 1:
                if not sm.dfile.id.valid:
                    sm._open_datafile(shared=True)
                parallel.scatter_rows(sm, rows, n, first, self._rank)
            else:
                if self._embed:   # (x, pad, y) rows of the 3-D device system -> the 2-D rows of the datasets
                    sm.update_ds_block(sm.rcoord_dsname, rows["r"][:n][:, self._embed], first)
                    sm.update_ds_block(sm.vcoord_dsname, rows["v"][:n][:, self._embed], first)
                else:
                    sm.update_ds_block(sm.rcoord_dsname, rows["r"][:n], first)
                    sm.update_ds_block(sm.vcoord_dsname, rows["v"][:n], first)
                sm.update_ds_block(sm.potenergy_dsname, rows["pe"][:n], first)
                sm.update_ds_block(sm.kinenergy_dsname, rows["ke"][:n], first)
                sm.update_ds_block(sm.temp_dsname, rows["temp"][:n], first)
            self.step = first + n - 1

        s = simrange.start
        if s + self._k < sm.STEPS and self._rows_alt is None:
            self._rows_alt = self._new_rows()
        bufs = [self._rows, self._rows_alt]
        in_flight = None
        b = 0
        while s < sm.STEPS:
            n = min(self._k, sm.STEPS - s)
            for i in range(s, s + n):
                self.pbarr.log_progress(i)
            rows = bufs[b % 2] if bufs[1] is not None else bufs[0]
            self._submit_batch(s, n, True, rows)
            if in_flight is not None:
                store(*in_flight)
            in_flight = (s, n, rows)
            s += n
            b += 1
        if in_flight is not None:
            store(*in_flight)
        bad = self._ctx.first_nonfinite_step()
        if bad >= 0:
            log.warning("Non-finite energies from step %d on (the reference writes NaN rows silently as well; "
                        "check Config.k_boltzmann / dt).", bad)
        self.pbarr.set_finish()
        self.pbarr.log_duration()

    def stats(self) -> dict:
        """Counters of the device context (force evaluations, pair interactions, launches, rebuilds)."""
        return self._ensure_ctx().stats()
