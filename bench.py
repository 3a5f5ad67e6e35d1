#!/usr/bin/env python
"""bench.py -- LJ pair-interactions/s and timesteps/s of the ContinuousPotentialSolver step loop on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c2|c3] [--precision fp64|fp32]
    python bench.py --impl reference ...      # the CPU implementation of the same path on the host cores

One bench "step" = one pass of the hot path over one batch: `md_steps` consecutive MD timesteps of the
workload, issued as ONE call of the C ABI (mdsea_b200_run).  Prints ONE JSON line (rank 0).

value    unique in-cutoff LJ pair interactions per second, whole job, state resident in HBM, timed with CUDA
         events on the library's stream (rows are still buffered in the device ring every step).
e2e      same metric through the C ABI with HOST buffers: every step uploads the batch's initial (r, v) from
         pinned memory and copies the batch's dataset rows back, inside the timed region.
roofline dominant kernel: the force kernel (CUDA-core FP pipe; peak = FMA micro-benchmark run here, because
         MEASURED_PEAKS.json holds only HBM and bf16-tensor peaks), plus the HBM-bound integrator.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FLOP_PER_UNIQUE_PAIR = {3: 45.0, 2: 34.0, 1: 23.0}
# CUDA-core FMA peak from the data sheet: 148 SMs x 128 fp32 (64 fp64) lanes x 2 flop x 1.965 GHz boost
DATASHEET_TFLOPS = {"fp32": 148 * 128 * 2 * 1.965e9 / 1e12, "fp64": 148 * 64 * 2 * 1.965e9 / 1e12}

WORKLOADS = {
    # configs[1] of BASELINE.json: 3-D LJ fluid, 4096 particles, vf 0.3, all-pairs tiled kernel
    "c2": dict(name="3D LJ fluid N=4096 vf=0.3 all-pairs (BASELINE configs[1])", ndim=3, per_row=16, vf=0.3,
               r_cutoff=0.0, skin=0.0, md_steps=200, precision="fp64"),
    # configs[2]: 3-D LJ liquid, 1M particles, cell list + Verlet skin
    # and, for N GPUs, configs[3] (weak scaling, ~1e6 particles per GPU on simple-cubic lattices: gen.py needs perfect cubes)
    "c3": dict(name="3D LJ liquid N=1e6 vf=0.3 rc=2.5 skin=0.3 cell+Verlet list (BASELINE configs[2])", ndim=3,
               per_row=100, vf=0.3, r_cutoff=2.5, skin=0.3, md_steps=20, precision="fp32",
               per_row_by_world={1: 100, 2: 126, 4: 159, 8: 200},
               name_multi="3D LJ liquid ~1e6 particles/GPU vf=0.3 rc=2.5 skin=0.3, spatial domain decomposition "
                          "(BASELINE configs[3], weak scaling)"),
}

# x (the contiguous index of the cell grid) is never split: see mdsea_b200.parallel.choose_proc_grid
PROC_GRIDS = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (1, 2, 4)}
if os.environ.get("MDSEA_PROC_GRID"):   # A/B: e.g. MDSEA_PROC_GRID=2,2,2
    _g = tuple(int(x) for x in os.environ["MDSEA_PROC_GRID"].split(","))
    PROC_GRIDS[_g[0] * _g[1] * _g[2]] = _g


# ---------------------------------------------------------------------------------------------------
def initial_state(ndim, per_row, vf, seed=0):
    """gen.py initial conditions (lattice + Maxwell-Boltzmann, k_B = 1) through the product's own generators."""
    from mdsea_b200 import quicker
    from mdsea_b200.gen import PosGen, VelGen
    from mdsea_b200.helpers import get_dt, nsphere_volume

    n = per_row ** ndim
    L = (n * nsphere_volume(ndim, 0.5) / vf) ** (1 / ndim)
    r0 = PosGen(ndim=ndim, boxlen=L, nparticles=n).simplecubic()
    np.random.seed(seed)
    v0 = VelGen(ndim=ndim, nparticles=n).mb(1.0, 1.0, 1.0)
    dt = get_dt(0.5, float(quicker.norm(v0, axis=0).mean()))
    return n, L, r0, v0, dt


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.lines, self.proc = device, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons, pw = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); pw.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ---------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU restatement of the reference's algorithm (oracle/), the only
# place bench.py executes oracle code.  The reference itself is Python and cannot travel to the GPU box.
# ---------------------------------------------------------------------------------------------------
def cpu_allpairs_numpy(wl, steps, warmup):
    from oracle.oracle_np import MiePotential, OracleSolver

    n, L, r0, v0, dt = initial_state(wl["ndim"], wl["per_row"], wl["vf"])
    o = OracleSolver(r=r0, v=v0, boxlen=L, dt=dt, pot=MiePotential())
    for _ in range(warmup):
        o.advance()
    p0 = o.pairs_evaluated
    t0 = time.perf_counter()
    for _ in range(steps):
        o.advance()
    t = time.perf_counter() - t0
    return dict(pairs=o.pairs_evaluated - p0, seconds=t, md_steps=steps, n=n, cores=1,
                sample=f"{steps} verlet step(s) = {2 * steps} force evaluations of the same N={n} state, NumPy float64 "
                       f"restatement of the reference (single-threaded like the reference)")


def run_reference_arm(args, wl):
    rank, world, _ = dist_env()
    if rank != 0:
        return
    name = wl["name"]
    if wl["r_cutoff"] > 0:
        from oracle import oracle_c

        if args.gpus > 1 and args.gpus in wl.get("per_row_by_world", {}):
            # the SAME system our arm decomposes over args.gpus GPUs (weak scaling: ~1e6 particles per GPU), on the host cores
            wl = dict(wl, per_row=wl["per_row_by_world"][args.gpus])
            name = wl["name_multi"]
        res = oracle_c.bench_cutoff(wl, steps=args.steps, warmup=args.warmup)
    else:
        res = cpu_allpairs_numpy(wl, args.steps, args.warmup)
    value = res["pairs"] / res["seconds"]
    line = {
        "impl": "reference", "metric": "LJ pair-interactions/s", "value": value, "unit": "pair-interactions/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * res["seconds"] / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "timesteps_per_s": res["md_steps"] / res["seconds"],
        "config": {"workload": name, "particles_total": res["n"], "md_steps_per_step": res["md_steps"] // max(args.steps, 1),
                   "k_boltzmann": 1.0},
        "cpu_baseline": {"value": value, "unit": "pair-interactions/s", "cores": res["cores"], "kind": "port",
                         "sample": res["sample"]},
        "e2e": {"value": value, "unit": "pair-interactions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": {"nproc": os.cpu_count()},
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------
# parity of the benched configuration itself: the same state advanced k steps by the library and by the C oracle
# (oracle/lj_oracle.c -- the checker, never the thing measured)
# ---------------------------------------------------------------------------------------------------
_H1, _H2 = np.uint64(0x9E3779B97F4A7C15), np.uint64(0xC2B2AE3D27D4EB4F)


def row_hashes(off, idx, n):
    """One 64-bit hash per particle of its (ascending) neighbour row: sum of per-entry hashes + a length term."""
    cnt = np.diff(off).astype(np.uint64)
    with np.errstate(over="ignore"):
        e = (idx.astype(np.uint64) + np.uint64(1)) * _H1
        e ^= e >> np.uint64(29)
        e *= _H2
        cs = np.concatenate([[np.uint64(0)], np.cumsum(e, dtype=np.uint64)])
        h = cs[off[1:]] - cs[off[:-1]] + cnt * _H2
    return h


def parity_allpairs(ctx, wl, L, dt, r_start, v_start, k, precision):
    """All-pairs workloads: k steps from (r_start, v_start) against the NumPy float64 restatement of the reference."""
    from oracle.oracle_np import MiePotential, OracleSolver

    tol = 1e-10 if precision == "fp64" else 1e-5
    rows = dict(r=None, v=None, pe=np.empty(k), ke=np.empty(k), temp=np.empty(k))
    ctx.set_state(r_start, v_start)
    ctx.run(k, kb=1.0, record=True, rows=rows)
    ref = OracleSolver(r=np.array(r_start), v=np.array(v_start), boxlen=L, dt=dt, pot=MiePotential()).run(k)
    pe_rel = float(np.max(np.abs(rows["pe"] - ref["pe"]) / np.abs(ref["pe"])))
    ke_rel = float(np.max(np.abs(rows["ke"] - ref["ke"]) / np.abs(ref["ke"])))
    return {"steps": k, "tolerance": tol, "pe_rel": pe_rel, "ke_rel": ke_rel, "oracle": "oracle/oracle_np.py (float64)",
            "ok": bool(pe_rel <= tol and ke_rel <= tol)}


def parity_check(ctx, wl, n, L, dt, r_start, v_start, k, precision, rank, world, dist, torch):
    """Advance (r_start, v_start) k steps on the GPU path under test and on the C oracle.  Must agree: the neighbour lists
    built from the start positions (bit for bit: same positions on both sides), the energies of every step (1e-10 fp64 /
    1e-5 fp32) and the number of list rebuilds.  The lists in force after the k steps are compared as well: with fp64
    pair math they must be identical again; with fp32 pair math the trajectories differ by ~1e-7 sigma by then, so a pair
    within that distance of the list radius may legitimately fall on the other side -- the differing entries are reported
    and bounded by 1e-6 of the list."""
    import zlib

    tol = 1e-10 if precision == "fp64" else 1e-5
    rows = dict(r=None, v=None, pe=np.empty(k), ke=np.empty(k), temp=np.empty(k))
    ctx.set_state(r_start, v_start)
    s0 = ctx.stats()

    def global_hashes():
        off, idx = ctx.get_neighbors()
        h = row_hashes(off, idx, n)
        ne = int(idx.shape[0])
        if world > 1:   # every rank holds the rows of the particles it owns
            t = torch.from_numpy(h.view(np.int64)).cuda()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            h = t.cpu().numpy().view(np.uint64)
            nt = torch.tensor([ne], dtype=torch.int64, device="cuda")
            dist.all_reduce(nt, op=dist.ReduceOp.SUM)
            ne = int(nt.item())
        return h, ne, (off, idx)

    h_start, ne_start, _ = global_hashes()
    ctx.run(k, kb=1.0, record=True, rows=rows)
    s1 = ctx.stats()
    h_end, ne_end, csr_end = global_hashes()
    if rank != 0:
        return None
    from oracle import oracle_c

    t0 = time.perf_counter()
    ref_off, ref_idx = oracle_c.neighbors(r_start, L, wl["r_cutoff"] + wl["skin"])
    o = oracle_c.CutoffSystem(r_start, v_start, L, wl["r_cutoff"], wl["skin"], dt)
    ref_h_start, ref_ne_start, crc_start = row_hashes(ref_off, ref_idx, n), int(ref_idx.shape[0]), zlib.crc32(ref_idx.tobytes())
    ref = o.run(k)
    oc = o.counters()
    ref_off, ref_idx = o.neighbor_list()
    o.close()
    t_oracle = time.perf_counter() - t0
    ref_h_end = row_hashes(ref_off, ref_idx, n)
    pe_rel = float(np.max(np.abs(rows["pe"] - ref["pe"]) / np.abs(ref["pe"])))
    ke_rel = float(np.max(np.abs(rows["ke"] - ref["ke"]) / np.abs(ref["ke"])))
    lists_equal = bool(ne_start == ref_ne_start and np.array_equal(h_start, ref_h_start))
    rows_diff_end = int(np.count_nonzero(h_end != ref_h_end))
    end_ok = rows_diff_end == 0 if precision == "fp64" else rows_diff_end <= max(4, int(1e-6 * ref_idx.shape[0]))
    reb = int(s1["list_rebuilds"] - s0["list_rebuilds"])
    out = {"steps": k, "tolerance": tol, "pe_rel": pe_rel, "ke_rel": ke_rel,
           "list_equal": lists_equal, "list_checksum_equal": lists_equal, "list_entries": ne_start, "list_crc32_oracle": crc_start,
           "lists_after_k_steps": {"entries": ne_end, "entries_oracle": int(ref_idx.shape[0]), "rows_that_differ": rows_diff_end,
                                   "ok": bool(end_ok)},
           "rebuilds": reb, "rebuilds_oracle": oc["rebuilds"], "rebuilds_equal": reb == oc["rebuilds"],
           "oracle": "oracle/lj_oracle.c (float64, cell + Verlet list, same trigger): the cut-off variant is THIS repository's "
                     "extension (the reference has no working cutoff), self-pinned -- to the reference-generated golden rows through "
                     "the all-pairs limit and to the NumPy restatement with a cutoff",
           "oracle_seconds": t_oracle}
    out["ok"] = bool(pe_rel <= tol and ke_rel <= tol and lists_equal and end_ok and out["rebuilds_equal"])
    return out


# ---------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------
def make_ctx(wl, n, L, dt, device, precision, ring, rank=0, world=1, proc_grid=(1, 1, 1), evals=1):
    from mdsea_b200 import _capi

    return _capi.Context(ndim=wl["ndim"], num_particles=n, boxlen=L, mass=1.0, radius=0.5, dt=dt, pbc=1, algorithm=0,
                         pot_kind=1, precision=0 if precision == "fp64" else 1, epsilon=1.0, sigma=1.0, mie_m=12.0,
                         mie_n=6.0, mie_a=0.0, r_cutoff=wl["r_cutoff"], skin=wl["skin"], force_evals_per_step=evals,
                         ring_rows=ring, device=device, rank=rank, world=world, proc_grid=proc_grid)


def run_ours(args, wl):
    import torch
    import torch.distributed as dist

    from mdsea_b200 import _capi, parallel

    rank, world, local = dist_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this framework has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    # host side of the row traffic: this rank's pinned buffers live on the NUMA node of its GPU
    all_cpus = os.sched_getaffinity(0)
    numa = None if os.environ.get("MDSEA_NO_NUMA_BIND") else parallel.bind_to_gpu_numa(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    precision = args.precision or wl["precision"]
    md_steps = args.md_steps or wl["md_steps"]
    ndim = wl["ndim"]
    decomposed = world > 1 and wl["r_cutoff"] > 0
    if decomposed:
        # one system, spatially decomposed over the ranks (every rank generates the same gen.py state)
        if world not in wl["per_row_by_world"]:
            raise SystemExit(f"bench.py: no weak-scaling lattice defined for {world} GPUs (have {sorted(wl['per_row_by_world'])})")
        n, L, r0, v0, dt = initial_state(ndim, wl["per_row_by_world"][world], wl["vf"], seed=0)
        ctx = make_ctx(wl, n, L, dt, local, precision, md_steps, rank=rank, world=world, proc_grid=PROC_GRIDS[world],
                       evals=args.evals)
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(_capi.Context.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()))
        width = ctx.rank_capacity()
    else:
        # replicas: every rank runs its own independent system (all-pairs does not shard)
        n, L, r0, v0, dt = initial_state(ndim, wl["per_row"], wl["vf"], seed=rank)
        ctx = make_ctx(wl, n, L, dt, local, precision, md_steps, evals=args.evals)
        width = n
    ctx.set_temperature(1.0)
    pin = lambda shape, dt_=torch.float64: torch.empty(shape, dtype=dt_, pin_memory=True).numpy()
    r_pin, v_pin = pin((ndim, n)), pin((ndim, n))
    r_pin[:], v_pin[:] = r0, v0
    rows = dict(r=pin((md_steps, ndim, width)), v=pin((md_steps, ndim, width)), pe=pin((md_steps,)), ke=pin((md_steps,)),
                temp=pin((md_steps,)))
    if decomposed:
        rows["ids"] = pin((md_steps, width), torch.int32)
        rows["n_owned"] = pin((md_steps,), torch.int64)
    rows_dev = dict(r=None, v=None, pe=rows["pe"], ke=rows["ke"], temp=rows["temp"])
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_loop(body, steps):
        """steps x body(), each bracketed by its own event pair on the library's stream, L2 flushed in between."""
        evs = []
        for _ in range(steps):
            flush_buf.zero_()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            body()
            e1.record(stream)
            evs.append((e0, e1))
        torch.cuda.synchronize()
        return [a.elapsed_time(b) for a, b in evs]

    # ---------------- device-resident throughput ("value") ----------------
    ctx.set_state(r_pin, v_pin)
    def body_dev():
        ctx.run(md_steps, kb=1.0, record=True, rows=rows_dev)
    for _ in range(args.warmup):
        body_dev()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    s0 = ctx.stats()
    t_wall0 = time.perf_counter()
    ms = timed_loop(body_dev, args.steps)
    t_wall = time.perf_counter() - t_wall0
    barrier()
    s1 = ctx.stats()
    clocks = sampler.stop() if rank == 0 else None
    pairs = s1["pair_evals_unique"] - s0["pair_evals_unique"]
    launches = s1["kernel_launches"] - s0["kernel_launches"]
    fevals = s1["force_evals"] - s0["force_evals"]
    t_dev = sum(ms) * 1e-3
    energy_ok = bool(np.isfinite(rows["pe"]).all() and np.isfinite(rows["ke"]).all())

    # the state the timed region ended in (a melted liquid after (warmup + steps) * md_steps timesteps): input of the e2e
    # batches, of the roofline sample and of the parity check below.  Decomposed runs: every rank owns a part of it.
    r_cur, v_cur, _, _ = ctx.get_state(want_acc=False)
    if decomposed:
        for a in (r_cur, v_cur):
            t = torch.from_numpy(np.nan_to_num(a, nan=0.0)).cuda()
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
            a[:] = t.cpu().numpy()
            del t
    r_pin[:], v_pin[:] = r_cur, v_cur
    del r_cur, v_cur

    # ---------------- end to end through the C ABI with host buffers ----------------
    # Every step: upload the batch's initial (r, v) from pinned memory, run md_steps, copy every dataset row of the
    # batch back into pinned host buffers.  Two batches are kept in flight (run_async / wait): the library
    # alternates between two device ring sets, so the rows of step s cross PCIe while step s+1 is computed.  The
    # timed region is the wall clock of all K steps including the final drain.  No L2 flush here: a flush needs a
    # device-wide synchronisation that would serialise the pipeline, and the working set (state + neighbour list
    # + row rings, > 1 GB at 1e6 particles) is far larger than the 126 MB L2 anyway.
    rows_b = dict(r=pin((md_steps, ndim, width)), v=pin((md_steps, ndim, width)), pe=pin((md_steps,)), ke=pin((md_steps,)),
                  temp=pin((md_steps,)))
    if decomposed:
        rows_b["ids"] = pin((md_steps, width), torch.int32)
        rows_b["n_owned"] = pin((md_steps,), torch.int64)
    row_sets = [rows, rows_b]

    def e2e_steps(k):
        in_flight = 0
        for s in range(k):
            ctx.set_state(r_pin, v_pin)
            ctx.run(md_steps, kb=1.0, record=True, rows=row_sets[s % 2], asynchronous=True)
            in_flight += 1
            if in_flight == 2:
                ctx.wait()
                in_flight -= 1
        while in_flight:
            ctx.wait()
            in_flight -= 1

    e2e_steps(2)
    barrier()
    p0 = ctx.stats()["pair_evals_unique"]
    t0 = time.perf_counter()
    e2e_steps(args.steps)
    torch.cuda.synchronize()
    e2e_times = [time.perf_counter() - t0]
    barrier()
    e2e_pairs = ctx.stats()["pair_evals_unique"] - p0
    t_e2e = sum(e2e_times)
    h2d = 2 * ndim * n * 8
    n_loc = ctx.stats()["n_local"] if decomposed else n
    d2h = md_steps * ((2 * ndim * 8 + (4 if decomposed else 0)) * n_loc + 3 * 8)

    # ---------------- per-kernel roofline numbers (library's own CUDA-event timers) ----------------
    ctx.set_state(r_pin, v_pin)
    ctx.run(md_steps, kb=1.0, record=True, rows=rows_dev)
    ctx.set_profiling(True)
    q0 = ctx.stats()
    ctx.run(md_steps, kb=1.0, record=True, rows=rows_dev)
    q1 = ctx.stats()
    ctx.set_profiling(False)
    f_evals = q1["force_evals"] - q0["force_evals"]
    f_ms = (q1["ms_force"] - q0["ms_force"]) / max(f_evals, 1)
    f_pairs = (q1["pair_evals_unique"] - q0["pair_evals_unique"]) / max(f_evals, 1)
    flops = FLOP_PER_UNIQUE_PAIR[ndim] * f_pairs
    fp64 = precision == "fp64"
    peak_tf = _capi.measure_fma_peak(local, fp64=fp64)
    achieved_tf = flops / (f_ms * 1e-3) / 1e12 if f_ms > 0 else 0.0
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = peaks.get("hbm_gbs", 6650.0)
    i_ms = (q1["ms_integrate"] - q0["ms_integrate"]) / max(md_steps, 1)
    if wl["r_cutoff"] > 0:
        # cut-off path: ONE kernel per step (second half kick + KE + rows, then wrap + first half kick + drift of the next
        # step): read r, v, a, build positions (4 x 24 B) + id (4 B); write r, v (48 B), the row (48 B) and the packed copy
        integ_bytes = n_loc * (4 * ndim * 8 + 4 + 2 * ndim * 8 + 2 * ndim * 8 + (16 if precision == "fp32" else 32))
        integ_name = "finish kernel fused with the next step's kick + drift (k_cut_kick_finish<1>)"
    else:
        # per component and step: read r,v,a + write r,v (kick_drift) ; read v,a,r + write v (kick_finish) ; write row r,v
        integ_bytes = n_loc * ndim * (5 * 8 + 4 * 8 + 2 * 8)
        integ_name = "fused integrator (kick_finish + kick_drift)"
    # DRAM bytes per launch of the force kernel from the tracked ncu --set full capture of this workload (same N)
    force_traffic, force_traffic_src = None, None
    if world == 1 and wl is WORKLOADS.get("c3"):
        src = "r01_ncu_c3_force_f64.txt" if fp64 else "r02/r02_ncu_c3_kernels.txt"
        force_traffic = ncu_traffic(src, "k_nbr_force_f64" if fp64 else "k_tile_force_f32")
        force_traffic_src = "profiles/" + src if force_traffic else None
    elif world == 1 and wl is WORKLOADS.get("c2"):
        src = "r02/r02_ncu_c2_allpairs_f64.txt" if fp64 else "r02/r02_ncu_c2_allpairs_f32.txt"
        force_traffic = ncu_traffic(src, "k_allpairs")
        force_traffic_src = "profiles/" + src if force_traffic else None
    roof_hbm = None
    if i_ms > 0:
        roof_hbm = {"bound": "hbm", "kernel": integ_name, "achieved": integ_bytes / (i_ms * 1e-3) / 1e9,
                    "peak": hbm_peak, "unit": "GB/s", "frac": integ_bytes / (i_ms * 1e-3) / 1e9 / hbm_peak,
                    "traffic": ncu_traffic("r02/r02_ncu_c3_kernels.txt", "k_cut_kick_finish<1>") if wl is WORKLOADS.get("c3") and not fp64 else None,
                    "bytes_per_launch": integ_bytes, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback"}

    # what the box's PCIe / host memory path gives when all ranks copy at once: one 256 MB device -> pinned-host copy per rank,
    # started together -- the ceiling of e2e.d2h_gbs_per_rank (8 GPUs of one node do not get 8 x the single-GPU rate)
    raw_d2h = None
    try:
        src_t = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        dst_t = torch.empty(256 << 20, dtype=torch.uint8, pin_memory=True)
        dst_t.copy_(src_t, non_blocking=True)
        barrier()
        t0 = time.perf_counter()
        for _ in range(4):
            dst_t.copy_(src_t, non_blocking=True)
        torch.cuda.synchronize()
        dt_raw = time.perf_counter() - t0
        rt = torch.tensor([dt_raw], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(rt, op=dist.ReduceOp.MAX)
        raw_d2h = 4 * (256 << 20) / float(rt.item()) / 1e9
        del src_t, dst_t
    except Exception:
        pass
    os.sched_setaffinity(0, all_cpus)   # the CPU legs below (oracle, cpu_baseline) use every host core
    # ---------------- parity of this very configuration against the CPU oracle ----------------
    parity = None
    if not args.no_parity:
        if wl["r_cutoff"] > 0:
            parity = parity_check(ctx, wl, n, L, dt, r_pin, v_pin, 10 if world == 1 else 6, precision, rank, world, dist, torch)
        elif rank == 0:
            parity = parity_allpairs(ctx, wl, L, dt, r_pin, v_pin, 3, precision)
    transport = {0: "single GPU", 1: "p2p (CUDA IPC stores over NVLink + arrival counters; NCCL for control collectives)",
                 2: "nccl (ncclSend/ncclRecv)", 3: "in-process"}[int(ctx.stats().get("halo_transport", 0))]

    # ---------------- aggregate over ranks ----------------
    t_dev_max, t_e2e_max, tot_pairs, tot_e2e_pairs = t_dev, t_e2e, pairs, e2e_pairs
    if world > 1:
        tt = torch.tensor([t_dev, t_e2e], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        pp = torch.tensor([pairs, e2e_pairs], dtype=torch.float64, device="cuda")
        dist.all_reduce(pp, op=dist.ReduceOp.SUM)
        t_dev_max, t_e2e_max = tt.tolist()
        tot_pairs, tot_e2e_pairs = pp.tolist()

    extra = None
    if world == 1 and not args.no_extra and args.workload != "c2":
        extra = allpairs_extra(local)
    # strong scaling beside the weak-scaling headline: 8 M particles at every N (at N = 8 that IS the headline system),
    # 64 M particles (BASELINE configs[4]) at N = 8
    strong = {}
    if wl is WORKLOADS["c3"] and precision == "fp32" and not args.no_strong:
        ctx.close()
        ctx = None
        if world < 8:
            strong["strong_8m"] = strong_scaling_extra(200, "strong_8m", rank, world, local, torch, dist)
        if world == 8:
            strong["strong_64m"] = strong_scaling_extra(400, "strong_64m", rank, world, local, torch, dist, dump_steps=1)
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            if wl["r_cutoff"] > 0:
                from oracle import oracle_c

                res = oracle_c.bench_cutoff(wl, steps=40, warmup=2)
            else:
                res = cpu_allpairs_numpy(wl, 2, 0)
            cpu = {"value": res["pairs"] / res["seconds"], "unit": "pair-interactions/s", "cores": res["cores"],
                   "kind": "port", "sample": res["sample"], "timesteps_per_s": res["md_steps"] / res["seconds"],
                   "host_nproc": os.cpu_count()}
        line = {
            "metric": "LJ pair-interactions/s", "value": tot_pairs / t_dev_max, "unit": "pair-interactions/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t_dev_max / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if fp64 else "f32 pair math, f64 state", "data": "synthetic",
            "timesteps_per_s": (1 if decomposed else world) * args.steps * md_steps / t_dev_max,
            "config": {"workload": wl["name_multi"] if decomposed else wl["name"], "particles_total": n if decomposed else n * world,
                       "particles_per_gpu": n // world if decomposed else n, "md_steps_per_step": md_steps,
                       "force_evals_per_md_step": fevals / (args.steps * md_steps), "precision": precision,
                       "k_boltzmann": 1.0, "l2": "flushed between timed steps (256 MiB write)",
                       "tolerances": "fp64 1e-10; fp32 pair math: energies 1e-5, per-particle forces 1e-5 of max(|F_i|, median|F|) on "
                                     "liquid states (the plain metric), of max(.., sum_j |f_ij|) near a lattice (tests/conftest.py)",
                       "parallelism": "replicas only (all-pairs does not shard)" if wl["r_cutoff"] <= 0 else
                       f"dd{world} grid {PROC_GRIDS.get(world)}", "list_rebuilds_in_timed_region": int(s1["list_rebuilds"] - s0["list_rebuilds"])},
            "e2e": {"value": tot_e2e_pairs / t_e2e_max, "unit": "pair-interactions/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "timesteps_per_s": (1 if decomposed else world) * args.steps * md_steps / t_e2e_max,
                    "ms_per_step": 1e3 * t_e2e_max / args.steps,
                    "d2h_gbs_per_rank": d2h / (t_e2e_max / args.steps) / 1e9, "raw_d2h_gbs_per_rank_all_ranks_copying": raw_d2h,
                    "host_numa_binding": numa,
                    "input": "the current (melted) state is uploaded from pinned memory before every batch",
                    "pipeline": "2 batches in flight (run_async/wait): D2H of step s overlaps the compute of step s+1; "
                                "wall clock over all K steps incl. the final drain; no L2 flush (working set >> L2)"},
            "gpu_launches": int(launches),
            "roofline": {"bound": "fp64" if fp64 else "fp32", "kernel": "all-pairs force (Newton's-third-law tile-pair kernel)" if wl["r_cutoff"] <= 0 else
                         ("neighbour-list force (k_nbr_force_f64)" if fp64 else "neighbour-list force (k_tile_force_f32: windows in shared memory)"),
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": achieved_tf / peak_tf if peak_tf else None, "traffic": force_traffic,
                         "traffic_source": force_traffic_src,
                         "flop_per_unique_pair": FLOP_PER_UNIQUE_PAIR[ndim], "unique_pairs_per_launch": f_pairs,
                         "ms_per_launch": f_ms,
                         "peak_source": "FMA micro-benchmark of the CUDA-core pipe measured in this run "
                                        "(MEASURED_PEAKS.json has no CUDA-core FP peak)",
                         "peak_datasheet": DATASHEET_TFLOPS["fp64" if fp64 else "fp32"],
                         "frac_of_datasheet": achieved_tf / DATASHEET_TFLOPS["fp64" if fp64 else "fp32"],
                         "sample_state": "melted liquid: the state the timed region ended in"},
            "roofline_hbm": roof_hbm,
            "parity": parity, "transport": transport,
            "cpu_baseline": cpu, "clocks": clocks, "energies_finite": energy_ok,
            "allpairs_configs1": extra,
            **strong,
            "host_wall_s_timed_region": t_wall,
        }
        if world == 8:
            line["strong_8m"] = {"same_as": "the headline line: at 8 GPUs the weak-scaling system is the 8 M-particle system",
                                 "pair_interactions_per_s": line["value"], "ms_per_md_step": line["ms_per_step"] / md_steps}
        print(json.dumps(line), flush=True)
    if ctx is not None:
        ctx.close()
    if world > 1:
        dist.destroy_process_group()
    if rank == 0 and parity is not None and not parity["ok"]:
        raise SystemExit("bench.py: parity against the CPU oracle FAILED: " + json.dumps(parity))


def strong_scaling_extra(per_row, label, rank, world, local, torch, dist, md=10, k_flush=4, dump_steps=0):
    """One FIXED-SIZE system (per_row^3 particles, vf 0.3, rc 2.5, skin 0.3, fp32 pair math) decomposed over `world` GPUs:
    BASELINE configs[3] "also report" (8 M particles on 1/2/4/8 GPUs) and configs[4] (64 M on 8).  The state is generated on
    the device (csrc/gen.cu: the reference lattice bit for bit, Maxwell-Boltzmann speeds sampled from the reference's table,
    isotropic directions) -- host generators would need minutes and 3 GB of host arrays per rank at 64 M.
    Returns device-resident throughput (rows buffered in the device ring every step, energies all-reduced per batch), the
    same with the rows of every step flushed to pinned host memory every k_flush steps, and optionally the time to scatter
    `dump_steps` rows of every rank into one HDF5 file."""
    import shutil

    from mdsea_b200 import _capi, parallel
    from mdsea_b200.gen import MONTECARLO_SPEEDRANGE, mb_cdf
    from mdsea_b200.helpers import get_dt, nsphere_volume

    wl = WORKLOADS["c3"]
    n = per_row ** 3
    L = (n * nsphere_volume(3, 0.5) / wl["vf"]) ** (1 / 3)
    grid = PROC_GRIDS[world]
    ctx = make_ctx(wl, n, L, 1e-3, local, "fp32", md, rank=rank, world=world, proc_grid=grid)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(_capi.Context.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, src=0)
        ctx.comm_init(bytes(uid.cpu().numpy().tobytes()))
    ctx.set_temperature(1.0)
    step = float(MONTECARLO_SPEEDRANGE[1] - MONTECARLO_SPEEDRANGE[0])
    mean_speed = ctx.generate_state(mb_cdf(1.0, 1.0, 1.0), step, seed=1234, isotropic=True)
    ctx.set_dt(get_dt(0.5, mean_speed))
    stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
    rows_dev = dict(r=None, v=None, pe=np.empty(md), ke=np.empty(md), temp=np.empty(md))

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(3):   # 30 steps: the lattice has started to melt, several list rebuilds behind us
        ctx.run(md, kb=1.0, record=True, rows=rows_dev)
    sync()
    s0 = ctx.stats()
    reps = 3
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(reps):
        ctx.run(md, kb=1.0, record=True, rows=rows_dev)
    e1.record(stream)
    sync()
    s1 = ctx.stats()
    t_dev = e0.elapsed_time(e1) * 1e-3
    pairs = float(s1["pair_evals_unique"] - s0["pair_evals_unique"])
    energies_ok = bool(np.isfinite(rows_dev["pe"]).all() and np.isfinite(rows_dev["ke"]).all())
    pe_last = float(rows_dev["pe"][-1])
    # rows of every step flushed to the host every k_flush steps, two batches in flight
    width = ctx.rank_capacity() if world > 1 else n
    pin = lambda shape, dt_=torch.float64: torch.empty(shape, dtype=dt_, pin_memory=True).numpy()
    sets = []
    for _ in range(2):
        rw = dict(r=pin((k_flush, 3, width)), v=pin((k_flush, 3, width)), pe=pin((k_flush,)), ke=pin((k_flush,)), temp=pin((k_flush,)))
        if world > 1:
            rw["ids"] = pin((k_flush, width), torch.int32)
            rw["n_owned"] = pin((k_flush,), torch.int64)
        sets.append(rw)
    ctx.run(k_flush, kb=1.0, record=True, rows=sets[0])
    sync()
    p0 = ctx.stats()["pair_evals_unique"]
    nb = 4
    t0 = time.perf_counter()
    in_flight = 0
    for b in range(nb):
        ctx.run(k_flush, kb=1.0, record=True, rows=sets[b % 2], asynchronous=True)
        in_flight += 1
        if in_flight == 2:
            ctx.wait(); in_flight -= 1
    while in_flight:
        ctx.wait(); in_flight -= 1
    torch.cuda.synchronize()
    t_flush = time.perf_counter() - t0
    sync()
    st = ctx.stats()
    flush_pairs = float(st["pair_evals_unique"] - p0)
    n_loc, n_gh = st["n_local"] if world > 1 else n, st["n_ghost"]
    row_bytes = k_flush * (48 + (4 if world > 1 else 0)) * n_loc
    # HDF5 dump of the last flushed rows: every rank scatters the columns of the particles it owns into ONE file
    dump = None
    if dump_steps > 0:
        from mdsea_b200.core import SysManager

        need = 2 * dump_steps * 3 * n * 8
        free = shutil.disk_usage(os.getcwd()).free
        if free < 2 * need:
            dump = {"skipped": f"{free / 1e9:.0f} GB free in the working directory, {need / 1e9:.0f} GB needed"}
        else:
            simid = f"bench_{label}"
            if rank == 0:
                sm = SysManager.new(simid=simid, ndim=3, num_particles=n, steps=dump_steps, vol_fraction=wl["vf"], radius_particle=0.5)
            sync()
            if rank != 0:
                sm = SysManager.load(simid)
            sm._open_datafile(shared=True)
            td = time.perf_counter()
            parallel.scatter_rows(sm, sets[(nb - 1) % 2], dump_steps, 0, rank)
            sm.dfile.flush() if hasattr(sm.dfile, "flush") else None
            sync()
            t_dump = time.perf_counter() - td
            ok = None
            if rank == 0:
                pe_file = sm.get_ds(sm.potenergy_dsname)
                ok = bool(np.allclose(pe_file[:dump_steps], sets[(nb - 1) % 2]["pe"][:dump_steps]))
            sm._close_datafile()
            sync()
            if rank == 0:
                sm.delete()
            dump = {"steps": dump_steps, "file_bytes": need, "seconds": t_dump, "gb_per_s": need / t_dump / 1e9,
                    "backend": "hdf5min (shared memory maps; every rank writes the columns it owns)", "scalars_read_back_ok": ok}
    ctx.close()
    agg = torch.tensor([pairs, flush_pairs, float(n_loc), float(n_gh), float(row_bytes)], dtype=torch.float64, device="cuda")
    mx = torch.tensor([t_dev, t_flush, float(n_gh) / max(n_loc, 1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(agg, op=dist.ReduceOp.SUM)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    agg, mx = agg.tolist(), mx.tolist()
    return {"workload": f"3D LJ liquid N={n} ({per_row}^3) vf=0.3 rc=2.5 skin=0.3, fixed total size over {world} GPU(s)",
            "particles_total": n, "n_gpus": world, "grid": list(grid), "md_steps_timed": reps * md,
            "pair_interactions_per_s": agg[0] / mx[0], "ms_per_md_step": 1e3 * mx[0] / (reps * md),
            "timesteps_per_s": reps * md / mx[0], "ghost_fraction_max": mx[2], "list_rebuilds": int(s1["list_rebuilds"] - s0["list_rebuilds"]),
            "energies": "pe / ke of every step all-reduced over the ranks once per batch", "energies_finite": energies_ok, "pe_last": pe_last,
            "flushed": {"rows_flushed_every": k_flush, "bytes_per_flush_all_ranks": agg[4], "pair_interactions_per_s": agg[1] / mx[1],
                        "ms_per_md_step": 1e3 * mx[1] / (nb * k_flush), "d2h_gbs_per_rank": (agg[4] / world) / (mx[1] / nb) / 1e9},
            "hdf5_dump": dump, "state": "generated on the device (gen.cu), 30 warm-up steps", "tile_lists": int(st["tile_lists"]),
            "transport": int(st.get("halo_transport", 0))}


_UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def ncu_traffic(summary, kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum (bytes per launch) of the first block of `kernel` in a tracked
    `ncu --set full` summary under profiles/ (written by profiles/summarize.py); None when there is no such capture."""
    try:
        lines = open(os.path.join(ROOT, "profiles", summary)).read().splitlines()
    except OSError:
        return None
    tot, inside, seen = 0.0, False, 0
    for ln in lines:
        if ln.startswith("== "):
            if inside:
                break
            inside = kernel in ln
        elif inside and ln.startswith("dram__bytes_"):
            f = ln.split()
            tot += float(f[1]) * _UNIT.get(f[2], 1.0)
            seen += 1
    return tot if seen == 2 else None


def allpairs_extra(local):
    """configs[1] (N=4096 all-pairs, fp64 and fp32) measured beside the default workload at N=1."""
    import torch

    from mdsea_b200 import _capi

    wl = WORKLOADS["c2"]
    n, L, r0, v0, dt = initial_state(wl["ndim"], wl["per_row"], wl["vf"])
    out = {"workload": wl["name"]}
    for precision in ("fp64", "fp32"):
        md = wl["md_steps"]
        ctx = make_ctx(wl, n, L, dt, local, precision, md)
        ctx.set_temperature(1.0)
        ctx.set_state(r0, v0)
        rows = dict(r=None, v=None, pe=np.empty(md), ke=np.empty(md), temp=np.empty(md))
        stream = torch.cuda.ExternalStream(ctx.stream(), device=torch.device("cuda", local))
        for _ in range(3):
            ctx.run(md, kb=1.0, record=True, rows=rows)
        torch.cuda.synchronize()
        s0 = ctx.stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            ctx.run(md, kb=1.0, record=True, rows=rows)
        e1.record(stream)
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) * 1e-3
        s1 = ctx.stats()
        ctx.set_profiling(True)
        q0 = ctx.stats()
        ctx.run(md, kb=1.0, record=True, rows=rows)
        q1 = ctx.stats()
        ctx.set_profiling(False)
        fe = q1["force_evals"] - q0["force_evals"]
        f_ms = (q1["ms_force"] - q0["ms_force"]) / max(fe, 1)
        peak = _capi.measure_fma_peak(local, fp64=precision == "fp64")
        ach = FLOP_PER_UNIQUE_PAIR[3] * (n * (n - 1) / 2) / (f_ms * 1e-3) / 1e12
        out[precision] = {"pair_interactions_per_s": (s1["pair_evals_unique"] - s0["pair_evals_unique"]) / t,
                          "timesteps_per_s": 5 * md / t, "force_ms_per_launch": f_ms,
                          "roofline": {"bound": precision, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak}}
        ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=None, choices=["fp64", "fp32"])
    ap.add_argument("--md-steps", type=int, default=None, help="MD timesteps per bench step (one C-ABI run call)")
    ap.add_argument("--evals", type=int, default=1, choices=[1, 2], help="force evaluations per MD step (2 = reference-faithful)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong", action="store_true", help="skip the strong-scaling ride-along systems (8 M / 64 M particles)")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity check of the benched configuration")
    ap.add_argument("--no-extra", action="store_true", help="skip the additional all-pairs (configs[1]) measurement at N=1")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    wl = WORKLOADS[args.workload]
    # a hung collective must not hold the box: dump every thread's Python stack and exit after MDSEA_BENCH_WATCHDOG seconds
    import faulthandler

    faulthandler.dump_traceback_later(float(os.environ.get("MDSEA_BENCH_WATCHDOG", "1500")), exit=True)
    if args.impl == "reference":
        run_reference_arm(args, wl)
    else:
        run_ours(args, wl)


if __name__ == "__main__":
    main()
