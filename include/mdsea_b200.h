/*
 * mdsea_b200.h -- C ABI of the B200 (sm_100a) implementation of mdsea's
 * ContinuousPotentialSolver step loop.
 *
 * The upstream reference (tpvasconcelos/mdsea) is pure Python/NumPy and has no FFI of
 * its own; the entry points below are what a Python maintainer would bind with ctypes
 * to replace the bodies of the methods cited next to each declaration (paths relative
 * to the reference tree).  INTEGRATION.md shows that binding.
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / numpy types cross this boundary.
 *   - every array of particle data is float64, SoA, shape (ndim, N) row-major -- the
 *     layout of the reference's r_vec / v_vec / acc (mdsea/simulator.py:54-70) and of one
 *     HDF5 row of r_coords / v_coords (mdsea/core.py:236-238).
 *   - host pointers unless stated otherwise; the caller owns every host buffer, the
 *     library owns all device memory, streams and events (freed by mdsea_b200_destroy).
 *   - every call returns MDSEA_OK (0) or a negative mdsea_status; the message for the
 *     most recent failure on a context is available from mdsea_b200_last_error.
 *   - a context is NOT thread-safe: one host thread drives one context (the reference
 *     is single-threaded and synchronous).  Multi-GPU = one context per rank/process.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails
 *     with MDSEA_ERR_CUDA.
 */
#ifndef MDSEA_B200_H
#define MDSEA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDSEA_B200_ABI_VERSION 2

typedef enum {
    MDSEA_OK = 0,
    MDSEA_ERR_BAD_ARG = -1,  /* invalid parameter / NULL pointer / unsupported combination */
    MDSEA_ERR_CUDA = -2,     /* CUDA runtime error (no device, launch failure, OOM)       */
    MDSEA_ERR_NCCL = -3,     /* NCCL error in the halo exchange / energy allreduce        */
    MDSEA_ERR_STATE = -4,    /* call out of order (e.g. run before set_state)             */
    MDSEA_ERR_NONFINITE = -5,/* reserved: NaN/Inf in the energy reduction (see first_nonfinite_step)   */
    MDSEA_ERR_CAPACITY = -6, /* neighbour / ghost / migration buffer capacity exceeded   */
    MDSEA_ERR_UNSUPPORTED = -7
} mdsea_status;

/* pair potential: the reference's bounded-Mie family, mdsea/potentials.py:15-64 */
typedef enum {
    MDSEA_POT_IDEAL = 0,      /* pf_ideal / ff_ideal  (potentials.py:33-35,62-64)            */
    MDSEA_POT_BOUNDED_MIE = 1 /* pf_/ff_boundedmie; mie = a 0; lennardjones = m 12, n 6, a 0 */
} mdsea_pot_kind;

typedef enum { MDSEA_ALGO_VERLET = 0, MDSEA_ALGO_SIMPLE = 1 } mdsea_algorithm; /* simulator.py:530-544 */
typedef enum { MDSEA_FP64 = 0, MDSEA_FP32 = 1 } mdsea_precision;

typedef struct mdsea_ctx mdsea_ctx;

typedef struct {
    int32_t abi_version;       /* MDSEA_B200_ABI_VERSION */
    int32_t ndim;              /* 1, 2 or 3 (SysManager.NDIM, core.py:68)                      */
    int64_t num_particles;     /* global N (SysManager.NUM_PARTICLES)                           */
    double  boxlen;            /* SysManager.LEN_BOX (core.py:113-115)                          */
    double  mass;              /* SysManager.MASS                                               */
    double  radius;            /* SysManager.RADIUS_PARTICLE (hard-wall clamp, simulator.py:168-179) */
    double  dt;                /* integration step (simulator.py:86-87)                         */
    int32_t pbc;               /* 1: apply_pbc (simulator.py:156-166); 0: apply_hbc (:168-179)  */
    int32_t algorithm;         /* mdsea_algorithm                                               */
    double  restitution;       /* restitution_coeff; only used by the hard-wall reflection      */
    int32_t pot_kind;          /* mdsea_pot_kind                                                */
    int32_t precision;         /* mdsea_precision: arithmetic of the PAIR math; state stays f64 */
    double  epsilon, sigma;    /* Potential.kwargs (potentials.py:109-122)                      */
    double  mie_m, mie_n, mie_a;
    double  r_cutoff;          /* <= 0: all pairs, no cutoff (the reference's only working mode).
                                  > 0 : EXTENSION -- plain truncation, pair counted iff dist < r_cutoff
                                  (strict, as simulator.py:283), cell list + Verlet list: 3-D boxes with >= 3
                                  cells of side r_cutoff + skin per edge.  pbc = 0 (hard walls): the
                                  box is embedded in a periodic cell grid of period boxlen + 2 (r_cutoff + skin),
                                  which is the grid mdsea_b200_get_cells reports; other boxes: all-pairs kernels */
    double  skin;              /* Verlet skin; list radius = r_cutoff + skin                    */
    int32_t force_evals_per_step; /* 2: evaluate a(r) twice per verlet step like simulator.py:540,544;
                                     1: reuse a(t+dt) as the next step's a(t) (pbc only; identical
                                        positions modulo the wrap, differences ~1e-15)            */
    int32_t ring_rows;         /* rows (steps) buffered on the device between flushes (k); >= 1 */
    int32_t device;            /* CUDA device ordinal                                           */
    int32_t max_neighbors;     /* per-particle neighbour capacity for the cutoff path; 0 = auto */
    /* spatial domain decomposition (cutoff path only).  world == 1: single GPU.              */
    int32_t rank, world;
    int32_t proc_grid[3];      /* ranks along x,y,z; product == world                           */
    int32_t reserved[8];
} mdsea_params;

/* per-call inputs of mdsea_b200_run */
typedef struct {
    int32_t nsteps;               /* 1..ring_rows                                                   */
    int32_t record_coords;        /* 1: buffer r/v rows of every step on the device (update_files, simulator.py:143-150) */
    double  k_boltzmann;          /* Config.k_boltzmann at call time (simulator.py:238)            */
    double  gravity_acceleration; /* Config.gravity_acceleration if gravity else 0 (simulator.py:201-205) */
    const double *quench_targets; /* NULL or [nsteps][2]: target temperatures of the (up to two) _quench
                                     calls of each step in call order (isothermal first, then the
                                     QUENCH_STEP entry; simulator.py:185-195); NaN = no quench     */
} mdsea_run_args;

/* caller-owned host buffers receiving the rows of one run() call, in HDF5 row layout
 * (core.py:236-242).  r/v may be NULL: with record_coords the rows are then still written to the
 * device ring (the on-device buffering of the dataset writes) but not copied to the host.
 * Pinned memory makes the copies asynchronous with respect to the host. */
typedef struct {
    double *r;    /* [nsteps][ndim][N]  -> r_coords rows */
    double *v;    /* [nsteps][ndim][N]  -> v_coords rows */
    double *pe;   /* [nsteps]           -> pot_energy    */
    double *ke;   /* [nsteps]           -> kin_energy    */
    double *temp; /* [nsteps]           -> temp          */
    /* world > 1 only: each rank returns the rows of the particles it owns, compact, in its slot order:
     * r / v are [nsteps][ndim][row_stride], ids [nsteps][row_stride] holds the particle id of every column,
     * n_owned[k] the number of valid columns of row k.  The host scatters column j of row k to
     * r_coords[step k][:, ids[k][j]] (every particle is owned by exactly one rank).  pe / ke are global. */
    int32_t *ids;
    int64_t *n_owned;
    int64_t  row_stride; /* >= mdsea_b200_rank_capacity(ctx) */
} mdsea_rows;

typedef struct {
    int64_t steps;               /* steps advanced since create                                  */
    int64_t force_evals;         /* force evaluations launched                                   */
    int64_t pair_evals_unique;   /* unique (i<j) pair interactions inside the cutoff, summed over evals */
    int64_t pair_evals_executed; /* directed pair evaluations the kernels executed (list entries / N*N) */
    int64_t kernel_launches;     /* kernels of this library launched since create                */
    int64_t list_rebuilds;       /* cell binning + neighbour list rebuilds                        */
    int64_t n_local, n_ghost;    /* owned / ghost particles on this rank (cutoff path)            */
    int64_t nbr_entries;         /* directed entries in the current neighbour list                */
    double  ms_force, ms_integrate, ms_rebuild, ms_comm; /* CUDA-event time sums when profiling is on */
    /* cut-off path, fp32 pair math: 1 while the current neighbour lists are tile lists (16-bit offsets into per-block
     * windows that the force kernel stages in shared memory), 0 for the 32-bit per-particle lists */
    int64_t tile_lists;
    int64_t tile_window_max;     /* largest window of the current list generation, entries of 16 bytes        */
    int64_t tile_fallbacks;      /* list generations that fell back to the per-particle lists (window did not fit) */
    /* data plane of the domain decomposition: 0 = single rank, 1 = peer memory (CUDA IPC stores over NVLink + arrival
     * counters), 2 = ncclSend/ncclRecv (MDSEA_HALO=nccl, or a peer mapping failed on some rank: reported on stderr),
     * 3 = in-process (host threads on one GPU, tests) */
    int64_t halo_transport;
} mdsea_stats;

const char *mdsea_b200_version(void);

/* replaces _BaseSimulator.__post_init__ device-side state (simulator.py:51-97); the O(N^2)
 * pair table of :122-137 is deliberately not reproduced. */
int mdsea_b200_create(const mdsea_params *params, mdsea_ctx **out);
void mdsea_b200_destroy(mdsea_ctx *ctx);
const char *mdsea_b200_last_error(const mdsea_ctx *ctx); /* ctx may be NULL: last create() error */

/* r_vec / v_vec upload (simulator.py:54-59): (ndim, N) float64 host arrays, GLOBAL particle order.
 * With world > 1 every rank passes the full arrays and keeps the particles it owns. */
int mdsea_b200_set_state(mdsea_ctx *ctx, const double *r_soa, const double *v_soa);
/* Initial conditions generated on the device instead of uploaded (opt-in; replaces PosGen.simplecubic + VelGen.mb,
 * mdsea/gen.py:98-123,144-156 as called by simulator.py:54-59).  The lattice is the reference's bit for bit; the
 * velocities are Maxwell-Boltzmann by inverse-CDF sampling of `speed_cdf` (the reference's table mb_cdf on the grid
 * arange(0, 50, speed_step), gen.py:27-37; NULL = all zero, the temp == 0 case) with the reference's direction
 * factors (3-D: vy, vz scaled by the product of sin(theta_0) over all particles, gen.py:83-92) or, with
 * isotropic != 0, uniformly distributed directions.  A counter-based generator keyed by (seed, particle id) replaces
 * the global NumPy stream, so the velocities are NOT the reference's for the same seed.  *mean_speed (may be NULL)
 * receives mean_i |v_i| over all particles, the input of get_dt (helpers.py:89-100).  With world > 1 every rank
 * generates its slice and the first rebuild routes it, like set_state. */
int mdsea_b200_generate_state(mdsea_ctx *ctx, const double *speed_cdf, int64_t n_cdf, double speed_step, uint64_t seed,
                              int32_t isotropic, double *mean_speed);
/* time step of the following run calls (simulator.py:86-87 derives it from the initial velocities) */
int mdsea_b200_set_dt(mdsea_ctx *ctx, double dt);
/* replaces self.apply_bc() = apply_pbc / apply_hbc (mdsea/simulator.py:90,156-179) applied to the state on the device:
 * periodic: the strict one-shot wrap (x < 0 -> x + L, x > L -> x - L); hard walls: clamp to [radius, L - radius] and
 * v *= -restitution for the entries that crossed.  The next force evaluation sees the moved positions. */
int mdsea_b200_apply_bc(mdsea_ctx *ctx);
/* replaces the `com` and `rog` properties (mdsea/simulator.py:207-231): com[ndim] = mean position, *rog = sqrt(mean |r - com|^2)
 * (separations folded with rint(d / L) when the box is periodic).  Single-rank contexts. */
int mdsea_b200_com_rog(mdsea_ctx *ctx, double *com, double *rog);
/* download of r_vec, v_vec, acc in original particle order; any pointer may be NULL.
 * With world > 1 each rank fills only the entries of particles it owns (others untouched)
 * and returns their ids in `owned_ids` (capacity N, may be NULL) / count in *n_owned. */
int mdsea_b200_get_state(mdsea_ctx *ctx, double *r_soa, double *v_soa, double *acc_soa,
                         int64_t *owned_ids, int64_t *n_owned);

/* replaces `for step in simrange: advance(); update_files()` (simulator.py:583-586) for
 * args->nsteps consecutive steps: apply_bc, algorithm_verlet/simple, apply_field,
 * apply_special, update_energies (simulator.py:560-568) and the buffering of the five
 * dataset rows; then copies the buffered rows to `rows`.  Blocks until the copies landed. */
int mdsea_b200_run(mdsea_ctx *ctx, const mdsea_run_args *args, const mdsea_rows *rows);
/* same, but returns once the batch is enqueued and its row copies are under way.  Up to TWO batches can be in
 * flight: the library alternates between two device ring sets, so the device->host copies of one batch overlap
 * the compute of the next (a third run_async first completes the oldest batch).  mdsea_b200_wait completes the
 * OLDEST batch in flight -- its `rows` buffers are valid afterwards -- and is a no-op when nothing is pending.
 * mdsea_b200_set_state may be called while batches are in flight (it is ordered after their compute and does
 * not touch the rings); every other call drains all pending batches first.  The caller keeps the `rows`
 * buffers (pinned host memory for true overlap) alive and untouched until the batch has been waited for. */
int mdsea_b200_run_async(mdsea_ctx *ctx, const mdsea_run_args *args, const mdsea_rows *rows);
int mdsea_b200_wait(mdsea_ctx *ctx);

/* replaces _BaseSimulator.update_acc (simulator.py:295-311) on the current positions: evaluates
 * the accelerations, keeps them as ctx state (sim.acc) and optionally returns them and the mean
 * potential energy of update_mean_pe (simulator.py:247-250). acc_soa / mean_pe may be NULL. */
int mdsea_b200_update_acc(mdsea_ctx *ctx, double *acc_soa, double *mean_pe);

/* replaces _BaseSimulator.update_dists(radius=None) (simulator.py:257-280) for small N:
 * dists[P], drunits[P][ndim], P = N(N-1)/2 in itertools.combinations order (simulator.py:122-126). */
int mdsea_b200_update_dists(mdsea_ctx *ctx, double *dists, double *drunits);

/* cutoff-path test hooks (bit-exact contracts):
 *   cells:     cell_of[i] = canonical linear cell id of particle i; order[k] = id of the k-th
 *              particle after the stable (cell id, particle id) sort.
 *   neighbors: CSR over ORIGINAL particle ids, each row ascending ids j != i with
 *              |r_i - r_j|^2_minimg < (r_cutoff+skin)^2 at the last rebuild.
 * Two-call pattern: pass idx == NULL to obtain *n_entries first. offsets has N+1 entries. */
int mdsea_b200_get_cells(mdsea_ctx *ctx, int32_t *cell_of, int32_t *order, int32_t ncell_dim[3]);
int mdsea_b200_get_neighbors(mdsea_ctx *ctx, int64_t *n_entries, int64_t *offsets, int32_t *idx);
/* diagnostics of the tile lists (fp32 cut-off path): per block of 128 slots the number of window entries and of
 * contiguous slot ranges that the force kernel stages in shared memory every step.  Pass NULL arrays to obtain
 * *n_blocks first; *n_blocks == 0 while the current list generation is not in tile format. */
int mdsea_b200_get_tile_windows(mdsea_ctx *ctx, int64_t *n_blocks, int32_t *n_win, int32_t *n_rng);

/* seeds self.temp / temp_init (simulator.py:79-80) -- the value written to the `temp` dataset until a
 * quench changes it; reset_ke != 0 puts mean_ke back to None (simulator.py:75). */
int mdsea_b200_set_temperature(mdsea_ctx *ctx, double temp, int32_t reset_ke);

int mdsea_b200_get_stats(mdsea_ctx *ctx, mdsea_stats *out);
/* index of the first step whose energies were NaN/Inf, or -1.  The reference writes NaN rows
 * silently; this library does the same and lets the host decide. */
int mdsea_b200_first_nonfinite_step(mdsea_ctx *ctx, int64_t *step);
/* 1: record CUDA-event timings per kernel family into mdsea_stats.ms_* (adds event overhead) */
int mdsea_b200_set_profiling(mdsea_ctx *ctx, int32_t on);
/* the CUDA stream the kernels are launched on (cudaStream_t as void*), for event timing by the caller */
void *mdsea_b200_stream(mdsea_ctx *ctx);

/* multi-GPU plumbing: rank 0 obtains an id (128 bytes), the host side broadcasts it
 * (torch.distributed / file / env), every rank calls comm_init before set_state. */
int mdsea_b200_nccl_unique_id(void *out128);
int mdsea_b200_comm_init(mdsea_ctx *ctx, const void *unique_id128);
/* in-process transport: `world` contexts of ONE process, each driven by its own host thread, exchange through
 * device-to-device copies behind host barriers (all contexts of a group pass the same group_key).  Lets the
 * multi-rank code run on a single GPU; the collective calls (set_state, run, update_acc, get_cells,
 * get_neighbors, destroy) must then be issued by all ranks concurrently, one thread per rank. */
int mdsea_b200_comm_init_local(mdsea_ctx *ctx, int64_t group_key);
/* per-rank slot capacity (owned + ghost particles) of the cutoff path; the row_stride to allocate rows with */
int64_t mdsea_b200_rank_capacity(mdsea_ctx *ctx);

/* Analyser-side reductions (stand-alone, no context needed): streams the `v_coords` dataset
 * v_coords[steps][ndim][n] (host memory, the HDF5 row layout) through the GPU and returns what mdsea's Analyser
 * derives from it (mdsea/analytics.py:53-56, quicker.py:44-46): speeds[steps][n] (may be NULL), their global mean
 * and population standard deviation (maxspeed = mean + 3 std) and the largest speed.  speeds are bit-identical to
 * quicker.norm; mean / std agree to ~1e-14 (different summation order). */
int mdsea_b200_analyse_speeds(int32_t device, const double *v_coords, int64_t steps, int32_t ndim, int64_t n,
                              double *speeds, double *mean, double *stddev, double *max_speed);

/* device micro-benchmarks used by bench.py for the roofline denominators that
 * MEASURED_PEAKS.json does not hold: dependent-free FMA throughput of the fp32 / fp64
 * CUDA-core pipes in TFLOP/s (FMA = 2 flop). */
int mdsea_b200_measure_fma_peak(int32_t device, int32_t fp64, double *tflops);

#ifdef __cplusplus
}
#endif
#endif /* MDSEA_B200_H */
