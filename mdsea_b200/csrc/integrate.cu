// K5: integrator kernels.  Every operation of apply_pbc / apply_hbc (simulator.py:156-179),
// algorithm_verlet / algorithm_simple (:530-544), apply_field (:201-205), the _quench velocity
// scaling (:192-195) and update_mean_ke (:241-245) is per COMPONENT, so the kernels run flat over
// the ndim*N elements of the SoA arrays: fully coalesced, no gathers, positions read once.
//
//   kick_drift : [boundary condition] ; [v += a dt/2] ; r += v dt        (before a force evaluation)
//   kick_finish: v += a dt/2 ; gravity ; quench scale ; sum v^2 ; row(r, v) into the device ring
//   (finalize) : fixed-order reduction of the per-block partial sums -> mean_pe, mean_ke, temp row, temperature
//                bookkeeping for the next step's quench: done by the LAST block of kick_finish (finalize.cuh)
#include "ctx.h"
#include "finalize.cuh"

#define IG_THREADS 256

__device__ __forceinline__ double md_sum_parts(const double *__restrict__ f, int nparts, size_t stride, size_t e) {
    double a = f[e];
    for (int p = 1; p < nparts; ++p) a += f[(size_t)p * stride + e];
    return a;
}

__global__ void __launch_bounds__(IG_THREADS)
k_kick_drift(double *__restrict__ r, double *__restrict__ v, const double *__restrict__ fsrc, int nparts, size_t stride,
             double *__restrict__ acc_out, size_t E, int do_bc, int do_kick, int pbc, double L, double radius,
             double restitution, double dt) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    double x = r[e], u = v[e];
    if (do_bc) {
        if (pbc) {  // one-shot strict wrap, simulator.py:165-166
            if (x < 0.0) x += L;
            if (x > L) x -= L;
        } else {  // clamp + reflect, simulator.py:172-179
            if (x - radius < 0.0) { x = radius; u *= -restitution; }
            if (x + radius > L) { x = L - radius; u *= -restitution; }
        }
    }
    if (do_kick) {
        const double a = md_sum_parts(fsrc, nparts, stride, e);
        if (acc_out) acc_out[e] = a;
        u += __dmul_rn(__dmul_rn(0.5, a), dt);  // v += 0.5 * acc * dt, evaluation order of simulator.py:540
    }
    x += __dmul_rn(u, dt);                      // r += v * dt, simulator.py:542
    r[e] = x;
    v[e] = u;
}

__global__ void __launch_bounds__(IG_THREADS)
k_apply_bc(double *__restrict__ r, double *__restrict__ v, size_t E, int pbc, double L, double radius, double restitution) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    double x = r[e];
    if (pbc) {
        if (x < 0.0) x += L;
        if (x > L) x -= L;
        r[e] = x;
    } else {
        double u = v[e];
        if (x - radius < 0.0) { x = radius; u *= -restitution; }
        if (x + radius > L) { x = L - radius; u *= -restitution; }
        r[e] = x;
        v[e] = u;
    }
}

__global__ void __launch_bounds__(IG_THREADS)
k_reduce_acc(const double *__restrict__ fsrc, int nparts, size_t stride, double *__restrict__ acc, size_t E) {
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) acc[e] = md_sum_parts(fsrc, nparts, stride, e);
}

// kick_finish of step k, optionally followed (DRIFT_NEXT) by the wrap + first half kick + drift of step k+1, which
// use the same accelerations when a(t+dt) is reused (periodic box, one force evaluation per step, no quench in the
// batch): r, v and the force partials are read once and one launch per step disappears.  The LAST block to publish
// its KE partial also does the step's reductions and bookkeeping (k_finalize), saving another launch.
template <bool DRIFT_NEXT>
__global__ void __launch_bounds__(IG_THREADS)
k_kick_finish(double *__restrict__ r, double *__restrict__ v, const double *__restrict__ fsrc, int nparts,
              size_t stride, double *__restrict__ acc_out, size_t E, size_t last_dim_start, double dt, double g_dt,
              StepScalars *__restrict__ sc, double *__restrict__ ke_part, double *__restrict__ r_row,
              double *__restrict__ v_row, double L, FinalizeArgs fin) {
    __shared__ double red[32];
    __shared__ bool   s_last;
    const size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    double v2 = 0.0;
    if (e < E) {
        const double a = md_sum_parts(fsrc, nparts, stride, e);
        if (acc_out) acc_out[e] = a;
        double u = v[e] + __dmul_rn(__dmul_rn(0.5, a), dt);  // simulator.py:544
        if (g_dt != 0.0 && e >= last_dim_start) u -= g_dt;   // v[-1] -= g * dt, simulator.py:205
        const double scale = sc->scale;                       // (T_target / T)^0.5, simulator.py:195
        if (scale != 1.0) u *= scale;
        v2 = u * u;
        double x = r[e];
        if (r_row) {
            r_row[e] = x;
            v_row[e] = u;
        }
        if (DRIFT_NEXT) {   // step k+1: apply_pbc (simulator.py:165-166), v += a dt/2 (:540), r += v dt (:542)
            if (x < 0.0) x += L;
            if (x > L) x -= L;
            u += __dmul_rn(__dmul_rn(0.5, a), dt);
            x += __dmul_rn(u, dt);
            r[e] = x;
        }
        v[e] = u;
    }
    const double b = md_block_sum(v2, red);
    if (threadIdx.x == 0) {
        ke_part[blockIdx.x] = b;
        __threadfence();
        s_last = atomicAdd(&sc->finish_blocks_done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        md_finalize_body(fin, sc, red);
        if (threadIdx.x == 0) sc->finish_blocks_done = 0u;
    }
}

__global__ void k_commit_pairs(StepScalars *sc) {
    sc->pairs_directed += sc->pairs_pending;
    sc->pairs_pending = 0ull;
}

// quench bookkeeping for the first step of a batch (uses the scalars left by the previous batch)
__global__ void k_prepare_first(StepScalars *sc, const double *quench_first, double kb) {
    double scale = 1.0, temp = sc->temp;
    if (quench_first) {
        for (int t = 0; t < 2; ++t) {
            const double target = quench_first[t];
            if (target == target) {
                if (sc->ke_valid) temp = (2.0 / 3.0) * sc->ke_prev / kb;
                if (temp != 0.0) scale *= sqrt(target / temp);
            }
        }
    }
    sc->temp = temp;
    sc->scale = scale;
}

// update_dists(radius=None) for small N (simulator.py:257-280): one thread per unique pair in
// itertools.combinations order.
__global__ void k_pair_dists(const double *__restrict__ r, int n, int ndim, long long P, int pbc, double L, double invL,
                             double *__restrict__ dists, double *__restrict__ units) {
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= P) return;
    // invert p = i*n - i(i+1)/2 + (j - i - 1)
    const double nn = (double)n - 0.5;
    long long i = (long long)floor(nn - sqrt(nn * nn - 2.0 * (double)p));
    while (i > 0 && i * n - i * (i + 1) / 2 > p) --i;
    while ((i + 1) * n - (i + 1) * (i + 2) / 2 <= p) ++i;
    const long long j = p - (i * n - i * (i + 1) / 2) + i + 1;
    double d[3], r2 = 0.0;
    for (int k = 0; k < ndim; ++k) {
        double x = __dsub_rn(r[(size_t)k * n + j], r[(size_t)k * n + i]);
        if (pbc) x = __dsub_rn(x, __dmul_rn(rint(__ddiv_rn(x, L)), L));
        d[k] = x;
        r2 = __dadd_rn(r2, __dmul_rn(x, x));
    }
    const double dist = __dsqrt_rn(r2);
    dists[p] = dist;
    for (int k = 0; k < ndim; ++k) units[(size_t)p * ndim + k] = __ddiv_rn(d[k], dist);
}

// ---------------------------------------------------------------------------------------------
// host wrappers
// ---------------------------------------------------------------------------------------------
static inline int ig_grid(const mdsea_ctx *c) { return md_div_up((long long)c->E, IG_THREADS); }

int md_apply_bc(mdsea_ctx *c) {
    k_apply_bc<<<ig_grid(c), IG_THREADS, 0, c->stream>>>(c->d_r, c->d_v, c->E, c->box.pbc, c->box.L, c->p.radius,
                                                         c->p.restitution);
    c->stats.kernel_launches += 1;
    MD_CUDA(c, cudaGetLastError());
    return MDSEA_OK;
}

// the same boundary condition over E entries of arbitrary arrays (the cut-off path keeps one padded row per dimension)
int md_apply_bc_range(mdsea_ctx *c, double *r, double *v, size_t E) {
    if (E == 0) return MDSEA_OK;
    // hard walls on the cut-off path: c->box.L is the period of the padded cell grid, the walls are those of the real box
    k_apply_bc<<<md_div_up((long long)E, IG_THREADS), IG_THREADS, 0, c->stream>>>(r, v, E, c->box.pbc, c->box.pbc ? c->box.L : c->p.boxlen,
                                                                                   c->p.radius, c->p.restitution);
    c->stats.kernel_launches += 1;
    MD_CUDA(c, cudaGetLastError());
    return MDSEA_OK;
}

// ---------------------------------------------------------------------------------------------
// centre of mass and radius of gyration (simulator.py:208-231): com = mean(r) per dimension, rog^2 = mean |r - com|^2,
// with the rint image of the separation when periodic.  Two passes of block sums, reduced in a fixed order.
// ---------------------------------------------------------------------------------------------
#define ROG_BLOCKS 256
__global__ void __launch_bounds__(256)
k_com_part(const double *__restrict__ r, long long n, size_t stride, int ndim, double *__restrict__ part /* [ndim][ROG_BLOCKS] */) {
    __shared__ double red[32];
    for (int d = 0; d < ndim; ++d) {
        double s = 0.0;
        for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
            s += r[(size_t)d * stride + i];
        s = md_block_sum(s, red);
        if (threadIdx.x == 0) part[d * ROG_BLOCKS + blockIdx.x] = s;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(256)
k_rog_part(const double *__restrict__ r, long long n, size_t stride, int ndim, const double *__restrict__ part, int pbc, double L,
           double *__restrict__ out /* [ROG_BLOCKS] */, double *__restrict__ com_out /* [ndim] */) {
    __shared__ double red[32];
    __shared__ double com[3];
    if (threadIdx.x < ndim) {
        double s = 0.0;
        for (int b = 0; b < ROG_BLOCKS; ++b) s += part[threadIdx.x * ROG_BLOCKS + b];
        com[threadIdx.x] = s / (double)n;
        if (blockIdx.x == 0) com_out[threadIdx.x] = com[threadIdx.x];
    }
    __syncthreads();
    double s = 0.0;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        double q = 0.0;
        for (int d = 0; d < ndim; ++d) {
            double x = r[(size_t)d * stride + i] - com[d];
            if (pbc) x -= rint(x / L) * L;
            q = fma(x, x, q);
        }
        s += q;
    }
    s = md_block_sum(s, red);
    if (threadIdx.x == 0) out[blockIdx.x] = s;
}

int md_com_rog(mdsea_ctx *c, const double *r, long long n, size_t stride, double *com, double *rog) {
    double *d_tmp = nullptr;
    const size_t words = (size_t)(3 + 1) * ROG_BLOCKS + 3;
    MD_CUDA(c, cudaMalloc(&d_tmp, sizeof(double) * words));
    double *d_part = d_tmp, *d_out = d_tmp + 3 * ROG_BLOCKS, *d_com = d_out + ROG_BLOCKS;
    k_com_part<<<ROG_BLOCKS, 256, 0, c->stream>>>(r, n, stride, c->ndim, d_part);
    k_rog_part<<<ROG_BLOCKS, 256, 0, c->stream>>>(r, n, stride, c->ndim, d_part, c->box.pbc, c->box.L, d_out, d_com);
    c->stats.kernel_launches += 2;
    double h[ROG_BLOCKS + 3];
    cudaError_t e = cudaMemcpyAsync(h, d_out, sizeof(double) * (ROG_BLOCKS + 3), cudaMemcpyDeviceToHost, c->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
    cudaFree(d_tmp);
    MD_CUDA(c, e);
    double s = 0.0;
    for (int b = 0; b < ROG_BLOCKS; ++b) s += h[b];
    for (int d = 0; d < c->ndim; ++d) com[d] = h[ROG_BLOCKS + d];
    *rog = sqrt(s / (double)n);
    return MDSEA_OK;
}

int md_kick_drift(mdsea_ctx *c, bool do_bc, bool do_kick, const double *fsrc, int nparts, size_t stride) {
    double *acc_out = (do_kick && fsrc != c->d_acc) ? c->d_acc : nullptr;
    md_prof_begin(c);
    k_kick_drift<<<ig_grid(c), IG_THREADS, 0, c->stream>>>(c->d_r, c->d_v, fsrc, nparts, stride, acc_out, c->E,
                                                           do_bc ? 1 : 0, do_kick ? 1 : 0, c->box.pbc, c->box.L,
                                                           c->p.radius, c->p.restitution, c->p.dt);
    md_prof_end(c, &c->stats.ms_integrate);
    c->stats.kernel_launches += 1;
    MD_CUDA(c, cudaGetLastError());
    return MDSEA_OK;
}

int md_reduce_acc(mdsea_ctx *c, const double *fsrc, int nparts, size_t stride) {
    k_reduce_acc<<<ig_grid(c), IG_THREADS, 0, c->stream>>>(fsrc, nparts, stride, c->d_acc, c->E);
    c->stats.kernel_launches += 1;
    MD_CUDA(c, cudaGetLastError());
    return MDSEA_OK;
}

// kick_finish of step `row` (+ the step's reductions); drift_next also starts step row+1 (see k_kick_finish)
int md_kick_finish(mdsea_ctx *c, const double *fsrc, int nparts, size_t stride, int row, int record, double g_dt,
                   const FinalizeArgs &fin, bool drift_next) {
    double *acc_out = (fsrc != c->d_acc) ? c->d_acc : nullptr;
    double *rr = record ? c->d_r_rows + (size_t)row * c->E : nullptr;
    double *vr = record ? c->d_v_rows + (size_t)row * c->E : nullptr;
    md_prof_begin(c);
    if (drift_next)
        k_kick_finish<true><<<ig_grid(c), IG_THREADS, 0, c->stream>>>(c->d_r, c->d_v, fsrc, nparts, stride, acc_out, c->E,
                                                                      (size_t)(c->ndim - 1) * c->n, c->p.dt, g_dt, c->d_sc, c->d_ke_part, rr,
                                                                      vr, c->box.L, fin);
    else
        k_kick_finish<false><<<ig_grid(c), IG_THREADS, 0, c->stream>>>(c->d_r, c->d_v, fsrc, nparts, stride, acc_out, c->E,
                                                                       (size_t)(c->ndim - 1) * c->n, c->p.dt, g_dt, c->d_sc, c->d_ke_part, rr,
                                                                       vr, c->box.L, fin);
    md_prof_end(c, &c->stats.ms_integrate);
    c->stats.kernel_launches += 1;
    MD_CUDA(c, cudaGetLastError());
    return MDSEA_OK;
}

FinalizeArgs md_finalize_args(mdsea_ctx *c, int row, int next_row_valid, double kb, long long step_index) {
    const double N = (double)c->n_global;
    FinalizeArgs A;
    A.pe_part = c->d_pe_part; A.n_pe = c->n_pe_part;
    A.ke_part = c->d_ke_part; A.n_ke = c->n_ke_part;
    A.pe_factor = (c->pot.fast == POTF_IDEAL) ? 0.0 : c->pot.lam_eps * 0.5 / N;
    A.ke_factor = 0.5 * c->p.mass / N;
    A.kb = kb;
    A.pe_rows = c->d_pe_rows; A.ke_rows = c->d_ke_rows; A.temp_rows = c->d_temp_rows;
    A.row = row;
    A.quench_next = next_row_valid ? c->d_quench + (size_t)(row + 1) * 2 : nullptr;
    A.step_index = step_index;
    return A;
}

int md_commit_pairs(mdsea_ctx *c) {
    k_commit_pairs<<<1, 1, 0, c->stream>>>(c->d_sc);
    c->stats.kernel_launches += 1;
    MD_CUDA(c, cudaGetLastError());
    return MDSEA_OK;
}

int md_prepare_first(mdsea_ctx *c, double kb) {
    k_prepare_first<<<1, 1, 0, c->stream>>>(c->d_sc, c->d_quench, kb);
    c->stats.kernel_launches += 1;
    MD_CUDA(c, cudaGetLastError());
    return MDSEA_OK;
}

int md_pair_dists(mdsea_ctx *c, double *d_dists, double *d_units) {
    const long long P = c->n * (c->n - 1) / 2;
    if (P == 0) return MDSEA_OK;
    k_pair_dists<<<md_div_up(P, 256), 256, 0, c->stream>>>(c->d_r, (int)c->n, c->ndim, P, c->box.pbc, c->box.L,
                                                           c->box.invL, d_dists, d_units);
    c->stats.kernel_launches += 1;
    MD_CUDA(c, cudaGetLastError());
    return MDSEA_OK;
}
