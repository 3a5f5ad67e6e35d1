// K5 (cut-off flavour): drift / finish integrators, quench bookkeeping for world > 1, small publish / peek / iota helpers.  Included by cells.cu only.
#pragma once

// ---------------------------------------------------------------------------------------------------
// K5 (cut-off flavour): particle-wise integrators.  Same per-component arithmetic as integrate.cu, plus
// the skin-displacement trigger, the packed position copy and the id-scatter of the dataset rows.
// ---------------------------------------------------------------------------------------------------
struct WallBC { double L, radius, restitution; };   // the REAL box of a hard-wall run (L below is the period of the cell grid)

// do_bc: 0 none, 1 periodic one-shot strict wrap (simulator.py:156-166), 2 hard walls: clamp + reflect (simulator.py:168-179)
__global__ void __launch_bounds__(256)
k_cut_kick_drift(double *__restrict__ r, double *__restrict__ v, const double *__restrict__ acc,
                 const double *__restrict__ rb, long long n, long long cap, int do_bc, int do_kick, int do_drift, double L,
                 double halfL, double dt, double trig2, int fp32, double inv_h, void *__restrict__ pos4,
                 StepScalars *__restrict__ sc, int tag, WallBC wall) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // all loads first (the stores below go through the same base pointers; see k_cut_kick_finish)
    double rr[3], vv[3], aa[3], bb[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        rr[d] = r[d * cap + i];
        vv[d] = v[d * cap + i];
        aa[d] = do_kick ? acc[d * cap + i] : 0.0;
        bb[d] = rb[d * cap + i];
    }
    double x[3], d2 = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double xx = rr[d], u = vv[d];
        if (do_bc == 1) {  // one-shot strict wrap, simulator.py:165-166
            if (xx < 0.0) xx += L;
            if (xx > L) xx -= L;
        } else if (do_bc == 2) {  // same two tests, in the same order, as apply_hbc (k_apply_bc, integrate.cu)
            if (xx - wall.radius < 0.0) { xx = wall.radius; u *= -wall.restitution; }
            if (xx + wall.radius > wall.L) { xx = wall.L - wall.radius; u *= -wall.restitution; }
        }
        if (do_kick) u += __dmul_rn(__dmul_rn(0.5, aa[d]), dt);
        if (do_drift) xx += __dmul_rn(u, dt);
        r[d * cap + i] = xx;
        if (do_kick || do_bc == 2) v[d * cap + i] = u;
        x[d] = xx;
        const double dd = md_minimg_exact(__dsub_rn(xx, bb[d]), L, halfL);
        d2 = __dadd_rn(d2, __dmul_rn(dd, dd));
    }
    if (do_drift && d2 > trig2) sc->rebuild_flag = tag;   // the skin trigger of THIS drift (tags: see StepScalars)
    if (fp32) ((uint4 *)pos4)[i] = make_uint4(md_to_fixed(x[0], inv_h), md_to_fixed(x[1], inv_h), md_to_fixed(x[2], inv_h), 0u);
    else ((double4 *)pos4)[i] = make_double4(x[0], x[1], x[2], 0.0);
}

// What the finish kernel needs to also START the next step (wrap + first half kick + drift, i.e. k_cut_kick_drift
// fused behind the second half kick): with a(t+dt) reused as the next step's a(t), v and acc are already in registers,
// so the fusion saves one launch and re-reading v / acc / re-writing v (96 B per particle and step).
struct NextDrift {
    const double *rb;     // build positions (skin trigger)
    double L, halfL, trig2, inv_h;
    void  *pos4;
    int    fp32;
    int    tag;           // trigger tag of the fused drift
    int   *host_flag;     // single GPU: mapped host word that receives (trigger fired) from the last block; else nullptr
};

template <bool FUSE>
__global__ void __launch_bounds__(256, FUSE ? 3 : 5)
k_cut_kick_finish(double *__restrict__ r, double *__restrict__ v, const double *__restrict__ acc,
                  const int *__restrict__ id, long long n, long long cap, long long row_stride, int *__restrict__ id_row,
                  double dt, double g_dt, StepScalars *__restrict__ sc, double *__restrict__ ke_part,
                  double *__restrict__ r_row, double *__restrict__ v_row, int guard_tag, FinalizeArgs fin,
                  const double *__restrict__ pe_force, int n_pe_force, double *__restrict__ pe_fold, NextDrift nd) {
    __shared__ double red[32];
    __shared__ bool   s_last;
    if (guard_tag && sc->rebuild_flag == guard_tag) return;
    double v2 = 0.0;
    const double scale = sc->scale;
    // grid-stride: a bounded number of blocks (8 per SM) keeps the partial-sum array and the number of same-address
    // atomics of the last-block protocol below small
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int    gid = id[i];
        // single GPU: rows are written in ORIGINAL particle order (core.py:236-238), scattered by id.
        // multi-rank: compact rows in slot order plus the ids; the host scatters them into the shared file.
        const long long col = id_row ? i : (long long)gid;
        if (id_row) id_row[i] = gid;
        // all loads first: the stores below go through the same base pointers (r, v), so the compiler would otherwise
        // order the next component's loads behind them and serialise three memory round trips per particle
        double rr[3], vv[3], aa[3], bb[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            rr[d] = r[d * cap + i];
            vv[d] = v[d * cap + i];
            aa[d] = acc[d * cap + i];
            bb[d] = FUSE ? nd.rb[d * cap + i] : 0.0;
        }
        double x[3], d2 = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double hk = __dmul_rn(__dmul_rn(0.5, aa[d]), dt);
            double u = vv[d] + hk;
            if (d == 2 && g_dt != 0.0) u -= g_dt;
            if (scale != 1.0) u *= scale;
            v2 += u * u;
            double xx = rr[d];
            if (r_row) {
                r_row[d * row_stride + col] = xx;
                v_row[d * row_stride + col] = u;
            }
            if (FUSE) {   // the next step: one-shot strict wrap, v += a dt / 2, r += v dt (same arithmetic as k_cut_kick_drift)
                if (xx < 0.0) xx += nd.L;
                if (xx > nd.L) xx -= nd.L;
                u += hk;
                xx += __dmul_rn(u, dt);
                r[d * cap + i] = xx;
                x[d] = xx;
                const double dd = md_minimg_exact(__dsub_rn(xx, bb[d]), nd.L, nd.halfL);
                d2 = __dadd_rn(d2, __dmul_rn(dd, dd));
            }
            v[d * cap + i] = u;
        }
        if (FUSE) {
            if (d2 > nd.trig2) sc->rebuild_flag = nd.tag;
            if (nd.fp32) ((uint4 *)nd.pos4)[i] = make_uint4(md_to_fixed(x[0], nd.inv_h), md_to_fixed(x[1], nd.inv_h), md_to_fixed(x[2], nd.inv_h), 0u);
            else ((double4 *)nd.pos4)[i] = make_double4(x[0], x[1], x[2], 0.0);
        }
    }
    const double b = md_block_sum(v2, red);
    // every block also folds its slice of the force kernel's PE partials, so that the last block reads gridDim.x values
    double pe_loc = 0.0;
    {
        const int chunk = (n_pe_force + (int)gridDim.x - 1) / (int)gridDim.x;
        const int k1 = min(n_pe_force, ((int)blockIdx.x + 1) * chunk);
        for (int k = (int)blockIdx.x * chunk + (int)threadIdx.x; k < k1; k += blockDim.x) pe_loc += pe_force[k];
    }
    const double bpe = md_block_sum(pe_loc, red);
    // the step's reductions and bookkeeping (finalize.cuh) are done by whichever block publishes
    // its partials last: no separate single-block launch on the critical path of every step
    if (threadIdx.x == 0) {
        ke_part[blockIdx.x] = b;
        pe_fold[blockIdx.x] = bpe;
        __threadfence();
        s_last = atomicAdd(&sc->finish_blocks_done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        md_finalize_body(fin, sc, red);
        if (threadIdx.x == 0) {
            sc->finish_blocks_done = 0u;
            // every block raised its trigger before bumping the counter: the flag of the fused drift is final here, so the
            // host learns it without a separate k_publish_flag launch
            if (FUSE && nd.host_flag) *nd.host_flag = *(volatile int *)&sc->rebuild_flag == nd.tag;
        }
    }
}

// world > 1: the temperature / quench-scale bookkeeping of md_finalize_body redone from the all-reduced mean_ke
__global__ void k_mr_quench_fix(StepScalars *__restrict__ sc, const double *__restrict__ ke_glob, const double *__restrict__ quench_next,
                                double kb, int redo) {
    const double mean_ke = *ke_glob;
    sc->ke_prev = mean_ke;
    if (!redo) return;
    // sc->temp was only overwritten by the finish kernel if the next step has a target, in which case it is
    // overwritten again here; otherwise it still holds self.temp
    double scale = 1.0, temp = sc->temp;
    if (quench_next) {
        for (int t = 0; t < 2; ++t) {
            const double target = quench_next[t];
            if (target == target) {
                temp = (2.0 / 3.0) * mean_ke / kb;
                if (temp != 0.0) scale *= sqrt(target / temp);
            }
        }
    }
    sc->temp = temp;
    sc->scale = scale;
}

__global__ void k_unsort(const double *__restrict__ src, const int *__restrict__ id, long long n, long long cap,
                         long long n_global, double *__restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int gid = id[i];
#pragma unroll
    for (int d = 0; d < 3; ++d) dst[d * n_global + gid] = src[d * cap + i];
}

// The host learns the skin trigger through a store into mapped pinned memory instead of a D2H memcpy: a copy would
// queue on the copy engine behind the bulk row transfers of the previous batch (run_async pipelining) and stall
// every step for milliseconds.
__global__ void k_publish_flag(const StepScalars *__restrict__ sc, int *__restrict__ host_flag, int tag) { *host_flag = sc->rebuild_flag == tag; }

// Small device -> host read-backs of the rebuild (counts, class boundaries) are stores into mapped pinned memory for
// the same reason: dst[i] = src[i * stride].
__global__ void k_peek(const int *__restrict__ src, long long stride, int n, int *__restrict__ host_dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) host_dst[i] = src[(long long)i * stride];
}

__global__ void k_iota_from(int *a, long long n, int first) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = first + (int)i;
}
__global__ void k_iota(int *a, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (int)i;
}
