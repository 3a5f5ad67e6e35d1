// K3: sliced-ELLPACK neighbour-list layout and the Verlet list build kernels.  Included by cells.cu only.
#pragma once

// Neighbour-list layout: sliced ELLPACK.  Particles are grouped in tiles of 32 consecutive slots (= one warp of
// the force kernel); tile t owns max_nbr rows of 32 ints, entry (k, lane) at t*max_nbr*32 + k*32 + lane.  The
// force kernel reads one 128-byte row per k (coalesced); the build kernels' stores of one warp all land in the
// same ~15 KB tile, so partially written sectors merge in L2 instead of being evicted half empty.
// (A quad-sliced layout with 8 lanes per particle was measured in r01: 1.6x fewer L1 wavefronts per gather, but
// 1.7x more instructions per pair -- 208 us against 136 us for this one; see profiles/README.md.)
__host__ __device__ __forceinline__ size_t md_nbr_row_base(long long i, int max_nbr) {
    return (size_t)(i >> 5) * (size_t)max_nbr * 32u + (size_t)(i & 31);
}
__host__ __device__ __forceinline__ unsigned md_nbr_off(unsigned k) { return k * 32u; }

// ---------------------------------------------------------------------------------------------------
// K3: Verlet neighbour list.
//
// Two kernels produce the SAME canonical list (j != i with canonical float64 r^2 < (rc+skin)^2):
//   k_nbr_build_warp  (default)  one warp per tile of 32 consecutive slots.  The warp walks the UNION of the 27-cell
//                     stencils of its particles (they share an x-row of cells, so the union is 9 short slot ranges),
//                     stages 32 candidates at a time in shared memory as floats relative to a common origin (the
//                     fixed-point subtraction is exact and wraps to the minimum image for free) and every lane tests
//                     the broadcast candidates against its own particle: no divergence, no per-lane gathers.  In-range
//                     candidates are collected in a 32-bit mask per lane and appended after each chunk.
//   k_nbr_build       one thread per particle over its own 27 cells (MDSEA_NBR_BUILD=thread; kept as the A/B baseline).
// Both decide in float with a guard band around the list radius; a particle that saw a candidate inside the band is
// redone with the canonical float64 test for every candidate, so the list is the canonical one bit for bit.
// ---------------------------------------------------------------------------------------------------
template <bool EXACT>
__device__ __forceinline__ void nb_scan_range(int j0, int j1, const uint4 qi, long long i, const double *__restrict__ r,
                                              const uint4 *__restrict__ q, long long cap, double L, double halfL, double rl2,
                                              float rl2s_lo, float rl2s_hi, int max_nbr, int *__restrict__ row0, int &cnt,
                                              bool &band) {
    if (EXACT) {
        const double xi = r[i], yi = r[cap + i], zi = r[2 * cap + i];
        for (int j = j0; j < j1; ++j) {
            const double ddx = md_minimg_exact(__dsub_rn(r[j], xi), L, halfL);
            const double ddy = md_minimg_exact(__dsub_rn(r[cap + j], yi), L, halfL);
            const double ddz = md_minimg_exact(__dsub_rn(r[2 * cap + j], zi), L, halfL);
            if (md_r2_canonical(ddx, ddy, ddz) < rl2 && j != (int)i) {
                if (cnt < max_nbr) row0[md_nbr_off((unsigned)cnt)] = j;
                ++cnt;
            }
        }
    } else {
        // clamped write cursor: one predicated store per candidate, no bounds branch (an over-full row keeps
        // overwriting its last slot; cnt still counts, the caller reports the overflow)
        const unsigned k_max = (unsigned)(max_nbr - 1);
#pragma unroll 4
        for (int j = j0; j < j1; ++j) {
            const uint4 qj = q[j];
            const float fx = (float)(int)(qj.x - qi.x), fy = (float)(int)(qj.y - qi.y), fz = (float)(int)(qj.z - qi.z);
            const float r2f = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
            const bool in = r2f < rl2s_lo;
            band = band || (!in && r2f < rl2s_hi);
            const bool st = in && (j != (int)i);
            const unsigned off = md_nbr_off(min((unsigned)cnt, k_max));
            if (st) row0[off] = j;
            cnt += st ? 1 : 0;
        }
    }
}

template <bool EXACT>
__device__ __forceinline__ int nb_particle(long long i, const uint4 qi, int c, const double *__restrict__ r,
                                           const uint4 *__restrict__ q, long long cap, const int *__restrict__ cell_start,
                                           const int *__restrict__ cell_end, int nc, double L, double halfL, double rl2,
                                           float rl2s_lo, float rl2s_hi, int max_nbr, int *__restrict__ row0, bool &band) {
    const int cx = c % nc, cy = (c / nc) % nc, cz = c / (nc * nc);
    const int xm = (cx + nc - 1) % nc, xp = (cx + 1) % nc;
    int cnt = 0;
    for (int dz = -1; dz <= 1; ++dz) {
        const int z = (cz + dz + nc) % nc;
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = (cy + dy + nc) % nc;
            const int row = (z * nc + y) * nc;
            // the three x-adjacent cells: merge slot ranges that happen to be contiguous (the usual case away from
            // the periodic wrap and from the owned/ghost boundary)
            int cs = cell_start[row + xm], ce = cell_end[row + xm];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int cc = row + (t == 0 ? cx : xp);
                const int s1 = cell_start[cc], e1 = cell_end[cc];
                if (ce == cs) {            // nothing pending
                    cs = s1; ce = e1;
                } else if (e1 != s1) {
                    if (ce == s1) ce = e1; // contiguous: extend
                    else {
                        nb_scan_range<EXACT>(cs, ce, qi, i, r, q, cap, L, halfL, rl2, rl2s_lo, rl2s_hi, max_nbr, row0, cnt, band);
                        cs = s1; ce = e1;
                    }
                }
            }
            nb_scan_range<EXACT>(cs, ce, qi, i, r, q, cap, L, halfL, rl2, rl2s_lo, rl2s_hi, max_nbr, row0, cnt, band);
        }
    }
    return cnt;
}

__global__ void __launch_bounds__(NB_THREADS)
k_nbr_build(const double *__restrict__ r, const uint4 *__restrict__ q, long long i_first, long long n_own, long long cap,
            const int *__restrict__ cell, const int *__restrict__ cell_start, const int *__restrict__ cell_end, int nc,
            double L, double halfL, double rl2, float rl2s_lo, float rl2s_hi, int max_nbr, int *__restrict__ nbr,
            int *__restrict__ nbr_cnt, StepScalars *__restrict__ sc) {
    const long long i = i_first + (long long)blockIdx.x * blockDim.x + threadIdx.x;   // lists of the owned slots [i_first, n_own)
    if (i >= n_own) return;
    const uint4 qi = q[i];
    const int c = cell[i];
    int *row0 = nbr + md_nbr_row_base(i, max_nbr);
    bool band = false;
    int cnt = nb_particle<false>(i, qi, c, r, q, cap, cell_start, cell_end, nc, L, halfL, rl2, rl2s_lo, rl2s_hi, max_nbr, row0, band);
    if (__builtin_expect(band, 0))
        cnt = nb_particle<true>(i, qi, c, r, q, cap, cell_start, cell_end, nc, L, halfL, rl2, rl2s_lo, rl2s_hi, max_nbr, row0, band);
    if (cnt > max_nbr) {
        sc->overflow_flag = 1;
        cnt = max_nbr;
    }
    nbr_cnt[i] = cnt;
    for (int k = cnt; k < ((cnt + NBG - 1) / NBG) * NBG; ++k) row0[k * 32] = (int)i;  // padding, masked by j != i
}

// ---- warp-cooperative build ------------------------------------------------------------------------
#define NBW_WARPS 4      // warps (= tiles of 32 slots) per block
#define NBW_RUN   6      // at most this many x-consecutive cells of i-particles per pass (bounds the float error)
// NBW_G (template parameter): lane groups per warp with their own candidate streams (32 / NBW_G consecutive particles
// each).  More groups = shorter streams (fewer tests) but more staging and more shared-memory wavefronts per broadcast
// LDS.128 (4 groups: 4 wavefronts, one address per warp: 2).  MDSEA_NBW_GROUPS selects 4 / 2 / 1 for A/B.
// Offsets from the common origin are plain floats (no wrap), so they only reproduce the minimum image while the
// candidate window stays inside half a period: (run + 2) cells < nc in x, 3 cells < nc in y and z.  The host
// therefore passes run_max = min(NBW_RUN, nc - 3) and uses the thread-per-particle kernel when nc == 3.

// test one staged chunk (<= 32 candidates, padded with far-away dummies) against this lane's particle.
// Expanded form r^2 - rl^2 = |c|^2 + (|p|^2 - rl^2) - 2 p.c: the candidate carries w = |c|^2 and the lane k = |p|^2 - rl^2
// (both rounded ONCE from float64), so a test costs 3 FFMA + 1 FADD on the FMA pipe instead of 3 FADD + 3 FFMA -- the
// kernel is bound by that pipe.  Offsets stay below ~4 cells in x and 1.5 cells in y / z, which bounds the rounding
// error of the chain (largest term last) by ~5e-6 rl^2; the guard band (2e-5 rl^2) covers it.
template <bool EXPANDED>
__device__ __forceinline__ unsigned nbw_test_chunk(const float4 *__restrict__ cand, int ncand, float a, float b, float cc, float k,
                                                   float neg_bw, float &bandmin) {
    // EXPANDED: (a, b, cc) = -2 p, k = |p|^2 - rl^2; direct form: (a, b, cc) = p, k = -rl^2
    unsigned m = 0u;
    for (int t0 = 0; t0 < ncand; t0 += 8) {
        unsigned m8 = 0u;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 cj = cand[t0 + u];   // broadcast LDS.128
            float d;                          // r^2 - rl^2
            if (EXPANDED) {
                d = fmaf(a, cj.x, fmaf(b, cj.y, fmaf(cc, cj.z, k))) + cj.w;
            } else {
                const float dx = cj.x - a, dy = cj.y - b, dz = cj.z - cc;
                d = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, k)));
            }
            if (d < neg_bw) m8 |= 1u << u;
            bandmin = fminf(bandmin, fabsf(d));
        }
        m |= m8 << t0;
    }
    return m;
}

// staged candidate: float offsets from the pass origin (+ their squared norm, one rounding, for the expanded form)
template <bool EXPANDED>
__device__ __forceinline__ float4 nbw_stage(const uint4 qj, uint32_t ox, uint32_t oy, uint32_t oz) {
    const float x = (float)(int)(qj.x - ox), y = (float)(int)(qj.y - oy), z = (float)(int)(qj.z - oz);
    if (!EXPANDED) return make_float4(x, y, z, 0.0f);
    const double w = fma((double)x, (double)x, fma((double)y, (double)y, (double)z * (double)z));
    return make_float4(x, y, z, (float)w);
}
template <bool EXPANDED>
__device__ __forceinline__ float4 nbw_dummy() {   // a candidate infinitely far away
    return EXPANDED ? make_float4(0.0f, 0.0f, 0.0f, 1.0e30f) : make_float4(1.0e30f, 1.0e30f, 1.0e30f, 0.0f);
}

// append one chunk's hits.  (Two alternatives were measured in r01 and dropped, profiles/README.md section 4: 8 lanes per
// particle testing 8 candidates at a time with ballot / popc compaction -- no staging, no divergence, but its 4-byte
// stores hit 8 different list rows per particle and instruction, 1670 us against 750 us; and deferring the appends
// through shared-memory hit masks, so that a warp does not wait for its busiest lane in every chunk, 13 % slower: the
// kernel's shared-memory pipe is already the bottleneck of the test loop -- every group-broadcast LDS.128 costs 4
// wavefronts.)
__device__ __forceinline__ void nbw_append_now(unsigned m, int base, unsigned long long row_addr, int k_max, int &cnt) {
    while (m) {
        const int t1 = __ffs(m) - 1;
        m &= m - 1;
        // entry cnt of the tile-sliced row (md_nbr_off): 32-bit index, one IMAD.WIDE + STG.  A full row keeps counting (the
        // overflow is reported by the caller) and overwrites its last slot, so every entry below nbr_cnt is a valid slot.
        asm volatile("st.global.b32 [%0], %1;" ::"l"(row_addr + (unsigned long long)(unsigned)min(cnt, k_max) * 128ull), "r"(base + t1) : "memory");
        ++cnt;
    }
}

template <bool EXPANDED, int NBW_G>
__global__ void __launch_bounds__(NBW_WARPS * 32, 9)
k_nbr_build_warp(const double *__restrict__ r, const uint4 *__restrict__ q, long long i_first /* multiple of 32 */, long long n_own, long long cap,
                 const int *__restrict__ cell, const int *__restrict__ cell_start, const int *__restrict__ cell_end, int nc,
                 double L, double halfL, double cell_len, double inv_h, double rl2, double rl2s_d, float bw, float rl2s_lo,
                 float rl2s_hi, int max_nbr, int run_max, int *__restrict__ nbr, int *__restrict__ nbr_cnt,
                 StepScalars *__restrict__ sc) {
    __shared__ float4 s_cand[NBW_WARPS][NBW_G][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long i0 = i_first + ((long long)blockIdx.x * NBW_WARPS + w) * 32;
    if (i0 >= n_own) return;   // whole warp out of range (only __syncwarp below)
    const long long i = i0 + lane;
    const bool valid = i < n_own;
    const long long il = valid ? i : n_own - 1;
    const uint4 qi = q[il];
    const int c = cell[il];
    const int cx = c % nc, cy = (c / nc) % nc, cz = c / (nc * nc);
    float4 *cand = s_cand[w][0];
    const int grp = lane / (32 / NBW_G);            // lane group: each group walks its own candidate stream
    float4 *cand_g = s_cand[w][grp];
    int *row0 = nbr + md_nbr_row_base(il, max_nbr);
    const unsigned long long row_addr = (unsigned long long)row0;
    const int k_max = max_nbr - 1;
    const float neg_bw = -bw;
    float bandmin = 3.0e38f;
    int cnt = 0;
    unsigned todo = __ballot_sync(0xffffffffu, valid);
    while (todo) {
        // ---- pick the particles of this pass: same (cy, cz) row as the first pending lane, cx within NBW_RUN cells
        const int leader = __ffs(todo) - 1;
        const int cl = __shfl_sync(0xffffffffu, c, leader);
        const int cxl = cl % nc, cyl = (cl / nc) % nc, czl = cl / (nc * nc);
        const bool mine = ((todo >> lane) & 1u) && cy == cyl && cz == czl && cx >= cxl && cx < cxl + run_max;
        todo &= ~__ballot_sync(0xffffffffu, mine);
        const int cxb = __reduce_max_sync(0xffffffffu, mine ? cx : cxl);
        const int ncx = cxb - cxl + 3;   // candidate cells cxl-1 .. cxb+1 (periodic); run_max <= nc - 3 keeps them distinct
        // common origin = centre of the i-cells of the pass, in fixed point
        const uint32_t ox = md_to_fixed((cxl + 0.5 * (double)(cxb - cxl + 1)) * cell_len, inv_h);
        const uint32_t oy = md_to_fixed((cyl + 0.5) * cell_len, inv_h), oz = md_to_fixed((czl + 0.5) * cell_len, inv_h);
        const float xi = (float)(int)(qi.x - ox), yi = (float)(int)(qi.y - oy), zi = (float)(int)(qi.z - oz);
        const float px2 = EXPANDED ? -2.0f * xi : xi, py2 = EXPANDED ? -2.0f * yi : yi, pz2 = EXPANDED ? -2.0f * zi : zi;
        const float ki = EXPANDED ? (float)(fma((double)xi, (double)xi, fma((double)yi, (double)yi, (double)zi * (double)zi)) - rl2s_d)
                                  : -(float)rl2s_d;
        // Candidate streams per lane group: the 8 consecutive particles of a group span 1-2 cells, so the group only
        // needs the cells [its first cell - 1, its last cell + 1] of every row -- about 3.6 cells instead of the 5.5
        // the whole tile needs.  tlo / thi index the pass's candidate cells (t = 0 is cell cxl - 1).
        int tlo = mine ? cx - cxl : 64, thi = mine ? cx - cxl + 2 : -1;
#pragma unroll
        for (int o = 1; o < 32 / NBW_G; o <<= 1) {
            tlo = min(tlo, __shfl_xor_sync(0xffffffffu, tlo, o));
            thi = max(thi, __shfl_xor_sync(0xffffffffu, thi, o));
        }
        const bool grp_on = thi >= 0;
        for (int dz = -1; dz <= 1; ++dz) {
            const int z = (czl + dz + nc) % nc;
            // lanes 0..23 fetch the slot ranges of the 3 x ncx candidate cells of this z layer
            int cs_l = 0, ce_l = 0;
            {
                const int dyi = lane >> 3, t = lane & 7;
                if (dyi < 3 && t < ncx) {
                    const int y = (cyl + dyi - 1 + nc) % nc;
                    const int x = (cxl - 1 + t + nc) % nc;
                    const int cc = (z * nc + y) * nc + x;
                    cs_l = cell_start[cc];
                    ce_l = cell_end[cc];
                }
            }
            // a row whose ncx cells are all occupied and slot-contiguous can be walked per group
            unsigned row_good;
            {
                const int ce_prev = __shfl_up_sync(0xffffffffu, ce_l, 1);
                const int t = lane & 7;
                const bool good = lane >= 24 || t >= ncx || (ce_l > cs_l && (t == 0 || cs_l == ce_prev));
                row_good = __ballot_sync(0xffffffffu, good);
            }
            for (int dyi = 0; dyi < 3; ++dyi) {
                if (((row_good >> (dyi * 8)) & 0xffu) == 0xffu) {
                    // ---- grouped walk: every lane group streams its own slot range [S, E) ----
                    // (all lanes take part in the shuffles; groups without a particle in this pass get an empty range)
                    const int S_all = __shfl_sync(0xffffffffu, cs_l, dyi * 8 + (grp_on ? tlo : 0));
                    const int E_all = __shfl_sync(0xffffffffu, ce_l, dyi * 8 + (grp_on ? thi : 0));
                    const int S = grp_on ? S_all : 0, E = grp_on ? E_all : 0;
                    const int len_max = __reduce_max_sync(0xffffffffu, E - S);
                    int Sg[NBW_G], Eg[NBW_G];
#pragma unroll
                    for (int k = 0; k < NBW_G; ++k) {
                        Sg[k] = __shfl_sync(0xffffffffu, S, k * (32 / NBW_G));
                        Eg[k] = __shfl_sync(0xffffffffu, E, k * (32 / NBW_G));
                    }
                    for (int c0 = 0; c0 < len_max; c0 += 32) {
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < NBW_G; ++k) {   // lane L stages candidate L of every group's chunk
                            const int Sk = Sg[k], Ek = Eg[k];
                            const int jj = Sk + c0 + lane;
                            float4 cd = nbw_dummy<EXPANDED>();
                            if (jj < Ek) cd = nbw_stage<EXPANDED>(q[jj], ox, oy, oz);
                            s_cand[w][k][lane] = cd;
                        }
                        __syncwarp();
                        const int base = S + c0;
                        float bm = 3.0e38f;
                        unsigned m = nbw_test_chunk<EXPANDED>(cand_g, min(32, len_max - c0), px2, py2, pz2, ki, neg_bw, bm);
                        if (mine) bandmin = fminf(bandmin, bm);
                        else m = 0u;
                        const long long self = i - base;
                        if (self >= 0 && self < 32) m &= ~(1u << self);
                        nbw_append_now(m, base, row_addr, k_max, cnt);
                    }
                    continue;
                }
                int rs = 0, re = 0;   // pending (merged) slot range
                for (int t = 0; t <= ncx; ++t) {
                    int s1 = 0, e1 = 0;
                    if (t < ncx) {
                        s1 = __shfl_sync(0xffffffffu, cs_l, dyi * 8 + t);
                        e1 = __shfl_sync(0xffffffffu, ce_l, dyi * 8 + t);
                        if (e1 == s1) continue;
                        if (re > rs && s1 == re) { re = e1; continue; }   // contiguous with the pending range
                    }
                    // flush the pending range, 32 candidates at a time
                    for (int base = rs; base < re; base += 32) {
                        const int jj = base + lane;
                        float4 cd = nbw_dummy<EXPANDED>();
                        if (jj < re) cd = nbw_stage<EXPANDED>(q[jj], ox, oy, oz);
                        __syncwarp();
                        cand[lane] = cd;
                        __syncwarp();
                        float bm = 3.0e38f;
                        unsigned m = nbw_test_chunk<EXPANDED>(cand, min(32, re - base), px2, py2, pz2, ki, neg_bw, bm);
                        if (mine) bandmin = fminf(bandmin, bm);
                        else m = 0u;
                        const long long self = i - base;
                        if (self >= 0 && self < 32) m &= ~(1u << self);
                        nbw_append_now(m, base, row_addr, k_max, cnt);
                    }
                    rs = s1; re = e1;
                }
            }
        }
    }
    if (valid) {
        if (__builtin_expect(bandmin <= bw, 0)) {   // a candidate sat in the guard band: canonical float64 redo of this particle
            bool band = false;
            cnt = nb_particle<true>(i, qi, c, r, q, cap, cell_start, cell_end, nc, L, halfL, rl2, rl2s_lo, rl2s_hi, max_nbr, row0, band);
        }
        if (cnt > max_nbr) {
            sc->overflow_flag = 1;
            cnt = max_nbr;
        }
        nbr_cnt[i] = cnt;
        for (int k = cnt; k < ((cnt + NBG - 1) / NBG) * NBG; ++k) row0[k * 32] = (int)i;  // padding, masked by j != i
    }
}
