// K3t / K4t: the TILE path of the cut-off family (fp32 pair math).  Included by cells.cu only.
//
// Why: the per-particle list kernels (cells_force.cuh) spend their time in the L1: every listed pair is a 16-byte
// gather from the packed position array and a warp's 32 gathers touch ~20 cache lines (77 % of the L1TEX peak at 28 %
// FMA-pipe utilisation, profiles/README.md section 3).  Here the gathers come from SHARED memory instead:
//
//   block   = 128 consecutive cell-sorted slots (4 warps, one particle per lane)
//   window  = every slot a particle of the block can have on its list: for each RUN of the block (consecutive cells
//             of one x-row of the cell grid) the 9 stencil rows, cells [first - 1, last + 1] -- a handful of
//             contiguous slot ranges, ~1400 entries = 22 KB at the liquid density of configs[2].  The force kernel
//             copies the ranges into shared memory once per step (coalesced 16-byte loads of the current fixed-point
//             positions) and then gathers from there.
//   list    = per particle, 16-bit BYTE OFFSETS into the window (entry index * 16 <= 65520), 8 entries per 16-byte
//             vector, tile-sliced so that a warp reads 512 contiguous bytes per group of 8 entries.  Half the bytes of
//             the 32-bit slot lists (the list stream is the kernel's DRAM traffic).  Window entry 0 is a dummy whose 4th
//             word is 1e30f: the 4th word of every record is ADDED to r^2 (0.0f for real particles), so rows are padded to
//             a multiple of 8 with offset 0 at no cost in the pair loop, and the own slot is simply not listed.
//
// The list is the CANONICAL Verlet list: same float prefilter + guard band + float64 redo as
// k_nbr_build_warp, so mdsea_b200_get_neighbors reads it back bit for bit from these tile lists (host side: cells.cu).
// What changes against k_nbr_build_warp: hits are appended as 16-bit window offsets into per-lane rows in shared
// memory (coalesced write-out at the end), one candidate stream per warp, and the float64 redo of a particle in the
// guard band is done by the whole warp (32 candidates per step) instead of by its lane alone.
//
// Fallback: a block whose window does not fit (more than TL_MAXRUN runs, TL_MAXRNG ranges or TL_WIN_MAX entries -- dilute
// gases, voids) raises a flag; the host then rebuilds with the per-particle kernels for this list generation.
#pragma once

#define TL_BLOCK 128        // slots per block
#define TL_WARPS 4
#define TL_MAXRUN 8         // runs per block
#define TL_ROWC 16          // cells per (run, stencil row) incl. the two flanking cells: runs of <= 14 cells
#define TL_MAXRNG 128       // contiguous slot ranges per window (<= TL_BLOCK: one thread copies one range descriptor)
#define TL_WIN_MAX 4095     // window entries (16-bit byte offsets)
#define TL_TAB (TL_MAXRUN * 9 * TL_ROWC)
#define TL_MAXCHK 96        // chunks of <= 32 candidates per pass of the build

struct TileDesc {           // window of one block: written by k_tile_build, read by the force kernel every step
    int n_rng;
    int n_win;              // entries incl. the dummy at index 0
    int rng_slot[TL_MAXRNG];   // first slot of range k
    int rng_lb[TL_MAXRNG];     // window index of its first entry (ranges tile the window: length = next lb - lb)
};

// block-wide exclusive scan of one int per thread (TL_BLOCK threads); total returned to every thread
__device__ __forceinline__ int tl_block_exscan(int v, int *s_w /* TL_WARPS + 1 */, int &total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
    }
    __syncthreads();
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    int off = 0, tot = 0;
#pragma unroll
    for (int k = 0; k < TL_WARPS; ++k) {
        const int t = s_w[k];
        if (k < w) off += t;
        tot += t;
    }
    total = tot;
    return off + inc - v;
}

// Shared-memory list rows of the build: entry k of thread tid is the 16-bit element k * TL_BLOCK + tid; a lane keeps a
// POINTER to its next entry (one STS + one clamped add per hit).  Row max_nbr is a sink: a lane that fills its list keeps
// writing there, and a list that reaches max_nbr entries is reported as an overflow.
// Masks come from tl_test_chunk: `nbits` candidates were shifted in, candidate u sits at bit nbits - 1 - u (first candidate =
// highest bit, found with one FLO); its window index is lbase + u.
__device__ __forceinline__ void tl_append(unsigned m, int nbits, int lbase, unsigned short *__restrict__ rows, unsigned &wb /* byte offset of the next entry */,
                                          unsigned wsink) {
    const int top = (lbase + nbits - 1) << 4;
    while (m) {
        const int p = 31 - __clz(m);
        m ^= 1u << p;
        *(unsigned short *)((char *)rows + wb) = (unsigned short)(top - (p << 4));
        wb = min(wb + 2u * TL_BLOCK, wsink);
    }
}

// test one staged chunk (<= 32 candidates, padded with far-away dummies) against this lane's particle: expanded form
// d' = |c|^2 + (|p|^2 - rl^2 + bw) - 2 p.c = r^2 - rl^2 + bw (see nbw_test_chunk for the error bound).  A candidate is a
// certain hit iff d' < 0, i.e. iff its SIGN BIT is set: one funnel shift collects it (no compare, no select).  The guard
// band |r^2 - rl^2| <= bw is 0 <= d' <= 2 bw: as unsigned integers negative floats are larger than all positive ones, so an
// unsigned minimum over the bit patterns tracks the smallest non-negative d'.  3 FFMA + 1 FADD + SHF + VIMNMX per test.
__device__ __forceinline__ unsigned tl_test_chunk(const float4 *__restrict__ cand, int ncand, float a, float b, float cc, float kp,
                                                  unsigned &bandbits, int &nbits) {
    unsigned m = 0u;
    int t0 = 0;
    for (; t0 < ncand; t0 += 8) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 cj = cand[t0 + u];   // broadcast LDS.128
            const unsigned d = __float_as_uint(fmaf(a, cj.x, fmaf(b, cj.y, fmaf(cc, cj.z, kp))) + cj.w);
            m = __funnelshift_l(d, m, 1);
            bandbits = min(bandbits, d);
        }
    }
    nbits = t0;
    return m;
}

// tile_stats: [0] fallback requested, [1] max window, [2] max ranges
__global__ void __launch_bounds__(TL_BLOCK, 6)
k_tile_build(const double *__restrict__ r, const uint4 *__restrict__ q, long long n_own, long long cap,
             const int *__restrict__ cell, const int *__restrict__ cell_start, const int *__restrict__ cell_end, int nc,
             double L, double halfL, double cell_len, double inv_h, double rl2, double rl2s_d, float bw, int max_nbr,
             int run_max, TileDesc *__restrict__ desc, uint4 *__restrict__ list, int *__restrict__ nbr_cnt,
             StepScalars *__restrict__ sc, int *__restrict__ tile_stats) {
    extern __shared__ unsigned short s_rows[];                 // [max_nbr + 1][TL_BLOCK]
    __shared__ float4 s_cand[TL_WARPS][32];
    __shared__ int   s_chk[TL_WARPS][2 * TL_MAXCHK];
    __shared__ int   s_cell[TL_BLOCK];
    __shared__ int   s_aux[TL_BLOCK];
    __shared__ unsigned char s_brk[TL_BLOCK + 1];
    __shared__ int   s_w[TL_WARPS + 1];
    __shared__ int   run_cxa[TL_MAXRUN], run_cxb[TL_MAXRUN], run_cy[TL_MAXRUN], run_cz[TL_MAXRUN], run_ncx[TL_MAXRUN];
    __shared__ int   tab_cs[TL_TAB];
    __shared__ unsigned short tab_lb[TL_TAB];   // window indices (<= TL_WIN_MAX)
    __shared__ unsigned short tab_n[TL_TAB];
    __shared__ int   s_bad;

    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const long long s0 = (long long)blockIdx.x * TL_BLOCK;
    const long long i = s0 + tid;
    const bool valid = i < n_own;
    TileDesc *__restrict__ D = desc + blockIdx.x;

    // ---- runs of the block: consecutive slots whose cells are x-adjacent in one row of the grid, cut every TL_ROWC - 2 cells
    const int c = valid ? cell[i] : -1;
    s_cell[tid] = c;
    if (tid == 0) s_bad = 0;
    __syncthreads();
    int cx = 0, row = -1;
    if (valid) { cx = c % nc; row = c / nc; }
    bool brk = false;
    if (valid) {
        if (tid == 0) brk = true;
        else {
            const int pc = s_cell[tid - 1], prow = pc / nc, pcx = pc % nc;
            brk = prow != row || (cx != pcx && cx != pcx + 1);
        }
    }
    int tot = 0;
    int run1 = tl_block_exscan(brk ? 1 : 0, s_w, tot) + (brk ? 1 : 0) - 1;   // index of the uncut run this slot belongs to
    if (brk) s_aux[run1] = cx;                                                 // its first cell (run1 < TL_BLOCK)
    __syncthreads();
    int sub = 0;
    if (valid) sub = (cx - s_aux[run1]) / (TL_ROWC - 2);
    bool brk2 = brk;
    if (valid && !brk) {
        const int pcx = s_cell[tid - 1] % nc;
        brk2 = (pcx - s_aux[run1]) / (TL_ROWC - 2) != sub;
    }
    __syncthreads();   // s_aux is reused by the scan's callers below
    int nruns = 0;
    const int rn = tl_block_exscan(brk2 ? 1 : 0, s_w, nruns) + (brk2 ? 1 : 0) - 1;
    s_brk[tid] = brk2 ? 1 : 0;
    if (tid == 0) s_brk[TL_BLOCK] = 1;
    __syncthreads();
    if (nruns > TL_MAXRUN) {   // uniform: dilute block, the per-particle kernels take this list generation
        if (tid == 0) atomicExch(&tile_stats[0], 1);
        return;
    }
    if (valid) {
        if (brk2) { run_cxa[rn] = cx; run_cy[rn] = row % nc; run_cz[rn] = row / nc; }
        const bool last = s_brk[tid + 1] || i + 1 >= n_own;
        if (last) {
            run_cxb[rn] = cx;
        }
    }
    __syncthreads();
    if (tid < nruns) {
        int ncx = run_cxb[tid] - run_cxa[tid] + 3;
        run_ncx[tid] = ncx > nc ? nc : ncx;   // a run that (with its flanks) covers the whole period: every cell once
    }
    __syncthreads();

    // ---- window table: slot range and window base of every (run, stencil row, cell)
    const int E = nruns * 9 * TL_ROWC;
    const int per = (E + TL_BLOCK - 1) / TL_BLOCK;
    int mysum = 0;
    for (int k = 0; k < per; ++k) {
        const int e = tid * per + k;
        if (e >= E) break;
        const int rr = e / (9 * TL_ROWC), rem = e % (9 * TL_ROWC), rho = rem / TL_ROWC, t = rem % TL_ROWC;
        int cs = 0, n = 0;
        if (t < run_ncx[rr]) {
            const int x = (run_cxa[rr] - 1 + t + nc) % nc;
            const int y = (run_cy[rr] + (rho % 3) - 1 + nc) % nc;
            const int z = (run_cz[rr] + (rho / 3) - 1 + nc) % nc;
            const int cc = (z * nc + y) * nc + x;
            cs = cell_start[cc];
            n = cell_end[cc] - cs;
        }
        tab_cs[e] = cs;
        tab_n[e] = (unsigned short)min(n, 65535);
        if (n > 65535) s_bad = 1;
        mysum += n;
    }
    int n_win = 0;
    int base = tl_block_exscan(mysum, s_w, n_win) + 1;   // entry 0 is the dummy
    n_win += 1;
    for (int k = 0; k < per; ++k) {
        const int e = tid * per + k;
        if (e >= E) break;
        tab_lb[e] = (unsigned short)min(base, 65535);
        base += tab_n[e];
    }
    __syncthreads();
    // ranges: maximal slot-contiguous pieces of a (run, row)
    int nst = 0;
    for (int k = 0; k < per; ++k) {
        const int e = tid * per + k;
        if (e >= E) break;
        const int t = e % TL_ROWC;
        const bool st = tab_n[e] > 0 && !(t > 0 && tab_n[e - 1] > 0 && tab_cs[e - 1] + (int)tab_n[e - 1] == tab_cs[e]);
        nst += st ? 1 : 0;
    }
    int n_rng = 0;
    int rk = tl_block_exscan(nst, s_w, n_rng);
    const bool bad = n_rng > TL_MAXRNG || n_win > TL_WIN_MAX || s_bad;
    if (bad) {   // uniform
        if (tid == 0) atomicExch(&tile_stats[0], 1);
        return;
    }
    for (int k = 0; k < per; ++k) {
        const int e = tid * per + k;
        if (e >= E) break;
        const int t = e % TL_ROWC;
        const bool st = tab_n[e] > 0 && !(t > 0 && tab_n[e - 1] > 0 && tab_cs[e - 1] + (int)tab_n[e - 1] == tab_cs[e]);
        if (st) {
            D->rng_slot[rk] = tab_cs[e];
            D->rng_lb[rk] = tab_lb[e];
            ++rk;
        }
    }
    if (tid == 0) {
        D->n_rng = n_rng;
        D->n_win = n_win;
        atomicMax(&tile_stats[1], n_win);
        atomicMax(&tile_stats[2], n_rng);
    }

    // ---- test phase: one warp per tile of 32 slots, one candidate stream per pass (same row of cells, <= run_max cells)
    const long long i0w = s0 + (long long)w * 32;
    if (i0w >= n_own) return;   // whole warp out of range (only __syncwarp below)
    const long long il = valid ? i : n_own - 1;
    const uint4 qi = q[il];
    float4 *cand = s_cand[w];
    int *chk = s_chk[w];
    unsigned short *const wrow = s_rows + tid;
    const unsigned wb0 = 2u * (unsigned)tid, wsink = wb0 + 2u * TL_BLOCK * (unsigned)max_nbr;
    unsigned wb = wb0;
    unsigned bandbits = 0x7f800000u;   // smallest non-negative d' seen (bit pattern)
    unsigned todo = __ballot_sync(0xffffffffu, valid);
    while (todo) {
        const int leader = __ffs(todo) - 1;
        const int rl = __shfl_sync(0xffffffffu, rn, leader);
        const int cxl = __shfl_sync(0xffffffffu, cx, leader);
        const bool mine = ((todo >> lane) & 1u) && rn == rl && cx >= cxl && cx < cxl + run_max;
        todo &= ~__ballot_sync(0xffffffffu, mine);
        const int cxb = __reduce_max_sync(0xffffffffu, mine ? cx : cxl);
        const int cxa = run_cxa[rl], ncx = run_ncx[rl], cyl = run_cy[rl], czl = run_cz[rl];
        const int t0 = cxl - cxa;
        const int nt = min(cxb - cxl + 3, ncx);   // candidate cells t0 .. t0 + nt - 1 (mod ncx) of every stencil row
        const uint32_t ox = md_to_fixed((cxl + 0.5 * (double)(cxb - cxl + 1)) * cell_len, inv_h);
        const uint32_t oy = md_to_fixed((cyl + 0.5) * cell_len, inv_h), oz = md_to_fixed((czl + 0.5) * cell_len, inv_h);
        const float xi = (float)(int)(qi.x - ox), yi = (float)(int)(qi.y - oy), zi = (float)(int)(qi.z - oz);
        const float px2 = -2.0f * xi, py2 = -2.0f * yi, pz2 = -2.0f * zi;
        const float kp = (float)(fma((double)xi, (double)xi, fma((double)yi, (double)yi, (double)zi * (double)zi)) - rl2s_d + (double)bw);
        // chunk list of the pass: (first slot, window index of it << 6 | candidates) per chunk of <= 32 candidates, recorded
        // by the (warp-uniform) walk over the 9 stencil rows, then processed with the NEXT chunk's candidates already in flight
        int nchk = 0;
        auto add_range = [&](int rs, int re, int lb) {
            for (int b0 = rs; b0 < re; b0 += 32) {
                if (nchk < TL_MAXCHK && lane == 0) {
                    chk[2 * nchk] = b0;
                    chk[2 * nchk + 1] = ((lb + (b0 - rs)) << 6) | min(32, re - b0);
                }
                ++nchk;
            }
        };
        for (int rho = 0; rho < 9; ++rho) {
            const int tb = (rl * 9 + rho) * TL_ROWC;
            {   // usual case: the nt cells are one contiguous piece of slots (all in one class, no periodic wrap): since the
                // cells of a row are slot-adjacent within their class, (slot span == number of window entries) decides it
                const int tl = t0 + nt - 1;
                if (tl < ncx) {
                    const int nf = tab_n[tb + t0], nl = tab_n[tb + tl];
                    const int sf = tab_cs[tb + t0], lf = tab_lb[tb + t0];
                    const int sspan = tab_cs[tb + tl] + nl - sf, lspan = tab_lb[tb + tl] + nl - lf;
                    if (nf > 0 && nl > 0 && sspan == lspan) {
                        add_range(sf, sf + sspan, lf);
                        continue;
                    }
                }
            }
            int rs = 0, re = 0, lb = 0;   // pending slot range [rs, re), window index of rs
            for (int k = 0; k <= nt; ++k) {
                int s1 = 0, n1 = 0, l1 = 0;
                if (k < nt) {
                    int t = t0 + k;
                    if (t >= ncx) t -= ncx;
                    s1 = tab_cs[tb + t];
                    n1 = tab_n[tb + t];
                    l1 = tab_lb[tb + t];
                    if (n1 == 0) continue;
                    if (re > rs && s1 == re && l1 == lb + (re - rs)) { re += n1; continue; }   // contiguous in slots and in the window
                }
                add_range(rs, re, lb);
                rs = s1; re = s1 + n1; lb = l1;
            }
        }
        if (nchk > TL_MAXCHK) {   // uniform; cannot happen at liquid densities (<= 9 rows x 7 chunks)
            if (lane == 0) atomicExch(&tile_stats[0], 1);
            nchk = TL_MAXCHK;
        }
        __syncwarp();
        uint4 qn = make_uint4(0u, 0u, 0u, 0u);
        if (nchk > 0 && lane < (chk[1] & 63)) qn = q[chk[0] + lane];
        for (int ck = 0; ck < nchk; ++ck) {
            const int b0 = chk[2 * ck], pk = chk[2 * ck + 1], nc_ = pk & 63, lw = pk >> 6;
            float4 cd = nbw_dummy<true>();
            if (lane < nc_) cd = nbw_stage<true>(qn, ox, oy, oz);
            if (ck + 1 < nchk && lane < (chk[2 * ck + 3] & 63)) qn = q[chk[2 * ck + 2] + lane];   // next chunk's candidates
            __syncwarp();
            cand[lane] = cd;
            __syncwarp();
            unsigned bb = 0x7f800000u;
            int nbits;
            unsigned m = tl_test_chunk(cand, nc_, px2, py2, pz2, kp, bb, nbits);
            if (mine) bandbits = min(bandbits, bb);
            else m = 0u;
            const long long self = i - b0;
            if (self >= 0 && self < nbits) m &= ~(1u << (nbits - 1 - (int)self));
            tl_append(m, nbits, lw, s_rows, wb, wsink);
        }
        __syncwarp();   // the chunk list is rewritten by the next pass
    }
    // ---- a candidate sat in the guard band: canonical float64 redo of that particle by the whole warp, over its own
    //      three cells of every stencil row (the same canonical set, in the same order)
    unsigned flagged = __ballot_sync(0xffffffffu, valid && bandbits <= __float_as_uint(2.0f * bw));
    while (flagged) {
        const int f = __ffs(flagged) - 1;
        flagged &= flagged - 1;
        const int rl = __shfl_sync(0xffffffffu, rn, f);
        const int cxf = __shfl_sync(0xffffffffu, cx, f);
        const long long jf = i0w + f;
        const double xf = r[jf], yf = r[cap + jf], zf = r[2 * cap + jf];
        const int cxa = run_cxa[rl], ncx = run_ncx[rl];
        const int t0 = cxf - cxa;
        const int nt = min(3, ncx);
        unsigned wf = wb0;
        for (int rho = 0; rho < 9; ++rho) {
            const int tb = (rl * 9 + rho) * TL_ROWC;
            int rs = 0, re = 0, lb = 0;
            for (int k = 0; k <= nt; ++k) {
                int s1 = 0, n1 = 0, l1 = 0;
                if (k < nt) {
                    int t = t0 + k;
                    if (t >= ncx) t -= ncx;
                    s1 = tab_cs[tb + t];
                    n1 = tab_n[tb + t];
                    l1 = tab_lb[tb + t];
                    if (n1 == 0) continue;
                    if (re > rs && s1 == re && l1 == lb + (re - rs)) { re += n1; continue; }
                }
                for (int b0 = rs; b0 < re; b0 += 32) {
                    const int jj = b0 + lane;
                    bool in = false;
                    if (jj < re && jj != (int)jf) {
                        const double ddx = md_minimg_exact(__dsub_rn(r[jj], xf), L, halfL);
                        const double ddy = md_minimg_exact(__dsub_rn(r[cap + jj], yf), L, halfL);
                        const double ddz = md_minimg_exact(__dsub_rn(r[2 * cap + jj], zf), L, halfL);
                        in = md_r2_canonical(ddx, ddy, ddz) < rl2;
                    }
                    const unsigned m = __brev(__ballot_sync(0xffffffffu, in));   // candidate u at bit 31 - u
                    if (lane == f) tl_append(m, 32, lb + (b0 - rs), s_rows, wf, wsink);
                }
                rs = s1; re = s1 + n1; lb = l1;
            }
        }
        if (lane == f) wb = wf;
    }
    // ---- write-out: rows padded to a multiple of 8 with the dummy (offset 0), 16 bytes per lane and group of 8
    int cnt = (int)((wb - wb0) / (2u * TL_BLOCK));
    if (valid) {
        if (cnt >= max_nbr) {   // the sink row was (or would next be) used: capacity must exceed the longest list
            sc->overflow_flag = 1;
            cnt = max_nbr - 8;
        }
        nbr_cnt[i] = cnt;
        for (int k = cnt; k < ((cnt + 7) & ~7); ++k) wrow[(size_t)k * TL_BLOCK] = 0;
    }
    __syncwarp();
    const int cnt8 = valid ? (cnt + 7) >> 3 : 0;
    const int m8 = __reduce_max_sync(0xffffffffu, cnt8);
    uint4 *__restrict__ lrow = list + ((size_t)(i0w >> 5) * (size_t)(max_nbr >> 3)) * 32u + lane;
    for (int k8 = 0; k8 < m8; ++k8) {
        if (k8 < cnt8) {
            const unsigned short *__restrict__ g = wrow + (size_t)(8 * k8) * TL_BLOCK;
            uint4 e;
            e.x = (unsigned)g[0 * TL_BLOCK] | ((unsigned)g[1 * TL_BLOCK] << 16);
            e.y = (unsigned)g[2 * TL_BLOCK] | ((unsigned)g[3 * TL_BLOCK] << 16);
            e.z = (unsigned)g[4 * TL_BLOCK] | ((unsigned)g[5 * TL_BLOCK] << 16);
            e.w = (unsigned)g[6 * TL_BLOCK] | ((unsigned)g[7 * TL_BLOCK] << 16);
            lrow[(size_t)k8 * 32u] = e;
        }
    }
}

// ---------------------------------------------------------------------------------------------------
// K4t: forces from the tile lists.  One block per 128 slots: the window is copied into shared memory (coalesced
// 16-byte loads of the CURRENT fixed-point positions), then every lane walks its list 8 entries at a time:
// one 16-byte list vector, 8 independent LDS.128 gathers, 8 pair evaluations.  Same arithmetic and the same exact
// treatment of the cutoff band as k_nbr_force_f32 (borderline pairs are re-decided from the float64 positions).
// Launched over the blocks that cover the owned slots [i_begin, n_own); lanes outside that range stay idle, so the
// interior / boundary split of the multi-rank step needs no alignment.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int tl_slot_of(int lidx, const int *__restrict__ s_slot, const int *__restrict__ s_lb, int nr) {
    int k = 0;
    while (k + 1 < nr && s_lb[k + 1] <= lidx) ++k;
    return s_slot[k] + (lidx - s_lb[k]);
}

template <int POTK>
__global__ void __launch_bounds__(TL_BLOCK, 6)
k_tile_force_f32(const uint4 *__restrict__ pos, const double *__restrict__ r64, const TileDesc *__restrict__ desc,
                 const uint4 *__restrict__ list, const int *__restrict__ nbr_cnt, int max8, long long blk0, long long i_begin,
                 long long n_own, long long cap, double L, double halfL, double rc2, float rc2s, float bw, double inv_h, PotF32 P,
                 double *__restrict__ acc, double *__restrict__ pe_part, StepScalars *__restrict__ sc, int guard_tag) {
    extern __shared__ uint4 s_win[];
    __shared__ int      s_slot[TL_MAXRNG];
    __shared__ int      s_lb[TL_MAXRNG + 1];
    __shared__ double   red[32];
    __shared__ unsigned redu[32];
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    // speculative launch: if the skin trigger of the drift this launch follows has fired, the list is stale and the host
    // will redo the step.  The decision must be BLOCK-UNIFORM: with world > 1 the flag can flip while the kernel runs (the
    // halo unpack on the communication stream merges the neighbours' flags), and the threads of a block depend on each other
    // here (window staging, range tables) -- a block that half-returned would gather through unwritten range descriptors.
    __shared__ int s_skip;
    if (tid == 0) s_skip = guard_tag && *(volatile const int *)&sc->rebuild_flag == guard_tag;
    __syncthreads();
    if (s_skip) return;
    const long long b = blk0 + blockIdx.x;
    const TileDesc *__restrict__ D = desc + b;
    const int nr = D->n_rng, nw = D->n_win;
    if (tid < nr) {
        s_slot[tid] = D->rng_slot[tid];
        s_lb[tid] = D->rng_lb[tid];
    }
    if (tid == 0) {
        s_lb[nr] = nw;
        s_win[0] = make_uint4(0u, 0u, 0u, __float_as_uint(1.0e30f));
    }
    // this lane's particle, list length and first list vector: issued before the window copy so that their latency overlaps it
    const long long i = b * TL_BLOCK + tid;
    const bool act = i >= i_begin && i < n_own;
    const uint4 *__restrict__ lrow = list + ((size_t)(i >> 5) * (size_t)max8) * 32u + lane;
    uint4 qi = make_uint4(0u, 0u, 0u, 0u), e_next = make_uint4(0u, 0u, 0u, 0u);
    int   nn8 = 0;
    if (act) {
        qi = pos[i];
        nn8 = (nbr_cnt[i] + 7) >> 3;
        e_next = __ldcs(lrow);   // group 0 exists in memory for every owned slot (unused if the list is empty)
    }
    __syncthreads();
    // window copy with cp.async (LDGSTS): 16 bytes per lane and request, global -> shared without a register round trip, all
    // requests of a thread in flight at once (the plain load + store loop was 21 % of the kernel's stall samples: one
    // exposed memory latency per range and lane)
    for (int k = w; k < nr; k += TL_WARPS) {
        const int sl = s_slot[k], lb = s_lb[k], len = s_lb[k + 1] - lb;
        for (int e = lane; e < len; e += 32) {
            const unsigned dst = (unsigned)__cvta_generic_to_shared(s_win + lb + e);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(pos + sl + e) : "memory");
        }
    }
    asm volatile("cp.async.commit_group;\n" ::: "memory");
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    __syncthreads();

    double   Fx = 0.0, Fy = 0.0, Fz = 0.0, PE = 0.0;
    unsigned cnt = 0;
    if (act) {
        const char  *__restrict__ wb = (const char *)s_win;
        const float rc2m = rc2s - bw;        // t' = r^2 - (rc^2 - bw): pair certainly inside <=> t' < 0 (sign bit); band <=> 0 <= t' <= 2 bw
        const unsigned band_hi = __float_as_uint(2.0f * bw);
        const float c8 = P.c / P.sigma2_s;   // LJ: s = t8 * (c8 - 2 c8 t6)
        const float m2c8 = -2.0f * c8;
        float fx = 0.f, fy = 0.f, fz = 0.f, pe = 0.f;
        const uint4 *__restrict__ lp = lrow;
        for (int k8 = 0; k8 < nn8; ++k8) {
            const uint4 e = e_next;
            lp += 32;
            if (k8 + 1 < nn8) e_next = __ldcs(lp);   // next group's vector: in flight during this group's arithmetic
            unsigned off[8];
            off[0] = e.x & 0xffffu; off[1] = e.x >> 16;
            off[2] = e.y & 0xffffu; off[3] = e.y >> 16;
            off[4] = e.z & 0xffffu; off[5] = e.z >> 16;
            off[6] = e.w & 0xffffu; off[7] = e.w >> 16;
            uint4 qj[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) qj[u] = *(const uint4 *)(wb + off[u]);
            unsigned tb = 0x7f800000u;   // smallest non-negative t' of the group (unsigned compare: negative floats are larger)
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float dx = (float)(int)(qj[u].x - qi.x), dy = (float)(int)(qj[u].y - qi.y), dz = (float)(int)(qj[u].z - qi.z);
                const float r2 = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, __uint_as_float(qj[u].w))));
                const unsigned tp = __float_as_uint(r2 - rc2m);
                const int okm = (int)tp >> 31;          // all ones iff inside
                tb = min(tb, tp);
                float s, uu;
                if (POTK == POTF_LJ) {
                    // r^2 if inside, +inf otherwise (rcp -> 0: no force, no energy): one LOP3
                    const float inv = md_rcpf(__uint_as_float((__float_as_uint(r2) & (unsigned)okm) | (~(unsigned)okm & 0x7f800000u)));
                    const float t2 = P.sigma2_s * inv;
                    const float t4 = t2 * t2;
                    const float t6 = t4 * t2;
                    const float t8 = t4 * t4;
                    s = t8 * fmaf(m2c8, t6, c8);
                    uu = fmaf(t6, t6, -t6);
                } else {
                    md_pair<POTK>(r2, okm != 0, P, s, uu);
                }
                fx = fmaf(dx, s, fx);
                fy = fmaf(dy, s, fy);
                fz = fmaf(dz, s, fz);
                pe += uu;
                cnt -= (unsigned)okm;
            }
            if (__builtin_expect(tb <= band_hi, 0)) {   // rare: decide and evaluate the band pairs from the float64 positions
#pragma unroll
                for (int u = 0; u < 8; ++u) {   // unrolled: off[] / qj[] stay in registers
                    const float dxs = (float)(int)(qj[u].x - qi.x), dys = (float)(int)(qj[u].y - qi.y), dzs = (float)(int)(qj[u].z - qi.z);
                    const float r2s = fmaf(dzs, dzs, fmaf(dys, dys, fmaf(dxs, dxs, __uint_as_float(qj[u].w))));
                    if (__float_as_uint(r2s - rc2m) > band_hi) continue;
                    const int jj = tl_slot_of((int)(off[u] >> 4), s_slot, s_lb, nr);
                    const double ex = md_minimg_list(r64[jj] - r64[i], L, halfL);
                    const double ey = md_minimg_list(r64[cap + jj] - r64[cap + i], L, halfL);
                    const double ez = md_minimg_list(r64[2 * cap + jj] - r64[2 * cap + i], L, halfL);
                    const double r2d = fma(ez, ez, fma(ey, ey, ex * ex));
                    if (r2d < rc2) {
                        const float dx = (float)(ex * inv_h), dy = (float)(ey * inv_h), dz = (float)(ez * inv_h);
                        const float r2 = (float)(r2d * inv_h * inv_h);
                        float s, uu;
                        md_pair<POTK>(r2, true, P, s, uu);
                        fx = fmaf(dx, s, fx);
                        fy = fmaf(dy, s, fy);
                        fz = fmaf(dz, s, fz);
                        pe += uu;
                        cnt += 1u;
                    }
                }
            }
            if ((k8 & 3) == 3) {   // fp32 partial sums of 32 entries are folded into the float64 accumulators
                Fx += (double)fx; Fy += (double)fy; Fz += (double)fz; PE += (double)pe;
                fx = fy = fz = pe = 0.f;
            }
        }
        Fx += (double)fx; Fy += (double)fy; Fz += (double)fz; PE += (double)pe;
        acc[i] = Fx;
        acc[cap + i] = Fy;
        acc[2 * cap + i] = Fz;
    }
    const double   bpe = md_block_sum(PE, red);
    const unsigned bc = md_block_sum(cnt, redu);
    if (tid == 0) {
        pe_part[blockIdx.x] = bpe;
        atomicAdd(&sc->pairs_pending, (unsigned long long)bc);
    }
}
