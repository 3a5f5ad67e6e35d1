// Cut-off path (large N): K2 cell binning + radix sort + reorder, K3 Verlet neighbour list with skin trigger,
// K4 neighbour-list force/energy kernels, and the particle-wise integrators that go with them.
//
// None of this exists in the reference (its r_cutoff is a TODO, simulator.py:48-49; update_acc(radius=)
// is broken, :289-309).  The physics is the reference's pair math restricted to dist < r_cutoff (strict,
// simulator.py:283), plain truncation; the geometry definitions are the canonical ones of SURVEY.md 8c,
// restated on the CPU by oracle/lj_oracle.c and matched here BIT FOR BIT:
//   cell id    w = x - floor(x / L) * L ; c = min(nc-1, floor(w / cell_len)) ; id = (cz*nc + cy)*nc + cx
//   order      stable sort by (cell id, particle id)                     -> LSD radix sort, 8-bit digits
//   list of i  j != i with dx*dx + dy*dy + dz*dz < (rc+skin)^2, d = minimage(r_j - r_i), no FMA
//   rebuild    max_i |minimage(r_i - r_i^build)|^2 > (skin/2)^2, evaluated after every drift
//
// Layout in HBM: particles live in CELL-SORTED SLOT ORDER between rebuilds (SoA float64 r, v, acc with row
// stride `cap`, plus their global id); the force kernels gather neighbours from a packed AoS copy of the
// positions that the drift kernel refreshes every step (double4 in fp64 mode; uint4 32-bit fixed point in
// fp32 mode, where integer subtraction gives the minimum image for free).  The neighbour list is a sliced
// ELLPACK (tiles of 32 particles x max_nbr rows) so that a warp of consecutive particles reads it coalesced.
#include <algorithm>

#include "ctx.h"

#define SORT_THREADS 256
#define SORT_ITEMS 8
#define SORT_CHUNK (SORT_THREADS * SORT_ITEMS)
#define NB_THREADS 128

struct CellState {
    int       nc = 0, ncell = 0;
    double    cell_len = 0, rlist = 0, rl2 = 0, rc2 = 0, trig2 = 0;
    long long cap = 0, n_pad = 0;
    int       max_nbr = 0;
    bool      need_rebuild = true;
    // double-buffered state for the reorder
    double *r_alt = nullptr, *v_alt = nullptr;
    int    *id = nullptr, *id_alt = nullptr;
    double *rb = nullptr;                 // positions at the last build, slot order
    void   *pos4 = nullptr;               // double4[cap] (fp64) or uint4[cap] (fp32)
    unsigned long long *key = nullptr, *key_alt = nullptr;
    int    *val = nullptr, *val_alt = nullptr;
    int    *hist = nullptr, *hist_tot = nullptr;
    uint4  *qfix = nullptr;               // 32-bit fixed-point positions, refreshed at every rebuild (list prefilter)
    int     sort_blocks = 0;
    int    *cell = nullptr;               // cell id per slot (sorted)
    int    *cell_start = nullptr;         // ncell + 1
    int    *nbr = nullptr, *nbr_cnt = nullptr;
    double *scratch = nullptr;            // [3][N] for id-ordered downloads
    int    *h_flag = nullptr;             // pinned
    int     id_bits = 0, cell_bits = 0;
};

// ---------------------------------------------------------------------------------------------------
// canonical geometry helpers (exact IEEE operations, no contraction)
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int md_cell_coord(double x, double L, double cl, int nc) {
    const double w = __dsub_rn(x, __dmul_rn(floor(__ddiv_rn(x, L)), L));
    int k = (int)floor(__ddiv_rn(w, cl));
    k = k > nc - 1 ? nc - 1 : k;
    return k < 0 ? 0 : k;
}

// minimum image, bit-identical to d - rint(d / L) * L: for |d| < 1.5 L away from the half-box tie the
// quotient rounds to -1, 0 or +1 and n * L is exact, so subtracting copysign(L, d) is the same single
// rounding; near the tie (or for absurd separations) the exact expression is evaluated.
__device__ __forceinline__ double md_minimg_exact(double d, double L, double halfL) {
    const double a = fabs(d);
    double x = d;
    if (a > halfL) x = __dsub_rn(d, copysign(L, d));
    if (__builtin_expect(fabs(fabs(x) - halfL) < 1e-6 * L || a >= 1.4 * L, 0))
        x = __dsub_rn(d, __dmul_rn(rint(__ddiv_rn(d, L)), L));
    return x;
}

// pairs taken from the neighbour list are closer than rc + skin <= L/3 after imaging, so the raw separation is
// either that small or within rc + skin of +-L: no tie is possible and the subtraction below IS the canonical
// minimum image, bit for bit.
__device__ __forceinline__ double md_minimg_list(double d, double L, double halfL) {
    return fabs(d) > halfL ? __dsub_rn(d, copysign(L, d)) : d;
}

__device__ __forceinline__ double md_r2_canonical(double dx, double dy, double dz) {
    return __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
}

// ---------------------------------------------------------------------------------------------------
// K2a: keys.  key = cell id << 32 | particle id ; value = current slot
// ---------------------------------------------------------------------------------------------------
__global__ void k_cell_keys(const double *__restrict__ r, const int *__restrict__ id, long long n, long long cap, double L,
                            double cl, int nc, unsigned long long *__restrict__ key, int *__restrict__ val) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = md_cell_coord(r[i], L, cl, nc);
    const int cy = md_cell_coord(r[cap + i], L, cl, nc);
    const int cz = md_cell_coord(r[2 * cap + i], L, cl, nc);
    const unsigned long long c = (unsigned long long)((cz * nc + cy) * nc + cx);
    key[i] = (c << 32) | (unsigned)id[i];
    val[i] = (int)i;
}

// ---------------------------------------------------------------------------------------------------
// K2b: stable LSD radix sort, one 8-bit digit per pass (histogram -> scan -> ranked scatter)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(SORT_THREADS)
k_sort_hist(const unsigned long long *__restrict__ key, long long n, int shift, int *__restrict__ hist, int nblk) {
    __shared__ int h[256];
    h[threadIdx.x] = 0;
    __syncthreads();
    const long long base = (long long)blockIdx.x * SORT_CHUNK;
#pragma unroll
    for (int t = 0; t < SORT_ITEMS; ++t) {
        const long long i = base + t * SORT_THREADS + threadIdx.x;
        if (i < n) atomicAdd(&h[(int)((key[i] >> shift) & 255ull)], 1);
    }
    __syncthreads();
    hist[threadIdx.x * nblk + blockIdx.x] = h[threadIdx.x];  // digit-major: the scan yields global offsets
}

// per-digit exclusive scan over the blocks: one CTA per digit row (nblk entries); row totals go to tot[256]
__global__ void __launch_bounds__(256) k_sort_rowscan(int *__restrict__ hist, int nblk, int *__restrict__ tot) {
    __shared__ int part[256];
    int *row = hist + (size_t)blockIdx.x * nblk;
    const int per = (nblk + 255) / 256;
    const int lo = threadIdx.x * per, hi = min(nblk, lo + per);
    int s = 0;
    for (int k = lo; k < hi; ++k) s += row[k];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 256; o <<= 1) {
        const int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;
    for (int k = lo; k < hi; ++k) {
        const int v = row[k];
        row[k] = run;
        run += v;
    }
    if (threadIdx.x == 255) tot[blockIdx.x] = part[255];
}

__global__ void __launch_bounds__(SORT_THREADS)
k_sort_scatter(const unsigned long long *__restrict__ key_in, const int *__restrict__ val_in,
               unsigned long long *__restrict__ key_out, int *__restrict__ val_out, long long n, int shift,
               const int *__restrict__ hist, int nblk, const int *__restrict__ tot) {
    __shared__ int warp_cnt[SORT_THREADS / 32][256];
    __shared__ int digit_base[256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    {   // global offset of digit d = sum of the totals of all smaller digits (256-entry block scan) + row scan
        const int mine = tot[threadIdx.x];
        digit_base[threadIdx.x] = mine;
        __syncthreads();
        for (int o = 1; o < 256; o <<= 1) {
            const int v = threadIdx.x >= o ? digit_base[threadIdx.x - o] : 0;
            __syncthreads();
            digit_base[threadIdx.x] += v;
            __syncthreads();
        }
        const int excl = digit_base[threadIdx.x] - mine;
        __syncthreads();
        digit_base[threadIdx.x] = excl + hist[threadIdx.x * nblk + blockIdx.x];
    }
    const long long base = (long long)blockIdx.x * SORT_CHUNK;
    for (int t = 0; t < SORT_ITEMS; ++t) {
#pragma unroll
        for (int w = 0; w < SORT_THREADS / 32; ++w) warp_cnt[w][threadIdx.x] = 0;
        __syncthreads();
        const long long i = base + t * SORT_THREADS + threadIdx.x;
        const bool valid = i < n;
        unsigned long long k = 0;
        int v = 0, d = 0;
        if (valid) {
            k = key_in[i];
            v = val_in[i];
            d = (int)((k >> shift) & 255ull);
        }
        // rank among the lanes of this warp holding the same digit (lane order == element order)
        const unsigned peers = __match_any_sync(0xffffffffu, valid ? d : 256 + lane);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (valid && rank == 0) warp_cnt[warp][d] = __popc(peers);
        __syncthreads();
        // exclusive prefix over the warps for digit = threadIdx.x, on top of the running base
        {
            int run = digit_base[threadIdx.x];
#pragma unroll
            for (int w = 0; w < SORT_THREADS / 32; ++w) {
                const int c = warp_cnt[w][threadIdx.x];
                warp_cnt[w][threadIdx.x] = run;
                run += c;
            }
            digit_base[threadIdx.x] = run;
        }
        __syncthreads();
        if (valid) {
            const int pos = warp_cnt[warp][d] + rank;
            key_out[pos] = k;
            val_out[pos] = v;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------------
// K2c: reorder the state into sorted slot order; cell boundaries
// ---------------------------------------------------------------------------------------------------
__global__ void k_reorder(const unsigned long long *__restrict__ key, const int *__restrict__ val, long long n, long long cap,
                          const double *__restrict__ r_in, const double *__restrict__ v_in, double *__restrict__ r_out,
                          double *__restrict__ v_out, double *__restrict__ rb, int *__restrict__ id_out,
                          int *__restrict__ cell) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const unsigned long long k = key[s];
    const int o = val[s];
    id_out[s] = (int)(k & 0xffffffffull);
    cell[s] = (int)(k >> 32);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double x = r_in[d * cap + o];
        r_out[d * cap + s] = x;
        rb[d * cap + s] = x;
        v_out[d * cap + s] = v_in[d * cap + o];
    }
}

__global__ void k_cell_bounds(const int *__restrict__ cell, long long n, int ncell, int *__restrict__ cell_start) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int c = cell[s];
    const int prev = s == 0 ? -1 : cell[s - 1];
    for (int cc = prev + 1; cc <= c; ++cc) cell_start[cc] = (int)s;
    if (s == n - 1)
        for (int cc = c + 1; cc <= ncell; ++cc) cell_start[cc] = (int)n;
}

// packed position copies for the force gathers
__device__ __forceinline__ uint32_t md_to_fixed(double x, double inv_h) {
    return (uint32_t)(unsigned long long)__double2ll_rn(x * inv_h);
}
__global__ void k_pack_positions(const double *__restrict__ r, long long n, long long cap, int fp32, double inv_h,
                                 void *__restrict__ pos4, uint4 *__restrict__ qfix) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = r[i], y = r[cap + i], z = r[2 * cap + i];
    const uint4 q = make_uint4(md_to_fixed(x, inv_h), md_to_fixed(y, inv_h), md_to_fixed(z, inv_h), 0u);
    qfix[i] = q;
    if (fp32) ((uint4 *)pos4)[i] = q;
    else ((double4 *)pos4)[i] = make_double4(x, y, z, 0.0);
}

// Neighbour-list layout: sliced ELLPACK.  Particles are grouped in tiles of 32 consecutive slots (= one warp of
// the force kernel); tile t owns max_nbr rows of 32 ints, entry (k, lane) at t*max_nbr*32 + k*32 + lane.  The
// force kernel reads one 128-byte row per k (coalesced); the build kernel's stores of one warp all land in the
// same ~15 KB tile, so partially written sectors merge in L2 instead of being evicted half empty.
__device__ __forceinline__ size_t md_nbr_index(long long i, int k, int max_nbr) {
    return ((size_t)(i >> 5) * (size_t)max_nbr + (size_t)k) * 32u + (size_t)(i & 31);
}

// ---------------------------------------------------------------------------------------------------
// K3: Verlet neighbour list.  One thread per particle; the 27-cell stencil of consecutive (cell-sorted)
// particles overlaps almost completely, so the candidate loads are L1 broadcasts.
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ void nb_scan_range(int j0, int j1, const uint4 qi, long long i, const double *__restrict__ r,
                                              const uint4 *__restrict__ q, long long cap, double L, double halfL, double rl2,
                                              float rl2s_lo, float rl2s_hi, int max_nbr, int *__restrict__ row0, int &cnt) {
#pragma unroll 2
    for (int j = j0; j < j1; ++j) {
        // prefilter in 32-bit fixed point: exact integer separations (minimum image by wrap-around), r^2 in
        // float; only candidates within 1e-5 (relative) of the list radius -- far more than the ~5e-7 the float
        // evaluation can be off -- go through the canonical float64 test
        const uint4 qj = q[j];
        const float fx = (float)(int)(qj.x - qi.x), fy = (float)(int)(qj.y - qi.y), fz = (float)(int)(qj.z - qi.z);
        const float r2f = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
        bool in = r2f < rl2s_lo;
        if (__builtin_expect(!in && r2f < rl2s_hi, 0)) {
            const double ddx = md_minimg_exact(__dsub_rn(r[j], r[i]), L, halfL);
            const double ddy = md_minimg_exact(__dsub_rn(r[cap + j], r[cap + i]), L, halfL);
            const double ddz = md_minimg_exact(__dsub_rn(r[2 * cap + j], r[2 * cap + i]), L, halfL);
            in = md_r2_canonical(ddx, ddy, ddz) < rl2;
        }
        if (in && j != (int)i) {
            if (cnt < max_nbr) row0[cnt * 32] = j;
            ++cnt;
        }
    }
}

__global__ void __launch_bounds__(NB_THREADS)
k_nbr_build(const double *__restrict__ r, const uint4 *__restrict__ q, long long n_own, long long cap,
            const int *__restrict__ cell, const int *__restrict__ cell_start, int nc, double L, double halfL, double rl2,
            float rl2s_lo, float rl2s_hi, int max_nbr, int *__restrict__ nbr, int *__restrict__ nbr_cnt,
            StepScalars *__restrict__ sc) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_own) return;
    const uint4 qi = q[i];
    const int c = cell[i];
    const int cx = c % nc, cy = (c / nc) % nc, cz = c / (nc * nc);
    int *row0 = nbr + md_nbr_index(i, 0, max_nbr);
    int cnt = 0;
    for (int dz = -1; dz <= 1; ++dz) {
        const int z = (cz + dz + nc) % nc;
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = (cy + dy + nc) % nc;
            const int row = (z * nc + y) * nc;
            // the three x-adjacent cells are one contiguous slot range unless the stencil wraps around the box
            if (cx > 0 && cx < nc - 1) {
                nb_scan_range(cell_start[row + cx - 1], cell_start[row + cx + 2], qi, i, r, q, cap, L, halfL, rl2, rl2s_lo,
                              rl2s_hi, max_nbr, row0, cnt);
            } else if (cx == 0) {
                nb_scan_range(cell_start[row + nc - 1], cell_start[row + nc], qi, i, r, q, cap, L, halfL, rl2, rl2s_lo, rl2s_hi,
                              max_nbr, row0, cnt);
                nb_scan_range(cell_start[row], cell_start[row + 2], qi, i, r, q, cap, L, halfL, rl2, rl2s_lo, rl2s_hi, max_nbr,
                              row0, cnt);
            } else {
                nb_scan_range(cell_start[row + nc - 2], cell_start[row + nc], qi, i, r, q, cap, L, halfL, rl2, rl2s_lo, rl2s_hi,
                              max_nbr, row0, cnt);
                nb_scan_range(cell_start[row], cell_start[row + 1], qi, i, r, q, cap, L, halfL, rl2, rl2s_lo, rl2s_hi, max_nbr,
                              row0, cnt);
            }
        }
    }
    nbr_cnt[i] = cnt < max_nbr ? cnt : max_nbr;
    if (cnt > max_nbr) sc->overflow_flag = 1;
}

// ---------------------------------------------------------------------------------------------------
// K4: neighbour-list forces.  One thread per particle, no atomics on forces; the list is read coalesced,
// neighbour positions are 16 B (fp32 mode) or 32 B (fp64 mode) gathers that mostly hit L1/L2 because
// slots are cell-sorted.  Borderline decisions (image tie, |r2 - rc2| within the fp32 guard band) are
// masked out of the hot loop and redone from the float64 positions, like in the all-pairs kernels.
// ---------------------------------------------------------------------------------------------------
struct LjF64 { double c6, sigma2; };   // c6 = 6 lambda eps / m
struct LjF32 { float c6, sigma2; };    // in fixed-point units (see allpairs.cu)

__global__ void __launch_bounds__(NB_THREADS)
k_nbr_force_f64(const double4 *__restrict__ pos, long long n_own, long long cap, const int *__restrict__ nbr,
                const int *__restrict__ nbr_cnt, int max_nbr, double L, double halfL, double rc2, LjF64 P,
                double *__restrict__ acc, double *__restrict__ pe_part, StepScalars *__restrict__ sc) {
    __shared__ double   red[32];
    __shared__ unsigned redu[32];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double   fx = 0.0, fy = 0.0, fz = 0.0, pe = 0.0;
    unsigned cnt = 0;
    if (i < n_own) {
        const double4 pi = pos[i];
        const int nn = nbr_cnt[i];
        const int *__restrict__ row0 = nbr + md_nbr_index(i, 0, max_nbr);
#pragma unroll 4
        for (int k = 0; k < nn; ++k) {
            const int j = row0[k * 32];
            const double4 pj = pos[j];
            const double dx = md_minimg_list(pj.x - pi.x, L, halfL);
            const double dy = md_minimg_list(pj.y - pi.y, L, halfL);
            const double dz = md_minimg_list(pj.z - pi.z, L, halfL);
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            const bool ok = r2 < rc2;
            const double inv = ok ? md_rcp(r2) : 0.0;
            const double t2 = P.sigma2 * inv;
            const double t6 = t2 * t2 * t2;
            const double s = (P.c6 * inv) * (t6 * fma(-2.0, t6, 1.0));
            fx = fma(dx, s, fx);
            fy = fma(dy, s, fy);
            fz = fma(dz, s, fz);
            pe += fma(t6, t6, -t6);
            cnt += ok ? 1u : 0u;
        }
        acc[i] = fx;
        acc[cap + i] = fy;
        acc[2 * cap + i] = fz;
    }
    const double   bpe = md_block_sum(pe, red);
    const unsigned bc = md_block_sum(cnt, redu);
    if (threadIdx.x == 0) {
        pe_part[blockIdx.x] = bpe;
        atomicAdd(&sc->pairs_directed, (unsigned long long)bc);
    }
}

__global__ void __launch_bounds__(NB_THREADS)
k_nbr_force_f32(const uint4 *__restrict__ pos, const double *__restrict__ r64, long long n_own, long long cap,
                const int *__restrict__ nbr, const int *__restrict__ nbr_cnt, int max_nbr, double L, double halfL,
                double rc2, float rc2s, double inv_h, LjF32 P, double *__restrict__ acc, double *__restrict__ pe_part,
                StepScalars *__restrict__ sc) {
    __shared__ double   red[32];
    __shared__ unsigned redu[32];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double   Fx = 0.0, Fy = 0.0, Fz = 0.0, PE = 0.0;
    unsigned cnt = 0;
    if (i < n_own) {
        const uint4 qi = pos[i];
        const int   nn = nbr_cnt[i];
        const int *__restrict__ row0 = nbr + md_nbr_index(i, 0, max_nbr);
        const float cut_lo = rc2s * (1.0f - 2e-6f), cut_hi = rc2s * (1.0f + 2e-6f);
        for (int k0 = 0; k0 < nn; k0 += 32) {
            float    fx = 0.f, fy = 0.f, fz = 0.f, pe = 0.f;
            unsigned flagged = 0u;
            const int kn = min(32, nn - k0);
#pragma unroll 8
            for (int b = 0; b < kn; ++b) {
                const int   j = row0[(k0 + b) * 32];
                const uint4 qj = pos[j];
                const int   ix = (int)(qj.x - qi.x), iy = (int)(qj.y - qi.y), iz = (int)(qj.z - qi.z);
                const float dx = (float)ix, dy = (float)iy, dz = (float)iz;
                const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                // an in-cutoff pair can only sit at the image tie if rc ~ L/2; rc + skin <= L/3 here, so only the
                // cutoff band needs the exact redo
                const bool slow = r2 > cut_lo && r2 < cut_hi;
                flagged |= (slow ? 1u : 0u) << b;
                const bool  ok = !slow && r2 < rc2s;
                const float inv = ok ? md_rcpf(r2) : 0.0f;
                const float t2 = P.sigma2 * inv;
                const float t6 = t2 * t2 * t2;
                const float s = (P.c6 * inv) * (t6 * fmaf(-2.0f, t6, 1.0f));
                fx = fmaf(dx, s, fx);
                fy = fmaf(dy, s, fy);
                fz = fmaf(dz, s, fz);
                pe += fmaf(t6, t6, -t6);
                cnt += ok ? 1u : 0u;
            }
            while (flagged) {  // rare: decide and evaluate from the float64 positions
                const int b = __ffs(flagged) - 1;
                flagged &= flagged - 1;
                const int    j = row0[(k0 + b) * 32];
                const double ex = md_minimg_list(r64[j] - r64[i], L, halfL);
                const double ey = md_minimg_list(r64[cap + j] - r64[cap + i], L, halfL);
                const double ez = md_minimg_list(r64[2 * cap + j] - r64[2 * cap + i], L, halfL);
                const double r2d = fma(ez, ez, fma(ey, ey, ex * ex));
                if (r2d < rc2) {
                    const float dx = (float)(ex * inv_h), dy = (float)(ey * inv_h), dz = (float)(ez * inv_h);
                    const float r2 = (float)(r2d * inv_h * inv_h);
                    const float inv = md_rcpf(r2);
                    const float t2 = P.sigma2 * inv;
                    const float t6 = t2 * t2 * t2;
                    const float s = (P.c6 * inv) * (t6 * fmaf(-2.0f, t6, 1.0f));
                    fx = fmaf(dx, s, fx);
                    fy = fmaf(dy, s, fy);
                    fz = fmaf(dz, s, fz);
                    pe += fmaf(t6, t6, -t6);
                    cnt += 1u;
                }
            }
            Fx += (double)fx; Fy += (double)fy; Fz += (double)fz; PE += (double)pe;
        }
        acc[i] = Fx;
        acc[cap + i] = Fy;
        acc[2 * cap + i] = Fz;
    }
    const double   bpe = md_block_sum(PE, red);
    const unsigned bc = md_block_sum(cnt, redu);
    if (threadIdx.x == 0) {
        pe_part[blockIdx.x] = bpe;
        atomicAdd(&sc->pairs_directed, (unsigned long long)bc);
    }
}

// ---------------------------------------------------------------------------------------------------
// K5 (cut-off flavour): particle-wise integrators.  Same per-component arithmetic as integrate.cu, plus
// the skin-displacement trigger, the packed position copy and the id-scatter of the dataset rows.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_cut_kick_drift(double *__restrict__ r, double *__restrict__ v, const double *__restrict__ acc,
                 const double *__restrict__ rb, long long n, long long cap, int do_bc, int do_kick, int do_drift, double L,
                 double halfL, double dt, double trig2, int fp32, double inv_h, void *__restrict__ pos4,
                 StepScalars *__restrict__ sc) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double x[3], d2 = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double xx = r[d * cap + i], u = v[d * cap + i];
        if (do_bc) {  // one-shot strict wrap, simulator.py:165-166
            if (xx < 0.0) xx += L;
            if (xx > L) xx -= L;
        }
        if (do_kick) u += __dmul_rn(__dmul_rn(0.5, acc[d * cap + i]), dt);
        if (do_drift) xx += __dmul_rn(u, dt);
        r[d * cap + i] = xx;
        if (do_kick) v[d * cap + i] = u;
        x[d] = xx;
        const double dd = md_minimg_exact(__dsub_rn(xx, rb[d * cap + i]), L, halfL);
        d2 = __dadd_rn(d2, __dmul_rn(dd, dd));
    }
    if (do_drift && d2 > trig2) sc->rebuild_flag = 1;
    if (fp32) ((uint4 *)pos4)[i] = make_uint4(md_to_fixed(x[0], inv_h), md_to_fixed(x[1], inv_h), md_to_fixed(x[2], inv_h), 0u);
    else ((double4 *)pos4)[i] = make_double4(x[0], x[1], x[2], 0.0);
}

__global__ void __launch_bounds__(256)
k_cut_kick_finish(const double *__restrict__ r, double *__restrict__ v, const double *__restrict__ acc,
                  const int *__restrict__ id, long long n, long long cap, long long n_global, double dt, double g_dt,
                  const StepScalars *__restrict__ sc, double *__restrict__ ke_part, double *__restrict__ r_row,
                  double *__restrict__ v_row) {
    __shared__ double red[32];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double v2 = 0.0;
    if (i < n) {
        const double scale = sc->scale;
        const int    gid = id[i];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            double u = v[d * cap + i] + __dmul_rn(__dmul_rn(0.5, acc[d * cap + i]), dt);
            if (d == 2 && g_dt != 0.0) u -= g_dt;
            if (scale != 1.0) u *= scale;
            v[d * cap + i] = u;
            v2 += u * u;
            if (r_row) {  // rows are kept in ORIGINAL particle order (core.py:236-238)
                r_row[d * n_global + gid] = r[d * cap + i];
                v_row[d * n_global + gid] = u;
            }
        }
    }
    const double b = md_block_sum(v2, red);
    if (threadIdx.x == 0) ke_part[blockIdx.x] = b;
}

__global__ void k_unsort(const double *__restrict__ src, const int *__restrict__ id, long long n, long long cap,
                         long long n_global, double *__restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int gid = id[i];
#pragma unroll
    for (int d = 0; d < 3; ++d) dst[d * n_global + gid] = src[d * cap + i];
}

__global__ void k_iota(int *a, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (int)i;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int bits_for(unsigned long long maxval) {
    int b = 1;
    while ((maxval >> b) != 0) ++b;
    return b;
}

int md_cells_setup(mdsea_ctx *c) {
    CellState *s = new CellState();
    c->cells = s;
    const mdsea_params &p = c->p;
    const double skin = p.skin > 0 ? p.skin : 0.0;
    s->rlist = p.r_cutoff + skin;
    s->rl2 = s->rlist * s->rlist;
    s->rc2 = p.r_cutoff * p.r_cutoff;
    s->trig2 = (0.5 * skin) * (0.5 * skin);
    s->nc = (int)floor(p.boxlen / s->rlist);
    s->cell_len = p.boxlen / s->nc;
    s->ncell = s->nc * s->nc * s->nc;
    const long long n = c->n;
    s->cap = ((n + 31) / 32) * 32;
    s->n_pad = s->cap;
    if (p.max_neighbors > 0) s->max_nbr = p.max_neighbors;
    else {
        const double rho = (double)c->n_global / (p.boxlen * p.boxlen * p.boxlen);
        const double expect = 4.0 / 3.0 * 3.14159265358979 * s->rlist * s->rlist * s->rlist * rho;
        s->max_nbr = ((int)(expect * 1.6) + 32 + 7) / 8 * 8;
    }
    s->id_bits = bits_for((unsigned long long)(c->n_global - 1));
    s->cell_bits = bits_for((unsigned long long)(s->ncell - 1));
    s->sort_blocks = md_div_up(s->cap, SORT_CHUNK);
    const size_t S = sizeof(double) * 3 * (size_t)s->cap;
    MD_CUDA(c, cudaMalloc(&c->d_r, S));
    MD_CUDA(c, cudaMalloc(&c->d_v, S));
    MD_CUDA(c, cudaMalloc(&c->d_acc, S));
    MD_CUDA(c, cudaMemset(c->d_acc, 0, S));
    MD_CUDA(c, cudaMalloc(&s->r_alt, S));
    MD_CUDA(c, cudaMalloc(&s->v_alt, S));
    MD_CUDA(c, cudaMalloc(&s->rb, S));
    MD_CUDA(c, cudaMalloc(&s->id, sizeof(int) * s->cap));
    MD_CUDA(c, cudaMalloc(&s->id_alt, sizeof(int) * s->cap));
    MD_CUDA(c, cudaMalloc(&s->pos4, (p.precision == MDSEA_FP32 ? sizeof(uint4) : sizeof(double4)) * (size_t)s->cap));
    MD_CUDA(c, cudaMalloc(&s->key, sizeof(unsigned long long) * s->cap));
    MD_CUDA(c, cudaMalloc(&s->key_alt, sizeof(unsigned long long) * s->cap));
    MD_CUDA(c, cudaMalloc(&s->val, sizeof(int) * s->cap));
    MD_CUDA(c, cudaMalloc(&s->val_alt, sizeof(int) * s->cap));
    MD_CUDA(c, cudaMalloc(&s->hist, sizeof(int) * 256 * (size_t)s->sort_blocks));
    MD_CUDA(c, cudaMalloc(&s->hist_tot, sizeof(int) * 256));
    MD_CUDA(c, cudaMalloc(&s->qfix, sizeof(uint4) * (size_t)s->cap));
    MD_CUDA(c, cudaMalloc(&s->cell, sizeof(int) * s->cap));
    MD_CUDA(c, cudaMalloc(&s->cell_start, sizeof(int) * ((size_t)s->ncell + 1)));
    MD_CUDA(c, cudaMalloc(&s->nbr, sizeof(int) * (size_t)s->max_nbr * (size_t)s->n_pad));
    MD_CUDA(c, cudaMalloc(&s->nbr_cnt, sizeof(int) * s->cap));
    MD_CUDA(c, cudaMalloc(&s->scratch, sizeof(double) * 3 * (size_t)c->n_global));
    MD_CUDA(c, cudaMallocHost(&s->h_flag, sizeof(int) * 2));
    const size_t rowE = (size_t)3 * (size_t)c->n_global;
    MD_CUDA(c, cudaMalloc(&c->d_r_rows, sizeof(double) * rowE * c->ring));
    MD_CUDA(c, cudaMalloc(&c->d_v_rows, sizeof(double) * rowE * c->ring));
    c->n_ke_part = md_div_up(n, 256);
    MD_CUDA(c, cudaMalloc(&c->d_ke_part, sizeof(double) * c->n_ke_part));
    c->n_pe_part = md_div_up(n, NB_THREADS);
    MD_CUDA(c, cudaMalloc(&c->d_pe_part, sizeof(double) * c->n_pe_part));
    MD_CUDA(c, cudaMalloc(&c->d_cnt_part, sizeof(unsigned long long) * c->n_pe_part));
    MD_CUDA(c, cudaMemset(c->d_pe_part, 0, sizeof(double) * c->n_pe_part));
    MD_CUDA(c, cudaMemset(c->d_cnt_part, 0, sizeof(unsigned long long) * c->n_pe_part));
    if (c->pot.fast != POTF_LJ) {
        md_set_error(c, "cutoff path: only the Lennard-Jones member of the Mie family is implemented on the list kernels");
        return MDSEA_ERR_UNSUPPORTED;
    }
    return MDSEA_OK;
}

void md_cells_free(mdsea_ctx *c) {
    CellState *s = c->cells;
    if (!s) return;
    cudaFree(s->r_alt); cudaFree(s->v_alt); cudaFree(s->rb); cudaFree(s->id); cudaFree(s->id_alt); cudaFree(s->pos4);
    cudaFree(s->key); cudaFree(s->key_alt); cudaFree(s->val); cudaFree(s->val_alt); cudaFree(s->hist); cudaFree(s->hist_tot); cudaFree(s->qfix);
    cudaFree(s->cell); cudaFree(s->cell_start); cudaFree(s->nbr); cudaFree(s->nbr_cnt); cudaFree(s->scratch);
    if (s->h_flag) cudaFreeHost(s->h_flag);
    delete s;
    c->cells = nullptr;
}

static inline double fixed_inv_h(const mdsea_ctx *c) { return 4294967296.0 / c->box.L; }

int md_cutoff_set_state(mdsea_ctx *c, const double *r, const double *v) {
    CellState *s = c->cells;
    const long long n = c->n;
    for (int d = 0; d < 3; ++d) {
        MD_CUDA(c, cudaMemcpyAsync(c->d_r + d * s->cap, r + d * n, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
        MD_CUDA(c, cudaMemcpyAsync(c->d_v + d * s->cap, v + d * n, sizeof(double) * n, cudaMemcpyHostToDevice, c->stream));
    }
    k_iota<<<md_div_up(n, 256), 256, 0, c->stream>>>(s->id, n);
    c->stats.kernel_launches += 1;
    MD_CUDA(c, cudaStreamSynchronize(c->stream));
    s->need_rebuild = true;
    c->stats.n_local = n;
    return MDSEA_OK;
}

static int rebuild(mdsea_ctx *c) {
    CellState *s = c->cells;
    const long long n = c->n;
    md_prof_begin(c);
    const int g256 = md_div_up(n, 256);
    k_cell_keys<<<g256, 256, 0, c->stream>>>(c->d_r, s->id, n, s->cap, c->box.L, s->cell_len, s->nc, s->key, s->val);
    int launches = 1;
    // LSD passes over the id bits, then over the cell bits (= stable sort by cell id, then by particle id)
    std::vector<int> shifts;
    for (int b = 0; b < s->id_bits; b += 8) shifts.push_back(b);
    for (int b = 0; b < s->cell_bits; b += 8) shifts.push_back(32 + b);
    const int nblk = md_div_up(n, SORT_CHUNK);
    for (int sh : shifts) {
        k_sort_hist<<<nblk, SORT_THREADS, 0, c->stream>>>(s->key, n, sh, s->hist, nblk);
        k_sort_rowscan<<<256, 256, 0, c->stream>>>(s->hist, nblk, s->hist_tot);
        k_sort_scatter<<<nblk, SORT_THREADS, 0, c->stream>>>(s->key, s->val, s->key_alt, s->val_alt, n, sh, s->hist, nblk,
                                                             s->hist_tot);
        std::swap(s->key, s->key_alt);
        std::swap(s->val, s->val_alt);
        launches += 3;
    }
    k_reorder<<<g256, 256, 0, c->stream>>>(s->key, s->val, n, s->cap, c->d_r, c->d_v, s->r_alt, s->v_alt, s->rb, s->id_alt,
                                           s->cell);
    std::swap(c->d_r, s->r_alt);
    std::swap(c->d_v, s->v_alt);
    std::swap(s->id, s->id_alt);
    k_cell_bounds<<<g256, 256, 0, c->stream>>>(s->cell, n, s->ncell, s->cell_start);
    k_pack_positions<<<g256, 256, 0, c->stream>>>(c->d_r, n, s->cap, c->p.precision == MDSEA_FP32, fixed_inv_h(c), s->pos4, s->qfix);
    {
        const double ih = fixed_inv_h(c);
        const double rl2s = s->rl2 * ih * ih;
        k_nbr_build<<<md_div_up(n, NB_THREADS), NB_THREADS, 0, c->stream>>>(
            c->d_r, s->qfix, n, s->cap, s->cell, s->cell_start, s->nc, c->box.L, c->box.halfL, s->rl2,
            (float)(rl2s * (1.0 - 1e-5)), (float)(rl2s * (1.0 + 1e-5)), s->max_nbr, s->nbr, s->nbr_cnt, c->d_sc);
    }
    launches += 4;
    MD_CUDA(c, cudaMemsetAsync(&c->d_sc->rebuild_flag, 0, sizeof(int), c->stream));
    md_prof_end(c, &c->stats.ms_rebuild);
    MD_CUDA(c, cudaGetLastError());
    c->stats.kernel_launches += launches;
    c->stats.list_rebuilds += 1;
    s->need_rebuild = false;
    return MDSEA_OK;
}

int md_cutoff_force(mdsea_ctx *c) {
    CellState *s = c->cells;
    if (s->need_rebuild) MD_TRY(rebuild(c));
    const long long n = c->n;
    md_prof_begin(c);
    const int grid = md_div_up(n, NB_THREADS);
    if (c->p.precision == MDSEA_FP64) {
        LjF64 P;
        P.c6 = 6.0 * c->pot.lam_eps * c->pot.inv_mass;
        P.sigma2 = c->pot.sigma2;
        k_nbr_force_f64<<<grid, NB_THREADS, 0, c->stream>>>((const double4 *)s->pos4, n, s->cap, s->nbr, s->nbr_cnt, s->max_nbr,
                                                            c->box.L, c->box.halfL, s->rc2, P, c->d_acc, c->d_pe_part, c->d_sc);
    } else {
        const double inv_h = fixed_inv_h(c);
        LjF32 P;
        P.c6 = (float)(6.0 * c->pot.lam_eps * c->pot.inv_mass * inv_h);
        P.sigma2 = (float)(c->pot.sigma2 * inv_h * inv_h);
        k_nbr_force_f32<<<grid, NB_THREADS, 0, c->stream>>>((const uint4 *)s->pos4, c->d_r, n, s->cap, s->nbr, s->nbr_cnt,
                                                            s->max_nbr, c->box.L, c->box.halfL, s->rc2,
                                                            (float)(s->rc2 * inv_h * inv_h), inv_h, P, c->d_acc, c->d_pe_part,
                                                            c->d_sc);
    }
    md_prof_end(c, &c->stats.ms_force);
    MD_CUDA(c, cudaGetLastError());
    c->stats.kernel_launches += 1;
    c->stats.force_evals += 1;
    return MDSEA_OK;
}

static int cut_kick_drift(mdsea_ctx *c, bool do_bc, bool do_kick, bool do_drift) {
    CellState *s = c->cells;
    md_prof_begin(c);
    k_cut_kick_drift<<<md_div_up(c->n, 256), 256, 0, c->stream>>>(c->d_r, c->d_v, c->d_acc, s->rb, c->n, s->cap, do_bc, do_kick,
                                                                  do_drift, c->box.L, c->box.halfL, c->p.dt, s->trig2,
                                                                  c->p.precision == MDSEA_FP32, fixed_inv_h(c), s->pos4, c->d_sc);
    md_prof_end(c, &c->stats.ms_integrate);
    MD_CUDA(c, cudaGetLastError());
    c->stats.kernel_launches += 1;
    return MDSEA_OK;
}

// after a drift: fetch the device-side trigger; rebuild before the next force evaluation if it fired
static int check_trigger(mdsea_ctx *c) {
    CellState *s = c->cells;
    if (s->need_rebuild) return MDSEA_OK;  // no list yet: the trigger was computed against stale build positions
    MD_CUDA(c, cudaMemcpyAsync(s->h_flag, &c->d_sc->rebuild_flag, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    MD_CUDA(c, cudaStreamSynchronize(c->stream));
    if (s->h_flag[0]) {
        s->need_rebuild = true;
        MD_CUDA(c, cudaMemsetAsync(&c->d_sc->rebuild_flag, 0, sizeof(int), c->stream));
    }
    return MDSEA_OK;
}

int md_cutoff_steps(mdsea_ctx *c, const mdsea_run_args *a, double g_dt) {
    CellState *s = c->cells;
    const bool verlet = c->p.algorithm == MDSEA_ALGO_VERLET;
    const bool reuse = verlet && c->p.force_evals_per_step == 1;
    const size_t rowE = (size_t)3 * (size_t)c->n_global;
    for (int k = 0; k < a->nsteps; ++k) {
        if (verlet) {
            if (reuse && c->acc_valid) {
                MD_TRY(cut_kick_drift(c, true, true, true));
            } else {
                MD_TRY(cut_kick_drift(c, true, false, false));   // apply_bc
                MD_TRY(md_cutoff_force(c));                       // simulator.py:540
                MD_TRY(cut_kick_drift(c, false, true, true));
            }
        } else {
            MD_TRY(cut_kick_drift(c, true, false, true));        // simulator.py:533
        }
        MD_TRY(check_trigger(c));
        MD_TRY(md_cutoff_force(c));                               // simulator.py:544
        double *rr = a->record_coords ? c->d_r_rows + (size_t)k * rowE : nullptr;
        double *vr = a->record_coords ? c->d_v_rows + (size_t)k * rowE : nullptr;
        md_prof_begin(c);
        k_cut_kick_finish<<<md_div_up(c->n, 256), 256, 0, c->stream>>>(c->d_r, c->d_v, c->d_acc, s->id, c->n, s->cap,
                                                                       c->n_global, c->p.dt, g_dt, c->d_sc, c->d_ke_part, rr, vr);
        md_prof_end(c, &c->stats.ms_integrate);
        MD_CUDA(c, cudaGetLastError());
        c->stats.kernel_launches += 1;
        c->acc_valid = true;
        MD_TRY(md_finalize_step(c, k, k + 1 < a->nsteps, a->k_boltzmann, c->stats.steps + k));
    }
    return MDSEA_OK;
}

int md_cutoff_finish_rows(mdsea_ctx *c) { return MDSEA_OK; }

int md_cutoff_get_state(mdsea_ctx *c, double *r, double *v, double *a, int64_t *ids, int64_t *n_owned) {
    CellState *s = c->cells;
    const long long n = c->n, N = c->n_global;
    const int g = md_div_up(n, 256);
    const double *src[3] = {c->d_r, c->d_v, c->d_acc};
    double *dst[3] = {r, v, a};
    for (int q = 0; q < 3; ++q) {
        if (!dst[q]) continue;
        k_unsort<<<g, 256, 0, c->stream>>>(src[q], s->id, n, s->cap, N, s->scratch);
        c->stats.kernel_launches += 1;
        MD_CUDA(c, cudaMemcpyAsync(dst[q], s->scratch, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, c->stream));
        MD_CUDA(c, cudaStreamSynchronize(c->stream));
    }
    if (ids || n_owned) {
        std::vector<int> h(n);
        MD_CUDA(c, cudaMemcpy(h.data(), s->id, sizeof(int) * n, cudaMemcpyDeviceToHost));
        if (ids) {
            for (long long i = 0; i < n; ++i) ids[i] = h[i];
            std::sort(ids, ids + n);
        }
        if (n_owned) *n_owned = n;
    }
    return MDSEA_OK;
}

// ---------------------------------------------------------------------------------------------------
// test hooks (bit-exact contracts)
// ---------------------------------------------------------------------------------------------------
extern "C" int mdsea_b200_get_cells(mdsea_ctx *c, int32_t *cell_of, int32_t *order, int32_t ncell_dim[3]) {
    if (!c || !cell_of || !order) return MDSEA_ERR_BAD_ARG;
    if (!c->cutoff_path || !c->have_state) { md_set_error(c, "get_cells: needs the cutoff path and a state"); return MDSEA_ERR_STATE; }
    MD_CUDA(c, cudaSetDevice(c->p.device));
    CellState *s = c->cells;
    if (s->need_rebuild) MD_TRY(rebuild(c));
    const long long n = c->n;
    std::vector<int> id(n), cell(n);
    MD_CUDA(c, cudaStreamSynchronize(c->stream));
    MD_CUDA(c, cudaMemcpy(id.data(), s->id, sizeof(int) * n, cudaMemcpyDeviceToHost));
    MD_CUDA(c, cudaMemcpy(cell.data(), s->cell, sizeof(int) * n, cudaMemcpyDeviceToHost));
    for (long long k = 0; k < n; ++k) {
        order[k] = id[k];
        cell_of[id[k]] = cell[k];
    }
    if (ncell_dim) ncell_dim[0] = ncell_dim[1] = ncell_dim[2] = s->nc;
    return MDSEA_OK;
}

extern "C" int mdsea_b200_get_neighbors(mdsea_ctx *c, int64_t *n_entries, int64_t *offsets, int32_t *idx) {
    if (!c || !n_entries || !offsets) return MDSEA_ERR_BAD_ARG;
    if (!c->cutoff_path || !c->have_state) { md_set_error(c, "get_neighbors: needs the cutoff path and a state"); return MDSEA_ERR_STATE; }
    MD_CUDA(c, cudaSetDevice(c->p.device));
    CellState *s = c->cells;
    if (s->need_rebuild) MD_TRY(rebuild(c));
    const long long n = c->n;
    MD_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<int> id(n), cnt(n);
    MD_CUDA(c, cudaMemcpy(id.data(), s->id, sizeof(int) * n, cudaMemcpyDeviceToHost));
    MD_CUDA(c, cudaMemcpy(cnt.data(), s->nbr_cnt, sizeof(int) * n, cudaMemcpyDeviceToHost));
    // CSR over ORIGINAL ids
    std::vector<int64_t> count_by_id(c->n_global, 0);
    for (long long k = 0; k < n; ++k) count_by_id[id[k]] = cnt[k];
    offsets[0] = 0;
    for (long long g = 0; g < c->n_global; ++g) offsets[g + 1] = offsets[g] + count_by_id[g];
    *n_entries = offsets[c->n_global];
    if (!idx) return MDSEA_OK;
    std::vector<int> nbr((size_t)s->max_nbr * (size_t)s->n_pad);
    MD_CUDA(c, cudaMemcpy(nbr.data(), s->nbr, sizeof(int) * nbr.size(), cudaMemcpyDeviceToHost));
    for (long long k = 0; k < n; ++k) {
        int32_t *row = idx + offsets[id[k]];
        for (int m = 0; m < cnt[k]; ++m) row[m] = id[nbr[((size_t)(k >> 5) * s->max_nbr + m) * 32 + (k & 31)]];
        std::sort(row, row + cnt[k]);
    }
    return MDSEA_OK;
}

extern "C" int mdsea_b200_nccl_unique_id(void *) { return MDSEA_ERR_UNSUPPORTED; }
extern "C" int mdsea_b200_comm_init(mdsea_ctx *c, const void *) {
    md_set_error(c, "comm_init: multi-GPU domain decomposition is not built yet");
    return MDSEA_ERR_UNSUPPORTED;
}
