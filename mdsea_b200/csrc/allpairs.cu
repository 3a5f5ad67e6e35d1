// K1: tiled all-pairs force + potential-energy kernels (small N, no neighbour list).
//
// Replaces, per force evaluation, the reference's update_dists (simulator.py:257-280: pair table
// gathers, minimum image through rint of a true division, norm), Potential.force
// (potentials.py:43-48) and the bincount scatter of update_acc (simulator.py:295-311), plus the
// sum of update_mean_pe (simulator.py:247-250).  Nothing of size O(N^2) is ever stored:
// every thread owns one particle i, j-particles are staged tile by tile in shared memory
// (broadcast reads), and the j range is split over blockIdx.y so that N = 4096 still fills
// 148 SMs.  Each (i, j-chunk) partial force is written once, coalesced, to
// fpart[chunk][dim][i]; the integrator sums the chunks in fixed order (no atomics, run-to-run
// deterministic).  Sign convention: dr = r_j - r_i, a_i += dr * (dV/dr)/(r m)  (simulator.py:266,301-305).
#include <stdlib.h>
#include <string.h>

#include "ctx.h"
#include "pair.cuh"

#define AP_TI 128  // particles i per block == j tile length

// ---------------------------------------------------------------------------------------------
// fp64 kernel
// ---------------------------------------------------------------------------------------------
// exact reference image of one component (rare path, out of line): d - rint(d / L) * L, IEEE division
__device__ __noinline__ double ap_exact_image(double x0, double L) {
    return __dsub_rn(x0, __dmul_rn(rint(__ddiv_rn(x0, L)), L));
}

// one j tile against this thread's particle.  SELF: the tile contains particle i itself (diagonal tile).
// The hot loop is branch-free so that ptxas can interleave the unrolled pairs (the fp64 pipe has a long
// dependent chain per pair): the minimum image uses rint(x * (1/L)); a pair whose imaged component lands
// within ~1e-6 L of the half-box tie -- where rint(x / L) of the reference could choose the other image --
// is masked out of the hot loop, remembered in a 128-bit per-thread mask, and redone afterwards with the
// exact IEEE expression of simulator.py:272.
template <int NDIM, int POTK, bool CUT, bool PBC, bool SELF>
__device__ __forceinline__ void ap_tile_f64(const double (*sj)[AP_TI], int nj, int self_k, const double *ri,
                                            const BoxDev &box, const PotF64 &P, double rc2, int tie_hi,
                                            double *f, double &pe, unsigned &cnt) {
    const int L_hi = __double2hiint(box.L), L_lo = __double2loint(box.L);
#pragma unroll 1
    for (int k0 = 0; k0 < nj; k0 += 32) {
        unsigned flagged = 0u;
        if (k0 + 32 <= nj) {
#pragma unroll 8
            for (int b = 0; b < 32; ++b) {
                const int k = k0 + b;
                double dd[NDIM], r2 = 0.0;
                int    amax = 0;
#pragma unroll
                for (int d = 0; d < NDIM; ++d) {
                    double x = sj[d][k] - ri[d];
                    if (PBC) {
                        // |x| < 1.5 L for any sane state, so rint(x/L) is -1, 0 or +1: subtract copysign(L, x)
                        // when |x| > L/2 (sign transplant on the integer pipe, one DSETP + one predicated DADD).
                        // Anything that is still >= ~L/2 afterwards (near-tie, or |x| >= 1.5 L) is flagged.
                        const int    hi = __double2hiint(x);
                        const double Ls = __hiloint2double((hi & 0x80000000) | L_hi, L_lo);
                        if (fabs(x) > box.halfL) x -= Ls;
                        amax = max(amax, __double2hiint(x) & 0x7fffffff);
                    }
                    dd[d] = x;
                    r2 = fma(x, x, r2);
                }
                bool ok = true;
                if (PBC) {
                    const bool tie = amax >= tie_hi;
                    flagged |= (tie ? 1u : 0u) << b;
                    ok = !tie;
                }
                if (SELF) ok = ok && (k != self_k);
                if (CUT) ok = ok && (r2 < rc2);
                double s, u;
                md_pair<POTK>(r2, ok, P, s, u);
#pragma unroll
                for (int d = 0; d < NDIM; ++d) f[d] = fma(dd[d], s, f[d]);
                pe += u;
                cnt += ok ? 1u : 0u;
            }
        } else {  // ragged tail of the last tile: everything through the exact path below
            flagged = (1u << (nj - k0)) - 1u;
        }
        // rare: exact redo of the masked pairs
        while (flagged) {
            const int b = __ffs(flagged) - 1;
            flagged &= flagged - 1;
            const int k = k0 + b;
            double dd[NDIM], r2 = 0.0;
#pragma unroll
            for (int d = 0; d < NDIM; ++d) {
                double x = sj[d][k] - ri[d];
                if (PBC) x = ap_exact_image(x, box.L);
                dd[d] = x;
                r2 = fma(x, x, r2);
            }
            bool ok = !SELF || (k != self_k);
            if (CUT) ok = ok && (r2 < rc2);
            double s, u;
            md_pair<POTK>(r2, ok, P, s, u);
#pragma unroll
            for (int d = 0; d < NDIM; ++d) f[d] = fma(dd[d], s, f[d]);
            pe += u;
            cnt += ok ? 1u : 0u;
        }
    }
}

template <int NDIM, int POTK, bool CUT, bool PBC>
__global__ void __launch_bounds__(AP_TI)
k_allpairs_f64(const double *__restrict__ r, int n, int jchunk, BoxDev box, PotF64 P, double rc2,
               double *__restrict__ fpart, double *__restrict__ pe_part,
               unsigned long long *__restrict__ cnt_part, StepScalars *__restrict__ sc) {
    __shared__ double sj[NDIM][AP_TI];
    __shared__ double red[32];
    __shared__ unsigned redu[32];

    const int  i = blockIdx.x * AP_TI + threadIdx.x;
    const bool active = i < n;
    const int  ic = active ? i : n - 1;
    double ri[NDIM], f[NDIM];
#pragma unroll
    for (int d = 0; d < NDIM; ++d) {
        ri[d] = r[(size_t)d * n + ic];
        f[d] = 0.0;
    }
    double   pe = 0.0;
    unsigned cnt = 0;
    const int j0 = blockIdx.y * jchunk;   // jchunk is a multiple of AP_TI: tiles are aligned with the i blocks
    const int j1 = min(n, j0 + jchunk);
    // any |x| >= L/2 (1 - 2e-6) shares or exceeds this high word (20 mantissa bits: band ~1e-6 L, conservative)
    const int tie_hi = __double2hiint(box.halfL * (1.0 - 2e-6));

    for (int jt = j0; jt < j1; jt += AP_TI) {
        __syncthreads();
        const int jl = jt + threadIdx.x;
        if (jl < j1) {
#pragma unroll
            for (int d = 0; d < NDIM; ++d) sj[d][threadIdx.x] = r[(size_t)d * n + jl];
        }
        __syncthreads();
        const int nj = min(AP_TI, j1 - jt);
        if (jt == (int)(blockIdx.x * AP_TI))
            ap_tile_f64<NDIM, POTK, CUT, PBC, true>(sj, nj, ic - jt, ri, box, P, rc2, tie_hi, f, pe, cnt);
        else
            ap_tile_f64<NDIM, POTK, CUT, PBC, false>(sj, nj, -1, ri, box, P, rc2, tie_hi, f, pe, cnt);
    }
    if (active) {
#pragma unroll
        for (int d = 0; d < NDIM; ++d) fpart[((size_t)blockIdx.y * NDIM + d) * n + i] = f[d];
    } else {
        pe = 0.0;
        cnt = 0;
    }
    double   bpe = md_block_sum(pe, red);
    unsigned bc = md_block_sum(cnt, redu);
    if (threadIdx.x == 0) {
        const int b = blockIdx.y * gridDim.x + blockIdx.x;
        pe_part[b] = bpe;
        cnt_part[b] = bc;
        atomicAdd(&sc->pairs_pending, (unsigned long long)bc);  // integer: order-independent, stats only
    }
}

// ---------------------------------------------------------------------------------------------
// fp64 kernel, Newton's third law (small N: up to AP3_MAXT tiles): one CTA per UNORDERED pair of
// 128-particle tiles (I <= J), every pair evaluated ONCE -- half the fp64 work of the kernel above, which is what
// bounds it.  Thread t owns particle i = I*128 + t (force in registers).  The reaction -f goes to the j particle
// through shared memory without atomics: in batch b, lane l of a warp visits the four j = u*32 + ((l + b) & 31),
// u = 0..3 -- distinct for the 32 lanes of the warp -- and adds into the WARP's private copy of the tile's force
// array; __syncwarp() orders the batches; the four copies are summed in fixed order at the end.  Partial forces go
// to fpart[partner tile][dim][particle]: every (partner, particle) slot is written by exactly one CTA per evaluation,
// and the integrator sums the T partners in fixed order (deterministic, like the split-j kernel).
// ---------------------------------------------------------------------------------------------
#define AP3_MAXT 64
template <int NDIM, int POTK, bool CUT, bool PBC>
__global__ void __launch_bounds__(AP_TI)
k_allpairs_n3_f64(const double *__restrict__ r, int n, BoxDev box, PotF64 P, double rc2, double *__restrict__ fpart,
                  double *__restrict__ pe_part, unsigned long long *__restrict__ cnt_part, StepScalars *__restrict__ sc) {
    __shared__ double sj[NDIM][AP_TI];
    __shared__ double sfj[AP_TI / 32][NDIM][AP_TI];
    __shared__ double red[32];
    __shared__ unsigned redu[32];
    const int J = blockIdx.x, I = blockIdx.y, bidx = blockIdx.y * gridDim.x + blockIdx.x;
    if (I > J) {   // lower triangle: nothing to do, but the partial-sum slots are read by the finalize kernel
        if (threadIdx.x == 0) { pe_part[bidx] = 0.0; cnt_part[bidx] = 0ull; }
        return;
    }
    const int  t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int  i = I * AP_TI + t, jg = J * AP_TI + t;
    const bool active = i < n, diag = I == J;
    const int  ic = active ? i : n - 1, nj = min(AP_TI, n - J * AP_TI);
    double ri[NDIM], f[NDIM];
#pragma unroll
    for (int d = 0; d < NDIM; ++d) {
        ri[d] = r[(size_t)d * n + ic];
        f[d] = 0.0;
        sj[d][t] = r[(size_t)d * n + min(jg, n - 1)];
#pragma unroll
        for (int ww = 0; ww < AP_TI / 32; ++ww) sfj[ww][d][t] = 0.0;
    }
    __syncthreads();
    double   pe = 0.0;
    unsigned cnt = 0;
    const int L_hi = __double2hiint(box.L), L_lo = __double2loint(box.L);
    const int tie_hi = __double2hiint(box.halfL * (1.0 - 2e-6));
    for (int b = 0; b < 32; ++b) {
        const int g = (lane + b) & 31;
        double   dd[4][NDIM], s4[4];
        unsigned tie_mask = 0u;
#pragma unroll
        for (int u = 0; u < 4; ++u) {   // four independent pairs: keeps the fp64 pipe fed
            const int jl = u * 32 + g;
            double r2 = 0.0;
            int    amax = 0;
#pragma unroll
            for (int d = 0; d < NDIM; ++d) {
                double x = sj[d][jl] - ri[d];
                if (PBC) {   // same branch-free image as ap_tile_f64; near-tie components are redone exactly below
                    const int    hi = __double2hiint(x);
                    const double Ls = __hiloint2double((hi & 0x80000000) | L_hi, L_lo);
                    if (fabs(x) > box.halfL) x -= Ls;
                    amax = max(amax, __double2hiint(x) & 0x7fffffff);
                }
                dd[u][d] = x;
                r2 = fma(x, x, r2);
            }
            bool ok = active && jl < nj && (!diag || jl > t);
            if (PBC && amax >= tie_hi) {
                tie_mask |= ok ? (1u << u) : 0u;
                ok = false;
            }
            if (CUT) ok = ok && (r2 < rc2);
            double s, uu;
            md_pair<POTK>(r2, ok, P, s, uu);
            s4[u] = s;
            pe += uu;
            cnt += ok ? 2u : 0u;
        }
        if (tie_mask) {   // rare: exact IEEE image (simulator.py:272) for a pair at the half-box tie
#pragma unroll
            for (int u = 0; u < 4; ++u) {   // static indices keep dd / s4 in registers
                if (!((tie_mask >> u) & 1u)) continue;
                const int jl = u * 32 + g;
                double r2 = 0.0;
#pragma unroll
                for (int d = 0; d < NDIM; ++d) {
                    const double x = ap_exact_image(sj[d][jl] - ri[d], box.L);
                    dd[u][d] = x;
                    r2 = fma(x, x, r2);
                }
                const bool ok = !CUT || r2 < rc2;
                double s, uu;
                md_pair<POTK>(r2, ok, P, s, uu);
                s4[u] = s;
                pe += uu;
                cnt += ok ? 2u : 0u;
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int jl = u * 32 + g;
#pragma unroll
            for (int d = 0; d < NDIM; ++d) {
                const double fd = dd[u][d] * s4[u];
                f[d] += fd;                 // a_i += dr * s   (dr = r_j - r_i)
                sfj[w][d][jl] -= fd;        // a_j -= dr * s   (this warp's copy; lanes hit distinct j)
            }
        }
        __syncwarp();
    }
    __syncthreads();
    double fj[NDIM];
#pragma unroll
    for (int d = 0; d < NDIM; ++d) {
        double a = sfj[0][d][t];
#pragma unroll
        for (int ww = 1; ww < AP_TI / 32; ++ww) a += sfj[ww][d][t];
        fj[d] = a;
    }
    if (diag) {
        if (active) {
#pragma unroll
            for (int d = 0; d < NDIM; ++d) fpart[((size_t)I * NDIM + d) * n + i] = f[d] + fj[d];
        }
    } else {
        if (active) {
#pragma unroll
            for (int d = 0; d < NDIM; ++d) fpart[((size_t)J * NDIM + d) * n + i] = f[d];
        }
        if (jg < n) {
#pragma unroll
            for (int d = 0; d < NDIM; ++d) fpart[((size_t)I * NDIM + d) * n + jg] = fj[d];
        }
    }
    // every unordered pair was evaluated once: count its energy for both directions (the finalize factor is
    // lambda eps / 2 / N over DIRECTED pairs, like the split-j kernel)
    const double   bpe = md_block_sum(pe, red);
    const unsigned bc = md_block_sum(cnt, redu);
    if (threadIdx.x == 0) {
        pe_part[bidx] = 2.0 * bpe;
        cnt_part[bidx] = bc;
        atomicAdd(&sc->pairs_pending, (unsigned long long)bc);
    }
}

// ---------------------------------------------------------------------------------------------
// fp32 kernel: positions as 32-bit fixed point (unit h = L/2^32 periodic, L/2^31 walls).
// The integer subtraction of two fixed-point coordinates is exact and -- for the periodic box --
// wraps to the minimum image for free; I2F then rounds RELATIVE to the separation, which keeps
// the pair geometry at ~6e-8 relative where fp32 absolute coordinates would lose 1e-6 * L.
// Pairs within 2^12 units of the half-box tie, and pairs within 2e-6 (relative) of the cutoff,
// are re-evaluated from the fp64 positions so that the image and in/out-of-cutoff decisions are
// exactly those of the fp64 path.  Per-tile fp32 partial sums are folded into fp64 accumulators.
// ---------------------------------------------------------------------------------------------
// exact (fp64) separation of one component of one pair, rare path
template <bool PBC>
__device__ __noinline__ double ap_exact_sep(const double *__restrict__ r64, size_t j, size_t i, double L, double invL) {
    double x = r64[j] - r64[i];
    if (PBC) x = md_minimg(x, L, invL);
    return x;
}

template <int NDIM, int POTK, bool CUT, bool PBC, bool SELF>
__device__ __forceinline__ void ap_tile_f32(const uint32_t (*sj)[AP_TI], int nj, int self_k, int jt, int ic, int n,
                                            const uint32_t *qi, const double *__restrict__ r64, const BoxDev &box,
                                            const PotF32 &P, float rc2s, float cut_lo, float cut_hi, double rc2,
                                            double inv_h, float *f, float &pe, unsigned &cnt) {
#pragma unroll 1
    for (int k0 = 0; k0 < nj; k0 += 32) {
        unsigned flagged = 0u;
        if (k0 + 32 <= nj) {
#pragma unroll 8
            for (int b = 0; b < 32; ++b) {
                const int k = k0 + b;
                float    dd[NDIM], r2 = 0.0f;
                uint32_t amax = 0;
#pragma unroll
                for (int d = 0; d < NDIM; ++d) {
                    const int di = (int)(sj[d][k] - qi[d]);
                    if (PBC) amax = max(amax, (uint32_t)abs(di));
                    const float x = (float)di;
                    dd[d] = x;
                    r2 = fmaf(x, x, r2);
                }
                bool slow = PBC && (amax >= 0x7ffff000u);
                if (CUT) slow = slow || (r2 > cut_lo && r2 < cut_hi);
                flagged |= (slow ? 1u : 0u) << b;
                bool ok = !slow;
                if (SELF) ok = ok && (k != self_k);
                if (CUT) ok = ok && (r2 < rc2s);
                float s, u;
                md_pair<POTK>(r2, ok, P, s, u);
#pragma unroll
                for (int d = 0; d < NDIM; ++d) f[d] = fmaf(dd[d], s, f[d]);
                pe += u;
                cnt += ok ? 1u : 0u;
            }
        } else {
            flagged = (1u << (nj - k0)) - 1u;
        }
        while (flagged) {
            const int b = __ffs(flagged) - 1;
            flagged &= flagged - 1;
            const int k = k0 + b;
            float  dd[NDIM];
            double r2d = 0.0;
#pragma unroll
            for (int d = 0; d < NDIM; ++d) {
                const double x = ap_exact_sep<PBC>(r64, (size_t)d * n + (jt + k), (size_t)d * n + ic, box.L, box.invL);
                r2d = fma(x, x, r2d);
                dd[d] = (float)(x * inv_h);
            }
            const float r2 = (float)(r2d * inv_h * inv_h);
            bool ok = !SELF || (k != self_k);
            if (CUT) ok = ok && (r2d < rc2);
            float s, u;
            md_pair<POTK>(r2, ok, P, s, u);
#pragma unroll
            for (int d = 0; d < NDIM; ++d) f[d] = fmaf(dd[d], s, f[d]);
            pe += u;
            cnt += ok ? 1u : 0u;
        }
    }
}

template <int NDIM, int POTK, bool CUT, bool PBC>
__global__ void __launch_bounds__(AP_TI)
k_allpairs_f32(const uint32_t *__restrict__ q, const double *__restrict__ r64, int n, int jchunk, BoxDev box,
               PotF32 P, float rc2s, double rc2, double inv_h, double *__restrict__ fpart,
               double *__restrict__ pe_part, unsigned long long *__restrict__ cnt_part, StepScalars *__restrict__ sc) {
    __shared__ uint32_t sj[NDIM][AP_TI];
    __shared__ double   red[32];
    __shared__ unsigned redu[32];

    const int  i = blockIdx.x * AP_TI + threadIdx.x;
    const bool active = i < n;
    const int  ic = active ? i : n - 1;
    uint32_t qi[NDIM];
    double   F[NDIM];
#pragma unroll
    for (int d = 0; d < NDIM; ++d) {
        qi[d] = q[(size_t)d * n + ic];
        F[d] = 0.0;
    }
    double   PE = 0.0;
    unsigned cnt = 0;
    const int j0 = blockIdx.y * jchunk;
    const int j1 = min(n, j0 + jchunk);
    const float cut_lo = rc2s * (1.0f - 2e-6f), cut_hi = rc2s * (1.0f + 2e-6f);

    for (int jt = j0; jt < j1; jt += AP_TI) {
        __syncthreads();
        const int jl = jt + threadIdx.x;
        if (jl < j1) {
#pragma unroll
            for (int d = 0; d < NDIM; ++d) sj[d][threadIdx.x] = q[(size_t)d * n + jl];
        }
        __syncthreads();
        const int nj = min(AP_TI, j1 - jt);
        float f[NDIM], pe = 0.0f;
#pragma unroll
        for (int d = 0; d < NDIM; ++d) f[d] = 0.0f;
        if (jt == (int)(blockIdx.x * AP_TI))
            ap_tile_f32<NDIM, POTK, CUT, PBC, true>(sj, nj, ic - jt, jt, ic, n, qi, r64, box, P, rc2s, cut_lo, cut_hi, rc2,
                                                    inv_h, f, pe, cnt);
        else
            ap_tile_f32<NDIM, POTK, CUT, PBC, false>(sj, nj, -1, jt, ic, n, qi, r64, box, P, rc2s, cut_lo, cut_hi, rc2,
                                                     inv_h, f, pe, cnt);
#pragma unroll
        for (int d = 0; d < NDIM; ++d) F[d] += (double)f[d];
        PE += (double)pe;
    }
    if (active) {
#pragma unroll
        for (int d = 0; d < NDIM; ++d) fpart[((size_t)blockIdx.y * NDIM + d) * n + i] = F[d];
    } else {
        PE = 0.0;
        cnt = 0;
    }
    double   bpe = md_block_sum(PE, red);
    unsigned bc = md_block_sum(cnt, redu);
    if (threadIdx.x == 0) {
        const int b = blockIdx.y * gridDim.x + blockIdx.x;
        pe_part[b] = bpe;
        cnt_part[b] = bc;
        atomicAdd(&sc->pairs_pending, (unsigned long long)bc);  // integer: order-independent, stats only
    }
}

// ---------------------------------------------------------------------------------------------
// fp32 kernel, Newton's third law (up to AP3_MAXT tiles): the fp32 twin of k_allpairs_n3_f64.  One CTA per UNORDERED
// pair of 128-particle tiles, every pair evaluated once; the reaction goes to the j particle through the warp's private
// copy of the tile's force array (lanes of a warp visit distinct j in every batch: no atomics, fixed order).  Positions
// are converted to 32-bit fixed point inside the kernel (no separate conversion launch) and staged as one 16-byte record
// per particle; the warp-private reaction arrays are float4 as well, so a pair costs one LDS.128 for the position and
// one LDS.128 + STS.128 for the reaction.  Pairs at the half-box tie or within 2e-6 of the cutoff are redone from the
// float64 positions exactly like in k_allpairs_f32.  fp32 sums cover one tile (128 pairs per particle) and are written
// out as float64 partials: fpart[partner tile][dim][particle], summed in fixed order by the integrator.
// ---------------------------------------------------------------------------------------------
// AP3_JS: every tile pair is shared by AP3_JS CTAs (blockIdx.z), each taking 128 / AP3_JS of the j tile: 528 tile pairs of
// 4 warps would leave a B200 at 14 resident warps per SM; partner slots are (J, z), so the integrator sums AP3_JS * T partials.
#define AP3_JS 2
#define AP3_JN (AP_TI / AP3_JS)   // j particles per CTA
#define AP3_U (AP3_JN / 32)       // independent pairs per batch
template <int NDIM, int POTK, bool CUT, bool PBC>
__global__ void __launch_bounds__(AP_TI)
k_allpairs_n3_f32(const double *__restrict__ r64, int n, BoxDev box, PotF32 P, float rc2s, double rc2, double inv_h,
                  double *__restrict__ fpart, double *__restrict__ pe_part, unsigned long long *__restrict__ cnt_part,
                  StepScalars *__restrict__ sc) {
    __shared__ uint4    sj[AP3_JN];
    __shared__ float4   sfj[AP_TI / 32][AP3_JN];
    __shared__ double   red[32];
    __shared__ unsigned redu[32];
    const int J = blockIdx.x, I = blockIdx.y, h = blockIdx.z;
    const int bidx = (blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
    if (I > J) {   // lower triangle: nothing to do, but the partial-sum slots are read by the finalize kernel
        if (threadIdx.x == 0) { pe_part[bidx] = 0.0; cnt_part[bidx] = 0ull; }
        return;
    }
    const int  t = threadIdx.x, lane = t & 31, w = t >> 5;
    const int  i = I * AP_TI + t;
    const int  j0 = J * AP_TI + h * AP3_JN;            // first j of this CTA
    const bool active = i < n, diag = I == J;
    const int  ic = active ? i : n - 1, nj = min(AP3_JN, n - j0);   // may be <= 0 for the ragged last tile
    const int  tj = t - h * AP3_JN;                    // this thread's own index inside the j range (diagonal tiles: pairs jl > tj)
    uint32_t qi[3] = {0u, 0u, 0u};
#pragma unroll
    for (int d = 0; d < NDIM; ++d) qi[d] = (uint32_t)(unsigned long long)__double2ll_rn(r64[(size_t)d * n + ic] * inv_h);
    if (t < AP3_JN) {
        const int jc = min(j0 + t, n - 1);
        uint32_t qj[3] = {0u, 0u, 0u};
#pragma unroll
        for (int d = 0; d < NDIM; ++d) qj[d] = (uint32_t)(unsigned long long)__double2ll_rn(r64[(size_t)d * n + jc] * inv_h);
        sj[t] = make_uint4(qj[0], qj[1], qj[2], 0u);
#pragma unroll
        for (int ww = 0; ww < AP_TI / 32; ++ww) sfj[ww][t] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    float    f[3] = {0.f, 0.f, 0.f}, pe = 0.f;
    unsigned cnt = 0;
    const float cut_lo = rc2s * (1.0f - 2e-6f), cut_hi = rc2s * (1.0f + 2e-6f);
    float4 *__restrict__ myf = sfj[w];
    // two batches per iteration (b and b + 16): their pair arithmetic is independent (2 * AP3_U pairs in flight); only the
    // reaction updates of the two batches are ordered by __syncwarp (lane l's slot of batch b + 16 is lane l+16's slot of batch b)
    for (int b = 0; b < 16; ++b) {
        float    dd[2 * AP3_U][3], s4[2 * AP3_U];
        int      jls[2 * AP3_U];
        unsigned slow_mask = 0u;
#pragma unroll
        for (int u = 0; u < 2 * AP3_U; ++u) {
            const int   g = (lane + b + (u >= AP3_U ? 16 : 0)) & 31;
            const int   jl = (u % AP3_U) * 32 + g;
            jls[u] = jl;
            const uint4 q = sj[jl];
            const uint32_t qq[3] = {q.x, q.y, q.z};
            float    r2 = 0.0f;
            uint32_t band = 0xffffffffu;
#pragma unroll
            for (int d = 0; d < NDIM; ++d) {
                const uint32_t du = qq[d] - qi[d];
                if (PBC) band = min(band, du - 0x7ffff000u);   // |di| within 2^12 units of the half-box tie <=> du - 0x7ffff000 <= 0x2000
                const float x = (float)(int)du;
                dd[u][d] = x;
                r2 = fmaf(x, x, r2);
            }
            bool ok = active && jl < nj && (!diag || jl > tj);
            bool slow = PBC && band <= 0x2000u;
            if (CUT) slow = slow || (r2 > cut_lo && r2 < cut_hi);
            if (slow) {
                slow_mask |= ok ? (1u << u) : 0u;
                ok = false;
            }
            if (CUT) ok = ok && (r2 < rc2s);
            float s, uu;
            md_pair<POTK>(r2, ok, P, s, uu);
            s4[u] = s;
            pe += uu;
            cnt += ok ? 2u : 0u;
        }
        if (slow_mask) {   // rare: image and in/out-of-cutoff decision from the float64 positions
#pragma unroll
            for (int u = 0; u < 2 * AP3_U; ++u) {   // static indices keep dd / s4 in registers
                if (!((slow_mask >> u) & 1u)) continue;
                double r2d = 0.0;
#pragma unroll
                for (int d = 0; d < NDIM; ++d) {
                    const double x = ap_exact_sep<PBC>(r64, (size_t)d * n + (j0 + jls[u]), (size_t)d * n + ic, box.L, box.invL);
                    r2d = fma(x, x, r2d);
                    dd[u][d] = (float)(x * inv_h);
                }
                const bool ok = !CUT || r2d < rc2;
                float s, uu;
                md_pair<POTK>((float)(r2d * inv_h * inv_h), ok, P, s, uu);
                s4[u] = s;
                pe += uu;
                cnt += ok ? 2u : 0u;
            }
        }
#pragma unroll
        for (int u = 0; u < 2 * AP3_U; ++u) {
            if (u == AP3_U) __syncwarp();
            float4 a = myf[jls[u]];
            const float fx = dd[u][0] * s4[u];
            f[0] += fx; a.x -= fx;           // a_i += dr * s, a_j -= dr * s   (dr = r_j - r_i)
            if (NDIM > 1) { const float fy = dd[u][1] * s4[u]; f[1] += fy; a.y -= fy; }
            if (NDIM > 2) { const float fz = dd[u][2] * s4[u]; f[2] += fz; a.z -= fz; }
            myf[jls[u]] = a;                 // this warp's copy; within a batch its lanes hit distinct j
        }
        __syncwarp();
    }
    __syncthreads();
    const int slot = J * AP3_JS + h;         // partner slot of the i particles; the j particles use slot I * AP3_JS + h' (h' = their own half)
    if (active && !diag) {
#pragma unroll
        for (int d = 0; d < NDIM; ++d) fpart[((size_t)slot * NDIM + d) * n + i] = (double)f[d];
    }
    double fj[3] = {0.0, 0.0, 0.0};
    if (t < AP3_JN) {
#pragma unroll
        for (int ww = 0; ww < AP_TI / 32; ++ww) {
            const float4 a = sfj[ww][t];
            fj[0] += (double)a.x; fj[1] += (double)a.y; fj[2] += (double)a.z;
        }
    }
    if (diag) {
        // diagonal tile pair: slot (I, h) of particle i holds its action over this CTA's j half; the reactions of that half go
        // to the other slot family of the SAME tile, (I, h) as seen from the j side -- fold both into one write per particle:
        // particle i gets f (all threads) and, if it lies in this CTA's j half, the reaction sum as well
        __shared__ double sfold[3][AP3_JN];
        if (t < AP3_JN) {
#pragma unroll
            for (int d = 0; d < NDIM; ++d) sfold[d][t] = fj[d];
        }
        __syncthreads();
        if (active) {
            const bool mine = tj >= 0 && tj < AP3_JN;
#pragma unroll
            for (int d = 0; d < NDIM; ++d) fpart[((size_t)slot * NDIM + d) * n + i] = (double)f[d] + (mine ? sfold[d][tj] : 0.0);
        }
    } else if (t < AP3_JN) {
        // slot (I, h) of the j tile's particles: the reaction for this CTA's half, zero for the other halves (every
        // (slot, particle) entry is written by exactly one CTA per evaluation: the integrator sums all slots)
#pragma unroll
        for (int hh = 0; hh < AP3_JS; ++hh) {
            const int jp = J * AP_TI + hh * AP3_JN + t;
            if (jp < n) {
#pragma unroll
                for (int d = 0; d < NDIM; ++d) fpart[((size_t)(I * AP3_JS + h) * NDIM + d) * n + jp] = hh == h ? fj[d] : 0.0;
            }
        }
    }
    const double   bpe = md_block_sum((double)pe, red);
    const unsigned bc = md_block_sum(cnt, redu);
    if (threadIdx.x == 0) {
        pe_part[bidx] = 2.0 * bpe;   // every unordered pair once: counted for both directions like the split-j kernel
        cnt_part[bidx] = bc;
        atomicAdd(&sc->pairs_pending, (unsigned long long)bc);
    }
}

// positions -> 32-bit fixed point (flat over ndim*n elements)
__global__ void k_to_fixed(const double *__restrict__ r, uint32_t *__restrict__ q, size_t E, double scale) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e < E) q[e] = (uint32_t)(unsigned long long)__double2ll_rn(r[e] * scale);
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int md_allpairs_setup(mdsea_ctx *c) {
    const int n = (int)c->n;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->p.device);
    c->ap_ti = AP_TI;
    c->ap_blocks_x = md_div_up(n, AP_TI);
    // j split: enough blocks for ~8 resident 128-thread CTAs per SM, a whole multiple of the SM count,
    // but never chunks shorter than one tile.
    int want = sms * 8;
    int nsplit = want / c->ap_blocks_x;
    if (const char *e = getenv("MDSEA_AP_NSPLIT")) nsplit = atoi(e);  // tuning hook
    if (nsplit < 1) nsplit = 1;
    int max_split = md_div_up(n, AP_TI);
    if (nsplit > max_split) nsplit = max_split;
    c->ap_nsplit = nsplit;
    c->ap_jchunk = md_div_up(md_div_up(n, nsplit), AP_TI) * AP_TI;  // whole tiles: the diagonal tile is block-uniform
    c->ap_nsplit = md_div_up(n, c->ap_jchunk);
    // up to AP3_MAXT tiles: the Newton's-third-law kernels (one CTA per unordered tile pair, T partner slots), both precisions
    c->ap_n3 = c->ap_blocks_x <= AP3_MAXT && !(getenv("MDSEA_AP_N3") && !strcmp(getenv("MDSEA_AP_N3"), "0"));
    if (c->ap_n3) c->ap_nsplit = c->ap_blocks_x * (c->p.precision == MDSEA_FP32 ? AP3_JS : 1);   // partner slots
    c->n_pe_part = c->ap_blocks_x * c->ap_nsplit;
    MD_CUDA(c, cudaMalloc(&c->d_fpart, sizeof(double) * c->E * c->ap_nsplit));
    MD_CUDA(c, cudaMalloc(&c->d_pe_part, sizeof(double) * c->n_pe_part));
    MD_CUDA(c, cudaMalloc(&c->d_cnt_part, sizeof(unsigned long long) * c->n_pe_part));
    MD_CUDA(c, cudaMemset(c->d_pe_part, 0, sizeof(double) * c->n_pe_part));
    MD_CUDA(c, cudaMemset(c->d_cnt_part, 0, sizeof(unsigned long long) * c->n_pe_part));
    if (c->p.precision == MDSEA_FP32) MD_CUDA(c, cudaMalloc(&c->d_rhi, sizeof(uint32_t) * c->E));
    return MDSEA_OK;
}

template <int NDIM, int POTK, bool CUT, bool PBC>
static void launch_f64(mdsea_ctx *c, dim3 grid, const PotF64 &P, double rc2) {
    k_allpairs_f64<NDIM, POTK, CUT, PBC><<<grid, AP_TI, 0, c->stream>>>(
        c->d_r, (int)c->n, c->ap_jchunk, c->box, P, rc2, c->d_fpart, c->d_pe_part, c->d_cnt_part, c->d_sc);
}
template <int NDIM, int POTK, bool CUT, bool PBC>
static void launch_n3_f64(mdsea_ctx *c, dim3 grid, const PotF64 &P, double rc2) {
    k_allpairs_n3_f64<NDIM, POTK, CUT, PBC><<<grid, AP_TI, 0, c->stream>>>(c->d_r, (int)c->n, c->box, P, rc2, c->d_fpart, c->d_pe_part,
                                                                        c->d_cnt_part, c->d_sc);
}
template <int NDIM, int POTK, bool CUT, bool PBC>
static void launch_n3_f32(mdsea_ctx *c, dim3 grid, const PotF32 &P, float rc2s, double rc2, double inv_h) {
    k_allpairs_n3_f32<NDIM, POTK, CUT, PBC><<<grid, AP_TI, 0, c->stream>>>(c->d_r, (int)c->n, c->box, P, rc2s, rc2, inv_h, c->d_fpart,
                                                                        c->d_pe_part, c->d_cnt_part, c->d_sc);
}
template <int NDIM, int POTK, bool CUT, bool PBC>
static void launch_f32(mdsea_ctx *c, dim3 grid, const PotF32 &P, float rc2s, double rc2, double inv_h) {
    k_allpairs_f32<NDIM, POTK, CUT, PBC><<<grid, AP_TI, 0, c->stream>>>(
        (const uint32_t *)c->d_rhi, c->d_r, (int)c->n, c->ap_jchunk, c->box, P, rc2s, rc2, inv_h, c->d_fpart,
        c->d_pe_part, c->d_cnt_part, c->d_sc);
}

#define AP_DISPATCH(LAUNCH, ...)                                                              \
    do {                                                                                      \
        const bool cut = c->p.r_cutoff > 0, pbc = c->box.pbc != 0;                            \
        const int  pk = c->pot.fast;                                                          \
        const int  key = (c->ndim - 1) * 8 + (pk == POTF_LJ ? 0 : 4) + (cut ? 2 : 0) + (pbc ? 1 : 0); \
        switch (key) {                                                                        \
            case 0: LAUNCH<1, POTF_LJ, false, false>(__VA_ARGS__); break;                     \
            case 1: LAUNCH<1, POTF_LJ, false, true>(__VA_ARGS__); break;                      \
            case 2: LAUNCH<1, POTF_LJ, true, false>(__VA_ARGS__); break;                      \
            case 3: LAUNCH<1, POTF_LJ, true, true>(__VA_ARGS__); break;                       \
            case 4: LAUNCH<1, POTF_MIE, false, false>(__VA_ARGS__); break;                    \
            case 5: LAUNCH<1, POTF_MIE, false, true>(__VA_ARGS__); break;                     \
            case 6: LAUNCH<1, POTF_MIE, true, false>(__VA_ARGS__); break;                     \
            case 7: LAUNCH<1, POTF_MIE, true, true>(__VA_ARGS__); break;                      \
            case 8: LAUNCH<2, POTF_LJ, false, false>(__VA_ARGS__); break;                     \
            case 9: LAUNCH<2, POTF_LJ, false, true>(__VA_ARGS__); break;                      \
            case 10: LAUNCH<2, POTF_LJ, true, false>(__VA_ARGS__); break;                     \
            case 11: LAUNCH<2, POTF_LJ, true, true>(__VA_ARGS__); break;                      \
            case 12: LAUNCH<2, POTF_MIE, false, false>(__VA_ARGS__); break;                   \
            case 13: LAUNCH<2, POTF_MIE, false, true>(__VA_ARGS__); break;                    \
            case 14: LAUNCH<2, POTF_MIE, true, false>(__VA_ARGS__); break;                    \
            case 15: LAUNCH<2, POTF_MIE, true, true>(__VA_ARGS__); break;                     \
            case 16: LAUNCH<3, POTF_LJ, false, false>(__VA_ARGS__); break;                    \
            case 17: LAUNCH<3, POTF_LJ, false, true>(__VA_ARGS__); break;                     \
            case 18: LAUNCH<3, POTF_LJ, true, false>(__VA_ARGS__); break;                     \
            case 19: LAUNCH<3, POTF_LJ, true, true>(__VA_ARGS__); break;                      \
            case 20: LAUNCH<3, POTF_MIE, false, false>(__VA_ARGS__); break;                   \
            case 21: LAUNCH<3, POTF_MIE, false, true>(__VA_ARGS__); break;                    \
            case 22: LAUNCH<3, POTF_MIE, true, false>(__VA_ARGS__); break;                    \
            default: LAUNCH<3, POTF_MIE, true, true>(__VA_ARGS__); break;                     \
        }                                                                                     \
    } while (0)

int md_allpairs_force(mdsea_ctx *c) {
    const dim3 grid(c->ap_blocks_x, c->ap_nsplit);
    const PotDev &pd = c->pot;
    const double rc2 = c->p.r_cutoff > 0 ? c->p.r_cutoff * c->p.r_cutoff : 0.0;
    md_prof_begin(c);
    if (c->p.precision == MDSEA_FP64) {
        const PotF64 P = md_pot_f64(pd);
        if (c->ap_n3) AP_DISPATCH(launch_n3_f64, c, dim3(c->ap_blocks_x, c->ap_blocks_x), P, rc2);
        else AP_DISPATCH(launch_f64, c, grid, P, rc2);
        c->stats.kernel_launches += 1;
    } else {
        const double units = c->box.pbc ? 4294967296.0 : 2147483648.0;
        const double h = c->box.L / units, inv_h = units / c->box.L;
        const PotF32 P = md_pot_f32(pd, inv_h);
        (void)h;
        if (c->ap_n3) {
            AP_DISPATCH(launch_n3_f32, c, dim3(c->ap_blocks_x, c->ap_blocks_x, AP3_JS), P, (float)(rc2 * inv_h * inv_h), rc2, inv_h);
            c->stats.kernel_launches += 1;
        } else {
            k_to_fixed<<<md_div_up((long long)c->E, 256), 256, 0, c->stream>>>(c->d_r, (uint32_t *)c->d_rhi, c->E, inv_h);
            AP_DISPATCH(launch_f32, c, grid, P, (float)(rc2 * inv_h * inv_h), rc2, inv_h);
            c->stats.kernel_launches += 2;
        }
    }
    md_prof_end(c, &c->stats.ms_force);
    MD_CUDA(c, cudaGetLastError());
    c->stats.force_evals += 1;
    c->stats.pair_evals_executed += (long long)c->n * (long long)c->n;
    return MDSEA_OK;
}
