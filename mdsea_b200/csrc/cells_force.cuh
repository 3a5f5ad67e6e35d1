// K4: neighbour-list force / energy kernels.  Included by cells.cu only.
#pragma once

// ---------------------------------------------------------------------------------------------------
// K4: neighbour-list forces.  One thread per particle, no atomics on forces; the list is read coalesced,
// neighbour positions are 16 B (fp32 mode) or 32 B (fp64 mode) gathers that mostly hit L1/L2 because
// slots are cell-sorted.  Borderline decisions (image tie, |r2 - rc2| within the fp32 guard band) are
// masked out of the hot loop and redone from the float64 positions, like in the all-pairs kernels.
// ---------------------------------------------------------------------------------------------------
// POTK = POTF_LJ: sqrt-free Lennard-Jones; POTF_MIE: the generic bounded-Mie member (pair.cuh).
template <int POTK, bool CS>   // CS: streaming (evict-first) loads of the list rows, see k_nbr_force_f32
__global__ void __launch_bounds__(NB_THREADS)
k_nbr_force_f64(const double4 *__restrict__ pos, long long i_begin, long long n_own, long long cap, const int *__restrict__ nbr,
                const int *__restrict__ nbr_cnt, int max_nbr, double L, double halfL, double rc2, PotF64 P,
                double *__restrict__ acc, double *__restrict__ pe_part, StepScalars *__restrict__ sc, int guard_tag) {
    __shared__ double   red[32];
    __shared__ unsigned redu[32];
    if (guard_tag && sc->rebuild_flag == guard_tag) return;  // speculative launch: the list is stale, the host will redo this step
    const long long i = i_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double   fx = 0.0, fy = 0.0, fz = 0.0, pe = 0.0;
    unsigned cnt = 0;
    if (i < n_own) {
        const double4 pi = pos[i];
        const int nn = nbr_cnt[i];
        const int *__restrict__ row0 = nbr + md_nbr_row_base(i, max_nbr);
#pragma unroll 4
        for (int k = 0; k < nn; ++k) {
            const int j = CS ? __ldcs(row0 + k * 32) : row0[k * 32];
            const double4 pj = pos[j];
            const double dx = md_minimg_list(pj.x - pi.x, L, halfL);
            const double dy = md_minimg_list(pj.y - pi.y, L, halfL);
            const double dz = md_minimg_list(pj.z - pi.z, L, halfL);
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            const bool ok = r2 < rc2;
            double s, u;
            md_pair<POTK>(r2, ok, P, s, u);
            fx = fma(dx, s, fx);
            fy = fma(dy, s, fy);
            fz = fma(dz, s, fz);
            pe += u;
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
        atomicAdd(&sc->pairs_pending, (unsigned long long)bc);
    }
}

// CS: the list rows are read once per launch -- streaming loads (evict-first) keep them from displacing the packed
// positions, which every thread gathers ~56 times, out of L1 / L2.
template <bool TEX, int POTK, bool CS>
__global__ void __launch_bounds__(NB_THREADS)
k_nbr_force_f32(const uint4 *__restrict__ pos, cudaTextureObject_t tex, const double *__restrict__ r64, long long i_begin, long long n_own, long long cap,
                const int *__restrict__ nbr, const int *__restrict__ nbr_cnt, int max_nbr, double L, double halfL,
                double rc2, float rc2s, double inv_h, PotF32 P, double *__restrict__ acc, double *__restrict__ pe_part,
                StepScalars *__restrict__ sc, int guard_tag) {
    __shared__ double   red[32];
    __shared__ unsigned redu[32];
    if (guard_tag && sc->rebuild_flag == guard_tag) return;  // speculative launch: the list is stale, the host will redo this step
    const long long i = i_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double   Fx = 0.0, Fy = 0.0, Fz = 0.0, PE = 0.0;
    unsigned cnt = 0;
    if (i < n_own) {
        const uint4 qi = pos[i];
        const int   nn = nbr_cnt[i];
        const int *__restrict__ row0 = nbr + md_nbr_row_base(i, max_nbr);
        const float cut_lo = rc2s * (1.0f - 2e-6f), cut_hi = rc2s * (1.0f + 2e-6f);
        // Groups of NBG entries: all index loads, then all position gathers, then the arithmetic -- NBG independent
        // gathers in flight per thread (the kernel is bound by the latency of these L1/L2 gathers, not by issue).
        // Rows are padded by the build kernel to a multiple of NBG with the particle's own slot (masked by j != i).
        for (int k0 = 0; k0 < nn; k0 += NBG) {
            int   j[NBG];
            uint4 qj[NBG];
#pragma unroll
            for (int u = 0; u < NBG; ++u) j[u] = CS ? __ldcs(row0 + (k0 + u) * 32) : row0[(k0 + u) * 32];
#pragma unroll
            for (int u = 0; u < NBG; ++u) qj[u] = TEX ? tex1Dfetch<uint4>(tex, j[u]) : pos[j[u]];
            float    fx = 0.f, fy = 0.f, fz = 0.f, pe = 0.f;
            unsigned flagged = 0u;
#pragma unroll
            for (int u = 0; u < NBG; ++u) {
                const float dx = (float)(int)(qj[u].x - qi.x), dy = (float)(int)(qj[u].y - qi.y), dz = (float)(int)(qj[u].z - qi.z);
                const float r2 = fmaf(dz, dz, fmaf(dy, dy, dx * dx));
                // an in-cutoff pair can only sit at the image tie if rc ~ L/2; rc + skin <= L/3 here, so only the
                // cutoff band needs the exact redo
                const bool inside = r2 < cut_lo, maybe = r2 < cut_hi;
                flagged |= ((maybe && !inside) ? 1u : 0u) << u;
                const bool  ok = inside && (j[u] != (int)i);
                float s, w;
                md_pair<POTK>(r2, ok, P, s, w);
                fx = fmaf(dx, s, fx);
                fy = fmaf(dy, s, fy);
                fz = fmaf(dz, s, fz);
                pe += w;
                cnt += ok ? 1u : 0u;
            }
            while (flagged) {  // rare: decide and evaluate from the float64 positions
                const int u = __ffs(flagged) - 1;
                flagged &= flagged - 1;
                const int jj = row0[(k0 + u) * 32];
                if (jj == (int)i) continue;
                const double ex = md_minimg_list(r64[jj] - r64[i], L, halfL);
                const double ey = md_minimg_list(r64[cap + jj] - r64[cap + i], L, halfL);
                const double ez = md_minimg_list(r64[2 * cap + jj] - r64[2 * cap + i], L, halfL);
                const double r2d = fma(ez, ez, fma(ey, ey, ex * ex));
                if (r2d < rc2) {
                    const float dx = (float)(ex * inv_h), dy = (float)(ey * inv_h), dz = (float)(ez * inv_h);
                    const float r2 = (float)(r2d * inv_h * inv_h);
                    float s, w;
                    md_pair<POTK>(r2, true, P, s, w);
                    fx = fmaf(dx, s, fx);
                    fy = fmaf(dy, s, fy);
                    fz = fmaf(dz, s, fz);
                    pe += w;
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
        atomicAdd(&sc->pairs_pending, (unsigned long long)bc);
    }
}
