// Per-step reductions and bookkeeping, run by the LAST block of the finish kernels (k_kick_finish of the all-pairs
// path, k_cut_kick_finish of the cut-off path): no single-block launch on the critical path of a step.  Deterministic: every thread sums a fixed strided
// subset of the partials, then a fixed shuffle tree -- the result does not depend on which block runs it.
#pragma once
#include "common.cuh"

struct FinalizeArgs {
    const double *pe_part; int n_pe;
    const double *ke_part; int n_ke;
    double pe_factor;   // lam_eps / 2 / N
    double ke_factor;   // 0.5 m / N
    double kb;
    double *pe_rows, *ke_rows, *temp_rows;
    int row;
    const double *quench_next;
    long long step_index;
};

__device__ __forceinline__ void md_finalize_body(const FinalizeArgs &A, StepScalars *__restrict__ sc, double *red /* >= 32 */) {
    double pe = 0.0, ke = 0.0;
    {   // eight independent accumulators per thread: the loads are latency-, not bandwidth-bound (L2 round trips)
        const int st = blockDim.x;
        double p[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) p[u] = 0.0;
        int k = threadIdx.x;
        for (; k + 7 * st < A.n_pe; k += 8 * st) {
#pragma unroll
            for (int u = 0; u < 8; ++u) p[u] += __ldcg(A.pe_part + k + u * st);
        }
        for (; k < A.n_pe; k += st) p[0] += __ldcg(A.pe_part + k);
        pe = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
#pragma unroll
        for (int u = 0; u < 8; ++u) p[u] = 0.0;
        k = threadIdx.x;
        for (; k + 7 * st < A.n_ke; k += 8 * st) {
#pragma unroll
            for (int u = 0; u < 8; ++u) p[u] += __ldcg(A.ke_part + k + u * st);
        }
        for (; k < A.n_ke; k += st) p[0] += __ldcg(A.ke_part + k);
        ke = ((p[0] + p[1]) + (p[2] + p[3])) + ((p[4] + p[5]) + (p[6] + p[7]));
    }
    pe = md_block_sum(pe, red);
    ke = md_block_sum(ke, red);
    if (threadIdx.x == 0) {
        const double mean_pe = pe * A.pe_factor;  // sum over unique pairs / N, simulator.py:249
        const double mean_ke = ke * A.ke_factor;  // 0.5 m mean(v.v), simulator.py:244
        A.pe_rows[A.row] = mean_pe;
        A.ke_rows[A.row] = mean_ke;
        A.temp_rows[A.row] = sc->temp;            // self.temp as written by update_files, simulator.py:147
        if (!(isfinite(mean_pe) && isfinite(mean_ke)) && sc->first_nonfinite_step < 0) sc->first_nonfinite_step = A.step_index;
        sc->ke_prev = mean_ke;
        sc->ke_valid = 1;
        sc->steps_done += 1;
        sc->pairs_directed += sc->pairs_pending;
        sc->pairs_pending = 0ull;
        // quench bookkeeping for the NEXT step (simulator.py:185-195, 233-239)
        double scale = 1.0, temp = sc->temp;
        if (A.quench_next) {
            for (int t = 0; t < 2; ++t) {
                const double target = A.quench_next[t];
                if (target == target) {  // not NaN
                    temp = (2.0 / 3.0) * mean_ke / A.kb;
                    if (temp != 0.0) scale *= sqrt(target / temp);
                }
            }
        }
        sc->temp = temp;
        sc->scale = scale;
    }
}
