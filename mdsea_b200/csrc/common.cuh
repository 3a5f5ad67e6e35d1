// Shared declarations of the mdsea_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/mdsea_b200.h"

#define MD_WARP 32

// ---------------------------------------------------------------------------------------------
// error plumbing: every host-side helper returns mdsea_status and records a message in the ctx
// ---------------------------------------------------------------------------------------------
std::string &md_global_error();

#define MD_CUDA(ctx, expr)                                                                       \
    do {                                                                                         \
        cudaError_t _e = (expr);                                                                 \
        if (_e != cudaSuccess) {                                                                 \
            char _b[512];                                                                        \
            snprintf(_b, sizeof(_b), "CUDA error %s at %s:%d: %s", cudaGetErrorName(_e), __FILE__, \
                     __LINE__, cudaGetErrorString(_e));                                          \
            md_set_error(ctx, _b);                                                               \
            return MDSEA_ERR_CUDA;                                                               \
        }                                                                                        \
    } while (0)

#define MD_TRY(expr)                     \
    do {                                 \
        int _s = (expr);                 \
        if (_s != MDSEA_OK) return _s;   \
    } while (0)

// ---------------------------------------------------------------------------------------------
// device-side constants of one context
// ---------------------------------------------------------------------------------------------
enum PotFast { POTF_IDEAL = 0, POTF_LJ = 1, POTF_MIE = 2 };

struct PotDev {
    int    fast;      // PotFast
    double lam_eps;   // lambda * epsilon  (lambda = (m/(m-n)) (m/n)^(n/(m-n)), potentials.py:17,45)
    double sigma;     // sigma
    double sigma2;    // sigma^2
    double a2;        // a^2
    double m, n;      // exponents
    int    mi, ni;    // integer exponents when integral and in [0, 64], else -1
    double inv_mass;  // 1 / MASS  (simulator.py:301)
};

struct BoxDev {
    double L, invL, halfL;
    float  Lh, Ll, invLf;  // L = Lh + Ll split for the fp32 path
    int    pbc;
};

// scalars that live on the device across steps (temperature bookkeeping of _quench/update_temp,
// simulator.py:192-195,233-239) plus per-step reduction outputs.
struct StepScalars {
    double temp;       // self.temp
    double ke_prev;    // self.mean_ke of the previous step
    int    ke_valid;   // mean_ke is not None
    double scale;      // velocity scale of the NEXT kick_finish (product of the pending quenches)
    long long first_nonfinite_step;  // -1 while energies are finite
    unsigned long long pairs_directed; // running count of directed in-cutoff pair evaluations (2 x unique pairs)
    unsigned long long pairs_pending;  // contributions of the current step's force launches; committed by the step's finalize (finalize.cuh)
                                       // (a speculative launch that is redone after a rebuild must not count twice)
    unsigned long long steps_done;
    int    rebuild_flag;             // cutoff path: TAG of the drift after which the max displacement exceeded skin/2 (0: none;
                                     // every drift gets its own tag, cells.cu CellState::drift_tag)
    int    overflow_flag;            // capacity overflow reported by a kernel
    unsigned finish_blocks_done;     // cut-off path: blocks of k_cut_kick_finish that have published their KE partial
};

// ---------------------------------------------------------------------------------------------
// small device helpers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double md_warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float md_warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned md_warp_sum(unsigned v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// block-wide sum; result valid in thread 0.  Deterministic (fixed shuffle tree + fixed warp order).
template <typename T>
__device__ __forceinline__ T md_block_sum(T v, T *smem /* >= 32 entries */) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = md_warp_sum(v);
    __syncthreads();
    if (lane == 0) smem[warp] = v;
    __syncthreads();
    T r = T(0);
    if (warp == 0) {
        const int nw = (blockDim.x + 31) >> 5;
        r = (lane < nw) ? smem[lane] : T(0);
        r = md_warp_sum(r);
    }
    return r;
}

// fp64 reciprocal: MUFU.RCP64H seed (>= 20 bits) + one cubic Newton step (-> > 53 bits).
// Relative error <~ 2 ulp; callers never rely on correct rounding (parity bound is 1e-10).
__device__ __forceinline__ double md_rcp(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double e = fma(-x, y, 1.0);
    double p = fma(e, e, e);
    return fma(y, p, y);
}

// fp32 reciprocal: one MUFU.RCP (<= 1 ulp), no denormal/special-case slow path
__device__ __forceinline__ float md_rcpf(float x) {
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// round-half-even to integer for |x| < 2^51 (RN add of 1.5*2^52), the semantics of np.rint
__device__ __forceinline__ double md_rint_magic(double x) {
    const double M = 6755399441055744.0;
    return __dsub_rn(__dadd_rn(x, M), M);
}
__device__ __forceinline__ float md_rint_magic(float x) {
    const float M = 12582912.0f;
    return __fsub_rn(__fadd_rn(x, M), M);
}

// reference minimum image on one component (simulator.py:271-272): d - rint(d / L) * L with a TRUE
// division.  Fast path multiplies by 1/L; when the quotient is within 1e-9 of a half-integer the
// image choice could differ from the IEEE quotient's, so that (rare) case redoes it exactly.
__device__ __forceinline__ double md_minimg(double d, const double L, const double invL) {
    double q = d * invL;
    double n = md_rint_magic(q);
    double f = q - n;
    if (__builtin_expect(fabs(f) > 0.5 - 1e-9 || fabs(q) > 1e15, 0)) {
        n = rint(__ddiv_rn(d, L));
        return __dsub_rn(d, __dmul_rn(n, L));
    }
    return fma(-n, L, d);
}

static inline int md_div_up(long long a, long long b) { return (int)((a + b - 1) / b); }
