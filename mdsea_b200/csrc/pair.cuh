// Pair math shared by the all-pairs (allpairs.cu) and the neighbour-list (cells.cu) force kernels: the bounded-Mie
// family of potentials.py:15-19 (V) and :43-48 (dV/dr), with the Lennard-Jones member (m, n, a) = (12, 6, 0) as a
// sqrt-free fast path.
#pragma once
#include "common.cuh"

// ---------------------------------------------------------------------------------------------
// pair kernels (sqrt-free where the exponents allow it)
//   s = (dV/dr) / (r m)   so that   a_i += dr * s
//   u = V / (lambda eps)
// ---------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T md_ipow(T b, int e) {
    T r = T(1);
    while (e > 0) {
        if (e & 1) r *= b;
        b *= b;
        e >>= 1;
    }
    return r;
}

struct PotF64 {
    double c;      // LJ: 6 lambda eps / m ; Mie: lambda eps / m
    double sigma, sigma2, a2, m, n;
    int    mi, ni;
};
struct PotF32 {
    float c;       // same as PotF64::c but divided by the fixed-point unit h (forces come out in real units)
    float sigma_s, sigma2_s, a2_s, m, n;  // lengths in fixed-point units
    int   mi, ni;
};

template <int POTK>
__device__ __forceinline__ void md_pair(double r2, bool ok, const PotF64 &P, double &s, double &u) {
    if (POTK == POTF_LJ) {
        double inv = ok ? md_rcp(r2) : 0.0;
        double t2 = P.sigma2 * inv;
        double t6 = t2 * t2 * t2;
        double w = fma(-2.0, t6, 1.0);
        s = (P.c * inv) * (t6 * w);
        u = fma(t6, t6, -t6);
    } else {
        double aprs = P.a2 + r2;
        double inv = ok ? md_rcp(aprs) : 0.0;
        double tm, tn;
        if (P.mi >= 0 && P.ni >= 0 && !((P.mi | P.ni) & 1)) {
            double t2 = P.sigma2 * inv;
            tm = md_ipow(t2, P.mi >> 1);
            tn = md_ipow(t2, P.ni >> 1);
        } else {
            double t = P.sigma * sqrt(inv);
            tm = (P.mi >= 0) ? md_ipow(t, P.mi) : pow(t, P.m);
            tn = (P.ni >= 0) ? md_ipow(t, P.ni) : pow(t, P.n);
        }
        tm = ok ? tm : 0.0;
        tn = ok ? tn : 0.0;
        s = P.c * inv * (P.n * tn - P.m * tm);
        u = tm - tn;
    }
}

template <int POTK>
__device__ __forceinline__ void md_pair(float r2, bool ok, const PotF32 &P, float &s, float &u) {
    if (POTK == POTF_LJ) {
        float inv = ok ? md_rcpf(r2) : 0.0f;
        float t2 = P.sigma2_s * inv;
        float t6 = t2 * t2 * t2;
        float w = fmaf(-2.0f, t6, 1.0f);
        s = (P.c * inv) * (t6 * w);
        u = fmaf(t6, t6, -t6);
    } else {
        float aprs = P.a2_s + r2;
        float inv = ok ? md_rcpf(aprs) : 0.0f;
        float tm, tn;
        if (P.mi >= 0 && P.ni >= 0 && !((P.mi | P.ni) & 1)) {
            float t2 = P.sigma2_s * inv;
            tm = md_ipow(t2, P.mi >> 1);
            tn = md_ipow(t2, P.ni >> 1);
        } else {
            float t = P.sigma_s * sqrtf(inv);
            tm = (P.mi >= 0) ? md_ipow(t, P.mi) : powf(t, P.m);
            tn = (P.ni >= 0) ? md_ipow(t, P.ni) : powf(t, P.n);
        }
        tm = ok ? tm : 0.0f;
        tn = ok ? tn : 0.0f;
        s = P.c * inv * (P.n * tn - P.m * tm);
        u = tm - tn;
    }
}

// host side: kernel constants from the context's potential.  inv_h = fixed-point units per length (fp32 kernels).
static inline PotF64 md_pot_f64(const PotDev &pd) {
    PotF64 P;
    P.c = (pd.fast == POTF_LJ ? 6.0 : 1.0) * pd.lam_eps * pd.inv_mass;
    P.sigma = pd.sigma; P.sigma2 = pd.sigma2; P.a2 = pd.a2; P.m = pd.m; P.n = pd.n; P.mi = pd.mi; P.ni = pd.ni;
    return P;
}
static inline PotF32 md_pot_f32(const PotDev &pd, double inv_h) {
    PotF32 P;
    P.c = (float)((pd.fast == POTF_LJ ? 6.0 : 1.0) * pd.lam_eps * pd.inv_mass * inv_h);
    P.sigma_s = (float)(pd.sigma * inv_h);
    P.sigma2_s = (float)(pd.sigma2 * inv_h * inv_h);
    P.a2_s = (float)(pd.a2 * inv_h * inv_h);
    P.m = (float)pd.m; P.n = (float)pd.n; P.mi = pd.mi; P.ni = pd.ni;
    return P;
}
