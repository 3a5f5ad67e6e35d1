// Generic exclusive scan over ints (bucket starts of the binning, ordered halo lists).  Included by cells.cu only.
#pragma once

// ---------------------------------------------------------------------------------------------------
// generic exclusive scan over ints (three kernels; used for the ordered halo lists)
// ---------------------------------------------------------------------------------------------------
#define SCAN_THREADS 256
#define SCAN_ITEMS 8
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_block(const int *__restrict__ in, int *__restrict__ out, long long n, int *__restrict__ block_sum) {
    __shared__ int wsum[SCAN_THREADS / 32];
    const long long base = ((long long)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
    int v[SCAN_ITEMS], t = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        v[k] = base + k < n ? in[base + k] : 0;
        t += v[k];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int wbase = 0;
    for (int w = 0; w < warp; ++w) wbase += wsum[w];
    int run = wbase + incl - t;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) {
        if (base + k < n) out[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == SCAN_THREADS - 1) block_sum[blockIdx.x] = run;
}
__global__ void __launch_bounds__(1024) k_scan_top(int *__restrict__ a, int m) {
    __shared__ int part[1024];
    const int per = (m + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(m, lo + per);
    int s = 0;
    for (int k = lo; k < hi; ++k) s += a[k];
    part[threadIdx.x] = s;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
        const int v = threadIdx.x >= o ? part[threadIdx.x - o] : 0;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    int run = part[threadIdx.x] - s;
    for (int k = lo; k < hi; ++k) {
        const int v = a[k];
        a[k] = run;
        run += v;
    }
}
__global__ void __launch_bounds__(SCAN_THREADS)
k_scan_add(int *__restrict__ out, long long n, const int *__restrict__ block_sum) {
    const long long base = ((long long)blockIdx.x * SCAN_THREADS + threadIdx.x) * SCAN_ITEMS;
    const int add = block_sum[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k)
        if (base + k < n) out[base + k] += add;
}
