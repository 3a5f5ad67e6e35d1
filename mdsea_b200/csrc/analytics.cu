// Analyser-side reductions on the device: the per-particle speeds of every stored step and
// their global mean / standard deviation -- what mdsea's Analyser derives from `v_coords` right after loading a
// simfile (mdsea/analytics.py:53-56: speeds = norm(v_vecs, axis=-1); maxspeed = mean + 3 std).  The dataset is
// streamed through the GPU in chunks of steps (H2D of chunk c+1 overlaps the kernel and the D2H of chunk c), so the
// multi-GB (STEPS, NDIM, N) array makes one trip over PCIe instead of several NumPy passes over host memory.
// Stand-alone entry point: it needs no solver context.
#include <math.h>

#include <algorithm>

#include "common.cuh"

// speeds[s][i] = sqrt((vx*vx + vy*vy) + vz*vz), each operation rounded on its own (no FMA): bit-identical to
// np.sqrt(np.add.reduce(np.square(v), axis=-1)) of quicker.norm (mdsea/quicker.py:44-46)
__global__ void __launch_bounds__(256)
k_speeds(const double *__restrict__ v, long long steps, int ndim, long long n, double *__restrict__ speeds,
         double *__restrict__ part_sum, double *__restrict__ part_sq, double *__restrict__ part_max) {
    __shared__ double red[32];
    const long long total = steps * n;
    double s1 = 0.0, s2 = 0.0, mx = 0.0;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long long)gridDim.x * blockDim.x) {
        const long long s = t / n, i = t - s * n;
        const double *row = v + s * ndim * n + i;
        double acc = __dmul_rn(row[0], row[0]);
        for (int d = 1; d < ndim; ++d) acc = __dadd_rn(acc, __dmul_rn(row[d * n], row[d * n]));
        const double sp = sqrt(acc);
        if (speeds) speeds[t] = sp;
        s1 += sp;
        s2 += sp * sp;
        mx = fmax(mx, sp);
    }
    const double b1 = md_block_sum(s1, red);
    const double b2 = md_block_sum(s2, red);
    // block max through the same shuffle tree
    for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); ++w) m = fmax(m, red[w]);
        part_sum[blockIdx.x] = b1;
        part_sq[blockIdx.x] = b2;
        part_max[blockIdx.x] = m;
    }
}

#define AN_CUDA(expr)                                                                                         \
    do {                                                                                                      \
        cudaError_t _e = (expr);                                                                              \
        if (_e != cudaSuccess) {                                                                              \
            char _b[256];                                                                                     \
            snprintf(_b, sizeof(_b), "analyse_speeds: CUDA error %s at %s:%d", cudaGetErrorName(_e), __FILE__, __LINE__); \
            md_global_error() = _b;                                                                           \
            st = MDSEA_ERR_CUDA;                                                                              \
            goto done;                                                                                        \
        }                                                                                                     \
    } while (0)

extern "C" int mdsea_b200_analyse_speeds(int32_t device, const double *v_coords, int64_t steps, int32_t ndim, int64_t n,
                                         double *speeds, double *mean, double *stddev, double *max_speed) {
    if (!v_coords || steps < 1 || n < 1 || ndim < 1 || ndim > 3) {
        md_global_error() = "analyse_speeds: bad argument";
        return MDSEA_ERR_BAD_ARG;
    }
    int st = MDSEA_OK;
    const size_t row_in = sizeof(double) * (size_t)ndim * (size_t)n, row_out = sizeof(double) * (size_t)n;
    const long long chunk = std::max<long long>(1, std::min<long long>(steps, (long long)((256u << 20) / row_in)));
    const int blocks = 148 * 8;
    double *d_in[2] = {nullptr, nullptr}, *d_out[2] = {nullptr, nullptr}, *d_part = nullptr;
    std::vector<double> h_part;
    cudaStream_t stream[2] = {nullptr, nullptr};
    double sum = 0.0, sq = 0.0, mx = 0.0;
    long long nchunks = 0;
    AN_CUDA(cudaSetDevice(device));
    for (int b = 0; b < 2; ++b) {
        AN_CUDA(cudaMalloc(&d_in[b], row_in * chunk));
        if (speeds) AN_CUDA(cudaMalloc(&d_out[b], row_out * chunk));
        AN_CUDA(cudaStreamCreateWithFlags(&stream[b], cudaStreamNonBlocking));
    }
    nchunks = (steps + chunk - 1) / chunk;
    AN_CUDA(cudaMalloc(&d_part, sizeof(double) * 3 * blocks * nchunks));
    for (long long c = 0; c < nchunks; ++c) {
        const int b = (int)(c & 1);
        const long long s0 = c * chunk, ns = std::min<long long>(chunk, steps - s0);
        // the two streams alternate: the copies of one chunk overlap the kernel of the other
        AN_CUDA(cudaMemcpyAsync(d_in[b], v_coords + (size_t)s0 * ndim * n, row_in * ns, cudaMemcpyHostToDevice, stream[b]));
        double *part = d_part + (size_t)c * 3 * blocks;
        k_speeds<<<blocks, 256, 0, stream[b]>>>(d_in[b], ns, ndim, n, speeds ? d_out[b] : nullptr, part, part + blocks, part + 2 * blocks);
        if (speeds) AN_CUDA(cudaMemcpyAsync(speeds + (size_t)s0 * n, d_out[b], row_out * ns, cudaMemcpyDeviceToHost, stream[b]));
    }
    AN_CUDA(cudaStreamSynchronize(stream[0]));
    AN_CUDA(cudaStreamSynchronize(stream[1]));
    AN_CUDA(cudaGetLastError());
    h_part.resize((size_t)3 * blocks * nchunks);
    AN_CUDA(cudaMemcpy(h_part.data(), d_part, sizeof(double) * h_part.size(), cudaMemcpyDeviceToHost));
    for (long long c = 0; c < nchunks; ++c)
        for (int k = 0; k < blocks; ++k) {   // fixed order: deterministic
            sum += h_part[(size_t)c * 3 * blocks + k];
            sq += h_part[(size_t)c * 3 * blocks + blocks + k];
            mx = fmax(mx, h_part[(size_t)c * 3 * blocks + 2 * blocks + k]);
        }
    {
        const double M = (double)steps * (double)n, mu = sum / M;
        if (mean) *mean = mu;
        if (stddev) *stddev = sqrt(fmax(0.0, sq / M - mu * mu));   // population std, like np.std
        if (max_speed) *max_speed = mx;
    }
done:
    for (int b = 0; b < 2; ++b) {
        cudaFree(d_in[b]);
        cudaFree(d_out[b]);
        if (stream[b]) cudaStreamDestroy(stream[b]);
    }
    cudaFree(d_part);
    return st;
}
