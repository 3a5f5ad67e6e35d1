// Initial conditions generated on the device: the step immediately BEFORE the hot path.  Opt-in:
// the reference draws from the global NumPy Mersenne Twister in a fixed call order (mdsea/gen.py:98-123), which a
// device generator cannot reproduce bit for bit, so parity runs keep the host generators of mdsea_b200/gen.py.  What
// this mode keeps from the reference:
//   * the simple-cubic lattice of gen.py:144-156, BIT FOR BIT: coords[d][g] = (L / ppr) * (0.5 + (g / ppr^d) % ppr);
//   * Maxwell-Boltzmann speeds by inverse-CDF sampling on the reference's own Monte-Carlo grid (gen.py:19-37,104-105):
//     the CDF table is computed by the host mirror (scipy erf, 50 000 entries) and only SAMPLED here, with the linear
//     interpolation of scipy's interp1d;
//   * the direction factors of gen.py:83-92,106-121, including the whole-array np.prod that multiplies the last two
//     components in 3-D by the product of sin(theta_0) over ALL particles (vy, vz ~ 0).  `isotropic` replaces them by
//     uniformly distributed directions.
// Random numbers: Philox4x32-10 keyed by the seed, counter = (global particle id, draw), so the state does not depend
// on how the particles are spread over ranks or blocks.
#include "ctx.h"

// ---------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): counter-based, 4 x 32 random bits per call
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 md_philox(uint4 ctr, uint2 key) {
#pragma unroll
    for (int round = 0; round < 10; ++round) {
        const unsigned hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
        ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
        key.x += 0x9E3779B9u;
        key.y += 0xBB67AE85u;
    }
    return ctr;
}
// 53-bit uniform in [0, 1) from two words, the construction of numpy's random_sample
__device__ __forceinline__ double md_u01(unsigned a, unsigned b) {
    return ((double)(a >> 5) * 67108864.0 + (double)(b >> 6)) * (1.0 / 9007199254740992.0);
}
// draws of particle g: .x = speed, .y = theta_0, .z = theta_1 (only the ones gen.py would draw are used)
__device__ __forceinline__ void md_draws(long long g, unsigned long long seed, double &u_speed, double &u_a, double &u_b) {
    const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
    const uint4 w0 = md_philox(make_uint4((unsigned)g, (unsigned)((unsigned long long)g >> 32), 0u, 0u), key);
    const uint4 w1 = md_philox(make_uint4((unsigned)g, (unsigned)((unsigned long long)g >> 32), 1u, 0u), key);
    u_speed = md_u01(w0.x, w0.y);
    u_a = md_u01(w0.z, w0.w);
    u_b = md_u01(w1.x, w1.y);
}

#define MD_PI 3.141592653589793

// ---------------------------------------------------------------------------------------------------
// lattice (gen.py:144-156): particle g = i0 + k goes to column k of r (row stride `stride`)
// ---------------------------------------------------------------------------------------------------
__global__ void k_gen_lattice(double *__restrict__ r, long long stride, long long i0, long long m, int ndim, long long ppr, double span) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    long long g = i0 + k;
    for (int d = 0; d < ndim; ++d) {
        r[d * stride + k] = __dmul_rn(span, 0.5 + (double)(g % ppr));   // span * ticks, one rounding like NumPy
        g /= ppr;
    }
}

// ---------------------------------------------------------------------------------------------------
// product of sin(theta_0) over ALL particles (3-D reference semantics): block partials, then one block
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gen_sinprod(long long n, unsigned long long seed, double *__restrict__ part) {
    __shared__ double red[256];
    double p = 1.0;
    for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n; g += (long long)gridDim.x * blockDim.x) {
        double us, ua, ub;
        md_draws(g, seed, us, ua, ub);
        p *= sin(MD_PI * ua);
    }
    red[threadIdx.x] = p;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) red[threadIdx.x] *= red[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}
__global__ void k_gen_sinprod_top(const double *__restrict__ part, int nb, double *__restrict__ out) {
    double p = 1.0;
    for (int k = 0; k < nb; ++k) p *= part[k];
    *out = p;
}

// ---------------------------------------------------------------------------------------------------
// velocities: inverse-CDF speed times direction factors; also the per-block sums of |v| (for get_dt, helpers.py:89-100)
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_gen_mb(double *__restrict__ v, long long stride, long long i0, long long m, int ndim, const double *__restrict__ cdf, int ncdf,
         double ds, unsigned long long seed, int isotropic, const double *__restrict__ sinprod, double *__restrict__ speed_part) {
    __shared__ double red[32];
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double norm = 0.0;
    if (k < m) {
        const long long g = i0 + k;
        double us, ua, ub;
        md_draws(g, seed, us, ua, ub);
        // interp1d(cdf, speeds)(u): largest j with cdf[j] <= u, linear inside [cdf[j], cdf[j + 1])
        int lo = 0, hi = ncdf - 1;
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (cdf[mid] <= us) lo = mid; else hi = mid;
        }
        const double c0 = cdf[lo], c1 = cdf[hi];
        const double f = (c1 > c0) ? fmin(1.0, (us - c0) / (c1 - c0)) : 0.0;   // (u beyond the table: its last speed)
        const double s = ds * ((double)lo + f);
        double c[3] = {0.0, 0.0, 0.0};
        if (ndim == 1) {
            c[0] = s * cos(2.0 * MD_PI * ua);                       // thetas = [U(0, 2 pi)]
        } else if (ndim == 2) {
            const double th = 2.0 * MD_PI * ua;                     // thetas = [U(0, 2 pi)]; empty sine products are 1
            c[0] = s * sin(th);
            c[1] = s * cos(th);
        } else if (isotropic) {
            const double ct = 2.0 * ua - 1.0, st = sqrt(fmax(0.0, 1.0 - ct * ct)), ph = 2.0 * MD_PI * ub;
            c[0] = s * st * cos(ph);
            c[1] = s * st * sin(ph);
            c[2] = s * ct;
        } else {
            const double t0 = MD_PI * ua, t1 = 2.0 * MD_PI * ub, P = *sinprod;   // gen.py:83-92 with the whole-array np.prod
            c[0] = s * cos(t0);
            c[1] = s * sin(t1) * P;
            c[2] = s * cos(t1) * P;
        }
        for (int d = 0; d < ndim; ++d) {
            v[d * stride + k] = c[d];
            norm += c[d] * c[d];
        }
        norm = sqrt(norm);
    }
    const double b = md_block_sum(norm, red);
    if (threadIdx.x == 0) speed_part[blockIdx.x] = b;
}
__global__ void k_gen_zero(double *__restrict__ v, long long stride, long long m, int ndim) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= m) return;
    for (int d = 0; d < ndim; ++d) v[d * stride + k] = 0.0;
}

int md_cutoff_adopt_generated(mdsea_ctx *c, long long lo, long long m);   // cells.cu

// mean_speed_out: mean over ALL particles of |v_i| (every rank generates the norms of the whole system: N hashes)
int md_generate_state(mdsea_ctx *c, const double *cdf, long long ncdf, double ds, unsigned long long seed, int isotropic,
                      double *mean_speed_out) {
    const long long N = c->n_global;
    const int nd = c->ndim;
    const long long ppr = llround(pow((double)N, 1.0 / nd));
    long long chk = 1;
    for (int d = 0; d < nd; ++d) chk *= ppr;
    if (chk != N) {   // gen.py:147-148
        md_set_error(c, "generate_state: num_particles must be a perfect power of ndim (simple-cubic lattice)");
        return MDSEA_ERR_BAD_ARG;
    }
    if (cdf && ncdf < 2) { md_set_error(c, "generate_state: the CDF table needs at least two entries"); return MDSEA_ERR_BAD_ARG; }
    const int W = c->p.world;
    long long lo = 0, m = N, stride = N;
    if (c->cutoff_path) {
        stride = md_cutoff_capacity(c);
        if (W > 1) {
            if (!c->comm) { md_set_error(c, "generate_state: world > 1 needs comm_init first"); return MDSEA_ERR_STATE; }
            lo = (N * c->p.rank) / W;
            m = (N * (c->p.rank + 1)) / W - lo;   // the same slice set_state would upload; the first rebuild routes it
        }
        if (m > stride) { md_set_error(c, "generate_state: particle capacity of this rank exceeded"); return MDSEA_ERR_CAPACITY; }
    }
    const int g = md_div_up(m > 0 ? m : 1, 256);
    k_gen_lattice<<<g, 256, 0, c->stream>>>(c->d_r, stride, lo, m, nd, ppr, c->p.boxlen / (double)ppr);
    c->stats.kernel_launches += 1;
    double mean_speed = 0.0;
    if (!cdf) {   // temp == 0: gen.py:100-103
        k_gen_zero<<<g, 256, 0, c->stream>>>(c->d_v, stride, m, nd);
        c->stats.kernel_launches += 1;
    } else {
        struct Scratch {   // freed on every exit path
            double *p[3] = {nullptr, nullptr, nullptr};
            ~Scratch() { for (double *q : p) cudaFree(q); }
        } scratch;
        double *&d_cdf = scratch.p[0], *&d_part = scratch.p[1], *&d_tmp = scratch.p[2];
        const int gN = md_div_up(N, 256), nbp = 1024;
        MD_CUDA(c, cudaMalloc(&d_cdf, sizeof(double) * (size_t)ncdf));
        MD_CUDA(c, cudaMalloc(&d_part, sizeof(double) * ((size_t)gN + nbp + 1)));
        MD_CUDA(c, cudaMemcpyAsync(d_cdf, cdf, sizeof(double) * (size_t)ncdf, cudaMemcpyHostToDevice, c->stream));
        double *d_prod = d_part + gN + nbp;
        if (nd == 3 && !isotropic) {
            k_gen_sinprod<<<nbp, 256, 0, c->stream>>>(N, seed, d_part + gN);
            k_gen_sinprod_top<<<1, 1, 0, c->stream>>>(d_part + gN, nbp, d_prod);
            c->stats.kernel_launches += 2;
        }
        // this rank's slice into the state arrays
        k_gen_mb<<<g, 256, 0, c->stream>>>(c->d_v, stride, lo, m, nd, d_cdf, (int)ncdf, ds, seed, isotropic, d_prod, d_part);
        c->stats.kernel_launches += 1;
        std::vector<double> h((size_t)gN);
        if (W > 1) {   // the mean speed of the WHOLE system: regenerate all norms into scratch (the state slice is untouched)
            MD_CUDA(c, cudaMalloc(&d_tmp, sizeof(double) * (size_t)nd * (size_t)N));
            k_gen_mb<<<gN, 256, 0, c->stream>>>(d_tmp, N, 0, N, nd, d_cdf, (int)ncdf, ds, seed, isotropic, d_prod, d_part);
            c->stats.kernel_launches += 1;
            MD_CUDA(c, cudaMemcpyAsync(h.data(), d_part, sizeof(double) * (size_t)gN, cudaMemcpyDeviceToHost, c->stream));
            MD_CUDA(c, cudaStreamSynchronize(c->stream));
        } else {
            MD_CUDA(c, cudaMemcpyAsync(h.data(), d_part, sizeof(double) * (size_t)g, cudaMemcpyDeviceToHost, c->stream));
            MD_CUDA(c, cudaStreamSynchronize(c->stream));
            h.resize((size_t)g);
        }
        double sum = 0.0;
        for (double x : h) sum += x;   // fixed order: the same bits on every rank
        mean_speed = sum / (double)N;
    }
    MD_CUDA(c, cudaGetLastError());
    if (c->cutoff_path) MD_TRY(md_cutoff_adopt_generated(c, lo, m));
    MD_CUDA(c, cudaStreamSynchronize(c->stream));
    if (mean_speed_out) *mean_speed_out = mean_speed;
    c->have_state = true;
    c->acc_valid = false;
    return MDSEA_OK;
}
