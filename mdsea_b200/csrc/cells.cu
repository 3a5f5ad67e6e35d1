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
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include "ctx.h"
#include "comm.h"
#include "finalize.cuh"
#include "pair.cuh"

#define SORT_THREADS 256
#define SORT_ITEMS 8
#define SORT_CHUNK (SORT_THREADS * SORT_ITEMS)
#define NB_THREADS 128       // thread-per-particle list build
#define NBG 8                // list entries per gather group of the force kernels

// Spatial domain decomposition on the canonical cell grid: rank (bx, by, bz) of a px*py*pz process grid owns the
// cells with bnd[d][b_d] <= c_d < bnd[d][b_d + 1]; its halo is the one-cell layer around that block (periodic).
// Ownership and halo membership are functions of the canonical cell id alone, so every rank can evaluate them
// for any particle and all ranks agree without talking to each other.
struct Decomp {
    int world, rank, nc;
    int p[3];
    int bnd[3][MD_MAXW + 1];
};
__host__ __device__ inline int dc_block(const Decomp &D, int d, int c) {
    int b = 0;
    while (b + 1 < D.p[d] && c >= D.bnd[d][b + 1]) ++b;
    return b;
}
__host__ __device__ inline int dc_owner(const Decomp &D, int cx, int cy, int cz) {
    return (dc_block(D, 2, cz) * D.p[1] + dc_block(D, 1, cy)) * D.p[0] + dc_block(D, 0, cx);
}
// is the cell inside rank q's block extended by the one-cell halo?  (true for q's own cells too)
__host__ __device__ inline bool dc_in_ext(const Decomp &D, int q, int cx, int cy, int cz) {
    const int b[3] = {q % D.p[0], (q / D.p[0]) % D.p[1], q / (D.p[0] * D.p[1])};
    const int cc[3] = {cx, cy, cz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int lo = D.bnd[d][b[d]], len = D.bnd[d][b[d] + 1] - lo;
        if (len + 2 >= D.nc) continue;  // block + halo covers the whole period
        const int dd = (cc[d] - (lo - 1) + D.nc) % D.nc;
        if (dd >= len + 2) return false;
    }
    return true;
}

// does the 27-cell stencil of this OWNED cell reach into the halo?  (cells on a face of the block in a decomposed
// dimension)  Interior cells only have owned neighbours, so their forces do not wait for the halo refresh.
__host__ __device__ inline bool dc_is_boundary(const Decomp &D, int cx, int cy, int cz) {
    const int b[3] = {D.rank % D.p[0], (D.rank / D.p[0]) % D.p[1], D.rank / (D.p[0] * D.p[1])};
    const int cc[3] = {cx, cy, cz};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (D.p[d] == 1) continue;
        if (cc[d] == D.bnd[d][b[d]] || cc[d] == D.bnd[d][b[d] + 1] - 1) return true;
    }
    return false;
}

struct PMsg {  // one particle in a rebuild-time message
    double    r[3], v[3];
    long long id;
};

struct CellState {
    Decomp    D;
    long long n_tot = 0;                  // owned + ghost slots in use
    long long n_int = 0;                  // owned slots [0, n_int) are interior (no ghost neighbours), [n_int, n) boundary
    cudaEvent_t ev_drift = nullptr;       // main stream -> comm stream
    // multi-rank staging
    PMsg  *sendpool = nullptr, *recvpool = nullptr;
    long long sendpool_cap = 0;
    int   *d_cnt = nullptr, *d_cnt_all = nullptr, *d_cursor = nullptr;  // [MAXW], [MAXW*MAXW], [MAXW]
    int   *h_cnt_all = nullptr;           // pinned + mapped [MAXW*MAXW + 8]: small read-backs of the rebuild (k_peek)
    int   *d_cnt_all_mapped = nullptr;    // device alias of h_cnt_all
    int   *halo_flag = nullptr, *halo_scan = nullptr, *halo_idx = nullptr;  // [world][cap]
    int   *scan_tmp = nullptr;
    long long halo_nsend[MD_MAXW], halo_nrecv[MD_MAXW], halo_base[MD_MAXW];
    double *halo_send = nullptr, *halo_recv = nullptr;   // packed positions, 3 doubles per entry
    // peer-memory halo (NCCL transport on one node): every rank exposes p2p_buf = [2 parities][p2p_stride doubles] +
    // [MD_MAXW arrival counters]; the neighbours' kernels store positions and counters straight into it over NVLink
    bool      p2p = false;
    double   *p2p_buf = nullptr;
    long long p2p_stride = 0;
    double   *peer_buf[MD_MAXW];
    unsigned long long halo_seq = 0;      // exchanges done so far (identical on all ranks)
    unsigned long long asm_seq = 0;       // rebuilds done so far (identical on all ranks)
    long long seg_cap = 0;                // rebuild inbox: one segment of seg_cap messages per source rank
    size_t    ctrl_off = 0, inbox_off = 0; // byte offsets of the control block / the inbox inside p2p_buf
    int      *d_roff = nullptr;           // [MD_MAXW + 2] staging offsets of the received segments, n_tot, n_self
    std::vector<long long> row_n_own;     // owned count of every buffered row
    int   *cell_end = nullptr;
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
    cudaTextureObject_t tex_pos = 0;      // fp32 positions as a linear texture (default gather path of the fp32 force kernel)
    unsigned long long *key = nullptr, *key_alt = nullptr;
    int    *val = nullptr, *val_alt = nullptr;
    int    *hist = nullptr, *hist_tot = nullptr;
    int    *bkt = nullptr, *arrival = nullptr, *bcount = nullptr, *bstart = nullptr;   // counting-sort binning
    int    *tmp_id = nullptr, *tmp_src = nullptr, *tmp_bkt = nullptr;
    long long nbucket = 0;
    bool    radix_binning = false;
    uint4  *qfix = nullptr;               // 32-bit fixed-point positions, refreshed at every rebuild (list prefilter)
    int     sort_blocks = 0;
    int    *cell = nullptr;               // cell id per slot (sorted)
    int    *cell_start = nullptr;         // ncell + 1
    int    *nbr = nullptr, *nbr_cnt = nullptr;
    double *scratch = nullptr;            // [3][N] for id-ordered downloads
    int    *h_flag = nullptr;             // pinned + mapped: the skin trigger is published with a store over PCIe
    int    *d_flag_mapped = nullptr;      // device alias of h_flag
    cudaEvent_t ev_flag = nullptr;
    int     id_bits = 0, cell_bits = 0;
    int     finish_blocks = 1480;         // grid of the finish kernel: two full waves of its resident blocks per SM (5 -> 10 per SM)
    int     finish_blocks_fused = 888;    // same for the variant that also starts the next step (3 resident -> 6 per SM)
    double *pe_fold = nullptr;            // [finish_blocks] PE partials of the force kernel folded by the finish blocks
    double *ke_glob = nullptr;            // world > 1: one double, the all-reduced mean_ke of a step (quench bookkeeping)
    // Skin-trigger tags.  StepScalars::rebuild_flag holds the TAG of the drift whose trigger fired (0: none); every drift
    // gets the next tag (same sequence on all ranks).  Guards compare against the tag of the drift they depend on, so
    // the finish kernel of step k can be guarded by drift k's trigger while its fused drift k+1 raises its own.
    int     drift_tag = 0;
    int     published_tag = 0;            // single GPU: the fused finish kernel already published this tag's trigger (h_flag[2 + (tag & 1)])
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

// 32-bit fixed point (unit L / 2^32): wraps like the periodic box
__device__ __forceinline__ uint32_t md_to_fixed(double x, double inv_h) {
    return (uint32_t)(unsigned long long)__double2ll_rn(x * inv_h);
}

// ---------------------------------------------------------------------------------------------------
// K2a: keys.  key = ghost << 63 | boundary << 62 | cell id << 32 | particle id ; value = current (staging) slot.
// Slot order: owned interior, owned boundary, ghosts; within each class the canonical (cell id, particle id).
// ---------------------------------------------------------------------------------------------------
__global__ void k_cell_keys(const double *__restrict__ r, const int *__restrict__ id, long long n, long long cap, double L,
                            double cl, int nc, Decomp D, unsigned long long *__restrict__ key, int *__restrict__ val,
                            int *__restrict__ n_owned) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = md_cell_coord(r[i], L, cl, nc);
    const int cy = md_cell_coord(r[cap + i], L, cl, nc);
    const int cz = md_cell_coord(r[2 * cap + i], L, cl, nc);
    const unsigned long long c = (unsigned long long)((cz * nc + cy) * nc + cx);
    unsigned long long ghost = 0ull, bnd = 0ull;
    if (D.world > 1) {
        ghost = dc_owner(D, cx, cy, cz) != D.rank ? 1ull : 0ull;
        if (!ghost) {
            atomicAdd(n_owned, 1);
            bnd = dc_is_boundary(D, cx, cy, cz) ? 1ull : 0ull;
            if (!bnd) atomicAdd(n_owned + 1, 1);
        }
    }
    key[i] = (ghost << 63) | (bnd << 62) | (c << 32) | (unsigned)id[i];
    val[i] = (int)i;
}

// ---------------------------------------------------------------------------------------------------
// K2 (default): binning as ONE counting pass over the cells -- a radix sort whose single digit is the bucket
// (class, cell) -- followed by a rank-by-id inside each bucket.  Buckets: class 0 = owned interior, 1 = owned
// boundary, 2 = ghost (single GPU: class 0 only); bucket = class * ncell + cell id.  The arrival order inside a
// bucket (atomicAdd) is not deterministic, so the final slot of a particle is bucket start + the number of bucket
// members with a smaller particle id: the result is the canonical stable order by (class, cell id, particle id),
// identical to what the LSD radix passes below produce (kept as MDSEA_BINNING=radix for A/B).
// ---------------------------------------------------------------------------------------------------
__global__ void k_bin_count(const double *__restrict__ r, long long n, long long cap, double L, double cl, int nc, int ncell,
                            Decomp D, int *__restrict__ bkt, int *__restrict__ arrival, int *__restrict__ bcount) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cx = md_cell_coord(r[i], L, cl, nc);
    const int cy = md_cell_coord(r[cap + i], L, cl, nc);
    const int cz = md_cell_coord(r[2 * cap + i], L, cl, nc);
    int b = (cz * nc + cy) * nc + cx;
    if (D.world > 1) {
        const int cls = dc_owner(D, cx, cy, cz) != D.rank ? 2 : (dc_is_boundary(D, cx, cy, cz) ? 1 : 0);
        b += cls * ncell;
    }
    bkt[i] = b;
    arrival[i] = atomicAdd(&bcount[b], 1);
}

__global__ void k_bin_scatter(long long n, const int *__restrict__ bkt, const int *__restrict__ arrival,
                              const int *__restrict__ bstart, const int *__restrict__ id, int *__restrict__ tmp_id,
                              int *__restrict__ tmp_src, int *__restrict__ tmp_bkt) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int b = bkt[i];
    const int u = bstart[b] + arrival[i];
    tmp_id[u] = id[i];
    tmp_src[u] = (int)i;
    tmp_bkt[u] = b;
}

// rank inside the bucket -> final slot; gather the state into it (+ build positions, ids, cell ids, cell bounds
// and the packed position copies: everything the rest of the rebuild reads)
__global__ void __launch_bounds__(256)
k_bin_place(long long n, long long cap, int ncell, const int *__restrict__ bstart, const int *__restrict__ tmp_id,
            const int *__restrict__ tmp_src, const int *__restrict__ tmp_bkt, const double *__restrict__ r_in,
            const double *__restrict__ v_in, double *__restrict__ r_out, double *__restrict__ v_out, double *__restrict__ rb,
            int *__restrict__ id_out, int *__restrict__ cell, int *__restrict__ cell_start, int *__restrict__ cell_end, int fp32,
            double inv_h, void *__restrict__ pos4, uint4 *__restrict__ qfix) {
    const long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const int b = tmp_bkt[u], lo = bstart[b], hi = bstart[b + 1];
    const int my = tmp_id[u];
    int rank = 0;
    for (int t = lo; t < hi; ++t) rank += tmp_id[t] < my ? 1 : 0;
    const int s = lo + rank, o = tmp_src[u];
    const int cc = b % ncell;
    id_out[s] = my;
    cell[s] = cc;
    if (rank == 0) {
        cell_start[cc] = lo;
        cell_end[cc] = hi;
    }
    double x[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        x[d] = r_in[d * cap + o];
        r_out[d * cap + s] = x[d];
        rb[d * cap + s] = x[d];
        v_out[d * cap + s] = v_in[d * cap + o];
    }
    const uint4 q = make_uint4(md_to_fixed(x[0], inv_h), md_to_fixed(x[1], inv_h), md_to_fixed(x[2], inv_h), 0u);
    qfix[s] = q;
    if (fp32) ((uint4 *)pos4)[s] = q;
    else ((double4 *)pos4)[s] = make_double4(x[0], x[1], x[2], 0.0);
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
    cell[s] = (int)((k >> 32) & 0x3fffffffull);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const double x = r_in[d * cap + o];
        r_out[d * cap + s] = x;
        rb[d * cap + s] = x;
        v_out[d * cap + s] = v_in[d * cap + o];
    }
}

// [cell_start[c], cell_end[c]) = slots of cell c; both arrays are zeroed first (empty cells: 0, 0).  With ghosts
// sorted after the owned particles the slot order is no longer globally monotone in the cell id, hence two arrays.
__global__ void k_cell_bounds(const int *__restrict__ cell, long long n, int *__restrict__ cell_start,
                              int *__restrict__ cell_end) {
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n) return;
    const int c = cell[s];
    if (s == 0 || cell[s - 1] != c) cell_start[c] = (int)s;
    if (s == n - 1 || cell[s + 1] != c) cell_end[c] = (int)s + 1;
}

// packed position copies for the force gathers
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
// force kernel reads one 128-byte row per k (coalesced); the build kernels' stores of one warp all land in the
// same ~15 KB tile, so partially written sectors merge in L2 instead of being evicted half empty.
// (A quad-sliced layout with 8 lanes per particle was measured in r01: 1.6x fewer L1 wavefronts per gather, but
// 1.7x more instructions per pair -- 208 us against 136 us for this one; see profiles/README.md.)
__host__ __device__ __forceinline__ size_t md_nbr_row_base(long long i, int max_nbr) {
    return (size_t)(i >> 5) * (size_t)max_nbr * 32u + (size_t)(i & 31);
}
__host__ __device__ __forceinline__ unsigned md_nbr_off(unsigned k) { return k * 32u; }

// ---------------------------------------------------------------------------------------------------
// K3: Verlet neighbour list.
//
// Two kernels produce the SAME canonical list (j != i with canonical float64 r^2 < (rc+skin)^2):
//   k_nbr_build_warp  (default)  one warp per tile of 32 consecutive slots.  The warp walks the UNION of the 27-cell
//                     stencils of its particles (they share an x-row of cells, so the union is 9 short slot ranges),
//                     stages 32 candidates at a time in shared memory as floats relative to a common origin (the
//                     fixed-point subtraction is exact and wraps to the minimum image for free) and every lane tests
//                     the broadcast candidates against its own particle: no divergence, no per-lane gathers.  In-range
//                     candidates are collected in a 32-bit mask per lane and appended after each chunk.
//   k_nbr_build       one thread per particle over its own 27 cells (MDSEA_NBR_BUILD=thread; kept as the A/B baseline).
// Both decide in float with a guard band around the list radius; a particle that saw a candidate inside the band is
// redone with the canonical float64 test for every candidate, so the list is the canonical one bit for bit.
// ---------------------------------------------------------------------------------------------------
template <bool EXACT>
__device__ __forceinline__ void nb_scan_range(int j0, int j1, const uint4 qi, long long i, const double *__restrict__ r,
                                              const uint4 *__restrict__ q, long long cap, double L, double halfL, double rl2,
                                              float rl2s_lo, float rl2s_hi, int max_nbr, int *__restrict__ row0, int &cnt,
                                              bool &band) {
    if (EXACT) {
        const double xi = r[i], yi = r[cap + i], zi = r[2 * cap + i];
        for (int j = j0; j < j1; ++j) {
            const double ddx = md_minimg_exact(__dsub_rn(r[j], xi), L, halfL);
            const double ddy = md_minimg_exact(__dsub_rn(r[cap + j], yi), L, halfL);
            const double ddz = md_minimg_exact(__dsub_rn(r[2 * cap + j], zi), L, halfL);
            if (md_r2_canonical(ddx, ddy, ddz) < rl2 && j != (int)i) {
                if (cnt < max_nbr) row0[md_nbr_off((unsigned)cnt)] = j;
                ++cnt;
            }
        }
    } else {
        // clamped write cursor: one predicated store per candidate, no bounds branch (an over-full row keeps
        // overwriting its last slot; cnt still counts, the caller reports the overflow)
        const unsigned k_max = (unsigned)(max_nbr - 1);
#pragma unroll 4
        for (int j = j0; j < j1; ++j) {
            const uint4 qj = q[j];
            const float fx = (float)(int)(qj.x - qi.x), fy = (float)(int)(qj.y - qi.y), fz = (float)(int)(qj.z - qi.z);
            const float r2f = fmaf(fz, fz, fmaf(fy, fy, fx * fx));
            const bool in = r2f < rl2s_lo;
            band = band || (!in && r2f < rl2s_hi);
            const bool st = in && (j != (int)i);
            const unsigned off = md_nbr_off(min((unsigned)cnt, k_max));
            if (st) row0[off] = j;
            cnt += st ? 1 : 0;
        }
    }
}

template <bool EXACT>
__device__ __forceinline__ int nb_particle(long long i, const uint4 qi, int c, const double *__restrict__ r,
                                           const uint4 *__restrict__ q, long long cap, const int *__restrict__ cell_start,
                                           const int *__restrict__ cell_end, int nc, double L, double halfL, double rl2,
                                           float rl2s_lo, float rl2s_hi, int max_nbr, int *__restrict__ row0, bool &band) {
    const int cx = c % nc, cy = (c / nc) % nc, cz = c / (nc * nc);
    const int xm = (cx + nc - 1) % nc, xp = (cx + 1) % nc;
    int cnt = 0;
    for (int dz = -1; dz <= 1; ++dz) {
        const int z = (cz + dz + nc) % nc;
        for (int dy = -1; dy <= 1; ++dy) {
            const int y = (cy + dy + nc) % nc;
            const int row = (z * nc + y) * nc;
            // the three x-adjacent cells: merge slot ranges that happen to be contiguous (the usual case away from
            // the periodic wrap and from the owned/ghost boundary)
            int cs = cell_start[row + xm], ce = cell_end[row + xm];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
                const int cc = row + (t == 0 ? cx : xp);
                const int s1 = cell_start[cc], e1 = cell_end[cc];
                if (ce == cs) {            // nothing pending
                    cs = s1; ce = e1;
                } else if (e1 != s1) {
                    if (ce == s1) ce = e1; // contiguous: extend
                    else {
                        nb_scan_range<EXACT>(cs, ce, qi, i, r, q, cap, L, halfL, rl2, rl2s_lo, rl2s_hi, max_nbr, row0, cnt, band);
                        cs = s1; ce = e1;
                    }
                }
            }
            nb_scan_range<EXACT>(cs, ce, qi, i, r, q, cap, L, halfL, rl2, rl2s_lo, rl2s_hi, max_nbr, row0, cnt, band);
        }
    }
    return cnt;
}

__global__ void __launch_bounds__(NB_THREADS)
k_nbr_build(const double *__restrict__ r, const uint4 *__restrict__ q, long long n_own, long long cap,
            const int *__restrict__ cell, const int *__restrict__ cell_start, const int *__restrict__ cell_end, int nc,
            double L, double halfL, double rl2, float rl2s_lo, float rl2s_hi, int max_nbr, int *__restrict__ nbr,
            int *__restrict__ nbr_cnt, StepScalars *__restrict__ sc) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_own) return;
    const uint4 qi = q[i];
    const int c = cell[i];
    int *row0 = nbr + md_nbr_row_base(i, max_nbr);
    bool band = false;
    int cnt = nb_particle<false>(i, qi, c, r, q, cap, cell_start, cell_end, nc, L, halfL, rl2, rl2s_lo, rl2s_hi, max_nbr, row0, band);
    if (__builtin_expect(band, 0))
        cnt = nb_particle<true>(i, qi, c, r, q, cap, cell_start, cell_end, nc, L, halfL, rl2, rl2s_lo, rl2s_hi, max_nbr, row0, band);
    if (cnt > max_nbr) {
        sc->overflow_flag = 1;
        cnt = max_nbr;
    }
    nbr_cnt[i] = cnt;
    for (int k = cnt; k < ((cnt + NBG - 1) / NBG) * NBG; ++k) row0[k * 32] = (int)i;  // padding, masked by j != i
}

// ---- warp-cooperative build ------------------------------------------------------------------------
#define NBW_WARPS 4      // warps (= tiles of 32 slots) per block
#define NBW_RUN   6      // at most this many x-consecutive cells of i-particles per pass (bounds the float error)
// NBW_G (template parameter): lane groups per warp with their own candidate streams (32 / NBW_G consecutive particles
// each).  More groups = shorter streams (fewer tests) but more staging and more shared-memory wavefronts per broadcast
// LDS.128 (4 groups: 4 wavefronts, one address per warp: 2).  MDSEA_NBW_GROUPS selects 4 / 2 / 1 for A/B.
// Offsets from the common origin are plain floats (no wrap), so they only reproduce the minimum image while the
// candidate window stays inside half a period: (run + 2) cells < nc in x, 3 cells < nc in y and z.  The host
// therefore passes run_max = min(NBW_RUN, nc - 3) and uses the thread-per-particle kernel when nc == 3.

// test one staged chunk (<= 32 candidates, padded with far-away dummies) against this lane's particle.
// Expanded form r^2 - rl^2 = |c|^2 + (|p|^2 - rl^2) - 2 p.c: the candidate carries w = |c|^2 and the lane k = |p|^2 - rl^2
// (both rounded ONCE from float64), so a test costs 3 FFMA + 1 FADD on the FMA pipe instead of 3 FADD + 3 FFMA -- the
// kernel is bound by that pipe.  Offsets stay below ~4 cells in x and 1.5 cells in y / z, which bounds the rounding
// error of the chain (largest term last) by ~5e-6 rl^2; the guard band (2e-5 rl^2) covers it.
template <bool EXPANDED>
__device__ __forceinline__ unsigned nbw_test_chunk(const float4 *__restrict__ cand, int ncand, float a, float b, float cc, float k,
                                                   float neg_bw, float &bandmin) {
    // EXPANDED: (a, b, cc) = -2 p, k = |p|^2 - rl^2; direct form: (a, b, cc) = p, k = -rl^2
    unsigned m = 0u;
    for (int t0 = 0; t0 < ncand; t0 += 8) {
        unsigned m8 = 0u;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const float4 cj = cand[t0 + u];   // broadcast LDS.128
            float d;                          // r^2 - rl^2
            if (EXPANDED) {
                d = fmaf(a, cj.x, fmaf(b, cj.y, fmaf(cc, cj.z, k))) + cj.w;
            } else {
                const float dx = cj.x - a, dy = cj.y - b, dz = cj.z - cc;
                d = fmaf(dz, dz, fmaf(dy, dy, fmaf(dx, dx, k)));
            }
            if (d < neg_bw) m8 |= 1u << u;
            bandmin = fminf(bandmin, fabsf(d));
        }
        m |= m8 << t0;
    }
    return m;
}

// staged candidate: float offsets from the pass origin (+ their squared norm, one rounding, for the expanded form)
template <bool EXPANDED>
__device__ __forceinline__ float4 nbw_stage(const uint4 qj, uint32_t ox, uint32_t oy, uint32_t oz) {
    const float x = (float)(int)(qj.x - ox), y = (float)(int)(qj.y - oy), z = (float)(int)(qj.z - oz);
    if (!EXPANDED) return make_float4(x, y, z, 0.0f);
    const double w = fma((double)x, (double)x, fma((double)y, (double)y, (double)z * (double)z));
    return make_float4(x, y, z, (float)w);
}
template <bool EXPANDED>
__device__ __forceinline__ float4 nbw_dummy() {   // a candidate infinitely far away
    return EXPANDED ? make_float4(0.0f, 0.0f, 0.0f, 1.0e30f) : make_float4(1.0e30f, 1.0e30f, 1.0e30f, 0.0f);
}

// append one chunk's hits.  (Two alternatives were measured in r01 and dropped, profiles/README.md section 4: 8 lanes per
// particle testing 8 candidates at a time with ballot / popc compaction -- no staging, no divergence, but its 4-byte
// stores hit 8 different list rows per particle and instruction, 1670 us against 750 us; and deferring the appends
// through shared-memory hit masks, so that a warp does not wait for its busiest lane in every chunk, 13 % slower: the
// kernel's shared-memory pipe is already the bottleneck of the test loop -- every group-broadcast LDS.128 costs 4
// wavefronts.)
__device__ __forceinline__ void nbw_append_now(unsigned m, int base, unsigned long long row_addr, int k_max, int &cnt) {
    while (m) {
        const int t1 = __ffs(m) - 1;
        m &= m - 1;
        // entry cnt of the tile-sliced row (md_nbr_off): 32-bit index, one IMAD.WIDE + STG.  A full row keeps counting (the
        // overflow is reported by the caller) and overwrites its last slot, so every entry below nbr_cnt is a valid slot.
        asm volatile("st.global.b32 [%0], %1;" ::"l"(row_addr + (unsigned long long)(unsigned)min(cnt, k_max) * 128ull), "r"(base + t1) : "memory");
        ++cnt;
    }
}

template <bool EXPANDED, int NBW_G>
__global__ void __launch_bounds__(NBW_WARPS * 32, 9)
k_nbr_build_warp(const double *__restrict__ r, const uint4 *__restrict__ q, long long n_own, long long cap,
                 const int *__restrict__ cell, const int *__restrict__ cell_start, const int *__restrict__ cell_end, int nc,
                 double L, double halfL, double cell_len, double inv_h, double rl2, double rl2s_d, float bw, float rl2s_lo,
                 float rl2s_hi, int max_nbr, int run_max, int *__restrict__ nbr, int *__restrict__ nbr_cnt,
                 StepScalars *__restrict__ sc) {
    __shared__ float4 s_cand[NBW_WARPS][NBW_G][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const long long i0 = ((long long)blockIdx.x * NBW_WARPS + w) * 32;
    if (i0 >= n_own) return;   // whole warp out of range (only __syncwarp below)
    const long long i = i0 + lane;
    const bool valid = i < n_own;
    const long long il = valid ? i : n_own - 1;
    const uint4 qi = q[il];
    const int c = cell[il];
    const int cx = c % nc, cy = (c / nc) % nc, cz = c / (nc * nc);
    float4 *cand = s_cand[w][0];
    const int grp = lane / (32 / NBW_G);            // lane group: each group walks its own candidate stream
    float4 *cand_g = s_cand[w][grp];
    int *row0 = nbr + md_nbr_row_base(il, max_nbr);
    const unsigned long long row_addr = (unsigned long long)row0;
    const int k_max = max_nbr - 1;
    const float neg_bw = -bw;
    float bandmin = 3.0e38f;
    int cnt = 0;
    unsigned todo = __ballot_sync(0xffffffffu, valid);
    while (todo) {
        // ---- pick the particles of this pass: same (cy, cz) row as the first pending lane, cx within NBW_RUN cells
        const int leader = __ffs(todo) - 1;
        const int cl = __shfl_sync(0xffffffffu, c, leader);
        const int cxl = cl % nc, cyl = (cl / nc) % nc, czl = cl / (nc * nc);
        const bool mine = ((todo >> lane) & 1u) && cy == cyl && cz == czl && cx >= cxl && cx < cxl + run_max;
        todo &= ~__ballot_sync(0xffffffffu, mine);
        const int cxb = __reduce_max_sync(0xffffffffu, mine ? cx : cxl);
        const int ncx = cxb - cxl + 3;   // candidate cells cxl-1 .. cxb+1 (periodic); run_max <= nc - 3 keeps them distinct
        // common origin = centre of the i-cells of the pass, in fixed point
        const uint32_t ox = md_to_fixed((cxl + 0.5 * (double)(cxb - cxl + 1)) * cell_len, inv_h);
        const uint32_t oy = md_to_fixed((cyl + 0.5) * cell_len, inv_h), oz = md_to_fixed((czl + 0.5) * cell_len, inv_h);
        const float xi = (float)(int)(qi.x - ox), yi = (float)(int)(qi.y - oy), zi = (float)(int)(qi.z - oz);
        const float px2 = EXPANDED ? -2.0f * xi : xi, py2 = EXPANDED ? -2.0f * yi : yi, pz2 = EXPANDED ? -2.0f * zi : zi;
        const float ki = EXPANDED ? (float)(fma((double)xi, (double)xi, fma((double)yi, (double)yi, (double)zi * (double)zi)) - rl2s_d)
                                  : -(float)rl2s_d;
        // Candidate streams per lane group: the 8 consecutive particles of a group span 1-2 cells, so the group only
        // needs the cells [its first cell - 1, its last cell + 1] of every row -- about 3.6 cells instead of the 5.5
        // the whole tile needs.  tlo / thi index the pass's candidate cells (t = 0 is cell cxl - 1).
        int tlo = mine ? cx - cxl : 64, thi = mine ? cx - cxl + 2 : -1;
#pragma unroll
        for (int o = 1; o < 32 / NBW_G; o <<= 1) {
            tlo = min(tlo, __shfl_xor_sync(0xffffffffu, tlo, o));
            thi = max(thi, __shfl_xor_sync(0xffffffffu, thi, o));
        }
        const bool grp_on = thi >= 0;
        for (int dz = -1; dz <= 1; ++dz) {
            const int z = (czl + dz + nc) % nc;
            // lanes 0..23 fetch the slot ranges of the 3 x ncx candidate cells of this z layer
            int cs_l = 0, ce_l = 0;
            {
                const int dyi = lane >> 3, t = lane & 7;
                if (dyi < 3 && t < ncx) {
                    const int y = (cyl + dyi - 1 + nc) % nc;
                    const int x = (cxl - 1 + t + nc) % nc;
                    const int cc = (z * nc + y) * nc + x;
                    cs_l = cell_start[cc];
                    ce_l = cell_end[cc];
                }
            }
            // a row whose ncx cells are all occupied and slot-contiguous can be walked per group
            unsigned row_good;
            {
                const int ce_prev = __shfl_up_sync(0xffffffffu, ce_l, 1);
                const int t = lane & 7;
                const bool good = lane >= 24 || t >= ncx || (ce_l > cs_l && (t == 0 || cs_l == ce_prev));
                row_good = __ballot_sync(0xffffffffu, good);
            }
            for (int dyi = 0; dyi < 3; ++dyi) {
                if (((row_good >> (dyi * 8)) & 0xffu) == 0xffu) {
                    // ---- grouped walk: every lane group streams its own slot range [S, E) ----
                    // (all lanes take part in the shuffles; groups without a particle in this pass get an empty range)
                    const int S_all = __shfl_sync(0xffffffffu, cs_l, dyi * 8 + (grp_on ? tlo : 0));
                    const int E_all = __shfl_sync(0xffffffffu, ce_l, dyi * 8 + (grp_on ? thi : 0));
                    const int S = grp_on ? S_all : 0, E = grp_on ? E_all : 0;
                    const int len_max = __reduce_max_sync(0xffffffffu, E - S);
                    int Sg[NBW_G], Eg[NBW_G];
#pragma unroll
                    for (int k = 0; k < NBW_G; ++k) {
                        Sg[k] = __shfl_sync(0xffffffffu, S, k * (32 / NBW_G));
                        Eg[k] = __shfl_sync(0xffffffffu, E, k * (32 / NBW_G));
                    }
                    for (int c0 = 0; c0 < len_max; c0 += 32) {
                        __syncwarp();
#pragma unroll
                        for (int k = 0; k < NBW_G; ++k) {   // lane L stages candidate L of every group's chunk
                            const int Sk = Sg[k], Ek = Eg[k];
                            const int jj = Sk + c0 + lane;
                            float4 cd = nbw_dummy<EXPANDED>();
                            if (jj < Ek) cd = nbw_stage<EXPANDED>(q[jj], ox, oy, oz);
                            s_cand[w][k][lane] = cd;
                        }
                        __syncwarp();
                        const int base = S + c0;
                        float bm = 3.0e38f;
                        unsigned m = nbw_test_chunk<EXPANDED>(cand_g, min(32, len_max - c0), px2, py2, pz2, ki, neg_bw, bm);
                        if (mine) bandmin = fminf(bandmin, bm);
                        else m = 0u;
                        const long long self = i - base;
                        if (self >= 0 && self < 32) m &= ~(1u << self);
                        nbw_append_now(m, base, row_addr, k_max, cnt);
                    }
                    continue;
                }
                int rs = 0, re = 0;   // pending (merged) slot range
                for (int t = 0; t <= ncx; ++t) {
                    int s1 = 0, e1 = 0;
                    if (t < ncx) {
                        s1 = __shfl_sync(0xffffffffu, cs_l, dyi * 8 + t);
                        e1 = __shfl_sync(0xffffffffu, ce_l, dyi * 8 + t);
                        if (e1 == s1) continue;
                        if (re > rs && s1 == re) { re = e1; continue; }   // contiguous with the pending range
                    }
                    // flush the pending range, 32 candidates at a time
                    for (int base = rs; base < re; base += 32) {
                        const int jj = base + lane;
                        float4 cd = nbw_dummy<EXPANDED>();
                        if (jj < re) cd = nbw_stage<EXPANDED>(q[jj], ox, oy, oz);
                        __syncwarp();
                        cand[lane] = cd;
                        __syncwarp();
                        float bm = 3.0e38f;
                        unsigned m = nbw_test_chunk<EXPANDED>(cand, min(32, re - base), px2, py2, pz2, ki, neg_bw, bm);
                        if (mine) bandmin = fminf(bandmin, bm);
                        else m = 0u;
                        const long long self = i - base;
                        if (self >= 0 && self < 32) m &= ~(1u << self);
                        nbw_append_now(m, base, row_addr, k_max, cnt);
                    }
                    rs = s1; re = e1;
                }
            }
        }
    }
    if (valid) {
        if (__builtin_expect(bandmin <= bw, 0)) {   // a candidate sat in the guard band: canonical float64 redo of this particle
            bool band = false;
            cnt = nb_particle<true>(i, qi, c, r, q, cap, cell_start, cell_end, nc, L, halfL, rl2, rl2s_lo, rl2s_hi, max_nbr, row0, band);
        }
        if (cnt > max_nbr) {
            sc->overflow_flag = 1;
            cnt = max_nbr;
        }
        nbr_cnt[i] = cnt;
        for (int k = cnt; k < ((cnt + NBG - 1) / NBG) * NBG; ++k) row0[k * 32] = (int)i;  // padding, masked by j != i
    }
}

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

// ---------------------------------------------------------------------------------------------------
// K5 (cut-off flavour): particle-wise integrators.  Same per-component arithmetic as integrate.cu, plus
// the skin-displacement trigger, the packed position copy and the id-scatter of the dataset rows.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_cut_kick_drift(double *__restrict__ r, double *__restrict__ v, const double *__restrict__ acc,
                 const double *__restrict__ rb, long long n, long long cap, int do_bc, int do_kick, int do_drift, double L,
                 double halfL, double dt, double trig2, int fp32, double inv_h, void *__restrict__ pos4,
                 StepScalars *__restrict__ sc, int tag) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    // all loads first (the stores below go through the same base pointers; see k_cut_kick_finish)
    double rr[3], vv[3], aa[3], bb[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        rr[d] = r[d * cap + i];
        vv[d] = v[d * cap + i];
        aa[d] = do_kick ? acc[d * cap + i] : 0.0;
        bb[d] = rb[d * cap + i];
    }
    double x[3], d2 = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        double xx = rr[d], u = vv[d];
        if (do_bc) {  // one-shot strict wrap, simulator.py:165-166
            if (xx < 0.0) xx += L;
            if (xx > L) xx -= L;
        }
        if (do_kick) u += __dmul_rn(__dmul_rn(0.5, aa[d]), dt);
        if (do_drift) xx += __dmul_rn(u, dt);
        r[d * cap + i] = xx;
        if (do_kick) v[d * cap + i] = u;
        x[d] = xx;
        const double dd = md_minimg_exact(__dsub_rn(xx, bb[d]), L, halfL);
        d2 = __dadd_rn(d2, __dmul_rn(dd, dd));
    }
    if (do_drift && d2 > trig2) sc->rebuild_flag = tag;   // the skin trigger of THIS drift (tags: see StepScalars)
    if (fp32) ((uint4 *)pos4)[i] = make_uint4(md_to_fixed(x[0], inv_h), md_to_fixed(x[1], inv_h), md_to_fixed(x[2], inv_h), 0u);
    else ((double4 *)pos4)[i] = make_double4(x[0], x[1], x[2], 0.0);
}

// What the finish kernel needs to also START the next step (wrap + first half kick + drift, i.e. k_cut_kick_drift
// fused behind the second half kick): with a(t+dt) reused as the next step's a(t), v and acc are already in registers,
// so the fusion saves one launch and re-reading v / acc / re-writing v (96 B per particle and step).
struct NextDrift {
    const double *rb;     // build positions (skin trigger)
    double L, halfL, trig2, inv_h;
    void  *pos4;
    int    fp32;
    int    tag;           // trigger tag of the fused drift
    int   *host_flag;     // single GPU: mapped host word that receives (trigger fired) from the last block; else nullptr
};

template <bool FUSE>
__global__ void __launch_bounds__(256, FUSE ? 3 : 5)
k_cut_kick_finish(double *__restrict__ r, double *__restrict__ v, const double *__restrict__ acc,
                  const int *__restrict__ id, long long n, long long cap, long long row_stride, int *__restrict__ id_row,
                  double dt, double g_dt, StepScalars *__restrict__ sc, double *__restrict__ ke_part,
                  double *__restrict__ r_row, double *__restrict__ v_row, int guard_tag, FinalizeArgs fin,
                  const double *__restrict__ pe_force, int n_pe_force, double *__restrict__ pe_fold, NextDrift nd) {
    __shared__ double red[32];
    __shared__ bool   s_last;
    if (guard_tag && sc->rebuild_flag == guard_tag) return;
    double v2 = 0.0;
    const double scale = sc->scale;
    // grid-stride: a bounded number of blocks (8 per SM) keeps the partial-sum array and the number of same-address
    // atomics of the last-block protocol below small
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int    gid = id[i];
        // single GPU: rows are written in ORIGINAL particle order (core.py:236-238), scattered by id.
        // multi-rank: compact rows in slot order plus the ids; the host scatters them into the shared file.
        const long long col = id_row ? i : (long long)gid;
        if (id_row) id_row[i] = gid;
        // all loads first: the stores below go through the same base pointers (r, v), so the compiler would otherwise
        // order the next component's loads behind them and serialise three memory round trips per particle
        double rr[3], vv[3], aa[3], bb[3];
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            rr[d] = r[d * cap + i];
            vv[d] = v[d * cap + i];
            aa[d] = acc[d * cap + i];
            bb[d] = FUSE ? nd.rb[d * cap + i] : 0.0;
        }
        double x[3], d2 = 0.0;
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            const double hk = __dmul_rn(__dmul_rn(0.5, aa[d]), dt);
            double u = vv[d] + hk;
            if (d == 2 && g_dt != 0.0) u -= g_dt;
            if (scale != 1.0) u *= scale;
            v2 += u * u;
            double xx = rr[d];
            if (r_row) {
                r_row[d * row_stride + col] = xx;
                v_row[d * row_stride + col] = u;
            }
            if (FUSE) {   // the next step: one-shot strict wrap, v += a dt / 2, r += v dt (same arithmetic as k_cut_kick_drift)
                if (xx < 0.0) xx += nd.L;
                if (xx > nd.L) xx -= nd.L;
                u += hk;
                xx += __dmul_rn(u, dt);
                r[d * cap + i] = xx;
                x[d] = xx;
                const double dd = md_minimg_exact(__dsub_rn(xx, bb[d]), nd.L, nd.halfL);
                d2 = __dadd_rn(d2, __dmul_rn(dd, dd));
            }
            v[d * cap + i] = u;
        }
        if (FUSE) {
            if (d2 > nd.trig2) sc->rebuild_flag = nd.tag;
            if (nd.fp32) ((uint4 *)nd.pos4)[i] = make_uint4(md_to_fixed(x[0], nd.inv_h), md_to_fixed(x[1], nd.inv_h), md_to_fixed(x[2], nd.inv_h), 0u);
            else ((double4 *)nd.pos4)[i] = make_double4(x[0], x[1], x[2], 0.0);
        }
    }
    const double b = md_block_sum(v2, red);
    // every block also folds its slice of the force kernel's PE partials, so that the last block reads gridDim.x values
    double pe_loc = 0.0;
    {
        const int chunk = (n_pe_force + (int)gridDim.x - 1) / (int)gridDim.x;
        const int k1 = min(n_pe_force, ((int)blockIdx.x + 1) * chunk);
        for (int k = (int)blockIdx.x * chunk + (int)threadIdx.x; k < k1; k += blockDim.x) pe_loc += pe_force[k];
    }
    const double bpe = md_block_sum(pe_loc, red);
    // the step's reductions and bookkeeping (finalize.cuh) are done by whichever block publishes
    // its partials last: no separate single-block launch on the critical path of every step
    if (threadIdx.x == 0) {
        ke_part[blockIdx.x] = b;
        pe_fold[blockIdx.x] = bpe;
        __threadfence();
        s_last = atomicAdd(&sc->finish_blocks_done, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (s_last) {
        __threadfence();
        md_finalize_body(fin, sc, red);
        if (threadIdx.x == 0) {
            sc->finish_blocks_done = 0u;
            // every block raised its trigger before bumping the counter: the flag of the fused drift is final here, so the
            // host learns it without a separate k_publish_flag launch
            if (FUSE && nd.host_flag) *nd.host_flag = *(volatile int *)&sc->rebuild_flag == nd.tag;
        }
    }
}

// world > 1: the temperature / quench-scale bookkeeping of md_finalize_body redone from the all-reduced mean_ke
__global__ void k_mr_quench_fix(StepScalars *__restrict__ sc, const double *__restrict__ ke_glob, const double *__restrict__ quench_next,
                                double kb, int redo) {
    const double mean_ke = *ke_glob;
    sc->ke_prev = mean_ke;
    if (!redo) return;
    // sc->temp was only overwritten by the finish kernel if the next step has a target, in which case it is
    // overwritten again here; otherwise it still holds self.temp
    double scale = 1.0, temp = sc->temp;
    if (quench_next) {
        for (int t = 0; t < 2; ++t) {
            const double target = quench_next[t];
            if (target == target) {
                temp = (2.0 / 3.0) * mean_ke / kb;
                if (temp != 0.0) scale *= sqrt(target / temp);
            }
        }
    }
    sc->temp = temp;
    sc->scale = scale;
}

__global__ void k_unsort(const double *__restrict__ src, const int *__restrict__ id, long long n, long long cap,
                         long long n_global, double *__restrict__ dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int gid = id[i];
#pragma unroll
    for (int d = 0; d < 3; ++d) dst[d * n_global + gid] = src[d * cap + i];
}

// The host learns the skin trigger through a store into mapped pinned memory instead of a D2H memcpy: a copy would
// queue on the copy engine behind the bulk row transfers of the previous batch (run_async pipelining) and stall
// every step for milliseconds.
__global__ void k_publish_flag(const StepScalars *__restrict__ sc, int *__restrict__ host_flag, int tag) { *host_flag = sc->rebuild_flag == tag; }

// Small device -> host read-backs of the rebuild (counts, class boundaries) are stores into mapped pinned memory for
// the same reason: dst[i] = src[i * stride].
__global__ void k_peek(const int *__restrict__ src, long long stride, int n, int *__restrict__ host_dst) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) host_dst[i] = src[(long long)i * stride];
}

__global__ void k_iota_from(int *a, long long n, int first) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = first + (int)i;
}
__global__ void k_iota(int *a, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = (int)i;
}

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

// ---------------------------------------------------------------------------------------------------
// K6: domain decomposition -- rebuild-time migration / ghost selection and per-step halo refresh
// ---------------------------------------------------------------------------------------------------
// pass 0 (count) / pass 1 (fill): every held particle goes to each rank whose extended block contains its cell
struct RouteDst { PMsg *pool[MD_MAXW]; };   // per destination rank: where its messages are written (fill pass)
__global__ void k_mr_route(const double *__restrict__ r, const double *__restrict__ v, const int *__restrict__ id, long long n,
                           long long cap, double L, double cl, int nc, Decomp D, int fill, int *__restrict__ counter,
                           RouteDst dst) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = r[i], y = r[cap + i], z = r[2 * cap + i];
    const int cx = md_cell_coord(x, L, cl, nc), cy = md_cell_coord(y, L, cl, nc), cz = md_cell_coord(z, L, cl, nc);
    for (int q = 0; q < D.world; ++q) {
        if (!dc_in_ext(D, q, cx, cy, cz)) continue;
        const int pos = atomicAdd(&counter[q], 1);
        if (fill) {
            PMsg m;
            m.r[0] = x; m.r[1] = y; m.r[2] = z;
            m.v[0] = v[i]; m.v[1] = v[cap + i]; m.v[2] = v[2 * cap + i];
            m.id = id[i];
            dst.pool[q][pos] = m;
        }
    }
}
__global__ void k_mr_unpack(const PMsg *__restrict__ pool, long long n, long long dst0, long long cap, double *__restrict__ r,
                            double *__restrict__ v, int *__restrict__ id) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const PMsg m = pool[k];
    const long long s = dst0 + k;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        r[d * cap + s] = m.r[d];
        v[d * cap + s] = m.v[d];
    }
    id[s] = (int)m.id;
}
// ---- rebuild-time routing over peer memory (NCCL transport on one node) ----
// Control block every rank exposes next to its halo buffers; peers store into it over NVLink.
struct PeerCtrl {
    int                cnt_from[MD_MAXW];       // messages rank p appended to my inbox segment p in the current rebuild
    unsigned long long asm_seq_from[MD_MAXW];   // ... valid once this reaches the rebuild number
    long long          halo_off_from[MD_MAXW];  // where MY halo segment starts inside rank p's receive buffer (doubles)
    unsigned long long off_seq_from[MD_MAXW];   // ... valid once this reaches the rebuild number
};
struct RouteP2P {
    PMsg *seg[MD_MAXW];   // my segment inside rank q's inbox (q == rank: unused)
    int   seg_cap;
};
// one pass: what this rank keeps goes straight into the staging arrays, everything else straight into the inbox of
// the rank that needs it (as owner or as ghost); no count pass, no send pool, no ncclSend/ncclRecv
__global__ void k_mr_route_p2p(const double *__restrict__ r, const double *__restrict__ v, const int *__restrict__ id, long long n,
                               long long cap, double L, double cl, int nc, Decomp D, int *__restrict__ cursor, RouteP2P dst,
                               double *__restrict__ r_st, double *__restrict__ v_st, int *__restrict__ id_st,
                               StepScalars *__restrict__ sc) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double x = r[i], y = r[cap + i], z = r[2 * cap + i];
    const int cx = md_cell_coord(x, L, cl, nc), cy = md_cell_coord(y, L, cl, nc), cz = md_cell_coord(z, L, cl, nc);
    for (int q = 0; q < D.world; ++q) {
        if (!dc_in_ext(D, q, cx, cy, cz)) continue;
        const int pos = atomicAdd(&cursor[q], 1);
        if (q == D.rank) {
            if (pos < cap) {
                r_st[pos] = x; r_st[cap + pos] = y; r_st[2 * cap + pos] = z;
                v_st[pos] = v[i]; v_st[cap + pos] = v[cap + i]; v_st[2 * cap + pos] = v[2 * cap + i];
                id_st[pos] = id[i];
            }
        } else if (pos < dst.seg_cap) {
            PMsg m;
            m.r[0] = x; m.r[1] = y; m.r[2] = z;
            m.v[0] = v[i]; m.v[1] = v[cap + i]; m.v[2] = v[2 * cap + i];
            m.id = id[i];
            dst.seg[q][pos] = m;
        } else {
            sc->overflow_flag = 1;
        }
    }
}
struct CtrlPtrs { PeerCtrl *p[MD_MAXW]; };
__global__ void k_mr_signal(const int *__restrict__ cursor, CtrlPtrs peers, int world, int rank, unsigned long long seq) {
    __threadfence_system();   // the messages of k_mr_route_p2p first
    const int q = threadIdx.x;
    if (q < world && q != rank) {
        peers.p[q]->cnt_from[rank] = cursor[q];
        __threadfence_system();
        *(volatile unsigned long long *)&peers.p[q]->asm_seq_from[rank] = seq;
    }
}
// wait for every peer's segment, lay the segments out behind the self-kept particles; roff[q] = first staging slot of
// source q, roff[W] = n_tot, roff[W + 1] = n_self; the two totals also go to the host through mapped memory
__global__ void k_mr_collect(const PeerCtrl *__restrict__ ctrl, const int *__restrict__ cursor, int world, int rank,
                             unsigned long long seq, int *__restrict__ roff, int *__restrict__ host_out,
                             StepScalars *__restrict__ sc) {
    const int q = threadIdx.x;
    if (q < world && q != rank) {
        const volatile unsigned long long *f = &ctrl->asm_seq_from[q];
        const long long t0 = clock64();
        while (*f < seq)
            if (clock64() - t0 > 8000000000LL) { sc->overflow_flag = 2; break; }
    }
    __syncthreads();
    __threadfence_system();
    if (threadIdx.x == 0) {
        int run = cursor[rank];
        for (int p = 0; p < world; ++p) {
            roff[p] = run;
            if (p != rank) run += ((const volatile PeerCtrl *)ctrl)->cnt_from[p];
        }
        roff[world] = run;
        roff[world + 1] = cursor[rank];
        host_out[0] = run;
        host_out[1] = cursor[rank];
    }
}
__global__ void k_mr_unpack_p2p(const PMsg *__restrict__ inbox, long long seg_cap, const int *__restrict__ roff, int world, int rank,
                                long long cap, double *__restrict__ r, double *__restrict__ v, int *__restrict__ id) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // index among the RECEIVED particles
    const int n_self = roff[world + 1], n_tot = roff[world];
    if (t >= n_tot - n_self) return;
    const int s = n_self + (int)t;
    int p = 0;
    for (int k = 0; k < world; ++k)
        if (k != rank && s >= roff[k]) p = k;   // sources are laid out in rank order: the last one starting at or before s
    const PMsg m = inbox[(long long)p * seg_cap + (s - roff[p])];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        r[d * cap + s] = m.r[d];
        v[d * cap + s] = m.v[d];
    }
    id[s] = (int)m.id;
}
// the receive layout of this rank's halo buffer, told to the senders: off_for[p] = 3 * (entries from ranks before p) + p
struct HaloOff { long long off_for[MD_MAXW]; };
__global__ void k_publish_halo_off(CtrlPtrs peers, HaloOff H, int world, int rank, unsigned long long seq) {
    const int p = threadIdx.x;
    if (p < world && p != rank) {
        peers.p[p]->halo_off_from[rank] = H.off_for[p];
        __threadfence_system();
        *(volatile unsigned long long *)&peers.p[p]->off_seq_from[rank] = seq;
    }
}

// flag[q][s]: owned slot s must be SENT to rank q every step (its cell lies in q's halo); ghost slot s is
// RECEIVED from rank q (q owns its cell).  One scan over the whole [world][n_tot] array then yields, per peer,
// the send list followed by the receive list, both in slot order = canonical (cell id, particle id) order --
// the same order on the sending and on the receiving rank, so no index exchange is needed.
__global__ void k_halo_flags(const int *__restrict__ cell, long long base, long long n_own, long long n_tot, int nc, Decomp D,
                             int *__restrict__ flag) {
    // only the owned BOUNDARY slots [base, n_own) can be in somebody's halo, and only the ghost slots [n_own, n_tot)
    // are received: the interior slots [0, base) are left out of the flag / scan arrays altogether
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long m = n_tot - base, s = base + t;
    if (t >= m) return;
    const int c = cell[s];
    const int cx = c % nc, cy = (c / nc) % nc, cz = c / (nc * nc);
    const int owner = dc_owner(D, cx, cy, cz);
    for (int q = 0; q < D.world; ++q) {
        int f;
        if (s < n_own) f = (q != D.rank && dc_in_ext(D, q, cx, cy, cz)) ? 1 : 0;
        else f = (owner == q) ? 1 : 0;
        flag[(long long)q * m + t] = f;
    }
}
__global__ void k_halo_lists(const int *__restrict__ flag, const int *__restrict__ scan, long long m, long long base, int world,
                             int *__restrict__ idx) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= m * world) return;
    if (!flag[t]) return;
    const long long q = t / m, k = t - q * m;
    idx[q * m + (scan[t] - scan[q * m])] = (int)(base + k);
}
// One message per peer and step: [3 * n positions][1 double: this rank's skin-trigger flag].  Every rank sends to
// every other rank every step (an empty halo still carries the flag), so after the unpack each rank holds the
// MAX of all flags -- the collective rebuild decision rides on the halo traffic instead of a separate allreduce.
struct HaloTab {
    int       world, rank;
    long long ent0[MD_MAXW + 1];  // prefix sum of the per-peer entry counts (self: 0 entries)
    long long idx0[MD_MAXW];      // start of the peer's index list inside halo_idx
};
__device__ __forceinline__ int halo_peer_of(const HaloTab &T, long long k) {
    int q = 0;
    while (q + 1 < T.world && k >= T.ent0[q + 1]) ++q;
    return q;
}
__global__ void k_halo_pack(const double *__restrict__ r, long long cap, const int *__restrict__ idx, HaloTab T,
                            const StepScalars *__restrict__ sc, double *__restrict__ buf, int tag) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = T.ent0[T.world];
    if (k < total) {
        const int q = halo_peer_of(T, k);
        const long long l = k - T.ent0[q];
        const int s = idx[T.idx0[q] + l];
        double *m = buf + 3 * T.ent0[q] + q;  // q flag slots precede this peer's segment
        m[3 * l] = r[s];
        m[3 * l + 1] = r[cap + s];
        m[3 * l + 2] = r[2 * cap + s];
    } else if (k < total + T.world) {
        const int q = (int)(k - total);
        if (q != T.rank) buf[3 * T.ent0[q + 1] + q] = (double)(sc->rebuild_flag == tag);
    }
}
// Arrival counters: rank p bumps counter[p] in rank q's buffer to `seq` after its stores of exchange number `seq`.
// The consumer polls its OWN memory.  A bounded spin (about 4 s) reports a comm error instead of hanging the GPU.
struct HaloWait {
    const unsigned long long *counters;   // nullptr: data came through ncclRecv, nothing to wait for
    unsigned long long        seq;
};
__device__ __forceinline__ void halo_wait_arrivals(const HaloWait &Wt, int world, int rank, StepScalars *__restrict__ sc) {
    if (!Wt.counters) return;
    if ((int)threadIdx.x < world && (int)threadIdx.x != rank) {
        const volatile unsigned long long *f = Wt.counters + threadIdx.x;
        const long long t0 = clock64();
        while (*f < Wt.seq) {
            if (clock64() - t0 > 8000000000LL) { sc->overflow_flag = 2; break; }
        }
    }
    __syncthreads();
    __threadfence_system();
}

__global__ void k_halo_unpack(const double *__restrict__ buf, const int *__restrict__ idx, HaloTab T, long long cap,
                              double *__restrict__ r, int fp32, double inv_h, void *__restrict__ pos4,
                              StepScalars *__restrict__ sc, HaloWait Wt, int tag) {
    halo_wait_arrivals(Wt, T.world, T.rank, sc);
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = T.ent0[T.world];
    if (k < total) {
        const int q = halo_peer_of(T, k);
        const long long l = k - T.ent0[q];
        const int s = idx[T.idx0[q] + l];
        const double *m = buf + 3 * T.ent0[q] + q;
        const double x = __ldcg(m + 3 * l), y = __ldcg(m + 3 * l + 1), z = __ldcg(m + 3 * l + 2);
        r[s] = x; r[cap + s] = y; r[2 * cap + s] = z;
        if (fp32) ((uint4 *)pos4)[s] = make_uint4(md_to_fixed(x, inv_h), md_to_fixed(y, inv_h), md_to_fixed(z, inv_h), 0u);
        else ((double4 *)pos4)[s] = make_double4(x, y, z, 0.0);
    } else if (k < total + T.world) {
        const int q = (int)(k - total);
        if (q != T.rank && __ldcg(buf + 3 * T.ent0[q + 1] + q) != 0.0) sc->rebuild_flag = tag;
    }
}

// fused pack + transfer: the owned boundary positions (and this rank's skin-trigger flag) are stored straight into
// the receive buffers of the ranks that hold them as ghosts -- no staging buffer, no ncclSend/ncclRecv
struct PeerDst { double *base[MD_MAXW]; };   // rank q's receive buffer (current parity); my segment starts at the offset q published
__global__ void k_halo_push(const double *__restrict__ r, long long cap, const int *__restrict__ idx, HaloTab T,
                            StepScalars *__restrict__ sc, PeerDst P, const PeerCtrl *__restrict__ ctrl, unsigned long long rebuild_seq,
                            int tag) {
    __shared__ long long s_off[MD_MAXW];
    if ((int)threadIdx.x < T.world && (int)threadIdx.x != T.rank) {   // the receivers' layouts of THIS list generation
        const volatile unsigned long long *f = &ctrl->off_seq_from[threadIdx.x];
        const long long t0 = clock64();
        while (*f < rebuild_seq)
            if (clock64() - t0 > 8000000000LL) { sc->overflow_flag = 2; break; }
        __threadfence_system();
        s_off[threadIdx.x] = ((const volatile PeerCtrl *)ctrl)->halo_off_from[threadIdx.x];
    }
    __syncthreads();
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = T.ent0[T.world];
    if (k < total) {
        const int q = halo_peer_of(T, k);
        const long long l = k - T.ent0[q];
        const int s = idx[T.idx0[q] + l];
        double *m = P.base[q] + s_off[q];
        m[3 * l] = r[s];
        m[3 * l + 1] = r[cap + s];
        m[3 * l + 2] = r[2 * cap + s];
    } else if (k < total + T.world) {
        const int q = (int)(k - total);
        if (q != T.rank) P.base[q][s_off[q] + 3 * (T.ent0[q + 1] - T.ent0[q])] = (double)(sc->rebuild_flag == tag);
    }
}
struct PeerCnt { unsigned long long *cnt[MD_MAXW]; };   // counter[this rank] inside peer q's buffer
__global__ void k_halo_signal(PeerCnt P, int world, int rank, unsigned long long seq) {
    __threadfence_system();   // the stores of k_halo_push (previous kernel on this stream) are visible system-wide first
    const int q = threadIdx.x;
    if (q < world && q != rank) *(volatile unsigned long long *)P.cnt[q] = seq;
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
static int bits_for(unsigned long long maxval) {
    int b = 1;
    while ((maxval >> b) != 0) ++b;
    return b;
}
static inline double fixed_inv_h(const mdsea_ctx *c) { return 4294967296.0 / c->box.L; }
static inline int g256(long long n) { return md_div_up(n > 0 ? n : 1, 256); }
// blocks of a list force launch over the owned slots [i0, i1)
static inline int force_grid(long long i0, long long i1) { return md_div_up(i1 - i0, NB_THREADS); }

static int scan_exclusive(mdsea_ctx *c, const int *in, int *out, long long n) {
    CellState *s = c->cells;
    const int nb = md_div_up(n, SCAN_THREADS * SCAN_ITEMS);
    k_scan_block<<<nb, SCAN_THREADS, 0, c->stream>>>(in, out, n, s->scan_tmp);
    k_scan_top<<<1, 1024, 0, c->stream>>>(s->scan_tmp, nb);
    k_scan_add<<<nb, SCAN_THREADS, 0, c->stream>>>(out, n, s->scan_tmp);
    c->stats.kernel_launches += 3;
    MD_CUDA(c, cudaGetLastError());
    return MDSEA_OK;
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
    const int W = p.world;
    // decomposition
    Decomp &D = s->D;
    memset(&D, 0, sizeof(D));
    D.world = W; D.rank = p.rank; D.nc = s->nc;
    for (int d = 0; d < 3; ++d) D.p[d] = W > 1 ? p.proc_grid[d] : 1;
    if (D.p[0] * D.p[1] * D.p[2] != W) { md_set_error(c, "cutoff path: proc_grid does not multiply to world"); return MDSEA_ERR_BAD_ARG; }
    for (int d = 0; d < 3; ++d) {
        if (D.p[d] < 1 || D.p[d] > MD_MAXW || D.p[d] > s->nc) { md_set_error(c, "cutoff path: proc_grid entry out of range (needs >= 1 cell per rank and dimension)"); return MDSEA_ERR_BAD_ARG; }
        for (int k = 0; k <= D.p[d]; ++k) D.bnd[d][k] = (int)(((long long)k * s->nc) / D.p[d]);
    }
    const long long N = c->n_global;
    s->cap = W == 1 ? ((N + 31) / 32) * 32 : std::min<long long>(((N + 31) / 32) * 32, ((2 * (N / W) + 65536 + 31) / 32) * 32);
    s->n_pad = s->cap;
    if (p.max_neighbors > 0) s->max_nbr = p.max_neighbors;
    else {
        const double rho = (double)N / (p.boxlen * p.boxlen * p.boxlen);
        const double expect = 4.0 / 3.0 * 3.14159265358979 * s->rlist * s->rlist * s->rlist * rho;
        s->max_nbr = ((int)(expect * 1.6) + 32 + 7) / 8 * 8;
    }
    s->max_nbr = (s->max_nbr + NBG - 1) / NBG * NBG;
    s->id_bits = bits_for((unsigned long long)(N - 1));
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
    MD_CUDA(c, cudaMalloc(&s->qfix, sizeof(uint4) * (size_t)s->cap));
    // fp32 gathers go through the texture path (measured 3 % faster than ld.global.nc on this access pattern);
    // MDSEA_FORCE_TEX=0 switches it off, and a linear texture cannot address more than 2^27 elements
    if (p.precision == MDSEA_FP32 && !(getenv("MDSEA_FORCE_TEX") && !strcmp(getenv("MDSEA_FORCE_TEX"), "0")) && s->cap <= (1LL << 27)) {
        cudaResourceDesc rd;
        memset(&rd, 0, sizeof(rd));
        rd.resType = cudaResourceTypeLinear;
        rd.res.linear.devPtr = s->pos4;
        rd.res.linear.desc = cudaCreateChannelDesc<uint4>();
        rd.res.linear.sizeInBytes = sizeof(uint4) * (size_t)s->cap;
        cudaTextureDesc td;
        memset(&td, 0, sizeof(td));
        td.readMode = cudaReadModeElementType;
        MD_CUDA(c, cudaCreateTextureObject(&s->tex_pos, &rd, &td, nullptr));
    }
    s->radix_binning = getenv("MDSEA_BINNING") && !strcmp(getenv("MDSEA_BINNING"), "radix");   // A/B switch
    s->nbucket = (long long)(W > 1 ? 3 : 1) * s->ncell;
    if (s->radix_binning) {
        MD_CUDA(c, cudaMalloc(&s->key, sizeof(unsigned long long) * s->cap));
        MD_CUDA(c, cudaMalloc(&s->key_alt, sizeof(unsigned long long) * s->cap));
        MD_CUDA(c, cudaMalloc(&s->val, sizeof(int) * s->cap));
        MD_CUDA(c, cudaMalloc(&s->val_alt, sizeof(int) * s->cap));
        MD_CUDA(c, cudaMalloc(&s->hist, sizeof(int) * 256 * (size_t)s->sort_blocks));
        MD_CUDA(c, cudaMalloc(&s->hist_tot, sizeof(int) * 256));
    } else {
        MD_CUDA(c, cudaMalloc(&s->bkt, sizeof(int) * s->cap));
        MD_CUDA(c, cudaMalloc(&s->arrival, sizeof(int) * s->cap));
        MD_CUDA(c, cudaMalloc(&s->tmp_id, sizeof(int) * s->cap));
        MD_CUDA(c, cudaMalloc(&s->tmp_src, sizeof(int) * s->cap));
        MD_CUDA(c, cudaMalloc(&s->tmp_bkt, sizeof(int) * s->cap));
        MD_CUDA(c, cudaMalloc(&s->bcount, sizeof(int) * ((size_t)s->nbucket + 1)));
        MD_CUDA(c, cudaMalloc(&s->bstart, sizeof(int) * ((size_t)s->nbucket + 1)));
    }
    {
        const long long scan_max = std::max<long long>(s->nbucket + 1, (long long)W * s->cap + 1);
        MD_CUDA(c, cudaMalloc(&s->scan_tmp, sizeof(int) * (md_div_up(scan_max, SCAN_THREADS * SCAN_ITEMS) + 1)));
    }
    MD_CUDA(c, cudaMalloc(&s->cell, sizeof(int) * s->cap));
    MD_CUDA(c, cudaMalloc(&s->cell_start, sizeof(int) * ((size_t)s->ncell + 1)));
    MD_CUDA(c, cudaMalloc(&s->cell_end, sizeof(int) * ((size_t)s->ncell + 1)));
    MD_CUDA(c, cudaMalloc(&s->nbr, sizeof(int) * (size_t)s->max_nbr * (size_t)s->n_pad));
    MD_CUDA(c, cudaMalloc(&s->nbr_cnt, sizeof(int) * s->cap));
    MD_CUDA(c, cudaMalloc(&s->d_cnt, sizeof(int) * (MD_MAXW + 8)));
    MD_CUDA(c, cudaHostAlloc(&s->h_flag, sizeof(int) * 4, cudaHostAllocMapped));
    MD_CUDA(c, cudaHostGetDevicePointer((void **)&s->d_flag_mapped, s->h_flag, 0));
    MD_CUDA(c, cudaEventCreateWithFlags(&s->ev_drift, cudaEventDisableTiming));
    MD_CUDA(c, cudaEventCreateWithFlags(&s->ev_flag, cudaEventDisableTiming));
    MD_CUDA(c, cudaHostAlloc(&s->h_cnt_all, sizeof(int) * (MD_MAXW * MD_MAXW + 8), cudaHostAllocMapped));
    MD_CUDA(c, cudaHostGetDevicePointer((void **)&s->d_cnt_all_mapped, s->h_cnt_all, 0));
    const size_t row_stride = W == 1 ? (size_t)N : (size_t)s->cap;
    c->coord_rows_bytes = sizeof(double) * 3 * row_stride * c->ring;
    MD_CUDA(c, cudaMalloc(&c->d_r_rows, c->coord_rows_bytes));
    MD_CUDA(c, cudaMalloc(&c->d_v_rows, c->coord_rows_bytes));
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, p.device);
        s->finish_blocks = 10 * sms;
        s->finish_blocks_fused = 6 * sms;
        MD_CUDA(c, cudaMalloc(&s->pe_fold, sizeof(double) * s->finish_blocks));
        MD_CUDA(c, cudaMalloc(&s->ke_glob, sizeof(double)));
    }
    c->n_ke_part = md_div_up(s->cap, 256);
    MD_CUDA(c, cudaMalloc(&c->d_ke_part, sizeof(double) * c->n_ke_part));
    MD_CUDA(c, cudaMemset(c->d_ke_part, 0, sizeof(double) * c->n_ke_part));
    c->n_pe_part = force_grid(0, s->cap) + 2;
    MD_CUDA(c, cudaMalloc(&c->d_pe_part, sizeof(double) * c->n_pe_part));
    MD_CUDA(c, cudaMemset(c->d_pe_part, 0, sizeof(double) * c->n_pe_part));
    if (W > 1) {
        MD_CUDA(c, cudaMalloc(&s->d_cnt_all, sizeof(int) * MD_MAXW * MD_MAXW));
        MD_CUDA(c, cudaMalloc(&s->d_cursor, sizeof(int) * MD_MAXW));
        MD_CUDA(c, cudaMalloc(&s->halo_flag, sizeof(int) * ((size_t)W * s->cap + 1)));
        MD_CUDA(c, cudaMalloc(&s->halo_scan, sizeof(int) * ((size_t)W * s->cap + 1)));
        MD_CUDA(c, cudaMalloc(&s->halo_idx, sizeof(int) * (size_t)W * s->cap));
        MD_CUDA(c, cudaMalloc(&s->halo_send, sizeof(double) * (3 * (size_t)s->cap * std::min(W - 1, 3) + MD_MAXW)));
        MD_CUDA(c, cudaMalloc(&s->halo_recv, sizeof(double) * (3 * (size_t)s->cap + MD_MAXW)));
        c->id_rows_bytes = sizeof(int) * (size_t)s->cap * c->ring;
        MD_CUDA(c, cudaMalloc(&c->d_id_rows, c->id_rows_bytes));
        s->row_n_own.assign(c->ring, 0);
    }
    c->n = 0;
    return MDSEA_OK;
}

void md_cells_free(mdsea_ctx *c) {
    CellState *s = c->cells;
    if (!s) return;
    md_comm_free(c);
    if (s->tex_pos) cudaDestroyTextureObject(s->tex_pos);
    cudaFree(s->r_alt); cudaFree(s->v_alt); cudaFree(s->rb); cudaFree(s->id); cudaFree(s->id_alt); cudaFree(s->pos4);
    cudaFree(s->key); cudaFree(s->key_alt); cudaFree(s->val); cudaFree(s->val_alt); cudaFree(s->hist); cudaFree(s->hist_tot);
    cudaFree(s->bkt); cudaFree(s->arrival); cudaFree(s->tmp_id); cudaFree(s->tmp_src); cudaFree(s->tmp_bkt); cudaFree(s->bcount); cudaFree(s->bstart);
    cudaFree(s->qfix); cudaFree(s->cell); cudaFree(s->cell_start); cudaFree(s->cell_end); cudaFree(s->nbr); cudaFree(s->nbr_cnt);
    cudaFree(s->scratch); cudaFree(s->d_cnt); cudaFree(s->sendpool); cudaFree(s->recvpool); cudaFree(s->d_cnt_all);
    cudaFree(s->d_cursor); cudaFree(s->halo_flag); cudaFree(s->halo_scan); cudaFree(s->halo_idx); cudaFree(s->scan_tmp);
    cudaFree(s->halo_send); cudaFree(s->halo_recv); cudaFree(s->p2p_buf); cudaFree(s->d_roff);
    cudaFree(s->pe_fold); cudaFree(s->ke_glob);
    if (s->h_flag) cudaFreeHost(s->h_flag);
    if (s->ev_flag) cudaEventDestroy(s->ev_flag);
    if (s->ev_drift) cudaEventDestroy(s->ev_drift);
    if (s->h_cnt_all) cudaFreeHost(s->h_cnt_all);
    delete s;
    c->cells = nullptr;
}

int md_cutoff_set_state(mdsea_ctx *c, const double *r, const double *v) {
    CellState *s = c->cells;
    const long long N = c->n_global;
    if (c->p.world == 1) {
        for (int d = 0; d < 3; ++d) {
            MD_CUDA(c, cudaMemcpyAsync(c->d_r + d * s->cap, r + d * N, sizeof(double) * N, cudaMemcpyHostToDevice, c->stream));
            MD_CUDA(c, cudaMemcpyAsync(c->d_v + d * s->cap, v + d * N, sizeof(double) * N, cudaMemcpyHostToDevice, c->stream));
        }
        k_iota<<<g256(N), 256, 0, c->stream>>>(s->id, N);
        c->stats.kernel_launches += 1;
        MD_CUDA(c, cudaStreamSynchronize(c->stream));
        c->n = N;
    } else {
        if (!c->comm) { md_set_error(c, "set_state: world > 1 needs comm_init first"); return MDSEA_ERR_STATE; }
        // Every rank uploads ONE CONTIGUOUS SLICE of the global arrays (particles [rank*N/W, (rank+1)*N/W)) and
        // simply holds it: the first rebuild routes every held particle to its owner and to the ranks that need it
        // as a ghost (k_mr_route_p2p / k_mr_route work on held particles, owned or not).  The host->device traffic
        // per rank is 1/W of the arrays instead of all of them, and no ownership test runs on the host.
        const int W = c->p.world;
        const long long lo = (N * c->p.rank) / W, hi = (N * (c->p.rank + 1)) / W, m = hi - lo;
        if (m > s->cap) { md_set_error(c, "set_state: particle capacity of this rank exceeded"); return MDSEA_ERR_CAPACITY; }
        for (int d = 0; d < 3; ++d) {
            MD_CUDA(c, cudaMemcpyAsync(c->d_r + d * s->cap, r + d * N + lo, sizeof(double) * m, cudaMemcpyHostToDevice, c->stream));
            MD_CUDA(c, cudaMemcpyAsync(c->d_v + d * s->cap, v + d * N + lo, sizeof(double) * m, cudaMemcpyHostToDevice, c->stream));
        }
        k_iota_from<<<g256(m), 256, 0, c->stream>>>(s->id, m, (int)lo);
        c->stats.kernel_launches += 1;
        MD_CUDA(c, cudaStreamSynchronize(c->stream));
        c->n = m;
    }
    s->n_tot = c->n;
    s->need_rebuild = true;
    c->stats.n_local = c->n;
    c->stats.n_ghost = 0;
    return MDSEA_OK;
}

// state generated in place by gen.cu: this rank holds the particles [lo, lo + m) in slots 0 .. m-1 (like the slice upload
// of md_cutoff_set_state); ids, counts and the pending rebuild follow
int md_cutoff_adopt_generated(mdsea_ctx *c, long long lo, long long m) {
    CellState *s = c->cells;
    k_iota_from<<<g256(m), 256, 0, c->stream>>>(s->id, m, (int)lo);
    c->stats.kernel_launches += 1;
    MD_CUDA(c, cudaGetLastError());
    c->n = m;
    s->n_tot = c->n;
    s->need_rebuild = true;
    c->stats.n_local = c->n;
    c->stats.n_ghost = 0;
    return MDSEA_OK;
}

// multi-rank front end of a rebuild: route every held particle to the ranks that need it (as owner or as ghost),
// exchange, and leave the union in the staging arrays (r_alt, v_alt, id_alt); returns the new n_tot.
static int mr_assemble(mdsea_ctx *c) {
    CellState *s = c->cells;
    const int W = c->p.world, me = c->p.rank;
    const long long n = c->n;
    s->asm_seq += 1;   // rebuild number (all ranks rebuild together)
    if (!s->p2p && !s->sendpool) {   // message pools of the ncclSend/ncclRecv (and in-process) transport, on first use
        s->sendpool_cap = (long long)std::min(W, 3) * s->cap;
        MD_CUDA(c, cudaMalloc(&s->sendpool, sizeof(PMsg) * (size_t)s->sendpool_cap));
        MD_CUDA(c, cudaMalloc(&s->recvpool, sizeof(PMsg) * (size_t)s->cap));
    }
    if (s->p2p) {
        RouteP2P dst;
        CtrlPtrs cp;
        for (int q = 0; q < MD_MAXW; ++q) {
            dst.seg[q] = q < W && q != me ? (PMsg *)((char *)s->peer_buf[q] + s->inbox_off) + (size_t)me * (size_t)s->seg_cap : nullptr;
            cp.p[q] = q < W ? (PeerCtrl *)((char *)s->peer_buf[q] + s->ctrl_off) : nullptr;
        }
        dst.seg_cap = (int)s->seg_cap;
        MD_CUDA(c, cudaMemsetAsync(s->d_cursor, 0, sizeof(int) * MD_MAXW, c->stream));
        k_mr_route_p2p<<<g256(n), 256, 0, c->stream>>>(c->d_r, c->d_v, s->id, n, s->cap, c->box.L, s->cell_len, s->nc, s->D, s->d_cursor, dst,
                                                       s->r_alt, s->v_alt, s->id_alt, c->d_sc);
        k_mr_signal<<<1, 32, 0, c->stream>>>(s->d_cursor, cp, W, me, s->asm_seq);
        k_mr_collect<<<1, 32, 0, c->stream>>>((const PeerCtrl *)((char *)s->p2p_buf + s->ctrl_off), s->d_cursor, W, me, s->asm_seq, s->d_roff,
                                              s->d_cnt_all_mapped, c->d_sc);
        MD_CUDA(c, cudaStreamSynchronize(c->stream));
        const long long n_tot = s->h_cnt_all[0], n_self = s->h_cnt_all[1];
        if (n_tot > s->cap || n_self > s->cap) {
            md_set_error(c, "rebuild: migration/ghost buffer capacity exceeded on this rank");
            return MDSEA_ERR_CAPACITY;
        }
        if (n_tot > n_self)
            k_mr_unpack_p2p<<<g256(n_tot - n_self), 256, 0, c->stream>>>((const PMsg *)((char *)s->p2p_buf + s->inbox_off), s->seg_cap, s->d_roff, W,
                                                                        me, s->cap, s->r_alt, s->v_alt, s->id_alt);
        c->stats.kernel_launches += 4;
        MD_CUDA(c, cudaGetLastError());
        s->n_tot = n_tot;
        return MDSEA_OK;
    }
    RouteDst none;
    memset(&none, 0, sizeof(none));
    MD_CUDA(c, cudaMemsetAsync(s->d_cnt, 0, sizeof(int) * MD_MAXW, c->stream));
    k_mr_route<<<g256(n), 256, 0, c->stream>>>(c->d_r, c->d_v, s->id, n, s->cap, c->box.L, s->cell_len, s->nc, s->D, 0, s->d_cnt,
                                               none);
    MD_TRY(md_comm_allgather_i32(c, s->d_cnt, MD_MAXW, s->d_cnt_all));
    k_peek<<<md_div_up(MD_MAXW * W, 256), 256, 0, c->stream>>>(s->d_cnt_all, 1, MD_MAXW * W, s->d_cnt_all_mapped);
    MD_CUDA(c, cudaMemsetAsync(s->d_cursor, 0, sizeof(int) * MD_MAXW, c->stream));
    MD_CUDA(c, cudaStreamSynchronize(c->stream));
    long long soff[MD_MAXW + 1], roff[MD_MAXW + 1];
    soff[0] = roff[0] = 0;
    for (int q = 0; q < W; ++q) {
        soff[q + 1] = soff[q] + (q == me ? 0 : s->h_cnt_all[me * MD_MAXW + q]);
        roff[q + 1] = roff[q] + s->h_cnt_all[q * MD_MAXW + me];
    }
    if (soff[W] > s->sendpool_cap || roff[W] > s->cap) {
        md_set_error(c, "rebuild: migration/ghost buffer capacity exceeded on this rank");
        return MDSEA_ERR_CAPACITY;
    }
    // fill pass: messages for peer q -> its segment of the send pool; what this rank keeps -> straight into its own
    // segment of the receive pool (the union is unpacked from there in one go; the sort canonicalises the order)
    RouteDst dst;
    for (int q = 0; q < MD_MAXW; ++q) dst.pool[q] = q < W ? (q == me ? s->recvpool + roff[me] : s->sendpool + soff[q]) : nullptr;
    k_mr_route<<<g256(n), 256, 0, c->stream>>>(c->d_r, c->d_v, s->id, n, s->cap, c->box.L, s->cell_len, s->nc, s->D, 1, s->d_cursor,
                                               dst);
    void  *sp[MD_MAXW], *rp[MD_MAXW];
    size_t sb[MD_MAXW], rbytes[MD_MAXW];
    for (int q = 0; q < W; ++q) {
        sp[q] = s->sendpool + soff[q];
        rp[q] = s->recvpool + roff[q];
        sb[q] = q == me ? 0 : sizeof(PMsg) * (size_t)(soff[q + 1] - soff[q]);
        rbytes[q] = q == me ? 0 : sizeof(PMsg) * (size_t)(roff[q + 1] - roff[q]);
    }
    MD_TRY(md_comm_exchange(c, sp, sb, rp, rbytes));
    if (roff[W] > 0) k_mr_unpack<<<g256(roff[W]), 256, 0, c->stream>>>(s->recvpool, roff[W], 0, s->cap, s->r_alt, s->v_alt, s->id_alt);
    c->stats.kernel_launches += 3;
    MD_CUDA(c, cudaGetLastError());
    s->n_tot = roff[W];
    return MDSEA_OK;
}

static int build_halo_lists(mdsea_ctx *c) {
    CellState *s = c->cells;
    const int W = c->p.world;
    const long long base = s->n_int, m = s->n_tot - base, tot = m * W;
    k_halo_flags<<<g256(m), 256, 0, c->stream>>>(s->cell, base, c->n, s->n_tot, s->nc, s->D, s->halo_flag);
    MD_CUDA(c, cudaMemsetAsync(s->halo_flag + tot, 0, sizeof(int), c->stream));  // sentinel: scan[tot] = grand total
    MD_TRY(scan_exclusive(c, s->halo_flag, s->halo_scan, tot + 1));
    k_halo_lists<<<g256(tot), 256, 0, c->stream>>>(s->halo_flag, s->halo_scan, m, base, W, s->halo_idx);
    c->stats.kernel_launches += 2;
    // list sizes from the scan at the segment starts (A) and at the owned/ghost split of each segment (B):
    // two strided copies (one int per segment)
    int *h = s->h_cnt_all;  // pinned scratch, >= 2 * MD_MAXW + 1 ints
    if (m > 0) {
        k_peek<<<1, 64, 0, c->stream>>>(s->halo_scan, m, W + 1, s->d_cnt_all_mapped);
        k_peek<<<1, 64, 0, c->stream>>>(s->halo_scan + (c->n - base), m, W, s->d_cnt_all_mapped + MD_MAXW + 1);
        MD_CUDA(c, cudaStreamSynchronize(c->stream));
    } else {
        for (int q = 0; q <= 2 * MD_MAXW + 1; ++q) h[q] = 0;
    }
    for (int q = 0; q < W; ++q) {
        s->halo_nsend[q] = h[MD_MAXW + 1 + q] - h[q];
        s->halo_nrecv[q] = h[q + 1] - h[MD_MAXW + 1 + q];
        s->halo_base[q] = (long long)q * m;
    }
    if (s->p2p) {
        // tell every sender where its segment starts inside my receive buffer (a store into its control block)
        HaloOff H;
        CtrlPtrs cp;
        long long ent = 0;
        for (int p = 0; p < MD_MAXW; ++p) {
            H.off_for[p] = 3 * ent + p;
            if (p < W && p != c->p.rank) ent += s->halo_nrecv[p];
            cp.p[p] = p < W ? (PeerCtrl *)((char *)s->peer_buf[p] + s->ctrl_off) : nullptr;
        }
        k_publish_halo_off<<<1, 32, 0, c->stream>>>(cp, H, W, c->p.rank, s->asm_seq);
        c->stats.kernel_launches += 1;
    }
    return MDSEA_OK;
}

// called by comm_init once the communicator exists: map every peer's receive buffer (CUDA IPC over NVLink).  Any
// failure (peer access impossible, MDSEA_HALO=nccl) leaves the ncclSend/ncclRecv path in place, on all ranks alike.
int md_cells_comm_ready(mdsea_ctx *c) {
    CellState *s = c->cells;
    if (!s || c->p.world < 2 || md_comm_backend(c) != 1) return MDSEA_OK;
    if (getenv("MDSEA_HALO") && !strcmp(getenv("MDSEA_HALO"), "nccl")) return MDSEA_OK;
    s->p2p_stride = 3 * s->cap + MD_MAXW;
    s->seg_cap = s->cap;   // worst case: everything a rank owns arrives from ONE source (first rebuild after set_state)
    s->ctrl_off = sizeof(double) * 2 * (size_t)s->p2p_stride + sizeof(unsigned long long) * (MD_MAXW + 2);
    s->ctrl_off = (s->ctrl_off + 255) / 256 * 256;
    s->inbox_off = (s->ctrl_off + sizeof(PeerCtrl) + 255) / 256 * 256;
    const size_t bytes = s->inbox_off + sizeof(PMsg) * (size_t)s->seg_cap * (size_t)c->p.world;
    MD_CUDA(c, cudaMalloc(&s->p2p_buf, bytes));
    MD_CUDA(c, cudaMemset(s->p2p_buf, 0, s->inbox_off));
    MD_CUDA(c, cudaMalloc(&s->d_roff, sizeof(int) * (MD_MAXW + 2)));
    const int st = md_comm_share_buffer(c, s->p2p_buf, (void **)s->peer_buf);
    if (st == MDSEA_ERR_UNSUPPORTED) return MDSEA_OK;   // keep NCCL send/recv
    if (st != MDSEA_OK) return st;
    s->p2p = true;
    return MDSEA_OK;
}

// per step: owned boundary positions (+ the trigger flag) -> the ranks that hold them as ghosts
static int halo_exchange(mdsea_ctx *c, cudaStream_t st) {
    CellState *s = c->cells;
    const int W = c->p.world, me = c->p.rank;
    if (st == c->stream) md_prof_begin(c);
    HaloTab ts, tr;
    ts.world = tr.world = W;
    ts.rank = tr.rank = me;
    ts.ent0[0] = tr.ent0[0] = 0;
    for (int q = 0; q < W; ++q) {
        ts.ent0[q + 1] = ts.ent0[q] + (q == me ? 0 : s->halo_nsend[q]);
        tr.ent0[q + 1] = tr.ent0[q] + (q == me ? 0 : s->halo_nrecv[q]);
        ts.idx0[q] = s->halo_base[q];
        tr.idx0[q] = s->halo_base[q] + s->halo_nsend[q];
    }
    void  *sp[MD_MAXW], *rp[MD_MAXW];
    size_t sb[MD_MAXW], rbytes[MD_MAXW];
    for (int q = 0; q < W; ++q) {
        sp[q] = s->halo_send + 3 * ts.ent0[q] + q;
        rp[q] = s->halo_recv + 3 * tr.ent0[q] + q;
        sb[q] = q == me ? 0 : sizeof(double) * (3 * (size_t)s->halo_nsend[q] + 1);
        rbytes[q] = q == me ? 0 : sizeof(double) * (3 * (size_t)s->halo_nrecv[q] + 1);
    }
    if (s->p2p) {
        const unsigned long long seq = ++s->halo_seq;
        const long long par = (long long)(seq & 1ull) * s->p2p_stride;
        PeerDst pd;
        PeerCnt pc;
        for (int q = 0; q < MD_MAXW; ++q) {
            pd.base[q] = q < W && q != me ? s->peer_buf[q] + par : nullptr;
            pc.cnt[q] = q < W && q != me ? (unsigned long long *)(s->peer_buf[q] + 2 * s->p2p_stride) + me : nullptr;
        }
        HaloWait wt;
        wt.counters = (const unsigned long long *)(s->p2p_buf + 2 * s->p2p_stride);
        wt.seq = seq;
        k_halo_push<<<g256(ts.ent0[W] + W), 256, 0, st>>>(c->d_r, s->cap, s->halo_idx, ts, c->d_sc, pd,
                                                          (const PeerCtrl *)((char *)s->p2p_buf + s->ctrl_off), s->asm_seq, s->drift_tag);
        k_halo_signal<<<1, 32, 0, st>>>(pc, W, me, seq);
        k_halo_unpack<<<g256(tr.ent0[W] + W), 256, 0, st>>>(s->p2p_buf + par, s->halo_idx, tr, s->cap, c->d_r,
                                                            c->p.precision == MDSEA_FP32, fixed_inv_h(c), s->pos4, c->d_sc, wt, s->drift_tag);
        c->stats.kernel_launches += 1;
    } else {
        HaloWait wt;
        wt.counters = nullptr;
        wt.seq = 0;
        k_halo_pack<<<g256(ts.ent0[W] + W), 256, 0, st>>>(c->d_r, s->cap, s->halo_idx, ts, c->d_sc, s->halo_send, s->drift_tag);
        MD_TRY(md_comm_exchange(c, sp, sb, rp, rbytes, st));
        k_halo_unpack<<<g256(tr.ent0[W] + W), 256, 0, st>>>(s->halo_recv, s->halo_idx, tr, s->cap, c->d_r,
                                                            c->p.precision == MDSEA_FP32, fixed_inv_h(c), s->pos4, c->d_sc, wt, s->drift_tag);
    }
    c->stats.kernel_launches += 2;
    if (st == c->stream) md_prof_end(c, &c->stats.ms_comm);
    MD_CUDA(c, cudaGetLastError());
    return MDSEA_OK;
}

#include <chrono>
static bool g_trace_rebuild = getenv("MDSEA_TRACE_REBUILD") != nullptr;
static int  g_nbw_groups = getenv("MDSEA_NBW_GROUPS") ? atoi(getenv("MDSEA_NBW_GROUPS")) : 2;   // r01: 4: 848 us, 2: 832 us, 1: 851 us per rebuild
static int  g_nbw_variant = getenv("MDSEA_NBW_VARIANT") ? atoi(getenv("MDSEA_NBW_VARIANT")) : 1;   // bit 0: expanded distance test
static bool g_force_cs = !(getenv("MDSEA_FORCE_CS") && !strcmp(getenv("MDSEA_FORCE_CS"), "0"));   // A/B switch: streaming list loads
static bool g_cut_fuse = !(getenv("MDSEA_CUT_FUSE") && !strcmp(getenv("MDSEA_CUT_FUSE"), "0"));   // A/B switch
static bool g_build_thread = getenv("MDSEA_NBR_BUILD") && !strcmp(getenv("MDSEA_NBR_BUILD"), "thread");  // A/B switch
struct PhaseTimer {  // debugging aid: wall-clock per rebuild phase (synchronises the stream; off by default)
    mdsea_ctx *c;
    std::chrono::steady_clock::time_point t0;
    explicit PhaseTimer(mdsea_ctx *cc) : c(cc) { if (g_trace_rebuild) { cudaStreamSynchronize(c->stream); t0 = std::chrono::steady_clock::now(); } }
    void lap(const char *name) {
        if (!g_trace_rebuild) return;
        cudaStreamSynchronize(c->stream);
        auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[rebuild rank %d] %-12s %8.1f us\n", c->p.rank, name, std::chrono::duration<double, std::micro>(t1 - t0).count());
        t0 = t1;
    }
};

static int rebuild(mdsea_ctx *c) {
    CellState *s = c->cells;
    const int W = c->p.world;
    md_prof_begin(c);
    PhaseTimer pt(c);
    if (W > 1) MD_TRY(mr_assemble(c));
    else s->n_tot = s->n_int = c->n;
    pt.lap("assemble");
    const long long nt = s->n_tot;
    const double *src_r = W > 1 ? s->r_alt : c->d_r, *src_v = W > 1 ? s->v_alt : c->d_v;
    const int    *src_id = W > 1 ? s->id_alt : s->id;
    double *dst_r = W > 1 ? c->d_r : s->r_alt, *dst_v = W > 1 ? c->d_v : s->v_alt;
    int    *dst_id = W > 1 ? s->id : s->id_alt;
    int launches = 0;
    MD_CUDA(c, cudaMemsetAsync(s->cell_start, 0, sizeof(int) * ((size_t)s->ncell + 1), c->stream));
    MD_CUDA(c, cudaMemsetAsync(s->cell_end, 0, sizeof(int) * ((size_t)s->ncell + 1), c->stream));
    if (!s->radix_binning) {
        // counting pass over (class, cell) buckets -> bucket starts -> rank by id inside the bucket, gather into place
        MD_CUDA(c, cudaMemsetAsync(s->bcount, 0, sizeof(int) * ((size_t)s->nbucket + 1), c->stream));
        k_bin_count<<<g256(nt), 256, 0, c->stream>>>(src_r, nt, s->cap, c->box.L, s->cell_len, s->nc, s->ncell, s->D, s->bkt, s->arrival,
                                                     s->bcount);
        MD_TRY(scan_exclusive(c, s->bcount, s->bstart, s->nbucket + 1));
        k_bin_scatter<<<g256(nt), 256, 0, c->stream>>>(nt, s->bkt, s->arrival, s->bstart, src_id, s->tmp_id, s->tmp_src, s->tmp_bkt);
        pt.lap("count+scan");
        k_bin_place<<<g256(nt), 256, 0, c->stream>>>(nt, s->cap, s->ncell, s->bstart, s->tmp_id, s->tmp_src, s->tmp_bkt, src_r, src_v, dst_r,
                                                     dst_v, s->rb, dst_id, s->cell, s->cell_start, s->cell_end,
                                                     c->p.precision == MDSEA_FP32, fixed_inv_h(c), s->pos4, s->qfix);
        launches += 3;
        if (W > 1)   // owned interior | owned boundary | ghost: the class boundaries are bucket starts
            k_peek<<<1, 32, 0, c->stream>>>(s->bstart + s->ncell, s->ncell, 2, s->d_cnt_all_mapped);
    } else {
        MD_CUDA(c, cudaMemsetAsync(s->d_cnt + MD_MAXW, 0, 2 * sizeof(int), c->stream));
        k_cell_keys<<<g256(nt), 256, 0, c->stream>>>(src_r, src_id, nt, s->cap, c->box.L, s->cell_len, s->nc, s->D, s->key, s->val,
                                                     s->d_cnt + MD_MAXW);
        // LSD passes over the id bits, then the cell bits, then (multi-rank) the boundary / ghost bits:
        // = stable sort by (ghost, boundary, cell id, particle id)
        std::vector<int> shifts;
        for (int b = 0; b < s->id_bits; b += 8) shifts.push_back(b);
        for (int b = 0; b < s->cell_bits; b += 8) shifts.push_back(32 + b);
        if (W > 1) shifts.push_back(56);
        const int nblk = md_div_up(nt, SORT_CHUNK);
        for (int sh : shifts) {
            k_sort_hist<<<nblk, SORT_THREADS, 0, c->stream>>>(s->key, nt, sh, s->hist, nblk);
            k_sort_rowscan<<<256, 256, 0, c->stream>>>(s->hist, nblk, s->hist_tot);
            k_sort_scatter<<<nblk, SORT_THREADS, 0, c->stream>>>(s->key, s->val, s->key_alt, s->val_alt, nt, sh, s->hist, nblk,
                                                                 s->hist_tot);
            std::swap(s->key, s->key_alt);
            std::swap(s->val, s->val_alt);
            launches += 3;
        }
        pt.lap("keys+sort");
        k_reorder<<<g256(nt), 256, 0, c->stream>>>(s->key, s->val, nt, s->cap, src_r, src_v, dst_r, dst_v, s->rb, dst_id, s->cell);
        k_cell_bounds<<<g256(nt), 256, 0, c->stream>>>(s->cell, nt, s->cell_start, s->cell_end);
        k_pack_positions<<<g256(nt), 256, 0, c->stream>>>(dst_r, nt, s->cap, c->p.precision == MDSEA_FP32, fixed_inv_h(c), s->pos4,
                                                          s->qfix);
        launches += 4;
        if (W > 1)   // {owned, owned interior}
            k_peek<<<1, 32, 0, c->stream>>>(s->d_cnt + MD_MAXW, 1, 2, s->d_cnt_all_mapped);
    }
    if (W == 1) {
        std::swap(c->d_r, s->r_alt);
        std::swap(c->d_v, s->v_alt);
        std::swap(s->id, s->id_alt);
    } else {
        MD_CUDA(c, cudaStreamSynchronize(c->stream));
        if (s->radix_binning) { c->n = s->h_cnt_all[0]; s->n_int = s->h_cnt_all[1]; }
        else { s->n_int = s->h_cnt_all[0]; c->n = s->h_cnt_all[1]; }
        c->stats.n_local = c->n;
        c->stats.n_ghost = nt - c->n;
    }
    pt.lap("place");
    if (W > 1) MD_TRY(build_halo_lists(c));
    pt.lap("halo lists");
    {
        const double ih = fixed_inv_h(c);
        const double rl2s = s->rl2 * ih * ih;
        const float lo = (float)(rl2s * (1.0 - 2e-5)), hi = (float)(rl2s * (1.0 + 2e-5));
        const long long no = c->n > 0 ? c->n : 1;
        if (g_build_thread || s->nc < 4)
            k_nbr_build<<<md_div_up(no, NB_THREADS), NB_THREADS, 0, c->stream>>>(
                c->d_r, s->qfix, c->n, s->cap, s->cell, s->cell_start, s->cell_end, s->nc, c->box.L, c->box.halfL, s->rl2, lo, hi,
                s->max_nbr, s->nbr, s->nbr_cnt, c->d_sc);
        else {
#define NBW_LAUNCH(EXP, G)                                                                                                          \
    k_nbr_build_warp<EXP, G><<<md_div_up(no, NBW_WARPS * 32), NBW_WARPS * 32, 0, c->stream>>>(                                     \
        c->d_r, s->qfix, c->n, s->cap, s->cell, s->cell_start, s->cell_end, s->nc, c->box.L, c->box.halfL, s->cell_len, ih, s->rl2, \
        rl2s, (float)(rl2s * 2e-5), lo, hi, s->max_nbr, std::min(NBW_RUN, s->nc - 3), s->nbr, s->nbr_cnt, c->d_sc)
            if (!(g_nbw_variant & 1)) NBW_LAUNCH(false, 4);   // MDSEA_NBW_VARIANT=0: direct distance test (A/B)
            else if (g_nbw_groups == 1) NBW_LAUNCH(true, 1);
            else if (g_nbw_groups == 2) NBW_LAUNCH(true, 2);
            else NBW_LAUNCH(true, 4);
#undef NBW_LAUNCH
        }
    }
    launches += 1;
    pt.lap("nbr build");
    MD_CUDA(c, cudaMemsetAsync(&c->d_sc->rebuild_flag, 0, sizeof(int), c->stream));
    md_prof_end(c, &c->stats.ms_rebuild);
    MD_CUDA(c, cudaGetLastError());
    c->stats.kernel_launches += launches;
    c->stats.list_rebuilds += 1;
    s->need_rebuild = false;
    return MDSEA_OK;
}

// force kernel over the owned slots [i0, i1); its per-block PE partials go to pe_part[pe_off ...]
static int launch_force(mdsea_ctx *c, long long i0, long long i1, int pe_off, int guarded /* guard tag, 0 = unguarded */) {
    CellState *s = c->cells;
    if (i1 <= i0) return MDSEA_OK;
    const int grid = force_grid(i0, i1);
    const bool lj = c->pot.fast == POTF_LJ;
    if (c->p.precision == MDSEA_FP64) {
        const PotF64 P = md_pot_f64(c->pot);
#define NBR_F64(POTK)                                                                                                              \
    do {                                                                                                                           \
        if (g_force_cs) NBR_F64_(POTK, true);                                                                                      \
        else NBR_F64_(POTK, false);                                                                                                \
    } while (0)
#define NBR_F64_(POTK, CS)                                                                                                         \
    k_nbr_force_f64<POTK, CS><<<grid, NB_THREADS, 0, c->stream>>>((const double4 *)s->pos4, i0, i1, s->cap, s->nbr, s->nbr_cnt, s->max_nbr, \
                                                              c->box.L, c->box.halfL, s->rc2, P, c->d_acc, c->d_pe_part + pe_off,   \
                                                              c->d_sc, guarded)
        if (lj) NBR_F64(POTF_LJ);
        else NBR_F64(POTF_MIE);
#undef NBR_F64
#undef NBR_F64_
    } else {
        const double inv_h = fixed_inv_h(c);
        const PotF32 P = md_pot_f32(c->pot, inv_h);
#define NBR_F32(TEX, POTK)                                                                                                         \
    do {                                                                                                                           \
        if (g_force_cs) NBR_F32_(TEX, POTK, true);                                                                                 \
        else NBR_F32_(TEX, POTK, false);                                                                                           \
    } while (0)
#define NBR_F32_(TEX, POTK, CS)                                                                                                    \
    k_nbr_force_f32<TEX, POTK, CS><<<grid, NB_THREADS, 0, c->stream>>>((const uint4 *)s->pos4, s->tex_pos, c->d_r, i0, i1, s->cap, s->nbr, \
                                                                   s->nbr_cnt, s->max_nbr, c->box.L, c->box.halfL, s->rc2,         \
                                                                   (float)(s->rc2 * inv_h * inv_h), inv_h, P, c->d_acc,            \
                                                                   c->d_pe_part + pe_off, c->d_sc, guarded)
        if (s->tex_pos) {
            if (lj) NBR_F32(true, POTF_LJ);
            else NBR_F32(true, POTF_MIE);
        } else {
            if (lj) NBR_F32(false, POTF_LJ);
            else NBR_F32(false, POTF_MIE);
        }
#undef NBR_F32
#undef NBR_F32_
    }
    MD_CUDA(c, cudaGetLastError());
    c->stats.kernel_launches += 1;
    return MDSEA_OK;
}

static int cutoff_force(mdsea_ctx *c, int guarded) {
    CellState *s = c->cells;
    if (s->need_rebuild) MD_TRY(rebuild(c));
    md_prof_begin(c);
    c->n_pe_part = force_grid(0, c->n > 0 ? c->n : 1);
    MD_TRY(launch_force(c, 0, c->n, 0, guarded));
    md_prof_end(c, &c->stats.ms_force);
    c->stats.force_evals += 1;
    return MDSEA_OK;
}

int md_cutoff_force(mdsea_ctx *c) { return cutoff_force(c, 0); }

static int cut_kick_drift(mdsea_ctx *c, bool do_bc, bool do_kick, bool do_drift) {
    CellState *s = c->cells;
    md_prof_begin(c);
    if (do_drift) s->drift_tag = s->drift_tag == 0x7fffffff ? 1 : s->drift_tag + 1;
    k_cut_kick_drift<<<g256(c->n), 256, 0, c->stream>>>(c->d_r, c->d_v, c->d_acc, s->rb, c->n, s->cap, do_bc, do_kick, do_drift,
                                                        c->box.L, c->box.halfL, c->p.dt, s->trig2, c->p.precision == MDSEA_FP32,
                                                        fixed_inv_h(c), s->pos4, c->d_sc, s->drift_tag);
    md_prof_end(c, &c->stats.ms_integrate);
    MD_CUDA(c, cudaGetLastError());
    c->stats.kernel_launches += 1;
    return MDSEA_OK;
}

// guard: tag of the drift this launch depends on (0 = unguarded).  next_tag != 0: the kernel also does wrap + kick +
// drift of the NEXT step, raising that tag if the skin trigger fires.
static int cut_finish_and_finalize(mdsea_ctx *c, const mdsea_run_args *a, int k, double g_dt, int guarded, int next_tag) {
    CellState *s = c->cells;
    const int W = c->p.world;
    const size_t row_stride = W == 1 ? (size_t)c->n_global : (size_t)s->cap;
    double *rr = a->record_coords ? c->d_r_rows + (size_t)k * 3 * row_stride : nullptr;
    double *vr = a->record_coords ? c->d_v_rows + (size_t)k * 3 * row_stride : nullptr;
    int *idr = (a->record_coords && W > 1) ? c->d_id_rows + (size_t)k * s->cap : nullptr;
    if (W > 1) s->row_n_own[k] = c->n;
    md_prof_begin(c);
    c->n_ke_part = std::min(g256(c->n), next_tag ? s->finish_blocks_fused : s->finish_blocks);
    FinalizeArgs fin = md_finalize_args(c, k, k + 1 < a->nsteps, a->k_boltzmann, c->stats.steps + k);
    fin.pe_part = s->pe_fold;   // the force kernel's partials, folded per finish block
    fin.n_pe = c->n_ke_part;
    NextDrift nd;
    nd.rb = s->rb; nd.L = c->box.L; nd.halfL = c->box.halfL; nd.trig2 = s->trig2; nd.inv_h = fixed_inv_h(c);
    nd.pos4 = s->pos4; nd.fp32 = c->p.precision == MDSEA_FP32; nd.tag = next_tag;
    // two mapped words alternate (tag parity): the host may still be reading the previous drift's word
    nd.host_flag = (W == 1 && next_tag) ? s->d_flag_mapped + 2 + (next_tag & 1) : nullptr;
    if (nd.host_flag) s->published_tag = next_tag;
#define CUT_FINISH(FUSE)                                                                                                            \
    k_cut_kick_finish<FUSE><<<c->n_ke_part, 256, 0, c->stream>>>(c->d_r, c->d_v, c->d_acc, s->id, c->n, s->cap, (long long)row_stride, idr, \
                                                                 c->p.dt, g_dt, c->d_sc, c->d_ke_part, rr, vr, guarded, fin, c->d_pe_part,   \
                                                                 c->n_pe_part, s->pe_fold, nd)
    if (next_tag) CUT_FINISH(true);
    else CUT_FINISH(false);
#undef CUT_FINISH
    md_prof_end(c, &c->stats.ms_integrate);
    MD_CUDA(c, cudaGetLastError());
    c->stats.kernel_launches += 1;
    return MDSEA_OK;
}

// After a drift the skin trigger sits in a device flag.  To keep the GPU busy while the host learns its value,
// the force / kick_finish / finalize kernels of the step are launched SPECULATIVELY, guarded by that flag (they
// return at once when it is set); the host then reads the flag and, only if it fired (every ~6 steps at 1e6
// particles), rebuilds and launches the three kernels again.  Multi-rank: the per-step halo messages carry every
// rank's flag, so after the (unconditional) ghost refresh all ranks hold the same MAX and rebuild together.
// fuse_next: the finish kernel also starts the next step (wrap + kick + drift); on return s->drift_tag is that drift's tag.
static int force_after_drift(mdsea_ctx *c, const mdsea_run_args *a, int k, double g_dt, bool fuse_next) {
    CellState *s = c->cells;
    const int W = c->p.world;
    const int tag = s->drift_tag;                                     // the drift this step's force evaluation follows
    const int next_tag = fuse_next ? (tag == 0x7fffffff ? 1 : tag + 1) : 0;
    if (s->need_rebuild) {  // no list yet (first evaluation / new state): nothing to speculate on
        MD_TRY(cutoff_force(c, 0));
        MD_TRY(cut_finish_and_finalize(c, a, k, g_dt, 0, next_tag));
        if (fuse_next) s->drift_tag = next_tag;
        return MDSEA_OK;
    }
    const long long fe0 = c->stats.force_evals;
    const int flag_slot = (W == 1 && s->published_tag == tag) ? 2 + (tag & 1) : 0;
    if (W == 1) {
        if (!flag_slot) k_publish_flag<<<1, 1, 0, c->stream>>>(c->d_sc, s->d_flag_mapped, tag);
        MD_CUDA(c, cudaEventRecord(s->ev_flag, c->stream));
        MD_TRY(cutoff_force(c, tag));
    } else if (c->profiling) {  // per-family timers need everything on one stream: no overlap
        MD_TRY(halo_exchange(c, c->stream));
        k_publish_flag<<<1, 1, 0, c->stream>>>(c->d_sc, s->d_flag_mapped, tag);
        MD_CUDA(c, cudaEventRecord(s->ev_flag, c->stream));
        MD_TRY(cutoff_force(c, tag));
    } else {
        // comm stream: ghost refresh (+ flags) while the main stream computes the INTERIOR forces, which need no ghosts
        MD_CUDA(c, cudaEventRecord(s->ev_drift, c->stream));
        MD_CUDA(c, cudaStreamWaitEvent(c->comm_stream, s->ev_drift, 0));
        MD_TRY(halo_exchange(c, c->comm_stream));
        k_publish_flag<<<1, 1, 0, c->comm_stream>>>(c->d_sc, s->d_flag_mapped, tag);
        MD_CUDA(c, cudaEventRecord(s->ev_flag, c->comm_stream));
        const int gi = force_grid(0, s->n_int > 0 ? s->n_int : 1);
        c->n_pe_part = gi + force_grid(s->n_int, c->n > s->n_int ? c->n : s->n_int + 1);
        MD_CUDA(c, cudaMemsetAsync(c->d_pe_part, 0, sizeof(double) * c->n_pe_part, c->stream));
        MD_TRY(launch_force(c, 0, s->n_int, 0, tag));                     // guarded by the LOCAL flag only (the global one
                                                                          // is not known yet): redone if any rank fired
        MD_CUDA(c, cudaStreamWaitEvent(c->stream, s->ev_flag, 0));
        MD_TRY(launch_force(c, s->n_int, c->n, gi, tag));                 // boundary particles: ghosts are fresh now
        c->stats.force_evals += 1;
    }
    MD_TRY(cut_finish_and_finalize(c, a, k, g_dt, tag, next_tag));
    MD_CUDA(c, cudaEventSynchronize(s->ev_flag));
    if (((volatile int *)s->h_flag)[flag_slot]) {  // the guarded kernels did nothing; rebuild (clears the flag) and do the step for real
        c->stats.force_evals = fe0;
        s->need_rebuild = true;
        MD_CUDA(c, cudaMemsetAsync(&c->d_sc->pairs_pending, 0, sizeof(unsigned long long), c->stream));
        MD_TRY(cutoff_force(c, 0));
        MD_TRY(cut_finish_and_finalize(c, a, k, g_dt, 0, next_tag));
    }
    if (fuse_next) s->drift_tag = next_tag;
    return MDSEA_OK;
}

int md_cutoff_steps(mdsea_ctx *c, const mdsea_run_args *a, double g_dt) {
    const int  W = c->p.world;
    const bool verlet = c->p.algorithm == MDSEA_ALGO_VERLET;
    const bool reuse = verlet && c->p.force_evals_per_step == 1;
    // world > 1: the finish kernel's bookkeeping only sees this rank's share of mean_ke.  _quench needs the GLOBAL
    // value of the previous step (simulator.py:192-195, 233-239), so a batch that carries quench targets all-reduces
    // one double per step and redoes the temperature / scale bookkeeping from it; every batch leaves the global
    // mean_ke of its last step in the scalars for the first step of the next batch.
    const bool mr_quench = W > 1 && a->quench_targets;
    // reuse mode: the finish kernel of step k also does wrap + kick + drift of step k + 1 (never across a batch
    // boundary: the caller may read or replace the state between batches).  MDSEA_CUT_FUSE=0 keeps the separate kernels.
    const bool fuse = reuse && g_cut_fuse;
    bool drifted = false;   // the previous finish kernel already started this step
    for (int k = 0; k < a->nsteps; ++k) {
        if (verlet) {
            if (reuse && c->acc_valid) {
                if (!drifted) MD_TRY(cut_kick_drift(c, true, true, true));
            } else {
                MD_TRY(cut_kick_drift(c, true, false, false));   // apply_bc
                MD_TRY(md_cutoff_force(c));                       // simulator.py:540
                MD_TRY(md_commit_pairs(c));                       // counted for good: not part of the speculation below
                MD_TRY(cut_kick_drift(c, false, true, true));
            }
        } else {
            MD_TRY(cut_kick_drift(c, true, false, true));        // simulator.py:533
        }
        drifted = fuse && k + 1 < a->nsteps;
        MD_TRY(force_after_drift(c, a, k, g_dt, drifted));        // simulator.py:544 + kick + energies + row (+ next drift)
        c->acc_valid = true;
        if (mr_quench || (W > 1 && k == a->nsteps - 1)) {
            CellState *s = c->cells;
            MD_CUDA(c, cudaMemcpyAsync(s->ke_glob, c->d_ke_rows + k, sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
            MD_TRY(md_comm_allreduce_sum_f64(c, s->ke_glob, 1));
            k_mr_quench_fix<<<1, 1, 0, c->stream>>>(c->d_sc, s->ke_glob, (mr_quench && k + 1 < a->nsteps) ? c->d_quench + (size_t)(k + 1) * 2 : nullptr,
                                                    a->k_boltzmann, mr_quench ? 1 : 0);
            c->stats.kernel_launches += 1;
        }
    }
    if (W > 1) {  // global energies: one small allreduce per batch (the rows hold this rank's share of the means)
        MD_TRY(md_comm_allreduce_sum_f64(c, c->d_pe_rows, a->nsteps));
        MD_TRY(md_comm_allreduce_sum_f64(c, c->d_ke_rows, a->nsteps));
    }
    return MDSEA_OK;
}

// world > 1: compact rows (slot order) + ids -> caller's buffers [nsteps][3][cap] / [nsteps][cap]
int md_cutoff_copy_rows(mdsea_ctx *c, const mdsea_run_args *a, const mdsea_rows *rows) {
    CellState *s = c->cells;
    if (!a->record_coords || !rows->r || !rows->v) return MDSEA_OK;
    if (!rows->ids || !rows->n_owned || rows->row_stride < s->cap) {
        md_set_error(c, "run: world > 1 needs rows.ids, rows.n_owned and rows.row_stride >= the rank capacity");
        return MDSEA_ERR_BAD_ARG;
    }
    const size_t st = (size_t)rows->row_stride;
    for (int k = 0; k < a->nsteps; ++k) {
        const size_t n = (size_t)s->row_n_own[k];
        rows->n_owned[k] = (int64_t)n;
        for (int d = 0; d < 3; ++d) {
            MD_CUDA(c, cudaMemcpyAsync(rows->r + ((size_t)k * 3 + d) * st, c->d_r_rows + ((size_t)k * 3 + d) * s->cap, sizeof(double) * n,
                                       cudaMemcpyDeviceToHost, c->copy_stream));
            MD_CUDA(c, cudaMemcpyAsync(rows->v + ((size_t)k * 3 + d) * st, c->d_v_rows + ((size_t)k * 3 + d) * s->cap, sizeof(double) * n,
                                       cudaMemcpyDeviceToHost, c->copy_stream));
        }
        MD_CUDA(c, cudaMemcpyAsync(rows->ids + (size_t)k * st, c->d_id_rows + (size_t)k * s->cap, sizeof(int) * n, cudaMemcpyDeviceToHost,
                                   c->copy_stream));
    }
    return MDSEA_OK;
}

int md_cutoff_finish_rows(mdsea_ctx *c) { return MDSEA_OK; }
long long md_cutoff_capacity(mdsea_ctx *c) { return c->cells ? c->cells->cap : 0; }

int md_cutoff_get_state(mdsea_ctx *c, double *r, double *v, double *a, int64_t *ids, int64_t *n_owned) {
    CellState *s = c->cells;
    const long long n = c->n, N = c->n_global;
    if (!s->scratch) MD_CUDA(c, cudaMalloc(&s->scratch, sizeof(double) * 3 * (size_t)N));
    const double *src[3] = {c->d_r, c->d_v, c->d_acc};
    double *dst[3] = {r, v, a};
    std::vector<int> h(n > 0 ? n : 1);
    MD_CUDA(c, cudaStreamSynchronize(c->stream));
    MD_CUDA(c, cudaMemcpy(h.data(), s->id, sizeof(int) * n, cudaMemcpyDeviceToHost));
    for (int q = 0; q < 3; ++q) {
        if (!dst[q]) continue;
        k_unsort<<<g256(n), 256, 0, c->stream>>>(src[q], s->id, n, s->cap, N, s->scratch);
        c->stats.kernel_launches += 1;
        if (c->p.world == 1) {
            MD_CUDA(c, cudaMemcpyAsync(dst[q], s->scratch, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, c->stream));
            MD_CUDA(c, cudaStreamSynchronize(c->stream));
        } else {  // only the owned entries are defined: copy through a host bounce and scatter those
            std::vector<double> tmp((size_t)3 * N);
            MD_CUDA(c, cudaMemcpyAsync(tmp.data(), s->scratch, sizeof(double) * 3 * N, cudaMemcpyDeviceToHost, c->stream));
            MD_CUDA(c, cudaStreamSynchronize(c->stream));
            for (long long k = 0; k < n; ++k)
                for (int d = 0; d < 3; ++d) dst[q][(size_t)d * N + h[k]] = tmp[(size_t)d * N + h[k]];
        }
    }
    if (ids) {
        for (long long i = 0; i < n; ++i) ids[i] = h[i];
        std::sort(ids, ids + n);
    }
    if (n_owned) *n_owned = n;
    return MDSEA_OK;
}

// ---------------------------------------------------------------------------------------------------
// test hooks (bit-exact contracts).  With world > 1 each rank reports its OWNED particles: cell_of / offsets
// entries of particles owned elsewhere are left at -1 / zero-length, `order` lists the owned ids in slot order.
// ---------------------------------------------------------------------------------------------------
extern "C" int mdsea_b200_get_cells(mdsea_ctx *c, int32_t *cell_of, int32_t *order, int32_t ncell_dim[3]) {
    if (!c || !cell_of || !order) return MDSEA_ERR_BAD_ARG;
    if (!c->cutoff_path || !c->have_state) { md_set_error(c, "get_cells: needs the cutoff path and a state"); return MDSEA_ERR_STATE; }
    MD_CUDA(c, cudaSetDevice(c->p.device));
    CellState *s = c->cells;
    if (s->need_rebuild) MD_TRY(rebuild(c));
    const long long n = c->n;
    std::vector<int> id(n > 0 ? n : 1), cell(n > 0 ? n : 1);
    MD_CUDA(c, cudaStreamSynchronize(c->stream));
    MD_CUDA(c, cudaMemcpy(id.data(), s->id, sizeof(int) * n, cudaMemcpyDeviceToHost));
    MD_CUDA(c, cudaMemcpy(cell.data(), s->cell, sizeof(int) * n, cudaMemcpyDeviceToHost));
    for (long long g = 0; g < c->n_global; ++g) { cell_of[g] = -1; order[g] = -1; }
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
    const long long n = c->n, nt = s->n_tot;
    MD_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<int> id(nt > 0 ? nt : 1), cnt(n > 0 ? n : 1);
    MD_CUDA(c, cudaMemcpy(id.data(), s->id, sizeof(int) * nt, cudaMemcpyDeviceToHost));
    MD_CUDA(c, cudaMemcpy(cnt.data(), s->nbr_cnt, sizeof(int) * n, cudaMemcpyDeviceToHost));
    std::vector<int64_t> count_by_id(c->n_global, 0);
    for (long long k = 0; k < n; ++k) count_by_id[id[k]] = cnt[k];
    offsets[0] = 0;
    for (long long g = 0; g < c->n_global; ++g) offsets[g + 1] = offsets[g] + count_by_id[g];
    *n_entries = offsets[c->n_global];
    if (!idx) return MDSEA_OK;
    const size_t tiles = (size_t)((n + 31) / 32);
    std::vector<int> nbr(tiles * (size_t)s->max_nbr * 32);
    MD_CUDA(c, cudaMemcpy(nbr.data(), s->nbr, sizeof(int) * nbr.size(), cudaMemcpyDeviceToHost));
    for (long long k = 0; k < n; ++k) {
        int32_t *row = idx + offsets[id[k]];
        const size_t base = md_nbr_row_base(k, s->max_nbr);
        for (int m = 0; m < cnt[k]; ++m) row[m] = id[nbr[base + md_nbr_off((unsigned)m)]];
        std::sort(row, row + cnt[k]);
    }
    return MDSEA_OK;
}
