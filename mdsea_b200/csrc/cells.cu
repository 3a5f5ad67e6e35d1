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
//
// One translation unit; the kernels live in headers included below, this file keeps the shared state (Decomp,
// CellState), the canonical geometry helpers and the host side (set-up, rebuild, step loop, test hooks):
//   cells_binning.cuh    K2  counting-pass binning, radix passes (A/B), reorder, cell bounds, packed copies
//   cells_nbr_build.cuh  K3  list layout, thread-per-particle and warp-cooperative list build
//   cells_force.cuh      K4  list force / energy kernels (fp64, fixed-point fp32)
//   cells_integrate.cuh  K5  drift, finish (+ fused next drift), quench bookkeeping for world > 1
//   cells_scan.cuh           exclusive scan over ints
//   cells_decomp.cuh     K6  rebuild-time routing and per-step halo refresh
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
    cudaEvent_t ev_bnd = nullptr;         // comm stream -> main stream: the boundary particles' force kernel has run
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
    // tile path (cells_tile.cuh): fp32 pair math, windows in shared memory, 16-bit window offsets as list entries
    bool    tile_ok = false;              // available: fp32 mode, >= 4 cells per dimension, not switched off (MDSEA_TILE=0)
    bool    tile_active = false;          // the CURRENT list generation has tile lists for the owned slots [0, tile_n)
    long long tile_n = 0;                 // all of them unless the process grid splits x: then the whole blocks of INTERIOR slots (the
                                          // boundary class -- faces that cut the x-rows, many short runs per block -- keeps the per-particle lists)
    bool    tile_off = false;             // a block's window did not fit: per-particle kernels until the next set_state
    struct TileDesc *tl_desc = nullptr;   // [ceil(cap / 128)]
    uint4  *tl_list = nullptr;            // [cap / 32][tl_max_nbr / 8][32] vectors of 8 16-bit window offsets
    int     tl_max_nbr = 0;
    int    *tl_stats = nullptr;           // device: fallback flag, max window, max ranges of the last build
    int     tl_win = 0;                   // largest window of the current list generation (entries): shared memory of the force launch
    long long tl_fallbacks = 0;
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

#include "cells_binning.cuh"
#include "cells_nbr_build.cuh"
#include "cells_force.cuh"
#include "cells_tile.cuh"
#include "cells_integrate.cuh"
#include "cells_scan.cuh"
#include "cells_decomp.cuh"

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
// tile path: the blocks of 128 absolute slots that cover [i0, i1) (never more than force_grid(i0, i1) + 1)
static inline int tile_grid(long long i0, long long i1) { return (int)((i1 + TL_BLOCK - 1) / TL_BLOCK - i0 / TL_BLOCK); }

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
    // c->box.L is the PERIOD of the cell grid: the box edge when periodic, the padded period L + 2 rlist for hard walls (capi.cu)
    s->nc = (int)floor(c->box.L / s->rlist);
    s->cell_len = c->box.L / s->nc;
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
    MD_CUDA(c, cudaMalloc(&s->nbr_cnt, sizeof(int) * s->cap));
    // Tile path (fp32 pair math, >= 4 cells per dimension like the warp-cooperative build): windows in shared memory and
    // 16-bit list entries.  The 32-bit per-particle list is then only allocated if a list generation ever falls back to it.
    s->tile_ok = p.precision == MDSEA_FP32 && s->nc >= 4 && !(getenv("MDSEA_TILE") && !strcmp(getenv("MDSEA_TILE"), "0"));
    if (s->tile_ok) {
        if (p.max_neighbors > 0) s->tl_max_nbr = (p.max_neighbors + 7) / 8 * 8;
        else {
            const double rho = (double)N / (p.boxlen * p.boxlen * p.boxlen);
            const double expect = 4.0 / 3.0 * 3.14159265358979 * s->rlist * s->rlist * s->rlist * rho;
            s->tl_max_nbr = ((int)(expect * 1.3) + 24 + 7) / 8 * 8;
        }
        if (s->tl_max_nbr > 640) s->tl_max_nbr = 640;   // 160 KB of list rows in shared memory per build block
        const size_t nblk = (size_t)((s->cap + TL_BLOCK - 1) / TL_BLOCK);
        MD_CUDA(c, cudaMalloc(&s->tl_desc, sizeof(TileDesc) * nblk));
        MD_CUDA(c, cudaMalloc(&s->tl_list, sizeof(uint4) * (size_t)(s->tl_max_nbr / 8) * (size_t)s->n_pad));
        MD_CUDA(c, cudaMalloc(&s->tl_stats, sizeof(int) * 4));
        MD_CUDA(c, cudaFuncSetAttribute(k_tile_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (s->tl_max_nbr + 1) * 2 * TL_BLOCK));
        MD_CUDA(c, cudaFuncSetAttribute(k_tile_force_f32<POTF_LJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (TL_WIN_MAX + 1) * 16));
        MD_CUDA(c, cudaFuncSetAttribute(k_tile_force_f32<POTF_MIE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (TL_WIN_MAX + 1) * 16));
    } else {
        MD_CUDA(c, cudaMalloc(&s->nbr, sizeof(int) * (size_t)s->max_nbr * (size_t)s->n_pad));
    }
    MD_CUDA(c, cudaMalloc(&s->d_cnt, sizeof(int) * (MD_MAXW + 8)));
    MD_CUDA(c, cudaHostAlloc(&s->h_flag, sizeof(int) * 4, cudaHostAllocMapped));
    MD_CUDA(c, cudaHostGetDevicePointer((void **)&s->d_flag_mapped, s->h_flag, 0));
    MD_CUDA(c, cudaEventCreateWithFlags(&s->ev_drift, cudaEventDisableTiming));
    MD_CUDA(c, cudaEventCreateWithFlags(&s->ev_flag, cudaEventDisableTiming));
    MD_CUDA(c, cudaEventCreateWithFlags(&s->ev_bnd, cudaEventDisableTiming));
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
    c->n_pe_part = force_grid(0, s->cap) + 4;
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
    cudaFree(s->tl_desc); cudaFree(s->tl_list); cudaFree(s->tl_stats);
    if (s->h_flag) cudaFreeHost(s->h_flag);
    if (s->ev_flag) cudaEventDestroy(s->ev_flag);
    if (s->ev_bnd) cudaEventDestroy(s->ev_bnd);
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
    s->tile_off = false;
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
    // The capacity decision is COLLECTIVE: every rank holds every rank's counts after the all-gather and all pools have
    // the same size, so all ranks evaluate all ranks' totals and leave together -- a rank that returned alone would leave
    // its peers blocked in ncclSend / ncclRecv.
    for (int p = 0; p < W; ++p) {
        long long sp_tot = 0, rp_tot = 0;
        for (int q = 0; q < W; ++q) {
            if (q != p) sp_tot += s->h_cnt_all[p * MD_MAXW + q];
            rp_tot += s->h_cnt_all[q * MD_MAXW + p];
        }
        if (sp_tot > s->sendpool_cap || rp_tot > s->cap) {
            md_set_error(c, p == me ? "rebuild: migration/ghost buffer capacity exceeded on this rank"
                                    : "rebuild: migration/ghost buffer capacity exceeded on another rank (all ranks stop)");
            return MDSEA_ERR_CAPACITY;
        }
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
    if (!s || c->p.world < 2) return MDSEA_OK;
    c->stats.halo_transport = md_comm_backend(c) == 1 ? 2 : 3;
    if (md_comm_backend(c) != 1) return MDSEA_OK;
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
    if (st == MDSEA_ERR_UNSUPPORTED) {   // keep NCCL send/recv -- never silently: the peer-memory path is the fast one
        if (c->p.rank == 0)
            fprintf(stderr, "mdsea_b200: a rank could not map a peer's buffer (CUDA IPC); the halo and rebuild traffic falls back to "
                            "ncclSend/ncclRecv on all ranks\n");
        return MDSEA_OK;
    }
    if (st != MDSEA_OK) return st;
    s->p2p = true;
    c->stats.halo_transport = 1;
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
static bool g_bnd_overlap = !(getenv("MDSEA_BND_OVERLAP") && !strcmp(getenv("MDSEA_BND_OVERLAP"), "0"));   // A/B switch
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

// per-particle 32-bit Verlet lists (k_nbr_build_warp / k_nbr_build): fp64 mode, < 4 cells per dimension, MDSEA_TILE=0, and
// list generations whose tile windows did not fit.  The list itself is allocated on first use.
static int legacy_list_build(mdsea_ctx *c, long long i_first) {
    CellState *s = c->cells;
    if (i_first >= c->n && c->n > 0) return MDSEA_OK;
    if (!s->nbr) MD_CUDA(c, cudaMalloc(&s->nbr, sizeof(int) * (size_t)s->max_nbr * (size_t)s->n_pad));
    {
        const double ih = fixed_inv_h(c);
        const double rl2s = s->rl2 * ih * ih;
        const float lo = (float)(rl2s * (1.0 - 2e-5)), hi = (float)(rl2s * (1.0 + 2e-5));
        const long long no = c->n > i_first ? c->n - i_first : 1;
        if (g_build_thread || s->nc < 4)
            k_nbr_build<<<md_div_up(no, NB_THREADS), NB_THREADS, 0, c->stream>>>(
                c->d_r, s->qfix, i_first, c->n, s->cap, s->cell, s->cell_start, s->cell_end, s->nc, c->box.L, c->box.halfL, s->rl2, lo, hi,
                s->max_nbr, s->nbr, s->nbr_cnt, c->d_sc);
        else {
#define NBW_LAUNCH(EXP, G)                                                                                                          \
    k_nbr_build_warp<EXP, G><<<md_div_up(no, NBW_WARPS * 32), NBW_WARPS * 32, 0, c->stream>>>(                                     \
        c->d_r, s->qfix, i_first, c->n, s->cap, s->cell, s->cell_start, s->cell_end, s->nc, c->box.L, c->box.halfL, s->cell_len, ih, s->rl2, \
        rl2s, (float)(rl2s * 2e-5), lo, hi, s->max_nbr, std::min(NBW_RUN, s->nc - 3), s->nbr, s->nbr_cnt, c->d_sc)
            if (!(g_nbw_variant & 1)) NBW_LAUNCH(false, 4);   // MDSEA_NBW_VARIANT=0: direct distance test (A/B)
            else if (g_nbw_groups == 1) NBW_LAUNCH(true, 1);
            else if (g_nbw_groups == 2) NBW_LAUNCH(true, 2);
            else NBW_LAUNCH(true, 4);
#undef NBW_LAUNCH
        }
    }
    c->stats.kernel_launches += 1;
    return MDSEA_OK;
}

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
    s->tile_active = false;
    s->tile_n = 0;
    // world > 1: when x is not split (the default grids) every owned slot gets a tile list -- the domain faces are whole
    // x-rows of cells; faces that cut the rows (p[0] > 1) leave many short runs per block, so there only the whole blocks of
    // interior slots are tiled and the rest keeps the per-particle lists
    const long long tile_target = !(s->tile_ok && !s->tile_off) ? 0
                                  : ((W == 1 || s->D.p[0] == 1) ? c->n : (s->n_int / TL_BLOCK) * TL_BLOCK);
    if (tile_target > 0) {
        // tile lists + windows; the build reports (mapped memory, one host sync per rebuild) whether every block's window fit
        const double ih = fixed_inv_h(c);
        const double rl2s = s->rl2 * ih * ih;
        MD_CUDA(c, cudaMemsetAsync(s->tl_stats, 0, sizeof(int) * 4, c->stream));
        k_tile_build<<<md_div_up(tile_target, TL_BLOCK), TL_BLOCK, (size_t)(s->tl_max_nbr + 1) * 2 * TL_BLOCK, c->stream>>>(
            c->d_r, s->qfix, tile_target, s->cap, s->cell, s->cell_start, s->cell_end, s->nc, c->box.L, c->box.halfL, s->cell_len, ih, s->rl2,
            rl2s, (float)(rl2s * 2e-5), s->tl_max_nbr, std::min(NBW_RUN, s->nc - 3), s->tl_desc, s->tl_list, s->nbr_cnt, c->d_sc,
            s->tl_stats);
        k_peek<<<1, 32, 0, c->stream>>>(s->tl_stats, 1, 3, s->d_cnt_all_mapped + MD_MAXW * MD_MAXW + 4);
        launches += 2;
        MD_CUDA(c, cudaStreamSynchronize(c->stream));
        const volatile int *h = s->h_cnt_all + MD_MAXW * MD_MAXW + 4;
        if (h[0]) {
            s->tile_off = true;
            s->tl_fallbacks += 1;
        } else {
            s->tile_active = true;
            s->tile_n = tile_target;
            s->tl_win = h[1];
        }
    }
    MD_TRY(legacy_list_build(c, s->tile_n));   // the owned slots the tile lists do not cover (all of them after a fallback)
    c->stats.tile_lists = s->tile_active ? 1 : 0;
    c->stats.tile_window_max = s->tile_active ? s->tl_win : 0;
    c->stats.tile_fallbacks = s->tl_fallbacks;
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
// number of PE partials (= blocks) of the force launches over the owned slots [i0, i1): tile blocks for the part below
// tile_n, per-particle blocks for the rest
static inline int force_parts(const CellState *s, long long i0, long long i1) {
    if (i1 <= i0) return 0;
    const long long t1 = s->tile_active ? std::min(i1, s->tile_n) : i0;
    int g = 0;
    if (t1 > i0) g += tile_grid(i0, t1);
    const long long l0 = std::max(i0, s->tile_active ? s->tile_n : i0);
    if (i1 > l0) g += force_grid(l0, i1);
    return g;
}

static int launch_force(mdsea_ctx *c, long long i0, long long i1, int pe_off, int guarded /* guard tag, 0 = unguarded */,
                        cudaStream_t st = nullptr /* default: the compute stream */) {
    CellState *s = c->cells;
    if (!st) st = c->stream;
    if (i1 <= i0) return MDSEA_OK;
    const bool lj = c->pot.fast == POTF_LJ;
    if (s->tile_active && i0 < s->tile_n) {
        const long long t1 = std::min(i1, s->tile_n);
        const double inv_h = fixed_inv_h(c);
        const PotF32 P = md_pot_f32(c->pot, inv_h);
        const float  rc2s = (float)(s->rc2 * inv_h * inv_h);
        const size_t smem = (size_t)((s->tl_win + 63) / 64 * 64) * 16;
#define TILE_F32(POTK)                                                                                                              \
    k_tile_force_f32<POTK><<<tile_grid(i0, t1), TL_BLOCK, smem, st>>>((const uint4 *)s->pos4, c->d_r, s->tl_desc, s->tl_list, s->nbr_cnt, \
                                                                         s->tl_max_nbr / 8, i0 / TL_BLOCK, i0, t1, s->cap, c->box.L, c->box.halfL, \
                                                                         s->rc2, rc2s, rc2s * 2e-6f, inv_h, P, c->d_acc, c->d_pe_part + pe_off,    \
                                                                         c->d_sc, guarded)
        if (lj) TILE_F32(POTF_LJ);
        else TILE_F32(POTF_MIE);
#undef TILE_F32
        MD_CUDA(c, cudaGetLastError());
        c->stats.kernel_launches += 1;
        if (i1 <= t1) return MDSEA_OK;
        pe_off += tile_grid(i0, t1);
        i0 = t1;
    }
    const int grid = force_grid(i0, i1);
    if (c->p.precision == MDSEA_FP64) {
        const PotF64 P = md_pot_f64(c->pot);
#define NBR_F64(POTK)                                                                                                              \
    do {                                                                                                                           \
        if (g_force_cs) NBR_F64_(POTK, true);                                                                                      \
        else NBR_F64_(POTK, false);                                                                                                \
    } while (0)
#define NBR_F64_(POTK, CS)                                                                                                         \
    k_nbr_force_f64<POTK, CS><<<grid, NB_THREADS, 0, st>>>((const double4 *)s->pos4, i0, i1, s->cap, s->nbr, s->nbr_cnt, s->max_nbr, \
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
    k_nbr_force_f32<TEX, POTK, CS><<<grid, NB_THREADS, 0, st>>>((const uint4 *)s->pos4, s->tex_pos, c->d_r, i0, i1, s->cap, s->nbr, \
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
    c->n_pe_part = force_parts(s, 0, c->n > 0 ? c->n : 1);
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
    WallBC wall;
    wall.L = c->p.boxlen; wall.radius = c->p.radius; wall.restitution = c->p.restitution;
    k_cut_kick_drift<<<g256(c->n), 256, 0, c->stream>>>(c->d_r, c->d_v, c->d_acc, s->rb, c->n, s->cap, do_bc ? (c->box.pbc ? 1 : 2) : 0,
                                                        do_kick, do_drift, c->box.L, c->box.halfL, c->p.dt, s->trig2,
                                                        c->p.precision == MDSEA_FP32, fixed_inv_h(c), s->pos4, c->d_sc, s->drift_tag, wall);
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
        // The boundary particles' force kernel follows the ghost refresh ON THE COMM STREAM, so it runs beside the interior
        // kernel instead of after it (a small launch whose ramp-up and tail would otherwise sit on the critical path); the
        // compute stream joins before the finish kernel.
        const long long ni1 = s->n_int > 0 ? s->n_int : 1, nb1 = c->n > s->n_int ? c->n : s->n_int + 1;
        const int gi = force_parts(s, 0, ni1);
        c->n_pe_part = gi + force_parts(s, s->n_int, nb1);
        MD_CUDA(c, cudaMemsetAsync(c->d_pe_part, 0, sizeof(double) * c->n_pe_part, c->stream));
        MD_CUDA(c, cudaEventRecord(s->ev_drift, c->stream));
        MD_CUDA(c, cudaStreamWaitEvent(c->comm_stream, s->ev_drift, 0));
        MD_TRY(halo_exchange(c, c->comm_stream));
        k_publish_flag<<<1, 1, 0, c->comm_stream>>>(c->d_sc, s->d_flag_mapped, tag);
        MD_CUDA(c, cudaEventRecord(s->ev_flag, c->comm_stream));
        MD_TRY(launch_force(c, 0, s->n_int, 0, tag));                     // guarded by the LOCAL flag only (the global one
                                                                          // is not known yet): redone if any rank fired
        if (g_bnd_overlap) {
            MD_TRY(launch_force(c, s->n_int, c->n, gi, tag, c->comm_stream));   // boundary particles: ghosts are fresh there
            MD_CUDA(c, cudaEventRecord(s->ev_bnd, c->comm_stream));
            MD_CUDA(c, cudaStreamWaitEvent(c->stream, s->ev_bnd, 0));
        } else {
            MD_CUDA(c, cudaStreamWaitEvent(c->stream, s->ev_flag, 0));
            MD_TRY(launch_force(c, s->n_int, c->n, gi, tag));                   // boundary particles: ghosts are fresh now
        }
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

// apply_bc (simulator.py:156-179) on the slots this rank holds; positions moved, so the next force evaluation rebuilds
int md_cutoff_apply_bc(mdsea_ctx *c) {
    CellState *s = c->cells;
    for (int d = 0; d < 3; ++d) MD_TRY(md_apply_bc_range(c, c->d_r + (size_t)d * s->cap, c->d_v + (size_t)d * s->cap, (size_t)c->n));
    s->need_rebuild = true;
    return MDSEA_OK;
}

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
    const long long n_tile = s->tile_active ? s->tile_n : 0;
    if (n_tile > 0) {
        // the lists the tile force kernel walks: 16-bit window offsets -> window entry -> slot (through the block's ranges) -> id
        const size_t nblk = (size_t)((n_tile + TL_BLOCK - 1) / TL_BLOCK), max8 = (size_t)(s->tl_max_nbr / 8);
        const size_t tiles32 = (size_t)((n_tile + 31) / 32);
        std::vector<TileDesc> desc(nblk > 0 ? nblk : 1);
        std::vector<unsigned short> lst(tiles32 * max8 * 32 * 8 + 8);
        MD_CUDA(c, cudaMemcpy(desc.data(), s->tl_desc, sizeof(TileDesc) * nblk, cudaMemcpyDeviceToHost));
        MD_CUDA(c, cudaMemcpy(lst.data(), s->tl_list, sizeof(uint4) * tiles32 * max8 * 32, cudaMemcpyDeviceToHost));
        std::vector<int> slot_of;
        for (size_t b = 0; b < nblk; ++b) {
            const TileDesc &D = desc[b];
            slot_of.assign((size_t)D.n_win, -1);
            for (int k = 0; k < D.n_rng; ++k) {
                const int lb = D.rng_lb[k], len = (k + 1 < D.n_rng ? D.rng_lb[k + 1] : D.n_win) - lb;
                for (int e = 0; e < len; ++e) slot_of[(size_t)lb + e] = D.rng_slot[k] + e;
            }
            for (long long k = (long long)b * TL_BLOCK; k < std::min<long long>(n_tile, (long long)(b + 1) * TL_BLOCK); ++k) {
                int32_t *row = idx + offsets[id[k]];
                const size_t tile = (size_t)(k >> 5), lane = (size_t)(k & 31);
                for (int m = 0; m < cnt[k]; ++m) {
                    const unsigned off = lst[((tile * max8 + (size_t)(m >> 3)) * 32 + lane) * 8 + (size_t)(m & 7)];
                    const int sl = (off >> 4) < (unsigned)D.n_win ? slot_of[off >> 4] : -1;
                    if (sl < 0) { md_set_error(c, "get_neighbors: tile list entry outside its window"); return MDSEA_ERR_STATE; }
                    row[m] = id[sl];
                }
                std::sort(row, row + cnt[k]);
            }
        }
        if (n_tile >= n) return MDSEA_OK;
    }
    const size_t tiles = (size_t)((n + 31) / 32);
    std::vector<int> nbr(tiles * (size_t)s->max_nbr * 32);
    MD_CUDA(c, cudaMemcpy(nbr.data(), s->nbr, sizeof(int) * nbr.size(), cudaMemcpyDeviceToHost));
    for (long long k = n_tile; k < n; ++k) {
        int32_t *row = idx + offsets[id[k]];
        const size_t base = md_nbr_row_base(k, s->max_nbr);
        for (int m = 0; m < cnt[k]; ++m) row[m] = id[nbr[base + md_nbr_off((unsigned)m)]];
        std::sort(row, row + cnt[k]);
    }
    return MDSEA_OK;
}

extern "C" int mdsea_b200_get_tile_windows(mdsea_ctx *c, int64_t *n_blocks, int32_t *n_win, int32_t *n_rng) {
    if (!c || !n_blocks) return MDSEA_ERR_BAD_ARG;
    if (!c->cutoff_path || !c->have_state) { md_set_error(c, "get_tile_windows: needs the cutoff path and a state"); return MDSEA_ERR_STATE; }
    MD_CUDA(c, cudaSetDevice(c->p.device));
    CellState *s = c->cells;
    if (s->need_rebuild) MD_TRY(rebuild(c));
    const size_t nblk = s->tile_active ? (size_t)((s->tile_n + TL_BLOCK - 1) / TL_BLOCK) : 0;
    *n_blocks = (int64_t)nblk;
    if (!nblk || !n_win || !n_rng) return MDSEA_OK;
    std::vector<TileDesc> desc(nblk);
    MD_CUDA(c, cudaStreamSynchronize(c->stream));
    MD_CUDA(c, cudaMemcpy(desc.data(), s->tl_desc, sizeof(TileDesc) * nblk, cudaMemcpyDeviceToHost));
    for (size_t b = 0; b < nblk; ++b) { n_win[b] = desc[b].n_win; n_rng[b] = desc[b].n_rng; }
    return MDSEA_OK;
}
