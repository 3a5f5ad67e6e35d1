// K2: cell binning of the cut-off path (counting pass + rank-by-id placement; 8-bit LSD radix passes kept for A/B), reorder, cell bounds, packed position copies.  Included by cells.cu only (single translation unit).
#pragma once

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
