// K6: domain decomposition kernels -- rebuild-time routing (NCCL messages or peer-memory stores) and the per-step halo refresh.  Included by cells.cu only.
#pragma once

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
