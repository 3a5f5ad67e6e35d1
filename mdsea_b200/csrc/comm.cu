// Rank-to-rank plumbing of the spatial domain decomposition (cut-off path only).
//
// Two transports behind the same four primitives:
//   NCCL   one process per GPU (torchrun); ncclSend/ncclRecv grouped per exchange, ncclAllReduce /
//          ncclAllGather for the small control collectives; everything is enqueued on the context's stream.
//   LOCAL  several contexts inside ONE process, each driven by its own host thread, exchanging with plain
//          device-to-device copies behind host barriers.  Used to exercise the multi-rank code on a single
//          GPU (tests) -- no kernel ever waits on another kernel, only host threads wait on each other.
// The reference has no distributed code at all; this file is new capability.
#include <dlfcn.h>
#include <nccl.h>
#include <pthread.h>
#include <string.h>

#include <map>
#include <mutex>

#include "ctx.h"
#include "comm.h"

int md_cells_comm_ready(mdsea_ctx *c);  // cells.cu: peer-memory halo set-up once the communicator exists

struct LocalGroup {
    int               world = 0;
    pthread_barrier_t bar;
    const void       *ptr[MD_MAXW][MD_MAXW];  // ptr[src][dst]: what src offers to dst in the current exchange
    size_t            bytes[MD_MAXW][MD_MAXW];
    const void       *flat[MD_MAXW];
    int               refs = 0;
};

struct CommState {
    int         backend = 0;  // 1 NCCL, 2 LOCAL
    int         rank = 0, world = 1;
    ncclComm_t  nccl = nullptr;
    LocalGroup *lg = nullptr;
    long long   key = 0;
    double     *h_tmp = nullptr;  // pinned scratch for the LOCAL reductions
    size_t      h_tmp_bytes = 0;
    std::vector<void *> ipc_mapped;  // peer buffers mapped with cudaIpcOpenMemHandle
};

// NCCL is bound at run time (dlopen) instead of at link time: a process that also imports PyTorch must end up
// with ONE libnccl.so.2 -- PyTorch's bundled copy if it is already loaded, the system copy otherwise -- and a
// single-GPU user of this library should not load NCCL at all.
struct NcclApi {
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    bool ok = false;
};
static NcclApi g_nccl;
static bool nccl_load() {
    if (g_nccl.ok) return true;
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);  // already in the process (PyTorch)?
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return false;
#define LD(name) *(void **)(&g_nccl.name) = dlsym(h, "nccl" #name); if (!g_nccl.name) return false;
    LD(GetUniqueId) LD(CommInitRank) LD(CommDestroy) LD(GetErrorString) LD(AllGather) LD(AllReduce) LD(Send) LD(Recv)
    LD(GroupStart) LD(GroupEnd)
#undef LD
    g_nccl.ok = true;
    return true;
}

static std::mutex g_lg_mu;
static std::map<long long, LocalGroup *> g_groups;

#define MD_NCCL(ctx, expr)                                                                            \
    do {                                                                                              \
        ncclResult_t _r = (expr);                                                                     \
        if (_r != ncclSuccess) {                                                                      \
            char _b[256];                                                                             \
            snprintf(_b, sizeof(_b), "NCCL error at %s:%d: %s", __FILE__, __LINE__, g_nccl.GetErrorString(_r)); \
            md_set_error(ctx, _b);                                                                    \
            return MDSEA_ERR_NCCL;                                                                    \
        }                                                                                             \
    } while (0)

extern "C" int mdsea_b200_nccl_unique_id(void *out128) {
    if (!out128) return MDSEA_ERR_BAD_ARG;
    if (!nccl_load()) { md_global_error() = "nccl_unique_id: libnccl.so.2 not found"; return MDSEA_ERR_NCCL; }
    ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != ncclSuccess) {
        md_global_error() = "nccl_unique_id: ncclGetUniqueId failed";
        return MDSEA_ERR_NCCL;
    }
    static_assert(NCCL_UNIQUE_ID_BYTES == 128, "header promises 128 bytes");
    memcpy(out128, id.internal, 128);
    return MDSEA_OK;
}

static int comm_common_init(mdsea_ctx *c, CommState **out) {
    if (!c) return MDSEA_ERR_BAD_ARG;
    if (c->p.world < 2) { md_set_error(c, "comm_init: world must be >= 2"); return MDSEA_ERR_BAD_ARG; }
    if (c->p.world > MD_MAXW) { md_set_error(c, "comm_init: world exceeds MD_MAXW"); return MDSEA_ERR_BAD_ARG; }
    if (c->comm) { md_set_error(c, "comm_init: already initialised"); return MDSEA_ERR_STATE; }
    CommState *s = new CommState();
    s->rank = c->p.rank;
    s->world = c->p.world;
    *out = s;
    return MDSEA_OK;
}

extern "C" int mdsea_b200_comm_init(mdsea_ctx *c, const void *unique_id128) {
    if (!unique_id128) return MDSEA_ERR_BAD_ARG;
    if (!nccl_load()) { md_set_error(c, "comm_init: libnccl.so.2 not found"); return MDSEA_ERR_NCCL; }
    CommState *s = nullptr;
    MD_TRY(comm_common_init(c, &s));
    MD_CUDA(c, cudaSetDevice(c->p.device));
    ncclUniqueId id;
    memcpy(id.internal, unique_id128, 128);
    ncclResult_t r = g_nccl.CommInitRank(&s->nccl, s->world, id, s->rank);
    if (r != ncclSuccess) {
        delete s;
        md_set_error(c, "comm_init: ncclCommInitRank failed");
        return MDSEA_ERR_NCCL;
    }
    s->backend = 1;
    c->comm = s;
    return md_cells_comm_ready(c);
}

extern "C" int mdsea_b200_comm_init_local(mdsea_ctx *c, int64_t group_key) {
    CommState *s = nullptr;
    MD_TRY(comm_common_init(c, &s));
    {
        std::lock_guard<std::mutex> lk(g_lg_mu);
        LocalGroup *&g = g_groups[group_key];
        if (!g) {
            g = new LocalGroup();
            g->world = s->world;
            pthread_barrier_init(&g->bar, nullptr, s->world);
        }
        if (g->world != s->world) { delete s; md_set_error(c, "comm_init_local: world mismatch within the group"); return MDSEA_ERR_BAD_ARG; }
        g->refs++;
        s->lg = g;
    }
    s->backend = 2;
    s->key = group_key;
    c->comm = s;
    return MDSEA_OK;
}

void md_comm_free(mdsea_ctx *c) {
    CommState *s = c->comm;
    if (!s) return;
    for (void *p : s->ipc_mapped) cudaIpcCloseMemHandle(p);
    if (s->backend == 1 && s->nccl) g_nccl.CommDestroy(s->nccl);
    if (s->backend == 2) {
        std::lock_guard<std::mutex> lk(g_lg_mu);
        if (--s->lg->refs == 0) {
            pthread_barrier_destroy(&s->lg->bar);
            g_groups.erase(s->key);
            delete s->lg;
        }
    }
    if (s->h_tmp) cudaFreeHost(s->h_tmp);
    delete s;
    c->comm = nullptr;
}

int md_comm_world(const mdsea_ctx *c) { return c->comm ? c->comm->world : 1; }

static int local_tmp(mdsea_ctx *c, size_t bytes) {
    CommState *s = c->comm;
    if (s->h_tmp_bytes >= bytes) return MDSEA_OK;
    if (s->h_tmp) cudaFreeHost(s->h_tmp);
    MD_CUDA(c, cudaMallocHost(&s->h_tmp, bytes));
    s->h_tmp_bytes = bytes;
    return MDSEA_OK;
}

// every rank contributes `count` ints; d_recv gets world*count ints in rank order
int md_comm_allgather_i32(mdsea_ctx *c, const int *d_send, int count, int *d_recv) {
    CommState *s = c->comm;
    if (s->backend == 1) {
        MD_NCCL(c, g_nccl.AllGather(d_send, d_recv, (size_t)count, ncclInt32, s->nccl, c->stream));
        return MDSEA_OK;
    }
    LocalGroup *g = s->lg;
    g->flat[s->rank] = d_send;
    MD_CUDA(c, cudaStreamSynchronize(c->stream));
    pthread_barrier_wait(&g->bar);
    for (int p = 0; p < s->world; ++p)
        MD_CUDA(c, cudaMemcpyAsync(d_recv + (size_t)p * count, g->flat[p], sizeof(int) * count, cudaMemcpyDefault, c->stream));
    MD_CUDA(c, cudaStreamSynchronize(c->stream));
    pthread_barrier_wait(&g->bar);
    return MDSEA_OK;
}

template <typename T, bool MAX>
static int local_allreduce(mdsea_ctx *c, T *d_buf, int count) {
    CommState *s = c->comm;
    LocalGroup *g = s->lg;
    MD_TRY(local_tmp(c, sizeof(T) * (size_t)count * (s->world + 1)));
    g->flat[s->rank] = d_buf;
    MD_CUDA(c, cudaStreamSynchronize(c->stream));
    pthread_barrier_wait(&g->bar);
    T *h = (T *)s->h_tmp;
    for (int p = 0; p < s->world; ++p)
        MD_CUDA(c, cudaMemcpyAsync(h + (size_t)(p + 1) * count, g->flat[p], sizeof(T) * count, cudaMemcpyDefault, c->stream));
    MD_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int k = 0; k < count; ++k) {
        T acc = h[(size_t)count + k];
        for (int p = 1; p < s->world; ++p) {  // fixed rank order: same bits on every rank
            const T v = h[(size_t)(p + 1) * count + k];
            acc = MAX ? (v > acc ? v : acc) : acc + v;
        }
        h[k] = acc;
    }
    pthread_barrier_wait(&g->bar);  // everyone has read the inputs before anyone overwrites them
    MD_CUDA(c, cudaMemcpyAsync(d_buf, h, sizeof(T) * count, cudaMemcpyHostToDevice, c->stream));
    MD_CUDA(c, cudaStreamSynchronize(c->stream));
    return MDSEA_OK;
}

int md_comm_allreduce_max_i32(mdsea_ctx *c, int *d_buf, int count) {
    CommState *s = c->comm;
    if (s->backend == 1) {
        MD_NCCL(c, g_nccl.AllReduce(d_buf, d_buf, (size_t)count, ncclInt32, ncclMax, s->nccl, c->stream));
        return MDSEA_OK;
    }
    return local_allreduce<int, true>(c, d_buf, count);
}

int md_comm_allreduce_sum_f64(mdsea_ctx *c, double *d_buf, int count) {
    CommState *s = c->comm;
    if (s->backend == 1) {
        MD_NCCL(c, g_nccl.AllReduce(d_buf, d_buf, (size_t)count, ncclFloat64, ncclSum, s->nccl, c->stream));
        return MDSEA_OK;
    }
    return local_allreduce<double, false>(c, d_buf, count);
}

// Peer-memory plumbing (NCCL transport, one process per GPU on ONE node): every rank owns one cudaMalloc'ed buffer;
// its CUDA IPC handle travels to all peers through an all-gather and each peer maps it, so that kernels can store
// straight into the neighbour's memory over NVLink.  peers[q] receives the mapped base pointer of rank q's buffer
// (peers[rank] = own).  Returns MDSEA_ERR_UNSUPPORTED (on EVERY rank, agreed through an all-reduce) when some rank
// cannot map some peer, so that the caller can fall back to ncclSend/ncclRecv consistently.
int md_comm_share_buffer(mdsea_ctx *c, void *own, void **peers) {
    CommState *s = c->comm;
    if (!s || s->backend != 1) return MDSEA_ERR_UNSUPPORTED;
    const int W = s->world;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    cudaIpcMemHandle_t mine;
    int ok = cudaIpcGetMemHandle(&mine, own) == cudaSuccess ? 1 : 0;
    int *d_buf = nullptr;
    MD_CUDA(c, cudaMalloc(&d_buf, sizeof(int) * 16 * (W + 1) + sizeof(int) * 4));
    int *d_all = d_buf + 16, *d_ok = d_all + 16 * W;
    MD_CUDA(c, cudaMemcpyAsync(d_buf, &mine, 64, cudaMemcpyHostToDevice, c->stream));
    int st = md_comm_allgather_i32(c, d_buf, 16, d_all);
    std::vector<cudaIpcMemHandle_t> all(W);
    if (st == MDSEA_OK) {
        MD_CUDA(c, cudaMemcpyAsync(all.data(), d_all, 64 * (size_t)W, cudaMemcpyDeviceToHost, c->stream));
        MD_CUDA(c, cudaStreamSynchronize(c->stream));
        for (int q = 0; q < W && ok; ++q) {
            if (q == s->rank) { peers[q] = own; continue; }
            void *p = nullptr;
            if (cudaIpcOpenMemHandle(&p, all[q], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = 0; cudaGetLastError(); break; }
            peers[q] = p;
            s->ipc_mapped.push_back(p);
        }
        // all ranks take the same decision: min over ranks = -(max of the negated flags)
        int neg = -ok;
        MD_CUDA(c, cudaMemcpy(d_ok, &neg, sizeof(int), cudaMemcpyHostToDevice));
        st = md_comm_allreduce_max_i32(c, d_ok, 1);
        if (st == MDSEA_OK) {
            MD_CUDA(c, cudaStreamSynchronize(c->stream));
            MD_CUDA(c, cudaMemcpy(&neg, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
            ok = -neg;
        }
    }
    cudaFree(d_buf);
    if (st != MDSEA_OK) return st;
    return ok ? MDSEA_OK : MDSEA_ERR_UNSUPPORTED;
}

int md_comm_backend(const mdsea_ctx *c) { return c->comm ? c->comm->backend : 0; }

// personalised exchange: send[q] (send_bytes[q]) goes to rank q, recv[p] (recv_bytes[p]) comes from rank p.
// Entries for q == rank are ignored (the caller handles its own data in place).
int md_comm_exchange(mdsea_ctx *c, void *const *d_send, const size_t *send_bytes, void *const *d_recv,
                     const size_t *recv_bytes, cudaStream_t stream) {
    CommState *s = c->comm;
    if (!stream) stream = c->stream;
    if (s->backend == 1) {
        MD_NCCL(c, g_nccl.GroupStart());
        for (int q = 0; q < s->world; ++q) {
            if (q == s->rank) continue;
            if (send_bytes[q]) MD_NCCL(c, g_nccl.Send(d_send[q], send_bytes[q], ncclUint8, q, s->nccl, stream));
            if (recv_bytes[q]) MD_NCCL(c, g_nccl.Recv(d_recv[q], recv_bytes[q], ncclUint8, q, s->nccl, stream));
        }
        MD_NCCL(c, g_nccl.GroupEnd());
        return MDSEA_OK;
    }
    LocalGroup *g = s->lg;
    for (int q = 0; q < s->world; ++q) {
        g->ptr[s->rank][q] = d_send[q];
        g->bytes[s->rank][q] = send_bytes[q];
    }
    MD_CUDA(c, cudaStreamSynchronize(stream));
    pthread_barrier_wait(&g->bar);
    for (int p = 0; p < s->world; ++p) {
        if (p == s->rank || !recv_bytes[p]) continue;
        if (g->bytes[p][s->rank] != recv_bytes[p]) {
            md_set_error(c, "comm_exchange: size mismatch between sender and receiver");
            pthread_barrier_wait(&g->bar);
            return MDSEA_ERR_STATE;
        }
        MD_CUDA(c, cudaMemcpyAsync(d_recv[p], g->ptr[p][s->rank], recv_bytes[p], cudaMemcpyDefault, stream));
    }
    MD_CUDA(c, cudaStreamSynchronize(stream));
    pthread_barrier_wait(&g->bar);
    return MDSEA_OK;
}
