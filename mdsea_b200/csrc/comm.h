// Rank-to-rank primitives of the domain decomposition (comm.cu).
#pragma once
#include <stddef.h>

struct mdsea_ctx;
#define MD_MAXW 16  // ranks per job (one node: 8 GPUs)

void md_comm_free(mdsea_ctx *c);
int  md_comm_world(const mdsea_ctx *c);
int  md_comm_backend(const mdsea_ctx *c);   // 0 none, 1 NCCL, 2 in-process
int  md_comm_share_buffer(mdsea_ctx *c, void *own, void **peers);
int  md_comm_allgather_i32(mdsea_ctx *c, const int *d_send, int count, int *d_recv);
int  md_comm_allreduce_max_i32(mdsea_ctx *c, int *d_buf, int count);
int  md_comm_allreduce_sum_f64(mdsea_ctx *c, double *d_buf, int count);
#include <cuda_runtime.h>
int  md_comm_exchange(mdsea_ctx *c, void *const *d_send, const size_t *send_bytes, void *const *d_recv,
                      const size_t *recv_bytes, cudaStream_t stream = nullptr);
