// Host-side context of one solver instance (one GPU / one rank).
#pragma once
#include "common.cuh"

struct CellState;   // cutoff path (cells.cu / nbr.cu)
struct CommState;   // domain decomposition (comm.cu)

struct mdsea_ctx {
    mdsea_params p;
    std::string  err;
    int          ndim = 0;
    long long    n = 0;        // particles resident on this rank (== N when world == 1)
    long long    n_global = 0;
    size_t       E = 0;        // ndim * n  (flat SoA length)
    PotDev       pot;
    BoxDev       box;
    bool         cutoff_path = false;
    bool         have_state = false;
    bool         acc_valid = false;   // d_acc holds a(r) for the current positions
    bool         profiling = false;

    cudaStream_t stream = nullptr;       // compute stream
    cudaStream_t copy_stream = nullptr;  // D2H of buffered rows
    cudaStream_t comm_stream = nullptr;  // halo refresh, overlapped with the interior force kernel
    cudaEvent_t  ev_batch_done = nullptr;
    cudaEvent_t  ev_a = nullptr, ev_b = nullptr;

    // state, SoA (ndim, n) float64
    double *d_r = nullptr, *d_v = nullptr, *d_acc = nullptr;
    // all-pairs force partials [nsplit][ndim][n] and per-block reductions
    double *d_fpart = nullptr;
    int     ap_ti = 128, ap_nsplit = 1, ap_jchunk = 0, ap_blocks_x = 0;
    bool    ap_n3 = false;                 // fp64 Newton's-third-law kernel: fpart holds one slot per partner tile
    double *d_pe_part = nullptr;           // per force-kernel block
    unsigned long long *d_cnt_part = nullptr;
    int     n_pe_part = 0;
    double *d_ke_part = nullptr;           // per integrator block
    int     n_ke_part = 0;
    // fp32 hi/lo split of the positions for the fp32 all-pairs kernel
    float  *d_rhi = nullptr, *d_rlo = nullptr;

    // ring buffer of rows.  Two ring sets alternate between consecutive run_async calls, so the D2H copies of one
    // batch (copy stream) overlap the compute of the next; the pointers below alias the set of the CURRENT batch.
    // The second set is allocated on first use (a batch submitted while another one is still being copied out).
    int     ring = 0;
    double *d_r_rows = nullptr, *d_v_rows = nullptr;  // [ring][E]
    double *d_pe_rows = nullptr, *d_ke_rows = nullptr, *d_temp_rows = nullptr;  // [ring]
    int    *d_id_rows = nullptr;                      // [ring][cap] particle ids of the buffered rows (world > 1)
    struct RingSet {
        double *r_rows = nullptr, *v_rows = nullptr, *pe_rows = nullptr, *ke_rows = nullptr, *temp_rows = nullptr;
        int    *id_rows = nullptr;
        double *d_quench = nullptr, *h_quench = nullptr;  // quench targets of the batch (device array, pinned staging)
        StepScalars *h_sc = nullptr;                  // pinned mirror of the scalars at the end of the batch
        StepScalars *d_sc_snap = nullptr;             // device snapshot taken on the compute stream at the end of the batch
        cudaEvent_t  ev_copy_done = nullptr;
        bool    allocated = false, pending = false;
    } rings[2];
    size_t  coord_rows_bytes = 0, id_rows_bytes = 0;  // sizes of r_rows / v_rows and id_rows of one set
    int     ring_cur = 0, ring_oldest = 0, n_pending = 0;
    double *d_quench = nullptr;                       // [ring][2]
    StepScalars *d_sc = nullptr;
    StepScalars *h_sc = nullptr;                      // pinned mirror
    double *h_quench = nullptr;                       // pinned staging of the batch's quench targets

    mdsea_stats stats;
    CellState  *cells = nullptr;
    CommState  *comm = nullptr;

};

void md_set_error(mdsea_ctx *ctx, const char *msg);

// CUDA-event timers around a kernel family; only active in profiling mode (adds a sync per kernel)
static inline void md_prof_begin(mdsea_ctx *c) {
    if (c->profiling) cudaEventRecord(c->ev_a, c->stream);
}
static inline void md_prof_end(mdsea_ctx *c, double *acc_ms) {
    if (!c->profiling) return;
    cudaEventRecord(c->ev_b, c->stream);
    cudaEventSynchronize(c->ev_b);
    float ms = 0;
    cudaEventElapsedTime(&ms, c->ev_a, c->ev_b);
    *acc_ms += ms;
}

// allpairs.cu
int md_allpairs_setup(mdsea_ctx *c);
int md_allpairs_force(mdsea_ctx *c);          // fills d_fpart / d_pe_part / d_cnt_part from d_r
// integrate.cu
int md_apply_bc(mdsea_ctx *c);
int md_apply_bc_range(mdsea_ctx *c, double *r, double *v, size_t E);
int md_com_rog(mdsea_ctx *c, const double *r, long long n, size_t stride, double *com, double *rog);
int md_kick_drift(mdsea_ctx *c, bool do_bc, bool do_kick, const double *fsrc, int nparts, size_t stride);
struct FinalizeArgs;
int md_kick_finish(mdsea_ctx *c, const double *fsrc, int nparts, size_t stride, int row, int record, double g_dt,
                   const FinalizeArgs &fin, bool drift_next);
int md_reduce_acc(mdsea_ctx *c, const double *fsrc, int nparts, size_t stride);
FinalizeArgs md_finalize_args(mdsea_ctx *c, int row, int next_row_valid, double kb, long long step_index);
int md_prepare_first(mdsea_ctx *c, double kb);
int md_commit_pairs(mdsea_ctx *c);
int md_pair_dists(mdsea_ctx *c, double *d_dists, double *d_units);
// cells.cu / nbr.cu (cutoff path)
int md_cells_setup(mdsea_ctx *c);
void md_cells_free(mdsea_ctx *c);
int md_cutoff_force(mdsea_ctx *c);            // rebuilds the list when flagged, then forces
int md_cutoff_copy_rows(mdsea_ctx *c, const mdsea_run_args *a, const mdsea_rows *rows);
long long md_cutoff_capacity(mdsea_ctx *c);
int md_cutoff_apply_bc(mdsea_ctx *c);
