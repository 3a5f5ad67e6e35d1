// cutoff path (cell list + Verlet list) -- placeholder until K2-K4 land
#include "ctx.h"
static int unsup(mdsea_ctx *c) { md_set_error(c, "cutoff path not built"); return MDSEA_ERR_UNSUPPORTED; }
int md_cells_setup(mdsea_ctx *c) { return unsup(c); }
void md_cells_free(mdsea_ctx *) {}
int md_cutoff_force(mdsea_ctx *c) { return unsup(c); }
int md_cutoff_set_state(mdsea_ctx *c, const double *, const double *) { return unsup(c); }
int md_cutoff_get_state(mdsea_ctx *c, double *, double *, double *, int64_t *, int64_t *) { return unsup(c); }
int md_cutoff_steps(mdsea_ctx *c, const mdsea_run_args *, double) { return unsup(c); }
int md_cutoff_finish_rows(mdsea_ctx *c) { return unsup(c); }
extern "C" int mdsea_b200_get_cells(mdsea_ctx *c, int32_t *, int32_t *, int32_t *) { return unsup(c); }
extern "C" int mdsea_b200_get_neighbors(mdsea_ctx *c, int64_t *, int64_t *, int32_t *) { return unsup(c); }
extern "C" int mdsea_b200_nccl_unique_id(void *) { return MDSEA_ERR_UNSUPPORTED; }
extern "C" int mdsea_b200_comm_init(mdsea_ctx *c, const void *) { return unsup(c); }
