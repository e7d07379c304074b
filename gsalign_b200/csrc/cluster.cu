#include "gsa_internal.cuh"
int gsa_impl_cluster(gsa_ctx *ctx) { return gsa_fail(ctx, GSA_ERR_ARG, "cluster: not built yet"); }
extern "C" int64_t gsa_dump_blocks(gsa_ctx *ctx, int32_t stage, int64_t *out) { (void)stage; (void)out; return gsa_fail(ctx, GSA_ERR_ARG, "not built yet"); }
