#include "gsa_internal.cuh"
int gsa_impl_fill(gsa_ctx *ctx, gsa_alignment *out) { (void)out; return gsa_fail(ctx, GSA_ERR_ARG, "fill: not built yet"); }
int gsa_impl_dp_batch(gsa_ctx *ctx, int32_t, const char *, const int64_t *, const char *, const int64_t *, char *, char *, int32_t *, float *) { return gsa_fail(ctx, GSA_ERR_ARG, "not built yet"); }
