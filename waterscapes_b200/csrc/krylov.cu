#include "ctx.h"
void krylov_free(mpet_ctx*) {}
void krylov_solve(mpet_ctx* ctx, const double*, double*, double*, cudaStream_t) { MPET_REQUIRE(false, "krylov: not implemented"); }
void pc_setup(mpet_ctx* ctx, cudaStream_t) { MPET_REQUIRE(false, "pc: not implemented"); }
void pc_apply(mpet_ctx* ctx, const double*, double*, cudaStream_t) { MPET_REQUIRE(false, "pc: not implemented"); }
void amg_free(mpet_ctx*) {}
void dist_attach(mpet_ctx* ctx, const void*, int, int) { MPET_REQUIRE(false, "dist: not implemented"); }
void dist_free(mpet_ctx*) {}
