// extern "C" entry points of libmpet_b200.so (see include/mpet_b200.h for the contract).
#include "ctx.h"
#include <cstring>

// rhs.cu
void rhs_prev(mpet_ctx* ctx, const double* up, double* b, cudaStream_t st);
void export_values(mpet_ctx* ctx, int which, double* out, cudaStream_t st);
void set_dirichlet_dofs(mpet_ctx* ctx, const int32_t* dofs, int64_t n, cudaStream_t st);
void scatter_bc_values(mpet_ctx* ctx, double* out, cudaStream_t st);
void add_entries(mpet_ctx* ctx, const int32_t* r, const int32_t* c, const double* v, int64_t n, cudaStream_t st);
// krylov.cu / amg.cu
void krylov_free(mpet_ctx* ctx);
void krylov_solve(mpet_ctx* ctx, const double* b, double* x, double* info, cudaStream_t st);
void pc_setup(mpet_ctx* ctx, cudaStream_t st);
void pc_apply(mpet_ctx* ctx, const double* r, double* z, cudaStream_t st);
void spmv_api(mpet_ctx* ctx, const double* x, double* y, cudaStream_t st);
void amg_free(mpet_ctx* ctx);
// dist.cu
void dist_attach(mpet_ctx* ctx, const void* uid, int rank, int nranks);
void dist_free(mpet_ctx* ctx);
void dist_unique_id(void* out128);
void dist_set_halo(mpet_ctx* ctx, int nnbr, const int* ranks, const int64_t* send_off, const int32_t* send_idx_dev,
                   const int64_t* recv_off, const int32_t* recv_idx_dev, const uint8_t* owned_dev, cudaStream_t st);

namespace {
__global__ void k_cell_dofs(const int32_t* __restrict__ cell_nodes, int64_t nc, int64_t n2, int64_t nv,
                            int A, int32_t* __restrict__ out) {
    int nloc = 30 + 4 * A;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nc * nloc) return;
    int64_t c = i / nloc;
    int l = (int)(i - c * nloc);
    if (l < 30) {
        int k = l / 10, a = l - 10 * k;
        out[i] = (int32_t)(k * n2 + cell_nodes[c * 10 + a]);
    } else {
        int q = l - 30;
        int n = q / 4, m = q - 4 * n;
        out[i] = (int32_t)(3 * n2 + n * nv + cell_nodes[c * 10 + m]);
    }
}

// ---- external dof numbering (mpet_set_dof_permutation)
__global__ void k_perm_gather(int64_t n, const int32_t* __restrict__ perm, const double* __restrict__ ext,
                              double* __restrict__ con) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) con[i] = ext[perm[i]];
}
__global__ void k_perm_scatter(int64_t n, const int32_t* __restrict__ perm, const double* __restrict__ con,
                               double* __restrict__ ext) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) ext[perm[i]] = con[i];
}
__global__ void k_perm_invert(int64_t n, const int32_t* __restrict__ perm, int32_t* __restrict__ inv, int* __restrict__ bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t e = perm[i];
    if (e < 0 || e >= n) { atomicAdd(bad, 1); return; }
    inv[e] = (int32_t)i;
}
__global__ void k_perm_holes(int64_t n, const int32_t* __restrict__ inv, int* __restrict__ bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n && inv[i] < 0) atomicAdd(bad, 1);
}
__global__ void k_perm_map(int64_t n, int64_t ndof, const int32_t* __restrict__ inv, const int32_t* __restrict__ in,
                           int32_t* __restrict__ out, int* __restrict__ bad) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int32_t e = in[i];
    if (e < 0 || e >= ndof) { atomicAdd(bad, 1); out[i] = 0; return; }
    out[i] = inv[e];
}
}  // namespace

// caller's vector (external numbering, `tail` extra entries behind the N dofs pass through) -> contract numbering
static const double* to_contract(mpet_ctx* ctx, const double* ext, double* scratch, int64_t tail, cudaStream_t st) {
    if (!ctx->perm) return ext;
    k_perm_gather<<<grid_for(ctx->N, 256), 256, 0, st>>>(ctx->N, ctx->perm, ext, scratch);
    LAUNCH_CHECK(ctx);
    if (tail > 0)
        CUDA_CHECK(cudaMemcpyAsync(scratch + ctx->N, ext + ctx->N, sizeof(double) * tail, cudaMemcpyDeviceToDevice, st));
    return scratch;
}
static void from_contract(mpet_ctx* ctx, const double* con, double* ext, int64_t tail, cudaStream_t st) {
    if (!ctx->perm) return;
    k_perm_scatter<<<grid_for(ctx->N, 256), 256, 0, st>>>(ctx->N, ctx->perm, con, ext);
    LAUNCH_CHECK(ctx);
    if (tail > 0)
        CUDA_CHECK(cudaMemcpyAsync(ext + ctx->N, con + ctx->N, sizeof(double) * tail, cudaMemcpyDeviceToDevice, st));
}
// caller's dof indices -> contract indices (temporary device array, freed by the caller after a stream sync)
static int32_t* map_dofs(mpet_ctx* ctx, const int32_t* ext, int64_t n, cudaStream_t st) {
    int32_t* out = nullptr;
    int* bad = nullptr;
    CUDA_CHECK(cudaMalloc(&out, sizeof(int32_t) * std::max<int64_t>(n, 1)));
    CUDA_CHECK(cudaMalloc(&bad, sizeof(int)));
    CUDA_CHECK(cudaMemsetAsync(bad, 0, sizeof(int), st));
    if (n > 0) {
        k_perm_map<<<grid_for(n, 256), 256, 0, st>>>(n, ctx->N, ctx->perm_inv, ext, out, bad);
        LAUNCH_CHECK(ctx);
    }
    int hbad = 0;
    CUDA_CHECK(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(bad);
    if (hbad) { cudaFree(out); throw MpetError("dof index outside 0..N-1"); }
    return out;
}

static cudaEvent_t prof_event(mpet_ctx* ctx) {
    if (!ctx->prof.pool.empty()) {
        cudaEvent_t e = ctx->prof.pool.back();
        ctx->prof.pool.pop_back();
        return e;
    }
    cudaEvent_t e;
    CUDA_CHECK(cudaEventCreate(&e));
    return e;
}

cudaEvent_t prof_begin(mpet_ctx* ctx, cudaStream_t st) {
    if (!ctx->prof.on || ctx->prof_suspended || ctx->prof.pending.size() > 20000) return nullptr;
    cudaEvent_t a = prof_event(ctx);
    CUDA_CHECK(cudaEventRecord(a, st));
    return a;
}

void prof_end(mpet_ctx* ctx, int cat, cudaEvent_t a, cudaStream_t st) {
    if (!a) return;
    cudaEvent_t b = prof_event(ctx);
    CUDA_CHECK(cudaEventRecord(b, st));
    ctx->prof.pending.push_back({cat, a, b});
}

void prof_collect(mpet_ctx* ctx) {   // call after the stream has been synchronised
    for (auto& sp : ctx->prof.pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, sp.a, sp.b) == cudaSuccess) {
            ctx->prof.ms[sp.cat] += ms;
            ctx->prof.cnt[sp.cat] += 1;
        }
        ctx->prof.pool.push_back(sp.a);
        ctx->prof.pool.push_back(sp.b);
    }
    ctx->prof.pending.clear();
}

extern "C" {

int mpet_profile(mpet_ctx* ctx, int enable, double* out_host) {
    MPET_TRY(ctx)
    CUDA_CHECK(cudaDeviceSynchronize());
    prof_collect(ctx);
    if (out_host) {
        for (int i = 0; i < PROF_NCAT; ++i) { out_host[i] = ctx->prof.ms[i]; out_host[PROF_NCAT + i] = (double)ctx->prof.cnt[i]; }
    }
    if (enable >= 0) {
        ctx->prof.on = enable != 0;
        for (int i = 0; i < PROF_NCAT; ++i) { ctx->prof.ms[i] = 0; ctx->prof.cnt[i] = 0; }
    }
    MPET_CATCH(ctx)
}

int mpet_abi_version(void) { return 1; }

int mpet_create(int device, mpet_ctx** out) {
    if (!out) return -1;
    *out = nullptr;
    mpet_ctx* ctx = new (std::nothrow) mpet_ctx();
    if (!ctx) return -1;
    ctx->device = device;
    *out = ctx;
    MPET_TRY(ctx)
    CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    ctx->sm_count = prop.multiProcessorCount;
    MPET_REQUIRE(prop.major >= 10, "libmpet_b200 is built for sm_100a (Blackwell B200) only");
    init_reference_tables();
    MPET_CATCH(ctx)
}

void mpet_destroy(mpet_ctx* ctx) {
    if (!ctx) return;
    try {
        cudaSetDevice(ctx->device);
        cudaDeviceSynchronize();
        krylov_free(ctx);
        amg_free(ctx);
        dist_free(ctx);
        for (int i = 0; i < ctx->pc_streams_ready; ++i) { cudaStreamDestroy(ctx->pc_stream[i]); cudaEventDestroy(ctx->pc_join[i]); }
        if (ctx->pc_fork) cudaEventDestroy(ctx->pc_fork);
        if (ctx->solve_stream) { cudaStreamDestroy(ctx->solve_stream); cudaEventDestroy(ctx->solve_event); }
        for (void* p : ctx->allocs) cudaFree(p);
    } catch (...) {
    }
    delete ctx;
}

const char* mpet_last_error(mpet_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int mpet_set_mesh(mpet_ctx* ctx, const double* coords_dev, const int32_t* cells_dev, int64_t nv,
                  int64_t nc, int n_networks, void* stream) {
    MPET_TRY(ctx)
    MPET_REQUIRE(ctx->Nc == 0, "mesh already set on this context");
    MPET_REQUIRE(nv > 0 && nc > 0, "empty mesh");
    MPET_REQUIRE(n_networks >= 0 && n_networks <= MPET_MAX_NETWORKS, "unsupported number of networks");
    CUDA_CHECK(cudaSetDevice(ctx->device));
    cudaStream_t st = as_stream(stream);
    ctx->Nv = nv;
    ctx->Nc = nc;
    ctx->A = n_networks;
    ctx->coords = dev_alloc<double>(ctx, nv * 3);
    ctx->cells = dev_alloc<int32_t>(ctx, nc * 4);
    CUDA_CHECK(cudaMemcpyAsync(ctx->coords, coords_dev, sizeof(double) * nv * 3, cudaMemcpyDeviceToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(ctx->cells, cells_dev, sizeof(int32_t) * nc * 4, cudaMemcpyDeviceToDevice, st));
    build_space(ctx, st);
    compute_geometry(ctx, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
    staged_build_block_plans(ctx);
    CUDA_CHECK(cudaStreamSynchronize(st));
    MPET_CATCH(ctx)
}

int mpet_get_sizes(mpet_ctx* ctx, int64_t* s) {
    MPET_TRY(ctx)
    s[0] = ctx->Nv; s[1] = ctx->Ne; s[2] = ctx->N2; s[3] = ctx->Nc; s[4] = ctx->N; s[5] = ctx->nnz;
    s[6] = ctx->g22.nnz; s[7] = ctx->g21.nnz; s[8] = ctx->g11.nnz; s[9] = ctx->A;
    MPET_CATCH(ctx)
}

int mpet_get_edges(mpet_ctx* ctx, int32_t* out, void* stream) {
    MPET_TRY(ctx)
    CUDA_CHECK(cudaMemcpyAsync(out, ctx->edge_v, sizeof(int32_t) * 2 * ctx->Ne, cudaMemcpyDeviceToDevice,
                               as_stream(stream)));
    MPET_CATCH(ctx)
}

int mpet_get_cell_dofs(mpet_ctx* ctx, int32_t* out, void* stream) {
    MPET_TRY(ctx)
    int64_t n = ctx->Nc * ctx->nloc;
    k_cell_dofs<<<grid_for(n, 256), 256, 0, as_stream(stream)>>>(ctx->cell_nodes, ctx->Nc, ctx->N2, ctx->Nv,
                                                                 ctx->A, out);
    LAUNCH_CHECK(ctx);
    MPET_CATCH(ctx)
}

int mpet_get_pattern(mpet_ctx* ctx, int64_t* rowptr, int32_t* cols, void* stream) {
    MPET_TRY(ctx)
    cudaStream_t st = as_stream(stream);
    CUDA_CHECK(cudaMemcpyAsync(rowptr, ctx->rowptr, sizeof(int64_t) * (ctx->N + 1), cudaMemcpyDeviceToDevice, st));
    CUDA_CHECK(cudaMemcpyAsync(cols, ctx->cols, sizeof(int32_t) * ctx->nnz, cudaMemcpyDeviceToDevice, st));
    MPET_CATCH(ctx)
}

static void store_material(mpet_ctx* ctx, int J, double E, double nu, const double* alpha, const double* K,
                           const double* S, const double* c, double dt, double theta) {
    ctx->E = E;
    ctx->nu = nu;
    ctx->mu = E / (2.0 * (1.0 + nu));                           // mpetproblem.py:13-16
    ctx->lmbda = nu * E / ((1.0 - 2.0 * nu) * (1.0 + nu));
    for (int i = 0; i < J; ++i) {
        ctx->alpha[i] = alpha[i];
        ctx->K[i] = K[i];
        ctx->c[i] = c[i];
        for (int j = 0; j < J; ++j) ctx->S[i * J + j] = S[i * J + j];
    }
    ctx->dt = dt;
    ctx->theta = theta;
    ctx->coef = BlockCoefs();
}

int mpet_set_params(mpet_ctx* ctx, double E, double nu, const double* alpha, const double* K,
                    const double* S, const double* c, double dt, double theta) {
    MPET_TRY(ctx)
    const int A = ctx->A;
    MPET_REQUIRE(ctx->Nc > 0, "mpet_set_mesh must be called first");
    store_material(ctx, A, E, nu, alpha, K, S, c, dt, theta);
    // standard formulation, a = lhs(F) / L = rhs(F) of mpetsolver.py:196-201,260-261; prec :268-272
    BlockCoefs& C = ctx->coef;
    const double dth = dt * theta, d1 = dt * (1.0 - theta);
    C.uu_mu = ctx->mu;
    C.uu_lam = ctx->lmbda;
    C.p_mu = ctx->mu;
    for (int i = 0; i < A; ++i) {
        double offsum = 0;
        for (int j = 0; j < A; ++j)
            if (j != i) offsum += S[i * A + j];
        C.cup[i] = C.cpu[i] = -alpha[i];
        C.cl[i] = -dth * K[i];
        C.rl[i] = d1 * K[i];
        C.ru[i] = 1.0;                       // -alpha_i int psi div(u^-) is A's own (p_i ; u) block applied to u^-
        C.pm[i] = c[i] + dth * offsum;
        C.pk[i] = dth * K[i];
        for (int j = 0; j < A; ++j) {
            C.cm[i * A + j] = (i == j) ? (-c[i] - dth * offsum) : dth * S[i * A + j];
            C.rm[i * A + j] = (i == j) ? (-c[i] + d1 * offsum) : -d1 * S[i * A + j];
        }
    }
    ctx->formulation = 0;
    ctx->params_set = true;
    MPET_CATCH(ctx)
}

int mpet_set_params_total_pressure(mpet_ctx* ctx, double E, double nu, const double* alpha, const double* K,
                                   const double* S, const double* c, double dt, double theta) {
    MPET_TRY(ctx)
    MPET_REQUIRE(ctx->Nc > 0, "mpet_set_mesh must be called first");
    MPET_REQUIRE(ctx->A >= 1, "the total-pressure formulation needs n_networks + 1 P1 fields (mpet_set_mesh)");
    const int NP = ctx->A, J = NP - 1;       // field 0 = total pressure p0, field i+1 = network i
    store_material(ctx, J, E, nu, alpha, K, S, c, dt, theta);
    // F of mpettotalpressuresolver.py:267-274 split with lhs/rhs; prec :276-283
    BlockCoefs& C = ctx->coef;
    const double dth = dt * theta, d1 = dt * (1.0 - theta), il = 1.0 / ctx->lmbda;
    C.uu_mu = ctx->mu;                        // 2 mu eps(u):eps(v) = mu (grad u:grad v + grad u^T:grad v)
    C.uu_lam = 0.0;
    C.p_mu = ctx->mu;
    C.cup[0] = C.cpu[0] = 1.0;                // + p0 div v ;  + div u w0
    C.cm[0] = -il;                            // -1/lambda p0 w0
    C.pm[0] = 1.0;                            // ppt = p0 w0
    for (int i = 0; i < J; ++i) {
        const int r = i + 1;
        double offsum = 0;
        for (int j = 0; j < J; ++j)
            if (j != i) offsum += S[i * J + j];
        C.cm[0 * NP + r] = -alpha[i] * il;    // -alpha_i/lambda p_i w0
        C.cm[r * NP + 0] = -alpha[i] * il;    // -alpha_i/lambda p0 w_i
        C.rm[r * NP + 0] = -alpha[i] * il;    // -alpha_i/lambda p0^- w_i   (rhs sign: see DESIGN.md)
        C.cl[r] = -dth * K[i];
        C.rl[r] = d1 * K[i];
        C.pm[r] = alpha[i] * alpha[i] * il + c[i] + dth * offsum;
        C.pk[r] = dth * K[i];
        for (int j = 0; j < J; ++j) {
            const int q = j + 1;
            const double aa = alpha[i] * alpha[j] * il;
            C.cm[r * NP + q] = -aa + ((i == j) ? (-c[i] - dth * offsum) : dth * S[i * J + j]);
            C.rm[r * NP + q] = -aa + ((i == j) ? (-c[i] + d1 * offsum) : -d1 * S[i * J + j]);
        }
    }
    ctx->formulation = 1;
    ctx->params_set = true;
    MPET_CATCH(ctx)
}

int mpet_set_cell_coefficient(mpet_ctx* ctx, int field, const double* values_dev, void* stream) {
    MPET_TRY(ctx)
    MPET_REQUIRE(ctx->Nc > 0, "mpet_set_mesh must be called first");
    MPET_REQUIRE(field >= 0 && field < ctx->A, "field index out of range");
    cudaStream_t st = as_stream(stream);
    if (values_dev) {
        if (!ctx->kcell[field]) {
            ctx->kcell[field] = dev_alloc<double>(ctx, ctx->Nc);
            ctx->l11w[field] = dev_alloc<double>(ctx, ctx->g11.nnz);
        }
        CUDA_CHECK(cudaMemcpyAsync(ctx->kcell[field], values_dev, sizeof(double) * ctx->Nc, cudaMemcpyDeviceToDevice, st));
    } else if (ctx->kcell[field]) {
        dev_free(ctx, ctx->kcell[field]);
        dev_free(ctx, ctx->l11w[field]);
        ctx->kcell[field] = nullptr;
        ctx->l11w[field] = nullptr;
    }
    ctx->cell_coef_dirty = true;
    ctx->lhs_ready = false;
    ctx->prec_ready = false;
    MPET_CATCH(ctx)
}

int mpet_set_dof_permutation(mpet_ctx* ctx, const int32_t* ext_of_contract_dev, void* stream) {
    MPET_TRY(ctx)
    MPET_REQUIRE(ctx->N > 0, "mpet_set_mesh must be called first");
    cudaStream_t st = as_stream(stream);
    CUDA_CHECK(cudaStreamSynchronize(st));
    if (ctx->perm) {
        dev_free(ctx, ctx->perm); dev_free(ctx, ctx->perm_inv); dev_free(ctx, ctx->perm_a); dev_free(ctx, ctx->perm_b);
        ctx->perm = ctx->perm_inv = nullptr;
        ctx->perm_a = ctx->perm_b = nullptr;
    }
    if (ext_of_contract_dev) {
        int32_t* perm = dev_alloc<int32_t>(ctx, ctx->N);
        int32_t* inv = dev_alloc<int32_t>(ctx, ctx->N);
        int* bad = nullptr;
        CUDA_CHECK(cudaMalloc(&bad, sizeof(int)));
        CUDA_CHECK(cudaMemsetAsync(bad, 0, sizeof(int), st));
        CUDA_CHECK(cudaMemcpyAsync(perm, ext_of_contract_dev, sizeof(int32_t) * ctx->N, cudaMemcpyDeviceToDevice, st));
        CUDA_CHECK(cudaMemsetAsync(inv, 0xFF, sizeof(int32_t) * ctx->N, st));
        k_perm_invert<<<grid_for(ctx->N, 256), 256, 0, st>>>(ctx->N, perm, inv, bad);
        LAUNCH_CHECK(ctx);
        k_perm_holes<<<grid_for(ctx->N, 256), 256, 0, st>>>(ctx->N, inv, bad);
        LAUNCH_CHECK(ctx);
        int hbad = 0;
        CUDA_CHECK(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        cudaFree(bad);
        if (hbad) {
            dev_free(ctx, perm); dev_free(ctx, inv);
            MPET_REQUIRE(false, "the dof permutation is not a bijection of 0..N-1");
        }
        ctx->perm = perm;
        ctx->perm_inv = inv;
        ctx->perm_a = dev_alloc<double>(ctx, ctx->N + 16);
        ctx->perm_b = dev_alloc<double>(ctx, ctx->N + 16);
    }
    MPET_CATCH(ctx)
}

int mpet_assemble_lhs(mpet_ctx* ctx, void* stream) {
    MPET_TRY(ctx)
    cudaEvent_t pe = prof_begin(ctx, as_stream(stream));
    assemble_lhs(ctx, as_stream(stream));
    prof_end(ctx, PROF_ASM, pe, as_stream(stream));
    MPET_CATCH(ctx)
}

int mpet_add_entries(mpet_ctx* ctx, const int32_t* rows, const int32_t* cols, const double* vals,
                     int64_t n, void* stream) {
    MPET_TRY(ctx)
    if (ctx->perm) {
        int32_t* r = map_dofs(ctx, rows, n, as_stream(stream));
        int32_t* c = nullptr;
        try { c = map_dofs(ctx, cols, n, as_stream(stream)); } catch (...) { cudaFree(r); throw; }
        try { add_entries(ctx, r, c, vals, n, as_stream(stream)); CUDA_CHECK(cudaStreamSynchronize(as_stream(stream))); }
        catch (...) { cudaFree(r); cudaFree(c); throw; }
        cudaFree(r); cudaFree(c);
    } else {
        add_entries(ctx, rows, cols, vals, n, as_stream(stream));
    }
    MPET_CATCH(ctx)
}

int mpet_assemble_prec(mpet_ctx* ctx, void* stream) {
    MPET_TRY(ctx)
    assemble_prec(ctx, as_stream(stream));
    MPET_CATCH(ctx)
}

int mpet_get_values(mpet_ctx* ctx, int which, double* vals, void* stream) {
    MPET_TRY(ctx)
    MPET_REQUIRE(which >= 0 && which <= 3, "which must be 0..3");
    export_values(ctx, which, vals, as_stream(stream));
    MPET_CATCH(ctx)
}

int mpet_set_dirichlet_dofs(mpet_ctx* ctx, const int32_t* dofs, int64_t n, void* stream) {
    MPET_TRY(ctx)
    MPET_REQUIRE(ctx->N > 0, "mpet_set_mesh must be called first");
    if (ctx->perm) {
        int32_t* d = map_dofs(ctx, dofs, n, as_stream(stream));
        try { set_dirichlet_dofs(ctx, d, n, as_stream(stream)); CUDA_CHECK(cudaStreamSynchronize(as_stream(stream))); }
        catch (...) { cudaFree(d); throw; }
        cudaFree(d);
    } else {
        set_dirichlet_dofs(ctx, dofs, n, as_stream(stream));
    }
    MPET_CATCH(ctx)
}

int mpet_set_dirichlet_values(mpet_ctx* ctx, const double* vals, void* stream) {
    MPET_TRY(ctx)
    if (ctx->n_bc > 0)
        CUDA_CHECK(cudaMemcpyAsync(ctx->bc_vals, vals, sizeof(double) * ctx->n_bc, cudaMemcpyDeviceToDevice,
                                   as_stream(stream)));
    MPET_CATCH(ctx)
}

int mpet_rhs_prev(mpet_ctx* ctx, const double* up_prev, double* b, void* stream) {
    MPET_TRY(ctx)
    cudaEvent_t pe = prof_begin(ctx, as_stream(stream));
    if (ctx->perm) {
        const double* up_c = to_contract(ctx, up_prev, ctx->perm_a, 0, as_stream(stream));
        to_contract(ctx, b, ctx->perm_b, 0, as_stream(stream));
        rhs_prev(ctx, up_c, ctx->perm_b, as_stream(stream));
        from_contract(ctx, ctx->perm_b, b, 0, as_stream(stream));
    } else {
        rhs_prev(ctx, up_prev, b, as_stream(stream));
    }
    prof_end(ctx, PROF_RHS, pe, as_stream(stream));
    MPET_CATCH(ctx)
}

int mpet_mass_apply(mpet_ctx* ctx, int space, double scale, const double* x, double* y, void* stream) {
    MPET_TRY(ctx)
    cudaStream_t st = as_stream(stream);
    DevCsr M;
    if (space == 2) {
        ensure_m22(ctx, st);
        M.nrows = M.ncols = ctx->N2; M.nnz = ctx->g22.nnz;
        M.rowptr = ctx->g22.rowptr; M.col = ctx->g22.col; M.val = ctx->m22;
    } else if (space == 1) {
        MPET_REQUIRE(ctx->lhs_ready || ctx->prec_ready, "assemble first (P1 mass is a by-product)");
        M.nrows = M.ncols = ctx->Nv; M.nnz = ctx->g11.nnz;
        M.rowptr = ctx->g11.rowptr; M.col = ctx->g11.col; M.val = ctx->m11;
    } else {
        MPET_REQUIRE(false, "space must be 1 (P1) or 2 (P2)");
    }
    csr32_spmm(ctx, M, x, M.ncols, y, M.nrows, 1, scale, 1.0, st);
    MPET_CATCH(ctx)
}

int mpet_lumped(mpet_ctx* ctx, int space, double* w, void* stream) {
    MPET_TRY(ctx)
    cudaStream_t st = as_stream(stream);
    int64_t n = (space == 2) ? ctx->N2 : ctx->Nv;
    std::vector<double> ones((size_t)n, 1.0);
    double* d1 = nullptr;
    CUDA_CHECK(cudaMalloc(&d1, sizeof(double) * n));
    CUDA_CHECK(cudaMemcpyAsync(d1, ones.data(), sizeof(double) * n, cudaMemcpyHostToDevice, st));
    CUDA_CHECK(cudaMemsetAsync(w, 0, sizeof(double) * n, st));
    int rc = mpet_mass_apply(ctx, space, 1.0, d1, w, stream);
    cudaStreamSynchronize(st);
    cudaFree(d1);
    if (rc) return rc;
    MPET_CATCH(ctx)
}

int mpet_apply_dirichlet_rhs(mpet_ctx* ctx, double* b, void* stream) {
    MPET_TRY(ctx)
    if (ctx->perm) {
        to_contract(ctx, b, ctx->perm_a, 0, as_stream(stream));
        scatter_bc_values(ctx, ctx->perm_a, as_stream(stream));
        from_contract(ctx, ctx->perm_a, b, 0, as_stream(stream));
    } else {
        scatter_bc_values(ctx, b, as_stream(stream));
    }
    MPET_CATCH(ctx)
}

int mpet_spmv(mpet_ctx* ctx, const double* x, double* y, void* stream) {
    MPET_TRY(ctx)
    MPET_REQUIRE(ctx->lhs_ready, "mpet_assemble_lhs must run first");
    cudaEvent_t pe = prof_begin(ctx, as_stream(stream));
    if (ctx->perm) {
        spmv_api(ctx, to_contract(ctx, x, ctx->perm_a, 0, as_stream(stream)), ctx->perm_b, as_stream(stream));
        from_contract(ctx, ctx->perm_b, y, 0, as_stream(stream));
    } else {
        spmv_api(ctx, x, y, as_stream(stream));
    }
    prof_end(ctx, PROF_VEC, pe, as_stream(stream));
    MPET_CATCH(ctx)
}

int mpet_csr_spmv(mpet_ctx* ctx, int64_t nrows, const int64_t* rowptr, const int32_t* cols,
                  const double* vals, const double* x, double* y, double beta, void* stream) {
    MPET_TRY(ctx)
    csr_spmv(ctx, nrows, rowptr, cols, vals, x, y, beta, nullptr, as_stream(stream));
    MPET_CATCH(ctx)
}

int mpet_krylov_setup(mpet_ctx* ctx, int method, int pc, double rtol, double atol, int maxit, int restart) {
    MPET_TRY(ctx)
    MPET_REQUIRE(method == 0 || method == 1, "method: 0 = MINRES, 1 = GMRES");
    MPET_REQUIRE(pc >= 0 && pc <= 2, "pc: 0 = none, 1 = Jacobi, 2 = block AMG");
    ctx->method = method;
    ctx->pc = pc;
    ctx->rtol = rtol;
    ctx->atol = atol;
    ctx->maxit = maxit;
    ctx->restart = restart > 0 ? restart : 30;
    MPET_CATCH(ctx)
}

int mpet_krylov_reference_norm(mpet_ctx* ctx, int mode) {
    MPET_TRY(ctx)
    MPET_REQUIRE(mode == 0 || mode == 1, "mode: 0 = |b| (PETSc default), 1 = min(|b|, |r0|)");
    ctx->norm_mode = mode;
    MPET_CATCH(ctx)
}

int mpet_set_border(mpet_ctx* ctx, int nb, const double* columns_dev, void* stream) {
    MPET_TRY(ctx)
    MPET_REQUIRE(ctx->N > 0, "mpet_set_mesh must be called first");
    MPET_REQUIRE(nb >= 0 && nb <= 16, "0..16 multipliers");
    cudaStream_t st = as_stream(stream);
    if (ctx->border) { dev_free(ctx, ctx->border); ctx->border = nullptr; }
    ctx->nb = nb;
    ctx->border_scaled = false;
    ctx->graph_epoch++;
    if (nb > 0) {
        ctx->border = dev_alloc<double>(ctx, (int64_t)nb * ctx->Nint);
        for (int i = 0; i < nb; ++i)
            to_internal(ctx, to_contract(ctx, columns_dev + (int64_t)i * ctx->N, ctx->perm_a, 0, st),
                        ctx->border + (int64_t)i * ctx->Nint, st);
        CUDA_CHECK(cudaStreamSynchronize(st));
    }
    MPET_CATCH(ctx)
}

int mpet_set_prec_shift(mpet_ctx* ctx, double shift_u, const double* shift_p_host) {
    MPET_TRY(ctx)
    ctx->prec_shift_u = shift_u;
    for (int i = 0; i < ctx->A; ++i) ctx->prec_shift_p[i] = shift_p_host ? shift_p_host[i] : 0.0;
    ctx->prec_ready = false;           // the P1 blocks are re-formed by the next mpet_assemble_prec
    MPET_CATCH(ctx)
}

int mpet_pc_setup(mpet_ctx* ctx, void* stream) {
    MPET_TRY(ctx)
    pc_setup(ctx, as_stream(stream));
    MPET_CATCH(ctx)
}

int mpet_solve(mpet_ctx* ctx, const double* b, double* x, double* info, void* stream) {
    MPET_TRY(ctx)
    if (ctx->perm) {       // multipliers of a bordered system (behind the N dofs) pass through unpermuted
        const double* bc = to_contract(ctx, b, ctx->perm_a, ctx->nb, as_stream(stream));
        to_contract(ctx, x, ctx->perm_b, ctx->nb, as_stream(stream));
        krylov_solve(ctx, bc, ctx->perm_b, info, as_stream(stream));
        from_contract(ctx, ctx->perm_b, x, ctx->nb, as_stream(stream));
        CUDA_CHECK(cudaStreamSynchronize(as_stream(stream)));
    } else {
        krylov_solve(ctx, b, x, info, as_stream(stream));
    }
    MPET_CATCH(ctx)
}

int mpet_pc_apply(mpet_ctx* ctx, const double* r, double* z, void* stream) {
    MPET_TRY(ctx)
    if (ctx->perm) {
        pc_apply(ctx, to_contract(ctx, r, ctx->perm_a, 0, as_stream(stream)), ctx->perm_b, as_stream(stream));
        from_contract(ctx, ctx->perm_b, z, 0, as_stream(stream));
    } else {
        pc_apply(ctx, r, z, as_stream(stream));
    }
    MPET_CATCH(ctx)
}

int mpet_attach_comm(mpet_ctx* ctx, const void* uid, int rank, int nranks) {
    MPET_TRY(ctx)
    dist_attach(ctx, uid, rank, nranks);
    MPET_CATCH(ctx)
}

int mpet_nccl_unique_id(void* out128_host) {
    try {
        dist_unique_id(out128_host);
    } catch (const std::exception&) {
        return -1;
    }
    return 0;
}

int mpet_set_halo(mpet_ctx* ctx, int n_neighbours, const int* ranks_host, const int64_t* send_off_host,
                  const int32_t* send_dofs_dev, const int64_t* recv_off_host, const int32_t* recv_dofs_dev,
                  const uint8_t* owned_dev, void* stream) {
    MPET_TRY(ctx)
    dist_set_halo(ctx, n_neighbours, ranks_host, send_off_host, send_dofs_dev, recv_off_host, recv_dofs_dev,
                  owned_dev, as_stream(stream));
    MPET_CATCH(ctx)
}

int64_t mpet_launch_count(mpet_ctx* ctx, int reset) {
    int64_t n = ctx->launches;
    if (reset) ctx->launches = 0;
    return n;
}

int64_t mpet_device_bytes(mpet_ctx* ctx) { return ctx->bytes; }

int64_t mpet_pc_bytes(mpet_ctx* ctx) { return ctx->pc_bytes_last; }

int mpet_comm_kind(mpet_ctx* ctx) { return dist_comm_kind(ctx); }

}  // extern "C"
