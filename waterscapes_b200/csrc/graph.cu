// Function-space and sparsity-graph construction, entirely on device.
//
// Replaces, for the MPET mixed space [P2]^3 x [P1]^A (mpetsolver.py:100-132):
//   * DOLFIN's edge numbering + dof map build (FunctionSpace(mesh, M), mpetsolver.py:130),
//   * SparsityPatternBuilder / PETScMatrix::init inside assemble() (mpetsolver.py:335,412,496).
//
// Design: the block system's pattern is the Kronecker-like expansion of four *scalar* node graphs
// (P2xP2, P2xP1, P1xP2, P1xP1).  Each graph is built by radix-sorting one 64-bit key per
// (cell, local pair) together with its contribution id; equal keys become one CSR entry and the
// run of contribution ids (ascending cell id, because the sort is stable) becomes that entry's
// gather list.  Assembly then *gathers* per entry in a fixed order: deterministic, no atomics.
#include "ctx.h"
#include <cub/cub.cuh>

namespace {

__constant__ int c_ledge[6][2] = {{2, 3}, {1, 3}, {1, 2}, {0, 3}, {0, 2}, {0, 1}};  // UFC local edges

__global__ void k_edge_keys(const int32_t* __restrict__ cells, int64_t nc, int64_t nv,
                            unsigned long long* __restrict__ keys) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nc * 6) return;
    int64_t c = i / 6;
    int le = (int)(i - c * 6);
    unsigned long long lo = cells[c * 4 + c_ledge[le][0]];
    unsigned long long hi = cells[c * 4 + c_ledge[le][1]];
    keys[i] = lo * (unsigned long long)nv + hi;
}

__global__ void k_edge_vertices(const unsigned long long* __restrict__ ekeys, int64_t ne, int64_t nv,
                                int32_t* __restrict__ edge_v) {
    int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    edge_v[2 * e] = (int32_t)(ekeys[e] / (unsigned long long)nv);
    edge_v[2 * e + 1] = (int32_t)(ekeys[e] % (unsigned long long)nv);
}

__device__ __forceinline__ int64_t lower_bound_u64(const unsigned long long* a, int64_t n,
                                                   unsigned long long key) {
    int64_t lo = 0, hi = n;
    while (lo < hi) {
        int64_t mid = (lo + hi) >> 1;
        if (a[mid] < key) lo = mid + 1; else hi = mid;
    }
    return lo;
}

__global__ void k_cell_nodes(const int32_t* __restrict__ cells, int64_t nc, int64_t nv,
                             const unsigned long long* __restrict__ ekeys, int64_t ne,
                             int32_t* __restrict__ cell_nodes) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nc * 10) return;
    int64_t c = i / 10;
    int a = (int)(i - c * 10);
    if (a < 4) {
        cell_nodes[i] = cells[c * 4 + a];
    } else {
        int le = a - 4;
        unsigned long long lo = cells[c * 4 + c_ledge[le][0]];
        unsigned long long hi = cells[c * 4 + c_ledge[le][1]];
        cell_nodes[i] = (int32_t)(nv + lower_bound_u64(ekeys, ne, lo * (unsigned long long)nv + hi));
    }
}

// kind 0: P2xP2 (100 pairs), 1: P2xP1 (40 pairs), 2: P1xP1 (16 pairs)
__global__ void k_pair_keys(const int32_t* __restrict__ cell_nodes, int64_t nc, int kind,
                            int64_t ncols, unsigned long long* __restrict__ keys,
                            uint32_t* __restrict__ vals) {
    int npairs = kind == 0 ? 100 : (kind == 1 ? 40 : 16);
    int nb = kind == 0 ? 10 : 4;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= nc * npairs) return;
    int64_t c = i / npairs;
    int p = (int)(i - c * npairs);
    int a = p / nb, b = p - a * nb;
    unsigned long long r = cell_nodes[c * 10 + a];
    unsigned long long col = cell_nodes[c * 10 + b];
    keys[i] = r * (unsigned long long)ncols + col;
    vals[i] = (uint32_t)i;
}

__global__ void k_head_flags(const unsigned long long* __restrict__ keys, int64_t n,
                             int32_t* __restrict__ flags) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    flags[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

__global__ void k_fill_entries(const unsigned long long* __restrict__ keys,
                               const int32_t* __restrict__ flags, const int32_t* __restrict__ pos,
                               int64_t n, int64_t ncols, int32_t* __restrict__ col,
                               int32_t* __restrict__ gptr, int32_t* __restrict__ rowcnt) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (flags[i]) {
        int32_t e = pos[i];
        unsigned long long k = keys[i];
        col[e] = (int32_t)(k % (unsigned long long)ncols);
        if (gptr) gptr[e] = (int32_t)i;
        atomicAdd(&rowcnt[(int64_t)(k / (unsigned long long)ncols)], 1);  // integer: order-independent
    }
}

__global__ void k_transpose_keys(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ col,
                                 int64_t nrows, int64_t nrows_as_cols,
                                 unsigned long long* __restrict__ keys) {
    // one thread per row (rows are short): key = col * nrows + row
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= nrows) return;
    for (int32_t j = rowptr[r]; j < rowptr[r + 1]; ++j)
        keys[j] = (unsigned long long)col[j] * (unsigned long long)nrows_as_cols + (unsigned long long)r;
}

__global__ void k_t21to12(const int32_t* __restrict__ rowptr21, const int32_t* __restrict__ col21,
                          int64_t n2, const int32_t* __restrict__ rowptr12,
                          const int32_t* __restrict__ col12, int32_t* __restrict__ t) {
    int64_t r = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (r >= n2) return;
    for (int32_t j = rowptr21[r]; j < rowptr21[r + 1]; ++j) {
        int32_t v = col21[j];
        int32_t lo = rowptr12[v], hi = rowptr12[v + 1];
        while (lo < hi) {
            int32_t mid = (lo + hi) >> 1;
            if (col12[mid] < (int32_t)r) lo = mid + 1; else hi = mid;
        }
        t[j] = lo;
    }
}

__global__ void k_row_degrees(const int32_t* __restrict__ rp22, const int32_t* __restrict__ rp21,
                              const int32_t* __restrict__ rp12, const int32_t* __restrict__ rp11,
                              int64_t n2, int64_t nv, int A, int64_t* __restrict__ deg) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    int64_t N = 3 * n2 + A * nv;
    if (i > N) return;
    if (i == N) { deg[i] = 0; return; }
    if (i < 3 * n2) {
        int64_t a = i % n2;
        deg[i] = 3 * (int64_t)(rp22[a + 1] - rp22[a]) + A * (int64_t)(rp21[a + 1] - rp21[a]);
    } else {
        int64_t v = (i - 3 * n2) % nv;
        deg[i] = 3 * (int64_t)(rp12[v + 1] - rp12[v]) + A * (int64_t)(rp11[v + 1] - rp11[v]);
    }
}

// one warp per scalar node row: expand the scalar columns into the block columns
__global__ void k_fill_cols(const int32_t* __restrict__ rpA, const int32_t* __restrict__ colA,
                            const int32_t* __restrict__ rpB, const int32_t* __restrict__ colB,
                            int64_t nnodes, int64_t row_base, int nblocks_rows, int64_t row_stride,
                            int64_t n2, int64_t nv, int A, const int64_t* __restrict__ rowptr,
                            int32_t* __restrict__ cols) {
    int64_t w = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    int lane = threadIdx.x & 31;
    if (w >= nnodes) return;
    int32_t a0 = rpA[w], dA = rpA[w + 1] - a0;   // columns in the P2 node set
    int32_t b0 = rpB[w], dB = rpB[w + 1] - b0;   // columns in the vertex set
    for (int rb = 0; rb < nblocks_rows; ++rb) {
        int64_t row = row_base + rb * row_stride + w;
        int64_t base = rowptr[row];
        for (int l = 0; l < 3; ++l)
            for (int j = lane; j < dA; j += 32)
                cols[base + (int64_t)l * dA + j] = (int32_t)(l * n2 + colA[a0 + j]);
        base += 3 * (int64_t)dA;
        for (int i = 0; i < A; ++i)
            for (int j = lane; j < dB; j += 32)
                cols[base + (int64_t)i * dB + j] = (int32_t)(3 * n2 + i * nv + colB[b0 + j]);
    }
}

int bits_for(unsigned long long maxval) {
    int b = 1;
    while (b < 64 && (maxval >> b)) ++b;
    return b;
}

struct Scratch {
    std::vector<void*> ptrs;
    template <typename T> T* get(int64_t n) {
        void* p = nullptr;
        CUDA_CHECK(cudaMalloc(&p, (size_t)(n > 0 ? n : 1) * sizeof(T)));
        ptrs.push_back(p);
        return (T*)p;
    }
    void release() {
        for (void* p : ptrs) cudaFree(p);
        ptrs.clear();
    }
    ~Scratch() { release(); }
};

// sort keys (with optional payload) and compress equal keys into a NodeGraph
void compress_sorted(mpet_ctx* ctx, NodeGraph& g, const unsigned long long* keys, int64_t n,
                     bool want_gather, const uint32_t* sorted_vals, cudaStream_t st) {
    Scratch s;
    int32_t* flags = s.get<int32_t>(n);
    int32_t* pos = s.get<int32_t>(n + 1);
    k_head_flags<<<grid_for(n, 256), 256, 0, st>>>(keys, n, flags);
    LAUNCH_CHECK(ctx);
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, flags, pos, n, st);
    void* tmp = s.get<char>((int64_t)tb);
    cub::DeviceScan::ExclusiveSum(tmp, tb, flags, pos, n, st);
    int32_t last_pos = 0, last_flag = 0;
    CUDA_CHECK(cudaMemcpyAsync(&last_pos, pos + n - 1, 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaMemcpyAsync(&last_flag, flags + n - 1, 4, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    g.nnz = (int64_t)last_pos + last_flag;
    g.col = dev_alloc<int32_t>(ctx, g.nnz);
    g.rowptr = dev_alloc<int32_t>(ctx, g.nrows + 1);
    int32_t* rowcnt = s.get<int32_t>(g.nrows + 1);
    CUDA_CHECK(cudaMemsetAsync(rowcnt, 0, (g.nrows + 1) * 4, st));
    if (want_gather) {
        g.gptr = dev_alloc<int32_t>(ctx, g.nnz + 1);
        g.glist = dev_alloc<uint32_t>(ctx, n);
        g.ncontrib = n;
        CUDA_CHECK(cudaMemcpyAsync(g.glist, sorted_vals, n * 4, cudaMemcpyDeviceToDevice, st));
        int32_t n32 = (int32_t)n;
        CUDA_CHECK(cudaMemcpyAsync(g.gptr + g.nnz, &n32, 4, cudaMemcpyHostToDevice, st));
    }
    k_fill_entries<<<grid_for(n, 256), 256, 0, st>>>(keys, flags, pos, n, g.ncols, g.col,
                                                     want_gather ? g.gptr : nullptr, rowcnt);
    LAUNCH_CHECK(ctx);
    size_t tb2 = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb2, rowcnt, g.rowptr, g.nrows + 1, st);
    void* tmp2 = s.get<char>((int64_t)tb2);
    cub::DeviceScan::ExclusiveSum(tmp2, tb2, rowcnt, g.rowptr, g.nrows + 1, st);
    CUDA_CHECK(cudaStreamSynchronize(st));
}

void build_pair_graph(mpet_ctx* ctx, NodeGraph& g, int kind, cudaStream_t st) {
    int npairs = kind == 0 ? 100 : (kind == 1 ? 40 : 16);
    int64_t n = ctx->Nc * npairs;
    MPET_REQUIRE(n < 2147483647LL, "mesh too large for 32-bit contribution ids (shard it across GPUs)");
    g.nrows = (kind == 2) ? ctx->Nv : ctx->N2;
    g.ncols = (kind == 0) ? ctx->N2 : ctx->Nv;
    Scratch s;
    unsigned long long* k0 = s.get<unsigned long long>(n);
    unsigned long long* k1 = s.get<unsigned long long>(n);
    uint32_t* v0 = s.get<uint32_t>(n);
    uint32_t* v1 = s.get<uint32_t>(n);
    // P1 nodes are the first 4 entries of cell_nodes rows, so kind 1/2 index the same array
    k_pair_keys<<<grid_for(n, 256), 256, 0, st>>>(ctx->cell_nodes, ctx->Nc, kind, g.ncols, k0, v0);
    LAUNCH_CHECK(ctx);
    cub::DoubleBuffer<unsigned long long> dk(k0, k1);
    cub::DoubleBuffer<uint32_t> dv(v0, v1);
    int bits = bits_for((unsigned long long)g.nrows * (unsigned long long)g.ncols);
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, dk, dv, n, 0, bits, st);
    void* tmp = s.get<char>((int64_t)tb);
    cub::DeviceRadixSort::SortPairs(tmp, tb, dk, dv, n, 0, bits, st);
    compress_sorted(ctx, g, dk.Current(), n, true, dv.Current(), st);
}

}  // namespace

void dev_free(mpet_ctx* ctx, void* p) {
    if (!p) return;
    for (size_t i = 0; i < ctx->allocs.size(); ++i)
        if (ctx->allocs[i] == p) {
            ctx->allocs[i] = ctx->allocs.back();
            ctx->allocs.pop_back();
            break;
        }
    cudaFree(p);
}

void build_space(mpet_ctx* ctx, cudaStream_t st) {
    const int64_t nv = ctx->Nv, nc = ctx->Nc;
    const int A = ctx->A;
    // ---- edges: lexicographic numbering of sorted vertex pairs
    {
        Scratch s;
        int64_t n = nc * 6;
        unsigned long long* k0 = s.get<unsigned long long>(n);
        unsigned long long* k1 = s.get<unsigned long long>(n);
        unsigned long long* uq = s.get<unsigned long long>(n);
        int64_t* d_num = s.get<int64_t>(1);
        k_edge_keys<<<grid_for(n, 256), 256, 0, st>>>(ctx->cells, nc, nv, k0);
        LAUNCH_CHECK(ctx);
        cub::DoubleBuffer<unsigned long long> dk(k0, k1);
        int bits = bits_for((unsigned long long)nv * (unsigned long long)nv);
        size_t tb = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, tb, dk, n, 0, bits, st);
        void* tmp = s.get<char>((int64_t)tb);
        cub::DeviceRadixSort::SortKeys(tmp, tb, dk, n, 0, bits, st);
        size_t tb2 = 0;
        cub::DeviceSelect::Unique(nullptr, tb2, dk.Current(), uq, d_num, n, st);
        void* tmp2 = s.get<char>((int64_t)tb2);
        cub::DeviceSelect::Unique(tmp2, tb2, dk.Current(), uq, d_num, n, st);
        int64_t ne = 0;
        CUDA_CHECK(cudaMemcpyAsync(&ne, d_num, 8, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        ctx->Ne = ne;
        ctx->N2 = nv + ne;
        ctx->edge_v = dev_alloc<int32_t>(ctx, 2 * ne);
        ctx->cell_nodes = dev_alloc<int32_t>(ctx, nc * 10);
        k_edge_vertices<<<grid_for(ne, 256), 256, 0, st>>>(uq, ne, nv, ctx->edge_v);
        LAUNCH_CHECK(ctx);
        k_cell_nodes<<<grid_for(nc * 10, 256), 256, 0, st>>>(ctx->cells, nc, nv, uq, ne, ctx->cell_nodes);
        LAUNCH_CHECK(ctx);
        CUDA_CHECK(cudaStreamSynchronize(st));
    }
    const int64_t n2 = ctx->N2;
    ctx->N = 3 * n2 + A * nv;
    ctx->nloc = 30 + 4 * A;
    ctx->Nint = 4 * n2 + (int64_t)A * nv;
    MPET_REQUIRE(ctx->N < 2147483647LL, "more than 2^31 dofs on one GPU");

    // ---- scalar node graphs with gather lists
    build_pair_graph(ctx, ctx->g22, 0, st);
    build_pair_graph(ctx, ctx->g21, 1, st);
    build_pair_graph(ctx, ctx->g11, 2, st);
    // ---- g12 = transpose pattern of g21
    {
        NodeGraph& g = ctx->g12;
        g.nrows = nv;
        g.ncols = n2;
        Scratch s;
        int64_t n = ctx->g21.nnz;
        unsigned long long* k0 = s.get<unsigned long long>(n);
        unsigned long long* k1 = s.get<unsigned long long>(n);
        k_transpose_keys<<<grid_for(n2, 128), 128, 0, st>>>(ctx->g21.rowptr, ctx->g21.col, n2, n2, k0);
        LAUNCH_CHECK(ctx);
        cub::DoubleBuffer<unsigned long long> dk(k0, k1);
        int bits = bits_for((unsigned long long)nv * (unsigned long long)n2);
        size_t tb = 0;
        cub::DeviceRadixSort::SortKeys(nullptr, tb, dk, n, 0, bits, st);
        void* tmp = s.get<char>((int64_t)tb);
        cub::DeviceRadixSort::SortKeys(tmp, tb, dk, n, 0, bits, st);
        compress_sorted(ctx, g, dk.Current(), n, false, nullptr, st);
        MPET_REQUIRE(g.nnz == n, "transpose graph size mismatch");
        ctx->t21to12 = dev_alloc<int32_t>(ctx, n);
        k_t21to12<<<grid_for(n2, 128), 128, 0, st>>>(ctx->g21.rowptr, ctx->g21.col, n2, g.rowptr, g.col,
                                                     ctx->t21to12);
        LAUNCH_CHECK(ctx);
    }
    // ---- block-system CSR
    {
        Scratch s;
        int64_t N = ctx->N;
        int64_t* deg = s.get<int64_t>(N + 1);
        k_row_degrees<<<grid_for(N + 1, 256), 256, 0, st>>>(ctx->g22.rowptr, ctx->g21.rowptr,
                                                            ctx->g12.rowptr, ctx->g11.rowptr, n2, nv, A, deg);
        LAUNCH_CHECK(ctx);
        ctx->rowptr = dev_alloc<int64_t>(ctx, N + 1);
        size_t tb = 0;
        cub::DeviceScan::ExclusiveSum(nullptr, tb, deg, ctx->rowptr, N + 1, st);
        void* tmp = s.get<char>((int64_t)tb);
        cub::DeviceScan::ExclusiveSum(tmp, tb, deg, ctx->rowptr, N + 1, st);
        CUDA_CHECK(cudaMemcpyAsync(&ctx->nnz, ctx->rowptr + N, 8, cudaMemcpyDeviceToHost, st));
        CUDA_CHECK(cudaStreamSynchronize(st));
        ctx->cols = dev_alloc<int32_t>(ctx, ctx->nnz);
        ctx->vals = dev_alloc<double>(ctx, ctx->nnz);
        int threads = 256;
        k_fill_cols<<<grid_for(n2 * 32, threads), threads, 0, st>>>(
            ctx->g22.rowptr, ctx->g22.col, ctx->g21.rowptr, ctx->g21.col, n2, 0, 3, n2, n2, nv, A,
            ctx->rowptr, ctx->cols);
        LAUNCH_CHECK(ctx);
        if (A > 0) {
            k_fill_cols<<<grid_for(nv * 32, threads), threads, 0, st>>>(
                ctx->g12.rowptr, ctx->g12.col, ctx->g11.rowptr, ctx->g11.col, nv, 3 * n2, A, nv, n2, nv, A,
                ctx->rowptr, ctx->cols);
            LAUNCH_CHECK(ctx);
        }
        CUDA_CHECK(cudaStreamSynchronize(st));
    }
    ctx->m11 = dev_alloc<double>(ctx, ctx->g11.nnz);
    ctx->l11 = dev_alloc<double>(ctx, ctx->g11.nnz);
}
