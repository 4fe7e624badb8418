// Multi-GPU: one process per GPU, mesh partitioned by cells with owned / ghost dofs.
//
// Replaces what DOLFIN + PETSc do under `mpirun` (utils/jobscript.sh:43): VecScatter halo updates in
// MatMult and MPI_Allreduce in the Krylov dot products.  Here every rank holds its own cells plus a
// one-cell ghost layer, so its OWNED rows are complete without any assembly collective; vectors carry
// owned + ghost entries.  Two collectives remain, both over NCCL (NVLink 5 / NVSwitch):
//   * halo exchange: pack owned boundary values -> ncclSend/ncclRecv inside one group -> unpack into
//     the ghost entries (forward), or ghost contributions added into the owner (reverse, used by the
//     additive-Schwarz application of the per-rank V-cycles);
//   * all-reduce of the 1-3 Krylov scalars per iteration (in place on the device scalar).
// NCCL is resolved at run time from the library already loaded in the process (torch's bundled
// libnccl.so.2), so libmpet_b200.so has no link-time dependency on it.
#include "ctx.h"
#include "layout.cuh"
#include <dlfcn.h>
#include <cstring>

namespace {

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { NCCL_INT8 = 0, NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2 };

struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*) = nullptr;     // optional (NCCL >= 2.18)
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
} g_nccl;

void load_nccl() {
    if (g_nccl.handle) return;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
        g_nccl.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.handle) break;
    }
    if (!g_nccl.handle) throw MpetError("NCCL (libnccl.so.2) is not loadable in this process");
#define RESOLVE(field, sym)                                                        \
    *(void**)(&g_nccl.field) = dlsym(g_nccl.handle, sym);                            \
    if (!g_nccl.field) throw MpetError("NCCL symbol " sym " not found");
    RESOLVE(GetUniqueId, "ncclGetUniqueId")
    RESOLVE(CommInitRank, "ncclCommInitRank")
    RESOLVE(CommDestroy, "ncclCommDestroy")
    RESOLVE(GroupStart, "ncclGroupStart")
    RESOLVE(GroupEnd, "ncclGroupEnd")
    RESOLVE(Send, "ncclSend")
    RESOLVE(Recv, "ncclRecv")
    RESOLVE(AllReduce, "ncclAllReduce")
    RESOLVE(AllGather, "ncclAllGather")
    RESOLVE(Broadcast, "ncclBroadcast")
    RESOLVE(GetErrorString, "ncclGetErrorString")
#undef RESOLVE
    *(void**)(&g_nccl.CommSplit) = dlsym(g_nccl.handle, "ncclCommSplit");
}

#define NCCL_CHECK(call)                                                                     \
    do {                                                                                     \
        ncclResult_t r__ = (call);                                                           \
        if (r__ != 0) {                                                                      \
            char buf__[256];                                                                 \
            snprintf(buf__, sizeof buf__, "%s:%d: NCCL: %s", __FILE__, __LINE__, g_nccl.GetErrorString(r__)); \
            throw MpetError(buf__);                                                          \
        }                                                                                    \
    } while (0)

__global__ void k_pack(const int32_t* __restrict__ idx, int64_t n, const double* __restrict__ v,
                       double* __restrict__ buf, const int* __restrict__ done) {
    if (done && *done) return;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) buf[i] = v[idx[i]];
}

__global__ void k_unpack(const int32_t* __restrict__ idx, int64_t n, const double* __restrict__ buf,
                         double* __restrict__ v, int add, const int* __restrict__ done) {
    if (done && *done) return;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (add) v[idx[i]] += buf[i];   // indices of one message are unique, messages are applied in rank order
    else v[idx[i]] = buf[i];
}

__global__ void k_api_to_internal_idx(const int32_t* __restrict__ in, int64_t n, int64_t n2, int32_t* __restrict__ out) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = (int32_t)api_to_internal(in[i], n2);
}

__global__ void k_count_holders(const int32_t* __restrict__ idx, int64_t n, double* __restrict__ w) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&w[idx[i]], 1.0);       // integer-valued: order independent
}

__global__ void k_fill(double* __restrict__ w, int64_t n, double v) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) w[i] = v;
}

__global__ void k_rsqrt(double* __restrict__ w, int64_t n) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) w[i] = 1.0 / sqrt(w[i]);
}

__global__ void k_scale(const double* __restrict__ w, const double* __restrict__ in, double* __restrict__ out,
                        int64_t n, const int* __restrict__ done) {
    if (done && *done) return;
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) out[i] = w[i] * in[i];
}

__global__ void k_own_internal(const uint8_t* __restrict__ own_api, int64_t n, int64_t n2, uint8_t* __restrict__ own_int) {
    int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
    if (i < n) own_int[api_to_internal(i, n2)] = own_api[i];
}

}  // namespace

struct HaloPlan {
    std::vector<int64_t> send_off, recv_off;   // [nnbr+1]
    int32_t* send_idx = nullptr;               // entries of owned dofs a neighbour ghosts
    int32_t* recv_idx = nullptr;               // entries of my ghost dofs
};

struct DistState {
    ncclComm_t comm = nullptr;
    // second communicator + staging buffers ("lane 1"): the P1-field cycles of the preconditioner exchange their
    // halos on their own stream while the displacement cycle runs on the main one (both are latency-bound)
    ncclComm_t comm2 = nullptr;
    double* send_buf2 = nullptr;
    double* recv_buf2 = nullptr;
    int rank = 0, nranks = 1;
    std::vector<int> nbr;                 // neighbour ranks (ascending)
    HaloPlan plan[DIST_NPLANS];           // Krylov vectors, P2 x W4, P1 x W4, P1 x W1 level vectors
    double* send_buf = nullptr;
    double* recv_buf = nullptr;
    uint8_t* own = nullptr;               // [Nint] 1 on owned dofs (pad lanes 0)
    std::vector<uint8_t> own_nodes;       // host copy: ownership of the scalar P2 nodes (vertices first)
};

void dist_unique_id(void* out128) {
    load_nccl();
    ncclUniqueId id;
    NCCL_CHECK(g_nccl.GetUniqueId(&id));
    memcpy(out128, &id, 128);
}

void dist_attach(mpet_ctx* ctx, const void* uid, int rank, int nranks) {
    load_nccl();
    MPET_REQUIRE(ctx->dist == nullptr, "communicator already attached");
    DistState* d = new DistState();
    d->rank = rank;
    d->nranks = nranks;
    ncclUniqueId id;
    memcpy(&id, uid, 128);
    CUDA_CHECK(cudaSetDevice(ctx->device));
    NCCL_CHECK(g_nccl.CommInitRank(&d->comm, nranks, id, rank));
    if (g_nccl.CommSplit && !getenv("MPET_NO_LANE1")) {
        if (g_nccl.CommSplit(d->comm, 0, rank, &d->comm2, nullptr) != 0) d->comm2 = nullptr;
    }
    ctx->dist = d;
}

void dist_free(mpet_ctx* ctx) {
    if (!ctx->dist) return;
    if (ctx->dist->comm2) g_nccl.CommDestroy(ctx->dist->comm2);
    if (ctx->dist->comm) g_nccl.CommDestroy(ctx->dist->comm);
    delete ctx->dist;
    ctx->dist = nullptr;
}

// Node lists arrive per neighbour (scalar P2 node ids of the LOCAL mesh, vertices are ids < Nv), both
// sides listing a shared node in the same order.  Four index plans are derived from them.
void dist_set_halo(mpet_ctx* ctx, int nnbr, const int* ranks, const int64_t* send_off, const int32_t* send_nodes_dev,
                   const int64_t* recv_off, const int32_t* recv_nodes_dev, const uint8_t* owned_nodes_dev,
                   cudaStream_t st) {
    MPET_REQUIRE(ctx->dist != nullptr, "mpet_attach_comm must be called first");
    MPET_REQUIRE(ctx->N > 0, "mpet_set_mesh must be called first");
    DistState* d = ctx->dist;
    CUDA_CHECK(cudaStreamSynchronize(st));
    d->nbr.assign(ranks, ranks + nnbr);
    const int64_t n2 = ctx->N2, nv = ctx->Nv;
    const int A = ctx->A;
    std::vector<int32_t> sn(send_off[nnbr]), rn(recv_off[nnbr]);
    if (!sn.empty()) CUDA_CHECK(cudaMemcpy(sn.data(), send_nodes_dev, sizeof(int32_t) * sn.size(), cudaMemcpyDeviceToHost));
    if (!rn.empty()) CUDA_CHECK(cudaMemcpy(rn.data(), recv_nodes_dev, sizeof(int32_t) * rn.size(), cudaMemcpyDeviceToHost));
    d->own_nodes.resize(n2);
    CUDA_CHECK(cudaMemcpy(d->own_nodes.data(), owned_nodes_dev, n2, cudaMemcpyDeviceToHost));
    int64_t maxlen = 1;
    for (int pl = 0; pl < DIST_NPLANS; ++pl) {
        HaloPlan& P = d->plan[pl];
        auto build = [&](const std::vector<int32_t>& nodes, const int64_t* off, std::vector<int64_t>& poff,
                         int32_t** dev) {
            std::vector<int32_t> idx;
            poff.assign(nnbr + 1, 0);
            for (int q = 0; q < nnbr; ++q) {
                for (int64_t t = off[q]; t < off[q + 1]; ++t) {
                    const int32_t a = nodes[t];
                    const bool vertex = a < nv;
                    switch (pl) {
                        case DIST_PLAN_KRYLOV:
                        case DIST_PLAN_P2W4: for (int k = 0; k < 3; ++k) idx.push_back(4 * a + k); break;
                        case DIST_PLAN_P1W4: if (vertex) for (int k = 0; k < 3; ++k) idx.push_back(4 * a + k); break;
                        case DIST_PLAN_P1W1: if (vertex) idx.push_back(a); break;
                        default: break;
                    }
                }
                if (pl == DIST_PLAN_P1WA)
                    for (int i = 0; i < A; ++i)
                        for (int64_t t = off[q]; t < off[q + 1]; ++t)
                            if (nodes[t] < nv) idx.push_back((int32_t)((int64_t)i * nv + nodes[t]));
                if (pl == DIST_PLAN_KRYLOV)
                    for (int i = 0; i < A; ++i)
                        for (int64_t t = off[q]; t < off[q + 1]; ++t)
                            if (nodes[t] < nv) idx.push_back((int32_t)(4 * n2 + (int64_t)i * nv + nodes[t]));
                poff[q + 1] = (int64_t)idx.size();
            }
            *dev = dev_alloc<int32_t>(ctx, (int64_t)idx.size());
            if (!idx.empty()) CUDA_CHECK(cudaMemcpy(*dev, idx.data(), sizeof(int32_t) * idx.size(), cudaMemcpyHostToDevice));
            maxlen = std::max<int64_t>(maxlen, (int64_t)idx.size());
        };
        build(sn, send_off, P.send_off, &P.send_idx);
        build(rn, recv_off, P.recv_off, &P.recv_idx);
    }
    d->send_buf = dev_alloc<double>(ctx, maxlen);
    d->recv_buf = dev_alloc<double>(ctx, maxlen);
    if (d->comm2) {
        d->send_buf2 = dev_alloc<double>(ctx, maxlen);
        d->recv_buf2 = dev_alloc<double>(ctx, maxlen);
    }
    // ownership mask of the Krylov vectors (solver-internal layout; pad lanes stay 0)
    std::vector<uint8_t> own(ctx->Nint, 0);
    for (int64_t a = 0; a < n2; ++a)
        if (d->own_nodes[a]) {
            own[4 * a] = own[4 * a + 1] = own[4 * a + 2] = 1;
            if (a < nv)
                for (int i = 0; i < A; ++i) own[4 * n2 + (int64_t)i * nv + a] = 1;
        }
    d->own = dev_alloc<uint8_t>(ctx, ctx->Nint);
    CUDA_CHECK(cudaMemcpy(d->own, own.data(), ctx->Nint, cudaMemcpyHostToDevice));
}

bool dist_active(mpet_ctx* ctx) { return ctx->dist != nullptr && ctx->dist->nranks > 1; }
const uint8_t* dist_owned_mask(mpet_ctx* ctx) { return dist_active(ctx) ? ctx->dist->own : nullptr; }

// forward: owners -> ghosts (copy).  reverse: ghosts -> owners (add).
void dist_halo(mpet_ctx* ctx, int plan, double* v, bool reverse, const int* done, cudaStream_t st) {
    if (!dist_active(ctx)) return;
    DistState* d = ctx->dist;
    const HaloPlan& P = d->plan[plan];
    const int nn = (int)d->nbr.size();
    const int32_t* pack_idx = reverse ? P.recv_idx : P.send_idx;
    const int32_t* unpack_idx = reverse ? P.send_idx : P.recv_idx;
    const std::vector<int64_t>& poff = reverse ? P.recv_off : P.send_off;
    const std::vector<int64_t>& uoff = reverse ? P.send_off : P.recv_off;
    const int64_t np = poff[nn], nu = uoff[nn];
    const bool lane1 = ctx->dist_lane == 1 && d->comm2 != nullptr;
    cudaEvent_t pe = lane1 ? nullptr : prof_begin(ctx, st);      // main-stream exchanges are on the critical path
    ncclComm_t comm = lane1 ? d->comm2 : d->comm;
    double* sbuf = lane1 ? d->send_buf2 : d->send_buf;
    double* rbuf = lane1 ? d->recv_buf2 : d->recv_buf;
    if (np) { k_pack<<<grid_for(np, 256), 256, 0, st>>>(pack_idx, np, v, sbuf, done); LAUNCH_CHECK(ctx); }
    NCCL_CHECK(g_nccl.GroupStart());
    for (int q = 0; q < nn; ++q) {
        const int64_t cs = poff[q + 1] - poff[q], cr = uoff[q + 1] - uoff[q];
        if (cs) NCCL_CHECK(g_nccl.Send(sbuf + poff[q], (size_t)cs, NCCL_FLOAT64, d->nbr[q], comm, st));
        if (cr) NCCL_CHECK(g_nccl.Recv(rbuf + uoff[q], (size_t)cr, NCCL_FLOAT64, d->nbr[q], comm, st));
    }
    NCCL_CHECK(g_nccl.GroupEnd());
    if (reverse) {
        for (int q = 0; q < nn; ++q) {      // rank order: deterministic sums
            const int64_t cr = uoff[q + 1] - uoff[q];
            if (cr) { k_unpack<<<grid_for(cr, 256), 256, 0, st>>>(unpack_idx + uoff[q], cr, rbuf + uoff[q], v, 1, done); LAUNCH_CHECK(ctx); }
        }
    } else if (nu) {
        k_unpack<<<grid_for(nu, 256), 256, 0, st>>>(unpack_idx, nu, rbuf, v, 0, done);
        LAUNCH_CHECK(ctx);
    }
    prof_end(ctx, PROF_COMM, pe, st);
}

int dist_rank(mpet_ctx* ctx) { return ctx->dist ? ctx->dist->rank : 0; }
int dist_nranks(mpet_ctx* ctx) { return ctx->dist ? ctx->dist->nranks : 1; }
const std::vector<uint8_t>& dist_own_nodes(mpet_ctx* ctx) { return ctx->dist->own_nodes; }

// equal-count all-gather of doubles (device buffers)
void dist_allgather(mpet_ctx* ctx, const double* send, double* recv, int64_t count, cudaStream_t st) {
    DistState* d = ctx->dist;
    const bool lane1 = ctx->dist_lane == 1 && d->comm2;
    cudaEvent_t pe = lane1 ? nullptr : prof_begin(ctx, st);
    NCCL_CHECK(g_nccl.AllGather(send, recv, (size_t)count, NCCL_FLOAT64, lane1 ? d->comm2 : d->comm, st));
    prof_end(ctx, PROF_COMM, pe, st);
}
bool dist_has_lane1(mpet_ctx* ctx) { return dist_active(ctx) && ctx->dist->comm2 != nullptr; }
// broadcast of raw bytes (device buffer) from `root`
void dist_bcast_bytes(mpet_ctx* ctx, void* buf, int64_t nbytes, int root, cudaStream_t st) {
    NCCL_CHECK(g_nccl.Broadcast(buf, buf, (size_t)nbytes, NCCL_INT8, root, ctx->dist->comm, st));
}
void dist_allreduce_max(mpet_ctx* ctx, double* dev_scalars, int count, cudaStream_t st) {
    if (!dist_active(ctx)) return;
    NCCL_CHECK(g_nccl.AllReduce(dev_scalars, dev_scalars, (size_t)count, NCCL_FLOAT64, NCCL_MAX, ctx->dist->comm, st));
}

void dist_allreduce_sum(mpet_ctx* ctx, double* dev_scalars, int count, cudaStream_t st) {
    if (!dist_active(ctx)) return;
    cudaEvent_t pe = prof_begin(ctx, st);
    NCCL_CHECK(g_nccl.AllReduce(dev_scalars, dev_scalars, (size_t)count, NCCL_FLOAT64, NCCL_SUM, ctx->dist->comm, st));
    prof_end(ctx, PROF_COMM, pe, st);
}
