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

// =====================================================================================================
// Peer-memory channel: halo exchange, small all-reduce and all-gather as ONE kernel each, over NVLink.
//
// Every rank owns an IPC-shared arena (cudaIpcGetMemHandle / cudaIpcOpenMemHandle; handles travel once through an
// NCCL all-gather).  Per lane (lane 0: main stream, lane 1: the P1-field cycles' side stream) the arena holds
//   * flags   flag[kind][src]  = last epoch rank `src` finished pushing to me (monotone, 64-bit),
//   * all-reduce slots red[parity][src][j],
//   * staging areas (halo and gather, each double-buffered by epoch parity).
// A halo kernel (a) gathers the owned boundary values and STORES them straight into the neighbours' staging areas
// over NVLink, (b) fences at system scope; the last CTA to arrive releases one flag per neighbour, (c) every CTA
// acquires the neighbours' flags of this epoch, (d) scatters its own staging area into the ghost entries, (e) the
// last CTA to finish advances the lane's epoch counter.  The epoch lives in device memory, so the launch has no
// per-call host state (CUDA-graph friendly) and a launch skipped by the solver's `done` flag -- on every rank
// alike, the flag derives from all-reduced scalars -- leaves all ranks consistent.
// Why double buffering is enough: rank A writes epoch e+2 into B's buffer of parity(e) only after A's kernel e+1
// acquired B's flag e+1, which B released in its kernel e+1, i.e. after B's kernel e (the reader of that buffer)
// had completed in stream order.  Grids never exceed one CTA per SM, so all CTAs of a kernel are co-resident and
// the spin in (c) cannot starve (b).
// Replaces k_pack -> ncclGroup{Send,Recv} -> k_unpack (3 launches + NCCL's proxy latency per exchange).
constexpr int kLanes = 2;
constexpr int kMaxRanks = 16;
constexpr int kRedMax = 128;
constexpr int kMaxNbr = 64;
enum { PK_HALO = 0, PK_GATHER = 1, PK_REDUCE = 2, PK_KINDS = 3 };
constexpr unsigned long long kSpinTimeoutNs = 60ull * 1000ull * 1000ull * 1000ull;

struct PeerShared {
    unsigned long long flag[PK_KINDS][kMaxRanks];
    double red[2][kMaxRanks][kRedMax];
};
struct LaneLocal {
    unsigned long long epoch[PK_KINDS];
    unsigned int arrive[PK_KINDS][2];
    int error;
    int pad;
};
struct PeerLaneDev {
    PeerShared* hdr[kMaxRanks];
    double* halo[kMaxRanks];      // halo staging base of every rank (2 * halo_cap doubles)
    double* gather[kMaxRanks];    // gather staging base of every rank (2 * gather_cap doubles)
    LaneLocal* local;
    int64_t halo_cap, gather_cap;
    int rank, nranks;
};
struct HaloArgs {
    const int32_t* send_idx;
    const int32_t* recv_idx;
    const int64_t* send_off;
    const int64_t* recv_off;
    const int64_t* dst_off;       // where my message starts in the neighbour's staging area
    const int* nbr;
    int nn;
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// spin until *flag >= e; a dead peer ends in an error code instead of a hung GPU
__device__ __forceinline__ void wait_flag(const unsigned long long* flag, unsigned long long e, int* err) {
    if (ld_acquire_sys(flag) >= e) return;
    const unsigned long long t0 = global_ns();
    while (ld_acquire_sys(flag) < e) {
        __nanosleep(64);
        if (global_ns() - t0 > kSpinTimeoutNs) { atomicExch(err, 1); return; }
    }
}

__global__ void __launch_bounds__(256)
k_peer_halo(PeerLaneDev L, HaloArgs H, double* __restrict__ v, const int* __restrict__ done) {
    if (done && *done) return;
    LaneLocal* loc = L.local;
    const unsigned long long e = *(volatile unsigned long long*)&loc->epoch[PK_HALO] + 1ull;
    const int64_t base = (int64_t)(e & 1ull) * L.halo_cap;
    const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, gsize = (int64_t)gridDim.x * blockDim.x;
    __shared__ int s_last;
    // (a) push my boundary values into the neighbours' staging areas (NVLink stores)
    for (int q = 0; q < H.nn; ++q) {
        double* dst = L.halo[H.nbr[q]] + base + H.dst_off[q];
        const int64_t s0 = H.send_off[q], s1 = H.send_off[q + 1];
        for (int64_t i = s0 + gtid; i < s1; i += gsize) dst[i - s0] = v[H.send_idx[i]];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&loc->arrive[PK_HALO][0], 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {      // (b) all CTAs have fenced their stores: release one flag per neighbour
        __threadfence_system();
        if ((int)threadIdx.x < H.nn) st_release_sys(&L.hdr[H.nbr[threadIdx.x]]->flag[PK_HALO][L.rank], e);
        if (threadIdx.x == 0) loc->arrive[PK_HALO][0] = 0u;
    }
    // (c) wait for the neighbours' messages of this epoch
    if ((int)threadIdx.x < H.nn) wait_flag(&L.hdr[L.rank]->flag[PK_HALO][H.nbr[threadIdx.x]], e, &loc->error);
    __syncthreads();
    // (d) my staging area -> ghost entries (L2 is the coherence point of remote writes: bypass L1)
    const double* src = L.halo[L.rank] + base;
    const int64_t nu = H.recv_off[H.nn];
    for (int64_t i = gtid; i < nu; i += gsize) v[H.recv_idx[i]] = __ldcg(src + i);
    // (e) the last CTA to finish advances the epoch
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&loc->arrive[PK_HALO][1], 1u) == gridDim.x - 1) {
            loc->arrive[PK_HALO][1] = 0u;
            *(volatile unsigned long long*)&loc->epoch[PK_HALO] = e;
            __threadfence();
        }
    }
}

// Sum over ranks of `count` (<= kRedMax) doubles, in rank order on every rank (bit-identical results everywhere).
// from_partials: value j = sum of partials[j*nparts .. (j+1)*nparts) in a fixed order (the second stage of the
// Krylov dot products), else value j = inout[j].  One CTA.
__global__ void __launch_bounds__(256)
k_peer_allreduce(PeerLaneDev L, const double* __restrict__ partials, int nparts, int count, double* __restrict__ inout,
                 const int* __restrict__ done) {
    if (done && *done) return;
    LaneLocal* loc = L.local;
    __shared__ double s_val[kRedMax];
    __shared__ double s_red[32];
    const unsigned long long e = *(volatile unsigned long long*)&loc->epoch[PK_REDUCE] + 1ull;
    const int par = (int)(e & 1ull);
    if (partials) {
        for (int j = 0; j < count; ++j) {
            double s = 0.0;
            for (int i = threadIdx.x; i < nparts; i += blockDim.x) s += partials[(int64_t)j * nparts + i];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
            if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = s;
            __syncthreads();
            if (threadIdx.x < 32) {
                s = threadIdx.x < (blockDim.x >> 5) ? s_red[threadIdx.x] : 0.0;
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
                if (threadIdx.x == 0) s_val[j] = s;
            }
            __syncthreads();
        }
    } else {
        for (int j = threadIdx.x; j < count; j += blockDim.x) s_val[j] = inout[j];
        __syncthreads();
    }
    for (int t = threadIdx.x; t < L.nranks * count; t += blockDim.x) {
        const int r = t / count, j = t - r * count;
        L.hdr[r]->red[par][L.rank][j] = s_val[j];
    }
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < L.nranks) {
        st_release_sys(&L.hdr[threadIdx.x]->flag[PK_REDUCE][L.rank], e);
        wait_flag(&L.hdr[L.rank]->flag[PK_REDUCE][threadIdx.x], e, &loc->error);
    }
    __syncthreads();
    for (int j = threadIdx.x; j < count; j += blockDim.x) {
        double s = 0.0;
        for (int r = 0; r < L.nranks; ++r) s += __ldcg(&L.hdr[L.rank]->red[par][r][j]);
        inout[j] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) *(volatile unsigned long long*)&loc->epoch[PK_REDUCE] = e;
}

// recv[r*count + i] = send_r[i] for every rank r
__global__ void __launch_bounds__(256)
k_peer_allgather(PeerLaneDev L, const double* __restrict__ send, double* __restrict__ recv, int64_t count,
                 const int* __restrict__ done) {
    if (done && *done) return;
    LaneLocal* loc = L.local;
    const unsigned long long e = *(volatile unsigned long long*)&loc->epoch[PK_GATHER] + 1ull;
    const int64_t base = (int64_t)(e & 1ull) * L.gather_cap;
    const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, gsize = (int64_t)gridDim.x * blockDim.x;
    __shared__ int s_last;
    for (int r = 0; r < L.nranks; ++r) {
        double* dst = L.gather[r] + base + (int64_t)L.rank * count;
        for (int64_t i = gtid; i < count; i += gsize) dst[i] = send[i];
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) s_last = (atomicAdd(&loc->arrive[PK_GATHER][0], 1u) == gridDim.x - 1) ? 1 : 0;
    __syncthreads();
    if (s_last) {
        __threadfence_system();
        if ((int)threadIdx.x < L.nranks) st_release_sys(&L.hdr[threadIdx.x]->flag[PK_GATHER][L.rank], e);
        if (threadIdx.x == 0) loc->arrive[PK_GATHER][0] = 0u;
    }
    if ((int)threadIdx.x < L.nranks) wait_flag(&L.hdr[L.rank]->flag[PK_GATHER][threadIdx.x], e, &loc->error);
    __syncthreads();
    const double* src = L.gather[L.rank] + base;
    const int64_t total = count * L.nranks;
    for (int64_t i = gtid; i < total; i += gsize) recv[i] = __ldcg(src + i);
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        if (atomicAdd(&loc->arrive[PK_GATHER][1], 1u) == gridDim.x - 1) {
            loc->arrive[PK_GATHER][1] = 0u;
            *(volatile unsigned long long*)&loc->epoch[PK_GATHER] = e;
            __threadfence();
        }
    }
}

}  // namespace

struct HaloPlan {
    std::vector<int64_t> send_off, recv_off;   // [nnbr+1]
    int32_t* send_idx = nullptr;               // entries of owned dofs a neighbour ghosts
    int32_t* recv_idx = nullptr;               // entries of my ghost dofs
    // peer-memory path: the same tables on the device + where my messages land in the neighbours' staging areas
    int64_t* d_send_off = nullptr;
    int64_t* d_recv_off = nullptr;
    int64_t* d_dst_off = nullptr;
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
    // peer-memory channel
    bool peer_ok = false;
    bool want_peer = true;
    void* arena = nullptr;                // my IPC-shared arena
    void* peer_base[kMaxRanks] = {};      // mapped arenas of the other ranks (mine: arena)
    LaneLocal* lane_local = nullptr;      // [kLanes] private device state
    PeerLaneDev lane[kLanes];
    int64_t halo_cap = 0, gather_cap = 0;
    int* d_nbr = nullptr;
};

// ---------------------------------------------------------------------------------- peer channel: host side
static int64_t lane_bytes(int64_t halo_cap, int64_t gather_cap) {
    const int64_t hdr = ((int64_t)sizeof(PeerShared) + 255) & ~(int64_t)255;
    return hdr + 8 * (2 * halo_cap + 2 * gather_cap);
}

static void peer_close(DistState* d) {
    for (int r = 0; r < d->nranks; ++r)
        if (r != d->rank && d->peer_base[r]) { cudaIpcCloseMemHandle(d->peer_base[r]); d->peer_base[r] = nullptr; }
    if (d->arena) { cudaFree(d->arena); d->arena = nullptr; }
    d->peer_ok = false;
}

// (Re)allocate the IPC arena with the given staging capacities (doubles) and map every peer's arena.  Collective:
// every rank calls it with the same capacities at a quiescent point.  On any failure all ranks fall back to NCCL.
static void peer_arena(mpet_ctx* ctx, int64_t halo_cap, int64_t gather_cap, cudaStream_t st) {
    DistState* d = ctx->dist;
    if (!d->want_peer || d->nranks < 2) return;
    CUDA_CHECK(cudaDeviceSynchronize());
    peer_close(d);
    halo_cap = (halo_cap + 31) & ~(int64_t)31;
    gather_cap = (gather_cap + 31) & ~(int64_t)31;
    const int64_t lb = lane_bytes(halo_cap, gather_cap);
    int ok = 1;
    cudaIpcMemHandle_t mine;
    if (cudaMalloc(&d->arena, (size_t)(kLanes * lb)) != cudaSuccess) { d->arena = nullptr; ok = 0; }
    if (ok) {
        CUDA_CHECK(cudaMemset(d->arena, 0, (size_t)(kLanes * lb)));
        CUDA_CHECK(cudaDeviceSynchronize());
        if (cudaIpcGetMemHandle(&mine, d->arena) != cudaSuccess) ok = 0;
    }
    if (!ok) memset(&mine, 0, sizeof mine);
    cudaGetLastError();
    // handles of all ranks (the all-gather doubles as the barrier "everybody has zeroed its flags")
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "unexpected IPC handle size");
    char* dh = nullptr;
    CUDA_CHECK(cudaMalloc(&dh, (size_t)64 * (d->nranks + 1)));
    CUDA_CHECK(cudaMemcpy(dh + 64 * d->nranks, &mine, 64, cudaMemcpyHostToDevice));
    NCCL_CHECK(g_nccl.AllGather(dh + 64 * d->nranks, dh, 64, NCCL_INT8, d->comm, st));
    std::vector<cudaIpcMemHandle_t> all(d->nranks);
    CUDA_CHECK(cudaMemcpyAsync(all.data(), dh, (size_t)64 * d->nranks, cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(dh);
    for (int r = 0; r < d->nranks && ok; ++r) {
        if (r == d->rank) { d->peer_base[r] = d->arena; continue; }
        if (cudaIpcOpenMemHandle(&d->peer_base[r], all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            d->peer_base[r] = nullptr;
            ok = 0;
        }
    }
    cudaGetLastError();
    // unanimous decision
    double flag = ok ? 0.0 : 1.0, *dflag = nullptr;
    CUDA_CHECK(cudaMalloc(&dflag, sizeof(double)));
    CUDA_CHECK(cudaMemcpy(dflag, &flag, sizeof(double), cudaMemcpyHostToDevice));
    NCCL_CHECK(g_nccl.AllReduce(dflag, dflag, 1, NCCL_FLOAT64, NCCL_MAX, d->comm, st));
    CUDA_CHECK(cudaMemcpyAsync(&flag, dflag, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(dflag);
    if (flag != 0.0) {
        peer_close(d);
        if (getenv("MPET_AMG_VERBOSE")) fprintf(stderr, "[mpet dist] peer-memory channel unavailable, using NCCL send/recv\n");
        return;
    }
    if (!d->lane_local) d->lane_local = dev_alloc<LaneLocal>(ctx, kLanes);
    CUDA_CHECK(cudaMemset(d->lane_local, 0, sizeof(LaneLocal) * kLanes));
    const int64_t hdr = ((int64_t)sizeof(PeerShared) + 255) & ~(int64_t)255;
    for (int l = 0; l < kLanes; ++l) {
        PeerLaneDev& L = d->lane[l];
        memset(&L, 0, sizeof L);
        for (int r = 0; r < d->nranks; ++r) {
            char* base = (char*)d->peer_base[r] + l * lb;
            L.hdr[r] = (PeerShared*)base;
            L.halo[r] = (double*)(base + hdr);
            L.gather[r] = L.halo[r] + 2 * halo_cap;
        }
        L.local = d->lane_local + l;
        L.halo_cap = halo_cap;
        L.gather_cap = gather_cap;
        L.rank = d->rank;
        L.nranks = d->nranks;
    }
    d->halo_cap = halo_cap;
    d->gather_cap = gather_cap;
    d->peer_ok = true;
    ctx->graph_epoch++;           // captured launches hold the old arena pointers
    CUDA_CHECK(cudaDeviceSynchronize());
}

// after the halo plans are built: device copies of the offset tables, and for every neighbour the offset at which
// MY message starts in ITS staging area (= its recv offset for me), exchanged once through NCCL
static void peer_setup_halo(mpet_ctx* ctx, int64_t maxlen, cudaStream_t st) {
    DistState* d = ctx->dist;
    if (!d->want_peer || d->nranks < 2) return;
    const int nn = (int)d->nbr.size();
    MPET_REQUIRE(nn <= kMaxNbr, "too many neighbour ranks");
    d->d_nbr = dev_alloc<int>(ctx, nn);
    if (nn) CUDA_CHECK(cudaMemcpy(d->d_nbr, d->nbr.data(), sizeof(int) * nn, cudaMemcpyHostToDevice));
    // exchange (recv offset, recv length) per plan with every neighbour
    const int W = 2 * DIST_NPLANS;
    std::vector<int64_t> out((size_t)std::max(1, nn) * W), in((size_t)std::max(1, nn) * W);
    for (int q = 0; q < nn; ++q)
        for (int pl = 0; pl < DIST_NPLANS; ++pl) {
            out[(size_t)q * W + 2 * pl] = d->plan[pl].recv_off[q];
            out[(size_t)q * W + 2 * pl + 1] = d->plan[pl].recv_off[q + 1] - d->plan[pl].recv_off[q];
        }
    int64_t *dout = nullptr, *din = nullptr;
    CUDA_CHECK(cudaMalloc(&dout, sizeof(int64_t) * out.size()));
    CUDA_CHECK(cudaMalloc(&din, sizeof(int64_t) * in.size()));
    CUDA_CHECK(cudaMemcpy(dout, out.data(), sizeof(int64_t) * out.size(), cudaMemcpyHostToDevice));
    NCCL_CHECK(g_nccl.GroupStart());
    for (int q = 0; q < nn; ++q) {
        NCCL_CHECK(g_nccl.Send(dout + (size_t)q * W, (size_t)W * 8, NCCL_INT8, d->nbr[q], d->comm, st));
        NCCL_CHECK(g_nccl.Recv(din + (size_t)q * W, (size_t)W * 8, NCCL_INT8, d->nbr[q], d->comm, st));
    }
    NCCL_CHECK(g_nccl.GroupEnd());
    CUDA_CHECK(cudaMemcpyAsync(in.data(), din, sizeof(int64_t) * in.size(), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(dout);
    cudaFree(din);
    for (int pl = 0; pl < DIST_NPLANS; ++pl) {
        HaloPlan& P = d->plan[pl];
        std::vector<int64_t> dst(std::max(1, nn));
        for (int q = 0; q < nn; ++q) {
            dst[q] = in[(size_t)q * W + 2 * pl];
            MPET_REQUIRE(in[(size_t)q * W + 2 * pl + 1] == P.send_off[q + 1] - P.send_off[q],
                         "halo lists of two neighbouring ranks disagree in length");
        }
        P.d_send_off = dev_alloc<int64_t>(ctx, nn + 1);
        P.d_recv_off = dev_alloc<int64_t>(ctx, nn + 1);
        P.d_dst_off = dev_alloc<int64_t>(ctx, std::max(1, nn));
        CUDA_CHECK(cudaMemcpy(P.d_send_off, P.send_off.data(), sizeof(int64_t) * (nn + 1), cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(P.d_recv_off, P.recv_off.data(), sizeof(int64_t) * (nn + 1), cudaMemcpyHostToDevice));
        CUDA_CHECK(cudaMemcpy(P.d_dst_off, dst.data(), sizeof(int64_t) * std::max(1, nn), cudaMemcpyHostToDevice));
    }
    // the largest message volume over all ranks sizes everybody's staging area
    double cap = (double)maxlen, *dcap = nullptr;
    CUDA_CHECK(cudaMalloc(&dcap, sizeof(double)));
    CUDA_CHECK(cudaMemcpy(dcap, &cap, sizeof(double), cudaMemcpyHostToDevice));
    NCCL_CHECK(g_nccl.AllReduce(dcap, dcap, 1, NCCL_FLOAT64, NCCL_MAX, d->comm, st));
    CUDA_CHECK(cudaMemcpyAsync(&cap, dcap, sizeof(double), cudaMemcpyDeviceToHost, st));
    CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(dcap);
    peer_arena(ctx, (int64_t)cap, 1024, st);
}

// collective: make room for all-gathers of `doubles` values in total (AMG set-up, amg.cu)
void dist_reserve_gather(mpet_ctx* ctx, int64_t doubles, cudaStream_t st) {
    DistState* d = ctx->dist;
    if (!d || !d->peer_ok || doubles <= d->gather_cap) return;
    peer_arena(ctx, d->halo_cap, doubles, st);
}

static int peer_grid(mpet_ctx* ctx, int64_t work) {
    int64_t g = (work + 2047) / 2048;          // ~8 entries per thread
    if (g < 1) g = 1;
    if (g > ctx->sm_count) g = ctx->sm_count;   // all CTAs co-resident (the kernels spin on remote flags)
    return (int)g;
}

// device-side time-outs of the spin loops surface here (called after a solve has been synchronised)
void dist_check(mpet_ctx* ctx) {
    DistState* d = ctx->dist;
    if (!d || !d->peer_ok) return;
    LaneLocal h[kLanes];
    CUDA_CHECK(cudaMemcpy(h, d->lane_local, sizeof h, cudaMemcpyDeviceToHost));
    for (int l = 0; l < kLanes; ++l)
        MPET_REQUIRE(h[l].error == 0, "peer-memory exchange timed out waiting for a neighbour rank (60 s)");
}

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
    MPET_REQUIRE(nranks <= kMaxRanks, "at most 16 ranks per node");
    {
        const char* e = getenv("MPET_COMM");       // "nccl": keep ncclSend/ncclRecv halos (A/B measurements, fallback)
        d->want_peer = !(e && e[0] == 'n');
    }
    // A second NCCL communicator driven from a second stream is a documented deadlock hazard (concurrent
    // collectives of two communicators on one device are not ordered across ranks): opt-in only.  The
    // peer-memory channel gives the side stream its own lane without NCCL.
    if (g_nccl.CommSplit && getenv("MPET_NCCL_LANE1") && !d->want_peer) {
        if (g_nccl.CommSplit(d->comm, 0, rank, &d->comm2, nullptr) != 0) d->comm2 = nullptr;
    }
    ctx->dist = d;
}

void dist_free(mpet_ctx* ctx) {
    if (!ctx->dist) return;
    peer_close(ctx->dist);
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
    peer_setup_halo(ctx, maxlen, st);
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
int dist_comm_kind(mpet_ctx* ctx) { return dist_active(ctx) ? (ctx->dist->peer_ok ? 2 : 1) : 0; }
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
    if (d->peer_ok && !reverse) {
        const int lane = ctx->dist_lane == 1 ? 1 : 0;
        cudaEvent_t pe = lane ? nullptr : prof_begin(ctx, st);
        HaloArgs H{P.send_idx, P.recv_idx, P.d_send_off, P.d_recv_off, P.d_dst_off, d->d_nbr, nn};
        k_peer_halo<<<peer_grid(ctx, std::max(np, nu)), 256, 0, st>>>(d->lane[lane], H, v, done);
        LAUNCH_CHECK(ctx);
        prof_end(ctx, PROF_COMM, pe, st);
        return;
    }
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
void dist_allgather(mpet_ctx* ctx, const double* send, double* recv, int64_t count, const int* done, cudaStream_t st) {
    DistState* d = ctx->dist;
    if (d->peer_ok && count * d->nranks <= d->gather_cap) {
        const int lane = ctx->dist_lane == 1 ? 1 : 0;
        cudaEvent_t pe = lane ? nullptr : prof_begin(ctx, st);
        k_peer_allgather<<<peer_grid(ctx, count * d->nranks), 256, 0, st>>>(d->lane[lane], send, recv, count, done);
        LAUNCH_CHECK(ctx);
        prof_end(ctx, PROF_COMM, pe, st);
        return;
    }
    const bool lane1 = ctx->dist_lane == 1 && d->comm2;
    cudaEvent_t pe = lane1 ? nullptr : prof_begin(ctx, st);
    NCCL_CHECK(g_nccl.AllGather(send, recv, (size_t)count, NCCL_FLOAT64, lane1 ? d->comm2 : d->comm, st));
    prof_end(ctx, PROF_COMM, pe, st);
}
bool dist_has_lane1(mpet_ctx* ctx) { return dist_active(ctx) && (ctx->dist->peer_ok || ctx->dist->comm2 != nullptr); }
// broadcast of raw bytes (device buffer) from `root`
void dist_bcast_bytes(mpet_ctx* ctx, void* buf, int64_t nbytes, int root, cudaStream_t st) {
    NCCL_CHECK(g_nccl.Broadcast(buf, buf, (size_t)nbytes, NCCL_INT8, root, ctx->dist->comm, st));
}
void dist_allreduce_max(mpet_ctx* ctx, double* dev_scalars, int count, cudaStream_t st) {
    if (!dist_active(ctx)) return;
    NCCL_CHECK(g_nccl.AllReduce(dev_scalars, dev_scalars, (size_t)count, NCCL_FLOAT64, NCCL_MAX, ctx->dist->comm, st));
}

// the second stage of a dot product and the sum over ranks in ONE launch (peer channel); false: caller does both
bool dist_reduce_partials(mpet_ctx* ctx, const double* partials, int nparts, int count, double* out, const int* done,
                          cudaStream_t st) {
    if (!dist_active(ctx) || !ctx->dist->peer_ok || count > kRedMax) return false;
    cudaEvent_t pe = prof_begin(ctx, st);
    k_peer_allreduce<<<1, 256, 0, st>>>(ctx->dist->lane[ctx->dist_lane == 1 ? 1 : 0], partials, nparts, count, out, done);
    LAUNCH_CHECK(ctx);
    prof_end(ctx, PROF_COMM, pe, st);
    return true;
}

void dist_allreduce_sum(mpet_ctx* ctx, double* dev_scalars, int count, cudaStream_t st) {
    if (!dist_active(ctx)) return;
    if (ctx->dist->peer_ok && count <= kRedMax) {
        cudaEvent_t pe = prof_begin(ctx, st);
        k_peer_allreduce<<<1, 256, 0, st>>>(ctx->dist->lane[ctx->dist_lane == 1 ? 1 : 0], nullptr, 0, count, dev_scalars,
                                            nullptr);
        LAUNCH_CHECK(ctx);
        prof_end(ctx, PROF_COMM, pe, st);
        return;
    }
    cudaEvent_t pe = prof_begin(ctx, st);
    NCCL_CHECK(g_nccl.AllReduce(dev_scalars, dev_scalars, (size_t)count, NCCL_FLOAT64, NCCL_SUM, ctx->dist->comm, st));
    prof_end(ctx, PROF_COMM, pe, st);
}
