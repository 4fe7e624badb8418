// Shared-memory-staged sparse kernels: the matrix stream is moved by the bulk-async copy engine
// (cp.async.bulk -> SASS UBLKCP, completion on an mbarrier), warps only gather and reduce.
//
// Why (profiles/r01_*): warp-per-row kernels are latency-bound.  Each warp walks a chain of dependent
// DRAM round trips (row pointers -> values/indices -> gather -> epilogue operands), so only ~1/3 of HBM
// bandwidth is reached.  Here a persistent CTA owns a ring of shared-memory stages.  ONE producer thread
// asks the copy engine for chunk after chunk (a run of consecutive rows whose values, column indices and
// row pointers are CONTIGUOUS in memory) as stages drain; the consumer warps take rows out of shared
// memory.  The only long-latency operation left in a consumer is the 256-bit gather of the neighbour's
// vector entries.
//
// Pipeline (second design, profiles/r01_block_rows_staged_v1.md): every stage has a `full` mbarrier
// (producer's expect_tx + the copy engine's complete_tx) and an `empty` mbarrier (one arrival per consumer
// warp).  There is no CTA-wide barrier in the steady state: the first design ended every chunk with
// __syncthreads and ncu attributed 39 % of all stall samples to it, because a chunk holds fewer rows than
// the CTA has warps.  Rows are dealt to the consumer warps round-robin ACROSS chunks, so the load is even
// although a single chunk is not.
//
// Two kernels share the machinery:
//   k_block_rows_pipe : the MPET block system, one "node" = NR rows that share their column structure
//                       (3 displacement rows of a P2 node / A pressure rows of a vertex).
//   k_spmm_pipe       : scalar CSR matrix applied to W-wide vectors with the fused Chebyshev / residual
//                       epilogue of the AMG V-cycle.
#include "ctx.h"
#include "layout.cuh"
#include "staged.h"
#include <algorithm>
#include <cstdlib>

namespace {

// ------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok = 0;
    const uint32_t addr = smem_u32(bar);
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}

// ------------------------------------------------------------------------------- configurations
// THREADS = 32 * (consumer warps + 1 producer warp); MINB = CTAs per SM the shared-memory ring allows.
template <int THREADS, int STAGES, int CAPV, int MINB>
struct BlkCfg {
    static constexpr int kThreads = THREADS, kStages = STAGES, kCapV = CAPV, kMinBlocks = MINB;
    static constexpr int kConsumers = THREADS / 32 - 1;
};
template <int THREADS, int STAGES, int CAP, int ROWS, int MINB>
struct SpmCfg {
    static constexpr int kThreads = THREADS, kStages = STAGES, kCap = CAP, kRows = ROWS, kMinBlocks = MINB;
    static constexpr int kConsumers = THREADS / 32 - 1;
};

constexpr int kNnMax = 64;      // nodes per block chunk (row-pointer slice)

// values per row group of one block stage when the node has NR rows: the stage keeps NR * capv doubles
constexpr int blk_capv(int cfg_capv, int nr) { return nr <= 3 ? cfg_capv : ((cfg_capv * 3 / nr) & ~1); }
constexpr int blk_capc(int capv) { return (capv / 3 + 3) & ~3; }

// ------------------------------------------------------------------------------- block rows
template <int NR, int CAPV>
struct BlockStage {
    static constexpr int kCapC = blk_capc(CAPV);
    double vals[NR][CAPV + 2];
    int32_t colA[kCapC + 8];
    int32_t colB[kCapC + 8];
    int32_t rpA[kNnMax + 8];
    int32_t rpB[kNnMax + 8];
};

struct GroupBase { int64_t v[MPET_MAX_NETWORKS]; };

template <int NR, class CFG>
constexpr size_t blk_smem_bytes() {
    return CFG::kStages * sizeof(BlockStage<NR, blk_capv(CFG::kCapV, NR)>) + 2 * CFG::kStages * sizeof(uint64_t) +
           CFG::kStages * sizeof(BlockChunk);
}

// NR rows per node, NA scalar (pressure) column blocks.  U_OUT: rows are the displacement of a P2 node
// (output = one padded 256-bit store), else the pressures of a vertex (NR scalar stores, stride nv).
// LANES lanes of a warp share one node (32 / LANES nodes per warp step).
template <int NR, int NA, bool U_OUT, int LANES, class CFG>
__global__ void __launch_bounds__(CFG::kThreads, CFG::kMinBlocks)
k_block_rows_pipe(const BlockChunk* __restrict__ chunks, int nchunks, GroupBase gbase,
                  const int32_t* __restrict__ rpA, const int32_t* __restrict__ colA,
                  const int32_t* __restrict__ rpB, const int32_t* __restrict__ colB,
                  const double* __restrict__ vals, const double* __restrict__ x, double* __restrict__ y,
                  int64_t n2, int64_t nv, const uint8_t* __restrict__ mask, const int* __restrict__ done,
                  int contiguous) {
    if (done && *done) return;
    constexpr int kStages = CFG::kStages, NCW = CFG::kConsumers, GPW = 32 / LANES;
    using Stage = BlockStage<NR, blk_capv(CFG::kCapV, NR)>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage* stage = reinterpret_cast<Stage*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * sizeof(Stage));
    uint64_t* empty = full + kStages;
    BlockChunk* desc = reinterpret_cast<BlockChunk*>(empty + kStages);
    const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
    const int gw = wl / LANES, lane = wl % LANES;          // node slot inside the warp, lane inside the slot

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCW); }
        fence_barrier_init();
    }
    __syncthreads();

    // chunk order: interleaved over the CTAs, or one contiguous range per CTA (L1 reuse of the gathered
    // vector entries between neighbouring rows)
    int first = blockIdx.x, stride = gridDim.x;
    int nmine = first < nchunks ? (nchunks - first + stride - 1) / stride : 0;
    if (contiguous) {
        const int per = nchunks / (int)gridDim.x, rem = nchunks % (int)gridDim.x;
        first = (int)blockIdx.x * per + min((int)blockIdx.x, rem);
        nmine = per + ((int)blockIdx.x < rem ? 1 : 0);
        stride = 1;
    }

    if (warp == 0) {
        // ------------------------------------------------ producer: one thread feeds the copy engine
        if (wl != 0) return;
        for (int it = 0; it < nmine; ++it) {
            const int s = it % kStages;
            if (it >= kStages) mbar_wait(&empty[s], (uint32_t)((it / kStages - 1) & 1));
            const BlockChunk c = chunks[first + it * stride];
            desc[s] = c;                    // consumers read the descriptor from shared memory (released by the arrive)
            Stage& S = stage[s];
            const uint32_t rp_bytes = (uint32_t)c.rp_len * 4u;
            const uint32_t ca_bytes = (uint32_t)c.cA_len * 4u, cb_bytes = (uint32_t)c.cB_len * 4u;
            uint32_t total = 2 * rp_bytes + ca_bytes + cb_bytes;
            uint32_t vbytes[NR];
            int64_t vstart[NR];
#pragma unroll
            for (int g = 0; g < NR; ++g) {
                const int64_t s0 = gbase.v[g] + c.v_rel;
                vstart[g] = s0 & ~(int64_t)1;
                vbytes[g] = (uint32_t)((((s0 - vstart[g]) + c.v_len + 1) & ~(int64_t)1) * 8);
                total += vbytes[g];
            }
            mbar_expect_tx(&full[s], total);
            bulk_g2s(S.rpA, rpA + c.rp_off, rp_bytes, &full[s]);
            bulk_g2s(S.rpB, rpB + c.rp_off, rp_bytes, &full[s]);
            if (ca_bytes) bulk_g2s(S.colA, colA + c.cA_off, ca_bytes, &full[s]);
            if (cb_bytes) bulk_g2s(S.colB, colB + c.cB_off, cb_bytes, &full[s]);
#pragma unroll
            for (int g = 0; g < NR; ++g) bulk_g2s(S.vals[g], vals + vstart[g], vbytes[g], &full[s]);
        }
        return;
    }

    // ---------------------------------------------------- consumers: node slots dealt round-robin across chunks
    const int cw = warp - 1;
    int rot = 0;                               // warp steps (GPW nodes each) of earlier chunks, modulo NCW
    for (int it = 0; it < nmine; ++it) {
        const int s = it % kStages;
        mbar_wait(&full[s], (uint32_t)((it / kStages) & 1));
        const BlockChunk c = desc[s];
        const Stage& S = stage[s];
        const int rsk = c.n0 - c.rp_off;                    // skew of the row-pointer slices
        const int32_t a_base = S.rpA[rsk], b_base = S.rpB[rsk];
        const int ska = a_base - c.cA_off, skb = b_base - c.cB_off;
        int vsk[NR];
#pragma unroll
        for (int g = 0; g < NR; ++g) vsk[g] = (int)((gbase.v[g] + c.v_rel) & 1);

        const int nq = (c.nn + GPW - 1) / GPW;
        int q = cw - rot;
        if (q < 0) q += NCW;
        for (; q < nq; q += NCW) {
            const int nl = q * GPW + gw;
            const bool active = nl < c.nn;
            const int nls = active ? nl : 0;
            const int32_t ra = S.rpA[rsk + nls], rb = S.rpB[rsk + nls];
            const int dA = active ? S.rpA[rsk + nls + 1] - ra : 0, dB = active ? S.rpB[rsk + nls + 1] - rb : 0;
            const int eA = ra - a_base + ska, eB = rb - b_base + skb;
            const int rel = 3 * (ra - a_base) + NA * (rb - b_base);
            const int64_t node = (int64_t)c.n0 + nls;
            // every global load of the node's first sweep is issued before anything is consumed
            uint32_t mword = 0;
            if (U_OUT && mask && lane == 0 && active) mword = __ldg(reinterpret_cast<const uint32_t*>(mask + 4 * node));
            const bool hasA = lane < dA, hasB = (NA > 0) && lane < dB;
            d4 xv = {0, 0, 0, 0};
            if (hasA) xv = ld256_gather(x + 4 * (int64_t)S.colA[eA + lane]);
            double pv[NA > 0 ? NA : 1];
            const double* p = x + 4 * n2;
            if (NA > 0) {
                const int32_t colb = hasB ? S.colB[eB + lane] : 0;
#pragma unroll
                for (int i = 0; i < NA; ++i) pv[i] = hasB ? __ldg(p + (int64_t)i * nv + colb) : 0.0;
            }
            double acc[NR];
#pragma unroll
            for (int g = 0; g < NR; ++g) acc[g] = 0.0;
            if (hasA) {
#pragma unroll
                for (int g = 0; g < NR; ++g) {
                    const double* v = S.vals[g] + vsk[g] + rel + lane;
                    acc[g] += v[0] * xv.x + v[dA] * xv.y + v[2 * dA] * xv.z;
                }
            }
            if (NA > 0 && hasB) {
#pragma unroll
                for (int i = 0; i < NA; ++i)
#pragma unroll
                    for (int g = 0; g < NR; ++g) acc[g] += S.vals[g][vsk[g] + rel + 3 * dA + i * dB + lane] * pv[i];
            }
            for (int j = lane + LANES; j < dA; j += LANES) {       // nodes with more than LANES neighbours
                const d4 xw = ld256_gather(x + 4 * (int64_t)S.colA[eA + j]);
#pragma unroll
                for (int g = 0; g < NR; ++g) {
                    const double* v = S.vals[g] + vsk[g] + rel + j;
                    acc[g] += v[0] * xw.x + v[dA] * xw.y + v[2 * dA] * xw.z;
                }
            }
            if (NA > 0) {
                for (int j = lane + LANES; j < dB; j += LANES) {
                    const int32_t col = S.colB[eB + j];
#pragma unroll
                    for (int i = 0; i < NA; ++i) {
                        const double pw = __ldg(p + (int64_t)i * nv + col);
#pragma unroll
                        for (int g = 0; g < NR; ++g) acc[g] += S.vals[g][vsk[g] + rel + 3 * dA + i * dB + j] * pw;
                    }
                }
            }
#pragma unroll
            for (int g = 0; g < NR; ++g) {
#pragma unroll
                for (int o = LANES / 2; o > 0; o >>= 1) acc[g] += __shfl_xor_sync(0xffffffffu, acc[g], o, LANES);
            }
            if (lane == 0 && active) {
                if (U_OUT) {
                    d4 out = {acc[0], NR > 1 ? acc[NR > 1 ? 1 : 0] : 0.0, NR > 2 ? acc[NR > 2 ? 2 : 0] : 0.0, 0.0};
                    if (mask) {
                        const uint32_t m = mword;
                        if (m) {
                            const d4 xa = ld256(x + 4 * node);
                            if (m & 0x000000ffu) out.x = xa.x;
                            if (m & 0x0000ff00u) out.y = xa.y;
                            if (m & 0x00ff0000u) out.z = xa.z;
                        }
                    }
                    st256(y + 4 * node, out);
                } else {
#pragma unroll
                    for (int g = 0; g < NR; ++g) {
                        const int64_t idx = 4 * n2 + (int64_t)g * nv + node;
                        y[idx] = (mask && mask[idx]) ? x[idx] : acc[g];
                    }
                }
            }
        }
        rot = (rot + nq) % NCW;
        __syncwarp();
        if (wl == 0) mbar_arrive(&empty[s]);      // this warp is done reading stage s
    }
}

// ------------------------------------------------------------------------------- scalar SpMM
template <int CAP, int ROWS>
struct SpmmStage {
    double vals[CAP + 2];
    int32_t cols[CAP + 8];
    int32_t rp[ROWS + 8];
};

enum { SEPI_RESID = 0, SEPI_CHEB = 1, SEPI_PLAIN = 2 };

template <class CFG>
constexpr size_t spm_smem_bytes() {
    return CFG::kStages * sizeof(SpmmStage<CFG::kCap, CFG::kRows>) + 2 * CFG::kStages * sizeof(uint64_t) +
           CFG::kStages * sizeof(SpmmChunk);
}

// Row groups of LANES threads; epilogues as in amg.cu (RESID: out = b - Ax; CHEB: Chebyshev step;
// PLAIN: out = Ax + beta*out).
template <int W, int LANES, int EPI, class CFG>
__global__ void __launch_bounds__(CFG::kThreads, CFG::kMinBlocks)
k_spmm_pipe(const SpmmChunk* __restrict__ chunks, int nchunks, const int32_t* __restrict__ rowptr,
            const int32_t* __restrict__ cols, const double* __restrict__ vals, const double* __restrict__ x,
            const double* __restrict__ b, double* __restrict__ out, double* __restrict__ d,
            const double* __restrict__ dinv, double c1, double c2, const int* __restrict__ done, int contiguous) {
    if (done && *done) return;
    constexpr int kStages = CFG::kStages, NCW = CFG::kConsumers, GPW = 32 / LANES;
    using Stage = SpmmStage<CFG::kCap, CFG::kRows>;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    Stage* stage = reinterpret_cast<Stage*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + kStages * sizeof(Stage));
    uint64_t* empty = full + kStages;
    SpmmChunk* desc = reinterpret_cast<SpmmChunk*>(empty + kStages);
    const int warp = threadIdx.x >> 5, wl = threadIdx.x & 31;
    const int gw = wl / LANES, lane = wl % LANES;          // row group inside the warp, lane inside the group

    if (threadIdx.x == 0) {
        for (int s = 0; s < kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCW); }
        fence_barrier_init();
    }
    __syncthreads();

    // chunk order: interleaved over the CTAs, or one contiguous range per CTA (L1 reuse of the gathered
    // vector entries between neighbouring rows)
    int first = blockIdx.x, stride = gridDim.x;
    int nmine = first < nchunks ? (nchunks - first + stride - 1) / stride : 0;
    if (contiguous) {
        const int per = nchunks / (int)gridDim.x, rem = nchunks % (int)gridDim.x;
        first = (int)blockIdx.x * per + min((int)blockIdx.x, rem);
        nmine = per + ((int)blockIdx.x < rem ? 1 : 0);
        stride = 1;
    }

    if (warp == 0) {
        if (wl != 0) return;
        for (int it = 0; it < nmine; ++it) {
            const int s = it % kStages;
            if (it >= kStages) mbar_wait(&empty[s], (uint32_t)((it / kStages - 1) & 1));
            const SpmmChunk c = chunks[first + it * stride];
            desc[s] = c;
            Stage& S = stage[s];
            const uint32_t vb = (uint32_t)c.v_len * 8u, cb = (uint32_t)c.c_len * 4u, rb = (uint32_t)c.rp_len * 4u;
            mbar_expect_tx(&full[s], vb + cb + rb);
            bulk_g2s(S.rp, rowptr + c.rp_off, rb, &full[s]);
            if (vb) bulk_g2s(S.vals, vals + c.v_off, vb, &full[s]);
            if (cb) bulk_g2s(S.cols, cols + c.c_off, cb, &full[s]);
        }
        return;
    }

    const int cw = warp - 1;
    int rot = 0;                         // row quads (GPW rows) of earlier chunks, modulo NCW
    for (int it = 0; it < nmine; ++it) {
        const int s = it % kStages;
        mbar_wait(&full[s], (uint32_t)((it / kStages) & 1));
        const SpmmChunk c = desc[s];
        const Stage& S = stage[s];
        const int rsk = c.r0 - c.rp_off;
        const int32_t e_base = S.rp[rsk];
        const int vsk = e_base - c.v_off, csk = e_base - c.c_off;
        const int nq = (c.nrows + GPW - 1) / GPW;
        int q = cw - rot;
        if (q < 0) q += NCW;
        // warp-uniform trip count: every lane takes part in the shuffles even when its group has no row
        for (; q < nq; q += NCW) {
            const int rl = q * GPW + gw;
            const bool active = rl < c.nrows;
            const int64_t row = (int64_t)c.r0 + (active ? rl : 0);
            const int32_t e0 = active ? S.rp[rsk + rl] - e_base : 0;
            const int32_t e1 = active ? S.rp[rsk + rl + 1] - e_base : 0;
            double eb1 = 0, ed1 = 0, ex1 = 0, edi = 0;
            d4 eb4 = {0, 0, 0, 0}, ed4 = {0, 0, 0, 0}, ex4 = {0, 0, 0, 0};
            if (lane == 0 && active) {
                if constexpr (W == 1) {
                    if (EPI != SEPI_PLAIN) eb1 = b[row];
                    if (EPI == SEPI_CHEB) { edi = dinv[row]; ex1 = x[row]; if (c1 != 0.0) ed1 = d[row]; }
                    if (EPI == SEPI_PLAIN && c1 != 0.0) ex1 = out[row];
                } else {
                    if (EPI != SEPI_PLAIN) eb4 = ld256(b + 4 * row);
                    if (EPI == SEPI_CHEB) { edi = dinv[row]; ex4 = ld256(x + 4 * row); }   // d is read after the sweep (registers)
                    if (EPI == SEPI_PLAIN && c1 != 0.0) ex4 = ld256(out + 4 * row);
                }
            }
            double a1 = 0.0;
            d4 a4 = {0, 0, 0, 0};
            for (int e = e0 + lane; e < e1; e += LANES) {
                const double v = S.vals[vsk + e];
                const int32_t col = S.cols[csk + e];
                if constexpr (W == 1) {
                    a1 += v * __ldg(x + col);
                } else {
                    const d4 g = ld256_gather(x + 4 * (int64_t)col);
                    a4.x += v * g.x; a4.y += v * g.y; a4.z += v * g.z; a4.w += v * g.w;
                }
            }
#pragma unroll
            for (int o = LANES / 2; o > 0; o >>= 1) {
                if constexpr (W == 1) {
                    a1 += __shfl_xor_sync(0xffffffffu, a1, o, LANES);
                } else {
                    a4.x += __shfl_xor_sync(0xffffffffu, a4.x, o, LANES);
                    a4.y += __shfl_xor_sync(0xffffffffu, a4.y, o, LANES);
                    a4.z += __shfl_xor_sync(0xffffffffu, a4.z, o, LANES);
                    a4.w += __shfl_xor_sync(0xffffffffu, a4.w, o, LANES);
                }
            }
            if (lane == 0 && active) {
                if constexpr (W == 1) {
                    if (EPI == SEPI_PLAIN) {
                        out[row] = a1 + c1 * ex1;
                    } else {
                        const double res = eb1 - a1;
                        if (EPI == SEPI_RESID) out[row] = res;
                        else { const double dn = c2 * edi * res + c1 * ed1; d[row] = dn; out[row] = ex1 + dn; }
                    }
                } else {
                    if (EPI == SEPI_PLAIN) {
                        d4 o = {a4.x + c1 * ex4.x, a4.y + c1 * ex4.y, a4.z + c1 * ex4.z, a4.w + c1 * ex4.w};
                        st256(out + 4 * row, o);
                    } else {
                        d4 res = {eb4.x - a4.x, eb4.y - a4.y, eb4.z - a4.z, eb4.w - a4.w};
                        if (EPI == SEPI_RESID) {
                            st256(out + 4 * row, res);
                        } else {
                            const double s2 = c2 * edi;
                            if (c1 != 0.0) ed4 = ld256(d + 4 * row);
                            d4 dn = {s2 * res.x + c1 * ed4.x, s2 * res.y + c1 * ed4.y, s2 * res.z + c1 * ed4.z,
                                     s2 * res.w + c1 * ed4.w};
                            st256(d + 4 * row, dn);
                            d4 o = {ex4.x + dn.x, ex4.y + dn.y, ex4.z + dn.z, ex4.w + dn.w};
                            st256(out + 4 * row, o);
                        }
                    }
                }
            }
        }
        rot = (rot + nq) % NCW;
        __syncwarp();
        if (wl == 0) mbar_arrive(&empty[s]);
    }
}

// ------------------------------------------------------------------------------- configurations in use
// Measured on B200 (profiles/r01_pipeline_sweep.md); MPET_BLK_CFG / MPET_SPM_CFG select another one
// at plan-build time (development aid).
using BlkCfg0 = BlkCfg<512, 4, 960, 2>;      // 2.03 ms on cfg5 (old design: 2.76 ms)
using BlkCfg1 = BlkCfg<512, 3, 1344, 2>;     // 2.12 ms
using BlkCfg2 = BlkCfg<384, 4, 960, 2>;      // 2.00 ms
constexpr int kNumBlkCfg = 3;                // rejected: <1024,8,960,1> 2.63, <512,6,640,2> 2.81, <1024,6,1344,1> 2.29 ms
constexpr int kBlkCapV[kNumBlkCfg] = {BlkCfg0::kCapV, BlkCfg1::kCapV, BlkCfg2::kCapV};

using SpmCfg0 = SpmCfg<384, 4, 2048, 256, 2>;    // 768 threads per SM: 85 registers, no spills in the W = 4 Chebyshev step
using SpmCfg1 = SpmCfg<512, 4, 2048, 256, 2>;
using SpmCfg2 = SpmCfg<384, 3, 1024, 128, 2>;    // small ring: most of the 228 KB stays L1
using SpmCfg3 = SpmCfg<256, 3, 1024, 128, 4>;
using SpmCfg4 = SpmCfg<512, 3, 1024, 128, 2>;
constexpr int kNumSpmCfg = 5;
constexpr int kSpmCap[kNumSpmCfg] = {2048, 2048, 1024, 1024, 1024};
constexpr int kSpmRows[kNumSpmCfg] = {256, 256, 128, 128, 128};

int env_cfg(const char* name, int n, int dflt) {
    const char* e = getenv(name);
    if (!e || !*e) return dflt;
    int v = atoi(e);
    return (v >= 0 && v < n) ? v : dflt;
}

template <typename T>
T* upload_vec(mpet_ctx* ctx, const std::vector<T>& h) {
    T* d = dev_alloc<T>(ctx, (int64_t)h.size());
    if (!h.empty()) CUDA_CHECK(cudaMemcpy(d, h.data(), sizeof(T) * h.size(), cudaMemcpyHostToDevice));
    return d;
}

template <int NR, int NA, bool U_OUT, int LANES, class CFG>
void launch_block_rows_l(mpet_ctx* ctx, const BlockPlan& P, const GroupBase& gb, const NodeGraph& gA,
                         const NodeGraph& gB, const double* x, double* y, const uint8_t* mask, const int* done,
                         cudaStream_t st) {
    constexpr size_t smem = blk_smem_bytes<NR, CFG>();
    auto kern = k_block_rows_pipe<NR, NA, U_OUT, LANES, CFG>;
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int grid = std::min(P.nchunks, CFG::kMinBlocks * ctx->sm_count);
    kern<<<grid, CFG::kThreads, smem, st>>>(P.chunks, P.nchunks, gb, gA.rowptr, gA.col, gB.rowptr, gB.col, ctx->vals, x,
                                            y, ctx->N2, ctx->Nv, mask, done, env_cfg("MPET_BLK_ORDER", 2, 0));
    LAUNCH_CHECK(ctx);
}

// lanes per node: the displacement rows (28 node neighbours on average) and the pressure rows (63)
template <int NR, int NA, bool U_OUT, class CFG>
void launch_block_rows(mpet_ctx* ctx, const BlockPlan& P, const GroupBase& gb, const NodeGraph& gA,
                       const NodeGraph& gB, const double* x, double* y, const uint8_t* mask, const int* done,
                       cudaStream_t st) {
    // measured on cfg5 (spmv incl. layout conversions): 32/32 lanes 2.037 ms, 16 (u rows) / 32 (p rows) 1.885 ms,
    // 16/16 1.922 ms, 32/16 2.072 ms: half a warp per P2 node halves the shuffle work per node
    const int lanes = env_cfg(U_OUT ? "MPET_BLK_LANES_U" : "MPET_BLK_LANES_P", 64, U_OUT ? 16 : 32);
    if (lanes == 8) launch_block_rows_l<NR, NA, U_OUT, 8, CFG>(ctx, P, gb, gA, gB, x, y, mask, done, st);
    else if (lanes == 16) launch_block_rows_l<NR, NA, U_OUT, 16, CFG>(ctx, P, gb, gA, gB, x, y, mask, done, st);
    else launch_block_rows_l<NR, NA, U_OUT, 32, CFG>(ctx, P, gb, gA, gB, x, y, mask, done, st);
}

template <int NA, class CFG>
void block_rows_all(mpet_ctx* ctx, const double* x, double* y, const uint8_t* mask, const int* done, cudaStream_t st) {
    // value runs: row (k, a) starts at k*T + 3*rp22[a] + NA*rp21[a]; row (i, v) at 3*T + i*Tp + 3*rp12[v] + NA*rp11[v]
    const int64_t T = 3 * ctx->g22.nnz + (int64_t)NA * ctx->g21.nnz;
    const int64_t Tp = 3 * ctx->g12.nnz + (int64_t)NA * ctx->g11.nnz;
    GroupBase gu, gp;
    for (int k = 0; k < MPET_MAX_NETWORKS; ++k) { gu.v[k] = (int64_t)k * T; gp.v[k] = 3 * T + (int64_t)k * Tp; }
    launch_block_rows<3, NA, true, CFG>(ctx, ctx->plan_u, gu, ctx->g22, ctx->g21, x, y, mask, done, st);
    if (NA > 0)
        launch_block_rows<(NA > 0 ? NA : 1), NA, false, CFG>(ctx, ctx->plan_p, gp, ctx->g12, ctx->g11, x, y, mask, done, st);
}

template <class CFG>
bool block_rows_cfg(mpet_ctx* ctx, const double* x, double* y, const uint8_t* mask, const int* done, cudaStream_t st) {
    switch (ctx->A) {
        case 0: block_rows_all<0, CFG>(ctx, x, y, mask, done, st); break;
        case 1: block_rows_all<1, CFG>(ctx, x, y, mask, done, st); break;
        case 2: block_rows_all<2, CFG>(ctx, x, y, mask, done, st); break;
        case 3: block_rows_all<3, CFG>(ctx, x, y, mask, done, st); break;
        case 4: block_rows_all<4, CFG>(ctx, x, y, mask, done, st); break;
        case 5: block_rows_all<5, CFG>(ctx, x, y, mask, done, st); break;
        default: return false;
    }
    return true;
}

}  // namespace

// ------------------------------------------------------------------------------- host: plans
// Greedy chunking of consecutive nodes so that every staged array fits its shared-memory slot.
static bool build_block_plan(mpet_ctx* ctx, const NodeGraph& gA, const NodeGraph& gB, int NA, int NR, int cfg,
                             BlockPlan& plan) {
    const int capV = blk_capv(kBlkCapV[cfg], NR), capC = blk_capc(capV);
    const int64_t n = gA.nrows;
    std::vector<int32_t> rpA(n + 1), rpB(n + 1);
    CUDA_CHECK(cudaMemcpy(rpA.data(), gA.rowptr, sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToHost));
    CUDA_CHECK(cudaMemcpy(rpB.data(), gB.rowptr, sizeof(int32_t) * (n + 1), cudaMemcpyDeviceToHost));
    std::vector<BlockChunk> chunks;
    int64_t a0 = 0;
    while (a0 < n) {
        int64_t a1 = a0;
        const int rp_off = (int)(a0 & ~(int64_t)3);
        while (a1 < n) {
            const int64_t vlen = 3 * (int64_t)(rpA[a1 + 1] - rpA[a0]) + (int64_t)NA * (rpB[a1 + 1] - rpB[a0]);
            // column slices start at the aligned-down offset: up to 3 leading pad entries
            const int64_t la = rpA[a1 + 1] - (rpA[a0] & ~3), lb = rpB[a1 + 1] - (rpB[a0] & ~3);
            if (vlen > capV || la > capC || lb > capC || (a1 + 1 - rp_off) + 1 > kNnMax) break;
            ++a1;
        }
        if (a1 == a0) return false;      // a single node does not fit: caller falls back
        BlockChunk c;
        c.n0 = (int32_t)a0;
        c.nn = (int32_t)(a1 - a0);
        c.rp_off = rp_off;
        c.rp_len = (int32_t)((((a1 + 1) - rp_off) + 3) & ~(int64_t)3);
        c.cA_off = rpA[a0] & ~3;
        c.cA_len = ((rpA[a1] - c.cA_off) + 3) & ~3;
        c.cB_off = rpB[a0] & ~3;
        c.cB_len = ((rpB[a1] - c.cB_off) + 3) & ~3;
        c.v_rel = 3 * (int64_t)rpA[a0] + (int64_t)NA * rpB[a0];
        c.v_len = (int32_t)(3 * (int64_t)(rpA[a1] - rpA[a0]) + (int64_t)NA * (rpB[a1] - rpB[a0]));
        c.pad = 0;
        chunks.push_back(c);
        a0 = a1;
    }
    plan.nchunks = (int)chunks.size();
    plan.chunks = upload_vec(ctx, chunks);
    plan.cfg = cfg;
    return true;
}

void staged_build_block_plans(mpet_ctx* ctx) {
    const int cfg = env_cfg("MPET_BLK_CFG", kNumBlkCfg, 0);
    ctx->staged_ok = ctx->A <= 5 && build_block_plan(ctx, ctx->g22, ctx->g21, ctx->A, 3, cfg, ctx->plan_u);
    if (ctx->staged_ok && ctx->A > 0)
        ctx->staged_ok = build_block_plan(ctx, ctx->g12, ctx->g11, ctx->A, ctx->A, cfg, ctx->plan_p);
}

bool staged_block_spmv(mpet_ctx* ctx, const double* x, double* y, const uint8_t* mask, const int* done,
                       cudaStream_t st) {
    if (!ctx->staged_ok) return false;
    switch (ctx->plan_u.cfg) {
        case 0: return block_rows_cfg<BlkCfg0>(ctx, x, y, mask, done, st);
        case 1: return block_rows_cfg<BlkCfg1>(ctx, x, y, mask, done, st);
        case 2: return block_rows_cfg<BlkCfg2>(ctx, x, y, mask, done, st);
        default: return false;
    }
}

// scalar CSR: chunks of consecutive rows with at most cap entries / rows rows
bool staged_build_spmm_plan(mpet_ctx* ctx, const std::vector<int32_t>& rp, SpmmPlan& plan) {
    const int cfg = env_cfg("MPET_SPM_CFG", kNumSpmCfg, 0);
    const int cap = kSpmCap[cfg], rows = kSpmRows[cfg];
    const int64_t n = (int64_t)rp.size() - 1;
    std::vector<SpmmChunk> chunks;
    int64_t r0 = 0;
    while (r0 < n) {
        const int rp_off = (int)(r0 & ~(int64_t)3);
        const int32_t c_off = rp[r0] & ~3;          // aligned-down start of the staged slices (up to 3 pad entries)
        int64_t r1 = r0;
        while (r1 < n && (rp[r1 + 1] - c_off) <= cap && ((r1 + 1 - rp_off) + 1) <= rows) ++r1;
        if (r1 == r0) return false;
        SpmmChunk c;
        c.r0 = (int32_t)r0;
        c.nrows = (int32_t)(r1 - r0);
        c.rp_off = rp_off;
        c.rp_len = (int32_t)((((r1 + 1) - rp_off) + 3) & ~(int64_t)3);
        c.v_off = rp[r0] & ~1;
        c.v_len = ((rp[r1] - c.v_off) + 1) & ~1;
        c.c_off = c_off;
        c.c_len = ((rp[r1] - c.c_off) + 3) & ~3;
        chunks.push_back(c);
        r0 = r1;
    }
    plan.nchunks = (int)chunks.size();
    plan.chunks = upload_vec(ctx, chunks);
    plan.cfg = cfg;
    return true;
}

template <int W, int LANES, int EPI, class CFG>
static void launch_spmm_pipe(mpet_ctx* ctx, const SpmmPlan& P, const DevCsr& M, const double* x, const double* b,
                             double* out, double* d, const double* dinv, double c1, double c2, const int* done,
                             cudaStream_t st) {
    constexpr size_t smem = spm_smem_bytes<CFG>();
    auto kern = k_spmm_pipe<W, LANES, EPI, CFG>;
    static bool configured = false;
    if (!configured) {
        CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    int grid = std::min(P.nchunks, CFG::kMinBlocks * ctx->sm_count);
    kern<<<grid, CFG::kThreads, smem, st>>>(P.chunks, P.nchunks, M.rowptr, M.col, M.val, x, b, out, d, dinv, c1, c2, done,
                                            env_cfg("MPET_SPM_ORDER", 2, 0));
    LAUNCH_CHECK(ctx);
}

template <int W, int EPI, class CFG>
static void spmm_pipe_lanes(mpet_ctx* ctx, const SpmmPlan& P, const DevCsr& M, const double* x, const double* b,
                            double* out, double* d, const double* dinv, double c1, double c2, const int* done,
                            cudaStream_t st) {
    const double mean = M.nrows > 0 ? (double)M.nnz / (double)M.nrows : 0.0;
    // lanes per row.  Measured on cfg5 (one preconditioner application, profiles/r01_pipeline_sweep.md): the
    // kernels are bound by per-row work (shuffle reduction + epilogue), not by the gathers, so FEWER lanes win:
    // P2 level (28 entries/row) 32 lanes 7.08 ms, 16: 5.06, 8: 4.33, 4: 3.33, 2: 3.41, 1: 3.70;
    // P1 and aggregated levels (<= 20 entries/row) 8 lanes 3.49 ms, 4: 3.23, 2: 3.15, 1: 3.03.
    // (development aid: MPET_SPM_LANES_HI / _LO override.)
    // Longer rows (smoothed-aggregation levels, 50-200 entries) keep ~14-28 entries per lane.
    int lanes = mean > 20 ? env_cfg("MPET_SPM_LANES_HI", 64, 2) : env_cfg("MPET_SPM_LANES_LO", 64, 1);
    if (mean > 160) lanes = 16;
    else if (mean > 80) lanes = 8;
    else if (mean > 40) lanes = 4;
    switch (lanes) {
        case 1: launch_spmm_pipe<W, 1, EPI, CFG>(ctx, P, M, x, b, out, d, dinv, c1, c2, done, st); break;
        case 2: launch_spmm_pipe<W, 2, EPI, CFG>(ctx, P, M, x, b, out, d, dinv, c1, c2, done, st); break;
        case 8: launch_spmm_pipe<W, 8, EPI, CFG>(ctx, P, M, x, b, out, d, dinv, c1, c2, done, st); break;
        case 16: launch_spmm_pipe<W, 16, EPI, CFG>(ctx, P, M, x, b, out, d, dinv, c1, c2, done, st); break;
        default: launch_spmm_pipe<W, 4, EPI, CFG>(ctx, P, M, x, b, out, d, dinv, c1, c2, done, st); break;
    }
}

template <class CFG>
static void spmm_pipe_cfg(mpet_ctx* ctx, int W, int epi, const SpmmPlan& P, const DevCsr& M, const double* x,
                          const double* b, double* out, double* d, const double* dinv, double c1, double c2,
                          const int* done, cudaStream_t st) {
    if (W == 4) {
        if (epi == 0) spmm_pipe_lanes<4, SEPI_RESID, CFG>(ctx, P, M, x, b, out, d, dinv, c1, c2, done, st);
        else if (epi == 1) spmm_pipe_lanes<4, SEPI_CHEB, CFG>(ctx, P, M, x, b, out, d, dinv, c1, c2, done, st);
        else spmm_pipe_lanes<4, SEPI_PLAIN, CFG>(ctx, P, M, x, b, out, d, dinv, c1, c2, done, st);
    } else {
        if (epi == 0) spmm_pipe_lanes<1, SEPI_RESID, CFG>(ctx, P, M, x, b, out, d, dinv, c1, c2, done, st);
        else if (epi == 1) spmm_pipe_lanes<1, SEPI_CHEB, CFG>(ctx, P, M, x, b, out, d, dinv, c1, c2, done, st);
        else spmm_pipe_lanes<1, SEPI_PLAIN, CFG>(ctx, P, M, x, b, out, d, dinv, c1, c2, done, st);
    }
}

// epi: 0 = residual, 1 = Chebyshev step, 2 = plain (out = M x + c1 * out)
void staged_spmm(mpet_ctx* ctx, int W, int epi, const SpmmPlan& P, const DevCsr& M, const double* x, const double* b,
                 double* out, double* d, const double* dinv, double c1, double c2, const int* done, cudaStream_t st) {
    switch (P.cfg) {
        case 1: spmm_pipe_cfg<SpmCfg1>(ctx, W, epi, P, M, x, b, out, d, dinv, c1, c2, done, st); break;
        case 2: spmm_pipe_cfg<SpmCfg2>(ctx, W, epi, P, M, x, b, out, d, dinv, c1, c2, done, st); break;
        case 3: spmm_pipe_cfg<SpmCfg3>(ctx, W, epi, P, M, x, b, out, d, dinv, c1, c2, done, st); break;
        case 4: spmm_pipe_cfg<SpmCfg4>(ctx, W, epi, P, M, x, b, out, d, dinv, c1, c2, done, st); break;
        default: spmm_pipe_cfg<SpmCfg0>(ctx, W, epi, P, M, x, b, out, d, dinv, c1, c2, done, st); break;
    }
}
