// Solver-internal vector layout and 256-bit memory helpers.
//
// Across the C-ABI vectors use the UFC numbering of include/mpet_b200.h (component-major).  Inside
// mpet_solve every vector is stored "node-interleaved and padded":
//     displacement of scalar P2 node a  ->  4 consecutive doubles  [4a .. 4a+3] = (u0, u1, u2, 0)
//     pressure i at vertex v            ->  4*N2 + i*Nv + v
// so that ONE 32-byte load (sm_100a LDG.E.ENL2.256) gathers the whole displacement of a neighbour
// node.  One gather then feeds the nine matrix values of a node pair (block SpMV) or the three
// right-hand sides of the shared scalar preconditioner block, instead of one 8-byte gather -- and one
// L1TEX wavefront -- per matrix value.  The pad lane is kept identically zero by every kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

struct d4 { double x, y, z, w; };

__device__ __forceinline__ d4 ld256_gather(const double* p) {     // cached in L1: neighbours reuse it
    d4 r;
    asm volatile("ld.global.nc.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ d4 ld256(const double* p) {
    d4 r;
    asm volatile("ld.global.v4.f64 {%0,%1,%2,%3}, [%4];" : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ void st256(double* p, const d4& v) {
    asm volatile("st.global.v4.f64 [%0], {%1,%2,%3,%4};" ::"l"(p), "d"(v.x), "d"(v.y), "d"(v.z), "d"(v.w) : "memory");
}
__device__ __forceinline__ double ldg_stream_f64(const double* p) {
    double r;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ int32_t ldg_stream_s32(const int32_t* p) {
    int32_t r;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(r) : "l"(p));
    return r;
}

__host__ __device__ __forceinline__ int64_t api_to_internal(int64_t i, int64_t n2) {
    if (i < 3 * n2) {
        int64_t k = i / n2;
        return 4 * (i - k * n2) + k;
    }
    return i + n2;   // 4*n2 + (i - 3*n2)
}
