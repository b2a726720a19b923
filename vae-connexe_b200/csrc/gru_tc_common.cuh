// Helpers shared by the tcgen05 recurrent kernels (gru_tc.cu forward, gru_tc_bwd.cu BPTT): 16x256b TMEM
// fragments, 256-bit global accesses, the TMEM-operand MMA and the unit permutation that makes them fit together.
#pragma once
#include "common.cuh"
#include "umma.cuh"

namespace crvae {

// 16 lanes x 32 columns fragment: thread (tr = lane/4, tq = lane%4) register 4*i + 2*rr + e  <->
// TMEM lane (base + tr + 8*rr), column (col + 8*i + 2*tq + e)
__device__ __forceinline__ void tmem_ld_16x32(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st_16x32(uint32_t taddr, const float* v) {
    const uint32_t* r = reinterpret_cast<const uint32_t*>(v);
    asm volatile(
        "tcgen05.st.sync.aligned.16x256b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]), "r"(r[10]),
        "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__device__ __forceinline__ float tf32_rna(float v) {
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
    return __uint_as_float(t);
}

// 256-bit global accesses (sm_100+): 8 consecutive floats of one row per thread, so the 4 threads of a quad
// cover one full 128-byte line -> a warp request touches 8 lines instead of 8 x (bytes / 32)
__device__ __forceinline__ void ldg_v8_stream(const float* p, float* v) {
    asm volatile("ld.global.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}
__device__ __forceinline__ void stg_v8(float* p, float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v0), "f"(v1), "f"(v2), "f"(v3), "f"(v4),
                 "f"(v5), "f"(v6), "f"(v7)
                 : "memory");
}

// D[tmem] (+)= A[tmem] * B[smem], issued by ONE thread
__device__ __forceinline__ void mma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, bool accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(static_cast<uint32_t>(accumulate))
        : "memory");
}

// Column permutation inside every 32-block of hidden units.  A thread's 16x256b fragment holds the TMEM columns
// p = 8i + 2tq + e (i < 4, e < 2); storing hidden unit u = 8tq + 2i + e at position p makes those 8 values
// 8 CONSECUTIVE units, i.e. one 32-byte vector of the row-major activations.  The same permutation is applied
// to the K index (columns of the h operand / of W_hh) and to the N index (gate rows of W_hh), so the GEMM
// is unchanged: gh[:, p] = sum_p' h[:, u(p')] * W_hh[u(p), u(p')].
__device__ __forceinline__ int pos_of_unit(int u) { return (u & ~31) | (((u >> 1) & 3) << 3) | (((u >> 3) & 3) << 1) | (u & 1); }

}  // namespace crvae
