// Persistent multi-head GRU recurrence, FORWARD, on the 5th-generation tensor cores (tcgen05).
//
// One CTA owns one head and up to TWO 128-row batch tiles for ALL timesteps:
//   * W_hh of the head (tf32 hi and lo parts, 2 x 48 KB) is brought in ONCE by TMA and stays in
//     shared memory (K-major, SWIZZLE_128B);
//   * per step and tile ONE thread issues the gate GEMM  gh[128 x 192] = h[128 x 64] . W_hh^T  as
//     8 K-steps x 3 tcgen05.mma (3xTF32: h_lo.W_hi + h_hi.W_lo + h_hi.W_hi, fp32 accumulate in TMEM);
//   * a warpgroup per tile (thread = batch row, hidden state kept in 64 registers) reads the accumulator
//     with tcgen05.ld, streams gi from global, does the r/z/n gate math, the per-head Linear(H,1),
//     writes r|z|n, gh_n, h back (row-contiguous float4) and re-writes the tile's h as tf32 hi/lo
//     directly in the swizzled K-major operand layout (conflict-free 16-byte stores), then signals
//     the issuer through an mbarrier (fence.proxy.async);
//   * the two tiles ping-pong: while one warpgroup does its pointwise math the tensor core works on
//     the other tile.
// Same buffers / semantics as gru_fwd_kernel (gru_recurrent.cu), which remains the exact-fp32 path.
//
// Reference arithmetic replaced: nn.GRU per-step linear_hh + cell (CRVAE_lorenz96.py:119) and
// nn.Linear(H,1) (:120).
#include "common.cuh"
#include "umma.cuh"

namespace crvae {

int make_tmap_2d(CUtensorMap* m, const float* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems,
                 uint32_t box_inner, uint32_t box_rows, bool atom32b);

constexpr int GH = CRVAE_HIDDEN;                 // 64
constexpr int GG = CRVAE_G;                      // 192
constexpr int GT_ROWS = 128;                     // rows per tile (UMMA M)
constexpr int GT_W_HALF = GG * 128;              // 24576 B: [192 rows x 32 k] fp32, one K half
constexpr int GT_W_BYTES = 2 * GT_W_HALF;        // 49152 B per (hi | lo)
constexpr int GT_H_HALF = GT_ROWS * 128;         // 16384 B: [128 rows x 32 k]
constexpr int GT_H_BYTES = 2 * GT_H_HALF;        // 32768 B per (hi | lo) per tile
constexpr int GT_OFF_WHI = 0;
constexpr int GT_OFF_WLO = GT_W_BYTES;
constexpr int GT_OFF_H = 2 * GT_W_BYTES;         // tile s: + s * 2 * GT_H_BYTES ; hi then lo
constexpr int GT_OFF_CONST = GT_OFF_H + 4 * GT_H_BYTES;          // b_hh[192] | b_ih[192] | w_lin[64] | b_lin
constexpr int GT_CONST_BYTES = (2 * GG + GH + 4) * 4;
constexpr int GT_OFF_BAR = GT_OFF_CONST + ((GT_CONST_BYTES + 15) / 16) * 16;
constexpr int GT_SMEM_BYTES = GT_OFF_BAR + 128 + 1024;
constexpr int GT_TMEM_COLS = 512;
constexpr int GT_THREADS = 32 + 256;             // warp 0: TMA + MMA issue; warps 1-4: tile 0; warps 5-8: tile 1

struct GruTcArgs {
    float* gates; const float* b_ih; const float* b_hh;
    const float* h0; long long h0_stride;
    const float* w_lin; const float* b_lin;
    float* hs; float* ghn; float* pred;
    int P, T, B, t_skip;
};

__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}

__device__ __forceinline__ float tf32_rna(float v) {
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
    return __uint_as_float(t);
}

// one 16-byte chunk (4 consecutive k) of row `row` of a K-major SWIZZLE_128B tile pair (hi | lo)
__device__ __forceinline__ void store_h_chunk(uint8_t* h_hi_s, uint8_t* h_lo_s, int row, int c, float v0, float v1, float v2,
                                              float v3) {
    const float a0 = tf32_rna(v0), a1 = tf32_rna(v1), a2 = tf32_rna(v2), a3 = tf32_rna(v3);
    const int half = c >> 3, cc = c & 7;
    const int off = half * GT_H_HALF + row * 128 + ((cc ^ (row & 7)) << 4);
    *reinterpret_cast<float4*>(h_hi_s + off) = make_float4(a0, a1, a2, a3);
    *reinterpret_cast<float4*>(h_lo_s + off) = make_float4(__fsub_rn(v0, a0), __fsub_rn(v1, a1), __fsub_rn(v2, a2), __fsub_rn(v3, a3));
}

__global__ void __launch_bounds__(GT_THREADS, 1)
gru_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, GruTcArgs a) {
    using namespace umma;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
    float* cst = reinterpret_cast<float*>(smem + GT_OFF_CONST);      // b_hh | b_ih | w_lin | b_lin
    uint64_t* wbar = reinterpret_cast<uint64_t*>(smem + GT_OFF_BAR);
    uint64_t* hbar = wbar + 1;        // [2] h tile written (4 warp arrivals)
    uint64_t* mbar = hbar + 2;        // [2] accumulator complete (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    const int b_base = blockIdx.x * 2 * GT_ROWS;
    const int ntiles = (a.B - b_base > GT_ROWS) ? 2 : 1;

    // constants of this head -> smem
    for (int e = threadIdx.x; e < 2 * GG + GH + 1; e += GT_THREADS) {
        float v;
        if (e < GG) v = __ldg(a.b_hh + (long long)head * GG + e);
        else if (e < 2 * GG) v = __ldg(a.b_ih + (long long)head * GG + (e - GG));
        else if (e < 2 * GG + GH) v = a.w_lin ? __ldg(a.w_lin + (long long)head * GH + (e - 2 * GG)) : 0.f;
        else v = a.b_lin ? __ldg(a.b_lin + head) : 0.f;
        cst[e] = v;
    }
    if (warp == 0) {
        if (lane == 0) {
            prefetch_tmap(&tmW_hi); prefetch_tmap(&tmW_lo);
            mbar_init(wbar, 1);
            for (int s = 0; s < 2; ++s) { mbar_init(&hbar[s], 4); mbar_init(&mbar[s], 1); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<GT_TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            // W_hh hi / lo of this head: two K halves each, once
            mbar_arrive_expect_tx(wbar, 2 * GT_W_BYTES);
            tma_load_2d(smem + GT_OFF_WHI, &tmW_hi, wbar, 0, head * GG);
            tma_load_2d(smem + GT_OFF_WHI + GT_W_HALF, &tmW_hi, wbar, 32, head * GG);
            tma_load_2d(smem + GT_OFF_WLO, &tmW_lo, wbar, 0, head * GG);
            tma_load_2d(smem + GT_OFF_WLO + GT_W_HALF, &tmW_lo, wbar, 32, head * GG);
            mbar_wait(wbar, 0);
            constexpr uint32_t idesc = idesc_tf32(GT_ROWS, GG, false, false);
            const uint32_t w_hi = smem_u32(smem + GT_OFF_WHI), w_lo = smem_u32(smem + GT_OFF_WLO);
            for (int t = 0; t < a.T; ++t) {
                for (int s = 0; s < ntiles; ++s) {
                    mbar_wait(&hbar[s], t & 1);           // h_{t-1} of tile s is in smem, accumulator s has been drained
                    tc_fence_after();
                    const uint32_t h_hi = smem_u32(smem + GT_OFF_H + s * 2 * GT_H_BYTES), h_lo = h_hi + GT_H_BYTES;
                    const uint32_t acc = tmem_base + static_cast<uint32_t>(s * GG);
#pragma unroll
                    for (int kk = 0; kk < 8; ++kk) {
                        const uint32_t offA = (kk >> 2) * GT_H_HALF + (kk & 3) * 32;
                        const uint32_t offB = (kk >> 2) * GT_W_HALF + (kk & 3) * 32;
                        mma_tf32_ss(acc, smem_desc_k_sw128(h_lo + offA), smem_desc_k_sw128(w_hi + offB), idesc, kk != 0);
                        mma_tf32_ss(acc, smem_desc_k_sw128(h_hi + offA), smem_desc_k_sw128(w_lo + offB), idesc, true);
                        mma_tf32_ss(acc, smem_desc_k_sw128(h_hi + offA), smem_desc_k_sw128(w_hi + offB), idesc, true);
                    }
                    mma_commit(&mbar[s]);
                }
            }
        }
    } else {
        const int s = (warp - 1) >> 2;                     // tile of this warpgroup
        if (s < ntiles) {
            const int q = warp & 3;                        // TMEM lane quadrant accessible to this warp
            const int row = q * 32 + lane;                 // row inside the tile == TMEM lane
            const int gb = b_base + s * GT_ROWS + row;     // batch row
            const bool live = gb < a.B;
            const bool has_lin = a.w_lin != nullptr;
            uint8_t* h_hi_s = smem + GT_OFF_H + s * 2 * GT_H_BYTES;
            uint8_t* h_lo_s = h_hi_s + GT_H_BYTES;
            const uint32_t acc = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(s * GG);
            const float* bhh = cst;
            const float* bih = cst + GG;
            const float* wl = cst + 2 * GG;

            // h0 -> registers + operand tile
            float h[GH];
            {
                const float* h0 = a.h0 + (long long)head * a.h0_stride + (long long)gb * GH;
#pragma unroll
                for (int c = 0; c < GH / 4; ++c) {
                    float4 v = live ? __ldg(reinterpret_cast<const float4*>(h0) + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                    h[4 * c] = v.x; h[4 * c + 1] = v.y; h[4 * c + 2] = v.z; h[4 * c + 3] = v.w;
                }
            }
#pragma unroll
            for (int c = 0; c < 16; ++c) store_h_chunk(h_hi_s, h_lo_s, row, c, h[4 * c], h[4 * c + 1], h[4 * c + 2], h[4 * c + 3]);
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&hbar[s]);

            for (int t = 0; t < a.T; ++t) {
                const long long grow = ((long long)head * a.T + t) * a.B + gb;      // global row of (head, t, b)
                float* gdst = a.gates + grow * GG;
                const bool use_bias = t < a.t_skip;
                mbar_wait(&mbar[s], t & 1);
                tc_fence_after();
                float ps = 0.f;
#pragma unroll
                for (int jc = 0; jc < GH; jc += 16) {      // fully unrolled: h[] stays in registers
                    // gate by gate to bound register pressure: r, then z, then n / h'
                    float acc_v[16], gi_v[16], rr[16], zz[16];
                    auto load_gi = [&](int gate) {
                        if (live && !use_bias) {
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                float4 v = *reinterpret_cast<const float4*>(gdst + gate * GH + jc + 4 * c);
                                gi_v[4 * c] = v.x; gi_v[4 * c + 1] = v.y; gi_v[4 * c + 2] = v.z; gi_v[4 * c + 3] = v.w;
                            }
                        } else {
#pragma unroll
                            for (int e = 0; e < 16; ++e) gi_v[e] = bih[gate * GH + jc + e];
                        }
                    };
                    tmem_ld_32x16(acc + jc, acc_v);
                    load_gi(0);
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) rr[e] = sigmoidf_fast(gi_v[e] + (acc_v[e] + bhh[jc + e]));
                    tmem_ld_32x16(acc + GH + jc, acc_v);
                    load_gi(1);
                    if (live) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            *reinterpret_cast<float4*>(gdst + jc + 4 * c) = make_float4(rr[4 * c], rr[4 * c + 1], rr[4 * c + 2], rr[4 * c + 3]);
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) zz[e] = sigmoidf_fast(gi_v[e] + (acc_v[e] + bhh[GH + jc + e]));
                    tmem_ld_32x16(acc + 2 * GH + jc, acc_v);
                    load_gi(2);
                    if (live) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            *reinterpret_cast<float4*>(gdst + GH + jc + 4 * c) = make_float4(zz[4 * c], zz[4 * c + 1], zz[4 * c + 2], zz[4 * c + 3]);
                    }
                    tmem_ld_wait();
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float ghn_ = acc_v[e] + bhh[2 * GH + jc + e];
                        const float n = tanhf_fast(__fadd_rn(gi_v[e], __fmul_rn(rr[e], ghn_)));
                        const float hn = __fadd_rn(__fmul_rn(__fsub_rn(h[jc + e], n), zz[e]), n);
                        h[jc + e] = hn;
                        acc_v[e] = ghn_;      // reuse: gh_n
                        gi_v[e] = n;          // reuse: n
                        ps = fmaf(hn, wl[jc + e], ps);
                    }
                    if (live) {
                        float* gh_dst = a.ghn + grow * GH + jc;
                        float* hs_dst = a.hs + grow * GH + jc;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            *reinterpret_cast<float4*>(gdst + 2 * GH + jc + 4 * c) = make_float4(gi_v[4 * c], gi_v[4 * c + 1], gi_v[4 * c + 2], gi_v[4 * c + 3]);
                            *reinterpret_cast<float4*>(gh_dst + 4 * c) = make_float4(acc_v[4 * c], acc_v[4 * c + 1], acc_v[4 * c + 2], acc_v[4 * c + 3]);
                            *reinterpret_cast<float4*>(hs_dst + 4 * c) = make_float4(h[jc + 4 * c], h[jc + 4 * c + 1], h[jc + 4 * c + 2], h[jc + 4 * c + 3]);
                        }
                    }
                    if (t + 1 < a.T) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
                            store_h_chunk(h_hi_s, h_lo_s, row, (jc >> 2) + c, h[jc + 4 * c], h[jc + 4 * c + 1], h[jc + 4 * c + 2],
                                          h[jc + 4 * c + 3]);
                    }
                }
                if (live && has_lin) a.pred[grow] = ps + cst[2 * GG + GH];
                if (t + 1 < a.T) {
                    tc_fence_before();                 // accumulator reads of this step are complete
                    fence_proxy_async_smem();          // h tile (generic-proxy stores) visible to the tensor core
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&hbar[s]);
                }
            }
            tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<GT_TMEM_COLS>(tmem_base);
    }
}

}  // namespace crvae

using namespace crvae;

// Tensor-core form of crvae_gru_fwd; w_hh_hi / w_hh_lo = crvae_split_tf32(w_hh) ([P,G,H] each).
extern "C" int crvae_gru_fwd_tc(float* gates, const float* b_ih, const float* w_hh_hi, const float* w_hh_lo,
                                const float* b_hh, const float* h0, int64_t h0_head_stride, const float* w_lin,
                                const float* b_lin, float* hs, float* ghn, float* pred, int P, int T, int B, int t_skip,
                                void* stream) {
    CRVAE_REQUIRE(gates && b_ih && w_hh_hi && w_hh_lo && b_hh && h0 && hs && ghn, "null operand");
    CRVAE_REQUIRE((w_lin == nullptr) == (pred == nullptr), "w_lin and pred go together");
    CRVAE_REQUIRE(w_lin == nullptr || b_lin != nullptr, "b_lin missing");
    CRVAE_REQUIRE(P >= 0 && T > 0 && B > 0 && t_skip >= 0 && t_skip <= T, "bad size");
    CRVAE_REQUIRE(aligned16(gates) && aligned16(hs) && aligned16(ghn) && aligned16(h0) && aligned16(w_hh_hi) && aligned16(w_hh_lo),
                  "16-byte alignment");
    if (P == 0) return 0;
    CUtensorMap tW_hi, tW_lo;
    int rc;
    if ((rc = make_tmap_2d(&tW_hi, w_hh_hi, GH, (uint64_t)P * GG, GH, 32, GG, false))) return rc;
    if ((rc = make_tmap_2d(&tW_lo, w_hh_lo, GH, (uint64_t)P * GG, GH, 32, GG, false))) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gru_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM_BYTES);
        if (e != cudaSuccess) { set_error("gru_fwd_tc smem attr (%d B): %s", GT_SMEM_BYTES, cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    GruTcArgs a{gates, b_ih, b_hh, h0, (long long)h0_head_stride, w_lin, b_lin, hs, ghn, pred, P, T, B, t_skip};
    dim3 grid((B + 2 * GT_ROWS - 1) / (2 * GT_ROWS), P);
    gru_fwd_tc_kernel<<<grid, GT_THREADS, GT_SMEM_BYTES, (cudaStream_t)stream>>>(tW_hi, tW_lo, a);
    return check_launch("gru_fwd_tc_kernel");
}
