// Persistent multi-head GRU recurrence, FORWARD, on the 5th-generation tensor cores (tcgen05).
//
// One CTA owns one (head, 128-row batch tile) pair for ALL timesteps:
//   * W_hh of the head (tf32 hi and lo parts, 2 x 48 KB) is brought in ONCE by TMA and stays in shared
//     memory (K-major, SWIZZLE_128B) as the B operand;
//   * the hidden state is the A operand and lives in TENSOR MEMORY (tf32 hi | lo, 2 x 64 columns): the
//     pointwise threads write h_t there with tcgen05.st, no shared-memory round trip, no swizzle;
//   * per step ONE thread issues the gate GEMM  gh[128 x 192] = h[128 x 64] . W_hh^T  as 8 K-steps x 3
//     tcgen05.mma (3xTF32: h_lo.W_hi + h_hi.W_lo + h_hi.W_hi, fp32 accumulate in TMEM);
//   * 16 warps do the gate math on 16x256b TMEM fragments (warp = lane quadrant x 16-column group): a
//     thread owns 4 batch rows x 4 hidden units for the whole sequence, so h stays in 16 registers, and
//     every global access is a float2 whose quad covers one full 32-byte sector (8 sectors per request,
//     all bytes used) in the row-major activation layout -- no staging buffer needed;
//   * gi of step t+1 does not depend on the recurrence: its 24 float2 loads per thread are issued before
//     the step barrier, so ~96 KB of loads per SM are in flight while the tensor core runs step t+1.
// Same buffers / semantics as gru_fwd_kernel (gru_recurrent.cu), which remains the exact-fp32 path.
//
// Reference arithmetic replaced: nn.GRU per-step linear_hh + cell (CRVAE_lorenz96.py:119) and
// nn.Linear(H,1) (:120).
#include "gru_tc_common.cuh"

namespace crvae {

constexpr int GH = CRVAE_HIDDEN;                 // 64
constexpr int GG = CRVAE_G;                      // 192
constexpr int GT_ROWS = 128;                     // rows per tile (UMMA M)
constexpr int GT_W_HALF = GG * 128;              // 24576 B: [192 rows x 32 k] fp32, one K half
constexpr int GT_W_BYTES = 2 * GT_W_HALF;        // 49152 B per (hi | lo)
constexpr int GT_OFF_WHI = 0;
constexpr int GT_OFF_WLO = GT_W_BYTES;
constexpr int GT_OFF_PRED = 2 * GT_W_BYTES;                      // [2 parities][128 rows][2 column groups] (+ slack)
constexpr int GT_OFF_CONST = GT_OFF_PRED + 2 * GT_ROWS * 4 * 4;  // b_ih[192] | b_hh[192]
constexpr int GT_OFF_BAR = GT_OFF_CONST + (2 * GG + GH) * 4;          // ... | w_lin[64]
constexpr int GT_SMEM_BYTES = GT_OFF_BAR + 64 + 1024;
constexpr int GT_TMEM_COLS = 512;
constexpr int GT_ACOL_HI = 0, GT_ACOL_LO = GH, GT_DCOL = 2 * GH;  // TMEM columns: h_hi | h_lo | gh[192]
constexpr int GT_THREADS = 512;

struct GruTcArgs {
    float* gates; const float* b_ih; const float* b_hh;
    const float* h0; long long h0_stride;
    const float* w_lin; const float* b_lin;
    float* hs; float* ghn; float* pred;
    int P, T, B, t_skip;
};

__global__ void __launch_bounds__(GT_THREADS, 1)
gru_fwd_tc_kernel(GruTcArgs a, const float* __restrict__ w_hi, const float* __restrict__ w_lo) {
    using namespace umma;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset form keeps the shared address space (LDS/STS, not generic LD/ST)
    float* pred_s = reinterpret_cast<float*>(smem + GT_OFF_PRED);
    float* bih_s = reinterpret_cast<float*>(smem + GT_OFF_CONST);
    float* bhh_s = bih_s + GG;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + GT_OFF_BAR);      // [3] gate r | z | n accumulator complete (tcgen05.commit)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(mbar + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y;
    const int b_base = blockIdx.x * GT_ROWS;
    const int q = warp & 3, hh = (warp >> 2) & 1, cg = warp >> 3;   // TMEM lane quadrant, 16-lane half, 32-column group
    const int tr = lane >> 2, tq = lane & 3;
    const bool has_lin = a.w_lin != nullptr;

    for (int e = threadIdx.x; e < 2 * GG + GH; e += GT_THREADS)
        bih_s[e] = e < GG       ? __ldg(a.b_ih + (long long)head * GG + e)
                   : e < 2 * GG ? __ldg(a.b_hh + (long long)head * GG + (e - GG))
                                : (has_lin ? __ldg(a.w_lin + (long long)head * GH + (e - 2 * GG)) : 0.f);
    if (warp == 0) {
        if (lane == 0) {
            for (int g = 0; g < 3; ++g) mbar_init(&mbar[g], 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<GT_TMEM_COLS>(tmem_slot);
    }
    // W_hh hi / lo of this head -> shared memory B operand (K-major, SWIZZLE_128B), rows and columns permuted
    {
        const float4* src_hi = reinterpret_cast<const float4*>(w_hi + (long long)head * GG * GH);
        const float4* src_lo = reinterpret_cast<const float4*>(w_lo + (long long)head * GG * GH);
        for (int idx = threadIdx.x; idx < GG * (GH / 4); idx += GT_THREADS) {
            const int r = idx >> 4, c4 = idx & 15;                       // source gate row, 4-unit chunk
            const int n = (r / GH) * GH + pos_of_unit(r % GH);           // destination row
            const int kp = pos_of_unit(4 * c4);                          // units 4c4, 4c4+1 -> kp, kp+1 ; 4c4+2, 4c4+3 -> kp+8, kp+9
            const int half = kp >> 5, k0 = kp & 31, k1 = k0 + 8;
            const uint32_t rowoff = half * GT_W_HALF + n * 128;
            const uint32_t o0 = rowoff + ((((k0 >> 2) ^ (n & 7)) << 4) | ((k0 & 3) << 2));
            const uint32_t o1 = rowoff + ((((k1 >> 2) ^ (n & 7)) << 4) | ((k1 & 3) << 2));
            float4 vh = __ldg(src_hi + idx), vl;
            if (w_lo != nullptr) {
                vl = __ldg(src_lo + idx);
            } else {        // w_hi is the fp32 matrix: split here
                const float4 v = vh;
                vh = make_float4(tf32_rna(v.x), tf32_rna(v.y), tf32_rna(v.z), tf32_rna(v.w));
                vl = make_float4(__fsub_rn(v.x, vh.x), __fsub_rn(v.y, vh.y), __fsub_rn(v.z, vh.z), __fsub_rn(v.w, vh.w));
            }
            *reinterpret_cast<float2*>(smem + GT_OFF_WHI + o0) = make_float2(vh.x, vh.y);
            *reinterpret_cast<float2*>(smem + GT_OFF_WHI + o1) = make_float2(vh.z, vh.w);
            *reinterpret_cast<float2*>(smem + GT_OFF_WLO + o0) = make_float2(vl.x, vl.y);
            *reinterpret_cast<float2*>(smem + GT_OFF_WLO + o1) = make_float2(vl.z, vl.w);
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // this thread's elements: batch rows brow0 + 8rr, hidden units ucol + m (m < 8);  fragment register
    // k = 4(m >> 1) + 2rr + (m & 1)
    const int ucol = 32 * cg + 8 * tq;
    const int brow0 = b_base + 32 * q + 16 * hh + tr;
    const float* wl = bih_s + 2 * GG + ucol;
    const float blin = (has_lin && a.b_lin) ? __ldg(a.b_lin + head) : 0.f;

    float h[16], gi[3][16];
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(32 * q + 16 * hh) << 16) + static_cast<uint32_t>(32 * cg);

    // gi of gate g for step t -> gi[g] (bias only while t < t_skip)
    auto load_gi_gate = [&](int t, int g) {
        if (t < a.t_skip) {
#pragma unroll
            for (int k = 0; k < 16; ++k) gi[g][k] = bih_s[g * GH + ucol + 2 * (k >> 2) + (k & 1)];
            return;
        }
        const float* gbase = a.gates + ((long long)head * a.T + t) * a.B * GG + g * GH + ucol;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int b = brow0 + 8 * rr;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (b < a.B) ldg_v8_stream(gbase + (long long)b * GG, v);
#pragma unroll
            for (int m = 0; m < 8; ++m) gi[g][4 * (m >> 1) + 2 * rr + (m & 1)] = v[m];
        }
    };
    // a 16-register fragment (2 rows x 8 units) -> rows brow0, brow0 + 8 of a row-major buffer (256-bit stores)
    auto store_frag = [&](float* base, int ld, const float* f) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int b = brow0 + 8 * rr;
            const int k0 = 2 * rr;          // register of unit m: k0 + 4(m >> 1) + (m & 1)
            if (b < a.B) stg_v8(base + (long long)b * ld, f[k0], f[k0 + 1], f[k0 + 4], f[k0 + 5], f[k0 + 8], f[k0 + 9], f[k0 + 12], f[k0 + 13]);
        }
    };
    auto store_h_operand = [&]() {     // h (registers) -> TMEM A operand, tf32 hi | lo
        float hi[16], lo[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            hi[k] = tf32_rna(h[k]);
            lo[k] = __fsub_rn(h[k], hi[k]);
        }
        tmem_st_16x32(lane_addr + GT_ACOL_HI, hi);
        tmem_st_16x32(lane_addr + GT_ACOL_LO, lo);
        tmem_st_wait();
    };
    // thread 0: the gate GEMM of one step, gate by gate (3 x 8 K-steps x 3 MMAs of N = 64) with one commit per gate, so
    // the pointwise warps start on r while the tensor core is still working on z and n
    auto issue_step = [&]() {
        tc_fence_after();
        constexpr uint32_t idesc = idesc_tf32(GT_ROWS, GH, false, false);
        const uint32_t w_hi_s = smem_u32(smem + GT_OFF_WHI), w_lo_s = smem_u32(smem + GT_OFF_WLO);
#pragma unroll
        for (int g = 0; g < 3; ++g) {
            const uint32_t acc = tmem_base + GT_DCOL + g * GH;
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
                const uint32_t offB = (kk >> 2) * GT_W_HALF + g * GH * 128 + (kk & 3) * 32;
                const uint32_t a_hi = tmem_base + GT_ACOL_HI + kk * 8, a_lo = tmem_base + GT_ACOL_LO + kk * 8;
                mma_tf32_ts(acc, a_lo, smem_desc_k_sw128(w_hi_s + offB), idesc, kk != 0);
                mma_tf32_ts(acc, a_hi, smem_desc_k_sw128(w_lo_s + offB), idesc, true);
                mma_tf32_ts(acc, a_hi, smem_desc_k_sw128(w_hi_s + offB), idesc, true);
            }
            mma_commit(&mbar[g]);
        }
    };

    // h0 -> registers -> operand
    {
        const float* h0 = a.h0 + (long long)head * a.h0_stride + ucol;
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int b = brow0 + 8 * rr;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (b < a.B) {
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(h0 + (long long)b * GH));
                const float4 v1 = __ldg(reinterpret_cast<const float4*>(h0 + (long long)b * GH) + 1);
                v[0] = v0.x; v[1] = v0.y; v[2] = v0.z; v[3] = v0.w; v[4] = v1.x; v[5] = v1.y; v[6] = v1.z; v[7] = v1.w;
            }
#pragma unroll
            for (int m = 0; m < 8; ++m) h[4 * (m >> 1) + 2 * rr + (m & 1)] = v[m];
        }
    }
    store_h_operand();
#pragma unroll
    for (int g = 0; g < 3; ++g) load_gi_gate(0, g);
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x == 0) issue_step();

    for (int t = 0; t < a.T; ++t) {
        const long long trow = ((long long)head * a.T + t) * a.B;       // global row of (head, t, b = 0)
        float* pred_t = pred_s + (t & 1) * GT_ROWS * 2;
        {
            // The LSU queue is in order: a prefetch issued after the step's stores waits for all of them to drain.  So every
            // gate's result is stored, and the same gate's gi of step t+1 requested, as soon as that gate is done.
            float ar[16], az[16], an[16];
            const bool more = t + 1 < a.T;
            mbar_wait(&mbar[0], t & 1);
            tc_fence_after();
            tmem_ld_16x32(lane_addr + GT_DCOL, ar);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int m = 2 * (k >> 2) + (k & 1);      // unit offset of register k
                ar[k] = sigmoidf_fast(gi[0][k] + (ar[k] + bhh_s[ucol + m]));          // r
            }
            store_frag(a.gates + trow * GG + ucol, GG, ar);
            if (more) load_gi_gate(t + 1, 0);
            mbar_wait(&mbar[1], t & 1);
            tc_fence_after();
            tmem_ld_16x32(lane_addr + GT_DCOL + GH, az);
            tmem_ld_wait();
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                const int m = 2 * (k >> 2) + (k & 1);
                az[k] = sigmoidf_fast(gi[1][k] + (az[k] + bhh_s[GH + ucol + m]));     // z
            }
            store_frag(a.gates + trow * GG + GH + ucol, GG, az);
            if (more) load_gi_gate(t + 1, 1);
            mbar_wait(&mbar[2], t & 1);
            tc_fence_after();
            tmem_ld_16x32(lane_addr + GT_DCOL + 2 * GH, an);
            tmem_ld_wait();
            float ps[2] = {0.f, 0.f};
#pragma unroll
            for (int k = 0; k < 16; ++k) {                 // gh_n replaces the accumulator, n replaces gi_n
                const int m = 2 * (k >> 2) + (k & 1);
                const float ghn_ = an[k] + bhh_s[2 * GH + ucol + m];
                const float n = tanhf_fast(__fadd_rn(gi[2][k], __fmul_rn(ar[k], ghn_)));
                const float hn = __fadd_rn(__fmul_rn(__fsub_rn(h[k], n), az[k]), n);
                h[k] = hn;
                an[k] = ghn_;
                gi[2][k] = n;
                ps[(k >> 1) & 1] = fmaf(hn, wl[m], ps[(k >> 1) & 1]);
            }
            store_frag(a.gates + trow * GG + 2 * GH + ucol, GG, gi[2]);
            if (more) {
                load_gi_gate(t + 1, 2);
                store_h_operand();
            }
            store_frag(a.ghn + trow * GH + ucol, GH, an);
            store_frag(a.hs + trow * GH + ucol, GH, h);
            if (has_lin) {
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    float p = ps[rr];
                    p += __shfl_xor_sync(0xffffffffu, p, 1);
                    p += __shfl_xor_sync(0xffffffffu, p, 2);
                    if (tq == 0) pred_t[(32 * q + 16 * hh + tr + 8 * rr) * 2 + cg] = p;
                }
            }
        }
        tc_fence_before();             // accumulator reads / operand writes of this step are complete
        __syncthreads();
        if (threadIdx.x == 0 && t + 1 < a.T) issue_step();
        if (has_lin && threadIdx.x < GT_ROWS && b_base + threadIdx.x < a.B) {
            const float2 p2 = *reinterpret_cast<const float2*>(pred_t + threadIdx.x * 2);
            a.pred[trow + b_base + threadIdx.x] = (p2.x + p2.y) + blin;
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<GT_TMEM_COLS>(tmem_base);
    }
}

}  // namespace crvae

using namespace crvae;

// Tensor-core form of crvae_gru_fwd; w_hh_hi / w_hh_lo = crvae_split_tf32(w_hh) ([P,G,H] each), or w_hh_hi = the
// fp32 W_hh and w_hh_lo = NULL (the kernel splits while staging the operand).
extern "C" int crvae_gru_fwd_tc(float* gates, const float* b_ih, const float* w_hh_hi, const float* w_hh_lo,
                                const float* b_hh, const float* h0, int64_t h0_head_stride, const float* w_lin,
                                const float* b_lin, float* hs, float* ghn, float* pred, int P, int T, int B, int t_skip,
                                void* stream) {
    CRVAE_REQUIRE(gates && b_ih && w_hh_hi && b_hh && h0 && hs && ghn, "null operand");
    CRVAE_REQUIRE((w_lin == nullptr) == (pred == nullptr), "w_lin and pred go together");
    CRVAE_REQUIRE(w_lin == nullptr || b_lin != nullptr, "b_lin missing");
    CRVAE_REQUIRE(P >= 0 && T > 0 && B > 0 && t_skip >= 0 && t_skip <= T, "bad size");
    CRVAE_REQUIRE(aligned16(gates) && aligned16(hs) && aligned16(ghn) && aligned16(h0) && aligned16(w_hh_hi) && (w_hh_lo == nullptr || aligned16(w_hh_lo)),
                  "16-byte alignment");
    CRVAE_REQUIRE(h0_head_stride % 4 == 0, "h0 head stride must keep 16-byte alignment");
    CRVAE_REQUIRE((reinterpret_cast<uintptr_t>(gates) & 31u) == 0 && (reinterpret_cast<uintptr_t>(hs) & 31u) == 0 &&
                      (reinterpret_cast<uintptr_t>(ghn) & 31u) == 0,
                  "32-byte alignment of the activation buffers");
    if (P == 0) return 0;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gru_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM_BYTES);
        if (e != cudaSuccess) { set_error("gru_fwd_tc smem attr (%d B): %s", GT_SMEM_BYTES, cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    GruTcArgs a{gates, b_ih, b_hh, h0, (long long)h0_head_stride, w_lin, b_lin, hs, ghn, pred, P, T, B, t_skip};
    dim3 grid((B + GT_ROWS - 1) / GT_ROWS, P);
    gru_fwd_tc_kernel<<<grid, GT_THREADS, GT_SMEM_BYTES, (cudaStream_t)stream>>>(a, w_hh_hi, w_hh_lo);
    return check_launch("gru_fwd_tc_kernel");
}
