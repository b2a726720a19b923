// Weight gradient of the multi-head input projection on tcgen05 (autograd twin of umma_proj.cu):
//
//   dw_ih[i][g][k] = sum_{m = (t >= t_skip, b)} dgates[i][m][g] * x[m][k]        (CRVAE_lorenz96.py:497)
//
// (MN-major tf32 operands must use the SWIZZLE_128B_BASE32B shared-memory layout, which TMA produces with
//  CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B.)
// computed transposed, D[k][g] = X^T . dG, so that one accumulator tile is [128 input columns k
// (TMEM lanes) x 192 gate rows g (TMEM columns)] and BOTH operands are consumed in their natural
// global layouts as MN-major UMMA operands (the reduction index m is the slow index of x [m][k]
// and of dgates [m][g]): no transposed copy of the 196 MB gate-gradient buffer is ever made.
//
//   warp 0   : TMA producer.  A = x_hi / x_lo (the fixed batch, pre-split once) as 4+4 boxes
//              {32 k x 32 m}; B = raw fp32 dgates as ONE 4-D box {32 g, 32 m, 6 g-blocks, head}.
//   warps 2-5: converter: split the landed dgates tile into tf32 hi (in place) and lo (second
//              buffer) in shared memory -- the swizzled layout is preserved because the split is
//              element-wise -- then fence.proxy.async + arrive; afterwards they are the epilogue.
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer, 3xTF32 per 8-row K step.
// One CTA owns one (k-tile, head) pair and walks all row chunks; smem ring 2 x 80 KB.
#include "common.cuh"
#include "umma.cuh"

namespace crvae {

int make_tmap_2d(CUtensorMap* m, const float* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems,
                 uint32_t box_inner, uint32_t box_rows, bool atom32b);
int make_tmap_generic(CUtensorMap* m, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, bool atom32b);

constexpr int WG_BM = 128;                 // input columns k per tile (UMMA M, TMEM lanes)
constexpr int WG_BN = CRVAE_G;             // 192 gate rows (UMMA N, TMEM columns)
constexpr int WG_BK = 16;                  // reduction rows (m) per stage (more, smaller stages: deeper TMA pipeline)
constexpr int WG_STAGES = 5;                // the pipeline is (stage count x TMA->convert->MMA->release latency) bound: 4 -> 5 stages = 88 -> 80 us
constexpr int WG_A_BYTES = WG_BM * WG_BK * 4;          // 4 MN-blocks x [WG_BK rows x 128 B]
constexpr int WG_B_BYTES = WG_BN * WG_BK * 4;          // 6 MN-blocks x [WG_BK rows x 128 B]
constexpr int WG_BLOCK_BYTES = WG_BK * 128;            // 4096: one MN-block (32 M/N elements) of a stage = LBO
constexpr int WG_STAGE_BYTES = 2 * WG_A_BYTES + 2 * WG_B_BYTES;
constexpr int WG_TX_BYTES = 2 * WG_A_BYTES + WG_B_BYTES;   // bytes landed by TMA per stage (B_lo is produced in smem)
constexpr int WG_TMEM_COLS = 256;
constexpr int WG_CONV_WARPS = 8;
constexpr int WG_THREADS = 64 + 32 * WG_CONV_WARPS;
constexpr int WG_SMEM_BYTES = WG_STAGES * WG_STAGE_BYTES + 1024 + 256;

struct WgradTcArgs {
    float* dW;               // [P][G][K]  (splits == 1)  or partials [P][splits][G][K]
    const uint8_t* mask;     // [P][K] or null
    int K, R, row_off;       // R = rows reduced per head (starting at row_off within the head's T*B rows)
    int splits;              // the reduction range of a head is cut into `splits` contiguous parts (blockIdx.z)
};

__global__ void __launch_bounds__(WG_THREADS, 1)
proj_wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                     const __grid_constant__ CUtensorMap tmG, WgradTcArgs a) {
    using namespace umma;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset form keeps the shared address space (LDS/STS, not generic LD/ST)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + WG_STAGES * WG_STAGE_BYTES);
    uint64_t* conv = full + WG_STAGES;
    uint64_t* empty = conv + WG_STAGES;
    uint64_t* tmem_full = empty + WG_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k_tile = blockIdx.x, head = blockIdx.y, split = blockIdx.z;
    const int nchunks_all = (a.R + WG_BK - 1) / WG_BK;
    const int per = (nchunks_all + a.splits - 1) / a.splits;
    const int c_begin = split * per;
    const int c_end = (c_begin + per < nchunks_all) ? c_begin + per : nchunks_all;
    const int nchunks = c_end > c_begin ? c_end - c_begin : 0;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmX_hi); prefetch_tmap(&tmX_lo); prefetch_tmap(&tmG); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < WG_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&conv[s], WG_CONV_WARPS); mbar_init(&empty[s], 1); }
            mbar_init(tmem_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<WG_TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % WG_STAGES, ph = (c / WG_STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * WG_STAGE_BYTES;
                const int row = a.row_off + (c_begin + c) * WG_BK;
                mbar_arrive_expect_tx(&full[s], WG_TX_BYTES);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    tma_load_2d(st + j * WG_BLOCK_BYTES, &tmX_hi, &full[s], k_tile * WG_BM + 32 * j, row);
                    tma_load_2d(st + WG_A_BYTES + j * WG_BLOCK_BYTES, &tmX_lo, &full[s], k_tile * WG_BM + 32 * j, row);
                }
                tma_load_4d(st + 2 * WG_A_BYTES, &tmG, &full[s], 0, row, 0, head);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_tf32(WG_BM, WG_BN, true, true);
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % WG_STAGES, ph = (c / WG_STAGES) & 1;
                mbar_wait(&conv[s], ph);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + s * WG_STAGE_BYTES);
                const uint64_t a_hi = smem_desc_mn_sw128_32b(st, WG_BLOCK_BYTES), a_lo = smem_desc_mn_sw128_32b(st + WG_A_BYTES, WG_BLOCK_BYTES);
                const uint64_t b_hi = smem_desc_mn_sw128_32b(st + 2 * WG_A_BYTES, WG_BLOCK_BYTES);
                const uint64_t b_lo = smem_desc_mn_sw128_32b(st + 2 * WG_A_BYTES + WG_B_BYTES, WG_BLOCK_BYTES);
                int ksteps = (a.R - (c_begin + c) * WG_BK + 7) / 8;
                if (ksteps > WG_BK / 8) ksteps = WG_BK / 8;
                for (int k = 0; k < ksteps; ++k) {
                    const uint64_t adv = static_cast<uint64_t>(k * (1024 >> 4));   // next 8 reduction rows = next 1024-byte atom
                    mma_tf32_ss(tmem_base, a_lo + adv, b_hi + adv, idesc, (c | k) != 0);
                    mma_tf32_ss(tmem_base, a_hi + adv, b_lo + adv, idesc, true);
                    mma_tf32_ss(tmem_base, a_hi + adv, b_hi + adv, idesc, true);
                }
                mma_commit(&empty[s]);
            }
            mma_commit(tmem_full);               // (also fires when this split had no chunks)
        }
    } else {
        // converter warps 0..7: dG tile -> tf32 hi (in place) | lo.  The conversion is a latency chain (LDS -> cvt -> STS ->
        // proxy fence -> barrier) and was the pacing stage of the pipeline with four warps: eight warps, and every thread
        // issues all of its loads before the first conversion.
        const int cw = warp - 2;
        constexpr int PER = WG_B_BYTES / 16 / (32 * WG_CONV_WARPS);      // float4 per thread per chunk
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % WG_STAGES, ph = (c / WG_STAGES) & 1;
            mbar_wait(&full[s], ph);
            float4* hi = reinterpret_cast<float4*>(smem + s * WG_STAGE_BYTES + 2 * WG_A_BYTES);
            float4* lo = reinterpret_cast<float4*>(smem + s * WG_STAGE_BYTES + 2 * WG_A_BYTES + WG_B_BYTES);
            float4 v[PER];
#pragma unroll
            for (int i = 0; i < PER; ++i) v[i] = hi[cw * 32 + lane + i * 32 * WG_CONV_WARPS];
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                float4 h, l;
                uint32_t t;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v[i].x)); h.x = __uint_as_float(t); l.x = __fsub_rn(v[i].x, h.x);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v[i].y)); h.y = __uint_as_float(t); l.y = __fsub_rn(v[i].y, h.y);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v[i].z)); h.z = __uint_as_float(t); l.z = __fsub_rn(v[i].z, h.z);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v[i].w)); h.w = __uint_as_float(t); l.w = __fsub_rn(v[i].w, h.w);
                hi[cw * 32 + lane + i * 32 * WG_CONV_WARPS] = h;
                lo[cw * 32 + lane + i * 32 * WG_CONV_WARPS] = l;
            }
            fence_proxy_async_smem();                // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(&conv[s]);
        }
        if (warp < 6) {                              // the epilogue needs one warp per TMEM lane quadrant
        // epilogue: D[lane = k][col = g] -> dW[head][g][k]   (coalesced over lanes for every g)
        const int q = warp & 3;
        const int kcol = k_tile * WG_BM + q * 32 + lane;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
        float mval = 1.f;
        if (a.mask && kcol < a.K) mval = a.mask[static_cast<long long>(head) * a.K + kcol] ? 1.f : 0.f;
        float* out = a.dW + (static_cast<long long>(head) * a.splits + split) * WG_BN * a.K + kcol;
#pragma unroll 1
        for (int c0 = 0; c0 < WG_BN; c0 += 32) {
            float v[32];
            tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(c0), v);
            tmem_ld_wait();
            if (kcol < a.K) {
#pragma unroll
                for (int j = 0; j < 32; ++j) out[static_cast<long long>(c0 + j) * a.K] = nchunks > 0 ? v[j] * mval : 0.f;
            }
        }
        tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<WG_TMEM_COLS>(tmem_base);
    }
}

}  // namespace crvae

using namespace crvae;

namespace crvae {
// sum the per-split partials in fixed order: out[p][e] = sum_s ws[p][s][e]
__global__ void split_sum_kernel(const float* __restrict__ ws, float* __restrict__ out, int S, long long n) {
    const long long p = blockIdx.y;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int s = 0; s < S; ++s) acc += ws[(p * S + s) * n + e];
        out[p * n + e] = acc;
    }
}

int tc_splits_for(int P, int tiles_per_head, int nchunks, int ctas_per_sm) {
    // one CTA per (tile, head) leaves SMs idle when a rank holds few heads: cut the reduction so the grid approaches
    // the SM count without spilling into a second wave
    int s = (148 * ctas_per_sm) / (P * tiles_per_head);
    if (s < 1) s = 1;
    if (s > 16) s = 16;
    // accuracy: the tensor core accumulates in fp32 WITHOUT round-to-nearest, so one accumulator chain drifts by about
    // 2^-24 per MMA K-step (measured: 32,768 rows in one chain = 2.3e-4 relative, all entries low).  Chains are kept to
    // 256 chunks of 16 rows (512 K-steps, ~3e-5); the partial sums are added in fp32 round-to-nearest by split_sum_kernel.
    const int s_acc = (nchunks + 255) / 256;
    if (s < s_acc) s = s_acc;
    if (s > nchunks) s = nchunks;
    return s;
}
int launch_split_sum(const float* ws, float* out, int P, int S, long long n, cudaStream_t st) {
    int bx = (int)((n + 255) / 256);
    if (bx > 64) bx = 64;
    split_sum_kernel<<<dim3(bx, P), 256, 0, st>>>(ws, out, S, n);
    return check_launch("split_sum_kernel");
}
}  // namespace crvae

extern "C" size_t crvae_proj_wgrad_tc_workspace(int P, int T, int B, int K, int t_skip) {
    const int R = (T - t_skip) * B;
    const int S = min(tc_splits_for(P, (K + WG_BM - 1) / WG_BM, (R + WG_BK - 1) / WG_BK, 1), 64);      // gridDim.z <= 64
    return S > 1 ? (size_t)P * S * CRVAE_G * K * sizeof(float) : 16;
}

extern "C" int crvae_proj_wgrad_tc(const float* dgates, const float* x_hi, const float* x_lo, const uint8_t* mask,
                                   float* dw_ih, int P, int T, int B, int K, int t_skip, void* workspace, void* stream) {
    CRVAE_REQUIRE(dgates && x_hi && x_lo && dw_ih, "null operand");
    CRVAE_REQUIRE(P >= 0 && T > 0 && B > 0 && K > 0 && t_skip >= 0 && t_skip <= T, "bad size");
    CRVAE_REQUIRE(K % 4 == 0, "tensor-core weight gradient needs K % 4 == 0 (16-byte TMA row pitch); use crvae_proj_wgrad");
    CRVAE_REQUIRE(aligned16(dgates) && aligned16(x_hi) && aligned16(x_lo), "16-byte alignment");
    if (P == 0) return 0;
    const int R = (T - t_skip) * B;
    if (R == 0) return (int)cudaMemsetAsync(dw_ih, 0, (size_t)P * CRVAE_G * K * sizeof(float), (cudaStream_t)stream);
    CUtensorMap tX_hi, tX_lo, tG;
    int rc;
    if ((rc = make_tmap_2d(&tX_hi, x_hi, K, (uint64_t)T * B, K, 32, WG_BK, true))) return rc;
    if ((rc = make_tmap_2d(&tX_lo, x_lo, K, (uint64_t)T * B, K, 32, WG_BK, true))) return rc;
    // dgates [P][T*B][192] viewed as {32 g_in, T*B rows, 6 g-blocks, P heads}; box {32, 32, 6, 1}
    const uint64_t dims[4] = {32, (uint64_t)T * B, 6, (uint64_t)P};
    const uint64_t strides[3] = {(uint64_t)CRVAE_G * 4, 128, (uint64_t)T * B * CRVAE_G * 4};
    const uint32_t box[4] = {32, WG_BK, 6, 1};
    if ((rc = make_tmap_generic(&tG, dgates, 4, dims, strides, box, true))) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(proj_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM_BYTES);
        if (e != cudaSuccess) { set_error("proj_wgrad_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    const int S = min(tc_splits_for(P, (K + WG_BM - 1) / WG_BM, (R + WG_BK - 1) / WG_BK, 1), 64);      // gridDim.z <= 64
    if (S > 1) CRVAE_REQUIRE(workspace && aligned16(workspace), "workspace required (crvae_proj_wgrad_tc_workspace)");
    WgradTcArgs a{S > 1 ? (float*)workspace : dw_ih, mask, K, R, t_skip * B, S};
    dim3 grid((K + WG_BM - 1) / WG_BM, P, S);
    proj_wgrad_tc_kernel<<<grid, WG_THREADS, WG_SMEM_BYTES, (cudaStream_t)stream>>>(tX_hi, tX_lo, tG, a);
    rc = check_launch("proj_wgrad_tc_kernel");
    if (rc || S == 1) return rc;
    return launch_split_sum((const float*)workspace, dw_ih, P, S, (long long)CRVAE_G * K, (cudaStream_t)stream);
}
