// Recurrent weight gradient dW_hh on tcgen05, taken OUT of the BPTT kernel's time loop:
//
//   dw_hh[i][g][k] = sum_{t,b} dgh[i][t][b][g] * h_{t-1}[i][b][k]                (autograd of nn.GRU's linear_hh, :497)
//
// The BPTT kernel (gru_recurrent.cu) needs dh_{t-1} = dgh . W_hh sequentially, but the accumulation of dW_hh is a
// plain reduction over all (t, b) rows -- a [192 x T*B] . [T*B x 64] GEMM per head.  Doing it here halves the FFMA
// work of the BPTT kernel, frees its 48 accumulator registers and its h_{t-1} tile (-> 2 CTAs per SM).
//
// Operands are consumed in their natural row-major layouts as MN-major UMMA operands (SWIZZLE_128B_BASE32B):
//   A tile 0 = dgi[:, 0:128]  (r and z parts; dgh == dgi there) from the gate buffer   [m][192]
//   A tile 1 = dgh_n          (= da_n * r, written by the BPTT kernel over gh_n)       [m][64]
//   B        = h_{t-1}        = h0 for the first B rows, hs shifted by one step after  [m][64]
// Two accumulators [128 lanes (g) x 64 columns (k)] in TMEM; 3xTF32 with the tf32 hi/lo split of ALL operands done
// in shared memory by the converter warps.  One CTA per head walks all T*B rows; 3-stage 64 KB ring.
#include "common.cuh"
#include "umma.cuh"

namespace crvae {

int make_tmap_generic(CUtensorMap* m, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, bool atom32b);
int tc_splits_for(int P, int tiles_per_head, int nchunks, int ctas_per_sm);
int launch_split_sum(const float* ws, float* out, int P, int S, long long n, cudaStream_t st);

constexpr int DH_H = CRVAE_HIDDEN;
constexpr int DH_G = CRVAE_G;
constexpr int DH_BK = 16;                          // reduction rows per stage
constexpr int DH_STAGES = 3;                       // 3 x 32 KB: two CTAs share an SM
constexpr int DH_BLOCK = DH_BK * 128;              // 4096 B: one MN-block (32 elements wide) of a stage = LBO
constexpr int DH_OFF_A0 = 0;                       // 4 blocks: g 0..127
constexpr int DH_OFF_A1 = 4 * DH_BLOCK;            // 2 blocks: g 128..191
constexpr int DH_OFF_B = 6 * DH_BLOCK;             // 2 blocks: k 0..63
constexpr int DH_HALF = 8 * DH_BLOCK;              // 32768 B of raw/hi data, followed by 32768 B of lo data
constexpr int DH_STAGE_BYTES = 2 * DH_HALF;
constexpr int DH_TMEM_COLS = 128;
constexpr int DH_CONV_WARPS = 8;
constexpr int DH_THREADS = 64 + 32 * DH_CONV_WARPS;
constexpr int DH_SMEM_BYTES = DH_STAGES * DH_STAGE_BYTES + 1024 + 256;

struct DwhhArgs {
    float* dw_hh;      // [P][G][H]  (splits == 1)  or partials [P][splits][G][H]
    int rows, B;       // rows = T*B
    int h0_per_head;
    int splits;        // the T*B rows of a head are cut into `splits` contiguous parts (blockIdx.y)
};

__global__ void __launch_bounds__(DH_THREADS, 1)
gru_dwhh_tc_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmN,
                   const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmZ, DwhhArgs a) {
    using namespace umma;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset form keeps the shared address space (LDS/STS, not generic LD/ST)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + DH_STAGES * DH_STAGE_BYTES);
    uint64_t* conv = full + DH_STAGES;
    uint64_t* empty = conv + DH_STAGES;
    uint64_t* tmem_full = empty + DH_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.x, split = blockIdx.y;
    const int nchunks_all = (a.rows + DH_BK - 1) / DH_BK;
    const int per = (nchunks_all + a.splits - 1) / a.splits;
    const int c_begin = split * per;
    const int c_end = (c_begin + per < nchunks_all) ? c_begin + per : nchunks_all;
    const int nchunks = c_end > c_begin ? c_end - c_begin : 0;

    if (warp == 0 && lane == 0) { prefetch_tmap(&tmG); prefetch_tmap(&tmN); prefetch_tmap(&tmH); prefetch_tmap(&tmZ); }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < DH_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&conv[s], DH_CONV_WARPS); mbar_init(&empty[s], 1); }
            mbar_init(tmem_full, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<DH_TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % DH_STAGES, ph = (c / DH_STAGES) & 1;
                mbar_wait(&empty[s], ph ^ 1);
                uint8_t* st = smem + s * DH_STAGE_BYTES;
                const int m0 = (c_begin + c) * DH_BK;
                mbar_arrive_expect_tx(&full[s], DH_HALF);
                tma_load_4d(st + DH_OFF_A0, &tmG, &full[s], 0, m0, 0, head);
                tma_load_4d(st + DH_OFF_A1, &tmN, &full[s], 0, m0, 0, head);
                if (m0 < a.B) tma_load_4d(st + DH_OFF_B, &tmZ, &full[s], 0, m0, 0, a.h0_per_head ? head : 0);   // h_{-1} = h0
                else tma_load_4d(st + DH_OFF_B, &tmH, &full[s], 0, m0 - a.B, 0, head);                          // h_{t-1}
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_tf32(128, DH_H, true, true);
            for (int c = 0; c < nchunks; ++c) {
                const int s = c % DH_STAGES, ph = (c / DH_STAGES) & 1;
                mbar_wait(&conv[s], ph);
                tc_fence_after();
                const uint32_t st = smem_u32(smem + s * DH_STAGE_BYTES);
                const uint64_t b_hi = smem_desc_mn_sw128_32b(st + DH_OFF_B, DH_BLOCK), b_lo = smem_desc_mn_sw128_32b(st + DH_HALF + DH_OFF_B, DH_BLOCK);
                int ksteps = (a.rows - (c_begin + c) * DH_BK + 7) / 8;
                if (ksteps > DH_BK / 8) ksteps = DH_BK / 8;
#pragma unroll
                for (int tile = 0; tile < 2; ++tile) {
                    const uint32_t offA = tile ? DH_OFF_A1 : DH_OFF_A0;
                    const uint64_t a_hi = smem_desc_mn_sw128_32b(st + offA, DH_BLOCK), a_lo = smem_desc_mn_sw128_32b(st + DH_HALF + offA, DH_BLOCK);
                    const uint32_t acc = tmem_base + static_cast<uint32_t>(tile * DH_H);
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t adv = static_cast<uint64_t>(k * (1024 >> 4));    // next 8 reduction rows
                        mma_tf32_ss(acc, a_lo + adv, b_hi + adv, idesc, (c | k) != 0);
                        mma_tf32_ss(acc, a_hi + adv, b_lo + adv, idesc, true);
                        mma_tf32_ss(acc, a_hi + adv, b_hi + adv, idesc, true);
                    }
                }
                mma_commit(&empty[s]);
            }
            mma_commit(tmem_full);
        }
    } else {
        // converter warps 0..7 (the conversion chain LDS -> cvt -> STS -> proxy fence -> barrier paces the pipeline: eight
        // warps, all loads of a thread issued before the first conversion)
        const int cw = warp - 2;
        constexpr int PER = DH_HALF / 16 / (32 * DH_CONV_WARPS);      // float4 per thread per chunk
        for (int c = 0; c < nchunks; ++c) {
            const int s = c % DH_STAGES, ph = (c / DH_STAGES) & 1;
            mbar_wait(&full[s], ph);
            float4* hi = reinterpret_cast<float4*>(smem + s * DH_STAGE_BYTES);
            float4* lo = reinterpret_cast<float4*>(smem + s * DH_STAGE_BYTES + DH_HALF);
            float4 v[PER];
#pragma unroll
            for (int i = 0; i < PER; ++i) v[i] = hi[cw * 32 + lane + i * 32 * DH_CONV_WARPS];
#pragma unroll
            for (int i = 0; i < PER; ++i) {
                float4 h, l;
                uint32_t t;
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v[i].x)); h.x = __uint_as_float(t); l.x = __fsub_rn(v[i].x, h.x);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v[i].y)); h.y = __uint_as_float(t); l.y = __fsub_rn(v[i].y, h.y);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v[i].z)); h.z = __uint_as_float(t); l.z = __fsub_rn(v[i].z, h.z);
                asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v[i].w)); h.w = __uint_as_float(t); l.w = __fsub_rn(v[i].w, h.w);
                hi[cw * 32 + lane + i * 32 * DH_CONV_WARPS] = h;
                lo[cw * 32 + lane + i * 32 * DH_CONV_WARPS] = l;
            }
            fence_proxy_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&conv[s]);
        }
        if (warp < 6) {                              // the epilogue needs one warp per TMEM lane quadrant
        // epilogue: accumulator `tile` holds dW_hh rows g = tile*128 + lane-index, 64 columns k
        const int q = warp & 3;
        const int gl = q * 32 + lane;
        mbar_wait(tmem_full, 0);
        tc_fence_after();
#pragma unroll 1
        for (int tile = 0; tile < 2; ++tile) {
            const int g = tile * 128 + gl;
#pragma unroll 1
            for (int c0 = 0; c0 < DH_H; c0 += 32) {
                float v[32];
                tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q * 32) << 16) + static_cast<uint32_t>(tile * DH_H + c0), v);
                tmem_ld_wait();
                if (g < DH_G) {
                    float* dst = a.dw_hh + ((static_cast<long long>(head) * a.splits + split) * DH_G + g) * DH_H + c0;
                    if (nchunks == 0) {
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] = 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                }
            }
        }
        tc_fence_before();
        }
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<DH_TMEM_COLS>(tmem_base);
    }
}

}  // namespace crvae

using namespace crvae;

// dgates [P,T,B,G] (dgi, as left by crvae_gru_bwd), dghn [P,T,B,H] (= da_n*r, written by crvae_gru_bwd in defer mode),
// hs [P,T,B,H], h0 [B,H] (stride 0) or [P,B,H].  Needs B % 32 == 0.
extern "C" size_t crvae_gru_dwhh_tc_workspace(int P, int T, int B) {
    const int S = tc_splits_for(P, 1, (T * B + DH_BK - 1) / DH_BK, 2);
    return S > 1 ? (size_t)P * S * DH_G * DH_H * sizeof(float) : 16;
}

extern "C" int crvae_gru_dwhh_tc(const float* dgates, const float* dghn, const float* hs, const float* h0,
                                 int64_t h0_head_stride, float* dw_hh, int P, int T, int B, void* workspace, void* stream) {
    CRVAE_REQUIRE(dgates && dghn && hs && h0 && dw_hh, "null operand");
    CRVAE_REQUIRE(P >= 0 && T > 0 && B > 0 && B % 32 == 0, "bad size (B must be a multiple of 32)");
    CRVAE_REQUIRE(aligned16(dgates) && aligned16(dghn) && aligned16(hs) && aligned16(h0) && aligned16(dw_hh), "16-byte alignment");
    if (P == 0) return 0;
    const uint64_t TB = (uint64_t)T * B;
    CUtensorMap tG, tN, tH, tZ;
    int rc;
    {   // r|z part of the gate-gradient buffer: {32 g_in, rows, 4 g-blocks, P}
        const uint64_t dims[4] = {32, TB, 4, (uint64_t)P};
        const uint64_t str[3] = {(uint64_t)DH_G * 4, 128, TB * DH_G * 4};
        const uint32_t box[4] = {32, DH_BK, 4, 1};
        if ((rc = make_tmap_generic(&tG, dgates, 4, dims, str, box, true))) return rc;
    }
    {
        const uint64_t dims[4] = {32, TB, 2, (uint64_t)P};
        const uint64_t str[3] = {(uint64_t)DH_H * 4, 128, TB * DH_H * 4};
        const uint32_t box[4] = {32, DH_BK, 2, 1};
        if ((rc = make_tmap_generic(&tN, dghn, 4, dims, str, box, true))) return rc;
        if ((rc = make_tmap_generic(&tH, hs, 4, dims, str, box, true))) return rc;
    }
    {
        const int per_head = h0_head_stride != 0;
        const uint64_t dims[4] = {32, (uint64_t)B, 2, (uint64_t)(per_head ? P : 1)};
        const uint64_t str[3] = {(uint64_t)DH_H * 4, 128, (uint64_t)(per_head ? h0_head_stride : (int64_t)B * DH_H) * 4};
        const uint32_t box[4] = {32, DH_BK, 2, 1};
        if ((rc = make_tmap_generic(&tZ, h0, 4, dims, str, box, true))) return rc;
    }
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gru_dwhh_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, DH_SMEM_BYTES);
        if (e != cudaSuccess) { set_error("gru_dwhh_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    const int S = tc_splits_for(P, 1, ((int)TB + DH_BK - 1) / DH_BK, 2);
    if (S > 1) CRVAE_REQUIRE(workspace && aligned16(workspace), "workspace required (crvae_gru_dwhh_tc_workspace)");
    DwhhArgs a{S > 1 ? (float*)workspace : dw_hh, (int)TB, B, h0_head_stride != 0, S};
    gru_dwhh_tc_kernel<<<dim3(P, S), DH_THREADS, DH_SMEM_BYTES, (cudaStream_t)stream>>>(tG, tN, tH, tZ, a);
    rc = check_launch("gru_dwhh_tc_kernel");
    if (rc || S == 1) return rc;
    return launch_split_sum((const float*)workspace, dw_hh, P, S, (long long)DH_G * DH_H, (cudaStream_t)stream);
}
