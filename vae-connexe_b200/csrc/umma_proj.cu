// Multi-head input projection on the 5th-generation tensor cores (tcgen05), fed by TMA.
//
//   gates[i][t][b][:] = b_ih[i] + x[t][b][:] . w_ih[i]^T        (GRU.forward, CRVAE_lorenz96.py:115-119)
//
// tcgen05 has no fp32 MMA (kind::tf32 / f16 / f8 only), and the path must agree with the fp32
// reference to 1e-4, so the contraction is error-compensated "3xTF32": every operand is pre-split
// into hi = tf32(v) and lo = v - hi (exact in fp32) and  A.B ~= Alo.Bhi + Ahi.Blo + Ahi.Bhi
// is accumulated in fp32 in TMEM (the dropped Alo.Blo term is ~2^-22 relative).
//
// Persistent kernel, one CTA per SM; a tile = [128 (t,b) rows x 192 gate columns] of one head:
//   warp 0   : TMA producer  (4 boxes per 32-wide K chunk: A_hi, A_lo [128x32], B_hi, B_lo [192x32], SWIZZLE_128B)
//   warp 1   : TMEM allocator + single-thread tcgen05.mma issuer (3 MMAs per 8-wide K step)
//   warps 2-5: epilogue, tcgen05.ld of the accumulator (lane = row), per-warp smem transpose, + bias -> global
// smem ring: 2 stages x 80 KB, full/empty mbarriers; TWO accumulators of 128 lanes x 192 fp32 columns in TMEM
// (tmem_full / tmem_empty mbarriers) so the epilogue of one tile overlaps the MMAs of the next.
#include <stdlib.h>
#include "gru_tc_common.cuh"

namespace crvae {

constexpr int TC_BM = 128;          // rows per tile (UMMA M)
constexpr int TC_BN = CRVAE_G;      // 192 gate columns (UMMA N)
constexpr int TC_BK = 32;           // K elements per stage chunk (128-byte rows)
constexpr int TC_STAGES = 2;
constexpr int TC_A_BYTES = TC_BM * TC_BK * 4;    // 16384
constexpr int TC_B_BYTES = TC_BN * TC_BK * 4;    // 24576
constexpr int TC_STAGE_BYTES = 2 * TC_A_BYTES + 2 * TC_B_BYTES;   // 81920
constexpr int TC_TMEM_COLS = 512;   // two 192-column accumulators (double buffered)
constexpr int TC_EPI_WARPS = 8;     // lane quadrant x half of the gate columns
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_SMEM_BYTES = TC_STAGES * TC_STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;

struct ProjTcArgs {
    float* C;               // gates + t_skip*B*G
    const float* bias;      // [P][G]
    long long c_head_stride;
    int M, K, n_mtiles, n_tiles;
};

// Epilogue warps (shared by both projection kernels): accumulator `ab` of tile number i -> + bias -> 256-bit stores.
// Tiles are first, first + step, ... (< last).
__device__ __forceinline__ void proj_epilogue(const ProjTcArgs& a, uint32_t tmem_base, uint64_t* tmem_full, uint64_t* tmem_empty, int warp,
                                              int lane, int first, int step, int last) {
    using namespace umma;
        const int q = warp & 3;                     // TMEM lane quadrant this warp may access
        const int chalf = (warp - 2) >> 2;          // gate columns [96 chalf, 96 chalf + 96)
        const int tr = lane >> 2, tq = lane & 3;
        int i = 0;
        for (int tile = first; tile < last; tile += step, ++i) {
            const int head = tile / a.n_mtiles, m_tile = tile % a.n_mtiles;
            const int ab = i & 1;
            const float* bias = a.bias + static_cast<long long>(head) * TC_BN + 96 * chalf + 8 * tq;
            float bv[3][8];
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + 32 * f)), b1 = __ldg(reinterpret_cast<const float4*>(bias + 32 * f) + 1);
                bv[f][0] = b0.x; bv[f][1] = b0.y; bv[f][2] = b0.z; bv[f][3] = b0.w; bv[f][4] = b1.x; bv[f][5] = b1.y; bv[f][6] = b1.z; bv[f][7] = b1.w;
            }
            mbar_wait(&tmem_full[ab], (i >> 1) & 1);
            tc_fence_after();
            const int row_base = m_tile * TC_BM + q * 32 + tr;
            float* cbase = a.C + static_cast<long long>(head) * a.c_head_stride + 96 * chalf + 8 * tq;
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
                const uint32_t acc = tmem_base + (static_cast<uint32_t>(q * 32 + 16 * hh) << 16) + static_cast<uint32_t>(ab * TC_BN + 96 * chalf);
                float v[3][16];
#pragma unroll
                for (int f = 0; f < 3; ++f) tmem_ld_16x32(acc + static_cast<uint32_t>(32 * f), v[f]);
                tmem_ld_wait();
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int row = row_base + 16 * hh + 8 * rr;
                    if (row < a.M) {
                        float* dst = cbase + static_cast<long long>(row) * TC_BN;
                        const int k0 = 2 * rr;      // fragment register of unit m: k0 + 4(m >> 1) + (m & 1)
#pragma unroll
                        for (int f = 0; f < 3; ++f)
                            stg_v8(dst + 32 * f, v[f][k0] + bv[f][0], v[f][k0 + 1] + bv[f][1], v[f][k0 + 4] + bv[f][2], v[f][k0 + 5] + bv[f][3],
                                   v[f][k0 + 8] + bv[f][4], v[f][k0 + 9] + bv[f][5], v[f][k0 + 12] + bv[f][6], v[f][k0 + 13] + bv[f][7]);
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty[ab]);
        }
}

// PERSISTENT: grid = min(#tiles, #SMs); every CTA walks tiles blockIdx.x, +gridDim.x, ...  The TMA ring runs
// continuously across tiles and the accumulator is double buffered in TMEM, so the epilogue of tile i (TMEM ->
// registers -> 256-bit stores) overlaps the MMAs of tile i+1.  The gate rows of W_ih arrive PERMUTED inside every
// 32-block (crvae_split_tf32_gate_rows), so the 8 accumulator columns a thread's 16x256b fragment holds are 8
// consecutive gate columns: each quad stores one full 128-byte line per row, no shared-memory transpose.
__global__ void __launch_bounds__(TC_THREADS, 1)
proj_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo, ProjTcArgs a) {
    using namespace umma;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);   // offset form keeps the shared address space (LDS/STS, not generic LD/ST)
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + TC_STAGES * TC_STAGE_BYTES);
    uint64_t* empty = full + TC_STAGES;
    uint64_t* tmem_full = empty + TC_STAGES;      // [2]
    uint64_t* tmem_empty = tmem_full + 2;         // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nchunks = (a.K + TC_BK - 1) / TC_BK;

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tmA_hi); prefetch_tmap(&tmA_lo); prefetch_tmap(&tmB_hi); prefetch_tmap(&tmB_lo);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < TC_STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
            for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], TC_EPI_WARPS); }
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<TC_TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int c = 0;                                   // running chunk counter over all tiles of this CTA
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x) {
                const int head = tile / a.n_mtiles, m_tile = tile % a.n_mtiles;
                for (int kc = 0; kc < nchunks; ++kc, ++c) {
                    const int s = c % TC_STAGES, ph = (c / TC_STAGES) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + s * TC_STAGE_BYTES;
                    mbar_arrive_expect_tx(&full[s], TC_STAGE_BYTES);
                    tma_load_2d(st, &tmA_hi, &full[s], kc * TC_BK, m_tile * TC_BM);
                    tma_load_2d(st + TC_A_BYTES, &tmA_lo, &full[s], kc * TC_BK, m_tile * TC_BM);
                    tma_load_2d(st + 2 * TC_A_BYTES, &tmB_hi, &full[s], kc * TC_BK, head * TC_BN);
                    tma_load_2d(st + 2 * TC_A_BYTES + TC_B_BYTES, &tmB_lo, &full[s], kc * TC_BK, head * TC_BN);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_tf32(TC_BM, TC_BN, false, false);
            int c = 0, i = 0;
            for (int tile = blockIdx.x; tile < a.n_tiles; tile += gridDim.x, ++i) {
                const int ab = i & 1;
                mbar_wait(&tmem_empty[ab], ((i >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t acc = tmem_base + static_cast<uint32_t>(ab * TC_BN);
                for (int kc = 0; kc < nchunks; ++kc, ++c) {
                    const int s = c % TC_STAGES, ph = (c / TC_STAGES) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t st = smem_u32(smem + s * TC_STAGE_BYTES);
                    const uint64_t a_hi = smem_desc_k_sw128(st), a_lo = smem_desc_k_sw128(st + TC_A_BYTES);
                    const uint64_t b_hi = smem_desc_k_sw128(st + 2 * TC_A_BYTES), b_lo = smem_desc_k_sw128(st + 2 * TC_A_BYTES + TC_B_BYTES);
                    int ksteps = (a.K - kc * TC_BK + 7) / 8;
                    if (ksteps > TC_BK / 8) ksteps = TC_BK / 8;
                    for (int k = 0; k < ksteps; ++k) {
                        const uint64_t adv = static_cast<uint64_t>(2 * k);        // 8 tf32 = 32 B = 2 x 16 B
                        mma_tf32_ss(acc, a_lo + adv, b_hi + adv, idesc, (kc | k) != 0);
                        mma_tf32_ss(acc, a_hi + adv, b_lo + adv, idesc, true);
                        mma_tf32_ss(acc, a_hi + adv, b_hi + adv, idesc, true);
                    }
                    mma_commit(&empty[s]);          // smem stage is free once these MMAs have read it
                }
                mma_commit(&tmem_full[ab]);         // accumulator complete
            }
        }
    } else {
        proj_epilogue(a, tmem_base, tmem_full, tmem_empty, warp, lane, blockIdx.x, gridDim.x, a.n_tiles);
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TC_TMEM_COLS>(tmem_base);
    }
}


// ---------------------------------------------------------------------------------------------
// RESIDENT-W variant for shallow reductions (K = 32 nfull + tail, nfull <= 3, tail <= 8; K = 100 is 3 x 32 + 4).
// The streaming kernel above re-fetches W_ih hi | lo (154 KB at K = 100) for each of a head's row tiles -- 256 KB
// into the SM per 96 KB out, and it is bound by exactly that ingress.  Here every CTA walks a CONTIGUOUS range of
// tiles (= consecutive row tiles of one head, at most two heads per CTA), keeps the head's W_ih in shared memory
// (full chunks K-major SWIZZLE_128B; the tail as one 8-wide K-step, SWIZZLE_32B) and streams only x: 102 KB per tile.
// ---------------------------------------------------------------------------------------------
constexpr int RS_WCHUNK = 2 * TC_B_BYTES;                 // 49152: W hi | lo of one 32-wide chunk
constexpr int RS_WTAIL = TC_BN * 32;                      // 6144: [192 rows x 8 k]
constexpr int RS_OFF_WTAIL = 3 * RS_WCHUNK;               // 147456
constexpr int RS_OFF_X = RS_OFF_WTAIL + 2 * RS_WTAIL;     // 159744 (1024-aligned)
constexpr int RS_XSLOT = 2 * TC_A_BYTES;                  // 32768: x hi | lo of one chunk (tail: 4096 + 4096 used)
constexpr int RS_XTAIL = TC_BM * 32;                      // 4096
constexpr int RS_SLOTS = 2;
constexpr int RS_SMEM_BYTES = RS_OFF_X + RS_SLOTS * RS_XSLOT + 1024 + 256;

struct ProjResMaps {
    CUtensorMap a_hi, a_lo, b_hi, b_lo;          // 32-wide boxes, SWIZZLE_128B
    CUtensorMap at_hi, at_lo, bt_hi, bt_lo;      // 8-wide boxes, SWIZZLE_32B (tail)
};

__global__ void __launch_bounds__(TC_THREADS, 1)
proj_fwd_tc_res_kernel(const __grid_constant__ ProjResMaps tm, ProjTcArgs a, int nfull, int tail) {
    using namespace umma;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (umma::smem_u32(smem_raw) & 1023u)) & 1023u);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + RS_OFF_X + RS_SLOTS * RS_XSLOT);
    uint64_t* empty = full + RS_SLOTS;
    uint64_t* tmem_full = empty + RS_SLOTS;       // [2]
    uint64_t* tmem_empty = tmem_full + 2;         // [2]
    uint64_t* w_full = tmem_empty + 2;            // W of the current head landed
    uint64_t* w_empty = w_full + 1;               // every MMA that reads the current W has completed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_empty + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nch = nfull + (tail > 0 ? 1 : 0);
    const int t_begin = static_cast<int>((static_cast<long long>(blockIdx.x) * a.n_tiles) / gridDim.x);
    const int t_end = static_cast<int>((static_cast<long long>(blockIdx.x + 1) * a.n_tiles) / gridDim.x);

    if (warp == 0 && lane == 0) {
        prefetch_tmap(&tm.a_hi); prefetch_tmap(&tm.a_lo); prefetch_tmap(&tm.b_hi); prefetch_tmap(&tm.b_lo);
        prefetch_tmap(&tm.at_hi); prefetch_tmap(&tm.at_lo); prefetch_tmap(&tm.bt_hi); prefetch_tmap(&tm.bt_lo);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < RS_SLOTS; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
            for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], TC_EPI_WARPS); }
            mbar_init(w_full, 1); mbar_init(w_empty, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<TC_TMEM_COLS>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int c = 0, cur_head = -1, nsw = 0;
            for (int tile = t_begin; tile < t_end; ++tile) {
                const int head = tile / a.n_mtiles, m_tile = tile % a.n_mtiles;
                if (head != cur_head) {
                    if (nsw > 0) mbar_wait(w_empty, (nsw - 1) & 1);          // the previous head's MMAs are done with W
                    mbar_arrive_expect_tx(w_full, nfull * RS_WCHUNK + (tail > 0 ? 2 * RS_WTAIL : 0));
                    for (int kc = 0; kc < nfull; ++kc) {
                        tma_load_2d(smem + kc * RS_WCHUNK, &tm.b_hi, w_full, kc * 32, head * TC_BN);
                        tma_load_2d(smem + kc * RS_WCHUNK + TC_B_BYTES, &tm.b_lo, w_full, kc * 32, head * TC_BN);
                    }
                    if (tail > 0) {
                        tma_load_2d(smem + RS_OFF_WTAIL, &tm.bt_hi, w_full, nfull * 32, head * TC_BN);
                        tma_load_2d(smem + RS_OFF_WTAIL + RS_WTAIL, &tm.bt_lo, w_full, nfull * 32, head * TC_BN);
                    }
                    cur_head = head; ++nsw;
                }
                for (int kc = 0; kc < nch; ++kc, ++c) {
                    const int s = c % RS_SLOTS, ph = (c / RS_SLOTS) & 1;
                    mbar_wait(&empty[s], ph ^ 1);
                    uint8_t* st = smem + RS_OFF_X + s * RS_XSLOT;
                    if (kc < nfull) {
                        mbar_arrive_expect_tx(&full[s], RS_XSLOT);
                        tma_load_2d(st, &tm.a_hi, &full[s], kc * 32, m_tile * TC_BM);
                        tma_load_2d(st + TC_A_BYTES, &tm.a_lo, &full[s], kc * 32, m_tile * TC_BM);
                    } else {
                        mbar_arrive_expect_tx(&full[s], 2 * RS_XTAIL);
                        tma_load_2d(st, &tm.at_hi, &full[s], nfull * 32, m_tile * TC_BM);
                        tma_load_2d(st + RS_XTAIL, &tm.at_lo, &full[s], nfull * 32, m_tile * TC_BM);
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_tf32(TC_BM, TC_BN, false, false);
            int c = 0, i = 0, cur_head = -1, nsw = 0;
            for (int tile = t_begin; tile < t_end; ++tile, ++i) {
                const int head = tile / a.n_mtiles;
                if (head != cur_head) {
                    mbar_wait(w_full, nsw & 1);
                    cur_head = head; ++nsw;
                }
                const int ab = i & 1;
                mbar_wait(&tmem_empty[ab], ((i >> 1) & 1) ^ 1);       // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t acc = tmem_base + static_cast<uint32_t>(ab * TC_BN);
                for (int kc = 0; kc < nch; ++kc, ++c) {
                    const int s = c % RS_SLOTS, ph = (c / RS_SLOTS) & 1;
                    mbar_wait(&full[s], ph);
                    tc_fence_after();
                    const uint32_t xs = smem_u32(smem + RS_OFF_X + s * RS_XSLOT);
                    if (kc < nfull) {
                        const uint32_t ws = smem_u32(smem + kc * RS_WCHUNK);
                        const uint64_t a_hi = smem_desc_k_sw128(xs), a_lo = smem_desc_k_sw128(xs + TC_A_BYTES);
                        const uint64_t b_hi = smem_desc_k_sw128(ws), b_lo = smem_desc_k_sw128(ws + TC_B_BYTES);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            const uint64_t adv = static_cast<uint64_t>(2 * k);        // 8 tf32 = 32 B = 2 x 16 B
                            mma_tf32_ss(acc, a_lo + adv, b_hi + adv, idesc, (kc | k) != 0);
                            mma_tf32_ss(acc, a_hi + adv, b_lo + adv, idesc, true);
                            mma_tf32_ss(acc, a_hi + adv, b_hi + adv, idesc, true);
                        }
                    } else {        // the tail: one K = 8 step on SWIZZLE_32B tiles (columns beyond K are zero-filled by TMA)
                        const uint32_t ws = smem_u32(smem + RS_OFF_WTAIL);
                        const uint64_t a_hi = smem_desc_k_sw32(xs), a_lo = smem_desc_k_sw32(xs + RS_XTAIL);
                        const uint64_t b_hi = smem_desc_k_sw32(ws), b_lo = smem_desc_k_sw32(ws + RS_WTAIL);
                        mma_tf32_ss(acc, a_lo, b_hi, idesc, kc != 0);
                        mma_tf32_ss(acc, a_hi, b_lo, idesc, true);
                        mma_tf32_ss(acc, a_hi, b_hi, idesc, true);
                    }
                    mma_commit(&empty[s]);          // smem slot is free once these MMAs have read it
                }
                mma_commit(&tmem_full[ab]);         // accumulator complete
                const bool last_of_head = (tile + 1 == t_end) || ((tile + 1) / a.n_mtiles != head);
                if (last_of_head) mma_commit(w_empty);
            }
        }
    } else {
        proj_epilogue(a, tmem_base, tmem_full, tmem_empty, warp, lane, t_begin, 1, t_end);
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TC_TMEM_COLS>(tmem_base);
    }
}

// same split, rows permuted inside every 32-row block: row r of src lands at row pos_of_unit(r) (gru_tc_common.cuh)
__global__ void split_tf32_gate_rows_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, long long rows,
                                            int cols) {
    const long long n = rows * cols;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const long long r = e / cols;
        const int c = static_cast<int>(e - r * cols);
        const long long rd = (r & ~31LL) | pos_of_unit(static_cast<int>(r & 31));
        const float v = src[e];
        const float hf = tf32_rna(v);
        hi[rd * cols + c] = hf;
        lo[rd * cols + c] = __fsub_rn(v, hf);
    }
}

// hi = tf32(v) (round to nearest, ties away), lo = v - hi (exact)
__global__ void split_tf32_kernel(const float* __restrict__ src, float* __restrict__ hi, float* __restrict__ lo, long long n) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        float v = src[e];
        uint32_t h;
        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
        float hf = __uint_as_float(h);
        hi[e] = hf;
        lo[e] = __fsub_rn(v, hf);
    }
}

// ---------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) {
            set_error("cuTensorMapEncodeTiled entry point unavailable");
            return nullptr;
        }
        fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// fp32 matrix [rows][inner] (inner contiguous, row stride = row_stride_elems), box {32 x box_rows}, 128B swizzle
int make_tmap_2d(CUtensorMap* m, const float* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems,
                 uint32_t box_inner, uint32_t box_rows, bool atom32b) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return CRVAE_E_BADARG;
    cuuint64_t gdim[2] = {inner, rows};
    cuuint64_t gstr[1] = {row_stride_elems * sizeof(float)};
    cuuint32_t box[2] = {box_inner, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, atom32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (%d): inner=%llu rows=%llu stride=%llu", (int)r, (unsigned long long)inner,
                  (unsigned long long)rows, (unsigned long long)row_stride_elems);
        return CRVAE_E_BADARG;
    }
    return 0;
}

// 2D map, 32-byte swizzle (box_inner = 8 floats): the narrow tail chunk of the resident-W projection
int make_tmap_2d_sw32(CUtensorMap* m, const float* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems, uint32_t box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return CRVAE_E_BADARG;
    cuuint64_t gdim[2] = {inner, rows};
    cuuint64_t gstr[1] = {row_stride_elems * sizeof(float)};
    cuuint32_t box[2] = {8, box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_32B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (32B swizzle) failed (%d)", (int)r);
        return CRVAE_E_BADARG;
    }
    return 0;
}

// general tiled map (rank <= 5), fp32, 128-byte swizzle; strides_bytes has rank-1 entries (dims 1..rank-1)
int make_tmap_generic(CUtensorMap* m, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, bool atom32b) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return CRVAE_E_BADARG;
    cuuint64_t gdim[5], gstr[4];
    cuuint32_t bx[5], estr[5];
    for (int i = 0; i < rank; ++i) { gdim[i] = dims[i]; bx[i] = box[i]; estr[i] = 1; }
    for (int i = 0; i < rank - 1; ++i) gstr[i] = strides_bytes[i];
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), gdim, gstr, bx, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, atom32b ? CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B : CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled (rank %d) failed (%d)", rank, (int)r);
        return CRVAE_E_BADARG;
    }
    return 0;
}

}  // namespace crvae

using namespace crvae;

extern "C" int crvae_split_tf32(const float* src, float* hi, float* lo, int64_t n, void* stream) {
    CRVAE_REQUIRE(src && hi && lo && n >= 0, "bad argument");
    if (n == 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    split_tf32_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, hi, lo, n);
    return check_launch("split_tf32_kernel");
}

extern "C" int crvae_split_tf32_gate_rows(const float* src, float* hi, float* lo, int64_t rows, int cols, void* stream) {
    CRVAE_REQUIRE(src && hi && lo && rows >= 0 && cols > 0, "bad argument");
    CRVAE_REQUIRE(rows % 32 == 0, "rows must be a multiple of 32 (gate rows: 192 per head)");
    CRVAE_REQUIRE(src != hi && src != lo, "the permuting split cannot run in place");
    if (rows == 0) return 0;
    long long blocks = (rows * cols + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    split_tf32_gate_rows_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(src, hi, lo, rows, cols);
    return check_launch("split_tf32_gate_rows_kernel");
}

extern "C" int crvae_proj_fwd_tc(const float* x_hi, const float* x_lo, const float* w_hi, const float* w_lo,
                                 const float* b_ih, float* gates, int P, int T, int B, int K, int t_skip, void* stream) {
    CRVAE_REQUIRE(x_hi && x_lo && w_hi && w_lo && b_ih && gates, "null operand");
    CRVAE_REQUIRE(P >= 0 && T > 0 && B > 0 && K > 0 && t_skip >= 0 && t_skip <= T, "bad size");
    CRVAE_REQUIRE(K % 4 == 0, "tensor-core projection needs K % 4 == 0 (16-byte TMA row pitch); use crvae_proj_fwd");
    CRVAE_REQUIRE(aligned16(x_hi) && aligned16(x_lo) && aligned16(w_hi) && aligned16(w_lo) && aligned16(b_ih), "16-byte alignment");
    CRVAE_REQUIRE((reinterpret_cast<uintptr_t>(gates) & 31u) == 0, "gates must be 32-byte aligned (256-bit stores)");
    const int M = (T - t_skip) * B;
    if (P == 0 || M == 0) return 0;
    const long long xoff = (long long)t_skip * B * K;
    CUtensorMap tA_hi, tA_lo, tB_hi, tB_lo;
    int rc;
    if ((rc = make_tmap_2d(&tA_hi, x_hi + xoff, K, M, K, TC_BK, TC_BM, false))) return rc;
    if ((rc = make_tmap_2d(&tA_lo, x_lo + xoff, K, M, K, TC_BK, TC_BM, false))) return rc;
    if ((rc = make_tmap_2d(&tB_hi, w_hi, K, (uint64_t)P * TC_BN, K, TC_BK, TC_BN, false))) return rc;
    if ((rc = make_tmap_2d(&tB_lo, w_lo, K, (uint64_t)P * TC_BN, K, TC_BK, TC_BN, false))) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(proj_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM_BYTES);
        if (e != cudaSuccess) { set_error("proj_fwd_tc smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    const int n_mtiles = (M + TC_BM - 1) / TC_BM;
    const int n_tiles = n_mtiles * P;
    ProjTcArgs a{gates + (long long)t_skip * B * TC_BN, b_ih, (long long)T * B * TC_BN, M, K, n_mtiles, n_tiles};
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    // Leave a few SMs free when the grid would fill the machine: the latency-bound encoder chain runs concurrently
    // on a high-priority stream (engine.py) and a persistent grid would otherwise starve it until this kernel ends.
    static int reserve = -1;
    if (reserve < 0) {
        const char* e = getenv("CRVAE_PROJ_RESERVE_SMS");
        reserve = e ? atoi(e) : 16;
    }
    int grid = n_tiles < num_sms ? n_tiles : num_sms;
    if (n_tiles >= 2 * num_sms && grid > reserve + 8) grid -= reserve;
    // shallow reduction + several row tiles per head: keep W_ih resident, stream x only
    static int res_mode = -1;
    if (res_mode < 0) {
        const char* e = getenv("CRVAE_PROJ_RESIDENT");
        res_mode = e ? atoi(e) : 1;
    }
    const int nfull = K / 32, tail = K % 32;
    if (res_mode && nfull >= 1 && nfull <= 3 && tail <= 8 && n_mtiles >= 4 && n_tiles >= 2 * grid) {
        ProjResMaps tm;
        tm.a_hi = tA_hi; tm.a_lo = tA_lo; tm.b_hi = tB_hi; tm.b_lo = tB_lo;
        if (tail > 0) {
            if ((rc = make_tmap_2d_sw32(&tm.at_hi, x_hi + xoff, K, M, K, TC_BM))) return rc;
            if ((rc = make_tmap_2d_sw32(&tm.at_lo, x_lo + xoff, K, M, K, TC_BM))) return rc;
            if ((rc = make_tmap_2d_sw32(&tm.bt_hi, w_hi, K, (uint64_t)P * TC_BN, K, TC_BN))) return rc;
            if ((rc = make_tmap_2d_sw32(&tm.bt_lo, w_lo, K, (uint64_t)P * TC_BN, K, TC_BN))) return rc;
        } else {
            tm.at_hi = tA_hi; tm.at_lo = tA_lo; tm.bt_hi = tB_hi; tm.bt_lo = tB_lo;
        }
        static bool attr2_done = false;
        if (!attr2_done) {
            cudaError_t e = cudaFuncSetAttribute(proj_fwd_tc_res_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RS_SMEM_BYTES);
            if (e != cudaSuccess) { set_error("proj_fwd_tc_res smem attr (%d B): %s", RS_SMEM_BYTES, cudaGetErrorString(e)); return (int)e; }
            attr2_done = true;
        }
        proj_fwd_tc_res_kernel<<<grid, TC_THREADS, RS_SMEM_BYTES, (cudaStream_t)stream>>>(tm, a, nfull, tail);
        return check_launch("proj_fwd_tc_res_kernel");
    }
    proj_fwd_tc_kernel<<<grid, TC_THREADS, TC_SMEM_BYTES, (cudaStream_t)stream>>>(tA_hi, tA_lo, tB_hi, tB_lo, a);
    return check_launch("proj_fwd_tc_kernel");
}
