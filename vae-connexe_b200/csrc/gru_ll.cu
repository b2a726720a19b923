// Low-latency multi-head GRU recurrence (forward + BPTT) for SMALL head shards and the replicated encoder.
//
// When a rank holds few heads (p = 100 over 8 GPUs: 12-13 heads; the encoder: one), the recurrent kernels are
// bound by the latency of ONE step, not by bandwidth: 128-row tcgen05 tiles give 26 CTAs on 148 SMs at
// 7 us per step.  Here a CTA is a (head, 16-row) tile -- 16x more CTAs -- and the step's serial chain is
// stripped to  [smem matmul -> gate math -> one barrier]:
//   * W_hh stays in shared memory for the whole sequence (forward: transposed, backward: natural layout);
//   * NOTHING of the step's global traffic goes through the load/store unit: the gate slab of a step
//     (16 rows x 192 floats = 12 KB, contiguous in the [P,T,B,G] layout) arrives by cp.async.bulk into a
//     ring of shared-memory slots two steps ahead, is overwritten IN PLACE with the step's result
//     (forward: r|z|n, backward: dgi) and leaves by cp.async.bulk again; h / gh_n / dgh_n go the same way;
//   * 256 threads = 4 unit groups x 2 HALVES OF THE REDUCTION: a thread accumulates 4 rows x 2 hidden units of all
//     three gates over its half of k -- 12 packed fp32 FMAs (fma.rn.f32x2) per k on operands that are conflict-free
//     64-/128-bit shared-memory reads (8 unit pairs x 4 row groups per warp: every operand word is read once per warp)
//     -- then the two halves swap partial sums through shared memory, each keeping 2 of the 4 rows for the gate
//     math.  (ncu on the first version, one warp per scheduler over the whole k range: 66 % of the step spent in the
//     matmul with the warp waiting on its own shared-memory loads, issue slots 27 % used; two warps per scheduler
//     overlap one warp's load latency with the other's FMAs and halve the gate-math chain.)
//   * exact fp32 (sum over k in two fixed halves: deterministic, within 1 ulp-level rounding of gru_fwd_kernel).
// The BPTT defers dW_hh to crvae_gru_dwhh_tc (one tcgen05 GEMM per head over all steps) like
// crvae_gru_bwd_deferred / crvae_gru_bwd_tc.
//
// Reference arithmetic replaced: nn.GRU per-step linear_hh + cell (CRVAE_lorenz96.py:119, :208, :155, :166),
// nn.Linear(H,1) (:120) and autograd through them (:497).
#include "common.cuh"
#include "umma.cuh"

namespace crvae {

int launch_gru_bwd_finalize(const float* ws, float* db_hh, float* db_ih, float* dw_lin, float* db_lin, int P, int ntiles, cudaStream_t st);

constexpr int LH = CRVAE_HIDDEN;       // 64
constexpr int LG = CRVAE_G;            // 192
constexpr int LL_ROWS = 16;
constexpr int LL_THREADS = 256;
constexpr int LL_HALF = 128;           // threads per reduction half
constexpr int LL_WT_LD = LG + 2;       // transposed W_hh row (forward): 8-byte aligned rows, staging writes conflict-free
constexpr int LL_HT_LD = LL_ROWS + 4;  // h^T / dgh^T row: 16-byte aligned, 4 row groups land in different banks
constexpr int LL_SLAB = LL_ROWS * LG;  // floats of one gate slab
constexpr int LL_HSLAB = LL_ROWS * LH; // floats of one h-sized slab

// workspace layout of one (head, tile) partial -- shared with gru_bwd_finalize_kernel (gru_recurrent.cu)
constexpr int LWS_TILE = LG * LH + 512;
constexpr int LWS_DBIH = LG * LH;
constexpr int LWS_DBHH = LG * LH + LG;
constexpr int LWS_DWLIN = LG * LH + 2 * LG;
constexpr int LWS_DBLIN = LG * LH + 2 * LG + LH;

// ---------------------------------------------------------------- bulk (non-tensor) async copies
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(umma::smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(umma::smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(umma::smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
struct GruLlFwdArgs {
    float* gates; const float* b_ih; const float* w_hh; const float* b_hh;
    const float* h0; long long h0_stride;
    const float* w_lin; const float* b_lin;
    float* hs; float* ghn; float* pred;
    int P, T, B, t_skip;
};

template <int NS>
struct LlFwdSmem {
    float wt[LH * LL_WT_LD];            // wt[k][g] = W_hh[g][k]
    float hT[2][LH * LL_HT_LD];         // hT[buf][k][row]
    float slab[NS][LL_SLAB];            // gi -> r|z|n of a step, row-major [row][192]
    float hst[2][LL_HSLAB];             // h_t rows for the bulk store
    float gst[2][LL_HSLAB];             // gh_n rows for the bulk store
    float predp[2][LL_ROWS * 4];        // per-unit-group partial of the output Linear
    float2 xch[2][6][LL_HALF];          // partial sums handed to the other reduction half: [giver][gate*2 + row][thread]
    uint64_t bar[NS];
};

template <int NS>
__global__ void __launch_bounds__(LL_THREADS, 2) gru_fwd_ll_kernel(GruLlFwdArgs a) {
    using namespace umma;
    extern __shared__ __align__(128) uint8_t ll_smem_raw[];
    LlFwdSmem<NS>& s = *reinterpret_cast<LlFwdSmem<NS>*>(ll_smem_raw);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int uw = warp & 3, kh = warp >> 2;     // unit group, reduction half
    const int th = tid & (LL_HALF - 1);          // index inside the half: the two halves' threads `th` hold the same (rows, units)
    const int rg = lane >> 3, up = lane & 7;
    const int r0 = 4 * rg;                       // matmul rows r0 .. r0+3 (tile-local)
    const int u0 = 16 * uw + 2 * up;             // hidden units u0, u0+1
    const int rm = r0 + 2 * kh;                  // gate-math rows rm, rm+1
    const int head = blockIdx.y, b_tile = blockIdx.x * LL_ROWS;
    const int vrows = min(LL_ROWS, a.B - b_tile);
    const bool has_lin = a.w_lin != nullptr;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) mbar_init(&s.bar[i], 1);
        fence_barrier_init();
    }
    // W_hh -> wt (transposed): lanes run over consecutive gate rows, so the 4 scalar stores of a float4 are conflict-free
    {
        const float* __restrict__ W = a.w_hh + (long long)head * LG * LH;
        for (int e = tid; e < LG * (LH / 4); e += LL_THREADS) {
            const int c = e / LG, g = e - c * LG;
            const float4 v = __ldg(reinterpret_cast<const float4*>(W + (long long)g * LH + 4 * c));
            s.wt[(4 * c + 0) * LL_WT_LD + g] = v.x;
            s.wt[(4 * c + 1) * LL_WT_LD + g] = v.y;
            s.wt[(4 * c + 2) * LL_WT_LD + g] = v.z;
            s.wt[(4 * c + 3) * LL_WT_LD + g] = v.w;
        }
    }
    // h0 -> hT[0] (each thread its two gate-math rows)
    {
        const float* __restrict__ h0 = a.h0 + (long long)head * a.h0_stride;
        float2 v[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int gb = b_tile + rm + i;
            v[i] = gb < a.B ? __ldg(reinterpret_cast<const float2*>(h0 + (long long)gb * LH + u0)) : make_float2(0.f, 0.f);
        }
        *reinterpret_cast<float2*>(&s.hT[0][u0 * LL_HT_LD + rm]) = make_float2(v[0].x, v[1].x);
        *reinterpret_cast<float2*>(&s.hT[0][(u0 + 1) * LL_HT_LD + rm]) = make_float2(v[0].y, v[1].y);
    }
    float bhh[3][2], bih[3][2], wl[2] = {0.f, 0.f};
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            bhh[g][e] = __ldg(a.b_hh + (long long)head * LG + g * LH + u0 + e);
            bih[g][e] = __ldg(a.b_ih + (long long)head * LG + g * LH + u0 + e);
        }
    float blin = 0.f;
    if (has_lin) {
        wl[0] = __ldg(a.w_lin + (long long)head * LH + u0);
        wl[1] = __ldg(a.w_lin + (long long)head * LH + u0 + 1);
        blin = __ldg(a.b_lin + head);
    }
    const long long head_row0 = (long long)head * a.T * a.B;
    const uint32_t slab_bytes = (uint32_t)vrows * LG * 4u, h_bytes = (uint32_t)vrows * LH * 4u;
    __syncthreads();                              // barriers initialised, wt / hT[0] complete
    if (tid == 0) {
        // (slabs of the zero-input steps t < t_skip are fetched too -- their rows exist, are not used, and fetching them
        // keeps the slot / phase arithmetic uniform)
        for (int t = 0; t < NS && t < a.T; ++t) {
            mbar_arrive_expect_tx(&s.bar[t], slab_bytes);
            bulk_g2s(s.slab[t], a.gates + (head_row0 + (long long)t * a.B + b_tile) * LG, slab_bytes, &s.bar[t]);
        }
    }

    for (int t = 0; t < a.T; ++t) {
        const int slot = t % NS, hb = t & 1;
        // ---- partial gh = h . W_hh^T over this half of k, 4 rows x (3 gates x 2 units): packed fp32 FMAs over adjacent units ----
        float2 acc[3][4];
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[g][i] = make_float2(0.f, 0.f);
        const float* __restrict__ hT = s.hT[hb];
        const int k_lo = kh * (LH / 2);
#pragma unroll 8
        for (int kk = 0; kk < LH / 2; ++kk) {
            const int k = k_lo + kk;
            const float4 h4 = *reinterpret_cast<const float4*>(&hT[k * LL_HT_LD + r0]);
            const float2 w_r = *reinterpret_cast<const float2*>(&s.wt[k * LL_WT_LD + u0]);
            const float2 w_z = *reinterpret_cast<const float2*>(&s.wt[k * LL_WT_LD + LH + u0]);
            const float2 w_n = *reinterpret_cast<const float2*>(&s.wt[k * LL_WT_LD + 2 * LH + u0]);
            const float hv[4] = {h4.x, h4.y, h4.z, h4.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 h2 = make_float2(hv[i], hv[i]);
                acc[0][i] = __ffma2_rn(h2, w_r, acc[0][i]);
                acc[1][i] = __ffma2_rn(h2, w_z, acc[1][i]);
                acc[2][i] = __ffma2_rn(h2, w_n, acc[2][i]);
            }
        }
        // hand the two rows the OTHER half does the gate math for to it; keep rows rm, rm+1
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int i = 0; i < 2; ++i) s.xch[kh][g * 2 + i][th] = kh == 0 ? acc[g][2 + i] : acc[g][i];     // (selects, not indexing: registers)
        __syncthreads();
        float2 gh[3][2];                          // full sums of rows rm, rm+1: (k < 32 part) + (k >= 32 part), fixed order
#pragma unroll
        for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float2 other = s.xch[kh ^ 1][g * 2 + i][th], mine = kh == 0 ? acc[g][i] : acc[g][2 + i];
                const float2 lo = kh == 0 ? mine : other, hi = kh == 0 ? other : mine;
                gh[g][i] = make_float2(__fadd_rn(lo.x, hi.x), __fadd_rn(lo.y, hi.y));
            }
        const bool from_slab = t >= a.t_skip;
        mbar_wait(&s.bar[slot], (t / NS) & 1);
        // ---- gate math; operation order h' = (h - n) * z + n reproduces ATen's CPU GRU ----
        float* slab = s.slab[slot];
        float hn[2][2];
        float ps[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float* srow = slab + (rm + i) * LG + u0;
            float2 gi_r, gi_z, gi_n;
            if (from_slab) {
                gi_r = *reinterpret_cast<const float2*>(srow);
                gi_z = *reinterpret_cast<const float2*>(srow + LH);
                gi_n = *reinterpret_cast<const float2*>(srow + 2 * LH);
            } else {                              // zero input step: gi is the bias (the projection did not write these rows)
                gi_r = make_float2(bih[0][0], bih[0][1]); gi_z = make_float2(bih[1][0], bih[1][1]); gi_n = make_float2(bih[2][0], bih[2][1]);
            }
            const float gir[2] = {gi_r.x, gi_r.y}, giz[2] = {gi_z.x, gi_z.y}, gin[2] = {gi_n.x, gi_n.y};
            const float ar[2] = {gh[0][i].x, gh[0][i].y}, az[2] = {gh[1][i].x, gh[1][i].y}, an[2] = {gh[2][i].x, gh[2][i].y};
            float rr[2], zz[2], nn[2], gn[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float ghr = ar[e] + bhh[0][e];
                const float ghz = az[e] + bhh[1][e];
                gn[e] = an[e] + bhh[2][e];
                rr[e] = sigmoidf_fast(gir[e] + ghr);
                zz[e] = sigmoidf_fast(giz[e] + ghz);
                nn[e] = tanhf_fast(__fadd_rn(gin[e], __fmul_rn(rr[e], gn[e])));
                const float hold = hT[(u0 + e) * LL_HT_LD + rm + i];
                hn[i][e] = __fadd_rn(__fmul_rn(__fsub_rn(hold, nn[e]), zz[e]), nn[e]);
            }
            *reinterpret_cast<float2*>(srow) = make_float2(rr[0], rr[1]);
            *reinterpret_cast<float2*>(srow + LH) = make_float2(zz[0], zz[1]);
            *reinterpret_cast<float2*>(srow + 2 * LH) = make_float2(nn[0], nn[1]);
            *reinterpret_cast<float2*>(&s.hst[hb][(rm + i) * LH + u0]) = make_float2(hn[i][0], hn[i][1]);
            *reinterpret_cast<float2*>(&s.gst[hb][(rm + i) * LH + u0]) = make_float2(gn[0], gn[1]);
            ps[i] = fmaf(hn[i][1], wl[1], hn[i][0] * wl[0]);
        }
        float* hTn = s.hT[hb ^ 1];
        *reinterpret_cast<float2*>(&hTn[u0 * LL_HT_LD + rm]) = make_float2(hn[0][0], hn[1][0]);
        *reinterpret_cast<float2*>(&hTn[(u0 + 1) * LL_HT_LD + rm]) = make_float2(hn[0][1], hn[1][1]);
        if (has_lin) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                float p = ps[i];
                p += __shfl_xor_sync(0xffffffffu, p, 1);
                p += __shfl_xor_sync(0xffffffffu, p, 2);
                p += __shfl_xor_sync(0xffffffffu, p, 4);
                if (up == 0) s.predp[hb][(rm + i) * 4 + uw] = p;
            }
        }
        if (tid == 0) bulk_wait_read_all();       // the stores of step t-1 have finished reading their staging buffers
        fence_proxy_async_smem();                 // this thread's shared-memory writes -> visible to the bulk-copy engine
        __syncthreads();
        if (tid == 0) {
            const long long grow = head_row0 + (long long)t * a.B + b_tile;
            bulk_s2g(a.gates + grow * LG, slab, slab_bytes);
            bulk_s2g(a.hs + grow * LH, s.hst[hb], h_bytes);
            bulk_s2g(a.ghn + grow * LH, s.gst[hb], h_bytes);
            bulk_commit();
            // the slot of step t-1 (its store is complete) takes the slab of step t-1+NS
            const int tn = t - 1 + NS;
            if (t >= 1 && tn < a.T) {
                const int sl = (t - 1) % NS;
                mbar_arrive_expect_tx(&s.bar[sl], slab_bytes);
                bulk_g2s(s.slab[sl], a.gates + (head_row0 + (long long)tn * a.B + b_tile) * LG, slab_bytes, &s.bar[sl]);
            }
        }
        if (has_lin && tid >= 32 && tid < 32 + vrows) {      // a warp that does not also issue the bulk copies
            const int r = tid - 32;
            const float4 p4 = *reinterpret_cast<const float4*>(&s.predp[hb][r * 4]);
            a.pred[head_row0 + (long long)t * a.B + b_tile + r] = ((p4.x + p4.y) + (p4.z + p4.w)) + blin;
        }
    }
    if (tid == 0) bulk_wait_all();
}

// ------------------------------------------------------------------------------------------------------------
// backward (dW_hh deferred)
// ------------------------------------------------------------------------------------------------------------
struct GruLlBwdArgs {
    float* gates; float* ghn; const float* hs;
    const float* h0; long long h0_stride;
    const float* w_hh; const float* w_lin;
    const float* dpred; const float* dh_last; const float* dhs;
    float* dh0; float* ws;
    int P, T, B, ntiles;
};

// dynamic shared memory: ws[G*H] | ds[G*LL_HT_LD] | xch[2][2][128] float2 | NS x { slab[16*192] | gn[16*64] | hp[16*64] | (de[16*64]) } | bars
template <int NS, bool HAS_DHS>
__global__ void __launch_bounds__(LL_THREADS, 2) gru_bwd_ll_kernel(GruLlBwdArgs a) {
    using namespace umma;
    extern __shared__ __align__(128) uint8_t ll_smem_raw[];
    constexpr int SLOT = LL_SLAB + (HAS_DHS ? 3 : 2) * LL_HSLAB;
    float* w_s = reinterpret_cast<float*>(ll_smem_raw);           // [G][H] natural layout
    float* d_s = w_s + LG * LH;                                    // [G][LL_HT_LD]: dgh^T of the step
    float2* xch = reinterpret_cast<float2*>(d_s + LG * LL_HT_LD);  // [giver half][row][thread]
    float* ring = reinterpret_cast<float*>(xch + 2 * 2 * LL_HALF);
    uint64_t* bar = reinterpret_cast<uint64_t*>(ring + NS * SLOT);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int uw = warp & 3, kh = warp >> 2;
    const int th = tid & (LL_HALF - 1);
    const int rg = lane >> 3, up = lane & 7;
    const int r0 = 4 * rg, u0 = 16 * uw + 2 * up;
    const int rm = r0 + 2 * kh;                   // this thread's two pointwise rows
    const int head = blockIdx.y, tile = blockIdx.x, b_tile = tile * LL_ROWS;
    const int vrows = min(LL_ROWS, a.B - b_tile);
    const bool has_lin = a.w_lin != nullptr;
    const long long head_row0 = (long long)head * a.T * a.B;
    const float* __restrict__ h0 = a.h0 + (long long)head * a.h0_stride;
    const uint32_t slab_bytes = (uint32_t)vrows * LG * 4u, h_bytes = (uint32_t)vrows * LH * 4u;

    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < NS; ++i) mbar_init(&bar[i], 1);
        fence_barrier_init();
    }
    {
        const float4* __restrict__ W = reinterpret_cast<const float4*>(a.w_hh + (long long)head * LG * LH);
        float4* dst = reinterpret_cast<float4*>(w_s);
        for (int e = tid; e < LG * LH / 4; e += LL_THREADS) dst[e] = __ldg(W + e);
    }
    // rows of a partial tile that are never loaded must not hold garbage (they enter the column sums as exact zeros)
    if (vrows < LL_ROWS)
        for (int e = tid; e < NS * SLOT; e += LL_THREADS) ring[e] = 0.f;
    float wl[2] = {0.f, 0.f};
    if (has_lin) { wl[0] = __ldg(a.w_lin + (long long)head * LH + u0); wl[1] = __ldg(a.w_lin + (long long)head * LH + u0 + 1); }

    float dh[2][2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int gb = b_tile + rm + i;
        float2 v = make_float2(0.f, 0.f);
        if (a.dh_last && gb < a.B) v = __ldg(reinterpret_cast<const float2*>(a.dh_last + ((long long)head * a.B + gb) * LH + u0));
        dh[i][0] = v.x; dh[i][1] = v.y;
    }
    float dbih[3][2], dbhn[2] = {0.f, 0.f}, dwl[2] = {0.f, 0.f}, dbl = 0.f;
#pragma unroll
    for (int g = 0; g < 3; ++g) dbih[g][0] = dbih[g][1] = 0.f;
    // dw_lin needs h_t (the OUTPUT of step t); the loop only ever sees h_{t-1}: fold in h_{T-1} here, step t adds dpred[t-1]*h_{t-1}
    float dp_cur[2], dp_prev[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int gb = b_tile + rm + i;
        dp_cur[i] = 0.f; dp_prev[i] = 0.f;
        if (has_lin && gb < a.B) {
            const long long grow = head_row0 + (long long)(a.T - 1) * a.B + gb;
            dp_cur[i] = __ldg(a.dpred + grow);
            if (a.T > 1) dp_prev[i] = __ldg(a.dpred + grow - a.B);
            const float2 hv = __ldg(reinterpret_cast<const float2*>(a.hs + grow * LH + u0));
            dwl[0] = fmaf(dp_cur[i], hv.x, dwl[0]); dwl[1] = fmaf(dp_cur[i], hv.y, dwl[1]);
        }
    }
    fence_proxy_async_smem();                     // the zero fill above precedes bulk writes into the same slots
    __syncthreads();

    auto issue_loads = [&](int t) {               // thread 0: everything step t reads, into slot (T-1-t) % NS
        const int step = a.T - 1 - t, sl = step % NS;
        float* slot = ring + sl * SLOT;
        const long long grow = head_row0 + (long long)t * a.B + b_tile;
        mbar_arrive_expect_tx(&bar[sl], slab_bytes + (HAS_DHS ? 3u : 2u) * h_bytes);
        bulk_g2s(slot, a.gates + grow * LG, slab_bytes, &bar[sl]);
        bulk_g2s(slot + LL_SLAB, a.ghn + grow * LH, h_bytes, &bar[sl]);
        if (t > 0) bulk_g2s(slot + LL_SLAB + LL_HSLAB, a.hs + (grow - a.B) * LH, h_bytes, &bar[sl]);
        else       bulk_g2s(slot + LL_SLAB + LL_HSLAB, h0 + (long long)b_tile * LH, h_bytes, &bar[sl]);
        if (HAS_DHS) bulk_g2s(slot + LL_SLAB + 2 * LL_HSLAB, a.dhs + grow * LH, h_bytes, &bar[sl]);
    };
    if (tid == 0)
        for (int st = 0; st < NS && st < a.T; ++st) issue_loads(a.T - 1 - st);

    for (int t = a.T - 1, step = 0; t >= 0; --t, ++step) {
        const int sl = step % NS;
        float* slot = ring + sl * SLOT;
        float* gn_s = slot + LL_SLAB;
        const float* hp_s = gn_s + LL_HSLAB;
        const float* de_s = hp_s + LL_HSLAB;
        mbar_wait(&bar[sl], (step / NS) & 1);
        // ---- pointwise cell backward (rows rm, rm+1): dgi (slab, in place), dgh_n (in place), dgh^T (d_s), dh*z ----
        float dhz[2][2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            float* srow = slot + (rm + i) * LG + u0;
            const float2 r2 = *reinterpret_cast<const float2*>(srow), z2 = *reinterpret_cast<const float2*>(srow + LH),
                         n2 = *reinterpret_cast<const float2*>(srow + 2 * LH);
            const float2 g2 = *reinterpret_cast<const float2*>(gn_s + (rm + i) * LH + u0);
            const float2 p2 = *reinterpret_cast<const float2*>(hp_s + (rm + i) * LH + u0);
            float2 e2 = make_float2(0.f, 0.f);
            if (HAS_DHS) e2 = *reinterpret_cast<const float2*>(de_s + (rm + i) * LH + u0);
            const float r_[2] = {r2.x, r2.y}, z_[2] = {z2.x, z2.y}, n_[2] = {n2.x, n2.y}, gn_[2] = {g2.x, g2.y}, hp_[2] = {p2.x, p2.y},
                        de_[2] = {e2.x, e2.y};
            const float dp = dp_cur[i], dpm1 = dp_prev[i];
            float dar[2], daz[2], dan[2], dgn[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const float d = dh[i][e] + dp * wl[e] + de_[e];      // total dL/dh_t (same order as gru_bwd_kernel)
                const float dn = d * (1.f - z_[e]);
                const float dz = d * (hp_[e] - n_[e]);
                dan[e] = dn * (1.f - n_[e] * n_[e]);
                const float dr = dan[e] * gn_[e];
                dar[e] = dr * r_[e] * (1.f - r_[e]);
                daz[e] = dz * z_[e] * (1.f - z_[e]);
                dgn[e] = dan[e] * r_[e];
                dhz[i][e] = d * z_[e];
                dbih[0][e] += dar[e]; dbih[1][e] += daz[e]; dbih[2][e] += dan[e];
                dbhn[e] += dgn[e];
                dwl[e] = fmaf(dpm1, hp_[e], dwl[e]);
            }
            if (up == 0 && uw == 0) dbl += dp;
            *reinterpret_cast<float2*>(srow) = make_float2(dar[0], dar[1]);
            *reinterpret_cast<float2*>(srow + LH) = make_float2(daz[0], daz[1]);
            *reinterpret_cast<float2*>(srow + 2 * LH) = make_float2(dan[0], dan[1]);
            *reinterpret_cast<float2*>(gn_s + (rm + i) * LH + u0) = make_float2(dgn[0], dgn[1]);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                d_s[(u0 + e) * LL_HT_LD + rm + i] = dar[e];
                d_s[(LH + u0 + e) * LL_HT_LD + rm + i] = daz[e];
                d_s[(2 * LH + u0 + e) * LL_HT_LD + rm + i] = dgn[e];
            }
        }
        // dpred of the next step (t-1) and of the one before it
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int gb = b_tile + rm + i;
            dp_cur[i] = dp_prev[i];
            dp_prev[i] = (has_lin && t >= 2 && gb < a.B) ? __ldg(a.dpred + head_row0 + (long long)(t - 2) * a.B + gb) : 0.f;
        }
        if (tid == 0) bulk_wait_read_all();       // stores of the previous step are done with their slot
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            const long long grow = head_row0 + (long long)t * a.B + b_tile;
            bulk_s2g(a.gates + grow * LG, slot, slab_bytes);
            bulk_s2g(a.ghn + grow * LH, gn_s, h_bytes);
            bulk_commit();
            // the slot of the previous step (stores complete) takes the inputs of step (this step + NS - 1)
            const int tn = t + 1 - NS;
            if (step >= 1 && tn >= 0) issue_loads(tn);
        }
        // ---- partial of  sum_g dgh[row][g] * W_hh[g][u]  over this half of the gate rows, 4 rows x 2 units ----
        float2 acc[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[i] = make_float2(0.f, 0.f);
        const int g_lo = kh * (LG / 2);
#pragma unroll 8
        for (int gg = 0; gg < LG / 2; ++gg) {
            const int g = g_lo + gg;
            const float4 d4 = *reinterpret_cast<const float4*>(&d_s[g * LL_HT_LD + r0]);
            const float2 w2 = *reinterpret_cast<const float2*>(&w_s[g * LH + u0]);
            acc[0] = __ffma2_rn(make_float2(d4.x, d4.x), w2, acc[0]);
            acc[1] = __ffma2_rn(make_float2(d4.y, d4.y), w2, acc[1]);
            acc[2] = __ffma2_rn(make_float2(d4.z, d4.z), w2, acc[2]);
            acc[3] = __ffma2_rn(make_float2(d4.w, d4.w), w2, acc[3]);
        }
        xch[(kh * 2 + 0) * LL_HALF + th] = kh == 0 ? acc[2] : acc[0];      // (selects, not indexing: registers)
        xch[(kh * 2 + 1) * LL_HALF + th] = kh == 0 ? acc[3] : acc[1];
        __syncthreads();                          // partials visible; d_s free for the next step's pointwise phase
        // dh_{t-1} = dh_t * z + (low-half partial + high-half partial), fixed order
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float2 other = xch[((kh ^ 1) * 2 + i) * LL_HALF + th], mine = kh == 0 ? acc[i] : acc[2 + i];
            const float2 lo = kh == 0 ? mine : other, hi = kh == 0 ? other : mine;
            dh[i][0] = __fadd_rn(dhz[i][0], __fadd_rn(lo.x, hi.x));
            dh[i][1] = __fadd_rn(dhz[i][1], __fadd_rn(lo.y, hi.y));
        }
    }
    if (tid == 0) bulk_wait_all();

    // ---- outputs: dh0, per-tile partial column sums (fixed order: deterministic) ----
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int gb = b_tile + rm + i;
        if (gb < a.B) *reinterpret_cast<float2*>(a.dh0 + ((long long)head * a.B + gb) * LH + u0) = make_float2(dh[i][0], dh[i][1]);
    }
    float v[11] = {dbih[0][0], dbih[0][1], dbih[1][0], dbih[1][1], dbih[2][0], dbih[2][1], dbhn[0], dbhn[1], dwl[0], dwl[1], dbl};
#pragma unroll
    for (int q = 0; q < 11; ++q) {
        v[q] += __shfl_xor_sync(0xffffffffu, v[q], 8);
        v[q] += __shfl_xor_sync(0xffffffffu, v[q], 16);
    }
    // the two halves hold the sums of different rows: the high half hands its sums over through d_s (free now)
    __syncthreads();
    if (kh == 1 && rg == 0) {
#pragma unroll
        for (int q = 0; q < 11; ++q) d_s[(uw * 8 + up) * 12 + q] = v[q];
    }
    __syncthreads();
    if (kh == 0 && rg == 0) {
#pragma unroll
        for (int q = 0; q < 11; ++q) v[q] += d_s[(uw * 8 + up) * 12 + q];
        float* ws = a.ws + ((long long)head * a.ntiles + tile) * LWS_TILE;
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const int u = u0 + e;
            ws[LWS_DBIH + u] = v[e];            ws[LWS_DBHH + u] = v[e];                  // r: dgh == dgi
            ws[LWS_DBIH + LH + u] = v[2 + e];   ws[LWS_DBHH + LH + u] = v[2 + e];         // z
            ws[LWS_DBIH + 2 * LH + u] = v[4 + e];                                         // n: db_ih
            ws[LWS_DBHH + 2 * LH + u] = v[6 + e];                                         // n: db_hh (dgh_n)
            ws[LWS_DWLIN + u] = v[8 + e];
        }
        if (uw == 0 && up == 0) ws[LWS_DBLIN] = v[10];
    }
}

static int ll_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

template <int NS>
static int launch_fwd_ll(const GruLlFwdArgs& a, cudaStream_t st) {
    const int smem = (int)sizeof(LlFwdSmem<NS>);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gru_fwd_ll_kernel<NS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_error("gru_fwd_ll smem attr (%d B): %s", smem, cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    dim3 grid((a.B + LL_ROWS - 1) / LL_ROWS, a.P);
    gru_fwd_ll_kernel<NS><<<grid, LL_THREADS, smem, st>>>(a);
    return check_launch("gru_fwd_ll_kernel");
}

template <int NS, bool HAS_DHS>
static int launch_bwd_ll(const GruLlBwdArgs& a, cudaStream_t st) {
    const int smem = (LG * LH + LG * LL_HT_LD + NS * (LL_SLAB + (HAS_DHS ? 3 : 2) * LL_HSLAB)) * 4 + 2 * 2 * LL_HALF * 8 + NS * 8 + 64;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gru_bwd_ll_kernel<NS, HAS_DHS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_error("gru_bwd_ll smem attr (%d B): %s", smem, cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    gru_bwd_ll_kernel<NS, HAS_DHS><<<dim3(a.ntiles, a.P), LL_THREADS, smem, st>>>(a);
    return check_launch("gru_bwd_ll_kernel");
}

}  // namespace crvae

using namespace crvae;

// Low-latency form of crvae_gru_fwd: same arguments, bit-identical results; every row pointer must be 16-byte aligned
// (bulk copies), which holds for B-row tensors of 64 / 192 floats per row.
extern "C" int crvae_gru_fwd_ll(float* gates, const float* b_ih, const float* w_hh, const float* b_hh,
                                const float* h0, int64_t h0_head_stride, const float* w_lin, const float* b_lin,
                                float* hs, float* ghn, float* pred, int P, int T, int B, int t_skip, void* stream) {
    CRVAE_REQUIRE(gates && b_ih && w_hh && b_hh && h0 && hs && ghn, "null operand");
    CRVAE_REQUIRE((w_lin == nullptr) == (pred == nullptr), "w_lin and pred go together");
    CRVAE_REQUIRE(w_lin == nullptr || b_lin != nullptr, "b_lin missing");
    CRVAE_REQUIRE(P >= 0 && T > 0 && B > 0 && t_skip >= 0 && t_skip <= T, "bad size");
    CRVAE_REQUIRE(aligned16(gates) && aligned16(hs) && aligned16(ghn) && aligned16(h0) && aligned16(w_hh), "16-byte alignment");
    CRVAE_REQUIRE(h0_head_stride % 4 == 0, "h0 head stride must keep 16-byte alignment");
    if (P == 0) return 0;
    GruLlFwdArgs a{gates, b_ih, w_hh, b_hh, h0, (long long)h0_head_stride, w_lin, b_lin, hs, ghn, pred, P, T, B, t_skip};
    const long long ctas = (long long)P * ((B + LL_ROWS - 1) / LL_ROWS);
    // one CTA per SM: a deeper slab ring; more CTAs than SMs: the 2-slot ring (113 KB) lets two CTAs share an SM
    return ctas <= ll_num_sms() ? launch_fwd_ll<4>(a, (cudaStream_t)stream) : launch_fwd_ll<2>(a, (cudaStream_t)stream);
}

// Low-latency form of crvae_gru_bwd_deferred: same arguments and results (gates <- dgi, ghn <- dgh_n in place, dw_hh
// is produced afterwards by crvae_gru_dwhh_tc).
extern "C" int crvae_gru_bwd_ll(float* gates, float* ghn, const float* hs, const float* h0, int64_t h0_head_stride,
                                const float* w_hh, const float* w_lin, const float* dpred, const float* dh_last,
                                const float* dhs, float* db_hh, float* db_ih, float* dw_lin, float* db_lin, float* dh0,
                                int P, int T, int B, void* workspace, void* stream) {
    CRVAE_REQUIRE(gates && ghn && hs && h0 && w_hh && db_hh && db_ih && dh0 && workspace, "null operand");
    CRVAE_REQUIRE((w_lin == nullptr) == (dpred == nullptr), "w_lin and dpred go together");
    CRVAE_REQUIRE(w_lin == nullptr || (dw_lin && db_lin), "dw_lin/db_lin missing");
    CRVAE_REQUIRE(P >= 0 && T > 0 && B > 0, "bad size");
    CRVAE_REQUIRE(aligned16(gates) && aligned16(hs) && aligned16(ghn) && aligned16(h0) && aligned16(dh0) && aligned16(workspace) &&
                  aligned16(w_hh), "16-byte alignment");
    CRVAE_REQUIRE(dhs == nullptr || aligned16(dhs), "16-byte alignment");
    CRVAE_REQUIRE(h0_head_stride % 4 == 0, "h0 head stride must keep 16-byte alignment");
    if (P == 0) return 0;
    const int ntiles = (B + LL_ROWS - 1) / LL_ROWS;
    GruLlBwdArgs a{gates, ghn, hs, h0, (long long)h0_head_stride, w_hh, w_lin, dpred, dh_last, dhs, dh0, (float*)workspace, P, T, B, ntiles};
    cudaStream_t st = (cudaStream_t)stream;
    const bool roomy = (long long)P * ntiles <= ll_num_sms();
    int rc;
    if (dhs) rc = launch_bwd_ll<3, true>(a, st);
    else     rc = roomy ? launch_bwd_ll<3, false>(a, st) : launch_bwd_ll<2, false>(a, st);
    if (rc) return rc;
    return launch_gru_bwd_finalize((const float*)workspace, db_hh, db_ih, dw_lin, db_lin, P, ntiles, st);
}
