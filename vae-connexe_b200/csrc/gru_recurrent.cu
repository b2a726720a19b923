// Persistent multi-head GRU recurrence (forward + hand-written BPTT), exact-fp32 FFMA path.
//
// One CTA owns one (head, batch-tile) pair for ALL timesteps: the head's W_hh (48 KB fp32) is
// staged into shared memory once and stays there; the hidden state of the tile lives in shared
// memory between steps; the r/z/n gate math, the per-head Linear(H,1) and (backward) the dW_hh /
// bias / dw_lin accumulations are fused into the step.  Global traffic per step is only the
// streamed gate buffer (read gi / write r,z,n in place), gh_n and h.
//
// Reference arithmetic replaced: ATen's native GRU per-step linear_hh + pointwise cell under
// nn.GRU (CRVAE_lorenz96.py:119, :208), nn.Linear(H,1) (:120) and autograd through them (:497).
#include "common.cuh"

namespace crvae {

constexpr int H = CRVAE_HIDDEN;   // 64
constexpr int G = CRVAE_G;        // 192
constexpr int WT_LD = G + 4;      // padded row of the transposed W_hh in smem
constexpr int D_LD = G + 4;       // padded row of the dgh tile in smem
constexpr int HP_LD = H + 4;

struct GruFwdArgs {
    float* gates; const float* b_ih; const float* w_hh; const float* b_hh;
    const float* h0; long long h0_stride;
    const float* w_lin; const float* b_lin;
    float* hs; float* ghn; float* pred;
    int P, T, B, t_skip;
};

// ------------------------------------------------------------------------------------------
// forward
// thread (ty, tx): rows b = b_tile + ty*RB + i (i < RB), hidden units j = 4*tx + jj (jj < 4),
// for all three gates -> RB x 12 accumulators.
// ------------------------------------------------------------------------------------------
template <int RB>
__global__ void __launch_bounds__(256, 2) gru_fwd_kernel(GruFwdArgs a) {
    constexpr int BT = 16 * RB;
    constexpr int HT_LD = BT + 4;
    extern __shared__ __align__(16) float smem[];
    float* Wt = smem;                 // [H][WT_LD]   Wt[k][g] = W_hh[g][k]
    float* hT = smem + H * WT_LD;     // [H][HT_LD]   hT[k][b] = h[b][k]

    // 16-row tiles (RB == 1, the latency-bound P = 1 shapes): lanes run over the ROWS and a warp holds only two unit
    // groups, so the W_hh operand of the inner product is a two-address broadcast (96 B per k and warp instead of 768 B:
    // with lanes over unit groups every warp re-read the whole matrix each step and the step was shared-memory bound).
    constexpr bool SWAP = (RB == 1);
    const int tid = threadIdx.x, tx = SWAP ? (tid >> 4) : (tid & 15), ty = SWAP ? (tid & 15) : (tid >> 4);
    const int head = blockIdx.y, b_tile = blockIdx.x * BT;
    const int j0 = 4 * tx;
    __shared__ float pred_part[SWAP ? 16 * 17 : 1];

    const float* __restrict__ W = a.w_hh + (long long)head * G * H;
    for (int e = tid; e < G * H; e += 256) {
        int g = e / H, k = e % H;
        Wt[k * WT_LD + g] = __ldg(W + e);
    }
    const float* __restrict__ h0 = a.h0 + (long long)head * a.h0_stride;
    for (int e = tid; e < BT * H; e += 256) {
        int b = e / H, j = e % H;
        int gb = b_tile + b;
        hT[j * HT_LD + b] = gb < a.B ? __ldg(h0 + (long long)gb * H + j) : 0.f;
    }
    float bhh[12], bih[12], wl[4];
#pragma unroll
    for (int q = 0; q < 12; ++q) {
        int g = (q >> 2) * H + j0 + (q & 3);
        bhh[q] = __ldg(a.b_hh + (long long)head * G + g);
        bih[q] = __ldg(a.b_ih + (long long)head * G + g);
    }
    const bool has_lin = a.w_lin != nullptr;
    float blin = 0.f;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) wl[jj] = has_lin ? __ldg(a.w_lin + (long long)head * H + j0 + jj) : 0.f;
    if (has_lin) blin = __ldg(a.b_lin + head);
    __syncthreads();

    for (int t = 0; t < a.T; ++t) {
        const long long row0 = ((long long)head * a.T + t) * a.B;   // row index of b = 0
        // Blackwell issues fp32 FMAs at full rate only as packed pairs (FFMA2, fma.rn.f32x2): accumulators are float2
        // over adjacent gate columns, the hidden value is broadcast into both halves.  Per-component results are
        // bit-identical to scalar fmaf.
        float gi[RB][12];
        auto load_gi = [&]() {
#pragma unroll
            for (int i = 0; i < RB; ++i) {
                int gb = b_tile + ty * RB + i;
                if (t < a.t_skip || gb >= a.B) {
#pragma unroll
                    for (int q = 0; q < 12; ++q) gi[i][q] = bih[q];
                } else {
                    const float* src = a.gates + (row0 + gb) * G + j0;
#pragma unroll
                    for (int gt = 0; gt < 3; ++gt) {
                        float4 v = *reinterpret_cast<const float4*>(src + gt * H);
                        gi[i][gt * 4 + 0] = v.x; gi[i][gt * 4 + 1] = v.y; gi[i][gt * 4 + 2] = v.z; gi[i][gt * 4 + 3] = v.w;
                    }
                }
            }
        };
        // 16-row tiles (RB == 1) are the latency-bound shapes (encoder, P = 1): gi is requested BEFORE the matmul so the
        // global round trip hides behind it.  Larger tiles fetch it after (see below).
        if (RB == 1) load_gi();
        float2 acc2[RB][6];
#pragma unroll
        for (int i = 0; i < RB; ++i)
#pragma unroll
            for (int q = 0; q < 6; ++q) acc2[i][q] = make_float2(0.f, 0.f);
#pragma unroll 4
        for (int k = 0; k < H; ++k) {
            float2 hv2[RB];
#pragma unroll
            for (int i = 0; i < RB; ++i) { const float v = hT[k * HT_LD + ty * RB + i]; hv2[i] = make_float2(v, v); }
            float2 w2[6];
#pragma unroll
            for (int gt = 0; gt < 3; ++gt) {
                float4 v = *reinterpret_cast<const float4*>(&Wt[k * WT_LD + gt * H + j0]);
                w2[gt * 2] = make_float2(v.x, v.y); w2[gt * 2 + 1] = make_float2(v.z, v.w);
            }
#pragma unroll
            for (int i = 0; i < RB; ++i)
#pragma unroll
                for (int q = 0; q < 6; ++q) acc2[i][q] = __ffma2_rn(hv2[i], w2[q], acc2[i][q]);
        }
        float acc[RB][12];
#pragma unroll
        for (int i = 0; i < RB; ++i)
#pragma unroll
            for (int q = 0; q < 6; ++q) { acc[i][2 * q] = acc2[i][q].x; acc[i][2 * q + 1] = acc2[i][q].y; }
        // RB > 1: gi is fetched AFTER the matmul: its latency is covered by the second CTA resident on the SM, and not
        // holding 12*RB registers across the K loop is what lets two CTAs fit (<= 128 registers per thread)
        if (RB != 1) load_gi();
        // gate math; operation order h' = (h - n)*z + n reproduces ATen's CPU GRU
        float hnew[RB][4], rr[RB][4], zz[RB][4], nn[RB][4], gn[RB][4];
#pragma unroll
        for (int i = 0; i < RB; ++i) {
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                float ghr = acc[i][jj] + bhh[jj];
                float ghz = acc[i][4 + jj] + bhh[4 + jj];
                float ghn_ = acc[i][8 + jj] + bhh[8 + jj];
                float r = sigmoidf_fast(gi[i][jj] + ghr);
                float z = sigmoidf_fast(gi[i][4 + jj] + ghz);
                float n = tanhf_fast(__fadd_rn(gi[i][8 + jj], __fmul_rn(r, ghn_)));
                float hold = hT[(j0 + jj) * HT_LD + ty * RB + i];
                hnew[i][jj] = __fadd_rn(__fmul_rn(__fsub_rn(hold, n), z), n);
                rr[i][jj] = r; zz[i][jj] = z; nn[i][jj] = n; gn[i][jj] = ghn_;
            }
        }
        __syncthreads();   // every thread is done reading hT for this step
#pragma unroll
        for (int i = 0; i < RB; ++i) {
            int gb = b_tile + ty * RB + i;
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) hT[(j0 + jj) * HT_LD + ty * RB + i] = hnew[i][jj];
            float ps = 0.f;
            if (has_lin) {
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) ps = fmaf(hnew[i][jj], wl[jj], ps);
                if (SWAP) {
                    pred_part[ty * 17 + tx] = ps;          // the 16 unit groups of a row sit in 8 different warps: sum after the barrier
                } else {
#pragma unroll
                    for (int o = 8; o > 0; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
                }
            }
            if (gb < a.B) {
                float* gdst = a.gates + (row0 + gb) * G + j0;
                *reinterpret_cast<float4*>(gdst) = make_float4(rr[i][0], rr[i][1], rr[i][2], rr[i][3]);
                *reinterpret_cast<float4*>(gdst + H) = make_float4(zz[i][0], zz[i][1], zz[i][2], zz[i][3]);
                *reinterpret_cast<float4*>(gdst + 2 * H) = make_float4(nn[i][0], nn[i][1], nn[i][2], nn[i][3]);
                *reinterpret_cast<float4*>(a.ghn + (row0 + gb) * H + j0) = make_float4(gn[i][0], gn[i][1], gn[i][2], gn[i][3]);
                *reinterpret_cast<float4*>(a.hs + (row0 + gb) * H + j0) =
                    make_float4(hnew[i][0], hnew[i][1], hnew[i][2], hnew[i][3]);
                if (!SWAP && has_lin && tx == 0) a.pred[row0 + gb] = ps + blin;
            }
        }
        __syncthreads();
        if (SWAP && has_lin && tid < 16 && b_tile + tid < a.B) {
            float ps = 0.f;
#pragma unroll
            for (int x = 0; x < 16; ++x) ps += pred_part[tid * 17 + x];
            a.pred[row0 + b_tile + tid] = ps + blin;       // pred_part is rewritten only after the next step's first barrier
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
struct GruBwdArgs {
    float* gates; const float* ghn; const float* hs;
    const float* h0; long long h0_stride;
    const float* w_hh; const float* w_lin;
    const float* dpred; const float* dh_last; const float* dhs;
    float* dh0; float* ws;
    int P, T, B, ntiles;
    float* dghn_out;      // DEFER_DW: receives dgh_n = da_n * r (may alias ghn: each element is read before it is written)
};

constexpr int WS_TILE = G * H + 512;   // floats per (head, tile) partial: dW_hh | db_ih | db_hh | dw_lin | db_lin
constexpr int WS_DBIH = G * H;
constexpr int WS_DBHH = G * H + G;
constexpr int WS_DWLIN = G * H + 2 * G;
constexpr int WS_DBLIN = G * H + 2 * G + H;

// DEFER_DW = true: the dW_hh accumulation is NOT done here (crvae_gru_dwhh_tc does it as one tensor-core GEMM per head);
// the kernel then writes dgh_n for that GEMM, needs neither the 48 accumulator registers nor the h_{t-1} tile, and two
// CTAs fit on an SM.
template <int RB, bool DEFER_DW>
__global__ void __launch_bounds__(256, DEFER_DW ? 2 : 1) gru_bwd_kernel(GruBwdArgs a) {
    constexpr int BT = 16 * RB;
    extern __shared__ __align__(16) float smem[];
    float* Ws = smem;                        // [G][H]      natural layout
    float* Ds = Ws + G * H;                  // [BT][D_LD]  dgh tile, row-major in b
    float* Hp = Ds + BT * D_LD;              // [BT][HP_LD] h_{t-1} tile

    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;      // (the forward's row-lane mapping for 16-row tiles does not pay here: 82 vs 85 us)
    const int head = blockIdx.y, tile = blockIdx.x, b_tile = tile * BT;
    const int j0 = 4 * tx;       // pointwise / matmul-1 column group
    const int g0 = 12 * ty;      // matmul-2 row group (dW_hh rows), columns k = j0..j0+3

    const float* __restrict__ W = a.w_hh + (long long)head * G * H;
    for (int e = tid; e < G * H; e += 256) Ws[e] = __ldg(W + e);
    const bool has_lin = a.w_lin != nullptr;
    float wl[4];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) wl[jj] = has_lin ? __ldg(a.w_lin + (long long)head * H + j0 + jj) : 0.f;

    float dh[RB][4];           // dL/dh_t flowing backwards, (row, j) mapping
#pragma unroll
    for (int i = 0; i < RB; ++i) {
        int gb = b_tile + ty * RB + i;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
            dh[i][jj] = (a.dh_last && gb < a.B) ? __ldg(a.dh_last + ((long long)head * a.B + gb) * H + j0 + jj) : 0.f;
    }
    float dW[12][4];
#pragma unroll
    for (int q = 0; q < 12; ++q)
#pragma unroll
        for (int c = 0; c < 4; ++c) dW[q][c] = 0.f;
    float dbih[12], dbhn[4], dwl[4], dbl = 0.f;
#pragma unroll
    for (int q = 0; q < 12; ++q) dbih[q] = 0.f;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) { dbhn[jj] = 0.f; dwl[jj] = 0.f; }

    const float* __restrict__ h0 = a.h0 + (long long)head * a.h0_stride;

    // dw_lin needs h_t (the OUTPUT of step t); the loop below only ever loads h_{t-1}, so the
    // last output h_{T-1} is folded in here and step t adds dpred[t-1]*h_{t-1}.
    if (has_lin) {
        const long long rowL = ((long long)head * a.T + (a.T - 1)) * a.B;
#pragma unroll
        for (int i = 0; i < RB; ++i) {
            int gb = b_tile + ty * RB + i;
            if (gb < a.B) {
                float dp = __ldg(a.dpred + rowL + gb);
                float4 hv = *reinterpret_cast<const float4*>(a.hs + (rowL + gb) * H + j0);
                dwl[0] = fmaf(dp, hv.x, dwl[0]); dwl[1] = fmaf(dp, hv.y, dwl[1]);
                dwl[2] = fmaf(dp, hv.z, dwl[2]); dwl[3] = fmaf(dp, hv.w, dwl[3]);
            }
        }
    }
    __syncthreads();

    float4 pf[6];                    // RB == 1 only: r | z | n | gh_n | h_{t-2} | dhs of step t-1, requested one step ahead
    float pf_dp = 0.f, pf_dpm1 = 0.f;
#pragma unroll
    for (int q = 0; q < 6; ++q) pf[q] = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int t = a.T - 1; t >= 0; --t) {
        const long long row0 = ((long long)head * a.T + t) * a.B;
        const long long rowP = ((long long)head * a.T + (t - 1)) * a.B;   // rows of h_{t-1} when t > 0
        float dhz[RB][4];
        // ---- pointwise cell backward -> dgi (global, in place), dgh + h_prev (smem) ----
#pragma unroll
        for (int i = 0; i < RB; ++i) {
            const int lb = ty * RB + i, gb = b_tile + lb;
            float4 r4, z4, n4, gn4, hp4, de4 = make_float4(0.f, 0.f, 0.f, 0.f);
            float dp = 0.f, dpm1 = 0.f;
            if (RB == 1 && t != a.T - 1) {      // 16-row tiles: the inputs were requested during the previous step's matmuls
                r4 = pf[0]; z4 = pf[1]; n4 = pf[2]; gn4 = pf[3]; hp4 = pf[4]; de4 = pf[5]; dp = pf_dp; dpm1 = pf_dpm1;
            } else
            if (gb < a.B) {
                const float* gsrc = a.gates + (row0 + gb) * G + j0;
                r4 = *reinterpret_cast<const float4*>(gsrc);
                z4 = *reinterpret_cast<const float4*>(gsrc + H);
                n4 = *reinterpret_cast<const float4*>(gsrc + 2 * H);
                gn4 = *reinterpret_cast<const float4*>(a.ghn + (row0 + gb) * H + j0);
                hp4 = t > 0 ? *reinterpret_cast<const float4*>(a.hs + (rowP + gb) * H + j0)
                            : *reinterpret_cast<const float4*>(h0 + (long long)gb * H + j0);
                if (has_lin) {
                    dp = __ldg(a.dpred + row0 + gb);
                    if (t > 0) dpm1 = __ldg(a.dpred + rowP + gb);
                }
                if (a.dhs) de4 = *reinterpret_cast<const float4*>(a.dhs + (row0 + gb) * H + j0);
            } else {
                r4 = z4 = n4 = gn4 = hp4 = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            const float r_[4] = {r4.x, r4.y, r4.z, r4.w}, z_[4] = {z4.x, z4.y, z4.z, z4.w};
            const float n_[4] = {n4.x, n4.y, n4.z, n4.w}, gn_[4] = {gn4.x, gn4.y, gn4.z, gn4.w};
            const float hp_[4] = {hp4.x, hp4.y, hp4.z, hp4.w}, de_[4] = {de4.x, de4.y, de4.z, de4.w};
            float dar[4], daz[4], dan[4], dghn[4];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                float d = dh[i][jj] + dp * wl[jj] + de_[jj];   // total dL/dh_t
                float dn = d * (1.f - z_[jj]);
                float dz = d * (hp_[jj] - n_[jj]);
                dan[jj] = dn * (1.f - n_[jj] * n_[jj]);
                float dr = dan[jj] * gn_[jj];
                dar[jj] = dr * r_[jj] * (1.f - r_[jj]);
                daz[jj] = dz * z_[jj] * (1.f - z_[jj]);
                dghn[jj] = dan[jj] * r_[jj];
                dhz[i][jj] = d * z_[jj];
                dbih[jj] += dar[jj]; dbih[4 + jj] += daz[jj]; dbih[8 + jj] += dan[jj];
                dbhn[jj] += dghn[jj];
                dwl[jj] = fmaf(dpm1, hp_[jj], dwl[jj]);
            }
            if (tx == 0) dbl += dp;
            if (gb < a.B) {
                float* gdst = a.gates + (row0 + gb) * G + j0;
                *reinterpret_cast<float4*>(gdst) = make_float4(dar[0], dar[1], dar[2], dar[3]);
                *reinterpret_cast<float4*>(gdst + H) = make_float4(daz[0], daz[1], daz[2], daz[3]);
                *reinterpret_cast<float4*>(gdst + 2 * H) = make_float4(dan[0], dan[1], dan[2], dan[3]);
            }
            float* drow = Ds + lb * D_LD + j0;
            *reinterpret_cast<float4*>(drow) = make_float4(dar[0], dar[1], dar[2], dar[3]);
            *reinterpret_cast<float4*>(drow + H) = make_float4(daz[0], daz[1], daz[2], daz[3]);
            *reinterpret_cast<float4*>(drow + 2 * H) = make_float4(dghn[0], dghn[1], dghn[2], dghn[3]);
            if (DEFER_DW) {
                if (gb < a.B) *reinterpret_cast<float4*>(a.dghn_out + (row0 + gb) * H + j0) = make_float4(dghn[0], dghn[1], dghn[2], dghn[3]);
            } else {
                *reinterpret_cast<float4*>(Hp + lb * HP_LD + j0) = hp4;
            }
        }
        __syncthreads();
        if (RB == 1 && t > 0) {          // inputs of step t-1: in flight across both matmuls of this step
            const int gb = b_tile + ty;
            const int tp = t - 1;
            const long long r0p = ((long long)head * a.T + tp) * a.B, rPp = ((long long)head * a.T + (tp - 1)) * a.B;
            pf_dp = 0.f; pf_dpm1 = 0.f;
#pragma unroll
            for (int q = 0; q < 6; ++q) pf[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gb < a.B) {
                const float* gsrc = a.gates + (r0p + gb) * G + j0;
                pf[0] = *reinterpret_cast<const float4*>(gsrc);
                pf[1] = *reinterpret_cast<const float4*>(gsrc + H);
                pf[2] = *reinterpret_cast<const float4*>(gsrc + 2 * H);
                pf[3] = *reinterpret_cast<const float4*>(a.ghn + (r0p + gb) * H + j0);
                pf[4] = tp > 0 ? *reinterpret_cast<const float4*>(a.hs + (rPp + gb) * H + j0)
                               : *reinterpret_cast<const float4*>(h0 + (long long)gb * H + j0);
                if (has_lin) {
                    pf_dp = __ldg(a.dpred + r0p + gb);
                    if (tp > 0) pf_dpm1 = __ldg(a.dpred + rPp + gb);
                }
                if (a.dhs) pf[5] = *reinterpret_cast<const float4*>(a.dhs + (r0p + gb) * H + j0);
            }
        }
        // ---- matmul 1: dh_{t-1}[b][k] = dh_t*z + sum_g dgh[b][g] * W_hh[g][k] ----
        float2 dh2[RB][2];
#pragma unroll
        for (int i = 0; i < RB; ++i) { dh2[i][0] = make_float2(dhz[i][0], dhz[i][1]); dh2[i][1] = make_float2(dhz[i][2], dhz[i][3]); }
#pragma unroll 2
        for (int g = 0; g < G; g += 4) {
            float2 w2[4][2];
#pragma unroll
            for (int gg = 0; gg < 4; ++gg) {
                float4 v = *reinterpret_cast<const float4*>(&Ws[(g + gg) * H + j0]);
                w2[gg][0] = make_float2(v.x, v.y); w2[gg][1] = make_float2(v.z, v.w);
            }
#pragma unroll
            for (int i = 0; i < RB; ++i) {
                float4 d = *reinterpret_cast<const float4*>(&Ds[(ty * RB + i) * D_LD + g]);
                const float dv[4] = {d.x, d.y, d.z, d.w};
#pragma unroll
                for (int gg = 0; gg < 4; ++gg) {
                    const float2 d2 = make_float2(dv[gg], dv[gg]);
                    dh2[i][0] = __ffma2_rn(d2, w2[gg][0], dh2[i][0]);     // packed fp32 FMA (FFMA2)
                    dh2[i][1] = __ffma2_rn(d2, w2[gg][1], dh2[i][1]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < RB; ++i) { dh[i][0] = dh2[i][0].x; dh[i][1] = dh2[i][0].y; dh[i][2] = dh2[i][1].x; dh[i][3] = dh2[i][1].y; }
        // ---- matmul 2: dW_hh[g][k] += sum_b dgh[b][g] * h_{t-1}[b][k]   (g = g0.., k = j0..) ----
        if (!DEFER_DW)
#pragma unroll 2
        for (int b = 0; b < BT; ++b) {
            float4 hp = *reinterpret_cast<const float4*>(&Hp[b * HP_LD + j0]);
            float dv[12];
#pragma unroll
            for (int q = 0; q < 12; q += 4) {
                float4 d = *reinterpret_cast<const float4*>(&Ds[b * D_LD + g0 + q]);
                dv[q] = d.x; dv[q + 1] = d.y; dv[q + 2] = d.z; dv[q + 3] = d.w;
            }
            const float2 hpa = make_float2(hp.x, hp.y), hpb = make_float2(hp.z, hp.w);
#pragma unroll
            for (int q = 0; q < 12; ++q) {
                const float2 d2 = make_float2(dv[q], dv[q]);
                float2 lo = __ffma2_rn(d2, hpa, make_float2(dW[q][0], dW[q][1]));
                float2 hi = __ffma2_rn(d2, hpb, make_float2(dW[q][2], dW[q][3]));
                dW[q][0] = lo.x; dW[q][1] = lo.y; dW[q][2] = hi.x; dW[q][3] = hi.y;
            }
        }
        __syncthreads();
    }

    // ---- outputs ----
#pragma unroll
    for (int i = 0; i < RB; ++i) {
        int gb = b_tile + ty * RB + i;
        if (gb < a.B)
            *reinterpret_cast<float4*>(a.dh0 + ((long long)head * a.B + gb) * H + j0) =
                make_float4(dh[i][0], dh[i][1], dh[i][2], dh[i][3]);
    }
    float* ws = a.ws + ((long long)head * a.ntiles + tile) * WS_TILE;
    if (!DEFER_DW) {
#pragma unroll
        for (int q = 0; q < 12; ++q)
            *reinterpret_cast<float4*>(ws + (g0 + q) * H + j0) = make_float4(dW[q][0], dW[q][1], dW[q][2], dW[q][3]);
    }
    // column sums over the 16 row groups (ty) through shared memory (Ds is free after the last sync)
    float* red = Ds;                       // [16 ty][16 tx][24]
    float* mine = red + (ty * 16 + tx) * 24;
#pragma unroll
    for (int q = 0; q < 12; ++q) mine[q] = dbih[q];
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) { mine[12 + jj] = dbhn[jj]; mine[16 + jj] = dwl[jj]; }
    mine[20] = dbl;
    __syncthreads();
    for (int e = tid; e < 16 * 21; e += 256) {
        int x = e / 21, q = e % 21;
        float s = 0.f;
        for (int y = 0; y < 16; ++y) s += red[(y * 16 + x) * 24 + q];
        if (q < 12) {
            int g = (q >> 2) * H + 4 * x + (q & 3);
            ws[WS_DBIH + g] = s;
            if (q < 8) ws[WS_DBHH + g] = s;     // r and z parts of dgh equal dgi
        } else if (q < 16) {
            ws[WS_DBHH + 2 * H + 4 * x + (q - 12)] = s;
        } else if (q < 20) {
            ws[WS_DWLIN + 4 * x + (q - 16)] = s;
        } else if (x == 0) {
            ws[WS_DBLIN] = s;
        }
    }
}

// Sum the per-tile partials in fixed order (deterministic) into the gradient buffers.
struct GruFinArgs {
    const float* ws; float* dw_hh; float* db_hh; float* db_ih; float* dw_lin; float* db_lin;
    int ntiles;
};
__global__ void gru_bwd_finalize_kernel(GruFinArgs a) {
    const int head = blockIdx.y;
    const float* ws = a.ws + (long long)head * a.ntiles * WS_TILE;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e <= WS_DBLIN; e += gridDim.x * blockDim.x) {
        if (e < WS_DBIH && !a.dw_hh) continue;
        float s = 0.f;
        for (int k = 0; k < a.ntiles; ++k) s += ws[(long long)k * WS_TILE + e];
        if (e < WS_DBIH) { if (a.dw_hh) a.dw_hh[(long long)head * G * H + e] = s; }
        else if (e < WS_DBHH) a.db_ih[(long long)head * G + (e - WS_DBIH)] = s;
        else if (e < WS_DWLIN) a.db_hh[(long long)head * G + (e - WS_DBHH)] = s;
        else if (e < WS_DBLIN) { if (a.dw_lin) a.dw_lin[(long long)head * H + (e - WS_DWLIN)] = s; }
        else if (a.db_lin) a.db_lin[head] = s;
    }
}

static int choose_rb(int P, int B) {
    // largest batch tile that still gives >= 2 CTAs per SM on 148 SMs, else the smallest tile
    for (int rb = 4; rb >= 2; rb >>= 1) {
        long long ctas = (long long)P * ((B + 16 * rb - 1) / (16 * rb));
        if (ctas >= 296) return rb;
    }
    return 1;
}

template <int RB>
static int launch_fwd(const GruFwdArgs& a, cudaStream_t st) {
    constexpr int BT = 16 * RB;
    size_t smem = (size_t)(H * WT_LD + H * (BT + 4)) * sizeof(float);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gru_fwd_kernel<RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("gru_fwd smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    dim3 grid((a.B + BT - 1) / BT, a.P);
    gru_fwd_kernel<RB><<<grid, 256, smem, st>>>(a);
    return check_launch("gru_fwd_kernel");
}

template <int RB, bool DEFER_DW>
static int launch_bwd(const GruBwdArgs& a, cudaStream_t st) {
    constexpr int BT = 16 * RB;
    size_t smem = (size_t)(G * H + BT * D_LD + (DEFER_DW ? 0 : BT * HP_LD)) * sizeof(float);
    size_t red = (size_t)(G * H + 256 * 24) * sizeof(float);
    if (smem < red) smem = red;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gru_bwd_kernel<RB, DEFER_DW>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) { set_error("gru_bwd smem attr: %s", cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    dim3 grid(a.ntiles, a.P);
    gru_bwd_kernel<RB, DEFER_DW><<<grid, 256, smem, st>>>(a);
    return check_launch("gru_bwd_kernel");
}

// used by the tensor-core BPTT (gru_tc_bwd.cu), which writes the same per-tile partial layout (without dW_hh)
int launch_gru_bwd_finalize(const float* ws, float* db_hh, float* db_ih, float* dw_lin, float* db_lin, int P, int ntiles, cudaStream_t st) {
    GruFinArgs f{ws, nullptr, db_hh, db_ih, dw_lin, db_lin, ntiles};
    gru_bwd_finalize_kernel<<<dim3((WS_DBLIN + 256) / 256, P), 256, 0, st>>>(f);
    return check_launch("gru_bwd_finalize_kernel");
}

}  // namespace crvae

using namespace crvae;

static int g_force_rb = 0;   // test/tuning hook: 0 = heuristic
extern "C" void crvae_debug_set_batch_tile(int rows) { g_force_rb = rows / 16; }

extern "C" int crvae_gru_fwd(float* gates, const float* b_ih, const float* w_hh, const float* b_hh,
                             const float* h0, int64_t h0_head_stride, const float* w_lin,
                             const float* b_lin, float* hs, float* ghn, float* pred, int P, int T, int B,
                             int t_skip, void* stream) {
    CRVAE_REQUIRE(gates && b_ih && w_hh && b_hh && h0 && hs && ghn, "null operand");
    CRVAE_REQUIRE((w_lin == nullptr) == (pred == nullptr), "w_lin and pred go together");
    CRVAE_REQUIRE(w_lin == nullptr || b_lin != nullptr, "b_lin missing");
    CRVAE_REQUIRE(P >= 0 && T > 0 && B > 0 && t_skip >= 0 && t_skip <= T, "bad size");
    CRVAE_REQUIRE(aligned16(gates) && aligned16(hs) && aligned16(ghn) && aligned16(h0), "16-byte alignment");
    if (P == 0) return 0;
    GruFwdArgs a{gates, b_ih, w_hh, b_hh, h0, (long long)h0_head_stride, w_lin, b_lin, hs, ghn, pred, P, T, B, t_skip};
    int rb = g_force_rb ? g_force_rb : choose_rb(P, B);
    switch (rb) {
        case 4: return launch_fwd<4>(a, (cudaStream_t)stream);
        case 2: return launch_fwd<2>(a, (cudaStream_t)stream);
        default: return launch_fwd<1>(a, (cudaStream_t)stream);
    }
}

extern "C" size_t crvae_gru_bwd_workspace(int P, int B) {
    return (size_t)(P > 0 ? P : 1) * ((B + 15) / 16) * WS_TILE * sizeof(float);
}

static int gru_bwd_impl(float* gates, float* ghn, const float* hs, const float* h0, int64_t h0_head_stride, const float* w_hh,
                        const float* w_lin, const float* dpred, const float* dh_last, const float* dhs, float* dw_hh,
                        float* db_hh, float* db_ih, float* dw_lin, float* db_lin, float* dh0, int P, int T, int B,
                        void* workspace, void* stream, bool defer_dw) {
    CRVAE_REQUIRE(gates && ghn && hs && h0 && w_hh && db_hh && db_ih && dh0 && workspace, "null operand");
    CRVAE_REQUIRE(defer_dw || dw_hh, "dw_hh missing");
    CRVAE_REQUIRE((w_lin == nullptr) == (dpred == nullptr), "w_lin and dpred go together");
    CRVAE_REQUIRE(w_lin == nullptr || (dw_lin && db_lin), "dw_lin/db_lin missing");
    CRVAE_REQUIRE(P >= 0 && T > 0 && B > 0, "bad size");
    CRVAE_REQUIRE(aligned16(gates) && aligned16(hs) && aligned16(ghn) && aligned16(h0) && aligned16(dh0) &&
                  aligned16(workspace), "16-byte alignment");
    CRVAE_REQUIRE(dhs == nullptr || aligned16(dhs), "16-byte alignment");
    if (P == 0) return 0;
    int rb = g_force_rb ? g_force_rb : choose_rb(P, B);
    int ntiles = (B + 16 * rb - 1) / (16 * rb);
    GruBwdArgs a{gates, ghn, hs, h0, (long long)h0_head_stride, w_hh, w_lin, dpred, dh_last, dhs, dh0,
                 (float*)workspace, P, T, B, ntiles, defer_dw ? ghn : nullptr};
    int rc;
    cudaStream_t st = (cudaStream_t)stream;
    if (defer_dw) {
        switch (rb) {
            case 4: rc = launch_bwd<4, true>(a, st); break;
            case 2: rc = launch_bwd<2, true>(a, st); break;
            default: rc = launch_bwd<1, true>(a, st); break;
        }
    } else {
        switch (rb) {
            case 4: rc = launch_bwd<4, false>(a, st); break;
            case 2: rc = launch_bwd<2, false>(a, st); break;
            default: rc = launch_bwd<1, false>(a, st); break;
        }
    }
    if (rc) return rc;
    GruFinArgs f{(const float*)workspace, defer_dw ? nullptr : dw_hh, db_hh, db_ih, dw_lin, db_lin, ntiles};
    gru_bwd_finalize_kernel<<<dim3((WS_DBLIN + 256) / 256, P), 256, 0, st>>>(f);
    return check_launch("gru_bwd_finalize_kernel");
}

extern "C" int crvae_gru_bwd(float* gates, const float* ghn, const float* hs, const float* h0,
                             int64_t h0_head_stride, const float* w_hh, const float* w_lin,
                             const float* dpred, const float* dh_last, const float* dhs, float* dw_hh,
                             float* db_hh, float* db_ih, float* dw_lin, float* db_lin, float* dh0, int P, int T,
                             int B, void* workspace, void* stream) {
    return gru_bwd_impl(gates, const_cast<float*>(ghn), hs, h0, h0_head_stride, w_hh, w_lin, dpred, dh_last, dhs, dw_hh, db_hh,
                        db_ih, dw_lin, db_lin, dh0, P, T, B, workspace, stream, false);
}

// Same BPTT with the dW_hh accumulation deferred: `ghn` is overwritten IN PLACE with dgh_n = da_n*r and dw_hh is not
// produced here -- follow with crvae_gru_dwhh_tc(gates, ghn, hs, h0, ...).
extern "C" int crvae_gru_bwd_deferred(float* gates, float* ghn, const float* hs, const float* h0,
                                      int64_t h0_head_stride, const float* w_hh, const float* w_lin,
                                      const float* dpred, const float* dh_last, const float* dhs,
                                      float* db_hh, float* db_ih, float* dw_lin, float* db_lin, float* dh0, int P, int T,
                                      int B, void* workspace, void* stream) {
    return gru_bwd_impl(gates, ghn, hs, h0, h0_head_stride, w_hh, w_lin, dpred, dh_last, dhs, nullptr, db_hh, db_ih, dw_lin,
                        db_lin, dh0, P, T, B, workspace, stream, true);
}
