// Persistent multi-head GRU BPTT on the 5th-generation tensor cores (tcgen05), dW_hh deferred.
//
// One CTA owns one (head, 128-row batch tile) pair for ALL timesteps, walking t = T-1 .. 0:
//   * W_hh of the head (tf32 hi | lo, 2 x 48 KB) is staged ONCE into shared memory as the B operand
//     [N = 64 hidden units x K = 192 gate rows] (K-major, SWIZZLE_128B, six 32-wide K blocks);
//   * the gate-gradient tile dgh = [da_r | da_z | da_n * r] of the step is the A operand and lives in TENSOR
//     MEMORY (tf32 hi | lo, 2 x 192 columns), written by the pointwise threads with tcgen05.st;
//   * the accumulator columns are pre-loaded with dh_t * z (tcgen05.st), then ONE thread issues
//     dh_{t-1}[128 x 64] = dh_t * z + dgh[128 x 192] . W_hh  as 24 K-steps x 3 tcgen05.mma (3xTF32);
//   * r | z | n of step t-1 arrive by TMA (six 128-byte-wide boxes, SWIZZLE_128B) while the tensor core works on
//     step t; gh_n and h_{t-1} are prefetched into registers with 256-bit loads; all stores are 256-bit;
//   * 8 warps (lane quadrant x 32-column group) do the cell backward on 16x256b TMEM fragments; a thread owns
//     4 batch rows x 8 hidden units for the whole sequence, so the column sums (db_ih, db_hh, dw_lin, db_lin) are
//     plain register accumulators, reduced once at the end in a fixed order (deterministic, no atomics).
// Same buffers / semantics as crvae_gru_bwd_deferred (gru_recurrent.cu): gates are overwritten with dgi,
// ghn with dgh_n, and crvae_gru_dwhh_tc computes dW_hh afterwards.
//
// Reference arithmetic replaced: autograd through nn.GRU + nn.Linear(H,1) (CRVAE_lorenz96.py:497).
#include "gru_tc_common.cuh"

namespace crvae {

int make_tmap_2d(CUtensorMap* m, const float* base, uint64_t inner, uint64_t rows, uint64_t row_stride_elems,
                 uint32_t box_inner, uint32_t box_rows, bool atom32b);

constexpr int BH = CRVAE_HIDDEN;                 // 64
constexpr int BG = CRVAE_G;                      // 192
constexpr int BT_ROWS = 128;                     // rows per tile (UMMA M)
constexpr int BT_WBLK = BH * 128;                // 8192 B: [64 units x 32 k] fp32, one K block
constexpr int BT_W_BYTES = 6 * BT_WBLK;          // 49152 B per (hi | lo)
constexpr int BT_GBOX = BT_ROWS * 128;           // 16384 B: [128 rows x 32 gate columns]
constexpr int BT_OFF_WHI = 0;
constexpr int BT_OFF_WLO = BT_W_BYTES;
constexpr int BT_OFF_GATE = 2 * BT_W_BYTES;      // six boxes: r | z | n, two 32-column halves each (also the final reduction scratch)
constexpr int BT_OFF_CONST = BT_OFF_GATE + 6 * BT_GBOX;      // w_lin[64]
constexpr int BT_OFF_BAR = BT_OFF_CONST + BH * 4;
constexpr int BT_SMEM_BYTES = BT_OFF_BAR + 64 + 1024;
constexpr int BT_TMEM_COLS = 512;
constexpr int BT_ACOL_HI = 0, BT_ACOL_LO = BG, BT_DCOL = 2 * BG;   // TMEM columns: dgh_hi[192] | dgh_lo[192] | dh[64]
constexpr int BT_THREADS = 256;
constexpr int BT_NACC = 41;                      // per-thread sums: da_r, da_z, da_n, dgh_n, dw_lin (8 units each) + db_lin

// workspace layout of one (head, tile) partial -- shared with gru_bwd_finalize_kernel (gru_recurrent.cu)
constexpr int BWS_TILE = BG * BH + 512;
constexpr int BWS_DBIH = BG * BH;
constexpr int BWS_DBHH = BG * BH + BG;
constexpr int BWS_DWLIN = BG * BH + 2 * BG;
constexpr int BWS_DBLIN = BG * BH + 2 * BG + BH;

struct GruBwdTcArgs {
    float* gates; float* ghn; const float* hs;
    const float* h0; long long h0_stride;
    const float* w_hh; const float* w_lin;
    const float* dpred; const float* dh_last;
    float* dh0; float* ws;
    int P, T, B, ntiles;
};

__device__ __forceinline__ void ldg_v8(const float* p, float* v) {
    asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
                 : "l"(p));
}

__global__ void __launch_bounds__(BT_THREADS, 1) gru_bwd_tc_kernel(const __grid_constant__ CUtensorMap tmG, GruBwdTcArgs a) {
    using namespace umma;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    uint8_t* gate_s = smem + BT_OFF_GATE;
    float* wl_s = reinterpret_cast<float*>(smem + BT_OFF_CONST);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem + BT_OFF_BAR);   // accumulator complete (tcgen05.commit)
    uint64_t* gbar = mbar + 1;                                         // gate tile landed (TMA)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gbar + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int head = blockIdx.y, tile = blockIdx.x;
    const int b_base = tile * BT_ROWS;
    const int q = warp & 3, cg = warp >> 2;       // TMEM lane quadrant of this warp, 32-column group
    const int tr = lane >> 2, tq = lane & 3;
    const bool has_lin = a.w_lin != nullptr;

    for (int e = threadIdx.x; e < BH; e += BT_THREADS) wl_s[e] = has_lin ? __ldg(a.w_lin + (long long)head * BH + e) : 0.f;
    if (warp == 0) {
        if (lane == 0) {
            prefetch_tmap(&tmG);
            mbar_init(mbar, 1);
            mbar_init(gbar, 1);
            fence_barrier_init();
        }
        __syncwarp();
        tmem_alloc<BT_TMEM_COLS>(tmem_slot);
    }
    // W_hh -> B operand: row = position of the hidden unit, K index = gate * 64 + position of the gate row's unit
    {
        const float4* src = reinterpret_cast<const float4*>(a.w_hh + (long long)head * BG * BH);
        for (int idx = threadIdx.x; idx < BG * (BH / 4); idx += BT_THREADS) {
            const int r = idx >> 4, c4 = idx & 15;
            const int kk = (r / BH) * BH + pos_of_unit(r % BH);
            const uint32_t kboff = (kk >> 5) * BT_WBLK, k0 = kk & 31;
            const float4 v = __ldg(src + idx);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = pos_of_unit(4 * c4 + j);
                const uint32_t off = kboff + n * 128 + ((((k0 >> 2) ^ (n & 7)) << 4) | ((k0 & 3) << 2));
                const float hi = tf32_rna(vv[j]);
                *reinterpret_cast<float*>(smem + BT_OFF_WHI + off) = hi;
                *reinterpret_cast<float*>(smem + BT_OFF_WLO + off) = __fsub_rn(vv[j], hi);
            }
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const long long head_row0 = (long long)head * a.T * a.B;
    auto issue_gate_tma = [&](int t) {            // thread 0
        mbar_arrive_expect_tx(gbar, 6 * BT_GBOX);
        const int row = static_cast<int>(head_row0 + (long long)t * a.B + b_base);
#pragma unroll
        for (int c = 0; c < 6; ++c) tma_load_2d(gate_s + c * BT_GBOX, &tmG, gbar, 32 * c, row);
    };
    if (threadIdx.x == 0) issue_gate_tma(a.T - 1);

    // this thread's elements: batch rows brow(hh, rr) = b_base + 32q + 16hh + tr + 8rr, hidden units ucol + m (m < 8);
    // fragment register of (rr, m) inside half hh: 4(m >> 1) + 2rr + (m & 1)
    const int ucol = 32 * cg + 8 * tq;
    const int lrow0 = 32 * q + tr;                 // tile-local row of (hh = 0, rr = 0)
    const uint32_t lane_addr = tmem_base + (static_cast<uint32_t>(32 * q) << 16) + static_cast<uint32_t>(32 * cg);
    const float* h0 = a.h0 + (long long)head * a.h0_stride + ucol;

    float acc[BT_NACC];
#pragma unroll
    for (int i = 0; i < BT_NACC; ++i) acc[i] = 0.f;
    float wl[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) wl[m] = wl_s[ucol + m];

    float pf_ghn[4][8], pf_hp[4][8], dp_cur[4], dp_prev[4];     // slot s = 2hh + rr
    auto prefetch_slot = [&](int t, int s) {       // gh_n[t], h_{t-1}, dpred[t-1] of row slot s
        const int b = b_base + lrow0 + 16 * (s >> 1) + 8 * (s & 1);
        if (b < a.B) {
            const long long grow = head_row0 + (long long)t * a.B + b;
            ldg_v8_stream(a.ghn + grow * BH + ucol, pf_ghn[s]);
            if (t > 0) ldg_v8(a.hs + (grow - a.B) * BH + ucol, pf_hp[s]);
            else       ldg_v8(h0 + (long long)b * BH, pf_hp[s]);
            dp_prev[s] = (has_lin && t > 0) ? __ldg(a.dpred + grow - a.B) : 0.f;
        } else {
#pragma unroll
            for (int m = 0; m < 8; ++m) { pf_ghn[s][m] = 0.f; pf_hp[s][m] = 0.f; }
            dp_prev[s] = 0.f;
        }
    };
    // dw_lin needs h_t (the OUTPUT of step t) while the loop only ever loads h_{t-1}: fold in h_{T-1} here
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const int b = b_base + lrow0 + 16 * (s >> 1) + 8 * (s & 1);
        dp_cur[s] = 0.f;
        if (has_lin && b < a.B) {
            const long long grow = head_row0 + (long long)(a.T - 1) * a.B + b;
            dp_cur[s] = __ldg(a.dpred + grow);
            float hv[8];
            ldg_v8(a.hs + grow * BH + ucol, hv);
#pragma unroll
            for (int m = 0; m < 8; ++m) acc[32 + m] = fmaf(dp_cur[s], hv[m], acc[32 + m]);
        }
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) prefetch_slot(a.T - 1, s);

    auto issue_step = [&]() {          // thread 0: the 72 MMAs of one step (accumulate onto the pre-loaded dh*z)
        tc_fence_after();
        constexpr uint32_t idesc = idesc_tf32(BT_ROWS, BH, false, false);
        const uint32_t w_hi_s = smem_u32(smem + BT_OFF_WHI), w_lo_s = smem_u32(smem + BT_OFF_WLO);
        const uint32_t dacc = tmem_base + BT_DCOL;
#pragma unroll 4
        for (int kk = 0; kk < 24; ++kk) {
            const uint32_t offB = (kk >> 2) * BT_WBLK + (kk & 3) * 32;
            const uint32_t a_hi = tmem_base + BT_ACOL_HI + kk * 8, a_lo = tmem_base + BT_ACOL_LO + kk * 8;
            mma_tf32_ts(dacc, a_lo, smem_desc_k_sw128(w_hi_s + offB), idesc, true);
            mma_tf32_ts(dacc, a_hi, smem_desc_k_sw128(w_lo_s + offB), idesc, true);
            mma_tf32_ts(dacc, a_hi, smem_desc_k_sw128(w_hi_s + offB), idesc, true);
        }
        mma_commit(mbar);
    };

    for (int t = a.T - 1, step = 0; t >= 0; --t, ++step) {
        if (step > 0) {
            mbar_wait(mbar, (step - 1) & 1);
            tc_fence_after();
        }
        mbar_wait(gbar, step & 1);
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {
            float dh[16];
            if (step > 0) {
                tmem_ld_16x32(lane_addr + (static_cast<uint32_t>(16 * hh) << 16) + BT_DCOL, dh);
                tmem_ld_wait();
            } else {
#pragma unroll
                for (int rr = 0; rr < 2; ++rr) {
                    const int b = b_base + lrow0 + 16 * hh + 8 * rr;
                    float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
                    if (a.dh_last && b < a.B) ldg_v8(a.dh_last + ((long long)head * a.B + b) * BH + ucol, v);
#pragma unroll
                    for (int m = 0; m < 8; ++m) dh[4 * (m >> 1) + 2 * rr + (m & 1)] = v[m];
                }
            }
            float f_r[16], f_z[16], f_n[16], f_dhz[16];      // da_r, da_z, dgh_n = da_n * r, dh * z (fragment order)
#pragma unroll
            for (int rr = 0; rr < 2; ++rr) {
                const int s = 2 * hh + rr;
                const int lrow = lrow0 + 16 * hh + 8 * rr;
                const int b = b_base + lrow;
                // r | z | n of (row, units ucol..ucol+7): box 2g + cg, 32-byte run at 32tq, 16-byte units XOR (row & 7)
                float rv[8], zv[8], nv[8];
                {
                    const uint32_t u0 = static_cast<uint32_t>(((2 * tq) ^ (lrow & 7)) << 4), u1 = static_cast<uint32_t>(((2 * tq + 1) ^ (lrow & 7)) << 4);
                    const uint8_t* base = gate_s + cg * BT_GBOX + lrow * 128;
                    const float4 r0 = *reinterpret_cast<const float4*>(base + u0), r1 = *reinterpret_cast<const float4*>(base + u1);
                    const float4 z0 = *reinterpret_cast<const float4*>(base + 2 * BT_GBOX + u0), z1 = *reinterpret_cast<const float4*>(base + 2 * BT_GBOX + u1);
                    const float4 n0 = *reinterpret_cast<const float4*>(base + 4 * BT_GBOX + u0), n1 = *reinterpret_cast<const float4*>(base + 4 * BT_GBOX + u1);
                    rv[0] = r0.x; rv[1] = r0.y; rv[2] = r0.z; rv[3] = r0.w; rv[4] = r1.x; rv[5] = r1.y; rv[6] = r1.z; rv[7] = r1.w;
                    zv[0] = z0.x; zv[1] = z0.y; zv[2] = z0.z; zv[3] = z0.w; zv[4] = z1.x; zv[5] = z1.y; zv[6] = z1.z; zv[7] = z1.w;
                    nv[0] = n0.x; nv[1] = n0.y; nv[2] = n0.z; nv[3] = n0.w; nv[4] = n1.x; nv[5] = n1.y; nv[6] = n1.z; nv[7] = n1.w;
                }
                const float dp = dp_cur[s], dpm1 = dp_prev[s];
                const bool live = b < a.B;
                float o_r[8], o_z[8], o_n[8], o_g[8];
#pragma unroll
                for (int m = 0; m < 8; ++m) {
                    const int k = 4 * (m >> 1) + 2 * rr + (m & 1);
                    const float r = rv[m], z = zv[m], n = nv[m];     // dead rows: d == 0 and the tile holds finite values -> all products 0
                    const float d = fmaf(dp, wl[m], dh[k]);              // total dL/dh_t
                    const float dn = d * (1.f - z);
                    const float dz = d * (pf_hp[s][m] - n);
                    const float dan = dn * (1.f - n * n);
                    const float dr = dan * pf_ghn[s][m];
                    const float dar = dr * r * (1.f - r);
                    const float daz = dz * z * (1.f - z);
                    const float dgn = dan * r;
                    o_r[m] = dar; o_z[m] = daz; o_n[m] = dan; o_g[m] = dgn;
                    f_r[k] = dar; f_z[k] = daz; f_n[k] = dgn; f_dhz[k] = d * z;
                    acc[m] += dar; acc[8 + m] += daz; acc[16 + m] += dan; acc[24 + m] += dgn;
                    acc[32 + m] = fmaf(dpm1, pf_hp[s][m], acc[32 + m]);
                }
                if (tq == 0 && cg == 0) acc[40] += dp;
                dp_cur[s] = dpm1;
                // the inputs of step t-1 go into the load queue AHEAD of this step's stores (their registers are free now)
                if (t > 0) prefetch_slot(t - 1, s);
                if (live) {
                    const long long grow = head_row0 + (long long)t * a.B + b;
                    float* gdst = a.gates + grow * BG + ucol;
                    stg_v8(gdst, o_r[0], o_r[1], o_r[2], o_r[3], o_r[4], o_r[5], o_r[6], o_r[7]);
                    stg_v8(gdst + BH, o_z[0], o_z[1], o_z[2], o_z[3], o_z[4], o_z[5], o_z[6], o_z[7]);
                    stg_v8(gdst + 2 * BH, o_n[0], o_n[1], o_n[2], o_n[3], o_n[4], o_n[5], o_n[6], o_n[7]);
                    stg_v8(a.ghn + grow * BH + ucol, o_g[0], o_g[1], o_g[2], o_g[3], o_g[4], o_g[5], o_g[6], o_g[7]);
                }
            }
            // A operand (tf32 hi | lo) and the dh * z pre-load of the accumulator
            const uint32_t la = lane_addr + (static_cast<uint32_t>(16 * hh) << 16);
            tmem_st_16x32(la + BT_DCOL, f_dhz);
            {
                float lo[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) { const float hi = tf32_rna(f_r[k]); lo[k] = __fsub_rn(f_r[k], hi); f_r[k] = hi; }
                tmem_st_16x32(la + BT_ACOL_HI, f_r);
                tmem_st_16x32(la + BT_ACOL_LO, lo);
#pragma unroll
                for (int k = 0; k < 16; ++k) { const float hi = tf32_rna(f_z[k]); lo[k] = __fsub_rn(f_z[k], hi); f_z[k] = hi; }
                tmem_st_16x32(la + BT_ACOL_HI + BH, f_z);
                tmem_st_16x32(la + BT_ACOL_LO + BH, lo);
#pragma unroll
                for (int k = 0; k < 16; ++k) { const float hi = tf32_rna(f_n[k]); lo[k] = __fsub_rn(f_n[k], hi); f_n[k] = hi; }
                tmem_st_16x32(la + BT_ACOL_HI + 2 * BH, f_n);
                tmem_st_16x32(la + BT_ACOL_LO + 2 * BH, lo);
            }
        }
        tmem_st_wait();
        tc_fence_before();                   // operand / accumulator writes and gate-tile reads of this step are complete
        __syncthreads();
        if (threadIdx.x == 0) {
            if (t > 0) issue_gate_tma(t - 1);
            issue_step();
        }
    }
    // dh0 = dL/dh_{-1}
    mbar_wait(mbar, (a.T - 1) & 1);
    tc_fence_after();
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
        float dh[16];
        tmem_ld_16x32(lane_addr + (static_cast<uint32_t>(16 * hh) << 16) + BT_DCOL, dh);
        tmem_ld_wait();
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
            const int b = b_base + lrow0 + 16 * hh + 8 * rr;
            const int k0 = 2 * rr;
            if (b < a.B)
                stg_v8(a.dh0 + ((long long)head * a.B + b) * BH + ucol, dh[k0], dh[k0 + 1], dh[k0 + 4], dh[k0 + 5], dh[k0 + 8], dh[k0 + 9],
                       dh[k0 + 12], dh[k0 + 13]);
        }
    }
    // column sums: over the 8 row groups of the warp (shuffles), then over the 4 lane quadrants (shared memory), fixed order
#pragma unroll
    for (int i = 0; i < BT_NACC; ++i) {
        float v = acc[i];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        acc[i] = v;
    }
    float* red = reinterpret_cast<float*>(gate_s);      // [4 q][2 cg][4 tq][41]; the gate tile is free (last TMA consumed)
    if (tr == 0) {
#pragma unroll
        for (int i = 0; i < BT_NACC; ++i) red[((q * 2 + cg) * 4 + tq) * BT_NACC + i] = acc[i];
    }
    tc_fence_before();
    __syncthreads();
    float* ws = a.ws + ((long long)head * a.ntiles + tile) * BWS_TILE;
    for (int e = threadIdx.x; e < 8 * BT_NACC; e += BT_THREADS) {
        const int grp = e / BT_NACC, i = e % BT_NACC;     // grp = cg * 4 + tq
        const int cgi = grp >> 2, tqi = grp & 3;
        float s = 0.f;
#pragma unroll
        for (int qq = 0; qq < 4; ++qq) s += red[((qq * 2 + cgi) * 4 + tqi) * BT_NACC + i];
        const int u = 32 * cgi + 8 * tqi + (i & 7);
        if (i < 8) { ws[BWS_DBIH + u] = s; ws[BWS_DBHH + u] = s; }                        // r: dgh == dgi
        else if (i < 16) { ws[BWS_DBIH + BH + u] = s; ws[BWS_DBHH + BH + u] = s; }        // z
        else if (i < 24) ws[BWS_DBIH + 2 * BH + u] = s;                                  // n: db_ih
        else if (i < 32) ws[BWS_DBHH + 2 * BH + u] = s;                                  // n: db_hh (dgh_n)
        else if (i < 40) ws[BWS_DWLIN + u] = s;
        else if (grp == 0) ws[BWS_DBLIN] = s;
    }
    if (warp == 0) {
        tc_fence_after();
        tmem_dealloc<BT_TMEM_COLS>(tmem_base);
    }
}

int launch_gru_bwd_finalize(const float* ws, float* db_hh, float* db_ih, float* dw_lin, float* db_lin, int P, int ntiles, cudaStream_t st);

}  // namespace crvae

using namespace crvae;

// Tensor-core form of crvae_gru_bwd_deferred (same buffers and results to fp32 rounding; dhs is not supported).
extern "C" int crvae_gru_bwd_tc(float* gates, float* ghn, const float* hs, const float* h0, int64_t h0_head_stride,
                                const float* w_hh, const float* w_lin, const float* dpred, const float* dh_last,
                                float* db_hh, float* db_ih, float* dw_lin, float* db_lin, float* dh0, int P, int T, int B,
                                void* workspace, void* stream) {
    CRVAE_REQUIRE(gates && ghn && hs && h0 && w_hh && db_hh && db_ih && dh0 && workspace, "null operand");
    CRVAE_REQUIRE((w_lin == nullptr) == (dpred == nullptr), "w_lin and dpred go together");
    CRVAE_REQUIRE(w_lin == nullptr || (dw_lin && db_lin), "dw_lin/db_lin missing");
    CRVAE_REQUIRE(P >= 0 && T > 0 && B > 0, "bad size");
    auto al32 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31u) == 0; };
    CRVAE_REQUIRE(al32(gates) && al32(ghn) && al32(hs) && al32(h0) && al32(dh0) && aligned16(workspace) && aligned16(w_hh),
                  "32-byte alignment of the activation buffers");
    CRVAE_REQUIRE(dh_last == nullptr || al32(dh_last), "32-byte alignment");
    CRVAE_REQUIRE(h0_head_stride % 8 == 0, "h0 head stride must keep 32-byte alignment");
    if (P == 0) return 0;
    CUtensorMap tG;
    int rc;
    if ((rc = make_tmap_2d(&tG, gates, BG, (uint64_t)P * T * B, BG, 32, BT_ROWS, false))) return rc;
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gru_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BT_SMEM_BYTES);
        if (e != cudaSuccess) { set_error("gru_bwd_tc smem attr (%d B): %s", BT_SMEM_BYTES, cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    const int ntiles = (B + BT_ROWS - 1) / BT_ROWS;
    GruBwdTcArgs a{gates, ghn, hs, h0, (long long)h0_head_stride, w_hh, w_lin, dpred, dh_last, dh0, (float*)workspace, P, T, B, ntiles};
    gru_bwd_tc_kernel<<<dim3(ntiles, P), BT_THREADS, BT_SMEM_BYTES, (cudaStream_t)stream>>>(tG, a);
    if ((rc = check_launch("gru_bwd_tc_kernel"))) return rc;
    return launch_gru_bwd_finalize((const float*)workspace, db_hh, db_ih, dw_lin, db_lin, P, ntiles, (cudaStream_t)stream);
}
