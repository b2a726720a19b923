// Register-resident multi-head GRU recurrence (forward + BPTT) on warp-level tensor-core MMAs (mma.sync m16n8k8, 3xTF32).
//
// The recurrent step is a [rows x 64] x [64 x 192] product per (head, batch tile) followed by gate math that needs the
// r, z and n pre-activations of one (row, unit) in ONE thread, then a dependency on the next step.  Its cost is the
// serial chain of one step.  The tcgen05 kernels (gru_tc.cu / gru_tc_bwd.cu) pay an mbarrier round trip, a TMEM load
// and a TMEM store on that chain (7-10 us per step, one 128-row tile per SM); the exact FFMA kernels (gru_ll.cu) are
// bound by shared-memory operand traffic (3-5 us per step per 16-row tile).  Here
//   * a CTA (8 warps) owns a (head, 16-row) tile: M = 16 is exactly the m16n8k8 shape, so the accumulators land in the
//     registers of the thread that does the gate math -- no TMEM, no mbarrier, ONE __syncthreads per step;
//   * W_hh lives in REGISTERS for the whole sequence as pre-split tf32 hi/lo B fragments (96 registers per thread):
//     warp w owns hidden units 8w..8w+7, i.e. three n-tiles (r, z, n) forward / one n-tile over K = 192 backward;
//   * the only shared-memory traffic of a step is the 16 x 64 hidden tile (forward, 4 KB) or the 16 x 192 gate-gradient
//     tile (BPTT, 12 KB) that every warp reads as its A operand -- as conflict-free 128-bit loads, because the K index
//     of the product is permuted (thread q of a quad reads units 16c+4q..16c+4q+3, c = 0..3) identically in A and B;
//   * 3xTF32 (A.lo*B.hi + A.hi*B.lo + A.hi*B.hi, fp32 accumulate) as in the tcgen05 kernels: <= ~3e-7 relative;
//   * global traffic goes straight between registers and HBM: a quad covers one 32-byte sector per access, the next
//     step's inputs are requested before the step's MMAs, stores are fire-and-forget;
//   * the grid is persistent (one CTA per SM, contiguous ranges of (head, tile)); W_hh fragments are reloaded only
//     when the head changes, the next tile's first inputs and h0 are prefetched during the last step of a tile.
// Measured rates that size this (tools/micro/mma_rate.cu, B200): mma.sync m16n8k8 tf32 = 512 MAC/clk/SM, so one
// 3xTF32 step of a 16-row tile is 576 MMAs = 0.59 us of tensor pipe; FFMA = 120 FMA/clk/SM would be 0.8 us exact and
// shared-memory bound well above that.
//
// Reference arithmetic replaced: nn.GRU per-step linear_hh + cell (CRVAE_lorenz96.py:119, :208, :155, :166),
// nn.Linear(H,1) (:120) and autograd through them (:497).  Same buffers and in-place conventions as
// crvae_gru_fwd_ll / crvae_gru_bwd_ll (gates: gi -> r|z|n -> dgi; ghn: gh_n -> dgh_n; dW_hh deferred to
// crvae_gru_dwhh_tc).
#include "common.cuh"
#include "umma.cuh"
#include <cuda_fp16.h>
#include <type_traits>
#include <cstdlib>

namespace crvae {

int make_tmap_generic(CUtensorMap* m, const float* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, bool atom32b);
int launch_gru_bwd_finalize(const float* ws, float* db_hh, float* db_ih, float* dw_lin, float* db_lin, int P, int ntiles, cudaStream_t st);

namespace {

constexpr int MH = CRVAE_HIDDEN;       // 64
constexpr int MG = CRVAE_G;            // 192
constexpr int M_ROWS = 16;
constexpr int M_THREADS = 256;
constexpr int MH_LD = MH + 16;         // row stride of the h tile: rows g, g+1 land 16 banks apart (128-bit reads conflict-free)
constexpr int MD_LD = MG + 16;         // same for the gate-gradient tile
// per-(head, tile) partial sums, layout shared with gru_bwd_finalize_kernel (gru_recurrent.cu)
constexpr int MWS_TILE = MG * MH + 512;
constexpr int MWS_DBIH = MG * MH;
constexpr int MWS_DBHH = MG * MH + MG;
constexpr int MWS_DWLIN = MG * MH + 2 * MG;
constexpr int MWS_DBLIN = MG * MH + 2 * MG + MH;

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split_tf32(float v, uint32_t& hi, uint32_t& lo) {
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
    lo = __float_as_uint(v - __uint_as_float(hi));
}
// round-to-nearest tf32 of a FINITE value (what cvt.rna.tf32.f32 compiles to, minus its Inf/NaN guard) and the residual
__device__ __forceinline__ void split_tf32_fast(float v, uint32_t& hi, uint32_t& lo) {
    hi = (__float_as_uint(v) + 0x1000u) & 0xffffe000u;
    lo = __float_as_uint(v - __uint_as_float(hi));
}
// A-operand tile in fragment order: 16-byte fragment (pair p of adjacent k indices, row group g) = {x[g][2p], x[g+8][2p],
// x[g][2p+1], x[g+8][2p+1]} = the four A registers of the k-step whose k-slots q, q+4 map to k = 2p, 2p+1 -- exactly what the
// thread (g, q) of warp w produces for its unit pair.  The row-group index is XORed with the consumer's quad lane so that the
// 8 lanes of a quarter warp (2 row groups x 4 quad lanes) hit 8 different 16-byte bank groups.
__device__ __forceinline__ int frag_idx(int p, int g) { return (p * 8 + (g ^ (((p >> 1) & 3) << 1))) * 4; }
// fp16 hi | lo split of a finite value of moderate magnitude (hidden states, recurrent weights): hi = fp16(v), lo = fp16((v - hi) * 2^11).
// fp16 x fp16 products are exact in the fp32 accumulator; the cross terms are accumulated apart and scaled back by 2^-11, so the
// result carries 22 significant bits like 3xTF32 -- at half the MMA count (m16n8k16 covers 16 reduction indices per issue slot).
__device__ __forceinline__ void split_f16x2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const __half2 l = __floats2half2_rn((x - __low2float(h)) * 2048.f, (y - __high2float(h)) * 2048.f);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ float2 ldg2(const float* p) { return *reinterpret_cast<const float2*>(p); }
__device__ __forceinline__ void stg2(float* p, float x, float y) { *reinterpret_cast<float2*>(p) = make_float2(x, y); }

}  // namespace

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
struct GruMmaFwdArgs {
    float* gates; const float* b_ih; const float* w_hh; const float* b_hh;
    const float* h0; long long h0_stride;
    const float* w_lin; const float* b_lin;
    float* hs; float* ghn; float* pred;
    int P, T, B, t_skip, ntiles;
};

__global__ void __launch_bounds__(M_THREADS, 1) gru_fwd_mma_kernel(GruMmaFwdArgs a) {
    __shared__ __align__(16) float hbuf[2][M_ROWS * MH_LD];
    __shared__ __align__(16) float predp[2][M_ROWS][8];

    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int total = a.P * a.ntiles;
    const int first = (int)((long long)blockIdx.x * total / gridDim.x);
    const int last = (int)((long long)(blockIdx.x + 1) * total / gridDim.x);
    if (first >= last) return;
    const int ucol = 8 * w + 2 * q;               // this thread's two hidden units (accumulator columns 2q, 2q+1 of the warp's n-tiles)
    const bool has_lin = a.w_lin != nullptr;
    const int T = a.T, B = a.B;

    // B fragments of W_hh^T: n-tile `gate` column g <-> gate row gate*64 + 8w + g; k-slot (s, q) <-> unit 16*(s/2) + 4q + 2*(s%2),
    // k-slot (s, q+4) <-> that unit + 1  (the same permutation the A loads below apply to the hidden tile)
    uint32_t bhi[3][8][2], blo[3][8][2];
    float bhh[3][2], wl[2] = {0.f, 0.f}, blin = 0.f;
    auto load_head = [&](int head) {
        const float* __restrict__ W = a.w_hh + (long long)head * MG * MH;
#pragma unroll
        for (int gate = 0; gate < 3; ++gate) {
            const float* row = W + (long long)(gate * MH + 8 * w + g) * MH + 4 * q;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 v = __ldg(reinterpret_cast<const float4*>(row + 16 * c));
                split_tf32(v.x, bhi[gate][2 * c][0], blo[gate][2 * c][0]);
                split_tf32(v.y, bhi[gate][2 * c][1], blo[gate][2 * c][1]);
                split_tf32(v.z, bhi[gate][2 * c + 1][0], blo[gate][2 * c + 1][0]);
                split_tf32(v.w, bhi[gate][2 * c + 1][1], blo[gate][2 * c + 1][1]);
            }
            const float2 b2 = __ldg(reinterpret_cast<const float2*>(a.b_hh + (long long)head * MG + gate * MH + ucol));
            bhh[gate][0] = b2.x; bhh[gate][1] = b2.y;
        }
        if (has_lin) {
            const float2 w2 = __ldg(reinterpret_cast<const float2*>(a.w_lin + (long long)head * MH + ucol));
            wl[0] = w2.x; wl[1] = w2.y;
            blin = __ldg(a.b_lin + head);
        }
    };
    // gi of (head, tile, t) at this thread's positions: rows g, g+8 x units ucol, ucol+1 x gates r, z, n
    auto load_gi = [&](float2 (&gi)[3][2], int head, int tile, int t) {
        if (t < a.t_skip) {                        // zero-input step: gi is the bias (the projection did not write these rows)
#pragma unroll
            for (int gate = 0; gate < 3; ++gate) {
                const float2 b2 = __ldg(reinterpret_cast<const float2*>(a.b_ih + (long long)head * MG + gate * MH + ucol));
                gi[gate][0] = b2; gi[gate][1] = b2;
            }
            return;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int b = tile * M_ROWS + g + 8 * i;
            if (b < B) {
                const float* p = a.gates + (((long long)head * T + t) * B + b) * MG + ucol;
                gi[0][i] = ldg2(p); gi[1][i] = ldg2(p + MH); gi[2][i] = ldg2(p + 2 * MH);
            } else {
                gi[0][i] = gi[1][i] = gi[2][i] = make_float2(0.f, 0.f);
            }
        }
    };
    auto load_h0 = [&](float2 (&h)[2], int head, int tile) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int b = tile * M_ROWS + g + 8 * i;
            h[i] = b < B ? __ldg(reinterpret_cast<const float2*>(a.h0 + (long long)head * a.h0_stride + (long long)b * MH + ucol))
                         : make_float2(0.f, 0.f);
        }
    };

    int idx = first, head = first / a.ntiles, tile = first - head * a.ntiles;
    load_head(head);
    float2 hreg[2], gi[3][2];
    load_h0(hreg, head, tile);
    load_gi(gi, head, tile, 0);
    int cur = 0;
#pragma unroll
    for (int i = 0; i < 2; ++i) *reinterpret_cast<float2*>(&hbuf[0][(g + 8 * i) * MH_LD + ucol]) = hreg[i];
    __syncthreads();

    for (; idx < last; ++idx) {
        const bool more = idx + 1 < last;
        int nhead = head, ntile = tile + 1;
        if (ntile == a.ntiles) { ntile = 0; ++nhead; }
        const int b_tile = tile * M_ROWS;
        for (int t = 0; t < T; ++t) {
            const bool last_step = t == T - 1;
            // ---- A operand: h_{t-1} rows g, g+8, this thread's 16 k-slots each (4 x 128-bit, conflict-free) ----
            float4 av[2][4];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) av[i][c] = *reinterpret_cast<const float4*>(&hbuf[cur][(g + 8 * i) * MH_LD + 16 * c + 4 * q]);
            // ---- request the next step's inputs before the MMAs ----
            float2 gin[3][2], h0n[2];
            h0n[0] = h0n[1] = make_float2(0.f, 0.f);
            if (!last_step) load_gi(gin, head, tile, t + 1);
            else if (more) { load_gi(gin, nhead, ntile, 0); load_h0(h0n, nhead, ntile); }
            else {
#pragma unroll
                for (int gate = 0; gate < 3; ++gate) gin[gate][0] = gin[gate][1] = make_float2(0.f, 0.f);
            }
            // ---- gh = h . W_hh^T for the warp's 8 units x 3 gates: 8 k-steps x 3 gates x 3 tf32 terms ----
            float acc[3][4];
#pragma unroll
            for (int gate = 0; gate < 3; ++gate)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[gate][e] = 0.f;
#pragma unroll
            for (int s = 0; s < 8; ++s) {
                const float4 v0 = av[0][s >> 1], v1 = av[1][s >> 1];
                const float x0 = (s & 1) ? v0.z : v0.x, x1 = (s & 1) ? v1.z : v1.x;     // k-slot q     : rows g, g+8
                const float y0 = (s & 1) ? v0.w : v0.y, y1 = (s & 1) ? v1.w : v1.y;     // k-slot q + 4 : rows g, g+8
                uint32_t ahi[4], alo[4];
                split_tf32(x0, ahi[0], alo[0]); split_tf32(x1, ahi[1], alo[1]);
                split_tf32(y0, ahi[2], alo[2]); split_tf32(y1, ahi[3], alo[3]);
#pragma unroll
                for (int gate = 0; gate < 3; ++gate) mma_tf32(acc[gate], alo, bhi[gate][s]);
#pragma unroll
                for (int gate = 0; gate < 3; ++gate) mma_tf32(acc[gate], ahi, blo[gate][s]);
#pragma unroll
                for (int gate = 0; gate < 3; ++gate) mma_tf32(acc[gate], ahi, bhi[gate][s]);
            }
            // ---- gate math; operation order h' = (h - n) * z + n reproduces ATen's CPU GRU ----
            float2 hn[2];
            float ps[2];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float gir[2] = {gi[0][i].x, gi[0][i].y}, giz[2] = {gi[1][i].x, gi[1][i].y}, gnn[2] = {gi[2][i].x, gi[2][i].y};
                const float hold[2] = {hreg[i].x, hreg[i].y};
                float rr[2], zz[2], nn[2], gn[2], hv[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const float ghr = acc[0][2 * i + e] + bhh[0][e];
                    const float ghz = acc[1][2 * i + e] + bhh[1][e];
                    gn[e] = acc[2][2 * i + e] + bhh[2][e];
                    rr[e] = sigmoidf_fast(gir[e] + ghr);
                    zz[e] = sigmoidf_fast(giz[e] + ghz);
                    nn[e] = tanhf_fast(__fadd_rn(gnn[e], __fmul_rn(rr[e], gn[e])));
                    hv[e] = __fadd_rn(__fmul_rn(__fsub_rn(hold[e], nn[e]), zz[e]), nn[e]);
                }
                hn[i] = make_float2(hv[0], hv[1]);
                ps[i] = fmaf(hv[1], wl[1], hv[0] * wl[0]);
                const int b = b_tile + g + 8 * i;
                if (b < B) {
                    const long long grow = ((long long)head * T + t) * B + b;
                    float* pg = a.gates + grow * MG + ucol;
                    stg2(pg, rr[0], rr[1]); stg2(pg + MH, zz[0], zz[1]); stg2(pg + 2 * MH, nn[0], nn[1]);
                    stg2(a.hs + grow * MH + ucol, hv[0], hv[1]);
                    stg2(a.ghn + grow * MH + ucol, gn[0], gn[1]);
                }
            }
            // ---- hand h_t (or, after the last step, the next tile's h0) to the other warps ----
            {
                const float2 o0 = last_step ? h0n[0] : hn[0], o1 = last_step ? h0n[1] : hn[1];
                *reinterpret_cast<float2*>(&hbuf[cur ^ 1][g * MH_LD + ucol]) = o0;
                *reinterpret_cast<float2*>(&hbuf[cur ^ 1][(g + 8) * MH_LD + ucol]) = o1;
                hreg[0] = o0; hreg[1] = o1;
            }
            if (has_lin) {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    float p = ps[i];
                    p += __shfl_xor_sync(0xffffffffu, p, 1);
                    p += __shfl_xor_sync(0xffffffffu, p, 2);
                    if (q == 0) predp[cur][g + 8 * i][w] = p;
                }
            }
            __syncthreads();
            if (has_lin && tid < M_ROWS && b_tile + tid < B) {
                const float4 p0 = *reinterpret_cast<const float4*>(&predp[cur][tid][0]);
                const float4 p1 = *reinterpret_cast<const float4*>(&predp[cur][tid][4]);
                a.pred[((long long)head * T + t) * B + b_tile + tid] = (((p0.x + p0.y) + (p0.z + p0.w)) + ((p1.x + p1.y) + (p1.z + p1.w))) + blin;
            }
            cur ^= 1;
#pragma unroll
            for (int gate = 0; gate < 3; ++gate) { gi[gate][0] = gin[gate][0]; gi[gate][1] = gin[gate][1]; }
        }
        if (more && nhead != head) load_head(nhead);
        head = nhead; tile = ntile;
    }
}

// ------------------------------------------------------------------------------------------------------------
// backward (dW_hh deferred to crvae_gru_dwhh_tc)
// ------------------------------------------------------------------------------------------------------------
struct GruMmaBwdArgs {
    float* gates; float* ghn; const float* hs;
    const float* h0; long long h0_stride;
    const float* w_hh; const float* w_lin;
    const float* dpred; const float* dh_last; const float* dhs;
    float* dh0; float* ws;
    int P, T, B, ntiles;
};

struct BwdIn {                // one step's inputs at a thread's positions (rows g, g+8 x units ucol, ucol+1)
    float2 r[2], z[2], n[2], gn[2], hp[2], de[2];
    float dp[2];
};

template <bool HAS_DHS>
__global__ void __launch_bounds__(M_THREADS, 1) gru_bwd_mma_kernel(GruMmaBwdArgs a) {
    __shared__ __align__(16) float dsm[2][M_ROWS * MG];       // dgh of the step in A-fragment order (frag_idx), double-buffered

    const int tid = threadIdx.x, w = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int total = a.P * a.ntiles;
    const int first = (int)((long long)blockIdx.x * total / gridDim.x);
    const int last = (int)((long long)(blockIdx.x + 1) * total / gridDim.x);
    if (first >= last) return;
    const int ucol = 8 * w + 2 * q;
    const bool has_lin = a.w_lin != nullptr;
    const int T = a.T, B = a.B;

    // B fragments of W_hh [192 x 64]: n-tile column g <-> output unit 8w + g; k-slot (s, q) <-> gate row 16*(s/2) + 4q + 2*(s%2)
    uint32_t bhi[24][2], blo[24][2];
    float wl[2] = {0.f, 0.f};
    auto load_head = [&](int head) {
        const float* __restrict__ W = a.w_hh + (long long)head * MG * MH + 8 * w + g;
#pragma unroll
        for (int c = 0; c < 12; ++c) {
            const float* p = W + (long long)(16 * c + 4 * q) * MH;
            const float v0 = __ldg(p), v1 = __ldg(p + MH), v2 = __ldg(p + 2 * MH), v3 = __ldg(p + 3 * MH);
            split_tf32(v0, bhi[2 * c][0], blo[2 * c][0]);
            split_tf32(v1, bhi[2 * c][1], blo[2 * c][1]);
            split_tf32(v2, bhi[2 * c + 1][0], blo[2 * c + 1][0]);
            split_tf32(v3, bhi[2 * c + 1][1], blo[2 * c + 1][1]);
        }
        if (has_lin) {
            const float2 w2 = __ldg(reinterpret_cast<const float2*>(a.w_lin + (long long)head * MH + ucol));
            wl[0] = w2.x; wl[1] = w2.y;
        }
    };
    auto load_in = [&](BwdIn& in, int head, int tile, int t) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int b = tile * M_ROWS + g + 8 * i;
            if (b < B) {
                const long long grow = ((long long)head * T + t) * B + b;
                const float* pg = a.gates + grow * MG + ucol;
                in.r[i] = ldg2(pg); in.z[i] = ldg2(pg + MH); in.n[i] = ldg2(pg + 2 * MH);
                in.gn[i] = ldg2(a.ghn + grow * MH + ucol);
                in.hp[i] = t > 0 ? ldg2(a.hs + (grow - B) * MH + ucol)
                                 : __ldg(reinterpret_cast<const float2*>(a.h0 + (long long)head * a.h0_stride + (long long)b * MH + ucol));
                in.de[i] = HAS_DHS ? __ldg(reinterpret_cast<const float2*>(a.dhs + grow * MH + ucol)) : make_float2(0.f, 0.f);
                in.dp[i] = has_lin ? __ldg(a.dpred + grow) : 0.f;
            } else {
                in.r[i] = in.z[i] = in.n[i] = in.gn[i] = in.hp[i] = in.de[i] = make_float2(0.f, 0.f);
                in.dp[i] = 0.f;
            }
        }
    };

    int idx = first, head = first / a.ntiles, tile = first - head * a.ntiles;
    load_head(head);
    BwdIn in;
    load_in(in, head, tile, T - 1);
    int buf = 0;

    for (; idx < last; ++idx) {
        const bool more = idx + 1 < last;
        int nhead = head, ntile = tile + 1;
        if (ntile == a.ntiles) { ntile = 0; ++nhead; }
        const int b_tile = tile * M_ROWS;
        // ---- per-tile state ----
        float dh[2][2];
        float2 hcur[2];                            // h_t of the step being processed (dw_lin += dpred[t] * h_t); starts as h_{T-1}
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int b = b_tile + g + 8 * i;
            float2 v = make_float2(0.f, 0.f);
            hcur[i] = make_float2(0.f, 0.f);
            if (b < B) {
                if (a.dh_last) v = __ldg(reinterpret_cast<const float2*>(a.dh_last + ((long long)head * B + b) * MH + ucol));
                if (has_lin) hcur[i] = ldg2(a.hs + (((long long)head * T + (T - 1)) * B + b) * MH + ucol);
            }
            dh[i][0] = v.x; dh[i][1] = v.y;
        }
        float dbih[3][2], dbhn[2] = {0.f, 0.f}, dwl[2] = {0.f, 0.f}, dbl = 0.f;
#pragma unroll
        for (int gate = 0; gate < 3; ++gate) dbih[gate][0] = dbih[gate][1] = 0.f;

        for (int t = T - 1; t >= 0; --t) {
            // ---- request the next step's inputs (or the next tile's first) ----
            BwdIn nx;
            if (t > 0) load_in(nx, head, tile, t - 1);
            else if (more) load_in(nx, nhead, ntile, T - 1);
            else {
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    nx.r[i] = nx.z[i] = nx.n[i] = nx.gn[i] = nx.hp[i] = nx.de[i] = make_float2(0.f, 0.f);
                    nx.dp[i] = 0.f;
                }
            }
            // ---- pointwise cell backward at this thread's positions ----
            float dhz[2][2];
            float fr[1][4], fz[1][4], fn_[1][4];
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float r_[2] = {in.r[i].x, in.r[i].y}, z_[2] = {in.z[i].x, in.z[i].y}, n_[2] = {in.n[i].x, in.n[i].y},
                            gn_[2] = {in.gn[i].x, in.gn[i].y}, hp_[2] = {in.hp[i].x, in.hp[i].y}, de_[2] = {in.de[i].x, in.de[i].y};
                const float dp = in.dp[i];
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
                }
                if (w == 0 && q == 0) dbl += dp;
                const int b = b_tile + g + 8 * i;
                if (b < B) {
                    const long long grow = ((long long)head * T + t) * B + b;
                    float* pg = a.gates + grow * MG + ucol;
                    stg2(pg, dar[0], dar[1]); stg2(pg + MH, daz[0], daz[1]); stg2(pg + 2 * MH, dan[0], dan[1]);
                    stg2(a.ghn + grow * MH + ucol, dgn[0], dgn[1]);
                }
                fr[0][i] = dar[0]; fr[0][2 + i] = dar[1];
                fz[0][i] = daz[0]; fz[0][2 + i] = daz[1];
                fn_[0][i] = dgn[0]; fn_[0][2 + i] = dgn[1];
            }
            {   // this thread's fragments {x[g][u], x[g+8][u], x[g][u+1], x[g+8][u+1]} of the three gate blocks
                const int pw = 4 * w + q;
                *reinterpret_cast<float4*>(&dsm[buf][frag_idx(pw, g)]) = make_float4(fr[0][0], fr[0][1], fr[0][2], fr[0][3]);
                *reinterpret_cast<float4*>(&dsm[buf][frag_idx(32 + pw, g)]) = make_float4(fz[0][0], fz[0][1], fz[0][2], fz[0][3]);
                *reinterpret_cast<float4*>(&dsm[buf][frag_idx(64 + pw, g)]) = make_float4(fn_[0][0], fn_[0][1], fn_[0][2], fn_[0][3]);
            }
            __syncthreads();
            // ---- dh_{t-1} = dh_t * z + dgh . W_hh : 24 k-steps x 3 tf32 terms, one accumulator per term ----
            float acc[3][4];
#pragma unroll
            for (int k = 0; k < 3; ++k)
#pragma unroll
                for (int e = 0; e < 4; ++e) acc[k][e] = 0.f;
#pragma unroll
            for (int grp = 0; grp < 6; ++grp) {
                float4 av[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int ks = 4 * grp + k;
                    av[k] = *reinterpret_cast<const float4*>(&dsm[buf][frag_idx(8 * (ks >> 1) + 2 * q + (ks & 1), g)]);
                }
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const int ks = 4 * grp + k;
                    uint32_t ahi[4], alo[4];
                    split_tf32_fast(av[k].x, ahi[0], alo[0]); split_tf32_fast(av[k].y, ahi[1], alo[1]);
                    split_tf32_fast(av[k].z, ahi[2], alo[2]); split_tf32_fast(av[k].w, ahi[3], alo[3]);
                    mma_tf32(acc[0], alo, bhi[ks]);
                    mma_tf32(acc[1], ahi, blo[ks]);
                    mma_tf32(acc[2], ahi, bhi[ks]);
                }
            }
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    dh[i][e] = __fadd_rn(dhz[i][e], __fadd_rn(__fadd_rn(acc[0][2 * i + e], acc[1][2 * i + e]), acc[2][2 * i + e]));
            // dw_lin += dpred[t] * h_t; h_t was this loop's previous h_{t-1} (the tile start loaded h_{T-1}) -- after the MMAs, so
            // that nothing on the step's critical path waits for a load that was just issued
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                dwl[0] = fmaf(in.dp[i], hcur[i].x, dwl[0]); dwl[1] = fmaf(in.dp[i], hcur[i].y, dwl[1]);
                hcur[i] = in.hp[i];
            }
            buf ^= 1;
            in = nx;
        }

        // ---- tile outputs: dh0, per-tile partial column sums (fixed order: deterministic) ----
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int b = b_tile + g + 8 * i;
            if (b < B) stg2(a.dh0 + ((long long)head * B + b) * MH + ucol, dh[i][0], dh[i][1]);
        }
        float v[11] = {dbih[0][0], dbih[0][1], dbih[1][0], dbih[1][1], dbih[2][0], dbih[2][1], dbhn[0], dbhn[1], dwl[0], dwl[1], dbl};
#pragma unroll
        for (int k = 0; k < 11; ++k) {
            v[k] += __shfl_xor_sync(0xffffffffu, v[k], 4);
            v[k] += __shfl_xor_sync(0xffffffffu, v[k], 8);
            v[k] += __shfl_xor_sync(0xffffffffu, v[k], 16);
        }
        if (g == 0) {
            float* ws = a.ws + ((long long)head * a.ntiles + tile) * MWS_TILE;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int u = ucol + e;
                ws[MWS_DBIH + u] = v[e];            ws[MWS_DBHH + u] = v[e];                  // r: dgh == dgi
                ws[MWS_DBIH + MH + u] = v[2 + e];   ws[MWS_DBHH + MH + u] = v[2 + e];         // z
                ws[MWS_DBIH + 2 * MH + u] = v[4 + e];                                         // n: db_ih
                ws[MWS_DBHH + 2 * MH + u] = v[6 + e];                                         // n: db_hh (dgh_n)
                ws[MWS_DWLIN + u] = v[8 + e];
            }
            if (w == 0 && q == 0) ws[MWS_DBLIN] = v[10];
        }
        if (more && nhead != head) load_head(nhead);
        head = nhead; tile = ntile;
    }
}


// ------------------------------------------------------------------------------------------------------------
// Warp-specialised forward: the tensor pipe and the gate math run in different warps.
//
// In the kernel above all 8 warps walk the step in lockstep -- operand loads, 72 MMAs per warp (tensor pipe: 8 clk per
// MMA per SM sub-partition, two warps each = 1152 clk), gate math (a 1000-1300 clk chain of MUFU / shuffle / shared
// memory latencies for 4 elements per thread, measured alone), stores, barrier -- so the tensor pipe idles during the
// gate math and vice versa (3000 clk per step).  Registers (W_hh fragments: 96) rule out a second CTA per SM.  Measured
// dead ends on the way here: interleaving one tile's MMAs with another tile's gate math in one instruction stream
// (3250 clk per tile-step: with two warps per scheduler the stream stalls on its own fixed-latency dependencies) and
// running the two warps of a scheduler one slot apart on two tiles (3500: a slot lasts as long as the gate math).
// So the roles are split (512 threads; with the accumulators handed over through shared memory both roles fit in the 128
// registers per thread that 512 threads leave -- the M warps need 122, the P warps under 100 -- so no setmaxnreg):
//   * warps 0-7, "M": hold W_hh, wait for a tile's hidden state, issue its 72 MMAs from pre-split A fragments (pure
//     LDS + HMMA), hand the accumulators to their partner warp through shared memory;
//   * warps 8-15, "P": gate math, staging of r|z|n, h_t, gh_n and of the next A operand (already split into tf32
//     hi | lo, in fragment order: one 128-bit load = one MMA operand), output Linear, TMA plumbing.
// A CTA works on TWO 16-row tiles of one head at a time (streams a, b): M_b(n) runs while P_a(n) does, M_a(n+1) while
// P_b(n) does; with few tiles (<= #SMs) stream b is idle.  Synchronisation is dataflow only (mbarriers): h_ready[s]
// (P -> M, whole tile), acc_ready[s][w] (M warp w -> P warp w), the TMA ring; the P warps sync among themselves with a
// named barrier.  Global traffic does not touch the load/store unit: the 16 x 192 gate slab of a step arrives by ONE
// TMA tile copy (SWIZZLE_128B: the 64-bit accesses of the accumulator layout are conflict-free) into a 3-deep ring, is
// overwritten in place with r|z|n and leaves by TMA together with the staged h / gh_n tiles; rows past B are clipped
// by the tensor map.
// ------------------------------------------------------------------------------------------------------------
constexpr int PF_NS = 3;               // gate-slab ring depth per stream
constexpr int WS_THREADS = 512;

// TMA-written tiles first (1024-byte aligned: the 128-byte swizzle is a function of the shared-memory address)
struct __align__(1024) FwdStreamSmem {
    float slab[PF_NS][M_ROWS * MG];        // [row][6 x 128 B] swizzled: gi -> r|z|n
    float hst[2][M_ROWS * MH];             // [row][2 x 128 B] swizzled: h_t rows for the store
    float gst[2][M_ROWS * MH];             // gh_n rows
    float hbuf[2][2][M_ROWS * MH];         // [step parity][hi | lo]: h_{t-1} as tf32 hi | lo in A-FRAGMENT order (frag_idx); double-buffered
                                           // because a P warp starts as soon as ITS M warp is done, while other M warps still read the tile
    float accb[8][3][32 * 4];              // gh accumulators, [M warp][gate][lane] x float4 (C-fragment order)
    float predp[2][M_ROWS][8];
    uint64_t slab_full[PF_NS];
    uint64_t h_ready;
    uint64_t acc_ready[8];
};
struct __align__(1024) FwdPipeSmem {
    FwdStreamSmem st[2];
    float bih[2][MG];                      // b_ih of the head (zero-input steps), by head parity
};
// float index of element (row, col) in a TMA tile of C 128-byte chunks per row written with SWIZZLE_128B:
// the 16-byte unit index inside a 128-byte line is XORed with the line index mod 8
__device__ __forceinline__ int sw_idx(int C, int row, int col) {
    const int line = row * C + (col >> 5), unit = (col & 31) >> 2;
    return line * 32 + (((unit ^ (line & 7)) << 2) | (col & 3));
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* m, const void* smem_src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(umma::smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void p_warps_sync() { asm volatile("bar.sync 1, 256;" ::: "memory"); }    // the 8 P warps only

#ifdef CRVAE_MMA_TIMING
__device__ long long g_mma_dbg[16 * 8];            // per-warp cycle counters of block 0 (tools/mma_timing.py)
#define MMA_CLK() clock64()
#else
#define MMA_CLK() 0ll
#endif
struct FwdPos { int j, head, tile, t, vrows; };   // (uniform) a stream's step: pair index in the CTA's range, head, tile, step; vrows == 0: idle

template <bool F16>
__global__ void __launch_bounds__(WS_THREADS, 1)
gru_fwd_mma_ws_kernel(const __grid_constant__ CUtensorMap tmG, const __grid_constant__ CUtensorMap tmH, const __grid_constant__ CUtensorMap tmN,
                      GruMmaFwdArgs a, int pf, int npph) {
    using namespace umma;
    extern __shared__ __align__(1024) uint8_t mma_smem_raw[];
    FwdPipeSmem& sm = *reinterpret_cast<FwdPipeSmem*>(mma_smem_raw);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int w = warp & 7;                       // unit group: hidden units 8w .. 8w+7 (M warp w and P warp 8 + w)
    const bool is_m = warp < 8;
    const int total = a.P * npph;                 // work units: (head, pair of tiles)
    const int first = (int)((long long)blockIdx.x * total / gridDim.x);
    const int last = (int)((long long)(blockIdx.x + 1) * total / gridDim.x);
    if (first >= last) return;
    const int ucol = 8 * w + 2 * q;
    const bool has_lin = a.w_lin != nullptr;
    const int T = a.T, B = a.B;
    const int nsteps = (last - first) * T;        // steps of each stream

    auto pos_of_pair = [&](int j, int s) {        // stream s at the start of pair j
        FwdPos p;
        p.j = j; p.t = 0;
        const int pj = first + j;
        p.head = pj / npph;
        p.tile = (pj - p.head * npph) * pf + s;
        const int b0 = p.tile * M_ROWS;
        p.vrows = (pj < last && s < pf && b0 < B) ? min(M_ROWS, B - b0) : 0;
        return p;
    };
    auto advance = [&](FwdPos& p, int s) {        // position of the stream's next step
        if (++p.t == T) p = pos_of_pair(p.j + 1, s);
    };

    if (tid == 32) { prefetch_tmap(&tmG); prefetch_tmap(&tmH); prefetch_tmap(&tmN); }
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
#pragma unroll
            for (int i = 0; i < PF_NS; ++i) mbar_init(&sm.st[s].slab_full[i], 1);
            mbar_init(&sm.st[s].h_ready, 1);
#pragma unroll
            for (int i = 0; i < 8; ++i) mbar_init(&sm.st[s].acc_ready[i], 1);
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (is_m) {
        // =============================================================== M warps ===============================================================
        // B fragments of W_hh^T: n-tile `gate` column g <-> gate row gate*64 + 8w + g; k-step s, k-slots q / q+4 <-> units 2p, 2p+1 with
        // p = 8*(s/2) + 2q + (s%2)  (the pair whose A fragment frag_idx(p, g) holds)
        uint32_t bhi[3][8][2], blo[3][8][2];      // (fp16 form: k-steps 0..3 only)
        int cur_head = -1;
        auto load_head = [&](int head) {
            const float* __restrict__ W = a.w_hh + (long long)head * MG * MH;
#pragma unroll
            for (int gate = 0; gate < 3; ++gate) {
                if (F16) {                        // k-step s (16 indices): k-slots 2q, 2q+1 <-> units 16s + 2q (+1), k-slots 2q+8, 2q+9 <-> units 16s + 8 + 2q (+1)
                    const float* row = W + (long long)(gate * MH + 8 * w + g) * MH + 2 * q;
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        const float2 v1 = __ldg(reinterpret_cast<const float2*>(row + 16 * ks));
                        const float2 v2 = __ldg(reinterpret_cast<const float2*>(row + 16 * ks + 8));
                        split_f16x2(v1.x, v1.y, bhi[gate][ks][0], blo[gate][ks][0]);
                        split_f16x2(v2.x, v2.y, bhi[gate][ks][1], blo[gate][ks][1]);
                    }
                } else {
                    const float* row = W + (long long)(gate * MH + 8 * w + g) * MH + 4 * q;
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const float4 v = __ldg(reinterpret_cast<const float4*>(row + 16 * c));
                        split_tf32_fast(v.x, bhi[gate][2 * c][0], blo[gate][2 * c][0]);
                        split_tf32_fast(v.y, bhi[gate][2 * c][1], blo[gate][2 * c][1]);
                        split_tf32_fast(v.z, bhi[gate][2 * c + 1][0], blo[gate][2 * c + 1][0]);
                        split_tf32_fast(v.w, bhi[gate][2 * c + 1][1], blo[gate][2 * c + 1][1]);
                    }
                }
            }
            cur_head = head;
        };
        FwdPos pos[2] = {pos_of_pair(0, 0), pos_of_pair(0, 1)};
        uint32_t cnt[2] = {0, 0};                 // MMAs done per stream: parity of h_ready / acc_ready
        long long dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const long long t_begin = MMA_CLK();
        for (int n = 0; n < nsteps; ++n) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (pos[s].vrows > 0) {
                    if (pos[s].head != cur_head) load_head(pos[s].head);
                    FwdStreamSmem& st = sm.st[s];
                    const long long mc0 = MMA_CLK();
                    mbar_wait(&st.h_ready, cnt[s] & 1u);
                    const long long mc1 = MMA_CLK();
                    const float* hb_hi = st.hbuf[cnt[s] & 1][0];
                    const float* hb_lo = st.hbuf[cnt[s] & 1][1];
                    float acc[3][4];
#pragma unroll
                    for (int gate = 0; gate < 3; ++gate)
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[gate][e] = 0.f;
                    if (F16) {
                        // four k-steps of 16: one 128-bit load per operand half; cross terms (scaled by 2^11) in their own accumulators
                        float accx[3][4];
#pragma unroll
                        for (int gate = 0; gate < 3; ++gate)
#pragma unroll
                            for (int e = 0; e < 4; ++e) accx[gate][e] = 0.f;
                        uint4 fh[4], fl[4];
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const int o = ((ks * 4 + q) * 8 + (g ^ (q << 1))) * 4;
                            fh[ks] = *reinterpret_cast<const uint4*>(&hb_hi[o]);
                            fl[ks] = *reinterpret_cast<const uint4*>(&hb_lo[o]);
                        }
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks) {
                            const uint32_t ahi[4] = {fh[ks].x, fh[ks].y, fh[ks].z, fh[ks].w};
                            const uint32_t alo[4] = {fl[ks].x, fl[ks].y, fl[ks].z, fl[ks].w};
#pragma unroll
                            for (int gate = 0; gate < 3; ++gate) mma_f16(accx[gate], alo, bhi[gate][ks]);
#pragma unroll
                            for (int gate = 0; gate < 3; ++gate) mma_f16(accx[gate], ahi, blo[gate][ks]);
#pragma unroll
                            for (int gate = 0; gate < 3; ++gate) mma_f16(acc[gate], ahi, bhi[gate][ks]);
                        }
#pragma unroll
                        for (int gate = 0; gate < 3; ++gate)
#pragma unroll
                            for (int e = 0; e < 4; ++e) acc[gate][e] = fmaf(accx[gate][e], 1.f / 2048.f, acc[gate][e]);
                    } else {
                    // A fragments in quarters (2 k-steps = 4 x 128-bit loads each), loaded one quarter ahead of the MMAs that use them
                    uint4 fa[2][4];               // [ping-pong][hi k0, hi k1, lo k0, lo k1]
                    auto load_q = [&](int qi, uint4 (&f)[4]) {
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const int ks = 2 * qi + k;
                            const int o = frag_idx(8 * (ks >> 1) + 2 * q + (ks & 1), g);
                            f[k] = *reinterpret_cast<const uint4*>(&hb_hi[o]);
                            f[2 + k] = *reinterpret_cast<const uint4*>(&hb_lo[o]);
                        }
                    };
                    load_q(0, fa[0]);
#pragma unroll
                    for (int qi = 0; qi < 4; ++qi) {
                        if (qi < 3) load_q(qi + 1, fa[(qi + 1) & 1]);
#pragma unroll
                        for (int k = 0; k < 2; ++k) {
                            const int ks = 2 * qi + k;
                            const uint4 vh = fa[qi & 1][k], vl = fa[qi & 1][2 + k];
                            const uint32_t ahi[4] = {vh.x, vh.y, vh.z, vh.w};
                            const uint32_t alo[4] = {vl.x, vl.y, vl.z, vl.w};
#pragma unroll
                            for (int gate = 0; gate < 3; ++gate) mma_tf32(acc[gate], alo, bhi[gate][ks]);
#pragma unroll
                            for (int gate = 0; gate < 3; ++gate) mma_tf32(acc[gate], ahi, blo[gate][ks]);
#pragma unroll
                            for (int gate = 0; gate < 3; ++gate) mma_tf32(acc[gate], ahi, bhi[gate][ks]);
                        }
                    }
                    }
#pragma unroll
                    for (int gate = 0; gate < 3; ++gate)
                        *reinterpret_cast<float4*>(&st.accb[w][gate][lane * 4]) = make_float4(acc[gate][0], acc[gate][1], acc[gate][2], acc[gate][3]);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&st.acc_ready[w]);
                    ++cnt[s];
#ifdef CRVAE_MMA_TIMING
                    { const long long mc2 = MMA_CLK(); dbg[0] += mc1 - mc0; dbg[1] += mc2 - mc1; dbg[4] += 1; }
#endif
                }
                advance(pos[s], s);
            }
        }
#ifdef CRVAE_MMA_TIMING
        if (blockIdx.x == 0 && lane == 0) { dbg[7] = MMA_CLK() - t_begin; for (int i = 0; i < 8; ++i) g_mma_dbg[warp * 8 + i] = dbg[i]; }
#endif
    } else {
        // =============================================================== P warps ===============================================================
        const int ptid = tid - 256;
        float bhh[3][2], wl[2] = {0.f, 0.f}, blin = 0.f;
        int cur_head = -1;
        auto load_head = [&](int head) {
#pragma unroll
            for (int gate = 0; gate < 3; ++gate) {
                const float2 b2 = __ldg(reinterpret_cast<const float2*>(a.b_hh + (long long)head * MG + gate * MH + ucol));
                bhh[gate][0] = b2.x; bhh[gate][1] = b2.y;
            }
            if (has_lin) {
                const float2 w2 = __ldg(reinterpret_cast<const float2*>(a.w_lin + (long long)head * MH + ucol));
                wl[0] = w2.x; wl[1] = w2.y;
                blin = __ldg(a.b_lin + head);
            }
            if (ptid < MG) sm.bih[head & 1][ptid] = __ldg(a.b_ih + (long long)head * MG + ptid);      // read after a P-warp barrier
            cur_head = head;
        };
        auto load_h0 = [&](float2 (&h)[2], const FwdPos& p) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = g + 8 * i;
                h[i] = r < p.vrows ? __ldg(reinterpret_cast<const float2*>(a.h0 + (long long)p.head * a.h0_stride +
                                                                           (long long)(p.tile * M_ROWS + r) * MH + ucol))
                                   : make_float2(0.f, 0.f);
            }
        };
        auto stage_h = [&](int s, int buf, const float2 (&h)[2]) {     // this thread's h values, split, into the A-operand tile
            if (F16) {                            // unit pair 4w + q = k-step w/2, fragment half w%2, quad lane q: {row g, row g+8} as two half2
                uint32_t hi[2], lo[2];
                split_f16x2(h[0].x, h[0].y, hi[0], lo[0]);
                split_f16x2(h[1].x, h[1].y, hi[1], lo[1]);
                const int o = (((w >> 1) * 4 + q) * 8 + (g ^ (q << 1))) * 4 + (w & 1) * 2;
                *reinterpret_cast<uint2*>(&sm.st[s].hbuf[buf][0][o]) = make_uint2(hi[0], hi[1]);
                *reinterpret_cast<uint2*>(&sm.st[s].hbuf[buf][1][o]) = make_uint2(lo[0], lo[1]);
                return;
            }
            uint32_t hi[4], lo[4];
            split_tf32_fast(h[0].x, hi[0], lo[0]); split_tf32_fast(h[1].x, hi[1], lo[1]);
            split_tf32_fast(h[0].y, hi[2], lo[2]); split_tf32_fast(h[1].y, hi[3], lo[3]);
            const int o = frag_idx(4 * w + q, g);
            *reinterpret_cast<uint4*>(&sm.st[s].hbuf[buf][0][o]) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            *reinterpret_cast<uint4*>(&sm.st[s].hbuf[buf][1][o]) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        };
        // ---- plumbing (thread 256): TMA tiles {32 floats, chunks, 16 rows, 1 (head, t)}; rows past B are clipped ----
        auto issue_load = [&](int s, int n) {         // gate slab of step n -> ring slot n % NS
            const int j = n / T;
            FwdPos p = pos_of_pair(j, s);
            p.t = n - j * T;
            const int slot = n % PF_NS;
            mbar_arrive_expect_tx(&sm.st[s].slab_full[slot], p.vrows > 0 ? (uint32_t)(M_ROWS * MG * 4) : 0u);
            if (p.vrows > 0) tma_load_4d(sm.st[s].slab[slot], &tmG, &sm.st[s].slab_full[slot], 0, 0, p.tile * M_ROWS, p.head * T + p.t);
        };
        auto plumb = [&](int s, int n, const FwdPos& p) {    // all P warps have staged step n of stream s
            // the previous store group of THIS stream has read its tiles (two streams: all but the latest group, which is the other stream's)
            if (pf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else         asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (p.vrows > 0) {
                const int c2 = p.tile * M_ROWS, c3 = p.head * T + p.t;
                tma_store_4d(&tmG, sm.st[s].slab[n % PF_NS], 0, 0, c2, c3);
                tma_store_4d(&tmH, sm.st[s].hst[n & 1], 0, 0, c2, c3);
                tma_store_4d(&tmN, sm.st[s].gst[n & 1], 0, 0, c2, c3);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (n >= 1 && n - 1 + PF_NS < nsteps) issue_load(s, n - 1 + PF_NS);
        };

        int o_r[2], o_z[2], o_n[2], o_h[2];       // this thread's (swizzled) positions in the slab / staging tiles
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            o_r[i] = sw_idx(6, g + 8 * i, ucol); o_z[i] = sw_idx(6, g + 8 * i, MH + ucol); o_n[i] = sw_idx(6, g + 8 * i, 2 * MH + ucol);
            o_h[i] = sw_idx(2, g + 8 * i, ucol);
        }
        FwdPos pos[2] = {pos_of_pair(0, 0), pos_of_pair(0, 1)};
        float2 hreg[2][2];
        uint32_t cnt[2] = {0, 0};                 // gate-math ops done per stream: parity of acc_ready
        long long dbg[8] = {0, 0, 0, 0, 0, 0, 0, 0}, pstart = 0;
        const long long t_begin = MMA_CLK();
        uint32_t hcnt[2] = {0, 0};                // hidden tiles staged per stream (= MMAs the M warps will have done before reading it)
        // ---- prologue: h0 of the first tiles, the first slabs ----
        load_head(pos[0].head);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            load_h0(hreg[s], pos[s]);
            stage_h(s, 0, hreg[s]);
            if (pos[s].vrows > 0) hcnt[s] = 1;
        }
        p_warps_sync();
        if (ptid == 0) {
            if (pos[0].vrows > 0) mbar_arrive(&sm.st[0].h_ready);
            if (pos[1].vrows > 0) mbar_arrive(&sm.st[1].h_ready);
            for (int n = 0; n < PF_NS && n < nsteps; ++n) { issue_load(0, n); if (pf == 2) issue_load(1, n); }
        }

        for (int n = 0; n < nsteps; ++n) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s == 1 && pf == 1) break;
                FwdStreamSmem& st = sm.st[s];
                const FwdPos pp = pos[s];
                FwdPos nx = pp;
                advance(nx, s);
                const bool last_step = pp.t == T - 1;
                const bool next_active = nx.vrows > 0;
                float2 h0n[2];
                h0n[0] = h0n[1] = make_float2(0.f, 0.f);
                if (last_step && next_active) load_h0(h0n, nx);
                if (pp.vrows > 0) {
                    if (pp.head != cur_head) {    // (uniform across the P warps) new head: biases, and its b_ih copy before anyone reads it
                        load_head(pp.head);
                        p_warps_sync();
                    }
                    const long long pc0 = MMA_CLK();
                    mbar_wait(&st.slab_full[n % PF_NS], (uint32_t)(n / PF_NS) & 1u);
                    const long long pc1 = MMA_CLK();
                    mbar_wait(&st.acc_ready[w], cnt[s] & 1u);
                    const long long pc2 = MMA_CLK();
#ifdef CRVAE_MMA_TIMING
                    dbg[0] += pc1 - pc0; dbg[1] += pc2 - pc1; dbg[4] += 1; pstart = pc2;
#endif
                    ++cnt[s];
                    float* slab = st.slab[n % PF_NS];
                    const bool from_slab = pp.t >= a.t_skip;
                    const float* bias = sm.bih[pp.head & 1];
                    float acc[3][4];
#pragma unroll
                    for (int gate = 0; gate < 3; ++gate) {
                        const float4 v = *reinterpret_cast<const float4*>(&st.accb[w][gate][lane * 4]);
                        acc[gate][0] = v.x; acc[gate][1] = v.y; acc[gate][2] = v.z; acc[gate][3] = v.w;
                    }
                    float2 hn[2];
                    float ps[2];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float2 gi_r = *reinterpret_cast<const float2*>(from_slab ? slab + o_r[i] : bias + ucol);
                        const float2 gi_z = *reinterpret_cast<const float2*>(from_slab ? slab + o_z[i] : bias + MH + ucol);
                        const float2 gi_n = *reinterpret_cast<const float2*>(from_slab ? slab + o_n[i] : bias + 2 * MH + ucol);
                        const float gir[2] = {gi_r.x, gi_r.y}, giz[2] = {gi_z.x, gi_z.y}, gnn[2] = {gi_n.x, gi_n.y};
                        const float hold[2] = {hreg[s][i].x, hreg[s][i].y};
                        float rr[2], zz[2], nn[2], gn[2], hv[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {         // h' = (h - n) * z + n in ATen's operation order
                            const float ghr = acc[0][2 * i + e] + bhh[0][e];
                            const float ghz = acc[1][2 * i + e] + bhh[1][e];
                            gn[e] = acc[2][2 * i + e] + bhh[2][e];
#ifdef CRVAE_WS_NOMATH
                            rr[e] = gir[e] + ghr; zz[e] = giz[e] + ghz; nn[e] = gnn[e] + gn[e]; hv[e] = hold[e] * 0.5f + nn[e] * 1e-3f;
#else
                            rr[e] = sigmoidf_fast(gir[e] + ghr);
                            zz[e] = sigmoidf_fast(giz[e] + ghz);
                            nn[e] = tanhf_fast(__fadd_rn(gnn[e], __fmul_rn(rr[e], gn[e])));
                            hv[e] = __fadd_rn(__fmul_rn(__fsub_rn(hold[e], nn[e]), zz[e]), nn[e]);
#endif
                        }
                        hn[i] = make_float2(hv[0], hv[1]);
                        ps[i] = fmaf(hv[1], wl[1], hv[0] * wl[0]);
                        *reinterpret_cast<float2*>(slab + o_r[i]) = make_float2(rr[0], rr[1]);
                        *reinterpret_cast<float2*>(slab + o_z[i]) = make_float2(zz[0], zz[1]);
                        *reinterpret_cast<float2*>(slab + o_n[i]) = make_float2(nn[0], nn[1]);
                        *reinterpret_cast<float2*>(&st.hst[n & 1][o_h[i]]) = hn[i];
                        *reinterpret_cast<float2*>(&st.gst[n & 1][o_h[i]]) = make_float2(gn[0], gn[1]);
                    }
                    if (last_step) { hn[0] = h0n[0]; hn[1] = h0n[1]; }
                    if (next_active) { stage_h(s, hcnt[s] & 1, hn); ++hcnt[s]; }
                    hreg[s][0] = hn[0]; hreg[s][1] = hn[1];
                    if (has_lin) {
#pragma unroll
                        for (int i = 0; i < 2; ++i) {
                            float pv = ps[i];
                            pv += __shfl_xor_sync(0xffffffffu, pv, 1);
                            pv += __shfl_xor_sync(0xffffffffu, pv, 2);
                            if (q == 0) st.predp[n & 1][g + 8 * i][w] = pv;
                        }
                    }
                    fence_proxy_async_smem();     // staged tiles -> visible to the TMA engine
                } else if (last_step && next_active) {        // idle stream that wakes up: stage the tile's h0
                    stage_h(s, hcnt[s] & 1, h0n);
                    ++hcnt[s];
                    hreg[s][0] = h0n[0]; hreg[s][1] = h0n[1];
                }
#ifdef CRVAE_MMA_TIMING
                const long long pc3 = MMA_CLK();
#endif
                p_warps_sync();
#ifdef CRVAE_MMA_TIMING
                const long long pc4 = MMA_CLK();
                dbg[2] += pc3 - pstart; dbg[3] += pc4 - pc3;
#endif
                if (ptid == 0) {
                    if (next_active) mbar_arrive(&st.h_ready);          // M_s(n+1) may start
                    plumb(s, n, pp);
                }
#ifdef CRVAE_MMA_TIMING
                dbg[5] += MMA_CLK() - pc4;
#endif
                if (warp == 9 && has_lin && lane < pp.vrows) {
                    const float4 p0 = *reinterpret_cast<const float4*>(&st.predp[n & 1][lane][0]);
                    const float4 p1 = *reinterpret_cast<const float4*>(&st.predp[n & 1][lane][4]);
                    a.pred[((long long)pp.head * T + pp.t) * B + pp.tile * M_ROWS + lane] =
                        (((p0.x + p0.y) + (p0.z + p0.w)) + ((p1.x + p1.y) + (p1.z + p1.w))) + blin;
                }
                pos[s] = nx;
            }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
#ifdef CRVAE_MMA_TIMING
        if (blockIdx.x == 0 && lane == 0) { dbg[7] = MMA_CLK() - t_begin; for (int i = 0; i < 8; ++i) g_mma_dbg[warp * 8 + i] = dbg[i]; }
#endif
    }
}

// ------------------------------------------------------------------------------------------------------------
// Warp-specialised BPTT (dW_hh deferred): the mirror image of gru_fwd_mma_ws_kernel, with the product split over K.
//   dh_{t-1} = dh_t * z + dgh . W_hh,   dgh = [da_r | da_z | da_n*r]  (16 x 192),  W_hh [192 x 64]
//   * warps 0-7, "M": warp w owns the K-SLICE of its 8 hidden units (3 gates x 8 = 24 of the 192 reduction indices) for ALL 64
//     outputs: W_hh rows of the slice as B fragments (3 k-steps x 8 n-tiles, 96 registers); its A operand is exactly what
//     the partner thread (same lane) of P warp w computes -- handed over pre-split into tf32 hi | lo as six 128-bit words,
//     so an M op is 6 LDS + 72 HMMA (8 independent accumulator chains) + 8 STS; the 16 x 64 partial product goes to shared
//     memory;
//   * warps 8-15, "P": cell backward at 4 elements per thread; dh_{t-1} = dh_t*z + the 8 partials summed in fixed order
//     (deterministic); staging of dgi (in place over r|z|n), dgh_n (in place over gh_n) and of the next operand; the
//     column sums (db_ih, db_hh, dw_lin, db_lin: 11 register accumulators per tile, reduced in fixed order); TMA plumbing.
//   (A first version split N instead -- every M warp reading the whole 12 KB operand and splitting it itself: 24 LDS + 288
//   ALU + 72 HMMA per op, M-bound at 3.8 us per step.)
// Per step a stream's inputs -- r|z|n (12 KB), gh_n, h_{t-1} (and dhs) tiles -- arrive by TMA into a ring on one mbarrier; dgi /
// dgh_n leave by TMA.  Two tiles of one head per CTA (streams a, b; 2-deep rings) once there are more tiles than SMs, else
// one stream with 3-deep rings.
// ------------------------------------------------------------------------------------------------------------
template <int NS>
struct __align__(1024) BwdStreamSmem {
    float slab[NS][M_ROWS * MG];           // r|z|n -> dgi                      [row][6 x 128 B] swizzled
    float gn[NS][M_ROWS * MH];             // gh_n -> dgh_n                     [row][2 x 128 B] swizzled
    float hp[NS][M_ROWS * MH];             // h_{t-1}
    float de[NS][M_ROWS * MH];             // dhs (per-step gradient into h_t), when present
    float dghb[8][3][2][32 * 4];           // A operand: [warp][gate][hi | lo][lane] x 4 words (the lane's MMA fragment)
    float accb[8][8][32 * 4];              // partial products: [M warp][n-tile][lane] x float4 (C-fragment order)
    uint64_t in_full[NS];
    uint64_t dgh_ready;
    uint64_t acc_ready;
};

struct BwdTmaps { CUtensorMap g, n, h, z, d; };

template <bool HAS_DHS, int NS>
__global__ void __launch_bounds__(WS_THREADS, 1)
gru_bwd_mma_ws_kernel(const __grid_constant__ BwdTmaps tm, GruMmaBwdArgs a, int pf, int npph, int h0_per_head) {
    using namespace umma;
    extern __shared__ __align__(1024) uint8_t mma_smem_raw[];
    BwdStreamSmem<NS>* sms = reinterpret_cast<BwdStreamSmem<NS>*>(mma_smem_raw);      // pf streams

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, q = lane & 3;
    const int w = warp & 7;
    const bool is_m = warp < 8;
    const int total = a.P * npph;
    const int first = (int)((long long)blockIdx.x * total / gridDim.x);
    const int last = (int)((long long)(blockIdx.x + 1) * total / gridDim.x);
    if (first >= last) return;
    const int ucol = 8 * w + 2 * q;
    const bool has_lin = a.w_lin != nullptr;
    const int T = a.T, B = a.B;
    const int nsteps = (last - first) * T;
    constexpr uint32_t IN_BYTES = (uint32_t)(M_ROWS * (MG + 2 * MH + (HAS_DHS ? MH : 0)) * 4);

    // step n of stream s: pair j = n / T, time t = T - 1 - n % T (the BPTT walks backwards)
    auto pos_of_pair = [&](int j, int s) {
        FwdPos p;
        p.j = j; p.t = T - 1;
        const int pj = first + j;
        p.head = pj / npph;
        p.tile = (pj - p.head * npph) * pf + s;
        const int b0 = p.tile * M_ROWS;
        p.vrows = (pj < last && s < pf && b0 < B) ? min(M_ROWS, B - b0) : 0;
        return p;
    };
    auto advance = [&](FwdPos& p, int s) {
        if (--p.t < 0) p = pos_of_pair(p.j + 1, s);
    };

    if (tid == 32) { prefetch_tmap(&tm.g); prefetch_tmap(&tm.n); prefetch_tmap(&tm.h); prefetch_tmap(&tm.z); if (HAS_DHS) prefetch_tmap(&tm.d); }
    if (tid == 0) {
        for (int s = 0; s < pf; ++s) {
#pragma unroll
            for (int i = 0; i < NS; ++i) mbar_init(&sms[s].in_full[i], 1);
            mbar_init(&sms[s].dgh_ready, 1);
            mbar_init(&sms[s].acc_ready, 8);
        }
        fence_barrier_init();
    }
    __syncthreads();

    if (is_m) {
        // =============================================================== M warps ===============================================================
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(160));       // W_hh 96 + accumulators 32 + operand 24 (the P warps give up 32 each)
        uint32_t bhi[3][8][2], blo[3][8][2];
        int cur_head = -1;
        auto load_head = [&](int head) {          // k-step `gate`: k-slots q / q+4 <-> W_hh rows gate*64 + 8w + 2q (+1); n-tile j column g <-> output unit 8j + g
            const float* __restrict__ W = a.w_hh + (long long)head * MG * MH;
#pragma unroll
            for (int gate = 0; gate < 3; ++gate) {
                const float* r0 = W + (long long)(gate * MH + 8 * w + 2 * q) * MH + g;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    split_tf32_fast(__ldg(r0 + 8 * j), bhi[gate][j][0], blo[gate][j][0]);
                    split_tf32_fast(__ldg(r0 + MH + 8 * j), bhi[gate][j][1], blo[gate][j][1]);
                }
            }
            cur_head = head;
        };
        FwdPos pos[2] = {pos_of_pair(0, 0), pos_of_pair(0, 1)};
        uint32_t cnt[2] = {0, 0};
        for (int n = 0; n < nsteps; ++n) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s == 1 && pf == 1) break;
                if (pos[s].vrows > 0) {
                    if (pos[s].head != cur_head) load_head(pos[s].head);
                    BwdStreamSmem<NS>& st = sms[s];
                    mbar_wait(&st.dgh_ready, cnt[s] & 1u);
                    uint4 ah[3], al[3];
#pragma unroll
                    for (int gate = 0; gate < 3; ++gate) {
                        ah[gate] = *reinterpret_cast<const uint4*>(&st.dghb[w][gate][0][lane * 4]);
                        al[gate] = *reinterpret_cast<const uint4*>(&st.dghb[w][gate][1][lane * 4]);
                    }
                    float acc[8][4];
#pragma unroll
                    for (int j = 0; j < 8; ++j)
#pragma unroll
                        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
#pragma unroll
                    for (int gate = 0; gate < 3; ++gate) {
                        const uint32_t ahi[4] = {ah[gate].x, ah[gate].y, ah[gate].z, ah[gate].w};
                        const uint32_t alo[4] = {al[gate].x, al[gate].y, al[gate].z, al[gate].w};
#pragma unroll
                        for (int j = 0; j < 8; ++j) mma_tf32(acc[j], alo, bhi[gate][j]);
#pragma unroll
                        for (int j = 0; j < 8; ++j) mma_tf32(acc[j], ahi, blo[gate][j]);
#pragma unroll
                        for (int j = 0; j < 8; ++j) mma_tf32(acc[j], ahi, bhi[gate][j]);
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(&st.accb[w][j][lane * 4]) = make_float4(acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&st.acc_ready);
                    ++cnt[s];
                }
                advance(pos[s], s);
            }
        }
    } else {
        // =============================================================== P warps ===============================================================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(96));
        const int ptid = tid - 256;
        float wl[2] = {0.f, 0.f};
        int cur_head = -1;
        auto load_head = [&](int head) {
            if (has_lin) {
                const float2 w2 = __ldg(reinterpret_cast<const float2*>(a.w_lin + (long long)head * MH + ucol));
                wl[0] = w2.x; wl[1] = w2.y;
            }
            cur_head = head;
        };
        // ---- plumbing (thread 256) ----
        auto issue_load = [&](int s, int n) {         // inputs of step n -> ring slot n % NS
            const int j = n / T;
            FwdPos p = pos_of_pair(j, s);
            p.t = T - 1 - (n - j * T);
            const int slot = n % NS;
            BwdStreamSmem<NS>& st = sms[s];
            mbar_arrive_expect_tx(&st.in_full[slot], p.vrows > 0 ? IN_BYTES : 0u);
            if (p.vrows > 0) {
                const int c2 = p.tile * M_ROWS, c3 = p.head * T + p.t;
                tma_load_4d(st.slab[slot], &tm.g, &st.in_full[slot], 0, 0, c2, c3);
                tma_load_4d(st.gn[slot], &tm.n, &st.in_full[slot], 0, 0, c2, c3);
                if (p.t > 0) tma_load_4d(st.hp[slot], &tm.h, &st.in_full[slot], 0, 0, c2, c3 - 1);
                else         tma_load_4d(st.hp[slot], &tm.z, &st.in_full[slot], 0, 0, c2, h0_per_head ? p.head : 0);
                if (HAS_DHS) tma_load_4d(st.de[slot], &tm.d, &st.in_full[slot], 0, 0, c2, c3);
            }
        };
        auto plumb = [&](int s, int n, const FwdPos& p) {    // all P warps have staged step n of stream s
            if (pf == 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            else         asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            if (p.vrows > 0) {
                const int c2 = p.tile * M_ROWS, c3 = p.head * T + p.t;
                tma_store_4d(&tm.g, sms[s].slab[n % NS], 0, 0, c2, c3);
                tma_store_4d(&tm.n, sms[s].gn[n % NS], 0, 0, c2, c3);
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (n >= 1 && n - 1 + NS < nsteps) issue_load(s, n - 1 + NS);
        };
        auto load_dp = [&](float (&dp)[2], const FwdPos& p) {
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = g + 8 * i;
                dp[i] = (has_lin && r < p.vrows) ? __ldg(a.dpred + ((long long)p.head * T + p.t) * B + p.tile * M_ROWS + r) : 0.f;
            }
        };
        auto load_tile_start = [&](float2 (&dhl)[2], float2 (&hl)[2], const FwdPos& p) {     // dh_last and h_{T-1} of a tile
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const int r = g + 8 * i, b = p.tile * M_ROWS + r;
                dhl[i] = make_float2(0.f, 0.f); hl[i] = make_float2(0.f, 0.f);
                if (r < p.vrows) {
                    if (a.dh_last) dhl[i] = __ldg(reinterpret_cast<const float2*>(a.dh_last + ((long long)p.head * B + b) * MH + ucol));
                    if (has_lin) hl[i] = __ldg(reinterpret_cast<const float2*>(a.hs + (((long long)p.head * T + (T - 1)) * B + b) * MH + ucol));
                }
            }
        };

        int o_r[2], o_z[2], o_n[2], o_h[2];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            o_r[i] = sw_idx(6, g + 8 * i, ucol); o_z[i] = sw_idx(6, g + 8 * i, MH + ucol); o_n[i] = sw_idx(6, g + 8 * i, 2 * MH + ucol);
            o_h[i] = sw_idx(2, g + 8 * i, ucol);
        }
        FwdPos pos[2] = {pos_of_pair(0, 0), pos_of_pair(0, 1)};
        // per-stream state
        float dh[2][2][2], dhz[2][2][2];          // [stream][row][unit]
        float2 hcur[2][2];                        // h_t of the step being processed (dw_lin += dpred[t] * h_t)
        float dpn[2][2];                          // dpred of the stream's next step (prefetched)
        float sums[2][11];                        // db_ih r,z,n (x2 units) | db_hh n (x2) | dw_lin (x2) | db_lin
        bool pending[2] = {false, false};         // an MMA result of this stream is outstanding
        FwdPos ppos[2] = {pos[0], pos[1]};        // position of the outstanding step
        uint32_t cnt[2] = {0, 0};
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            load_dp(dpn[s], pos[s]);
#pragma unroll
            for (int k = 0; k < 11; ++k) sums[s][k] = 0.f;
#pragma unroll
            for (int i = 0; i < 2; ++i) { dh[s][i][0] = dh[s][i][1] = 0.f; dhz[s][i][0] = dhz[s][i][1] = 0.f; hcur[s][i] = make_float2(0.f, 0.f); }
        }
        load_head(pos[0].head);
        if (ptid == 0) {
            for (int n = 0; n < NS && n < nsteps; ++n) { issue_load(0, n); if (pf == 2) issue_load(1, n); }
        }
        // finish the outstanding step of stream s: dh_{t-1} = dh_t * z + dgh . W_hh; at a tile end write dh0 and the tile's column sums
        auto finish = [&](int s) {
            BwdStreamSmem<NS>& st = sms[s];
            mbar_wait(&st.acc_ready, cnt[s] & 1u);
            ++cnt[s];
            float4 sum = *reinterpret_cast<const float4*>(&st.accb[0][w][lane * 4]);
#pragma unroll
            for (int v = 1; v < 8; ++v) {         // the 8 K-slices, fixed order
                const float4 x = *reinterpret_cast<const float4*>(&st.accb[v][w][lane * 4]);
                sum.x = __fadd_rn(sum.x, x.x); sum.y = __fadd_rn(sum.y, x.y); sum.z = __fadd_rn(sum.z, x.z); sum.w = __fadd_rn(sum.w, x.w);
            }
            dh[s][0][0] = __fadd_rn(dhz[s][0][0], sum.x); dh[s][0][1] = __fadd_rn(dhz[s][0][1], sum.y);
            dh[s][1][0] = __fadd_rn(dhz[s][1][0], sum.z); dh[s][1][1] = __fadd_rn(dhz[s][1][1], sum.w);
            pending[s] = false;
            const FwdPos& p = ppos[s];
            if (p.t == 0) {                       // tile end
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const int r = g + 8 * i;
                    if (r < p.vrows) stg2(a.dh0 + ((long long)p.head * B + p.tile * M_ROWS + r) * MH + ucol, dh[s][i][0], dh[s][i][1]);
                }
                float vv[11];
#pragma unroll
                for (int k = 0; k < 11; ++k) {
                    vv[k] = sums[s][k];
                    vv[k] += __shfl_xor_sync(0xffffffffu, vv[k], 4);
                    vv[k] += __shfl_xor_sync(0xffffffffu, vv[k], 8);
                    vv[k] += __shfl_xor_sync(0xffffffffu, vv[k], 16);
                    sums[s][k] = 0.f;
                }
                if (g == 0) {
                    float* ws = a.ws + ((long long)p.head * a.ntiles + p.tile) * MWS_TILE;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const int u = ucol + e;
                        ws[MWS_DBIH + u] = vv[e];            ws[MWS_DBHH + u] = vv[e];                  // r: dgh == dgi
                        ws[MWS_DBIH + MH + u] = vv[2 + e];   ws[MWS_DBHH + MH + u] = vv[2 + e];         // z
                        ws[MWS_DBIH + 2 * MH + u] = vv[4 + e];                                          // n: db_ih
                        ws[MWS_DBHH + 2 * MH + u] = vv[6 + e];                                          // n: db_hh (dgh_n)
                        ws[MWS_DWLIN + u] = vv[8 + e];
                    }
                    if (w == 0 && q == 0) ws[MWS_DBLIN] = vv[10];
                }
            }
        };

        for (int n = 0; n < nsteps; ++n) {
#pragma unroll
            for (int s = 0; s < 2; ++s) {
                if (s == 1 && pf == 1) break;
                BwdStreamSmem<NS>& st = sms[s];
                const FwdPos pp = pos[s];
                FwdPos nx = pp;
                advance(nx, s);
                if (pending[s]) finish(s);
                if (pp.vrows > 0) {
                    if (pp.head != cur_head) load_head(pp.head);
                    if (pp.t == T - 1) {          // tile start: dh_last, h_{T-1}
                        float2 dhl[2], hl[2];
                        load_tile_start(dhl, hl, pp);
#pragma unroll
                        for (int i = 0; i < 2; ++i) { dh[s][i][0] = dhl[i].x; dh[s][i][1] = dhl[i].y; hcur[s][i] = hl[i]; }
                    }
                    const float dp[2] = {dpn[s][0], dpn[s][1]};
                    // prefetch for the stream's next step
                    if (nx.vrows > 0) load_dp(dpn[s], nx);
                    const int slot = n % NS;
                    mbar_wait(&st.in_full[slot], (uint32_t)(n / NS) & 1u);
                    float* slab = st.slab[slot];
                    float* gns = st.gn[slot];
                    const float* hps = st.hp[slot];
                    float fr[4], fz[4], fn_[4], fa[4];         // da_r, da_z, da_n*r (the MMA operand) and da_n of the 4 elements
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const float2 r2 = *reinterpret_cast<const float2*>(slab + o_r[i]), z2 = *reinterpret_cast<const float2*>(slab + o_z[i]),
                                     n2 = *reinterpret_cast<const float2*>(slab + o_n[i]);
                        const float2 g2 = *reinterpret_cast<const float2*>(gns + o_h[i]);
                        const float2 p2 = *reinterpret_cast<const float2*>(hps + o_h[i]);
                        float2 e2 = make_float2(0.f, 0.f);
                        if (HAS_DHS) e2 = *reinterpret_cast<const float2*>(st.de[slot] + o_h[i]);
                        const float r_[2] = {r2.x, r2.y}, z_[2] = {z2.x, z2.y}, n_[2] = {n2.x, n2.y}, gn_[2] = {g2.x, g2.y}, hp_[2] = {p2.x, p2.y},
                                    de_[2] = {e2.x, e2.y};
                        const float hc_[2] = {hcur[s][i].x, hcur[s][i].y};
                        float dar[2], daz[2], dan[2], dgn[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) {
                            const float d = dh[s][i][e] + dp[i] * wl[e] + de_[e];      // total dL/dh_t (same order as gru_bwd_kernel)
                            const float dn = d * (1.f - z_[e]);
                            const float dz = d * (hp_[e] - n_[e]);
                            dan[e] = dn * (1.f - n_[e] * n_[e]);
                            const float dr = dan[e] * gn_[e];
                            dar[e] = dr * r_[e] * (1.f - r_[e]);
                            daz[e] = dz * z_[e] * (1.f - z_[e]);
                            dgn[e] = dan[e] * r_[e];
                            dhz[s][i][e] = d * z_[e];
                            sums[s][e] += dar[e]; sums[s][2 + e] += daz[e]; sums[s][4 + e] += dan[e]; sums[s][6 + e] += dgn[e];
                            sums[s][8 + e] = fmaf(dp[i], hc_[e], sums[s][8 + e]);
                        }
                        if (w == 0 && q == 0) sums[s][10] += dp[i];
                        hcur[s][i] = p2;
                        fr[i] = dar[0]; fr[2 + i] = dar[1];
                        fz[i] = daz[0]; fz[2 + i] = daz[1];
                        fn_[i] = dgn[0]; fn_[2 + i] = dgn[1];
                        fa[i] = dan[0]; fa[2 + i] = dan[1];
                    }
                    // the MMA operand first -- the partner lane's A fragments {x[g][u], x[g+8][u], x[g][u+1], x[g+8][u+1]} of the three gate
                    // blocks, split hi | lo -- it is all the M warps wait for; dgi / dgh_n are staged after they are released (measured:
                    // 193.5 -> 187.4 us at P = 100; the same reordering in the forward, whose gate-math warps are the bottleneck, costs
                    // more in the second barrier than it gains: 160 -> 172 us)
                    {
                        uint32_t hi[4], lo[4];
#pragma unroll
                        for (int e = 0; e < 4; ++e) split_tf32_fast(fr[e], hi[e], lo[e]);
                        *reinterpret_cast<uint4*>(&st.dghb[w][0][0][lane * 4]) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(&st.dghb[w][0][1][lane * 4]) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) split_tf32_fast(fz[e], hi[e], lo[e]);
                        *reinterpret_cast<uint4*>(&st.dghb[w][1][0][lane * 4]) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(&st.dghb[w][1][1][lane * 4]) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
#pragma unroll
                        for (int e = 0; e < 4; ++e) split_tf32_fast(fn_[e], hi[e], lo[e]);
                        *reinterpret_cast<uint4*>(&st.dghb[w][2][0][lane * 4]) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
                        *reinterpret_cast<uint4*>(&st.dghb[w][2][1][lane * 4]) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
                    }
                    pending[s] = true;
                    ppos[s] = pp;
                    p_warps_sync();               // every P warp has read the previous partials and staged its operand: M may overwrite accb
                    if (ptid == 0) mbar_arrive(&st.dgh_ready);
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        *reinterpret_cast<float2*>(slab + o_r[i]) = make_float2(fr[i], fr[2 + i]);
                        *reinterpret_cast<float2*>(slab + o_z[i]) = make_float2(fz[i], fz[2 + i]);
                        *reinterpret_cast<float2*>(slab + o_n[i]) = make_float2(fa[i], fa[2 + i]);
                        *reinterpret_cast<float2*>(gns + o_h[i]) = make_float2(fn_[i], fn_[2 + i]);
                    }
                    fence_proxy_async_smem();
                } else {
                    p_warps_sync();
                }
                p_warps_sync();                   // every P warp has staged dgi / dgh_n of step n
                if (ptid == 0) plumb(s, n, pp);
                pos[s] = nx;
            }
        }
#pragma unroll
        for (int s = 0; s < 2; ++s)
            if (pending[s]) finish(s);
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

static int mma_num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace crvae

using namespace crvae;

#ifdef CRVAE_MMA_TIMING
extern "C" int crvae_debug_mma_timing(long long* out) { return (int)cudaMemcpyFromSymbol(out, g_mma_dbg, sizeof(long long) * 128); }
#endif

// Tensor-core (warp-level MMA, 3xTF32) form of crvae_gru_fwd_ll: same arguments and buffers; results agree with the exact
// kernels to ~1e-6 relative (like crvae_gru_fwd_tc).
extern "C" int crvae_gru_fwd_mma(float* gates, const float* b_ih, const float* w_hh, const float* b_hh,
                                 const float* h0, int64_t h0_head_stride, const float* w_lin, const float* b_lin,
                                 float* hs, float* ghn, float* pred, int P, int T, int B, int t_skip, void* stream) {
    CRVAE_REQUIRE(gates && b_ih && w_hh && b_hh && h0 && hs && ghn, "null operand");
    CRVAE_REQUIRE((w_lin == nullptr) == (pred == nullptr), "w_lin and pred go together");
    CRVAE_REQUIRE(w_lin == nullptr || b_lin != nullptr, "b_lin missing");
    CRVAE_REQUIRE(P >= 0 && T > 0 && B > 0 && t_skip >= 0 && t_skip <= T, "bad size");
    CRVAE_REQUIRE(aligned16(gates) && aligned16(hs) && aligned16(ghn) && aligned16(h0) && aligned16(w_hh), "16-byte alignment");
    CRVAE_REQUIRE(((uintptr_t)b_ih & 7u) == 0 && ((uintptr_t)b_hh & 7u) == 0 && (w_lin == nullptr || ((uintptr_t)w_lin & 7u) == 0),
                  "8-byte alignment of the bias / output-weight rows");
    CRVAE_REQUIRE(h0_head_stride % 4 == 0, "h0 head stride must keep 16-byte alignment");
    if (P == 0) return 0;
    const int ntiles = (B + M_ROWS - 1) / M_ROWS;
    GruMmaFwdArgs a{gates, b_ih, w_hh, b_hh, h0, (long long)h0_head_stride, w_lin, b_lin, hs, ghn, pred, P, T, B, t_skip, ntiles};
    const long long total = (long long)P * ntiles;
    static const int variant = [] { const char* e = getenv("CRVAE_MMA_VARIANT"); return e ? atoi(e) : 1; }();
    if (variant == 0) {
        const int grid = (int)(total < mma_num_sms() ? total : mma_num_sms());
        gru_fwd_mma_kernel<<<grid, M_THREADS, 0, (cudaStream_t)stream>>>(a);
        return check_launch("gru_fwd_mma_kernel");
    }
    // warp-specialised kernel: pairs of tiles once there are more tiles than SMs, single tiles (stream b idle) below that
    const int pf = total > mma_num_sms() ? 2 : 1;
    const int npph = (ntiles + pf - 1) / pf;
    const long long units = (long long)P * npph;
    const int grid = (int)(units < mma_num_sms() ? units : mma_num_sms());
    const int smem = (int)sizeof(FwdPipeSmem);
    static bool attr_done = false;
    if (!attr_done) {
        cudaError_t e = cudaFuncSetAttribute(gru_fwd_mma_ws_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(gru_fwd_mma_ws_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) { set_error("gru_fwd_mma_ws smem attr (%d B): %s", smem, cudaGetErrorString(e)); return (int)e; }
        attr_done = true;
    }
    CUtensorMap tG, tH, tN;
    {
        int rc;
        const uint64_t PT = (uint64_t)P * T;
        const uint64_t dg[4] = {32, 6, (uint64_t)B, PT}, sg[3] = {128, (uint64_t)MG * 4, (uint64_t)B * MG * 4};
        const uint32_t bg[4] = {32, 6, M_ROWS, 1};
        if ((rc = make_tmap_generic(&tG, gates, 4, dg, sg, bg, false))) return rc;
        const uint64_t dh[4] = {32, 2, (uint64_t)B, PT}, sh[3] = {128, (uint64_t)MH * 4, (uint64_t)B * MH * 4};
        const uint32_t bh[4] = {32, 2, M_ROWS, 1};
        if ((rc = make_tmap_generic(&tH, hs, 4, dh, sh, bh, false))) return rc;
        if ((rc = make_tmap_generic(&tN, ghn, 4, dh, sh, bh, false))) return rc;
    }
    // operand format of the gate product: fp16 hi | lo (m16n8k16, 36 MMAs per warp-step) unless CRVAE_MMA_F16=0 (tf32 hi | lo, 72)
    static const int f16 = [] { const char* e = getenv("CRVAE_MMA_F16"); return e ? atoi(e) : 1; }();
    if (f16) gru_fwd_mma_ws_kernel<true><<<grid, WS_THREADS, smem, (cudaStream_t)stream>>>(tG, tH, tN, a, pf, npph);
    else     gru_fwd_mma_ws_kernel<false><<<grid, WS_THREADS, smem, (cudaStream_t)stream>>>(tG, tH, tN, a, pf, npph);
    return check_launch("gru_fwd_mma_ws_kernel");
}

// Tensor-core (warp-level MMA, 3xTF32) form of crvae_gru_bwd_ll: gates <- dgi, ghn <- dgh_n in place; dw_hh is produced
// afterwards by crvae_gru_dwhh_tc.  `workspace` >= crvae_gru_bwd_workspace(P, B).
extern "C" int crvae_gru_bwd_mma(float* gates, float* ghn, const float* hs, const float* h0, int64_t h0_head_stride,
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
    CRVAE_REQUIRE(dh_last == nullptr || aligned16(dh_last), "16-byte alignment");
    CRVAE_REQUIRE(w_lin == nullptr || ((uintptr_t)w_lin & 7u) == 0, "8-byte alignment of the output-weight rows");
    CRVAE_REQUIRE(h0_head_stride % 4 == 0, "h0 head stride must keep 16-byte alignment");
    if (P == 0) return 0;
    const int ntiles = (B + M_ROWS - 1) / M_ROWS;
    GruMmaBwdArgs a{gates, ghn, hs, h0, (long long)h0_head_stride, w_hh, w_lin, dpred, dh_last, dhs, dh0, (float*)workspace, P, T, B, ntiles};
    cudaStream_t st = (cudaStream_t)stream;
    const long long total = (long long)P * ntiles;
    const int grid = (int)(total < mma_num_sms() ? total : mma_num_sms());
    static const int variant = [] { const char* e = getenv("CRVAE_MMA_BWD_VARIANT"); return e ? atoi(e) : 1; }();
    int rc;
    if (variant == 0) {
        if (dhs) gru_bwd_mma_kernel<true><<<grid, M_THREADS, 0, st>>>(a);
        else     gru_bwd_mma_kernel<false><<<grid, M_THREADS, 0, st>>>(a);
        rc = check_launch("gru_bwd_mma_kernel");
    } else {
        const int pf = total > mma_num_sms() ? 2 : 1;
        const int npph = (ntiles + pf - 1) / pf;
        const long long units = (long long)P * npph;
        const int wgrid = (int)(units < mma_num_sms() ? units : mma_num_sms());
        BwdTmaps tm;
        const uint64_t PT = (uint64_t)P * T;
        const uint64_t dg[4] = {32, 6, (uint64_t)B, PT}, sg[3] = {128, (uint64_t)MG * 4, (uint64_t)B * MG * 4};
        const uint32_t bg[4] = {32, 6, M_ROWS, 1};
        if ((rc = make_tmap_generic(&tm.g, gates, 4, dg, sg, bg, false))) return rc;
        const uint64_t dh[4] = {32, 2, (uint64_t)B, PT}, sh[3] = {128, (uint64_t)MH * 4, (uint64_t)B * MH * 4};
        const uint32_t bh[4] = {32, 2, M_ROWS, 1};
        if ((rc = make_tmap_generic(&tm.n, ghn, 4, dh, sh, bh, false))) return rc;
        if ((rc = make_tmap_generic(&tm.h, hs, 4, dh, sh, bh, false))) return rc;
        if ((rc = make_tmap_generic(&tm.d, dhs ? dhs : hs, 4, dh, sh, bh, false))) return rc;
        const int per_head = h0_head_stride != 0;
        const uint64_t dz[4] = {32, 2, (uint64_t)B, (uint64_t)(per_head ? P : 1)};
        const uint64_t sz[3] = {128, (uint64_t)MH * 4, (uint64_t)(per_head ? h0_head_stride : (int64_t)B * MH) * 4};
        if ((rc = make_tmap_generic(&tm.z, h0, 4, dz, sz, bh, false))) return rc;
        // one stream: 3-deep rings; two streams: 2-deep rings (shared memory)
        const int smem = pf == 1 ? (int)sizeof(BwdStreamSmem<3>) : 2 * (int)sizeof(BwdStreamSmem<2>);
        static bool attr_done = false;
        if (!attr_done) {
            const int s3 = (int)sizeof(BwdStreamSmem<3>), s2 = 2 * (int)sizeof(BwdStreamSmem<2>);
            cudaError_t e = cudaFuncSetAttribute(gru_bwd_mma_ws_kernel<true, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, s3);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(gru_bwd_mma_ws_kernel<false, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, s3);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(gru_bwd_mma_ws_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2);
            if (e == cudaSuccess) e = cudaFuncSetAttribute(gru_bwd_mma_ws_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, s2);
            if (e != cudaSuccess) { set_error("gru_bwd_mma_ws smem attr (%d / %d B): %s", s3, s2, cudaGetErrorString(e)); return (int)e; }
            attr_done = true;
        }
        if (pf == 1) {
            if (dhs) gru_bwd_mma_ws_kernel<true, 3><<<wgrid, WS_THREADS, smem, st>>>(tm, a, pf, npph, per_head);
            else     gru_bwd_mma_ws_kernel<false, 3><<<wgrid, WS_THREADS, smem, st>>>(tm, a, pf, npph, per_head);
        } else {
            if (dhs) gru_bwd_mma_ws_kernel<true, 2><<<wgrid, WS_THREADS, smem, st>>>(tm, a, pf, npph, per_head);
            else     gru_bwd_mma_ws_kernel<false, 2><<<wgrid, WS_THREADS, smem, st>>>(tm, a, pf, npph, per_head);
        }
        rc = check_launch("gru_bwd_mma_ws_kernel");
    }
    if (rc) return rc;
    return launch_gru_bwd_finalize((const float*)workspace, db_hh, db_ih, dw_lin, db_lin, P, ntiles, st);
}
