// Exact-fp32 (FFMA) batched GEMM used for (a) every small dense product on the path (fc_mu|fc_std,
// encoder projection, their gradients) and (b) the "exact" mode of the multi-head input
// projection and its weight-gradient twin.  Register-tiled, shared-memory staged, register
// prefetch of the next k-tile; split-K with a deterministic second-pass reduction.
//
// Reference arithmetic replaced: ATen addmm/mm under nn.GRU's linear_ih (CRVAE_lorenz96.py:119,
// :208) and nn.Linear (:210-211) and their autograd twins (:497).
#include "common.cuh"

namespace crvae {

struct GemmArgs {
    const float* A;
    const float* B;
    float* C;
    const float* bias;
    const uint8_t* colmask;
    int M, N, K;
    int lda, ldb, ldc;
    long long sA, sB, sC, sBias, sMask;
    int splits;
    int accumulate;
    float* ws;  // split partials [batch][splits][M][N] when splits > 1
};

template <int T>
__device__ __forceinline__ int tile_index(int t, int i, int BMN) {
    // T == 8: two groups of four, half a tile apart (conflict-free float4 smem reads)
    if (T == 8) return (i >> 2) * (BMN / 2) + t * 4 + (i & 3);
    return t * T + i;
}

template <int BM, int BN, int TM, int TN, bool A_KC, bool B_KC>
__global__ void __launch_bounds__(256) gemm_f32_kernel(GemmArgs g) {
    constexpr int BK = 16;
    constexpr int NTX = BN / TN;
    static_assert((BM / TM) * (BN / TN) == 256, "256 threads per CTA");
    static_assert(TM % 4 == 0 && TN % 4 == 0, "float4 micro-tiles");
    __shared__ __align__(16) float As[BK][BM + 4];
    __shared__ __align__(16) float Bs[BK][BN + 4];

    const int tid = threadIdx.x;
    const int tx = tid % NTX, ty = tid / NTX;
    const int batch = blockIdx.z / g.splits, split = blockIdx.z % g.splits;
    const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
    int kchunk = (g.K + g.splits - 1) / g.splits;
    kchunk = (kchunk + BK - 1) / BK * BK;
    const int kbeg = split * kchunk;
    const int kend = min(g.K, kbeg + kchunk);

    const float* __restrict__ A = g.A + (long long)batch * g.sA;
    const float* __restrict__ B = g.B + (long long)batch * g.sB;

    constexpr int A_LD = BM * BK / 256, B_LD = BN * BK / 256;
    float ra[A_LD], rb[B_LD];
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    auto load_tiles = [&](int k0) {
#pragma unroll
        for (int i = 0; i < A_LD; ++i) {
            int e = tid + i * 256;
            int m, k;
            if (A_KC) { m = e / BK; k = e % BK; } else { k = e / BM; m = e % BM; }
            int gm = m0 + m, gk = k0 + k;
            float v = 0.f;
            if (gm < g.M && gk < kend)
                v = A_KC ? __ldg(A + (long long)gm * g.lda + gk) : __ldg(A + (long long)gk * g.lda + gm);
            ra[i] = v;
        }
#pragma unroll
        for (int i = 0; i < B_LD; ++i) {
            int e = tid + i * 256;
            int n, k;
            if (B_KC) { n = e / BK; k = e % BK; } else { k = e / BN; n = e % BN; }
            int gn = n0 + n, gk = k0 + k;
            float v = 0.f;
            if (gn < g.N && gk < kend)
                v = B_KC ? __ldg(B + (long long)gn * g.ldb + gk) : __ldg(B + (long long)gk * g.ldb + gn);
            rb[i] = v;
        }
    };
    auto store_tiles = [&]() {
#pragma unroll
        for (int i = 0; i < A_LD; ++i) {
            int e = tid + i * 256;
            int m, k;
            if (A_KC) { m = e / BK; k = e % BK; } else { k = e / BM; m = e % BM; }
            As[k][m] = ra[i];
        }
#pragma unroll
        for (int i = 0; i < B_LD; ++i) {
            int e = tid + i * 256;
            int n, k;
            if (B_KC) { n = e / BK; k = e % BK; } else { k = e / BN; n = e % BN; }
            Bs[k][n] = rb[i];
        }
    };

    if (kbeg < kend) {
        load_tiles(kbeg);
        for (int k0 = kbeg; k0 < kend; k0 += BK) {
            store_tiles();
            __syncthreads();
            if (k0 + BK < kend) load_tiles(k0 + BK);
#pragma unroll
            for (int kk = 0; kk < BK; ++kk) {
                float a[TM], b[TN];
#pragma unroll
                for (int i = 0; i < TM; i += 4) {
                    float4 v = *reinterpret_cast<const float4*>(&As[kk][tile_index<TM>(ty, i, BM)]);
                    a[i] = v.x; a[i + 1] = v.y; a[i + 2] = v.z; a[i + 3] = v.w;
                }
#pragma unroll
                for (int j = 0; j < TN; j += 4) {
                    float4 v = *reinterpret_cast<const float4*>(&Bs[kk][tile_index<TN>(tx, j, BN)]);
                    b[j] = v.x; b[j + 1] = v.y; b[j + 2] = v.z; b[j + 3] = v.w;
                }
                // packed fp32 FMA (FFMA2, fma.rn.f32x2): Blackwell's full-rate fp32 path; bit-identical per component
#pragma unroll
                for (int i = 0; i < TM; ++i) {
                    const float2 a2 = make_float2(a[i], a[i]);
#pragma unroll
                    for (int j = 0; j < TN; j += 2) {
                        float2 r = __ffma2_rn(a2, make_float2(b[j], b[j + 1]), make_float2(acc[i][j], acc[i][j + 1]));
                        acc[i][j] = r.x; acc[i][j + 1] = r.y;
                    }
                }
            }
            __syncthreads();
        }
    }

    // epilogue
    if (g.splits > 1) {
        float* W = g.ws + ((long long)batch * g.splits + split) * (long long)g.M * g.N;
#pragma unroll
        for (int i = 0; i < TM; ++i) {
            int gm = m0 + tile_index<TM>(ty, i, BM);
            if (gm >= g.M) continue;
#pragma unroll
            for (int j = 0; j < TN; ++j) {
                int gn = n0 + tile_index<TN>(tx, j, BN);
                if (gn < g.N) W[(long long)gm * g.N + gn] = acc[i][j];
            }
        }
        return;
    }
    float* C = g.C + (long long)batch * g.sC;
    const float* bias = g.bias ? g.bias + (long long)batch * g.sBias : nullptr;
    const uint8_t* cm = g.colmask ? g.colmask + (long long)batch * g.sMask : nullptr;
    const bool vec_ok = ((g.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15u) == 0) &&
                        !g.accumulate && !cm;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        int gm = m0 + tile_index<TM>(ty, i, BM);
        if (gm >= g.M) continue;
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
            int gn = n0 + tile_index<TN>(tx, j, BN);
            float v[4] = {acc[i][j], acc[i][j + 1], acc[i][j + 2], acc[i][j + 3]};
            if (vec_ok && gn + 3 < g.N) {
                if (bias) {
                    v[0] += bias[gn]; v[1] += bias[gn + 1]; v[2] += bias[gn + 2]; v[3] += bias[gn + 3];
                }
                *reinterpret_cast<float4*>(C + (long long)gm * g.ldc + gn) = make_float4(v[0], v[1], v[2], v[3]);
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    int n = gn + q;
                    if (n >= g.N) continue;
                    float o = v[q] + (bias ? bias[n] : 0.f);
                    if (cm && !cm[n]) o = 0.f;
                    float* dst = C + (long long)gm * g.ldc + n;
                    *dst = g.accumulate ? (*dst + o) : o;
                }
            }
        }
    }
}

// out[b][m][n] = (accumulate ? out : 0) + bias + sum_s ws[b][s][m][n]   (fixed order: deterministic)
__global__ void gemm_split_reduce_kernel(GemmArgs g) {
    long long total = (long long)g.M * g.N;
    int batch = blockIdx.y;
    const float* W = g.ws + (long long)batch * g.splits * total;
    float* C = g.C + (long long)batch * g.sC;
    const float* bias = g.bias ? g.bias + (long long)batch * g.sBias : nullptr;
    const uint8_t* cm = g.colmask ? g.colmask + (long long)batch * g.sMask : nullptr;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        int m = (int)(e / g.N), n = (int)(e % g.N);
        float s = 0.f;
        for (int k = 0; k < g.splits; ++k) s += W[(long long)k * total + e];
        if (bias) s += bias[n];
        if (cm && !cm[n]) s = 0.f;
        float* dst = C + (long long)m * g.ldc + n;
        *dst = g.accumulate ? (*dst + s) : s;
    }
}

template <int BM, int BN, int TM, int TN>
static int launch_form(int form, const GemmArgs& g, int batch, cudaStream_t st) {
    dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, batch * g.splits);
    switch (form) {
        case CRVAE_GEMM_NT: gemm_f32_kernel<BM, BN, TM, TN, true, true><<<grid, 256, 0, st>>>(g); break;
        case CRVAE_GEMM_TN: gemm_f32_kernel<BM, BN, TM, TN, false, false><<<grid, 256, 0, st>>>(g); break;
        case CRVAE_GEMM_NN: gemm_f32_kernel<BM, BN, TM, TN, true, false><<<grid, 256, 0, st>>>(g); break;
        default: set_error("gemm: bad form %d", form); return CRVAE_E_BADARG;
    }
    int rc = check_launch("gemm_f32_kernel");
    if (rc) return rc;
    if (g.splits > 1) {
        long long total = (long long)g.M * g.N;
        int bx = (int)((total + 255) / 256);
        if (bx > 1024) bx = 1024;
        gemm_split_reduce_kernel<<<dim3(bx, batch), 256, 0, st>>>(g);
        rc = check_launch("gemm_split_reduce_kernel");
    }
    return rc;
}

// Host-side dispatcher shared by the C-ABI entry points.
int gemm_f32(int form, int batch, int M, int N, int K, const float* A, int lda, long long sA,
             const float* B, int ldb, long long sB, float* C, int ldc, long long sC,
             const float* bias, long long sBias, const uint8_t* colmask, long long sMask,
             int accumulate, int splits, float* ws, cudaStream_t st) {
    if (batch <= 0 || M <= 0 || N <= 0) return 0;
    GemmArgs g;
    g.A = A; g.B = B; g.C = C; g.bias = bias; g.colmask = colmask;
    g.M = M; g.N = N; g.K = K; g.lda = lda; g.ldb = ldb; g.ldc = ldc;
    g.sA = sA; g.sB = sB; g.sC = sC; g.sBias = sBias; g.sMask = sMask;
    g.splits = splits < 1 ? 1 : splits; g.accumulate = accumulate; g.ws = ws;
    if (g.splits > 1 && ws == nullptr) { set_error("gemm: split-K needs a workspace"); return CRVAE_E_BADARG; }
    if (M >= 128 && N >= 96) return launch_form<128, 128, 8, 8>(form, g, batch, st);
    return launch_form<64, 64, 4, 4>(form, g, batch, st);
}

int pick_splits(int batch, int M, int N, int K) {
    // aim for >= 2 CTAs per SM (148 SMs); keep >= 4 k-tiles of 16 per split
    long long tiles = (long long)batch * ((M >= 128 && N >= 96) ? ((M + 127) / 128) * ((N + 127) / 128)
                                                                 : ((M + 63) / 64) * ((N + 63) / 64));
    int want = (int)((296 + tiles - 1) / tiles);
    int maxs = K / 64;
    if (maxs < 1) maxs = 1;
    if (want > maxs) want = maxs;
    if (want > 32) want = 32;
    return want < 1 ? 1 : want;
}

}  // namespace crvae

using namespace crvae;

extern "C" int crvae_gemm_f32(int form, int batch, int M, int N, int K, const float* A, int lda,
                              int64_t sA, const float* B, int ldb, int64_t sB, float* C, int ldc,
                              int64_t sC, const float* bias, int64_t sBias, int accumulate,
                              void* stream) {
    CRVAE_REQUIRE(A && B && C, "null operand");
    CRVAE_REQUIRE(M >= 0 && N >= 0 && K >= 0 && batch >= 0, "negative size");
    return gemm_f32(form, batch, M, N, K, A, lda, sA, B, ldb, sB, C, ldc, sC, bias, sBias, nullptr, 0,
                    accumulate, 1, nullptr, (cudaStream_t)stream);
}

// gates[i][t][b][:] = b_ih[i] + x[t][b][:] . w_ih[i]^T   for t >= t_skip   (GRU.forward :115-119)
extern "C" int crvae_proj_fwd(const float* x, const float* w_ih, const float* b_ih, float* gates,
                              int P, int T, int B, int K, int t_skip, void* stream) {
    CRVAE_REQUIRE(x && w_ih && b_ih && gates, "null operand");
    CRVAE_REQUIRE(P >= 0 && T > 0 && B > 0 && K > 0 && t_skip >= 0 && t_skip <= T, "bad size");
    const int G = CRVAE_G;
    int M = (T - t_skip) * B;
    return gemm_f32(CRVAE_GEMM_NT, P, M, G, K, x + (long long)t_skip * B * K, K, 0, w_ih, K,
                    (long long)G * K, gates + (long long)t_skip * B * G, G, (long long)T * B * G, b_ih, G,
                    nullptr, 0, 0, 1, nullptr, (cudaStream_t)stream);
}

extern "C" size_t crvae_proj_wgrad_workspace(int P, int T, int B, int K) {
    int splits = pick_splits(P, CRVAE_G, K, T * B);
    return splits > 1 ? (size_t)P * splits * CRVAE_G * K * sizeof(float) : 16;
}

// dw_ih[i] = sum_{t>=t_skip,b} dgates[i][t][b][:]^T x[t][b][:]   (autograd of the projection, :497)
extern "C" int crvae_proj_wgrad(const float* dgates, const float* x, const uint8_t* mask, float* dw_ih,
                                int P, int T, int B, int K, int t_skip, void* workspace, void* stream) {
    CRVAE_REQUIRE(dgates && x && dw_ih, "null operand");
    CRVAE_REQUIRE(P >= 0 && T > 0 && B > 0 && K > 0 && t_skip >= 0 && t_skip <= T, "bad size");
    const int G = CRVAE_G;
    int R = (T - t_skip) * B;  // reduction length
    int splits = pick_splits(P, G, K, T * B);
    if (splits > 1) CRVAE_REQUIRE(workspace, "workspace required");
    if (R == 0) {
        cudaError_t e = cudaMemsetAsync(dw_ih, 0, (size_t)P * G * K * sizeof(float), (cudaStream_t)stream);
        return (int)e;
    }
    return gemm_f32(CRVAE_GEMM_TN, P, G, K, R, dgates + (long long)t_skip * B * G, G, (long long)T * B * G,
                    x + (long long)t_skip * B * K, K, 0, dw_ih, K, (long long)G * K, nullptr, 0, mask, K,
                    0, splits, (float*)workspace, (cudaStream_t)stream);
}

// ------------------------------------------------------------------------------------------------------------------
// Gather-packed ragged heads (phase 2, CRVAE_lorenz96.py:115, :200-201, :788-790): head i reads only its k_i connected
// series, so its first-layer weight is (3H, k_i), not (3H, p).  Packed storage: w_ih [P, G, Kp] with Kp = the widest
// head's input count rounded up to 4, cols [P, Kp] the series index of every packed column (ascending; padding columns
// are masked), xg [P, T, B, Kp] the per-head gathered input.  The two GEMMs are the exact FFMA kernel with a per-head A /
// B operand; summing the packed columns in ascending series order is the masked-dense sum with its exact zeros removed,
// so the two forms agree BIT FOR BIT (the reduction of the weight gradient is cut into the same number of splits the
// masked-dense form would use for `K_dense` series).
// ------------------------------------------------------------------------------------------------------------------
namespace crvae {
__global__ void gather_cols_kernel(const float* __restrict__ x, const int* __restrict__ cols, const uint8_t* __restrict__ mask,
                                   float* __restrict__ xg, int P, long long rows, int K, int Kp) {
    const long long n = (long long)P * rows * Kp;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(e % Kp);
        const long long r = (e / Kp) % rows;
        const int i = (int)(e / ((long long)Kp * rows));
        const bool live = mask == nullptr || mask[(long long)i * Kp + c] != 0;
        xg[e] = live ? __ldg(x + r * K + cols[(long long)i * Kp + c]) : 0.f;
    }
}
}  // namespace crvae

extern "C" int crvae_gather_cols(const float* x, const int* cols, const uint8_t* mask, float* xg, int P, int64_t rows, int K,
                                 int Kp, void* stream) {
    CRVAE_REQUIRE(x && cols && xg && P >= 0 && rows >= 0 && K > 0 && Kp > 0, "bad argument");
    const long long n = (long long)P * rows * Kp;
    if (n == 0) return 0;
    long long blocks = (n + 255) / 256;
    if (blocks > 148 * 16) blocks = 148 * 16;
    gather_cols_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(x, cols, mask, xg, P, rows, K, Kp);
    return check_launch("gather_cols_kernel");
}

extern "C" int crvae_proj_fwd_packed(const float* xg, const float* w_ih, const float* b_ih, float* gates,
                                     int P, int T, int B, int Kp, int t_skip, void* stream) {
    CRVAE_REQUIRE(xg && w_ih && b_ih && gates, "null operand");
    CRVAE_REQUIRE(P >= 0 && T > 0 && B > 0 && Kp > 0 && t_skip >= 0 && t_skip <= T, "bad size");
    const int G = CRVAE_G;
    const int M = (T - t_skip) * B;
    return gemm_f32(CRVAE_GEMM_NT, P, M, G, Kp, xg + (long long)t_skip * B * Kp, Kp, (long long)T * B * Kp, w_ih, Kp,
                    (long long)G * Kp, gates + (long long)t_skip * B * G, G, (long long)T * B * G, b_ih, G,
                    nullptr, 0, 0, 1, nullptr, (cudaStream_t)stream);
}

extern "C" size_t crvae_proj_wgrad_packed_workspace(int P, int T, int B, int Kp, int K_dense) {
    int splits = pick_splits(P, CRVAE_G, K_dense, T * B);
    return splits > 1 ? (size_t)P * splits * CRVAE_G * Kp * sizeof(float) : 16;
}

extern "C" int crvae_proj_wgrad_packed(const float* dgates, const float* xg, const uint8_t* mask, float* dw_ih,
                                       int P, int T, int B, int Kp, int K_dense, int t_skip, void* workspace, void* stream) {
    CRVAE_REQUIRE(dgates && xg && dw_ih, "null operand");
    CRVAE_REQUIRE(P >= 0 && T > 0 && B > 0 && Kp > 0 && K_dense >= Kp / 4 && t_skip >= 0 && t_skip <= T, "bad size");
    const int G = CRVAE_G;
    const int R = (T - t_skip) * B;
    const int splits = pick_splits(P, G, K_dense, T * B);
    if (splits > 1) CRVAE_REQUIRE(workspace, "workspace required");
    if (R == 0) return (int)cudaMemsetAsync(dw_ih, 0, (size_t)P * G * Kp * sizeof(float), (cudaStream_t)stream);
    return gemm_f32(CRVAE_GEMM_TN, P, G, Kp, R, dgates + (long long)t_skip * B * G, G, (long long)T * B * G,
                    xg + (long long)t_skip * B * Kp, Kp, (long long)T * B * Kp, dw_ih, Kp, (long long)G * Kp, nullptr, 0, mask, Kp,
                    0, splits, (float*)workspace, (cudaStream_t)stream);
}
