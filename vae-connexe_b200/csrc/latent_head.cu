// Fused latent head of the CR-VAE encoder (H = Z = 64): the fc_mu | fc_std Linear pair around the
// reparameterisation, forward and backward, as ONE launch each.  These sit on the latency-bound encoder
// chain that every rank replicates; as separate tiny GEMM launches (3 GEMMs + a split reduction + a
// pointwise kernel) they cost ~80 us of that chain, fused they cost a few microseconds.
//
// Reference arithmetic replaced: self.fc_mu / self.fc_std (CRVAE_lorenz96.py:210-211), the
// reparameterisation (:213-216), the KL term (:486) and autograd through them (:497).
#include "common.cuh"

namespace crvae {

constexpr int LH = CRVAE_HIDDEN;      // 64: encoder hidden width = latent width
constexpr int LROWS = 16;             // batch rows per CTA

// ---------------------------------------------------------------------------------------------
// forward: lat[b] = hT[b] . W^T + bias  ([mu | log_var]);  z = mu + exp(0.5 log_var) eps;  KL partial per CTA,
// summed in CTA order by the last CTA to finish (fixed order -> deterministic)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) latent_head_fwd_kernel(const float* __restrict__ hT, const float* __restrict__ w,
                                                              const float* __restrict__ bias, const float* __restrict__ eps,
                                                              float* __restrict__ lat, float* __restrict__ z,
                                                              float* __restrict__ kl_out, double* __restrict__ part,
                                                              unsigned int* __restrict__ counter, int B, int kl_form) {
    __shared__ float Wt[LH][2 * LH + 4];      // W^T: [k][j]
    __shared__ float Hs[LROWS][LH + 1];
    __shared__ double red[8];
    __shared__ bool is_last;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    const int b = blockIdx.x * LROWS + ty;
    for (int e = tid; e < 2 * LH * LH; e += 256) {
        const int j = e / LH, k = e % LH;     // w is [2H][H]
        Wt[k][j] = __ldg(w + e);
    }
    for (int e = tid; e < LROWS * LH; e += 256) {
        const int r = e / LH, k = e % LH;
        const int bb = blockIdx.x * LROWS + r;
        Hs[r][k] = bb < B ? __ldg(hT + (long long)bb * LH + k) : 0.f;
    }
    __syncthreads();
    float mu[4], lv[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) { mu[c] = 0.f; lv[c] = 0.f; }
#pragma unroll 8
    for (int k = 0; k < LH; ++k) {
        const float a = Hs[ty][k];
        const float4 wm = *reinterpret_cast<const float4*>(&Wt[k][4 * tx]);
        const float4 wl = *reinterpret_cast<const float4*>(&Wt[k][LH + 4 * tx]);
        mu[0] = fmaf(a, wm.x, mu[0]); mu[1] = fmaf(a, wm.y, mu[1]); mu[2] = fmaf(a, wm.z, mu[2]); mu[3] = fmaf(a, wm.w, mu[3]);
        lv[0] = fmaf(a, wl.x, lv[0]); lv[1] = fmaf(a, wl.y, lv[1]); lv[2] = fmaf(a, wl.z, lv[2]); lv[3] = fmaf(a, wl.w, lv[3]);
    }
    double klp = 0.0;
    if (b < B) {
        float zz[4];
        const float4 e4 = *reinterpret_cast<const float4*>(eps + (long long)b * LH + 4 * tx);
        const float ev[4] = {e4.x, e4.y, e4.z, e4.w};
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            mu[c] = __fadd_rn(mu[c], __ldg(bias + 4 * tx + c));
            lv[c] = __fadd_rn(lv[c], __ldg(bias + LH + 4 * tx + c));
            const float sigma = expf(__fmul_rn(0.5f, lv[c]));
            zz[c] = __fadd_rn(mu[c], __fmul_rn(sigma, ev[c]));
            float term;
            if (kl_form == CRVAE_KL_SWAPPED) term = 1.f + mu[c] - lv[c] * lv[c] - expf(mu[c]);
            else term = 1.f + lv[c] - mu[c] * mu[c] - expf(lv[c]);
            klp += (double)(-0.5f * term);
        }
        *reinterpret_cast<float4*>(lat + (long long)b * 2 * LH + 4 * tx) = make_float4(mu[0], mu[1], mu[2], mu[3]);
        *reinterpret_cast<float4*>(lat + (long long)b * 2 * LH + LH + 4 * tx) = make_float4(lv[0], lv[1], lv[2], lv[3]);
        *reinterpret_cast<float4*>(z + (long long)b * LH + 4 * tx) = make_float4(zz[0], zz[1], zz[2], zz[3]);
    }
    klp = warp_sum(klp);
    if ((tid & 31) == 0) red[tid >> 5] = klp;
    __syncthreads();
    if (tid == 0) {
        double s = 0.0;
        for (int i = 0; i < 8; ++i) s += red[i];
        part[blockIdx.x] = s;
        __threadfence();
        is_last = atomicAdd(counter, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (is_last && tid == 0) {
        __threadfence();
        double s = 0.0;
        for (unsigned i = 0; i < gridDim.x; ++i) s += reinterpret_cast<volatile double*>(part)[i];
        kl_out[0] = (float)(s / (double)B);
        *counter = 0u;          // ready for the next launch / graph replay
    }
}

// ---------------------------------------------------------------------------------------------
// backward:  dhT = dlat . W            (CTAs [0, nA): 16 batch rows each)
//            dW  = dlat^T . hT, db = column sums of dlat   (CTAs [nA, nA + 16): 8 rows of dW each, batch walked in order)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) latent_head_bwd_kernel(const float* __restrict__ dlat, const float* __restrict__ hT,
                                                              const float* __restrict__ w, float* __restrict__ dW,
                                                              float* __restrict__ db, float* __restrict__ dhT, int B, int nA) {
    __shared__ __align__(16) float smem[2 * LH * LH + 64 * 8 + 16];
    const int tid = threadIdx.x;
    if ((int)blockIdx.x < nA) {
        float* Ws = smem;                         // [2H][H] natural layout (32 KB)
        float* Ds = smem + 2 * LH * LH;           // unused here beyond 0; dlat rows are read through registers
        (void)Ds;
        const int tx = tid & 15, ty = tid >> 4;
        const int b = blockIdx.x * LROWS + ty;
        for (int e = tid; e < 2 * LH * LH / 4; e += 256) reinterpret_cast<float4*>(Ws)[e] = __ldg(reinterpret_cast<const float4*>(w) + e);
        __syncthreads();
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        const float* drow = dlat + (long long)(b < B ? b : 0) * 2 * LH;
#pragma unroll 4
        for (int j4 = 0; j4 < 2 * LH; j4 += 4) {
            const float4 d4 = __ldg(reinterpret_cast<const float4*>(drow + j4));
            const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                const float4 wv = *reinterpret_cast<const float4*>(Ws + (j4 + jj) * LH + 4 * tx);
                acc[0] = fmaf(dv[jj], wv.x, acc[0]); acc[1] = fmaf(dv[jj], wv.y, acc[1]);
                acc[2] = fmaf(dv[jj], wv.z, acc[2]); acc[3] = fmaf(dv[jj], wv.w, acc[3]);
            }
        }
        if (b < B) *reinterpret_cast<float4*>(dhT + (long long)b * LH + 4 * tx) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    } else {
        float* Hs = smem;                         // [64 rows][64] chunk of hT
        float* Dl = smem + LH * LH;               // [64 rows][8]  slice of dlat
        const int c = blockIdx.x - nA;            // dW rows 8c .. 8c+7
        const int jj = tid >> 5, hx = tid & 31;   // row 8c + jj, columns 2hx, 2hx + 1
        float a0 = 0.f, a1 = 0.f, sb = 0.f;
        for (int b0 = 0; b0 < B; b0 += 64) {
            __syncthreads();
            for (int e = tid; e < 64 * LH / 4; e += 256) {
                const int r = e / (LH / 4);
                reinterpret_cast<float4*>(Hs)[e] = (b0 + r < B) ? __ldg(reinterpret_cast<const float4*>(hT + (long long)(b0 + r) * LH) + (e % (LH / 4)))
                                                                : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            for (int e = tid; e < 64 * 8; e += 256) {
                const int r = e >> 3, q = e & 7;
                Dl[e] = (b0 + r < B) ? __ldg(dlat + (long long)(b0 + r) * 2 * LH + 8 * c + q) : 0.f;
            }
            __syncthreads();
#pragma unroll 8
            for (int r = 0; r < 64; ++r) {
                const float d = Dl[r * 8 + jj];
                const float2 h2 = *reinterpret_cast<const float2*>(Hs + r * LH + 2 * hx);
                a0 = fmaf(d, h2.x, a0); a1 = fmaf(d, h2.y, a1);
                sb += d;
            }
        }
        *reinterpret_cast<float2*>(dW + (long long)(8 * c + jj) * LH + 2 * hx) = make_float2(a0, a1);
        if (hx == 0) db[8 * c + jj] = sb;
    }
}

}  // namespace crvae

using namespace crvae;

extern "C" size_t crvae_latent_head_workspace(int B) { return (size_t)((B + LROWS - 1) / LROWS + 2) * sizeof(double); }

extern "C" int crvae_latent_head_fwd(const float* hT, const float* lat_w, const float* lat_b, const float* eps, float* lat,
                                     float* z, float* kl_out, int B, int kl_form, void* workspace, void* stream) {
    CRVAE_REQUIRE(hT && lat_w && lat_b && eps && lat && z && kl_out && workspace && B > 0, "bad argument");
    CRVAE_REQUIRE(aligned16(hT) && aligned16(eps) && aligned16(lat) && aligned16(z) && aligned16(workspace), "16-byte alignment");
    const int n = (B + LROWS - 1) / LROWS;
    double* part = reinterpret_cast<double*>(workspace);
    unsigned int* counter = reinterpret_cast<unsigned int*>(part + n);     // zero before the first launch; the kernel resets it
    latent_head_fwd_kernel<<<n, 256, 0, (cudaStream_t)stream>>>(hT, lat_w, lat_b, eps, lat, z, kl_out, part, counter, B, kl_form);
    return check_launch("latent_head_fwd_kernel");
}

extern "C" int crvae_latent_head_bwd(const float* dlat, const float* hT, const float* lat_w, float* d_lat_w, float* d_lat_b,
                                     float* dhT, int B, void* stream) {
    CRVAE_REQUIRE(dlat && hT && lat_w && d_lat_w && d_lat_b && dhT && B > 0, "bad argument");
    CRVAE_REQUIRE(aligned16(dlat) && aligned16(hT) && aligned16(lat_w) && aligned16(d_lat_w) && aligned16(dhT), "16-byte alignment");
    const int nA = (B + LROWS - 1) / LROWS;
    latent_head_bwd_kernel<<<nA + 2 * LH / 8, 256, 0, (cudaStream_t)stream>>>(dlat, hT, lat_w, d_lat_w, d_lat_b, dhT, B, nA);
    return check_launch("latent_head_bwd_kernel");
}
