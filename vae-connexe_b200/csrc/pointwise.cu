// Memory-bound kernels of the CR-VAE hot path: fused reparameterisation + KL, fused MSE
// forward/backward, gradient step, fused GD + group-lasso prox + GC norms, Adam.
// All reductions are fixed-order (deterministic); nothing here uses atomics.
#include "common.cuh"

namespace crvae {

constexpr int H = CRVAE_HIDDEN;
constexpr int G = CRVAE_G;

// block-wide sum of doubles, fixed order; result valid in every thread
__device__ double block_sum(double v, double* sh /* >= 32 doubles */) {
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    double s = 0.0;
    for (int i = 0; i < nw; ++i) s += sh[i];
    return s;
}

// ---------------------------------------------------------------------------------------------
// z = mu + exp(0.5*log_var)*eps ; KL (CRVAE.forward :210-216, trainer :486)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) latent_fwd_kernel(const float* __restrict__ lat,
                                                          const float* __restrict__ eps,
                                                          float* __restrict__ z, float* __restrict__ kl_out,
                                                          int B, int Z, int kl_form) {
    __shared__ double sh[32];
    double part = 0.0;
    const int n = B * Z;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        int b = e / Z, h = e % Z;
        float mu = lat[b * 2 * Z + h], lv = lat[b * 2 * Z + Z + h];
        float term;
        if (kl_form == CRVAE_KL_LOGSIGMA) {          // lv holds log(sigma): z = mu + exp(s)*0.5*eps (CRVAE.py:74)
            z[e] = __fadd_rn(mu, __fmul_rn(__fmul_rn(expf(lv), 0.5f), eps[e]));
            term = 1.f + 2.f * lv - mu * mu - expf(2.f * lv);
        } else {
            float sigma = expf(__fmul_rn(0.5f, lv));
            z[e] = __fadd_rn(mu, __fmul_rn(sigma, eps[e]));
            if (kl_form == CRVAE_KL_SWAPPED) term = 1.f + mu - lv * lv - expf(mu);
            else term = 1.f + lv - mu * mu - expf(lv);
        }
        part += (double)(-0.5f * term);
    }
    double tot = block_sum(part, sh);
    if (threadIdx.x == 0) kl_out[0] = (float)(tot / (double)B);
}

// dz = sum over heads of dh0 (+ peer partial); gradient into [mu | log_var]
__global__ void latent_bwd_kernel(const float* __restrict__ dh0, int P, const float* __restrict__ dz_extra,
                                  const float* __restrict__ lat, const float* __restrict__ eps, float beta,
                                  int kl_form, float* __restrict__ dlat, float* __restrict__ dz_out, int B, int Z) {
    const int n = B * Z;
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    float dz = 0.f;
    int i = 0;
    for (; i + 16 <= P; i += 16) {              // 16 loads in flight, added in head order (fixed)
        float a[16];
#pragma unroll
        for (int u = 0; u < 16; ++u) a[u] = __ldcs(dh0 + (long long)(i + u) * n + e);
#pragma unroll
        for (int u = 0; u < 16; ++u) dz += a[u];
    }
    for (; i + 4 <= P; i += 4) {
        float a0 = dh0[(long long)i * n + e], a1 = dh0[(long long)(i + 1) * n + e];
        float a2 = dh0[(long long)(i + 2) * n + e], a3 = dh0[(long long)(i + 3) * n + e];
        dz += a0; dz += a1; dz += a2; dz += a3;     // fixed order
    }
    for (; i < P; ++i) dz += dh0[(long long)i * n + e];
    if (dz_out) dz_out[e] = dz;
    if (!dlat) return;
    if (dz_extra) dz += dz_extra[e];
    const int b = e / Z, h = e % Z;
    const float mu = lat[b * 2 * Z + h], lv = lat[b * 2 * Z + Z + h];
    const float invB = 1.f / (float)B;
    float dkl_mu, dkl_lv;
    if (kl_form == CRVAE_KL_LOGSIGMA) {
        dkl_mu = mu * invB;
        dkl_lv = -(1.f - expf(2.f * lv)) * invB;
        dlat[b * 2 * Z + h] = dz + beta * dkl_mu;
        dlat[b * 2 * Z + Z + h] = dz * eps[e] * 0.5f * expf(lv) + beta * dkl_lv;
        return;
    }
    if (kl_form == CRVAE_KL_SWAPPED) {
        dkl_mu = -0.5f * (1.f - expf(mu)) * invB;
        dkl_lv = lv * invB;
    } else {
        dkl_mu = mu * invB;
        dkl_lv = -0.5f * (1.f - expf(lv)) * invB;
    }
    const float sigma = expf(0.5f * lv);
    dlat[b * 2 * Z + h] = dz + beta * dkl_mu;
    dlat[b * 2 * Z + Z + h] = dz * eps[e] * 0.5f * sigma + beta * dkl_lv;
}

// ---------------------------------------------------------------------------------------------
// MSE forward + backward, one CTA per head (trainer :484 / :509, residual :599 / :639)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) mse_kernel(const float* __restrict__ pred, const float* __restrict__ target,
                                                  float* __restrict__ sse, float* __restrict__ dpred,
                                                  float* __restrict__ err, int n, float dscale) {
    __shared__ double sh[32];
    const long long base = (long long)blockIdx.x * n;
    const float scale = dscale > 0.f ? dscale : 2.f / (float)n;
    double part = 0.0;
    for (int e = threadIdx.x; e < n; e += blockDim.x) {
        float p = pred[base + e], t = target[base + e];
        float d = p - t;
        part += (double)d * (double)d;
        if (dpred) dpred[base + e] = scale * d;
        if (err) err[base + e] = t - p;
    }
    double tot = block_sum(part, sh);
    if (threadIdx.x == 0) sse[blockIdx.x] = (float)tot;
}

// ---------------------------------------------------------------------------------------------
// theta <- theta - lr*grad (:498-499); the product is rounded before the subtraction, as torch does
// ---------------------------------------------------------------------------------------------
__global__ void gd_step_kernel(float* __restrict__ theta, const float* __restrict__ grad, long long n, float lr) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        theta[e] = __fsub_rn(theta[e], __fmul_rn(lr, grad[e]));
}

__global__ void axpy_kernel(float* __restrict__ y, const float* __restrict__ x, long long n, float alpha) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        y[e] = fmaf(alpha, x[e], y[e]);
}

// ---------------------------------------------------------------------------------------------
// Fused GD + group-lasso prox + GC norms over w_ih [P,G,K]  (:498-499, prox_update :308-314, GC :297)
// CTA = 32 columns x 8 row groups of one head; a thread keeps its 24 column entries in registers,
// so W and dW are each read once and W written once (12 B per element).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) gd_prox_gc_kernel(float* __restrict__ w, const float* __restrict__ dw,
                                                         const uint8_t* __restrict__ mask,
                                                         float* __restrict__ col_norm, int K, float lr, float thr,
                                                         int do_prox) {
    __shared__ double red[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int head = blockIdx.y, j = blockIdx.x * 32 + tx;
    const bool live = j < K;
    const bool keep = live && (mask == nullptr || mask[(long long)head * K + j] != 0);
    float v[24], d[24];
    double ss = 0.0;
    // all 48 loads of the thread go out before the first use (one memory round trip instead of 24)
    const long long idx0 = ((long long)head * G + ty) * K + (keep ? j : 0);
#pragma unroll
    for (int q = 0; q < 24; ++q) {
        v[q] = keep ? __ldcs(w + idx0 + (long long)(8 * q) * K) : 0.f;
        d[q] = (keep && dw) ? __ldcs(dw + idx0 + (long long)(8 * q) * K) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < 24; ++q) {
        float x = v[q];
        if (dw) x = __fsub_rn(x, __fmul_rn(lr, d[q]));
        v[q] = x;
        ss += (double)x * (double)x;
    }
    red[ty][tx] = ss;
    __syncthreads();
    double tot = 0.0;
#pragma unroll
    for (int y = 0; y < 8; ++y) tot += red[y][tx];
    float nu = (float)sqrt(tot);
    if (do_prox) {
        const float den = fmaxf(nu, thr);
        const float fac = fmaxf(__fsub_rn(nu, thr), 0.f);
        ss = 0.0;
#pragma unroll
        for (int q = 0; q < 24; ++q) {
            v[q] = __fmul_rn(__fdiv_rn(v[q], den), fac);
            ss += (double)v[q] * (double)v[q];
        }
        __syncthreads();
        red[ty][tx] = ss;
        __syncthreads();
        tot = 0.0;
#pragma unroll
        for (int y = 0; y < 8; ++y) tot += red[y][tx];
        nu = (float)sqrt(tot);
    }
    if (live) {
#pragma unroll
        for (int q = 0; q < 24; ++q) {
            const int g = ty + 8 * q;
            w[((long long)head * G + g) * K + j] = v[q];
        }
        if (ty == 0 && col_norm) col_norm[(long long)head * K + j] = nu;
    }
}

// ---------------------------------------------------------------------------------------------
// Adam, torch.optim.Adam default semantics (:565, :612-614)
// ---------------------------------------------------------------------------------------------
__global__ void adam_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
                            float* __restrict__ v, long long n, float one_minus_b1, float b2, float one_minus_b2,
                            float eps, float neg_step_size, float bc2_sqrt) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const float g = grad[e];
        // exp_avg.lerp_(grad, 1-b1): start + w*(end-start) (w < 0.5 branch of ATen's lerp)
        float mm = __fadd_rn(m[e], __fmul_rn(one_minus_b1, __fsub_rn(g, m[e])));
        // exp_avg_sq.mul_(b2).addcmul_(grad, grad, value=1-b2)
        float vv = __fadd_rn(__fmul_rn(v[e], b2), __fmul_rn(__fmul_rn(one_minus_b2, g), g));
        float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv), bc2_sqrt), eps);
        // param.addcdiv_(exp_avg, denom, value=-step_size): self + (value*t1)/t2, left to right as ATen
        theta[e] = __fadd_rn(theta[e], __fdiv_rn(__fmul_rn(neg_step_size, mm), denom));
        m[e] = mm; v[e] = vv;
    }
}

// Adam with the step count held in device memory (CUDA-graph friendly): step = *counter + 1; the
// bias corrections are evaluated in double exactly as torch/optim/adam.py does on the host.
__global__ void adam_dev_kernel(float* __restrict__ theta, const float* __restrict__ grad, float* __restrict__ m,
                                float* __restrict__ v, long long n, double lr, double b1, double b2, float eps,
                                const int* __restrict__ counter) {
    __shared__ float s_neg_step, s_bc2_sqrt;
    if (threadIdx.x == 0) {
        const int step = *counter + 1;
        const double bc1 = 1.0 - pow(b1, (double)step), bc2 = 1.0 - pow(b2, (double)step);
        s_neg_step = (float)(-(lr / bc1));
        s_bc2_sqrt = (float)sqrt(bc2);
    }
    __syncthreads();
    const float one_minus_b1 = (float)(1.0 - b1), b2f = (float)b2, one_minus_b2 = (float)(1.0 - b2);
    const float neg_step_size = s_neg_step, bc2_sqrt = s_bc2_sqrt;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const float g = grad[e];
        float mm = __fadd_rn(m[e], __fmul_rn(one_minus_b1, __fsub_rn(g, m[e])));
        float vv = __fadd_rn(__fmul_rn(v[e], b2f), __fmul_rn(__fmul_rn(one_minus_b2, g), g));
        float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(vv), bc2_sqrt), eps);
        theta[e] = __fadd_rn(theta[e], __fdiv_rn(__fmul_rn(neg_step_size, mm), denom));
        m[e] = mm; v[e] = vv;
    }
}
__global__ void incr_kernel(int* counter) { *counter += 1; }

__global__ void __launch_bounds__(1024) sumsq_kernel(const float* __restrict__ x, long long n, float* __restrict__ out) {
    __shared__ double sh[32];
    double part = 0.0;
    for (long long e = threadIdx.x; e < n; e += blockDim.x) part += (double)x[e] * (double)x[e];
    double tot = block_sum(part, sh);
    if (threadIdx.x == 0) out[0] = (float)tot;
}

__global__ void __launch_bounds__(256) dot_small_kernel(const float* __restrict__ x, int n, float scale,
                                                        float* __restrict__ out) {
    __shared__ double sh[32];
    double part = 0.0;
    for (int e = threadIdx.x; e < n; e += blockDim.x) part += (double)x[e];
    double tot = block_sum(part, sh);
    if (threadIdx.x == 0) out[0] = (float)(tot * (double)scale);
}

// y = tanh(x)  /  dx = dy * (1 - y^2)      (VRAE4E: z = tanh(linear_hidden(z)), CRVAE_lorenz96.py:164)
__global__ void tanh_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        y[e] = tanhf(x[e]);
}
__global__ void tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long long n) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x)
        dx[e] = dy[e] * (1.f - y[e] * y[e]);
}
// generic output activation of VRAE.py's decoder (:60-68): kind 0 tanh, 1 sigmoid, 2 relu, 3 identity
__global__ void act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long n, int kind) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const float v = x[e];
        y[e] = kind == 0 ? tanhf(v) : kind == 1 ? sigmoidf_acc(v) : kind == 2 ? fmaxf(v, 0.f) : v;
    }
}
__global__ void act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, long long n,
                               int kind) {
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const float v = y[e];
        const float gy = kind == 0 ? 1.f - v * v : kind == 1 ? v * (1.f - v) : kind == 2 ? (v > 0.f ? 1.f : 0.f) : 1.f;
        dx[e] = dy[e] * gy;
    }
}
// out[c][r] = in[r][c]  (small 2-D transpose through shared memory; err [P][T*B] <-> [T*B][P])
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int rows, int cols) {
    __shared__ float tile[32][33];
    int c = blockIdx.x * 32 + threadIdx.x, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        if (r0 + i < rows && c < cols) tile[i][threadIdx.x] = in[(long long)(r0 + i) * cols + c];
    __syncthreads();
    int r = r0 + threadIdx.x, c0 = blockIdx.x * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y)
        if (c0 + i < cols && r < rows) out[(long long)(c0 + i) * rows + r] = tile[threadIdx.x][i];
}

static inline int grid_for(long long n, int block = 256, int cap = 148 * 8) {
    long long b = (n + block - 1) / block;
    if (b > cap) b = cap;
    if (b < 1) b = 1;
    return (int)b;
}


// ---------------------------------------------------------------------------------------------
// Batch binding: the window batch X [B, Te+Td, p] -> time-major encoder / decoder inputs (+ their tf32 hi | lo
// splits for the tensor-core projection) and the per-head targets, in one pass (arrange of :208 / :119 / :484).
// Block = (32 batch rows, one step, 128 columns) staged through shared memory so every global access is coalesced.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) bind_batch_kernel(const float* __restrict__ X, float* __restrict__ enc_in,
                                                         float* __restrict__ enc_hi, float* __restrict__ enc_lo,
                                                         float* __restrict__ dec_in, float* __restrict__ dec_hi,
                                                         float* __restrict__ dec_lo, float* __restrict__ target, int B, int p,
                                                         int Te, int Td, int head_lo, int P) {
    __shared__ float tile[32][129];
    const int b0 = blockIdx.x * 32, t = blockIdx.y, c0 = blockIdx.z * 128;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int b = b0 + r;
        for (int c = tx; c < 128; c += 32)
            tile[r][c] = (b < B && c0 + c < p) ? __ldg(X + ((long long)b * (Te + Td) + t) * p + c0 + c) : 0.f;
    }
    __syncthreads();
    float *dst = nullptr, *dhi = nullptr, *dlo = nullptr;
    if (t < Te) {
        dst = enc_in + (long long)t * B * p; dhi = enc_hi ? enc_hi + (long long)t * B * p : nullptr; dlo = enc_lo ? enc_lo + (long long)t * B * p : nullptr;
    } else if (t < Te + Td - 1) {        // decoder step s = t - Te + 1 sees x_{t}; step 0 stays zero
        const long long off = (long long)(t - Te + 1) * B * p;
        dst = dec_in + off; dhi = dec_hi ? dec_hi + off : nullptr; dlo = dec_lo ? dec_lo + off : nullptr;
    }
    if (dst) {
        for (int r = ty; r < 32; r += 8) {
            const int b = b0 + r;
            if (b >= B) continue;
            for (int c = tx; c < 128 && c0 + c < p; c += 32) {
                const float v = tile[r][c];
                const long long o = (long long)b * p + c0 + c;
                dst[o] = v;
                if (dhi) {
                    uint32_t h;
                    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(h) : "f"(v));
                    const float hf = __uint_as_float(h);
                    dhi[o] = hf;
                    dlo[o] = __fsub_rn(v, hf);
                }
            }
        }
    }
    if (t >= Te && target) {             // target[i][t - Te][b] = X[b][t][head_lo + i]
        for (int c = ty; c < 128; c += 8) {
            const int i = c0 + c - head_lo;
            if (i < 0 || i >= P || b0 + tx >= B) continue;
            target[((long long)i * Td + (t - Te)) * B + b0 + tx] = tile[tx][c];
        }
    }
}

}  // namespace crvae

using namespace crvae;

extern "C" int crvae_latent_fwd(const float* lat, const float* eps, float* z, float* kl_out, int B, int Z,
                                int kl_form, void* stream) {
    CRVAE_REQUIRE(lat && eps && z && kl_out && B > 0 && Z > 0, "bad argument");
    latent_fwd_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(lat, eps, z, kl_out, B, Z, kl_form);
    return check_launch("latent_fwd_kernel");
}

extern "C" int crvae_latent_bwd(const float* dh0, int P, const float* dz_extra, const float* lat,
                                const float* eps, float beta, int kl_form, float* dlat, float* dz_out, int B,
                                int Z, void* stream) {
    CRVAE_REQUIRE(B > 0 && Z > 0 && P >= 0 && (P == 0 || dh0), "bad argument");
    CRVAE_REQUIRE(dlat == nullptr || (lat && eps), "lat/eps required with dlat");
    int n = B * Z;
    latent_bwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(dh0, P, dz_extra, lat, eps, beta, kl_form,
                                                                         dlat, dz_out, B, Z);
    return check_launch("latent_bwd_kernel");
}

extern "C" int crvae_mse_fwd_bwd(const float* pred, const float* target, float* sse, float* dpred, float* err,
                                 int P, int T, int B, float dscale, void* stream) {
    CRVAE_REQUIRE(pred && target && sse && P >= 0 && T > 0 && B > 0, "bad argument");
    if (P == 0) return 0;
    mse_kernel<<<P, 1024, 0, (cudaStream_t)stream>>>(pred, target, sse, dpred, err, T * B, dscale);
    return check_launch("mse_kernel");
}

extern "C" int crvae_gd_step(float* theta, const float* grad, int64_t n, float lr, void* stream) {
    CRVAE_REQUIRE(theta && grad && n >= 0, "bad argument");
    if (n == 0) return 0;
    gd_step_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(theta, grad, n, lr);
    return check_launch("gd_step_kernel");
}

extern "C" int crvae_axpy(float* y, const float* x, int64_t n, float alpha, void* stream) {
    CRVAE_REQUIRE(y && x && n >= 0, "bad argument");
    if (n == 0) return 0;
    axpy_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(y, x, n, alpha);
    return check_launch("axpy_kernel");
}

extern "C" int crvae_gd_prox_gc(float* w_ih, const float* dw_ih, const uint8_t* mask, float* col_norm, int P,
                                int K, float lr, float thr, int do_prox, void* stream) {
    CRVAE_REQUIRE(w_ih && P >= 0 && K > 0, "bad argument");
    if (P == 0) return 0;
    gd_prox_gc_kernel<<<dim3((K + 31) / 32, P), 256, 0, (cudaStream_t)stream>>>(w_ih, dw_ih, mask, col_norm, K, lr,
                                                                                thr, do_prox);
    return check_launch("gd_prox_gc_kernel");
}

extern "C" int crvae_adam_step(float* theta, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                               double lr, double beta1, double beta2, double eps, int step, void* stream) {
    CRVAE_REQUIRE(theta && grad && exp_avg && exp_avg_sq && n >= 0 && step >= 1, "bad argument");
    if (n == 0) return 0;
    // scalars computed in double exactly as torch/optim/adam.py does, then rounded to fp32 at use
    double bc1 = 1.0 - pow(beta1, (double)step);
    double bc2 = 1.0 - pow(beta2, (double)step);
    double step_size = lr / bc1;
    double bc2_sqrt = sqrt(bc2);
    adam_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(theta, grad, exp_avg, exp_avg_sq, n,
                                                               (float)(1.0 - beta1), (float)beta2,
                                                               (float)(1.0 - beta2), (float)eps,
                                                               (float)(-step_size), (float)bc2_sqrt);
    return check_launch("adam_kernel");
}

extern "C" int crvae_sumsq(const float* x, int64_t n, float* out, void* stream) {
    CRVAE_REQUIRE(x && out && n >= 0, "bad argument");
    sumsq_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(x, n, out);
    return check_launch("sumsq_kernel");
}

extern "C" int crvae_dot_small(const float* x, int n, float scale, float* out, void* stream) {
    CRVAE_REQUIRE(x && out && n >= 0 && n <= (1 << 20), "bad argument");
    dot_small_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(x, n, scale, out);
    return check_launch("dot_small_kernel");
}

extern "C" int crvae_tanh_fwd(const float* x, float* y, int64_t n, void* stream) {
    CRVAE_REQUIRE(x && y && n >= 0, "bad argument");
    if (n == 0) return 0;
    tanh_fwd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, y, n);
    return check_launch("tanh_fwd_kernel");
}

extern "C" int crvae_tanh_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream) {
    CRVAE_REQUIRE(dy && y && dx && n >= 0, "bad argument");
    if (n == 0) return 0;
    tanh_bwd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(dy, y, dx, n);
    return check_launch("tanh_bwd_kernel");
}

namespace crvae {
// Bernoulli reconstruction loss of MixtureCSRAE (CSRAE_new.py:144): F.binary_cross_entropy_with_logits(reduction="sum"),
// numerically stable form max(l,0) - l*x + log1p(exp(-|l|)); dlogits = (sigmoid(l) - x) * dscale.  Per-CTA partial sums in
// fp64, summed in CTA order by bce_finish_kernel (deterministic).
__global__ void __launch_bounds__(256) bce_logits_kernel(const float* __restrict__ logits, const float* __restrict__ x,
                                                         double* __restrict__ part, float* __restrict__ dlogits, long long n, float dscale) {
    __shared__ double sh[32];
    double acc = 0.0;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const float l = logits[e], t = x[e];
        acc += (double)(fmaxf(l, 0.f) - l * t + log1pf(expf(-fabsf(l))));
        if (dlogits) dlogits[e] = (sigmoidf_acc(l) - t) * dscale;
    }
    const double tot = block_sum(acc, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = tot;
}
__global__ void bce_finish_kernel(const double* __restrict__ part, int nparts, float* __restrict__ out) {
    double s = 0.0;
    for (int i = 0; i < nparts; ++i) s += part[i];
    out[0] = (float)s;
}
// Family-B ISTA (CRVAE.py:134-150): one warp per row of W_in (one candidate parent), norm over the row's H entries.
__global__ void ista_rows_kernel(float* __restrict__ w, const float* __restrict__ dw, float* __restrict__ row_norm, long long rows,
                                 int cols, float lr, float thr, int do_prox) {
    const long long row = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (row >= rows) return;
    float* wr = w + row * cols;
    float ss = 0.f;
    for (int c = lane; c < cols; c += 32) {
        float v = wr[c];
        if (dw) v = __fsub_rn(v, __fmul_rn(lr, dw[row * cols + c]));       // W_tmp = W - lr * grad
        ss = fmaf(v, v, ss);
    }
    ss = warp_sum(ss);
    const float nu = sqrtf(ss);
    float shrink = 1.f;
    if (do_prox) shrink = fmaxf(__fsub_rn(1.f, __fdiv_rn(thr, nu)), 0.f);  // thr / 0 = inf -> shrink = 0
    if (do_prox || dw)
        for (int c = lane; c < cols; c += 32) {
            float v = wr[c];
            if (dw) v = __fsub_rn(v, __fmul_rn(lr, dw[row * cols + c]));
            wr[c] = do_prox ? __fmul_rn(v, shrink) : v;
        }
    if (lane == 0 && row_norm) row_norm[row] = nu * shrink;               // norm of the stored row
}
// Test-mode generation, the step between two recurrent updates (CRVAE_lorenz96.py:232-236, :281-283): the p heads' scalar
// outputs of this step become the next input row of EVERY head.  y [R][W][B]: outputs gathered over R head shards of at
// most W heads each (balanced contiguous partition: the first `rem` shards hold base+1 heads, the rest base); one pass
// writes x_next[b][j] = y_j[b] (+ scale * noise[b][t][j]), the stored sequence out[b][t][j], and the tf32 hi / lo split
// of x_next for the tensor-core projection of the next step.
__global__ void gen_scatter_kernel(const float* __restrict__ y, const float* __restrict__ noise, float* __restrict__ x,
                                   float* __restrict__ x_hi, float* __restrict__ x_lo, float* __restrict__ out, int B, int p, int t,
                                   int steps, int base, int rem, int widest, float scale) {
    const long long n = (long long)B * p;
    for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
        const int b = (int)(e / p), j = (int)(e - (long long)b * p);
        int r, l;
        if (j < rem * (base + 1)) { r = j / (base + 1); l = j - r * (base + 1); }
        else { const int jj = j - rem * (base + 1); r = rem + jj / base; l = jj - (r - rem) * base; }
        float v = y[((long long)r * widest + l) * B + b];
        if (noise) v = __fadd_rn(v, __fmul_rn(scale, noise[((long long)b * steps + t) * p + j]));
        x[e] = v;
        out[((long long)b * steps + t) * p + j] = v;
        if (x_hi) {
            uint32_t hi;
            asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(hi) : "f"(v));
            x_hi[e] = __uint_as_float(hi);
            x_lo[e] = __fsub_rn(v, __uint_as_float(hi));
        }
    }
}

}  // namespace crvae

extern "C" size_t crvae_bce_logits_workspace(int64_t n) {
    long long blocks = (n + 255) / 256;
    if (blocks > 592) blocks = 592;
    return (size_t)(blocks > 0 ? blocks : 1) * sizeof(double);
}
extern "C" int crvae_bce_logits_fwd_bwd(const float* logits, const float* x, float* sum_out, float* dlogits, int64_t n, float dscale,
                                        void* workspace, void* stream) {
    CRVAE_REQUIRE(logits && x && sum_out && workspace && n > 0, "bad argument");
    long long blocks = (n + 255) / 256;
    if (blocks > 592) blocks = 592;
    bce_logits_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(logits, x, (double*)workspace, dlogits, n, dscale);
    int rc = check_launch("bce_logits_kernel");
    if (rc) return rc;
    bce_finish_kernel<<<1, 1, 0, (cudaStream_t)stream>>>((const double*)workspace, (int)blocks, sum_out);
    return check_launch("bce_finish_kernel");
}

extern "C" int crvae_ista_rows(float* w, const float* dw, float* row_norm, int64_t rows, int cols, float lr, float thr, int do_prox,
                               void* stream) {
    CRVAE_REQUIRE(w && rows >= 0 && cols > 0 && lr >= 0.f && thr >= 0.f, "bad argument");
    if (rows == 0) return 0;
    const long long threads = rows * 32;
    ista_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(w, dw, row_norm, rows, cols, lr, thr, do_prox);
    return check_launch("ista_rows_kernel");
}

extern "C" int crvae_gen_scatter(const float* y, const float* noise, float* x, float* x_hi, float* x_lo, float* out, int B, int p,
                                 int t, int steps, int base, int rem, int widest, float scale, void* stream) {
    CRVAE_REQUIRE(y && x && out && B > 0 && p > 0 && t >= 0 && t < steps && base >= 0 && rem >= 0 && widest > 0, "bad argument");
    CRVAE_REQUIRE((x_hi == nullptr) == (x_lo == nullptr), "x_hi and x_lo go together");
    CRVAE_REQUIRE(base > 0 || rem * (base + 1) >= p, "partition does not cover the p heads");
    gen_scatter_kernel<<<grid_for((long long)B * p), 256, 0, (cudaStream_t)stream>>>(y, noise, x, x_hi, x_lo, out, B, p, t, steps, base, rem,
                                                                                  widest, scale);
    return check_launch("gen_scatter_kernel");
}

extern "C" int crvae_transpose(const float* in, float* out, int rows, int cols, void* stream) {
    CRVAE_REQUIRE(in && out && rows > 0 && cols > 0, "bad argument");
    transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, (cudaStream_t)stream>>>(in, out, rows, cols);
    return check_launch("transpose_kernel");
}

extern "C" int crvae_adam_step_dev(float* theta, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                                   double lr, double beta1, double beta2, double eps, int* step_counter, void* stream) {
    CRVAE_REQUIRE(theta && grad && exp_avg && exp_avg_sq && step_counter && n >= 0, "bad argument");
    if (n > 0) {
        adam_dev_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(theta, grad, exp_avg, exp_avg_sq, n, lr, beta1, beta2,
                                                                       (float)eps, step_counter);
        int rc = check_launch("adam_dev_kernel");
        if (rc) return rc;
    }
    incr_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(step_counter);
    return check_launch("incr_kernel");
}

extern "C" int crvae_act_fwd(const float* x, float* y, int64_t n, int kind, void* stream) {
    CRVAE_REQUIRE(x && y && n >= 0 && kind >= 0 && kind <= 3, "bad argument");
    if (n == 0) return 0;
    act_fwd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(x, y, n, kind);
    return check_launch("act_fwd_kernel");
}

extern "C" int crvae_act_bwd(const float* dy, const float* y, float* dx, int64_t n, int kind, void* stream) {
    CRVAE_REQUIRE(dy && y && dx && n >= 0 && kind >= 0 && kind <= 3, "bad argument");
    if (n == 0) return 0;
    act_bwd_kernel<<<grid_for(n), 256, 0, (cudaStream_t)stream>>>(dy, y, dx, n, kind);
    return check_launch("act_bwd_kernel");
}

extern "C" int crvae_bind_batch(const float* X, float* enc_in, float* enc_hi, float* enc_lo, float* dec_in, float* dec_hi,
                                float* dec_lo, float* target, int B, int p, int Te, int Td, int head_lo, int P, void* stream) {
    CRVAE_REQUIRE(X && enc_in && dec_in && B > 0 && p > 0 && Te > 0 && Td > 0 && P >= 0 && head_lo >= 0, "bad argument");
    CRVAE_REQUIRE((enc_hi == nullptr) == (enc_lo == nullptr) && (dec_hi == nullptr) == (dec_lo == nullptr), "hi/lo go together");
    CRVAE_REQUIRE(P == 0 || target, "target missing");
    dim3 grid((B + 31) / 32, Te + Td, (p + 127) / 128);
    bind_batch_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(X, enc_in, enc_hi, enc_lo, dec_in, dec_hi, dec_lo, P > 0 ? target : nullptr, B, p,
                                                              Te, Td, head_lo, P);
    return check_launch("bind_batch_kernel");
}
