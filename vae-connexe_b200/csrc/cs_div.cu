// Cauchy-Schwarz divergence between the Gaussian posterior and an equal-weight GMM prior, forward + backward
// (CR-CS-RAE.py:124-163: gaussian_overlap, cs_divergence_gmm; trainer use :568-582).
//
//   D_CS(b) = -log( mean_k N(mu_q; mu_k, var_q + var_k) )                     term 1
//             + 0.5 log( mean_{k,k'} N(mu_k; mu_k', var_k + var_k') )          term 2
//             + 0.5 log( N(mu_q; mu_q, 2 var_q) )                              term 3,   clamp(min = 0)
// evaluated exactly like the reference: every overlap is exp(log-density), then mean, then log (NOT
// logsumexp), so underflow behaves the same.  D = H = 64 latent dims, K <= 32 mixture components.
// Inputs follow the reference trainer's (swapped) roles: mu_q = the fc_std output, log var_q = the fc_mu
// output, i.e. mu_q = lat[:, H:2H], logvar_q = lat[:, 0:H].
#include "common.cuh"

namespace crvae {

constexpr int CH = CRVAE_HIDDEN;          // 64 = D
constexpr int CS_MAXK = 32;
constexpr float LOG_2PI = 1.8378770664093453f;

// sum over the 64 threads (2 warps) of a block; result broadcast
__device__ __forceinline__ float block64_sum(float v, float* sh) {
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    return sh[0] + sh[1];
}

// term 2 and its gradient w.r.t. the prior: ws[0] = term2; g2_mu / g2_var [K][D] = d term2 / d(mu_k, var_k)
__global__ void __launch_bounds__(64) cs_prior_kernel(const float* __restrict__ pmu, const float* __restrict__ plv, int K,
                                                      float* __restrict__ ws, float* __restrict__ g2_mu,
                                                      float* __restrict__ g2_var) {
    __shared__ float sh[2];
    __shared__ float P[CS_MAXK][CS_MAXK];
    const int d = threadIdx.x;
    for (int k = 0; k < K; ++k) {
        const float mk = pmu[k * CH + d], vk = expf(plv[k * CH + d]);
        for (int k2 = 0; k2 < K; ++k2) {
            const float s = vk + expf(plv[k2 * CH + d]);
            const float diff = mk - pmu[k2 * CH + d];
            const float sl = block64_sum(logf(s), sh);
            const float sq = block64_sum(diff * diff / s, sh);
            if (d == 0) P[k][k2] = expf((-0.5f * CH * LOG_2PI - 0.5f * sl) + (-0.5f * sq));
        }
    }
    __syncthreads();
    float t2 = 0.f;
    for (int k = 0; k < K; ++k)
        for (int k2 = 0; k2 < K; ++k2) t2 += P[k][k2];
    t2 /= (float)(K * K);
    if (d == 0) ws[0] = t2;
    const float c = 2.f / (float)(K * K);
    for (int k = 0; k < K; ++k) {
        const float mk = pmu[k * CH + d], vk = expf(plv[k * CH + d]);
        float gm = 0.f, gv = 0.f;
        for (int k2 = 0; k2 < K; ++k2) {
            const float s = vk + expf(plv[k2 * CH + d]);
            const float diff = mk - pmu[k2 * CH + d];
            gm += P[k][k2] * (-diff / s);
            gv += P[k][k2] * (-0.5f / s + 0.5f * diff * diff / (s * s));
        }
        g2_mu[k * CH + d] = c * gm;
        g2_var[k * CH + d] = c * gv;
    }
}

// one block per sample: cs[b], gradient into lat (swapped roles) and per-sample prior partials
__global__ void __launch_bounds__(64) cs_sample_kernel(const float* __restrict__ lat, const float* __restrict__ pmu,
                                                       const float* __restrict__ plv, int K, const float* __restrict__ ws,
                                                       float scale /* lambda_cs / B */, float* __restrict__ cs,
                                                       float* __restrict__ dlat, float* __restrict__ part_mu,
                                                       float* __restrict__ part_var, int B) {
    __shared__ float sh[2];
    __shared__ float O[CS_MAXK];
    const int b = blockIdx.x, d = threadIdx.x;
    const float mq = lat[b * 2 * CH + CH + d];             // mu_q     = fc_std output (the trainer's swapped names)
    const float lvq = lat[b * 2 * CH + d];                 // logvar_q = fc_mu output
    const float vq = expf(lvq);
    for (int k = 0; k < K; ++k) {
        const float s = vq + expf(plv[k * CH + d]);
        const float diff = mq - pmu[k * CH + d];
        const float sl = block64_sum(logf(s), sh);
        const float sq = block64_sum(diff * diff / s, sh);
        if (d == 0) O[k] = expf((-0.5f * CH * LOG_2PI - 0.5f * sl) + (-0.5f * sq));
    }
    const float l3sum = block64_sum(logf(2.f * vq), sh);
    __syncthreads();
    float osum = 0.f;
    for (int k = 0; k < K; ++k) osum += O[k];
    const float term1 = osum / (float)K;
    const float term3 = expf(-0.5f * CH * LOG_2PI - 0.5f * l3sum);
    const float raw = -logf(term1) + 0.5f * logf(ws[0]) + 0.5f * logf(term3);
    const bool pass = raw >= 0.f;                          // clamp(min=0) passes the gradient where x >= 0
    if (d == 0) cs[b] = pass ? raw : 0.f;
    float dmq = 0.f, dvq = 0.f;
    for (int k = 0; k < K; ++k) {
        const float s = vq + expf(plv[k * CH + d]);
        const float diff = mq - pmu[k * CH + d];
        const float w = pass ? O[k] / osum : 0.f;          // -d raw / d log O_k
        const float fm = diff / s;                         // -d logN/d mu_q = +d logN/d mu_k
        const float fv = -0.5f / s + 0.5f * diff * diff / (s * s);
        dmq += w * fm;
        dvq -= w * fv;
        part_mu[((long long)b * K + k) * CH + d] = -scale * w * fm;
        part_var[((long long)b * K + k) * CH + d] = -scale * w * fv;
    }
    if (pass) dvq -= 0.25f / vq;                           // 0.5 * d log term3 / d var_q
    dlat[b * 2 * CH + CH + d] = scale * dmq;               // gradient w.r.t. the fc_std output (mu_q)
    dlat[b * 2 * CH + d] = scale * dvq * vq;               // gradient w.r.t. the fc_mu output (log var_q)
}

// prior gradients: fixed-order sum over samples + the term-2 part; out_mean[0] = mean_b cs[b]
__global__ void __launch_bounds__(64) cs_reduce_kernel(const float* __restrict__ part_mu, const float* __restrict__ part_var,
                                                       const float* __restrict__ cs, const float* __restrict__ ws,
                                                       const float* __restrict__ g2_mu, const float* __restrict__ g2_var,
                                                       const float* __restrict__ plv, float scale, int B, int K,
                                                       float* __restrict__ dpmu, float* __restrict__ dplv,
                                                       float* __restrict__ out_mean) {
    const int k = blockIdx.x, d = threadIdx.x;
    double sm = 0.0, sv = 0.0;
    int npass = 0;
    double csum = 0.0;
    for (int b = 0; b < B; ++b) {
        sm += part_mu[((long long)b * K + k) * CH + d];
        sv += part_var[((long long)b * K + k) * CH + d];
        // a clamped sample has w == 0 for every k, so its partials are exactly 0; count passes from cs > 0 or partial != 0
        csum += cs[b];
    }
    // number of unclamped samples: recomputed from the stored weights would need another pass; the sample kernel
    // marks clamped samples by cs[b] == 0 AND zero partials, and an unclamped sample with raw == 0 contributes a
    // term-2 gradient of measure zero -- count cs[b] > 0
    for (int b = 0; b < B; ++b) npass += cs[b] > 0.f ? 1 : 0;
    const float t2c = scale * (float)npass * 0.5f / ws[0];
    const float vk = expf(plv[k * CH + d]);
    dpmu[k * CH + d] = (float)sm + t2c * g2_mu[k * CH + d];
    dplv[k * CH + d] = ((float)sv + t2c * g2_var[k * CH + d]) * vk;
    if (k == 0 && d == 0) out_mean[0] = (float)(csum / (double)B);
}

}  // namespace crvae

using namespace crvae;

extern "C" size_t crvae_cs_div_workspace(int B, int K) {
    return (size_t)(16 + 2 * K * CH + 2 * (size_t)B * K * CH + B) * sizeof(float);
}

// lat [B,2H] = [fc_mu out | fc_std out]; prior_mu / prior_logvar [K,H].
// cs_mean[0] = mean_b D_CS(b); dlat [B,2H], dprior_mu, dprior_logvar [K,H] = gradient of scale_loss * cs_mean.
extern "C" int crvae_cs_div_fwd_bwd(const float* lat, const float* prior_mu, const float* prior_logvar, int B, int K,
                                    float scale_loss, float* cs_mean, float* dlat, float* dprior_mu, float* dprior_logvar,
                                    void* workspace, void* stream) {
    CRVAE_REQUIRE(lat && prior_mu && prior_logvar && cs_mean && dlat && dprior_mu && dprior_logvar && workspace, "null operand");
    CRVAE_REQUIRE(B > 0 && K > 0 && K <= CS_MAXK, "bad size (K <= 32)");
    float* ws = (float*)workspace;
    float* g2_mu = ws + 16;
    float* g2_var = g2_mu + K * CH;
    float* part_mu = g2_var + K * CH;
    float* part_var = part_mu + (size_t)B * K * CH;
    float* cs = part_var + (size_t)B * K * CH;
    cudaStream_t st = (cudaStream_t)stream;
    cs_prior_kernel<<<1, 64, 0, st>>>(prior_mu, prior_logvar, K, ws, g2_mu, g2_var);
    int rc = check_launch("cs_prior_kernel");
    if (rc) return rc;
    const float scale = scale_loss / (float)B;
    cs_sample_kernel<<<B, 64, 0, st>>>(lat, prior_mu, prior_logvar, K, ws, scale, cs, dlat, part_mu, part_var, B);
    if ((rc = check_launch("cs_sample_kernel"))) return rc;
    cs_reduce_kernel<<<K, 64, 0, st>>>(part_mu, part_var, cs, ws, g2_mu, g2_var, prior_logvar, scale, B, K, dprior_mu,
                                       dprior_logvar, cs_mean);
    return check_launch("cs_reduce_kernel");
}
