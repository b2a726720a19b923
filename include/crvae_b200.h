/*
 * crvae_b200.h -- C ABI of libcrvae_b200.so: the B200 (sm_100a) kernels behind the CR-VAE training
 * hot path of anonyme-Zheng/VAE-connexe (reference file: CRVAE_lorenz96.py).
 *
 * The reference has no FFI of its own (it is pure Python on top of torch.nn.GRU / nn.Linear /
 * autograd); the drop-in boundary is its Python symbol surface (CRVAE, VRAE4E, GC(), prox_update,
 * train_phase1/2 ...), mirrored by the package vae-connexe_b200/.  This header is the layer below
 * that mirror: plain device pointers + sizes + a cudaStream_t passed as void*; no torch types.
 * Each entry point cites the reference statement(s) whose arithmetic it replaces.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 unless stated otherwise; all tensors are dense
 *     row-major with the index order written in the comment;
 *   - H (hidden) must be 64 (the only size any reference driver uses, CRVAE_lorenz96.py:768);
 *     G = 3H = 192 gate rows ordered [r; z; n] as in nn.GRU.weight_*;
 *   - P = heads in this call (a head shard on multi-GPU), T = timesteps, B = batch rows,
 *     K = projection depth (= number of series p);
 *   - return value: 0 on success, otherwise a negative CRVAE_E_* code or a positive cudaError_t;
 *     crvae_last_error() returns a static description for the calling thread;
 *   - every call is asynchronous on `stream`; nothing synchronises; no global state is kept
 *     (besides the launch counter below).  Re-entrant for distinct streams/buffers.
 */
#ifndef CRVAE_B200_H
#define CRVAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CRVAE_ABI_VERSION 1
#define CRVAE_HIDDEN 64

#define CRVAE_E_BADARG   (-1)   /* unsupported size / null pointer / misaligned buffer */
#define CRVAE_E_NODEVICE (-2)   /* no sm_100 device available */

/* GEMM operand forms for crvae_gemm_f32 */
#define CRVAE_GEMM_NT 0   /* C[m,n] = sum_k A[m,k] * B[n,k]   (F.linear: x @ W^T)          */
#define CRVAE_GEMM_TN 1   /* C[m,n] = sum_k A[k,m] * B[k,n]   (weight gradient: dY^T @ X)  */
#define CRVAE_GEMM_NN 2   /* C[m,n] = sum_k A[m,k] * B[k,n]   (input gradient: dY @ W)     */

/* KL forms for crvae_latent_fwd / crvae_latent_bwd */
#define CRVAE_KL_STANDARD 0 /* -0.5*sum(1 + log_var - mu^2 - exp(log_var))                        */
#define CRVAE_KL_SWAPPED  1 /* what the reference trainers evaluate: forward returns
                               (pred, log_var, mu) (CRVAE_lorenz96.py:221) but the trainer unpacks
                               (pred, mu, log_var) (:482,:508), so the roles are exchanged:
                               -0.5*sum(1 + mu - log_var^2 - exp(mu))                             */
#define CRVAE_KL_LOGSIGMA 2 /* Family-B CR-VAE (CRVAE.py:72-75, :169): the second half of `lat` is log(sigma), not log(var):
                              z = mu + 0.5*exp(s)*eps,  KL term = -0.5*(1 + 2s - mu^2 - exp(2s))                   */

int         crvae_abi_version(void);
const char* crvae_last_error(void);
/* number of kernel launches issued through this library since load / last reset */
uint64_t    crvae_launch_count(void);
void        crvae_launch_count_reset(void);
/* 0 when the current device is compute capability 10.x, else CRVAE_E_NODEVICE */
int         crvae_check_device(void);

/* ---------------------------------------------------------------------------------------------
 * Batched fp32 GEMM (exact FFMA path).  Replaces the ATen `addmm`/`mm` calls under nn.GRU's
 * input projection (CRVAE_lorenz96.py:119, :208), nn.Linear (:210-211) and their autograd twins.
 *   C[b] (M x N, ldc) = op(A[b]) * op(B[b]) (+ bias[b][n])        b = 0..batch-1
 * strides sA/sB/sC/sBias are in elements and may be 0 (operand shared by all batches).
 * accumulate != 0 adds into C instead of overwriting.
 * ------------------------------------------------------------------------------------------- */
int crvae_gemm_f32(int form, int batch, int M, int N, int K,
                   const float* A, int lda, int64_t sA,
                   const float* B, int ldb, int64_t sB,
                   float* C, int ldc, int64_t sC,
                   const float* bias, int64_t sBias,
                   int accumulate, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Multi-head input projection  (nn.GRU's `linear_ih` for all heads and timesteps at once;
 * GRU.forward, CRVAE_lorenz96.py:115-119).
 *   gates[i][t][b][:] = b_ih[i][:] + x[t][b][:] . w_ih[i]^T          for t >= t_skip
 * x [T,B,K]; w_ih [P,G,K] (masked-dense: structural zeros for unconnected inputs);
 * b_ih [P,G]; gates [P,T,B,G].  Steps t < t_skip are known-zero inputs (the prepended zero
 * step, :205/:119) and are not written: crvae_gru_fwd substitutes b_ih for them.
 * ------------------------------------------------------------------------------------------- */
int crvae_proj_fwd(const float* x, const float* w_ih, const float* b_ih, float* gates,
                   int P, int T, int B, int K, int t_skip, void* stream);

/* Tensor-core form of crvae_proj_fwd: tcgen05.mma kind::tf32 fed by TMA, accumulators in TMEM,
 * error-compensated 3xTF32 (A.B ~= Alo.Bhi + Ahi.Blo + Ahi.Bhi, fp32 accumulate) so the result agrees
 * with the fp32 reference to ~1e-6 relative.  Operands arrive pre-split: x_hi/x_lo [T,B,K] from
 * crvae_split_tf32, w_hi/w_lo [P,G,K] from crvae_split_tf32_gate_rows (gate rows permuted inside
 * every 32-block so that the epilogue's TMEM fragments store 256-bit row segments directly).
 * Needs K % 4 == 0 (TMA row pitch); gates 32-byte aligned.                                      */
int crvae_proj_fwd_tc(const float* x_hi, const float* x_lo, const float* w_hi, const float* w_lo,
                      const float* b_ih, float* gates, int P, int T, int B, int K, int t_skip, void* stream);
/* Tensor-core form of crvae_proj_wgrad (3xTF32, MN-major UMMA operands straight from the natural
 * layouts; the gate gradients are split into tf32 hi/lo inside the kernel).  x_hi/x_lo [T,B,K].  */
size_t crvae_proj_wgrad_tc_workspace(int P, int T, int B, int K, int t_skip);
int crvae_proj_wgrad_tc(const float* dgates, const float* x_hi, const float* x_lo, const uint8_t* mask,
                        float* dw_ih, int P, int T, int B, int K, int t_skip, void* workspace, void* stream);
/* hi[i] = tf32(src[i]) (round to nearest), lo[i] = src[i] - hi[i] (exact in fp32)               */
int crvae_split_tf32(const float* src, float* hi, float* lo, int64_t n, void* stream);
/* Same split of a [rows, cols] matrix with row r written to row (r & ~31) | perm(r & 31), where
 * perm(u) = 8*((u>>1)&3) + 2*(u>>3) + (u&1): the operand layout crvae_proj_fwd_tc expects for W_ih.
 * rows % 32 == 0; not in place.                                                                 */
int crvae_split_tf32_gate_rows(const float* src, float* hi, float* lo, int64_t rows, int cols, void* stream);

/* Weight gradient of the projection (autograd of the above, :497):
 *   dw_ih[i] (G x K) = sum_{t>=t_skip,b} dgates[i][t][b][:]^T x[t][b][:]   (x mask[i][k] if mask)
 * workspace: >= crvae_proj_wgrad_workspace(P,T,B,K) bytes (split-reduction partials).          */
size_t crvae_proj_wgrad_workspace(int P, int T, int B, int K);
int crvae_proj_wgrad(const float* dgates, const float* x, const uint8_t* mask, float* dw_ih,
                     int P, int T, int B, int K, int t_skip, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Persistent multi-head GRU recurrence, forward (ATen's per-step `linear_hh` + gate math under
 * nn.GRU, :119/:208, fused with the per-head Linear(H,1), :120).
 *   for t: gh = h W_hh^T + b_hh; r = sig(gi_r+gh_r); z = sig(gi_z+gh_z);
 *          n = tanh(gi_n + r*gh_n); h = (h - n)*z + n          (this operation order)
 * gates [P,T,B,G]  in: gi (steps < t_skip: ignored, b_ih used)   out: r | z | n  (in place)
 * h0 [B,H] shared by all heads (h0_head_stride = 0) or [P,B,H] (h0_head_stride = B*H)
 * hs [P,T,B,H] = h_1..h_T;  ghn [P,T,B,H] = gh_n (kept for the backward);
 * pred [P,T,B] = h_t . w_lin[i] + b_lin[i]   (w_lin/b_lin/pred may be NULL: encoder GRU).
 * ------------------------------------------------------------------------------------------- */
int crvae_gru_fwd(float* gates, const float* b_ih, const float* w_hh, const float* b_hh,
                  const float* h0, int64_t h0_head_stride,
                  const float* w_lin, const float* b_lin,
                  float* hs, float* ghn, float* pred,
                  int P, int T, int B, int t_skip, void* stream);

/* Tensor-core form of crvae_gru_fwd (same buffers and results to fp32 rounding): the per-step gate
 * GEMM h.W_hh^T runs on tcgen05 (3xTF32; h operand and accumulator in TMEM), W_hh stays resident in
 * shared memory for all timesteps.  w_hh_hi / w_hh_lo = crvae_split_tf32(W_hh) ([P,G,H] each), or
 * w_hh_hi = the fp32 W_hh itself and w_hh_lo = NULL (split while staging).  gates / hs / ghn must be
 * 32-byte aligned (256-bit vector accesses).                                                       */
int crvae_gru_fwd_tc(float* gates, const float* b_ih, const float* w_hh_hi, const float* w_hh_lo,
                     const float* b_hh, const float* h0, int64_t h0_head_stride,
                     const float* w_lin, const float* b_lin, float* hs, float* ghn, float* pred,
                     int P, int T, int B, int t_skip, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Hand-written BPTT for the above (replaces autograd through :119-:120, triggered at :497).
 * gates [P,T,B,G]  in: r | z | n   out: dgi = [da_r | da_z | da_n]  (in place)
 * dpred [P,T,B] (NULL if no output head);  dh_last [P,B,H] gradient flowing into h_T (NULL = 0);
 * dhs [P,T,B,H] extra gradient w.r.t. every h_t (NULL = 0; VRAE4E's Linear(H,p) head, :167)
 * outputs: dw_hh [P,G,H], db_hh [P,G], db_ih [P,G], dw_lin [P,H], db_lin [P] (NULL ok with
 *          w_lin NULL), dh0 [P,B,H].   workspace >= crvae_gru_bwd_workspace(P,B) bytes.
 * ------------------------------------------------------------------------------------------- */
size_t crvae_gru_bwd_workspace(int P, int B);
int crvae_gru_bwd(float* gates, const float* ghn, const float* hs,
                  const float* h0, int64_t h0_head_stride,
                  const float* w_hh, const float* w_lin,
                  const float* dpred, const float* dh_last, const float* dhs,
                  float* dw_hh, float* db_hh, float* db_ih, float* dw_lin, float* db_lin,
                  float* dh0, int P, int T, int B, void* workspace, void* stream);

/* BPTT with the recurrent weight gradient deferred to the tensor cores: identical to crvae_gru_bwd except that
 * dw_hh is NOT produced and `ghn` is overwritten in place with dgh_n = da_n*r; crvae_gru_dwhh_tc then computes
 *   dw_hh[i][g][k] = sum_{t,b} dgh[i][t][b][g] * h_{t-1}[i][b][k]
 * as one tcgen05 GEMM per head (3xTF32, MN-major operands read in place, tf32 split in shared memory).
 * Needs B % 32 == 0.  Halves the FFMA work of the BPTT kernel and lets two of its CTAs share an SM.
 * Both tensor-core gradient GEMMs cut a head's reduction range into several CTAs when a rank holds few heads
 * (head shards on multi-GPU) and sum the partials in fixed order; `workspace` >= the *_workspace() size.        */
int crvae_gru_bwd_deferred(float* gates, float* ghn, const float* hs, const float* h0, int64_t h0_head_stride,
                           const float* w_hh, const float* w_lin, const float* dpred, const float* dh_last,
                           const float* dhs, float* db_hh, float* db_ih, float* dw_lin, float* db_lin, float* dh0,
                           int P, int T, int B, void* workspace, void* stream);

/* Tensor-core form of crvae_gru_bwd_deferred without the dhs input (same buffers, results to fp32 rounding):
 * the per-step product dh_{t-1} = dh_t*z + dgh.W_hh runs on tcgen05 (3xTF32; dgh operand and accumulator in
 * TMEM, W_hh resident in shared memory), r|z|n tiles arrive by TMA, column sums are register accumulators.
 * One CTA per (head, 128-row tile).  gates / ghn / hs / h0 / dh0 / dh_last must be 32-byte aligned.
 * Replaces autograd through nn.GRU + nn.Linear(H,1) (CRVAE_lorenz96.py:497).                                  */
int crvae_gru_bwd_tc(float* gates, float* ghn, const float* hs, const float* h0, int64_t h0_head_stride,
                     const float* w_hh, const float* w_lin, const float* dpred, const float* dh_last,
                     float* db_hh, float* db_ih, float* dw_lin, float* db_lin, float* dh0, int P, int T, int B,
                     void* workspace, void* stream);

/* Low-latency forms of crvae_gru_fwd / crvae_gru_bwd_deferred (same arguments, same results in exact fp32) for the
 * shapes whose cost is the latency of ONE recurrent step rather than bandwidth: small head shards (p = 100 over 8 GPUs
 * = 12-13 heads), the replicated encoder (CRVAE_lorenz96.py:208), VRAE4E (:155, :166) and the long sequences of
 * VRAE.py.  One CTA per (head, 16-row tile); W_hh resident in shared memory; the step's gate / h / gh_n slabs move by
 * cp.async.bulk through a shared-memory ring (in place: gi -> r|z|n, r|z|n -> dgi), so nothing of the step sits in
 * the load/store queue; packed fp32 FMAs.  crvae_gru_bwd_ll leaves dw_hh to crvae_gru_dwhh_tc.                       */
int crvae_gru_fwd_ll(float* gates, const float* b_ih, const float* w_hh, const float* b_hh,
                     const float* h0, int64_t h0_head_stride, const float* w_lin, const float* b_lin,
                     float* hs, float* ghn, float* pred, int P, int T, int B, int t_skip, void* stream);
int crvae_gru_bwd_ll(float* gates, float* ghn, const float* hs, const float* h0, int64_t h0_head_stride,
                     const float* w_hh, const float* w_lin, const float* dpred, const float* dh_last,
                     const float* dhs, float* db_hh, float* db_ih, float* dw_lin, float* db_lin, float* dh0,
                     int P, int T, int B, void* workspace, void* stream);
/* Warp-level tensor-core forms of the two calls above (mma.sync; same arguments, buffers and in-place conventions; results
 * agree with the exact kernels to ~2e-6, like the *_tc kernels).  Warp-specialised: 8 MMA warps hold W_hh in registers as
 * pre-split B fragments, 8 gate-math warps do the cell math and the staging; (head, 16-row tile)s in a persistent grid, two
 * tiles of a head per CTA once there are more tiles than SMs; every global access is a TMA tile copy (4-D tensor maps over
 * gates / hs / ghn / h0 / dhs, rows past B clipped).  Forward: fp16 hi|lo operands on m16n8k16 (tf32 hi|lo with
 * CRVAE_MMA_F16=0); BPTT: tf32 hi|lo, product split over K, partial products summed in fixed order (deterministic).
 * The default recurrent path of the shapes bound by the latency of one step: head shards up to 26 heads, the encoder,
 * VRAE4E, VRAE.py (vae-connexe_b200/rec.py); the BPTT also of p = 100 on one rank when run as one launch.
 * All row pointers 16-byte aligned; `workspace` >= crvae_gru_bwd_workspace(P, B).                                       */
int crvae_gru_fwd_mma(float* gates, const float* b_ih, const float* w_hh, const float* b_hh,
                      const float* h0, int64_t h0_head_stride, const float* w_lin, const float* b_lin,
                      float* hs, float* ghn, float* pred, int P, int T, int B, int t_skip, void* stream);
int crvae_gru_bwd_mma(float* gates, float* ghn, const float* hs, const float* h0, int64_t h0_head_stride,
                      const float* w_hh, const float* w_lin, const float* dpred, const float* dh_last,
                      const float* dhs, float* db_hh, float* db_ih, float* dw_lin, float* db_lin, float* dh0,
                      int P, int T, int B, void* workspace, void* stream);
size_t crvae_gru_dwhh_tc_workspace(int P, int T, int B);
int crvae_gru_dwhh_tc(const float* dgates, const float* dghn, const float* hs, const float* h0,
                      int64_t h0_head_stride, float* dw_hh, int P, int T, int B, void* workspace, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused reparameterisation + KL  (CRVAE.forward :210-216, VRAE4E.forward :157-163, trainer :486)
 *   lat [B,2Z] = [mu | log_var] (output of the fc_mu|fc_std GEMM);  eps [B,Z] ~ N(0,1)
 *   (Z = latent width: H for CRVAE / VRAE4E, free for the generic VRAE of VRAE.py:105-147)
 *   z [B,Z] = mu + exp(0.5*log_var)*eps
 *   kl_out[0] = mean_b sum_h KL-term (form: CRVAE_KL_*)                 (single deterministic sum)
 * ------------------------------------------------------------------------------------------- */
int crvae_latent_fwd(const float* lat, const float* eps, float* z, float* kl_out,
                     int B, int Z, int kl_form, void* stream);

/* Backward of the above + the sum over heads of dh0 (every head's h0 is z, :218):
 *   dz[b][h] = sum_{i<P} dh0[i][b][h]  (+ dz_extra[b][h] if not NULL; a peer-reduced partial)
 *   dlat[b][0:H]  = dz + beta * dKL/dmu ;  dlat[b][H:2H] = dz*eps*0.5*exp(0.5*log_var) + beta * dKL/dlog_var
 * dh0 may be NULL with P = 0 (then dz = dz_extra).  dz_out (NULL ok) receives the head sum.     */
int crvae_latent_bwd(const float* dh0, int P, const float* dz_extra, const float* lat,
                     const float* eps, float beta, int kl_form, float* dlat, float* dz_out,
                     int B, int Z, void* stream);

/* The same latent head for H = Z = 64 (CR-VAE, VRAE4E) fused with its two Linear layers, one launch each:
 *   fwd: lat = hT . lat_w^T + lat_b (fc_mu | fc_std, :210-211), z, KL as crvae_latent_fwd
 *   bwd: d_lat_w = dlat^T . hT, d_lat_b = column sums of dlat, dhT = dlat . lat_w          (autograd, :497)
 * hT [B,64], lat_w [128,64], lat_b [128], lat/dlat [B,128].  workspace: crvae_latent_head_workspace(B) bytes,
 * zero-filled once by the caller (per-CTA KL partials + a completion counter the kernel resets itself).      */
size_t crvae_latent_head_workspace(int B);
int crvae_latent_head_fwd(const float* hT, const float* lat_w, const float* lat_b, const float* eps, float* lat,
                          float* z, float* kl_out, int B, int kl_form, void* workspace, void* stream);
int crvae_latent_head_bwd(const float* dlat, const float* hT, const float* lat_w, float* d_lat_w, float* d_lat_b,
                          float* dhT, int B, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused MSE loss forward + backward for all heads (trainer :484, :509; nn.MSELoss 'mean'):
 *   sse[i]         = sum_{t,b} (pred[i][t][b] - target[i][t][b])^2     (loss = sum_i sse[i]/(T*B))
 *   dpred[i][t][b] = dscale*(pred - target), dscale <= 0 meaning 2/(T*B) (nn.MSELoss 'mean'); VRAE.py's
 *                    sum-reduced loss / batch (VRAE.py:143) passes dscale = 2/batch
 *   err[i][t][b]   = target - pred          (NULL ok; the phase-2 residual, :599/:639)
 * pred, target, dpred, err [P,T,B]; target is X[:, 10:, i] of the fixed batch, head-major.
 * ------------------------------------------------------------------------------------------- */
int crvae_mse_fwd_bwd(const float* pred, const float* target, float* sse, float* dpred, float* err,
                      int P, int T, int B, float dscale, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Plain gradient step  theta <- theta - lr*grad  (:498-499), product rounded to fp32 first as
 * torch does (`lr * param.grad` then `-=`), no FMA contraction.
 * ------------------------------------------------------------------------------------------- */
int crvae_gd_step(float* theta, const float* grad, int64_t n, float lr, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused gradient step + group-lasso proximal update + GC norms on the first-layer weights
 * (:498-499 restricted to weight_ih_l0, then prox_update :308-314, then what GC() :297 reads):
 *   W <- W - lr*dW ;  nu_j = ||W[:,j]||_2 ;  W[:,j] <- (W[:,j] / max(nu_j, thr)) * max(nu_j - thr, 0)
 *   col_norm[i][j] = || updated W[i][:,j] ||_2          (thr = fp32(lam*lr); lam == 0: no prox)
 * w_ih, dw_ih [P,G,K]; mask [P,K] u8 or NULL (masked-out columns stay exactly 0).
 * dw_ih may be NULL (prox only, the stand-alone prox_update()).
 * ------------------------------------------------------------------------------------------- */
int crvae_gd_prox_gc(float* w_ih, const float* dw_ih, const uint8_t* mask, float* col_norm,
                     int P, int K, float lr, float thr, int do_prox, void* stream);

/* Adam step with torch.optim.Adam default semantics (:565, :612-614); step counts from 1.       */
int crvae_adam_step(float* theta, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                    double lr, double beta1, double beta2, double eps, int step, void* stream);

/* y = tanh(x) and its backward dx = dy*(1-y^2): VRAE4E's z = tanh(linear_hidden(z)) (:164)         */
int crvae_tanh_fwd(const float* x, float* y, int64_t n, void* stream);
int crvae_tanh_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream);
/* Output activation of the generic VRAE decoder (VRAE.py:60-68) and its backward dx = dy * act'(y);
 * kind: 0 tanh, 1 sigmoid, 2 relu, 3 identity.                                                       */
int crvae_act_fwd(const float* x, float* y, int64_t n, int kind, void* stream);
int crvae_act_bwd(const float* dy, const float* y, float* dx, int64_t n, int kind, void* stream);
/* out[c][r] = in[r][c]: residual [P][T*B] (head-major) <-> [T*B][P] (VRAE4E input, :599/:639)       */
int crvae_transpose(const float* in, float* out, int rows, int cols, void* stream);
/* Gather-packed ragged heads (phase 2: head i = GRU(k_i, H) on the k_i series selected by column i of the connection
 * matrix, CRVAE_lorenz96.py:115, :200-201, :788-790).  w_ih [P,G,Kp] packed (Kp = widest head's input count rounded up
 * to 4), cols [P,Kp] int32 = series index of every packed column (ascending), mask [P,Kp] u8 = 0 on padding columns.
 *   crvae_gather_cols        xg[i][r][c] = mask[i][c] ? x[r][cols[i][c]] : 0         (x [rows,K] -> xg [P,rows,Kp])
 *   crvae_proj_fwd_packed    gates[i] = b_ih[i] + xg[i] . w_ih[i]^T                  (exact fp32)
 *   crvae_proj_wgrad_packed  dw_ih[i] = dgates[i]^T . xg[i], padding columns zeroed  (exact fp32)
 * Results equal the masked-dense form (crvae_proj_fwd / crvae_proj_wgrad on [P,G,K_dense] with structural zeros) BIT
 * FOR BIT: ascending packed columns = the dense sum without its exact zeros; the gradient's reduction is cut into
 * the number of splits the dense form uses for K_dense series.                                                   */
int crvae_gather_cols(const float* x, const int* cols, const uint8_t* mask, float* xg, int P, int64_t rows, int K,
                      int Kp, void* stream);
int crvae_proj_fwd_packed(const float* xg, const float* w_ih, const float* b_ih, float* gates,
                          int P, int T, int B, int Kp, int t_skip, void* stream);
size_t crvae_proj_wgrad_packed_workspace(int P, int T, int B, int Kp, int K_dense);
int crvae_proj_wgrad_packed(const float* dgates, const float* xg, const uint8_t* mask, float* dw_ih,
                            int P, int T, int B, int Kp, int K_dense, int t_skip, void* workspace, void* stream);

/* Head-sharded training, the one data-path collective: dz = sum over ALL heads of dh0, needed by the
 * replicated encoder.  One kernel = local head sum + one-shot all-reduce over NVLink peer memory + the latent backward
 * (crvae_latent_bwd's arithmetic) in its epilogue.  peer_bufs[r] = rank r's SYMMETRIC buffer of crvae_dz_allreduce_bytes()
 * bytes as mapped into THIS process (e.g. torch.distributed._symmetric_memory rendezvous: buffer_ptrs); the buffers must
 * be zero before the first call on any rank.  Partials are summed in rank order on every rank: bit-identical results
 * everywhere.  Re-launchable back to back (epochs + two slots), CUDA-graph capturable, no NCCL involved.            */
size_t crvae_dz_allreduce_bytes(int B, int Z, int world);
int crvae_dz_allreduce_latent_bwd(const float* dh0, int P, void* const* peer_bufs, int rank, int world, const float* lat,
                                  const float* eps, float beta, int kl_form, float* dlat, float* dz_out, int B, int Z,
                                  void* stream);

/* MixtureCSRAE (CSRAE_new.py:113-150), the Bernoulli reconstruction term: sum_out[0] = sum over n elements of
 * binary_cross_entropy_with_logits(logits, x); dlogits (optional) = (sigmoid(logits) - x) * dscale.  workspace >=
 * crvae_bce_logits_workspace(n) bytes (per-CTA fp64 partials, added in CTA order).                                  */
size_t crvae_bce_logits_workspace(int64_t n);
int crvae_bce_logits_fwd_bwd(const float* logits, const float* x, float* sum_out, float* dlogits, int64_t n, float dscale,
                             void* workspace, void* stream);

/* Family-B CR-VAE (CRVAE.py:134-150): ISTA step on the per-head input maps W_in[i] (D x H), one group per ROW
 * (= one candidate parent series):  W_tmp = W - lr*dW;  W <- W_tmp * max(1 - thr/||W_tmp[row,:]||_2, 0)  with thr = lr*lambda
 * (a zero row stays zero, as in the reference: 1 - thr/0 = -inf -> 0).  w / dw [rows, cols] row-major; row_norm [rows]
 * receives ||W[row,:]||_2 of the RESULT (what granger_matrix thresholds, :126-131).  dw = NULL: no gradient step;
 * do_prox = 0: norms only.                                                                                          */
int crvae_ista_rows(float* w, const float* dw, float* row_norm, int64_t rows, int cols, float lr, float thr, int do_prox,
                    void* stream);

/* Test-mode generation (CRVAE.forward(mode='test'), CRVAE_lorenz96.py:223-243 / :264-284), the step between two
 * recurrent updates: every head's next input is the vector of ALL heads' outputs (:232-236).  y [R][W][B] = the step's
 * outputs gathered over R head shards of at most W heads (balanced contiguous partition: the first `rem` shards hold
 * base+1 heads, the others base; one GPU: R = 1, base = W = p, rem = 0).  Writes x_next [B][p] = y (+ scale*noise[b][t][j],
 * phase 1, :281-283), the stored sequence out [B][steps][p] at step t, and (optional) the tf32 hi / lo split of x_next.   */
int crvae_gen_scatter(const float* y, const float* noise, float* x, float* x_hi, float* x_lo, float* out, int B, int p,
                      int t, int steps, int base, int rem, int widest, float scale, void* stream);

/* Same update with the step count in device memory (step = *step_counter + 1, then incremented):
 * lets the whole phase-2 iteration be replayed from a CUDA graph.                                  */
int crvae_adam_step_dev(float* theta, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                        double lr, double beta1, double beta2, double eps, int* step_counter, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Cauchy-Schwarz divergence of the Gaussian posterior against a learnable equal-weight GMM prior,
 * forward + backward (CR-CS-RAE.py:124-163 gaussian_overlap / cs_divergence_gmm; trainer :568-582):
 *   cs_mean[0] = mean_b clamp(-log mean_k N(mu_q; mu_k, var_q+var_k) + 0.5 log mean_kk' N(mu_k; mu_k', var_k+var_k')
 *                             + 0.5 log N(mu_q; mu_q, 2 var_q), min=0)     (exp -> mean -> log, as the reference)
 * lat [B,2H] = [fc_mu out | fc_std out]; with the reference trainer's swapped unpacking (:563, :570-571)
 * mu_q = lat[:,H:2H] and var_q = exp(lat[:,0:H]).  prior_mu / prior_logvar [K,H], K <= 32.
 * dlat, dprior_mu, dprior_logvar = gradient of scale_loss*cs_mean (scale_loss = lambda_cs).
 * ------------------------------------------------------------------------------------------- */
size_t crvae_cs_div_workspace(int B, int K);
int crvae_cs_div_fwd_bwd(const float* lat, const float* prior_mu, const float* prior_logvar, int B, int K,
                         float scale_loss, float* cs_mean, float* dlat, float* dprior_mu, float* dprior_logvar,
                         void* workspace, void* stream);

/* Ridge penalty pieces (ridge_regularize :321-325): out[0] = sum(x^2) over n elements.          */
int crvae_sumsq(const float* x, int64_t n, float* out, void* stream);
/* out[0] = sum_i scale[i] * x[i] for n <= 4096 values (loss = sum_i sse[i]/(T*B) and friends)     */
int crvae_dot_small(const float* x, int n, float scale, float* out, void* stream);
/* test / tuning hook: force the recurrent kernels' batch tile (16, 32 or 64 rows; 0 = heuristic) */
void crvae_debug_set_batch_tile(int rows);
/* grad += 2*lam_ridge*theta (gradient of the ridge term in `smooth`, :488/:513-515)             */
int crvae_axpy(float* y, const float* x, int64_t n, float alpha, void* stream);

/* Batch binding (the layouts arrange_input's windows are streamed in; :208 encoder input X[:,0:Te], :119 decoder input
 * [0, X[:,Te:Te+Td-1]], :484 per-head targets X[:,Te:,i]) in one pass:
 *   X [B,Te+Td,p] -> enc_in [Te,B,p], dec_in [Td,B,p] (step 0 is left untouched = zeros), optional tf32 hi|lo splits of
 *   both (NULL, NULL to skip), target [P,Td,B] for heads head_lo .. head_lo+P-1.                                        */
int crvae_bind_batch(const float* X, float* enc_in, float* enc_hi, float* enc_lo, float* dec_in, float* dec_hi,
                     float* dec_lo, float* target, int B, int p, int Te, int Td, int head_lo, int P, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CRVAE_B200_H */
