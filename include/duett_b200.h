/*
 * duett_b200 — C ABI of the B200-native DuETT hot path (sm_100a).
 *
 * This header is the drop-in boundary.  The reference
 * (lastdancewithyou/multimodal_edema_prediction) has no FFI: its hot path is reached through Python
 * nn.Module calls that bottom out in stock ATen/cuBLAS kernels.  Each entry point below therefore cites
 * the reference Python site whose arithmetic it replaces; the Python host layer
 * (multimodal_edema_prediction_b200/) binds these symbols with ctypes and keeps the reference's
 * nn.Module / trainer surface.  See INTEGRATION.md for the binding a maintainer would add.
 *
 * Conventions
 *  - every function returns 0 on success, <0 on error (dx_last_error() gives the message);
 *  - all pointers are DEVICE pointers unless the name ends in _host; no allocation, no host sync,
 *    re-entrant per stream; `stream` is a cudaStream_t passed as void*;
 *  - dtype codes: DX_F32 = 0, DX_BF16 = 1.  "act dtype" is the storage type of activations
 *    (bf16 in the bf16 mode, f32 in the fp32 mode); statistics, parameters and gradients are f32.
 */
#ifndef DUETT_B200_H
#define DUETT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DX_F32 0
#define DX_BF16 1

/* activation codes for dx_gemm_desc.act */
#define DX_ACT_NONE 0
#define DX_ACT_GELU 1     /* out = gelu(pre), out2 = pre                     (x_transformers FeedForward, duett/duett.py:95-105) */
#define DX_ACT_RELU 2     /* out = relu(pre)                                  (simple_mlp, duett/duett.py:24-39) */
#define DX_ACT_TANH 3     /* out = tanh(pre)                                  (cve, duett/duett.py:151-157) */
#define DX_ACT_GELU_BWD 4 /* v = pre * gelu'(aux); out2 = v; out = v*row_scale2; row_dot += v*(aux-aux_bias) */
#define DX_ACT_RELU_BWD 5 /* v = pre * (aux > 0) */
#define DX_ACT_TANH_BWD 6 /* v = pre * (1 - aux^2), aux = tanh output */

const char* dx_last_error(void);
int dx_version(void);
/* 1 when the running device is sm_100 (tcgen05/TMEM/TMA paths usable) */
int dx_device_ok(void);

/*
 * Fused GEMM:  acc[m,n] = sum_k A(m,k) * B(n,k)     (i.e. A @ B^T, the nn.Linear contraction)
 *   a_mn = 0: A(m,k) = A[m*lda + k] (K-major)   a_mn = 1: A(m,k) = A[k*lda + m] (MN-major)
 *   b_mn = 0: B(n,k) = B[n*ldb + k]             b_mn = 1: B(n,k) = B[k*ldb + n]
 * in_dtype DX_BF16 -> tcgen05.mma kind::f16 with TMEM fp32 accumulators, TMA-fed (dx_gemm_tc.cu);
 * in_dtype DX_F32  -> fp32 FFMA kernel (dx_gemm_simt.cu), used by the fp32 precision mode.
 * Epilogue (all optional, null pointer = skipped), in this order:
 *   v = acc * row_scale[m] + bias[n];  act (see DX_ACT_*);  v += res[m,n];
 *   v -= cx[m,n] * coef_num[m] / max(coef_den[m], 1e-24);   (ScaleNorm backward projection)
 *   row_sumsq[m] += sum_n v^2;  out[m,n] (=|+=) v
 * Replaces: every nn.Linear / x_transformers projection on the path — duett/duett.py:95-105 (Encoder
 * to_q/to_k/to_v/to_out/ff), :84-86,106,110-125 (MLPs), models/main_architecture_duett.py:566,1027,1216-1219.
 */
typedef struct dx_gemm_desc {
  int32_t M, N, K;
  int32_t in_dtype;   /* dtype of A and B */
  int32_t a_mn, b_mn;
  const void* A; int64_t lda;
  const void* B; int64_t ldb;
  void* out; int64_t ldo; int32_t out_dtype; int32_t accumulate; /* accumulate: out += v (f32 out only) */
  void* out2; int64_t ldo2;          /* act dtype */
  int32_t act; int32_t act_dtype;    /* act dtype = dtype of out2/res/cx/aux */
  const float* row_scale;            /* [M] */
  const float* row_scale2;           /* [M] (DX_ACT_GELU_BWD only) */
  const float* bias;                 /* [N] */
  const void* res; int64_t ldr;      /* [M,N] residual */
  const void* aux; int64_t ldx;      /* [M,N] saved pre-activation / activation for *_BWD */
  const float* aux_bias;             /* [N] subtracted from aux in row_dot */
  const void* cx; int64_t ldc;       /* [M,N] */
  const float* coef_num; const float* coef_den; /* [M] */
  float* row_sumsq;                  /* [M] atomically accumulated */
  float* row_dot;                    /* [M] atomically accumulated (DX_ACT_GELU_BWD) */
  int32_t force_simt;                /* test hook: run the FFMA kernel even for bf16 inputs */
  int32_t batch;                     /* grouped mode: number of independent GEMMs of identical shape (0/1 = single) */
  /* element strides between consecutive batches (grouped mode; the per-variable embedding MLPs of duett/duett.py:84-86) */
  int64_t a_bs, b_bs, out_bs, out2_bs, res_bs, aux_bs, cx_bs;
  int32_t bias_bs;                   /* stride of bias / aux_bias between batches */
  int32_t rowvec_bs;                 /* stride of the [M] row vectors (row_scale, coef_num, row_sumsq, ...) between batches */
  int32_t split_k;                   /* FFMA kernel only: >1 splits K over blockIdx.z; needs accumulate into a zeroed f32 out,
                                        plain epilogue (bias allowed).  The tcgen05 kernel picks its own split for dW GEMMs. */
  int32_t allow_tf32;                /* fp32 operands: 1 = contract on the tensor cores (tcgen05 kind::tf32, fp32 accumulate) — the
                                        precision of the reference's SSL / fine-tune path (set_float32_matmul_precision('high'),
                                        duett/duett.py:9); 0 = exact fp32 FFMA products (the 1e-3 parity mode) */
} dx_gemm_desc;

int dx_gemm(const dx_gemm_desc* d, void* stream);

/* Data-parallel runs: the persistent tcgen05 grids leave n SMs (even) free for the NCCL kernels of the gradient all-reduce that
 * overlaps the backward pass (training_duett/trainer.py:217-218 / Lightning DDP bucket all-reduce).  Returns the previous n.
 * Environment default: DX_GEMM_SM_RESERVE. */
int dx_gemm_reserve_sms(int n);

/*
 * Test hook for bring-up: same as dx_gemm on the tcgen05 path but with the shared-memory
 * matrix-descriptor fields overridden (lbo/sbo in bytes for A and B; <0 keeps the built-in value).
 */
int dx_gemm_tc_debug(const dx_gemm_desc* d, int32_t block_n, int32_t stages, int32_t a_lbo, int32_t a_sbo,
                     int32_t b_lbo, int32_t b_sbo, void* stream);

/* ---------------------------------------------------------------------------------------------------------------
 * T<->V axis re-layout fused with ScaleNorm bookkeeping.  src is [B,P,Q,d] (tokens (b,p)), dst is [B,Q,P,d].
 *   fwd: dst[b,q,p,:] = src[b,p,q,:] * (sqrt(Q*d)*g[0]/||src row (b,p)||) + pos_bcast[q,p,:] + pos_batched[b,q,p,:];
 *        dst_rowsq[b*Q+q] = ||dst row||^2.  src_rowsq==NULL: no scaling.  pos_* may be NULL.
 *   bwd: dsrc = s*(gy - src*<src,gy>/||src||^2) with gy the transposed gdst; *dg += sum <gy,src>*c/||src||.
 * Replaces: psi.transpose(1,2).flatten(2) + full_event_embedding.weight, event_outs.flatten(2) + time_embeddings
 * (duett/duett.py:274-279; models/main_architecture_duett.py:71-91) and the x_transformers final/pre ScaleNorm
 * reductions around them.  d % 8 == 0. */
int dx_relayout_fwd(const void* src, const float* src_rowsq, const float* g, const float* pos_bcast,
                    const void* pos_batched, void* dst, float* dst_rowsq, int B, int P, int Q, int d, int act_dtype,
                    void* stream);
int dx_relayout_bwd(const void* gdst, const void* src, const float* src_rowsq, const float* g, void* dsrc, float* dg,
                    int B, int P, int Q, int d, int act_dtype, void* stream);

/* out[n] (+)= sum_m X[m*ld+n]  — bias / positional-embedding gradients (autograd of nn.Linear bias, duett.py:275). */
int dx_colsum(const void* X, int64_t ld, int M, int64_t N, float* out, int accumulate, int dtype, void* stream);
/* y (+)= alpha*x over n elements (n % 8 == 0) — accumulation of per-layer time-embedding gradients. */
int dx_axpy(const void* x, void* y, int64_t n, float alpha, int accumulate, int dtype, void* stream);
/* dtype conversion of a dense buffer (fp32 master weights -> bf16 operands; autocast's casts). */
/* y = xs[0] + ... + xs[count-1] (1 <= count <= 8 tensors of n elements, n % 8 == 0), summed in fp32 and rounded once: the
 * gradient of the time embedding is the sum of the time encoders' input gradients of all layers (autograd accumulation of
 * `time_embeddings` at duett/duett.py:278). */
int dx_sum_n(const void* const* xs, int count, void* y, int64_t n, int dtype, void* stream);
int dx_cast(const void* x, int x_dtype, void* y, int y_dtype, int64_t n, void* stream);
/* out[i] = c*g[0]/max(sqrt(rowsq[i]),1e-12): x_transformers ScaleNorm row scale (duett/duett.py:95-105). */
int dx_scalenorm_scale(const float* rowsq, const float* g, float c, float* out, int N, void* stream);
/* rowdot[n] = <a[n,:], g[n,:]>; g[n,:] *= row_scale[n] (ScaleNorm backward bookkeeping on [N,C] tensors). */
int dx_rowdot_scale(const void* a, void* g, const float* row_scale, float* rowdot, int N, int C, int dtype, void* stream);
/* sink[0] += sum(x)/g[0]: ScaleNorm gain gradient. */
int dx_sum_div_acc(const float* x, int64_t n, const float* g, float* sink, void* stream);

/* Unmasked multi-head attention core dropout(softmax(q k^T / sqrt(dh))) v, fp32 softmax.  Element (b,s,h,i) of a tensor
 * lives at ptr + b*bs + s*rs + h*dh + i (strides in elements).  lse: [B,H,Sq] f32.  dh in {4,8,12,16,32,64,128}.
 * Dropout on the attention probabilities (attn_dropout of the x_transformers Encoder, duett/duett.py:98,104; dropout of
 * nn.MultiheadAttention in _PerceiverBlock): drop_p in [0,1); probability (b,h,q,k) is kept iff
 * splitmix64(seed + idx * 0x9E3779B97F4A7C15) >> 32 >= drop_p * 2^32 with idx = ((b*H+h)*Sq+q)*Sk+k and
 * seed = drop_seed + (drop_seed_dev ? *drop_seed_dev : 0), and scaled by 1/(1-drop_p); the backward call regenerates the
 * mask from the same (drop_p, seed).  drop_p = 0 disables it.
 * Replaces: x_transformers Attention core (duett/duett.py:95-105 call sites :276,:279) and the nn.MultiheadAttention core
 * of _PerceiverBlock (models/main_architecture_duett.py:752,759). */
int dx_attn_fwd(const void* q, int64_t q_bs, int64_t q_rs, const void* k, int64_t k_bs, int64_t k_rs, const void* v,
                int64_t v_bs, int64_t v_rs, void* o, int64_t o_bs, int64_t o_rs, float* lse, int B, int H, int Sq, int Sk,
                int dh, int dtype, float drop_p, uint64_t drop_seed, const uint64_t* drop_seed_dev, void* stream);
int dx_attn_bwd(const void* q, int64_t q_bs, int64_t q_rs, const void* k, int64_t k_bs, int64_t k_rs, const void* v,
                int64_t v_bs, int64_t v_rs, const void* o, int64_t o_bs, int64_t o_rs, const void* go, int64_t go_bs,
                int64_t go_rs, void* dq, int64_t dq_bs, int64_t dq_rs, void* dk, int64_t dk_bs, int64_t dk_rs, void* dv,
                int64_t dv_bs, int64_t dv_rs, const float* lse, float* D_ws, int B, int H, int Sq, int Sk, int dh,
                int dtype, float drop_p, uint64_t drop_seed, const uint64_t* drop_seed_dev, void* stream);

/* Head-averaged attention probabilities out[b,i,j] = 1/H sum_h exp(<q_bih, k_bjh>/sqrt(dh) - lse[b,h,i]), out [B,Sq,Sk] f32,
 * from the q / k projections and the lse of a dropout-free dx_attn_fwd call (same addressing as dx_attn_fwd).
 * Replaces: the attention-weights return of nn.MultiheadAttention(need_weights=True, average_attn_weights=True) in
 * _PerceiverBlock.forward(return_attn=True) (models/main_architecture_duett.py:762,772), surfaced as img_attn / ts_attn by
 * PatchDualPathologyPerceiver.forward (:621-631,651-653) and TeacherModel.forward (:1123-1128) for visualisation. */
int dx_attn_probs_mean(const void* q, int64_t q_bs, int64_t q_rs, const void* k, int64_t k_bs, int64_t k_rs, const float* lse,
                       float* out, int B, int H, int Sq, int Sk, int dh, int dtype, void* stream);

/* Value/count embedding into psi[B,T+1,V+1,d] (duett/duett.py:245-266 == models/main_architecture_duett.py:31-65, the
 * per-variable Python loop): count lookup (n_obs_embedding, clip 0..15), V grouped MLPs
 * Linear(2,64)-ReLU-BatchNorm(batch stats over B*T)-Linear(64,d), static column, [REP] row, MASK substitution.
 * The 64->d contraction of all variables is ONE grouped dx_gemm (batch = V) writing into the strided psi view; these
 * entry points are the pieces around it.  xs: [B,T,2V+1] f32 (values | counts | masked-step flag); stacked parameters
 * W0 [V,64,2], b0/gamma/beta/run_* [V,64], nobs [16], special [8,d], tab [B,d] (tab_encoder output);
 * hn / dhn: [V, B*(T+1), 64] in the act dtype. */
int dx_embed_stats(const float* xs, int B, int T, int V, const float* W0, const float* b0, const float* nobs, float* run_mean,
                   float* run_var, double* stats_ws, float* mean, float* rstd, int training, void* stream);
int dx_embed_hidden(const float* xs, int B, int T, int V, const float* W0, const float* b0, const float* nobs,
                    const float* gamma, const float* beta, const float* mean, const float* rstd, void* hn, int act_dtype,
                    void* stream);
int dx_embed_special(const float* xs, int B, int T, int V, int d, const float* special, const float* tab, void* psi,
                     int act_dtype, void* stream);
/* backward: dspecial accumulated, dtab written, special cells of dpsi zeroed in place */
int dx_embed_special_bwd(const float* xs, int B, int T, int V, int d, void* dpsi, int act_dtype, float* dspecial, float* dtab,
                         void* stream);
int dx_embed_bn_reduce(const float* xs, int B, int T, int V, const float* W0, const float* b0, const float* nobs,
                       const float* mean, const float* rstd, const void* dhn, int act_dtype, float* dgb, void* stream);
int dx_embed_bwd_front(const float* xs, int B, int T, int V, const float* W0, const float* b0, const float* nobs,
                       const float* gamma, const float* mean, const float* rstd, const void* dhn, int act_dtype,
                       const float* dgb, float* dW0, float* db0, float* dnobs, int training, void* stream);

/* BatchNormLastDim on [R,C] f32 (duett/duett.py:11-22; tab_encoder, cve, head) and LayerNorm on [R,C]
 * (models/main_architecture_duett.py:745-774).  Backward accumulates dw/db; dx may be NULL. */
int dx_bn2d_fwd(const float* x, int R, int C, const float* w, const float* b, float* run_mean, float* run_var, float* y,
                float* mean, float* rstd, double* stats_ws /* [C,2] */, int training, void* stream);
int dx_bn2d_bwd(const float* dy, const float* x, int R, int C, const float* w, const float* mean, const float* rstd,
                float* dx, float* dw, float* db, float* ws /* [C,2] */, int training, void* stream);
int dx_layernorm_fwd(const void* x, int R, int C, const float* w, const float* b, void* y, float* mean, float* rstd,
                     int dtype, void* stream);
int dx_layernorm_bwd(const void* dy, const void* x, int R, int C, const float* w, const float* mean, const float* rstd,
                     void* dx, float* dw, float* db, int dtype, void* stream);
/* out = act(x) for act in DX_ACT_{GELU,RELU,TANH} (activation after a split-K GEMM, whose epilogue cannot apply it). */
int dx_act_fwd(const void* x, void* out, int64_t n, int act, int dtype, void* stream);
/* out = g * act'(aux), act in DX_ACT_{GELU,RELU,TANH}_BWD. */
int dx_act_bwd(const void* g, const void* aux, void* out, int64_t n, int act, int dtype, void* stream);

/* Pooling / gathers around the heads: mean over the T hourly tokens (models/main_architecture_duett.py:1228-1231,
 * duett/duett.py:297-298); row / column gathers of the SSL heads (duett/duett.py:287-296,310-313) and their scatter.
 * A negative offset selects nothing: its gathered row is zero and nothing is scattered for it (the zero-padded rows of
 * pretrain_masked_steps > 1, duett/duett.py:291). */
int dx_mean_rows(const void* x, float* y, int B, int T1, int T, int64_t E, int dtype, void* stream);
int dx_mean_rows_bwd(const float* dy, void* dx, int B, int T1, int T, int64_t E, int dtype, void* stream);
int dx_gather_vec(const void* src, const int64_t* offsets, float* out, int n, int L, int dtype, void* stream);
int dx_scatter_vec(const float* src, const int64_t* offsets, void* dst, int n, int L, int accumulate, int dtype, void* stream);

/* Losses: each writes the scalar terms AND d(loss)/d(logits) (pass NULL to skip the gradient).
 *  dx_kd_loss          out3 = {alpha*bce+(1-alpha)*kd, bce, kd}       loss/losses_duett.py:8-25,39-57
 *  dx_bce_logits       mean_i w_i*bce(z_i,y_i), w = y>0 ? w_pos:w_neg  duett/duett.py:360-365
 *  dx_masked_mse_bce   out2[0] += mse(yhat*m,y*m); out2[1] += w*bce(phat,m)   duett/duett.py:337-358
 *  dx_masked_bce_cols  per[k] = sum_b bce*m/(sum_b m+eps); dz scaled by coef[k] (NULL=1)  loss/losses_duett.py:152-194
 *  dx_aux_residual_kl  label-smoothed Bernoulli KL on sigmoid(img.detach()+corr)   training_duett/engine.py:149-165 */
int dx_kd_loss(const float* zs, const float* zt, const float* y, int B, float T, float alpha, float pos_weight, float eps,
               float* out3, float* dz, void* stream);
int dx_bce_logits(const float* z, const float* y, int n, float w_pos, float w_neg, float* out, float* dz, void* stream);
int dx_masked_mse_bce(const float* yhat, const float* phat, const float* y, const float* m, int n, float w_presence,
                      float* out2, float* d_yhat, float* d_phat, void* stream);
int dx_masked_bce_cols(const float* z, const float* y, const float* m, const float* pos_weight, const float* coef, int B,
                       int K, float eps, float* per, float* dz, void* stream);
int dx_aux_residual_kl(const float* img_logits, const float* scaled_corr, const float* y, const float* mask, int n,
                       float eps_smooth, float* out, float* dcorr, void* stream);
/* out[i] = x[i]*s[i % ns] (device-side scale by the upstream loss gradient; no host sync). */
int dx_scale_dev(const float* x, const float* s, float* out, int64_t n, int ns, void* stream);
/* Pathology-query logits (models/main_architecture_duett.py:631-639): img = hi+bias_i, ts = ht+bias_t, scaled = beta*corr,
 * fusion = img.detach()+scaled; backward accumulates dbeta/dbias_*, writes d_corr. */
int dx_fusion_logits(const float* hi, const float* ht, const float* corr, const float* bias_i, const float* bias_t,
                     const float* beta, float* img, float* ts, float* scaled, float* fusion, int B, int K, void* stream);
int dx_fusion_logits_bwd(const float* d_img, const float* d_ts, const float* d_scaled, const float* d_fus, const float* corr,
                         const float* beta, float* d_corr, float* dbeta, float* dbias_i, float* dbias_t, int B, int K,
                         void* stream);

/* Fused AdamW on flat f32 buffers + global-norm clipping pieces (training_duett/trainer.py:383,902; duett/duett.py:325-327;
 * duett/train_duett_ssl.py:191 gradient_clip_val).  Effective grad = g * grad_scale * (grad_scale_dev ? *grad_scale_dev : 1). */
/* step_dev (int, device) overrides `step` and lr_scale_dev (float, device) multiplies lr when non-NULL, so a captured CUDA
 * graph of the whole training step stays valid as the step counter / LR schedule advance.  shadow_bf16 (optional, n bf16
 * elements): receives the updated parameters rounded to bf16 in the same pass (the tensor-core operands of the next step). */
int dx_adamw(float* p, const float* g, float* m, float* v, int64_t n, float lr, float beta1, float beta2, float eps,
             float weight_decay, int step, const float* grad_scale_dev, float grad_scale, const int* step_dev,
             const float* lr_scale_dev, void* shadow_bf16, void* stream);
int dx_sumsq(const float* x, int64_t n, float* out, void* stream);
int dx_clip_factor(const float* sumsq, float max_norm, float* clip, void* stream);

/* nn.Dropout (heads, perceiver feed-forward, FFN hidden of the axis encoders: ff_dropout, duett/duett.py:99,105):
 * y[i] = x[i] * keep(i) / (1-p) with the generator of dx_attn_fwd over the flat index i; applying the same call to the
 * upstream gradient is the backward pass.  x == y allowed. */
int dx_dropout(const void* x, void* y, int64_t n, float p, uint64_t seed, const uint64_t* seed_dev, int dtype, void* stream);
/* out[n] = sum_c a[n,c] * (b[n,c] - bias[c])  (bias may be NULL): the FFN ScaleNorm-backward row dot when a dropout mask
 * sits between the GELU and W2 (the fused GELU' epilogue of dx_gemm computes it only for the mask-free case). */
int dx_rowdot_bias(const void* a, const void* b, const float* bias, float* out, int N, int C, int dtype, void* stream);

/* ---- evaluation metrics --------------------------------------------------------------------------------
 * Replaces the host-side scoring of training_duett/evaluator.py:22-35 (torch.sigmoid -> .numpy() -> sklearn
 * roc_auc_score / average_precision_score).  out[0] = AUROC (trapezoid over the ROC points of the distinct-score
 * thresholds), out[1] = AUPRC (step-wise average precision), out[2] = number of positives, out[3] = n; NaN where sklearn
 * raises (single-class input).  key_ws / lab_ws: npad floats of scratch each, npad = n rounded up to a power of two.
 * apply_sigmoid = 1 ranks sigmoid(logits) computed in fp32 like the reference, 0 ranks the raw scores. */
int dx_binary_auc(const float* logits, const float* labels, int64_t n, float* key_ws, float* lab_ws, int64_t npad,
                  int apply_sigmoid, double* out, void* stream);

/* ---- input binning ----------------------------------------------------------------------------------------
 * Replaces the python row walk of duett/mimic_dataset.py:33-46 (build_stay_tensor) for a whole batch of stays in one
 * launch.  Rows of all stays are concatenated: slot[r] = hourly slot index of row r (rows with slot >= T or < 0 are
 * ignored by construction: no thread owns them), vals / cnts [R, V] float64 (value and observation count of every
 * variable in that row), row_start [B+1] = first row of each stay.  x [B, T, 2V] f32 is fully written (zeros where nothing
 * was observed); a later row of the same slot overwrites an earlier one, like the reference's sequential assignment.
 * float64 arithmetic, one rounding to float32: bit-identical to the reference. */
int dx_bin_events(const int* slot, const double* vals, const double* cnts, const int64_t* row_start, const double* means,
                  const double* stds, int B, int T, int V, float* x, void* stream);

/* ---- SSL masking ------------------------------------------------------------------------------------------
 * Model.pretrain_prep_batch (duett/duett.py:189-237) as one launch.  The random draws stay on the host (numpy Generator,
 * same order as the reference: per sample K = pretrain_masked_steps timesteps — one rng.choice call, with replacement when
 * K > 1 — then one variable, then the [B,V] variable-dropout matrix) and arrive as index arrays: step [B,K] int32, ev [B]
 * int32 (NULL = predict_events off), keep [B,V] uint8 (NULL = pretrain_dropout 0).  xs [B,T,2V+1] f32 -> xc (masked copy),
 * y_ts / y_mask [B,K,V] (values and clipped counts of the masked timesteps, draw order, repeats included), y_ev / y_ev_mask
 * [B,T] (the masked variable's column).  Bit-exact selection. */
int dx_ssl_mask(const float* xs, const int* step, const int* ev, const unsigned char* keep, int B, int T, int V, int K,
                float* xc, float* y_ts, float* y_mask, float* y_ev, float* y_ev_mask, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DUETT_B200_H */
