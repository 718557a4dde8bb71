// Value/count embedding of the binned [B,T,V] grid into psi[B,T+1,V+1,d]  (duett/duett.py:245-266 ==
// models/main_architecture_duett.py:31-65).  The reference loops over V per-variable MLPs
//     Linear(2,64) -> ReLU -> BatchNorm1d(64, batch statistics over B*T rows) -> Linear(64,d)
// and scatters each result into psi (V CopySlices autograd nodes, ~30 % of its step).  Here the 64 -> d contraction of
// ALL variables is ONE grouped tensor-core GEMM (dx_gemm, batch = V, written straight into the strided psi view) and the
// cheap 2 -> 64 front is recomputed from the two input scalars wherever it is needed:
//   dx_embed_stats       : per-(variable, channel) sum / sum-of-squares of the ReLU hidden -> mean / rstd, running-stat
//                          update (momentum 0.1, unbiased running var); eval mode reads the running buffers
//   dx_embed_hidden      : count-embedding lookup (n_obs_embedding, clip 0..15) -> ReLU hidden -> BatchNorm ->
//                          hn[V, B*(T+1), 64] in the act dtype ([REP] rows zero) = A operand of the grouped GEMM
//   dx_embed_special     : static column, [REP] row and the MASK substitution for masked timesteps / masked variables
//   dx_embed_special_bwd : gradients of MASK / [REP] / tab_encoder output; zeroes those cells of dpsi in place so the
//                          grouped dW4 / dhn GEMMs see exact zeros there
//   dx_embed_bn_reduce   : dgamma / dbeta of this step from dhn (needed by every row of the BN backward)
//   dx_embed_bwd_front   : BN backward + ReLU + first Linear + count-embedding gradients
// Hidden width is fixed at 64 (the reference default d_hidden_mlp_embedding; every BASELINE config uses it).
#include "dx_common.cuh"
#include "../../include/duett_b200.h"

namespace {

constexpr int H = 64;
constexpr int NT = 256;
constexpr float BN_EPS = 1e-5f;

struct EmbedIn {
  const float* xs;  // [B,T,2V+1]
  int B, T, V, d;
  const float* W0;   // [V,H,2]
  const float* b0;   // [V,H]
  const float* nobs; // [16]
};

__device__ __forceinline__ void cell_inputs(const EmbedIn& in, int r, int v, float& val, float& cnte, int& idx) {
  const float* row = in.xs + (long long)r * (2 * in.V + 1);
  val = row[v];
  const float c = row[in.V + v];
  idx = min(max((int)c, 0), 15);
  cnte = in.nobs[idx];
}
__device__ __forceinline__ bool cell_masked(const EmbedIn& in, int b, int t, int v) {
  // (t < T and the timestep is masked) or (v < V and the variable is event-masked; the [REP] row copies row 0's mask)
  const int V = in.V, T = in.T;
  if (t < T && in.xs[((long long)b * T + t) * (2 * V + 1) + 2 * V] == 1.f) return true;
  if (v < V && in.xs[((long long)b * T + (t < T ? t : 0)) * (2 * V + 1) + V + v] == -1.f) return true;
  return false;
}

// ---- stats -------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NT) embed_stats_kernel(EmbedIn in, double* __restrict__ stats /*[V,H,2]*/, int rows_per_block) {
  __shared__ float sh[2][4][H];
  const int v = blockIdx.x;
  const int c = threadIdx.x & (H - 1), rl = threadIdx.x >> 6;
  const int R = in.B * in.T;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(R, r0 + rows_per_block);
  const float w0 = in.W0[(v * H + c) * 2], w1 = in.W0[(v * H + c) * 2 + 1], bb = in.b0[v * H + c];
  float s = 0.f, ss = 0.f;
  for (int r = r0 + rl; r < r1; r += 4) {
    float val, cnte; int idx;
    cell_inputs(in, r, v, val, cnte, idx);
    const float h = fmaxf(fmaf(w0, val, fmaf(w1, cnte, bb)), 0.f);
    s += h;
    ss = fmaf(h, h, ss);
  }
  sh[0][rl][c] = s;
  sh[1][rl][c] = ss;
  __syncthreads();
  if (rl == 0) {
    s = sh[0][0][c] + sh[0][1][c] + sh[0][2][c] + sh[0][3][c];
    ss = sh[1][0][c] + sh[1][1][c] + sh[1][2][c] + sh[1][3][c];
    atomicAdd(stats + ((long long)v * H + c) * 2, (double)s);
    atomicAdd(stats + ((long long)v * H + c) * 2 + 1, (double)ss);
  }
}

__global__ void embed_bnfinal_kernel(const double* __restrict__ stats, int n, int R, float* __restrict__ mean,
                                     float* __restrict__ rstd, float* __restrict__ run_mean, float* __restrict__ run_var,
                                     float momentum) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double m = stats[2 * i] / R;
  double var = stats[2 * i + 1] / R - m * m;
  if (var < 0) var = 0;
  mean[i] = (float)m;
  rstd[i] = (float)(1.0 / sqrt(var + (double)BN_EPS));
  if (run_mean) {
    const double unb = R > 1 ? var * R / (R - 1) : var;
    run_mean[i] = (1.f - momentum) * run_mean[i] + momentum * (float)m;
    run_var[i] = (1.f - momentum) * run_var[i] + momentum * (float)unb;
  }
}

__global__ void embed_bn_eval_kernel(const float* __restrict__ run_mean, const float* __restrict__ run_var, int n,
                                     float* __restrict__ mean, float* __restrict__ rstd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  mean[i] = run_mean[i];
  rstd[i] = rsqrtf(run_var[i] + BN_EPS);
}

// ---- hidden: hn[v, b*(T+1)+t, c] ----------------------------------------------------------------------
// grid (V, row chunks over B*(T+1)); thread = (group of 8 channels, row lane): one 16 B (bf16) store per row and thread,
// a warp covers 4 whole rows = 512 contiguous bytes
constexpr int CG = 8;            // channels per thread
constexpr int RL = NT / (H / CG);   // 32 row lanes per block

struct ChanConst {   // per-thread constants of its 8 channels
  float w0[CG], w1[CG], bb[CG];
};
__device__ __forceinline__ ChanConst load_chan(const EmbedIn& in, int v, int c0) {
  ChanConst k;
#pragma unroll
  for (int j = 0; j < CG; ++j) {
    const int i = v * H + c0 + j;
    k.w0[j] = in.W0[i * 2]; k.w1[j] = in.W0[i * 2 + 1]; k.bb[j] = in.b0[i];
  }
  return k;
}

template <typename T>
__global__ void __launch_bounds__(NT) embed_hidden_kernel(EmbedIn in, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                         const float* __restrict__ mean, const float* __restrict__ rstd,
                                                         T* __restrict__ hn, int rows_per_block) {
  const int v = blockIdx.x;
  const int T1 = in.T + 1;
  const int R1 = in.B * T1;
  const int c0 = (threadIdx.x & (H / CG - 1)) * CG, rl = threadIdx.x / (H / CG);
  const int r0 = blockIdx.y * rows_per_block, r1 = min(R1, r0 + rows_per_block);
  const ChanConst k = load_chan(in, v, c0);
  float sc[CG], sf[CG];
#pragma unroll
  for (int j = 0; j < CG; ++j) {
    const int i = v * H + c0 + j;
    sc[j] = gamma[i] * rstd[i];
    sf[j] = beta[i] - mean[i] * sc[j];
  }
  for (int rr = r0 + rl; rr < r1; rr += RL) {
    const int b = rr / T1, t = rr - b * T1;
    float o[CG];
#pragma unroll
    for (int j = 0; j < CG; ++j) o[j] = 0.f;
    if (t < in.T) {
      float val, cnte; int idx;
      cell_inputs(in, b * in.T + t, v, val, cnte, idx);
#pragma unroll
      for (int j = 0; j < CG; ++j) o[j] = fmaf(fmaxf(fmaf(k.w0[j], val, fmaf(k.w1[j], cnte, k.bb[j])), 0.f), sc[j], sf[j]);
    }
    dx_st8(hn + ((long long)v * R1 + rr) * H + c0, o);
  }
}

// ---- special cells ------------------------------------------------------------------------------------
// one thread per (cell, 8-element vector); writes only static column / [REP] row / masked cells
template <typename T>
__global__ void __launch_bounds__(NT) embed_special_kernel(EmbedIn in, const float* __restrict__ tab /*[B,d]*/,
                                                          const float* __restrict__ sp, T* __restrict__ psi) {
  const int d = in.d, V1 = in.V + 1, T1 = in.T + 1, dv = d >> 3;
  const long long n = (long long)in.B * T1 * V1 * dv;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const int k = (int)(i % dv);
    const long long cell = i / dv;
    const int v = (int)(cell % V1);
    const int t = (int)((cell / V1) % T1);
    const int b = (int)(cell / ((long long)V1 * T1));
    const float* srcp;
    if (cell_masked(in, b, t, v)) srcp = sp;
    else if (t == in.T) srcp = sp + d;
    else if (v == in.V) srcp = tab + (long long)b * d;
    else continue;
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = srcp[k * 8 + j];
    dx_st8(psi + cell * d + k * 8, o);
  }
}

// grads of special_embeddings[0] (MASK), [1] ([REP]) and of the tab_encoder output; one block per (sample, time row).
// Every special cell of dpsi is zeroed afterwards (the embedding MLP outputs there were overwritten in the forward).
// The row's V+1 cells are classified once into shared memory (1 = MASK, 2 = [REP], 3 = static column, 0 = ordinary);
// the d-vector loop then touches the special cells only.
template <typename T>
__global__ void __launch_bounds__(NT) embed_special_bwd_kernel(EmbedIn in, T* __restrict__ dpsi, float* __restrict__ dsp /*[8,d]*/,
                                                              float* __restrict__ dtab /*[B,d], zeroed by the caller*/) {
  extern __shared__ unsigned char kind[];   // [V+1]
  const int d = in.d, V = in.V, T_ = in.T;
  const int b = blockIdx.x / (T_ + 1), t = blockIdx.x % (T_ + 1);
  for (int v = threadIdx.x; v <= V; v += blockDim.x)
    kind[v] = cell_masked(in, b, t, v) ? 1 : (t == T_ ? 2 : (v == V ? 3 : 0));
  __syncthreads();
  T* row = dpsi + (((long long)b * (T_ + 1) + t) * (V + 1)) * d;
  for (int dd = threadIdx.x; dd < d; dd += blockDim.x) {  // threads over dd: coalesced across the d-vector
    float a0 = 0.f, a1 = 0.f, at = 0.f;
    for (int v = 0; v <= V; ++v) {
      const int kd = kind[v];
      if (kd) {
        T* p = row + (long long)v * d + dd;
        const float g = dx_ld(p);
        if (kd == 1) a0 += g;
        else if (kd == 2) a1 += g;
        else at += g;
        dx_st(p, 0.f);
      }
    }
    if (a0 != 0.f) atomicAdd(dsp + dd, a0);
    if (a1 != 0.f) atomicAdd(dsp + d + dd, a1);
    if (at != 0.f) atomicAdd(dtab + (long long)b * d + dd, at);
  }
}

// dgamma[v,c] = sum_r dhn*hhat ; dbeta[v,c] = sum_r dhn   (this step's values, [2,V,H] f32, zeroed by the caller)
// thread = (group of 8 channels, row lane): 16 B loads of dhn, a warp covers 4 whole rows
template <typename T>
__global__ void __launch_bounds__(NT) embed_bn_reduce_kernel(EmbedIn in, const float* __restrict__ mean, const float* __restrict__ rstd,
                                                            const T* __restrict__ dhn, float* __restrict__ dgb, int rows_per_block) {
  __shared__ float sh[2][RL][H + 1];
  const int v = blockIdx.x;
  const int c0 = (threadIdx.x & (H / CG - 1)) * CG, rl = threadIdx.x / (H / CG);
  const int T1 = in.T + 1, R1 = in.B * T1;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(R1, r0 + rows_per_block);
  const ChanConst k = load_chan(in, v, c0);
  float mu[CG], rs[CG], ag[CG], ab[CG];
#pragma unroll
  for (int j = 0; j < CG; ++j) {
    mu[j] = mean[v * H + c0 + j]; rs[j] = rstd[v * H + c0 + j];
    ag[j] = 0.f; ab[j] = 0.f;
  }
  for (int rr = r0 + rl; rr < r1; rr += RL) {
    const int b = rr / T1, t = rr - b * T1;
    if (t == in.T) continue;
    float val, cnte; int idx;
    cell_inputs(in, b * in.T + t, v, val, cnte, idx);
    float g[CG];
    dx_ld8(dhn + ((long long)v * R1 + rr) * H + c0, g);
#pragma unroll
    for (int j = 0; j < CG; ++j) {
      const float h = fmaxf(fmaf(k.w0[j], val, fmaf(k.w1[j], cnte, k.bb[j])), 0.f);
      ab[j] += g[j];
      ag[j] = fmaf(g[j], (h - mu[j]) * rs[j], ag[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < CG; ++j) {
    sh[0][rl][c0 + j] = ag[j];
    sh[1][rl][c0 + j] = ab[j];
  }
  __syncthreads();
  if (threadIdx.x < 2 * H) {
    const int which = threadIdx.x / H, c = threadIdx.x % H;
    float a = 0.f;
#pragma unroll 8
    for (int r = 0; r < RL; ++r) a += sh[which][r][c];
    atomicAdd(dgb + which * in.V * H + v * H + c, a);
  }
}

// BatchNorm backward + ReLU + first Linear + count-embedding grads.
//   dh = gamma*rstd*(dhn - (dbeta + hhat*dgamma)/R);  dpre = dh*(pre>0)
template <typename T>
__global__ void __launch_bounds__(NT) embed_bwd_front_kernel(EmbedIn in, const float* __restrict__ gamma, const float* __restrict__ mean,
                                                            const float* __restrict__ rstd, const T* __restrict__ dhn,
                                                            const float* __restrict__ dgb, float* __restrict__ dW0,
                                                            float* __restrict__ db0, float* __restrict__ dnobs, int rows_per_block,
                                                            int training) {
  __shared__ float sh[3][RL][H + 1];
  __shared__ float snobs[16];
  const int v = blockIdx.x;
  const int T1 = in.T + 1, R1 = in.B * T1, R = in.B * in.T;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(R1, r0 + rows_per_block);
  const int c0 = (threadIdx.x & (H / CG - 1)) * CG, rl = threadIdx.x / (H / CG);
  if (threadIdx.x < 16) snobs[threadIdx.x] = 0.f;
  __syncthreads();
  const ChanConst k = load_chan(in, v, c0);
  float mu[CG], rs[CG], gr[CG], mdg[CG], mdb[CG], a0[CG], a1[CG], ab[CG];
#pragma unroll
  for (int j = 0; j < CG; ++j) {
    const int i = v * H + c0 + j;
    mu[j] = mean[i]; rs[j] = rstd[i]; gr[j] = gamma[i] * rstd[i];
    mdg[j] = training ? dgb[i] / R : 0.f;
    mdb[j] = training ? dgb[in.V * H + i] / R : 0.f;
    a0[j] = a1[j] = ab[j] = 0.f;
  }
  // uniform trip count for the whole block: the 8 lanes of a row shuffle inside the loop
  for (int base = r0; base < r1; base += RL) {
    const int rr = base + rl;
    const int b = rr / T1, t = rr - b * T1;
    const bool valid = rr < r1 && t < in.T;
    float dc = 0.f;
    int idx = 0;
    if (valid) {
      float val, cnte;
      cell_inputs(in, b * in.T + t, v, val, cnte, idx);
      float g[CG];
      dx_ld8(dhn + ((long long)v * R1 + rr) * H + c0, g);
#pragma unroll
      for (int j = 0; j < CG; ++j) {
        const float pre = fmaf(k.w0[j], val, fmaf(k.w1[j], cnte, k.bb[j]));
        const float hhat = (fmaxf(pre, 0.f) - mu[j]) * rs[j];
        const float dh = gr[j] * (g[j] - mdb[j] - hhat * mdg[j]);
        const float dpre = pre > 0.f ? dh : 0.f;
        a0[j] = fmaf(dpre, val, a0[j]);
        a1[j] = fmaf(dpre, cnte, a1[j]);
        ab[j] += dpre;
        dc = fmaf(dpre, k.w1[j], dc);
      }
    }
    // d(count embedding): reduce over the row's 64 channels = 8 consecutive lanes
    dc += __shfl_xor_sync(0xffffffffu, dc, 1);
    dc += __shfl_xor_sync(0xffffffffu, dc, 2);
    dc += __shfl_xor_sync(0xffffffffu, dc, 4);
    if (valid && c0 == 0) atomicAdd(&snobs[idx], dc);
  }
#pragma unroll
  for (int j = 0; j < CG; ++j) {
    sh[0][rl][c0 + j] = a0[j];
    sh[1][rl][c0 + j] = a1[j];
    sh[2][rl][c0 + j] = ab[j];
  }
  __syncthreads();
  if (threadIdx.x < 3 * H) {
    const int which = threadIdx.x / H, c = threadIdx.x % H;
    float a = 0.f;
#pragma unroll 8
    for (int r = 0; r < RL; ++r) a += sh[which][r][c];
    if (which == 0) atomicAdd(dW0 + (v * H + c) * 2, a);
    else if (which == 1) atomicAdd(dW0 + (v * H + c) * 2 + 1, a);
    else atomicAdd(db0 + v * H + c, a);
  }
  if (threadIdx.x < 16 && snobs[threadIdx.x] != 0.f) atomicAdd(dnobs + threadIdx.x, snobs[threadIdx.x]);
}

int chunking(int V, int R, int& rows_per_block) {
  // every thread walks its rows with dependent 2-4 B loads (latency bound): many short blocks (~32 per SM) hide it
  int nch = max(1, (32 * 148 + V - 1) / V);
  nch = min(nch, max(1, R / 32));
  rows_per_block = (R + nch - 1) / nch;
  rows_per_block = ((rows_per_block + 7) / 8) * 8;
  return (R + rows_per_block - 1) / rows_per_block;
}

}  // namespace

extern "C" {

/* mean / rstd [V,64] of the ReLU hidden of every variable.  training=1: batch statistics over the B*T rows (stats_ws:
 * [V,64,2] doubles scratch; running buffers updated when non-NULL); training=0: from the running buffers. */
int dx_embed_stats(const float* xs, int B, int T, int V, const float* W0, const float* b0, const float* nobs, float* run_mean,
                   float* run_var, double* stats_ws, float* mean, float* rstd, int training, void* stream) {
  DX_CHECK_ARG(xs && W0 && b0 && nobs && mean && rstd, "dx_embed_stats: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  EmbedIn in{xs, B, T, V, 0, W0, b0, nobs};
  if (training) {
    DX_CHECK_ARG(stats_ws, "dx_embed_stats: stats workspace required in training mode");
    DX_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * V * H * 2, st));
    int rpb;
    const int nch = chunking(V, B * T, rpb);
    embed_stats_kernel<<<dim3(V, nch), NT, 0, st>>>(in, stats_ws, rpb);
    DX_LAUNCH_CHECK();
    embed_bnfinal_kernel<<<dx_ceil_div(V * H, 256), 256, 0, st>>>(stats_ws, V * H, B * T, mean, rstd, run_mean, run_var, 0.1f);
  } else {
    DX_CHECK_ARG(run_mean && run_var, "dx_embed_stats: eval mode needs running statistics");
    embed_bn_eval_kernel<<<dx_ceil_div(V * H, 256), 256, 0, st>>>(run_mean, run_var, V * H, mean, rstd);
  }
  DX_LAUNCH_CHECK();
  return DX_OK;
}

/* hn[V, B*(T+1), 64] (act dtype): BatchNorm'd ReLU hidden of every (variable, cell); [REP] rows are zero. */
int dx_embed_hidden(const float* xs, int B, int T, int V, const float* W0, const float* b0, const float* nobs,
                    const float* gamma, const float* beta, const float* mean, const float* rstd, void* hn, int act_dtype,
                    void* stream) {
  DX_CHECK_ARG(xs && W0 && b0 && nobs && gamma && beta && mean && rstd && hn, "dx_embed_hidden: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  EmbedIn in{xs, B, T, V, 0, W0, b0, nobs};
  int rpb;
  const int nch = chunking(V, B * (T + 1), rpb);
  if (act_dtype == DX_BF16) embed_hidden_kernel<bf16><<<dim3(V, nch), NT, 0, st>>>(in, gamma, beta, mean, rstd, (bf16*)hn, rpb);
  else embed_hidden_kernel<float><<<dim3(V, nch), NT, 0, st>>>(in, gamma, beta, mean, rstd, (float*)hn, rpb);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

/* psi[B,T+1,V+1,d]: overwrite the static column (tab), the [REP] row (special[1]) and every masked cell (special[0]). */
int dx_embed_special(const float* xs, int B, int T, int V, int d, const float* special, const float* tab, void* psi,
                     int act_dtype, void* stream) {
  DX_CHECK_ARG(xs && special && tab && psi && d % 8 == 0, "dx_embed_special: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  EmbedIn in{xs, B, T, V, d, nullptr, nullptr, nullptr};
  if (act_dtype == DX_BF16) embed_special_kernel<bf16><<<148 * 8, NT, 0, st>>>(in, tab, special, (bf16*)psi);
  else embed_special_kernel<float><<<148 * 8, NT, 0, st>>>(in, tab, special, (float*)psi);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

/* dspecial [8,d] accumulated, dtab [B,d] written; the special cells of dpsi are zeroed IN PLACE. */
int dx_embed_special_bwd(const float* xs, int B, int T, int V, int d, void* dpsi, int act_dtype, float* dspecial, float* dtab,
                         void* stream) {
  DX_CHECK_ARG(xs && dpsi && dspecial && dtab, "dx_embed_special_bwd: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  EmbedIn in{xs, B, T, V, d, nullptr, nullptr, nullptr};
  DX_CUDA(cudaMemsetAsync(dtab, 0, sizeof(float) * (size_t)B * d, st));
  const int nthr = d >= 256 ? 256 : (d >= 128 ? 128 : 64);
  const size_t ksm = (size_t)((V + 1 + 15) & ~15);
  if (act_dtype == DX_BF16) embed_special_bwd_kernel<bf16><<<B * (T + 1), nthr, ksm, st>>>(in, (bf16*)dpsi, dspecial, dtab);
  else embed_special_bwd_kernel<float><<<B * (T + 1), nthr, ksm, st>>>(in, (float*)dpsi, dspecial, dtab);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

/* dgb[0] = dgamma, dgb[1] = dbeta of this step ([2,V,64] f32, zeroed here) from dhn[V, B*(T+1), 64]. */
int dx_embed_bn_reduce(const float* xs, int B, int T, int V, const float* W0, const float* b0, const float* nobs,
                       const float* mean, const float* rstd, const void* dhn, int act_dtype, float* dgb, void* stream) {
  DX_CHECK_ARG(xs && W0 && b0 && nobs && mean && rstd && dhn && dgb, "dx_embed_bn_reduce: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  EmbedIn in{xs, B, T, V, 0, W0, b0, nobs};
  DX_CUDA(cudaMemsetAsync(dgb, 0, sizeof(float) * 2 * V * H, st));
  int rpb;
  const int nch = chunking(V, B * (T + 1), rpb);
  if (act_dtype == DX_BF16) embed_bn_reduce_kernel<bf16><<<dim3(V, nch), NT, 0, st>>>(in, mean, rstd, (const bf16*)dhn, dgb, rpb);
  else embed_bn_reduce_kernel<float><<<dim3(V, nch), NT, 0, st>>>(in, mean, rstd, (const float*)dhn, dgb, rpb);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

/* dW0 [V,64,2], db0 [V,64], dnobs [16] accumulated. training=0: frozen statistics (no mean terms). */
int dx_embed_bwd_front(const float* xs, int B, int T, int V, const float* W0, const float* b0, const float* nobs,
                       const float* gamma, const float* mean, const float* rstd, const void* dhn, int act_dtype,
                       const float* dgb, float* dW0, float* db0, float* dnobs, int training, void* stream) {
  DX_CHECK_ARG(xs && W0 && b0 && nobs && gamma && mean && rstd && dhn && dgb && dW0 && db0 && dnobs,
               "dx_embed_bwd_front: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  EmbedIn in{xs, B, T, V, 0, W0, b0, nobs};
  int rpb;
  const int nch = chunking(V, B * (T + 1), rpb);
  if (act_dtype == DX_BF16)
    embed_bwd_front_kernel<bf16><<<dim3(V, nch), NT, 0, st>>>(in, gamma, mean, rstd, (const bf16*)dhn, dgb, dW0, db0, dnobs, rpb, training);
  else
    embed_bwd_front_kernel<float><<<dim3(V, nch), NT, 0, st>>>(in, gamma, mean, rstd, (const float*)dhn, dgb, dW0, db0, dnobs, rpb, training);
  DX_LAUNCH_CHECK();
  return DX_OK;
}

}  // extern "C"
