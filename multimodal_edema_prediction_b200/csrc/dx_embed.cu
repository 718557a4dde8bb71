// Value/count embedding of the binned [B,T,V] grid into psi[B,T+1,V+1,d]  (duett/duett.py:245-266 ==
// models/main_architecture_duett.py:31-65).  The reference loops over V per-variable MLPs
//     Linear(2,64) -> ReLU -> BatchNorm1d(64, batch statistics over B*T rows) -> Linear(64,d)
// and scatters each result into psi (V CopySlices autograd nodes, ~30 % of its step).  Here all variables run as one
// grouped kernel family:
//   embed_stats   : per-(variable, channel) sum / sum-of-squares of the ReLU hidden (recomputed from the 2 inputs)
//   embed_bnfinal : mean / rstd, running-stat update (momentum 0.1, unbiased running var)
//   embed_apply   : BN folded into the second Linear, written straight into psi together with the count-embedding
//                   lookup (n_obs_embedding, clip 0..15), the MASK substitution for masked timesteps / masked
//                   variables, the static column and the [REP] row
//   backward      : embed_special_bwd (MASK/[REP]/static grads), embed_bwd_a (dW4, db4, dgamma, dbeta, dhn),
//                   embed_bwd_b (BN backward, ReLU, dW0, db0, count-embedding grads)
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

__device__ __forceinline__ void cell_inputs(const EmbedIn& in, int r, int v, float& val, float& cnte, int& idx, bool& ev_masked) {
  const float* row = in.xs + (long long)r * (2 * in.V + 1);
  val = row[v];
  const float c = row[in.V + v];
  ev_masked = (c == -1.f);
  idx = min(max((int)c, 0), 15);
  cnte = in.nobs[idx];
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
    float val, cnte; int idx; bool em;
    cell_inputs(in, r, v, val, cnte, idx, em);
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

// eval mode: statistics come from the running buffers
__global__ void embed_bn_eval_kernel(const float* __restrict__ run_mean, const float* __restrict__ run_var, int n,
                                     float* __restrict__ mean, float* __restrict__ rstd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  mean[i] = run_mean[i];
  rstd[i] = rsqrtf(run_var[i] + BN_EPS);
}

// ---- apply -------------------------------------------------------------------------------------
// grid (V, row tiles of RT rows); block NT threads.  smem: Wf[d][H+1], bf[d], hs[RT][H], flags[RT]
constexpr int RT = 64;

template <typename T>
__global__ void __launch_bounds__(NT) embed_apply_kernel(EmbedIn in, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                        const float* __restrict__ W4 /*[V,d,H]*/, const float* __restrict__ b4 /*[V,d]*/,
                                                        const float* __restrict__ sp /*[8,d]*/, T* __restrict__ psi) {
  extern __shared__ float smem[];
  const int d = in.d, V = in.V, T_ = in.T;
  float* Wf = smem;                    // d*(H+1)
  float* bf = Wf + d * (H + 1);        // d
  float* hs = bf + d;                  // RT*H
  int* flags = reinterpret_cast<int*>(hs + RT * H);  // RT
  __shared__ float sc[H], sf[H];       // per-channel scale / shift of the folded BN
  const int v = blockIdx.x;
  const int R = in.B * T_;
  const int r0 = blockIdx.y * RT;
  if (threadIdx.x < H) {
    const int c = threadIdx.x;
    const float a = gamma[v * H + c] * rstd[v * H + c];
    sc[c] = a;
    sf[c] = beta[v * H + c] - mean[v * H + c] * a;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < d * H; i += NT) {
    const int dd = i / H, c = i - dd * H;
    Wf[dd * (H + 1) + c] = W4[((long long)v * d + dd) * H + c] * sc[c];
  }
  for (int dd = threadIdx.x; dd < d; dd += NT) {
    float s = b4[v * d + dd];
    for (int c = 0; c < H; ++c) s = fmaf(W4[((long long)v * d + dd) * H + c], sf[c], s);
    bf[dd] = s;
  }
  {
    const int c = threadIdx.x & (H - 1), rl = threadIdx.x >> 6;
    const float w0 = in.W0[(v * H + c) * 2], w1 = in.W0[(v * H + c) * 2 + 1], bb = in.b0[v * H + c];
    for (int rr = rl; rr < RT; rr += 4) {
      const int r = r0 + rr;
      float h = 0.f;
      if (r < R) {
        float val, cnte; int idx; bool em;
        cell_inputs(in, r, v, val, cnte, idx, em);
        h = fmaxf(fmaf(w0, val, fmaf(w1, cnte, bb)), 0.f);
        if (c == 0) {
          const bool step_masked = in.xs[(long long)r * (2 * V + 1) + 2 * V] == 1.f;
          flags[rr] = (em || step_masked) ? 1 : 0;
        }
      }
      hs[rr * H + c] = h;
    }
  }
  __syncthreads();
  const int lanes = NT / d > 0 ? NT / d : 1;  // row lanes when d < NT
  for (int dd0 = 0; dd0 < d; dd0 += NT) {
    const int dd = dd0 + (threadIdx.x % min(d, NT));
    const int rl = threadIdx.x / min(d, NT);
    if (dd >= d || rl >= lanes) continue;
    float w[H];
#pragma unroll
    for (int c = 0; c < H; ++c) w[c] = Wf[dd * (H + 1) + c];
    const float bias = bf[dd];
    const float m0 = sp[dd];
    for (int rr = rl; rr < RT; rr += lanes) {
      const int r = r0 + rr;
      if (r >= R) break;
      float s = bias;
      const float4* h4 = reinterpret_cast<const float4*>(hs + rr * H);
#pragma unroll
      for (int c4 = 0; c4 < H / 4; ++c4) {
        const float4 hv = h4[c4];
        s = fmaf(hv.x, w[4 * c4], s);
        s = fmaf(hv.y, w[4 * c4 + 1], s);
        s = fmaf(hv.z, w[4 * c4 + 2], s);
        s = fmaf(hv.w, w[4 * c4 + 3], s);
      }
      if (flags[rr]) s = m0;
      const int b = r / T_, t = r - b * T_;
      dx_st(psi + ((((long long)b * (T_ + 1) + t) * (V + 1)) + v) * d + dd, s);
    }
  }
}

// static column (v == V, t < T) and [REP] row (t == T): elementwise over [B, T+1+V, d] "extra" cells
template <typename T>
__global__ void __launch_bounds__(NT) embed_special_kernel(EmbedIn in, const float* __restrict__ tab /*[B,d]*/,
                                                          const float* __restrict__ sp, T* __restrict__ psi) {
  const int d = in.d, V = in.V, T_ = in.T;
  const long long n = (long long)in.B * (T_ + V + 1) * d;
  for (long long i = (long long)blockIdx.x * NT + threadIdx.x; i < n; i += (long long)gridDim.x * NT) {
    const int dd = (int)(i % d);
    const long long cell = i / d;
    const int b = (int)(cell / (T_ + V + 1));
    const int j = (int)(cell % (T_ + V + 1));
    int t, v;
    float val;
    if (j < T_) {  // static column
      t = j; v = V;
      const bool step_masked = in.xs[((long long)b * T_ + t) * (2 * V + 1) + 2 * V] == 1.f;
      val = step_masked ? sp[dd] : tab[(long long)b * d + dd];
    } else {       // [REP] row, copies row 0's event mask (duett/duett.py:250)
      t = T_; v = j - T_;
      bool em = false;
      if (v < V) em = in.xs[((long long)b * T_) * (2 * V + 1) + V + v] == -1.f;
      val = em ? sp[dd] : sp[d + dd];
    }
    dx_st(psi + ((((long long)b * (T_ + 1) + t) * (V + 1)) + v) * d + dd, val);
  }
}

// ---- backward ----------------------------------------------------------------------------------
// grads of special_embeddings[0] (MASK), [1] ([REP]) and of the tab_encoder output; one block per sample.
template <typename T>
__global__ void __launch_bounds__(NT) embed_special_bwd_kernel(EmbedIn in, const T* __restrict__ dpsi, float* __restrict__ dsp /*[8,d]*/,
                                                              float* __restrict__ dtab /*[B,d]*/) {
  const int d = in.d, V = in.V, T_ = in.T;
  const int b = blockIdx.x;
  for (int dd = threadIdx.x; dd < d; dd += NT) {  // threads over dd: coalesced across the d-vector
    float a0 = 0.f, a1 = 0.f, at = 0.f;
    for (int t = 0; t <= T_; ++t) {
      const bool step_masked = t < T_ && in.xs[((long long)b * T_ + t) * (2 * V + 1) + 2 * V] == 1.f;
      for (int v = 0; v <= V; ++v) {
        bool em = false;
        if (v < V) em = in.xs[((long long)b * T_ + (t < T_ ? t : 0)) * (2 * V + 1) + V + v] == -1.f;
        const float g = dx_ld(dpsi + ((((long long)b * (T_ + 1) + t) * (V + 1)) + v) * d + dd);
        if (step_masked || em) a0 += g;
        else if (t == T_) a1 += g;
        else if (v == V) at += g;
      }
    }
    atomicAdd(dsp + dd, a0);
    atomicAdd(dsp + d + dd, a1);
    dtab[(long long)b * d + dd] = at;
  }
}

// Phase A: grid (V, NCH). Per block: loop over its row chunk in sub-tiles of RT rows.
//   dhn[r][c] = sum_dd W4[v,dd,c] * dout[r][dd]      -> stored to dhn_ws[v][r][c]
//   dbeta[c] += dhn ; dgamma[c] += dhn * hhat         (atomics at block end)
//   dW4[dd][c] += dout[r][dd] * hn[r][c] ; db4[dd] += dout[r][dd]
template <typename T>
__global__ void __launch_bounds__(NT) embed_bwd_a_kernel(EmbedIn in, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        const float* __restrict__ mean, const float* __restrict__ rstd,
                                                        const float* __restrict__ W4, const T* __restrict__ dpsi,
                                                        float* __restrict__ dhn_ws, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                                        float* __restrict__ dW4, float* __restrict__ db4, int rows_per_block) {
  extern __shared__ float smem[];
  const int d = in.d, V = in.V, T_ = in.T;
  float* Wt = smem;                 // [H][d+1]  (W4 transposed: Wt[c][dd])
  float* dout = Wt + H * (d + 1);   // [RT][d]
  float* hn = dout + RT * d;        // [RT][H]   normalised hidden (gamma*hhat+beta)
  float* hh = hn + RT * H;          // [RT][H]   hhat
  const int v = blockIdx.x;
  const int R = in.B * T_;
  const int rb0 = blockIdx.y * rows_per_block, rb1 = min(R, rb0 + rows_per_block);
  for (int i = threadIdx.x; i < d * H; i += NT) {
    const int dd = i / H, c = i - dd * H;
    Wt[c * (d + 1) + dd] = W4[((long long)v * d + dd) * H + c];
  }
  const int c = threadIdx.x & (H - 1), rl = threadIdx.x >> 6;
  const float w0 = in.W0[(v * H + c) * 2], w1 = in.W0[(v * H + c) * 2 + 1], bb = in.b0[v * H + c];
  const float mu = mean[v * H + c], rs = rstd[v * H + c], ga = gamma[v * H + c], be = beta[v * H + c];
  float acc_db = 0.f, acc_dg = 0.f;
  // dW4 accumulators: thread owns (dd = tid % d .. ) — handle d <= NT by lanes over rows, d > NT by looping dd chunks
  const int dthreads = min(d, NT);
  const int lanes = NT / dthreads;
  const int my_dd = threadIdx.x % dthreads, my_lane = threadIdx.x / dthreads;
  float aw[H];   // dW4[my_dd][0..63] partial over this block's rows (d <= NT)
  float ab = 0.f;
#pragma unroll
  for (int j = 0; j < H; ++j) aw[j] = 0.f;
  for (int r0 = rb0; r0 < rb1; r0 += RT) {
    __syncthreads();
    // stage dout (masked cells -> 0) and the recomputed hidden
    for (int i = threadIdx.x; i < RT * d; i += NT) {
      const int rr = i / d, dd = i - rr * d;
      const int r = r0 + rr;
      float g = 0.f;
      if (r < rb1) {
        const int b = r / T_, t = r - b * T_;
        const float* row = in.xs + (long long)r * (2 * V + 1);
        const bool masked = (row[2 * V] == 1.f) || (row[V + v] == -1.f);
        if (!masked) g = dx_ld(dpsi + ((((long long)b * (T_ + 1) + t) * (V + 1)) + v) * d + dd);
      }
      dout[i] = g;
    }
    for (int rr = rl; rr < RT; rr += 4) {
      const int r = r0 + rr;
      float hhat = 0.f, hnv = 0.f;
      if (r < rb1) {
        float val, cnte; int idx; bool em;
        cell_inputs(in, r, v, val, cnte, idx, em);
        const float h = fmaxf(fmaf(w0, val, fmaf(w1, cnte, bb)), 0.f);
        hhat = (h - mu) * rs;
        hnv = fmaf(hhat, ga, be);
      }
      hh[rr * H + c] = hhat;
      hn[rr * H + c] = hnv;
    }
    __syncthreads();
    // dhn for (row lane rl, channel c)
    for (int rr = rl; rr < RT; rr += 4) {
      const int r = r0 + rr;
      if (r >= rb1) break;
      const float* dr = dout + rr * d;
      const float* wc = Wt + c * (d + 1);
      float s = 0.f;
      for (int dd = 0; dd < d; ++dd) s = fmaf(wc[dd], dr[dd], s);
      dhn_ws[((long long)v * R + r) * H + c] = s;
      acc_db += s;
      acc_dg = fmaf(s, hh[rr * H + c], acc_dg);
    }
    // dW4 / db4 accumulation
    if (my_lane < lanes) {
      for (int rr = my_lane; rr < RT; rr += lanes) {
        if (r0 + rr >= rb1) break;
        const float g = dout[rr * d + my_dd];
        ab += g;
        const float4* h4 = reinterpret_cast<const float4*>(hn + rr * H);
#pragma unroll
        for (int c4 = 0; c4 < H / 4; ++c4) {
          const float4 hv = h4[c4];
          aw[4 * c4] = fmaf(g, hv.x, aw[4 * c4]);
          aw[4 * c4 + 1] = fmaf(g, hv.y, aw[4 * c4 + 1]);
          aw[4 * c4 + 2] = fmaf(g, hv.z, aw[4 * c4 + 2]);
          aw[4 * c4 + 3] = fmaf(g, hv.w, aw[4 * c4 + 3]);
        }
      }
    }
  }
  atomicAdd(dbeta + v * H + c, acc_db);
  atomicAdd(dgamma + v * H + c, acc_dg);
  if (my_lane < lanes) {
    atomicAdd(db4 + v * d + my_dd, ab);
#pragma unroll
    for (int j = 0; j < H; ++j) atomicAdd(dW4 + ((long long)v * d + my_dd) * H + j, aw[j]);
  }
}

// Phase B: BatchNorm backward + ReLU + first Linear + count-embedding grads.
//   dh = gamma*rstd*(dhn - (dbeta + hhat*dgamma)/R);  dpre = dh*(pre>0)
__global__ void __launch_bounds__(NT) embed_bwd_b_kernel(EmbedIn in, const float* __restrict__ gamma, const float* __restrict__ mean,
                                                        const float* __restrict__ rstd, const float* __restrict__ dhn_ws,
                                                        const float* __restrict__ dgamma_cur, const float* __restrict__ dbeta_cur,
                                                        float* __restrict__ dW0, float* __restrict__ db0, float* __restrict__ dnobs,
                                                        int rows_per_block, int training) {
  __shared__ float sh[3][4][H];
  __shared__ float snobs[16];
  const int V = in.V, T_ = in.T;
  const int v = blockIdx.x;
  const int R = in.B * T_;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(R, r0 + rows_per_block);
  const int c = threadIdx.x & (H - 1), rl = threadIdx.x >> 6;
  if (threadIdx.x < 16) snobs[threadIdx.x] = 0.f;
  __syncthreads();
  const float w0 = in.W0[(v * H + c) * 2], w1 = in.W0[(v * H + c) * 2 + 1], bb = in.b0[v * H + c];
  const float mu = mean[v * H + c], rs = rstd[v * H + c], ga = gamma[v * H + c];
  const float mdb = training ? dbeta_cur[v * H + c] / R : 0.f;
  const float mdg = training ? dgamma_cur[v * H + c] / R : 0.f;
  float a0 = 0.f, a1 = 0.f, ab = 0.f;
  for (int r = r0 + rl; r < r1; r += 4) {
    float val, cnte; int idx; bool em;
    cell_inputs(in, r, v, val, cnte, idx, em);
    const float pre = fmaf(w0, val, fmaf(w1, cnte, bb));
    const float h = fmaxf(pre, 0.f);
    const float hhat = (h - mu) * rs;
    const float dhn = dhn_ws[((long long)v * R + r) * H + c];
    const float dh = ga * rs * (dhn - mdb - hhat * mdg);
    const float dpre = pre > 0.f ? dh : 0.f;
    a0 = fmaf(dpre, val, a0);
    a1 = fmaf(dpre, cnte, a1);
    ab += dpre;
    // d(count embedding) = sum_c dpre*w1 : reduce over the 64 channels = 2 warps
    float dc = dx_warp_sum(dpre * w1);
    if ((threadIdx.x & 31) == 0) atomicAdd(&snobs[idx], dc);
  }
  sh[0][rl][c] = a0; sh[1][rl][c] = a1; sh[2][rl][c] = ab;
  __syncthreads();
  if (rl == 0) {
    a0 = sh[0][0][c] + sh[0][1][c] + sh[0][2][c] + sh[0][3][c];
    a1 = sh[1][0][c] + sh[1][1][c] + sh[1][2][c] + sh[1][3][c];
    ab = sh[2][0][c] + sh[2][1][c] + sh[2][2][c] + sh[2][3][c];
    atomicAdd(dW0 + (v * H + c) * 2, a0);
    atomicAdd(dW0 + (v * H + c) * 2 + 1, a1);
    atomicAdd(db0 + v * H + c, ab);
  }
  if (threadIdx.x < 16 && snobs[threadIdx.x] != 0.f) atomicAdd(dnobs + threadIdx.x, snobs[threadIdx.x]);
}

int chunking(int V, int R, int min_rows, int& rows_per_block) {
  int nch = max(1, (4 * 148 + V - 1) / V);
  nch = min(nch, max(1, R / min_rows));
  rows_per_block = (R + nch - 1) / nch;
  rows_per_block = ((rows_per_block + RT - 1) / RT) * RT;
  return (R + rows_per_block - 1) / rows_per_block;
}

}  // namespace

extern "C" {

/* Forward. stats_ws: [V,64,2] doubles (zeroed by this call); mean/rstd: [V,64] f32 outputs (saved for backward).
 * training=1: batch statistics (+ running update when run_mean != NULL); training=0: running statistics. */
int dx_embed_fwd(const float* xs, int B, int T, int V, int d, const float* W0, const float* b0, const float* gamma,
                 const float* beta, float* run_mean, float* run_var, const float* W4, const float* b4, const float* nobs,
                 const float* special, const float* tab, void* psi, int act_dtype, double* stats_ws, float* mean,
                 float* rstd, int training, void* stream) {
  DX_CHECK_ARG(xs && W0 && b0 && gamma && beta && W4 && b4 && nobs && special && tab && psi && mean && rstd,
               "dx_embed_fwd: null argument");
  DX_CHECK_ARG(d % 8 == 0 && d <= 256, "dx_embed_fwd: d_embedding must be a multiple of 8 and <= 256 (got %d)", d);
  cudaStream_t st = (cudaStream_t)stream;
  EmbedIn in{xs, B, T, V, d, W0, b0, nobs};
  const int R = B * T;
  if (training) {
    DX_CHECK_ARG(stats_ws, "dx_embed_fwd: stats workspace required in training mode");
    DX_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * V * H * 2, st));
    int rpb;
    const int nch = chunking(V, R, 256, rpb);
    embed_stats_kernel<<<dim3(V, nch), NT, 0, st>>>(in, stats_ws, rpb);
    DX_LAUNCH_CHECK();
    embed_bnfinal_kernel<<<dx_ceil_div(V * H, 256), 256, 0, st>>>(stats_ws, V * H, R, mean, rstd, run_mean, run_var, 0.1f);
    DX_LAUNCH_CHECK();
  } else {
    DX_CHECK_ARG(run_mean && run_var, "dx_embed_fwd: eval mode needs running statistics");
    embed_bn_eval_kernel<<<dx_ceil_div(V * H, 256), 256, 0, st>>>(run_mean, run_var, V * H, mean, rstd);
    DX_LAUNCH_CHECK();
  }
  const size_t smem = (size_t)(d * (H + 1) + d + RT * H + RT) * sizeof(float);
  dim3 grid(V, dx_ceil_div(R, RT));
  if (act_dtype == DX_BF16) {
    auto k = embed_apply_kernel<bf16>;
    if (smem > 48 * 1024) DX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, NT, smem, st>>>(in, gamma, beta, mean, rstd, W4, b4, special, (bf16*)psi);
    DX_LAUNCH_CHECK();
    embed_special_kernel<bf16><<<148 * 4, NT, 0, st>>>(in, tab, special, (bf16*)psi);
  } else {
    auto k = embed_apply_kernel<float>;
    if (smem > 48 * 1024) DX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<grid, NT, smem, st>>>(in, gamma, beta, mean, rstd, W4, b4, special, (float*)psi);
    DX_LAUNCH_CHECK();
    embed_special_kernel<float><<<148 * 4, NT, 0, st>>>(in, tab, special, (float*)psi);
  }
  DX_LAUNCH_CHECK();
  return DX_OK;
}

/* Backward. Gradients are ACCUMULATED into dW0,db0,dgamma,dbeta,dW4,db4,dnobs,dspecial (f32); dtab [B,d] is written.
 * dhn_ws: [V, B*T, 64] f32 scratch; dgb_ws: [2, V, 64] f32 scratch (this step's dgamma / dbeta). */
int dx_embed_bwd(const float* xs, int B, int T, int V, int d, const float* W0, const float* b0, const float* gamma,
                 const float* beta, const float* W4, const float* nobs, const float* mean, const float* rstd,
                 const void* dpsi, int act_dtype, float* dhn_ws, float* dgb_ws, float* dW0, float* db0, float* dgamma,
                 float* dbeta, float* dW4, float* db4, float* dnobs, float* dspecial, float* dtab, int training,
                 void* stream) {
  DX_CHECK_ARG(xs && dpsi && dhn_ws && dgb_ws && dW0 && db0 && dgamma && dbeta && dW4 && db4 && dnobs && dspecial && dtab,
               "dx_embed_bwd: null argument");
  DX_CHECK_ARG(d % 8 == 0 && d <= 256, "dx_embed_bwd: d_embedding must be a multiple of 8 and <= 256 (got %d)", d);
  cudaStream_t st = (cudaStream_t)stream;
  EmbedIn in{xs, B, T, V, d, W0, b0, nobs};
  const int R = B * T;
  float* dg_cur = dgb_ws;
  float* db_cur = dgb_ws + V * H;
  DX_CUDA(cudaMemsetAsync(dgb_ws, 0, sizeof(float) * 2 * V * H, st));
  int rpb;
  const int nch = chunking(V, R, RT, rpb);
  const size_t smem = (size_t)(H * (d + 1) + RT * d + 2 * RT * H) * sizeof(float);
  if (act_dtype == DX_BF16) {
    embed_special_bwd_kernel<bf16><<<B, NT, 0, st>>>(in, (const bf16*)dpsi, dspecial, dtab);
    DX_LAUNCH_CHECK();
    auto k = embed_bwd_a_kernel<bf16>;
    if (smem > 48 * 1024) DX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<dim3(V, nch), NT, smem, st>>>(in, gamma, beta, mean, rstd, W4, (const bf16*)dpsi, dhn_ws, dg_cur, db_cur, dW4, db4, rpb);
  } else {
    embed_special_bwd_kernel<float><<<B, NT, 0, st>>>(in, (const float*)dpsi, dspecial, dtab);
    DX_LAUNCH_CHECK();
    auto k = embed_bwd_a_kernel<float>;
    if (smem > 48 * 1024) DX_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k<<<dim3(V, nch), NT, smem, st>>>(in, gamma, beta, mean, rstd, W4, (const float*)dpsi, dhn_ws, dg_cur, db_cur, dW4, db4, rpb);
  }
  DX_LAUNCH_CHECK();
  embed_bwd_b_kernel<<<dim3(V, nch), NT, 0, st>>>(in, gamma, mean, rstd, dhn_ws, dg_cur, db_cur, dW0, db0, dnobs, rpb, training);
  DX_LAUNCH_CHECK();
  // fold this step's dgamma/dbeta into the accumulated parameter grads
  extern int dx_axpy(const void*, void*, int64_t, float, int, int, void*);
  int rc = dx_axpy(dg_cur, dgamma, (int64_t)V * H, 1.f, 1, DX_F32, stream);
  if (rc) return rc;
  return dx_axpy(db_cur, dbeta, (int64_t)V * H, 1.f, 1, DX_F32, stream);
}

}  // extern "C"
