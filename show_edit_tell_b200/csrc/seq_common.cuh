// Pieces shared by the EditNet and DCNet sequence drivers: workspace arena, token sampling for the
// rollouts (editnet_rl.py:514-546 / dcnet_rl.py:313-343), the packed cross-entropy
// (editnet.py:571-577) and the SCST criterion (editnet_rl.py:557-573).  Header-only, internal linkage.
#pragma once
#include <string.h>

#include <string>
#include <vector>

#include "../../include/set_b200.h"
#include "cells.cuh"
#include "gemm.cuh"

namespace set {
namespace {

struct Arena {
  char* base = nullptr;
  size_t off = 0;
  struct Entry { std::string name; size_t off, bytes; };
  std::vector<Entry> entries;
  template <typename T>
  T* take(const char* name, size_t n) {
    const size_t bytes = (n * sizeof(T) + 255) & ~size_t(255);
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    entries.push_back({name, off, n * sizeof(T)});
    off += bytes;
    return p;
  }
};

__global__ void fill_kernel(float* p, long n, float v) {
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (long)gridDim.x * blockDim.x) p[x] = v;
}
__global__ void fill_i64_kernel(int64_t* p, long n, int64_t v) {
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (long)gridDim.x * blockDim.x) p[x] = v;
}

int batch_sizes(const SetSeqShape& s, const int* dec_len_host, std::vector<int>& bt) {
  bt.assign(s.T, 0);
  for (int i = 0; i < s.B; ++i) {
    SET_REQUIRE(dec_len_host[i] >= 0 && dec_len_host[i] <= s.T, "decode_len out of range");
    if (i > 0) SET_REQUIRE(dec_len_host[i] <= dec_len_host[i - 1], "decode_len must be sorted descending");
    for (int t = 0; t < dec_len_host[i]; ++t) bt[t]++;
  }
  return SET_OK;
}

// ------------------------------------------------------------------ rollout sampling
// log_softmax + greedy / multinomial / forced choice + finished bookkeeping for one step
// (editnet_rl.py:514-546).  One block per row.
__global__ void __launch_bounds__(256) sample_step_kernel(const float* __restrict__ logits, int V, int B, int T, int t,
                                                          int mode, const int64_t* __restrict__ forced,
                                                          uint64_t seed, int64_t end_token,
                                                          int* __restrict__ unfinished, int* __restrict__ unf_count,
                                                          int64_t* __restrict__ it_out, int64_t* __restrict__ seq,
                                                          float* __restrict__ slp, float* __restrict__ lse_out,
                                                          int64_t* __restrict__ tok_raw) {
  __shared__ float red[40];
  __shared__ float csum[256];
  __shared__ int s_idx;
  const int i = blockIdx.x, tid = threadIdx.x;
  const bool live = (t == 0) || (unf_count[t] > 0);   // the reference leaves the loop once every row finished (:546)
  if (!live) {
    if (tid == 0) { it_out[i] = 0; tok_raw[(long)t * B + i] = -1; }
    return;
  }
  const float* x = logits + (long)i * V;
  float m = -INFINITY;
  int am = 0x7fffffff;
  for (int v = tid; v < V; v += blockDim.x) {
    const float xv = x[v];
    if (xv > m) { m = xv; am = v; }
  }
  // block arg-max (first index on ties)
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, m, o);
    const int oa = __shfl_xor_sync(0xffffffffu, am, o);
    if (om > m || (om == m && oa < am)) { m = om; am = oa; }
  }
  __shared__ float wm[8];
  __shared__ int wa[8];
  if ((tid & 31) == 0) { wm[tid >> 5] = m; wa[tid >> 5] = am; }
  __syncthreads();
  m = wm[0]; am = wa[0];
  for (int k = 1; k < (int)(blockDim.x >> 5); ++k)
    if (wm[k] > m || (wm[k] == m && wa[k] < am)) { m = wm[k]; am = wa[k]; }
  float part = 0.f;
  for (int v = tid; v < V; v += blockDim.x) part += expf(x[v] - m);
  const float sum = block_sum(part, red);
  const float lse = m + logf(sum);
  int tok = am;
  if (mode == 1) {
    // inverse CDF over contiguous per-thread chunks
    const float u = philox_uniform(seed, kSiteSample, (uint64_t)t * B + i) * sum;
    const int chunk = (V + blockDim.x - 1) / blockDim.x;
    const int v0 = tid * chunk, v1 = min(V, v0 + chunk);
    float loc = 0.f;
    for (int v = v0; v < v1; ++v) loc += expf(x[v] - m);
    csum[tid] = loc;
    if (tid == 0) s_idx = V - 1;
    __syncthreads();
    if (tid == 0) {
      float run = 0.f;
      int sel_t = blockDim.x - 1;
      for (int k = 0; k < (int)blockDim.x; ++k) {
        if (run + csum[k] > u) { sel_t = k; break; }
        run += csum[k];
      }
      const int a0 = sel_t * chunk, a1 = min(V, a0 + chunk);
      int pick = max(a1 - 1, 0);
      for (int v = a0; v < a1; ++v) {
        run += expf(x[v] - m);
        if (run > u) { pick = v; break; }
      }
      s_idx = pick;
    }
    __syncthreads();
    tok = s_idx;
  } else if (mode == 2) {
    tok = (int)forced[(long)i * T + t];
  }
  if (tid == 0) {
    const float lp = x[tok] - m - logf(sum);
    int64_t tk = (tok == end_token) ? 0 : tok;
    const int unf = (t == 0) ? (tk > 0) : (unfinished[i] && tk > 0);
    if (!unf) tk = 0;
    seq[(long)i * T + t] = tk;
    slp[(long)i * T + t] = lp;
    unfinished[i] = unf;
    it_out[i] = tk;
    lse_out[(long)t * B + i] = lse;
    tok_raw[(long)t * B + i] = tok;
    if (unf) atomicAdd(&unf_count[t + 1], 1);
  }
}

// Scheduled sampling (editnet.py:508-520): the token fed at step t >= 1 is, with probability ss_prob,
// a draw from multinomial(exp(logits of step t-1)) (the reference exponentiates raw logits, :517;
// multinomial normalises, so this is softmax sampling), else the ground-truth token.  `replay`
// (batch-major [B][Wc]) overrides the choice (tests).  One block per row; writes the time-major feed
// `it_t` and the batch-major record `fed` (consumed by the backward pass as "the captions").
__global__ void __launch_bounds__(256) ss_choose_kernel(const float* __restrict__ logits_prev, int V, int B, int Wc,
                                                        int t, float ss_prob, uint64_t seed,
                                                        const int64_t* __restrict__ caps,
                                                        const int64_t* __restrict__ replay,
                                                        int64_t* __restrict__ it_t, int64_t* __restrict__ fed) {
  __shared__ float red[40];
  __shared__ float csum[256];
  __shared__ float wm[8];
  const int i = blockIdx.x, tid = threadIdx.x;
  const int64_t gt = caps[(long)i * Wc + t];
  int64_t tok = gt;
  bool sample = false;
  if (replay) tok = replay[(long)i * Wc + t];
  else if (t >= 1 && ss_prob > 0.f) sample = philox_uniform(seed, kSiteSample, (uint64_t)(2 * t) * B + i) < ss_prob;
  if (sample) {   // block-uniform
    const float* x = logits_prev + (long)i * V;
    float m = -INFINITY;
    for (int v = tid; v < V; v += blockDim.x) m = fmaxf(m, x[v]);
    m = warp_max(m);
    if ((tid & 31) == 0) wm[tid >> 5] = m;
    __syncthreads();
    m = wm[0];
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) m = fmaxf(m, wm[k]);
    const int chunk = (V + blockDim.x - 1) / blockDim.x;
    const int v0 = tid * chunk, v1 = min(V, v0 + chunk);
    float loc = 0.f;
    for (int v = v0; v < v1; ++v) loc += expf(x[v] - m);
    csum[tid] = loc;
    const float sum = block_sum(loc, red);
    if (tid == 0) {
      const float u = philox_uniform(seed, kSiteSample, (uint64_t)(2 * t + 1) * B + i) * sum;
      float run = 0.f;
      int sel_t = blockDim.x - 1;
      for (int k = 0; k < (int)blockDim.x; ++k) {
        if (run + csum[k] > u) { sel_t = k; break; }
        run += csum[k];
      }
      const int a0 = sel_t * chunk, a1 = min(V, a0 + chunk);
      int pick = max(a1 - 1, 0);
      for (int v = a0; v < a1; ++v) {
        run += expf(x[v] - m);
        if (run > u) { pick = v; break; }
      }
      tok = pick;
    }
  }
  if (tid == 0) {
    it_t[i] = tok;
    fed[(long)i * Wc + t] = tok;
  }
}

// d logits[t][i][v] = d_slp[i][t] * (1[v == tok] - softmax_v)   (in place over the saved logits)
__global__ void rollout_dlogits_kernel(float* __restrict__ logits, int V, int B, int T,
                                       const float* __restrict__ d_slp, const float* __restrict__ lse,
                                       const int64_t* __restrict__ tok_raw) {
  const long row = blockIdx.x;  // t*B + i
  const int t = (int)(row / B), i = (int)(row % B);
  const int64_t tok = tok_raw[row];
  const float g = (tok >= 0) ? d_slp[(long)i * T + t] : 0.f;
  const float l = lse[row];
  float* x = logits + row * V;
  for (int v = threadIdx.x; v < V; v += blockDim.x) {
    const float pr = (tok >= 0) ? expf(x[v] - l) : 0.f;
    x[v] = g * ((v == tok ? 1.f : 0.f) - pr);
  }
}

// packed cross-entropy (editnet.py:571-577)
__global__ void __launch_bounds__(256) xe_loss_kernel(int B, int T, int V, int Wc, long stride_b, long stride_t,
                                                      const float* __restrict__ pred,
                                                      const int64_t* __restrict__ caps,
                                                      const int* __restrict__ dec_len, float inv_count,
                                                      float* __restrict__ loss_out, float* __restrict__ dpred) {
  __shared__ float red[40];
  const int i = blockIdx.x / T, t = blockIdx.x % T;
  const long off = (long)i * stride_b + (long)t * stride_t;
  const bool valid = dec_len[i] > t;
  if (!valid) {
    if (dpred) for (int v = threadIdx.x; v < V; v += blockDim.x) dpred[off + v] = 0.f;
    return;
  }
  if (inv_count <= 0.f) {
    int cnt = 0;
    for (int k = 0; k < B; ++k) cnt += min(dec_len[k], T);
    inv_count = 1.f / (float)cnt;
  }
  const float* x = pred + off;
  float m = -INFINITY;
  for (int v = threadIdx.x; v < V; v += blockDim.x) m = fmaxf(m, x[v]);
  m = warp_max(m);
  __shared__ float wm[8];
  if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = m;
  __syncthreads();
  m = wm[0];
  for (int k = 1; k < (int)(blockDim.x >> 5); ++k) m = fmaxf(m, wm[k]);
  float part = 0.f;
  for (int v = threadIdx.x; v < V; v += blockDim.x) part += expf(x[v] - m);
  const float sum = block_sum(part, red);
  const float lse = m + logf(sum);
  const int64_t tgt = caps[(long)i * Wc + t + 1];
  if (threadIdx.x == 0) atomicAdd(loss_out, (lse - x[tgt]) * inv_count);
  if (dpred) {
    for (int v = threadIdx.x; v < V; v += blockDim.x) {
      const float pr = expf(x[v] - lse);
      dpred[off + v] = (pr - (v == tgt ? 1.f : 0.f)) * inv_count;
    }
  }
}

__global__ void xe_count_kernel(int B, int T, const int* __restrict__ dec_len, float* __restrict__ loss_out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    int cnt = 0;
    for (int k = 0; k < B; ++k) cnt += min(dec_len[k], T);
    loss_out[1] = (float)cnt;
  }
}

// RewardCriterion (editnet_rl.py:557-573), single block
__global__ void __launch_bounds__(256) reward_kernel(int B, int T, const float* __restrict__ slp,
                                                     const int64_t* __restrict__ seq, const float* __restrict__ reward,
                                                     float* __restrict__ loss_out, float* __restrict__ dlp) {
  __shared__ float red[40];
  float num = 0.f, den = 0.f;
  for (int x = threadIdx.x; x < B * T; x += blockDim.x) {
    const int t = x % T;
    const float mk = (t == 0) ? 1.f : (seq[x - 1] > 0 ? 1.f : 0.f);
    num += -slp[x] * reward[x] * mk;
    den += mk;
  }
  num = block_sum(num, red);
  den = block_sum(den, red);
  if (threadIdx.x == 0) loss_out[0] = num / den;
  if (dlp)
    for (int x = threadIdx.x; x < B * T; x += blockDim.x) {
      const int t = x % T;
      const float mk = (t == 0) ? 1.f : (seq[x - 1] > 0 ? 1.f : 0.f);
      dlp[x] = -reward[x] * mk / den;
    }
}

}  // namespace
}  // namespace set
