// Batched beam search on the device: the expansion step of evaluate() (editnet.py:654-696) and of
// the EditNet + DCNet ensemble evaluate_full() (eval/eval xe/eval_full.py:151-191) for MANY images at once.  The
// reference searches one image at a time and keeps the bookkeeping on the host (python lists, a .tolist() per step);
// here every image owns `K` rows of one step session, and ONE kernel per step does, per image: log-softmax of the live
// beams' scores (or log of the mean of the two networks' softmaxes), cumulative add, top-k over (live beams x
// vocabulary), sequence extension, <end> bookkeeping (completed beams leave the search, k shrinks) and the row indices
// the state re-gather needs.  No host round trip inside the search.
#include "../../include/set_b200.h"
#include "common.cuh"

namespace set {
namespace {

constexpr int kBeamMaxK = 8;
constexpr int kBeamThreads = 256;

struct Cand { float v; int idx; };

__device__ __forceinline__ bool cand_better(float v, int idx, float bv, int bidx) {
  return v > bv || (v == bv && idx < bidx);     // ties: the smaller flat index
}

// one CTA per image
__global__ void __launch_bounds__(kBeamThreads) beam_expand_kernel(
    int K, int V, int step, int Lmax, long long end_tok, const float* __restrict__ logits_e,
    const float* __restrict__ logits_d, int* __restrict__ k_live, float* __restrict__ beam_scores,
    const long long* __restrict__ seq_in, long long* __restrict__ seq_out, long long* __restrict__ next_tokens,
    int* __restrict__ src_row, int* __restrict__ n_complete, float* __restrict__ complete_scores,
    long long* __restrict__ complete_seqs, int* __restrict__ complete_len, int* __restrict__ live_images) {
  __shared__ float red[40];
  __shared__ float s_m[2][kBeamMaxK], s_ls[2][kBeamMaxK];      // per beam: max and log-sum-exp shift of each network
  __shared__ float cv[kBeamThreads][kBeamMaxK];
  __shared__ int ci[kBeamThreads][kBeamMaxK];
  __shared__ float win_v[kBeamMaxK];
  __shared__ int win_i[kBeamMaxK];
  __shared__ float wred_v[8];
  __shared__ int wred_i[8], wred_t[8];
  const int img = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int k = k_live[img];
  if (k == 0) {
    for (int j = tid; j < K; j += blockDim.x) { src_row[img * K + j] = img * K + j; next_tokens[img * K + j] = 0; }
    return;
  }
  const int nb = (step == 1) ? 1 : k;      // step 1: all beams are identical, the reference expands beam 0 only (:660-661)
  // ---- softmax statistics of every considered beam, both networks
  for (int net = 0; net < (logits_d ? 2 : 1); ++net) {
    const float* lg = net ? logits_d : logits_e;
    for (int j = 0; j < nb; ++j) {
      const float* row = lg + (long)(img * K + j) * V;
      float m = -INFINITY;
      for (int v = tid; v < V; v += blockDim.x) m = fmaxf(m, row[v]);
      m = warp_max(m);
      __syncthreads();
      if (lane == 0) red[wid] = m;
      __syncthreads();
      m = red[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w) m = fmaxf(m, red[w]);
      float s = 0.f;
      for (int v = tid; v < V; v += blockDim.x) s += expf(row[v] - m);
      s = block_sum(s, red);
      if (tid == 0) { s_m[net][j] = m; s_ls[net][j] = logf(s); }
    }
  }
  __syncthreads();
  // ---- per-thread top-k over the flat (beam, word) candidates
  float lv[kBeamMaxK]; int li[kBeamMaxK];
#pragma unroll
  for (int r = 0; r < kBeamMaxK; ++r) { lv[r] = -INFINITY; li[r] = 0x7fffffff; }
  for (int j = 0; j < nb; ++j) {
    const float base = beam_scores[img * K + j];
    const float* re = logits_e + (long)(img * K + j) * V;
    const float* rd = logits_d ? logits_d + (long)(img * K + j) * V : nullptr;
    const float me = s_m[0][j], lse = s_ls[0][j];
    const float md = rd ? s_m[1][j] : 0.f, lsd = rd ? s_ls[1][j] : 0.f;
    for (int v = tid; v < V; v += blockDim.x) {
      float lp;
      if (rd) {
        // log((softmax_e + softmax_d) / 2), eval_full.py:151-153
        lp = logf(0.5f * (expf(re[v] - me - lse) + expf(rd[v] - md - lsd)));
      } else {
        lp = re[v] - me - lse;               // F.log_softmax, editnet.py:654
      }
      const float val = base + lp;           // :657
      const int idx = j * V + v;
      // (lists of the full kBeamMaxK entries keep the indices static, i.e. in registers)
      if (cand_better(val, idx, lv[kBeamMaxK - 1], li[kBeamMaxK - 1])) {
        lv[kBeamMaxK - 1] = val; li[kBeamMaxK - 1] = idx;
#pragma unroll
        for (int r = kBeamMaxK - 1; r > 0; --r)
          if (cand_better(lv[r], li[r], lv[r - 1], li[r - 1])) {
            const float tv = lv[r]; lv[r] = lv[r - 1]; lv[r - 1] = tv;
            const int ti = li[r]; li[r] = li[r - 1]; li[r - 1] = ti;
          }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < kBeamMaxK; ++r) { cv[tid][r] = lv[r]; ci[tid][r] = li[r]; }
  __syncthreads();
  // ---- merge: k rounds of block argmax over the threads' current heads (descending order, like topk(sorted=True))
  int head = 0;
  for (int r = 0; r < k; ++r) {
    float v = head < kBeamMaxK ? cv[tid][head] : -INFINITY;
    int idx = head < kBeamMaxK ? ci[tid][head] : 0x7fffffff;
    int t = tid;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
      const int ot = __shfl_xor_sync(0xffffffffu, t, o);
      if (cand_better(ov, oi, v, idx)) { v = ov; idx = oi; t = ot; }
    }
    if (lane == 0) { wred_v[wid] = v; wred_i[wid] = idx; wred_t[wid] = t; }
    __syncthreads();
    if (tid == 0) {
      float bv = wred_v[0]; int bi = wred_i[0], bt = wred_t[0];
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (cand_better(wred_v[w], wred_i[w], bv, bi)) { bv = wred_v[w]; bi = wred_i[w]; bt = wred_t[w]; }
      win_v[r] = bv; win_i[r] = bi; wred_t[0] = bt;
    }
    __syncthreads();
    if (tid == wred_t[0]) ++head;
    __syncthreads();
  }
  // ---- bookkeeping (editnet.py:666-696): extend sequences, set completed beams aside, compact the live ones
  __shared__ int s_slot[kBeamMaxK], s_prev[kBeamMaxK];
  if (tid == 0) {
    int live = 0, nc = n_complete[img];
    for (int r = 0; r < k; ++r) {
      const int prev = win_i[r] / V, word = win_i[r] % V;
      s_prev[r] = prev;
      if (word == (int)end_tok) {
        s_slot[r] = -1 - nc;                       // completed: goes to completion slot nc
        complete_scores[img * K + nc] = win_v[r];
        complete_len[img * K + nc] = step + 1;
        ++nc;
      } else {
        s_slot[r] = live;
        beam_scores[img * K + live] = win_v[r];
        next_tokens[img * K + live] = word;
        src_row[img * K + live] = img * K + prev;
        ++live;
      }
    }
    for (int j = live; j < K; ++j) { next_tokens[img * K + j] = 0; src_row[img * K + j] = img * K + j; }
    n_complete[img] = nc;
    k_live[img] = live;
    if (live == 0) atomicSub(live_images, 1);
  }
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    const long long* src = seq_in + (long)(img * K + s_prev[r]) * Lmax;
    long long* dst = s_slot[r] >= 0 ? seq_out + (long)(img * K + s_slot[r]) * Lmax
                                    : complete_seqs + (long)(img * K + (-1 - s_slot[r])) * Lmax;
    for (int p = tid; p < step && p < Lmax; p += blockDim.x) dst[p] = src[p];
    if (tid == 0 && step < Lmax) dst[step] = win_i[r] % V;
  }
}

// rows of the four state tensors follow their beams: out[r] = in[src_row[r]]
__global__ void beam_gather_kernel(int rows, int D4, const int* __restrict__ src_row, const float4* __restrict__ i0,
                                   const float4* __restrict__ i1, const float4* __restrict__ i2, const float4* __restrict__ i3,
                                   float4* __restrict__ o0, float4* __restrict__ o1, float4* __restrict__ o2,
                                   float4* __restrict__ o3) {
  const long total = (long)rows * D4;
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int r = (int)(x / D4), d = (int)(x % D4);
    const long s = (long)src_row[r] * D4 + d;
    o0[x] = i0[s]; o1[x] = i1[s]; o2[x] = i2[s]; o3[x] = i3[s];
  }
}

// result per image (editnet.py:702-713): runaway guard -> first 18 tokens of the first live beam; else the completed beam
// with the highest score (first occurrence)
__global__ void beam_finalize_kernel(int K, int Lmax, int steps_done, const int* __restrict__ k_live,
                                     const long long* __restrict__ seq_live, const float* __restrict__ beam_scores,
                                     const int* __restrict__ n_complete, const float* __restrict__ complete_scores,
                                     const long long* __restrict__ complete_seqs, const int* __restrict__ complete_len,
                                     long long* __restrict__ out_seq, int* __restrict__ out_len, float* __restrict__ out_score) {
  const int img = blockIdx.x;
  const long long* src; int len; float sc;
  if (k_live[img] > 0) {
    src = seq_live + (long)(img * K) * Lmax;
    len = steps_done + 1 < 18 ? steps_done + 1 : 18;
    sc = beam_scores[img * K];
  } else {
    int best = 0;
    for (int j = 1; j < n_complete[img]; ++j)
      if (complete_scores[img * K + j] > complete_scores[img * K + best]) best = j;
    src = complete_seqs + (long)(img * K + best) * Lmax;
    len = complete_len[img * K + best];
    sc = complete_scores[img * K + best];
  }
  for (int p = threadIdx.x; p < Lmax; p += blockDim.x) out_seq[(long)img * Lmax + p] = p < len ? src[p] : 0;
  if (threadIdx.x == 0) { out_len[img] = len; out_score[img] = sc; }
}

}  // namespace
}  // namespace set

using namespace set;

extern "C" {

int set_beam_expand(int N, int K, int V, int step, int Lmax, int64_t end_tok, const float* logits_e, const float* logits_d,
                    int* k_live, float* beam_scores, const int64_t* seq_in, int64_t* seq_out, int64_t* next_tokens,
                    int* src_row, int* n_complete, float* complete_scores, int64_t* complete_seqs, int* complete_len,
                    int* live_images, void* stream) {
  SET_REQUIRE(N > 0 && K >= 1 && K <= kBeamMaxK && V > 1 && step >= 1 && Lmax > step, "bad beam shape");
  SET_REQUIRE(logits_e && k_live && beam_scores && seq_in && seq_out && next_tokens && src_row && n_complete &&
              complete_scores && complete_seqs && complete_len && live_images, "null argument");
  beam_expand_kernel<<<N, kBeamThreads, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      K, V, step, Lmax, (long long)end_tok, logits_e, logits_d, k_live, beam_scores,
      reinterpret_cast<const long long*>(seq_in), reinterpret_cast<long long*>(seq_out),
      reinterpret_cast<long long*>(next_tokens), src_row, n_complete, complete_scores,
      reinterpret_cast<long long*>(complete_seqs), complete_len, live_images);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  return SET_OK;
}

int set_beam_gather(int rows, int D, const int* src_row, const float* in0, const float* in1, const float* in2,
                    const float* in3, float* out0, float* out1, float* out2, float* out3, void* stream) {
  SET_REQUIRE(rows > 0 && D > 0 && D % 4 == 0 && src_row && in0 && in1 && in2 && in3 && out0 && out1 && out2 && out3, "bad args");
  const long total = (long)rows * (D / 4);
  const int blocks = (int)((total + 255) / 256 < 1184 ? (total + 255) / 256 : 1184);
  beam_gather_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      rows, D / 4, src_row, reinterpret_cast<const float4*>(in0), reinterpret_cast<const float4*>(in1),
      reinterpret_cast<const float4*>(in2), reinterpret_cast<const float4*>(in3), reinterpret_cast<float4*>(out0),
      reinterpret_cast<float4*>(out1), reinterpret_cast<float4*>(out2), reinterpret_cast<float4*>(out3));
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  return SET_OK;
}

int set_beam_finalize(int N, int K, int Lmax, int steps_done, const int* k_live, const int64_t* seq_live,
                      const float* beam_scores, const int* n_complete, const float* complete_scores,
                      const int64_t* complete_seqs, const int* complete_len, int64_t* out_seq, int* out_len,
                      float* out_score, void* stream) {
  SET_REQUIRE(N > 0 && K >= 1 && K <= kBeamMaxK && Lmax > 0 && k_live && seq_live && beam_scores && n_complete &&
              complete_scores && complete_seqs && complete_len && out_seq && out_len && out_score, "bad args");
  beam_finalize_kernel<<<N, 64, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      K, Lmax, steps_done, k_live, reinterpret_cast<const long long*>(seq_live), beam_scores, n_complete, complete_scores,
      reinterpret_cast<const long long*>(complete_seqs), complete_len, reinterpret_cast<long long*>(out_seq), out_len,
      out_score);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  return SET_OK;
}

}  // extern "C"
