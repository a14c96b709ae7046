// Self-critical CIDEr-D reward on the device: replaces the GPU -> CPU -> strings -> CIDEr-D ->
// GPU round trip of get_self_critical_reward(), editnet_rl.py:611-646 (+ preprocess_gd :587-600, array_to_str :602-609)
// and the CiderD.compute_score() it calls (pyciderevalcap, un-vendored; algorithm restated in oracle/ciderd_oracle.py).
//
// Integer / hash work, tiny: one CTA per hypothesis (B sampled + B greedy), 256 threads = one per n-gram occurrence
// (a sentence of <= 64 tokens has <= 250 occurrences of n-grams with n = 1..4).  An n-gram of token ids is packed
// exactly into 64 bits (16 bits per token, +1 so that token 0 -- the <end> the reference keeps as a word -- is
// distinct from "absent"); document frequencies come from an open-addressing table built on the host from the
// reference's `coco-train-idxs` pickle format.  All arithmetic in fp64, like the numpy reference.
#include "../../include/set_b200.h"
#include "common.cuh"

namespace set {
namespace {

constexpr int kCdMaxLen = 64;   // tokens per sentence: the loader's caption width is 52 (max_len 50 + <start>/<end>)
constexpr int kCdMaxNg = 4 * kCdMaxLen;
constexpr int kCdThreads = 256;  // >= n-gram occurrences of a 64-token sentence (64 + 63 + 62 + 61 = 250)
constexpr unsigned long long kCdEmpty = ~0ull;

struct CdSent {
  int len, n_ent;
  int tok[kCdMaxLen];
  unsigned long long key[kCdMaxNg];
  double w[kCdMaxNg];          // tf-idf weight of the n-gram (same value on every occurrence)
  unsigned char ord[kCdMaxNg];   // n - 1
  unsigned char first[kCdMaxNg]; // 1 on the first occurrence of the n-gram in the sentence
  double norm[4];
};

__device__ __forceinline__ unsigned long long cd_mix(unsigned long long x) {   // splitmix64 finaliser
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

__device__ __forceinline__ float cd_df(const unsigned long long* __restrict__ keys, const float* __restrict__ vals,
                                       unsigned long long mask, unsigned long long key) {
  unsigned long long s = cd_mix(key) & mask;
  for (;;) {
    const unsigned long long k = keys[s];
    if (k == key) return vals[s];
    if (k == kCdEmpty) return 0.f;
    s = (s + 1) & mask;
  }
}

// n-gram occurrences, term frequencies, tf-idf weights and per-order norms of the sentence in S.tok[0..S.len)
__device__ void cd_build(CdSent& S, const unsigned long long* keys, const float* vals, unsigned long long mask,
                         double ref_len_log) {
  const int t = threadIdx.x, L = S.len;
  int n_ent = 0;
  for (int k = 1; k <= 4; ++k) n_ent += L - k + 1 > 0 ? L - k + 1 : 0;
  if (t == 0) S.n_ent = n_ent;
  if (t < n_ent) {
    int e = t, k = 1;
    while (e >= L - k + 1) { e -= L - k + 1; ++k; }
    unsigned long long key = 0;
    for (int j = 0; j < k; ++j) key |= (unsigned long long)(S.tok[e + j] + 1) << (16 * j);
    S.key[t] = key;
    S.ord[t] = (unsigned char)(k - 1);
  }
  __syncthreads();
  if (t < n_ent) {
    const unsigned long long key = S.key[t];
    int tf = 0, first = 1;
    for (int e = 0; e < n_ent; ++e)
      if (S.key[e] == key) { ++tf; if (e < t) first = 0; }
    const double df = log(fmax(1.0, (double)cd_df(keys, vals, mask, key)));
    S.w[t] = (double)tf * (ref_len_log - df);
    S.first[t] = (unsigned char)first;
  }
  __syncthreads();
  if (t < 4) {
    double s = 0.0;
    for (int e = 0; e < n_ent; ++e)
      if (S.first[e] && S.ord[e] == t) s += S.w[e] * S.w[e];
    S.norm[t] = sqrt(s);
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kCdThreads) ciderd_score_kernel(
    int B, int L, int R, int Wc, const int64_t* __restrict__ gen, const int64_t* __restrict__ greedy,
    const int64_t* __restrict__ all_caps, long long start_tok, long long end_tok, long long pad_tok,
    const unsigned long long* __restrict__ df_keys, const float* __restrict__ df_vals, unsigned long long df_mask,
    double ref_len_log, double sigma, float* __restrict__ scores) {
  __shared__ CdSent H, Rf;
  __shared__ double acc[4], score[4];
  const int h = blockIdx.x, t = threadIdx.x;
  const int img = h % B;
  if (t == 0) {
    // array_to_str (editnet_rl.py:602-609): tokens up to and including the first 0
    const int64_t* row = (h < B ? gen : greedy) + (long)img * L;
    int n = 0;
    for (int i = 0; i < L && n < kCdMaxLen; ++i) {
      H.tok[n++] = (int)row[i];
      if (row[i] == 0) break;
    }
    H.len = n;
  }
  if (t < 4) score[t] = 0.0;
  __syncthreads();
  cd_build(H, df_keys, df_vals, df_mask, ref_len_log);
  for (int r = 0; r < R; ++r) {
    if (t == 0) {
      // preprocess_gd (:587-600): drop <start> / <pad>, <end> -> 0; then array_to_str stops behind the first 0
      const int64_t* row = all_caps + ((long)img * R + r) * Wc;
      int n = 0;
      for (int i = 0; i < Wc && n < kCdMaxLen; ++i) {
        long long w = row[i];
        if (w == start_tok || w == pad_tok) continue;
        if (w == end_tok) w = 0;
        Rf.tok[n++] = (int)w;
        if (w == 0) break;
      }
      Rf.len = n;
    }
    if (t < 4) acc[t] = 0.0;
    __syncthreads();
    cd_build(Rf, df_keys, df_vals, df_mask, ref_len_log);
    if (t < H.n_ent && H.first[t]) {
      const unsigned long long key = H.key[t];
      for (int e = 0; e < Rf.n_ent; ++e)
        if (Rf.first[e] && Rf.key[e] == key) {
          const double wr = Rf.w[e];
          atomicAdd(&acc[H.ord[t]], fmin(H.w[t], wr) * wr);       // clipped: min(h, r) * r
          break;
        }
    }
    __syncthreads();
    if (t < 4) {
      double v = acc[t];
      if (H.norm[t] != 0.0 && Rf.norm[t] != 0.0) v /= H.norm[t] * Rf.norm[t];
      // length = number of bigram occurrences (the `if n == 1: length += term_freq` of the scorer)
      const double delta = (double)((H.len > 1 ? H.len - 1 : 0) - (Rf.len > 1 ? Rf.len - 1 : 0));
      v *= exp(-(delta * delta) / (2.0 * sigma * sigma));
      score[t] += v;
    }
    __syncthreads();
  }
  if (t == 0) scores[h] = (float)((score[0] + score[1] + score[2] + score[3]) / 4.0 / (double)R * 10.0);
}

__global__ void ciderd_reward_kernel(int B, int L, float weight, const float* __restrict__ scores, float* __restrict__ rewards) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x < B * L) {
    const int i = x / L;
    rewards[x] = weight * scores[i] - weight * scores[B + i];      // sample - greedy, broadcast over the steps (:642-644)
  }
}

}  // namespace
}  // namespace set

extern "C" int set_ciderd_reward(int B, int L, int R, int Wc, const int64_t* gen, const int64_t* greedy,
                                 const int64_t* all_caps, int64_t start_tok, int64_t end_tok, int64_t pad_tok,
                                 const uint64_t* df_keys, const float* df_vals, uint64_t df_capacity, double ref_len,
                                 double sigma, float cider_weight, float* scores, float* rewards, void* stream) {
  SET_REQUIRE(B > 0 && L > 0 && R > 0 && Wc > 0 && gen && greedy && all_caps && df_keys && df_vals && scores && rewards, "bad args");
  SET_REQUIRE(df_capacity > 0 && (df_capacity & (df_capacity - 1)) == 0, "table capacity must be a power of two");
  SET_REQUIRE(ref_len > 0 && sigma > 0, "ref_len / sigma");
  SET_REQUIRE(L <= set::kCdMaxLen, "rollouts longer than 64 tokens are not supported by the n-gram kernel");
  SET_REQUIRE(Wc <= set::kCdMaxLen, "reference captions wider than 64 tokens would be truncated by the n-gram kernel");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  set::ciderd_score_kernel<<<2 * B, set::kCdThreads, 0, st>>>(
      B, L, R, Wc, gen, greedy, all_caps, start_tok, end_tok, pad_tok, reinterpret_cast<const unsigned long long*>(df_keys),
      df_vals, df_capacity - 1, log(ref_len), sigma, scores);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  set::ciderd_reward_kernel<<<(B * L + 255) / 256, 256, 0, st>>>(B, L, cider_weight, scores, rewards);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  return SET_OK;
}
