// Element-wise / attention kernels of the decode step.  See cells.cuh for the contract
// of each launcher and the reference lines it reproduces.
#include "cells.cuh"

#include <mutex>

namespace set {

namespace {

constexpr int kThreads = 256;
constexpr int kAttnThreads = 256;    // CTAs are latency-bound: many warps, and several CTAs per sample (see kVisSlices)
constexpr int kVisSlices = 4;        // column slices of the region-feature context / unit slices of the score MLP
constexpr int kCapSlices = 2;        // same for the caption attention

inline int blocks_for(long n, int per_block) {
  long b = (n + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}

// ------------------------------------------------------------------------ embedding
__global__ void embed_fwd_kernel(const int64_t* __restrict__ tokens, long tok_ld, long tok_os,
                                 const float* __restrict__ table, int V, float* __restrict__ out, int n_outer,
                                 int n_inner, int D, int train, uint64_t seed, uint32_t site, long drop_row0,
                                 long drop_os, long drop_is) {
  const int D4 = D >> 2;
  const long total = (long)n_outer * n_inner * D4;
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int e4 = (int)(x % D4);
    const long row = x / D4;
    const int i = (int)(row % n_inner), o = (int)(row / n_inner);
    long tok = tokens[(long)i * tok_ld + (long)o * tok_os];
    tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
    float4 v = __ldg(reinterpret_cast<const float4*>(table + tok * D) + e4);
    v.x = fmaxf(v.x, 0.f); v.y = fmaxf(v.y, 0.f); v.z = fmaxf(v.z, 0.f); v.w = fmaxf(v.w, 0.f);
    if (train) {
      const uint64_t idx = (uint64_t)(drop_row0 + (long)o * drop_os + (long)i * drop_is) * D + (uint64_t)e4 * 4;
      const uint32_t k = drop_keep4(seed, site, idx);
      v.x = (k & 1) ? v.x * 2.f : 0.f; v.y = (k & 2) ? v.y * 2.f : 0.f;
      v.z = (k & 4) ? v.z * 2.f : 0.f; v.w = (k & 8) ? v.w * 2.f : 0.f;
    }
    reinterpret_cast<float4*>(out + row * D)[e4] = v;
  }
}

__global__ void embed_bwd_kernel(const int64_t* __restrict__ tokens, long tok_ld, long tok_os,
                                 const float* __restrict__ out, const float* __restrict__ dout,
                                 float* __restrict__ table_grad, int V, int n_outer, int n_inner, int D,
                                 float scale, const int* __restrict__ row_len) {
  const long total = (long)n_outer * n_inner * D;
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int e = (int)(x % D);
    const long row = x / D;
    const int i = (int)(row % n_inner), o = (int)(row / n_inner);
    if (row_len && row_len[i] <= o) continue;
    if (out[x] > 0.f) {
      const float g = dout[x] * scale;
      long tok = tokens[(long)i * tok_ld + (long)o * tok_os];
      tok = tok < 0 ? 0 : (tok >= V ? V - 1 : tok);
      if (g != 0.f) atomicAdd(table_grad + tok * D + e, g);
    }
  }
}

// ---------------------------------------------------------------------- region prep
__global__ void region_mean_kernel(const float* __restrict__ feats, float* __restrict__ out, int B, int R, int F) {
  const long total = (long)B * F;
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int f = (int)(x % F);
    const long i = x / F;
    float s = 0.f;
    for (int r = 0; r < R; ++r) s += feats[(i * R + r) * F + f];
    out[x] = s / (float)R;
  }
}

__global__ void region_count_kernel(const float* __restrict__ feats, int* __restrict__ nreg, int R, int F) {
  // one block per sample; warp per region row
  __shared__ int cnt;
  if (threadIdx.x == 0) cnt = 0;
  __syncthreads();
  const int i = blockIdx.x, lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int r = wid; r < R; r += nw) {
    float s = 0.f;
    for (int f = lane; f < F; f += 32) s += feats[((long)i * R + r) * F + f];
    s = warp_sum(s);
    if (lane == 0 && s != 0.f) atomicAdd(&cnt, 1);
  }
  __syncthreads();
  if (threadIdx.x == 0) nreg[i] = cnt;
}

__global__ void zero_pad_regions_kernel(float* __restrict__ fe, const int* __restrict__ nreg, int B, int R, int D) {
  const long total = (long)B * R * D;
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const long row = x / D;
    const int r = (int)(row % R), i = (int)(row / R);
    if (r >= nreg[i]) fe[x] = 0.f;
  }
}

__global__ void vis_dropout_fwd_kernel(const float* __restrict__ fe_pre, float* __restrict__ fe_t, int T, long n,
                                       uint64_t seed) {
  // n = B*R*D (multiple of 4); fe_t[t][x]
  const long n4 = n >> 2;
  const long total = (long)T * n4;
  for (long y = (long)blockIdx.x * blockDim.x + threadIdx.x; y < total; y += (long)gridDim.x * blockDim.x) {
    const long x4 = y % n4;
    float4 v = __ldg(reinterpret_cast<const float4*>(fe_pre) + x4);
    const uint32_t k = drop_keep4(seed, kSiteVis, (uint64_t)y * 4);
    v.x = (k & 1) ? v.x * 2.f : 0.f; v.y = (k & 2) ? v.y * 2.f : 0.f;
    v.z = (k & 4) ? v.z * 2.f : 0.f; v.w = (k & 8) ? v.w * 2.f : 0.f;
    reinterpret_cast<float4*>(fe_t)[y] = v;
  }
}

__global__ void vis_dropout_bwd_kernel(const float* __restrict__ fe_pre, const float* __restrict__ dfe_t,
                                       float* __restrict__ dfe_pre, const int* __restrict__ dec_len, int T, int B,
                                       long per_sample, uint64_t seed) {
  // per_sample = R*D (multiple of 4)
  const long ps4 = per_sample >> 2;
  const long total = (long)B * ps4;
  for (long y = (long)blockIdx.x * blockDim.x + threadIdx.x; y < total; y += (long)gridDim.x * blockDim.x) {
    const int i = (int)(y / ps4);
    const float4 f = __ldg(reinterpret_cast<const float4*>(fe_pre) + y);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int Ti = dec_len ? min(dec_len[i], T) : T;
    for (int t = 0; t < Ti; ++t) {
      const long z = (long)t * total + y;
      const float4 d = __ldg(reinterpret_cast<const float4*>(dfe_t) + z);
      const uint32_t k = drop_keep4(seed, kSiteVis, (uint64_t)z * 4);
      if (k & 1) acc.x += d.x; if (k & 2) acc.y += d.y; if (k & 4) acc.z += d.z; if (k & 8) acc.w += d.w;
    }
    acc.x = f.x > 0.f ? acc.x * 2.f : 0.f; acc.y = f.y > 0.f ? acc.y * 2.f : 0.f;
    acc.z = f.z > 0.f ? acc.z * 2.f : 0.f; acc.w = f.w > 0.f ? acc.w * 2.f : 0.f;
    reinterpret_cast<float4*>(dfe_pre)[y] = acc;
  }
}

__global__ void relu_bwd_kernel(float* __restrict__ dx, const float* __restrict__ y, long n) {
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (long)gridDim.x * blockDim.x)
    if (!(y[x] > 0.f)) dx[x] = 0.f;
}
__global__ void tanh_bwd_kernel(float* __restrict__ dy, const float* __restrict__ y, long n) {
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < n; x += (long)gridDim.x * blockDim.x) {
    const float v = y[x];
    dy[x] *= (1.f - v * v);
  }
}

// ------------------------------------------------------------------------ LSTM cells
__global__ void lstm_fwd_kernel(const float* __restrict__ pre, long ld_pre, const float* __restrict__ c_prev,
                                const float* __restrict__ h_prev, float* __restrict__ gates,
                                float* __restrict__ c_out, float* __restrict__ h_out, long ld_h, int rows, int D,
                                const int64_t* __restrict__ len, int t, float* __restrict__ seq_h,
                                float* __restrict__ seq_m, long seq_ld) {
  pdl_trigger();
  pdl_wait();
  const long total = (long)rows * D;
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int d = (int)(x % D);
    const long i = x / D;
    const float* p = pre + i * ld_pre;
    const bool active = (len == nullptr) || (len[i] > t);
    float gi = 0.f, gf = 0.f, gg = 0.f, go = 0.f, c, h;
    if (active) {
      gi = sigmoidf_(p[d]); gf = sigmoidf_(p[D + d]); gg = tanhf(p[2 * D + d]); go = sigmoidf_(p[3 * D + d]);
      c = gf * c_prev[x] + gi * gg;
      h = go * tanhf(c);
    } else {
      c = c_prev[x];
      h = h_prev[x];
    }
    float* g = gates + i * 4 * D;
    g[d] = gi; g[D + d] = gf; g[2 * D + d] = gg; g[3 * D + d] = go;
    c_out[x] = c;
    h_out[i * ld_h + d] = h;
    if (seq_h) {
      seq_h[i * seq_ld + (long)t * D + d] = active ? h : 0.f;
      seq_m[i * seq_ld + (long)t * D + d] = active ? c : 0.f;
    }
  }
}

__device__ __forceinline__ void lstm_bwd_core(float gi, float gf, float gg, float go, float c_prev, float c_cur,
                                              float dh, float dc_in, float* dg, int D, int d, float& dc_prev) {
  const float tc = tanhf(c_cur);
  const float dc = dc_in + dh * go * (1.f - tc * tc);
  dg[d] = dc * gg * gi * (1.f - gi);
  dg[D + d] = dc * c_prev * gf * (1.f - gf);
  dg[2 * D + d] = dc * gi * (1.f - gg * gg);
  dg[3 * D + d] = dh * tc * go * (1.f - go);
  dc_prev = dc * gf;
}

__global__ void lstm_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                const float* __restrict__ c_cur, const float* __restrict__ dh, long ld_dh,
                                const float* __restrict__ dh_b, float* __restrict__ dc_carry,
                                float* __restrict__ dgates, int rows, int D) {
  pdl_trigger();
  pdl_wait();
  const long total = (long)rows * D;
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int d = (int)(x % D);
    const long i = x / D;
    const float* g = gates + i * 4 * D;
    float dhv = dh[i * ld_dh + d];
    if (dh_b) dhv += dh_b[x];
    float dcp;
    lstm_bwd_core(g[d], g[D + d], g[2 * D + d], g[3 * D + d], c_prev[x], c_cur[x], dhv, dc_carry[x],
                  dgates + i * 4 * D, D, d, dcp);
    dc_carry[x] = dcp;
  }
}

__global__ void enc_lstm_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                    const float* __restrict__ c_cur, float* __restrict__ dh_run,
                                    float* __restrict__ dc_run, const float* __restrict__ dseq_h,
                                    const float* __restrict__ dseq_m, long seq_ld,
                                    const float* __restrict__ dh_last, const int64_t* __restrict__ len, int t,
                                    float* __restrict__ dgates, int rows, int D) {
  pdl_trigger();
  pdl_wait();
  const long total = (long)rows * D;
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int d = (int)(x % D);
    const long i = x / D;
    float* dg = dgates + i * 4 * D;
    const long L = len[i];
    if (L <= t) {
      dg[d] = 0.f; dg[D + d] = 0.f; dg[2 * D + d] = 0.f; dg[3 * D + d] = 0.f;
      continue;
    }
    const float* g = gates + i * 4 * D;
    float dhv = dh_run[x] + dseq_h[i * seq_ld + (long)t * D + d];
    if (L - 1 == t) dhv += dh_last[x];
    const float dcv = dc_run[x] + dseq_m[i * seq_ld + (long)t * D + d];
    float dcp;
    lstm_bwd_core(g[d], g[D + d], g[2 * D + d], g[3 * D + d], c_prev[x], c_cur[x], dhv, dcv, dg, D, d, dcp);
    dc_run[x] = dcp;
  }
}

__global__ void bilstm_fwd_kernel(const float* __restrict__ hh_pre, const float* __restrict__ xg,
                                  const int64_t* __restrict__ len, int s, int reverse,
                                  const float* __restrict__ h_prev, const float* __restrict__ c_prev,
                                  float* __restrict__ h_out, float* __restrict__ c_out, float* __restrict__ gates,
                                  float* __restrict__ out, long out_ld_row, long out_ld_pos, int B, int P, int C) {
  const long total = (long)B * C;
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int d = (int)(x % C);
    const long i = x / C;
    const long L = len[i];
    float* g = gates + i * 4 * C;
    if (L <= s) {
      g[d] = 0.f; g[C + d] = 0.f; g[2 * C + d] = 0.f; g[3 * C + d] = 0.f;
      h_out[x] = h_prev[x];
      c_out[x] = c_prev[x];
      continue;
    }
    const long pos = reverse ? L - 1 - s : s;
    const float* px = xg + (i * P + pos) * 4 * C;
    float pi = px[d], pf = px[C + d], pg = px[2 * C + d], po = px[3 * C + d];
    if (hh_pre) {
      const float* ph = hh_pre + i * 4 * C;
      pi += ph[d]; pf += ph[C + d]; pg += ph[2 * C + d]; po += ph[3 * C + d];
    }
    const float gi = sigmoidf_(pi), gf = sigmoidf_(pf), gg = tanhf(pg), go = sigmoidf_(po);
    const float c = gf * c_prev[x] + gi * gg;
    const float h = go * tanhf(c);
    g[d] = gi; g[C + d] = gf; g[2 * C + d] = gg; g[3 * C + d] = go;
    c_out[x] = c;
    h_out[x] = h;
    out[i * out_ld_row + pos * out_ld_pos + d] = h;
  }
}

__global__ void bilstm_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                  const float* __restrict__ c_cur, float* __restrict__ dh_run,
                                  float* __restrict__ dc_run, const float* __restrict__ dout, long out_ld_row,
                                  long out_ld_pos, const float* __restrict__ dh_last, long ld_dh_last,
                                  const int64_t* __restrict__ len, int s, int reverse, float* __restrict__ dgates,
                                  float* __restrict__ dxg, int B, int P, int C) {
  const long total = (long)B * C;
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int d = (int)(x % C);
    const long i = x / C;
    const long L = len[i];
    float* dg = dgates + i * 4 * C;
    if (L <= s) {
      dg[d] = 0.f; dg[C + d] = 0.f; dg[2 * C + d] = 0.f; dg[3 * C + d] = 0.f;
      continue;
    }
    const long pos = reverse ? L - 1 - s : s;
    const float* g = gates + i * 4 * C;
    float dhv = dh_run[x] + dout[i * out_ld_row + pos * out_ld_pos + d];
    if (L - 1 == s) dhv += dh_last[i * ld_dh_last + d];     // this direction's final state feeds `concat`
    float dcp;
    lstm_bwd_core(g[d], g[C + d], g[2 * C + d], g[3 * C + d], c_prev[x], c_cur[x], dhv, dc_run[x], dg, C, d, dcp);
    dc_run[x] = dcp;
    float* dx = dxg + (i * P + pos) * 4 * C;
    dx[d] = dg[d]; dx[C + d] = dg[C + d]; dx[2 * C + d] = dg[2 * C + d]; dx[3 * C + d] = dg[3 * C + d];
  }
}

__global__ void enc_mask_kernel(const float* __restrict__ prev_m, float* __restrict__ mask, long rows, int D) {
  // warp per (i,p) row
  const int lane = threadIdx.x & 31;
  const long w = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= rows) return;
  float s = 0.f;
  for (int d = lane; d < D; d += 32) s += prev_m[w * D + d];
  s = warp_sum(s);
  if (lane == 0) mask[w] = (s != 0.f) ? 1.f : 0.f;
}

// ------------------------------------------------------------------------ attention
// Every loop below is latency-bound (a sample's rows are HBM- or L2-cold and a CTA owns few of them), so the
// kernels are organised around memory-level parallelism: (1) several CTAs per sample -- kCapSlices column /
// unit slices for the caption attention, kVisSlices for the visual one, each recomputing the (cheap) scores;
// (2) every thread requests a whole batch of independent 128-bit loads before it consumes the first one.
constexpr int kRowBatch = 9;

// Pull [rows x bytes_per_row] (row stride in bytes) towards L2, one 128-byte line per thread and iteration.  The
// attention kernels call it BEFORE pdl_wait() on operands no in-flight kernel writes (region features, the hoisted
// score projections, encoder outputs): their HBM latency is then paid while the previous GEMM is still running.
__device__ __forceinline__ void l2_prefetch_rows(const void* base, long row_stride_bytes, int rows, int bytes_per_row) {
  const int lines = (bytes_per_row + 127) >> 7;
  const char* b = reinterpret_cast<const char*>(base);
  for (int e = threadIdx.x; e < rows * lines; e += blockDim.x) {
    const int r = e / lines, l = e % lines;
    asm volatile("prefetch.global.L2 [%0];" ::"l"(b + (long)r * row_stride_bytes + (long)l * 128));
  }
}

// scores s_j = w . act(att1_j + a2) + bias for the n rows of one sample, into sc[] (shared).  A warp takes up to 3
// rows at a time and requests every piece of them before using any.
template <bool kTanh>
__device__ __forceinline__ void attn_scores(const float* __restrict__ att1, int n, int A, const float* a2, const float* wv,
                                            float bias, float* sc) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const int A4 = A >> 2;
  const float4* a2v = reinterpret_cast<const float4*>(a2);
  const float4* wvv = reinterpret_cast<const float4*>(wv);
  for (int jb = wid; jb < n; jb += 3 * nw) {
    float acc[3] = {0.f, 0.f, 0.f};
    for (int xc = 0; xc < A4; xc += 128) {
      float4 v[3][4];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const int j = jb + r * nw;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int x4 = xc + lane + 32 * k;
          v[r][k] = (j < n && x4 < A4) ? __ldg(reinterpret_cast<const float4*>(att1 + (long)j * A) + x4)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int x4 = xc + lane + 32 * k;
          if (x4 >= A4) continue;
          const float4 b = a2v[x4], ww = wvv[x4], y = v[r][k];
          if (kTanh) {
            acc[r] += ww.x * tanhf(y.x + b.x) + ww.y * tanhf(y.y + b.y) + ww.z * tanhf(y.z + b.z) + ww.w * tanhf(y.w + b.w);
          } else {
            acc[r] += ww.x * fmaxf(y.x + b.x, 0.f) + ww.y * fmaxf(y.y + b.y, 0.f) + ww.z * fmaxf(y.z + b.z, 0.f) +
                      ww.w * fmaxf(y.w + b.w, 0.f);
          }
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      const int j = jb + r * nw;
      if (j >= n) continue;          // warp-uniform
      const float sv = warp_sum(acc[r]) + bias;
      if (lane == 0) sc[j] = sv;
    }
  }
}

// out[c] = sum_{r < nrows} alpha[r] * rows[r][c] for the float4 columns [c0, c1).  Threads = (column, row group);
// each thread batches kRowBatch row loads; the groups' partial sums meet in `part` (blockDim float4 of shared memory).
__device__ __forceinline__ void attn_weighted_rows(const float* __restrict__ rows, long row_stride, int nrows, int c0, int c1,
                                                   const float* alpha, float* __restrict__ out, float4* part) {
  const int tid = threadIdx.x;
  const int cw = min(c1 - c0, (int)blockDim.x);
  if (cw <= 0) return;
  const int G = blockDim.x / cw;
  const int g = tid / cw, c = tid % cw;
  for (int cb = c0; cb < c1; cb += cw) {
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const int col = cb + c;
    if (g < G && col < c1) {
      for (int r0 = g; r0 < nrows; r0 += G * kRowBatch) {
        float4 v[kRowBatch];
#pragma unroll
        for (int k = 0; k < kRowBatch; ++k) {
          const int r = r0 + k * G;
          v[k] = (r < nrows) ? __ldg(reinterpret_cast<const float4*>(rows + (long)r * row_stride) + col)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
        }
#pragma unroll
        for (int k = 0; k < kRowBatch; ++k) {
          const int r = r0 + k * G;
          const float al = (r < nrows) ? alpha[r] : 0.f;
          acc.x += al * v[k].x; acc.y += al * v[k].y; acc.z += al * v[k].z; acc.w += al * v[k].w;
        }
      }
    }
    part[tid] = acc;
    __syncthreads();
    if (g == 0 && col < c1) {
      for (int k = 1; k < G; ++k) {
        const float4 v = part[k * cw + c];
        acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
      }
      reinterpret_cast<float4*>(out)[col] = acc;
    }
    __syncthreads();
  }
}

// grid (b, cap_slices + vis_slices).  y < cap_slices: caption attention (editnet.py:370-376) + select (:409-421) for
// a column slice of D;  y >= cap_slices: visual attention (:442-446; adaptive :449-456) for a column slice of F.
// dynamic smem: a2[A] | wv[A] | sc[max(P,R) padded] | part[4 * blockDim]
__global__ void __launch_bounds__(kAttnThreads, 3) attention_fwd_kernel(const AttnFwdArgs a, int cap_slices) {
  pdl_trigger();
  extern __shared__ float sm[];
  const int A = a.A;
  const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const bool cap = ((int)blockIdx.y < cap_slices);
  const int n = cap ? a.P : a.R;
  {
    // constant operands of this CTA: its sample's hoisted score projection and its column slice of the values
    const int slices = cap ? cap_slices : (int)gridDim.y - cap_slices, sl = cap ? (int)blockIdx.y : (int)blockIdx.y - cap_slices;
    const int width = cap ? a.D : a.F;                       // value row length (floats)
    const int per4 = ((width >> 2) + slices - 1) / slices;   // float4 columns per slice
    const float* vals = cap ? a.prev_h + (long)i * a.P * a.D : a.feats + (long)i * a.R * a.F;
    l2_prefetch_rows(cap ? a.att1c + (long)i * a.P * A : a.att1v + (long)i * a.R * A, (long)A * 4, n, A * 4);
    l2_prefetch_rows(vals + (long)sl * per4 * 4, (long)width * 4, n, per4 * 16);
  }
  pdl_wait();
  float* a2 = sm;
  float* wv = sm + A;
  float* sc = sm + 2 * A;
  float4* part = reinterpret_cast<float4*>(sm + 2 * A + ((max(a.P, a.R) + 3) & ~3));
  const float* att1 = cap ? a.att1c + (long)i * a.P * A : a.att1v + (long)i * a.R * A;
  const float* s2row = a.s2 + (long)i * a.ld_s2 + (cap ? 0 : A);
  const float* w = cap ? a.cap_w : a.vis_w;
  const float bias = cap ? a.cap_b[0] : a.vis_b[0];
  for (int x = tid; x < A; x += blockDim.x) { a2[x] = s2row[x]; wv[x] = w[x]; }
  __syncthreads();
  const int nvalid = cap ? n : (a.nreg ? a.nreg[i] : n);
  if (cap) attn_scores<true>(att1, n, A, a2, wv, bias, sc);
  else attn_scores<false>(att1, n, A, a2, wv, bias, sc);
  __syncthreads();
  // mask, softmax by warp 0
  if (wid == 0) {
    float m = -INFINITY;
    for (int j = lane; j < n; j += 32) {
      float v = sc[j];
      if (cap) { if (a.mask[(long)i * a.P + j] == 0.f) v = kNegFill; }
      else if (j >= nvalid) v = kNegFill;
      sc[j] = v;
      m = fmaxf(m, v);
    }
    m = warp_max(m);
    float sum = 0.f;
    for (int j = lane; j < n; j += 32) { const float e = expf(sc[j] - m); sc[j] = e; sum += e; }
    sum = warp_sum(sum);
    for (int j = lane; j < n; j += 32) sc[j] = sc[j] / sum;
  }
  __syncthreads();
  if (cap) {
    const int sl = blockIdx.y;
    if (sl == 0)
      for (int j = tid; j < n; j += blockDim.x) a.alpha_c[(long)i * a.P + j] = sc[j];
    const int D4 = a.D >> 2;
    const int per = (D4 + cap_slices - 1) / cap_slices;
    const int c0 = sl * per, c1 = min(D4, c0 + per);
    attn_weighted_rows(a.prev_h + (long)i * a.P * a.D, a.D, n, c0, c1, sc, a.ctx + (long)i * (a.ld_ctx ? a.ld_ctx : a.D), part);
    if (a.prev_m) {
      int js = 0; float best = sc[0];
      for (int j = 1; j < n; ++j) if (sc[j] > best) { best = sc[j]; js = j; }
      const float wsel = best + (1.f - best);
      const float4* pm = reinterpret_cast<const float4*>(a.prev_m + ((long)i * a.P + js) * a.D);
      float4* so = reinterpret_cast<float4*>(a.sel + (long)i * a.D);
      for (int c = c0 + tid; c < c1; c += blockDim.x) {
        const float4 v = __ldg(pm + c);
        so[c] = make_float4(wsel * v.x, wsel * v.y, wsel * v.z, wsel * v.w);
      }
      if (sl == 0 && tid == 0) a.sel_idx[i] = js;
    }
  } else {
    const int slices = gridDim.y - cap_slices, sl = blockIdx.y - cap_slices;
    if (sl == 0)
      for (int j = tid; j < n; j += blockDim.x) a.alpha_v[(long)i * a.R + j] = sc[j];
    const int F4 = a.F >> 2;
    const int per = (F4 + slices - 1) / slices;
    const int c0 = sl * per, c1 = min(F4, c0 + per);
    attn_weighted_rows(a.feats + (long)i * a.R * a.F, a.F, nvalid, c0, c1, sc, a.att_img + (long)i * a.ld_img, part);
  }
}

// Attention backward runs as two chained kernels so that no CTA has to pull a whole sample through one SM:
//   (1) attention_bwd_dal_kernel: d alpha_j = <d context, value_j> (+ select term); the rows of a sample are dealt
//       round-robin to the sample's CTAs; results go to a small global scratch;
//   (2) attention_bwd_main_kernel: softmax backward (recomputed by every CTA of the sample: n <= 100 values), then a
//       unit slice of the score MLP; the caption CTAs also share the value gradients (d prev_h, d prev_m).
__global__ void __launch_bounds__(kAttnThreads, 3) attention_bwd_dal_kernel(const AttnBwdArgs a, float* __restrict__ dal_out,
                                                                            int cap_slices) {
  pdl_trigger();
  const bool cap = ((int)blockIdx.y < cap_slices);
  const int i = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, nw = blockDim.x >> 5;
  float* dal = dal_out + (long)i * (a.P + a.R);
  if (cap) {
    // this CTA's value rows (constant): rows blockIdx.y, + cap_slices, ...
    const int rows = (a.P - (int)blockIdx.y + cap_slices - 1) / cap_slices;
    l2_prefetch_rows(a.prev_h + ((long)i * a.P + blockIdx.y) * a.D, (long)cap_slices * a.D * 4, rows, a.D * 4);
  } else {
    const int slices = gridDim.y - cap_slices, sl = blockIdx.y - cap_slices;
    const int rows = (a.R - sl + slices - 1) / slices;
    l2_prefetch_rows(a.feats + ((long)i * a.R + sl) * a.F, (long)slices * a.F * 4, rows, a.F * 4);
  }
  pdl_wait();
  // <x, y_r> for this CTA's rows: a warp per row, 8 x 128-bit loads per lane in flight
  auto dots = [&](const float* __restrict__ x, const float* __restrict__ ybase, long ystride, int len4, int r_first, int r_step,
                  int nrows, int nvalid, float* out, int js, const float* __restrict__ xs, const float* __restrict__ ys) {
    const float4* x4 = reinterpret_cast<const float4*>(x);
    for (int r = r_first + r_step * wid; r < nrows; r += r_step * nw) {
      float s = 0.f;
      if (r < nvalid) {
        const float4* y4 = reinterpret_cast<const float4*>(ybase + (long)r * ystride);
        for (int d0 = lane; d0 < len4; d0 += 32 * 8) {
          float4 yv[8], xv[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const int d = d0 + 32 * k;
            yv[k] = d < len4 ? __ldg(y4 + d) : make_float4(0.f, 0.f, 0.f, 0.f);
            xv[k] = d < len4 ? x4[d] : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) s += xv[k].x * yv[k].x + xv[k].y * yv[k].y + xv[k].z * yv[k].z + xv[k].w * yv[k].w;
        }
        if (r == js) {   // straight-through select term: <d sel, prev_m[js]>
          const float4* y4s = reinterpret_cast<const float4*>(ys);
          const float4* x4s = reinterpret_cast<const float4*>(xs);
          for (int d = lane; d < len4; d += 32) {
            const float4 xx = x4s[d], yy = __ldg(y4s + d);
            s += xx.x * yy.x + xx.y * yy.y + xx.z * yy.z + xx.w * yy.w;
          }
        }
      }
      s = warp_sum(s);
      if (lane == 0) out[r] = s;
    }
  };
  if (cap) {
    const int js = a.dsel ? a.sel_idx[i] : -1;
    dots(a.dctx + (long)i * (a.ld_dctx ? a.ld_dctx : a.D), a.prev_h + (long)i * a.P * a.D, a.D, a.D >> 2, blockIdx.y, cap_slices,
         a.P, a.P, dal, js, a.dsel ? a.dsel + (long)i * a.D : nullptr,
         (a.dsel && js >= 0) ? a.prev_m + ((long)i * a.P + js) * a.D : nullptr);
  } else {
    const int slices = gridDim.y - cap_slices, sl = blockIdx.y - cap_slices;
    const int nvalid = a.nreg ? a.nreg[i] : a.R;
    dots(a.datt_img + (long)i * a.ld_dimg, a.feats + (long)i * a.R * a.F, a.F, a.F >> 2, sl, slices, a.R, nvalid, dal + a.P, -1,
         nullptr, nullptr);
  }
}

// dynamic smem: a2[A] | wv[A] | al[n] | ds[n] | red[40] | part[2 * blockDim]
__global__ void __launch_bounds__(kAttnThreads, 3) attention_bwd_main_kernel(const AttnBwdArgs a, const float* __restrict__ dal_in,
                                                                             int cap_slices) {
  pdl_trigger();
  extern __shared__ float sm[];
  const int A = a.A;
  const int y = blockIdx.y;
  const bool cap = (y < cap_slices);
  {
    // the unit slice of the hoisted score projection this CTA differentiates (constant)
    const int slices_ = cap ? cap_slices : (int)gridDim.y - cap_slices, sl_ = cap ? y : y - cap_slices;
    const int per_ = (A + slices_ - 1) / slices_;
    const float* att1_ = cap ? a.att1c + (long)blockIdx.x * a.P * A : a.att1v + (long)blockIdx.x * a.R * A;
    l2_prefetch_rows(att1_ + (long)sl_ * per_, (long)A * 4, cap ? a.P : a.R, per_ * 4);
  }
  pdl_wait();
  const int n = cap ? a.P : a.R;
  const int npad = (max(a.P, a.R) + 3) & ~3;
  float* a2 = sm;
  float* wv = sm + A;
  float* al = sm + 2 * A;
  float* ds = al + npad;
  float* red = ds + npad;
  float* part = red + 40;
  const int i = blockIdx.x, tid = threadIdx.x;
  const float* s2row = a.s2 + (long)i * a.ld_s2 + (cap ? 0 : A);
  const float* w = cap ? a.cap_w : a.vis_w;
  const float* alpha = cap ? a.alpha_c + (long)i * a.P : a.alpha_v + (long)i * a.R;
  const float* dal = dal_in + (long)i * (a.P + a.R) + (cap ? 0 : a.P);
  for (int x = tid; x < A; x += blockDim.x) { a2[x] = s2row[x]; wv[x] = w[x]; }
  for (int j = tid; j < n; j += blockDim.x) al[j] = alpha[j];
  __syncthreads();
  const int nvalid = cap ? n : (a.nreg ? a.nreg[i] : n);
  // softmax backward
  float p = 0.f;
  for (int j = tid; j < n; j += blockDim.x) p += al[j] * dal[j];
  const float dot = block_sum(p, red);
  float dsum = 0.f;
  for (int j = tid; j < n; j += blockDim.x) {
    float v = al[j] * (dal[j] - dot);
    if (cap) { if (a.mask[(long)i * a.P + j] == 0.f) v = 0.f; }
    else if (j >= nvalid) v = 0.f;
    ds[j] = v;
    dsum += v;
  }
  const float dbias = block_sum(dsum, red);  // also orders ds[] writes before the reads below
  const int slices = cap ? cap_slices : (int)gridDim.y - cap_slices;
  const int sl = cap ? y : y - cap_slices;
  if (sl == 0 && tid == 0) atomicAdd(cap ? a.dcap_b : a.dvis_b, dbias);
  if (cap) {
    // value gradients, rows dealt round-robin to the caption CTAs: d prev_h[j] += alpha_j * d ctx
    const float* __restrict__ dc = a.dctx + (long)i * (a.ld_dctx ? a.ld_dctx : a.D);
    float* __restrict__ dph = a.dprev_h + (long)i * a.P * a.D;
    const int D4 = a.D >> 2;
    const int nrow = (n - sl + slices - 1) / slices;          // rows sl, sl + slices, ...
    const int items = nrow * D4;
    for (int e0 = tid; e0 < items; e0 += blockDim.x * kRowBatch) {
      float4 old[kRowBatch], g[kRowBatch];
#pragma unroll
      for (int k = 0; k < kRowBatch; ++k) {
        const int e = e0 + k * blockDim.x;
        if (e < items) {
          const int j = sl + slices * (e / D4), d4 = e % D4;
          old[k] = *(reinterpret_cast<const float4*>(dph + (long)j * a.D) + d4);
          g[k] = reinterpret_cast<const float4*>(dc)[d4];
        }
      }
#pragma unroll
      for (int k = 0; k < kRowBatch; ++k) {
        const int e = e0 + k * blockDim.x;
        if (e < items) {
          const int j = sl + slices * (e / D4), d4 = e % D4;
          const float aj = al[j];
          float4 v = old[k];
          v.x += aj * g[k].x; v.y += aj * g[k].y; v.z += aj * g[k].z; v.w += aj * g[k].w;
          *(reinterpret_cast<float4*>(dph + (long)j * a.D) + d4) = v;
        }
      }
    }
    const int js = a.dsel ? a.sel_idx[i] : -1;
    if (js >= 0 && sl == 0) {
      const float best = al[js];
      const float wsel = best + (1.f - best);
      float* dpm = a.dprev_m + ((long)i * a.P + js) * a.D;
      const float* dsl = a.dsel + (long)i * a.D;
      for (int d = tid; d < a.D; d += blockDim.x) dpm[d] += wsel * dsl[d];
    }
  }
  // score-MLP backward for a slice of the attention units: threads = (unit, row group), rows in batches
  const int per = (A + slices - 1) / slices;
  const int x0 = sl * per, x1 = min(A, x0 + per);
  const float* __restrict__ att1 = cap ? a.att1c + (long)i * a.P * A : a.att1v + (long)i * a.R * A;
  float* __restrict__ datt1 = cap ? a.datt1c + (long)i * a.P * A : a.datt1v + (long)i * a.R * A;
  const bool accum = cap ? true : (a.datt1v_accum != 0);
  float* ds2row = a.ds2 + (long)i * a.ld_ds2 + (cap ? 0 : A);
  float* dwv = cap ? a.dcap_w : a.dvis_w;
  const int cw = min(per, (int)blockDim.x);
  const int G = blockDim.x / cw;
  const int g = tid / cw, c = tid % cw;
  for (int xb = x0; xb < x1; xb += cw) {
    const int x = xb + c;
    float d2 = 0.f, dw = 0.f;
    if (g < G && x < x1) {
      const float av = a2[x], wx = wv[x];
      for (int j0 = g; j0 < n; j0 += G * kRowBatch) {
        float pre[kRowBatch], old[kRowBatch];
#pragma unroll
        for (int k = 0; k < kRowBatch; ++k) {
          const int j = j0 + k * G;
          pre[k] = (j < n) ? att1[(long)j * A + x] : 0.f;
          old[k] = (accum && j < n) ? datt1[(long)j * A + x] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < kRowBatch; ++k) {
          const int j = j0 + k * G;
          if (j >= n) continue;
          const float pv = pre[k] + av;
          float yv, dpre;
          if (cap) { yv = tanhf(pv); dpre = ds[j] * wx * (1.f - yv * yv); }
          else { yv = fmaxf(pv, 0.f); dpre = (pv > 0.f) ? ds[j] * wx : 0.f; }
          dw += ds[j] * yv;
          d2 += dpre;
          datt1[(long)j * A + x] = old[k] + dpre;
        }
      }
    }
    part[2 * tid] = d2; part[2 * tid + 1] = dw;
    __syncthreads();
    if (g == 0 && x < x1) {
      for (int k = 1; k < G; ++k) { d2 += part[2 * (k * cw + c)]; dw += part[2 * (k * cw + c) + 1]; }
      ds2row[x] = d2;
      atomicAdd(dwv + x, dw);
    }
    __syncthreads();
  }
}

// --------------------------------------------------------------------- context gate
__global__ void ctx_gate_fwd_kernel(const float* __restrict__ s4, long ld_s4, const float* __restrict__ th,
                                    long ld_th, float* __restrict__ zst, float* __restrict__ att_cap,
                                    long ld_cap, int rows, int D) {
  pdl_trigger();
  pdl_wait();
  const long total = (long)rows * D;
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int d = (int)(x % D);
    const long i = x / D;
    const float z = sigmoidf_(s4[i * ld_s4 + d]);
    const float tsc = tanhf(s4[i * ld_s4 + D + d]);
    const float ttc = tanhf(th[i * ld_th + d]);
    float* o = zst + i * 3 * D;
    o[d] = z; o[D + d] = tsc; o[2 * D + d] = ttc;
    att_cap[i * ld_cap + d] = z * tsc + (1.f - z) * ttc;
  }
}
__global__ void ctx_gate_bwd_kernel(const float* __restrict__ zst, const float* __restrict__ datt_cap,
                                    long ld_dcap, float* __restrict__ dz_out, float* __restrict__ dtc_out,
                                    long ld_ds2, float* __restrict__ dsc, int rows, int D) {
  pdl_trigger();
  pdl_wait();
  const long total = (long)rows * D;
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int d = (int)(x % D);
    const long i = x / D;
    const float* o = zst + i * 3 * D;
    const float z = o[d], tsc = o[D + d], ttc = o[2 * D + d];
    const float g = datt_cap[i * ld_dcap + d];
    dz_out[i * ld_ds2 + d] = g * (tsc - ttc) * z * (1.f - z);
    dsc[x] = g * z * (1.f - tsc * tsc);
    dtc_out[i * ld_ds2 + d] = g * (1.f - z) * (1.f - ttc * ttc);
  }
}

// ------------------------------------------------------------------------ copy-LSTM
__global__ void copy1_fwd_kernel(float* __restrict__ g2, const float* __restrict__ c2_prev,
                                 float* __restrict__ cnew, int rows, int D) {
  pdl_trigger();
  pdl_wait();
  const long total = (long)rows * D;
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int d = (int)(x % D);
    const long i = x / D;
    float* g = g2 + i * 4 * D;
    const float gi = sigmoidf_(g[d]), gf = sigmoidf_(g[D + d]), gg = tanhf(g[2 * D + d]), go = sigmoidf_(g[3 * D + d]);
    g[d] = gi; g[D + d] = gf; g[2 * D + d] = gg; g[3 * D + d] = go;
    cnew[x] = gf * c2_prev[x] + gi * gg;
  }
}
__global__ void copy2_fwd_kernel(const float* __restrict__ kpre, long ld_k, const float* __restrict__ g2,
                                 const float* __restrict__ sel, const float* __restrict__ cnew,
                                 float* __restrict__ kgate, float* __restrict__ c2, float* __restrict__ h2,
                                 float* __restrict__ h2drop, int rows, int D, int train, uint64_t seed,
                                 long drop_base) {
  pdl_trigger();
  pdl_wait();
  const long total = (long)rows * D;
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int d = (int)(x % D);
    const long i = x / D;
    const float k = sigmoidf_(kpre[i * ld_k + d]);
    const float c = k * sel[x] + (1.f - k) * cnew[x];
    const float h = g2[i * 4 * D + 3 * D + d] * tanhf(c);
    kgate[x] = k; c2[x] = c; h2[x] = h;
    if (h2drop) {
      float hd = h;
      if (train) hd = drop_keep(seed, kSiteFc, (uint64_t)(drop_base + x)) ? h * 2.f : 0.f;
      h2drop[x] = hd;
    }
  }
}
__global__ void copy2_bwd_kernel(const float* __restrict__ dh2_carry, const float* __restrict__ dh2drop_raw,
                                 const float* __restrict__ dc2_carry, const float* __restrict__ g2,
                                 const float* __restrict__ c2, const float* __restrict__ kgate,
                                 const float* __restrict__ sel, const float* __restrict__ cnew,
                                 float* __restrict__ dg2, float* __restrict__ dkpre, float* __restrict__ dsel,
                                 float* __restrict__ dcnew, int rows, int D, int train, uint64_t seed,
                                 long drop_base) {
  pdl_trigger();
  pdl_wait();
  const long total = (long)rows * D;
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int d = (int)(x % D);
    const long i = x / D;
    float dfc = dh2drop_raw ? dh2drop_raw[x] : 0.f;
    if (train && dh2drop_raw) dfc = drop_keep(seed, kSiteFc, (uint64_t)(drop_base + x)) ? dfc * 2.f : 0.f;
    const float dh = dh2_carry[x] + dfc;
    const float go = g2[i * 4 * D + 3 * D + d];
    const float tc = tanhf(c2[x]);
    dg2[i * 4 * D + 3 * D + d] = dh * tc * go * (1.f - go);
    const float dc = dc2_carry[x] + dh * go * (1.f - tc * tc);
    const float k = kgate[x];
    dkpre[x] = dc * (sel[x] - cnew[x]) * k * (1.f - k);
    dsel[x] = dc * k;
    dcnew[x] = dc * (1.f - k);
  }
}
__global__ void copy1_bwd_kernel(const float* __restrict__ dcnew, const float* __restrict__ g2,
                                 const float* __restrict__ c2_prev, float* __restrict__ dg2,
                                 float* __restrict__ dc2_carry, int rows, int D) {
  pdl_trigger();
  pdl_wait();
  const long total = (long)rows * D;
  for (long x = (long)blockIdx.x * blockDim.x + threadIdx.x; x < total; x += (long)gridDim.x * blockDim.x) {
    const int d = (int)(x % D);
    const long i = x / D;
    const float* g = g2 + i * 4 * D;
    const float gi = g[d], gf = g[D + d], gg = g[2 * D + d];
    const float dc = dcnew[x];
    float* dg = dg2 + i * 4 * D;
    dg[d] = dc * gg * gi * (1.f - gi);
    dg[D + d] = dc * c2_prev[x] * gf * (1.f - gf);
    dg[2 * D + d] = dc * gi * (1.f - gg * gg);
    dc2_carry[x] = dc * gf;
  }
}

__global__ void dropout_fwd_kernel(const float* __restrict__ x, float* __restrict__ out, long n, int train,
                                   uint64_t seed, uint32_t site, long base) {
  for (long y = (long)blockIdx.x * blockDim.x + threadIdx.x; y < n; y += (long)gridDim.x * blockDim.x) {
    float v = x[y];
    if (train) v = drop_keep(seed, site, (uint64_t)(base + y)) ? v * 2.f : 0.f;
    out[y] = v;
  }
}
__global__ void transpose_kernel(const float* __restrict__ in, long ld_in, float* __restrict__ out, long ld_out,
                                 int rows, int cols) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int r = r0 + j, c = c0 + threadIdx.x;
    tile[j][threadIdx.x] = (r < rows && c < cols) ? in[(long)r * ld_in + c] : 0.f;
  }
  __syncthreads();
  for (int j = threadIdx.y; j < 32; j += 8) {
    const int c = c0 + j, r = r0 + threadIdx.x;
    if (c < cols && r < rows) out[(long)c * ld_out + r] = tile[threadIdx.x][j];
  }
}
// up to kTrMaxJobs transposes in one grid: 64x64 tiles through shared memory, 128-bit accesses on both sides
__global__ void __launch_bounds__(256) transpose_batch_kernel(const __grid_constant__ TrBatch b) {
  __shared__ float tile[64][65];
  int ji = 0;
  while (ji + 1 < b.n && (int)blockIdx.x >= b.j[ji + 1].tile0) ++ji;
  const TrJob& J = b.j[ji];
  const int t = blockIdx.x - J.tile0;
  const int r0 = (t / J.tiles_c) * 64, c0 = (t % J.tiles_c) * 64;
  const int tid = threadIdx.x, q = tid >> 4, l4 = (tid & 15) * 4;
  const bool vin = (J.ld_in % 4 == 0) && ((reinterpret_cast<uintptr_t>(J.in) & 15) == 0);
  const bool vout = (J.ld_out % 4 == 0) && ((reinterpret_cast<uintptr_t>(J.out) & 15) == 0);
  float4 v[4];
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int r = r0 + ps * 16 + q, c = c0 + l4;
    v[ps] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < J.rows) {
      const float* src = J.in + (long)r * J.ld_in + c;
      if (vin && c + 3 < J.cols) v[ps] = __ldg(reinterpret_cast<const float4*>(src));
      else {
        if (c < J.cols) v[ps].x = src[0];
        if (c + 1 < J.cols) v[ps].y = src[1];
        if (c + 2 < J.cols) v[ps].z = src[2];
        if (c + 3 < J.cols) v[ps].w = src[3];
      }
    }
  }
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    float* trow = tile[ps * 16 + q];
    trow[l4] = v[ps].x; trow[l4 + 1] = v[ps].y; trow[l4 + 2] = v[ps].z; trow[l4 + 3] = v[ps].w;
  }
  __syncthreads();
#pragma unroll
  for (int ps = 0; ps < 4; ++ps) {
    const int c = c0 + ps * 16 + q, r = r0 + l4;
    if (c >= J.cols) continue;
    const int cc = ps * 16 + q;
    const float4 o = make_float4(tile[l4][cc], tile[l4 + 1][cc], tile[l4 + 2][cc], tile[l4 + 3][cc]);
    float* dst = J.out + (long)c * J.ld_out + r;
    if (vout && r + 3 < J.rows) *reinterpret_cast<float4*>(dst) = o;
    else {
      if (r < J.rows) dst[0] = o.x;
      if (r + 1 < J.rows) dst[1] = o.y;
      if (r + 2 < J.rows) dst[2] = o.z;
      if (r + 3 < J.rows) dst[3] = o.w;
    }
  }
}
__global__ void sum_time_kernel(const float* __restrict__ x, float* __restrict__ out, int T, long n) {
  for (long y = (long)blockIdx.x * blockDim.x + threadIdx.x; y < n; y += (long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int t = 0; t < T; ++t) s += x[(long)t * n + y];
    out[y] = s;
  }
}
__global__ void zero_batch_kernel(const ZeroBatch zb) {
  const ZeroJob j = zb.j[blockIdx.y];
  for (long y = (long)blockIdx.x * blockDim.x + threadIdx.x; y < j.n; y += (long)gridDim.x * blockDim.x) j.p[y] = 0.f;
}
__global__ void keep_mask_kernel(float* __restrict__ out, long n, uint64_t seed, uint32_t site, long base) {
  for (long y = (long)blockIdx.x * blockDim.x + threadIdx.x; y < n; y += (long)gridDim.x * blockDim.x)
    out[y] = drop_keep(seed, site, (uint64_t)(base + y)) ? 1.f : 0.f;
}

// The tensor-core GEMM needs the maximum shared-memory carveout (197 KB of dynamic smem).  If the small
// kernels that run between two GEMM launches ask for the default split, every SM is drained and
// re-partitioned at each boundary (~10 us per launch on B200).  All kernels of the path therefore
// declare the same preference once.
static std::once_flag g_carveout_once;
template <typename K>
static void prefer_smem(K kern) {
  cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}
static void init_carveout() {
  prefer_smem(embed_fwd_kernel); prefer_smem(embed_bwd_kernel); prefer_smem(region_mean_kernel);
  prefer_smem(region_count_kernel); prefer_smem(zero_pad_regions_kernel); prefer_smem(vis_dropout_fwd_kernel);
  prefer_smem(vis_dropout_bwd_kernel); prefer_smem(relu_bwd_kernel); prefer_smem(tanh_bwd_kernel);
  prefer_smem(lstm_fwd_kernel); prefer_smem(lstm_bwd_kernel); prefer_smem(enc_lstm_bwd_kernel);
  prefer_smem(bilstm_fwd_kernel); prefer_smem(bilstm_bwd_kernel); prefer_smem(enc_mask_kernel);
  prefer_smem(attention_fwd_kernel); prefer_smem(attention_bwd_dal_kernel); prefer_smem(attention_bwd_main_kernel); prefer_smem(ctx_gate_fwd_kernel);
  prefer_smem(ctx_gate_bwd_kernel); prefer_smem(copy1_fwd_kernel); prefer_smem(copy2_fwd_kernel);
  prefer_smem(copy2_bwd_kernel); prefer_smem(copy1_bwd_kernel); prefer_smem(dropout_fwd_kernel);
  prefer_smem(transpose_kernel); prefer_smem(sum_time_kernel); prefer_smem(keep_mask_kernel);
  cudaGetLastError();
}

#define LAUNCH_OK()                         \
  std::call_once(g_carveout_once, init_carveout); \
  SET_CHECK_CUDA(cudaGetLastError());       \
  set_count_launch(1);                      \
  return SET_OK

}  // namespace

int embed_fwd(const int64_t* tokens, long tok_ld, long tok_os, const float* table, int V, float* out,
              int n_outer, int n_inner, int D, int train, uint64_t seed, uint32_t site, long drop_row0,
              long drop_os, long drop_is, cudaStream_t s) {
  SET_REQUIRE(D % 4 == 0, "D % 4");
  const long n = (long)n_outer * n_inner * (D / 4);
  if (n == 0) return SET_OK;
  embed_fwd_kernel<<<blocks_for(n, kThreads), kThreads, 0, s>>>(tokens, tok_ld, tok_os, table, V, out, n_outer,
                                                                n_inner, D, train, seed, site, drop_row0, drop_os, drop_is);
  LAUNCH_OK();
}
int embed_bwd(const int64_t* tokens, long tok_ld, long tok_os, const float* out, const float* dout,
              float* table_grad, int V, int n_outer, int n_inner, int D, int train, const int* row_len,
              cudaStream_t s) {
  const long n = (long)n_outer * n_inner * D;
  if (n == 0) return SET_OK;
  embed_bwd_kernel<<<blocks_for(n, kThreads), kThreads, 0, s>>>(tokens, tok_ld, tok_os, out, dout, table_grad,
                                                                V, n_outer, n_inner, D, train ? 2.f : 1.f, row_len);
  LAUNCH_OK();
}
int region_mean(const float* feats, float* out, int B, int R, int F, cudaStream_t s) {
  region_mean_kernel<<<blocks_for((long)B * F, kThreads), kThreads, 0, s>>>(feats, out, B, R, F);
  LAUNCH_OK();
}
int region_count(const float* feats, int* nreg, int B, int R, int F, cudaStream_t s) {
  region_count_kernel<<<B, kThreads, 0, s>>>(feats, nreg, R, F);
  LAUNCH_OK();
}
int zero_pad_regions(float* fe, const int* nreg, int B, int R, int D, cudaStream_t s) {
  zero_pad_regions_kernel<<<blocks_for((long)B * R * D, kThreads), kThreads, 0, s>>>(fe, nreg, B, R, D);
  LAUNCH_OK();
}
int vis_dropout_fwd(const float* fe_pre, float* fe_t, int T, int B, int R, int D, uint64_t seed, cudaStream_t s) {
  SET_REQUIRE(D % 4 == 0, "D % 4");
  const long n = (long)B * R * D;
  vis_dropout_fwd_kernel<<<blocks_for((long)T * n / 4, kThreads), kThreads, 0, s>>>(fe_pre, fe_t, T, n, seed);
  LAUNCH_OK();
}
int vis_dropout_bwd(const float* fe_pre, const float* dfe_t, float* dfe_pre, const int* dec_len, int T, int B,
                    int R, int D, uint64_t seed, cudaStream_t s) {
  const long ps = (long)R * D;
  vis_dropout_bwd_kernel<<<blocks_for((long)B * ps / 4, kThreads), kThreads, 0, s>>>(fe_pre, dfe_t, dfe_pre,
                                                                                    dec_len, T, B, ps, seed);
  LAUNCH_OK();
}
int relu_bwd_inplace(float* dx, const float* y, long n, cudaStream_t s) {
  relu_bwd_kernel<<<blocks_for(n, kThreads), kThreads, 0, s>>>(dx, y, n);
  LAUNCH_OK();
}
int tanh_bwd_inplace(float* dy, const float* y, long n, cudaStream_t s) {
  tanh_bwd_kernel<<<blocks_for(n, kThreads), kThreads, 0, s>>>(dy, y, n);
  LAUNCH_OK();
}
int lstm_fwd(const float* pre, long ld_pre, const float* c_prev, const float* h_prev, float* gates,
             float* c_out, float* h_out, long ld_h, int rows, int D, const int64_t* len, int t, float* seq_h,
             float* seq_m, long seq_ld, cudaStream_t s) {
  if (rows <= 0) return SET_OK;
  SET_CHECK_CUDA(launch_chain(lstm_fwd_kernel, dim3(blocks_for((long)rows * D, kThreads)), dim3(kThreads), 0, s, pre, ld_pre, c_prev, h_prev, gates, c_out, h_out, ld_h, rows, D, len, t, seq_h, seq_m, seq_ld));
  LAUNCH_OK();
}
int lstm_bwd(const float* gates, const float* c_prev, const float* c_cur, const float* dh, long ld_dh,
             const float* dh_b, float* dc_carry, float* dgates, int rows, int D, cudaStream_t s) {
  if (rows <= 0) return SET_OK;
  SET_CHECK_CUDA(launch_chain(lstm_bwd_kernel, dim3(blocks_for((long)rows * D, kThreads)), dim3(kThreads), 0, s, gates, c_prev, c_cur, dh, ld_dh, dh_b, dc_carry, dgates, rows, D));
  LAUNCH_OK();
}
int enc_lstm_bwd(const float* gates, const float* c_prev, const float* c_cur, float* dh_run, float* dc_run,
                 const float* dseq_h, const float* dseq_m, long seq_ld, const float* dh_last,
                 const int64_t* len, int t, float* dgates, int rows, int D, cudaStream_t s) {
  if (rows <= 0) return SET_OK;
  SET_CHECK_CUDA(launch_chain(enc_lstm_bwd_kernel, dim3(blocks_for((long)rows * D, kThreads)), dim3(kThreads), 0, s, gates, c_prev, c_cur, dh_run, dc_run, dseq_h, dseq_m, seq_ld, dh_last, len, t, dgates, rows, D));
  LAUNCH_OK();
}
int bilstm_fwd(const float* hh_pre, const float* xg, const int64_t* len, int s, int reverse, const float* h_prev,
               const float* c_prev, float* h_out, float* c_out, float* gates, float* out, long out_ld_row,
               long out_ld_pos, int B, int P, int C, cudaStream_t st) {
  bilstm_fwd_kernel<<<blocks_for((long)B * C, kThreads), kThreads, 0, st>>>(
      hh_pre, xg, len, s, reverse, h_prev, c_prev, h_out, c_out, gates, out, out_ld_row, out_ld_pos, B, P, C);
  LAUNCH_OK();
}
int bilstm_bwd(const float* gates, const float* c_prev, const float* c_cur, float* dh_run, float* dc_run,
               const float* dout, long out_ld_row, long out_ld_pos, const float* dh_last, long ld_dh_last,
               const int64_t* len, int s, int reverse, float* dgates, float* dxg, int B, int P, int C,
               cudaStream_t st) {
  bilstm_bwd_kernel<<<blocks_for((long)B * C, kThreads), kThreads, 0, st>>>(
      gates, c_prev, c_cur, dh_run, dc_run, dout, out_ld_row, out_ld_pos, dh_last, ld_dh_last, len, s, reverse,
      dgates, dxg, B, P, C);
  LAUNCH_OK();
}
int enc_mask(const float* prev_m, float* mask, int B, int P, int D, cudaStream_t s) {
  const long rows = (long)B * P;
  enc_mask_kernel<<<(int)((rows * 32 + kThreads - 1) / kThreads), kThreads, 0, s>>>(prev_m, mask, rows, D);
  LAUNCH_OK();
}
int attention_fwd(const AttnFwdArgs& a, cudaStream_t s) {
  if (a.b <= 0) return SET_OK;
  SET_REQUIRE(a.F % 4 == 0 && a.D % 4 == 0 && a.A % 4 == 0, "D, A, F % 4");
  const int n = a.P > a.R ? a.P : a.R;
  const size_t smem = sizeof(float) * (2 * a.A + ((n + 3) & ~3) + 4 * kAttnThreads);
  SET_REQUIRE(smem <= 48 * 1024, "attention smem");
  const int cs = a.att1c ? kCapSlices : 0, vs = a.att1v ? kVisSlices : 0;
  if (cs + vs == 0) return SET_OK;
  SET_CHECK_CUDA(launch_chain(attention_fwd_kernel, dim3(a.b, cs + vs), dim3(kAttnThreads), smem, s, a, cs));
  LAUNCH_OK();
}
int attention_bwd(const AttnBwdArgs& a, cudaStream_t s) {
  if (a.b <= 0) return SET_OK;
  const int n = a.P > a.R ? a.P : a.R;
  const size_t smem = sizeof(float) * (2 * a.A + 2 * ((n + 3) & ~3) + 40 + 2 * kAttnThreads);
  SET_REQUIRE(smem <= 48 * 1024, "attention smem");
  // d alpha scratch between the two kernels (library-owned per (device, stream), grown on demand)
  const size_t need = (size_t)a.b * (a.P + a.R);
  float* dal = static_cast<float*>(lib_scratch(kScratchAttnDal, s, sizeof(float) * (need < 65536 ? 65536 : 2 * need), false));
  if (!dal) return SET_ERR_CUDA;
  const int cs = a.att1c ? kCapSlices : 0, vs = a.att1v ? kVisSlices : 0;
  if (cs + vs == 0) return SET_OK;
  SET_CHECK_CUDA(launch_chain(attention_bwd_dal_kernel, dim3(a.b, cs + vs), dim3(kAttnThreads), 0, s, a, dal, cs));
  set_count_launch(1);
  SET_CHECK_CUDA(launch_chain(attention_bwd_main_kernel, dim3(a.b, cs + vs), dim3(kAttnThreads), smem, s, a,
                              (const float*)dal, cs));
  LAUNCH_OK();
}
int ctx_gate_fwd(const float* s4, long ld_s4, const float* th, long ld_th, float* zst, float* att_cap,
                 long ld_cap, int rows, int D, cudaStream_t s) {
  if (rows <= 0) return SET_OK;
  SET_CHECK_CUDA(launch_chain(ctx_gate_fwd_kernel, dim3(blocks_for((long)rows * D, kThreads)), dim3(kThreads), 0, s, s4, ld_s4, th, ld_th, zst, att_cap, ld_cap, rows, D));
  LAUNCH_OK();
}
int ctx_gate_bwd(const float* zst, const float* datt_cap, long ld_dcap, float* dz_out, float* dtc_out,
                 long ld_ds2, float* dsc, int rows, int D, cudaStream_t s) {
  if (rows <= 0) return SET_OK;
  SET_CHECK_CUDA(launch_chain(ctx_gate_bwd_kernel, dim3(blocks_for((long)rows * D, kThreads)), dim3(kThreads), 0, s, zst, datt_cap, ld_dcap, dz_out, dtc_out, ld_ds2, dsc, rows, D));
  LAUNCH_OK();
}
int copy1_fwd(float* g2, const float* c2_prev, float* cnew, int rows, int D, cudaStream_t s) {
  if (rows <= 0) return SET_OK;
  SET_CHECK_CUDA(launch_chain(copy1_fwd_kernel, dim3(blocks_for((long)rows * D, kThreads)), dim3(kThreads), 0, s, g2, c2_prev, cnew, rows, D));
  LAUNCH_OK();
}
int copy2_fwd(const float* kpre, long ld_k, const float* g2, const float* sel, const float* cnew, float* kgate,
              float* c2, float* h2, float* h2drop, int rows, int D, int train, uint64_t seed, long drop_base,
              cudaStream_t s) {
  if (rows <= 0) return SET_OK;
  SET_CHECK_CUDA(launch_chain(copy2_fwd_kernel, dim3(blocks_for((long)rows * D, kThreads)), dim3(kThreads), 0, s, kpre, ld_k, g2, sel, cnew, kgate, c2, h2, h2drop, rows, D, train, seed, drop_base));
  LAUNCH_OK();
}
int copy2_bwd(const float* dh2_carry, const float* dh2drop_raw, const float* dc2_carry, const float* g2,
              const float* c2, const float* kgate, const float* sel, const float* cnew, float* dg2,
              float* dkpre, float* dsel, float* dcnew, int rows, int D, int train, uint64_t seed, long drop_base,
              cudaStream_t s) {
  if (rows <= 0) return SET_OK;
  SET_CHECK_CUDA(launch_chain(copy2_bwd_kernel, dim3(blocks_for((long)rows * D, kThreads)), dim3(kThreads), 0, s, dh2_carry, dh2drop_raw, dc2_carry, g2, c2, kgate, sel, cnew, dg2, dkpre, dsel, dcnew, rows, D, train, seed, drop_base));
  LAUNCH_OK();
}
int copy1_bwd(const float* dcnew, const float* g2, const float* c2_prev, float* dg2, float* dc2_carry, int rows,
              int D, cudaStream_t s) {
  if (rows <= 0) return SET_OK;
  SET_CHECK_CUDA(launch_chain(copy1_bwd_kernel, dim3(blocks_for((long)rows * D, kThreads)), dim3(kThreads), 0, s, dcnew, g2, c2_prev, dg2, dc2_carry, rows, D));
  LAUNCH_OK();
}
int dropout_fwd(const float* x, float* out, int rows, int D, int train, uint64_t seed, uint32_t site,
                long drop_base, cudaStream_t s) {
  const long n = (long)rows * D;
  if (n <= 0) return SET_OK;
  dropout_fwd_kernel<<<blocks_for(n, kThreads), kThreads, 0, s>>>(x, out, n, train, seed, site, drop_base);
  LAUNCH_OK();
}
int transpose(const float* in, float* out, int rows, int cols, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return SET_OK;
  transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, s>>>(in, cols, out, rows, rows, cols);
  LAUNCH_OK();
}
int transpose_ld(const float* in, long ld_in, float* out, long ld_out, int rows, int cols, cudaStream_t s) {
  if (rows <= 0 || cols <= 0) return SET_OK;
  transpose_kernel<<<dim3((cols + 31) / 32, (rows + 31) / 32), dim3(32, 8), 0, s>>>(in, ld_in, out, ld_out, rows, cols);
  LAUNCH_OK();
}
int transpose_batch(const TrJob* jobs, int n, cudaStream_t s) {
  int i = 0;
  while (i < n) {
    TrBatch b;
    b.n = 0;
    int tiles = 0;
    for (; i < n && b.n < kTrMaxJobs; ++i) {
      if (jobs[i].rows <= 0 || jobs[i].cols <= 0) continue;
      TrJob j = jobs[i];
      j.tiles_c = (j.cols + 63) / 64;
      j.tile0 = tiles;
      tiles += j.tiles_c * ((j.rows + 63) / 64);
      b.j[b.n++] = j;
    }
    if (b.n == 0) break;
    transpose_batch_kernel<<<tiles, 256, 0, s>>>(b);
    SET_CHECK_CUDA(cudaGetLastError());
    set_count_launch(1);
  }
  return SET_OK;
}
int sum_time(const float* x, float* out, int T, long BN, cudaStream_t s) {
  sum_time_kernel<<<blocks_for(BN, kThreads), kThreads, 0, s>>>(x, out, T, BN);
  LAUNCH_OK();
}
int zero_batch(const ZeroJob* jobs, int n, cudaStream_t s) {
  if (n <= 0) return SET_OK;
  if (n > kZeroMaxJobs) { set_record_error("zero_batch: too many ranges"); return SET_ERR_ARG; }
  ZeroBatch zb;
  long longest = 0;
  for (int k = 0; k < n; ++k) { zb.j[k] = jobs[k]; longest = jobs[k].n > longest ? jobs[k].n : longest; }
  for (int k = n; k < kZeroMaxJobs; ++k) zb.j[k] = ZeroJob{nullptr, 0};
  int bx = blocks_for(longest, kThreads);
  if (bx > 1184) bx = 1184;
  zero_batch_kernel<<<dim3(bx, n), kThreads, 0, s>>>(zb);
  LAUNCH_OK();
}
int dropout_keep_mask(float* out, long n, uint64_t seed, uint32_t site, long base, cudaStream_t s) {
  if (n <= 0) return SET_OK;
  keep_mask_kernel<<<blocks_for(n, kThreads), kThreads, 0, s>>>(out, n, seed, site, base);
  LAUNCH_OK();
}

}  // namespace set
