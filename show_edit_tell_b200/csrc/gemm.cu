// FP32 CUDA-core GEMM used for every dense contraction of the path in this round.
// Tiled 64 x {64,32} x 16, register-prefetch double buffering, grouped launch (up to 8
// independent problems per grid), K-segments, fused bias/addend/relu epilogue.  Exact
// fp32 FMA accumulation keeps the 1e-4 parity budget trivially; the tcgen05 3xTF32
// path that replaces it for the weight-streaming shapes is described in DESIGN.md.
#include <string.h>

#include "gemm.cuh"

namespace set {

namespace {

constexpr int BM = 64;
constexpr int BK = 16;

__device__ __forceinline__ long row_off(long x, long ld, int inner, long ld_inner) {
  return inner > 0 ? (x / inner) * ld + (x % inner) * ld_inner : x * ld;
}
__device__ __forceinline__ bool row_valid(const int* row_len, int vi, long x) {
  return row_len == nullptr || row_len[x % vi] > (int)(x / vi);
}

// 4 consecutive floats starting at p, element q valid iff q < nvalid; vectorised when possible
__device__ __forceinline__ float4 load4(const float* p, int nvalid) {
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (nvalid >= 4 && ((reinterpret_cast<uintptr_t>(p) & 15) == 0)) {
    v = __ldg(reinterpret_cast<const float4*>(p));
  } else {
    if (nvalid > 0) v.x = __ldg(p);
    if (nvalid > 1) v.y = __ldg(p + 1);
    if (nvalid > 2) v.z = __ldg(p + 2);
    if (nvalid > 3) v.w = __ldg(p + 3);
  }
  return v;
}

template <int MODE, int BN>
__global__ void __launch_bounds__(256) gemm_kernel(const __grid_constant__ GemmGroup g) {
  constexpr int TNW = BN / 16;  // output columns per thread
  __shared__ __align__(16) float As[2][BK][BM + 4];
  __shared__ __align__(16) float Bs[2][BK][BN + 4];

  int tile = blockIdx.x, pi = 0;
  while (pi + 1 < g.n && tile >= g.tile_start[pi + 1]) ++pi;
  const GemmProblem& P = g.p[pi];
  tile -= g.tile_start[pi];
  const int tiles_n = (P.N + BN - 1) / BN;
  const int m0 = (tile / tiles_n) * BM, n0 = (tile % tiles_n) * BN;
  const int tid = threadIdx.x;
  const int ty = tid >> 4, tx = tid & 15;

  float acc[4][TNW];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < TNW; ++j) acc[i][j] = 0.f;

  int ntiles = 0;
  for (int s = 0; s < P.nseg; ++s) ntiles += (P.seg[s].K + BK - 1) / BK;

  // loader slots
  const int a_i = (MODE == kTN) ? (tid & 15) * 4 : (tid >> 2);   // m offset
  const int a_r = (MODE == kTN) ? (tid >> 4) : (tid & 3) * 4;    // k offset
  constexpr int BROW_T = BN / 4;                                  // threads per B row (j-contig)
  const int b_j = (MODE == kNT) ? (tid >> 2) : (tid % BROW_T) * 4;
  const int b_r = (MODE == kNT) ? (tid & 3) * 4 : (tid / BROW_T);
  const bool b_active = (MODE == kNT) ? (b_j < BN) : (b_r < BK);

  bool a_ok = false;  // NT/NN: this thread's A row exists and is valid
  if (MODE != kTN) {
    const long x = m0 + a_i;
    a_ok = (x < P.M) && row_valid(P.a_row_len, P.a_valid_inner, x);
  }

  float4 ra, rb;
  int seg = 0, k0 = 0;

  auto gload = [&](int sg, int kk) {
    const GemmSeg& S = P.seg[sg];
    ra = make_float4(0.f, 0.f, 0.f, 0.f);
    rb = make_float4(0.f, 0.f, 0.f, 0.f);
    if (MODE != kTN) {
      if (a_ok) {
        const long off = row_off(m0 + a_i, S.lda, P.a_inner, P.a_ld_inner);
        ra = load4(S.A + off + kk + a_r, S.K - (kk + a_r));
      }
    } else {
      const long x = kk + a_r;
      if (x < S.K && row_valid(P.a_row_len, P.a_valid_inner, x)) {
        const long off = row_off(x, S.lda, P.a_inner, P.a_ld_inner);
        ra = load4(S.A + off + m0 + a_i, P.M - (m0 + a_i));
      }
    }
    if (b_active) {
      if (MODE == kNT) {
        if (n0 + b_j < P.N) rb = load4(S.B + (long)(n0 + b_j) * S.ldb + kk + b_r, S.K - (kk + b_r));
      } else {
        if (kk + b_r < S.K) rb = load4(S.B + (long)(kk + b_r) * S.ldb + n0 + b_j, P.N - (n0 + b_j));
      }
    }
  };
  auto sstore = [&](int buf) {
    if (MODE != kTN) {
      As[buf][a_r + 0][a_i] = ra.x; As[buf][a_r + 1][a_i] = ra.y;
      As[buf][a_r + 2][a_i] = ra.z; As[buf][a_r + 3][a_i] = ra.w;
    } else {
      *reinterpret_cast<float4*>(&As[buf][a_r][a_i]) = ra;
    }
    if (b_active) {
      if (MODE == kNT) {
        Bs[buf][b_r + 0][b_j] = rb.x; Bs[buf][b_r + 1][b_j] = rb.y;
        Bs[buf][b_r + 2][b_j] = rb.z; Bs[buf][b_r + 3][b_j] = rb.w;
      } else {
        *reinterpret_cast<float4*>(&Bs[buf][b_r][b_j]) = rb;
      }
    }
  };
  auto advance = [&]() {
    k0 += BK;
    if (k0 >= P.seg[seg].K) { ++seg; k0 = 0; }
  };

  if (ntiles > 0) {
    gload(seg, k0);
    sstore(0);
    advance();
  }
  __syncthreads();
  for (int it = 0; it < ntiles; ++it) {
    const int buf = it & 1;
    const bool more = (it + 1 < ntiles);
    if (more) gload(seg, k0);
#pragma unroll
    for (int r = 0; r < BK; ++r) {
      const float4 a = *reinterpret_cast<const float4*>(&As[buf][r][ty * 4]);
      float b[TNW];
      if constexpr (TNW == 4) {
        const float4 bv = *reinterpret_cast<const float4*>(&Bs[buf][r][tx * 4]);
        b[0] = bv.x; b[1] = bv.y; b[2] = bv.z; b[3] = bv.w;
      } else {
        const float2 bv = *reinterpret_cast<const float2*>(&Bs[buf][r][tx * 2]);
        b[0] = bv.x; b[1] = bv.y;
      }
      const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < TNW; ++j) acc[i][j] = fmaf(av[i], b[j], acc[i][j]);
    }
    if (more) {
      sstore(buf ^ 1);
      advance();
    }
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= P.M) continue;
    if (!row_valid(P.c_row_len, P.c_valid_inner, m)) continue;
    float* crow = P.C + row_off(m, P.ldc, P.c_inner, P.c_ld_inner);
    const float* addrow = P.add ? P.add + (long)(P.add_mod ? m % P.add_mod : m) * P.ldadd : nullptr;
#pragma unroll
    for (int j = 0; j < TNW; ++j) {
      const int n = n0 + tx * TNW + j;
      if (n >= P.N) continue;
      float v = acc[i][j];
      if (P.bias) v += __ldg(P.bias + n);
      if (P.bias2) v += __ldg(P.bias2 + n);
      if (addrow) v += addrow[n];
      if (P.act == 1) v = fmaxf(v, 0.f);
      else if (P.act == 2) v = tanhf(v);
      if (P.beta) v += crow[n];
      crow[n] = v;
    }
  }
}

template <int MODE, int BN>
int launch(const GemmGroup& g, int total_tiles, cudaStream_t stream) {
  static const bool once = [] {   // same carveout preference as the tensor-core kernel (see cells.cu)
    cudaFuncSetAttribute(gemm_kernel<MODE, BN>, cudaFuncAttributePreferredSharedMemoryCarveout,
                         cudaSharedmemCarveoutMaxShared);
    return true;
  }();
  (void)once;
  gemm_kernel<MODE, BN><<<total_tiles, 256, 0, stream>>>(g);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  return SET_OK;
}

// several column sums in one grid; a block = 32 columns x (8 row lanes x 8 rows each) of one job
__global__ void __launch_bounds__(256) colsum_batch_kernel(const __grid_constant__ ColBatch b) {
  __shared__ float red[8][33];
  int ji = 0;
  while (ji + 1 < b.n && (int)blockIdx.x >= b.j[ji + 1].block0) ++ji;
  const ColJob& J = b.j[ji];
  const int t = blockIdx.x - J.block0;
  const int cb = t % J.col_blocks, rb = t / J.col_blocks;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = cb * 32 + tx;
  const int m0 = rb * 64 + ty;
  float v[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const int m = m0 + 8 * k;
    v[k] = (n < J.N && m < J.M) ? J.X[(long)m * J.ld + n] : 0.f;
  }
  float sacc = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) sacc += v[k];
  red[ty][tx] = sacc;
  __syncthreads();
  if (ty == 0 && n < J.N) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) tot += red[k][tx];
    atomicAdd(J.out + n, tot);
  }
}

}  // namespace

int colsum_batch(const ColJob* jobs, int n, cudaStream_t stream) {
  int i = 0;
  while (i < n) {
    ColBatch b;
    b.n = 0;
    int blocks = 0;
    for (; i < n && b.n < kColMaxJobs; ++i) {
      if (jobs[i].M <= 0 || jobs[i].N <= 0) continue;
      ColJob j = jobs[i];
      j.col_blocks = (j.N + 31) / 32;
      j.block0 = blocks;
      blocks += j.col_blocks * ((j.M + 63) / 64);
      b.j[b.n++] = j;
    }
    if (b.n == 0) break;
    colsum_batch_kernel<<<blocks, 256, 0, stream>>>(b);
    SET_CHECK_CUDA(cudaGetLastError());
    set_count_launch(1);
  }
  return SET_OK;
}

int g_backend = 0;              // 0: tensor cores where eligible, 1: CUDA cores only
int g_pdl = getenv("SET_PDL") ? atoi(getenv("SET_PDL")) : 1;
long long g_tc_launches = 0, g_simt_launches = 0, g_tc_twin_launches = 0;

int gemm_group(int mode, const GemmProblem* probs_in, int n, cudaStream_t stream) {
  SET_REQUIRE(n >= 1 && n <= 8, "1..8 problems per group");
  GemmProblem probs[8];
  bool taken[8] = {false, false, false, false, false, false, false, false};
  if (g_backend == 0) {
    SET_PROPAGATE(gemm_tc_try_group(mode, probs_in, n, taken, stream));
    for (int i = 0; i < n; ++i)
      if (taken[i]) { ++g_tc_launches; break; }
  }
  int kept = 0;
  for (int i = 0; i < n; ++i) {
    if (probs_in[i].M <= 0 || probs_in[i].N <= 0 || taken[i]) continue;
    probs[kept++] = probs_in[i];
  }
  n = kept;
  if (n == 0) return SET_OK;
  ++g_simt_launches;
  GemmGroup g;
  memset(&g, 0, sizeof(g));
  // skinny problems get 32-wide column tiles so that more CTAs stream weights
  long tiles64 = 0;
  for (int i = 0; i < n; ++i)
    tiles64 += (long)((probs[i].M + BM - 1) / BM) * ((probs[i].N + 63) / 64);
  const int bn = (tiles64 < 2 * 148) ? 32 : 64;
  int total = 0, cnt = 0;
  for (int i = 0; i < n; ++i) {
    const GemmProblem& p = probs[i];
    if (p.M <= 0 || p.N <= 0) continue;
    SET_REQUIRE(p.C != nullptr, "null C");
    SET_REQUIRE(p.nseg >= 0 && p.nseg <= 4, "segments");
    g.p[cnt] = p;
    g.tile_start[cnt] = total;
    total += ((p.M + BM - 1) / BM) * ((p.N + bn - 1) / bn);
    ++cnt;
  }
  g.n = cnt;
  g.tile_start[cnt] = total;
  if (total == 0) return SET_OK;
  if (bn == 64) {
    if (mode == kNT) return launch<kNT, 64>(g, total, stream);
    if (mode == kNN) return launch<kNN, 64>(g, total, stream);
    if (mode == kTN) return launch<kTN, 64>(g, total, stream);
  } else {
    if (mode == kNT) return launch<kNT, 32>(g, total, stream);
    if (mode == kNN) return launch<kNN, 32>(g, total, stream);
    if (mode == kTN) return launch<kTN, 32>(g, total, stream);
  }
  SET_REQUIRE(false, "bad gemm mode");
  return SET_OK;
}

int colsum(const float* X, long ld, int M, int N, float* out, int beta, cudaStream_t stream) {
  if (N <= 0) return SET_OK;
  if (!beta) SET_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, stream));
  if (M <= 0) return SET_OK;
  const ColJob j{X, ld, M, N, out, 0, 0};
  return colsum_batch(&j, 1, stream);
}

}  // namespace set
