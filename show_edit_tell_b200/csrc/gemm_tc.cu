// fp32-accurate tensor-core GEMM for sm_100a: tcgen05.mma kind::tf32 with the 3xTF32 split,
// operands staged by TMA (128B-swizzled tiles), accumulator in TMEM.
//
//   D[p, q] = sum_k P[p, k] * Q[q, k]           P tile = UMMA "A" (128 rows), Q tile = UMMA "B" (QN rows)
//
// Every fp32 operand x is split on chip into hi = tf32(x) (low 13 mantissa bits cleared, written
// back in place over the TMA'd tile) and lo = x - hi (exact in fp32; written to a sibling tile), and
// each K-step issues  D += P_lo*Q_hi ;  D += P_hi*Q_lo ;  D += P_hi*Q_hi  -- the dropped lo*lo term is
// 2^-22 relative, so results match fp32 FMA accumulation to ~1e-6 (one-pass
// TF32 misses the 1e-4 parity budget by 10x, 3xTF32 meets it).
//
// Either operand may be K-major (global rows = tile rows, reduction contiguous) or MN-major (global
// rows = reduction index, tile rows contiguous), so x@W^T (NT), dy@W (NN) and dy^T@x (TN) all read
// the row-major buffers in place; no transposed copies exist.
//
// "swap" mode puts the weight matrix on the 128-row P side and the (<=64..128 row) activation batch
// on the Q side: that is how the skinny per-step GEMMs (M = batch) fill the tensor core's M=128
// datapath, with split-K spreading one weight matrix over all 148 SMs.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2-5 = hi/lo converters during the main loop, then the epilogue (TMEM -> registers -> global).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <type_traits>

#include "gemm.cuh"

namespace set {

namespace {

constexpr int kBlockK = 32;        // fp32 per smem row: 128 B = one swizzle span
constexpr int kTileP = 128;        // UMMA M
constexpr int kConvWarps = 8;       // hi/lo converters (two per TMEM lane quarter), also the epilogue warps
constexpr int kThreadsTc = 64 + 32 * kConvWarps;

struct TcParams {
  CUtensorMap mapP[4];
  CUtensorMap mapQ[4];
  int K[4];
  int nseg;
  int p_mn, q_mn;                  // 1: operand is MN-major in global memory
  int Pr, Qr;                      // row extents of the two operands
  int swap;                        // 0: (m,n) = (p,q);  1: (m,n) = (q,p)
  int split_k;
  int tiles_p, tiles_q;
  float* C; long ldc; int c_inner; long c_ld_inner; const int* c_row_len; int c_valid_inner;
  const float* bias; const float* bias2;
  const float* add; long ldadd; int add_mod;
  int beta, act;
  int a_tmem;                      // 1: the P operand is fed to the MMA from tensor memory (hi/lo written by tcgen05.st)
  unsigned idesc_xor;              // debugging aid (SET_TC_IDESC_XOR)
  unsigned long long* trace;       // debugging aid: per-phase %globaltimer stamps of CTA 0 (SET_TC_TRACE)
};

__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TC_STAMP(slot)                                                         \
  do {                                                                         \
    if (prm.trace && blockIdx.x == 0) prm.trace[slot] = gtimer();              \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  } while (!done);
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
// K-major / MN-major SWIZZLE_128B shared-memory matrix descriptor (sm_100 format, version 1)
__device__ __forceinline__ uint64_t umma_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((addr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;   // SWIZZLE_128B
  return d;
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
// A operand from tensor memory: [128 lanes] x [8 columns of tf32] at `tmem_a`
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(b_desc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}
constexpr uint32_t kTmemABase = 128;   // columns [0,128): accumulator; [128 + 64*stage, +64): P_hi | P_lo of a stage

__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,"
      "%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr) : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

template <int QN>
struct TcCfg {
  static constexpr int kStages = (QN <= 64) ? 4 : 3;
  static constexpr int kPBytes = kTileP * 128;              // one P tile (hi or lo)
  static constexpr int kQBytes = QN * 128;
  static constexpr int kStageBytes = 2 * kPBytes + 2 * kQBytes;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <int G>
struct TcGroup {
  int n;
  int cta_start[9];        // first CTA of each problem (tiles * split_k each)
  TcParams p[G];
};

template <int QN, int G>
__global__ void __launch_bounds__(kThreadsTc, 1) gemm_tc_kernel(const __grid_constant__ TcGroup<G> grp) {
  using Cfg = TcCfg<QN>;
  int pi = 0;
  while (pi + 1 < grp.n && (int)blockIdx.x >= grp.cta_start[pi + 1]) ++pi;
  const TcParams& prm = grp.p[pi];
  constexpr int S = Cfg::kStages;
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t bar_base = base + S * Cfg::kStageBytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto conv_bar = [&](int s) { return bar_base + 8u * (S + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (2 * S + s); };
  const uint32_t accum_bar = bar_base + 8u * (3 * S);
  const uint32_t tmem_slot = bar_base + 8u * (3 * S + 1);
  uint8_t* gen_base = smem_dyn + (base - smem_u32(smem_dyn));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 64) TC_STAMP(0);
  if (prm.trace && threadIdx.x == 0 && blockIdx.x < 1000) prm.trace[16 + 2 * blockIdx.x] = gtimer();   // per-CTA start

  // tile / split decode
  int bid = blockIdx.x - grp.cta_start[pi];
  const int ks = bid % prm.split_k; bid /= prm.split_k;
  const int qt = bid % prm.tiles_q;
  const int pt = bid / prm.tiles_q;
  const int p0 = pt * kTileP, q0 = qt * QN;

  int nkb_total = 0;
  for (int s = 0; s < prm.nseg; ++s) nkb_total += (prm.K[s] + kBlockK - 1) / kBlockK;
  const int kb_begin = (int)((long)nkb_total * ks / prm.split_k);
  const int kb_end = (int)((long)nkb_total * (ks + 1) / prm.split_k);
  const int nkb = kb_end - kb_begin;

  if (threadIdx.x == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(conv_bar(s), kConvWarps);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"(prm.a_tmem ? 512u : (uint32_t)QN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));
  if (threadIdx.x == 64) TC_STAMP(1);

  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (lane == 0 && nkb > 0) {
      int seg = 0, kb_in_seg = kb_begin;
      while (kb_in_seg >= (prm.K[seg] + kBlockK - 1) / kBlockK) { kb_in_seg -= (prm.K[seg] + kBlockK - 1) / kBlockK; ++seg; }
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S;
        const uint32_t ph = (uint32_t)(i / S) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t st = base + s * Cfg::kStageBytes;
        mbar_expect_tx(full_bar(s), Cfg::kPBytes + Cfg::kQBytes);
        const int k0 = kb_in_seg * kBlockK;
        if (!prm.p_mn) {
          tma_load_2d(st, &prm.mapP[seg], full_bar(s), k0, p0);
        } else {
#pragma unroll
          for (int c = 0; c < kTileP / 32; ++c) tma_load_2d(st + c * 4096, &prm.mapP[seg], full_bar(s), p0 + 32 * c, k0);
        }
        const uint32_t sq = st + 2 * Cfg::kPBytes;
        if (!prm.q_mn) {
          tma_load_2d(sq, &prm.mapQ[seg], full_bar(s), k0, q0);
        } else {
#pragma unroll
          for (int c = 0; c < QN / 32; ++c) tma_load_2d(sq + c * 4096, &prm.mapQ[seg], full_bar(s), q0 + 32 * c, k0);
        }
        if (++kb_in_seg >= (prm.K[seg] + kBlockK - 1) / kBlockK) { kb_in_seg = 0; ++seg; }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (lane == 0 && nkb > 0) {
      // instruction descriptor: D=f32, A=B=tf32, majors, N, M=128
      const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(prm.p_mn ? 1 : 0) << 15) |
                             ((uint32_t)(prm.q_mn ? 1 : 0) << 16) | ((uint32_t)(QN >> 3) << 17) |
                             ((uint32_t)(kTileP >> 4) << 24);
      const uint32_t idesc_final = idesc ^ prm.idesc_xor;
      // per K-step (8 tf32) descriptor advance and strides
      const uint32_t p_step = prm.p_mn ? 1024u : 32u, q_step = prm.q_mn ? 1024u : 32u;
      const uint32_t p_lbo = prm.p_mn ? 4096u : 16u, q_lbo = prm.q_mn ? 4096u : 16u;
      for (int i = 0; i < nkb; ++i) {
        const int s = i % S;
        const uint32_t ph = (uint32_t)(i / S) & 1u;
        mbar_wait(conv_bar(s), ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t st = base + s * Cfg::kStageBytes;
        const uint32_t p_hi = st, p_lo = st + Cfg::kPBytes, q_hi = st + 2 * Cfg::kPBytes,
                       q_lo = st + 2 * Cfg::kPBytes + Cfg::kQBytes;
#pragma unroll
        for (int k = 0; k < kBlockK / 8; ++k) {
          const uint64_t a_hi = umma_desc(p_hi + k * p_step, p_lbo, 1024u);
          const uint64_t a_lo = umma_desc(p_lo + k * p_step, p_lbo, 1024u);
          const uint64_t b_hi = umma_desc(q_hi + k * q_step, q_lbo, 1024u);
          const uint64_t b_lo = umma_desc(q_lo + k * q_step, q_lbo, 1024u);
          if (prm.a_tmem) {
            const uint32_t ta_hi = tmem_base + kTmemABase + (uint32_t)s * 64u + (uint32_t)k * 8u;
            umma_tf32_ts(tmem_base, ta_hi + 32u, b_hi, idesc_final, (i > 0 || k > 0) ? 1u : 0u);
            umma_tf32_ts(tmem_base, ta_hi, b_lo, idesc_final, 1u);
            umma_tf32_ts(tmem_base, ta_hi, b_hi, idesc_final, 1u);
          } else {
            umma_tf32(tmem_base, a_lo, b_hi, idesc_final, (i > 0 || k > 0) ? 1u : 0u);
            umma_tf32(tmem_base, a_hi, b_lo, idesc_final, 1u);
            umma_tf32(tmem_base, a_hi, b_hi, idesc_final, 1u);
          }
        }
        umma_commit(empty_bar(s));   // smem slot reusable once these MMAs retire
      }
      umma_commit(accum_bar);
    }
  } else {
    // ============================== converters, then epilogue ==============================
    const int ct = threadIdx.x - 64;   // 0 .. 32*kConvWarps-1
    constexpr int kCT = 32 * kConvWarps;
    const int khalf = (warp - 2) >> 2;  // which 16-column half of the K-block this warp converts (A-from-TMEM)
    for (int i = 0; i < nkb; ++i) {
      const int s = i % S;
      const uint32_t ph = (uint32_t)(i / S) & 1u;
      mbar_wait(full_bar(s), ph);
      if (ct == 0 && i == 0) TC_STAMP(2);
      uint8_t* st = gen_base + s * Cfg::kStageBytes;
      float4* p_hi = reinterpret_cast<float4*>(st);
      float4* p_lo = reinterpret_cast<float4*>(st + Cfg::kPBytes);
      float4* q_hi = reinterpret_cast<float4*>(st + 2 * Cfg::kPBytes);
      float4* q_lo = reinterpret_cast<float4*>(st + 2 * Cfg::kPBytes + Cfg::kQBytes);
      auto split = [](float x, float& hi, float& lo) {
        hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        lo = x - hi;
      };
      if (prm.a_tmem) {
        // P (the 128-row operand) goes to tensor memory: this thread owns tile row `prow` = its TMEM lane,
        // reads the row's 32 fp32 out of the 128B-swizzled tile (16-byte chunk c of row r sits at chunk
        // c ^ (r & 7)) and stores hi | lo as 2 x 32 columns.  No shared-memory write-back, and the MMA
        // no longer re-reads P from shared memory -- the two largest smem streams of the SS form.
        const int prow = (warp & 3) * 32 + lane;
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int cch = khalf * 4 + cc;
          const float4 v = p_hi[prow * 8 + (cch ^ (prow & 7))];
          float h, l;
          split(v.x, h, l); hi[cc * 4 + 0] = __float_as_uint(h); lo[cc * 4 + 0] = __float_as_uint(l);
          split(v.y, h, l); hi[cc * 4 + 1] = __float_as_uint(h); lo[cc * 4 + 1] = __float_as_uint(l);
          split(v.z, h, l); hi[cc * 4 + 2] = __float_as_uint(h); lo[cc * 4 + 2] = __float_as_uint(l);
          split(v.w, h, l); hi[cc * 4 + 3] = __float_as_uint(h); lo[cc * 4 + 3] = __float_as_uint(l);
        }
        const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + kTmemABase + (uint32_t)s * 64u +
                            (uint32_t)khalf * 16u;
        tmem_st16(ta, hi);
        tmem_st16(ta + 32u, lo);
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      } else {
#pragma unroll
        for (int j = 0; j < Cfg::kPBytes / 16 / kCT; ++j) {
          const float4 v = p_hi[ct + kCT * j];
          float4 h, l;
          split(v.x, h.x, l.x); split(v.y, h.y, l.y); split(v.z, h.z, l.z); split(v.w, h.w, l.w);
          p_hi[ct + kCT * j] = h;
          p_lo[ct + kCT * j] = l;
        }
      }
#pragma unroll
      for (int j = 0; j < Cfg::kQBytes / 16 / kCT; ++j) {
        const float4 v = q_hi[ct + kCT * j];
        float4 h, l;
        split(v.x, h.x, l.x); split(v.y, h.y, l.y); split(v.z, h.z, l.z); split(v.w, h.w, l.w);
        q_hi[ct + kCT * j] = h;
        q_lo[ct + kCT * j] = l;
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the MMA (async proxy)
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(conv_bar(s));
      if (ct == 0 && i == 0) TC_STAMP(3);
      if (ct == 0 && i == nkb - 1) TC_STAMP(4);
    }
    // ---- epilogue
    if (nkb > 0) {
      mbar_wait(accum_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    if (ct == 0) TC_STAMP(5);
    // Accumulator -> registers -> shared (the pipeline stages are idle now).  The tile is staged in the
    // orientation of the OUTPUT rows (transposed for swap mode) so that the second phase reads float4
    // along the contiguous global direction: 128-bit global loads/stores/reductions, several rows in
    // flight per thread, no serial latency chain.
    const int quarter = warp & 3;              // TMEM lane quarter this warp may access
    const bool stager = (warp - 2) < 4;         // one warp per quarter moves TMEM -> smem
    constexpr int EPW_N = QN + 4;              // staged row pitch, non-swap: [128 p][QN q]
    constexpr int EPW_S = kTileP + 4;          // staged row pitch, swap:     [QN q][128 p]
    float* ep = reinterpret_cast<float*>(gen_base);
    const int prow = quarter * 32 + lane;      // tile row held by this thread
#pragma unroll 1
    for (int cb = 0; stager && cb < QN / 32; ++cb) {
      uint32_t r[32];
      if (nkb > 0) {
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cb * 32), r);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) r[j] = 0u;
      }
      if (prm.swap) {
#pragma unroll
        for (int j = 0; j < 32; ++j) ep[(cb * 32 + j) * EPW_S + prow] = __uint_as_float(r[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(&ep[prow * EPW_N + cb * 32 + j]) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]),
                          __uint_as_float(r[j + 3]));
      }
    }
    asm volatile("bar.sync 1, %0;" ::"n"(32 * kConvWarps) : "memory");   // converter/epilogue warps only
    if (ct == 0) TC_STAMP(6);
    const bool lead = (ks == 0);               // split 0 carries bias / addend
    const bool atomic = prm.split_k > 1;
    const int width = prm.swap ? kTileP : QN;  // contiguous extent of a staged row
    const int pitch = prm.swap ? EPW_S : EPW_N;
    const int m_base = prm.swap ? q0 : p0, n_base = prm.swap ? p0 : q0;
    const int m_lim = prm.swap ? prm.Qr : prm.Pr, n_lim = prm.swap ? prm.Pr : prm.Qr;
    const int vec_per_row = width / 4;
    const int total_vec = (prm.swap ? QN : kTileP) * vec_per_row;
#pragma unroll 4
    for (int e = ct; e < total_vec; e += kCT) {
      const int o = e / vec_per_row, i4 = (e % vec_per_row) * 4;
      const int m = m_base + o, n = n_base + i4;
      if (m >= m_lim || n >= n_lim) continue;
      if (prm.c_row_len && !(prm.c_row_len[m % prm.c_valid_inner] > m / prm.c_valid_inner)) continue;
      const float4 acc = *reinterpret_cast<const float4*>(&ep[o * pitch + i4]);
      float v[4] = {acc.x, acc.y, acc.z, acc.w};
      float* cp = prm.C + (prm.c_inner > 0 ? (long)(m / prm.c_inner) * prm.ldc + (long)(m % prm.c_inner) * prm.c_ld_inner
                                           : (long)m * prm.ldc) + n;
      const int nv = min(4, n_lim - n);
      const bool vec_ok = (nv == 4) && ((reinterpret_cast<uintptr_t>(cp) & 15) == 0);
      if (!atomic || lead) {
        if (prm.bias)
          for (int k = 0; k < nv; ++k) v[k] += __ldg(prm.bias + n + k);
        if (prm.bias2)
          for (int k = 0; k < nv; ++k) v[k] += __ldg(prm.bias2 + n + k);
        if (prm.add) {
          const float* ap = prm.add + (long)(prm.add_mod ? m % prm.add_mod : m) * prm.ldadd + n;
          for (int k = 0; k < nv; ++k) v[k] += ap[k];
        }
      }
      if (atomic) {
        if (vec_ok) {
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(cp), "f"(v[0]), "f"(v[1]), "f"(v[2]),
                       "f"(v[3]) : "memory");
        } else {
          for (int k = 0; k < nv; ++k) atomicAdd(cp + k, v[k]);
        }
        continue;
      }
      if (prm.act == 1) { for (int k = 0; k < 4; ++k) v[k] = fmaxf(v[k], 0.f); }
      else if (prm.act == 2) { for (int k = 0; k < 4; ++k) v[k] = tanhf(v[k]); }
      if (vec_ok) {
        if (prm.beta) {
          const float4 old = *reinterpret_cast<const float4*>(cp);
          v[0] += old.x; v[1] += old.y; v[2] += old.z; v[3] += old.w;
        }
        *reinterpret_cast<float4*>(cp) = make_float4(v[0], v[1], v[2], v[3]);
      } else {
        for (int k = 0; k < nv; ++k) cp[k] = prm.beta ? cp[k] + v[k] : v[k];
      }
    }
  }
  if (threadIdx.x == 64) TC_STAMP(7);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 64) TC_STAMP(8);
  if (prm.trace && threadIdx.x == 0 && blockIdx.x < 1000) prm.trace[17 + 2 * blockIdx.x] = gtimer();   // per-CTA end
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(prm.a_tmem ? 512u : (uint32_t)QN) : "memory");
  }
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
bool g_tc_ready = false, g_tc_failed = false;
std::once_flag g_tc_once;

void tc_init() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
    g_tc_failed = true;
    return;
  }
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  bool ok = true;
  auto set_attr = [&](auto kern, int bytes) {
    ok = ok && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
  };
  set_attr(gemm_tc_kernel<64, 1>, TcCfg<64>::kSmemBytes);  set_attr(gemm_tc_kernel<128, 1>, TcCfg<128>::kSmemBytes);
  set_attr(gemm_tc_kernel<64, 2>, TcCfg<64>::kSmemBytes);  set_attr(gemm_tc_kernel<128, 2>, TcCfg<128>::kSmemBytes);
  set_attr(gemm_tc_kernel<64, 5>, TcCfg<64>::kSmemBytes);  set_attr(gemm_tc_kernel<128, 5>, TcCfg<128>::kSmemBytes);
  set_attr(gemm_tc_kernel<64, 8>, TcCfg<64>::kSmemBytes);  set_attr(gemm_tc_kernel<128, 8>, TcCfg<128>::kSmemBytes);
  if (!ok) {
    cudaGetLastError();
    g_tc_failed = true;
    return;
  }
  g_tc_ready = true;
}

// rows x cols fp32 matrix with row stride ld (elements); box = box_cols x box_rows
bool make_map(CUtensorMap* m, const float* ptr, long rows, long cols, long ld, int box_cols, int box_rows) {
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

unsigned long long* g_tc_trace = nullptr;

bool aligned_ok(const float* p, long ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld % 4) == 0; }

}  // namespace

// Fills `prm` for one problem; returns false if the problem is not eligible for the tensor-core path
// (caller falls back to the CUDA-core kernel).  `QN` is the Q-tile width chosen for the whole group.
static bool tc_plan(int mode, const GemmProblem& g, int QN, TcParams& prm) {
  if (g.nseg < 1 || g.M <= 0 || g.N <= 0) return false;
  // MN-major operands (the NN / TN forms) are wired through the kernel but read back as zeros on
  // sm_100a with these descriptors (tools/debug_tc2.py) -- until that is understood only the
  // K-major/K-major (NT) form runs on tensor cores; callers present NN/TN work in NT form on
  // transposed copies (editnet.cu backward_core) or fall back to the CUDA-core kernel.
  if (mode != kNT && getenv("SET_TC_ALLOW_MN") == nullptr) return false;
  if (g.a_inner > 0 || g.a_row_len) return false;         // two-level / masked A rows stay on the CUDA-core path
  long ktot = 0;
  for (int s = 0; s < g.nseg; ++s) {
    if (!aligned_ok(g.seg[s].A, g.seg[s].lda) || !aligned_ok(g.seg[s].B, g.seg[s].ldb)) return false;
    ktot += g.seg[s].K;
  }
  if (ktot < 64) return false;
  memset(&prm, 0, sizeof(prm));
  // operand roles: the A-side (M rows) and the B-side (N rows) of the logical GEMM
  //   kNT: A[m][k] K-major,   B[n][k] K-major
  //   kNN: A[m][k] K-major,   B[k][n] MN-major
  //   kTN: A[k][m] MN-major,  B[k][n] MN-major
  const int a_mn = (mode == kTN), b_mn = (mode != kNT);
  // skinny M: weights (B side, N rows) take the 128-row P role
  const bool swap = (g.M <= QN && g.N > g.M);
  const int Pr = swap ? g.N : g.M, Qr = swap ? g.M : g.N;
  if (swap && (g.c_inner > 0)) return false;
  prm.swap = swap; prm.Pr = Pr; prm.Qr = Qr;
  prm.p_mn = swap ? b_mn : a_mn;
  prm.q_mn = swap ? a_mn : b_mn;
  prm.nseg = g.nseg;
  for (int s = 0; s < g.nseg; ++s) {
    const GemmSeg& sg = g.seg[s];
    prm.K[s] = sg.K;
    const float* Pp = swap ? sg.B : sg.A; const long Pld = swap ? sg.ldb : sg.lda;
    const float* Qp = swap ? sg.A : sg.B; const long Qld = swap ? sg.lda : sg.ldb;
    bool ok;
    if (!prm.p_mn) ok = make_map(&prm.mapP[s], Pp, Pr, sg.K, Pld, kBlockK, kTileP);
    else ok = make_map(&prm.mapP[s], Pp, sg.K, Pr, Pld, 32, kBlockK);
    if (!prm.q_mn) ok = ok && make_map(&prm.mapQ[s], Qp, Qr, sg.K, Qld, kBlockK, QN);
    else ok = ok && make_map(&prm.mapQ[s], Qp, sg.K, Qr, Qld, 32, kBlockK);
    if (!ok) return false;
  }
  prm.tiles_p = (Pr + kTileP - 1) / kTileP;
  prm.tiles_q = (Qr + QN - 1) / QN;
  prm.split_k = 1;
  prm.C = g.C; prm.ldc = g.ldc; prm.c_inner = g.c_inner; prm.c_ld_inner = g.c_ld_inner;
  prm.c_row_len = g.c_row_len; prm.c_valid_inner = g.c_valid_inner > 0 ? g.c_valid_inner : 1;
  prm.bias = g.bias; prm.bias2 = g.bias2; prm.add = g.add; prm.ldadd = g.ldadd; prm.add_mod = g.add_mod;
  prm.beta = g.beta; prm.act = g.act;
  { const char* e = getenv("SET_TC_IDESC_XOR"); prm.idesc_xor = e ? (unsigned)strtoul(e, nullptr, 0) : 0u; }
  prm.trace = g_tc_trace;
  {
    static const int atmem = getenv("SET_TC_ATMEM") ? atoi(getenv("SET_TC_ATMEM")) : 1;
    prm.a_tmem = (atmem && !prm.p_mn) ? 1 : 0;
  }
  return true;
}

// Launches every eligible problem of the group in ONE grid (taken[i] = true); the others are left to
// the CUDA-core kernel.  Independent GEMMs of one phase of the decode step (everything that consumes
// h1, say) thereby stream their weights concurrently instead of paying a launch each.
int gemm_tc_try_group(int mode, const GemmProblem* probs, int n, bool* taken, cudaStream_t stream) {
  for (int i = 0; i < n; ++i) taken[i] = false;
  std::call_once(g_tc_once, tc_init);
  if (!g_tc_ready) return SET_OK;
  // one Q-tile width per launch: 64 if every problem's small side fits, else 128
  int QN = 64;
  for (int i = 0; i < n; ++i) {
    const GemmProblem& g = probs[i];
    if (g.M <= 0 || g.N <= 0) continue;
    const int small = g.M < g.N ? g.M : g.N;
    if (small > 64) QN = 128;
  }
  static TcGroup<8> grp;  // host staging (launch copies it); calls are serialised by the caller's stream use
  grp.n = 0;
  long tiles_total = 0;
  int idx[8];
  for (int i = 0; i < n; ++i) {
    if (probs[i].M <= 0 || probs[i].N <= 0) { taken[i] = true; continue; }
    if (!tc_plan(mode, probs[i], QN, grp.p[grp.n])) continue;
    idx[grp.n] = i;
    tiles_total += (long)grp.p[grp.n].tiles_p * grp.p[grp.n].tiles_q;
    ++grp.n;
  }
  if (grp.n == 0) return SET_OK;
  // split-K: spread the group over ~all SMs (partials meet in global reductions)
  int cta = 0;
  for (int k = 0; k < grp.n; ++k) {
    TcParams& prm = grp.p[k];
    const GemmProblem& g = probs[idx[k]];
    long nkb = 0;
    for (int s = 0; s < g.nseg; ++s) nkb += (g.seg[s].K + kBlockK - 1) / kBlockK;
    int split = 1;
    if (g.act == 0 && tiles_total < 148 && !(g.c_inner > 0 || g.c_row_len)) {
      // fill one wave of the 148 SMs (a second wave would repeat every CTA's fixed prologue/epilogue)
      split = (int)(148 / tiles_total);
      const int max_split = (int)(nkb / 4 > 0 ? nkb / 4 : 1);
      if (split > max_split) split = max_split;
      if (split > 16) split = 16;
      if (split < 1) split = 1;
      // a long-K problem that fills only ~half the machine: three partials over two waves is ~1.5x faster
      if (split == 1 && tiles_total * 3 <= 2 * 148 && nkb >= 96) split = 3;
    }
    prm.split_k = split;
    if (split > 1 && !g.beta && !g.c_zeroed)   // partial sums are reduced into C: it must start at zero
      SET_CHECK_CUDA(cudaMemset2DAsync(g.C, sizeof(float) * g.ldc, 0, sizeof(float) * g.N, g.M, stream));
    grp.cta_start[k] = cta;
    cta += prm.tiles_p * prm.tiles_q * split;
    taken[idx[k]] = true;
  }
  grp.cta_start[grp.n] = cta;
  auto launch = [&](auto tag) {
    constexpr int G = decltype(tag)::value;
    TcGroup<G> small;
    small.n = grp.n;
    memcpy(small.cta_start, grp.cta_start, sizeof(small.cta_start));
    memcpy(small.p, grp.p, sizeof(TcParams) * grp.n);
    if (QN == 64) gemm_tc_kernel<64, G><<<cta, kThreadsTc, TcCfg<64>::kSmemBytes, stream>>>(small);
    else gemm_tc_kernel<128, G><<<cta, kThreadsTc, TcCfg<128>::kSmemBytes, stream>>>(small);
  };
  if (grp.n == 1) launch(std::integral_constant<int, 1>{});
  else if (grp.n == 2) launch(std::integral_constant<int, 2>{});
  else if (grp.n <= 5) launch(std::integral_constant<int, 5>{});
  else launch(std::integral_constant<int, 8>{});
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  return SET_OK;
}

// debugging: device buffer of >= 16 u64 that CTA 0 of every following tensor-core launch stamps
void gemm_tc_set_trace(unsigned long long* buf) { g_tc_trace = buf; }

}  // namespace set
