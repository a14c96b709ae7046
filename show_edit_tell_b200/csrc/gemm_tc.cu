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
// Both operands are K-major (global rows = tile rows, reduction contiguous): x@W^T (NT) reads the row-major
// buffers in place; callers present dX / dW work in NT form on transposed copies (MN-major tf32 operands
// read back as zeros with sm_100a descriptors built this way, tools/debug_tc2.py).
//
// "swap" mode puts the weight matrix on the 128-row P side and the (<=64..128 row) activation batch
// on the Q side: that is how the skinny per-step GEMMs (M = batch) fill the tensor core's M=128
// datapath, with split-K spreading one weight matrix over all 148 SMs.
//
// The 128-row operand never makes a second trip through shared memory: converter warps read the TMA'd
// fp32 tile once, split it in registers and write hi | lo into TENSOR memory (tcgen05.st), from where the
// MMA takes its A operand; only the small Q tile is split in place in shared memory.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer, warps 2.. = hi/lo converters
// (groups of 4 warps alternate K-blocks) during the main loop, then the epilogue (TMEM -> registers ->
// shared -> global).
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
#ifndef SET_TC_CONV_WARPS
#define SET_TC_CONV_WARPS 8
#endif
constexpr int kConvWarps = SET_TC_CONV_WARPS;   // hi/lo converters, also the epilogue warps
constexpr int kConvGroups = kConvWarps / 4;     // a group = 4 warps = the 4 TMEM lane quarters; K-block i belongs to
                                                // group i % kConvGroups, so the groups' latency chains overlap
constexpr int kGT = 128;                        // threads per converter group
constexpr int kThreadsTc = 64 + 32 * kConvWarps;

struct TcParams {
  CUtensorMap mapP[4];
  CUtensorMap mapQ[4];
  int K[4];
  int nseg;
  int Pr, Qr;                      // row extents of the two operands
  int swap;                        // 0: (m,n) = (p,q);  1: (m,n) = (q,p)
  int split_k;
  int tiles_p, tiles_q;
  float* C; long ldc; int c_inner; long c_ld_inner; const int* c_row_len; int c_valid_inner;
  const float* bias; const float* bias2;
  const float* add; long ldadd; int add_mod;
  int beta, act;
  int pre_p, pre_q;                // operand is a constant weight: its first pipeline stages load before pdl_wait()
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

// per-K-block SM-clock stamps of CTA 0 (slots: 0 producer issued Q, 1 producer issued P, 2 converter saw Q,
// 3 converter saw P, 4 converter done, 5 MMA saw converted block, 6 MMA issued)
#define KB_STAMP(i, slot)                                                                       \
  do {                                                                                          \
    if (prm.trace && blockIdx.x == 0 && (i) < 48) prm.trace[400 + 8 * (i) + (slot)] = clock64(); \
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
// tensor-memory columns [0,QN): accumulator; [QN + 64*slot, +64): P_hi | P_lo of a Q/TMEM slot

// one lane of a converged warp; ptxas then treats the guarded block as warp-uniform (tcgen05.mma issues
// straight from uniform registers, no per-lane replay loop around it)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "elect.sync _|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
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
  // Two rings.  P: raw fp32 tiles of the 128-row operand; a K-block of P is converted straight into tensor
  // memory, so the ring only has to cover the HBM latency -- it is the deep one (Little: 148 SMs x kNP x 16 KB
  // in flight ~ 19 MB >= 6.5 TB/s x 2 us).  Q: hi | lo tiles the MMA reads from shared memory, paired with the
  // tensor-memory slots of P_hi | P_lo; both are released by the MMA's commit.
#ifndef SET_TC_NQ64
#define SET_TC_NQ64 6
#define SET_TC_NP64 6
#endif
  static constexpr int kNQ = (QN <= 64) ? SET_TC_NQ64 : 3;
  static constexpr int kNP = (QN <= 64) ? SET_TC_NP64 : 6;
  static constexpr int kPBytes = kTileP * 128;
  static constexpr int kQBytes = QN * 128;
  static constexpr int kQSlot = 2 * kQBytes;
  static constexpr int kRingBytes = kNP * kPBytes + kNQ * kQSlot;
  static constexpr int kSmemBytes = kRingBytes + 1024 /*align*/ + 256 /*barriers*/;
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
  constexpr int NP = Cfg::kNP, NQ = Cfg::kNQ;
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  const uint32_t q_base = base + NP * Cfg::kPBytes;
  const uint32_t bar_base = base + Cfg::kRingBytes;
  auto p_full = [&](int s) { return bar_base + 8u * s; };
  auto q_full = [&](int s) { return bar_base + 8u * (NP + s); };
  auto conv_bar = [&](int s) { return bar_base + 8u * (NP + NQ + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (NP + 2 * NQ + s); };
  const uint32_t accum_bar = bar_base + 8u * (NP + 3 * NQ);
  const uint32_t tmem_slot = bar_base + 8u * (NP + 3 * NQ + 1);
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
    for (int s = 0; s < NP; ++s) mbar_init(p_full(s), 1);
    for (int s = 0; s < NQ; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(conv_bar(s), 4);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));
  if (threadIdx.x == 64) TC_STAMP(1);

  pdl_trigger();
  if (warp == 0) {
    // ============================== TMA producer ==============================
    if (nkb > 0) {   // warp-uniform control flow; one elected lane issues the bulk copies
      struct KbIter { int seg, kb; };
      auto nkb_of = [&](int sg) { return (prm.K[sg] + kBlockK - 1) / kBlockK; };
      auto advance = [&](KbIter& it) { if (++it.kb >= nkb_of(it.seg)) { it.kb = 0; ++it.seg; } };
      KbIter it0{0, kb_begin};
      while (it0.kb >= nkb_of(it0.seg)) { it0.kb -= nkb_of(it0.seg); ++it0.seg; }
      auto load_p = [&](int j, const KbIter& it) {   // K-block j of this CTA
        const int s = j % NP;
        if (elect_one()) {
          mbar_expect_tx(p_full(s), Cfg::kPBytes);
          tma_load_2d(base + s * Cfg::kPBytes, &prm.mapP[it.seg], p_full(s), it.kb * kBlockK, p0);
          KB_STAMP(j, 1);
        }
        __syncwarp();
      };
      auto load_q = [&](int j, const KbIter& it) {
        const int s = j % NQ;
        if (elect_one()) {
          mbar_expect_tx(q_full(s), Cfg::kQBytes);
          tma_load_2d(q_base + s * Cfg::kQSlot, &prm.mapQ[it.seg], q_full(s), it.kb * kBlockK, q0);
          KB_STAMP(j, 0);
        }
        __syncwarp();
      };
      // Fill both (empty) rings; the operand that is a constant weight goes out before the grid
      // dependency resolves, i.e. while the previous kernels of the chain are still running.
      const int np0 = nkb < NP ? nkb : NP, nq0 = nkb < NQ ? nkb : NQ;
      KbIter itp = it0, itq = it0;
      if (prm.pre_p) for (int j = 0; j < np0; ++j) { load_p(j, itp); advance(itp); }
      if (prm.pre_q) for (int j = 0; j < nq0; ++j) { load_q(j, itq); advance(itq); }
      pdl_wait();
      if (!prm.pre_p) for (int j = 0; j < np0; ++j) { load_p(j, itp); advance(itp); }
      if (!prm.pre_q) for (int j = 0; j < nq0; ++j) { load_q(j, itq); advance(itq); }
      // Steady state: the MMA's commit for K-block j - NQ frees Q/TMEM slot j % NQ -- and, a fortiori, the P
      // slot of K-block j - NQ (its conversion preceded that MMA), which K-block j - NQ + NP reuses.
      for (int j = NQ; j < nkb; ++j) {
        mbar_wait(empty_bar(j % NQ), ((uint32_t)(j / NQ) & 1u) ^ 1u);
        load_q(j, itq); advance(itq);
        const int jp = j - NQ + NP;
        if (jp < nkb) { load_p(jp, itp); advance(itp); }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    if (nkb > 0) {
      // instruction descriptor: D=f32, A=B=tf32, both K-major, N, M=128
      const uint32_t idesc = ((1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(QN >> 3) << 17) |
                              ((uint32_t)(kTileP >> 4) << 24)) ^ prm.idesc_xor;
      // the whole warp walks the K-blocks (warp-uniform control flow); one elected lane issues
      for (int i = 0; i < nkb; ++i) {
        const int s = i % NQ;
        mbar_wait(conv_bar(s), (uint32_t)(i / NQ) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          KB_STAMP(i, 5);
          const uint32_t q_hi = q_base + s * Cfg::kQSlot, q_lo = q_hi + Cfg::kQBytes;
          const uint64_t b_hi0 = umma_desc(q_hi, 16u, 1024u), b_lo0 = umma_desc(q_lo, 16u, 1024u);
          const uint32_t ta0 = tmem_base + (uint32_t)QN + (uint32_t)s * 64u;
#pragma unroll
          for (int k = 0; k < kBlockK / 8; ++k) {
            // +32 bytes along K = +2 in the descriptor's (addr >> 4) field; +8 tensor-memory columns
            const uint64_t b_hi = b_hi0 + (uint64_t)(2 * k), b_lo = b_lo0 + (uint64_t)(2 * k);
            const uint32_t ta_hi = ta0 + (uint32_t)k * 8u;
            umma_tf32_ts(tmem_base, ta_hi + 32u, b_hi, idesc, (i > 0 || k > 0) ? 1u : 0u);   // P_lo * Q_hi
            umma_tf32_ts(tmem_base, ta_hi, b_lo, idesc, 1u);                                 // P_hi * Q_lo
            umma_tf32_ts(tmem_base, ta_hi, b_hi, idesc, 1u);                                 // P_hi * Q_hi
          }
          umma_commit(empty_bar(s));   // Q slot + tensor-memory slot reusable once these MMAs retire
          KB_STAMP(i, 6);
        }
        __syncwarp();
      }
      if (elect_one()) umma_commit(accum_bar);
      __syncwarp();
    }
  } else {
    // ============================== converters, then epilogue ==============================
    const int ct = threadIdx.x - 64;   // 0 .. 32*kConvWarps-1
    constexpr int kCT = 32 * kConvWarps;
    const int grp_id = (warp - 2) >> 2;  // converter group of this warp
    const int gt = ct & (kGT - 1);       // thread index within the group
    for (int i = grp_id; i < nkb; i += kConvGroups) {
      const int sq = i % NQ, sp = i % NP;
      auto split = [](float x, float& hi, float& lo) {
        hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
        lo = x - hi;
      };
      // Q: hi written back in place over the TMA'd tile, lo to the sibling tile
      mbar_wait(q_full(sq), (uint32_t)(i / NQ) & 1u);
      if (gt == 0) KB_STAMP(i, 2);
      float4* q_hi = reinterpret_cast<float4*>(gen_base + (q_base - base) + sq * Cfg::kQSlot);
      float4* q_lo = reinterpret_cast<float4*>(gen_base + (q_base - base) + sq * Cfg::kQSlot + Cfg::kQBytes);
#pragma unroll
      for (int j = 0; j < Cfg::kQBytes / 16 / kGT; ++j) {
        const float4 v = q_hi[gt + kGT * j];
        float4 h, l;
        split(v.x, h.x, l.x); split(v.y, h.y, l.y); split(v.z, h.z, l.z); split(v.w, h.w, l.w);
        q_hi[gt + kGT * j] = h;
        q_lo[gt + kGT * j] = l;
      }
      // P (the 128-row operand) goes to tensor memory: this thread owns tile row `prow` = its TMEM lane, reads
      // the row's 32 fp32 out of the 128B-swizzled tile (16-byte chunk c of row r sits at chunk c ^ (r & 7)) and
      // stores hi | lo as 2 x 32 columns.  No shared-memory write-back, and the MMA never reads P from shared
      // memory.
      mbar_wait(p_full(sp), (uint32_t)(i / NP) & 1u);
      if (gt == 0) KB_STAMP(i, 3);
      if (ct == 0 && i == 0) TC_STAMP(2);
      const float4* p_raw = reinterpret_cast<const float4*>(gen_base + sp * Cfg::kPBytes);
      const int prow = (warp & 3) * 32 + lane;
      const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)QN + (uint32_t)sq * 64u;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const int cch = half * 4 + cc;
          const float4 v = p_raw[prow * 8 + (cch ^ (prow & 7))];
          float h, l;
          split(v.x, h, l); hi[cc * 4 + 0] = __float_as_uint(h); lo[cc * 4 + 0] = __float_as_uint(l);
          split(v.y, h, l); hi[cc * 4 + 1] = __float_as_uint(h); lo[cc * 4 + 1] = __float_as_uint(l);
          split(v.z, h, l); hi[cc * 4 + 2] = __float_as_uint(h); lo[cc * 4 + 2] = __float_as_uint(l);
          split(v.w, h, l); hi[cc * 4 + 3] = __float_as_uint(h); lo[cc * 4 + 3] = __float_as_uint(l);
        }
        tmem_st16(ta + (uint32_t)half * 16u, hi);
        tmem_st16(ta + 32u + (uint32_t)half * 16u, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the MMA (async proxy)
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(conv_bar(sq));
      if (gt == 0) KB_STAMP(i, 4);
      if (ct == 0 && i == 0) TC_STAMP(3);
      if (i == nkb - 1 && gt == 0) TC_STAMP(4);
    }
    // ---- epilogue
    if (nkb > 0) {
      mbar_wait(accum_bar, 0);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    }
    pdl_wait();   // C, `add` and c_row_len may be produced by the preceding kernels
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
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
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
long g_tc_trace_stride = 0;      // > 0: launch n of a traced sequence stamps buf + n * stride (u64 units)
int g_tc_trace_left = 0;

bool aligned_ok(const float* p, long ld) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0 && (ld % 4) == 0; }

}  // namespace

// Fills `prm` for one problem; returns false if the problem is not eligible for the tensor-core path
// (caller falls back to the CUDA-core kernel).  `QN` is the Q-tile width chosen for the whole group.
static bool tc_plan(int mode, const GemmProblem& g, int QN, TcParams& prm) {
  if (g.nseg < 1 || g.M <= 0 || g.N <= 0) return false;
  // only the K-major/K-major (NT) form runs on tensor cores; NN/TN work arrives in NT form on transposed
  // copies (editnet.cu backward_core) or falls back to the CUDA-core kernel
  if (mode != kNT) return false;
  if (g.a_inner > 0 || g.a_row_len) return false;         // two-level / masked A rows stay on the CUDA-core path
  long ktot = 0;
  for (int s = 0; s < g.nseg; ++s) {
    if (!aligned_ok(g.seg[s].A, g.seg[s].lda) || !aligned_ok(g.seg[s].B, g.seg[s].ldb)) return false;
    ktot += g.seg[s].K;
  }
  if (ktot < 64) return false;
  memset(&prm, 0, sizeof(prm));
  // skinny M: weights (B side, N rows) take the 128-row P role
  const bool swap = (g.M <= QN && g.N > g.M);
  const int Pr = swap ? g.N : g.M, Qr = swap ? g.M : g.N;
  if (swap && (g.c_inner > 0)) return false;
  prm.swap = swap; prm.Pr = Pr; prm.Qr = Qr;
  prm.nseg = g.nseg;
  for (int s = 0; s < g.nseg; ++s) {
    const GemmSeg& sg = g.seg[s];
    prm.K[s] = sg.K;
    const float* Pp = swap ? sg.B : sg.A; const long Pld = swap ? sg.ldb : sg.lda;
    const float* Qp = swap ? sg.A : sg.B; const long Qld = swap ? sg.lda : sg.ldb;
    if (!make_map(&prm.mapP[s], Pp, Pr, sg.K, Pld, kBlockK, kTileP)) return false;
    if (!make_map(&prm.mapQ[s], Qp, Qr, sg.K, Qld, kBlockK, QN)) return false;
  }
  prm.tiles_p = (Pr + kTileP - 1) / kTileP;
  prm.tiles_q = (Qr + QN - 1) / QN;
  prm.split_k = 1;
  prm.C = g.C; prm.ldc = g.ldc; prm.c_inner = g.c_inner; prm.c_ld_inner = g.c_ld_inner;
  prm.c_row_len = g.c_row_len; prm.c_valid_inner = g.c_valid_inner > 0 ? g.c_valid_inner : 1;
  prm.bias = g.bias; prm.bias2 = g.bias2; prm.add = g.add; prm.ldadd = g.ldadd; prm.add_mod = g.add_mod;
  prm.beta = g.beta; prm.act = g.act;
  { const char* e = getenv("SET_TC_IDESC_XOR"); prm.idesc_xor = e ? (unsigned)strtoul(e, nullptr, 0) : 0u; }
  prm.trace = g_tc_trace;   // (a sequence trace advances per launch, see gemm_tc_try_group)
  prm.pre_p = (g.w_const && swap) ? 1 : 0;
  prm.pre_q = (g.w_const && !swap) ? 1 : 0;
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
  if (g_tc_trace && g_tc_trace_stride > 0) {
    if (g_tc_trace_left > 0) {
      // header: [8] grid size, [9] problems in the group, [10] K-blocks per CTA of problem 0
      for (int k = 0; k < grp.n; ++k) grp.p[k].trace = g_tc_trace;
      g_tc_trace += g_tc_trace_stride;
      --g_tc_trace_left;
    } else {
      for (int k = 0; k < grp.n; ++k) grp.p[k].trace = nullptr;
    }
  }
  auto launch = [&](auto tag) {
    constexpr int G = decltype(tag)::value;
    TcGroup<G> small;
    small.n = grp.n;
    memcpy(small.cta_start, grp.cta_start, sizeof(small.cta_start));
    memcpy(small.p, grp.p, sizeof(TcParams) * grp.n);
    if (QN == 64) return launch_chain(gemm_tc_kernel<64, G>, dim3(cta), dim3(kThreadsTc), TcCfg<64>::kSmemBytes, stream, small);
    return launch_chain(gemm_tc_kernel<128, G>, dim3(cta), dim3(kThreadsTc), TcCfg<128>::kSmemBytes, stream, small);
  };
  cudaError_t lerr;
  if (grp.n == 1) lerr = launch(std::integral_constant<int, 1>{});
  else if (grp.n == 2) lerr = launch(std::integral_constant<int, 2>{});
  else if (grp.n <= 5) lerr = launch(std::integral_constant<int, 5>{});
  else lerr = launch(std::integral_constant<int, 8>{});
  SET_CHECK_CUDA(lerr);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  return SET_OK;
}

// debugging: device buffer of >= 16 u64 that CTA 0 of every following tensor-core launch stamps
void gemm_tc_set_trace(unsigned long long* buf) { g_tc_trace = buf; g_tc_trace_stride = 0; g_tc_trace_left = 0; }
void gemm_tc_set_trace_seq(unsigned long long* buf, long stride, int launches) {
  g_tc_trace = buf; g_tc_trace_stride = stride; g_tc_trace_left = launches;
}

}  // namespace set
