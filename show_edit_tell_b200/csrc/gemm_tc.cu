// fp32-accurate tensor-core GEMM for sm_100a: tcgen05.mma kind::tf32 with the 3xTF32 split,
// operands staged by TMA (128B-swizzled tiles), accumulator in TMEM.
//
//   D[p, q] = sum_k P[p, k] * Q[q, k]           P tile = UMMA "A" (128 rows), Q tile = UMMA "B" (QN rows)
//
// Every fp32 operand x is used as hi = tf32(x) -- the raw word: kind::tf32 ignores the low 13 mantissa bits -- and
// lo = x - (x with those bits cleared) (exact in fp32), and each K-step issues
// D += P_lo*Q_hi ;  D += P_hi*Q_lo ;  D += P_hi*Q_hi  -- the dropped lo*lo term is 2^-22 relative, so results match
// fp32 FMA accumulation to ~1e-6 (one-pass TF32 misses the 1e-4 parity budget by 10x, 3xTF32
// meets it).
//
// Both operands are K-major (global rows = tile rows, reduction contiguous): x@W^T (NT) reads the row-major
// buffers in place; callers present dX / dW work in NT form on transposed copies (MN-major tf32 operands
// read back as zeros with the sm_100a descriptors tried in round 1; that experiment is not kept in the tree).
//
// "swap" mode puts the weight matrix on the 128-row P side and the (<=64..128 row) activation batch
// on the Q side: that is how the skinny per-step GEMMs (M = batch) fill the tensor core's M=128
// datapath, with split-K spreading one weight matrix over all 148 SMs.
//
// The 128-row operand never makes a second trip through shared memory: converter warps read the TMA'd
// fp32 tile once, split it in registers and write hi | lo into TENSOR memory (tcgen05.st), from where the
// MMA takes its A operand; only the small Q tile is split in place in shared memory.
//
// Warp roles: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer (both walk the K-blocks warp-uniformly and
// issue from one elect.sync lane), warps 2.. = hi/lo converters (groups of 4 warps alternate K-blocks) during the main
// loop, then the epilogue: TMEM -> registers -> shared, then one of three finishes -- plain stores / red.global.add
// (split-K into pre-zeroed or accumulating C), the scratch-slab cooperative finish, or the thread-block-cluster finish
// through distributed shared memory -- the last two optionally applying a fused LSTM-family cell (GemmEpi).
// Launches carry the programmatic-dependent-launch attribute: constant weight tiles are requested before
// griddepcontrol.wait (common.cuh).
#include <cuda.h>
#include <stdlib.h>
#include <string.h>

#include <mutex>
#include <type_traits>

#include "gemm.cuh"
#include "tc_common.cuh"
#include "step_kernel.cuh"

namespace set {

namespace {

constexpr int kBlockK = 32;        // fp32 per smem row: 128 B = one swizzle span
constexpr int kTileP = 128;        // UMMA M
// 1: the "hi" operand is the raw fp32 word -- kind::tf32 reads only sign, exponent and the top 10 mantissa bits
// of each 32-bit container, so masking the low 13 bits first changes nothing and the write-back of Q_hi is saved.
// (Verified on B200 by the parity tests: a rounding datapath would show as ~1e-3 errors.)
#ifndef SET_TC_RAW_HI
#define SET_TC_RAW_HI 1
#endif
#ifndef SET_TC_CONV_WARPS
#define SET_TC_CONV_WARPS 8
#endif
constexpr int kConvWarps = SET_TC_CONV_WARPS;   // hi/lo converters, also the epilogue warps
constexpr int kConvGroups = kConvWarps / 4;     // a group = 4 warps = the 4 TMEM lane quarters; K-block i belongs to
                                                // group i % kConvGroups, so the groups' latency chains overlap
constexpr int kGT = 128;                        // threads per converter group
constexpr int kMaxFusedSplit = 8;               // split-K ways the fused epilogue reduces
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
  int fuse_ok;                     // (host) the problem may take the fused epilogue
  int fused;                       // swap mode: split-K partials meet in a scratch slab; the tile's last CTA reduces,
                                   // adds bias/add/C and applies `epi`
  int nblk, blk_stride;            // P tile = nblk blocks of 128/nblk rows; block j starts at global row
                                   // j * blk_stride + tile * (128 / nblk)   (nblk = 4: the four gates of 32 units)
  float* scratch; int* counters;   // library-owned split-K scratch: one [QN][128] slab and two counters per CTA
  const float* zero16;             // 16 bytes of zeros in global memory (stand-in for absent epilogue operands)
  int cluster;                     // > 1: the `split_k` CTAs of a tile form a thread-block cluster; partials stay in shared
                                   // memory and meet through DSMEM (no scratch slab, no atomics, no global fences)
  int coop;                        // 1: every CTA of a tile takes a share of the finish (needs a one-wave grid)
  GemmEpi epi;
  int pre_p, pre_q;                // operand is a constant weight: its first pipeline stages load before pdl_wait()
  unsigned idesc_xor;              // debugging aid (SET_TC_IDESC_XOR)
  unsigned long long* trace;       // debugging aid: per-phase %globaltimer stamps of CTA 0 (SET_TC_TRACE)
};

#define TC_STAMP(slot)                                                         \
  do {                                                                         \
    if (prm.trace && blockIdx.x == 0) prm.trace[slot] = gtimer();              \
  } while (0)

// per-K-block SM-clock stamps of CTA 0 (slots: 0 producer issued Q, 1 producer issued P, 2 converter saw Q,
// 3 converter saw P, 4 converter done, 5 MMA saw converted block, 6 MMA issued)
#define KB_STAMP(i, slot)                                                                       \
  do {                                                                                          \
    if (prm.trace && blockIdx.x == 0 && (i) < 48) prm.trace[2100 + 8 * (i) + (slot)] = clock64(); \
  } while (0)

// TWIN = 1: a shallower pipeline sized so that TWO CTAs share an SM (half the shared memory, 256 tensor-memory
// columns each).  Used for the big time-batched GEMMs (many tiles per SM): one CTA's prologue / epilogue / MMA
// drain overlaps the other's main loop, and twice as many converter warps feed the tensor core.
template <int QN, int TWIN = 0>
struct TcCfg {
  // Two rings.  P: raw fp32 tiles of the 128-row operand; a K-block of P is converted straight into tensor
  // memory, so the ring only has to cover the HBM latency -- it is the deep one (Little: 148 SMs x kNP x 16 KB
  // in flight ~ 19 MB >= 6.5 TB/s x 2 us).  Q: hi | lo tiles the MMA reads from shared memory, paired with the
  // tensor-memory slots of P_hi | P_lo; both are released by the MMA's commit.
#ifndef SET_TC_NQ64
#define SET_TC_NQ64 6
#define SET_TC_NP64 6
#endif
  static constexpr int kNQ = TWIN ? 2 : ((QN <= 64) ? SET_TC_NQ64 : 4);
  static constexpr int kNP = TWIN ? 3 : ((QN <= 64) ? SET_TC_NP64 : 4);
  // Converter groups take alternate K-blocks and wait on q_full by PARITY, which only tells a phase from its
  // predecessor: a group must therefore visit every slot on consecutive phases, i.e. kNQ % groups == 0.  (With 3 Q
  // slots and 2 groups a group came back to a slot two phases later; TMA completes out of order, so the phase in
  // between could still be pending and the wait returned at once on stale data -- a 1-in-3000 corruption of whole
  // output tiles, found with tools/stress_repeat.py.)  The P ring needs no such rule: a group reaches the P wait of
  // K-block i only after Q(i) landed, which the producer issues after the MMA of K-block i - kNQ, i.e. after every
  // earlier P tile was consumed.
  static_assert(kNQ % kConvGroups == 0, "a converter group must revisit a Q slot on consecutive barrier phases");
  static_assert(kNP >= kNQ, "the producer refills P slot (j - kNQ + kNP) % kNP behind the MMA of K-block j - kNQ");
  static constexpr uint32_t kTmemCols = TWIN ? 256u : 512u;
  static_assert(!TWIN || QN == 128, "the twin configuration is built for 128-wide Q tiles");
  static constexpr int kPBytes = kTileP * 128;
  static constexpr int kQBytes = QN * 128;
  static constexpr int kQSlot = 2 * kQBytes;
  static constexpr int kRingBytes = kNP * kPBytes + kNQ * kQSlot;
  // (the twin configuration has no room for alignment slack: the kernel keeps no static shared memory, so the
  // dynamic window starts 1024-aligned; the kernel traps if that ever stops being true)
  static constexpr int kSmemBytes = kRingBytes + (TWIN ? 0 : 1024) /*align*/ + 256 /*barriers*/;
};

template <int G>
struct TcGroup {
  int n;
  int cta_start[9];        // first CTA of each problem (tiles * split_k each)
  TcParams p[G];
};

template <int QN, int G, int TWIN = 0>
__global__ void __launch_bounds__(kThreadsTc, TWIN ? 2 : 1) gemm_tc_kernel(const __grid_constant__ TcGroup<G> grp) {
  using Cfg = TcCfg<QN, TWIN>;
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
  if (TWIN && base != smem_u32(smem_dyn)) __trap();
  volatile int* s_last_p = reinterpret_cast<volatile int*>(gen_base + (bar_base - base) + 8 * (NP + 3 * NQ + 2));

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
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(Cfg::kTmemCols)
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
          if (prm.nblk == 4) {
#pragma unroll
            for (int g = 0; g < 4; ++g)
              tma_load_2d(base + s * Cfg::kPBytes + g * 4096, &prm.mapP[it.seg], p_full(s), it.kb * kBlockK,
                          g * prm.blk_stride + pt * 32);
          } else {
            tma_load_2d(base + s * Cfg::kPBytes, &prm.mapP[it.seg], p_full(s), it.kb * kBlockK, p0);
          }
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
      // (A parity wait only tells phase n from phase n-1, so a waiter must see EVERY phase of a barrier.  The groups
      // alternate K-blocks: kNQ is a multiple of the group count, hence a group revisits a Q slot on consecutive
      // phases -- see the static_assert in TcCfg.  The P wait below is gated by this one.)
      mbar_wait(q_full(sq), (uint32_t)(i / NQ) & 1u);
      if (gt == 0) KB_STAMP(i, 2);
      float4* q_hi = reinterpret_cast<float4*>(gen_base + (q_base - base) + sq * Cfg::kQSlot);
      float4* q_lo = reinterpret_cast<float4*>(gen_base + (q_base - base) + sq * Cfg::kQSlot + Cfg::kQBytes);
#pragma unroll
      for (int j = 0; j < Cfg::kQBytes / 16 / kGT; ++j) {
        const float4 v = q_hi[gt + kGT * j];
        float4 h, l;
        split(v.x, h.x, l.x); split(v.y, h.y, l.y); split(v.z, h.z, l.z); split(v.w, h.w, l.w);
#if !SET_TC_RAW_HI
        q_hi[gt + kGT * j] = h;
#endif
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
          split(v.x, h, l); hi[cc * 4 + 0] = __float_as_uint(SET_TC_RAW_HI ? v.x : h); lo[cc * 4 + 0] = __float_as_uint(l);
          split(v.y, h, l); hi[cc * 4 + 1] = __float_as_uint(SET_TC_RAW_HI ? v.y : h); lo[cc * 4 + 1] = __float_as_uint(l);
          split(v.z, h, l); hi[cc * 4 + 2] = __float_as_uint(SET_TC_RAW_HI ? v.z : h); lo[cc * 4 + 2] = __float_as_uint(l);
          split(v.w, h, l); hi[cc * 4 + 3] = __float_as_uint(SET_TC_RAW_HI ? v.w : h); lo[cc * 4 + 3] = __float_as_uint(l);
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
    const int stage_grp = (warp - 2) >> 2;     // the warps sharing a quarter split the column blocks
    constexpr int EPW_N = QN + 4;              // staged row pitch, non-swap: [128 p][QN q]
    constexpr int EPW_S = kTileP + 4;          // staged row pitch, swap:     [QN q][128 p]
    float* ep = reinterpret_cast<float*>(gen_base);
    const int prow = quarter * 32 + lane;      // tile row held by this thread
#pragma unroll 1
    for (int cb = stage_grp; cb < QN / 32; cb += kConvGroups) {
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
    if (prm.fused) {
      // ---------------- fused epilogue (swap mode: staged tile = ep[q][p], q batch rows, p weight rows)
      // Split-K partials go to per-CTA scratch slabs (plain vector stores; no atomics on C, no pre-zeroed C).
      // Cooperative finish (grid fits one wave, so all partners are resident): every CTA of the tile waits for
      // its partners and finishes 1/split of the tile's items; otherwise the last CTA to arrive finishes all.
      // An item gathers every partial + bias/add/C operand it needs with independent 128-bit loads issued
      // together (one L2 round trip), applies the cell, and stores.  Summation order is fixed (split 0, 1, ..).
      constexpr int kSlab = QN * kTileP;
      const int rows = prm.Qr < QN ? prm.Qr : QN;
      const int split = prm.split_k;
      int nfin = 1, fin = 0;
      bool finisher = true;
      int* cnt = prm.counters + 2 * (blockIdx.x - ks);
      const bool dsm = prm.cluster > 1;
      uint32_t peer[kMaxFusedSplit];   // the partners' staged tiles in distributed shared memory
      if (dsm) {
        // every thread of every CTA of the cluster (the idle producer / MMA warps included, see below) meets here
        // once the partial tiles are staged; CTA `ks` then finishes 1/split of the tile out of its partners' shared memory
        cluster_sync_all();
#pragma unroll
        for (int k2 = 0; k2 < kMaxFusedSplit; ++k2) peer[k2] = dsmem_addr(smem_u32(ep), (uint32_t)(k2 < split ? k2 : 0));
        nfin = split; fin = ks;
        if (ct == 0) TC_STAMP(10);
      } else if (split > 1) {
        float* part = prm.scratch + (size_t)blockIdx.x * kSlab;
        for (int e = ct; e < rows * 32; e += kCT) {
          const int q = e >> 5, p4 = (e & 31) * 4;
          __stcg(reinterpret_cast<float4*>(part + q * kTileP + p4), *reinterpret_cast<const float4*>(&ep[q * EPW_S + p4]));
        }
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kConvWarps) : "memory");
        if (ct == 0) TC_STAMP(9);
        if (ct == 0) {
          const int old = atomicAdd(cnt, 1);
          if (prm.coop) {
            int seen = old + 1;
            while (seen < split) {
              __nanosleep(40);
              asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(cnt) : "memory");
            }
            *s_last_p = 1;
          } else {
            const int last = (old == split - 1) ? 1 : 0;
            if (last) *cnt = 0;   // every partner has arrived; the next launch finds the counter at zero
            *s_last_p = last;
          }
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kConvWarps) : "memory");
        finisher = (*s_last_p != 0);
        if (prm.coop) { nfin = split; fin = ks; }
        if (ct == 0) TC_STAMP(10);
        __threadfence();
        if (ct == 0) TC_STAMP(11);
      }
      if (finisher) {
        const GemmEpi& eo = prm.epi;
        const float* pbase = prm.scratch + (size_t)(blockIdx.x - ks) * kSlab;
        auto add4 = [](float4& a, const float4& b) { a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w; };
        // Absent operands read a 16-byte zero block instead of branching: an item's loads then form one
        // straight-line batch the compiler issues back to back (one L2 round trip, not one per operand).
        const float* zero4 = prm.zero16;
        const int addq_mod = prm.add_mod > 0 ? prm.add_mod : (1 << 30);
        struct Pre { float4 v[kMaxFusedSplit]; float4 b1, b2, ad, old; };
        // request the pre-activation of 4 consecutive tile rows p4.. of batch row q (global column n..n+3)
        auto request = [&](Pre& r, int q, int p4, int n) {
          if (dsm) {
            const uint32_t off = (uint32_t)(q * EPW_S + p4) * 4u;
#pragma unroll
            for (int k2 = 0; k2 < kMaxFusedSplit; ++k2) r.v[k2] = dsmem_ld4(peer[k2] + off, k2 < split);
          } else {
            const float* src = pbase + q * kTileP + p4;
#pragma unroll
            for (int k2 = 0; k2 < kMaxFusedSplit; ++k2)
              r.v[k2] = __ldcg(reinterpret_cast<const float4*>(k2 < split ? src + (size_t)k2 * kSlab : zero4));
          }
          r.b1 = __ldg(reinterpret_cast<const float4*>(prm.bias ? prm.bias + n : zero4));
          r.b2 = __ldg(reinterpret_cast<const float4*>(prm.bias2 ? prm.bias2 + n : zero4));
          r.ad = *reinterpret_cast<const float4*>(prm.add ? prm.add + (long)(q % addq_mod) * prm.ldadd + n : zero4);
          r.old = *reinterpret_cast<const float4*>(prm.beta ? prm.C + (long)q * prm.ldc + n : zero4);
        };
        auto resolve = [&](const Pre& r, int q, int p4) -> float4 {
          float4 acc;
          if (split > 1) {
            acc = r.v[0];
#pragma unroll
            for (int k2 = 1; k2 < kMaxFusedSplit; ++k2) add4(acc, r.v[k2]);   // absent splits contribute exact zeros
          } else {
            acc = *reinterpret_cast<const float4*>(&ep[q * EPW_S + p4]);
          }
          add4(acc, r.b1); add4(acc, r.b2); add4(acc, r.ad); add4(acc, r.old);
          return acc;
        };
        if (eo.op != kEpiLstm && eo.op != kEpiCopy1) {
          const int items = rows * 32;   // (q, 4 consecutive rows of the tile)
          const int i0 = (int)((long)items * fin / nfin), i1 = (int)((long)items * (fin + 1) / nfin);
          const int D = eo.D;
          for (int it = i0 + ct; it < i1; it += kCT) {
            const int q = it >> 5, p4 = (it & 31) * 4;
            const int n = pt * kTileP + p4;
            if (n >= prm.Pr) continue;
            Pre pr;
            request(pr, q, p4, n);
            const long x = (long)q * D + n;
            float4 sel, cn, og;
            if (eo.op == kEpiCopy2) {
              sel = *reinterpret_cast<const float4*>(eo.sel + x);
              cn = *reinterpret_cast<const float4*>(eo.cnew + x);
              og = *reinterpret_cast<const float4*>(eo.gates + (long)q * eo.ld_gates + 3 * D + n);
            }
            // reverse-pass cells: their own operands join the same batch of loads
            float4 r0, r1, r2, r3, r4, r5, r6;
            const bool cg = (eo.op == kEpiCtxGateBwd) && n >= eo.col0 && n < eo.col0 + D;   // CTA-uniform (tile-aligned range)
            if (eo.op == kEpiLstmBwd) {
              const float* gt4 = eo.gates + (long)q * eo.ld_gates + n;
              r0 = *reinterpret_cast<const float4*>(gt4); r1 = *reinterpret_cast<const float4*>(gt4 + D);
              r2 = *reinterpret_cast<const float4*>(gt4 + 2 * D); r3 = *reinterpret_cast<const float4*>(gt4 + 3 * D);
              r4 = *reinterpret_cast<const float4*>(eo.c_prev + x);
              r5 = *reinterpret_cast<const float4*>(eo.x1 + x);
              r6 = *reinterpret_cast<const float4*>(eo.y0 + x);
            } else if (cg) {
              const float* z4 = eo.x0 + (long)q * 3 * D + (n - eo.col0);
              r0 = *reinterpret_cast<const float4*>(z4); r1 = *reinterpret_cast<const float4*>(z4 + D);
              r2 = *reinterpret_cast<const float4*>(z4 + 2 * D);
            }
            else if (eo.op == kEpiCopy1Bwd) {
              const float* gt4 = eo.gates + (long)q * eo.ld_gates + n;
              r0 = *reinterpret_cast<const float4*>(gt4); r1 = *reinterpret_cast<const float4*>(gt4 + D);
              r2 = *reinterpret_cast<const float4*>(gt4 + 2 * D);
              r4 = *reinterpret_cast<const float4*>(eo.c_prev + x);
            } else if (eo.op == kEpiCopy2Bwd) {
              r0 = *reinterpret_cast<const float4*>(eo.gates + (long)q * eo.ld_gates + 3 * D + n);   // o gate
              r1 = *reinterpret_cast<const float4*>(eo.x1 + x);      // c2
              r2 = *reinterpret_cast<const float4*>(eo.kgate + x);
              r3 = *reinterpret_cast<const float4*>(eo.sel + x);
              r4 = *reinterpret_cast<const float4*>(eo.cnew + x);
              r5 = *reinterpret_cast<const float4*>(eo.y0 + x);      // d c2 carry
              r6 = *reinterpret_cast<const float4*>(eo.x0 ? eo.x0 + x : zero4);   // d dropout(h2), raw
            }
            float4 dhb = (eo.op == kEpiLstmBwd) ? *reinterpret_cast<const float4*>(eo.x0 ? eo.x0 + x : zero4)
                                                : make_float4(0.f, 0.f, 0.f, 0.f);
            // length-masked variant (encoder BPTT, enc_lstm_bwd_kernel in cells.cu): extra d h / d c from the sequence
            // outputs at position t, d h_last where the row ends at t; rows that ended earlier produce zero gates
            bool enc_inactive = false;
            float4 dcx = make_float4(0.f, 0.f, 0.f, 0.f);
            if (eo.op == kEpiLstmBwd && eo.len) {
              const long long Lq = eo.len[q];
              enc_inactive = (Lq <= (long long)eo.t);
              const long so = (long)q * eo.seq_ld + (long)eo.t * D + n;
              add4(dhb, *reinterpret_cast<const float4*>(eo.seq_h + so));
              dcx = *reinterpret_cast<const float4*>(eo.seq_m + so);
              add4(dhb, *reinterpret_cast<const float4*>((Lq - 1 == (long long)eo.t) ? eo.h_prev + x : zero4));
            }
            const float4 v = resolve(pr, q, p4);
            if (eo.op == kEpiNone) {
              *reinterpret_cast<float4*>(prm.C + (long)q * prm.ldc + n) = v;
            } else if (eo.op == kEpiLstmBwd) {
              // lstm_bwd_core (cells.cu) on 4 units: gates r0..r3 = i,f,g,o; r4 = c_prev, r5 = c_cur, r6 = d c carry
              float4 di, df, dg, dgo, dcp;
              auto cell = [](float gi, float gf, float gg, float go, float cp, float cc, float dh, float dcin, float& a, float& b,
                             float& c, float& d, float& e) {
                const float tc = tanhf(cc);
                const float dc = dcin + dh * go * (1.f - tc * tc);
                a = dc * gg * gi * (1.f - gi);
                b = dc * cp * gf * (1.f - gf);
                c = dc * gi * (1.f - gg * gg);
                d = dh * tc * go * (1.f - go);
                e = dc * gf;
              };
              cell(r0.x, r1.x, r2.x, r3.x, r4.x, r5.x, v.x + dhb.x, r6.x + dcx.x, di.x, df.x, dg.x, dgo.x, dcp.x);
              cell(r0.y, r1.y, r2.y, r3.y, r4.y, r5.y, v.y + dhb.y, r6.y + dcx.y, di.y, df.y, dg.y, dgo.y, dcp.y);
              cell(r0.z, r1.z, r2.z, r3.z, r4.z, r5.z, v.z + dhb.z, r6.z + dcx.z, di.z, df.z, dg.z, dgo.z, dcp.z);
              cell(r0.w, r1.w, r2.w, r3.w, r4.w, r5.w, v.w + dhb.w, r6.w + dcx.w, di.w, df.w, dg.w, dgo.w, dcp.w);
              float* dgt = eo.y1 + (long)q * 4 * D + n;
              if (enc_inactive) {
                di = df = dg = dgo = make_float4(0.f, 0.f, 0.f, 0.f);
              } else {
                *reinterpret_cast<float4*>(eo.y0 + x) = dcp;
              }
              *reinterpret_cast<float4*>(dgt) = di; *reinterpret_cast<float4*>(dgt + D) = df;
              *reinterpret_cast<float4*>(dgt + 2 * D) = dg; *reinterpret_cast<float4*>(dgt + 3 * D) = dgo;
            } else if (eo.op == kEpiCopy1Bwd) {
              // copy1_bwd_kernel (cells.cu): v = d c_new; r0,r1,r2 = i,f,g; r4 = c2_prev
              float4 di, df, dg, dcp;
              di.x = v.x * r2.x * r0.x * (1.f - r0.x); di.y = v.y * r2.y * r0.y * (1.f - r0.y);
              di.z = v.z * r2.z * r0.z * (1.f - r0.z); di.w = v.w * r2.w * r0.w * (1.f - r0.w);
              df.x = v.x * r4.x * r1.x * (1.f - r1.x); df.y = v.y * r4.y * r1.y * (1.f - r1.y);
              df.z = v.z * r4.z * r1.z * (1.f - r1.z); df.w = v.w * r4.w * r1.w * (1.f - r1.w);
              dg.x = v.x * r0.x * (1.f - r2.x * r2.x); dg.y = v.y * r0.y * (1.f - r2.y * r2.y);
              dg.z = v.z * r0.z * (1.f - r2.z * r2.z); dg.w = v.w * r0.w * (1.f - r2.w * r2.w);
              dcp.x = v.x * r1.x; dcp.y = v.y * r1.y; dcp.z = v.z * r1.z; dcp.w = v.w * r1.w;
              float* dgt = eo.y1 + (long)q * 4 * D + n;
              *reinterpret_cast<float4*>(dgt) = di; *reinterpret_cast<float4*>(dgt + D) = df;
              *reinterpret_cast<float4*>(dgt + 2 * D) = dg;
              *reinterpret_cast<float4*>(eo.y0 + x) = dcp;
            } else if (eo.op == kEpiCopy2Bwd) {
              // copy2_bwd_kernel (cells.cu): v = carried d h2; r0 = o, r1 = c2, r2 = k, r3 = sel, r4 = c_new, r5 = d c2 carry
              float4 dfc = r6;
              if (eo.train && eo.x0) {
                const uint32_t keep = drop_keep4(eo.seed, kSiteFc, (uint64_t)(eo.drop_base + x));
                dfc.x = (keep & 1u) ? dfc.x * 2.f : 0.f; dfc.y = (keep & 2u) ? dfc.y * 2.f : 0.f;
                dfc.z = (keep & 4u) ? dfc.z * 2.f : 0.f; dfc.w = (keep & 8u) ? dfc.w * 2.f : 0.f;
              }
              float4 dgo, dk, dsl, dcn;
              auto cell = [](float dh, float go, float c2v, float k, float sl, float cn, float dcin, float& a, float& b, float& c,
                             float& d) {
                const float tc = tanhf(c2v);
                a = dh * tc * go * (1.f - go);
                const float dc = dcin + dh * go * (1.f - tc * tc);
                b = dc * (sl - cn) * k * (1.f - k);
                c = dc * k;
                d = dc * (1.f - k);
              };
              cell(v.x + dfc.x, r0.x, r1.x, r2.x, r3.x, r4.x, r5.x, dgo.x, dk.x, dsl.x, dcn.x);
              cell(v.y + dfc.y, r0.y, r1.y, r2.y, r3.y, r4.y, r5.y, dgo.y, dk.y, dsl.y, dcn.y);
              cell(v.z + dfc.z, r0.z, r1.z, r2.z, r3.z, r4.z, r5.z, dgo.z, dk.z, dsl.z, dcn.z);
              cell(v.w + dfc.w, r0.w, r1.w, r2.w, r3.w, r4.w, r5.w, dgo.w, dk.w, dsl.w, dcn.w);
              *reinterpret_cast<float4*>(eo.y1 + (long)q * 4 * D + 3 * D + n) = dgo;
              *reinterpret_cast<float4*>(eo.y2 + x) = dk;
              *reinterpret_cast<float4*>(eo.x2 + x) = dsl;
              *reinterpret_cast<float4*>(eo.x3 + x) = dcn;
            } else if (eo.op == kEpiCtxGateBwd) {
              *reinterpret_cast<float4*>(prm.C + (long)q * prm.ldc + n) = v;
              if (cg) {
                // ctx_gate_bwd_kernel (cells.cu): r0 = z, r1 = tanh(sc), r2 = tanh(tc); v = d att_cap
                const int d0 = n - eo.col0;
                float4 dz, dsc, dtc;
                dz.x = v.x * (r1.x - r2.x) * r0.x * (1.f - r0.x); dz.y = v.y * (r1.y - r2.y) * r0.y * (1.f - r0.y);
                dz.z = v.z * (r1.z - r2.z) * r0.z * (1.f - r0.z); dz.w = v.w * (r1.w - r2.w) * r0.w * (1.f - r0.w);
                dsc.x = v.x * r0.x * (1.f - r1.x * r1.x); dsc.y = v.y * r0.y * (1.f - r1.y * r1.y);
                dsc.z = v.z * r0.z * (1.f - r1.z * r1.z); dsc.w = v.w * r0.w * (1.f - r1.w * r1.w);
                dtc.x = v.x * (1.f - r0.x) * (1.f - r2.x * r2.x); dtc.y = v.y * (1.f - r0.y) * (1.f - r2.y * r2.y);
                dtc.z = v.z * (1.f - r0.z) * (1.f - r2.z * r2.z); dtc.w = v.w * (1.f - r0.w) * (1.f - r2.w * r2.w);
                *reinterpret_cast<float4*>(eo.y0 + (long)q * eo.ldy + d0) = dz;
                *reinterpret_cast<float4*>(eo.y1 + (long)q * eo.ldy + d0) = dtc;
                *reinterpret_cast<float4*>(eo.y2 + (long)q * D + d0) = dsc;
              }
            } else {
              // copy gate (editnet.py:281-285): k = sigmoid(pre); c2 = k sel + (1-k) c_new; h2 = o tanh(c2)
              float4 k, c, h;
              k.x = sigmoidf_(v.x); k.y = sigmoidf_(v.y); k.z = sigmoidf_(v.z); k.w = sigmoidf_(v.w);
              c.x = k.x * sel.x + (1.f - k.x) * cn.x; c.y = k.y * sel.y + (1.f - k.y) * cn.y;
              c.z = k.z * sel.z + (1.f - k.z) * cn.z; c.w = k.w * sel.w + (1.f - k.w) * cn.w;
              h.x = og.x * tanhf(c.x); h.y = og.y * tanhf(c.y); h.z = og.z * tanhf(c.z); h.w = og.w * tanhf(c.w);
              *reinterpret_cast<float4*>(eo.kgate + x) = k;
              *reinterpret_cast<float4*>(eo.c_out + x) = c;
              *reinterpret_cast<float4*>(eo.h_out + (long)q * eo.ld_h + n) = h;
              if (eo.h2drop) {
                float4 hd = h;
                if (eo.train) {
                  const uint32_t keep = drop_keep4(eo.seed, kSiteFc, (uint64_t)(eo.drop_base + x));
                  hd.x = (keep & 1u) ? h.x * 2.f : 0.f; hd.y = (keep & 2u) ? h.y * 2.f : 0.f;
                  hd.z = (keep & 4u) ? h.z * 2.f : 0.f; hd.w = (keep & 8u) ? h.w * 2.f : 0.f;
                }
                *reinterpret_cast<float4*>(eo.h2drop + x) = hd;
              }
            }
          }
        } else {
          // 4-gate cells; tile rows: [i | f | g | o] x 32 units (editnet.py:233-244 LSTMCellC / nn.LSTMCell;
          // :272-279 copy-LSTM stage 1).  Item = (q, 2 consecutive units): with the usual 4-way split that is one
          // item per epilogue thread -- the cell's ~20 transcendentals are latency-bound per thread, so the work is
          // spread as thin as the thread count allows.
          const int items = rows * 16;
          const int i0 = (int)((long)items * fin / nfin), i1 = (int)((long)items * (fin + 1) / nfin);
          const int D = eo.D;
          auto ld2 = [](const float* ptr) { return __ldcg(reinterpret_cast<const float2*>(ptr)); };
          for (int it = i0 + ct; it < i1; it += kCT) {
            const int q = it >> 4, u0 = (it & 15) * 2;
            const int unit = pt * 32 + u0;
            if (unit >= D) continue;
            float2 v[4][kMaxFusedSplit], b1[4], b2[4], ad[4], old[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              const int n = g * D + unit;
              if (dsm) {
                const uint32_t off = (uint32_t)(q * EPW_S + g * 32 + u0) * 4u;
#pragma unroll
                for (int k2 = 0; k2 < kMaxFusedSplit; ++k2) v[g][k2] = dsmem_ld2(peer[k2] + off, k2 < split);
              } else {
                const float* src = pbase + q * kTileP + g * 32 + u0;
#pragma unroll
                for (int k2 = 0; k2 < kMaxFusedSplit; ++k2) v[g][k2] = ld2(k2 < split ? src + (size_t)k2 * kSlab : zero4);
              }
              b1[g] = ld2(prm.bias ? prm.bias + n : zero4);
              b2[g] = ld2(prm.bias2 ? prm.bias2 + n : zero4);
              ad[g] = ld2(prm.add ? prm.add + (long)(q % addq_mod) * prm.ldadd + n : zero4);
              old[g] = ld2(prm.beta ? prm.C + (long)q * prm.ldc + n : zero4);
            }
            const float2 cp = ld2(eo.c_prev + (long)q * D + unit);
            float2 pre[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float2 acc;
              if (split > 1) {
                acc = v[g][0];
#pragma unroll
                for (int k2 = 1; k2 < kMaxFusedSplit; ++k2) { acc.x += v[g][k2].x; acc.y += v[g][k2].y; }
              } else {
                acc = *reinterpret_cast<const float2*>(&ep[q * EPW_S + g * 32 + u0]);
              }
              acc.x += b1[g].x; acc.y += b1[g].y; acc.x += b2[g].x; acc.y += b2[g].y;
              acc.x += ad[g].x; acc.y += ad[g].y; acc.x += old[g].x; acc.y += old[g].y;
              pre[g] = acc;
            }
            float2 gi, gf, gg, go, c;
            gi.x = sigmoidf_(pre[0].x); gi.y = sigmoidf_(pre[0].y);
            gf.x = sigmoidf_(pre[1].x); gf.y = sigmoidf_(pre[1].y);
            gg.x = tanhf(pre[2].x); gg.y = tanhf(pre[2].y);
            go.x = sigmoidf_(pre[3].x); go.y = sigmoidf_(pre[3].y);
            c.x = gf.x * cp.x + gi.x * gg.x; c.y = gf.y * cp.y + gi.y * gg.y;
            const bool active = (eo.len == nullptr) || (eo.len[q] > (long long)eo.t);
            if (!active) {   // carried through unchanged; zero gates (lstm_fwd_kernel, cells.cu)
              gi = gf = gg = go = make_float2(0.f, 0.f);
              c = cp;
            }
            float* g = eo.gates + (long)q * eo.ld_gates + unit;
            *reinterpret_cast<float2*>(g) = gi; *reinterpret_cast<float2*>(g + D) = gf;
            *reinterpret_cast<float2*>(g + 2 * D) = gg; *reinterpret_cast<float2*>(g + 3 * D) = go;
            *reinterpret_cast<float2*>(eo.c_out + (long)q * D + unit) = c;
            if (eo.op == kEpiLstm) {
              float2 h;
              if (active) { h.x = go.x * tanhf(c.x); h.y = go.y * tanhf(c.y); }
              else h = ld2(eo.h_prev + (long)q * D + unit);
              *reinterpret_cast<float2*>(eo.h_out + (long)q * eo.ld_h + unit) = h;
              if (eo.seq_h) {
                const long so = (long)q * eo.seq_ld + (long)eo.t * D + unit;
                *reinterpret_cast<float2*>(eo.seq_h + so) = active ? h : make_float2(0.f, 0.f);
                *reinterpret_cast<float2*>(eo.seq_m + so) = active ? c : make_float2(0.f, 0.f);
              }
            }
          }
        }
      }
      if (ct == 0) TC_STAMP(12);
      if (dsm) cluster_sync_all();   // no CTA leaves (and frees its shared memory) while a partner may still read it
      if (!dsm && split > 1 && prm.coop) {
        // the last CTA to finish reading the slabs re-arms both counters for the next launch
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kConvWarps) : "memory");
        if (ct == 0) {
          const int old = atomicAdd(cnt + 1, 1);
          if (old == split - 1) { cnt[0] = 0; cnt[1] = 0; }
        }
      }
    } else {
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
    }   // !fused
  }
  if (warp < 2 && prm.fused && prm.cluster > 1) {
    cluster_sync_all();   // partials staged
    cluster_sync_all();   // partners done reading
  }
  if (threadIdx.x == 64) TC_STAMP(7);
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 64) TC_STAMP(8);
  if (prm.trace && threadIdx.x == 0 && blockIdx.x < 1000) prm.trace[17 + 2 * blockIdx.x] = gtimer();   // per-CTA end
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(Cfg::kTmemCols) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Persistent kernel for the big time-batched GEMMs (>= one 128x128 tile per SM, plain epilogue, no split-K): one CTA
// per SM walks the tile list (tile = blockIdx.x + k * gridDim.x over the concatenated tiles of the group's problems).
// The K-block stream never stops at a tile boundary: the producer, the converters and the MMA issuer count K-blocks
// ACROSS tiles, so the rings keep their phase and the first loads of tile n+1 are in flight while tile n still
// multiplies.  The accumulator is double-buffered in tensor memory (2 x 128 columns + 4 operand slots of 64 columns =
// all 512) and four dedicated warps drain tile n -- tensor memory -> a per-warp 32x32 staging block -> 128-byte
// row segments in global memory -- while the MMA issuer is already into tile n+1.  Measured before this kernel
// existed (tools/tc_kb_trace.py): the one-tile-per-CTA kernel spends 3 us in front of and 7 us behind a 35 us main
// loop, and the two-CTAs-per-SM variant starves on its 2-deep Q ring (1800 cycles per K-block per CTA against 800 of
// tensor-pipe work).
// Tiles are handed out dynamically (when there are more tiles than CTAs): the producer warp draws tiles from a
// global counter (one atomic per tile, issued a tile ahead) and publishes it to the other warps through a 4-deep ring
// in shared memory.  A static stride would tie the kernel's duration to its slowest CTA -- and a CTA whose SM is
// still held by a concurrent kernel (the data-parallel step all-reduces gradient buckets under these GEMMs) starts
// late by that kernel's whole duration; with the counter a late CTA simply finds nothing left.  The last CTA out
// re-arms the counter for the next launch of the stream.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kBigEpiWarps = 4;
constexpr int kBigFirstEpiWarp = 2 + kConvWarps;
constexpr int kBigThreads = 32 * (kBigFirstEpiWarp + kBigEpiWarps);
struct BigCfg {
  static constexpr int kNS = 4;                      // ring depth: raw P tiles, Q raw|lo tiles, tensor-memory operand slots
  static_assert(kNS % kConvGroups == 0, "a converter group must revisit a slot on consecutive barrier phases");
  static constexpr int kPBytes = kTileP * 128;
  static constexpr int kQBytes = 128 * 128;
  static constexpr int kQSlot = 2 * kQBytes;
  static constexpr int kRingBytes = kNS * (kPBytes + kQSlot);
  static constexpr int kEpiPitch = 36;               // floats per staged row: 16-byte aligned, conflict-free both ways
  static constexpr int kEpiBytes = kBigEpiWarps * 32 * kEpiPitch * 4;
  static constexpr int kSmemBytes = kRingBytes + kEpiBytes + 1024 /*align*/ + 256 /*barriers*/;
};

template <int G>
__global__ void __launch_bounds__(kBigThreads, 1) gemm_big_kernel(const __grid_constant__ TcGroup<G> grp) {
  using Cfg = BigCfg;
  constexpr int NS = Cfg::kNS;
  constexpr int QN = 128;
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_dyn + (base - smem_u32(smem_dyn));
  const uint32_t q_base = base + NS * Cfg::kPBytes;
  const uint32_t epi_base = base + Cfg::kRingBytes;
  const uint32_t bar_base = epi_base + Cfg::kEpiBytes;
  auto p_full = [&](int s) { return bar_base + 8u * s; };
  auto q_full = [&](int s) { return bar_base + 8u * (NS + s); };
  auto conv_bar = [&](int s) { return bar_base + 8u * (2 * NS + s); };
  auto empty_bar = [&](int s) { return bar_base + 8u * (3 * NS + s); };
  auto acc_full = [&](int a) { return bar_base + 8u * (4 * NS + a); };
  auto acc_empty = [&](int a) { return bar_base + 8u * (4 * NS + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (4 * NS + 4);
  constexpr int kSched = 4;                             // tile-id ring: producer -> MMA issuer, converters, epilogue
  constexpr int kSchedReaders = 1 + kConvWarps + kBigEpiWarps;
  auto sched_full = [&](int s) { return bar_base + 8u * (4 * NS + 5 + s); };
  auto sched_empty = [&](int s) { return bar_base + 8u * (4 * NS + 5 + kSched + s); };
  volatile int* sched_tile = reinterpret_cast<volatile int*>(gen_base + (bar_base - base) + 8 * (4 * NS + 5 + 2 * kSched));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ntiles = grp.cta_start[grp.n];
  int* const sched_ctr = grp.p[0].counters;             // [0] next tile, [1] CTAs that left
  static_assert(kConvGroups == 2, "the converters split the K-block stream by parity");
  // consumer side of the tile ring: the it-th tile of this CTA (-1: no more)
  auto next_tile = [&](uint32_t it) {
    const int s = (int)(it % kSched);
    mbar_wait(sched_full(s), (it / kSched) & 1u);
    const int t = sched_tile[s];
    __syncwarp();
    if (lane == 0) mbar_arrive(sched_empty(s));
    return t;
  };

  struct Tile { int pi, p0, q0, nkb; };
  auto decode = [&](int t) {
    Tile tl;
    tl.pi = 0;
    while (tl.pi + 1 < grp.n && t >= grp.cta_start[tl.pi + 1]) ++tl.pi;
    const TcParams& prm = grp.p[tl.pi];
    const int bid = t - grp.cta_start[tl.pi];
    tl.q0 = (bid % prm.tiles_q) * QN;
    tl.p0 = (bid / prm.tiles_q) * kTileP;
    tl.nkb = 0;
    for (int s = 0; s < prm.nseg; ++s) tl.nkb += (prm.K[s] + kBlockK - 1) / kBlockK;
    return tl;
  };

  if (threadIdx.x == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(p_full(s), 1);
      mbar_init(q_full(s), 1);
      mbar_init(conv_bar(s), 4);
      mbar_init(empty_bar(s), 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(acc_full(a), 1);
      mbar_init(acc_empty(a), kBigEpiWarps);
    }
    for (int k = 0; k < kSched; ++k) {
      mbar_init(sched_full(k), 1);
      mbar_init(sched_empty(k), kSchedReaders);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));
  const uint32_t tmem_ops = tmem_base + 2u * QN;   // operand slots behind the two accumulators

  pdl_trigger();
  if (warp == 0) {
    // ============================== TMA producer ==============================
    pdl_wait();   // (also orders this launch's draws from the tile counter behind the previous launch's re-arm)
    uint32_t g = 0;
    const bool dyn = ntiles > (int)gridDim.x;   // one tile per CTA: nothing to balance, no atomics
    int t = blockIdx.x;
    if (dyn) {
      if (lane == 0) t = atomicAdd(sched_ctr, 1);
      t = __shfl_sync(0xffffffffu, t, 0);
    }
    for (uint32_t it = 0;; ++it) {
      // publish tile `t` (or the end mark), then draw the one after it while this tile's K-blocks stream
      const int ss = (int)(it % kSched);
      if (it >= (uint32_t)kSched) mbar_wait(sched_empty(ss), ((it / kSched) & 1u) ^ 1u);
      const bool live = t < ntiles;
      if (lane == 0) {
        sched_tile[ss] = live ? t : -1;
        mbar_arrive(sched_full(ss));   // (release: the tile id is visible to whoever sees the phase complete)
      }
      __syncwarp();
      if (!live) break;
      int t_next = ntiles;
      if (dyn && lane == 0) t_next = atomicAdd(sched_ctr, 1);
      const Tile tl = decode(t);
      const TcParams& prm = grp.p[tl.pi];
      for (int sg = 0; sg < prm.nseg; ++sg) {
        const int nk = (prm.K[sg] + kBlockK - 1) / kBlockK;
        for (int kb = 0; kb < nk; ++kb, ++g) {
          const int s = (int)(g % NS);
          if (g >= (uint32_t)NS) mbar_wait(empty_bar(s), ((g / NS) & 1u) ^ 1u);   // MMA of K-block g - NS retired
          if (elect_one()) {
            mbar_expect_tx(q_full(s), Cfg::kQBytes);
            tma_load_2d(q_base + s * Cfg::kQSlot, &prm.mapQ[sg], q_full(s), kb * kBlockK, tl.q0);
            mbar_expect_tx(p_full(s), Cfg::kPBytes);
            tma_load_2d(base + s * Cfg::kPBytes, &prm.mapP[sg], p_full(s), kb * kBlockK, tl.p0);
          }
          __syncwarp();
        }
      }
      t = __shfl_sync(0xffffffffu, t_next, 0);
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(QN >> 3) << 17) | ((uint32_t)(kTileP >> 4) << 24);
    uint32_t g = 0, tc = 0;
    for (int t = next_tile(0); t >= 0; t = next_tile(++tc)) {
      const Tile tl = decode(t);
      const uint32_t ab = tc & 1u;
      if (tc >= 2u) mbar_wait(acc_empty(ab), ((tc >> 1) & 1u) ^ 1u);   // the epilogue drained this accumulator
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t acc = tmem_base + ab * (uint32_t)QN;
      for (int i = 0; i < tl.nkb; ++i, ++g) {
        const int s = (int)(g % NS);
        mbar_wait(conv_bar(s), (g / NS) & 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint32_t q_hi = q_base + s * Cfg::kQSlot, q_lo = q_hi + Cfg::kQBytes;
          const uint64_t b_hi0 = umma_desc(q_hi, 16u, 1024u), b_lo0 = umma_desc(q_lo, 16u, 1024u);
          const uint32_t ta0 = tmem_ops + (uint32_t)s * 64u;
#pragma unroll
          for (int k = 0; k < kBlockK / 8; ++k) {
            const uint64_t b_hi = b_hi0 + (uint64_t)(2 * k), b_lo = b_lo0 + (uint64_t)(2 * k);
            const uint32_t ta_hi = ta0 + (uint32_t)k * 8u;
            umma_tf32_ts(acc, ta_hi + 32u, b_hi, idesc, (i > 0 || k > 0) ? 1u : 0u);   // P_lo * Q_hi
            umma_tf32_ts(acc, ta_hi, b_lo, idesc, 1u);                                 // P_hi * Q_lo
            umma_tf32_ts(acc, ta_hi, b_hi, idesc, 1u);                                 // P_hi * Q_hi
          }
          umma_commit(empty_bar(s));
          if (i == tl.nkb - 1) umma_commit(acc_full(ab));
        }
        __syncwarp();
      }
    }
  } else if (warp < kBigFirstEpiWarp) {
    // ============================== converters ==============================
    const int grp_id = (warp - 2) >> 2;
    const int gt = (threadIdx.x - 64) & (kGT - 1);
    auto split = [](float x, float& lo) { lo = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); };
    uint32_t g = 0, tcount = 0;   // g: K-blocks of all earlier tiles of this CTA
    for (int t = next_tile(0); t >= 0; t = next_tile(++tcount)) {
    const uint32_t g_end = g + (uint32_t)decode(t).nkb;
    for (g += (g + (uint32_t)grp_id) & 1u ? 1u : 0u; g < g_end; g += kConvGroups) {   // this group's K-blocks: g % 2 == grp_id
      const int s = (int)(g % NS);
      const uint32_t ph = (g / NS) & 1u;
      // Q: the raw tile doubles as the hi operand (kind::tf32 ignores the low 13 mantissa bits); lo goes to the sibling
      mbar_wait(q_full(s), ph);
      const float4* q_hi = reinterpret_cast<const float4*>(gen_base + (q_base - base) + s * Cfg::kQSlot);
      float4* q_lo = reinterpret_cast<float4*>(gen_base + (q_base - base) + s * Cfg::kQSlot + Cfg::kQBytes);
#pragma unroll
      for (int j = 0; j < Cfg::kQBytes / 16 / kGT; ++j) {
        const float4 v = q_hi[gt + kGT * j];
        float4 l;
        split(v.x, l.x); split(v.y, l.y); split(v.z, l.z); split(v.w, l.w);
        q_lo[gt + kGT * j] = l;
      }
      // P: tile row = tensor-memory lane; hi | lo as 2 x 32 columns of operand slot s
      mbar_wait(p_full(s), ph);
      const float4* p_raw = reinterpret_cast<const float4*>(gen_base + s * Cfg::kPBytes);
      const int prow = (warp & 3) * 32 + lane;
      const uint32_t ta = tmem_ops + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)s * 64u;
#pragma unroll
      for (int half = 0; half < 2; ++half) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const float4 v = p_raw[prow * 8 + ((half * 4 + cc) ^ (prow & 7))];
          float l;
          split(v.x, l); hi[cc * 4 + 0] = __float_as_uint(v.x); lo[cc * 4 + 0] = __float_as_uint(l);
          split(v.y, l); hi[cc * 4 + 1] = __float_as_uint(v.y); lo[cc * 4 + 1] = __float_as_uint(l);
          split(v.z, l); hi[cc * 4 + 2] = __float_as_uint(v.z); lo[cc * 4 + 2] = __float_as_uint(l);
          split(v.w, l); hi[cc * 4 + 3] = __float_as_uint(v.w); lo[cc * 4 + 3] = __float_as_uint(l);
        }
        tmem_st16(ta + (uint32_t)half * 16u, hi);
        tmem_st16(ta + 32u + (uint32_t)half * 16u, lo);
      }
      asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(conv_bar(s));
    }
    g = g_end;
    }
  } else {
    // ============================== epilogue ==============================
    pdl_wait();   // C, `add` and c_row_len may be produced by the preceding kernels
    const int quarter = warp & 3;   // tensor-memory lane quarter this warp may read
    float* stg = reinterpret_cast<float*>(gen_base + (epi_base - base)) + (warp - kBigFirstEpiWarp) * 32 * Cfg::kEpiPitch;
    const int rsub = lane >> 3, c4 = (lane & 7) * 4;   // write-out: 8 lanes cover one 128-byte row segment
    uint32_t tc = 0;
    for (int t = next_tile(0); t >= 0; t = next_tile(++tc)) {
      const Tile tl = decode(t);
      const TcParams& prm = grp.p[tl.pi];
      const uint32_t ab = tc & 1u;
      const int m_lim = prm.Pr, n_lim = prm.Qr;
      float* const Cp = prm.C;
      const long ldc = prm.ldc, c_ld_inner = prm.c_ld_inner, ldadd = prm.ldadd;
      const int c_inner = prm.c_inner, vin = prm.c_valid_inner, add_mod = prm.add_mod, beta = prm.beta, act = prm.act;
      const int* row_len = prm.c_row_len;
      const float* bias = prm.bias; const float* bias2 = prm.bias2; const float* add = prm.add;
      auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
      const bool fast = al16(Cp) && ldc % 4 == 0 && (c_inner == 0 || c_ld_inner % 4 == 0) && al16(bias) && al16(bias2) &&
                        al16(add) && ldadd % 4 == 0;
      mbar_wait(acc_full(ab), (tc >> 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
      for (int cb = 0; cb < QN / 32; ++cb) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + ab * (uint32_t)QN + (uint32_t)(cb * 32), r);
        if (cb == QN / 32 - 1) {   // the accumulator is in registers: hand it back to the MMA issuer
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) mbar_arrive(acc_empty(ab));
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          *reinterpret_cast<float4*>(&stg[lane * Cfg::kEpiPitch + j]) =
              make_float4(__uint_as_float(r[j]), __uint_as_float(r[j + 1]), __uint_as_float(r[j + 2]), __uint_as_float(r[j + 3]));
        __syncwarp();
        const int n = tl.q0 + cb * 32 + c4;
        const int nv = min(4, n_lim - n);
        const int m0 = tl.p0 + quarter * 32 + rsub;   // this lane's rows: m0 + 4 * pass
        auto row_ok = [&](int m) {
          return m < m_lim && (row_len == nullptr || row_len[m % vin] > m / vin);
        };
        auto c_off = [&](int m) {
          return (c_inner > 0 ? (long)(m / c_inner) * ldc + (long)(m % c_inner) * c_ld_inner : (long)m * ldc) + n;
        };
        if (fast && nv == 4) {
          // every operand is 16-byte aligned: the 8 rows' reads of C (beta) and `add` go out together, one round trip
          const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const float4 b1 = bias ? __ldg(reinterpret_cast<const float4*>(bias + n)) : z4;
          const float4 b2 = bias2 ? __ldg(reinterpret_cast<const float4*>(bias2 + n)) : z4;
          float4 ov[8], av[8];
          long off[8];
          unsigned okm = 0u;
#pragma unroll
          for (int pass = 0; pass < 8; ++pass) {
            const int m = m0 + 4 * pass;
            const bool ok = row_ok(m);
            okm |= ok ? (1u << pass) : 0u;
            off[pass] = ok ? c_off(m) : 0;
            ov[pass] = (ok && beta) ? *reinterpret_cast<const float4*>(Cp + off[pass]) : z4;
            av[pass] = (ok && add) ? *reinterpret_cast<const float4*>(add + (long)(add_mod ? m % add_mod : m) * ldadd + n) : z4;
          }
#pragma unroll
          for (int pass = 0; pass < 8; ++pass) {
            if (!((okm >> pass) & 1u)) continue;
            const float4 a4 = *reinterpret_cast<const float4*>(&stg[(pass * 4 + rsub) * Cfg::kEpiPitch + c4]);
            float v[4] = {a4.x, a4.y, a4.z, a4.w};
            if (bias) { v[0] += b1.x; v[1] += b1.y; v[2] += b1.z; v[3] += b1.w; }
            if (bias2) { v[0] += b2.x; v[1] += b2.y; v[2] += b2.z; v[3] += b2.w; }
            if (add) { v[0] += av[pass].x; v[1] += av[pass].y; v[2] += av[pass].z; v[3] += av[pass].w; }
            if (act == 1) { for (int k = 0; k < 4; ++k) v[k] = fmaxf(v[k], 0.f); }
            else if (act == 2) { for (int k = 0; k < 4; ++k) v[k] = tanhf(v[k]); }
            if (beta) { v[0] += ov[pass].x; v[1] += ov[pass].y; v[2] += ov[pass].z; v[3] += ov[pass].w; }
            *reinterpret_cast<float4*>(Cp + off[pass]) = make_float4(v[0], v[1], v[2], v[3]);
          }
        } else if (nv > 0) {
          // ragged right edge or unaligned operands: element-wise
#pragma unroll 1
          for (int pass = 0; pass < 8; ++pass) {
            const int m = m0 + 4 * pass;
            if (!row_ok(m)) continue;
            const float4 a4 = *reinterpret_cast<const float4*>(&stg[(pass * 4 + rsub) * Cfg::kEpiPitch + c4]);
            float v[4] = {a4.x, a4.y, a4.z, a4.w};
            float* cp = Cp + c_off(m);
            const float* ap = add ? add + (long)(add_mod ? m % add_mod : m) * ldadd + n : nullptr;
            for (int k = 0; k < nv; ++k) {
              float x = v[k];
              if (bias) x += __ldg(bias + n + k);
              if (bias2) x += __ldg(bias2 + n + k);
              if (ap) x += ap[k];
              if (act == 1) x = fmaxf(x, 0.f);
              else if (act == 2) x = tanhf(x);
              cp[k] = beta ? cp[k] + x : x;
            }
          }
        }
        __syncwarp();   // the staging block is rewritten by the next column block
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
  if (threadIdx.x == 0) {
    // every CTA has drawn its last tile before it gets here: the last one out re-arms the counters
    __threadfence();
    if (atomicAdd(sched_ctr + 1, 1) == (int)gridDim.x - 1) {
      sched_ctr[0] = 0;
      sched_ctr[1] = 0;
      __threadfence();
    }
  }
}

// ------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn g_encode = nullptr;
// split-K scratch of the fused epilogue: one [128][128] fp32 slab and one arrival counter per CTA of a launch.
// Launches of one stream are ordered (a launch's epilogue begins after its grid dependency resolved), so one slab set
// serves every launch of a (device, stream) pair; the set is keyed by that pair (lib_scratch), so several streams,
// threads or devices of one process never share slabs or counters.
constexpr int kScratchSlots = 320;
constexpr int kMaxDevices = 64;
struct TcDevice {
  std::once_flag once;
  bool ready = false;
  int sm_count = 0;
  int max_clusters[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};   // [c]: clusters of c CTAs (one CTA per SM) resident at once
};
TcDevice g_tc_dev[kMaxDevices];

// per-device set-up: function attributes live in the device's context, cluster occupancy depends on its SM layout
void tc_init(TcDevice* d, int dev) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) return;
  g_encode = reinterpret_cast<EncodeTiledFn>(fn);
  bool ok = true;
  auto set_attr = [&](auto kern, int bytes) {
    ok = ok && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes) == cudaSuccess;
  };
  set_attr(gemm_tc_kernel<64, 1>, TcCfg<64>::kSmemBytes);  set_attr(gemm_tc_kernel<128, 1>, TcCfg<128>::kSmemBytes);
  set_attr(gemm_tc_kernel<64, 2>, TcCfg<64>::kSmemBytes);  set_attr(gemm_tc_kernel<128, 2>, TcCfg<128>::kSmemBytes);
  set_attr(gemm_tc_kernel<64, 5>, TcCfg<64>::kSmemBytes);  set_attr(gemm_tc_kernel<128, 5>, TcCfg<128>::kSmemBytes);
  set_attr(gemm_tc_kernel<64, 8>, TcCfg<64>::kSmemBytes);  set_attr(gemm_tc_kernel<128, 8>, TcCfg<128>::kSmemBytes);
  set_attr(gemm_tc_kernel<128, 1, 1>, TcCfg<128, 1>::kSmemBytes);  set_attr(gemm_tc_kernel<128, 2, 1>, TcCfg<128, 1>::kSmemBytes);
  set_attr(gemm_tc_kernel<128, 5, 1>, TcCfg<128, 1>::kSmemBytes);  set_attr(gemm_tc_kernel<128, 8, 1>, TcCfg<128, 1>::kSmemBytes);
  set_attr(gemm_big_kernel<1>, BigCfg::kSmemBytes);  set_attr(gemm_big_kernel<2>, BigCfg::kSmemBytes);
  set_attr(gemm_big_kernel<5>, BigCfg::kSmemBytes);  set_attr(gemm_big_kernel<8>, BigCfg::kSmemBytes);
  ok = ok && cudaDeviceGetAttribute(&d->sm_count, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess;
  for (int c = 2; ok && c <= 8; c <<= 1) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(c * 64); cfg.blockDim = dim3(kThreadsTc); cfg.dynamicSmemBytes = TcCfg<64>::kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, gemm_tc_kernel<64, 1>, &cfg) != cudaSuccess) { cudaGetLastError(); nc = 0; }
    d->max_clusters[c] = nc;
  }
  if (getenv("SET_TC_VERBOSE"))
    fprintf(stderr, "libset_b200: device %d: %d SMs, resident clusters of 2/4/8 CTAs: %d/%d/%d\n", dev, d->sm_count,
            d->max_clusters[2], d->max_clusters[4], d->max_clusters[8]);
  if (!ok) { cudaGetLastError(); return; }
  d->ready = true;
}

// the current device's state (nullptr: tensor-core path unavailable -> callers use the CUDA-core kernel)
TcDevice* tc_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDevices) { cudaGetLastError(); return nullptr; }
  TcDevice* d = &g_tc_dev[dev];
  std::call_once(d->once, tc_init, d, dev);
  return d->ready ? d : nullptr;
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

bool tc_encode_map(CUtensorMap* m, const float* ptr, int rank, const unsigned long long* dims,
                   const unsigned long long* strides_bytes, const unsigned int* box, bool swizzle128) {
  if (!tc_device() || rank < 2 || rank > 3) return false;
  cuuint64_t d[3]; cuuint64_t st[2]; cuuint32_t bx[3]; cuuint32_t es[3] = {1, 1, 1};
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; bx[i] = box[i]; }
  for (int i = 0; i + 1 < rank; ++i) st[i] = strides_bytes[i];
  return g_encode(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, rank, const_cast<float*>(ptr), d, st, bx, es,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

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
  // fused epilogue (scratch-slab split-K + optional cell): swap mode, plain row-major C, 16-byte aligned vectors
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  static const int fuse_on = getenv("SET_TC_FUSE") ? atoi(getenv("SET_TC_FUSE")) : 1;
  const bool fuse_ok = fuse_on && swap && g.act == 0 && g.c_inner == 0 && !g.c_row_len && g.N % 4 == 0 && al16(g.C) &&
                       g.ldc % 4 == 0 && al16(g.bias) && al16(g.bias2) && al16(g.add) && g.ldadd % 4 == 0;
  const int op = fuse_ok ? g.epi.op : kEpiNone;
  const bool gates4 = (op == kEpiLstm || op == kEpiCopy1);
  if (gates4 && !(g.epi.D % 32 == 0 && g.N == 4 * g.epi.D)) return false;
  if (op == kEpiCopy2 && g.N != g.epi.D) return false;
  if ((op == kEpiLstmBwd || op == kEpiCopy1Bwd || op == kEpiCopy2Bwd) && g.N != g.epi.D) return false;
  if (op == kEpiCtxGateBwd && !(g.epi.col0 % kTileP == 0 && g.epi.D % kTileP == 0 && g.epi.col0 + g.epi.D <= g.N)) return false;
  // Plain split-K keeps the fire-and-forget red.global.add epilogue (measured faster than the slab protocol
  // when there is no cell to apply); the slab path is for problems that carry a cell.
  // ... and for plain problems whose C is neither pre-zeroed nor accumulated into: the slab path needs no memset
  // in front of the launch (SET_TC_PLAIN_FUSED=0 restores memset + red.global.add for them).
  static const int plain_fused = getenv("SET_TC_PLAIN_FUSED") ? atoi(getenv("SET_TC_PLAIN_FUSED")) : 1;
  prm.fused = (fuse_ok && (op != kEpiNone || plain_fused == 2 || (plain_fused && !g.beta && !g.c_zeroed))) ? 1 : 0;
  prm.fuse_ok = fuse_ok ? 1 : 0;
  prm.nblk = gates4 ? 4 : 1;
  prm.blk_stride = gates4 ? g.epi.D : 0;
  prm.epi = g.epi; prm.epi.op = op;
  // (scratch / counters / zero16 are filled per (device, stream) by gemm_tc_try_group)
  for (int s = 0; s < g.nseg; ++s) {
    const GemmSeg& sg = g.seg[s];
    prm.K[s] = sg.K;
    const float* Pp = swap ? sg.B : sg.A; const long Pld = swap ? sg.ldb : sg.lda;
    const float* Qp = swap ? sg.A : sg.B; const long Qld = swap ? sg.lda : sg.ldb;
    if (!make_map(&prm.mapP[s], Pp, Pr, sg.K, Pld, kBlockK, kTileP / prm.nblk)) return false;
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
  TcDevice* tcd = tc_device();
  if (!tcd) return SET_OK;
  const int g_sm_count = tcd->sm_count;
  const int* g_max_clusters = tcd->max_clusters;
  // one Q-tile width per launch: 64 if every problem's small side fits, else 128
  int QN = 64;
  for (int i = 0; i < n; ++i) {
    const GemmProblem& g = probs[i];
    if (g.M <= 0 || g.N <= 0) continue;
    const int small = g.M < g.N ? g.M : g.N;
    if (small > 64) QN = 128;
  }
  static thread_local TcGroup<8> grp;  // host staging (the launch copies it)
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
  // split-K: spread the group over one wave of the 148 SMs (a second wave would repeat every CTA's fixed
  // prologue/epilogue).  Greedy balance: the next split goes to the problem whose CTAs carry the most K-blocks.
  long nkbs[8]; int tiles[8], splits[8]; bool splittable[8];
  static const int max_fused_split = []{
    int v = getenv("SET_TC_MAX_FSPLIT") ? atoi(getenv("SET_TC_MAX_FSPLIT")) : 6;
    return v < 1 ? 1 : (v > kMaxFusedSplit ? kMaxFusedSplit : v);
  }();
  for (int k = 0; k < grp.n; ++k) {
    const GemmProblem& g = probs[idx[k]];
    nkbs[k] = 0;
    for (int s = 0; s < g.nseg; ++s) nkbs[k] += (g.seg[s].K + kBlockK - 1) / kBlockK;
    tiles[k] = grp.p[k].tiles_p * grp.p[k].tiles_q;
    splits[k] = 1;
    splittable[k] = (g.act == 0 && !(g.c_inner > 0 || g.c_row_len));
  }
  if (tiles_total < 148) {
    long ctas = tiles_total;
    for (;;) {
      int best = -1; double best_load = 0.0;
      for (int k = 0; k < grp.n; ++k) {
        if (!splittable[k] || ctas + tiles[k] > 148) continue;
        const int cap = grp.p[k].fused ? max_fused_split : 16;
        if (splits[k] >= cap || nkbs[k] / (splits[k] + 1) < 4) continue;
        const double load = (double)nkbs[k] / splits[k];
        if (load > best_load) { best_load = load; best = k; }
      }
      if (best < 0) break;
      ++splits[best];
      ctas += tiles[best];
    }
    // a single long-K problem that fills only ~half the machine: three partials over two waves is ~1.5x faster
    if (grp.n == 1 && splits[0] == 1 && splittable[0] && tiles_total * 3 <= 2 * 148 && nkbs[0] >= 96) splits[0] = 3;
  }
  // A single skinny problem: its split-K partners form a thread-block cluster and reduce through distributed shared
  // memory (power-of-two cluster sizes; the grid must fit the clusters the device can hold at once).
  int cluster = 1;
  {
    static const int cluster_on = getenv("SET_TC_CLUSTER") ? atoi(getenv("SET_TC_CLUSTER")) : 1;
    bool all_ok = cluster_on && tiles_total < 148;
    long min_nkb = 1L << 40, max_nkb = 0;
    double greedy_max = 0.0;
    for (int k = 0; k < grp.n; ++k) {
      all_ok = all_ok && grp.p[k].fuse_ok && splittable[k];
      min_nkb = nkbs[k] < min_nkb ? nkbs[k] : min_nkb;
      max_nkb = nkbs[k] > max_nkb ? nkbs[k] : max_nkb;
      const double l = (double)nkbs[k] / splits[k];
      greedy_max = l > greedy_max ? l : greedy_max;
    }
    if (all_ok) {
      // one split for the whole launch (= the cluster size).  Groups: off by default (measured on B200: the three-GEMM
      // group on ctx/sel and the reverse-pass pairs run 2-3 us SLOWER as clusters than with per-problem splits +
      // red.global.add -- uniform splits unbalance them and big clusters wait for whole GPC slices to drain); when
      // enabled (SET_TC_GROUP_CLUSTER=1), only if it does not cost more than ~12
      // K-blocks of load balance against the greedy plan (the cluster epilogue is worth about that much)
      static const int max_cluster = getenv("SET_TC_MAX_CLUSTER") ? atoi(getenv("SET_TC_MAX_CLUSTER")) : 8;
      static const int group_cluster = getenv("SET_TC_GROUP_CLUSTER") ? atoi(getenv("SET_TC_GROUP_CLUSTER")) : 0;
      int sp = max_cluster >= 8 ? 8 : (max_cluster >= 4 ? 4 : (max_cluster >= 2 ? 2 : 1));
      while (sp > 1 && (tiles_total * sp > 148 || min_nkb / sp < 4 || tiles_total > g_max_clusters[sp])) sp >>= 1;
      if (sp > 1 && (grp.n == 1 || (group_cluster && (double)max_nkb / sp <= greedy_max + 12.0))) {
        cluster = sp;
        for (int k = 0; k < grp.n; ++k) splits[k] = sp;
      }
    }
  }
  // library-owned scratch of this (device, stream): slabs, two counters per CTA (+ 16 bytes of zeros: TcParams::zero16)
  float* tc_scratch = static_cast<float*>(lib_scratch(kScratchTcSlabs, stream, sizeof(float) * (size_t)kScratchSlots * kTileP * 128, false));
  int* tc_counters = static_cast<int*>(lib_scratch(kScratchTcCounters, stream, sizeof(int) * (2 * kScratchSlots + 8), true));
  if (!tc_scratch || !tc_counters) return SET_ERR_CUDA;
  int cta = 0;
  for (int k = 0; k < grp.n; ++k) {
    TcParams& prm = grp.p[k];
    prm.scratch = tc_scratch; prm.counters = tc_counters;
    prm.zero16 = reinterpret_cast<const float*>(tc_counters + 2 * kScratchSlots);
    const GemmProblem& g = probs[idx[k]];
    const int split = splits[k];
    prm.split_k = split;
    prm.cluster = cluster;
    if (cluster > 1) prm.fused = 1;
    if (prm.fused && split == 1 && prm.epi.op == kEpiNone) prm.fused = 0;   // nothing to reduce, nothing to apply
    if (prm.fused && cta + tiles[k] * split > kScratchSlots) {
      // (cannot happen with the one-wave heuristic above; guard the slab indexing anyway)
      set_record_error("fused GEMM launch exceeds the split-K scratch");
      return SET_ERR_ARG;
    }
    if (!prm.fused && split > 1 && !g.beta && !g.c_zeroed)   // partial sums are reduced into C: it must start at zero
      SET_CHECK_CUDA(cudaMemset2DAsync(g.C, sizeof(float) * g.ldc, 0, sizeof(float) * g.N, g.M, stream));
    if (prm.fused && prm.epi.op != kEpiNone && g.epi_done) *g.epi_done = 1;
    grp.cta_start[k] = cta;
    cta += tiles[k] * split;
    taken[idx[k]] = true;
  }
  grp.cta_start[grp.n] = cta;
  {
    // cooperative finish spins on partner CTAs: only when the whole grid is resident at once (one CTA per SM)
    static const int coop_on = getenv("SET_TC_COOP") ? atoi(getenv("SET_TC_COOP")) : 1;
    const int coop = (coop_on && cta <= g_sm_count) ? 1 : 0;
    for (int k = 0; k < grp.n; ++k) grp.p[k].coop = coop;
  }
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
    if (cluster > 1) {
      if (QN == 64)
        return launch_chain_cluster(gemm_tc_kernel<64, G>, dim3(cta), dim3(kThreadsTc), TcCfg<64>::kSmemBytes, stream, cluster, small);
      return launch_chain_cluster(gemm_tc_kernel<128, G>, dim3(cta), dim3(kThreadsTc), TcCfg<128>::kSmemBytes, stream, cluster, small);
    }
    if (QN == 64) return launch_chain(gemm_tc_kernel<64, G>, dim3(cta), dim3(kThreadsTc), TcCfg<64>::kSmemBytes, stream, small);
    // plain 128x128-tile problems without split-K: the persistent kernel (one CTA per SM walks the tile list, the
    // epilogue of a tile overlaps the main loop of the next).  SET_TC_BIG=0 restores the one-tile-per-CTA kernels.
    static const int big_on = getenv("SET_TC_BIG") ? atoi(getenv("SET_TC_BIG")) : 1;
    bool big_ok = big_on != 0;
    for (int k = 0; k < grp.n; ++k)
      big_ok = big_ok && !grp.p[k].fused && grp.p[k].split_k == 1 && !grp.p[k].swap && grp.p[k].nblk == 1;
    if (big_ok) {
      ++g_tc_twin_launches;
      small.p[0].counters = tc_counters + 2 * kScratchSlots + 4;   // tile counter + exit counter of the persistent kernel
    }
    if (big_ok)
      return launch_chain(gemm_big_kernel<G>, dim3(cta < g_sm_count ? cta : g_sm_count), dim3(kBigThreads),
                          BigCfg::kSmemBytes, stream, small);
    // many tiles per SM and no fused epilogue in play: two CTAs per SM (see TcCfg)
    static const int twin_on = getenv("SET_TC_TWIN") ? atoi(getenv("SET_TC_TWIN")) : 1;
    bool any_fused = false;
    for (int k = 0; k < grp.n; ++k) any_fused = any_fused || grp.p[k].fused;
    if (twin_on && !any_fused && cta > g_sm_count && (++g_tc_twin_launches, true))
      return launch_chain(gemm_tc_kernel<128, G, 1>, dim3(cta), dim3(kThreadsTc), TcCfg<128, 1>::kSmemBytes, stream, small);
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
