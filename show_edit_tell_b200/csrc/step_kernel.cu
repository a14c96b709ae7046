// Persistent decode-step kernel (sm_100a): the whole per-timestep stack of EditNet -- attention-LSTM, both attentions
// with the select, context gate, copy-LSTM (editnet.py:527-543) -- for `nt`
// consecutive timesteps in ONE cooperative launch of one CTA per SM.
//
// A step is seven phases separated by grid-wide barriers (an atomic counter in L2, ~1 us) instead of kernel
// boundaries:
//   A   attention-LSTM gates  [h2_prev ; h1_prev] x [W_ih[:,2D:3D] ; W_hh]   + LSTM cell            -> h1, c1
//   B   everything that consumes h1: cap_decoder_att, decoder_att, context_gate / tc_affine h1 parts,
//       copy-LSTM x2h[:, 0:D] h1 + h2h h2_prev                                                       -> s2, g2
//   C1  attention scores (both attentions): w . act(att1_j + att2), one warp per (sample, row)        -> raw scores
//   C2  masked softmax, context = sum alpha_j value_j with the value rows TMA-staged in shared memory,
//       argmax + memory-row gather (SelectC)                                              -> ctx, sel, att_img, alphas
//   D   [context_gate ctx part | sc_affine] as two-block tiles + context-gate cell -> att_cap;  gate_cmem(sel);
//       x2h[:, 2D:] att_img                                                                           -> g2 +=
//   E   x2h[:, D:2D] att_cap + copy-LSTM stage 1 (four-gate tiles)                                    -> c_new
//   F   gate_cnew(c_new) + copy gate, c2, h2, dropout(h2)                                             -> h2, c2
//
// GEMM phases run the 3xTF32 tcgen05 pipeline of gemm_tc.cu in "swap" form (weights = 128-row P tiles fed to the MMA
// from tensor memory as hi | lo, batch = one 64-row Q tile in shared memory), one (tile, K-split) job per CTA and
// phase.  What the persistent form buys over the launch chain:
//   * tensor memory, mbarriers and tensor maps are set up once per launch, not once per GEMM;
//   * the weight (P) ring has its own producer warp that never waits for a phase boundary: weights are constants, so
//     the ring refills with the NEXT phase's tiles while this phase drains, reduces and synchronises;
//   * the split-K partners of a tile are the CTAs of one thread-block cluster (splits 1, 2, 4): partial tiles stay in
//     shared memory, the partners signal each other with remote mbarrier arrives and every CTA finishes 1/split of the
//     tile out of its partners' shared memory (DSMEM; fixed summation order, no atomics, no scratch in L2) and applies
//     the cell that consumes the GEMM;
//   * the attention value rows (36 x 2048 region features, 18 x 1024 encoder states per sample) are constants too:
//     their first chunks are requested by TMA before the scores exist and stream through a 3-deep shared-memory ring.
//
// Warp roles (512 threads): 0 = weight producer, 1 = tensor-memory owner + MMA issuer, 2 = activation producer
// (the only TMA reader of data other CTAs wrote: it waits for the grid barrier), 3 = spare, 4..7 = weight converters
// (fp32 tile -> hi | lo in tensor memory, running ahead of the activations), 8..15 = activation lo-split, epilogue /
// finish, attention, and the CTA's arrival at the grid barriers.
#include "step_kernel.cuh"

#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "tc_common.cuh"

namespace set {

namespace {

#ifndef SET_STEP_NP
#define SET_STEP_NP 4
#endif
constexpr int kNP = SET_STEP_NP;       // weight ring slots (16 KB each): what keeps HBM requests in flight
constexpr int kNQ = 12;                // raw activation (Q) tiles in flight, 8 KB each: a job's Q operand lands in one round trip
constexpr int kNL = 4;                 // lo tiles (x - hi), 8 KB each, recycled at the MMA's pace
constexpr int kNT = 7;                 // tensor-memory slots of converted weight tiles (hi | lo, 64 columns each)
constexpr int kQN = 64;                // batch rows of the Q tile
constexpr int kTileP = 128;
constexpr int kBlockK = 32;
constexpr int kThreads = 512;
constexpr int kEpiWarps = 8;
constexpr int kEpiThreads = 32 * kEpiWarps;
constexpr int kFirstConvWarp = 4;      // warps 4..7: weight converters (one warp per tensor-memory lane quarter)
constexpr int kConvGroups = 1;         // (two groups alternating K-blocks -- 608 threads, 96 registers -- measured slower: 114 vs 94 us per step)
constexpr int kFirstEpiWarp = 8;       // warps 8..15: Q lo-split, epilogue / finish, attention, grid barriers
constexpr int kMaxSplit = 4;            // split-K partners are the CTAs of one thread-block cluster
constexpr int kPBytes = kTileP * 128;  // 16 KB
constexpr int kQBytes = kQN * 128;     // 8 KB
constexpr int kAttnCols = 256;         // columns of a staged value chunk
constexpr int kAttnRows = 36;          // rows of a staged value chunk (box rows <= this)
constexpr int kAttnBuf = kAttnRows * kAttnCols * 4;
constexpr int kAttnBufs = 3;
constexpr int kAttnBatch = 8;          // softmaxes computed per round (one per epilogue warp)
constexpr int kAttnMaxN = 128;         // max(P, R) supported by the persistent path
constexpr int kAttnMaxItems = 32;      // (sample, column slice) items of one CTA in phase C2
constexpr int kAttnMaxUnits = 96;      // value chunks of one CTA in phase C2
constexpr int kEpPitch = kTileP + 4;   // staged accumulator tile: [64 q][128 p], padded
constexpr int kStageBytes = (kNQ + kNL) * kQBytes;           // 131072: Q raw + lo rings; aliased by the accumulator tile (33 KB)
static_assert(kStageBytes >= kAttnBufs * kAttnBuf && kStageBytes >= kQN * kEpPitch * 4, "staging region too small");  // and the attention chunks
static_assert(64 + 64 * kNT <= 512, "tensor memory: accumulator + converted weight slots");
constexpr int kNumBars = 2 * kNP + 2 * kNT + 2 * kNQ + 2 * kNL + 1 + kAttnBufs + kMaxSplit;
constexpr int kSmallBytes = 8 * kNumBars + 16 /*tmem slot, flags*/ + 4 * (kAttnBatch * kAttnMaxN + 4 * kAttnBatch) + 256 /*jobs*/ + 16 * (kAttnMaxUnits + kAttnMaxItems);
constexpr int kSmemBytes = 1024 + kNP * kPBytes + kStageBytes + ((kSmallBytes + 127) & ~127);
constexpr uint32_t kTmemCols = 512;

#define EPI_BAR() asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory")
// fine-grained stamps of CTA 0 behind the per-phase table of the trace buffer: [(step * 5 + gemm phase) * 16 + k]
// (compiled in with -DSET_STEP_FINE_TRACE: tests/tools/step_trace.py prints them; off by default, the stamps sit on
// the critical path)
#ifdef SET_STEP_FINE_TRACE
#define FS(s_, ph_, k_)                                                                                   \
  do {                                                                                                    \
    if (P.trace && (int)blockIdx.x == P.trace_cta) P.trace[(long)P.nt * 8 * gridDim.x + ((s_) * 5 + (ph_)) * 16 + (k_)] = gtimer(); \
  } while (0)
#else
#define FS(s_, ph_, k_) do { } while (0)
#endif

struct Job {
  int prob;      // -1: no job in this phase
  int tile, ks, split, first_cta;
  int kb_begin, kb_end;
};

__device__ __forceinline__ Job get_job(const StepParams& P, int ph, int cta) {
  Job j;
  j.prob = -1;
  const StepPhase& phs = P.phase[ph];
  if (cta >= phs.ncta) return j;
  int k = 0;
  while (k + 1 < phs.nprob && cta >= P.prob[phs.prob[k + 1]].cta0) ++k;
  const StepProb& pr = P.prob[phs.prob[k]];
  const int bid = cta - pr.cta0;
  j.prob = phs.prob[k];
  j.split = pr.split;
  j.ks = bid % pr.split;
  j.tile = bid / pr.split;
  j.first_cta = cta - j.ks;
  const int total = pr.nkb[0] + (pr.nseg > 1 ? pr.nkb[1] : 0);
  j.kb_begin = (int)((long)total * j.ks / pr.split);
  j.kb_end = (int)((long)total * (j.ks + 1) / pr.split);
  return j;
}

__device__ __forceinline__ float* tp(const TPtr& x, int t) { return x.p + (long)t * x.st; }

// Compact activations (ex2.approx + fast reciprocal, absolute error ~1e-7): the persistent kernel's code has to fit the
// instruction cache, and the library forms of tanhf / expf / division inline to hundreds of bytes per call site.
__device__ __forceinline__ float sig_(float x) { return __fdividef(1.0f, 1.0f + __expf(-x)); }
__device__ __forceinline__ float tanh_(float x) {
  const float xc = fminf(fmaxf(x, -15.f), 15.f);
  return 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * xc));
}

__device__ __forceinline__ float2 ldcg2(const float* p) { return __ldcg(reinterpret_cast<const float2*>(p)); }

}  // namespace

static_assert(sizeof(StepParams) <= 16 * 1024, "kernel parameter block too large");

// ring position: slot index + phase parity, advanced without divisions
struct RingPos {
  uint32_t slot = 0, phase = 0, wrapped = 0;
  __device__ __forceinline__ void advance(uint32_t n) {
    if (++slot == n) { slot = 0; phase ^= 1u; wrapped = 1u; }
  }
};

// Code size matters here: the roles of a persistent kernel share one instruction cache, and everything the epilogue
// group runs between two grid barriers is on the step's critical path.  Hence ONE finish loop for every cell, one
// grid-barrier site, jobs decoded once into shared memory, ring positions advanced without divisions, compact
// activations.
__global__ void __launch_bounds__(kThreads, 1) step_kernel(const __grid_constant__ StepParams P) {
  extern __shared__ uint8_t smem_dyn[];
  const uint32_t base = (smem_u32(smem_dyn) + 1023u) & ~1023u;
  uint8_t* gen_base = smem_dyn + (base - smem_u32(smem_dyn));
  const uint32_t stage = base + kNP * kPBytes;            // Q ring | accumulator tile | attention chunks
  uint8_t* gen_stage = gen_base + kNP * kPBytes;
  const uint32_t bar_base = stage + kStageBytes;
  uint8_t* gen_small = gen_stage + kStageBytes;
  auto p_full = [&](int s) { return bar_base + 8u * s; };                                  // weight tile landed (TMA)
  auto p_empty = [&](int s) { return bar_base + 8u * (kNP + s); };                         // weight tile converted
  auto pconv_bar = [&](int s) { return bar_base + 8u * (2 * kNP + s); };                   // hi | lo in tensor memory
  auto t_empty = [&](int s) { return bar_base + 8u * (2 * kNP + kNT + s); };               // tensor-memory slot consumed
  auto q_full = [&](int s) { return bar_base + 8u * (2 * kNP + 2 * kNT + s); };                   // raw activation tile landed
  auto q_empty = [&](int s) { return bar_base + 8u * (2 * kNP + 2 * kNT + kNQ + s); };            // raw tile consumed
  auto qconv_bar = [&](int s) { return bar_base + 8u * (2 * kNP + 2 * kNT + 2 * kNQ + s); };      // lo tile written
  auto l_empty = [&](int s) { return bar_base + 8u * (2 * kNP + 2 * kNT + 2 * kNQ + kNL + s); };  // lo tile consumed
  const uint32_t accum_bar = bar_base + 8u * (2 * kNP + 2 * kNT + 2 * kNQ + 2 * kNL);
  auto attn_full = [&](int s) { return bar_base + 8u * (2 * kNP + 2 * kNT + 2 * kNQ + 2 * kNL + 1 + s); };
  auto xch_bar = [&](int r) { return bar_base + 8u * (2 * kNP + 2 * kNT + 2 * kNQ + 2 * kNL + 1 + kAttnBufs + r); };   // "partial tile of cluster rank r is staged"
  const uint32_t lo_base = stage + kNQ * kQBytes;        // lo ring behind the raw ring
  const uint32_t tmem_slot = bar_base + 8u * kNumBars;
  float* alpha_s = reinterpret_cast<float*>(gen_small + 8 * kNumBars + 16);      // [kAttnBatch][kAttnMaxN]
  int* js_s = reinterpret_cast<int*>(alpha_s + kAttnBatch * kAttnMaxN);            // [kAttnBatch] argmax
  float* wsel_s = reinterpret_cast<float*>(js_s + kAttnBatch);                     // [kAttnBatch] straight-through weight
  Job* jobs = reinterpret_cast<Job*>(wsel_s + kAttnBatch);                         // [kStepGemmPhases] this CTA's jobs
  int4* aitems = reinterpret_cast<int4*>(reinterpret_cast<uint8_t*>(jobs) + 256);  // [kAttnMaxItems] {sample, slice, vis, first unit}
  int4* aunits = aitems + kAttnMaxItems;                                           // [kAttnMaxUnits] {col, row, sample, item ordinal | last << 16}

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cta = blockIdx.x, G = gridDim.x;

  if (threadIdx.x == 0) {
    for (int s = 0; s < kNP; ++s) { mbar_init(p_full(s), 1); mbar_init(p_empty(s), 4); }
    for (int s = 0; s < kNT; ++s) { mbar_init(pconv_bar(s), 4); mbar_init(t_empty(s), 1); }
    for (int s = 0; s < kNQ; ++s) { mbar_init(q_full(s), 1); mbar_init(q_empty(s), 1); }
    for (int s = 0; s < kNL; ++s) { mbar_init(qconv_bar(s), kEpiWarps); mbar_init(l_empty(s), 1); }
    mbar_init(accum_bar, 1);
    for (int s = 0; s < kAttnBufs; ++s) mbar_init(attn_full(s), 1);
    for (int r = 0; r < kMaxSplit; ++r) mbar_init(xch_bar(r), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (threadIdx.x >= 32 && threadIdx.x < 32 + kStepGemmPhases) jobs[threadIdx.x - 32] = get_job(P, threadIdx.x - 32, cta);
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(kTmemCols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_small + 8 * kNumBars);
  const uint32_t crank = cluster_ctarank();
  if (cluster_nctarank() > 1) cluster_sync_all();   // the partners' exchange barriers are initialised

  // Barrier numbering: barrier g (0-based, counted over the whole launch) is complete when sync[0] >= (g + 1) * G.
  // Step s (= t - t0) owns barriers 7s .. 7s+6 = end of A, B, C1, C2, D, E, F.  A GEMM phase may read what every
  // earlier phase wrote once the barrier in front of it is complete; phase A of the first step depends on nothing
  // inside the launch.

  if (warp == 0) {
    // ============================== weight producer ==============================
    RingPos rp;
    for (int s = 0; s < P.nt; ++s) {
      for (int ph = 0; ph < kStepGemmPhases; ++ph) {
        const Job j = jobs[ph];
        if (j.prob < 0) continue;
        const StepProb& pr = P.prob[j.prob];
        for (int kb = j.kb_begin; kb < j.kb_end; ++kb) {
          const int sg = (kb < pr.nkb[0]) ? 0 : 1;
          const int kk = kb - (sg ? pr.nkb[0] : 0);
          if (rp.wrapped) mbar_wait_guarded(p_empty(rp.slot), rp.phase ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(p_full(rp.slot), kPBytes);
            const uint32_t dst = base + rp.slot * kPBytes;
            const int nb = pr.nblk;                       // 1, 2 or 4 row blocks of 128 / nb rows each
            for (int g = 0; g < nb; ++g)
              tma_load_2d(dst + g * (kPBytes / nb), &P.maps[pr.pmap[sg][nb == 2 ? g : 0]], p_full(rp.slot),
                          pr.pcol0[sg][nb == 2 ? g : 0] + kk * kBlockK, g * pr.blk_stride + j.tile * (kTileP / nb));
          }
          __syncwarp();
          rp.advance(kNP);
        }
      }
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    const uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kQN >> 3) << 17) | ((uint32_t)(kTileP >> 4) << 24);
    RingPos rt, rq, rl;
    for (int s = 0; s < P.nt; ++s) {
      for (int ph = 0; ph < kStepGemmPhases; ++ph) {
        const Job j = jobs[ph];
        if (j.prob < 0) continue;
        for (int kb = j.kb_begin; kb < j.kb_end; ++kb) {
          mbar_wait_guarded(pconv_bar(rt.slot), rt.phase);
          mbar_wait_guarded(qconv_bar(rl.slot), rl.phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            const uint32_t q_hi = stage + rq.slot * kQBytes, q_lo = lo_base + rl.slot * kQBytes;
            const uint64_t b_hi0 = umma_desc(q_hi, 16u, 1024u), b_lo0 = umma_desc(q_lo, 16u, 1024u);
            const uint32_t ta0 = tmem_base + (uint32_t)kQN + rt.slot * 64u;
#pragma unroll
            for (int k = 0; k < kBlockK / 8; ++k) {
              const uint64_t b_hi = b_hi0 + (uint64_t)(2 * k), b_lo = b_lo0 + (uint64_t)(2 * k);
              const uint32_t ta_hi = ta0 + (uint32_t)k * 8u;
              umma_tf32_ts(tmem_base, ta_hi + 32u, b_hi, idesc, (kb > j.kb_begin || k > 0) ? 1u : 0u);   // P_lo * Q_hi
              umma_tf32_ts(tmem_base, ta_hi, b_lo, idesc, 1u);                                           // P_hi * Q_lo
              umma_tf32_ts(tmem_base, ta_hi, b_hi, idesc, 1u);                                           // P_hi * Q_hi
            }
            umma_commit(t_empty(rt.slot));
            umma_commit(q_empty(rq.slot));
            umma_commit(l_empty(rl.slot));
            if (kb == j.kb_end - 1) umma_commit(accum_bar);
          }
          __syncwarp();
          rt.advance(kNT); rq.advance(kNQ); rl.advance(kNL);
        }
      }
    }
  } else if (warp == 2) {
    // ============================== activation producer ==============================
    RingPos rq;
    uint32_t gate_target = 0;   // arrivals that complete the barrier in front of the current phase
    for (int s = 0; s < P.nt; ++s) {
      const int t = P.t0 + s;
      for (int ph = 0; ph < kStepGemmPhases; ++ph) {
        // barriers completed before phase ph of step s: 7s + {0, 1, 4, 5, 6}[ph]
        const uint32_t nbar = (uint32_t)(kStepBarriersPerStep * s) + (ph < 2 ? (uint32_t)ph : (uint32_t)ph + 2u);
        gate_target = nbar * (uint32_t)G;
        const Job j = jobs[ph];
        if (j.prob < 0) continue;
        const StepProb& pr = P.prob[j.prob];
        if (nbar > 0) {
          if (lane == 0) { spin_until_ge(P.sync, gate_target); FS(s, ph, 0); }
          __syncwarp();
          fence_proxy_async_global();
          asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        }
        for (int kb = j.kb_begin; kb < j.kb_end; ++kb) {
          const int sg = (kb < pr.nkb[0]) ? 0 : 1;
          const int kk = kb - (sg ? pr.nkb[0] : 0);
          if (rq.wrapped) mbar_wait_guarded(q_empty(rq.slot), rq.phase ^ 1u);
          if (elect_one()) {
            mbar_expect_tx(q_full(rq.slot), kQBytes);
            tma_load_3d(stage + rq.slot * kQBytes, &P.maps[pr.qmap[sg]], q_full(rq.slot), pr.qcol0[sg] + kk * kBlockK, 0, t + pr.qtoff[sg]);
          }
          __syncwarp();
          rq.advance(kNQ);
        }
      }
    }
  } else if (warp >= kFirstConvWarp && warp < kFirstEpiWarp) {
    // ============================== weight converters ==============================
    // A landed fp32 weight tile goes to tensor memory as hi | lo (the MMA's A operand) as soon as a slot is free --
    // independent of the activations, so the tiles of the NEXT phase are converted while this phase still drains,
    // reduces and synchronises: up to kNT converted + kNP landed tiles wait on chip when a phase opens.
    // Two groups alternate K-blocks.  A group therefore sees only every other use of a slot; that is safe because each
    // wait names the exact use (parity from the global K-block position) and no barrier can run two phases ahead of its
    // waiter: the next load of a weight slot needs this conversion's p_empty arrival, the next MMA on a tensor-memory
    // slot needs this conversion's pconv arrival.
    const int quarter = warp & 3;
    const uint32_t grp = (uint32_t)(warp - kFirstConvWarp) >> 2;
    RingPos rp, rt;
    uint32_t seq = 0;
    for (int s = 0; s < P.nt; ++s) {
      for (int ph = 0; ph < kStepGemmPhases; ++ph) {
        const Job j = jobs[ph];
        if (j.prob < 0) continue;
        for (int kb = j.kb_begin; kb < j.kb_end; ++kb, ++seq) {
          if ((seq & (kConvGroups - 1)) != grp) { rp.advance(kNP); rt.advance(kNT); continue; }
          if (rt.wrapped) mbar_wait_guarded(t_empty(rt.slot), rt.phase ^ 1u);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          mbar_wait_guarded(p_full(rp.slot), rp.phase);
          const float4* p_raw = reinterpret_cast<const float4*>(gen_base + rp.slot * kPBytes);
          const int prow = quarter * 32 + lane;
          const uint32_t ta = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)kQN + rt.slot * 64u;
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            uint32_t hi[16], lo[16];
#pragma unroll
            for (int cc = 0; cc < 4; ++cc) {
              const int cch = half * 4 + cc;
              const float4 v = p_raw[prow * 8 + (cch ^ (prow & 7))];
              hi[cc * 4 + 0] = __float_as_uint(v.x); lo[cc * 4 + 0] = __float_as_uint(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u));
              hi[cc * 4 + 1] = __float_as_uint(v.y); lo[cc * 4 + 1] = __float_as_uint(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u));
              hi[cc * 4 + 2] = __float_as_uint(v.z); lo[cc * 4 + 2] = __float_as_uint(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u));
              hi[cc * 4 + 3] = __float_as_uint(v.w); lo[cc * 4 + 3] = __float_as_uint(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u));
            }
            tmem_st16(ta + (uint32_t)half * 16u, hi);
            tmem_st16(ta + 32u + (uint32_t)half * 16u, lo);
          }
          asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) { mbar_arrive(pconv_bar(rt.slot)); mbar_arrive(p_empty(rp.slot)); }
          rp.advance(kNP); rt.advance(kNT);
        }
      }
    }
  } else if (warp >= kFirstEpiWarp) {
    // ============================== Q lo-split / epilogue / attention ==============================
    const int ct = threadIdx.x - 32 * kFirstEpiWarp;   // 0 .. 255
    const int ew = warp - kFirstEpiWarp;               // 0 .. 7
    const int quarter = warp & 3;                      // tensor-memory lane quarter of this warp
    float* ep = reinterpret_cast<float*>(gen_stage);
    RingPos rq, rl;
    uint32_t jobseq = 0, barseq = 0;
    uint32_t xph = 0;                            // bit r: parity of the next completion of xch_bar(r)
    uint32_t attn_units_done = 0;                // value chunks consumed since launch (ring slot / parity)
    const StepAttn& AT = P.attn;
    const int D = P.D;

    // ---- attention work split (phase C2): value-chunk items ordered [visual (i, slice) ..., caption (i, slice) ...]
    // cost-balanced contiguous ranges (cost = rows of the item)
    auto attn_range = [&](int b, int& v0, int& v1, int& c0, int& c1) {
      const int sv = AT.F / kAttnCols, scn = AT.D / kAttnCols;
      const unsigned nv = (unsigned)(b * sv), nc = (unsigned)(b * scn);
      const unsigned cv = (unsigned)AT.R, cc = (unsigned)AT.P + 1u;
      const unsigned tot = nv * cv + nc * cc;                 // <= 64 * (8 * 128 + 4 * 129): fits 32 bits with room
      const unsigned lo = tot * (unsigned)cta / (unsigned)G, hi = tot * (unsigned)(cta + 1) / (unsigned)G;
      auto first_at = [&](unsigned x) -> unsigned {   // number of items whose cost-start lies before x
        if (x <= nv * cv) return (x + cv - 1u) / cv;
        return nv + (x - nv * cv + cc - 1u) / cc;
      };
      const unsigned a = first_at(lo), e = first_at(hi);
      v0 = (int)(a < nv ? a : nv); v1 = (int)(e < nv ? e : nv);
      c0 = (int)(a > nv ? a - nv : 0); c1 = (int)(e > nv ? e - nv : 0);
    };
    // this CTA's C2 work as tables in shared memory (rebuilt when the decoded batch size changes): items = (sample,
    // column slice) of one attention, units = the value chunks (<= 36 rows x 256 columns) they stream through the ring
    auto build_attn_tables = [&](int v0, int v1, int c0, int c1, int& nitems, int& nunits) {
      const int nchv = (AT.R + AT.chunk_v - 1) / AT.chunk_v, nchc = (AT.P + AT.chunk_c - 1) / AT.chunk_c;
      const int nvi = v1 - v0;
      nitems = nvi + (c1 - c0);
      nunits = nvi * nchv + (c1 - c0) * nchc;
      for (int k = ct; k < nitems; k += kEpiThreads) {
        const bool vis = k < nvi;
        const int item = vis ? v0 + k : c0 + (k - nvi);
        const int nsl = (vis ? AT.F : AT.D) / kAttnCols;
        aitems[k] = make_int4(item / nsl, item % nsl, vis ? 1 : 0, vis ? k * nchv : nvi * nchv + (k - nvi) * nchc);
      }
      for (int u = ct; u < nunits; u += kEpiThreads) {
        const bool vis = u < nvi * nchv;
        const int uu = vis ? u : u - nvi * nchv;
        const int nch = vis ? nchv : nchc;
        const int k = vis ? uu / nch : nvi + uu / nch, ch = uu % nch;
        const int item = vis ? v0 + k : c0 + (k - nvi);
        const int nsl = (vis ? AT.F : AT.D) / kAttnCols;
        aunits[u] = make_int4((item % nsl) * kAttnCols, ch * (vis ? AT.chunk_v : AT.chunk_c), item / nsl,
                              k | (ch == nch - 1 ? 1 << 16 : 0) | (vis ? 1 << 17 : 0));
      }
    };
    auto issue_attn_unit = [&](int u, uint32_t unit_seq) {
      const int4 e = aunits[u];
      const bool vis = (e.w >> 17) & 1;
      const uint32_t slot = unit_seq % kAttnBufs;
      mbar_expect_tx(attn_full(slot), (uint32_t)((vis ? AT.chunk_v : AT.chunk_c) * kAttnCols * 4));
      tma_load_3d(stage + slot * kAttnBuf, &P.maps[vis ? AT.map_feats : AT.map_prevh], attn_full(slot), e.x, e.y, e.z);
    };
    int nitems = 0, nunits = 0, b_tables = -1;

    for (int s = 0; s < P.nt; ++s) {
      const int t = P.t0 + s;
      const int b = P.bt[s];
      const int rows = b < kQN ? b : kQN;
      for (int p7 = 0; p7 < kStepBarriersPerStep; ++p7) {
        if (p7 != 2 && p7 != 3) {
          // ================= GEMM phase (A, B, D, E, F) =================
          const int ph = p7 < 2 ? p7 : p7 - 2;
          const Job j = jobs[ph];
          if (j.prob >= 0) {
            const StepProb& pr = P.prob[j.prob];
            const int nkb = j.kb_end - j.kb_begin;
            for (int i = 0; i < nkb; ++i) {
              // the activation tile: raw words serve as "hi" (kind::tf32 ignores the low mantissa bits), lo = x - hi
              // goes to the sibling tile
              mbar_wait_guarded(q_full(rq.slot), rq.phase);
              if (i == 0 && ct == 0) FS(s, ph, 1);
              if (i == nkb - 1 && ct == 0) FS(s, ph, 3);
              if (rl.wrapped) mbar_wait_guarded(l_empty(rl.slot), rl.phase ^ 1u);
              const float4* q_hi = reinterpret_cast<const float4*>(gen_stage + rq.slot * kQBytes);
              float4* q_lo = reinterpret_cast<float4*>(gen_stage + (kNQ + rl.slot) * kQBytes);
#pragma unroll
              for (int jj = 0; jj < kQBytes / 16 / kEpiThreads; ++jj) {
                const float4 v = q_hi[ct + kEpiThreads * jj];
                float4 l;
                l.x = v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u);
                l.y = v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u);
                l.z = v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u);
                l.w = v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u);
                q_lo[ct + kEpiThreads * jj] = l;
              }
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes -> visible to the MMA (async proxy)
              __syncwarp();
              if (lane == 0) mbar_arrive(qconv_bar(rl.slot));
              rq.advance(kNQ); rl.advance(kNL);
            }
            // ---- accumulator -> registers -> shared (staged as [q][p]; the Q ring is idle now)
            mbar_wait_guarded(accum_bar, jobseq & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            ++jobseq;
            if (ct == 0) FS(s, ph, 4);
            {
              const int prow = quarter * 32 + lane;
              const int cb = ew >> 2;                    // warps 0..3: columns 0..31, warps 4..7: columns 32..63
              uint32_t r[32];
              tmem_ld32(tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(cb * 32), r);
#pragma unroll
              for (int x = 0; x < 32; ++x) ep[(cb * 32 + x) * kEpPitch + prow] = __uint_as_float(r[x]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            EPI_BAR();
            if (ct == 0) FS(s, ph, 5);
            // ---- split-K: the partners of a tile are CTAs of one cluster; each tells the others (remote mbarrier
            // arrive) that its partial tile is staged, then reads theirs through distributed shared memory
            const int split = j.split;
            const uint32_t g0 = crank - (uint32_t)j.ks;        // cluster rank of the tile's first CTA
            uint32_t peer[kMaxSplit];
#pragma unroll
            for (int k2 = 0; k2 < kMaxSplit; ++k2) peer[k2] = dsmem_addr(smem_u32(ep), g0 + (uint32_t)(k2 < split ? k2 : 0));
            if (split > 1) {
              if (ct == 0) {
                // Relaxed arrives, no cluster-scope fence (1.3 us measured): the staged tile was written with st.shared
                // by threads that have all passed the bar.sync above, which drains their shared-memory stores -- the
                // data sits in this SM's shared memory before the arrive leaves it, and the partners only ever read it
                // there (ld.shared::cluster after an acquire wait).
                for (int k2 = 0; k2 < split; ++k2)
                  if (k2 != j.ks) mbar_arrive_remote_relaxed(xch_bar((int)crank), g0 + (uint32_t)k2);
                FS(s, ph, 6);
              }
              for (int k2 = 0; k2 < split; ++k2) {
                if (k2 == j.ks) continue;
                const uint32_t r = g0 + (uint32_t)k2;
                mbar_wait_cluster_guarded(xch_bar((int)r), (xph >> r) & 1u);
                xph ^= (1u << r);
              }
              if (ct == 0) FS(s, ph, 7);
            }
            // ---- finish 1/split of the tile.  One loop for every cell.  An item is (batch row q, two consecutive units
            // u, u+1) of the tile's nb row blocks (1: plain columns, 2: [context gate | sc_affine], 4: the four gates);
            // block g of an item is output column g * blk_stride + tile * (128 / nb) + u.  The finish is a chain of L2 /
            // DSMEM round trips, so a thread takes 4 / nb items per pass and requests everything the pass needs -- four
            // (item, block) "slots" of partials and addends plus the cells' own operands -- before it consumes anything.
            {
              const int nb = pr.nblk;
              const int lnb = nb == 1 ? 0 : (nb == 2 ? 1 : 2);
              const int ub = kTileP >> lnb;               // units per block
              const int sh = 6 - lnb;                     // log2(items per batch row) = log2(ub / 2)
              const int ipb = 4 >> lnb;                   // items per pass
              const int items = rows << sh;
              const int i0 = (int)((unsigned)(items * j.ks) / (unsigned)split), i1 = (int)((unsigned)(items * (j.ks + 1)) / (unsigned)split);
              const int epi = pr.epi;
              float* Ct = pr.C.p ? tp(pr.C, t) : nullptr;
              const float* addt = pr.add.p ? tp(pr.add, t) : nullptr;
              // which blocks take which addend (bit = block): the context-gate tile adds `add` to block 0, `bias2` to block 1
              const uint32_t m_add = addt ? (epi == kSEpiCtxGate ? 1u : 15u) : 0u;
              const uint32_t m_b1 = pr.bias ? 15u : 0u;
              const uint32_t m_b2 = pr.bias2 ? (epi == kSEpiCtxGate ? 2u : 15u) : 0u;
              const uint32_t m_old = (pr.beta && Ct) ? 15u : 0u;
              const float2 z2 = make_float2(0.f, 0.f);
              for (int it0 = i0 + ct; it0 < i1; it0 += ipb * kEpiThreads) {
                float2 pre[4];
                int qs[4], us[4];
                bool on[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                  const int blk = g & (nb - 1);
                  const int it = it0 + (g >> lnb) * kEpiThreads;
                  const int q = it >> sh, u = (it & ((1 << sh) - 1)) * 2;
                  const int unit = j.tile * ub + u;
                  on[g] = it < i1 && unit < pr.N;
                  qs[g] = q; us[g] = unit;
                  const uint32_t off = (uint32_t)(q * kEpPitch + blk * ub + u) * 4u;
                  float2 acc = z2;
#pragma unroll
                  for (int k2 = 0; k2 < kMaxSplit; ++k2) {
                    const float2 v = dsmem_ld2(peer[k2] + off, on[g] && k2 < split);
                    acc.x += v.x; acc.y += v.y;
                  }
                  const int n = blk * pr.blk_stride + unit;
                  const float2 b1 = (on[g] && ((m_b1 >> blk) & 1u)) ? __ldg(reinterpret_cast<const float2*>(pr.bias + n)) : z2;
                  const float2 b2 = (on[g] && ((m_b2 >> blk) & 1u)) ? __ldg(reinterpret_cast<const float2*>(pr.bias2 + n)) : z2;
                  const float2 ad = (on[g] && ((m_add >> blk) & 1u)) ? ldcg2(addt + (long)q * pr.ldadd + n) : z2;
                  const float2 od = (on[g] && ((m_old >> blk) & 1u)) ? ldcg2(Ct + (long)q * pr.ldc + n) : z2;
                  acc.x += b1.x + b2.x + ad.x + od.x; acc.y += b1.y + b2.y + ad.y + od.y;
                  pre[g] = acc;
                }
                if (epi == kSEpiPlain) {
#pragma unroll
                  for (int g = 0; g < 4; ++g)
                    if (on[g]) *reinterpret_cast<float2*>(Ct + (long)qs[g] * pr.ldc + us[g]) = pre[g];
                } else if (epi == kSEpiLstm || epi == kSEpiCopy1) {
                  // nn.LSTMCell (editnet.py:532) / copy-LSTM stage 1 (:272-279): gates i, f, g, o; c = f c_prev + i g
                  if (on[0]) {
                    const int q = qs[0], unit = us[0];
                    const long x = (long)q * D + unit;
                    const float2 cp = ldcg2(tp(pr.a0, t) + x);
                    float2 gi, gf, gg, go, c;
                    gi.x = sig_(pre[0].x); gi.y = sig_(pre[0].y);
                    gf.x = sig_(pre[1].x); gf.y = sig_(pre[1].y);
                    gg.x = tanh_(pre[2].x); gg.y = tanh_(pre[2].y);
                    go.x = sig_(pre[3].x); go.y = sig_(pre[3].y);
                    c.x = gf.x * cp.x + gi.x * gg.x; c.y = gf.y * cp.y + gi.y * gg.y;
                    float* gp = Ct + (long)q * pr.ldc + unit;
                    *reinterpret_cast<float2*>(gp) = gi; *reinterpret_cast<float2*>(gp + D) = gf;
                    *reinterpret_cast<float2*>(gp + 2 * D) = gg; *reinterpret_cast<float2*>(gp + 3 * D) = go;
                    *reinterpret_cast<float2*>(tp(pr.a1, t) + x) = c;
                    if (epi == kSEpiLstm) {
                      float2 h;
                      h.x = go.x * tanh_(c.x); h.y = go.y * tanh_(c.y);
                      *reinterpret_cast<float2*>(tp(pr.a2, t) + (long)q * pr.ld0 + unit) = h;
                    }
                  }
                } else if (epi == kSEpiCtxGate) {
                  // editnet.py:378-380: z = sigmoid(gate), att_cap = z tanh(sc) + (1 - z) tanh(tc); slots (0,1), (2,3)
                  float2 tcp[2];
#pragma unroll
                  for (int k = 0; k < 2; ++k)
                    tcp[k] = on[2 * k] ? ldcg2(tp(pr.a0, t) + (long)qs[2 * k] * pr.ld0 + us[2 * k]) : z2;
#pragma unroll
                  for (int k = 0; k < 2; ++k) {
                    if (!on[2 * k]) continue;
                    const int q = qs[2 * k], unit = us[2 * k];
                    float2 z, ts, tt, o;
                    z.x = sig_(pre[2 * k].x); z.y = sig_(pre[2 * k].y);
                    ts.x = tanh_(pre[2 * k + 1].x); ts.y = tanh_(pre[2 * k + 1].y);
                    tt.x = tanh_(tcp[k].x); tt.y = tanh_(tcp[k].y);
                    o.x = z.x * ts.x + (1.f - z.x) * tt.x; o.y = z.y * ts.y + (1.f - z.y) * tt.y;
                    float* zr = tp(pr.a1, t) + (long)q * 3 * D + unit;
                    *reinterpret_cast<float2*>(zr) = z;
                    *reinterpret_cast<float2*>(zr + D) = ts;
                    *reinterpret_cast<float2*>(zr + 2 * D) = tt;
                    *reinterpret_cast<float2*>(tp(pr.a2, t) + (long)q * pr.ld1 + unit) = o;
                  }
                } else {
                  // copy gate (editnet.py:281-285): k = sigmoid(pre); c2 = k sel + (1-k) c_new; h2 = o tanh(c2)
                  float2 sel[4], cn[4], og[4];
#pragma unroll
                  for (int g = 0; g < 4; ++g) {
                    const long x = (long)qs[g] * D + us[g];
                    sel[g] = on[g] ? ldcg2(tp(pr.a1, t) + x) : z2;
                    cn[g] = on[g] ? ldcg2(tp(pr.a2, t) + x) : z2;
                    og[g] = on[g] ? ldcg2(tp(pr.a0, t) + (long)qs[g] * pr.ld0 + 3 * D + us[g]) : z2;
                  }
#pragma unroll
                  for (int g = 0; g < 4; ++g) {
                    if (!on[g]) continue;
                    const long x = (long)qs[g] * D + us[g];
                    float2 k, c, h;
                    k.x = sig_(pre[g].x); k.y = sig_(pre[g].y);
                    c.x = k.x * sel[g].x + (1.f - k.x) * cn[g].x; c.y = k.y * sel[g].y + (1.f - k.y) * cn[g].y;
                    h.x = og[g].x * tanh_(c.x); h.y = og[g].y * tanh_(c.y);
                    *reinterpret_cast<float2*>(tp(pr.a3, t) + x) = k;
                    *reinterpret_cast<float2*>(tp(pr.a4, t) + x) = c;
                    *reinterpret_cast<float2*>(tp(pr.a5, t) + x) = h;
                    float2 hd = h;
                    if (P.train) {
                      const uint64_t idx = (uint64_t)((long)t * P.B * D + x);
                      const uint32_t keep = drop_keep4(P.seed, kSiteFc, idx & ~3ull) >> (idx & 2ull);
                      hd.x = (keep & 1u) ? h.x * 2.f : 0.f; hd.y = (keep & 2u) ? h.y * 2.f : 0.f;
                    }
                    *reinterpret_cast<float2*>(tp(pr.a6, t) + x) = hd;
                  }
                }
              }
            }
            if (ct == 0) FS(s, ph, 8);
          }
        } else if (p7 == 2) {
          // ================= phase C1: raw attention scores =================
          // The value rows of the attention are constants: this CTA's first chunks are requested now (the staging
          // region is free: the barrier behind phase B is complete), i.e. before the scores exist.
          if (ct == 0) FS(s, 1, 11);
          if (b != b_tables) {
            int v0, v1, c0, c1;
            attn_range(b, v0, v1, c0, c1);
            build_attn_tables(v0, v1, c0, c1, nitems, nunits);
            b_tables = b;
            EPI_BAR();
          }
          if (ct == 0) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int u = 0; u < nunits && u < kAttnBufs; ++u) issue_attn_unit(u, attn_units_done + u);
          }
          if (ct == 0) FS(s, 1, 12);
          const int nv = b * AT.R, nc = b * AT.P, total = nv + nc;
          const int per = (total + G - 1) / G;
          const int r0 = cta * per, r1 = (r0 + per < total) ? r0 + per : total;
          const int A4 = AT.A >> 2;
          const float* s2t = tp(AT.s2, t);
          const float* att1v = tp(AT.att1v, t);
          // One (sample, row) per warp and pass, two passes in flight (compact code: the whole phase is a few hundred
          // instructions; the rows were pulled into L2 by the spare warp while phase B ran).
#pragma unroll 1
          for (int rb = r0 + ew; rb < r1; rb += 2 * kEpiWarps) {
            float acc[2];
            const float* a1p[2]; const float* a2p[2]; const float* wp[2]; bool tanh_sel[2]; bool on[2];
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int r = rb + k * kEpiWarps;
              on[k] = r < r1;
              const int rr = on[k] ? r : r0;
              const bool vis = rr < nv;
              const int x = vis ? rr : rr - nv;
              const int i = vis ? x / AT.R : x / AT.P;
              a1p[k] = (vis ? att1v : AT.att1c) + (long)x * AT.A;
              a2p[k] = s2t + (long)i * AT.ld_s2 + (vis ? AT.A : 0);
              wp[k] = vis ? AT.vis_w : AT.cap_w;
              tanh_sel[k] = !vis;
              acc[k] = 0.f;
            }
#pragma unroll 1
            for (int x4 = lane; x4 < A4; x4 += 128) {
              float4 vv[2][4], bb[2][4];
#pragma unroll
              for (int k = 0; k < 2; ++k)
#pragma unroll
                for (int m = 0; m < 4; ++m) {
                  const bool ok = on[k] && x4 + 32 * m < A4;
                  vv[k][m] = ok ? __ldg(reinterpret_cast<const float4*>(a1p[k]) + x4 + 32 * m) : make_float4(0.f, 0.f, 0.f, 0.f);
                  bb[k][m] = ok ? __ldcg(reinterpret_cast<const float4*>(a2p[k]) + x4 + 32 * m) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
#pragma unroll
              for (int k = 0; k < 2; ++k)
#pragma unroll 1
                for (int m = 0; m < 4; ++m) {
                  if (!(on[k] && x4 + 32 * m < A4)) continue;
                  const float4 w = __ldg(reinterpret_cast<const float4*>(wp[k]) + x4 + 32 * m);
                  // (dynamic m: select the register with a small switch instead of spilling the arrays)
                  float4 y = m == 0 ? vv[k][0] : (m == 1 ? vv[k][1] : (m == 2 ? vv[k][2] : vv[k][3]));
                  const float4 a2 = m == 0 ? bb[k][0] : (m == 1 ? bb[k][1] : (m == 2 ? bb[k][2] : bb[k][3]));
                  y.x += a2.x; y.y += a2.y; y.z += a2.z; y.w += a2.w;
                  // tanh for the caption attention (editnet.py:372), relu for the visual one (:444)
                  if (tanh_sel[k]) { y.x = tanh_(y.x); y.y = tanh_(y.y); y.z = tanh_(y.z); y.w = tanh_(y.w); }
                  else { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
                  acc[k] += w.x * y.x + w.y * y.y + w.z * y.z + w.w * y.w;
                }
            }
            if (ct == 0) FS(s, 1, (rb == r0 ? 13 : 14));
#pragma unroll
            for (int k = 0; k < 2; ++k) {
              const int r = rb + k * kEpiWarps;
              const float sv = warp_sum(acc[k]);
              if (on[k] && lane == 0) {
                const bool vis = r < nv;
                const int x = vis ? r : r - nv;
                const int n = vis ? AT.R : AT.P;
                const int i = x / n, jr = x - i * n;
                AT.sc[(long)i * (AT.P + AT.R) + (vis ? AT.P : 0) + jr] = sv + __ldg(vis ? AT.vis_b : AT.cap_b);
              }
            }
          }
          if (ct == 0) FS(s, 1, 15);
        } else {
          // ================= phase C2: softmax, contexts, select =================
          int u = 0;                     // units consumed in this phase
          for (int ib = 0; ib < nitems; ib += kAttnBatch) {
            // one softmax per warp for the next kAttnBatch items
            EPI_BAR();                   // the previous batch's alpha_s readers are done
            if (ib + ew < nitems) {
              const int4 itm = aitems[ib + ew];
              const int i = itm.x, sl = itm.y;
              const bool vis = itm.z != 0;
              const int n = vis ? AT.R : AT.P;
              const float* scr = AT.sc + (long)i * (AT.P + AT.R) + (vis ? AT.P : 0);
              const int nvalid = vis ? (AT.nreg ? AT.nreg[i] : n) : n;
              float vals[kAttnMaxN / 32];
              float m = -INFINITY;
#pragma unroll
              for (int x = 0; x < kAttnMaxN / 32; ++x) {
                const int jr = lane + 32 * x;
                float v = -INFINITY;
                if (jr < n) {
                  v = __ldcg(scr + jr);
                  if (vis) { if (jr >= nvalid) v = kNegFill; }
                  else if (__ldg(AT.mask + (long)i * AT.P + jr) == 0.f) v = kNegFill;
                }
                vals[x] = v;
                m = fmaxf(m, v);
              }
              m = warp_max(m);
              float sum = 0.f;
#pragma unroll
              for (int x = 0; x < kAttnMaxN / 32; ++x) {
                const int jr = lane + 32 * x;
                const float e = (jr < n) ? expf(vals[x] - m) : 0.f;
                vals[x] = e;
                sum += e;
              }
              sum = warp_sum(sum);
              float best = -1.f; int bj = 0;
              float* aout = vis ? tp(AT.alpha_v, t) + (long)i * AT.R : tp(AT.alpha_c, t) + (long)i * AT.P;
#pragma unroll
              for (int x = 0; x < kAttnMaxN / 32; ++x) {
                const int jr = lane + 32 * x;
                if (jr < n) {
                  const float al = vals[x] / sum;
                  alpha_s[ew * kAttnMaxN + jr] = al;
                  if (al > best) { best = al; bj = jr; }      // first maximum within the lane (ascending jr)
                  if (sl == 0) aout[jr] = al;
                }
              }
              if (!vis) {
                // argmax over the row, first occurrence (torch.max semantics of SelectC, editnet.py:410)
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                  const float ob = __shfl_xor_sync(0xffffffffu, best, o);
                  const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
                  if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
                }
                if (lane == 0) {
                  js_s[ew] = bj;
                  wsel_s[ew] = best + (1.f - best);
                  if (sl == 0) (AT.sel_idx + (long)t * AT.sel_idx_st)[i] = bj;
                }
              }
            }
            EPI_BAR();
            const int u_end = (ib + kAttnBatch < nitems) ? aitems[ib + kAttnBatch].w : nunits;
            float acc = 0.f;               // column `ct` of the current item, accumulated over its row chunks
            for (; u < u_end; ++u) {
              const int4 e = aunits[u];
              const int k = e.w & 0xffff;
              const bool vis = (e.w >> 17) & 1;
              const int n = vis ? AT.R : AT.P;
              const int chunk = vis ? AT.chunk_v : AT.chunk_c;
              const float* al = alpha_s + (k - ib) * kAttnMaxN + e.y;
              const uint32_t useq = attn_units_done + (uint32_t)u;
              const uint32_t slot = useq % kAttnBufs;
              mbar_wait_guarded(attn_full(slot), (useq / kAttnBufs) & 1u);
              const float* buf = reinterpret_cast<const float*>(gen_stage + slot * kAttnBuf);
              const int rmax = (n - e.y < chunk) ? n - e.y : chunk;
#pragma unroll 4
              for (int r = 0; r < rmax; ++r) acc += al[r] * buf[r * kAttnCols + ct];
              EPI_BAR();               // everyone is done with the buffer: refill it with the unit 3 ahead
              if (ct == 0 && u + kAttnBufs < nunits) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                issue_attn_unit(u + kAttnBufs, useq + kAttnBufs);
              }
              if ((e.w >> 16) & 1) {     // last chunk of the item
                float* out = vis ? tp(AT.att_img, t) + (long)e.z * AT.ld_img : tp(AT.ctx, t) + (long)e.z * AT.D;
                out[e.x + ct] = acc;
                if (!vis) {
                  // select: the memory row at the argmax, straight-through weight alpha + (1 - alpha) (editnet.py:410-420)
                  (tp(AT.sel, t) + (long)e.z * AT.D)[e.x + ct] =
                      wsel_s[k - ib] * __ldg(AT.prev_m + ((long)e.z * AT.P + js_s[k - ib]) * AT.D + e.x + ct);
                }
                acc = 0.f;
              }
            }
          }
          attn_units_done += (uint32_t)nunits;
        }
        // ================= grid barrier: arrive (after this CTA's writes of the phase), wait for everybody =========
        if (P.trace && ct == 0) P.trace[((long)s * 8 + p7) * G + cta] = gtimer();
        fence_proxy_async_global();
        EPI_BAR();
        if (ct == 0) {
          // (red.release.gpu is cumulative over the CTA's writes ordered by the barrier above)
          red_release_gpu_add(P.sync, 1u);
          spin_until_ge(P.sync, (barseq + 1u) * (uint32_t)G);
        }
        ++barseq;
        EPI_BAR();
      }
    }
    // ---- re-arm the counters for the next launch: the last CTA out zeroes them
    if (ct == 0) {
      const unsigned old = atomicAdd(P.sync + 1, 1u);
      if (old == (unsigned)G - 1u) {
        P.sync[0] = 0u;
        __threadfence();
        P.sync[1] = 0u;
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (cluster_nctarank() > 1) cluster_sync_all();   // no CTA leaves while a partner may still address its shared memory
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------ host side
namespace {
struct StepDevice {
  std::once_flag once;
  bool ok = false;
  int grid = 0;        // CTAs of the persistent launch (a multiple of the cluster size, all co-resident)
  int cluster = 0;     // CTAs per cluster = maximum split-K fan-in
};
StepDevice g_step_dev[64];
unsigned long long* g_step_trace = nullptr;

void step_init(StepDevice* d, int dev) {
  if (cudaFuncSetAttribute(step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes) != cudaSuccess) { cudaGetLastError(); return; }
  int coop = 0, sms = 0;
  if (cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev) != cudaSuccess || !coop) { cudaGetLastError(); return; }
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess) { cudaGetLastError(); return; }
  // the largest cluster size (4, then 2) whose co-resident clusters cover at least 128 CTAs
  static const int want = getenv("SET_STEP_CLUSTER") ? atoi(getenv("SET_STEP_CLUSTER")) : 4;
  for (int c = (want >= 4 ? 4 : 2); c >= 2; c >>= 1) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((sms / c) * c); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = c; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, step_kernel, &cfg) != cudaSuccess) { cudaGetLastError(); nc = 0; }
    if (getenv("SET_TC_VERBOSE")) fprintf(stderr, "libset_b200: step kernel: %d co-resident clusters of %d CTAs\n", nc, c);
    if (nc * c >= 128 || (c == 2 && nc * c >= 64)) { d->cluster = c; d->grid = nc * c < sms ? nc * c : (sms / c) * c; break; }
  }
  if (d->cluster == 0) return;
  d->ok = true;
}
}  // namespace

bool step_kernel_available(int* grid, int* cluster) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) { cudaGetLastError(); return false; }
  StepDevice* d = &g_step_dev[dev];
  std::call_once(d->once, step_init, d, dev);
  if (grid) *grid = d->grid;
  if (cluster) *cluster = d->cluster;
  return d->ok;
}

void step_set_trace(unsigned long long* buf) { g_step_trace = buf; }
long long g_step_launches = 0, g_step_steps = 0;

// One job (tile x K-split) per CTA and phase; the split partners of a tile are consecutive CTAs of one cluster, so
// splits are powers of two up to the cluster size.  Greedy: the problem whose CTAs carry the most K-blocks doubles its
// split while the phase still fits the grid.  Problems are laid out by descending split so that every tile starts at
// a cluster rank that is a multiple of its split.
int step_plan_splits(StepParams& prm, int grid, int cluster) {
  for (int ph = 0; ph < kStepGemmPhases; ++ph) {
    StepPhase& phs = prm.phase[ph];
    SET_REQUIRE(phs.nprob >= 1 && phs.nprob <= kStepMaxPhaseProbs, "problems per phase");
    int ctas = 0;
    for (int k = 0; k < phs.nprob; ++k) {
      StepProb& p = prm.prob[phs.prob[k]];
      const int units_per_tile = p.nblk == 4 ? 32 : (p.nblk == 2 ? 64 : kTileP);
      p.tiles = (p.N + units_per_tile - 1) / units_per_tile;
      p.split = 1;
      ctas += p.tiles;
    }
    SET_REQUIRE(ctas <= grid, "a decode-step phase has more tiles than the persistent grid has CTAs");
    for (;;) {
      int best = -1; double best_load = 0.0;
      for (int k = 0; k < phs.nprob; ++k) {
        const StepProb& p = prm.prob[phs.prob[k]];
        const int nkb = p.nkb[0] + (p.nseg > 1 ? p.nkb[1] : 0);
        if (ctas + p.tiles * p.split > grid || p.split * 2 > cluster || p.split * 2 > kMaxSplit || nkb / (p.split * 2) < 2) continue;
        const double load = (double)nkb / p.split;
        if (load > best_load) { best_load = load; best = k; }
      }
      if (best < 0) break;
      StepProb& p = prm.prob[phs.prob[best]];
      ctas += p.tiles * p.split;
      p.split *= 2;
    }
    // order by descending split (stable), assign CTA ranges
    for (int a = 1; a < phs.nprob; ++a)
      for (int b2 = a; b2 > 0 && prm.prob[phs.prob[b2]].split > prm.prob[phs.prob[b2 - 1]].split; --b2) {
        const int tmp = phs.prob[b2]; phs.prob[b2] = phs.prob[b2 - 1]; phs.prob[b2 - 1] = tmp;
      }
    int c = 0;
    for (int k = 0; k < phs.nprob; ++k) {
      StepProb& p = prm.prob[phs.prob[k]];
      SET_REQUIRE(c % p.split == 0, "split alignment");
      p.cta0 = c;
      c += p.tiles * p.split;
    }
    phs.ncta = c;
    if (getenv("SET_TC_VERBOSE")) {
      fprintf(stderr, "libset_b200: step phase %d: %d CTAs:", ph, c);
      for (int k = 0; k < phs.nprob; ++k) {
        const StepProb& p = prm.prob[phs.prob[k]];
        fprintf(stderr, " [%d tiles x split %d, %d kb]", p.tiles, p.split, (p.nkb[0] + (p.nseg > 1 ? p.nkb[1] : 0)) / p.split);
      }
      fprintf(stderr, "\n");
    }
  }
  return SET_OK;
}

int step_launch(StepParams& prm, int grid, int cluster, cudaStream_t stream) {
  SET_REQUIRE(prm.nt >= 1 && prm.nt <= 64, "1..64 timesteps per launch");
  prm.sync = static_cast<unsigned*>(lib_scratch(kScratchStepBarrier, stream, sizeof(unsigned) * 16, true));
  if (!prm.sync) return SET_ERR_CUDA;
  prm.slabs = nullptr;
  prm.trace = g_step_trace;
  prm.trace_cta = getenv("SET_STEP_TRACE_CTA") ? atoi(getenv("SET_STEP_TRACE_CTA")) : 0;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(kThreads); cfg.dynamicSmemBytes = kSmemBytes; cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeCooperative;
  attr[1].val.cooperative = 1;
  // Cooperative launch: the runtime guarantees that all CTAs are co-resident (the grid barriers spin).  SET_STEP_COOP=0
  // drops the attribute (co-residency then rests on the grid being the number of clusters the idle device holds).
  static const int coop = getenv("SET_STEP_COOP") ? atoi(getenv("SET_STEP_COOP")) : 1;
  cfg.attrs = attr; cfg.numAttrs = coop ? 2 : 1;
  SET_CHECK_CUDA(cudaLaunchKernelEx(&cfg, step_kernel, prm));
  set_count_launch(1);
  ++g_step_launches;
  g_step_steps += prm.nt;
  return SET_OK;
}

}  // namespace set
