// Persistent decode-step kernel of the EditNet path: ONE cooperative launch runs `nt` consecutive timesteps
// (editnet.py:527-543) on all SMs.  See step_kernel.cu for the design.
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace set {

constexpr int kStepMaxMaps = 24;
constexpr int kStepMaxProbs = 12;
constexpr int kStepGemmPhases = 5;     // A, B, D, E, F (the attention phases C1, C2 sit between B and D)
constexpr int kStepMaxPhaseProbs = 5;
constexpr int kStepBarriersPerStep = 7;   // A, B, C1, C2, D, E, F

// pointer that advances with the timestep: address at step t = p + t * st (floats)
struct TPtr { float* p; long st; };

enum StepEpi { kSEpiPlain = 0, kSEpiLstm = 1, kSEpiCtxGate = 2, kSEpiCopy1 = 3, kSEpiCopy2 = 4 };

// One GEMM problem of a phase in "swap" form: the weight matrix supplies the 128-row P tiles, the batch (<= 64 rows)
// is the single Q tile.  out[q][n] = sum_k act[q][k] * W[n][k] (+ bias + add + old C), then the cell `epi`.
struct StepProb {
  int nseg;
  int nkb[2];          // K-blocks (32 fp32) per segment
  int pmap[2][2];      // [seg][block]: tensor map of the weight operand (block 1 only for 2-block tiles)
  int pcol0[2][2];     // [seg][block]: first K column inside the weight matrix
  int qmap[2];         // activation tensor map (3D: k, row, t)
  int qcol0[2];        // first K column inside an activation row
  int qtoff[2];        // time coordinate = t + qtoff
  int nblk;            // 1: plain 128-row tiles; 2: [64 rows of map A | 64 rows of map B]; 4: the four gates of 32 units
  int blk_stride;      // nblk == 4: rows between two gates (= D)
  int N;               // output features of the problem (plain: weight rows; 2-block / 4-gate: units)
  int tiles, split, cta0;   // (host planner) tiles, split-K ways, first CTA of the problem inside its phase
  int epi;
  int beta;            // 1: the finished value adds the old C
  const float* bias; const float* bias2;   // [N-space of the staged tile row], may be null
  TPtr add; long ldadd;                    // + add[q * ldadd + n] (null p: none)
  TPtr C; long ldc;
  // cell operands, meaning per `epi` (step_kernel.cu, finish_*):
  //  Lstm:    C = activated gates out [q][4D]; add = hoisted pre-activation; a0 = c_prev, a1 = c_out, a2 = h_out (ld0)
  //  CtxGate: add = gate pre-activation part from phase B (ldadd), bias2 = sc_affine bias; a0 = tc pre-activation (ld0),
  //           a1 = zst [q][3D], a2 = att_cap out (ld1)
  //  Copy1:   C = g2 (accumulated pre-activations in, activated gates out) [q][4D]; a0 = c2_prev, a1 = c_new out
  //  Copy2:   C = copy-gate pre-activation (read, ldc); a0 = g2 (o gate at +3D, ld0), a1 = sel, a2 = c_new, a3 = k out,
  //           a4 = c2 out, a5 = h2 out, a6 = dropout(h2) out
  TPtr a0, a1, a2, a3, a4, a5, a6;
  long ld0, ld1;
};

struct StepPhase {
  int nprob;
  int prob[kStepMaxPhaseProbs];     // indices into StepParams::prob, in CTA order (cta0 ascending)
  int ncta;                         // CTAs with a job in this phase
};

// both attentions (editnet.py:370-376, 409-421, 442-446; adaptive editnet_adaptive.py:449-456)
struct StepAttn {
  int P, R, A, D, F;
  TPtr s2; long ld_s2;              // row i: [cap_decoder_att(h1) (A) | decoder_att(h1) (A) | ...]
  const float* att1c;               // [B][P][A] hoisted cap_features_att(prev_h)
  TPtr att1v;                       // [B][R][A] hoisted features_att(fe) (per step in train mode)
  const float* cap_w; const float* cap_b; const float* vis_w; const float* vis_b;
  const float* mask;                // [B][P]
  const int* nreg;                  // [B] valid regions (adaptive) or null
  float* sc;                        // [B][P + R] raw scores (scratch between the two attention phases)
  int map_feats, map_prevh;         // 3D maps (cols, rows, sample), box {256, chunk rows, 1}, no swizzle
  int chunk_v, chunk_c;             // rows per staged chunk (<= 36)
  const float* prev_m;              // [B][P][D]
  TPtr alpha_c, ctx, sel, alpha_v, att_img;
  long ld_img;
  int* sel_idx; long sel_idx_st;
};

struct StepParams {
  CUtensorMap maps[kStepMaxMaps];
  StepProb prob[kStepMaxProbs];
  StepPhase phase[kStepGemmPhases];
  StepAttn attn;
  int t0, nt;                       // timesteps [t0, t0 + nt)
  int bt[64];                       // decoded rows at step t (index t - t0); nt <= 64
  int B;                            // rows of a time-major slab
  int D;
  int train; unsigned long long seed;
  unsigned int* sync;               // [0] grid barrier, [1] exit counter
  float* slabs;                     // (unused: split-K partials meet in distributed shared memory)
  int trace_cta;                    // CTA whose fine-grained stamps are recorded (-DSET_STEP_FINE_TRACE builds)
  unsigned long long* trace;        // optional: [(step * 8 + phase) * grid + cta] %globaltimer at each phase boundary
};

// host side (step_kernel.cu)
// device supports the cooperative cluster launch (one CTA per SM); *grid = CTAs to launch, *cluster = CTAs per cluster
bool step_kernel_available(int* grid, int* cluster);
int step_plan_splits(StepParams& prm, int grid, int cluster);   // fills tiles / split / cta0 / phase[].ncta
int step_launch(StepParams& prm, int grid, int cluster, cudaStream_t stream);
void step_set_trace(unsigned long long* buf);
extern long long g_step_launches, g_step_steps;   // persistent launches / timesteps they covered

// tensor-map encoders (gemm_tc.cu): fp32, rank 2 or 3, dims/strides innermost first (strides in bytes, rank - 1 of them)
bool tc_encode_map(CUtensorMap* m, const float* ptr, int rank, const unsigned long long* dims,
                   const unsigned long long* strides_bytes, const unsigned int* box, bool swizzle128);

}  // namespace set
