// GEMM front-end for the decode path: every dense contraction on the path goes
// through set::gemm_group().  Three operand layouts cover forward (NT), the dX pass
// (NN) and the dW pass (TN); up to four K-segments let a logical concatenation
// ([h2;h1], [att_cap|att_img], ...) be consumed in place without materialising it.
#pragma once
#include "common.cuh"

namespace set {

enum GemmMode { kNT = 0, kNN = 1, kTN = 2 };
// kNT: C[m,n] = sum_k A[row(m) + k]   * B[n*ldb + k]      (x @ W^T, W row-major [N,K])
// kNN: C[m,n] = sum_k A[row(m) + k]   * B[k*ldb + n]      (dy @ W)
// kTN: C[m,n] = sum_k A[row(k) + m]   * B[k*ldb + n]      (dy^T @ x)
// row(x) = x*lda, or with a_inner>0: (x / a_inner)*lda + (x % a_inner)*a_ld_inner
// (time-major row index over a batch-major [B][T][.] tensor).

struct GemmSeg {
  const float* A; long lda;
  const float* B; long ldb;
  int K;
};

// Pointwise cell fused behind a skinny (M = batch) GEMM: the CTA that completes a tile's split-K reduction
// applies it to the finished pre-activations (gemm_tc.cu "fused epilogue").  A 4-gate op makes the kernel
// compose each 128-row weight tile from the 32-row blocks of the four gates of the same 32 hidden units.
enum GemmEpiOp { kEpiNone = 0, kEpiLstm = 1, kEpiCopy1 = 2, kEpiCopy2 = 3, kEpiLstmBwd = 4, kEpiCtxGateBwd = 5, kEpiCopy1Bwd = 6, kEpiCopy2Bwd = 7 };
struct GemmEpi {
  int op;
  int D;                        // hidden size = gate stride along N
  const float* c_prev;          // lstm, copy1: previous cell state [rows][D]
  float* c_out;                 // lstm: c;  copy1: c_new;  copy2: c2
  float* h_out; long ld_h;      // lstm: h;  copy2: h2
  float* gates; long ld_gates;  // lstm: activated gates out;  copy1: g2 in place (C);  copy2: g2 (o gate, read)
  const float* sel; const float* cnew;   // copy2
  float* kgate; float* h2drop;           // copy2
  int train; unsigned long long seed; long drop_base;
  // reverse-pass cells (elementwise on the GEMM's output columns):
  //   lstm_bwd:     C = d h (this step's consumers); x0 = carried d h (may be null), x1 = c_cur, c_prev, gates;
  //                 y0 = d c carry (in/out), y1 = d gates out [rows][4D]
  //   ctx_gate_bwd: columns [col0, col0 + D) of C are d att_cap; x0 = zst [rows][3D]; y0 = d z-pre, y1 = d tc-pre
  //                 (row stride ldy), y2 = d sc-pre [rows][D]
  //   copy1_bwd:    C = d c_new; gates = g2 (i,f,g read), c_prev = c2_prev; y0 = d c2 carry (written), y1 = d g2 [rows][4D]
  //   copy2_bwd:    C = carried d h2; x0 = d dropout(h2) raw (may be null), y0 = d c2 carry (read), gates = g2 (o read),
  //                 x1 = c2, kgate (read), sel, cnew; y1 = d g2 (o gate written), y2 = d k-pre, x2 = d sel out, x3 = d c_new out
  const float* x0; const float* x1;
  float* y0; float* y1; float* y2; long ldy; int col0;
  float* x2; float* x3;
  // length-masked LSTM (the previous-caption encoder, editnet.py:333-338): row i is active at step t iff len[i] > t;
  // inactive rows carry c/h through and store zero gates; seq_h/seq_m (optional) get h/c of active rows, else 0
  const long long* len; int t; const float* h_prev; float* seq_h; float* seq_m; long seq_ld;
};

struct GemmProblem {
  int M, N, nseg;
  GemmSeg seg[4];
  int a_inner; long a_ld_inner;          // two-level A rows (all segments)
  const int* a_row_len; int a_valid_inner;  // A row x valid iff a_row_len[x % vi] > x / vi, else reads 0
  const float* bias; const float* bias2; // [N], added once (ignored when null)
  const float* add; long ldadd; int add_mod;  // + add[(add_mod ? m % add_mod : m)*ldadd + n]
  float* C; long ldc;
  int c_inner; long c_ld_inner;          // two-level C rows
  const int* c_row_len; int c_valid_inner;  // invalid C rows are not written
  int beta;                              // 1: C += result
  int c_zeroed;                          // caller guarantees C is all-zero (lets split-K skip its memset)
  int act;                               // 0 none, 1 relu, 2 tanh
  GemmEpi epi;                           // optional fused cell (tensor-core swap mode only)
  int* epi_done;                         // set to 1 when the launch applies `epi`; else the caller runs the cell kernel
  int w_const;                           // the B operands are weights no in-flight kernel writes: the tensor-core
                                         // kernel may stream them before its grid dependency resolves (common.cuh)
};

struct GemmGroup {
  int n;
  int tile_start[9];
  GemmProblem p[8];
};

inline GemmProblem gemm_problem(int M, int N, float* C, long ldc) {
  GemmProblem p;
  memset(&p, 0, sizeof(p));
  p.M = M; p.N = N; p.C = C; p.ldc = ldc;
  return p;
}
inline void gemm_add_seg(GemmProblem& p, const float* A, long lda, const float* B, long ldb, int K) {
  if (K <= 0 || A == nullptr) return;
  GemmSeg& s = p.seg[p.nseg++];
  s.A = A; s.lda = lda; s.B = B; s.ldb = ldb; s.K = K;
}

// tensor-core path (gemm_tc.cu): launches the eligible problems of a group in one grid
int gemm_tc_try_group(int mode, const GemmProblem* probs, int n, bool* taken, cudaStream_t stream);
void gemm_tc_set_trace(unsigned long long* buf);
void gemm_tc_set_trace_seq(unsigned long long* buf, long stride, int launches);
extern int g_backend;
extern long long g_tc_launches, g_simt_launches, g_tc_twin_launches;

// Launch up to 8 independent problems of one mode (tensor cores where eligible, else one grouped
// CUDA-core grid).
int gemm_group(int mode, const GemmProblem* probs, int n, cudaStream_t stream);
inline int gemm(int mode, const GemmProblem& p, cudaStream_t stream) { return gemm_group(mode, &p, 1, stream); }

// column sums: out[n] (+)= sum_m X[m*ld + n]
int colsum(const float* X, long ld, int M, int N, float* out, int beta, cudaStream_t stream);

// several accumulating column sums in one launch: out[n] += sum_m X[m*ld + n]
constexpr int kColMaxJobs = 24;
struct ColJob { const float* X; long ld; int M, N; float* out; int block0, col_blocks; };
struct ColBatch { int n; ColJob j[kColMaxJobs]; };
int colsum_batch(const ColJob* jobs, int n, cudaStream_t stream);

}  // namespace set
