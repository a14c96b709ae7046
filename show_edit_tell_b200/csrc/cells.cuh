// Element-wise / attention kernels of the EditNet + DCNet decode step (forward and
// backward).  Each launcher states which reference lines it reproduces.
#pragma once
#include "common.cuh"

namespace set {

// dropout(relu(Emb[tok])) for a [n_outer][n_inner] grid of tokens -> out[(o*n_inner+i)*D + e]
// token(o,i) = tokens[i*tok_ld + o*tok_os]; keep-bit index = (drop_row0 + o*drop_os + i*drop_is)*D + e
// (editnet.py:300-304)
int embed_fwd(const int64_t* tokens, long tok_ld, long tok_os, const float* table, int V, float* out,
              int n_outer, int n_inner, int D, int train, uint64_t seed, uint32_t site, long drop_row0,
              long drop_os, long drop_is, cudaStream_t s);
// table_grad[tok] += dout * (out > 0 ? scale : 0)  (relu + dropout backward, scatter-add)
int embed_bwd(const int64_t* tokens, long tok_ld, long tok_os, const float* out, const float* dout,
              float* table_grad, int V, int n_outer, int n_inner, int D, int train, const int* row_len,
              cudaStream_t s);

// mean over regions (editnet.py:503)
int region_mean(const float* feats, float* out, int B, int R, int F, cudaStream_t s);
// adaptive: nreg[i] = #rows with non-zero sum (editnet_adaptive.py:440-441); fe rows >= nreg zeroed
int region_count(const float* feats, int* nreg, int B, int R, int F, cudaStream_t s);
int zero_pad_regions(float* fe, const int* nreg, int B, int R, int D, cudaStream_t s);

// fe_t[t][i][r][e] = keep ? 2*fe_pre[i][r][e] : 0   (att_embed dropout, editnet.py:432,441)
int vis_dropout_fwd(const float* fe_pre, float* fe_t, int T, int B, int R, int D, uint64_t seed, cudaStream_t s);
// dfe_pre[i][r][e] = (fe_pre>0) * sum_t keep_t*2*dfe_t[t][i][r][e]
int vis_dropout_bwd(const float* fe_pre, const float* dfe_t, float* dfe_pre, const int* dec_len, int T, int B,
                    int R, int D, uint64_t seed, cudaStream_t s);
// x = (y > 0) ? x : 0  (relu backward in place on x given the forward output y)
int relu_bwd_inplace(float* dx, const float* y, long n, cudaStream_t s);

// LSTM cell pointwise.  pre[rows][4D] (i,f,g,o pre-activations) -> gates (post-activation, saved),
// c_out, h_out.  With `len` != null this is the caption-encoder step `t` (editnet.py:333-338):
// rows with len[i] <= t keep (h_prev,c_prev), emit zeros into seq_h/seq_m and zero gates.
int lstm_fwd(const float* pre, long ld_pre, const float* c_prev, const float* h_prev, float* gates,
             float* c_out, float* h_out, long ld_h, int rows, int D, const int64_t* len, int t, float* seq_h,
             float* seq_m, long seq_ld, cudaStream_t s);
// dgates (pre-activation grads) from dh (+dh2 optional second addend), dc carry (in/out: becomes dc_prev)
int lstm_bwd(const float* gates, const float* c_prev, const float* c_cur, const float* dh, long ld_dh,
             const float* dh_b, float* dc_carry, float* dgates, int rows, int D, cudaStream_t s);
// encoder BPTT step t (reverse of the masked lstm_fwd): dh_run/dc_run carries, dprev_h/dprev_m inputs,
// dh_last injected at t == len-1
int enc_lstm_bwd(const float* gates, const float* c_prev, const float* c_cur, float* dh_run, float* dc_run,
                 const float* dseq_h, const float* dseq_m, long seq_ld, const float* dh_last,
                 const int64_t* len, int t, float* dgates, int rows, int D, cudaStream_t s);
// One step of one direction of the packed bidirectional nn.LSTM of DCNet's caption encoder
// (dcnet.py:217,233).  Row i works on position pos = reverse ? len[i]-1-s : s and is active iff
// s < len[i]; inactive rows keep their state.  pre = xg[i][pos] (+ hh_pre[i]); out[i][pos][col0..col0+C) = h.
int bilstm_fwd(const float* hh_pre, const float* xg, const int64_t* len, int s, int reverse, const float* h_prev,
               const float* c_prev, float* h_out, float* c_out, float* gates, float* out, long out_ld_row,
               long out_ld_pos, int B, int P, int C, cudaStream_t st);
// reverse of the above: dgates (also scattered to dxg[i][pos]), carries dh_run / dc_run
int bilstm_bwd(const float* gates, const float* c_prev, const float* c_cur, float* dh_run, float* dc_run,
               const float* dout, long out_ld_row, long out_ld_pos, const float* dh_last, long ld_dh_last,
               const int64_t* len, int s, int reverse, float* dgates, float* dxg, int B, int P, int C,
               cudaStream_t st);
// mask[i][p] = (sum_d prev_m[i][p][d] != 0)  (editnet.py:340)
int enc_mask(const float* prev_m, float* mask, int B, int P, int D, cudaStream_t s);
// y = dy * (1 - y^2) in place on dy (tanh backward)
int tanh_bwd_inplace(float* dy, const float* y, long n, cudaStream_t s);

struct AttnFwdArgs {
  int b, P, R, D, A, F;
  // caption attention (editnet.py:370-376) + select (editnet.py:409-421)
  const float* att1c;   // [B][P][A]
  const float* s2;      // row i: [att2c(A) | att2(A) | ...], ld_s2
  long ld_s2;
  const float* cap_w;   // [A]
  const float* cap_b;   // [1]
  const float* mask;    // [B][P]
  const float* prev_h;  // [B][P][D]
  const float* prev_m;  // [B][P][D] (null: no select -- DCNet)
  float* alpha_c;       // [b][P]
  float* ctx;           // row i at ctx + i*ld_ctx (ld_ctx == 0 means D)
  long ld_ctx;
  float* sel;           // [b][D]
  int* sel_idx;         // [b]
  // visual attention (editnet.py:442-446; adaptive :449-456)
  const float* att1v;   // [B][R][A]
  const float* vis_w;   // [A]
  const float* vis_b;   // [1]
  const float* feats;   // [B][R][F]
  const int* nreg;      // [B] or null
  float* alpha_v;       // [b][R]
  float* att_img;       // row i at att_img + i*ld_img
  long ld_img;
};
int attention_fwd(const AttnFwdArgs& a, cudaStream_t s);

struct AttnBwdArgs {
  int b, P, R, D, A, F;
  const float* att1c; const float* s2; long ld_s2; const float* cap_w; const float* mask;
  const float* prev_h; const float* prev_m; const float* alpha_c; const int* sel_idx;
  const float* dctx;      // row i at dctx + i*ld_dctx (0 means D)
  long ld_dctx;
  const float* dsel;      // [b][D] (null: no select)
  float* dprev_h;         // [B][P][D] +=
  float* dprev_m;         // [B][P][D] +=
  float* datt1c;          // [B][P][A] +=
  float* ds2;             // row i: [datt2c(A) | datt2(A) | ...] written
  long ld_ds2;
  float* dcap_w;          // [A] += (atomic)
  float* dcap_b;          // [1] += (atomic)
  const float* att1v; const float* vis_w; const float* feats; const int* nreg; const float* alpha_v;
  const float* datt_img; long ld_dimg;
  float* datt1v;          // [B][R][A]; written (=) when datt1v_accum==0, else +=
  int datt1v_accum;
  float* dvis_w; float* dvis_b;
};
int attention_bwd(const AttnBwdArgs& a, cudaStream_t s);

// context gate (editnet.py:378-380): z = sigmoid(zc), att_cap = z*tanh(sc) + (1-z)*tanh(th)
// s4 row: [zc(D) | sc(D) | ...] ; th from s2 row (offset th_off); saves zst row [z | tsc | ttc]
int ctx_gate_fwd(const float* s4, long ld_s4, const float* th, long ld_th, float* zst, float* att_cap,
                 long ld_cap, int rows, int D, cudaStream_t s);
// writes dzpre -> dz_out, dtcpre -> dtc_out (both in the dS2 row), dscpre -> dsc
int ctx_gate_bwd(const float* zst, const float* datt_cap, long ld_dcap, float* dz_out, float* dtc_out,
                 long ld_ds2, float* dsc, int rows, int D, cudaStream_t s);

// copy-LSTM (editnet.py:272-285), stage 1: gates in place, c_new = f*c2 + i*g
int copy1_fwd(float* g2, const float* c2_prev, float* cnew, int rows, int D, cudaStream_t s);
// stage 2: k = sigmoid(kpre); c2 = k*sel + (1-k)*cnew; h2 = o*tanh(c2); h2drop = dropout(h2)
int copy2_fwd(const float* kpre, long ld_k, const float* g2, const float* sel, const float* cnew, float* kgate,
              float* c2, float* h2, float* h2drop, int rows, int D, int train, uint64_t seed, long drop_base,
              cudaStream_t s);
// backward stage 2: from dh2 (= dh2_carry + keep*2*dh2drop_raw) and dc2_carry
int copy2_bwd(const float* dh2_carry, const float* dh2drop_raw, const float* dc2_carry, const float* g2,
              const float* c2, const float* kgate, const float* sel, const float* cnew, float* dg2,
              float* dkpre, float* dsel, float* dcnew, int rows, int D, int train, uint64_t seed, long drop_base,
              cudaStream_t s);
// backward stage 1: dcnew -> di,df,dg pre-activation grads; dc2_carry = dcnew*f
int copy1_bwd(const float* dcnew, const float* g2, const float* c2_prev, float* dg2, float* dc2_carry, int rows,
              int D, cudaStream_t s);

// plain LSTM (DCNet language_lstm / both attention LSTMs) share lstm_fwd/lstm_bwd.
// h2drop for a plain cell: out = dropout(h)
int dropout_fwd(const float* x, float* out, int rows, int D, int train, uint64_t seed, uint32_t site,
                long drop_base, cudaStream_t s);

// out[c][r] = in[r][c]   (in is [rows][cols], dense)
int transpose(const float* in, float* out, int rows, int cols, cudaStream_t s);
// strided form: out[c*ld_out + r] = in[r*ld_in + c]
int transpose_ld(const float* in, long ld_in, float* out, long ld_out, int rows, int cols, cudaStream_t s);
// several transposes in one launch: out[c * ld_out + r] = in[r * ld_in + c]
constexpr int kTrMaxJobs = 24;
struct TrJob { const float* in; long ld_in; float* out; long ld_out; int rows, cols; int tile0, tiles_c; };
struct TrBatch { int n; TrJob j[kTrMaxJobs]; };
int transpose_batch(const TrJob* jobs, int n, cudaStream_t s);
// sum over time of a [T][B][N] buffer -> [B][N]
int sum_time(const float* x, float* out, int T, long BN, cudaStream_t s);
// several ranges set to zero in one launch (the accumulated / scattered gradient tensors of an overwriting reverse pass)
constexpr int kZeroMaxJobs = 40;
struct ZeroJob { float* p; long n; };
struct ZeroBatch { ZeroJob j[kZeroMaxJobs]; };
int zero_batch(const ZeroJob* jobs, int n, cudaStream_t s);
// materialise keep bits as floats (tests): out[i] = keep(seed, site, base+i)
int dropout_keep_mask(float* out, long n, uint64_t seed, uint32_t site, long base, cudaStream_t s);

}  // namespace set
