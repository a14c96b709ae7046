// Sub-module call surface of the reference: the beam-search code the reference ships does not go through
// DecoderC.forward -- evaluate() (editnet.py:613, 645-653) and evaluate_full() (eval/eval xe/eval_full.py:107-149) call
// decoder.caption_encoder / .embed / .attention_lstm / .caption_attention / .visual_attention / .select / .copy_lstm /
// .fc one by one.  Each entry point below is ONE of those forwards as the reference writes it (nothing hoisted: the
// caller owns the loop), built from the same GEMM engine and cell kernels as the fused path.  Inference only (no saved
// activations); scratch comes from the caller.
#include "../../include/set_b200.h"
#include "cells.cuh"
#include "gemm.cuh"

namespace set {
namespace {

// SelectC.forward (editnet.py:403-421): one-hot at argmax(alpha) with the straight-through weight alpha + (1 - alpha)
__global__ void select_fwd_kernel(const float* __restrict__ prev_m, const float* __restrict__ alpha, float* __restrict__ out,
                                  int P, int D) {
  const int i = blockIdx.x;
  __shared__ int js_s;
  __shared__ float w_s;
  if (threadIdx.x == 0) {
    int js = 0; float best = alpha[(long)i * P];
    for (int j = 1; j < P; ++j) { const float v = alpha[(long)i * P + j]; if (v > best) { best = v; js = j; } }
    js_s = js; w_s = best + (1.f - best);
  }
  __syncthreads();
  const float* row = prev_m + ((long)i * P + js_s) * D;
  for (int d = threadIdx.x; d < D; d += blockDim.x) out[(long)i * D + d] = w_s * row[d];
}

}  // namespace
}  // namespace set

using namespace set;

extern "C" {

int set_embed_forward(const int64_t* tokens, long n, const float* table, int V, int D, int train, uint64_t seed,
                      float* out, void* stream) {
  SET_REQUIRE(tokens && table && out && n > 0 && V > 0 && D > 0 && D % 4 == 0, "bad args");
  return embed_fwd(tokens, 1, 0, table, V, out, 1, (int)n, D, train, seed, kSiteEmb, 0, 0, 1,
                   reinterpret_cast<cudaStream_t>(stream));
}

int set_lstm_cell_forward(int rows, int I, int D, const float* x, const float* h, const float* c, const float* w_ih,
                          const float* w_hh, const float* b_ih, const float* b_hh, float* gates_scratch, float* h_out,
                          float* c_out, void* stream) {
  SET_REQUIRE(rows > 0 && I > 0 && D > 0 && D % 4 == 0 && x && h && c && w_ih && w_hh && gates_scratch && h_out && c_out, "bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  GemmProblem p = gemm_problem(rows, 4 * D, gates_scratch, 4 * D);
  gemm_add_seg(p, x, I, w_ih, I, I);
  gemm_add_seg(p, h, D, w_hh, D, D);
  p.bias = b_ih; p.bias2 = b_hh;
  SET_PROPAGATE(gemm(kNT, p, st));
  return lstm_fwd(gates_scratch, 4 * D, c, nullptr, gates_scratch, c_out, h_out, D, rows, D, nullptr, 0, nullptr, nullptr, 0, st);
}

size_t set_caption_attention_scratch_floats(const SetDims* d, int rows, int P) {
  if (!d || rows <= 0 || P <= 0) return 0;
  return (size_t)rows * P * d->A + (size_t)rows * 2 * d->A + (size_t)rows * d->D * 7 + 64;
}

int set_caption_attention_forward(const SetDims* d, int rows, int P, const SetEditNetParams* w, const float* prev_h,
                                  const float* h1, const float* emb, const float* mask, float* scratch,
                                  size_t scratch_floats, float* out, float* alpha, void* stream) {
  SET_REQUIRE(d && w && prev_h && h1 && emb && mask && scratch && out && alpha && rows > 0 && P > 0, "bad args");
  SET_REQUIRE(scratch_floats >= set_caption_attention_scratch_floats(d, rows, P), "scratch too small");
  const int D = d->D, A = d->A;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* att1c = scratch;
  float* s2 = att1c + (size_t)rows * P * A;      // [rows][2A] (caption half used)
  float* ctx = s2 + (size_t)rows * 2 * A;        // [rows][D]
  float* s4 = ctx + (size_t)rows * D;            // [rows][3D]: zc | sc | tc
  float* zst = s4 + (size_t)rows * 3 * D;        // [rows][3D]
  {
    GemmProblem p[2];
    p[0] = gemm_problem(rows * P, A, att1c, A);                       // cap_features_att(prev_h), editnet.py:370
    gemm_add_seg(p[0], prev_h, D, w->ca_feat_w, D, D); p[0].bias = w->ca_feat_b;
    p[1] = gemm_problem(rows, A, s2, 2 * A);                          // cap_decoder_att(h1), :371
    gemm_add_seg(p[1], h1, D, w->ca_dec_w, D, D); p[1].bias = w->ca_dec_b;
    SET_PROPAGATE(gemm(kNT, p[0], st));
    SET_PROPAGATE(gemm(kNT, p[1], st));
  }
  AttnFwdArgs a;
  memset(&a, 0, sizeof(a));
  a.b = rows; a.P = P; a.R = 1; a.D = D; a.A = A; a.F = d->F;
  a.att1c = att1c; a.s2 = s2; a.ld_s2 = 2 * A; a.cap_w = w->ca_full_w; a.cap_b = w->ca_full_b;
  a.mask = mask; a.prev_h = prev_h; a.prev_m = nullptr; a.alpha_c = alpha; a.ctx = ctx;
  SET_PROPAGATE(attention_fwd(a, st));                                 // :372-376
  {
    GemmProblem p[3];
    p[0] = gemm_problem(rows, D, s4, 3 * D);                          // context_gate([emb; h1; ctx]), :378
    gemm_add_seg(p[0], emb, D, w->ca_gate_w, 3 * D, D);
    gemm_add_seg(p[0], h1, D, w->ca_gate_w + D, 3 * D, D);
    gemm_add_seg(p[0], ctx, D, w->ca_gate_w + 2 * D, 3 * D, D);
    p[0].bias = w->ca_gate_b;
    p[1] = gemm_problem(rows, D, s4 + D, 3 * D);                      // sc_affine(ctx), :380
    gemm_add_seg(p[1], ctx, D, w->ca_sc_w, D, D); p[1].bias = w->ca_sc_b;
    p[2] = gemm_problem(rows, D, s4 + 2 * D, 3 * D);                  // tc_affine([emb; h1]), :379
    gemm_add_seg(p[2], emb, D, w->ca_tc_w, 2 * D, D);
    gemm_add_seg(p[2], h1, D, w->ca_tc_w + D, 2 * D, D);
    p[2].bias = w->ca_tc_b;
    SET_PROPAGATE(gemm_group(kNT, p, 3, st));
  }
  return ctx_gate_fwd(s4, 3 * D, s4 + 2 * D, 3 * D, zst, out, D, rows, D, st);
}

size_t set_visual_attention_scratch_floats(const SetDims* d, int rows, int R) {
  if (!d || rows <= 0 || R <= 0) return 0;
  return (size_t)rows * R * (d->D + d->A) + (size_t)rows * 2 * d->A + (size_t)rows + 64;
}

int set_visual_attention_forward(const SetDims* d, int rows, int R, const SetEditNetParams* w, const float* feats,
                                 const float* h1, int adaptive, int train, uint64_t seed, float* scratch,
                                 size_t scratch_floats, float* out, void* stream) {
  SET_REQUIRE(d && w && feats && h1 && scratch && out && rows > 0 && R > 0, "bad args");
  SET_REQUIRE(scratch_floats >= set_visual_attention_scratch_floats(d, rows, R), "scratch too small");
  const int D = d->D, A = d->A, F = d->F;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* fe = scratch;                                  // [rows][R][D]
  float* att1 = fe + (size_t)rows * R * D;              // [rows][R][A]
  float* s2 = att1 + (size_t)rows * R * A;              // [rows][2A] (visual half used)
  int* nreg = reinterpret_cast<int*>(s2 + (size_t)rows * 2 * A);
  {
    GemmProblem p = gemm_problem(rows * R, D, fe, D);   // relu(att_embed.0(feats)) -- every call, as the reference (:441)
    gemm_add_seg(p, feats, F, w->va_emb_w, F, F);
    p.bias = w->va_emb_b; p.act = 1;
    SET_PROPAGATE(gemm(kNT, p, st));
  }
  if (adaptive) {
    SET_PROPAGATE(region_count(feats, nreg, rows, R, F, st));
    SET_PROPAGATE(zero_pad_regions(fe, nreg, rows, R, D, st));
  }
  if (train) SET_PROPAGATE(vis_dropout_fwd(fe, fe, 1, rows, R, D, seed, st));   // dropout of att_embed (:432), in place
  {
    GemmProblem p[2];
    p[0] = gemm_problem(rows * R, A, att1, A);          // features_att(fe), :442
    gemm_add_seg(p[0], fe, D, w->va_feat_w, D, D); p[0].bias = w->va_feat_b;
    p[1] = gemm_problem(rows, A, s2 + A, 2 * A);        // decoder_att(h1), :443
    gemm_add_seg(p[1], h1, D, w->va_dec_w, D, D); p[1].bias = w->va_dec_b;
    SET_PROPAGATE(gemm(kNT, p[0], st));
    SET_PROPAGATE(gemm(kNT, p[1], st));
  }
  AttnFwdArgs a;
  memset(&a, 0, sizeof(a));
  a.b = rows; a.P = 1; a.R = R; a.D = D; a.A = A; a.F = F;
  a.s2 = s2; a.ld_s2 = 2 * A;
  a.att1v = att1; a.vis_w = w->va_full_w; a.vis_b = w->va_full_b; a.feats = feats;
  a.nreg = adaptive ? nreg : nullptr;
  a.alpha_v = reinterpret_cast<float*>(fe);             // (alpha is not part of this forward's result: park it in fe, consumed already)
  a.att_img = out; a.ld_img = F;
  return attention_fwd(a, st);                          // :444-446
}

size_t set_dcnet_caption_attention_scratch_floats(const SetDims* d, int rows, int P) {
  if (!d || rows <= 0 || P <= 0) return 0;
  return (size_t)rows * P * d->A + (size_t)rows * 2 * d->A + (size_t)rows * P + 64;
}

// DCNet CaptionAttention.forward, dcnet.py:254-270: the EditNet caption attention without gate and select
int set_dcnet_caption_attention_forward(const SetDims* d, int rows, int P, const SetDcNetParams* w, const float* enc,
                                        const float* h1, const float* mask, float* scratch, size_t scratch_floats,
                                        float* out, void* stream) {
  SET_REQUIRE(d && w && enc && h1 && mask && scratch && out && rows > 0 && P > 0, "bad args");
  SET_REQUIRE(scratch_floats >= set_dcnet_caption_attention_scratch_floats(d, rows, P), "scratch too small");
  const int D = d->D, A = d->A;     // encoder outputs are 2 * caption_features_dim = D wide (dcnet.py:286-291)
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* att1c = scratch;
  float* s2 = att1c + (size_t)rows * P * A;
  float* alpha = s2 + (size_t)rows * 2 * A;
  GemmProblem p[2];
  p[0] = gemm_problem(rows * P, A, att1c, A);
  gemm_add_seg(p[0], enc, D, w->ca_feat_w, D, D); p[0].bias = w->ca_feat_b;
  p[1] = gemm_problem(rows, A, s2, 2 * A);
  gemm_add_seg(p[1], h1, D, w->ca_dec_w, D, D); p[1].bias = w->ca_dec_b;
  SET_PROPAGATE(gemm(kNT, p[0], st));
  SET_PROPAGATE(gemm(kNT, p[1], st));
  AttnFwdArgs a;
  memset(&a, 0, sizeof(a));
  a.b = rows; a.P = P; a.R = 1; a.D = D; a.A = A; a.F = d->F;
  a.att1c = att1c; a.s2 = s2; a.ld_s2 = 2 * A; a.cap_w = w->ca_full_w; a.cap_b = w->ca_full_b;
  a.mask = mask; a.prev_h = enc; a.prev_m = nullptr; a.alpha_c = alpha; a.ctx = out;
  return attention_fwd(a, st);
}

int set_select_forward(int rows, int P, int D, const float* prev_m, const float* alpha, float* out, void* stream) {
  SET_REQUIRE(rows > 0 && P > 0 && D > 0 && prev_m && alpha && out, "bad args");
  select_fwd_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(prev_m, alpha, out, P, D);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  return SET_OK;
}

size_t set_copy_lstm_scratch_floats(const SetDims* d, int rows) {
  if (!d || rows <= 0) return 0;
  return (size_t)rows * d->D * 7 + 64;
}

int set_copy_lstm_forward(const SetDims* d, int rows, const SetEditNetParams* w, const float* x, const float* h,
                          const float* c, const float* mem, float* scratch, size_t scratch_floats, float* h_out,
                          float* c_out, void* stream) {
  SET_REQUIRE(d && w && x && h && c && mem && scratch && h_out && c_out && rows > 0, "bad args");
  SET_REQUIRE(scratch_floats >= set_copy_lstm_scratch_floats(d, rows), "scratch too small");
  const int D = d->D, LX2 = 2 * d->D + d->F;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  float* g2 = scratch;                      // [rows][4D]
  float* cnew = g2 + (size_t)rows * 4 * D;  // [rows][D]
  float* kpre = cnew + (size_t)rows * D;    // [rows][D]
  float* kgate = kpre + (size_t)rows * D;   // [rows][D]
  {
    GemmProblem p = gemm_problem(rows, 4 * D, g2, 4 * D);             // x2h(x) + h2h(h), editnet.py:272
    gemm_add_seg(p, x, LX2, w->cl_x2h_w, LX2, LX2);
    gemm_add_seg(p, h, D, w->cl_h2h_w, D, D);
    p.bias = w->cl_x2h_b; p.bias2 = w->cl_h2h_b;
    SET_PROPAGATE(gemm(kNT, p, st));
  }
  SET_PROPAGATE(copy1_fwd(g2, c, cnew, rows, D, st));                 // :273-279
  {
    GemmProblem p = gemm_problem(rows, D, kpre, D);                   // gate_cnew(c_new) + gate_cmem(mem), :281
    gemm_add_seg(p, cnew, D, w->cl_gcn_w, D, D);
    gemm_add_seg(p, mem, D, w->cl_gcm_w, D, D);
    p.bias = w->cl_gcn_b; p.bias2 = w->cl_gcm_b;
    SET_PROPAGATE(gemm(kNT, p, st));
  }
  return copy2_fwd(kpre, D, g2, mem, cnew, kgate, c_out, h_out, nullptr, rows, D, 0, 0, 0, st);   // :282-285
}

}  // extern "C"
