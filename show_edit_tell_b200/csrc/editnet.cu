// Host orchestration + C ABI of the EditNet path: caption encoder, hoisted projections,
// the per-step kernel chain, batched vocabulary projection, and the full reverse pass.
// Reference: DecoderC.forward editnet.py:479-548, editnet_rl.py:485-549,
// adaptive_features/editnet_adaptive.py:489-562 and the cells they call (editnet.py:210-447).
//
// Data layout in HBM (fp32).  Per-sequence tensors are batch-major like the reference's
// ([B][P][D] encoder outputs, [B][R][F] features); everything the recurrence saves is
// time-major ([T][B][.]) so a step's rows are contiguous and the time-batched contractions
// (vocabulary projection, every dW, the hoisted word/feature projections) see one dense
// [T*B, K] operand.  The concatenations of the reference are never materialised: X2[t] =
// [h1 | att_cap | att_img] is written in place by the producing kernels, and GEMM
// K-segments read [h2 ; h1] or column blocks of weight_ih where they lie.
#include "seq_common.cuh"
#include "step_kernel.cuh"

namespace set {

namespace {

struct Ws {
  // ---- zeroed at the start of forward (region A)
  float *emb_prev, *enc_xg, *enc_gates, *enc_h, *enc_c, *prev_h, *prev_m, *mask, *fh, *att1c;
  float *image_mean, *fe_pre, *att1, *fe_t, *pre1s, *emb_all, *pre1, *gw, *tw;
  float *gates1, *c1, *X2, *s2, *g2, *alpha_c, *ctx_c, *sel, *alpha_v, *zst, *cnew, *kgate, *c2, *h2, *h2drop;
  float *logits, *lse, *scratch4d, *s4, *ones, *att_sc;
  int *sel_idx, *nreg, *dec_len, *unfinished, *unf_count;
  int64_t *it, *tok_raw;
  // ---- zeroed at the start of backward (region B)
  float *dG1, *dS2, *dG2, *dsc, *dK, *dh2raw, *datt1, *datt1c, *dprev_h, *dprev_m, *dfh, *dfe_pre, *dfe_t;
  float *demb_all, *demb_prev, *denc_g, *dh_last, *dh_run, *dc_run, *sumG1;
  float *dh2c, *dc2c, *dh1c, *dc1c, *dX2, *dctx, *dsel, *dcnew;
  // ---- transposed weight copies for the dX pass (written at the start of backward, never zeroed)
  float *t_fc, *t_cl_gcn, *t_cl_gcm, *t_cl_x2h, *t_ca_gate, *t_ca_sc, *t_ca_dec, *t_va_dec, *t_ca_tc, *t_al_whh,
      *t_al_wih, *t_cl_h2h, *t_ca_feat, *t_va_feat, *t_enc_aff, *t_enc_h2h, *t_enc_x2h;
  float* tscratch; size_t tscratch_floats;   // transposed activations for the dW pass
  size_t regionA_end = 0, regionB_begin = 0, regionB_end = 0, total = 0;
};

struct Ctx {
  SetDims d;
  SetSeqShape s;
  const SetEditNetParams* w;
  Ws ws;
  cudaStream_t st;
  uint64_t seed;
  int LS2, LX2;
  int fresh = 1;   // 1: GEMM output slabs were zeroed by the region memset of this call (split-K skips memsets)
};

void layout(const SetDims& d, const SetSeqShape& s, Arena& ar, Ws& w) {
  const size_t B = s.B, P = s.P, T = s.T, R = s.R, D = d.D, A = d.A, F = d.F, V = d.V;
  const size_t Tv = s.train ? T : 1;  // per-step visual tensors exist only with dropout
  w.emb_prev = ar.take<float>("emb_prev", P * B * D);
  w.enc_xg = ar.take<float>("enc_xg", P * B * 4 * D);
  w.enc_gates = ar.take<float>("enc_gates", P * B * 4 * D);
  w.enc_h = ar.take<float>("enc_h", (P + 1) * B * D);
  w.enc_c = ar.take<float>("enc_c", (P + 1) * B * D);
  w.prev_h = ar.take<float>("prev_h", B * P * D);
  w.prev_m = ar.take<float>("prev_m", B * P * D);
  w.mask = ar.take<float>("mask", B * P);
  w.fh = ar.take<float>("final_hidden", B * D);
  w.att1c = ar.take<float>("att1c", B * P * A);
  w.image_mean = ar.take<float>("image_mean", B * F);
  w.fe_pre = ar.take<float>("fe_pre", B * R * D);
  w.att1 = ar.take<float>("att1", Tv * B * R * A);
  w.fe_t = ar.take<float>("fe_t", s.train ? T * B * R * D : 1);
  w.pre1s = ar.take<float>("pre1s", B * 4 * D);
  w.emb_all = ar.take<float>("emb_all", T * B * D);
  w.pre1 = ar.take<float>("pre1", T * B * 4 * D);
  w.gw = ar.take<float>("gw", T * B * D);
  w.tw = ar.take<float>("tw", T * B * D);
  w.gates1 = ar.take<float>("gates1", T * B * 4 * D);
  w.c1 = ar.take<float>("c1", (T + 1) * B * D);
  w.X2 = ar.take<float>("X2", T * B * (2 * D + F));
  w.s2 = ar.take<float>("s2", T * B * (2 * A + 2 * D));
  w.g2 = ar.take<float>("g2", T * B * 4 * D);
  w.alpha_c = ar.take<float>("alpha_c", T * B * P);
  w.ctx_c = ar.take<float>("ctx_c", T * B * D);
  w.sel = ar.take<float>("sel", T * B * D);
  w.alpha_v = ar.take<float>("alpha_v", T * B * R);
  w.zst = ar.take<float>("zst", T * B * 3 * D);
  w.cnew = ar.take<float>("cnew", T * B * D);
  w.kgate = ar.take<float>("kgate", T * B * D);
  w.c2 = ar.take<float>("c2", (T + 1) * B * D);
  w.h2 = ar.take<float>("h2", (T + 1) * B * D);
  w.h2drop = ar.take<float>("h2drop", T * B * D);
  w.logits = ar.take<float>("logits", Tv * B * V);   // rollouts only use it; XE writes to `predictions`
  w.lse = ar.take<float>("lse", T * B);
  w.scratch4d = ar.take<float>("scratch4d", B * 4 * D);
  w.s4 = ar.take<float>("s4", T * B * 3 * D);
  w.ones = ar.take<float>("ones", T * B);
  w.att_sc = ar.take<float>("att_sc", B * (P + R));
  w.sel_idx = ar.take<int>("sel_idx", T * B);
  w.nreg = ar.take<int>("nreg", B);
  w.dec_len = ar.take<int>("dec_len", B);
  w.unfinished = ar.take<int>("unfinished", B);
  w.unf_count = ar.take<int>("unf_count", T + 2);
  w.it = ar.take<int64_t>("it", (T + 1) * B);
  w.tok_raw = ar.take<int64_t>("tok_raw", T * B);
  w.regionA_end = ar.off;
  w.regionB_begin = ar.off;
  w.dG1 = ar.take<float>("dG1", T * B * 4 * D);
  w.dS2 = ar.take<float>("dS2", T * B * (2 * A + 2 * D));
  w.dG2 = ar.take<float>("dG2", T * B * 4 * D);
  w.dsc = ar.take<float>("dsc", T * B * D);
  w.dK = ar.take<float>("dK", T * B * D);
  w.dh2raw = ar.take<float>("dh2raw", T * B * D);
  w.datt1 = ar.take<float>("datt1", Tv * B * R * A);
  w.datt1c = ar.take<float>("datt1c", B * P * A);
  w.dprev_h = ar.take<float>("dprev_h", B * P * D);
  w.dprev_m = ar.take<float>("dprev_m", B * P * D);
  w.dfh = ar.take<float>("dfh", B * D);
  w.dfe_pre = ar.take<float>("dfe_pre", B * R * D);
  w.dfe_t = ar.take<float>("dfe_t", s.train ? T * B * R * D : 1);
  w.demb_all = ar.take<float>("demb_all", T * B * D);
  w.demb_prev = ar.take<float>("demb_prev", P * B * D);
  w.denc_g = ar.take<float>("denc_g", P * B * 4 * D);
  w.dh_last = ar.take<float>("dh_last", B * D);
  w.dh_run = ar.take<float>("dh_run", P * B * D);
  w.dc_run = ar.take<float>("dc_run", B * D);
  w.sumG1 = ar.take<float>("sumG1", B * 4 * D);
  w.dh2c = ar.take<float>("dh2c", (T + 1) * B * D);
  w.dc2c = ar.take<float>("dc2c", B * D);
  w.dh1c = ar.take<float>("dh1c", (T + 1) * B * D);
  w.dc1c = ar.take<float>("dc1c", B * D);
  w.dX2 = ar.take<float>("dX2", T * B * (2 * D + F));
  w.dctx = ar.take<float>("dctx", T * B * D);
  w.dsel = ar.take<float>("dsel", B * D);
  w.dcnew = ar.take<float>("dcnew", B * D);
  w.regionB_end = ar.off;
  w.t_fc = ar.take<float>("t_fc", V * D);
  w.t_cl_gcn = ar.take<float>("t_cl_gcn", D * D);
  w.t_cl_gcm = ar.take<float>("t_cl_gcm", D * D);
  w.t_cl_x2h = ar.take<float>("t_cl_x2h", 4 * D * (2 * D + F));
  w.t_ca_gate = ar.take<float>("t_ca_gate", D * 3 * D);
  w.t_ca_sc = ar.take<float>("t_ca_sc", D * D);
  w.t_ca_dec = ar.take<float>("t_ca_dec", A * D);
  w.t_va_dec = ar.take<float>("t_va_dec", A * D);
  w.t_ca_tc = ar.take<float>("t_ca_tc", D * 2 * D);
  w.t_al_whh = ar.take<float>("t_al_whh", 4 * D * D);
  w.t_al_wih = ar.take<float>("t_al_wih", 4 * D * (3 * D + F));
  w.t_cl_h2h = ar.take<float>("t_cl_h2h", 4 * D * D);
  w.t_ca_feat = ar.take<float>("t_ca_feat", A * D);
  w.t_va_feat = ar.take<float>("t_va_feat", A * D);
  w.t_enc_aff = ar.take<float>("t_enc_aff", D * D);
  w.t_enc_h2h = ar.take<float>("t_enc_h2h", 4 * D * D);
  w.t_enc_x2h = ar.take<float>("t_enc_x2h", 4 * D * D);
  {
    // upper bound of everything backward_core transposes for the dW pass (rows padded to 4)
    const size_t TB = T * B + 4, BRr = B * R + 4, TBR = (s.train ? T : 1) * B * R + 4, BP = B * P + 4, PB = P * B + 4;
    w.tscratch_floats = TB * (4 * D + D + D + (2 * D + F) + 4 * D + (2 * A + 2 * D) + 5 * D + V + D) +
                        (B + 4) * (4 * D + D + F + 2 * D) + BP * (A + D) + TBR * (A + D) + BRr * (D + F) +
                        PB * (4 * D + 2 * D) + 4096;
    w.tscratch_floats = w.tscratch_floats * 2 + TB * (4 * D + (2 * D + F));   // column-block / row-range views repeat some matrices
    w.tscratch = ar.take<float>("tscratch", w.tscratch_floats);
  }
  w.total = ar.off;
}

int check_args(const SetDims* d, const SetSeqShape* s) {
  SET_REQUIRE(d && s, "null dims/shape");
  SET_REQUIRE(d->D > 0 && d->D % 4 == 0 && d->A > 0 && d->A % 4 == 0 && d->F > 0 && d->F % 4 == 0 && d->V > 1,
              "D, A, F must be positive multiples of 4");
  SET_REQUIRE(s->B > 0 && s->R > 0 && s->R <= 1024 && s->Wp > 0 && s->P > 0 && s->P <= s->Wp && s->T > 0,
              "bad sequence shape");
  SET_REQUIRE(d->A <= 4096, "attention_dim too large for the attention kernel's shared memory");
  return SET_OK;
}

int make_ctx(Ctx& c, const SetDims* d, const SetSeqShape* s, const SetEditNetParams* w, void* workspace,
             size_t workspace_bytes, uint64_t seed, void* stream) {
  SET_PROPAGATE(check_args(d, s));
  SET_REQUIRE(w != nullptr && workspace != nullptr, "null params/workspace");
  SET_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  c.d = *d; c.s = *s; c.w = w; c.seed = seed;
  c.st = reinterpret_cast<cudaStream_t>(stream);
  Arena ar;
  ar.base = reinterpret_cast<char*>(workspace);
  layout(*d, *s, ar, c.ws);
  if (c.ws.total > workspace_bytes) {
    set_record_error("workspace too small: see set_editnet_workspace_bytes()");
    return SET_ERR_WORKSPACE;
  }
  c.LS2 = 2 * d->A + 2 * d->D;
  c.LX2 = 2 * d->D + d->F;
  return SET_OK;
}

// optional timing of the two recurrent loops (bench.py's roofline line): events on the launch stream
bool g_profile = false;
cudaEvent_t g_ev[4] = {nullptr, nullptr, nullptr, nullptr};
int g_ev_state = 0;  // bit0: forward pair recorded, bit1: backward pair recorded
int profile_mark(int idx, cudaStream_t st) {
  if (!g_profile) return SET_OK;
  if (!g_ev[idx]) SET_CHECK_CUDA(cudaEventCreate(&g_ev[idx]));
  SET_CHECK_CUDA(cudaEventRecord(g_ev[idx], st));
  if (idx == 1) g_ev_state |= 1;
  if (idx == 3) g_ev_state |= 2;
  return SET_OK;
}

// --------------------------------------------------------------------- shared prologue
// caption encoder (editnet.py:319-348), image mean (:503), hoisted time-invariant projections
int encode_prev(Ctx& c, const int64_t* prev, const int64_t* prev_len) {
  const int B = c.s.B, P = c.s.P, T = c.s.T, D = c.d.D, A = c.d.A;
  const SetEditNetParams& w = *c.w;
  Ws& s = c.ws;
  cudaStream_t st = c.st;
  SET_CHECK_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(s.emb_prev), 0, s.regionA_end, st));
  fill_kernel<<<8, 256, 0, st>>>(s.ones, (long)T * B, 1.0f);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  // --- previous-caption encoder
  SET_PROPAGATE(embed_fwd(prev, c.s.Wp, 1, w.embed, c.d.V, s.emb_prev, P, B, D, c.s.train, c.seed, kSiteEnc, 0, 1,
                          c.s.Wp, st));
  {
    GemmProblem p = gemm_problem(P * B, 4 * D, s.enc_xg, 4 * D);
    gemm_add_seg(p, s.emb_prev, D, w.enc_x2h_w, D, D);
    p.bias = w.enc_x2h_b; p.bias2 = w.enc_h2h_b;
    SET_PROPAGATE(gemm(kNT, p, st));
  }
  for (int t = 0; t < P; ++t) {
    // Every GEMM output below lands in a slab of the workspace that the region memset already zeroed
    // (c_zeroed): split-K launches then need no memset of their own.
    float* pre = s.enc_gates + (size_t)t * B * 4 * D;   // pre-activations, converted in place by lstm_fwd
    GemmProblem p = gemm_problem(B, 4 * D, pre, 4 * D);
    if (t > 0) gemm_add_seg(p, s.enc_h + (size_t)t * B * D, D, w.enc_h2h_w, D, D);
    p.add = s.enc_xg + (size_t)t * B * 4 * D; p.ldadd = 4 * D; p.c_zeroed = c.fresh; p.w_const = 1;
    int fused = 0;   // the length-masked encoder cell in the GEMM's epilogue
    p.epi.op = kEpiLstm; p.epi.D = D; p.epi.c_prev = s.enc_c + (size_t)t * B * D; p.epi.c_out = s.enc_c + (size_t)(t + 1) * B * D;
    p.epi.h_out = s.enc_h + (size_t)(t + 1) * B * D; p.epi.ld_h = D; p.epi.gates = pre; p.epi.ld_gates = 4 * D;
    p.epi.len = reinterpret_cast<const long long*>(prev_len); p.epi.t = t; p.epi.h_prev = s.enc_h + (size_t)t * B * D;
    p.epi.seq_h = s.prev_h; p.epi.seq_m = s.prev_m; p.epi.seq_ld = (long)P * D;
    p.epi_done = &fused;
    SET_PROPAGATE(gemm(kNT, p, st));
    if (!fused)
      SET_PROPAGATE(lstm_fwd(pre, 4 * D, s.enc_c + (size_t)t * B * D, s.enc_h + (size_t)t * B * D,
                             s.enc_gates + (size_t)t * B * 4 * D, s.enc_c + (size_t)(t + 1) * B * D,
                             s.enc_h + (size_t)(t + 1) * B * D, D, B, D, prev_len, t, s.prev_h, s.prev_m,
                             (long)P * D, st));
  }
  SET_PROPAGATE(enc_mask(s.prev_m, s.mask, B, P, D, st));
  {
    GemmProblem p[2];
    p[0] = gemm_problem(B, D, s.fh, D);   // final_hidden = tanh(affine_hn(h_last)), editnet.py:341
    gemm_add_seg(p[0], s.enc_h + (size_t)P * B * D, D, w.enc_aff_w, D, D);
    p[0].bias = w.enc_aff_b; p[0].act = 2;
    SET_PROPAGATE(gemm(kNT, p[0], st));
    p[1] = gemm_problem(B * P, A, s.att1c, A);  // cap_features_att(prev_h), editnet.py:370 (time-invariant)
    gemm_add_seg(p[1], s.prev_h, D, w.ca_feat_w, D, D);
    p[1].bias = w.ca_feat_b;
    SET_PROPAGATE(gemm(kNT, p[1], st));
  }
  return SET_OK;
}

int prepare_common(Ctx& c, const float* feats, const float* image_mean_in, const int64_t* prev,
                   const int64_t* prev_len) {
  const int B = c.s.B, T = c.s.T, R = c.s.R, D = c.d.D, A = c.d.A, F = c.d.F;
  const SetEditNetParams& w = *c.w;
  Ws& s = c.ws;
  cudaStream_t st = c.st;
  SET_PROPAGATE(encode_prev(c, prev, prev_len));
  // --- image side
  if (c.s.adaptive) {
    SET_REQUIRE(image_mean_in != nullptr, "adaptive needs image_mean");
    SET_CHECK_CUDA(cudaMemcpyAsync(s.image_mean, image_mean_in, sizeof(float) * B * F, cudaMemcpyDeviceToDevice, st));
    SET_PROPAGATE(region_count(feats, s.nreg, B, R, F, st));
  } else {
    SET_PROPAGATE(region_mean(feats, s.image_mean, B, R, F, st));
  }
  {
    GemmProblem p = gemm_problem(B * R, D, s.fe_pre, D);  // relu(att_embed.0(feats)), editnet.py:430-431,441
    gemm_add_seg(p, feats, F, w.va_emb_w, F, F);
    p.bias = w.va_emb_b; p.act = 1;
    SET_PROPAGATE(gemm(kNT, p, st));
  }
  if (c.s.adaptive) SET_PROPAGATE(zero_pad_regions(s.fe_pre, s.nreg, B, R, D, st));
  if (c.s.train) {
    // dropout is redrawn every step (editnet.py:432), so features_att is re-projected per step --
    // but it does not depend on the recurrence: all T steps go through one time-batched GEMM.
    SET_PROPAGATE(vis_dropout_fwd(s.fe_pre, s.fe_t, T, B, R, D, c.seed, st));
    GemmProblem p = gemm_problem(T * B * R, A, s.att1, A);
    gemm_add_seg(p, s.fe_t, D, w.va_feat_w, D, D);
    p.bias = w.va_feat_b;
    SET_PROPAGATE(gemm(kNT, p, st));
  } else {
    GemmProblem p = gemm_problem(B * R, A, s.att1, A);
    gemm_add_seg(p, s.fe_pre, D, w.va_feat_w, D, D);
    p.bias = w.va_feat_b;
    SET_PROPAGATE(gemm(kNT, p, st));
  }
  {
    // attention-LSTM input columns that do not change over time: final_hidden and image_mean
    GemmProblem p = gemm_problem(B, 4 * D, s.pre1s, 4 * D);
    gemm_add_seg(p, s.fh, D, w.al_wih + D, 3 * D + F, D);
    gemm_add_seg(p, s.image_mean, F, w.al_wih + 3 * D, 3 * D + F, F);
    p.bias = w.al_bih; p.bias2 = w.al_bhh;
    SET_PROPAGATE(gemm(kNT, p, st));
  }
  return SET_OK;
}

// word-dependent projections for rows [t0, t0+nt) x B of emb_all (hoisted for teacher forcing,
// per step for rollouts)
int project_words(Ctx& c, int t0, int nt) {
  const int B = c.s.B, D = c.d.D, F = c.d.F;
  const SetEditNetParams& w = *c.w;
  Ws& s = c.ws;
  const size_t r0 = (size_t)t0 * B;
  GemmProblem p[3];
  memset(p, 0, sizeof(p));
  p[0] = gemm_problem(nt * B, 4 * D, s.pre1 + r0 * 4 * D, 4 * D);
  gemm_add_seg(p[0], s.emb_all + r0 * D, D, w.al_wih, 3 * D + F, D);
  p[0].add = s.pre1s; p[0].ldadd = 4 * D; p[0].add_mod = B;
  p[1] = gemm_problem(nt * B, D, s.gw + r0 * D, D);
  gemm_add_seg(p[1], s.emb_all + r0 * D, D, w.ca_gate_w, 3 * D, D);
  p[1].bias = w.ca_gate_b;
  p[2] = gemm_problem(nt * B, D, s.tw + r0 * D, D);
  gemm_add_seg(p[2], s.emb_all + r0 * D, D, w.ca_tc_w, 2 * D, D);
  p[2].bias = w.ca_tc_b;
  p[0].c_zeroed = p[1].c_zeroed = p[2].c_zeroed = c.fresh;
  p[0].w_const = p[1].w_const = p[2].w_const = 1;
  return gemm_group(kNT, p, 3, c.st);
}

// one decode step on rows [0,b): SURVEY.md Appendix A steps 2-7 (editnet.py:527-543)
int step_forward(Ctx& c, const float* feats, int t, int b) {
  const int B = c.s.B, P = c.s.P, R = c.s.R, D = c.d.D, A = c.d.A, F = c.d.F;
  const int LS2 = c.LS2, LX2 = c.LX2;
  const SetEditNetParams& w = *c.w;
  Ws& s = c.ws;
  cudaStream_t st = c.st;
  float* X2t = s.X2 + (size_t)t * B * LX2;
  float* s2t = s.s2 + (size_t)t * B * LS2;
  float* g2t = s.g2 + (size_t)t * B * 4 * D;
  const float* h2prev = s.h2 + (size_t)t * B * D;
  float* s4t = s.s4 + (size_t)t * B * 3 * D;   // per-step [zc | sc | kc] (pre-zeroed slab)
  {  // attention-LSTM recurrent part: W_ih[:,2D:3D] h2 + W_hh h1 + hoisted terms (editnet.py:532)
    float* pre = s.gates1 + (size_t)t * B * 4 * D;   // pre-activations, converted in place by lstm_fwd
    GemmProblem p = gemm_problem(b, 4 * D, pre, 4 * D);
    if (t > 0) {
      gemm_add_seg(p, h2prev, D, w.al_wih + 2 * D, 3 * D + F, D);
      gemm_add_seg(p, s.X2 + (size_t)(t - 1) * B * LX2, LX2, w.al_whh, D, D);
    }
    p.add = s.pre1 + (size_t)t * B * 4 * D; p.ldadd = 4 * D; p.c_zeroed = c.fresh; p.w_const = 1;
    // the LSTM cell rides in the GEMM's epilogue when the tensor-core path takes the problem
    int fused = 0;
    p.epi.op = kEpiLstm; p.epi.D = D; p.epi.c_prev = s.c1 + (size_t)t * B * D; p.epi.c_out = s.c1 + (size_t)(t + 1) * B * D;
    p.epi.h_out = X2t; p.epi.ld_h = LX2; p.epi.gates = pre; p.epi.ld_gates = 4 * D; p.epi_done = &fused;
    SET_PROPAGATE(gemm(kNT, p, st));
    if (!fused)
      SET_PROPAGATE(lstm_fwd(pre, 4 * D, s.c1 + (size_t)t * B * D, nullptr, s.gates1 + (size_t)t * B * 4 * D,
                             s.c1 + (size_t)(t + 1) * B * D, X2t, LX2, b, D, nullptr, 0, nullptr, nullptr, 0, st));
  }
  {  // everything that consumes h1 in one grouped launch
    GemmProblem p[5];
    p[0] = gemm_problem(b, A, s2t, LS2);                     // cap_decoder_att(h1), :371
    gemm_add_seg(p[0], X2t, LX2, w.ca_dec_w, D, D); p[0].bias = w.ca_dec_b;
    p[1] = gemm_problem(b, A, s2t + A, LS2);                 // decoder_att(h1), :443
    gemm_add_seg(p[1], X2t, LX2, w.va_dec_w, D, D); p[1].bias = w.va_dec_b;
    p[2] = gemm_problem(b, D, s2t + 2 * A, LS2);             // context_gate[:, D:2D] h1 + word part, :378
    gemm_add_seg(p[2], X2t, LX2, w.ca_gate_w + D, 3 * D, D);
    p[2].add = s.gw + (size_t)t * B * D; p[2].ldadd = D;
    p[3] = gemm_problem(b, D, s2t + 2 * A + D, LS2);         // tc_affine[:, D:2D] h1 + word part, :379-380
    gemm_add_seg(p[3], X2t, LX2, w.ca_tc_w + D, 2 * D, D);
    p[3].add = s.tw + (size_t)t * B * D; p[3].ldadd = D;
    p[4] = gemm_problem(b, 4 * D, g2t, 4 * D);               // copy_lstm: x2h[:, 0:D] h1 + h2h h2, :272
    gemm_add_seg(p[4], X2t, LX2, w.cl_x2h_w, LX2, D);
    if (t > 0) gemm_add_seg(p[4], h2prev, D, w.cl_h2h_w, D, D);
    p[4].bias = w.cl_x2h_b; p[4].bias2 = w.cl_h2h_b;
    for (int k = 0; k < 5; ++k) { p[k].c_zeroed = c.fresh; p[k].w_const = 1; }
    SET_PROPAGATE(gemm_group(kNT, p, 5, st));
  }
  {
    AttnFwdArgs a;
    memset(&a, 0, sizeof(a));
    a.b = b; a.P = P; a.R = R; a.D = D; a.A = A; a.F = F;
    a.att1c = s.att1c; a.s2 = s2t; a.ld_s2 = LS2; a.cap_w = w.ca_full_w; a.cap_b = w.ca_full_b;
    a.mask = s.mask; a.prev_h = s.prev_h; a.prev_m = s.prev_m;
    a.alpha_c = s.alpha_c + (size_t)t * B * P; a.ctx = s.ctx_c + (size_t)t * B * D;
    a.sel = s.sel + (size_t)t * B * D; a.sel_idx = s.sel_idx + (size_t)t * B;
    a.att1v = c.s.train ? s.att1 + (size_t)t * B * R * A : s.att1;
    a.vis_w = w.va_full_w; a.vis_b = w.va_full_b; a.feats = feats;
    a.nreg = c.s.adaptive ? s.nreg : nullptr;
    a.alpha_v = s.alpha_v + (size_t)t * B * R; a.att_img = X2t + 2 * D; a.ld_img = LX2;
    SET_PROPAGATE(attention_fwd(a, st));
  }
  {
    const float* ctx = s.ctx_c + (size_t)t * B * D;
    GemmProblem p[3];
    p[0] = gemm_problem(b, D, s4t, 3 * D);                  // context_gate[:, 2D:3D] ctx (+ h1/word parts)
    gemm_add_seg(p[0], ctx, D, w.ca_gate_w + 2 * D, 3 * D, D);
    p[0].add = s2t + 2 * A; p[0].ldadd = LS2;
    p[1] = gemm_problem(b, D, s4t + D, 3 * D);              // sc_affine(ctx), :380
    gemm_add_seg(p[1], ctx, D, w.ca_sc_w, D, D); p[1].bias = w.ca_sc_b;
    p[2] = gemm_problem(b, D, s4t + 2 * D, 3 * D);          // gate_cmem(sel) (+ both copy-gate biases), :281
    gemm_add_seg(p[2], s.sel + (size_t)t * B * D, D, w.cl_gcm_w, D, D);
    p[2].bias = w.cl_gcm_b; p[2].bias2 = w.cl_gcn_b;
    p[0].c_zeroed = p[1].c_zeroed = p[2].c_zeroed = c.fresh;
    p[0].w_const = p[1].w_const = p[2].w_const = 1;
    SET_PROPAGATE(gemm_group(kNT, p, 3, st));
    SET_PROPAGATE(ctx_gate_fwd(s4t, 3 * D, s2t + 2 * A + D, LS2, s.zst + (size_t)t * B * 3 * D, X2t + D, LX2, b, D, st));
  }
  {
    GemmProblem p = gemm_problem(b, 4 * D, g2t, 4 * D);      // x2h[:, D:] [att_cap | att_img], :272
    gemm_add_seg(p, X2t + D, LX2, w.cl_x2h_w + D, LX2, D + F);
    p.beta = 1; p.w_const = 1;
    int fused = 0;   // copy-LSTM stage 1 (gates, c_new) in the epilogue
    p.epi.op = kEpiCopy1; p.epi.D = D; p.epi.c_prev = s.c2 + (size_t)t * B * D; p.epi.c_out = s.cnew + (size_t)t * B * D;
    p.epi.gates = g2t; p.epi.ld_gates = 4 * D; p.epi_done = &fused;
    SET_PROPAGATE(gemm(kNT, p, st));
    if (!fused) SET_PROPAGATE(copy1_fwd(g2t, s.c2 + (size_t)t * B * D, s.cnew + (size_t)t * B * D, b, D, st));
  }
  {
    GemmProblem p = gemm_problem(b, D, s4t + 2 * D, 3 * D);  // gate_cnew(c_new), :281
    gemm_add_seg(p, s.cnew + (size_t)t * B * D, D, w.cl_gcn_w, D, D);
    p.beta = 1; p.w_const = 1;
    int fused = 0;   // copy gate, c2, h2, dropout(h2) in the epilogue
    p.epi.op = kEpiCopy2; p.epi.D = D; p.epi.gates = g2t; p.epi.ld_gates = 4 * D;
    p.epi.sel = s.sel + (size_t)t * B * D; p.epi.cnew = s.cnew + (size_t)t * B * D;
    p.epi.kgate = s.kgate + (size_t)t * B * D; p.epi.c_out = s.c2 + (size_t)(t + 1) * B * D;
    p.epi.h_out = s.h2 + (size_t)(t + 1) * B * D; p.epi.ld_h = D; p.epi.h2drop = s.h2drop + (size_t)t * B * D;
    p.epi.train = c.s.train; p.epi.seed = c.seed; p.epi.drop_base = (long)t * B * D; p.epi_done = &fused;
    SET_PROPAGATE(gemm(kNT, p, st));
    if (!fused)
      SET_PROPAGATE(copy2_fwd(s4t + 2 * D, 3 * D, g2t, s.sel + (size_t)t * B * D, s.cnew + (size_t)t * B * D,
                              s.kgate + (size_t)t * B * D, s.c2 + (size_t)(t + 1) * B * D,
                              s.h2 + (size_t)(t + 1) * B * D, s.h2drop + (size_t)t * B * D, b, D, c.s.train, c.seed,
                              (long)t * B * D, st));
  }
  return SET_OK;
}

// ------------------------------------------------------------------ persistent decode-step kernel (step_kernel.cu)
// Steps [t0, t0 + nt) in ONE cooperative launch.  *launched = 0 (and SET_OK) when the shape is outside what the
// persistent kernel covers: the caller then runs the launch chain (step_forward).
int steps_persistent(Ctx& c, const float* feats, int t0, int nt, const int* bt, int* launched) {
  *launched = 0;
  static const int on = getenv("SET_STEP_PERSIST") ? atoi(getenv("SET_STEP_PERSIST")) : 1;
  const int B = c.s.B, P = c.s.P, R = c.s.R, T = c.s.T, D = c.d.D, A = c.d.A, F = c.d.F;
  const int LS2 = c.LS2, LX2 = c.LX2;
  if (!on || g_backend != 0 || nt < 1) return SET_OK;
  if (B > 64 || D % 256 != 0 || A % 128 != 0 || F % 256 != 0 || P > 128 || R > 128 || P < 1 || R < 1) return SET_OK;
  int grid = 0, cluster = 0;
  if (!step_kernel_available(&grid, &cluster)) return SET_OK;
  {
    // per-CTA table sizes of the attention phase (step_kernel.cu: kAttnMaxItems / kAttnMaxUnits)
    const int items = (B * (F / 256 + D / 256) + grid - 1) / grid + 2;
    const int nchv = (R + 35) / 36, nchc = (P + 35) / 36;
    if (items > 32 || items * (nchv > nchc ? nchv : nchc) > 96) return SET_OK;
  }
  const SetEditNetParams& w = *c.w;
  Ws& s = c.ws;
  static thread_local StepParams prm;
  memset(&prm, 0, sizeof(prm));
  bool ok = true;
  auto map2 = [&](int idx, const float* ptr, int rows, int cols, int box_rows) {
    const unsigned long long dims[2] = {(unsigned long long)cols, (unsigned long long)rows};
    const unsigned long long st[1] = {(unsigned long long)cols * 4ull};
    const unsigned int box[2] = {32u, (unsigned)box_rows};
    ok = ok && tc_encode_map(&prm.maps[idx], ptr, 2, dims, st, box, true);
  };
  auto map3 = [&](int idx, const float* ptr, int cols, int rows, int outer, long ld_row, long ld_outer, int box_cols, int box_rows,
                  bool swz) {
    const unsigned long long dims[3] = {(unsigned long long)cols, (unsigned long long)rows, (unsigned long long)outer};
    const unsigned long long st[2] = {(unsigned long long)ld_row * 4ull, (unsigned long long)ld_outer * 4ull};
    const unsigned int box[3] = {(unsigned)box_cols, (unsigned)box_rows, 1u};
    ok = ok && tc_encode_map(&prm.maps[idx], ptr, 3, dims, st, box, swz);
  };
  enum { mWih = 0, mWhh, mCaDec, mVaDec, mGate128, mTc, mX2h128, mH2h, mGate64, mSc64, mGcm, mX2h32, mGcn,
         mQh2, mQX2, mQctx, mQsel, mQcnew, mFeats, mPrevH };
  map2(mWih, w.al_wih, 4 * D, 3 * D + F, 32);   map2(mWhh, w.al_whh, 4 * D, D, 32);
  map2(mCaDec, w.ca_dec_w, A, D, 128);          map2(mVaDec, w.va_dec_w, A, D, 128);
  map2(mGate128, w.ca_gate_w, D, 3 * D, 128);   map2(mTc, w.ca_tc_w, D, 2 * D, 128);
  map2(mX2h128, w.cl_x2h_w, 4 * D, LX2, 128);   map2(mH2h, w.cl_h2h_w, 4 * D, D, 128);
  map2(mGate64, w.ca_gate_w, D, 3 * D, 64);     map2(mSc64, w.ca_sc_w, D, D, 64);
  map2(mGcm, w.cl_gcm_w, D, D, 128);            map2(mX2h32, w.cl_x2h_w, 4 * D, LX2, 32);
  map2(mGcn, w.cl_gcn_w, D, D, 128);
  map3(mQh2, s.h2, D, B, T + 1, D, (long)B * D, 32, 64, true);
  map3(mQX2, s.X2, LX2, B, T, LX2, (long)B * LX2, 32, 64, true);
  map3(mQctx, s.ctx_c, D, B, T, D, (long)B * D, 32, 64, true);
  map3(mQsel, s.sel, D, B, T, D, (long)B * D, 32, 64, true);
  map3(mQcnew, s.cnew, D, B, T, D, (long)B * D, 32, 64, true);
  const int chunk_v = R < 36 ? R : 36, chunk_c = P < 36 ? P : 36;
  map3(mFeats, feats, F, R, B, F, (long)R * F, 256, chunk_v, false);
  map3(mPrevH, s.prev_h, D, P, B, D, (long)P * D, 256, chunk_c, false);
  if (!ok) return SET_OK;   // (tensor-core path unavailable: the chain decides what to do)
  auto TP = [](float* p, long st) { TPtr x; x.p = p; x.st = st; return x; };
  auto seg = [&](StepProb& p, int K, int pm0, int pc0, int qm, int qc0, int qtoff, int pm1 = 0, int pc1 = 0) {
    const int sg = p.nseg++;
    p.nkb[sg] = (K + 31) / 32;
    p.pmap[sg][0] = pm0; p.pcol0[sg][0] = pc0; p.pmap[sg][1] = pm1; p.pcol0[sg][1] = pc1;
    p.qmap[sg] = qm; p.qcol0[sg] = qc0; p.qtoff[sg] = qtoff;
  };
  const long BD = (long)B * D;
  int np = 0;
  {  // phase A: attention-LSTM recurrent part (editnet.py:532) + cell
    StepProb& p = prm.prob[np];
    seg(p, D, mWih, 2 * D, mQh2, 0, 0);
    seg(p, D, mWhh, 0, mQX2, 0, -1);
    p.nblk = 4; p.blk_stride = D; p.N = D; p.epi = kSEpiLstm;
    p.C = TP(s.gates1, 4 * BD); p.ldc = 4 * D;
    p.add = TP(s.pre1, 4 * BD); p.ldadd = 4 * D;
    p.a0 = TP(s.c1, BD); p.a1 = TP(s.c1 + BD, BD); p.a2 = TP(s.X2, (long)B * LX2); p.ld0 = LX2;
    prm.phase[0].nprob = 1; prm.phase[0].prob[0] = np++;
  }
  {  // phase B: everything that consumes h1
    StepPhase& ph = prm.phase[1];
    const long s2st = (long)B * LS2;
    StepProb* p = &prm.prob[np];
    seg(*p, D, mCaDec, 0, mQX2, 0, 0); p->nblk = 1; p->N = A; p->bias = w.ca_dec_b; p->C = TP(s.s2, s2st); p->ldc = LS2;
    ph.prob[ph.nprob++] = np++;
    p = &prm.prob[np];
    seg(*p, D, mVaDec, 0, mQX2, 0, 0); p->nblk = 1; p->N = A; p->bias = w.va_dec_b; p->C = TP(s.s2 + A, s2st); p->ldc = LS2;
    ph.prob[ph.nprob++] = np++;
    p = &prm.prob[np];
    seg(*p, D, mGate128, D, mQX2, 0, 0); p->nblk = 1; p->N = D; p->C = TP(s.s2 + 2 * A, s2st); p->ldc = LS2;
    p->add = TP(s.gw, BD); p->ldadd = D;
    ph.prob[ph.nprob++] = np++;
    p = &prm.prob[np];
    seg(*p, D, mTc, D, mQX2, 0, 0); p->nblk = 1; p->N = D; p->C = TP(s.s2 + 2 * A + D, s2st); p->ldc = LS2;
    p->add = TP(s.tw, BD); p->ldadd = D;
    ph.prob[ph.nprob++] = np++;
    p = &prm.prob[np];
    seg(*p, D, mX2h128, 0, mQX2, 0, 0);
    seg(*p, D, mH2h, 0, mQh2, 0, 0);
    p->nblk = 1; p->N = 4 * D; p->bias = w.cl_x2h_b; p->bias2 = w.cl_h2h_b; p->C = TP(s.g2, 4 * BD); p->ldc = 4 * D;
    ph.prob[ph.nprob++] = np++;
  }
  {  // phase D: context gate (two-block tiles), gate_cmem(sel), x2h[:, 2D:] att_img
    StepPhase& ph = prm.phase[2];
    StepProb* p = &prm.prob[np];
    seg(*p, D, mGate64, 2 * D, mQctx, 0, 0, mSc64, 0);
    p->nblk = 2; p->N = D; p->epi = kSEpiCtxGate;
    p->add = TP(s.s2 + 2 * A, (long)B * LS2); p->ldadd = LS2; p->bias2 = w.ca_sc_b;
    p->a0 = TP(s.s2 + 2 * A + D, (long)B * LS2); p->ld0 = LS2;
    p->a1 = TP(s.zst, 3 * BD); p->a2 = TP(s.X2 + D, (long)B * LX2); p->ld1 = LX2;
    ph.prob[ph.nprob++] = np++;
    p = &prm.prob[np];
    seg(*p, D, mGcm, 0, mQsel, 0, 0); p->nblk = 1; p->N = D; p->bias = w.cl_gcm_b; p->bias2 = w.cl_gcn_b;
    p->C = TP(s.s4 + 2 * D, 3 * BD); p->ldc = 3 * D;
    ph.prob[ph.nprob++] = np++;
    // x2h[:, 2D:] att_img: the first half of its K range here, the second half rides with phase E (whose own K is
    // short), so that no CTA of this phase carries more than ~16 K-blocks
    const int Fh = (F / 2) / 32 * 32;
    p = &prm.prob[np];
    seg(*p, Fh, mX2h128, 2 * D, mQX2, 2 * D, 0); p->nblk = 1; p->N = 4 * D; p->C = TP(s.g2, 4 * BD); p->ldc = 4 * D; p->beta = 1;
    ph.prob[ph.nprob++] = np++;
  }
  {  // phase E: x2h[:, D:2D] att_cap (+ the second half of x2h[:, 2D:] att_img) + copy-LSTM stage 1
    const int Fh = (F / 2) / 32 * 32;
    StepProb& p = prm.prob[np];
    seg(p, D, mX2h32, D, mQX2, D, 0);
    seg(p, F - Fh, mX2h32, 2 * D + Fh, mQX2, 2 * D + Fh, 0);
    p.nblk = 4; p.blk_stride = D; p.N = D; p.epi = kSEpiCopy1; p.beta = 1;
    p.C = TP(s.g2, 4 * BD); p.ldc = 4 * D;
    p.a0 = TP(s.c2, BD); p.a1 = TP(s.cnew, BD);
    prm.phase[3].nprob = 1; prm.phase[3].prob[0] = np++;
  }
  {  // phase F: gate_cnew(c_new) + copy gate, c2, h2, dropout(h2)
    StepProb& p = prm.prob[np];
    seg(p, D, mGcn, 0, mQcnew, 0, 0);
    p.nblk = 1; p.N = D; p.epi = kSEpiCopy2; p.beta = 1;
    p.C = TP(s.s4 + 2 * D, 3 * BD); p.ldc = 3 * D;
    p.a0 = TP(s.g2, 4 * BD); p.ld0 = 4 * D;
    p.a1 = TP(s.sel, BD); p.a2 = TP(s.cnew, BD); p.a3 = TP(s.kgate, BD);
    p.a4 = TP(s.c2 + BD, BD); p.a5 = TP(s.h2 + BD, BD); p.a6 = TP(s.h2drop, BD);
    prm.phase[4].nprob = 1; prm.phase[4].prob[0] = np++;
  }
  StepAttn& a = prm.attn;
  a.P = P; a.R = R; a.A = A; a.D = D; a.F = F;
  a.s2 = TP(s.s2, (long)B * LS2); a.ld_s2 = LS2;
  a.att1c = s.att1c; a.att1v = TP(s.att1, c.s.train ? (long)B * R * A : 0);
  a.cap_w = w.ca_full_w; a.cap_b = w.ca_full_b; a.vis_w = w.va_full_w; a.vis_b = w.va_full_b;
  a.mask = s.mask; a.nreg = c.s.adaptive ? s.nreg : nullptr; a.sc = s.att_sc;
  a.map_feats = mFeats; a.map_prevh = mPrevH; a.chunk_v = chunk_v; a.chunk_c = chunk_c;
  a.prev_m = s.prev_m;
  a.alpha_c = TP(s.alpha_c, (long)B * P); a.ctx = TP(s.ctx_c, BD); a.sel = TP(s.sel, BD);
  a.alpha_v = TP(s.alpha_v, (long)B * R); a.att_img = TP(s.X2 + 2 * D, (long)B * LX2); a.ld_img = LX2;
  a.sel_idx = s.sel_idx; a.sel_idx_st = B;
  prm.B = B; prm.D = D; prm.train = c.s.train; prm.seed = c.seed;
  if (step_plan_splits(prm, grid, cluster) != SET_OK) return SET_OK;   // (fewer SMs than a phase has tiles: chain)
  for (int done = 0; done < nt; done += 64) {
    const int n = nt - done < 64 ? nt - done : 64;
    prm.t0 = t0 + done; prm.nt = n;
    for (int k = 0; k < n; ++k) prm.bt[k] = bt[done + k];
    SET_PROPAGATE(step_launch(prm, grid, cluster, c.st));
  }
  *launched = 1;
  return SET_OK;
}

// ------------------------------------------------------------------------ reverse pass
// Data-parallel overlap.  The reverse pass finishes the parameter gradients in five groups, in this order -- the
// Python side lays the flat gradient buffer out in the same order, one contiguous range ("bucket") per group:
//   0  fc.*                                                   before the per-step loop (needs only d logits and the dropped h2)
//   1  attention_lstm.*, copy_lstm.*, cap_features_att.*      after the loop: the two big weight-gradient groups
//   2  embed.*, caption_encoder.*                             input-gradient tail + encoder BPTT
//   3  caption attention, visual decoder_att / full_att       third weight-gradient group
//   4  visual features_att / att_embed                        last (visual feature path)
// When armed (set_backward_bucket_events), backward_core records the caller's event k on its stream as soon as
// bucket k is final; the caller's communication stream waits for event k and all-reduces bucket k underneath the
// rest of the pass.  Only the last bucket (10 of 355 MB) is reduced after the pass.  Measured on 8 B200s
// (tools/dp_timeline.py): every overlapped all-reduce ends well before the next bucket is final; what the overlap
// costs is the SMs and memory bandwidth the collective takes from the kernels it runs beside (the pass is ~0.4 ms
// longer than on one GPU).
// Overwriting reverse pass (set_backward_overwrite_grads): the next backward call initialises the gradient tensors
// itself -- every weight matrix is written by exactly one GEMM (or by GEMMs on disjoint column blocks), so those run
// with beta = 0; only the tensors that are accumulated or scattered into (biases, the embedding table, the two full_att
// rows) are zeroed, in one small launch.  Saves the caller's memset of the whole flat buffer (355 MB) and the
// read-modify-write of every weight gradient.
thread_local bool g_overwrite_armed = false;
constexpr int kMaxBuckets = 8;
thread_local cudaEvent_t g_bucket_ev[kMaxBuckets];
thread_local int g_bucket_n = 0;       // events armed for the next reverse pass
thread_local int g_bucket_next = 0;    // first event not yet recorded
int bucket_notify(int k, cudaStream_t st) {
  // events are recorded in order: bucket k final implies every earlier bucket is final (or absent in this path)
  while (g_bucket_next <= k && g_bucket_next < g_bucket_n) {
    SET_CHECK_CUDA(cudaEventRecord(g_bucket_ev[g_bucket_next], st));
    ++g_bucket_next;
  }
  return SET_OK;
}
// first statement of the backward entry points: whatever armed the call (bucket events, overwrite mode) is disarmed when
// the call returns, on error paths that never reach backward_core too
struct BackwardArms { ~BackwardArms() { g_overwrite_armed = false; g_bucket_n = g_bucket_next = 0; } };
int bucket_finish(cudaStream_t st) {   // end of the pass: whatever was not signalled is final now; disarm
  const int r = bucket_notify(kMaxBuckets, st);
  g_bucket_n = g_bucket_next = 0;
  return r;
}

struct DLogits {
  const float* p; long ld; int inner; long ld_inner; const int* row_len;  // rows are time-major m = t*B + i
};

int backward_core(Ctx& c, const SetEditNetParams& g, const float* feats, const int64_t* caps_tok, long tok_ld,
                  long tok_os, const int64_t* prev, const int64_t* prev_len, const int* bt_host, DLogits dl,
                  const int* dec_len_dev) {
  const int B = c.s.B, P = c.s.P, T = c.s.T, R = c.s.R, D = c.d.D, A = c.d.A, F = c.d.F, V = c.d.V;
  const int LS2 = c.LS2, LX2 = c.LX2;
  const SetEditNetParams& w = *c.w;
  Ws& s = c.ws;
  cudaStream_t st = c.st;
  const int TB = T * B;
  struct BucketGuard { ~BucketGuard() { g_bucket_n = g_bucket_next = 0; } } bucket_guard;   // an error return disarms too
  const bool overwrite = g_overwrite_armed;
  g_overwrite_armed = false;
  if (overwrite) {
    const ZeroJob zj[] = {
        {g.embed, (long)V * D}, {g.enc_x2h_b, 4L * D}, {g.enc_h2h_b, 4L * D}, {g.enc_aff_b, D}, {g.ca_feat_b, A},
        {g.ca_dec_b, A}, {g.ca_full_w, A}, {g.ca_full_b, 1}, {g.ca_gate_b, D}, {g.ca_sc_b, D}, {g.ca_tc_b, D},
        {g.va_emb_b, D}, {g.va_feat_b, A}, {g.va_dec_b, A}, {g.va_full_w, A}, {g.va_full_b, 1}, {g.al_bih, 4L * D},
        {g.al_bhh, 4L * D}, {g.cl_x2h_b, 4L * D}, {g.cl_h2h_b, 4L * D}, {g.cl_gcn_b, D}, {g.cl_gcm_b, D}, {g.fc_b, V},
        {g.al_whh, T > 1 ? 0L : 4L * D * D}};   // (no recurrent step: nothing writes d W_hh)
    SET_PROPAGATE(zero_batch(zj, (int)(sizeof(zj) / sizeof(zj[0])), st));
  }
  const int wbeta = overwrite ? 0 : 1;   // beta of every weight-gradient GEMM
  SET_CHECK_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(s.emb_prev) + s.regionB_begin, 0,
                                 s.regionB_end - s.regionB_begin, st));
  // The dX pass contracts over a weight's OUTPUT features.  The tensor-core kernel wants both operands
  // with the same major-ness (mixed K-major x MN-major tf32 operands read back as zeros on sm_100a), so
  // when it is enabled each weight gets a transposed copy W^T [in][out] and dX = dY @ W becomes the NT
  // form dY @ (W^T)^T.  ~0.3 GB of traffic per train step; the CUDA-core path reads W in place (NN).
  const bool use_wt = (g_backend == 0);
  const int dxm = use_wt ? kNT : kNN;
  if (use_wt) {
    struct { const float* w; float* t; int O, I; } tr[] = {
        {w.fc_w, s.t_fc, V, D}, {w.cl_gcn_w, s.t_cl_gcn, D, D}, {w.cl_gcm_w, s.t_cl_gcm, D, D},
        {w.cl_x2h_w, s.t_cl_x2h, 4 * D, LX2}, {w.ca_gate_w, s.t_ca_gate, D, 3 * D}, {w.ca_sc_w, s.t_ca_sc, D, D},
        {w.ca_dec_w, s.t_ca_dec, A, D}, {w.va_dec_w, s.t_va_dec, A, D}, {w.ca_tc_w, s.t_ca_tc, D, 2 * D},
        {w.al_whh, s.t_al_whh, 4 * D, D}, {w.al_wih, s.t_al_wih, 4 * D, 3 * D + F}, {w.cl_h2h_w, s.t_cl_h2h, 4 * D, D},
        {w.ca_feat_w, s.t_ca_feat, A, D}, {w.va_feat_w, s.t_va_feat, A, D}, {w.enc_aff_w, s.t_enc_aff, D, D},
        {w.enc_h2h_w, s.t_enc_h2h, 4 * D, D}, {w.enc_x2h_w, s.t_enc_x2h, 4 * D, D}};
    TrJob jobs[17];
    int nj = 0;
    for (auto& e : tr) jobs[nj++] = TrJob{e.w, (long)e.I, e.t, (long)e.O, e.O, e.I, 0, 0};
    SET_PROPAGATE(transpose_batch(jobs, nj, st));
  }
  // one dX term: out[m][j] += sum_o dY[m][o] * W[o][c0 + j]   (W is [O][I])
  auto dx = [&](GemmProblem& p, const float* dY, long ldy, const float* W, const float* WTp, int O, int I, int c0) {
    if (use_wt) gemm_add_seg(p, dY, ldy, WTp + (size_t)c0 * O, O, O);
    else gemm_add_seg(p, dY, ldy, W + c0, I, O);
    p.w_const = 1;   // W / W^T are not written again before the optimizer step
  };
  // ---- time-batched tail: weight gradients (dY^T X over all T*B rows; undecoded rows are zero)
  // With the tensor-core engine dW = dY^T X is issued in NT form on explicit transposes of the two
  // activation matrices (K-major operands only, see the note on W^T above); each distinct matrix is
  // transposed once per backward call into `tscratch`.
  struct TrEntry { const float* src; long ld; int rows, cols; const float* dst; };
  std::vector<TrEntry> tr_cache;
  std::vector<TrJob> tr_pending;
  size_t tr_used = 0;
  int tr_err = SET_OK;
  auto TR = [&](const float* X, long ld, int rows, int cols) -> const float* {
    for (const auto& e : tr_cache)
      if (e.src == X && e.ld == ld && e.rows == rows && e.cols == cols) return e.dst;
    const size_t rp = ((size_t)rows + 3) & ~size_t(3);
    if (tr_used + rp * cols > s.tscratch_floats) { tr_err = SET_ERR_WORKSPACE; return nullptr; }
    float* dst = s.tscratch + tr_used;
    tr_used += rp * cols;
    tr_pending.push_back(TrJob{X, ld, dst, (long)rp, rows, cols, 0, 0});   // launched by flush_tr(), batched
    tr_cache.push_back({X, ld, rows, cols, dst});
    return dst;
  };
  auto flush_tr = [&]() -> int {
    if (tr_pending.empty()) return SET_OK;
    const int r = transpose_batch(tr_pending.data(), (int)tr_pending.size(), st);
    tr_pending.clear();
    return r;
  };
  const int dwm = use_wt ? kNT : kTN;
  auto TN = [&](float* C, long ldc, int M, int N, const float* dY, long ldy, const float* X, long ldx, int K) {
    GemmProblem p = gemm_problem(M, N, C, ldc);
    if (use_wt) {
      const long rp = ((long)K + 3) & ~3L;
      gemm_add_seg(p, TR(dY, ldy, K, M), rp, TR(X, ldx, K, N), rp, K);
    } else {
      gemm_add_seg(p, dY, ldy, X, ldx, K);
    }
    p.beta = wbeta;
    return p;
  };
  {  // d(dropout(h2)) for every step at once: dlogits @ fc.weight
    GemmProblem p = gemm_problem(TB, D, s.dh2raw, D);
    dx(p, dl.p, dl.ld, w.fc_w, s.t_fc, V, D, 0);
    p.a_inner = dl.inner; p.a_ld_inner = dl.ld_inner; p.a_row_len = dl.row_len; p.a_valid_inner = B;
    p.w_const = 0;   // directly follows the kernels that write W^T
    SET_PROPAGATE(gemm(dxm, p, st));
  }
  {  // bucket 0: fc.* needs only d logits and the dropped h2 -- first, so that its all-reduce runs under the per-step loop
    const bool dl_plain = (dl.inner == 0 && dl.row_len == nullptr);   // time-major d logits (trainer / rollout)
    if (dl_plain) {
      GemmProblem p = TN(g.fc_w, D, V, D, dl.p, dl.ld, s.h2drop, D, TB);
      SET_REQUIRE(tr_err == SET_OK, "transpose scratch exhausted");
      SET_PROPAGATE(flush_tr());
      SET_PROPAGATE(gemm(dwm, p, st));
      SET_PROPAGATE(colsum(dl.p, dl.ld, TB, V, g.fc_b, 1, st));
    } else {
      // batch-major upstream gradient (autograd drop-in path): strided, masked rows -> CUDA-core TN kernel
      GemmProblem q[2];
      for (int k = 0; k < 2; ++k) {
        q[k] = k == 0 ? gemm_problem(V, D, g.fc_w, D) : gemm_problem(V, 1, g.fc_b, 1);
        gemm_add_seg(q[k], dl.p, dl.ld, k == 0 ? s.h2drop : s.ones, k == 0 ? D : 1, TB);
        q[k].beta = (k == 0) ? wbeta : 1;   // (fc.bias is zeroed with the other biases)
        q[k].a_inner = dl.inner; q[k].a_ld_inner = dl.ld_inner; q[k].a_row_len = dl.row_len; q[k].a_valid_inner = B;
      }
      SET_PROPAGATE(gemm_group(kTN, q, 2, st));
    }
    SET_PROPAGATE(bucket_notify(0, st));
  }
  SET_PROPAGATE(profile_mark(2, st));
  bool copy2_done = false;
  for (int t = T - 1; t >= 0; --t) {
    const int b = bt_host[t];
    if (b <= 0) continue;
    const size_t tb = (size_t)t * B;
    float* dG1t = s.dG1 + tb * 4 * D;
    float* dG2t = s.dG2 + tb * 4 * D;
    float* dS2t = s.dS2 + tb * LS2;
    float* dKt = s.dK + tb * D;
    const float* g2t = s.g2 + tb * 4 * D;
    // per-step slabs (pre-zeroed by the region memset): carries INTO step t live at index t
    float* dX2t = s.dX2 + tb * LX2;
    float* dctxt = s.dctx + tb * D;
    float* dh1c_t = s.dh1c + tb * D;
    float* dh2c_t = s.dh2c + tb * D;
    if (!copy2_done)   // else: applied by the epilogue of the previous iteration's last GEMM
      SET_PROPAGATE(copy2_bwd(dh2c_t, s.dh2raw + tb * D, s.dc2c, g2t, s.c2 + (tb + B) * D, s.kgate + tb * D,
                              s.sel + tb * D, s.cnew + tb * D, dG2t, dKt, s.dsel, s.dcnew, b, D, c.s.train, c.seed,
                              (long)tb * D, st));
    copy2_done = false;
    {
      GemmProblem p[2];
      p[0] = gemm_problem(b, D, s.dcnew, D); dx(p[0], dKt, D, w.cl_gcn_w, s.t_cl_gcn, D, D, 0); p[0].beta = 1;
      p[1] = gemm_problem(b, D, s.dsel, D); dx(p[1], dKt, D, w.cl_gcm_w, s.t_cl_gcm, D, D, 0); p[1].beta = 1;
      // (optional, off: in a two-problem group the fused cell takes the scratch-slab epilogue, which costs more than the
      // stand-alone kernel saves -- 110.6 vs 108.5 us per reverse step on B200)
      int fused = 0;   // copy-LSTM stage-1 backward on the finished d c_new
      static const int fuse_copy1_bwd = getenv("SET_FUSE_COPY1_BWD") ? atoi(getenv("SET_FUSE_COPY1_BWD")) : 0;
      if (fuse_copy1_bwd) p[0].epi.op = kEpiCopy1Bwd; p[0].epi.D = D; p[0].epi.gates = const_cast<float*>(g2t); p[0].epi.ld_gates = 4 * D;
      p[0].epi.c_prev = s.c2 + tb * D; p[0].epi.y0 = s.dc2c; p[0].epi.y1 = dG2t; p[0].epi_done = &fused;
      SET_PROPAGATE(gemm_group(dxm, p, 2, st));
      if (!fused) SET_PROPAGATE(copy1_bwd(s.dcnew, g2t, s.c2 + tb * D, dG2t, s.dc2c, b, D, st));
    }
    {
      GemmProblem p = gemm_problem(b, LX2, dX2t, LX2);      // d[h1 | att_cap | att_img]
      dx(p, dG2t, 4 * D, w.cl_x2h_w, s.t_cl_x2h, 4 * D, LX2, 0);
      p.c_zeroed = c.fresh;
      int fused = 0;   // the context-gate backward rides on the d att_cap columns of this GEMM's output
      p.epi.op = kEpiCtxGateBwd; p.epi.D = D; p.epi.col0 = D; p.epi.x0 = s.zst + tb * 3 * D;
      p.epi.y0 = dS2t + 2 * A; p.epi.y1 = dS2t + 2 * A + D; p.epi.ldy = LS2; p.epi.y2 = s.dsc + tb * D;
      p.epi_done = &fused;
      SET_PROPAGATE(gemm(dxm, p, st));
      if (!fused)
        SET_PROPAGATE(ctx_gate_bwd(s.zst + tb * 3 * D, dX2t + D, LX2, dS2t + 2 * A, dS2t + 2 * A + D, LS2,
                                   s.dsc + tb * D, b, D, st));
    }
    {
      GemmProblem p = gemm_problem(b, D, dctxt, D);
      p.c_zeroed = c.fresh;
      dx(p, dS2t + 2 * A, LS2, w.ca_gate_w, s.t_ca_gate, D, 3 * D, 2 * D);
      dx(p, s.dsc + tb * D, D, w.ca_sc_w, s.t_ca_sc, D, D, 0);
      SET_PROPAGATE(gemm(dxm, p, st));
    }
    {
      AttnBwdArgs a;
      memset(&a, 0, sizeof(a));
      a.b = b; a.P = P; a.R = R; a.D = D; a.A = A; a.F = F;
      a.att1c = s.att1c; a.s2 = s.s2 + tb * LS2; a.ld_s2 = LS2; a.cap_w = w.ca_full_w; a.mask = s.mask;
      a.prev_h = s.prev_h; a.prev_m = s.prev_m; a.alpha_c = s.alpha_c + tb * P; a.sel_idx = s.sel_idx + tb;
      a.dctx = dctxt; a.dsel = s.dsel; a.dprev_h = s.dprev_h; a.dprev_m = s.dprev_m; a.datt1c = s.datt1c;
      a.ds2 = dS2t; a.ld_ds2 = LS2; a.dcap_w = g.ca_full_w; a.dcap_b = g.ca_full_b;
      a.att1v = c.s.train ? s.att1 + tb * R * A : s.att1; a.vis_w = w.va_full_w; a.feats = feats;
      a.nreg = c.s.adaptive ? s.nreg : nullptr; a.alpha_v = s.alpha_v + tb * R;
      a.datt_img = dX2t + 2 * D; a.ld_dimg = LX2;
      a.datt1v = c.s.train ? s.datt1 + tb * R * A : s.datt1; a.datt1v_accum = c.s.train ? 0 : 1;
      a.dvis_w = g.va_full_w; a.dvis_b = g.va_full_b;
      SET_PROPAGATE(attention_bwd(a, st));
    }
    {
      GemmProblem p = gemm_problem(b, D, dX2t, LX2);        // dh1 += every consumer of h1
      dx(p, dS2t, LS2, w.ca_dec_w, s.t_ca_dec, A, D, 0);
      dx(p, dS2t + A, LS2, w.va_dec_w, s.t_va_dec, A, D, 0);
      dx(p, dS2t + 2 * A, LS2, w.ca_gate_w, s.t_ca_gate, D, 3 * D, D);
      dx(p, dS2t + 2 * A + D, LS2, w.ca_tc_w, s.t_ca_tc, D, 2 * D, D);
      p.beta = 1;
      int fused = 0;   // attention-LSTM backward in the epilogue (d h1 is complete once this GEMM has added its terms)
      p.epi.op = kEpiLstmBwd; p.epi.D = D; p.epi.gates = s.gates1 + tb * 4 * D; p.epi.ld_gates = 4 * D;
      p.epi.c_prev = s.c1 + tb * D; p.epi.x1 = s.c1 + (tb + B) * D; p.epi.x0 = dh1c_t; p.epi.y0 = s.dc1c; p.epi.y1 = dG1t;
      p.epi_done = &fused;
      SET_PROPAGATE(gemm(dxm, p, st));
      if (!fused)
        SET_PROPAGATE(lstm_bwd(s.gates1 + tb * 4 * D, s.c1 + tb * D, s.c1 + (tb + B) * D, dX2t, LX2, dh1c_t, s.dc1c,
                               dG1t, b, D, st));
    }
    if (t > 0) {
      GemmProblem p[2];
      p[0] = gemm_problem(b, D, dh1c_t - (size_t)B * D, D); dx(p[0], dG1t, 4 * D, w.al_whh, s.t_al_whh, 4 * D, D, 0);
      p[1] = gemm_problem(b, D, dh2c_t - (size_t)B * D, D);
      p[0].c_zeroed = p[1].c_zeroed = c.fresh;
      dx(p[1], dG1t, 4 * D, w.al_wih, s.t_al_wih, 4 * D, 3 * D + F, 2 * D);
      dx(p[1], dG2t, 4 * D, w.cl_h2h_w, s.t_cl_h2h, 4 * D, D, 0);
      // step t-1 opens with the copy-LSTM stage-2 backward on exactly this GEMM's d h2 carry: it can ride in this
      // epilogue when the two steps decode the same rows (optional, off: the fused path caps the split of this 50 MB
      // GEMM and the reverse step gets 12 us slower)
      int fused = 0;
      static const int fuse_copy2_bwd = getenv("SET_FUSE_COPY2_BWD") ? atoi(getenv("SET_FUSE_COPY2_BWD")) : 0;
      if (fuse_copy2_bwd && bt_host[t - 1] == b) {
        const size_t tp = tb - B;
        GemmEpi& e = p[1].epi;
        e.op = kEpiCopy2Bwd; e.D = D; e.x0 = s.dh2raw + tp * D; e.y0 = s.dc2c; e.gates = s.g2 + tp * 4 * D; e.ld_gates = 4 * D;
        e.x1 = s.c2 + (tp + B) * D; e.kgate = s.kgate + tp * D; e.sel = s.sel + tp * D; e.cnew = s.cnew + tp * D;
        e.y1 = s.dG2 + tp * 4 * D; e.y2 = s.dK + tp * D; e.x2 = s.dsel; e.x3 = s.dcnew;
        e.train = c.s.train; e.seed = c.seed; e.drop_base = (long)tp * D;
        p[1].epi_done = &fused;
      }
      SET_PROPAGATE(gemm_group(dxm, p, 2, st));
      copy2_done = fused != 0;
    }
  }
  SET_PROPAGATE(profile_mark(3, st));
  // ---- time-batched tail.  Order = the order in which the data-parallel step wants the gradient buckets: the two big
  // weight-gradient groups first (their all-reduce then runs under everything that follows), the input-gradient tail and
  // the encoder BPTT -- a chain of small kernels that leaves most SMs to the collective -- next, the attention
  // weights and the visual feature path last.
  SET_PROPAGATE(sum_time(s.dG1, s.sumG1, T, (long)B * 4 * D, st));   // (first: a weight-gradient operand and the d final_hidden source)
  {  // every bias gradient that is a column sum of a per-step buffer
    const ColJob cj[] = {
        {s.dG1, 4L * D, TB, 4 * D, g.al_bih, 0, 0},        {s.dG1, 4L * D, TB, 4 * D, g.al_bhh, 0, 0},
        {s.dG2, 4L * D, TB, 4 * D, g.cl_x2h_b, 0, 0},      {s.dG2, 4L * D, TB, 4 * D, g.cl_h2h_b, 0, 0},
        {s.dS2, (long)LS2, TB, A, g.ca_dec_b, 0, 0},        {s.dS2 + A, (long)LS2, TB, A, g.va_dec_b, 0, 0},
        {s.dS2 + 2 * A, (long)LS2, TB, D, g.ca_gate_b, 0, 0}, {s.dS2 + 2 * A + D, (long)LS2, TB, D, g.ca_tc_b, 0, 0},
        {s.dsc, (long)D, TB, D, g.ca_sc_b, 0, 0},           {s.dK, (long)D, TB, D, g.cl_gcn_b, 0, 0},
        {s.dK, (long)D, TB, D, g.cl_gcm_b, 0, 0},           {s.datt1c, (long)A, B * P, A, g.ca_feat_b, 0, 0}};
    SET_PROPAGATE(colsum_batch(cj, (int)(sizeof(cj) / sizeof(cj[0])), st));
  }
  // weight gradients (dY^T X over all T*B rows; undecoded rows are zero), one group per bucket
  {
    GemmProblem p[8];
    int n = 0;
    if (T > 1) p[n++] = TN(g.al_whh, D, 4 * D, D, s.dG1 + (size_t)B * 4 * D, 4 * D, s.X2, LX2, (T - 1) * B);
    p[n++] = TN(g.al_wih, 3 * D + F, 4 * D, D, s.dG1, 4 * D, s.emb_all, D, TB);
    p[n++] = TN(g.al_wih + D, 3 * D + F, 4 * D, D, s.sumG1, 4 * D, s.fh, D, B);
    p[n++] = TN(g.al_wih + 2 * D, 3 * D + F, 4 * D, D, s.dG1, 4 * D, s.h2, D, TB);
    p[n++] = TN(g.al_wih + 3 * D, 3 * D + F, 4 * D, F, s.sumG1, 4 * D, s.image_mean, F, B);
    SET_PROPAGATE(flush_tr());
    SET_PROPAGATE(gemm_group(dwm, p, n, st));
  }
  {
    GemmProblem p[8];
    int n = 0;
    p[n++] = TN(g.cl_x2h_w, LX2, 4 * D, LX2, s.dG2, 4 * D, s.X2, LX2, TB);
    p[n++] = TN(g.cl_h2h_w, D, 4 * D, D, s.dG2, 4 * D, s.h2, D, TB);
    p[n++] = TN(g.cl_gcn_w, D, D, D, s.dK, D, s.cnew, D, TB);
    p[n++] = TN(g.cl_gcm_w, D, D, D, s.dK, D, s.sel, D, TB);
    p[n++] = TN(g.ca_feat_w, D, A, D, s.datt1c, A, s.prev_h, D, B * P);
    SET_REQUIRE(tr_err == SET_OK, "transpose scratch exhausted");
    SET_PROPAGATE(flush_tr());
    SET_PROPAGATE(gemm_group(dwm, p, n, st));
  }
  SET_PROPAGATE(bucket_notify(1, st));   // attention_lstm.*, copy_lstm.*, cap_features_att.* final
  // ---- time-batched tail: input gradients
  {
    GemmProblem p[2];
    p[0] = gemm_problem(B, D, s.dfh, D);                      // d final_hidden
    dx(p[0], s.sumG1, 4 * D, w.al_wih, s.t_al_wih, 4 * D, 3 * D + F, D);
    p[1] = gemm_problem(TB, D, s.demb_all, D);                // d embeddings (three consumers)
    dx(p[1], s.dG1, 4 * D, w.al_wih, s.t_al_wih, 4 * D, 3 * D + F, 0);
    dx(p[1], s.dS2 + 2 * A, LS2, w.ca_gate_w, s.t_ca_gate, D, 3 * D, 0);
    dx(p[1], s.dS2 + 2 * A + D, LS2, w.ca_tc_w, s.t_ca_tc, D, 2 * D, 0);
    SET_PROPAGATE(gemm_group(dxm, p, 2, st));
  }
  SET_PROPAGATE(embed_bwd(caps_tok, tok_ld, tok_os, s.emb_all, s.demb_all, g.embed, V, T, B, D, c.s.train,
                          dec_len_dev, st));
  {  // d prev_h also flows through cap_features_att
    GemmProblem p = gemm_problem(B * P, D, s.dprev_h, D);
    dx(p, s.datt1c, A, w.ca_feat_w, s.t_ca_feat, A, D, 0);
    p.beta = 1;
    SET_PROPAGATE(gemm(dxm, p, st));
  }
  // ---- caption encoder BPTT (reverse of editnet.py:333-341)
  SET_PROPAGATE(tanh_bwd_inplace(s.dfh, s.fh, (long)B * D, st));
  {
    GemmProblem pw = TN(g.enc_aff_w, D, D, D, s.dfh, D, s.enc_h + (size_t)P * B * D, D, B);
    SET_PROPAGATE(flush_tr());
    SET_PROPAGATE(gemm(dwm, pw, st));
    SET_PROPAGATE(colsum(s.dfh, D, B, D, g.enc_aff_b, 1, st));
    GemmProblem px = gemm_problem(B, D, s.dh_last, D);
    dx(px, s.dfh, D, w.enc_aff_w, s.t_enc_aff, D, D, 0);
    SET_PROPAGATE(gemm(dxm, px, st));
  }
  bool enc_cell_done = false;   // step t's cell already applied by the epilogue of the previous iteration's GEMM
  for (int t = P - 1; t >= 0; --t) {
    const size_t tb = (size_t)t * B;
    if (!enc_cell_done)
      SET_PROPAGATE(enc_lstm_bwd(s.enc_gates + tb * 4 * D, s.enc_c + tb * D, s.enc_c + (tb + B) * D, s.dh_run + tb * D,
                                 s.dc_run, s.dprev_h, s.dprev_m, (long)P * D, s.dh_last, prev_len, t,
                                 s.denc_g + tb * 4 * D, B, D, st));
    enc_cell_done = false;
    if (t > 0) {
      GemmProblem p = gemm_problem(B, D, s.dh_run + (tb - B) * D, D);
      p.c_zeroed = c.fresh;
      dx(p, s.denc_g + tb * 4 * D, 4 * D, w.enc_h2h_w, s.t_enc_h2h, 4 * D, D, 0);
      // this GEMM's output is the d h that step t-1's cell backward starts from: apply that cell in the epilogue
      int fused = 0;
      const size_t tp = tb - B;
      GemmEpi& e = p.epi;
      e.op = kEpiLstmBwd; e.D = D; e.gates = s.enc_gates + tp * 4 * D; e.ld_gates = 4 * D; e.c_prev = s.enc_c + tp * D;
      e.x1 = s.enc_c + (tp + B) * D; e.x0 = nullptr; e.y0 = s.dc_run; e.y1 = s.denc_g + tp * 4 * D;
      e.len = reinterpret_cast<const long long*>(prev_len); e.t = t - 1; e.seq_h = s.dprev_h; e.seq_m = s.dprev_m;
      e.seq_ld = (long)P * D; e.h_prev = s.dh_last;
      p.epi_done = &fused;
      SET_PROPAGATE(gemm(dxm, p, st));
      enc_cell_done = fused != 0;
    }
  }
  {
    GemmProblem p[2];
    p[0] = TN(g.enc_x2h_w, D, 4 * D, D, s.denc_g, 4 * D, s.emb_prev, D, P * B);
    p[1] = TN(g.enc_h2h_w, D, 4 * D, D, s.denc_g, 4 * D, s.enc_h, D, P * B);
    SET_PROPAGATE(flush_tr());
    SET_PROPAGATE(gemm_group(dwm, p, 2, st));
    SET_PROPAGATE(colsum(s.denc_g, 4 * D, P * B, 4 * D, g.enc_x2h_b, 1, st));
    SET_PROPAGATE(colsum(s.denc_g, 4 * D, P * B, 4 * D, g.enc_h2h_b, 1, st));
    GemmProblem px = gemm_problem(P * B, D, s.demb_prev, D);
    dx(px, s.denc_g, 4 * D, w.enc_x2h_w, s.t_enc_x2h, 4 * D, D, 0);
    SET_PROPAGATE(gemm(dxm, px, st));
    SET_PROPAGATE(embed_bwd(prev, c.s.Wp, 1, s.emb_prev, s.demb_prev, g.embed, V, P, B, D, c.s.train, nullptr, st));
  }
  SET_PROPAGATE(bucket_notify(2, st));   // embed.*, caption_encoder.* final
  {
    GemmProblem p[8];
    int n = 0;
    p[n++] = TN(g.ca_dec_w, D, A, D, s.dS2, LS2, s.X2, LX2, TB);
    p[n++] = TN(g.va_dec_w, D, A, D, s.dS2 + A, LS2, s.X2, LX2, TB);
    p[n++] = TN(g.ca_gate_w, 3 * D, D, D, s.dS2 + 2 * A, LS2, s.emb_all, D, TB);
    p[n++] = TN(g.ca_gate_w + D, 3 * D, D, D, s.dS2 + 2 * A, LS2, s.X2, LX2, TB);
    p[n++] = TN(g.ca_gate_w + 2 * D, 3 * D, D, D, s.dS2 + 2 * A, LS2, s.ctx_c, D, TB);
    p[n++] = TN(g.ca_sc_w, D, D, D, s.dsc, D, s.ctx_c, D, TB);
    p[n++] = TN(g.ca_tc_w, 2 * D, D, D, s.dS2 + 2 * A + D, LS2, s.emb_all, D, TB);
    p[n++] = TN(g.ca_tc_w + D, 2 * D, D, D, s.dS2 + 2 * A + D, LS2, s.X2, LX2, TB);
    SET_PROPAGATE(flush_tr());
    SET_PROPAGATE(gemm_group(dwm, p, n, st));
  }
  SET_PROPAGATE(bucket_notify(3, st));   // caption attention, decoder_att / full_att of the visual attention final
  // ---- visual feature path
  if (c.s.train) {
    const int TBR = T * B * R;
    GemmProblem pw = TN(g.va_feat_w, D, A, D, s.datt1, A, s.fe_t, D, TBR);
    SET_PROPAGATE(flush_tr());
    SET_PROPAGATE(gemm(dwm, pw, st));
    SET_PROPAGATE(colsum(s.datt1, A, TBR, A, g.va_feat_b, 1, st));
    GemmProblem px = gemm_problem(TBR, D, s.dfe_t, D);
    dx(px, s.datt1, A, w.va_feat_w, s.t_va_feat, A, D, 0);
    SET_PROPAGATE(gemm(dxm, px, st));
    SET_PROPAGATE(vis_dropout_bwd(s.fe_pre, s.dfe_t, s.dfe_pre, dec_len_dev, T, B, R, D, c.seed, st));
  } else {
    GemmProblem pw = TN(g.va_feat_w, D, A, D, s.datt1, A, s.fe_pre, D, B * R);
    SET_PROPAGATE(flush_tr());
    SET_PROPAGATE(gemm(dwm, pw, st));
    SET_PROPAGATE(colsum(s.datt1, A, B * R, A, g.va_feat_b, 1, st));
    GemmProblem px = gemm_problem(B * R, D, s.dfe_pre, D);
    dx(px, s.datt1, A, w.va_feat_w, s.t_va_feat, A, D, 0);
    SET_PROPAGATE(gemm(dxm, px, st));
    SET_PROPAGATE(relu_bwd_inplace(s.dfe_pre, s.fe_pre, (long)B * R * D, st));
  }
  {
    GemmProblem pw = TN(g.va_emb_w, F, D, F, s.dfe_pre, D, feats, F, B * R);
    SET_PROPAGATE(flush_tr());
    SET_PROPAGATE(gemm(dwm, pw, st));
    SET_PROPAGATE(colsum(s.dfe_pre, D, B * R, D, g.va_emb_b, 1, st));
  }
  SET_REQUIRE(tr_err == SET_OK, "transpose scratch exhausted");
  SET_PROPAGATE(bucket_finish(st));
  return SET_OK;
}

}  // namespace
}  // namespace set

using namespace set;

extern "C" {

size_t set_editnet_workspace_bytes(const SetDims* dims, const SetSeqShape* shape) {
  if (check_args(dims, shape) != SET_OK) return 0;
  Arena ar;
  Ws ws;
  layout(*dims, *shape, ar, ws);
  return ws.total;
}

int set_editnet_workspace_lookup(const SetDims* dims, const SetSeqShape* shape, const char* name, size_t* offset,
                                 size_t* bytes) {
  SET_PROPAGATE(check_args(dims, shape));
  Arena ar;
  Ws ws;
  layout(*dims, *shape, ar, ws);
  for (const auto& e : ar.entries)
    if (e.name == name) { *offset = e.off; *bytes = e.bytes; return SET_OK; }
  set_record_error("unknown workspace buffer name");
  return SET_ERR_ARG;
}

int set_editnet_encode(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w, const int64_t* seq,
                       const int64_t* seq_len, uint64_t seed, float* hidden_states, float* memory_states,
                       float* final_hidden, float* mask, void* workspace, size_t workspace_bytes, void* stream) {
  Ctx c;
  SET_PROPAGATE(make_ctx(c, dims, shape, w, workspace, workspace_bytes, seed, stream));
  SET_REQUIRE(seq && seq_len, "null input");
  SET_PROPAGATE(encode_prev(c, seq, seq_len));
  const size_t B = shape->B, P = shape->P, D = dims->D;
  if (hidden_states)
    SET_CHECK_CUDA(cudaMemcpyAsync(hidden_states, c.ws.prev_h, sizeof(float) * B * P * D, cudaMemcpyDeviceToDevice, c.st));
  if (memory_states)
    SET_CHECK_CUDA(cudaMemcpyAsync(memory_states, c.ws.prev_m, sizeof(float) * B * P * D, cudaMemcpyDeviceToDevice, c.st));
  if (final_hidden)
    SET_CHECK_CUDA(cudaMemcpyAsync(final_hidden, c.ws.fh, sizeof(float) * B * D, cudaMemcpyDeviceToDevice, c.st));
  if (mask) SET_CHECK_CUDA(cudaMemcpyAsync(mask, c.ws.mask, sizeof(float) * B * P, cudaMemcpyDeviceToDevice, c.st));
  return SET_OK;
}

int set_editnet_step_begin(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w,
                           const float* feats, const float* image_mean, const int64_t* prev,
                           const int64_t* prev_len, void* workspace, size_t workspace_bytes, void* stream) {
  Ctx c;
  SET_PROPAGATE(make_ctx(c, dims, shape, w, workspace, workspace_bytes, 0, stream));
  SET_REQUIRE(shape->T == 2 && shape->train == 0, "step sessions use T == 2, eval mode");
  SET_REQUIRE(feats && prev && prev_len, "null input");
  return prepare_common(c, feats, image_mean, prev, prev_len);
}

int set_editnet_step(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w, const float* feats,
                     const int64_t* tokens, int rows, float* h1, float* c1, float* h2, float* c2, float* scores,
                     void* workspace, size_t workspace_bytes, void* stream) {
  Ctx c;
  SET_PROPAGATE(make_ctx(c, dims, shape, w, workspace, workspace_bytes, 0, stream));
  SET_REQUIRE(shape->T == 2 && shape->train == 0, "step sessions use T == 2, eval mode");
  SET_REQUIRE(feats && tokens && h1 && c1 && h2 && c2 && scores && rows >= 1 && rows <= shape->B, "bad args");
  c.fresh = 0;   // the t = 1 slabs are reused by every step of the session
  const size_t B = shape->B, D = dims->D, V = dims->V, LX2 = c.LX2;
  Ws& s = c.ws;
  cudaStream_t st = c.st;
  const size_t row = sizeof(float) * D;
  // state in: the step runs as t = 1, so "previous" state lives in the t = 0 outputs
  SET_CHECK_CUDA(cudaMemcpy2DAsync(s.X2, sizeof(float) * LX2, h1, row, row, rows, cudaMemcpyDeviceToDevice, st));
  SET_CHECK_CUDA(cudaMemcpyAsync(s.c1 + B * D, c1, row * rows, cudaMemcpyDeviceToDevice, st));
  SET_CHECK_CUDA(cudaMemcpyAsync(s.h2 + B * D, h2, row * rows, cudaMemcpyDeviceToDevice, st));
  SET_CHECK_CUDA(cudaMemcpyAsync(s.c2 + B * D, c2, row * rows, cudaMemcpyDeviceToDevice, st));
  SET_PROPAGATE(embed_fwd(tokens, 1, 0, w->embed, dims->V, s.emb_all + B * D, 1, rows, D, 0, 0, kSiteEmb, 0, 0, 1, st));
  SET_PROPAGATE(project_words(c, 1, 1));
  SET_PROPAGATE(step_forward(c, feats, 1, rows));
  GemmProblem p = gemm_problem(rows, V, scores, V);          // fc(h2), editnet.py:653 (eval: dropout is identity)
  gemm_add_seg(p, s.h2drop + B * D, D, w->fc_w, D, D);
  p.bias = w->fc_b;
  SET_PROPAGATE(gemm(kNT, p, st));
  SET_CHECK_CUDA(cudaMemcpy2DAsync(h1, row, s.X2 + B * LX2, sizeof(float) * LX2, row, rows, cudaMemcpyDeviceToDevice, st));
  SET_CHECK_CUDA(cudaMemcpyAsync(c1, s.c1 + 2 * B * D, row * rows, cudaMemcpyDeviceToDevice, st));
  SET_CHECK_CUDA(cudaMemcpyAsync(h2, s.h2 + 2 * B * D, row * rows, cudaMemcpyDeviceToDevice, st));
  SET_CHECK_CUDA(cudaMemcpyAsync(c2, s.c2 + 2 * B * D, row * rows, cudaMemcpyDeviceToDevice, st));
  return SET_OK;
}

static int xe_forward_impl(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w,
                           const float* feats, const float* image_mean, const int64_t* caps,
                           const int* decode_len_host, const int64_t* prev, const int64_t* prev_len, uint64_t seed,
                           float ss_prob, const int64_t* ss_replay, int64_t* fed_tokens, float* predictions,
                           void* workspace, size_t workspace_bytes, void* stream) {
  Ctx c;
  SET_PROPAGATE(make_ctx(c, dims, shape, w, workspace, workspace_bytes, seed, stream));
  SET_REQUIRE(feats && caps && decode_len_host && prev && prev_len, "null input");
  SET_REQUIRE(predictions != nullptr || shape->train, "time-major logits live in the train-mode workspace");
  SET_REQUIRE(shape->Wc > shape->T, "caption width must exceed T");
  const bool ss = (ss_prob > 0.f) || (ss_replay != nullptr);
  SET_REQUIRE(!ss || (shape->train && fed_tokens != nullptr), "scheduled sampling needs train mode and fed_tokens");
  std::vector<int> bt;
  SET_PROPAGATE(batch_sizes(*shape, decode_len_host, bt));
  const int B = shape->B, T = shape->T, D = dims->D, V = dims->V;
  Ws& s = c.ws;
  SET_PROPAGATE(prepare_common(c, feats, image_mean, prev, prev_len));
  SET_CHECK_CUDA(cudaMemcpyAsync(s.dec_len, decode_len_host, sizeof(int) * B, cudaMemcpyHostToDevice, c.st));
  if (!ss) {
    SET_PROPAGATE(embed_fwd(caps, shape->Wc, 1, w->embed, V, s.emb_all, T, B, D, shape->train, seed, kSiteEmb, 0, B, 1,
                            c.st));
    SET_PROPAGATE(project_words(c, 0, T));
    SET_PROPAGATE(profile_mark(0, c.st));
    int persistent = 0;
    {
      int nsteps = 0;   // steps that decode at least one row (bt is non-increasing)
      while (nsteps < T && bt[nsteps] > 0) ++nsteps;
      SET_PROPAGATE(steps_persistent(c, feats, 0, nsteps, bt.data(), &persistent));
    }
    if (!persistent)
      for (int t = 0; t < T; ++t) SET_PROPAGATE(step_forward(c, feats, t, bt[t]));
    SET_PROPAGATE(profile_mark(1, c.st));
  } else {
    // scheduled sampling (editnet.py:508-520): step t's input may be drawn from step t-1's scores, so
    // the word projections and the vocabulary projection run per step, as in a rollout
    SET_CHECK_CUDA(cudaMemcpy2DAsync(fed_tokens, sizeof(int64_t) * shape->Wc, caps, sizeof(int64_t) * shape->Wc,
                                     sizeof(int64_t) * shape->Wc, B, cudaMemcpyDeviceToDevice, c.st));
    for (int t = 0; t < T; ++t) {
      const int b = bt[t];
      ss_choose_kernel<<<b, 256, 0, c.st>>>(t > 0 ? s.logits + (size_t)(t - 1) * B * V : s.logits, V, B, shape->Wc, t,
                                            ss_prob, seed, caps, ss_replay, s.it + (size_t)t * B, fed_tokens);
      SET_CHECK_CUDA(cudaGetLastError());
      set_count_launch(1);
      SET_PROPAGATE(embed_fwd(s.it + (size_t)t * B, 1, 0, w->embed, V, s.emb_all + (size_t)t * B * D, 1, b, D,
                              shape->train, seed, kSiteEmb, (long)t * B, 0, 1, c.st));
      SET_PROPAGATE(project_words(c, t, 1));
      SET_PROPAGATE(step_forward(c, feats, t, b));
      GemmProblem p = gemm_problem(b, V, s.logits + (size_t)t * B * V, V);
      gemm_add_seg(p, s.h2drop + (size_t)t * B * D, D, w->fc_w, D, D);
      p.bias = w->fc_b; p.c_zeroed = c.fresh;
      SET_PROPAGATE(gemm(kNT, p, c.st));
    }
    if (predictions == nullptr) return SET_OK;   // time-major logits are already in place
  }
  // vocabulary projection for all decoded rows at once (editnet.py:545-546)
  if (predictions == nullptr) {
    // fused-trainer form: logits stay time-major [T][B][V] in the workspace (one dense GEMM operand later)
    GemmProblem p = gemm_problem(T * B, V, s.logits, V);
    gemm_add_seg(p, s.h2drop, D, w->fc_w, D, D);
    p.bias = w->fc_b; p.c_zeroed = c.fresh;
    SET_PROPAGATE(gemm(kNT, p, c.st));
    return SET_OK;
  }
  SET_CHECK_CUDA(cudaMemsetAsync(predictions, 0, sizeof(float) * (size_t)B * T * V, c.st));
  GemmProblem p = gemm_problem(T * B, V, predictions, V);   // written batch-major, undecoded rows stay zero
  gemm_add_seg(p, s.h2drop, D, w->fc_w, D, D);
  p.bias = w->fc_b;
  p.c_inner = B; p.c_ld_inner = (long)T * V; p.c_row_len = s.dec_len; p.c_valid_inner = B;
  SET_PROPAGATE(gemm(kNT, p, c.st));
  return SET_OK;
}

int set_editnet_xe_forward(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w,
                           const float* feats, const float* image_mean, const int64_t* caps,
                           const int* decode_len_host, const int64_t* prev, const int64_t* prev_len,
                           uint64_t seed, float* predictions, void* workspace, size_t workspace_bytes,
                           void* stream) {
  return xe_forward_impl(dims, shape, w, feats, image_mean, caps, decode_len_host, prev, prev_len, seed, 0.f, nullptr,
                         nullptr, predictions, workspace, workspace_bytes, stream);
}

int set_editnet_xe_forward_ss(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w,
                              const float* feats, const float* image_mean, const int64_t* caps,
                              const int* decode_len_host, const int64_t* prev, const int64_t* prev_len,
                              uint64_t seed, float ss_prob, const int64_t* ss_replay, int64_t* fed_tokens,
                              float* predictions, void* workspace, size_t workspace_bytes, void* stream) {
  return xe_forward_impl(dims, shape, w, feats, image_mean, caps, decode_len_host, prev, prev_len, seed, ss_prob,
                         ss_replay, fed_tokens, predictions, workspace, workspace_bytes, stream);
}

int set_editnet_xe_backward(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w,
                            const SetEditNetParams* grads, const float* feats, const int64_t* caps,
                            const int* decode_len_host, const int64_t* prev, const int64_t* prev_len,
                            uint64_t seed, const float* d_predictions, void* workspace,
                            size_t workspace_bytes, void* stream) {
  BackwardArms arms_guard;
  Ctx c;
  SET_PROPAGATE(make_ctx(c, dims, shape, w, workspace, workspace_bytes, seed, stream));
  SET_REQUIRE(grads && feats && caps && decode_len_host && prev && prev_len, "null input");
  std::vector<int> bt;
  SET_PROPAGATE(batch_sizes(*shape, decode_len_host, bt));
  DLogits dl;
  if (d_predictions == nullptr) {
    // d logits were written time-major over the workspace logits by set_xe_loss_time_major()
    SET_REQUIRE(shape->train, "time-major logits live in the train-mode workspace");
    dl.p = c.ws.logits; dl.ld = dims->V; dl.inner = 0; dl.ld_inner = 0; dl.row_len = nullptr;
    return backward_core(c, *grads, feats, caps, shape->Wc, 1, prev, prev_len, bt.data(), dl, c.ws.dec_len);
  }
  dl.p = d_predictions; dl.ld = dims->V; dl.inner = shape->B; dl.ld_inner = (long)shape->T * dims->V;
  dl.row_len = c.ws.dec_len;  // uploaded by the forward call
  return backward_core(c, *grads, feats, caps, shape->Wc, 1, prev, prev_len, bt.data(), dl, c.ws.dec_len);
}

int set_xe_loss(int B, int T, int V, int Wc, const float* predictions, const int64_t* caps,
                const int* decode_len_dev, float inv_count, float* loss_out, float* d_predictions,
                void* stream) {
  SET_REQUIRE(B > 0 && T > 0 && V > 1 && Wc > T && predictions && caps && decode_len_dev && loss_out, "bad args");
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  SET_CHECK_CUDA(cudaMemsetAsync(loss_out, 0, 2 * sizeof(float), st));
  xe_loss_kernel<<<B * T, 256, 0, st>>>(B, T, V, Wc, (long)T * V, (long)V, predictions, caps, decode_len_dev,
                                        inv_count, loss_out, d_predictions);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  xe_count_kernel<<<1, 32, 0, st>>>(B, T, decode_len_dev, loss_out);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  return SET_OK;
}

int set_editnet_xe_loss_time_major(const SetDims* dims, const SetSeqShape* shape, const int64_t* caps,
                                   float inv_count, float* loss_out, void* workspace, size_t workspace_bytes,
                                   void* stream) {
  Ctx c;
  SetEditNetParams dummy;
  SET_PROPAGATE(make_ctx(c, dims, shape, &dummy, workspace, workspace_bytes, 0, stream));
  SET_REQUIRE(shape->train && caps && loss_out, "bad args");
  const int B = shape->B, T = shape->T, V = dims->V;
  SET_CHECK_CUDA(cudaMemsetAsync(loss_out, 0, 2 * sizeof(float), c.st));
  xe_loss_kernel<<<B * T, 256, 0, c.st>>>(B, T, V, shape->Wc, (long)V, (long)B * V, c.ws.logits, caps, c.ws.dec_len,
                                          inv_count, loss_out, c.ws.logits);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  xe_count_kernel<<<1, 32, 0, c.st>>>(B, T, c.ws.dec_len, loss_out);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  return SET_OK;
}

int set_editnet_rollout(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w,
                        const float* feats, const float* image_mean, const int64_t* prev,
                        const int64_t* prev_len, int64_t start_token, int64_t end_token, int mode,
                        const int64_t* forced, uint64_t seed, int64_t* seq, float* seq_logprobs,
                        void* workspace, size_t workspace_bytes, void* stream) {
  Ctx c;
  SET_PROPAGATE(make_ctx(c, dims, shape, w, workspace, workspace_bytes, seed, stream));
  SET_REQUIRE(feats && prev && prev_len && seq && seq_logprobs, "null input");
  SET_REQUIRE(mode >= 0 && mode <= 2 && (mode != 2 || forced != nullptr), "bad mode");
  const int B = shape->B, T = shape->T, D = dims->D, V = dims->V;
  Ws& s = c.ws;
  SET_PROPAGATE(prepare_common(c, feats, image_mean, prev, prev_len));
  SET_CHECK_CUDA(cudaMemsetAsync(seq, 0, sizeof(int64_t) * (size_t)B * T, c.st));
  SET_CHECK_CUDA(cudaMemsetAsync(seq_logprobs, 0, sizeof(float) * (size_t)B * T, c.st));
  fill_i64_kernel<<<1, 256, 0, c.st>>>(s.it, B, start_token);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  for (int t = 0; t < T; ++t) {
    SET_PROPAGATE(embed_fwd(s.it + (size_t)t * B, 1, 0, w->embed, V, s.emb_all + (size_t)t * B * D, 1, B, D,
                            shape->train, seed, kSiteEmb, (long)t * B, 0, 1, c.st));
    SET_PROPAGATE(project_words(c, t, 1));
    SET_PROPAGATE(step_forward(c, feats, t, B));
    float* lg = shape->train ? s.logits + (size_t)t * B * V : s.logits;
    GemmProblem p = gemm_problem(B, V, lg, V);
    gemm_add_seg(p, s.h2drop + (size_t)t * B * D, D, w->fc_w, D, D);
    p.bias = w->fc_b;
    SET_PROPAGATE(gemm(kNT, p, c.st));
    sample_step_kernel<<<B, 256, 0, c.st>>>(lg, V, B, T, t, mode, forced, seed, end_token, s.unfinished,
                                            s.unf_count, s.it + (size_t)(t + 1) * B, seq, seq_logprobs, s.lse, s.tok_raw);
    SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  }
  return SET_OK;
}

int set_editnet_rollout_backward(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w,
                                 const SetEditNetParams* grads, const float* feats, const int64_t* prev,
                                 const int64_t* prev_len, uint64_t seed, const float* d_seq_logprobs,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  BackwardArms arms_guard;
  Ctx c;
  SET_PROPAGATE(make_ctx(c, dims, shape, w, workspace, workspace_bytes, seed, stream));
  SET_REQUIRE(shape->train, "rollout backward needs a train-mode forward (activations kept)");
  SET_REQUIRE(grads && feats && prev && prev_len && d_seq_logprobs, "null input");
  const int B = shape->B, T = shape->T, V = dims->V;
  Ws& s = c.ws;
  rollout_dlogits_kernel<<<T * B, 256, 0, c.st>>>(s.logits, V, B, T, d_seq_logprobs, s.lse, s.tok_raw);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  std::vector<int> bt(T, B);
  DLogits dl;
  dl.p = s.logits; dl.ld = V; dl.inner = 0; dl.ld_inner = 0; dl.row_len = nullptr;
  // the embedding of step t was looked up from the token fed at step t (`it[t]`: <start>, then the
  // previous step's output token after the <end>/finished rewrite, editnet_rl.py:531-540)
  return backward_core(c, *grads, feats, s.it, 1, B, prev, prev_len, bt.data(), dl, nullptr);
}

int set_reward_criterion(int B, int T, const float* seq_logprobs, const int64_t* seq, const float* reward,
                         float* loss_out, float* d_logprobs, void* stream) {
  SET_REQUIRE(B > 0 && T > 0 && seq_logprobs && seq && reward && loss_out, "bad args");
  reward_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(B, T, seq_logprobs, seq, reward, loss_out,
                                                                       d_logprobs);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  return SET_OK;
}

int set_profile_enable(int on) {
  g_profile = on != 0;
  g_ev_state = 0;
  return SET_OK;
}

int set_profile_read(float* fwd_loop_ms, float* bwd_loop_ms) {
  if (fwd_loop_ms) *fwd_loop_ms = -1.f;
  if (bwd_loop_ms) *bwd_loop_ms = -1.f;
  if ((g_ev_state & 1) && fwd_loop_ms) {
    SET_CHECK_CUDA(cudaEventSynchronize(g_ev[1]));
    SET_CHECK_CUDA(cudaEventElapsedTime(fwd_loop_ms, g_ev[0], g_ev[1]));
  }
  if ((g_ev_state & 2) && bwd_loop_ms) {
    SET_CHECK_CUDA(cudaEventSynchronize(g_ev[3]));
    SET_CHECK_CUDA(cudaEventElapsedTime(bwd_loop_ms, g_ev[2], g_ev[3]));
  }
  return SET_OK;
}

int set_dropout_keep_mask(float* out, size_t n, uint64_t seed, int site, size_t base, void* stream) {
  SET_REQUIRE(out != nullptr, "null out");
  return dropout_keep_mask(out, (long)n, seed, (uint32_t)site, (long)base, reinterpret_cast<cudaStream_t>(stream));
}

int set_gemm_trace(void* buf) {
  gemm_tc_set_trace(reinterpret_cast<unsigned long long*>(buf));
  return SET_OK;
}

int set_gemm_trace_seq(void* buf, long stride_u64, int launches) {
  gemm_tc_set_trace_seq(reinterpret_cast<unsigned long long*>(buf), stride_u64, launches);
  return SET_OK;
}

int set_gemm_backend(int backend) {
  g_backend = backend;
  return SET_OK;
}

int set_gemm_stats(long long* tc_launches, long long* simt_launches, int reset) {
  if (tc_launches) *tc_launches = g_tc_launches;
  if (simt_launches) *simt_launches = g_simt_launches;
  if (reset) g_tc_launches = g_simt_launches = 0;
  return SET_OK;
}

int set_step_stats(long long* launches, long long* steps, int reset) {
  if (launches) *launches = g_step_launches;
  if (steps) *steps = g_step_steps;
  if (reset) g_step_launches = g_step_steps = 0;
  return SET_OK;
}

int set_backward_overwrite_grads(int on) {
  g_overwrite_armed = on != 0;
  return SET_OK;
}

int set_backward_bucket_events(void* const* events, int n) {
  SET_REQUIRE(n >= 0 && n <= kMaxBuckets && (n == 0 || events), "0..8 events");
  for (int k = 0; k < n; ++k) {
    SET_REQUIRE(events[k], "null event");
    g_bucket_ev[k] = reinterpret_cast<cudaEvent_t>(events[k]);
  }
  g_bucket_n = n;
  g_bucket_next = 0;
  return SET_OK;
}

int set_step_geometry(int* grid, int* cluster) {
  int g = 0, c = 0;
  const bool ok = step_kernel_available(&g, &c);
  if (grid) *grid = ok ? g : 0;
  if (cluster) *cluster = ok ? c : 0;
  return SET_OK;
}

int set_step_trace(void* buf) {
  step_set_trace(reinterpret_cast<unsigned long long*>(buf));
  return SET_OK;
}

long long set_gemm_twin_launches(int reset) {
  const long long v = g_tc_twin_launches;
  if (reset) g_tc_twin_launches = 0;
  return v;
}

int set_gemm(int mode, int M, int N, int K, const float* A, long lda, const float* Bm, long ldb,
             const float* bias, float* C, long ldc, int beta, int act, void* stream) {
  GemmProblem p = gemm_problem(M, N, C, ldc);
  gemm_add_seg(p, A, lda, Bm, ldb, K);
  p.bias = bias; p.beta = beta; p.act = act;
  return gemm(mode, p, reinterpret_cast<cudaStream_t>(stream));
}

}  // extern "C"
