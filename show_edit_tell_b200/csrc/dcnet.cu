// Host orchestration + C ABI of the DCNet (text-only denoising auto-encoder) path.
// Reference: DAE.forward dcnet.py:303-350, dcnet_rl.py:286-346; cells dcnet.py:147-270.
// DCNet is the EditNet step without image attention, context gate, memory select and copy gate:
//   h1,c1 = LSTMCell([emb; final_hidden; h2]) ; ctx = soft-attention over the bi-LSTM encoding ;
//   h2,c2 = LSTMCell([h1; ctx]) ; logits = fc(dropout(h2)).
// Layout conventions are those of editnet.cu (time-major saved activations, X2[t] = [h1 | ctx]).
#include "seq_common.cuh"

namespace set {
namespace {

struct DWs {
  float *emb_prev, *xg_f, *xg_r, *hh_pre, *enc_h_f, *enc_c_f, *enc_h_r, *enc_c_r, *gates_f, *gates_r, *enc_out, *hcat,
      *mask, *fh, *att1c;
  float *pre1s, *emb_all, *pre1, *gates1, *c1, *X2, *s2, *g2, *alpha_c, *c2, *h2, *h2drop, *logits, *lse, *scratch4d,
      *ones;
  int *dec_len, *unfinished, *unf_count;
  int64_t *it, *tok_raw;
  float *dG1, *dS2, *dG2, *dh2raw, *datt1c, *denc_out, *dfh, *dhcat, *demb_all, *demb_prev, *dgates_f, *dgates_r, *dxg_f,
      *dxg_r, *dh_run, *dc_run, *sumG1, *dh2c, *dc2c, *dh1c, *dc1c, *dX2;
  size_t regionA_end = 0, regionB_begin = 0, total = 0;
};

struct DCtx {
  SetDims d;
  SetSeqShape s;
  const SetDcNetParams* w;
  DWs ws;
  cudaStream_t st;
  uint64_t seed;
  int C;    // encoder hidden per direction (caption_features_dim), D == 2*C
  int LX2;  // D + 2C
};

void dlayout(const SetDims& d, const SetSeqShape& s, Arena& ar, DWs& w) {
  const size_t B = s.B, P = s.P, T = s.T, D = d.D, A = d.A, V = d.V, C = d.D / 2;
  const size_t Tv = s.train ? T : 1;
  w.emb_prev = ar.take<float>("emb_prev", B * P * D);
  w.xg_f = ar.take<float>("xg_f", B * P * 4 * C);
  w.xg_r = ar.take<float>("xg_r", B * P * 4 * C);
  w.hh_pre = ar.take<float>("hh_pre", B * 4 * C);
  w.enc_h_f = ar.take<float>("enc_h_f", (P + 1) * B * C);
  w.enc_c_f = ar.take<float>("enc_c_f", (P + 1) * B * C);
  w.enc_h_r = ar.take<float>("enc_h_r", (P + 1) * B * C);
  w.enc_c_r = ar.take<float>("enc_c_r", (P + 1) * B * C);
  w.gates_f = ar.take<float>("gates_f", P * B * 4 * C);
  w.gates_r = ar.take<float>("gates_r", P * B * 4 * C);
  w.enc_out = ar.take<float>("enc_out", B * P * 2 * C);
  w.hcat = ar.take<float>("hcat", B * 2 * C);
  w.mask = ar.take<float>("mask", B * P);
  w.fh = ar.take<float>("final_hidden", B * 2 * C);
  w.att1c = ar.take<float>("att1c", B * P * A);
  w.pre1s = ar.take<float>("pre1s", B * 4 * D);
  w.emb_all = ar.take<float>("emb_all", T * B * D);
  w.pre1 = ar.take<float>("pre1", T * B * 4 * D);
  w.gates1 = ar.take<float>("gates1", T * B * 4 * D);
  w.c1 = ar.take<float>("c1", (T + 1) * B * D);
  w.X2 = ar.take<float>("X2", T * B * (D + 2 * C));
  w.s2 = ar.take<float>("s2", T * B * A);
  w.g2 = ar.take<float>("g2", T * B * 4 * D);
  w.alpha_c = ar.take<float>("alpha_c", T * B * P);
  w.c2 = ar.take<float>("c2", (T + 1) * B * D);
  w.h2 = ar.take<float>("h2", (T + 1) * B * D);
  w.h2drop = ar.take<float>("h2drop", T * B * D);
  w.logits = ar.take<float>("logits", Tv * B * V);
  w.lse = ar.take<float>("lse", T * B);
  w.scratch4d = ar.take<float>("scratch4d", B * 4 * D);
  w.ones = ar.take<float>("ones", T * B);
  w.dec_len = ar.take<int>("dec_len", B);
  w.unfinished = ar.take<int>("unfinished", B);
  w.unf_count = ar.take<int>("unf_count", T + 2);
  w.it = ar.take<int64_t>("it", (T + 1) * B);
  w.tok_raw = ar.take<int64_t>("tok_raw", T * B);
  w.regionA_end = ar.off;
  w.regionB_begin = ar.off;
  w.dG1 = ar.take<float>("dG1", T * B * 4 * D);
  w.dS2 = ar.take<float>("dS2", T * B * A);
  w.dG2 = ar.take<float>("dG2", T * B * 4 * D);
  w.dh2raw = ar.take<float>("dh2raw", T * B * D);
  w.datt1c = ar.take<float>("datt1c", B * P * A);
  w.denc_out = ar.take<float>("denc_out", B * P * 2 * C);
  w.dfh = ar.take<float>("dfh", B * 2 * C);
  w.dhcat = ar.take<float>("dhcat", B * 2 * C);
  w.demb_all = ar.take<float>("demb_all", T * B * D);
  w.demb_prev = ar.take<float>("demb_prev", B * P * D);
  w.dgates_f = ar.take<float>("dgates_f", P * B * 4 * C);
  w.dgates_r = ar.take<float>("dgates_r", P * B * 4 * C);
  w.dxg_f = ar.take<float>("dxg_f", B * P * 4 * C);
  w.dxg_r = ar.take<float>("dxg_r", B * P * 4 * C);
  w.dh_run = ar.take<float>("dh_run", B * C);
  w.dc_run = ar.take<float>("dc_run", B * C);
  w.sumG1 = ar.take<float>("sumG1", B * 4 * D);
  w.dh2c = ar.take<float>("dh2c", B * D);
  w.dc2c = ar.take<float>("dc2c", B * D);
  w.dh1c = ar.take<float>("dh1c", B * D);
  w.dc1c = ar.take<float>("dc1c", B * D);
  w.dX2 = ar.take<float>("dX2", B * (D + 2 * C));
  w.total = ar.off;
}

int dcheck(const SetDims* d, const SetSeqShape* s) {
  SET_REQUIRE(d && s, "null dims/shape");
  SET_REQUIRE(d->D > 0 && d->D % 8 == 0 && d->A > 0 && d->A % 4 == 0 && d->V > 1,
              "DCNet: decoder_dim must be a multiple of 8 (= 2 * caption_features_dim), A a multiple of 4");
  SET_REQUIRE(s->B > 0 && s->Wp > 0 && s->P > 0 && s->P <= s->Wp && s->T > 0, "bad sequence shape");
  return SET_OK;
}

int make_dctx(DCtx& c, const SetDims* d, const SetSeqShape* s, const SetDcNetParams* w, void* workspace,
              size_t workspace_bytes, uint64_t seed, void* stream) {
  SET_PROPAGATE(dcheck(d, s));
  SET_REQUIRE(w != nullptr && workspace != nullptr, "null params/workspace");
  SET_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "workspace must be 256-byte aligned");
  c.d = *d; c.s = *s; c.w = w; c.seed = seed;
  c.st = reinterpret_cast<cudaStream_t>(stream);
  Arena ar;
  ar.base = reinterpret_cast<char*>(workspace);
  dlayout(*d, *s, ar, c.ws);
  if (c.ws.total > workspace_bytes) {
    set_record_error("workspace too small: see set_dcnet_workspace_bytes()");
    return SET_ERR_WORKSPACE;
  }
  c.C = d->D / 2;
  c.LX2 = d->D + 2 * c.C;
  return SET_OK;
}

// bi-LSTM caption encoder (dcnet.py:220-243) + hoisted projections
int dprepare(DCtx& c, const int64_t* prev, const int64_t* prev_len) {
  const int B = c.s.B, P = c.s.P, T = c.s.T, D = c.d.D, A = c.d.A, C = c.C;
  const SetDcNetParams& w = *c.w;
  DWs& s = c.ws;
  cudaStream_t st = c.st;
  SET_CHECK_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(s.emb_prev), 0, s.regionA_end, st));
  fill_kernel<<<8, 256, 0, st>>>(s.ones, (long)T * B, 1.0f);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  // embeddings of the previous caption, batch-major [B][P][D] (keep bits indexed (i*Wp + p)*D + e)
  SET_PROPAGATE(embed_fwd(prev, 1, c.s.Wp, w.embed, c.d.V, s.emb_prev, B, P, D, c.s.train, c.seed, kSiteEnc, 0,
                          c.s.Wp, 1, st));
  {
    GemmProblem p[2];
    p[0] = gemm_problem(B * P, 4 * C, s.xg_f, 4 * C);
    gemm_add_seg(p[0], s.emb_prev, D, w.enc_wih_f, D, D); p[0].bias = w.enc_bih_f; p[0].bias2 = w.enc_bhh_f;
    p[1] = gemm_problem(B * P, 4 * C, s.xg_r, 4 * C);
    gemm_add_seg(p[1], s.emb_prev, D, w.enc_wih_r, D, D); p[1].bias = w.enc_bih_r; p[1].bias2 = w.enc_bhh_r;
    SET_PROPAGATE(gemm_group(kNT, p, 2, st));
  }
  for (int dir = 0; dir < 2; ++dir) {
    float* hs = dir ? s.enc_h_r : s.enc_h_f;
    float* cs = dir ? s.enc_c_r : s.enc_c_f;
    float* gs = dir ? s.gates_r : s.gates_f;
    const float* whh = dir ? w.enc_whh_r : w.enc_whh_f;
    for (int t = 0; t < P; ++t) {
      const size_t tb = (size_t)t * B;
      if (t > 0) {
        GemmProblem p = gemm_problem(B, 4 * C, s.hh_pre, 4 * C);
        gemm_add_seg(p, hs + tb * C, C, whh, C, C);
        SET_PROPAGATE(gemm(kNT, p, st));
      }
      SET_PROPAGATE(bilstm_fwd(t > 0 ? s.hh_pre : nullptr, dir ? s.xg_r : s.xg_f, prev_len, t, dir, hs + tb * C,
                               cs + tb * C, hs + (tb + B) * C, cs + (tb + B) * C, gs + tb * 4 * C,
                               s.enc_out + dir * C, (long)P * 2 * C, 2 * C, B, P, C, st));
    }
  }
  SET_PROPAGATE(enc_mask(s.enc_out, s.mask, B, P, 2 * C, st));     // outputs.sum(2) != 0, dcnet.py:239
  {
    // final_hidden = tanh(concat([h_fwd_last ; h_bwd_last])), dcnet.py:241-242
    GemmProblem p = gemm_problem(B, 2 * C, s.fh, 2 * C);
    gemm_add_seg(p, s.enc_h_f + (size_t)P * B * C, C, w.enc_cat_w, 2 * C, C);
    gemm_add_seg(p, s.enc_h_r + (size_t)P * B * C, C, w.enc_cat_w + C, 2 * C, C);
    p.bias = w.enc_cat_b; p.act = 2;
    SET_PROPAGATE(gemm(kNT, p, st));
    GemmProblem q = gemm_problem(B * P, A, s.att1c, A);               // cap_features_att, time-invariant
    gemm_add_seg(q, s.enc_out, 2 * C, w.ca_feat_w, 2 * C, 2 * C);
    q.bias = w.ca_feat_b;
    SET_PROPAGATE(gemm(kNT, q, st));
    GemmProblem r = gemm_problem(B, 4 * D, s.pre1s, 4 * D);           // W_ih[:, D:2D] final_hidden + biases
    gemm_add_seg(r, s.fh, 2 * C, w.al_wih + D, 3 * D, 2 * C);
    r.bias = w.al_bih; r.bias2 = w.al_bhh;
    SET_PROPAGATE(gemm(kNT, r, st));
  }
  return SET_OK;
}

int dproject_words(DCtx& c, int t0, int nt) {
  const int B = c.s.B, D = c.d.D;
  DWs& s = c.ws;
  const size_t r0 = (size_t)t0 * B;
  GemmProblem p = gemm_problem(nt * B, 4 * D, s.pre1 + r0 * 4 * D, 4 * D);
  gemm_add_seg(p, s.emb_all + r0 * D, D, c.w->al_wih, 3 * D, D);
  p.add = s.pre1s; p.ldadd = 4 * D; p.add_mod = B;
  return gemm(kNT, p, c.st);
}

int dstep_forward(DCtx& c, int t, int b) {
  const int B = c.s.B, P = c.s.P, D = c.d.D, A = c.d.A, C = c.C, LX2 = c.LX2;
  const SetDcNetParams& w = *c.w;
  DWs& s = c.ws;
  cudaStream_t st = c.st;
  const size_t tb = (size_t)t * B;
  float* X2t = s.X2 + tb * LX2;
  const float* h2prev = s.h2 + tb * D;
  {
    GemmProblem p = gemm_problem(b, 4 * D, s.scratch4d, 4 * D);        // attention_lstm, dcnet.py:340
    if (t > 0) {
      gemm_add_seg(p, h2prev, D, w.al_wih + 2 * D, 3 * D, D);
      gemm_add_seg(p, s.X2 + (tb - B) * LX2, LX2, w.al_whh, D, D);
    }
    p.add = s.pre1 + tb * 4 * D; p.ldadd = 4 * D;
    SET_PROPAGATE(gemm(kNT, p, st));
    SET_PROPAGATE(lstm_fwd(s.scratch4d, 4 * D, s.c1 + tb * D, nullptr, s.gates1 + tb * 4 * D, s.c1 + (tb + B) * D,
                           X2t, LX2, b, D, nullptr, 0, nullptr, nullptr, 0, st));
  }
  {
    GemmProblem p[2];
    p[0] = gemm_problem(b, A, s.s2 + tb * A, A);                       // cap_decoder_att(h1), dcnet.py:262
    gemm_add_seg(p[0], X2t, LX2, w.ca_dec_w, D, D); p[0].bias = w.ca_dec_b;
    p[1] = gemm_problem(b, 4 * D, s.g2 + tb * 4 * D, 4 * D);           // language_lstm: W_ih[:, 0:D] h1 + W_hh h2
    gemm_add_seg(p[1], X2t, LX2, w.ll_wih, 2 * D, D);
    if (t > 0) gemm_add_seg(p[1], h2prev, D, w.ll_whh, D, D);
    p[1].bias = w.ll_bih; p[1].bias2 = w.ll_bhh;
    SET_PROPAGATE(gemm_group(kNT, p, 2, st));
  }
  {
    AttnFwdArgs a;
    memset(&a, 0, sizeof(a));
    a.b = b; a.P = P; a.R = 0; a.D = 2 * C; a.A = A; a.F = 4;
    a.att1c = s.att1c; a.s2 = s.s2 + tb * A; a.ld_s2 = A; a.cap_w = w.ca_full_w; a.cap_b = w.ca_full_b;
    a.mask = s.mask; a.prev_h = s.enc_out; a.prev_m = nullptr;
    a.alpha_c = s.alpha_c + tb * P; a.ctx = X2t + D; a.ld_ctx = LX2;
    SET_PROPAGATE(attention_fwd(a, st));
  }
  {
    GemmProblem p = gemm_problem(b, 4 * D, s.g2 + tb * 4 * D, 4 * D);  // W_ih[:, D:] ctx, dcnet.py:346
    gemm_add_seg(p, X2t + D, LX2, w.ll_wih + D, 2 * D, 2 * C);
    p.beta = 1;
    SET_PROPAGATE(gemm(kNT, p, st));
    SET_PROPAGATE(lstm_fwd(s.g2 + tb * 4 * D, 4 * D, s.c2 + tb * D, nullptr, s.g2 + tb * 4 * D, s.c2 + (tb + B) * D,
                           s.h2 + (tb + B) * D, D, b, D, nullptr, 0, nullptr, nullptr, 0, st));
    SET_PROPAGATE(dropout_fwd(s.h2 + (tb + B) * D, s.h2drop + tb * D, b, D, c.s.train, c.seed, kSiteFc, (long)tb * D, st));
  }
  return SET_OK;
}

struct DLg { const float* p; long ld; int inner; long ld_inner; const int* row_len; };

int dbackward(DCtx& c, const SetDcNetParams& g, const int64_t* tok, long tok_ld, long tok_os, const int64_t* prev,
              const int64_t* prev_len, const int* bt, DLg dl, const int* dec_len_dev) {
  const int B = c.s.B, P = c.s.P, T = c.s.T, D = c.d.D, A = c.d.A, V = c.d.V, C = c.C, LX2 = c.LX2;
  const SetDcNetParams& w = *c.w;
  DWs& s = c.ws;
  cudaStream_t st = c.st;
  const int TB = T * B;
  SET_CHECK_CUDA(cudaMemsetAsync(reinterpret_cast<char*>(s.emb_prev) + s.regionB_begin, 0, s.total - s.regionB_begin, st));
  {
    GemmProblem p = gemm_problem(TB, D, s.dh2raw, D);
    gemm_add_seg(p, dl.p, dl.ld, w.fc_w, D, V);
    p.a_inner = dl.inner; p.a_ld_inner = dl.ld_inner; p.a_row_len = dl.row_len; p.a_valid_inner = B;
    SET_PROPAGATE(gemm(kNN, p, st));
    // through the fc dropout (same keep bits as the forward)
    SET_PROPAGATE(dropout_fwd(s.dh2raw, s.dh2raw, TB, D, c.s.train, c.seed, kSiteFc, 0, st));
  }
  for (int t = T - 1; t >= 0; --t) {
    const int b = bt[t];
    if (b <= 0) continue;
    const size_t tb = (size_t)t * B;
    float* dG1t = s.dG1 + tb * 4 * D;
    float* dG2t = s.dG2 + tb * 4 * D;
    SET_PROPAGATE(lstm_bwd(s.g2 + tb * 4 * D, s.c2 + tb * D, s.c2 + (tb + B) * D, s.dh2raw + tb * D, D, s.dh2c, s.dc2c,
                           dG2t, b, D, st));
    {
      GemmProblem p = gemm_problem(b, LX2, s.dX2, LX2);                // d[h1 | ctx]
      gemm_add_seg(p, dG2t, 4 * D, w.ll_wih, 2 * D, 4 * D);
      SET_PROPAGATE(gemm(kNN, p, st));
    }
    {
      AttnBwdArgs a;
      memset(&a, 0, sizeof(a));
      a.b = b; a.P = P; a.R = 0; a.D = 2 * C; a.A = A; a.F = 4;
      a.att1c = s.att1c; a.s2 = s.s2 + tb * A; a.ld_s2 = A; a.cap_w = w.ca_full_w; a.mask = s.mask;
      a.prev_h = s.enc_out; a.alpha_c = s.alpha_c + tb * P;
      a.dctx = s.dX2 + D; a.ld_dctx = LX2; a.dprev_h = s.denc_out; a.datt1c = s.datt1c;
      a.ds2 = s.dS2 + tb * A; a.ld_ds2 = A; a.dcap_w = g.ca_full_w; a.dcap_b = g.ca_full_b;
      SET_PROPAGATE(attention_bwd(a, st));
    }
    {
      GemmProblem p = gemm_problem(b, D, s.dX2, LX2);                  // dh1 += d att2 @ cap_decoder_att
      gemm_add_seg(p, s.dS2 + tb * A, A, w.ca_dec_w, D, A);
      p.beta = 1;
      SET_PROPAGATE(gemm(kNN, p, st));
    }
    SET_PROPAGATE(lstm_bwd(s.gates1 + tb * 4 * D, s.c1 + tb * D, s.c1 + (tb + B) * D, s.dX2, LX2, s.dh1c, s.dc1c, dG1t,
                           b, D, st));
    if (t > 0) {
      GemmProblem p[2];
      p[0] = gemm_problem(b, D, s.dh1c, D); gemm_add_seg(p[0], dG1t, 4 * D, w.al_whh, D, 4 * D);
      p[1] = gemm_problem(b, D, s.dh2c, D);
      gemm_add_seg(p[1], dG1t, 4 * D, w.al_wih + 2 * D, 3 * D, 4 * D);
      gemm_add_seg(p[1], dG2t, 4 * D, w.ll_whh, D, 4 * D);
      SET_PROPAGATE(gemm_group(kNN, p, 2, st));
    }
  }
  auto TN = [&](float* Cm, long ldc, int M, int N, const float* dY, long ldy, const float* X, long ldx, int K) {
    GemmProblem p = gemm_problem(M, N, Cm, ldc);
    gemm_add_seg(p, dY, ldy, X, ldx, K);
    p.beta = 1;
    return p;
  };
  SET_PROPAGATE(sum_time(s.dG1, s.sumG1, T, (long)B * 4 * D, st));
  {
    GemmProblem p[2];
    p[0] = gemm_problem(B, 2 * C, s.dfh, 2 * C);
    gemm_add_seg(p[0], s.sumG1, 4 * D, w.al_wih + D, 3 * D, 4 * D);
    p[1] = gemm_problem(TB, D, s.demb_all, D);
    gemm_add_seg(p[1], s.dG1, 4 * D, w.al_wih, 3 * D, 4 * D);
    SET_PROPAGATE(gemm_group(kNN, p, 2, st));
  }
  SET_PROPAGATE(embed_bwd(tok, tok_ld, tok_os, s.emb_all, s.demb_all, g.embed, V, T, B, D, c.s.train, dec_len_dev, st));
  {
    GemmProblem p[8];
    int n = 0;
    if (T > 1) p[n++] = TN(g.al_whh, D, 4 * D, D, s.dG1 + (size_t)B * 4 * D, 4 * D, s.X2, LX2, (T - 1) * B);
    p[n++] = TN(g.al_wih, 3 * D, 4 * D, D, s.dG1, 4 * D, s.emb_all, D, TB);
    p[n++] = TN(g.al_wih + D, 3 * D, 4 * D, 2 * C, s.sumG1, 4 * D, s.fh, 2 * C, B);
    p[n++] = TN(g.al_wih + 2 * D, 3 * D, 4 * D, D, s.dG1, 4 * D, s.h2, D, TB);
    p[n++] = TN(g.ll_wih, 2 * D, 4 * D, LX2, s.dG2, 4 * D, s.X2, LX2, TB);
    p[n++] = TN(g.ll_whh, D, 4 * D, D, s.dG2, 4 * D, s.h2, D, TB);
    p[n++] = TN(g.ca_dec_w, D, A, D, s.dS2, A, s.X2, LX2, TB);
    p[n++] = TN(g.ca_feat_w, 2 * C, A, 2 * C, s.datt1c, A, s.enc_out, 2 * C, B * P);
    SET_PROPAGATE(gemm_group(kTN, p, n, st));
    GemmProblem q[2];
    for (int k = 0; k < 2; ++k) {
      q[k] = k == 0 ? gemm_problem(V, D, g.fc_w, D) : gemm_problem(V, 1, g.fc_b, 1);
      gemm_add_seg(q[k], dl.p, dl.ld, k == 0 ? s.h2drop : s.ones, k == 0 ? D : 1, TB);
      q[k].beta = 1;
      q[k].a_inner = dl.inner; q[k].a_ld_inner = dl.ld_inner; q[k].a_row_len = dl.row_len; q[k].a_valid_inner = B;
    }
    SET_PROPAGATE(gemm_group(kTN, q, 2, st));
  }
  SET_PROPAGATE(colsum(s.dG1, 4 * D, TB, 4 * D, g.al_bih, 1, st));
  SET_PROPAGATE(colsum(s.dG1, 4 * D, TB, 4 * D, g.al_bhh, 1, st));
  SET_PROPAGATE(colsum(s.dG2, 4 * D, TB, 4 * D, g.ll_bih, 1, st));
  SET_PROPAGATE(colsum(s.dG2, 4 * D, TB, 4 * D, g.ll_bhh, 1, st));
  SET_PROPAGATE(colsum(s.dS2, A, TB, A, g.ca_dec_b, 1, st));
  SET_PROPAGATE(colsum(s.datt1c, A, B * P, A, g.ca_feat_b, 1, st));
  {
    GemmProblem p = gemm_problem(B * P, 2 * C, s.denc_out, 2 * C);     // d enc_out also through cap_features_att
    gemm_add_seg(p, s.datt1c, A, w.ca_feat_w, 2 * C, A);
    p.beta = 1;
    SET_PROPAGATE(gemm(kNN, p, st));
  }
  // ---- encoder: final_hidden = tanh(concat(hcat))
  SET_PROPAGATE(tanh_bwd_inplace(s.dfh, s.fh, (long)B * 2 * C, st));
  {
    GemmProblem p[2];
    p[0] = TN(g.enc_cat_w, 2 * C, 2 * C, C, s.dfh, 2 * C, s.enc_h_f + (size_t)P * B * C, C, B);
    p[1] = TN(g.enc_cat_w + C, 2 * C, 2 * C, C, s.dfh, 2 * C, s.enc_h_r + (size_t)P * B * C, C, B);
    SET_PROPAGATE(gemm_group(kTN, p, 2, st));
    SET_PROPAGATE(colsum(s.dfh, 2 * C, B, 2 * C, g.enc_cat_b, 1, st));
    GemmProblem px = gemm_problem(B, 2 * C, s.dhcat, 2 * C);
    gemm_add_seg(px, s.dfh, 2 * C, w.enc_cat_w, 2 * C, 2 * C);
    SET_PROPAGATE(gemm(kNN, px, st));
  }
  for (int dir = 0; dir < 2; ++dir) {
    const float* hs = dir ? s.enc_h_r : s.enc_h_f;
    const float* cs = dir ? s.enc_c_r : s.enc_c_f;
    const float* gs = dir ? s.gates_r : s.gates_f;
    float* dgs = dir ? s.dgates_r : s.dgates_f;
    float* dxg = dir ? s.dxg_r : s.dxg_f;
    const float* whh = dir ? w.enc_whh_r : w.enc_whh_f;
    SET_CHECK_CUDA(cudaMemsetAsync(s.dh_run, 0, sizeof(float) * B * C, st));
    SET_CHECK_CUDA(cudaMemsetAsync(s.dc_run, 0, sizeof(float) * B * C, st));
    for (int t = P - 1; t >= 0; --t) {
      const size_t tb = (size_t)t * B;
      SET_PROPAGATE(bilstm_bwd(gs + tb * 4 * C, cs + tb * C, cs + (tb + B) * C, s.dh_run, s.dc_run, s.denc_out + dir * C,
                               (long)P * 2 * C, 2 * C, s.dhcat + dir * C, 2 * C, prev_len, t, dir, dgs + tb * 4 * C, dxg,
                               B, P, C, st));
      if (t > 0) {
        GemmProblem p = gemm_problem(B, C, s.dh_run, C);
        gemm_add_seg(p, dgs + tb * 4 * C, 4 * C, whh, C, 4 * C);
        SET_PROPAGATE(gemm(kNN, p, st));
      }
    }
    GemmProblem p[2];
    p[0] = TN(dir ? g.enc_wih_r : g.enc_wih_f, D, 4 * C, D, dxg, 4 * C, s.emb_prev, D, B * P);
    p[1] = TN(dir ? g.enc_whh_r : g.enc_whh_f, C, 4 * C, C, dgs, 4 * C, hs, C, P * B);
    SET_PROPAGATE(gemm_group(kTN, p, 2, st));
    SET_PROPAGATE(colsum(dxg, 4 * C, B * P, 4 * C, dir ? g.enc_bih_r : g.enc_bih_f, 1, st));
    SET_PROPAGATE(colsum(dxg, 4 * C, B * P, 4 * C, dir ? g.enc_bhh_r : g.enc_bhh_f, 1, st));
    GemmProblem px = gemm_problem(B * P, D, s.demb_prev, D);
    gemm_add_seg(px, dxg, 4 * C, dir ? w.enc_wih_r : w.enc_wih_f, D, 4 * C);
    px.beta = dir;   // second direction accumulates
    SET_PROPAGATE(gemm(kNN, px, st));
  }
  SET_PROPAGATE(embed_bwd(prev, 1, c.s.Wp, s.emb_prev, s.demb_prev, g.embed, V, B, P, D, c.s.train, nullptr, st));
  return SET_OK;
}

}  // namespace
}  // namespace set

using namespace set;

extern "C" {

size_t set_dcnet_workspace_bytes(const SetDims* dims, const SetSeqShape* shape) {
  if (dcheck(dims, shape) != SET_OK) return 0;
  Arena ar;
  DWs ws;
  dlayout(*dims, *shape, ar, ws);
  return ws.total;
}

int set_dcnet_workspace_lookup(const SetDims* dims, const SetSeqShape* shape, const char* name, size_t* offset,
                               size_t* bytes) {
  SET_PROPAGATE(dcheck(dims, shape));
  Arena ar;
  DWs ws;
  dlayout(*dims, *shape, ar, ws);
  for (const auto& e : ar.entries)
    if (e.name == name) { *offset = e.off; *bytes = e.bytes; return SET_OK; }
  set_record_error("unknown workspace buffer name");
  return SET_ERR_ARG;
}

int set_dcnet_xe_forward(const SetDims* dims, const SetSeqShape* shape, const SetDcNetParams* w, const int64_t* caps,
                         const int* decode_len_host, const int64_t* prev, const int64_t* prev_len, uint64_t seed,
                         float* predictions, void* workspace, size_t workspace_bytes, void* stream) {
  DCtx c;
  SET_PROPAGATE(make_dctx(c, dims, shape, w, workspace, workspace_bytes, seed, stream));
  SET_REQUIRE(caps && decode_len_host && prev && prev_len && predictions, "null input");
  SET_REQUIRE(shape->Wc > shape->T, "caption width must exceed T");
  std::vector<int> bt;
  SET_PROPAGATE(batch_sizes(*shape, decode_len_host, bt));
  const int B = shape->B, T = shape->T, D = dims->D, V = dims->V;
  SET_PROPAGATE(dprepare(c, prev, prev_len));
  SET_CHECK_CUDA(cudaMemcpyAsync(c.ws.dec_len, decode_len_host, sizeof(int) * B, cudaMemcpyHostToDevice, c.st));
  SET_PROPAGATE(embed_fwd(caps, shape->Wc, 1, w->embed, V, c.ws.emb_all, T, B, D, shape->train, seed, kSiteEmb, 0, B, 1,
                          c.st));
  SET_PROPAGATE(dproject_words(c, 0, T));
  for (int t = 0; t < T; ++t) SET_PROPAGATE(dstep_forward(c, t, bt[t]));
  SET_CHECK_CUDA(cudaMemsetAsync(predictions, 0, sizeof(float) * (size_t)B * T * V, c.st));
  GemmProblem p = gemm_problem(T * B, V, predictions, V);
  gemm_add_seg(p, c.ws.h2drop, D, w->fc_w, D, D);
  p.bias = w->fc_b;
  p.c_inner = B; p.c_ld_inner = (long)T * V; p.c_row_len = c.ws.dec_len; p.c_valid_inner = B;
  SET_PROPAGATE(gemm(kNT, p, c.st));
  return SET_OK;
}

int set_dcnet_xe_backward(const SetDims* dims, const SetSeqShape* shape, const SetDcNetParams* w,
                          const SetDcNetParams* grads, const int64_t* caps, const int* decode_len_host,
                          const int64_t* prev, const int64_t* prev_len, uint64_t seed, const float* d_predictions,
                          void* workspace, size_t workspace_bytes, void* stream) {
  DCtx c;
  SET_PROPAGATE(make_dctx(c, dims, shape, w, workspace, workspace_bytes, seed, stream));
  SET_REQUIRE(grads && caps && decode_len_host && prev && prev_len && d_predictions, "null input");
  std::vector<int> bt;
  SET_PROPAGATE(batch_sizes(*shape, decode_len_host, bt));
  DLg dl;
  dl.p = d_predictions; dl.ld = dims->V; dl.inner = shape->B; dl.ld_inner = (long)shape->T * dims->V;
  dl.row_len = c.ws.dec_len;
  return dbackward(c, *grads, caps, shape->Wc, 1, prev, prev_len, bt.data(), dl, c.ws.dec_len);
}

int set_dcnet_step_begin(const SetDims* dims, const SetSeqShape* shape, const SetDcNetParams* w, const int64_t* prev,
                         const int64_t* prev_len, void* workspace, size_t workspace_bytes, void* stream) {
  DCtx c;
  SET_PROPAGATE(make_dctx(c, dims, shape, w, workspace, workspace_bytes, 0, stream));
  SET_REQUIRE(shape->T == 2 && shape->train == 0, "step sessions use T == 2, eval mode");
  SET_REQUIRE(prev && prev_len, "null input");
  return dprepare(c, prev, prev_len);
}

int set_dcnet_step(const SetDims* dims, const SetSeqShape* shape, const SetDcNetParams* w, const int64_t* tokens,
                   int rows, float* h1, float* c1, float* h2, float* c2, float* scores, void* workspace,
                   size_t workspace_bytes, void* stream) {
  DCtx c;
  SET_PROPAGATE(make_dctx(c, dims, shape, w, workspace, workspace_bytes, 0, stream));
  SET_REQUIRE(shape->T == 2 && shape->train == 0, "step sessions use T == 2, eval mode");
  SET_REQUIRE(tokens && h1 && c1 && h2 && c2 && scores && rows >= 1 && rows <= shape->B, "bad args");
  const size_t B = shape->B, D = dims->D, V = dims->V, LX2 = c.LX2;
  DWs& s = c.ws;
  cudaStream_t st = c.st;
  const size_t row = sizeof(float) * D;
  // the step runs as t = 1: the "previous" state lives where step 0 would have left it
  SET_CHECK_CUDA(cudaMemcpy2DAsync(s.X2, sizeof(float) * LX2, h1, row, row, rows, cudaMemcpyDeviceToDevice, st));
  SET_CHECK_CUDA(cudaMemcpyAsync(s.c1 + B * D, c1, row * rows, cudaMemcpyDeviceToDevice, st));
  SET_CHECK_CUDA(cudaMemcpyAsync(s.h2 + B * D, h2, row * rows, cudaMemcpyDeviceToDevice, st));
  SET_CHECK_CUDA(cudaMemcpyAsync(s.c2 + B * D, c2, row * rows, cudaMemcpyDeviceToDevice, st));
  SET_PROPAGATE(embed_fwd(tokens, 1, 0, w->embed, dims->V, s.emb_all + B * D, 1, rows, D, 0, 0, kSiteEmb, 0, 0, 1, st));
  SET_PROPAGATE(dproject_words(c, 1, 1));
  SET_PROPAGATE(dstep_forward(c, 1, rows));
  GemmProblem p = gemm_problem(rows, V, scores, V);          // fc(h2), eval: dropout is identity (eval_full.py:149)
  gemm_add_seg(p, s.h2drop + B * D, D, w->fc_w, D, D);
  p.bias = w->fc_b;
  SET_PROPAGATE(gemm(kNT, p, st));
  SET_CHECK_CUDA(cudaMemcpy2DAsync(h1, row, s.X2 + B * LX2, sizeof(float) * LX2, row, rows, cudaMemcpyDeviceToDevice, st));
  SET_CHECK_CUDA(cudaMemcpyAsync(c1, s.c1 + 2 * B * D, row * rows, cudaMemcpyDeviceToDevice, st));
  SET_CHECK_CUDA(cudaMemcpyAsync(h2, s.h2 + 2 * B * D, row * rows, cudaMemcpyDeviceToDevice, st));
  SET_CHECK_CUDA(cudaMemcpyAsync(c2, s.c2 + 2 * B * D, row * rows, cudaMemcpyDeviceToDevice, st));
  return SET_OK;
}

int set_dcnet_rollout(const SetDims* dims, const SetSeqShape* shape, const SetDcNetParams* w, const int64_t* prev,
                      const int64_t* prev_len, int64_t start_token, int64_t end_token, int mode,
                      const int64_t* forced, uint64_t seed, int64_t* seq, float* seq_logprobs, void* workspace,
                      size_t workspace_bytes, void* stream) {
  DCtx c;
  SET_PROPAGATE(make_dctx(c, dims, shape, w, workspace, workspace_bytes, seed, stream));
  SET_REQUIRE(prev && prev_len && seq && seq_logprobs, "null input");
  SET_REQUIRE(mode >= 0 && mode <= 2 && (mode != 2 || forced != nullptr), "bad mode");
  const int B = shape->B, T = shape->T, D = dims->D, V = dims->V;
  DWs& s = c.ws;
  SET_PROPAGATE(dprepare(c, prev, prev_len));
  SET_CHECK_CUDA(cudaMemsetAsync(seq, 0, sizeof(int64_t) * (size_t)B * T, c.st));
  SET_CHECK_CUDA(cudaMemsetAsync(seq_logprobs, 0, sizeof(float) * (size_t)B * T, c.st));
  fill_i64_kernel<<<1, 256, 0, c.st>>>(s.it, B, start_token);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  for (int t = 0; t < T; ++t) {
    SET_PROPAGATE(embed_fwd(s.it + (size_t)t * B, 1, 0, w->embed, V, s.emb_all + (size_t)t * B * D, 1, B, D,
                            shape->train, seed, kSiteEmb, (long)t * B, 0, 1, c.st));
    SET_PROPAGATE(dproject_words(c, t, 1));
    SET_PROPAGATE(dstep_forward(c, t, B));
    float* lg = shape->train ? s.logits + (size_t)t * B * V : s.logits;
    GemmProblem p = gemm_problem(B, V, lg, V);
    gemm_add_seg(p, s.h2drop + (size_t)t * B * D, D, w->fc_w, D, D);
    p.bias = w->fc_b;
    SET_PROPAGATE(gemm(kNT, p, c.st));
    sample_step_kernel<<<B, 256, 0, c.st>>>(lg, V, B, T, t, mode, forced, seed, end_token, s.unfinished, s.unf_count,
                                            s.it + (size_t)(t + 1) * B, seq, seq_logprobs, s.lse, s.tok_raw);
    SET_CHECK_CUDA(cudaGetLastError());
    set_count_launch(1);
  }
  return SET_OK;
}

int set_dcnet_rollout_backward(const SetDims* dims, const SetSeqShape* shape, const SetDcNetParams* w,
                               const SetDcNetParams* grads, const int64_t* prev, const int64_t* prev_len,
                               uint64_t seed, const float* d_seq_logprobs, void* workspace, size_t workspace_bytes,
                               void* stream) {
  DCtx c;
  SET_PROPAGATE(make_dctx(c, dims, shape, w, workspace, workspace_bytes, seed, stream));
  SET_REQUIRE(shape->train, "rollout backward needs a train-mode forward (activations kept)");
  SET_REQUIRE(grads && prev && prev_len && d_seq_logprobs, "null input");
  const int B = shape->B, T = shape->T, V = dims->V;
  DWs& s = c.ws;
  rollout_dlogits_kernel<<<T * B, 256, 0, c.st>>>(s.logits, V, B, T, d_seq_logprobs, s.lse, s.tok_raw);
  SET_CHECK_CUDA(cudaGetLastError());
  set_count_launch(1);
  std::vector<int> bt(T, B);
  DLg dl;
  dl.p = s.logits; dl.ld = V; dl.inner = 0; dl.ld_inner = 0; dl.row_len = nullptr;
  return dbackward(c, *grads, s.it, 1, B, prev, prev_len, bt.data(), dl, nullptr);
}

}  // extern "C"
