/*
 * set_b200.h -- C ABI of the B200-native EditNet/DCNet decode path.
 *
 * The reference (fawazsammani/show-edit-tell) has no FFI/plugin layer: its hot path sits
 * behind `nn.Module.forward` methods.  This header is the boundary a
 * binding would target instead: stateless functions over caller-owned, contiguous device
 * buffers (fp32 data, int64 tokens/lengths exactly as the reference's tensors hold them),
 * explicit sizes, an explicit CUDA stream, caller-provided workspace.  No torch types.
 * Every entry point names the reference interface it replaces.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the name ends in `_host`;
 *   - matrices are row-major [out_features, in_features] exactly like `nn.Linear.weight`;
 *   - E = D = C (emb_dim = decoder_dim = caption_features_dim), which the reference's own
 *     concatenations force (editnet.py:359-361,468-469); A = attention_dim, F = image
 *     feature dim, V = vocabulary size.  D, A, F must be multiples of 4;
 *   - rows of a batch are in the order the reference computes in: sorted by caption
 *     length, descending, for the teacher-forced path (editnet.py:488-492); caller order
 *     for the rollout path (editnet_rl.py:485-549);
 *   - return value 0 = success; otherwise set_last_error() describes the failure.  The
 *     library never falls back to a CPU path.
 */
#ifndef SET_B200_H_
#define SET_B200_H_

#include <stddef.h>
#include <stdint.h>

#if defined(__GNUC__)
#define SET_API __attribute__((visibility("default")))
#else
#define SET_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef struct SetDims {
  int V; /* vocabulary (len(word_map)) */
  int D; /* emb_dim = decoder_dim = caption_features_dim (1024) */
  int A; /* attention_dim (512) */
  int F; /* image_features_dim (2048) */
} SetDims;

/* EditNet parameters, one pointer per `state_dict` entry of DecoderC
 * (editnet.py:451-471; key names in SURVEY.md Appendix B).  The same struct, pointing
 * at gradient buffers, receives d(loss)/d(param) from the backward entry points
 * (accumulated: the caller zeroes them). */
typedef struct SetEditNetParams {
  float* embed;                          /* embed.embedding.weight                  [V,D]      */
  float *enc_x2h_w, *enc_x2h_b;          /* caption_encoder.lstm_encoder_cell.x2h   [4D,D],[4D]*/
  float *enc_h2h_w, *enc_h2h_b;          /* caption_encoder.lstm_encoder_cell.h2h   [4D,D],[4D]*/
  float *enc_aff_w, *enc_aff_b;          /* caption_encoder.affine_hn               [D,D],[D]  */
  float *ca_feat_w, *ca_feat_b;          /* caption_attention.cap_features_att      [A,D],[A]  */
  float *ca_dec_w, *ca_dec_b;            /* caption_attention.cap_decoder_att       [A,D],[A]  */
  float *ca_full_w, *ca_full_b;          /* caption_attention.cap_full_att          [1,A],[1]  */
  float *ca_gate_w, *ca_gate_b;          /* caption_attention.context_gate          [D,3D],[D] */
  float *ca_sc_w, *ca_sc_b;              /* caption_attention.sc_affine             [D,D],[D]  */
  float *ca_tc_w, *ca_tc_b;              /* caption_attention.tc_affine             [D,2D],[D] */
  float *va_emb_w, *va_emb_b;            /* visual_attention.att_embed.0            [D,F],[D]  */
  float *va_feat_w, *va_feat_b;          /* visual_attention.features_att           [A,D],[A]  */
  float *va_dec_w, *va_dec_b;            /* visual_attention.decoder_att            [A,D],[A]  */
  float *va_full_w, *va_full_b;          /* visual_attention.full_att               [1,A],[1]  */
  float *al_wih, *al_whh, *al_bih, *al_bhh; /* attention_lstm.{weight_ih [4D,3D+F], weight_hh [4D,D], bias_*} */
  float *cl_x2h_w, *cl_x2h_b;            /* copy_lstm.x2h                           [4D,2D+F]  */
  float *cl_h2h_w, *cl_h2h_b;            /* copy_lstm.h2h                           [4D,D]     */
  float *cl_gcn_w, *cl_gcn_b;            /* copy_lstm.gate_cnew                     [D,D]      */
  float *cl_gcm_w, *cl_gcm_b;            /* copy_lstm.gate_cmem                     [D,D]      */
  float *fc_w, *fc_b;                    /* fc                                      [V,D],[V]  */
} SetEditNetParams;

/* Shape of one teacher-forced (XE) or rollout call. */
typedef struct SetSeqShape {
  int B;        /* captions in the batch                                                  */
  int R;        /* regions per image (36; <=128)                                          */
  int Wc;       /* width of the caption tensor (20)      -- XE only                       */
  int Wp;       /* width of the previous-caption tensor (18)                              */
  int P;        /* max previous-caption length in the batch (encoder steps, <=Wp)         */
  int T;        /* decode steps: max(decode_lengths) for XE, max_len for rollouts         */
  int train;    /* 1: dropout active (Philox keyed by `seed`), 0: eval                    */
  int adaptive; /* 1: ragged-region variant (adaptive_features/editnet_adaptive.py:438-457):
                   all-zero region rows are padding, image_mean is an input              */
} SetSeqShape;

SET_API const char* set_last_error(void);
SET_API int set_version(void);

/* Workspace (saved activations + scratch) needed by set_editnet_xe_forward/backward and
 * by set_editnet_rollout for this shape. */
SET_API size_t set_editnet_workspace_bytes(const SetDims* dims, const SetSeqShape* shape);

/* Previous-caption encoder alone: replaces CaptionEncoderC.forward, editnet.py:319-348 (rows in caller
 * order; the reference's internal sort/unsort is unobservable).  Outputs (any may be NULL):
 * hidden_states / memory_states [B,P,D] (zero beyond each row's length), final_hidden [B,D], mask [B,P].
 * shape: B, Wp, P, train are used (R = T = 1 is fine). */
SET_API int set_editnet_encode(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w,
                       const int64_t* seq, const int64_t* seq_len, uint64_t seed, float* hidden_states,
                       float* memory_states, float* final_hidden, float* mask, void* workspace,
                       size_t workspace_bytes, void* stream);

/* One decode step on explicit state -- what beam search needs (evaluate(), editnet.py:639-653 calls the
 * cells one by one on k <= beam_size rows; here the whole step is one call).  set_editnet_step_begin runs
 * the per-image/per-caption precomputation once (encoder, attention projections) for B rows into the
 * workspace; set_editnet_step then advances the first `rows` rows: embeds `tokens`, updates
 * (h1,c1,h2,c2) [rows,D] IN PLACE and writes the vocabulary scores fc(h2) [rows,V].  Rows keep their
 * identity across steps (the reference's beams share one image, so re-ordering beams only permutes the
 * caller's state tensors).  shape->T must be 2 and shape->train 0. */
SET_API int set_editnet_step_begin(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w,
                           const float* feats, const float* image_mean, const int64_t* prev,
                           const int64_t* prev_len, void* workspace, size_t workspace_bytes, void* stream);
SET_API int set_editnet_step(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w,
                     const float* feats, const int64_t* tokens, int rows, float* h1, float* c1, float* h2,
                     float* c2, float* scores, void* workspace, size_t workspace_bytes, void* stream);

/* Teacher-forced forward: replaces DecoderC.forward, editnet.py:479-548 (use_ss=False)
 * and, with shape->adaptive, adaptive_features/editnet_adaptive.py:489-562.
 *   feats            [B,R,F]   image_features[sort_ind]
 *   image_mean       [B,F]     adaptive only (else NULL: computed as feats.mean(1), editnet.py:503)
 *   caps             [B,Wc]    encoded_captions[sort_ind] (int64)
 *   decode_len_host  [B]       HOST ints, caption_lengths-1, non-increasing (editnet.py:497)
 *   prev, prev_len   [B,Wp],[B] encoded_previous_captions / previous_cap_length, sorted rows (int64)
 *   predictions      [B,T,V]   written in full: rows t >= decode_len[i] are zero (editnet.py:499,546);
 *                              NULL (train mode only): the logits stay time-major [T,B,V] inside the
 *                              workspace for set_editnet_xe_loss_time_major / backward(d_predictions=NULL)
 * The workspace keeps what set_editnet_xe_backward needs; it must stay untouched between
 * the two calls. */
SET_API int set_editnet_xe_forward(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w,
                           const float* feats, const float* image_mean, const int64_t* caps,
                           const int* decode_len_host, const int64_t* prev, const int64_t* prev_len,
                           uint64_t seed, float* predictions, void* workspace, size_t workspace_bytes,
                           void* stream);

/* The same with scheduled sampling (use_ss=True, editnet.py:508-520; train mode only): at steps t >= 1 a
 * row's input token is, with probability ss_prob, drawn from multinomial(exp(scores of step t-1)) instead of
 * the ground truth (Philox keyed by `seed`; torch's RNG stream cannot be reproduced).  fed_tokens [B,Wc]
 * receives the tokens actually fed (pass it as `caps` to set_editnet_xe_backward; the loss still uses the
 * true captions).  ss_replay [B,Wc] (optional) forces the fed tokens -- used to validate against the oracle. */
SET_API int set_editnet_xe_forward_ss(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w,
                              const float* feats, const float* image_mean, const int64_t* caps,
                              const int* decode_len_host, const int64_t* prev, const int64_t* prev_len,
                              uint64_t seed, float ss_prob, const int64_t* ss_replay, int64_t* fed_tokens,
                              float* predictions, void* workspace, size_t workspace_bytes, void* stream);

/* Backward of the above for an upstream gradient d_predictions [B,T,V] (entries at
 * t >= decode_len[i] are ignored, as the reference's slice-assignment does).  Replaces the
 * autograd replay that `loss.backward()` (editnet.py:579) performs through DecoderC.forward.
 * d_predictions == NULL takes d(logits) from the workspace (written there by
 * set_editnet_xe_loss_time_major).  Gradients are ACCUMULATED into *grads. */
SET_API int set_editnet_xe_backward(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w,
                            const SetEditNetParams* grads, const float* feats, const int64_t* caps,
                            const int* decode_len_host, const int64_t* prev, const int64_t* prev_len,
                            uint64_t seed, const float* d_predictions, void* workspace,
                            size_t workspace_bytes, void* stream);

/* Packed cross-entropy over the decoded positions: replaces pack_padded_sequence x2 +
 * CrossEntropyLoss, editnet.py:571-577.  Writes the mean loss to loss_out[0],
 * sum(decode_len) to loss_out[1], and (if d_predictions != NULL) d(mean loss)/d(predictions)
 * -- zero at undecoded positions; may alias `predictions`.  `inv_count` <= 0 means
 * 1/sum(decode_len); data-parallel callers pass 1/global_count instead. */
SET_API int set_xe_loss(int B, int T, int V, int Wc, const float* predictions, const int64_t* caps,
                const int* decode_len_dev, float inv_count, float* loss_out, float* d_predictions,
                void* stream);

/* Same loss on the time-major logits a set_editnet_xe_forward(predictions=NULL) call left in the
 * workspace; overwrites them with d(loss)/d(logits).  loss_out as above. */
SET_API int set_editnet_xe_loss_time_major(const SetDims* dims, const SetSeqShape* shape, const int64_t* caps,
                                   float inv_count, float* loss_out, void* workspace, size_t workspace_bytes,
                                   void* stream);

/* Autoregressive rollout: replaces DecoderC.forward of editnet_rl.py:485-549.
 *   mode 0 = greedy (sample_max, :521), 1 = multinomial sample (sample_rl, :525-527; inverse-CDF
 *   on Philox uniforms keyed by `seed`), 2 = forced (replay `forced` [B,T] tokens and return
 *   their log-probs; used to validate mode 1 and to re-materialise activations).
 *   seq [B,T] int64 and seq_logprobs [B,T] are written in full (zeros after a row finishes).
 * With shape->train the activations for set_editnet_rollout_backward are kept in the workspace. */
SET_API int set_editnet_rollout(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w,
                        const float* feats, const float* image_mean, const int64_t* prev,
                        const int64_t* prev_len, int64_t start_token, int64_t end_token, int mode,
                        const int64_t* forced, uint64_t seed, int64_t* seq, float* seq_logprobs,
                        void* workspace, size_t workspace_bytes, void* stream);

/* Backward of a rollout for upstream d(loss)/d(seq_logprobs) [B,T] (RewardCriterion,
 * editnet_rl.py:557-573, produces it).  Gradients are accumulated into *grads. */
SET_API int set_editnet_rollout_backward(const SetDims* dims, const SetSeqShape* shape, const SetEditNetParams* w,
                                 const SetEditNetParams* grads, const float* feats, const int64_t* prev,
                                 const int64_t* prev_len, uint64_t seed, const float* d_seq_logprobs,
                                 void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------
 * DCNet (text-only denoising auto-encoder): DAE of dcnet.py:273-350 / dcnet_rl.py:273-346.
 * SetDims: D = decoder_dim = emb_dim = 2 * caption_features_dim (the reference's concatenations force
 * this, dcnet.py:286-291), A = attention_dim, F unused.  SetSeqShape: R and adaptive unused. */
typedef struct SetDcNetParams {
  float* embed;                                          /* embed.embedding.weight            [V,D]    */
  float *enc_wih_f, *enc_whh_f, *enc_bih_f, *enc_bhh_f;  /* caption_encoder.lstm_encoder.*_l0 [4C,D],[4C,C] */
  float *enc_wih_r, *enc_whh_r, *enc_bih_r, *enc_bhh_r;  /* caption_encoder.lstm_encoder.*_l0_reverse   */
  float *enc_cat_w, *enc_cat_b;                          /* caption_encoder.concat            [2C,2C]  */
  float *ca_feat_w, *ca_feat_b;                          /* caption_attention.cap_features_att [A,2C]  */
  float *ca_dec_w, *ca_dec_b;                            /* caption_attention.cap_decoder_att  [A,D]   */
  float *ca_full_w, *ca_full_b;                          /* caption_attention.cap_full_att     [1,A]   */
  float *al_wih, *al_whh, *al_bih, *al_bhh;              /* attention_lstm [4D,3D],[4D,D]              */
  float *ll_wih, *ll_whh, *ll_bih, *ll_bhh;              /* language_lstm  [4D,2D],[4D,D]              */
  float *fc_w, *fc_b;                                    /* fc [V,D]                                   */
} SetDcNetParams;

SET_API size_t set_dcnet_workspace_bytes(const SetDims* dims, const SetSeqShape* shape);
SET_API int set_dcnet_workspace_lookup(const SetDims* dims, const SetSeqShape* shape, const char* name,
                               size_t* offset, size_t* bytes);
/* Teacher-forced forward / backward: replace DAE.forward (dcnet.py:303-350) and its autograd replay.
 * Arguments as for the EditNet entry points, without image inputs. */
SET_API int set_dcnet_xe_forward(const SetDims* dims, const SetSeqShape* shape, const SetDcNetParams* w,
                         const int64_t* caps, const int* decode_len_host, const int64_t* prev,
                         const int64_t* prev_len, uint64_t seed, float* predictions, void* workspace,
                         size_t workspace_bytes, void* stream);
SET_API int set_dcnet_xe_backward(const SetDims* dims, const SetSeqShape* shape, const SetDcNetParams* w,
                          const SetDcNetParams* grads, const int64_t* caps, const int* decode_len_host,
                          const int64_t* prev, const int64_t* prev_len, uint64_t seed,
                          const float* d_predictions, void* workspace, size_t workspace_bytes, void* stream);
/* One DCNet decode step on explicit state: the DCNet half of the ensemble beam search, eval/eval xe/eval_full.py:141-149
 * (embed, attention_lstm, caption_attention, language_lstm, fc called one by one on k rows there).  Same contract as
 * set_editnet_step_begin / set_editnet_step: begin runs the bi-LSTM encoder and the hoisted projections for B rows,
 * step advances the first `rows` rows in place and writes fc(h2) [rows,V].  shape->T == 2, shape->train == 0. */
SET_API int set_dcnet_step_begin(const SetDims* dims, const SetSeqShape* shape, const SetDcNetParams* w,
                         const int64_t* prev, const int64_t* prev_len, void* workspace, size_t workspace_bytes,
                         void* stream);
SET_API int set_dcnet_step(const SetDims* dims, const SetSeqShape* shape, const SetDcNetParams* w, const int64_t* tokens,
                   int rows, float* h1, float* c1, float* h2, float* c2, float* scores, void* workspace,
                   size_t workspace_bytes, void* stream);
/* Rollout / its backward: replace DAE.forward of dcnet_rl.py:286-346 (modes as set_editnet_rollout). */
SET_API int set_dcnet_rollout(const SetDims* dims, const SetSeqShape* shape, const SetDcNetParams* w,
                      const int64_t* prev, const int64_t* prev_len, int64_t start_token, int64_t end_token,
                      int mode, const int64_t* forced, uint64_t seed, int64_t* seq, float* seq_logprobs,
                      void* workspace, size_t workspace_bytes, void* stream);
SET_API int set_dcnet_rollout_backward(const SetDims* dims, const SetSeqShape* shape, const SetDcNetParams* w,
                               const SetDcNetParams* grads, const int64_t* prev, const int64_t* prev_len,
                               uint64_t seed, const float* d_seq_logprobs, void* workspace,
                               size_t workspace_bytes, void* stream);

/* SCST loss: replaces RewardCriterion.forward, editnet_rl.py:557-573.  loss_out[0] = loss;
 * d_logprobs [B,T] (optional) = d loss / d seq_logprobs. */
SET_API int set_reward_criterion(int B, int T, const float* seq_logprobs, const int64_t* seq, const float* reward,
                         float* loss_out, float* d_logprobs, void* stream);

/* Self-critical CIDEr-D reward from token ids: replaces get_self_critical_reward() of editnet_rl.py:611-646 together with
 * preprocess_gd (:587-600), array_to_str (:602-609) and the CiderD.compute_score call (pyciderevalcap, `df='coco-train-idxs'`).
 * gen / greedy [B][L] int64 rollouts (0 = <end>/pad as the rollout writes them); all_caps [B][R][Wc] int64 reference captions
 * as the data loader delivers them (<start> .. <end> <pad>..); df_keys / df_vals: open-addressing table (capacity a power of
 * two, empty key = ~0, linear probing behind splitmix64(key)) mapping a packed n-gram -- sum_j (token_j + 1) << (16 j) --
 * to its document frequency; ref_len = number of documents.  Writes scores [2B] (CIDEr-D x 10 of the B sampled then the B
 * greedy captions) and rewards [B][L] = weight * (score_sample - score_greedy) broadcast over the steps. */
SET_API int set_ciderd_reward(int B, int L, int R, int Wc, const int64_t* gen, const int64_t* greedy,
                      const int64_t* all_caps, int64_t start_tok, int64_t end_tok, int64_t pad_tok,
                      const uint64_t* df_keys, const float* df_vals, uint64_t df_capacity, double ref_len, double sigma,
                      float cider_weight, float* scores, float* rewards, void* stream);

/* Global-norm clip + Adam over one flat parameter buffer: replaces clip_grad_norm_(0.25) +
 * Adam.step(), editnet.py:580-581.  Gradients are first multiplied by grad_scale and, when
 * count_dev != NULL, divided by *count_dev (a device float: the all-reduced token count of a
 * data-parallel step whose ranks contributed loss SUMS; a zero count leaves the
 * parameters unchanged).  `scratch` holds >= SET_CLIP_ADAM_SCRATCH_FLOATS floats: scratch[1]
 * receives the pre-clip total norm, scratch[8..] the per-block partial sums of squares, which
 * every block then adds in one fixed order -- the clip coefficient is a deterministic function
 * of the gradient bits, so data-parallel replicas stay bit-identical. */
#define SET_CLIP_ADAM_SCRATCH_FLOATS 2048
SET_API int set_clip_adam(float* params, const float* grads, float* exp_avg, float* exp_avg_sq, size_t n,
                  int step, float lr, float beta1, float beta2, float eps, float max_norm, float grad_scale,
                  const float* count_dev, float* scratch, void* stream);

/* Number of CUDA kernels this library has launched since the last reset (bench.py reports it). */
SET_API long long set_launch_count(int reset);

/* Optional timing of the two recurrent loops of the teacher-forced path (CUDA events recorded on
 * the launch stream around the T forward steps and the T reverse steps of the most recent calls).
 * set_profile_read() blocks until those events completed; a value of -1 means "not recorded". */
SET_API int set_profile_enable(int on);
SET_API int set_profile_read(float* fwd_loop_ms, float* bwd_loop_ms);

/* Test / debugging helpers. */
/* keep flags (0/1 floats) of dropout site `site` (1 enc-embed, 2 embed, 3 att_embed, 4 fc) for
 * linear element indices [base, base+n) -- the exact bits the kernels use. */
SET_API int set_dropout_keep_mask(float* out, size_t n, uint64_t seed, int site, size_t base, void* stream);
/* byte offset and byte size of a named workspace buffer for this shape (returns 0 if found). */
SET_API int set_editnet_workspace_lookup(const SetDims* dims, const SetSeqShape* shape, const char* name,
                                 size_t* offset, size_t* bytes);
/* GEMM engine selection: 0 = tcgen05 3xTF32 tensor-core kernel where a problem is eligible (16-byte
 * aligned operands, K >= 64), CUDA-core fp32 kernel otherwise; 1 = CUDA-core kernel only.  Both are
 * this library's own kernels; the switch exists for A/B parity tests and profiling. */
SET_API int set_gemm_backend(int backend);
/* debugging: device buffer (>= 16 x uint64) stamped with %globaltimer by CTA 0 of each tensor-core launch */
SET_API int set_gemm_trace(void* buf);
/* debugging: the next `launches` tensor-core launches stamp consecutive slices (stride_u64 x uint64 each,
   >= 2500) of `buf`: [0..8] phase stamps of CTA 0, [16 + 2c], [17 + 2c] entry/exit time of CTA c < 1000,
   [2100 + 8k + s] SM-clock stamps of K-block k of CTA 0 */
SET_API int set_gemm_trace_seq(void* buf, long stride_u64, int launches);
SET_API int set_gemm_stats(long long* tc_launches, long long* simt_launches, int reset);
/* ---- Sub-module call surface: the reference's beam searches call the decoder's sub-modules one by one
 * (evaluate(), editnet.py:613,645-653; evaluate_full(), eval/eval xe/eval_full.py:107-149).  Each entry point is one
 * of those forwards exactly as the reference computes it (nothing hoisted), inference only, on contiguous fp32 / int64
 * device buffers; scratch comes from the caller (size queries below, in floats). */
/* EmbeddingC.forward, editnet.py:300-304: out[n][D] = dropout(relu(table[tokens[n]])) */
SET_API int set_embed_forward(const int64_t* tokens, long n, const float* table, int V, int D, int train, uint64_t seed,
                              float* out, void* stream);
/* nn.LSTMCell.forward (attention_lstm, editnet.py:532; DCNet cells dcnet.py:338,346): x [rows][I], gates_scratch [rows][4D] */
SET_API int set_lstm_cell_forward(int rows, int I, int D, const float* x, const float* h, const float* c,
                                  const float* w_ih, const float* w_hh, const float* b_ih, const float* b_hh,
                                  float* gates_scratch, float* h_out, float* c_out, void* stream);
/* CaptionAttentionC.forward, editnet.py:364-381: prev_h [rows][P][D], h1 / emb [rows][D], mask [rows][P] ->
 * out [rows][D] (gated context), alpha [rows][P] */
SET_API size_t set_caption_attention_scratch_floats(const SetDims* dims, int rows, int P);
SET_API int set_caption_attention_forward(const SetDims* dims, int rows, int P, const SetEditNetParams* w,
                                          const float* prev_h, const float* h1, const float* emb, const float* mask,
                                          float* scratch, size_t scratch_floats, float* out, float* alpha, void* stream);
/* VisualAttentionC.forward, editnet.py:439-447 (adaptive: editnet_adaptive.py:438-457): feats [rows][R][F], h1 [rows][D]
 * -> out [rows][F]; att_embed is recomputed on every call, as the reference does */
SET_API size_t set_visual_attention_scratch_floats(const SetDims* dims, int rows, int R);
SET_API int set_visual_attention_forward(const SetDims* dims, int rows, int R, const SetEditNetParams* w,
                                         const float* feats, const float* h1, int adaptive, int train, uint64_t seed,
                                         float* scratch, size_t scratch_floats, float* out, void* stream);
/* DCNet CaptionAttention.forward, dcnet.py:254-270: enc [rows][P][D], h1 [rows][D], mask [rows][P] -> out [rows][D] */
SET_API size_t set_dcnet_caption_attention_scratch_floats(const SetDims* dims, int rows, int P);
SET_API int set_dcnet_caption_attention_forward(const SetDims* dims, int rows, int P, const SetDcNetParams* w,
                                                const float* enc, const float* h1, const float* mask, float* scratch,
                                                size_t scratch_floats, float* out, void* stream);
/* SelectC.forward, editnet.py:403-421: prev_m [rows][P][D], alpha [rows][P] -> out [rows][D] */
SET_API int set_select_forward(int rows, int P, int D, const float* prev_m, const float* alpha, float* out, void* stream);
/* CopyLSTMCellC.forward, editnet.py:265-285: x [rows][2D+F], h / c / mem [rows][D] -> h_out, c_out */
SET_API size_t set_copy_lstm_scratch_floats(const SetDims* dims, int rows);
SET_API int set_copy_lstm_forward(const SetDims* dims, int rows, const SetEditNetParams* w, const float* x,
                                  const float* h, const float* c, const float* mem, float* scratch,
                                  size_t scratch_floats, float* h_out, float* c_out, void* stream);

/* ---- Batched beam search on the device: the expansion step of evaluate(), editnet.py:654-696,
 * and of the ensemble evaluate_full(), eval/eval xe/eval_full.py:151-191, for N images x K beams at once (rows
 * image * K + beam of one step session).  Per image: log-softmax of the live beams' scores (logits_d != NULL:
 * log((softmax_e + softmax_d) / 2)), cumulative add, top-k over (live beams x V) -- beam 0 only at step 1 -- sequence
 * extension, <end> bookkeeping (completed beams move to the completion store, k_live shrinks), next input tokens and
 * the rows the state re-gather reads (src_row).  live_images counts images that still have live beams. */
SET_API int set_beam_expand(int N, int K, int V, int step, int Lmax, int64_t end_tok, const float* logits_e,
                            const float* logits_d, int* k_live, float* beam_scores, const int64_t* seq_in, int64_t* seq_out,
                            int64_t* next_tokens, int* src_row, int* n_complete, float* complete_scores,
                            int64_t* complete_seqs, int* complete_len, int* live_images, void* stream);
/* out_s[r] = in_s[src_row[r]] for the four state tensors (rows x D) */
SET_API int set_beam_gather(int rows, int D, const int* src_row, const float* in0, const float* in1, const float* in2,
                            const float* in3, float* out0, float* out1, float* out2, float* out3, void* stream);
/* per image: the completed beam with the highest score (first occurrence), or -- runaway guard, editnet.py:702-713 --
 * the first 18 tokens of the first live beam */
SET_API int set_beam_finalize(int N, int K, int Lmax, int steps_done, const int* k_live, const int64_t* seq_live,
                              const float* beam_scores, const int* n_complete, const float* complete_scores,
                              const int64_t* complete_seqs, const int* complete_len, int64_t* out_seq, int* out_len,
                              float* out_score, void* stream);

/* Data-parallel overlap (new: the reference is single-process).  The reverse pass finishes the parameter gradients in
 * five groups ("buckets"), in this order:
 *   0  fc.*                                                          before the per-step loop
 *   1  attention_lstm.*, copy_lstm.*, caption_attention.cap_features_att.*   after the loop (big weight-gradient groups)
 *   2  embed.*, caption_encoder.*                                    input-gradient tail, encoder BPTT
 *   3  the rest of caption_attention.*, visual_attention.decoder_att / full_att
 *   4  visual_attention.features_att / att_embed, final when the call returns.
 * `events` = n <= 8 cudaEvent_t handles.  Arms the NEXT set_editnet_xe_backward / set_editnet_rollout_backward call of
 * this thread: the call records events[k] on its stream as soon as bucket k is final (events of buckets the call does
 * not know, and every event not recorded earlier, are recorded at its end).  A communication stream that waits for
 * events[k] before all-reducing bucket k overlaps the collective with the rest of the pass; the Python side
 * (train.XETrainer) keeps each bucket contiguous in its flat gradient buffer. */
SET_API int set_backward_bucket_events(void* const* events, int n);
/* Arms the NEXT set_editnet_xe_backward / set_editnet_rollout_backward call of this thread to INITIALISE the gradient
 * buffers instead of accumulating into them: weight-matrix gradients are written with beta = 0 and the accumulated /
 * scattered ones (biases, embedding table, the full_att rows) are zeroed by the call itself -- the caller need not
 * zero anything (trainers: saves a 355 MB memset and the read of every old gradient per step).  Default (not armed):
 * gradients accumulate, as autograd expects. */
SET_API int set_backward_overwrite_grads(int on);
/* persistent decode-step kernel (csrc/step_kernel.cu): launches since the last reset and the timesteps they covered
   (0 launches: the shape fell outside the persistent path and the per-step launch chain ran) */
SET_API int set_step_stats(long long* launches, long long* steps, int reset);
/* CTAs and CTAs-per-cluster of the persistent launch on the current device (0, 0: not available) */
SET_API int set_step_geometry(int* grid, int* cluster);
/* debugging: device buffer of >= steps * 8 * grid uint64; every CTA stamps %globaltimer at each phase boundary of each
   timestep: [(step * 8 + phase) * grid + cta], phases 0..6 = end of A, B, C1, C2, D, E, F.  NULL switches it off. */
SET_API int set_step_trace(void* buf);
/* tensor-core launches that took one of the many-tile configurations -- the persistent kernel of the big time-batched
   GEMMs (gemm_big_kernel) or the two-CTAs-per-SM ("twin") launch -- since the last reset (parity tests assert that the
   big time-batched GEMMs of the benchmarked configuration really ran on them) */
SET_API long long set_gemm_twin_launches(int reset);
/* C = A @ B^T style contraction through the library's GEMM engine (mode 0 NT, 1 NN, 2 TN). */
SET_API int set_gemm(int mode, int M, int N, int K, const float* A, long lda, const float* Bm, long ldb,
             const float* bias, float* C, long ldc, int beta, int act, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SET_B200_H_ */
