/* vrdone_b200 -- C ABI of the B200-native MaskVRD inference kernels (libvrdone_b200.so).
 *
 * The reference (lucaspk512/vrdone) has no FFI layer: its "operator API" for this path is the Python class
 * models/maskvrd.py:16 `MaskVRD` whose arithmetic is PyTorch library calls.  Each entry
 * point below replaces one group of those library calls on the packed varlen layout of vrdone_b200/layout.py; the
 * Python mirror `vrdone_b200.MaskVRD` binds them with ctypes (vrdone_b200/cuda_ops.py) and keeps the reference's
 * constructor / config / checkpoint / forward contract.  INTEGRATION.md shows the binding a maintainer adds.
 *
 * Conventions: every pointer is a DEVICE pointer unless stated otherwise; matrices are row-major `[rows, cols]` with a
 * leading dimension `ld*` in ELEMENTS; `*_dtype` is VRD_F32 or VRD_BF16; `stream` is a cudaStream_t; kernels are
 * enqueued asynchronously.  Every function returns 0 on success and non-zero on error, with text in vrd_last_error().
 *
 * A pyramid level of the row layout is passed as (row_seq, seqinfo, R, B):
 *   row_seq[r]  int32, owning pair of row r, -1 for a separator row        (R entries, R % 128 == 0)
 *   seqinfo[i]  int32x4 (first row, valid rows, first-pad-column-exists, 0) (B entries, 16-byte aligned)
 * Kernels that take `streams` operate on `streams` stacked copies of the level (subject rows then object rows).
 */
#ifndef VRDONE_B200_H
#define VRDONE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VRD_F32 0
#define VRD_BF16 1
#define VRD_ACT_NONE 0
#define VRD_ACT_RELU 1
#define VRD_ACT_GELU 2
#define VRD_ABI_VERSION 4

typedef void* vrd_stream_t; /* cudaStream_t */

int vrd_abi_version(void);
const char* vrd_last_error(void);
/* compute capability major*10+minor of the current device (100 on B200); <0 on error */
int vrd_device_arch(void);
/* experiment switches of the launchers -- "pdl": programmatic dependent launch on / off, "dw_cfg": dwconv_ln_tile variant,
 * "gemm_spec": tcgen05 GEMM epilogue modes, "embed_ln" / "proj_ln": LayerNorm as the epilogue of the embedding convs / of the encoder
 * blocks' attention projection; the defaults come from the environment (VRD_PDL, VRD_DW_CFG, VRD_GEMM_SPEC, VRD_EMBED_LN,
 * VRD_PROJ_LN).  vrd_set_option returns the previous value, vrd_get_option the current one, <0 for an unknown
 * name.  No reference counterpart: they exist so that one process can A/B a switch on the same inputs. */
int vrd_set_option(const char* name, int value);
int vrd_get_option(const char* name);

/* a0 -- replaces utils.dict_to_device for the pair features (eval.py:144, utils/misc.py:98-112): n asynchronous
 * host->device copies on `stream` (the copy engine; src[i] HOST pointers, pinned for true asynchrony) of bytes[i] bytes to
 * dst_base + dst_offset[i].  The caller overlaps them with the kernels of the previous chunk on another stream. */
int vrd_h2d_pairs(const void* const* src, const int64_t* bytes, void* dst_base, const int64_t* dst_offset, int n,
                  vrd_stream_t stream);

/* a1 -- one pyramid level of n_chunks (<= 16) consecutive chunk layouts merged into the layout of ONE batch on the device (rows
 * of chunk c follow those of chunk c - 1, pair ids renumbered, row offsets shifted): the slices of MaskVRD.forward_test
 * (maskvrd.py:201-240) seen as one batch by the query decoder and the heads.  row_seq / seqinfo / R / B are HOST arrays holding
 * the device pointers and the row / pair counts of the chunk layouts; the outputs have sum(R) and 4 * sum(B) int32 entries. */
int vrd_merge_layout(int n_chunks, const int32_t* const* row_seq, const int32_t* const* seqinfo, const int32_t* R, const int32_t* B,
                     int32_t* row_seq_out, int32_t* seqinfo_out, vrd_stream_t stream);

/* a0 -- small host->device upload done by a kernel (host_src: PINNED host memory, read by the SMs through UVA) instead of the
 * copy engine, for per-video bookkeeping arrays on a stream whose kernels must not wait behind bulk copies of other streams
 * (replaces the .to(device) of small index tensors, utils/misc.py:98-112).  bytes % 4 == 0, both pointers 16-byte aligned. */
int vrd_upload(const void* host_src, void* dev_dst, int64_t bytes, vrd_stream_t stream);

/* SURVEY 8f row 2 -- replaces the duplicate-tracklet filter of the data loader (dataloaders/vidor.py:583-641, vidvrd.py same
 * code): boxes [T, 4] fp32 hold every tracklet's (already clamped) boxes back to back, trk_base[i] the first row of tracklet
 * i, durations [N, 2] int32 (start, end frame), cat_ids [N] int32.  For base < ref of one category with overlapping durations
 * the intersection / base / ref box volumes over the common frames (fp64 sums of the reference's fp32 per-frame terms) decide
 * rule 1 (ref dropped: inter / vol_ref > thr and base covers ref in time) or rule 2 (base dropped); flags [N, N] uint8 receives
 * the decision (0 / 1 / 2), sums [N, N, 3] fp64 (optional, may be NULL) the three volumes, valid [N] int32 the result of the
 * reference's greedy scan in tracklet order.  N <= 1024. */
int vrd_viou_filter(const float* boxes, const int32_t* trk_base, const int32_t* durations, const int32_t* cat_ids, int n_tracklets,
                    float viou_threshold, double* sums, uint8_t* flags, int32_t* valid, vrd_stream_t stream);

/* k9 -- replaces MaskVRD.preprocessing (maskvrd.py:363-414) + the channel split of backbones.py:161-166 / 329-341.
 * pair_ptrs[i] -> fp32 (C, L_i) tensor with element strides pair_strides[2i] (channel), pair_strides[2i+1] (time).
 * Writes vis [2R, nv], clip [2R, nc] (or NULL when nc == 0) in act_dtype, bbox_so [R, 8] and bbox_ent [2R, 8] in fp32.
 * token_major != 0 promises that every pair has channel stride 1 (the loader's (L, C) buffer): warp-per-row fast path. */
int vrd_pack_pairs(const void* pair_ptrs, const int64_t* pair_strides, const int32_t* row_seq, const int32_t* seqinfo, int R,
                   int B, int nv, int nc, int nbs, int nbe, void* vis, void* clip, int act_dtype, float* bbox_so,
                   float* bbox_ent, int token_major, vrd_stream_t stream);

/* k1/k2 -- replaces nn.Conv1d (k=1, and dense k=3 as three row-shifted K-slabs; blocks.py:85, 46, 728-737, 1054-1060):
 * out = act(A * W^T + bias [+ corr on the last valid row of pairs with a pad column]) + res1 + res2, separator rows = 0.
 * a_dtype VRD_BF16 -> tcgen05/TMEM tensor-core kernel (W bf16); VRD_F32 -> fp32 CUDA-core kernel (W fp32).
 * W is [N, taps*K]; row_seq may be NULL (no layout: all rows valid). */
int vrd_gemm(const void* A, int a_dtype, int64_t lda, const void* W, const float* bias, void* out, int out_dtype, int64_t ldo,
             int M, int N, int K, int taps, int act, const float* res1, int64_t ldr1, const float* res2, int64_t ldr2,
             const float* corr, const int32_t* row_seq, const int32_t* seqinfo, int R, vrd_stream_t stream);

/* a7: the attention output projection of an encoder block with the residual and the LayerNorm that feeds the MLP
 * (blocks.py:1070-1076): out[M, 512] (fp32) = (A * W^T + bias + res) * row validity -- the new residual stream -- and
 * ln_out[M, 512] (bf16) = LayerNorm_channels(out) * row validity, from one pass over the accumulator row.  bf16 operands.  With
 * "proj_ln" = 0 (or N != 512) the call runs the projection and the LayerNorm as two launches. */
int vrd_gemm_res_ln(const void* A, int64_t lda, const void* W, const float* bias, const float* res, int64_t ldr, const float* gamma,
                    const float* beta, float* out, int64_t ldo, void* ln_out, int64_t ld_ln, int M, int N, int K,
                    const int32_t* row_seq, const int32_t* seqinfo, int R, vrd_stream_t stream);

/* k1 + k6 fused -- the embedding convs followed by their channel LayerNorm and ReLU (backbones.py:184-197 with blocks.py:143-158):
 * out[M, 512] (bf16) = [relu](LN_channels(A * W^T + bias [+ corr on the last row of padded pairs])) * row validity.  bf16 operands
 * only (tcgen05 path), N must be 512: the tile spans the row, the statistics are taken on the fp32 accumulator in tensor memory and
 * the fp32 conv output never reaches HBM.  Same operand conventions as vrd_gemm. */
int vrd_gemm_ln(const void* A, int64_t lda, const void* W, const float* bias, const float* corr, const float* gamma, const float* beta,
                int relu, void* out, int64_t ldo, int M, int N, int K, int taps, const int32_t* row_seq, const int32_t* seqinfo, int R,
                vrd_stream_t stream);

/* k6 -- channel LayerNorm (blocks.py:143-158) [+ReLU]; C in {256, 512}; separator rows -> 0 when row_seq != NULL. */
int vrd_layernorm(const void* x, int x_dtype, int64_t ldx, const float* gamma, const float* beta, void* out, int out_dtype,
                  int64_t ldo, int rows, int C, int relu, const int32_t* row_seq, int R, vrd_stream_t stream);

/* tiny-K k=3 conv of the box-geometry channels (backbones.py:199-202, 232): x [rows, 8] fp32, wt [3*cin, 512]. */
int vrd_small_conv(const float* x, int cin, const float* wt, const float* bias, const float* gamma, const float* beta, int relu,
                   void* out, int out_dtype, int64_t ldo, int rows, int N, const int32_t* row_seq, int R, vrd_stream_t stream);

/* k3 -- [LN_pre] -> depthwise k=3 conv (stride 1|2) -> mask -> LN for up to 3 branches (q/k/v) sharing one read of x
 * (blocks.py:927-933, local_transformer.py:150-156, 561-567).  w[b] is [3, C] tap-major. */
int vrd_dwconv_ln(const void* x, int x_dtype, int64_t ldx, const int32_t* row_seq_in, const int32_t* seqinfo_in, int R_in,
                  const int32_t* row_seq_out, const int32_t* seqinfo_out, int R_out, int B, int stride, const float* pre_gamma,
                  const float* pre_beta, int n_branches, const float* const* w, const int32_t* use_pre,
                  const float* const* gamma, const float* const* beta, void* const* out, const int64_t* ldo, int out_dtype,
                  int C, int streams, vrd_stream_t stream);

/* k4 -- windowed attention, softmax over keys |i-j| <= w inside each pair (blocks.py:950-986). C = 512. */
int vrd_window_attn(const void* q, const void* k, const void* v, void* out, int dtype, int64_t ld, const int32_t* row_seq,
                    const int32_t* seqinfo, int R, int B, int n_head, int C, int w, int streams, vrd_stream_t stream);

/* k5 -- full attention inside each pair (local_transformer.py:170-183); max_len = longest pair of the level.  attn_tiles
 * (optional, may be NULL): int32x4 (first row, pair's first row, pair length, 0) per 128-row query tile of every pair, longest
 * pairs first (vrdone_b200/layout.py) -- with it the bf16 / head_dim 64 case runs on the tcgen05 / TMEM kernel. */
int vrd_full_attn(const void* q, const void* k, const void* v, void* out, int dtype, int64_t ld, const int32_t* row_seq,
                  const int32_t* seqinfo, int R, int B, int n_head, int C, int max_len, const int32_t* attn_tiles, int n_attn_tiles,
                  vrd_stream_t stream);

/* k7 -- MaxPool1d(3, 2, 1) skip path of the stride-2 blocks (blocks.py:1040-1046, 1074). fp32, C = 512. */
int vrd_maxpool_skip(const float* x, int64_t ldx, const int32_t* row_seq_in, const int32_t* seqinfo_in, int R_in,
                     const int32_t* row_seq_out, const int32_t* seqinfo_out, int R_out, int B, float* out, int64_t ldo, int C,
                     vrd_stream_t stream);

/* FPN (fpns.py:229-257): top level grouped conv; lower levels lateral-LN + nearest-up2 add + depthwise conv + LN;
 * mask_features depthwise conv + bias.  fp32, fpn_dim = 256, input channels 512. */
int vrd_fpn_top(const float* x, int64_t ldx, const int32_t* row_seq, const int32_t* seqinfo, int R, int B, const float* pre_gamma,
                const float* pre_beta, const float* wt, const float* gamma, const float* beta, float* out, int64_t ldo,
                vrd_stream_t stream);
int vrd_fpn_level(const float* cur, int64_t ldc, const float* y_up, int64_t ldu, const int32_t* row_seq, const int32_t* seqinfo,
                  int R, const int32_t* row_seq_up, const int32_t* seqinfo_up, int R_up, int B, const float* lat_gamma,
                  const float* lat_beta, const float* beta_up, const float* w, const float* gamma, const float* beta,
                  float* out, int64_t ldo, vrd_stream_t stream);
int vrd_mask_features(const float* y, int64_t ldy, const int32_t* row_seq, const int32_t* seqinfo, int R, int B,
                      const float* beta, const float* w, const float* bias, float* out, int64_t ldo, vrd_stream_t stream);

/* predictor query rows (local_transformer.py:807-835, 956-976): out = LN2(dw * (LN(x) + pos[row % Q])), stages optional. */
int vrd_query_ln(const float* x, int64_t ldx, const float* gamma, const float* beta, const float* pos, int Q, int nrows,
                 int total_rows, const float* dw, const float* gamma2, const float* beta2, void* out, int out_dtype,
                 int64_t ldo, int C, vrd_stream_t stream);
int vrd_query_self_attn(const void* q, const void* k, const void* v, void* out, int dtype, int64_t ld, int B, int Q, int n_head,
                        int C, vrd_stream_t stream);
int vrd_query_cross_attn(const void* q, const void* k, const void* v, void* out, int dtype, int64_t ld, const int32_t* row_seq,
                         const int32_t* seqinfo, int R, int B, int Q, int n_head, int C, vrd_stream_t stream);

/* k10 -- heads epilogue (predictor.py:103-105, maskvrd.py:247-248, 287-295): mask logits <mask_embed, mask_features>,
 * sigmoid(.) > 0.5 in fp32, first/last active frame per (pair, query) (-1, -1 if none); class softmax + top-k over
 * classes 1..n_cls-1 (1-based ids; ties -> lower id).  masks may be NULL. */
int vrd_mask_logits(const float* mask_embed, int64_t ldm, const float* mask_feat, int64_t ldf, const int32_t* row_seq,
                    const int32_t* seqinfo, int R, int B, int Q, float* masks, int64_t ldk, int32_t* first_last,
                    vrd_stream_t stream);
int vrd_softmax_topk(const float* logits, int64_t ldl, int nrows, int n_cls, int topk, float* scores, int32_t* ids,
                     vrd_stream_t stream);

/* k10, device-side compaction -- replaces the candidate loop, the mean-score ranking and the top-n_max_pair cut of
 * MaskVRD.forward_test (maskvrd.py:262-328) on the outputs of vrd_softmax_topk / vrd_mask_logits.  Candidate
 * c = (pair * Q + query) * topk + j is kept iff last >= 0 and (last - first) * feat_stride + 1 >= pred_min_frames; its score
 * is ((cat_scores[sid] + topk_score) + cat_scores[oid]) / 3 in fp32; the n_max best (descending score, ties to the lower c)
 * are written in rank order as records[i] = (c, score bits, predicate score bits, 1-based predicate id, first, last).
 * sids / oids / so_offset [B] int64, cat_scores [N] fp32, traj_durations [N, 2] int64 are the reference's own per-video
 * tensors (vidor.py:716-734) on the device.  header[0] = number of records, header[1] != 0 if a kept candidate violates the
 * reference's assert 0 <= start, end <= overlap (maskvrd.py:297).  keys: scratch of B*Q*topk uint64.  n_max <= 1024. */
int vrd_rank_triplets(const float* topk_scores, const int32_t* topk_ids, const int32_t* first_last, const int64_t* sids,
                      const int64_t* oids, const float* cat_scores, const int64_t* traj_durations, const int64_t* so_offset, int B,
                      int Q, int topk, int feat_stride, int pred_min_frames, int n_max, uint64_t* keys, int32_t* header,
                      int32_t* records, vrd_stream_t stream);

/* ---- native backbone schedule ------------------------------------------------------------------------------------
 * Replaces the per-operator Python schedule for MaskConvTransformerBackbone.forward + FPN1D_Fuse.forward
 * (backbones.py:154-248 / 323-436, fpns.py:229-257) of one chunk of packed pairs: the same kernels in the same order,
 * issued from C++ (one call instead of ~100 ctypes calls).  The query decoder and the heads stay per-operator. */
#define VRD_MAX_LEVELS 8
typedef struct {
    int32_t visual_dim, clip_dim /* 0: no CLIP stream */, bbox_so_dim, bbox_entity_dim, embd_dim, n_head, fuse_head;
    int32_t n_conv, n_stem, n_branch, win /* n_mha_win_size */, use_local, fpn_dim, act_dtype /* VRD_F32 | VRD_BF16 */;
} vrd_model_cfg_t;
typedef struct {
    const int32_t* row_seq; /* device, R entries */
    const int32_t* seqinfo; /* device, B x 4 */
    int32_t R, B, max_len;
    int32_t n_attn_tiles;       /* level 0 only (0 elsewhere): query tiles of vrd_full_attn */
    const int32_t* attn_tiles;  /* device, n_attn_tiles x 4, or NULL */
} vrd_level_t;
typedef struct vrd_engine vrd_engine_t;

const char* vrd_engine_last_error(void);
/* names[i] / ptrs[i]: kernel-layout weights of engine.PackedWeights (device pointers) with their 2-D shape rows[i] x cols[i]. */
int vrd_engine_create(const vrd_model_cfg_t* cfg, const char* const* names, const void* const* ptrs, const int32_t* rows,
                      const int32_t* cols, int n, vrd_engine_t** out);
void vrd_engine_destroy(vrd_engine_t* engine);
int64_t vrd_engine_launches(const vrd_engine_t* engine);   /* kernels launched so far through this engine */
/* levels: n_branch + 1 entries.  Workspace need of vrd_backbone_pack + vrd_backbone_compute for this chunk (-1 on error). */
int64_t vrd_backbone_workspace_bytes(vrd_engine_t* engine, const vrd_level_t* levels);
/* pack kernel only (vrd_pack_pairs into the head of the workspace); the caller may recycle the pair tensors once it has run */
int vrd_backbone_pack(vrd_engine_t* engine, const vrd_level_t* levels, const void* pair_ptrs, const int64_t* pair_strides,
                      int token_major, void* workspace, int64_t workspace_bytes, vrd_stream_t stream);
/* SURVEY 8f row 1 -- the data loader's pair construction on the device (dataloaders/vidor.py:659-711, utils/misc.py:158-217)
 * instead of vrd_backbone_pack: vis_all [T, visual_dim] / clip_all [T, clip_dim] / boxes_all [T, 4] fp32 hold every tracklet's
 * frames back to back; pair_tab[i] = (row of the subject's first sub-sampled frame, same for the object, feat_stride, 0).
 * Copies the two feature rows of every packed row and computes the 5-d relative and 8-d entity box features. */
int vrd_backbone_pack_tracklets(vrd_engine_t* engine, const vrd_level_t* levels, const float* vis_all, const float* clip_all,
                                const float* boxes_all, const int32_t* pair_tab, float video_w, float video_h, void* workspace,
                                int64_t workspace_bytes, vrd_stream_t stream);
/* everything after the pack: e_top [R_top, embd_dim] fp32 (coarsest level), mask_feat [R_0, fpn_dim] fp32 */
int vrd_backbone_compute(vrd_engine_t* engine, const vrd_level_t* levels, void* workspace, int64_t workspace_bytes, float* e_top,
                         float* mask_feat, vrd_stream_t stream);

/* a15-a16 -- the query decoder and the heads (predictor.py:85-115, local_transformer.py:773-835, 875-976) as ONE native
 * schedule over a batch described by its level 0 (mask features, R_0 rows) and its coarsest level (e_top, R_top rows): the
 * ~80 launches of MaskedTransformerPredictor.forward + the top-k / mask epilogue of maskvrd.py:242-299.  Outputs: logits
 * [ceil128(B*Q), n_cls_pad] fp32, topk_scores / topk_ids [B*Q, topk], first_last [B, Q, 2] int32 (first / last frame with
 * sigmoid(mask) > 0.5, -1 if none), masks [R_0, Q] fp32 mask logits or NULL. */
typedef struct {
    int32_t n_embd, num_queries, n_head, num_layers, n_hidden, n_cls /* K + 1 */, n_cls_pad /* rows of the padded class head */;
} vrd_predictor_cfg_t;
int64_t vrd_predict_workspace_bytes(vrd_engine_t* engine, const vrd_predictor_cfg_t* cfg, const vrd_level_t* level0,
                                    const vrd_level_t* level_top);
int vrd_predict(vrd_engine_t* engine, const vrd_predictor_cfg_t* cfg, const vrd_level_t* level0, const vrd_level_t* level_top,
                const float* e_top, const float* mask_feat, int topk, void* workspace, int64_t workspace_bytes, float* logits,
                float* topk_scores, int32_t* topk_ids, int32_t* first_last, float* masks, vrd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* VRDONE_B200_H */
