// extern "C" boundary of libvrdone_b200.so: argument checking, error text, dispatch to the kernel launchers.
#include <cstdio>
#include <cstring>
#include "../../include/vrdone_b200.h"
#include "kernels.h"

namespace {
thread_local char t_err[512] = {0};

int fail(const char* what) {
    snprintf(t_err, sizeof t_err, "%s", what);
    return 1;
}
int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(t_err, sizeof t_err, "%s: %s", what, cudaGetErrorString(e));
        return 1;
    }
    return 0;
}
Lay make_lay(const int32_t* row_seq, const int32_t* seqinfo, int R, int B) {
    Lay l;
    l.row_seq = row_seq;
    l.seqinfo = reinterpret_cast<const int4*>(seqinfo);
    l.R = R;
    l.B = B;
    return l;
}
}  // namespace

VrdOptions& vrd_options() {
    static VrdOptions o = [] {
        VrdOptions d;
        d.pdl = (getenv("VRD_PDL") != nullptr && atoi(getenv("VRD_PDL")) == 0) ? 0 : 1;
        d.dw_cfg = getenv("VRD_DW_CFG") != nullptr ? atoi(getenv("VRD_DW_CFG")) : 2;
        d.gemm_spec = getenv("VRD_GEMM_SPEC") != nullptr ? atoi(getenv("VRD_GEMM_SPEC")) : 1;
        d.embed_ln = (getenv("VRD_EMBED_LN") != nullptr && atoi(getenv("VRD_EMBED_LN")) == 0) ? 0 : 1;
        d.proj_ln = (getenv("VRD_PROJ_LN") != nullptr && atoi(getenv("VRD_PROJ_LN")) == 0) ? 0 : 1;
        return d;
    }();
    return o;
}

extern "C" {

int vrd_abi_version(void) { return VRD_ABI_VERSION; }

static int* option_slot(const char* name) {
    if (name == nullptr) return nullptr;
    VrdOptions& o = vrd_options();
    if (strcmp(name, "pdl") == 0) return &o.pdl;
    if (strcmp(name, "dw_cfg") == 0) return &o.dw_cfg;
    if (strcmp(name, "gemm_spec") == 0) return &o.gemm_spec;
    if (strcmp(name, "embed_ln") == 0) return &o.embed_ln;
    if (strcmp(name, "proj_ln") == 0) return &o.proj_ln;
    return nullptr;
}

int vrd_set_option(const char* name, int value) {
    int* slot = option_slot(name);
    if (slot == nullptr) { fail("vrd_set_option: unknown option"); return -1; }
    const int old = *slot;
    *slot = value;
    return old;
}

int vrd_get_option(const char* name) {
    const int* slot = option_slot(name);
    if (slot == nullptr) { fail("vrd_get_option: unknown option"); return -1; }
    return *slot;
}
const char* vrd_last_error(void) { return t_err; }

int vrd_device_arch(void) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { fail("cudaGetDevice failed (no CUDA device?)"); return -1; }
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
    return major * 10 + minor;
}

int vrd_h2d_pairs(const void* const* src, const int64_t* bytes, void* dst_base, const int64_t* dst_offset, int n,
                  vrd_stream_t stream) {
    if (n < 0 || (n > 0 && (src == nullptr || bytes == nullptr || dst_base == nullptr || dst_offset == nullptr)))
        return fail("vrd_h2d_pairs: bad arguments");
    for (int i = 0; i < n; ++i) {
        cudaError_t e = cudaMemcpyAsync((char*)dst_base + dst_offset[i], src[i], (size_t)bytes[i], cudaMemcpyHostToDevice,
                                        (cudaStream_t)stream);
        if (e != cudaSuccess) {
            snprintf(t_err, sizeof t_err, "vrd_h2d_pairs: copy %d of %d failed: %s", i, n, cudaGetErrorString(e));
            return 1;
        }
    }
    return 0;
}

int vrd_merge_layout(int n_chunks, const int32_t* const* row_seq, const int32_t* const* seqinfo, const int32_t* R, const int32_t* B,
                     int32_t* row_seq_out, int32_t* seqinfo_out, vrd_stream_t stream) {
    if (row_seq == nullptr || seqinfo == nullptr || R == nullptr || B == nullptr || row_seq_out == nullptr || seqinfo_out == nullptr)
        return fail("vrd_merge_layout: null argument");
    if (vrd::merge_layout(n_chunks, row_seq, seqinfo, R, B, row_seq_out, seqinfo_out, (cudaStream_t)stream) != 0)
        return fail("vrd_merge_layout: 1..16 chunks per call");
    return check_launch("vrd_merge_layout");
}

int vrd_upload(const void* host_src, void* dev_dst, int64_t bytes, vrd_stream_t stream) {
    if (host_src == nullptr || dev_dst == nullptr) return fail("vrd_upload: null argument");
    if (vrd::upload(host_src, dev_dst, bytes, (cudaStream_t)stream) != 0)
        return fail("vrd_upload: bytes must be a positive multiple of 4, both pointers 16-byte aligned");
    return check_launch("vrd_upload");
}

int vrd_viou_filter(const float* boxes, const int32_t* trk_base, const int32_t* durations, const int32_t* cat_ids, int n_tracklets,
                    float viou_threshold, double* sums, uint8_t* flags, int32_t* valid, vrd_stream_t stream) {
    if (boxes == nullptr || trk_base == nullptr || durations == nullptr || cat_ids == nullptr || flags == nullptr || valid == nullptr)
        return fail("vrd_viou_filter: null argument");
    if (n_tracklets < 1 || n_tracklets > 1024) return fail("vrd_viou_filter: n_tracklets must be in [1, 1024]");
    if (vrd::viou_filter(boxes, trk_base, durations, cat_ids, n_tracklets, viou_threshold, sums, flags, valid, (cudaStream_t)stream) != 0)
        return fail("vrd_viou_filter: bad arguments");
    return check_launch("vrd_viou_filter");
}

int vrd_pack_pairs(const void* pair_ptrs, const int64_t* pair_strides, const int32_t* row_seq, const int32_t* seqinfo, int R,
                   int B, int nv, int nc, int nbs, int nbe, void* vis, void* clip, int act_dtype, float* bbox_so,
                   float* bbox_ent, int token_major, vrd_stream_t stream) {
    if (nbs > 8 || nbe > 8 || R <= 0 || B <= 0 || (nv & 1) || (nc & 1)) return fail("vrd_pack_pairs: bad sizes (nv, nc must be even)");
    if (nc > 0 && clip == nullptr) return fail("vrd_pack_pairs: clip output missing");
    vrd::pack_pairs(pair_ptrs, (const long long*)pair_strides, make_lay(row_seq, seqinfo, R, B), nv, nc, nbs, nbe, vis, clip,
                    act_dtype, bbox_so, bbox_ent, token_major, (cudaStream_t)stream);
    return check_launch("vrd_pack_pairs");
}

int vrd_gemm(const void* A, int a_dtype, int64_t lda, const void* W, const float* bias, void* out, int out_dtype, int64_t ldo,
             int M, int N, int K, int taps, int act, const float* res1, int64_t ldr1, const float* res2, int64_t ldr2,
             const float* corr, const int32_t* row_seq, const int32_t* seqinfo, int R, vrd_stream_t stream) {
    if (taps != 1 && taps != 3) return fail("vrd_gemm: taps must be 1 or 3");
    if (corr != nullptr && row_seq == nullptr) return fail("vrd_gemm: corr needs a layout");
    vrd::GemmArgs g;
    g.A = A; g.lda = lda; g.W = W; g.bias = bias; g.out = out; g.out_dtype = out_dtype; g.ldo = ldo;
    g.M = M; g.N = N; g.K = K; g.taps = taps; g.act = act;
    g.res1 = res1; g.ldr1 = ldr1; g.res2 = res2; g.ldr2 = ldr2; g.corr = corr;
    g.row_seq = row_seq; g.seqinfo = reinterpret_cast<const int4*>(seqinfo); g.R = R;
    if (a_dtype == VRD_BF16) {
        if (vrd::gemm_tcgen05_bf16(g, (cudaStream_t)stream) != 0) return fail(vrd::gemm_tcgen05_error());
    } else if (a_dtype == VRD_F32) {
        const int rc = vrd::gemm_f32(g, (cudaStream_t)stream);
        if (rc == 2) return fail(vrd::gemm_tcgen05_error());
        if (rc != 0) return fail("vrd_gemm(fp32): K must be a multiple of 16 and lda of 4");
    } else {
        return fail("vrd_gemm: bad a_dtype");
    }
    return check_launch("vrd_gemm");
}

int vrd_gemm_ln(const void* A, int64_t lda, const void* W, const float* bias, const float* corr, const float* gamma, const float* beta,
                int relu, void* out, int64_t ldo, int M, int N, int K, int taps, const int32_t* row_seq, const int32_t* seqinfo, int R,
                vrd_stream_t stream) {
    if (taps != 1 && taps != 3) return fail("vrd_gemm_ln: taps must be 1 or 3");
    if (gamma == nullptr || beta == nullptr) return fail("vrd_gemm_ln: gamma and beta are required");
    if (corr != nullptr && row_seq == nullptr) return fail("vrd_gemm_ln: corr needs a layout");
    vrd::GemmArgs g;
    g.A = A; g.lda = lda; g.W = W; g.bias = bias; g.out = out; g.out_dtype = VRD_BF16; g.ldo = ldo;
    g.M = M; g.N = N; g.K = K; g.taps = taps; g.act = 0;
    g.res1 = nullptr; g.ldr1 = 0; g.res2 = nullptr; g.ldr2 = 0; g.corr = corr;
    g.row_seq = row_seq; g.seqinfo = reinterpret_cast<const int4*>(seqinfo); g.R = R;
    g.ln_gamma = gamma; g.ln_beta = beta; g.ln_relu = relu;
    if (vrd::gemm_tcgen05_bf16(g, (cudaStream_t)stream) != 0) return fail(vrd::gemm_tcgen05_error());
    return check_launch("vrd_gemm_ln");
}

int vrd_gemm_res_ln(const void* A, int64_t lda, const void* W, const float* bias, const float* res, int64_t ldr, const float* gamma,
                    const float* beta, float* out, int64_t ldo, void* ln_out, int64_t ld_ln, int M, int N, int K,
                    const int32_t* row_seq, const int32_t* seqinfo, int R, vrd_stream_t stream) {
    if (gamma == nullptr || beta == nullptr || res == nullptr || ln_out == nullptr) return fail("vrd_gemm_res_ln: null argument");
    vrd::GemmArgs g;
    g.A = A; g.lda = lda; g.W = W; g.bias = bias; g.out = out; g.out_dtype = VRD_F32; g.ldo = ldo;
    g.M = M; g.N = N; g.K = K; g.taps = 1; g.act = 0;
    g.res1 = res; g.ldr1 = ldr; g.res2 = nullptr; g.ldr2 = 0; g.corr = nullptr;
    g.row_seq = row_seq; g.seqinfo = reinterpret_cast<const int4*>(seqinfo); g.R = R;
    if (vrd::gemm_res_ln_fused_ok(g)) {
        g.ln_gamma = gamma; g.ln_beta = beta; g.ln_relu = 0; g.ln_out = ln_out; g.ld_ln = ld_ln;
        if (vrd::gemm_tcgen05_bf16(g, (cudaStream_t)stream) != 0) return fail(vrd::gemm_tcgen05_error());
        return check_launch("vrd_gemm_res_ln");
    }
    // switched off, or another row width: two launches
    if (vrd::gemm_tcgen05_bf16(g, (cudaStream_t)stream) != 0) return fail(vrd::gemm_tcgen05_error());
    if (vrd::layernorm(out, VRD_F32, ldo, gamma, beta, ln_out, VRD_BF16, ld_ln, M, N, 0, row_seq, R, (cudaStream_t)stream))
        return fail("vrd_gemm_res_ln: unsupported width for the LayerNorm");
    return check_launch("vrd_gemm_res_ln");
}

int vrd_layernorm(const void* x, int x_dtype, int64_t ldx, const float* gamma, const float* beta, void* out, int out_dtype,
                  int64_t ldo, int rows, int C, int relu, const int32_t* row_seq, int R, vrd_stream_t stream) {
    if (vrd::layernorm(x, x_dtype, ldx, gamma, beta, out, out_dtype, ldo, rows, C, relu, row_seq, R, (cudaStream_t)stream))
        return fail("vrd_layernorm: unsupported C / dtype combination");
    return check_launch("vrd_layernorm");
}

int vrd_small_conv(const float* x, int cin, const float* wt, const float* bias, const float* gamma, const float* beta, int relu,
                   void* out, int out_dtype, int64_t ldo, int rows, int N, const int32_t* row_seq, int R, vrd_stream_t stream) {
    if (row_seq == nullptr) return fail("vrd_small_conv: layout required");
    if (vrd::small_conv(x, cin, wt, bias, gamma, beta, relu, out, out_dtype, ldo, rows, N, row_seq, R, (cudaStream_t)stream))
        return fail("vrd_small_conv: N must be 512 and cin <= 8");
    return check_launch("vrd_small_conv");
}

int vrd_dwconv_ln(const void* x, int x_dtype, int64_t ldx, const int32_t* row_seq_in, const int32_t* seqinfo_in, int R_in,
                  const int32_t* row_seq_out, const int32_t* seqinfo_out, int R_out, int B, int stride, const float* pre_gamma,
                  const float* pre_beta, int n_branches, const float* const* w, const int32_t* use_pre,
                  const float* const* gamma, const float* const* beta, void* const* out, const int64_t* ldo, int out_dtype,
                  int C, int streams, vrd_stream_t stream) {
    if (n_branches < 1 || n_branches > 3) return fail("vrd_dwconv_ln: 1..3 branches");
    if (stride != 1 && stride != 2) return fail("vrd_dwconv_ln: stride must be 1 or 2");
    vrd::DwBranches br;
    memset(&br, 0, sizeof br);
    br.n = n_branches;
    for (int i = 0; i < n_branches; ++i) {
        br.w[i] = w[i]; br.use_pre[i] = use_pre[i]; br.g[i] = gamma[i]; br.b[i] = beta[i]; br.out[i] = out[i]; br.ldo[i] = ldo[i];
        if (use_pre[i] && pre_gamma == nullptr) return fail("vrd_dwconv_ln: pre-LN branch without pre-LN parameters");
    }
    if (vrd::dwconv_ln(x, x_dtype, ldx, make_lay(row_seq_in, seqinfo_in, R_in, B), make_lay(row_seq_out, seqinfo_out, R_out, B),
                       stride, pre_gamma, pre_beta, br, out_dtype, C, streams, (cudaStream_t)stream))
        return fail("vrd_dwconv_ln: unsupported C / dtype / branch combination (branches: qkv all pre-LN, q,k pre-LN + v raw, 2 raw, 1 pre-LN)");
    return check_launch("vrd_dwconv_ln");
}

int vrd_window_attn(const void* q, const void* k, const void* v, void* out, int dtype, int64_t ld, const int32_t* row_seq,
                    const int32_t* seqinfo, int R, int B, int n_head, int C, int w, int streams, vrd_stream_t stream) {
    if (vrd::window_attn(q, k, v, out, dtype, ld, make_lay(row_seq, seqinfo, R, B), n_head, C, w, streams, (cudaStream_t)stream))
        return fail("vrd_window_attn: needs C == 512, head_dim in {64, 128}, 1 <= w <= 4");
    return check_launch("vrd_window_attn");
}

int vrd_full_attn(const void* q, const void* k, const void* v, void* out, int dtype, int64_t ld, const int32_t* row_seq,
                  const int32_t* seqinfo, int R, int B, int n_head, int C, int max_len, const int32_t* attn_tiles, int n_attn_tiles,
                  vrd_stream_t stream) {
    Lay l = make_lay(row_seq, seqinfo, R, B);
    l.tiles = reinterpret_cast<const int4*>(attn_tiles);
    l.n_tiles = attn_tiles != nullptr ? n_attn_tiles : 0;
    if (vrd::full_attn(q, k, v, out, dtype, ld, l, n_head, C, max_len, (cudaStream_t)stream))
        return fail("vrd_full_attn: needs head_dim in {64, 128} and B <= 65535");
    return check_launch("vrd_full_attn");
}

int vrd_maxpool_skip(const float* x, int64_t ldx, const int32_t* row_seq_in, const int32_t* seqinfo_in, int R_in,
                     const int32_t* row_seq_out, const int32_t* seqinfo_out, int R_out, int B, float* out, int64_t ldo, int C,
                     vrd_stream_t stream) {
    if (vrd::maxpool_skip(x, ldx, make_lay(row_seq_in, seqinfo_in, R_in, B), make_lay(row_seq_out, seqinfo_out, R_out, B), out, ldo,
                          C, (cudaStream_t)stream))
        return fail("vrd_maxpool_skip: C must be 512");
    return check_launch("vrd_maxpool_skip");
}

int vrd_fpn_top(const float* x, int64_t ldx, const int32_t* row_seq, const int32_t* seqinfo, int R, int B, const float* pre_gamma,
                const float* pre_beta, const float* wt, const float* gamma, const float* beta, float* out, int64_t ldo,
                vrd_stream_t stream) {
    vrd::fpn_top(x, ldx, make_lay(row_seq, seqinfo, R, B), pre_gamma, pre_beta, wt, gamma, beta, out, ldo, (cudaStream_t)stream);
    return check_launch("vrd_fpn_top");
}

int vrd_fpn_level(const float* cur, int64_t ldc, const float* y_up, int64_t ldu, const int32_t* row_seq, const int32_t* seqinfo,
                  int R, const int32_t* row_seq_up, const int32_t* seqinfo_up, int R_up, int B, const float* lat_gamma,
                  const float* lat_beta, const float* beta_up, const float* w, const float* gamma, const float* beta,
                  float* out, int64_t ldo, vrd_stream_t stream) {
    vrd::fpn_level(cur, ldc, y_up, ldu, make_lay(row_seq, seqinfo, R, B), make_lay(row_seq_up, seqinfo_up, R_up, B), lat_gamma,
                   lat_beta, beta_up, w, gamma, beta, out, ldo, (cudaStream_t)stream);
    return check_launch("vrd_fpn_level");
}

int vrd_mask_features(const float* y, int64_t ldy, const int32_t* row_seq, const int32_t* seqinfo, int R, int B,
                      const float* beta, const float* w, const float* bias, float* out, int64_t ldo, vrd_stream_t stream) {
    vrd::mask_features(y, ldy, make_lay(row_seq, seqinfo, R, B), beta, w, bias, out, ldo, (cudaStream_t)stream);
    return check_launch("vrd_mask_features");
}

int vrd_query_ln(const float* x, int64_t ldx, const float* gamma, const float* beta, const float* pos, int Q, int nrows,
                 int total_rows, const float* dw, const float* gamma2, const float* beta2, void* out, int out_dtype,
                 int64_t ldo, int C, vrd_stream_t stream) {
    if (vrd::query_ln(x, ldx, gamma, beta, pos, Q, nrows, total_rows, dw, gamma2, beta2, out, out_dtype, ldo, C,
                      (cudaStream_t)stream))
        return fail("vrd_query_ln: C must be 256");
    return check_launch("vrd_query_ln");
}

int vrd_query_self_attn(const void* q, const void* k, const void* v, void* out, int dtype, int64_t ld, int B, int Q, int n_head,
                        int C, vrd_stream_t stream) {
    if (vrd::query_self_attn(q, k, v, out, dtype, ld, B, Q, n_head, C, (cudaStream_t)stream))
        return fail("vrd_query_self_attn: needs C == 256, Q <= 12, n_head <= 8");
    return check_launch("vrd_query_self_attn");
}

int vrd_query_cross_attn(const void* q, const void* k, const void* v, void* out, int dtype, int64_t ld, const int32_t* row_seq,
                         const int32_t* seqinfo, int R, int B, int Q, int n_head, int C, vrd_stream_t stream) {
    if (vrd::query_cross_attn(q, k, v, out, dtype, ld, make_lay(row_seq, seqinfo, R, B), Q, n_head, C, (cudaStream_t)stream))
        return fail("vrd_query_cross_attn: needs C == 256, Q <= 12, head_dim in {32, 64}");
    return check_launch("vrd_query_cross_attn");
}

int vrd_mask_logits(const float* mask_embed, int64_t ldm, const float* mask_feat, int64_t ldf, const int32_t* row_seq,
                    const int32_t* seqinfo, int R, int B, int Q, float* masks, int64_t ldk, int32_t* first_last,
                    vrd_stream_t stream) {
    if (vrd::mask_logits(mask_embed, ldm, mask_feat, ldf, make_lay(row_seq, seqinfo, R, B), Q, masks, ldk, first_last,
                         (cudaStream_t)stream))
        return fail("vrd_mask_logits: Q must be <= 16");
    return check_launch("vrd_mask_logits");
}

int vrd_softmax_topk(const float* logits, int64_t ldl, int nrows, int n_cls, int topk, float* scores, int32_t* ids,
                     vrd_stream_t stream) {
    if (vrd::softmax_topk(logits, ldl, nrows, n_cls, topk, scores, ids, (cudaStream_t)stream))
        return fail("vrd_softmax_topk: needs n_cls <= 256 and topk < n_cls");
    return check_launch("vrd_softmax_topk");
}

int vrd_rank_triplets(const float* topk_scores, const int32_t* topk_ids, const int32_t* first_last, const int64_t* sids,
                      const int64_t* oids, const float* cat_scores, const int64_t* traj_durations, const int64_t* so_offset, int B,
                      int Q, int topk, int feat_stride, int pred_min_frames, int n_max, uint64_t* keys, int32_t* header,
                      int32_t* records, vrd_stream_t stream) {
    if (!topk_scores || !topk_ids || !first_last || !sids || !oids || !cat_scores || !traj_durations || !so_offset || !keys || !header ||
        !records)
        return fail("vrd_rank_triplets: null argument");
    if (vrd::rank_triplets(topk_scores, topk_ids, first_last, (const long long*)sids, (const long long*)oids, cat_scores,
                           (const long long*)traj_durations, (const long long*)so_offset, B, Q, topk, feat_stride, pred_min_frames, n_max,
                           (unsigned long long*)keys, header, records, (cudaStream_t)stream))
        return fail("vrd_rank_triplets: need 1 <= n_max <= 1024 and B * Q * topk < 2^32 - 1");
    return check_launch("vrd_rank_triplets");
}

}  // extern "C"
