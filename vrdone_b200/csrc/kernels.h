// Internal C++ launch interface between the C-ABI (cabi.cu) and the kernel translation units.
#pragma once
#include <cuda_runtime.h>
#include "common.cuh"

namespace vrd {

struct DwBranches {
    int n;
    const float* w[3];     // depthwise taps, [3, C] tap-major
    int use_pre[3];        // 1: the branch convolves LN_pre(x), 0: raw x
    const float* g[3];     // post-conv LayerNorm gamma / beta
    const float* b[3];
    void* out[3];
    long long ldo[3];
};

struct GemmArgs;
// true when gemm_tcgen05_bf16 takes the residual + LayerNorm epilogue for this problem (N = 512; any number of rows, so that results
// do not depend on how a video is chunked)
bool gemm_res_ln_fused_ok(const GemmArgs& g);

struct GemmArgs {
    const void* A; long long lda;      // [M, K] activations (row r, tap d reads row r + d - 1 when taps == 3)
    const void* W;                     // [N, taps*K], K contiguous
    const float* bias;                 // [N] or null
    void* out; int out_dtype; long long ldo;
    int M, N, K, taps, act;
    const float* res1; long long ldr1; // optional fp32 residuals added after the activation
    const float* res2; long long ldr2;
    const float* corr;                 // optional [N] vector added (before the activation) on the last row of pairs with haspad
    const int* row_seq; const int4* seqinfo; int R;   // optional layout: separator rows are written as zeros
    // optional channel LayerNorm (+ ReLU) of the N = 512 outputs in the epilogue (tcgen05 bf16 path only; bf16 output, no residual)
    const float* ln_gamma = nullptr; const float* ln_beta = nullptr; int ln_relu = 0;
    // with a residual (res1): out = fp32 sum (the residual stream), ln_out = bf16 LayerNorm of it
    void* ln_out = nullptr; long long ld_ln = 0;
};

void pack_pairs(const void* ptrs, const long long* strides, Lay lay, int nv, int nc, int nbs, int nbe, void* vis, void* clip,
                int adt, float* bso, float* bent, int token_major, cudaStream_t st);
int pack_tracklets(const float* vis_all, const float* clip_all, const float* boxes_all, const int* pair_tab, Lay lay, int nv, int nc,
                   float vw, float vh, void* vis, void* clip, int adt, float* bso, float* bent, cudaStream_t st);
int viou_filter(const float* boxes, const int* trk_base, const int* durs, const int* cat_ids, int N, float thr, double* sums,
                unsigned char* flags, int* valid, cudaStream_t st);
int merge_layout(int n, const int* const* rs, const int* const* si, const int* R, const int* B, int* rs_out, int* si_out,
                 cudaStream_t st);
int upload(const void* host_src, void* dev_dst, long long bytes, cudaStream_t st);
int layernorm(const void* x, int xdt, long long ldx, const float* g, const float* b, void* out, int odt, long long ldo, int rows,
              int C, int relu, const int* row_seq, int R, cudaStream_t st);
int small_conv(const float* x, int cin, const float* wt, const float* bias, const float* g, const float* b, int relu, void* out,
               int odt, long long ldo, int rows, int N, const int* row_seq, int R, cudaStream_t st);
int dwconv_ln(const void* x, int xdt, long long ldx, Lay lin, Lay lout, int stride, const float* pre_g, const float* pre_b,
              const DwBranches& br, int odt, int C, int streams, cudaStream_t st);
int maxpool_skip(const float* x, long long ldx, Lay lin, Lay lout, float* out, long long ldo, int C, cudaStream_t st);
int fpn_top(const float* x, long long ldx, Lay lay, const float* pre_g, const float* pre_b, const float* wt, const float* g,
            const float* b, float* out, long long ldo, cudaStream_t st);
int fpn_level(const float* cur, long long ldc, const float* yup, long long ldu, Lay lay, Lay lup, const float* lat_g,
              const float* lat_b, const float* beta_up, const float* w, const float* g, const float* b, float* out,
              long long ldo, cudaStream_t st);
int mask_features(const float* y, long long ldy, Lay lay, const float* beta, const float* w, const float* bias, float* out,
                  long long ldo, cudaStream_t st);
int query_ln(const float* x, long long ldx, const float* g, const float* b, const float* pos, int Q, int nrows, int total_rows,
             const float* dw, const float* g2, const float* b2, void* out, int odt, long long ldo, int C, cudaStream_t st);
int mask_logits(const float* me, long long ldm, const float* mf, long long ldf, Lay lay, int Q, float* masks, long long ldk,
                int* first_last, cudaStream_t st);
int softmax_topk(const float* logits, long long ldl, int nrows, int n_cls, int topk, float* scores, int* ids, cudaStream_t st);
int rank_triplets(const float* topk_scores, const int* topk_ids, const int* first_last, const long long* sids, const long long* oids,
                  const float* cat_scores, const long long* durs, const long long* so_offset, int B, int Q, int topk, int feat_stride,
                  int pred_min_frames, int n_max, unsigned long long* keys, int* header, int* records, cudaStream_t st);

int window_attn(const void* q, const void* k, const void* v, void* out, int dt, long long ld, Lay lay, int n_head, int C, int w,
                int streams, cudaStream_t st);
int full_attn(const void* q, const void* k, const void* v, void* out, int dt, long long ld, Lay lay, int n_head, int C,
              int max_len, cudaStream_t st);
// tcgen05 / TMEM / TMA varlen attention (bf16, head_dim 64): 0 ok, 1 unsupported shape (caller falls back), 2 launch error
int full_attn_tcgen05(const void* q, const void* k, const void* v, void* out, long long ld, Lay lay, int n_head, int C, cudaStream_t st);
int query_self_attn(const void* q, const void* k, const void* v, void* out, int dt, long long ld, int B, int Q, int n_head, int C,
                    cudaStream_t st);
int query_cross_attn(const void* q, const void* k, const void* v, void* out, int dt, long long ld, Lay lay, int Q, int n_head,
                     int C, cudaStream_t st);

int gemm_simt_f32(const GemmArgs& a, cudaStream_t st);
// fp32 operands on tcgen05 as 3 x bf16 split products (0 ok, 1 shape unsupported -> use gemm_simt_f32, 2 error: gemm_tcgen05_error())
int gemm_tcgen05_f32split(const GemmArgs& a, cudaStream_t st);
// the fp32 path's GEMM: tensor cores when the shape fits (VRD_FP32_GEMM=simt forces the CUDA-core kernel), else CUDA cores
int gemm_f32(const GemmArgs& a, cudaStream_t st);
int gemm_tcgen05_bf16(const GemmArgs& a, cudaStream_t st);   // returns non-zero + sets error text on failure
const char* gemm_tcgen05_error();

}  // namespace vrd
