// Native schedule of the MaskVRD backbone + FPN over one chunk of packed pairs: the host-side runtime of the hot path.
// It issues the same kernels in the same order as the Python schedule vrdone_b200/engine.py (Engine.backbone, which stays
// the readable specification and the path the CPU-emulated tests run), but without a Python / ctypes round trip per launch:
// ~100 launches cost ~0.3 ms of host time instead of ~2.5 ms, which is what lets host-resident inputs be pipelined in small
// chunks (see vrdone_b200/maskvrd.py).  Data flow follows the reference forward: models/backbones.py:154-248 / 323-436,
// models/blocks.py:1070-1080, models/local_transformer.py:807-835, models/fpns.py:229-257.
//
// Weights are looked up by the names of engine.PackedWeights (a name -> device pointer table handed over at creation).
// Activations live in a caller-provided workspace, carved by a stack allocator (mark / release mirrors the Python scopes);
// vrd_backbone_workspace_bytes() runs the same schedule without launching to size it.
#include <cstdio>
#include <cstring>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>
#include "../../include/vrdone_b200.h"
#include "kernels.h"

namespace {

struct Mat {
    char* p; long long ld; int rows, cols, dt;
    int esize() const { return dt == VRD_BF16 ? 2 : 4; }
    Mat cols_from(int c0, int n) const { Mat m = *this; m.p = p + (long long)c0 * esize(); m.cols = n; return m; }
    Mat rows_from(long long r0, int n) const { Mat m = *this; m.p = p + r0 * ld * esize(); m.rows = n; return m; }
};

struct Weight { const void* p; int rows, cols; };

struct Arena {
    char* base; size_t cap, top, peak; bool dry;
    Mat alloc(long long rows, int cols, int dt) {
        const size_t bytes = ((size_t)rows * cols * (dt == VRD_BF16 ? 2 : 4) + 1023) & ~(size_t)1023;
        Mat m{dry ? nullptr : base + top, cols, (int)rows, cols, dt};
        top += bytes;
        if (top > peak) peak = top;
        return m;
    }
};

}  // namespace

struct vrd_engine {
    vrd_model_cfg_t cfg;
    std::unordered_map<std::string, Weight> w;
    long long launches;
    char err[512];
};

namespace {

struct Run {
    vrd_engine* E; const vrd_level_t* L; cudaStream_t st; Arena* A; int adt; bool fail;

    bool dry() const { return A->dry; }
    Lay lay(int l) const {
        Lay x; x.row_seq = L[l].row_seq; x.seqinfo = reinterpret_cast<const int4*>(L[l].seqinfo); x.R = L[l].R; x.B = L[l].B;
        x.tiles = reinterpret_cast<const int4*>(L[l].attn_tiles); x.n_tiles = L[l].attn_tiles != nullptr ? L[l].n_attn_tiles : 0;
        return x;
    }
    void error(const char* what, const std::string& name) {
        if (!fail) snprintf(E->err, sizeof E->err, "%s: %s", what, name.c_str());
        fail = true;
    }
    const Weight* find(const std::string& name, bool required = true) {
        auto it = E->w.find(name);
        if (it == E->w.end()) { if (required) error("missing weight", name); return nullptr; }
        return &it->second;
    }
    const float* F(const std::string& name, bool required = true) { const Weight* x = find(name, required); return x ? (const float*)x->p : nullptr; }

    void check(const char* what) {
        if (dry() || fail) return;
        ++E->launches;
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) { snprintf(E->err, sizeof E->err, "%s: %s", what, cudaGetErrorString(e)); fail = true; }
    }

    // out = act(a * W^T + bias [+corr]) + res1 + res2 with the layout of level l (or none when l < 0)
    void gemm(const Mat& a, const std::string& wname, const Mat& out, int l, int taps = 1, int act = 0, const Mat* res1 = nullptr,
              const Mat* res2 = nullptr, const float* corr = nullptr, const float* ln_g = nullptr, const float* ln_b = nullptr,
              int ln_relu = 0, const Mat* ln_out = nullptr) {
        const Weight* W = find(wname + ".W");
        if (fail || dry()) return;
        vrd::GemmArgs g;
        g.A = a.p; g.lda = a.ld; g.W = W->p; g.bias = F(wname + ".b", false);
        g.out = out.p; g.out_dtype = out.dt; g.ldo = out.ld;
        g.M = a.rows; g.N = W->rows; g.K = a.cols; g.taps = taps; g.act = act;
        g.res1 = res1 ? (const float*)res1->p : nullptr; g.ldr1 = res1 ? res1->ld : 0;
        g.res2 = res2 ? (const float*)res2->p : nullptr; g.ldr2 = res2 ? res2->ld : 0;
        g.corr = corr;
        g.ln_gamma = ln_g; g.ln_beta = ln_b; g.ln_relu = ln_relu;
        if (ln_out != nullptr) { g.ln_out = ln_out->p; g.ld_ln = ln_out->ld; }
        if (l >= 0) { g.row_seq = L[l].row_seq; g.seqinfo = reinterpret_cast<const int4*>(L[l].seqinfo); g.R = L[l].R; }
        else { g.row_seq = nullptr; g.seqinfo = nullptr; g.R = 0; }
        if (W->cols != taps * a.cols || out.rows != a.rows || out.cols != W->rows) { error("gemm shape mismatch", wname); return; }
        int rc = a.dt == VRD_BF16 ? vrd::gemm_tcgen05_bf16(g, st) : vrd::gemm_f32(g, st);
        if (rc != 0) { error((a.dt == VRD_BF16 || rc == 2) ? vrd::gemm_tcgen05_error() : "gemm_simt: unsupported shape", wname); return; }
        check("gemm");
    }

    void layernorm(const Mat& x, const std::string& p, const Mat& out, bool relu, int l) {
        const float* g = F(p + ".g"); const float* b = F(p + ".be");
        if (fail || dry()) return;
        if (vrd::layernorm(x.p, x.dt, x.ld, g, b, out.p, out.dt, out.ld, x.rows, x.cols, relu ? 1 : 0, L[l].row_seq, L[l].R, st))
            error("layernorm: unsupported shape", p);
        check("layernorm");
    }

    struct Branch { const char* name; bool use_pre; };
    // q/k/v style pre-projections: dwconv_ln over x, then one GEMM per branch.  outs[i] receives the projected branch i.
    void attention_qkv(const std::string& p, const Mat& x, const char* pre, const Branch* br, int nb, int lin, int lout, int stride,
                       int streams, Mat* outs) {
        const int C = x.cols;
        const long long rows = (long long)streams * L[lout].R;
        vrd::DwBranches d;
        memset(&d, 0, sizeof d);
        d.n = nb;
        Mat pre_out[3];
        for (int i = 0; i < nb; ++i) {
            pre_out[i] = A->alloc(rows, C, adt);
            d.w[i] = F(p + "." + br[i].name + "_conv.w");
            d.use_pre[i] = br[i].use_pre ? 1 : 0;
            d.g[i] = F(p + "." + br[i].name + "_norm.g");
            d.b[i] = F(p + "." + br[i].name + "_norm.be");
            d.out[i] = pre_out[i].p;
            d.ldo[i] = pre_out[i].ld;
        }
        const float* pg = pre ? F(std::string(pre) + ".g") : nullptr;
        const float* pb = pre ? F(std::string(pre) + ".be") : nullptr;
        if (!fail && !dry()) {
            if (vrd::dwconv_ln(x.p, x.dt, x.ld, lay(lin), lay(lout), stride, pg, pb, d, adt, C, streams, st)) error("dwconv_ln: unsupported", p);
            check("dwconv_ln");
        }
        for (int i = 0; i < nb; ++i) {
            outs[i] = A->alloc(rows, C, adt);
            gemm(pre_out[i], p + "." + br[i].name, outs[i], lout);
        }
    }

    // pre-LN block: windowed attention + (identity | max-pool) skip + MLP  (blocks.py:1070-1080); `out` is fp32 [streams*R_out, C]
    void encoder_block(const Mat& x, const std::string& p, int lin, int lout, int stride, int streams, int n_head, int win, const Mat& out) {
        const int C = x.cols;
        const long long rows = (long long)streams * L[lout].R;
        const size_t mark = A->top;
        const std::string pre = p + ".ln1";
        const Branch br[3] = {{"query", true}, {"key", true}, {"value", true}};
        Mat qkv[3];
        attention_qkv(p + ".attn", x, pre.c_str(), br, 3, lin, lout, stride, streams, qkv);
        Mat a = A->alloc(rows, C, adt);
        if (!fail && !dry()) {
            if (vrd::window_attn(qkv[0].p, qkv[1].p, qkv[2].p, a.p, adt, a.ld, lay(lout), n_head, C, win / 2, streams, st)) error("window_attn: unsupported", p);
            check("window_attn");
        }
        Mat skip = x;
        if (stride != 1) {
            skip = A->alloc(rows, C, VRD_F32);
            if (!fail && !dry()) {
                if (vrd::maxpool_skip((const float*)x.p, x.ld, lay(lin), lay(lout), (float*)skip.p, skip.ld, C, st)) error("maxpool_skip: unsupported", p);
                check("maxpool_skip");
            }
        }
        Mat y = A->alloc(rows, C, VRD_F32);
        Mat h = A->alloc(rows, C, adt);
        // bf16 path: projection + residual + the LayerNorm that feeds the MLP as ONE launch
        bool fused = false;
        if (adt == VRD_BF16 && C == 512 && !fail && !dry()) {
            vrd::GemmArgs t;
            t.M = (int)rows; t.N = C; t.K = C; t.taps = 1; t.act = 0; t.res1 = (const float*)skip.p; t.res2 = nullptr; t.corr = nullptr;
            fused = vrd::gemm_res_ln_fused_ok(t);
        }
        if (fused) gemm(a, p + ".attn.proj", y, lout, 1, 0, &skip, nullptr, nullptr, F(p + ".ln2.g"), F(p + ".ln2.be"), 0, &h);
        else {
            gemm(a, p + ".attn.proj", y, lout, 1, 0, &skip);
            layernorm(y, p + ".ln2", h, false, lout);
        }
        Mat h2 = A->alloc(rows, 4 * C, adt);
        gemm(h, p + ".mlp.0", h2, lout, 1, VRD_ACT_GELU);
        gemm(h2, p + ".mlp.3", out, lout, 1, 0, &y);
        A->top = mark;
    }

    void attend(const Mat& q, const Mat& k, const Mat& v, const Mat& a, int n_head, int window) {
        if (fail || dry()) return;
        const int C = q.cols;
        if (window <= 0) {
            if (vrd::full_attn(q.p, k.p, v.p, a.p, adt, a.ld, lay(0), n_head, C, L[0].max_len, st)) error("full_attn: unsupported", "");
            check("full_attn");
        } else {
            if (vrd::window_attn(q.p, k.p, v.p, a.p, adt, a.ld, lay(0), n_head, C, window / 2, 1, st)) error("window_attn: unsupported", "");
            check("window_attn");
        }
    }

    // out = tgt + decoder_layer(tgt, mem) (local_transformer.py:807-835 with the extra residual of backbones.py:217-221)
    void sos_layer(const Mat& tgt, const Mat& mem, const std::string& p, const Mat& out, int n_head, int window) {
        const int C = tgt.cols;
        const long long R = L[0].R;
        const size_t mark = A->top;
        {
            const std::string pre = p + ".ln1";
            const Branch br[3] = {{"query", true}, {"key", true}, {"value", false}};
            Mat qkv[3];
            attention_qkv(p + ".self_attn", tgt, pre.c_str(), br, 3, 0, 0, 1, 1, qkv);
            Mat a = A->alloc(R, C, adt);
            attend(qkv[0], qkv[1], qkv[2], a, n_head, window);
            Mat tgt1 = A->alloc(R, C, VRD_F32);
            gemm(a, p + ".self_attn.proj", tgt1, 0, 1, 0, &tgt);
            const std::string pre2 = p + ".ln2";
            const Branch bq[1] = {{"query", true}};
            const Branch bkv[2] = {{"key", false}, {"value", false}};
            Mat q[1], kv[2];
            attention_qkv(p + ".multihead_attn", tgt1, pre2.c_str(), bq, 1, 0, 0, 1, 1, q);
            attention_qkv(p + ".multihead_attn", mem, nullptr, bkv, 2, 0, 0, 1, 1, kv);
            Mat a2 = A->alloc(R, C, adt);
            attend(q[0], kv[0], kv[1], a2, n_head, window);
            gemm(a2, p + ".multihead_attn.proj", out, 0, 1, 0, &tgt1, &tgt);
        }
        A->top = mark;
    }

    // 2 x (k=3 conv-as-GEMM -> LN -> ReLU) on stacked s/o rows; result written into `out` (a column slice)
    void embed(Mat x, const std::string& conv, const std::string& norm, const Mat& out) {
        const int C = E->cfg.embd_dim, n_conv = E->cfg.n_conv;
        const long long rows = 2LL * L[0].R;
        const size_t mark = A->top;
        const bool fuse = adt == VRD_BF16 && C == 512 && vrd_options().embed_ln != 0;   // LayerNorm + ReLU as the GEMM's epilogue
        for (int i = 0; i < n_conv; ++i) {
            const std::string cn = "backbone." + conv + "." + std::to_string(i);
            const std::string nn = "backbone." + norm + "." + std::to_string(i);
            if (fuse) {
                Mat nx = (i == n_conv - 1) ? out : A->alloc(rows, C, adt);
                gemm(x, cn, nx, 0, 3, 0, nullptr, nullptr, F(cn + ".corr", false), F(nn + ".g"), F(nn + ".be"), 1);
                x = nx;
                continue;
            }
            Mat e = A->alloc(rows, C, VRD_F32);
            gemm(x, cn, e, 0, 3, 0, nullptr, nullptr, F(cn + ".corr", false));
            Mat nx = (i == n_conv - 1) ? out : A->alloc(rows, C, adt);
            layernorm(e, nn, nx, true, 0);
            x = nx;
        }
        A->top = mark;
    }

    void small_conv(const Mat& x, int cin, const std::string& p, const char* norm, bool relu, const Mat& out) {
        const float* w = F(p + ".w"); const float* b = F(p + ".b");
        const float* g = norm ? F(std::string(norm) + ".g") : nullptr;
        const float* be = norm ? F(std::string(norm) + ".be") : nullptr;
        const Weight* W = find(p + ".w");
        if (fail || dry()) return;
        if (vrd::small_conv((const float*)x.p, cin, w, b, g, be, relu ? 1 : 0, out.p, out.dt, out.ld, x.rows, W->cols, L[0].row_seq, L[0].R, st))
            error("small_conv: unsupported", p);
        check("small_conv");
    }

    // workspace prefix shared by vrd_backbone_pack and vrd_backbone_compute
    void pack_buffers(Mat& vis, Mat& clp, Mat& bso, Mat& bent) {
        const vrd_model_cfg_t& c = E->cfg;
        const long long R0 = L[0].R;
        vis = A->alloc(2 * R0, c.visual_dim, adt);
        clp = c.clip_dim > 0 ? A->alloc(2 * R0, c.clip_dim, adt) : Mat{nullptr, 0, 0, 0, adt};
        bso = A->alloc(R0, 8, VRD_F32);
        bent = A->alloc(2 * R0, 8, VRD_F32);
    }

    void compute(float* e_top_out, float* mf_out) {
        const vrd_model_cfg_t& c = E->cfg;
        const int C = c.embd_dim, Fd = c.fpn_dim, n_lev = c.n_branch + 1;
        const long long R0 = L[0].R;
        const std::string bb = "backbone";
        Mat vis, clp, bso, bent;
        pack_buffers(vis, clp, bso, bent);

        // 2. embedding convs, entity-box embedding, fuse MLPs (s rows then o rows; weights are shared)
        Mat P = A->alloc(2 * R0, C, VRD_F32);           // x / xn of the stem loop
        Mat EB = A->alloc(2 * R0, C, VRD_F32);          // encoder-block output
        {
            const size_t mark = A->top;
            Mat cat = A->alloc(2 * R0, 2 * C, adt);
            if (c.clip_dim > 0) {
                Mat vcat = A->alloc(2 * R0, 2 * C, adt);
                embed(vis, "visual_embd", "visual_embd_norm", vcat.cols_from(0, C));
                embed(clp, "clip_embd", "clip_embd_norm", vcat.cols_from(C, C));
                Mat h = A->alloc(2 * R0, C, adt);
                gemm(vcat, bb + ".visual_clip_fuse.0", h, 0, 1, VRD_ACT_GELU);
                gemm(h, bb + ".visual_clip_fuse.1", cat.cols_from(0, C), 0);
            } else {
                embed(vis, "visual_embd", "visual_embd_norm", cat.cols_from(0, C));
            }
            small_conv(bent, c.bbox_entity_dim, bb + ".bbox_entity_embd", "backbone.bbox_entity_norm", true, cat.cols_from(C, C));
            Mat h = A->alloc(2 * R0, C, adt);
            gemm(cat, bb + ".visual_bbox_fuse.0", h, 0, 1, VRD_ACT_GELU);
            gemm(h, bb + ".visual_bbox_fuse.1", P, 0);
            A->top = mark;
        }
        // 3. stem blocks (shared weights, both streams at once) interleaved with subject-object synergy layers
        const int sos_window = c.use_local ? c.win : 0;
        for (int i = 0; i < c.n_stem; ++i) {
            const std::string is = std::to_string(i);
            encoder_block(P, bb + ".stem." + is, 0, 0, 1, 2, c.n_head, c.win, EB);
            Mat s = EB.rows_from(0, (int)R0), o = EB.rows_from(R0, (int)R0);
            sos_layer(s, o, bb + ".s_attn." + is, P.rows_from(0, (int)R0), c.fuse_head, sos_window);
            sos_layer(o, s, bb + ".o_attn." + is, P.rows_from(R0, (int)R0), c.fuse_head, sos_window);
        }
        // 4. fuse the two streams and the relative-box embedding into one
        std::vector<Mat> e(n_lev);
        e[0] = A->alloc(R0, C, VRD_F32);
        {
            const size_t mark = A->top;
            Mat cat2 = A->alloc(R0, 2 * C, adt);
            layernorm(P.rows_from(0, (int)R0), bb + ".s_fuse_norm", cat2.cols_from(0, C), false, 0);
            layernorm(P.rows_from(R0, (int)R0), bb + ".o_fuse_norm", cat2.cols_from(C, C), false, 0);
            Mat h = A->alloc(R0, C, adt);
            gemm(cat2, bb + ".so_fuse.0", h, 0, 1, VRD_ACT_GELU);
            Mat cat3 = A->alloc(R0, 2 * C, adt);
            gemm(h, bb + ".so_fuse.1", cat3.cols_from(0, C), 0);
            small_conv(bso, c.bbox_so_dim, bb + ".bbox_so_embd", nullptr, false, cat3.cols_from(C, C));
            Mat h2 = A->alloc(R0, C, adt);
            gemm(cat3, bb + ".so_visual_bbox_fuse.0", h2, 0, 1, VRD_ACT_GELU);
            gemm(h2, bb + ".so_visual_bbox_fuse.1", e[0], 0);
            A->top = mark;
        }
        // 5. stride-2 pyramid; the coarsest level goes straight to the caller's buffer
        for (int i = 0; i < c.n_branch; ++i) {
            const long long Ri = L[i + 1].R;
            e[i + 1] = (i + 1 == n_lev - 1) ? Mat{(char*)e_top_out, C, (int)Ri, C, VRD_F32} : A->alloc(Ri, C, VRD_F32);
            encoder_block(e[i], bb + ".branch." + std::to_string(i), i, i + 1, 2, 1, c.n_head, c.win, e[i + 1]);
        }
        // 6. top-down FPN -> mask features at full temporal resolution
        const int top = c.n_branch;
        Mat y = A->alloc(L[top].R, Fd, VRD_F32);
        {
            const std::string t = std::to_string(top);
            const float* pg = F("neck.input_norms." + t + ".g"); const float* pb = F("neck.input_norms." + t + ".be");
            const float* w = F("neck.fpn_convs." + t + ".w");
            const float* g = F("neck.fpn_norms." + t + ".g"); const float* b = F("neck.fpn_norms." + t + ".be");
            if (!fail && !dry()) { vrd::fpn_top((const float*)e[top].p, e[top].ld, lay(top), pg, pb, w, g, b, (float*)y.p, y.ld, st); check("fpn_top"); }
        }
        for (int l = top - 1; l >= 0; --l) {
            const std::string ls = std::to_string(l), us = std::to_string(l + 1);
            Mat n = A->alloc(L[l].R, C, adt);
            layernorm(e[l], "neck.input_norms." + ls, n, false, l);
            Mat cur = A->alloc(L[l].R, Fd, VRD_F32);
            gemm(n, "neck.lateral_convs." + ls, cur, l);
            Mat yl = A->alloc(L[l].R, Fd, VRD_F32);
            const float* lg = F("neck.lateral_norms." + ls + ".g"); const float* lb = F("neck.lateral_norms." + ls + ".be");
            const float* bu = F("neck.fpn_norms." + us + ".be"); const float* w = F("neck.fpn_convs." + ls + ".w");
            const float* g = F("neck.fpn_norms." + ls + ".g"); const float* b = F("neck.fpn_norms." + ls + ".be");
            if (!fail && !dry()) {
                vrd::fpn_level((const float*)cur.p, cur.ld, (const float*)y.p, y.ld, lay(l), lay(l + 1), lg, lb, bu, w, g, b, (float*)yl.p, yl.ld, st);
                check("fpn_level");
            }
            y = yl;
        }
        {
            const float* beta = F("neck.fpn_norms.0.be"); const float* w = F("neck.mask_features.w"); const float* b = F("neck.mask_features.b");
            if (!fail && !dry()) { vrd::mask_features((const float*)y.p, y.ld, lay(0), beta, w, b, mf_out, Fd, st); check("mask_features"); }
        }
    }

    // Query decoder + heads (predictor.py:85-115, local_transformer.py:773-835 / 875-976): L[0] = level 0 (mask features),
    // L[1] = the coarsest level.  Same kernels and order as engine.Engine._predictor.
    void query_ln(const Mat& x, const char* ln, const float* pos, int Q, int nrows, const float* dw, const char* ln2, const Mat& out,
                  const std::string& p) {
        const float* g = ln ? F(p + ln + ".g") : nullptr; const float* b = ln ? F(p + ln + ".be") : nullptr;
        const float* g2 = ln2 ? F(p + ln2 + ".g") : nullptr; const float* b2 = ln2 ? F(p + ln2 + ".be") : nullptr;
        if (fail || dry()) return;
        if (vrd::query_ln((const float*)x.p, x.ld, g, b, pos, Q, nrows, out.rows, dw, g2, b2, out.p, out.dt, out.ld, x.cols, st))
            error("query_ln: unsupported shape", p);
        check("query_ln");
    }

    void predict(const vrd_predictor_cfg_t& pc, const float* e_top, int e_top_cols, const float* mf, int mf_cols, int topk, float* logits,
                 float* scores, int* ids, int* first_last, float* masks) {
        const int D = pc.n_embd, Q = pc.num_queries, nh = pc.n_head;
        const int B = L[0].B;
        const long long Rt = L[1].R, R0 = L[0].R;
        const long long MQ = ((long long)B * Q + 127) / 128 * 128;
        const Mat top{(char*)e_top, e_top_cols, (int)Rt, e_top_cols, VRD_F32};
        Mat n = A->alloc(Rt, e_top_cols, adt);
        layernorm(top, "predictor.input_norm", n, false, 1);
        Mat src = A->alloc(Rt, D, VRD_F32);
        gemm(n, "predictor.input_proj", src, 1);
        Mat tgt = A->alloc(MQ, D, VRD_F32), tgt_next = A->alloc(MQ, D, VRD_F32);
        if (!fail && !dry()) cudaMemsetAsync(tgt.p, 0, (size_t)MQ * D * 4, st);
        const float* pos = F("predictor.query_pos");
        for (int j = 0; j < pc.num_layers; ++j) {
            const size_t mark = A->top;
            const std::string p = "predictor.transformer.decoder.layers." + std::to_string(j);
            const std::string sa = p + ".self_attn", ca = p + ".multihead_attn";
            // self-attention among the Q queries of each pair: q = k = LN1(tgt) + pos, v = tgt
            Mat qk = A->alloc(MQ, D, adt), tv = A->alloc(MQ, D, adt);
            query_ln(tgt, ".ln1", pos, Q, B * Q, nullptr, nullptr, qk, p);
            query_ln(tgt, nullptr, nullptr, Q, B * Q, nullptr, nullptr, tv, p);
            Mat q = A->alloc(MQ, D, adt), k = A->alloc(MQ, D, adt), v = A->alloc(MQ, D, adt), a = A->alloc(MQ, D, adt);
            gemm(qk, sa + ".query", q, -1);
            gemm(qk, sa + ".key", k, -1);
            gemm(tv, sa + ".value", v, -1);
            if (!fail && !dry()) {
                if (vrd::query_self_attn(q.p, k.p, v.p, a.p, adt, a.ld, B, Q, nh, D, st)) error("query_self_attn: unsupported", p);
                check("query_self_attn");
            }
            Mat tgt1 = A->alloc(MQ, D, VRD_F32);
            gemm(a, sa + ".proj", tgt1, -1, 1, 0, &tgt);
            // cross-attention to the coarsest pyramid level
            const Branch bkv[2] = {{"key", false}, {"value", false}};
            Mat kv[2];
            attention_qkv(ca, src, nullptr, bkv, 2, 1, 1, 1, 1, kv);
            Mat h = A->alloc(MQ, D, adt);
            query_ln(tgt1, ".ln2", pos, Q, B * Q, F(ca + ".query_conv.w"), ".multihead_attn.query_norm", h, p);
            Mat q2 = A->alloc(MQ, D, adt), a2 = A->alloc(MQ, D, adt);
            gemm(h, ca + ".query", q2, -1);
            if (!fail && !dry()) {
                if (vrd::query_cross_attn(q2.p, kv[0].p, kv[1].p, a2.p, adt, a2.ld, lay(1), Q, nh, D, st)) error("query_cross_attn: unsupported", p);
                check("query_cross_attn");
            }
            Mat tgt2 = A->alloc(MQ, D, VRD_F32);
            gemm(a2, ca + ".proj", tgt2, -1, 1, 0, &tgt1);
            // FFN
            Mat h3 = A->alloc(MQ, D, adt), h4 = A->alloc(MQ, pc.n_hidden, adt);
            query_ln(tgt2, ".ln3", nullptr, Q, B * Q, nullptr, nullptr, h3, p);
            gemm(h3, p + ".mlp.0", h4, -1, 1, VRD_ACT_GELU);
            gemm(h4, p + ".mlp.3", tgt_next, -1, 1, 0, &tgt2);
            std::swap(tgt, tgt_next);
            A->top = mark;
        }
        Mat hs = A->alloc(MQ, D, adt);
        query_ln(tgt, ".norm", nullptr, Q, B * Q, nullptr, nullptr, hs, "predictor.transformer.decoder");
        const Mat lg{(char*)logits, pc.n_cls_pad, (int)MQ, pc.n_cls_pad, VRD_F32};
        gemm(hs, "predictor.class_embed", lg, -1);
        Mat m0 = A->alloc(MQ, D, adt), m1 = A->alloc(MQ, D, adt), me = A->alloc(MQ, D, VRD_F32);
        gemm(hs, "predictor.mask_embed.0", m0, -1, 1, VRD_ACT_GELU);
        gemm(m0, "predictor.mask_embed.1", m1, -1, 1, VRD_ACT_GELU);
        gemm(m1, "predictor.mask_embed.2", me, -1);
        if (!fail && !dry()) {
            if (vrd::mask_logits((const float*)me.p, me.ld, mf, mf_cols, lay(0), Q, masks, Q, first_last, st)) error("mask_logits: unsupported", "");
            check("mask_logits");
            if (vrd::softmax_topk(logits, pc.n_cls_pad, B * Q, pc.n_cls, topk, scores, ids, st)) error("softmax_topk: unsupported", "");
            check("softmax_topk");
        }
        (void)R0;
    }
};

thread_local char t_eng_err[512] = {0};

}  // namespace

extern "C" {

const char* vrd_engine_last_error(void) { return t_eng_err; }

int vrd_engine_create(const vrd_model_cfg_t* cfg, const char* const* names, const void* const* ptrs, const int32_t* rows,
                      const int32_t* cols, int n, vrd_engine_t** out) {
    if (cfg == nullptr || out == nullptr || (n > 0 && (names == nullptr || ptrs == nullptr || rows == nullptr || cols == nullptr))) {
        snprintf(t_eng_err, sizeof t_eng_err, "vrd_engine_create: bad arguments");
        return 1;
    }
    if (cfg->n_branch + 1 > VRD_MAX_LEVELS || cfg->n_branch < 1) {
        snprintf(t_eng_err, sizeof t_eng_err, "vrd_engine_create: 1..%d pyramid branches supported", VRD_MAX_LEVELS - 1);
        return 1;
    }
    vrd_engine* e = new vrd_engine();
    e->cfg = *cfg;
    e->launches = 0;
    e->err[0] = 0;
    for (int i = 0; i < n; ++i) e->w[names[i]] = Weight{ptrs[i], rows[i], cols[i]};
    *out = e;
    return 0;
}

void vrd_engine_destroy(vrd_engine_t* e) { delete e; }
int64_t vrd_engine_launches(const vrd_engine_t* e) { return e->launches; }

int64_t vrd_backbone_workspace_bytes(vrd_engine_t* e, const vrd_level_t* levels) {
    Arena a{nullptr, 0, 0, 0, true};
    Run r{e, levels, nullptr, &a, e->cfg.act_dtype, false};
    r.compute(nullptr, nullptr);
    if (r.fail) { snprintf(t_eng_err, sizeof t_eng_err, "%s", e->err); return -1; }
    return (int64_t)a.peak;
}

int vrd_backbone_pack(vrd_engine_t* e, const vrd_level_t* levels, const void* pair_ptrs, const int64_t* pair_strides, int token_major,
                      void* workspace, int64_t workspace_bytes, vrd_stream_t stream) {
    Arena a{(char*)workspace, (size_t)workspace_bytes, 0, 0, false};
    Run r{e, levels, (cudaStream_t)stream, &a, e->cfg.act_dtype, false};
    Mat vis, clp, bso, bent;
    r.pack_buffers(vis, clp, bso, bent);
    if (a.peak > a.cap) { snprintf(t_eng_err, sizeof t_eng_err, "vrd_backbone_pack: workspace too small"); return 1; }
    const vrd_model_cfg_t& c = e->cfg;
    vrd::pack_pairs(pair_ptrs, (const long long*)pair_strides, r.lay(0), c.visual_dim, c.clip_dim, c.bbox_so_dim, c.bbox_entity_dim, vis.p,
                    clp.p, e->cfg.act_dtype, (float*)bso.p, (float*)bent.p, token_major, (cudaStream_t)stream);
    r.check("pack_pairs");
    if (r.fail) { snprintf(t_eng_err, sizeof t_eng_err, "%s", e->err); return 1; }
    return 0;
}

int vrd_backbone_pack_tracklets(vrd_engine_t* e, const vrd_level_t* levels, const float* vis_all, const float* clip_all,
                                const float* boxes_all, const int32_t* pair_tab, float video_w, float video_h, void* workspace,
                                int64_t workspace_bytes, vrd_stream_t stream) {
    Arena a{(char*)workspace, (size_t)workspace_bytes, 0, 0, false};
    Run r{e, levels, (cudaStream_t)stream, &a, e->cfg.act_dtype, false};
    Mat vis, clp, bso, bent;
    r.pack_buffers(vis, clp, bso, bent);
    if (a.peak > a.cap) { snprintf(t_eng_err, sizeof t_eng_err, "vrd_backbone_pack_tracklets: workspace too small"); return 1; }
    const vrd_model_cfg_t& c = e->cfg;
    if (c.bbox_so_dim != 5 || c.bbox_entity_dim != 8) {
        snprintf(t_eng_err, sizeof t_eng_err, "vrd_backbone_pack_tracklets: geometry features are 5 + 8 + 8 channels");
        return 1;
    }
    if (c.clip_dim > 0 && clip_all == nullptr) { snprintf(t_eng_err, sizeof t_eng_err, "vrd_backbone_pack_tracklets: CLIP features missing"); return 1; }
    if (vrd::pack_tracklets(vis_all, clip_all, boxes_all, pair_tab, r.lay(0), c.visual_dim, c.clip_dim, video_w, video_h, vis.p, clp.p,
                            e->cfg.act_dtype, (float*)bso.p, (float*)bent.p, (cudaStream_t)stream)) {
        snprintf(t_eng_err, sizeof t_eng_err, "vrd_backbone_pack_tracklets: feature dims must be multiples of 4");
        return 1;
    }
    r.check("pack_tracklets");
    if (r.fail) { snprintf(t_eng_err, sizeof t_eng_err, "%s", e->err); return 1; }
    return 0;
}

int vrd_backbone_compute(vrd_engine_t* e, const vrd_level_t* levels, void* workspace, int64_t workspace_bytes, float* e_top,
                         float* mask_feat, vrd_stream_t stream) {
    {   // size check with the same schedule
        Arena d{nullptr, 0, 0, 0, true};
        Run rd{e, levels, nullptr, &d, e->cfg.act_dtype, false};
        rd.compute(nullptr, nullptr);
        if (rd.fail) { snprintf(t_eng_err, sizeof t_eng_err, "%s", e->err); return 1; }
        if ((int64_t)d.peak > workspace_bytes) {
            snprintf(t_eng_err, sizeof t_eng_err, "vrd_backbone_compute: workspace of %lld bytes needed, %lld given", (long long)d.peak,
                     (long long)workspace_bytes);
            return 1;
        }
    }
    Arena a{(char*)workspace, (size_t)workspace_bytes, 0, 0, false};
    Run r{e, levels, (cudaStream_t)stream, &a, e->cfg.act_dtype, false};
    r.compute(e_top, mask_feat);
    if (r.fail) { snprintf(t_eng_err, sizeof t_eng_err, "%s", e->err); return 1; }
    return 0;
}

int64_t vrd_predict_workspace_bytes(vrd_engine_t* e, const vrd_predictor_cfg_t* pc, const vrd_level_t* level0, const vrd_level_t* level_top) {
    const vrd_level_t lv[2] = {*level0, *level_top};
    Arena a{nullptr, 0, 0, 0, true};
    Run r{e, lv, nullptr, &a, e->cfg.act_dtype, false};
    r.predict(*pc, nullptr, e->cfg.embd_dim, nullptr, e->cfg.fpn_dim, 1, nullptr, nullptr, nullptr, nullptr, nullptr);
    if (r.fail) { snprintf(t_eng_err, sizeof t_eng_err, "%s", e->err); return -1; }
    return (int64_t)a.peak;
}

int vrd_predict(vrd_engine_t* e, const vrd_predictor_cfg_t* pc, const vrd_level_t* level0, const vrd_level_t* level_top, const float* e_top,
                const float* mask_feat, int topk, void* workspace, int64_t workspace_bytes, float* logits, float* topk_scores,
                int32_t* topk_ids, int32_t* first_last, float* masks, vrd_stream_t stream) {
    if (pc == nullptr || level0 == nullptr || level_top == nullptr || e_top == nullptr || mask_feat == nullptr || logits == nullptr ||
        topk_scores == nullptr || topk_ids == nullptr || first_last == nullptr) {
        snprintf(t_eng_err, sizeof t_eng_err, "vrd_predict: null argument");
        return 1;
    }
    if (pc->n_embd != e->cfg.fpn_dim) { snprintf(t_eng_err, sizeof t_eng_err, "vrd_predict: predictor width must equal fpn_dim"); return 1; }
    const int64_t need = vrd_predict_workspace_bytes(e, pc, level0, level_top);
    if (need < 0) return 1;
    if (need > workspace_bytes) {
        snprintf(t_eng_err, sizeof t_eng_err, "vrd_predict: workspace of %lld bytes needed, %lld given", (long long)need, (long long)workspace_bytes);
        return 1;
    }
    const vrd_level_t lv[2] = {*level0, *level_top};
    Arena a{(char*)workspace, (size_t)workspace_bytes, 0, 0, false};
    Run r{e, lv, (cudaStream_t)stream, &a, e->cfg.act_dtype, false};
    r.predict(*pc, e_top, e->cfg.embd_dim, mask_feat, e->cfg.fpn_dim, topk, logits, topk_scores, topk_ids, first_last, masks);
    if (r.fail) { snprintf(t_eng_err, sizeof t_eng_err, "%s", e->err); return 1; }
    return 0;
}

}  // extern "C"
