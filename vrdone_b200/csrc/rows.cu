// Memory-bound row kernels: one warp owns one token row of C = 128*NCH channels, lanes own 16-byte chunks so
// every warp load/store instruction touches one contiguous 512 B (fp32) / 256 B (bf16) span.
// Reference semantics (file:line under /root/reference/models/): channel LayerNorm blocks.py:143-158,
// masked depthwise conv blocks.py:91-113 + 706-724, max-pool skip blocks.py:1040-1046 + 1074,
// FPN fpns.py:229-257, input contract maskvrd.py:363-414 (padding, re-derived analytically per SURVEY appendix B).
#include <type_traits>
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace vrd {

constexpr int WARPS = 8;   // warps per block for the row kernels

// ------------------------------------------------------------------------------------------------------------
// pack_pairs: ragged (C, L) fp32 pair tensors with arbitrary (channel, time) strides -> token-major operand rows
// ------------------------------------------------------------------------------------------------------------
// One block owns 32 rows and walks over 64-channel tiles (blockIdx.y strides the channel tiles, so the per-row source
// lookup is done once and a block moves ~64 KB instead of 4 KB).  Rows whose pair tensor is contiguous along time (dense
// (C, L): the data loader's transposed view after pickling / pin_memory) are read with lanes along time and transposed in
// the tile; rows of token-major tensors (the (L, C) buffer the reference builds, vidor.py:708-711) are read with lanes
// along channels.  Either way the global reads are coalesced 128-byte segments and the operand rows are written as packed
// channel pairs (bf16x2 / float2).  nv and nc must be even.
template <typename TA> struct Pair2;
template <> struct Pair2<float> { static __device__ __forceinline__ void st(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); } };
template <> struct Pair2<__nv_bfloat16> {
    static __device__ __forceinline__ void st(__nv_bfloat16* p, float a, float b) { *reinterpret_cast<__nv_bfloat162*>(p) = __floats2bfloat162_rn(a, b); }
};

template <typename TA>
__global__ void __launch_bounds__(256) pack_pairs_kernel(const float* const* __restrict__ ptrs, const long long* __restrict__ strides,
                                                         Lay lay, int nv, int nc, int nbs, int nbe, TA* __restrict__ vis,
                                                         TA* __restrict__ clip, float* __restrict__ bso, float* __restrict__ bent) {
    __shared__ float tile[32][65];
    __shared__ const float* s_src[32];     // element (t, channel 0) of the row's pair, or nullptr for separator rows
    __shared__ long long s_sc[32];
    __shared__ int s_tmode[32];            // 1: time-contiguous source (read lanes along time)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
    const int r0 = blockIdx.x * 32;
    const long long R = lay.R;
    const int c0 = 2 * nv + 2 * nc;
    const int C = c0 + nbs + 2 * nbe;
    if (threadIdx.x < 32) {
        const int r = r0 + threadIdx.x;
        const int seq = (r < lay.R) ? lay.row_seq[r] : -1;
        if (seq >= 0) {
            const int4 si = lay.seqinfo[seq];
            const long long sc = strides[2 * seq], st = strides[2 * seq + 1];
            s_src[threadIdx.x] = ptrs[seq] + (long long)(r - si.x) * st;
            s_sc[threadIdx.x] = sc;
            s_tmode[threadIdx.x] = (st == 1 && sc != 1) ? 1 : 0;
        } else {
            s_src[threadIdx.x] = nullptr;
            s_sc[threadIdx.x] = 0;
            s_tmode[threadIdx.x] = 0;
        }
    }
    __syncthreads();
    for (int cb = blockIdx.y * 64; cb < C; cb += gridDim.y * 64) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            // lanes along time: row = tx, channel = ty + 8k
            const int c = cb + ty + 8 * k;
            if (s_tmode[tx] && c < C) tile[tx][ty + 8 * k] = __ldg(s_src[tx] + (long long)c * s_sc[tx]);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            // lanes along channels: row = ty + 8k, channels = tx and tx + 32
            const int rr = ty + 8 * k;
            if (!s_tmode[rr] && s_src[rr] != nullptr) {
                if (cb + tx < C) tile[rr][tx] = __ldg(s_src[rr] + (long long)(cb + tx) * s_sc[rr]);
                if (cb + 32 + tx < C) tile[rr][tx + 32] = __ldg(s_src[rr] + (long long)(cb + 32 + tx) * s_sc[rr]);
            }
        }
        __syncthreads();
        const int c = cb + 2 * tx;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int rr = ty + 8 * k;
            const long long r = r0 + rr;
            if (r >= lay.R) continue;
            const bool live = s_src[rr] != nullptr;
            const float v0 = (live && c < C) ? tile[rr][2 * tx] : 0.f;
            const float v1 = (live && c + 1 < C) ? tile[rr][2 * tx + 1] : 0.f;
            if (c + 1 < nv) Pair2<TA>::st(vis + r * nv + c, v0, v1);
            else if (c + 1 < 2 * nv) Pair2<TA>::st(vis + (R + r) * nv + (c - nv), v0, v1);
            else if (c + 1 < 2 * nv + nc) Pair2<TA>::st(clip + r * nc + (c - 2 * nv), v0, v1);
            else if (c + 1 < c0) Pair2<TA>::st(clip + (R + r) * nc + (c - 2 * nv - nc), v0, v1);
            else {
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int cc = c + u;
                    const float v = u == 0 ? v0 : v1;
                    if (cc < c0 + nbs) bso[r * 8 + (cc - c0)] = v;
                    else if (cc < c0 + nbs + nbe) bent[r * 8 + (cc - c0 - nbs)] = v;
                    else if (cc < C) bent[(R + r) * 8 + (cc - c0 - nbs - nbe)] = v;
                }
            }
        }
        __syncthreads();
    }
    // unused tail columns of the 8-wide geometry rows (written once, by the blocks of the first channel slice)
    if (blockIdx.y == 0 && threadIdx.x < 32) {
        const long long r = r0 + threadIdx.x;
        if (r < lay.R) {
            for (int j = nbs; j < 8; ++j) bso[r * 8 + j] = 0.f;
            for (int j = nbe; j < 8; ++j) { bent[r * 8 + j] = 0.f; bent[(R + r) * 8 + j] = 0.f; }
        }
    }
}

// Every source token-major (channel stride 1; the (L, C) buffer the reference's loader builds): one warp per row, lanes walk
// the row's channel pairs, no shared memory, 8 independent loads in flight per lane.
template <typename TA>
__global__ void __launch_bounds__(256) pack_rows_kernel(const float* const* __restrict__ ptrs, const long long* __restrict__ strides,
                                                        Lay lay, int nv, int nc, int nbs, int nbe, TA* __restrict__ vis,
                                                        TA* __restrict__ clip, float* __restrict__ bso, float* __restrict__ bent) {
    const int lane = threadIdx.x & 31;
    const long long r = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (r >= lay.R) return;
    const long long R = lay.R;
    const int c0 = 2 * nv + 2 * nc;
    const int C = c0 + nbs + 2 * nbe;
    const int seq = lay.row_seq[r];
    const float* src = nullptr;
    if (seq >= 0) {
        const int4 si = lay.seqinfo[seq];
        src = ptrs[seq] + (long long)(r - si.x) * strides[2 * seq + 1];
    }
    auto put = [&](int c, float v0, float v1) {
        if (c + 1 < nv) Pair2<TA>::st(vis + r * nv + c, v0, v1);
        else if (c + 1 < 2 * nv) Pair2<TA>::st(vis + (R + r) * nv + (c - nv), v0, v1);
        else if (c + 1 < 2 * nv + nc) Pair2<TA>::st(clip + r * nc + (c - 2 * nv), v0, v1);
        else if (c + 1 < c0) Pair2<TA>::st(clip + (R + r) * nc + (c - 2 * nv - nc), v0, v1);
        else {
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int cc = c + u;
                const float v = u == 0 ? v0 : v1;
                if (cc < c0 + nbs) bso[r * 8 + (cc - c0)] = v;
                else if (cc < c0 + nbs + nbe) bent[r * 8 + (cc - c0 - nbs)] = v;
                else if (cc < C) bent[(R + r) * 8 + (cc - c0 - nbs - nbe)] = v;
            }
        }
    };
    constexpr int UNROLL = 4;
    for (int cb = 0; cb < C; cb += 64 * UNROLL) {
        float v0[UNROLL], v1[UNROLL];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int c = cb + 64 * u + 2 * lane;
            v0[u] = (src != nullptr && c < C) ? __ldg(src + c) : 0.f;
            v1[u] = (src != nullptr && c + 1 < C) ? __ldg(src + c + 1) : 0.f;
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const int c = cb + 64 * u + 2 * lane;
            if (c < C) put(c, v0[u], v1[u]);
        }
    }
    if (lane == 0) {
        for (int j = nbs; j < 8; ++j) bso[r * 8 + j] = 0.f;
        for (int j = nbe; j < 8; ++j) { bent[r * 8 + j] = 0.f; bent[(R + r) * 8 + j] = 0.f; }
    }
}

void pack_pairs(const void* ptrs, const long long* strides, Lay lay, int nv, int nc, int nbs, int nbe, void* vis,
                void* clip, int adt, float* bso, float* bent, int token_major, cudaStream_t st) {
    if (token_major) {
        const int grid = (lay.R + WARPS - 1) / WARPS;
        if (adt == VRD_BF16)
            pack_rows_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const float* const*)ptrs, strides, lay, nv, nc, nbs, nbe,
                                                                  (__nv_bfloat16*)vis, (__nv_bfloat16*)clip, bso, bent);
        else
            pack_rows_kernel<float><<<grid, 256, 0, st>>>((const float* const*)ptrs, strides, lay, nv, nc, nbs, nbe,
                                                          (float*)vis, (float*)clip, bso, bent);
        return;
    }
    const int C = 2 * nv + 2 * nc + nbs + 2 * nbe;
    const int ctiles = (C + 63) / 64;
    const dim3 grid((lay.R + 31) / 32, (ctiles + 7) / 8);      // ~8 channel tiles per block
    if (adt == VRD_BF16)
        pack_pairs_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const float* const*)ptrs, strides, lay, nv, nc, nbs, nbe,
                                                               (__nv_bfloat16*)vis, (__nv_bfloat16*)clip, bso, bent);
    else
        pack_pairs_kernel<float><<<grid, 256, 0, st>>>((const float* const*)ptrs, strides, lay, nv, nc, nbs, nbe,
                                                       (float*)vis, (float*)clip, bso, bent);
}

// ------------------------------------------------------------------------------------------------------------
// pack_tracklets: the pair enumeration + feature gather of the reference DATA LOADER on the device (SURVEY 8f row 1;
// dataloaders/vidor.py:659-711, utils/misc.py:158-217).  Per-tracklet features are uploaded once; a pair (s, o) is
// described by the rows of its first sub-sampled frame in the concatenated tracklet arrays and the sub-sampling stride.
// One warp per packed row: the lanes copy the two visual (and CLIP) rows, lanes 0..2 compute the 5-d relative and the two
// 8-d entity box features with the reference's formulas and operation order (true divisions, logf).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void box_cwh(const float4 b, float& cx, float& cy, float& w, float& h) {
    cx = (b.z + b.x) / 2; cy = (b.w + b.y) / 2; w = b.z - b.x; h = b.w - b.y;
}
// 8-d entity feature of frame t (sub-sampled sequence of length L): [cx, dcx, cy, dcy, w, dw, h, dh] of the box normalised by
// the video size; d[t] = v[t] - v[t-1], d[0] extrapolated as d[1] - (d[2] - d[1]) (or d[1] when L == 2).
__device__ __forceinline__ void entity_feature(const float4* __restrict__ boxes, long long row0, int stride, int t, int L, float vw,
                                               float vh, float* out) {
    auto norm = [&](int tt, float (&v)[4]) {
        float4 b = boxes[row0 + (long long)tt * stride];
        b.x = b.x / vw; b.z = b.z / vw; b.y = b.y / vh; b.w = b.w / vh;
        box_cwh(b, v[0], v[1], v[2], v[3]);
    };
    float cur[4];
    norm(t, cur);
    float d[4];
    if (t >= 1) {
        float prev[4];
        norm(t - 1, prev);
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = cur[i] - prev[i];
    } else {
        float v1[4];
        norm(1, v1);
#pragma unroll
        for (int i = 0; i < 4; ++i) d[i] = v1[i] - cur[i];
        if (L > 2) {
            float v2[4];
            norm(2, v2);
#pragma unroll
            for (int i = 0; i < 4; ++i) { const float d2 = v2[i] - v1[i]; d[i] = d[i] - (d2 - d[i]); }
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) { out[2 * i] = cur[i]; out[2 * i + 1] = d[i]; }
}

template <typename TA>
__global__ void __launch_bounds__(256) pack_tracklets_kernel(const float* __restrict__ vis_all, const float* __restrict__ clip_all,
                                                             const float4* __restrict__ boxes_all, const int4* __restrict__ pair_tab,
                                                             Lay lay, int nv, int nc, float vw, float vh, TA* __restrict__ vis,
                                                             TA* __restrict__ clip, float* __restrict__ bso, float* __restrict__ bent) {
    const int lane = threadIdx.x & 31;
    const long long r = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (r >= lay.R) return;
    const long long R = lay.R;
    const int seq = lay.row_seq[r];
    if (seq < 0) {
        const float z[4] = {0.f, 0.f, 0.f, 0.f};
        for (int c = lane * 4; c < nv; c += 128) { st4(vis + r * nv + c, z); st4(vis + (R + r) * nv + c, z); }
        for (int c = lane * 4; c < nc; c += 128) { st4(clip + r * nc + c, z); st4(clip + (R + r) * nc + c, z); }
        if (lane < 8) { bso[r * 8 + lane] = 0.f; bent[r * 8 + lane] = 0.f; bent[(R + r) * 8 + lane] = 0.f; }
        return;
    }
    const int4 si = lay.seqinfo[seq];
    const int4 pt = pair_tab[seq];                   // (subject row of frame 0, object row of frame 0, stride, -)
    const int t = (int)r - si.x, L = si.y, stride = pt.z;
    const long long fs = (long long)pt.x + (long long)t * stride, fo = (long long)pt.y + (long long)t * stride;
    for (int c = lane * 4; c < nv; c += 128) {
        float a[4], b[4];
        ld4(vis_all + fs * nv + c, a);
        ld4(vis_all + fo * nv + c, b);
        st4(vis + r * nv + c, a);
        st4(vis + (R + r) * nv + c, b);
    }
    for (int c = lane * 4; c < nc; c += 128) {
        float a[4], b[4];
        ld4(clip_all + fs * nc + c, a);
        ld4(clip_all + fo * nc + c, b);
        st4(clip + r * nc + c, a);
        st4(clip + (R + r) * nc + c, b);
    }
    if (lane == 0) {
        float scx, scy, sw, sh, ocx, ocy, ow, oh;
        box_cwh(boxes_all[fs], scx, scy, sw, sh);
        box_cwh(boxes_all[fo], ocx, ocy, ow, oh);
        float* o = bso + r * 8;
        o[0] = (scx - ocx) / ocx;
        o[1] = (scy - ocy) / ocy;
        o[2] = logf(sw / ow);
        o[3] = logf(sh / oh);
        o[4] = logf((sw * sh) / (ow * oh));
        o[5] = o[6] = o[7] = 0.f;
    } else if (lane == 1) {
        entity_feature(boxes_all, pt.x, stride, t, L, vw, vh, bent + r * 8);
    } else if (lane == 2) {
        entity_feature(boxes_all, pt.y, stride, t, L, vw, vh, bent + (R + r) * 8);
    }
}

int pack_tracklets(const float* vis_all, const float* clip_all, const float* boxes_all, const int* pair_tab, Lay lay, int nv, int nc,
                   float vw, float vh, void* vis, void* clip, int adt, float* bso, float* bent, cudaStream_t st) {
    if ((nv & 3) || (nc & 3)) return 1;
    const int grid = (lay.R + WARPS - 1) / WARPS;
    if (adt == VRD_BF16)
        pack_tracklets_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(vis_all, clip_all, (const float4*)boxes_all, (const int4*)pair_tab, lay, nv,
                                                                   nc, vw, vh, (__nv_bfloat16*)vis, (__nv_bfloat16*)clip, bso, bent);
    else
        pack_tracklets_kernel<float><<<grid, 256, 0, st>>>(vis_all, clip_all, (const float4*)boxes_all, (const int4*)pair_tab, lay, nv, nc,
                                                           vw, vh, (float*)vis, (float*)clip, bso, bent);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// layernorm (+ReLU): out = [relu](LN(x)), separator rows -> 0
// ------------------------------------------------------------------------------------------------------------
constexpr int LN_RPW = 2;   // rows per warp: both rows' loads are in flight before the first reduction starts (HBM-bound kernel)

template <typename TI, typename TO, int NCH>
__global__ void layernorm_kernel(const TI* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                                 const float* __restrict__ beta, TO* __restrict__ out, long long ldo, int rows, int relu,
                                 const int* __restrict__ row_seq, int R) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int row0 = (blockIdx.x * WARPS + (threadIdx.x >> 5)) * LN_RPW;
    if (row0 >= rows) return;
    float v[LN_RPW][NCH][4];
    bool live[LN_RPW];
#pragma unroll
    for (int u = 0; u < LN_RPW; ++u) {
        const int row = row0 + u;
        live[u] = row < rows && !(row_seq != nullptr && row_seq[row % R] < 0);
        if (live[u]) load_row<TI, NCH>(x + (long long)row * ldx, lane, v[u]);
    }
#pragma unroll
    for (int u = 0; u < LN_RPW; ++u) {
        const int row = row0 + u;
        if (row >= rows) break;
        TO* o = out + (long long)row * ldo;
        if (!live[u]) { zero_row<TO, NCH>(o, lane); continue; }      // warp-uniform
        row_normalize<NCH>(v[u], lane, gamma, beta);
        if (relu) {
#pragma unroll
            for (int j = 0; j < NCH; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) v[u][j][i] = fmaxf(v[u][j][i], 0.f);
        }
        store_row<TO, NCH>(o, lane, v[u]);
    }
}

template <typename TI, typename TO>
static int layernorm_t(const void* x, long long ldx, const float* g, const float* b, void* out, long long ldo, int rows,
                       int C, int relu, const int* row_seq, int R, cudaStream_t st) {
    const int grid = (rows + WARPS * LN_RPW - 1) / (WARPS * LN_RPW);
    if (C == 512)
        launch_k(layernorm_kernel<TI, TO, 4>, dim3(grid), dim3(WARPS * 32), 0, st, (const TI*)x, ldx, g, b, (TO*)out, ldo, rows, relu, row_seq, R);
    else if (C == 256)
        launch_k(layernorm_kernel<TI, TO, 2>, dim3(grid), dim3(WARPS * 32), 0, st, (const TI*)x, ldx, g, b, (TO*)out, ldo, rows, relu, row_seq, R);
    else
        return 1;
    return 0;
}

int layernorm(const void* x, int xdt, long long ldx, const float* g, const float* b, void* out, int odt, long long ldo,
              int rows, int C, int relu, const int* row_seq, int R, cudaStream_t st) {
    if (xdt == VRD_F32 && odt == VRD_F32) return layernorm_t<float, float>(x, ldx, g, b, out, ldo, rows, C, relu, row_seq, R, st);
    if (xdt == VRD_F32 && odt == VRD_BF16) return layernorm_t<float, __nv_bfloat16>(x, ldx, g, b, out, ldo, rows, C, relu, row_seq, R, st);
    if (xdt == VRD_BF16 && odt == VRD_BF16) return layernorm_t<__nv_bfloat16, __nv_bfloat16>(x, ldx, g, b, out, ldo, rows, C, relu, row_seq, R, st);
    return 1;
}

// ------------------------------------------------------------------------------------------------------------
// upload: device copy of a small pinned host array made by a KERNEL (the SMs read the host memory through UVA), not by the copy
// engine.  Used for per-video layout arrays on the compute stream: a cudaMemcpyAsync there queues behind whatever bulk
// host->device copies other streams have in flight (measured: the network of a resident video slowed from 18.5 to 32.7 ms while a
// side stream copied 1.1 GB, although neither GEMMs, LayerNorms nor empty launches are slowed by such copies).
// ------------------------------------------------------------------------------------------------------------
__global__ void upload_kernel(const int4* __restrict__ src, int4* __restrict__ dst, long long n16, const int* __restrict__ src4,
                              int* __restrict__ dst4, int tail4) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n16; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i];
    if (blockIdx.x == 0 && (int)threadIdx.x < tail4) dst4[threadIdx.x] = src4[threadIdx.x];
}

int upload(const void* host_src, void* dev_dst, long long bytes, cudaStream_t st) {
    if (bytes <= 0 || (bytes & 3) != 0 || ((uintptr_t)host_src & 15) != 0 || ((uintptr_t)dev_dst & 15) != 0) return 1;
    const long long n16 = bytes / 16;
    const int tail4 = (int)((bytes - n16 * 16) / 4);
    const int grid = (int)((n16 + 255) / 256 < 64 ? ((n16 + 255) / 256 > 0 ? (n16 + 255) / 256 : 1) : 64);
    upload_kernel<<<grid, 256, 0, st>>>((const int4*)host_src, (int4*)dev_dst, n16, (const int*)host_src + n16 * 4, (int*)dev_dst + n16 * 4, tail4);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// merge_layout: one level of several consecutive chunk layouts viewed as ONE batch (rows of chunk c follow those of chunk c - 1,
// pair ids are renumbered), built on the device from the chunk layouts that are already there.  The query decoder and the
// heads run once per video over this merged layout; building it on the host meant one more small host->device copy per video
// on the compute stream, which queues behind the bulk pair copies of the next video.
// ------------------------------------------------------------------------------------------------------------
constexpr int MERGE_MAX = 16;
struct MergeArgs {
    const int* rs[MERGE_MAX];
    const int4* si[MERGE_MAX];
    int row_base[MERGE_MAX + 1], pair_base[MERGE_MAX + 1];
    int n;
};

__global__ void merge_layout_kernel(MergeArgs a, int* __restrict__ rs_out, int4* __restrict__ si_out) {
    const int total_rows = a.row_base[a.n], total_pairs = a.pair_base[a.n];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total_rows; i += gridDim.x * blockDim.x) {
        int c = 0;
        while (c + 1 < a.n && i >= a.row_base[c + 1]) ++c;
        const int v = a.rs[c][i - a.row_base[c]];
        rs_out[i] = v >= 0 ? v + a.pair_base[c] : -1;
    }
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total_pairs; i += gridDim.x * blockDim.x) {
        int c = 0;
        while (c + 1 < a.n && i >= a.pair_base[c + 1]) ++c;
        int4 s = a.si[c][i - a.pair_base[c]];
        s.x += a.row_base[c];
        si_out[i] = s;
    }
}

// rs / si / R / B: HOST arrays (device pointers of the chunk layouts, their row and pair counts); n <= MERGE_MAX
int merge_layout(int n, const int* const* rs, const int* const* si, const int* R, const int* B, int* rs_out, int* si_out,
                 cudaStream_t st) {
    if (n < 1 || n > MERGE_MAX) return 1;
    MergeArgs a;
    a.n = n;
    a.row_base[0] = a.pair_base[0] = 0;
    for (int c = 0; c < n; ++c) {
        a.rs[c] = rs[c];
        a.si[c] = reinterpret_cast<const int4*>(si[c]);
        a.row_base[c + 1] = a.row_base[c] + R[c];
        a.pair_base[c + 1] = a.pair_base[c] + B[c];
    }
    const int work = a.row_base[n] > a.pair_base[n] ? a.row_base[n] : a.pair_base[n];
    int grid = (work + 255) / 256;
    grid = grid < 1 ? 1 : (grid > 296 ? 296 : grid);
    merge_layout_kernel<<<grid, 256, 0, st>>>(a, rs_out, reinterpret_cast<int4*>(si_out));
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// small_conv: k=3 conv with cin <= 8 input channels on 8-wide fp32 rows (+bias [+LN] [+ReLU]); N = 512 outputs
// wt is [3*cin, N] (tap-major rows) so that lanes read contiguous output channels.
// ------------------------------------------------------------------------------------------------------------
constexpr int SC_RPW = 4;   // consecutive rows per warp: every weight vector read from L1 serves four rows

template <typename TO, int NCH>
__global__ void small_conv_kernel(const float* __restrict__ x, int cin, const float* __restrict__ wt,
                                  const float* __restrict__ bias, const float* __restrict__ gamma,
                                  const float* __restrict__ beta, int relu, TO* __restrict__ out, long long ldo, int rows,
                                  const int* __restrict__ row_seq, int R) {
    pdl_wait();
    constexpr int N = NCH * 128;
    const int lane = threadIdx.x & 31;
    const int row0 = (blockIdx.x * WARPS + (threadIdx.x >> 5)) * SC_RPW;
    if (row0 >= rows) return;
    float v[SC_RPW][NCH][4];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        float bj[4];
        ld4(bias + (j * 32 + lane) * 4, bj);
#pragma unroll
        for (int u = 0; u < SC_RPW; ++u)
#pragma unroll
            for (int i = 0; i < 4; ++i) v[u][j][i] = bj[i];
    }
    for (int tap = 0; tap < 3; ++tap) {
        for (int c = 0; c < cin; ++c) {
            float xv[SC_RPW];
#pragma unroll
            for (int u = 0; u < SC_RPW; ++u) {
                const int rr = row0 + u + tap - 1;
                xv[u] = (rr >= 0 && rr < rows) ? __ldg(x + (long long)rr * 8 + c) : 0.f;
            }
            const float* w = wt + (long long)(tap * cin + c) * N;
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                float wv[4];
                ld4(w + (j * 32 + lane) * 4, wv);
#pragma unroll
                for (int u = 0; u < SC_RPW; ++u)
#pragma unroll
                    for (int i = 0; i < 4; ++i) v[u][j][i] = fmaf(xv[u], wv[i], v[u][j][i]);
            }
        }
    }
#pragma unroll
    for (int u = 0; u < SC_RPW; ++u) {
        const int row = row0 + u;
        if (row >= rows) break;
        TO* o = out + (long long)row * ldo;
        if (row_seq[row % R] < 0) { zero_row<TO, NCH>(o, lane); continue; }
        if (gamma != nullptr) row_normalize<NCH>(v[u], lane, gamma, beta);
        if (relu) {
#pragma unroll
            for (int j = 0; j < NCH; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) v[u][j][i] = fmaxf(v[u][j][i], 0.f);
        }
        store_row<TO, NCH>(o, lane, v[u]);
    }
}

int small_conv(const float* x, int cin, const float* wt, const float* bias, const float* g, const float* b, int relu,
               void* out, int odt, long long ldo, int rows, int N, const int* row_seq, int R, cudaStream_t st) {
    if (N != 512 || cin > 8) return 1;
    const int grid = (rows + WARPS * SC_RPW - 1) / (WARPS * SC_RPW);
    if (odt == VRD_BF16)
        launch_k(small_conv_kernel<__nv_bfloat16, 4>, dim3(grid), dim3(WARPS * 32), 0, st, x, cin, wt, bias, g, b, relu, (__nv_bfloat16*)out, ldo, rows, row_seq, R);
    else
        launch_k(small_conv_kernel<float, 4>, dim3(grid), dim3(WARPS * 32), 0, st, x, cin, wt, bias, g, b, relu, (float*)out, ldo, rows, row_seq, R);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// dwconv_ln: [LN_pre] -> depthwise k=3 (stride 1|2) -> LN, for up to 3 branches sharing one read of the input rows.
// A warp walks DW_RUN consecutive output rows and keeps the three input rows of the stencil in registers (raw and/or
// pre-normalised, as the compile-time branch mask needs), so every input row is loaded and normalised once per run.
// ------------------------------------------------------------------------------------------------------------
constexpr int DW_RUN = 16;
constexpr int DW_WARPS = 4;

template <int NCH>
__device__ __forceinline__ void row_copy(float (&d)[NCH][4], const float (&s)[NCH][4]) {
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) d[j][i] = s[j][i];
}
template <int NCH>
__device__ __forceinline__ void row_zero(float (&d)[NCH][4]) {
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) d[j][i] = 0.f;
}

// PREMASK bit b: branch b convolves LN_pre(x) (else raw x)
template <typename TI, typename TO, int NCH, int STRIDE, int NB, int PREMASK>
__global__ void __launch_bounds__(DW_WARPS * 32) dwconv_ln_kernel(const TI* __restrict__ x, long long ldx, Lay lin, Lay lout,
                                                                 const float* __restrict__ pre_g, const float* __restrict__ pre_b,
                                                                 DwBranches br, int streams) {
    pdl_wait();
    constexpr bool ANY_PRE = PREMASK != 0;
    constexpr bool ANY_RAW = PREMASK != ((1 << NB) - 1);
    constexpr int C = NCH * 128;
    const int lane = threadIdx.x & 31;
    const int runs_per_stream = (lout.R + DW_RUN - 1) / DW_RUN;
    const int run = blockIdx.x * DW_WARPS + (threadIdx.x >> 5);
    if (run >= streams * runs_per_stream) return;
    const int s = run / runs_per_stream;
    const int r_begin = (run - s * runs_per_stream) * DW_RUN;
    const int r_end = min(r_begin + DW_RUN, lout.R);

    // stencil window: index 0/1/2 = time t-1 / t / t+1 of the current output row
    float raw[ANY_RAW ? 3 : 1][NCH][4];
    float nrm[ANY_PRE ? 3 : 1][NCH][4];
    int cur_seq = -1, cur_t = 0;
    int4 so = make_int4(0, 0, 0, 0), si = make_int4(0, 0, 0, 0);
    const TI* base = x;

    // Software prefetch: the STRIDE input rows the NEXT output row will add to the window are requested before the arithmetic
    // of the current row starts, so their latency hides behind ~10^3 instructions instead of stalling the first use.
    float pre[STRIDE][NCH][4];
    int pre_seq = -1, pre_t = 0;               // pre[i] holds input time pre_t + i of pair pre_seq (if that time is < len)

    auto fetch = [&](int slot, int tt, int seq) {    // input time index tt of the current pair -> window slot `slot`
        const int len = si.y;
        if (tt >= 0 && tt < len) {
            float v[NCH][4];
            if (seq == pre_seq && tt == pre_t) row_copy<NCH>(v, pre[0]);
            else if (STRIDE == 2 && seq == pre_seq && tt == pre_t + 1) row_copy<NCH>(v, pre[STRIDE - 1]);
            else load_row<TI, NCH>(base + (long long)tt * ldx, lane, v);
            if constexpr (ANY_RAW) row_copy<NCH>(raw[slot], v);
            if constexpr (ANY_PRE) { row_normalize<NCH>(v, lane, pre_g, pre_b); row_copy<NCH>(nrm[slot], v); }
        } else {
            if constexpr (ANY_RAW) row_zero<NCH>(raw[slot]);
            if constexpr (ANY_PRE) {
                // the first pad column (tt == len) carries the LN bias if it exists; stride 2 only reads it when it exists
                if (tt == len && (si.z != 0 || STRIDE == 2)) {
#pragma unroll
                    for (int j = 0; j < NCH; ++j) ld4(pre_b + (j * 32 + lane) * 4, nrm[slot][j]);
                } else {
                    row_zero<NCH>(nrm[slot]);
                }
            }
        }
    };

    for (int r = r_begin; r < r_end; ++r) {
        const long long grow = (long long)s * lout.R + r;
        const int seq = lout.row_seq[r];
        if (seq < 0) {
#pragma unroll
            for (int b = 0; b < NB; ++b) zero_row<TO, NCH>((TO*)br.out[b] + grow * br.ldo[b], lane);
            cur_seq = -1;
            continue;
        }
        if (seq != cur_seq) {
            so = lout.seqinfo[seq];
            si = lin.seqinfo[seq];
            base = x + ((long long)s * lin.R + si.x) * ldx;
        }
        const int t = (r - so.x) * STRIDE;
        if (seq == cur_seq && t == cur_t + STRIDE) {          // slide the window
            if constexpr (STRIDE == 1) {
                if constexpr (ANY_RAW) { row_copy<NCH>(raw[0], raw[1]); row_copy<NCH>(raw[1], raw[2]); }
                if constexpr (ANY_PRE) { row_copy<NCH>(nrm[0], nrm[1]); row_copy<NCH>(nrm[1], nrm[2]); }
                fetch(2, t + 1, seq);
            } else {
                if constexpr (ANY_RAW) row_copy<NCH>(raw[0], raw[2]);
                if constexpr (ANY_PRE) row_copy<NCH>(nrm[0], nrm[2]);
                fetch(1, t, seq);
                fetch(2, t + 1, seq);
            }
        } else {
            fetch(0, t - 1, seq);
            fetch(1, t, seq);
            fetch(2, t + 1, seq);
        }
        cur_seq = seq;
        cur_t = t;
        if (r + 1 < r_end && (r + 1 - so.x) < so.y) {         // the next output row belongs to the same pair
            pre_seq = seq;
            pre_t = t + 2;
#pragma unroll
            for (int i = 0; i < STRIDE; ++i)
                if (pre_t + i < si.y) load_row<TI, NCH>(base + (long long)(pre_t + i) * ldx, lane, pre[i]);
        } else {
            pre_seq = -1;
        }
        float y[NB][NCH][4];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const float* w = br.w[b];   // [3, C] tap-major
            constexpr int dummy = 0; (void)dummy;
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                float w0[4], w1[4], w2[4];
                const int c = (j * 32 + lane) * 4;
                ld4(w + c, w0); ld4(w + C + c, w1); ld4(w + 2 * C + c, w2);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    float a0, a1, a2;
                    if ((PREMASK >> b) & 1) { a0 = nrm[0][j][i]; a1 = nrm[ANY_PRE ? 1 : 0][j][i]; a2 = nrm[ANY_PRE ? 2 : 0][j][i]; }
                    else { a0 = raw[0][j][i]; a1 = raw[ANY_RAW ? 1 : 0][j][i]; a2 = raw[ANY_RAW ? 2 : 0][j][i]; }
                    y[b][j][i] = a0 * w0[i] + a1 * w1[i] + a2 * w2[i];
                }
            }
        }
        // post-conv LayerNorm of all branches, reductions interleaved for instruction-level parallelism
        float mean[NB], rstd[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            float sm = 0.f;
#pragma unroll
            for (int j = 0; j < NCH; ++j) sm += (y[b][j][0] + y[b][j][1]) + (y[b][j][2] + y[b][j][3]);
            mean[b] = sm;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int b = 0; b < NB; ++b) mean[b] += __shfl_xor_sync(FULL_MASK, mean[b], o);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            mean[b] *= (1.0f / C);
            float q = 0.f;
#pragma unroll
            for (int j = 0; j < NCH; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) { const float d = y[b][j][i] - mean[b]; q += d * d; }
            rstd[b] = q;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1)
#pragma unroll
            for (int b = 0; b < NB; ++b) rstd[b] += __shfl_xor_sync(FULL_MASK, rstd[b], o);
#pragma unroll
        for (int b = 0; b < NB; ++b) {
            const float rs = 1.0f / sqrtf(rstd[b] * (1.0f / C) + VRD_EPS);
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                float g[4], be[4];
                ld4(br.g[b] + (j * 32 + lane) * 4, g);
                ld4(br.b[b] + (j * 32 + lane) * 4, be);
#pragma unroll
                for (int i = 0; i < 4; ++i) y[b][j][i] = (y[b][j][i] - mean[b]) * rs * g[i] + be[i];
            }
            store_row<TO, NCH>((TO*)br.out[b] + grow * br.ldo[b], lane, y[b]);
        }
    }
}

// ------------------------------------------------------------------------------------------------------------
// dwconv_ln_tile: the stride-1, C = 512, fp32-input case (every level-0 call: encoder q/k/v, SOS self and cross), shared-
// memory tiled and persistent.  One CTA per SM (16 warps) walks over tiles of DWT_TILE consecutive rows of the stacked
// streams; one thread stages the DWT_TILE + 2 input rows of the NEXT tile with a single TMA bulk copy (cp.async.bulk +
// mbarrier, ~70 KB in flight per SM, no register staging) while all warps work on the current one.  Per tile: the warps
// first normalise every staged row once (in place when every branch reads LN_pre(x); statistics only when a branch also
// needs the raw row), then each warp produces output rows from the three staged neighbours, so no state is carried
// between rows.  A staged separator row normalises to exactly beta, which is the value of the first pad column.
// ------------------------------------------------------------------------------------------------------------
// TILE output rows per CTA iteration (TILE + 2 staged rows).  Each tile is a chain of dependent phases (TMA wait, normalise,
// barrier, per-branch conv + two shuffle reductions): with ONE tile in flight per SM the chain's latency is exposed whatever the
// number of warps working on it.  TILE = 16 halves the shared memory so that two independent CTAs share an SM and their phases
// interleave.
// NW warps per CTA, DWT_TILE / NW consecutive output rows per warp (parameter loads are shared between a warp's rows).  8 warps x 4
// rows left the SM with two warps per scheduler (issue active 31 %: every shuffle / shared-memory latency exposed); 16 warps x 2
// rows doubles the warps that can cover for each other at the price of ~30 % more L1 parameter traffic.
// two TMA buffers of raw rows [+ one tile of normalised rows when a branch needs both] + a zero row + a beta row + barriers
constexpr int dwt_smem_bytes(bool mixed, int tile) { return (mixed ? 3 : 2) * (tile + 2) * 512 * 4 + 2 * 512 * 4 + 32; }

template <typename TO, int NB, int PREMASK, int NW, int TILE>
__global__ void __launch_bounds__(NW * 32, 32 / TILE) dwconv_ln_tile_kernel(const float* __restrict__ x, Lay lay,
                                                                           const float* __restrict__ pre_g,
                                                                           const float* __restrict__ pre_b, DwBranches br,
                                                                           int total_rows) {
    pdl_wait();
    constexpr int NCH = 4, C = 512;
    constexpr int DWT_TILE = TILE, DWT_ROWS = TILE + 2, DWT_WARPS = NW, DWT_RPW = TILE / NW;
    constexpr bool ANY_PRE = PREMASK != 0;
    constexpr bool ALL_PRE = PREMASK == ((1 << NB) - 1);         // no branch reads the raw row: normalise in place
    constexpr bool MIXED = ANY_PRE && !ALL_PRE;                  // normalised rows go to their own tile
    extern __shared__ __align__(128) uint8_t dwt_smem[];
    float* sx_all = reinterpret_cast<float*>(dwt_smem);          // [2][DWT_ROWS][C]; row i <-> physical row r0 - 1 + i
    float* s_nrm = sx_all + 2 * DWT_ROWS * C;                    // [DWT_ROWS][C] (MIXED only)
    float* s_zero = s_nrm + (MIXED ? DWT_ROWS * C : 0);          // [C] zeros: taps outside the pair
    float* s_beta = s_zero + C;                                  // [C] LN_pre bias: the first pad column
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_beta + C);    // [2]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int n_tiles = total_rows / DWT_TILE;
    const uint32_t bar_u = (uint32_t)__cvta_generic_to_shared(bars);
    auto issue = [&](int tile, int buf) {                         // one thread: stage the rows of `tile` into buffer `buf`
        const int r0 = tile * DWT_TILE;
        const int lo = max(r0 - 1, 0), hi = min(r0 + DWT_TILE + 1, total_rows);
        const uint32_t bytes = (uint32_t)(hi - lo) * C * 4;
        const uint32_t b = bar_u + 8 * buf;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(sx_all + buf * DWT_ROWS * C + (lo - (r0 - 1)) * C)),
                       "l"(x + (long long)lo * C), "r"(bytes), "r"(b) : "memory");
    };
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if ((int)blockIdx.x < n_tiles) issue(blockIdx.x, 0);
    }
    for (int i = threadIdx.x; i < C; i += DWT_WARPS * 32) {
        s_zero[i] = 0.f;
        s_beta[i] = ANY_PRE ? pre_b[i] : 0.f;
    }
    __syncthreads();                                             // barriers initialised before anyone polls them
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        float* sx = sx_all + buf * DWT_ROWS * C;
        const float* sn = MIXED ? s_nrm : sx;                    // where the normalised rows live
        const int r0 = tile * DWT_TILE;
        const int lo = max(r0 - 1, 0), hi = min(r0 + DWT_TILE + 1, total_rows);
        if (threadIdx.x == 0 && tile + (int)gridDim.x < n_tiles) issue(tile + gridDim.x, buf ^ 1);   // freed by the trailing sync
        {
            const uint32_t parity = (it >> 1) & 1;
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(bar_u + 8 * buf), "r"(parity) : "memory");
            }
        }
        if (ANY_PRE) {
            // normalise every staged row once.  A warp works on its rows of a round together (loads, the two shuffle reductions
            // and the parameter reads are interleaved, not one row after the other: the serial version spent 22-40 % of the
            // kernel here on exposed latency).  DWT_ROWS is not a multiple of the warp count: the full rounds run on every warp,
            // the remainder only on the warps that have a row in it (a warp-uniform branch, not predicated-off work).
            constexpr int NFULL = DWT_ROWS / DWT_WARPS, NREM = DWT_ROWS % DWT_WARPS;
            auto normalise = [&](auto nr_tag, int first) {
                constexpr int NR = decltype(nr_tag)::value;
                f2 v[NR][NCH][2];
                bool on[NR];
                const uint32_t sx_base = (uint32_t)__cvta_generic_to_shared(sx) + lane * 16;
                const uint32_t dst_base = (uint32_t)__cvta_generic_to_shared(MIXED ? s_nrm : sx) + lane * 16;
#pragma unroll
                for (int u = 0; u < NR; ++u) {
                    const int i = first + warp + u * DWT_WARPS;
                    const int p = r0 - 1 + i;
                    on[u] = p >= lo && p < hi;
#pragma unroll
                    for (int j = 0; j < NCH; ++j) {
                        if (on[u]) f2_lds(sx_base + (uint32_t)i * C * 4 + j * 512, v[u][j]);
                        else v[u][j][0] = v[u][j][1] = 0ull;
                    }
                }
                float mean[NR], rstd[NR];
#pragma unroll
                for (int u = 0; u < NR; ++u) {
                    f2 sm = f2_add(v[u][0][0], v[u][0][1]);
#pragma unroll
                    for (int j = 1; j < NCH; ++j) sm = f2_add(sm, f2_add(v[u][j][0], v[u][j][1]));
                    mean[u] = f2_hsum(sm);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int u = 0; u < NR; ++u) mean[u] += __shfl_xor_sync(FULL_MASK, mean[u], o);
#pragma unroll
                for (int u = 0; u < NR; ++u) {
                    mean[u] *= (1.0f / C);
                    const f2 mm = f2_splat(mean[u]);
                    f2 q = 0ull;
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
#pragma unroll
                        for (int i = 0; i < 2; ++i) { v[u][j][i] = f2_sub(v[u][j][i], mm); q = f2_fma(v[u][j][i], v[u][j][i], q); }
                    rstd[u] = f2_hsum(q);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int u = 0; u < NR; ++u) rstd[u] += __shfl_xor_sync(FULL_MASK, rstd[u], o);
#pragma unroll
                for (int u = 0; u < NR; ++u) rstd[u] = rsqrtf(rstd[u] * (1.0f / C) + VRD_EPS);
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    f2 g[2], be[2];
                    f2_ldg(pre_g + (j * 32 + lane) * 4, g);
                    f2_ldg(pre_b + (j * 32 + lane) * 4, be);
#pragma unroll
                    for (int u = 0; u < NR; ++u) {
                        const f2 rs = f2_splat(rstd[u]);
#pragma unroll
                        for (int i = 0; i < 2; ++i) v[u][j][i] = f2_fma(f2_mul(v[u][j][i], rs), g[i], be[i]);
                    }
                }
#pragma unroll
                for (int u = 0; u < NR; ++u) {
                    const int i = first + warp + u * DWT_WARPS;
                    if (on[u]) {
#pragma unroll
                        for (int j = 0; j < NCH; ++j) f2_sts(dst_base + (uint32_t)i * C * 4 + j * 512, v[u][j]);
                    }
                }
            };
            normalise(std::integral_constant<int, NFULL>{}, 0);
            if (NREM > 0 && warp < NREM) normalise(std::integral_constant<int, 1>{}, NFULL * DWT_WARPS);
            __syncthreads();
        }
        // Each warp owns DWT_RPW consecutive output rows and reads the DWT_RPW + 2 staged rows around them ONCE per branch and
        // channel chunk (a register window shared by the rows' taps): the shared-memory pipe, not the FMA pipe, bounds this
        // kernel (per tile ~4600 of its ~6900 LSU cycles were the 3 x DWT_RPW tap reads per branch).  The per-channel parameters
        // (conv taps, LayerNorm gamma / beta) are re-read from L1 per use and shared between the warp's rows.  A window row
        // that is a separator reads the zero row; the one case where a separator carries a value -- the first pad column of a
        // pair (LN_pre bias) seen by the tap +1 of the pair's last row -- is added afterwards under a warp-uniform branch.
        {
            constexpr int NWIN = DWT_RPW + 2;
            const int rr0 = warp * DWT_RPW;
            uint32_t araw[NWIN], anrm[NWIN];
            bool live[DWT_RPW];
            uint32_t hib = 0;                                    // bit u: tap +1 of output row u is a first pad column
            const uint32_t sx_u = (uint32_t)__cvta_generic_to_shared(sx), sn_u = (uint32_t)__cvta_generic_to_shared(sn);
            const uint32_t zero_u = (uint32_t)__cvta_generic_to_shared(s_zero), beta_u = (uint32_t)__cvta_generic_to_shared(s_beta);
            int wseq[NWIN];
#pragma unroll
            for (int w = 0; w < NWIN; ++w) {
                const int grow = r0 + rr0 - 1 + w;
                int seq = -1;
                if (grow >= 0 && grow < total_rows) seq = lay.row_seq[grow % lay.R];
                wseq[w] = seq;
                const uint32_t row_off = (uint32_t)(rr0 + w) * C * 4;
                araw[w] = (seq >= 0 ? sx_u + row_off : zero_u) + lane * 16;
                anrm[w] = (seq >= 0 ? sn_u + row_off : zero_u) + lane * 16;
            }
#pragma unroll
            for (int u = 0; u < DWT_RPW; ++u) {
                live[u] = wseq[u + 1] >= 0;
                if (ANY_PRE && live[u] && wseq[u + 2] < 0 && lay.seqinfo[wseq[u + 1]].z != 0) hib |= 1u << u;
            }
#pragma unroll 1
            for (int b = 0; b < NB; ++b) {
                const bool pre = (PREMASK >> b) & 1;
                const float* wb = br.w[b];
                f2 y[DWT_RPW][NCH][2];
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const int c = (j * 32 + lane) * 4;
                    f2 w0[2], w1[2], w2[2];
                    f2_ldg(wb + c, w0); f2_ldg(wb + C + c, w1); f2_ldg(wb + 2 * C + c, w2);
                    f2 a[NWIN][2];
#pragma unroll
                    for (int w = 0; w < NWIN; ++w) f2_lds((pre ? anrm[w] : araw[w]) + j * 512, a[w]);
#pragma unroll
                    for (int u = 0; u < DWT_RPW; ++u)
#pragma unroll
                        for (int i = 0; i < 2; ++i) y[u][j][i] = f2_fma(a[u + 2][i], w2[i], f2_fma(a[u + 1][i], w1[i], f2_mul(a[u][i], w0[i])));
                    if (pre && hib != 0) {                       // warp-uniform and rare (one row per padded pair)
                        f2 bt[2];
                        f2_lds(beta_u + lane * 16 + j * 512, bt);
#pragma unroll
                        for (int u = 0; u < DWT_RPW; ++u)
                            if ((hib >> u) & 1) {
#pragma unroll
                                for (int i = 0; i < 2; ++i) y[u][j][i] = f2_fma(bt[i], w2[i], y[u][j][i]);   // the tap read 0 above
                            }
                    }
                }
                // post-conv LayerNorm of the warp's rows, reductions interleaved for instruction-level parallelism
                float mean[DWT_RPW], rstd[DWT_RPW];
#pragma unroll
                for (int u = 0; u < DWT_RPW; ++u) {
                    f2 sm = f2_add(y[u][0][0], y[u][0][1]);
#pragma unroll
                    for (int j = 1; j < NCH; ++j) sm = f2_add(sm, f2_add(y[u][j][0], y[u][j][1]));
                    mean[u] = f2_hsum(sm);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int u = 0; u < DWT_RPW; ++u) mean[u] += __shfl_xor_sync(FULL_MASK, mean[u], o);
#pragma unroll
                for (int u = 0; u < DWT_RPW; ++u) {
                    mean[u] *= (1.0f / C);
                    const f2 mm = f2_splat(mean[u]);
                    f2 q = 0ull;
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
#pragma unroll
                        for (int i = 0; i < 2; ++i) { y[u][j][i] = f2_sub(y[u][j][i], mm); q = f2_fma(y[u][j][i], y[u][j][i], q); }
                    rstd[u] = f2_hsum(q);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int u = 0; u < DWT_RPW; ++u) rstd[u] += __shfl_xor_sync(FULL_MASK, rstd[u], o);
#pragma unroll
                for (int u = 0; u < DWT_RPW; ++u) rstd[u] = rsqrtf(rstd[u] * (1.0f / C) + VRD_EPS);
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    f2 g[2], be[2];
                    f2_ldg(br.g[b] + (j * 32 + lane) * 4, g);
                    f2_ldg(br.b[b] + (j * 32 + lane) * 4, be);
#pragma unroll
                    for (int u = 0; u < DWT_RPW; ++u) {
                        const f2 rs = f2_splat(rstd[u]);
#pragma unroll
                        for (int i = 0; i < 2; ++i) y[u][j][i] = f2_fma(f2_mul(y[u][j][i], rs), g[i], be[i]);
                    }
                }
#pragma unroll
                for (int u = 0; u < DWT_RPW; ++u) {
                    TO* o = (TO*)br.out[b] + (long long)(r0 + rr0 + u) * br.ldo[b];
                    if (live[u]) {                                          // warp-uniform
#pragma unroll
                        for (int j = 0; j < NCH; ++j) f2_stg(o + (j * 32 + lane) * 4, y[u][j]);
                    } else {
                        zero_row<TO, NCH>(o, lane);                         // separator rows -> 0
                    }
                }
            }
        }
        if constexpr (ALL_PRE) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes before the next TMA fill
        __syncthreads();                                         // this buffer may be refilled from now on
    }
}

// ------------------------------------------------------------------------------------------------------------
// dwconv_ln_bw ("branch warps"): same staging as dwconv_ln_tile, different division of labour.  ncu on dwconv_ln_tile: the
// shared-memory / L1 data pipe is 63-72 % busy while issue slots (35 %) and DRAM (28-39 %) idle, and MORE THAN HALF of its
// wavefronts are the per-channel parameters (three conv taps, LayerNorm gamma / beta: 10 KB per branch) that every warp
// re-reads from L1 for every two rows.  Here a warp belongs to ONE branch for the lifetime of the persistent CTA and keeps that
// branch's 5 x 16 parameters per lane in registers: parameter traffic disappears, what is left per output row is the window
// (32 wavefronts per branch) and the one-off pre-LayerNorm pass.  WPB warps per branch; a warp walks over row pairs of the tile.
// One CTA per SM (the parameters cost 80 registers), 32-row tiles.
// ------------------------------------------------------------------------------------------------------------
template <typename TO, int NB, int PREMASK, int WPB, int TILE>
__global__ void __launch_bounds__(NB * WPB * 32, 1) dwconv_ln_bw_kernel(const float* __restrict__ x, Lay lay,
                                                                         const float* __restrict__ pre_g,
                                                                         const float* __restrict__ pre_b, DwBranches br,
                                                                         int total_rows) {
    pdl_wait();
    constexpr int NCH = 4, C = 512;
    constexpr int DWT_ROWS = TILE + 2, NW = NB * WPB;
    constexpr int PAIRS_PER_WARP = TILE / 2 / WPB;
    static_assert(TILE % (2 * WPB) == 0, "row pairs must divide evenly over a branch's warps");
    constexpr bool ANY_PRE = PREMASK != 0;
    constexpr bool ALL_PRE = PREMASK == ((1 << NB) - 1);
    constexpr bool MIXED = ANY_PRE && !ALL_PRE;
    extern __shared__ __align__(128) uint8_t dwt_smem[];
    float* sx_all = reinterpret_cast<float*>(dwt_smem);          // [2][DWT_ROWS][C]; row i <-> physical row r0 - 1 + i
    float* s_nrm = sx_all + 2 * DWT_ROWS * C;                    // [DWT_ROWS][C] (MIXED only)
    float* s_zero = s_nrm + (MIXED ? DWT_ROWS * C : 0);          // [C] zeros: taps outside the pair
    float* s_beta = s_zero + C;                                  // [C] LN_pre bias: the first pad column
    uint64_t* bars = reinterpret_cast<uint64_t*>(s_beta + C);    // [2]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = warp / WPB, sub = warp - b * WPB;              // this warp's branch, and its index among the branch's warps
    const bool pre = (PREMASK >> b) & 1;
    const int n_tiles = total_rows / TILE;
    const uint32_t bar_u = (uint32_t)__cvta_generic_to_shared(bars);
    auto issue = [&](int tile, int buf) {                         // one thread: stage the rows of `tile` into buffer `buf`
        const int r0 = tile * TILE;
        const int lo = max(r0 - 1, 0), hi = min(r0 + TILE + 1, total_rows);
        const uint32_t bytes = (uint32_t)(hi - lo) * C * 4;
        const uint32_t bb = bar_u + 8 * buf;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bb), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(sx_all + buf * DWT_ROWS * C + (lo - (r0 - 1)) * C)),
                       "l"(x + (long long)lo * C), "r"(bytes), "r"(bb) : "memory");
    };
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if ((int)blockIdx.x < n_tiles) issue(blockIdx.x, 0);
    }
    for (int i = threadIdx.x; i < C; i += NW * 32) {
        s_zero[i] = 0.f;
        s_beta[i] = ANY_PRE ? pre_b[i] : 0.f;
    }
    // the branch's parameters, resident in registers: lane owns channels (j * 32 + lane) * 4 .. + 3 of every 128-channel chunk j
    f2 w0[NCH][2], w1[NCH][2], w2[NCH][2], gm[NCH][2], bt[NCH][2];
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int c = (j * 32 + lane) * 4;
        f2_ldg(br.w[b] + c, w0[j]); f2_ldg(br.w[b] + C + c, w1[j]); f2_ldg(br.w[b] + 2 * C + c, w2[j]);
        f2_ldg(br.g[b] + c, gm[j]); f2_ldg(br.b[b] + c, bt[j]);
    }
    TO* const out_b = (TO*)br.out[b];
    const long long ldo_b = br.ldo[b];
    __syncthreads();                                             // barriers initialised before anyone polls them
    int it = 0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int buf = it & 1;
        float* sx = sx_all + buf * DWT_ROWS * C;
        const float* sn = MIXED ? s_nrm : sx;                    // where the normalised rows live
        const int r0 = tile * TILE;
        const int lo = max(r0 - 1, 0), hi = min(r0 + TILE + 1, total_rows);
        if (threadIdx.x == 0 && tile + (int)gridDim.x < n_tiles) issue(tile + gridDim.x, buf ^ 1);   // freed by the trailing sync
        {
            const uint32_t parity = (it >> 1) & 1;
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(bar_u + 8 * buf), "r"(parity) : "memory");
            }
        }
        if (ANY_PRE) {
            // every warp of the CTA normalises its share of the staged rows once (rows i = warp, warp + NW, ...)
            const uint32_t sx_base = (uint32_t)__cvta_generic_to_shared(sx) + lane * 16;
            const uint32_t dst_base = (uint32_t)__cvta_generic_to_shared(MIXED ? s_nrm : sx) + lane * 16;
            for (int i = warp; i < DWT_ROWS; i += NW) {
                const int p = r0 - 1 + i;
                if (p < lo || p >= hi) continue;                 // warp-uniform
                f2 v[NCH][2];
#pragma unroll
                for (int j = 0; j < NCH; ++j) f2_lds(sx_base + (uint32_t)i * C * 4 + j * 512, v[j]);
                f2 sm = f2_add(v[0][0], v[0][1]);
#pragma unroll
                for (int j = 1; j < NCH; ++j) sm = f2_add(sm, f2_add(v[j][0], v[j][1]));
                const float mean = warp_sum(f2_hsum(sm)) * (1.0f / C);
                const f2 mm = f2_splat(mean);
                f2 q = 0ull;
#pragma unroll
                for (int j = 0; j < NCH; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) { v[j][e] = f2_sub(v[j][e], mm); q = f2_fma(v[j][e], v[j][e], q); }
                const f2 rs = f2_splat(rsqrtf(warp_sum(f2_hsum(q)) * (1.0f / C) + VRD_EPS));
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    f2 g[2], be[2];
                    f2_ldg(pre_g + (j * 32 + lane) * 4, g);
                    f2_ldg(pre_b + (j * 32 + lane) * 4, be);
#pragma unroll
                    for (int e = 0; e < 2; ++e) v[j][e] = f2_fma(f2_mul(v[j][e], rs), g[e], be[e]);
                    f2_sts(dst_base + (uint32_t)i * C * 4 + j * 512, v[j]);
                }
            }
            __syncthreads();
        }
        {
            const uint32_t sx_u = (uint32_t)__cvta_generic_to_shared(sx), sn_u = (uint32_t)__cvta_generic_to_shared(sn);
            const uint32_t zero_u = (uint32_t)__cvta_generic_to_shared(s_zero), beta_u = (uint32_t)__cvta_generic_to_shared(s_beta);
            const uint32_t src_u = pre ? sn_u : sx_u;
#pragma unroll 1
            for (int pp = 0; pp < PAIRS_PER_WARP; ++pp) {
                const int rr0 = 2 * (sub * PAIRS_PER_WARP + pp);                     // first of the two output rows, within the tile
                uint32_t awin[4];
                int wseq[4];
#pragma unroll
                for (int w = 0; w < 4; ++w) {
                    const int grow = r0 + rr0 - 1 + w;
                    int seq = -1;
                    if (grow >= 0 && grow < total_rows) seq = lay.row_seq[grow % lay.R];
                    wseq[w] = seq;
                    awin[w] = (seq >= 0 ? src_u + (uint32_t)(rr0 + w) * C * 4 : zero_u) + lane * 16;
                }
                bool live[2];
                uint32_t hib = 0;                                // bit u: tap +1 of output row u is a first pad column
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    live[u] = wseq[u + 1] >= 0;
                    if (pre && live[u] && wseq[u + 2] < 0 && lay.seqinfo[wseq[u + 1]].z != 0) hib |= 1u << u;
                }
                f2 y[2][NCH][2];
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    f2 a[4][2];
#pragma unroll
                    for (int w = 0; w < 4; ++w) f2_lds(awin[w] + j * 512, a[w]);
#pragma unroll
                    for (int u = 0; u < 2; ++u)
#pragma unroll
                        for (int e = 0; e < 2; ++e) y[u][j][e] = f2_fma(a[u + 2][e], w2[j][e], f2_fma(a[u + 1][e], w1[j][e], f2_mul(a[u][e], w0[j][e])));
                    if (hib != 0) {                              // warp-uniform and rare (one row per padded pair)
                        f2 pb[2];
                        f2_lds(beta_u + lane * 16 + j * 512, pb);
#pragma unroll
                        for (int u = 0; u < 2; ++u)
                            if ((hib >> u) & 1) {
#pragma unroll
                                for (int e = 0; e < 2; ++e) y[u][j][e] = f2_fma(pb[e], w2[j][e], y[u][j][e]);   // the tap read 0 above
                            }
                    }
                }
                float mean[2], rstd[2];
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    f2 sm = f2_add(y[u][0][0], y[u][0][1]);
#pragma unroll
                    for (int j = 1; j < NCH; ++j) sm = f2_add(sm, f2_add(y[u][j][0], y[u][j][1]));
                    mean[u] = f2_hsum(sm);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int u = 0; u < 2; ++u) mean[u] += __shfl_xor_sync(FULL_MASK, mean[u], o);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    mean[u] *= (1.0f / C);
                    const f2 mm = f2_splat(mean[u]);
                    f2 q = 0ull;
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
#pragma unroll
                        for (int e = 0; e < 2; ++e) { y[u][j][e] = f2_sub(y[u][j][e], mm); q = f2_fma(y[u][j][e], y[u][j][e], q); }
                    rstd[u] = f2_hsum(q);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                    for (int u = 0; u < 2; ++u) rstd[u] += __shfl_xor_sync(FULL_MASK, rstd[u], o);
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const f2 rs = f2_splat(rsqrtf(rstd[u] * (1.0f / C) + VRD_EPS));
                    TO* o = out_b + (long long)(r0 + rr0 + u) * ldo_b;
                    if (live[u]) {                                          // warp-uniform
#pragma unroll
                        for (int j = 0; j < NCH; ++j) {
#pragma unroll
                            for (int e = 0; e < 2; ++e) y[u][j][e] = f2_fma(f2_mul(y[u][j][e], rs), gm[j][e], bt[j][e]);
                            f2_stg(o + (j * 32 + lane) * 4, y[u][j]);
                        }
                    } else {
                        zero_row<TO, NCH>(o, lane);                         // separator rows -> 0
                    }
                }
            }
        }
        if constexpr (ALL_PRE) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic writes before the next TMA fill
        __syncthreads();                                         // this buffer may be refilled from now on
    }
}

// ------------------------------------------------------------------------------------------------------------
// dwconv_ln_qr ("quarter-row warps", VRD_DW_CFG=4): same arithmetic per element as dwconv_ln_tile, different ownership.  ncu on
// dwconv_ln_tile: the shared-memory / L1 data pipe is 63-72 % busy and more than half of its ~280 wavefronts per row are the
// per-channel parameters (three conv taps, LayerNorm gamma / beta of three branches) re-read for every two rows, the rest is the
// stencil window re-read per branch.  Here a lane owns FOUR channels (a warp: one 128-channel quarter of a row), so the parameters
// of ALL branches fit in registers (20 per branch) for the lifetime of the persistent CTA, and a warp walks DOWN a strip of
// QR_S rows with the stencil window in registers: every staged row is read once for the pre-LayerNorm statistics and once for the
// window (~70 wavefronts per row instead of ~280).  The price is that a LayerNorm row is spread over four warps: the post-conv
// statistics (sum and sum of squares per row and branch, one pass) of two rows at a time are reduced inside each warp through a
// transposition in shared memory (12 conflict-free stores, then 24 lanes add 16 values each), exchanged between the four warps
// behind a 128-thread named barrier, finished by one lane per (row, branch) and broadcast by shuffles.  The four warps of a group form an independent pipeline (own TMA double
// buffer, own mbarriers, own named barrier): NG groups per CTA drift out of phase and cover each other's dependent chains
// (TMA wait -> statistics -> barrier -> conv -> exchange -> normalise), which dwconv_ln_tile needed two CTAs per SM for.
// ------------------------------------------------------------------------------------------------------------
constexpr int QR_S = 8;                 // output rows per strip
constexpr int QR_ROWS = QR_S + 2;       // staged rows per strip (one halo row on either side)
constexpr int QR_NG = 4;                // groups of four warps per CTA

constexpr int qr_smem_bytes(int ng) {
    // per group: two TMA buffers of QR_ROWS rows, (mean, rstd) and the row codes of a strip (double-buffered), the exchange
    // buffer of two batches (12 values x 4 quarters each), two mbarriers, four per-warp transposition scratches of 12 x 36 floats
    return ng * (2 * QR_ROWS * 512 * 4 + 2 * QR_ROWS * 8 + 2 * QR_ROWS * 4 + 2 * 12 * 4 * 4 + 16 + 4 * 12 * 36 * 4) + 128;
}

template <typename TO, int NB, int PREMASK, int NG>
__global__ void __launch_bounds__(NG * 128, 1) dwconv_ln_qr_kernel(const float* __restrict__ x, Lay lay,
                                                                    const float* __restrict__ pre_g,
                                                                    const float* __restrict__ pre_b, DwBranches br,
                                                                    int total_rows) {
    pdl_wait();
    constexpr int C = 512;
    constexpr bool ANY_PRE = PREMASK != 0;
    constexpr bool ANY_RAW = PREMASK != ((1 << NB) - 1);
    constexpr int NV = 2 * NB * 2;                                // values exchanged per batch of two rows
    extern __shared__ __align__(128) uint8_t qr_smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int g = warp >> 2, wq = warp & 3;                       // group, channel quarter
    const int gt = threadIdx.x & 127;                             // thread index inside the group
    float* s_rows = reinterpret_cast<float*>(qr_smem) + (size_t)g * 2 * QR_ROWS * C;                  // [2][QR_ROWS][C]
    constexpr int AUX = 2 * QR_ROWS * 8 + 2 * QR_ROWS * 4 + 2 * 12 * 4 * 4 + 16 + 4 * 12 * 36 * 4;
    uint8_t* aux = qr_smem + (size_t)NG * 2 * QR_ROWS * C * 4 + (size_t)g * AUX;
    float2* s_stats = reinterpret_cast<float2*>(aux);                                               // [2][QR_ROWS]
    int* s_code = reinterpret_cast<int*>(aux + 2 * QR_ROWS * 8);                                     // [2][QR_ROWS]
    float* s_part = reinterpret_cast<float*>(aux + 2 * QR_ROWS * 8 + 2 * QR_ROWS * 4);               // [2][12][4]
    uint64_t* bars = reinterpret_cast<uint64_t*>(aux + 2 * QR_ROWS * 8 + 2 * QR_ROWS * 4 + 2 * 12 * 4 * 4);   // [2]
    float* s_scr = reinterpret_cast<float*>(aux + 2 * QR_ROWS * 8 + 2 * QR_ROWS * 4 + 2 * 12 * 4 * 4 + 16) + wq * 12 * 36;   // [12][36], this warp's
    const uint32_t bar_u = (uint32_t)__cvta_generic_to_shared(bars);
    const uint32_t rows_u = (uint32_t)__cvta_generic_to_shared(s_rows);
    const int n_strips = total_rows / QR_S;
    const int strip0 = blockIdx.x * NG + g, strip_step = gridDim.x * NG;

    auto issue = [&](int strip, int buf) {                         // one thread of the group: stage the rows of `strip`
        const int r0 = strip * QR_S;
        const int lo = max(r0 - 1, 0), hi = min(r0 + QR_S + 1, total_rows);
        const uint32_t bytes = (uint32_t)(hi - lo) * C * 4;
        const uint32_t b = bar_u + 8 * buf;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(rows_u + (uint32_t)((buf * QR_ROWS + (lo - (r0 - 1))) * C * 4)),
                       "l"(x + (long long)lo * C), "r"(bytes), "r"(b) : "memory");
    };
    if (gt == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u));
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u + 8));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        if (strip0 < n_strips) issue(strip0, 0);
    }
    // the parameters of every branch for this lane's four channels, resident in registers
    const int c0 = wq * 128 + lane * 4;
    f2 w0[NB][2], w1[NB][2], w2[NB][2], gm[NB][2], bt[NB][2], pg[2], pb[2];
#pragma unroll
    for (int b = 0; b < NB; ++b) {
        f2_ldg(br.w[b] + c0, w0[b]); f2_ldg(br.w[b] + C + c0, w1[b]); f2_ldg(br.w[b] + 2 * C + c0, w2[b]);
        f2_ldg(br.g[b] + c0, gm[b]); f2_ldg(br.b[b] + c0, bt[b]);
    }
    if constexpr (ANY_PRE) { f2_ldg(pre_g + c0, pg); f2_ldg(pre_b + c0, pb); }
    else { pg[0] = pg[1] = pb[0] = pb[1] = 0ull; }
    __syncthreads();                                               // barriers initialised before anyone polls them

    int it = 0;
    for (int strip = strip0; strip < n_strips; strip += strip_step, ++it) {
        const int buf = it & 1;
        const int r0 = strip * QR_S;
        const int lo = max(r0 - 1, 0), hi = min(r0 + QR_S + 1, total_rows);
        // every warp of the group has passed the last exchange barrier of the previous strip: its rows (buffer buf ^ 1) are dead
        if (gt == 0 && strip + strip_step < n_strips) issue(strip + strip_step, buf ^ 1);
        {
            const uint32_t parity = (it >> 1) & 1;
            uint32_t ok = 0;
            while (!ok) {
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                             : "=r"(ok) : "r"(bar_u + 8 * buf), "r"(parity) : "memory");
            }
        }
        const uint32_t sx_u = rows_u + (uint32_t)(buf * QR_ROWS * C * 4);
        float2* stats = s_stats + buf * QR_ROWS;
        int* code = s_code + buf * QR_ROWS;
        // ---- phase 1: row codes (bit 0: the row belongs to a pair, bit 1: that pair has a first pad column) and the
        //      pre-LayerNorm statistics of every staged row (two-pass, as dwconv_ln_tile)
        if (wq == 0 && lane < QR_ROWS) {
            const int p = r0 - 1 + lane;
            int cd = 0;
            if (p >= lo && p < hi) {
                const int seq = lay.row_seq[p % lay.R];
                if (seq >= 0) cd = 1 | (lay.seqinfo[seq].z != 0 ? 2 : 0);
            }
            code[lane] = cd;
        }
        if constexpr (ANY_PRE) {
            f2 v[3][4][2];
            bool on[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int i = wq + 4 * u;
                const int p = r0 - 1 + i;
                on[u] = i < QR_ROWS && p >= lo && p < hi;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    if (on[u]) f2_lds(sx_u + (uint32_t)i * C * 4 + j * 512 + lane * 16, v[u][j]);
                    else v[u][j][0] = v[u][j][1] = 0ull;
                }
            }
            float mean[3], rstd[3];
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                f2 sm = f2_add(v[u][0][0], v[u][0][1]);
#pragma unroll
                for (int j = 1; j < 4; ++j) sm = f2_add(sm, f2_add(v[u][j][0], v[u][j][1]));
                mean[u] = f2_hsum(sm);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int u = 0; u < 3; ++u) mean[u] += __shfl_xor_sync(FULL_MASK, mean[u], o);
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                mean[u] *= (1.0f / C);
                const f2 mm = f2_splat(mean[u]);
                f2 q = 0ull;
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int e = 0; e < 2; ++e) { const f2 d = f2_sub(v[u][j][e], mm); q = f2_fma(d, d, q); }
                rstd[u] = f2_hsum(q);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1)
#pragma unroll
                for (int u = 0; u < 3; ++u) rstd[u] += __shfl_xor_sync(FULL_MASK, rstd[u], o);
#pragma unroll
            for (int u = 0; u < 3; ++u) {
                const int i = wq + 4 * u;
                if (lane == 0 && i < QR_ROWS) stats[i] = make_float2(mean[u], rsqrtf(rstd[u] * (1.0f / C) + VRD_EPS));
            }
        }
        asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");

        // ---- phase 2: walk down the strip, two output rows per batch (fully unrolled: the window rotates by renaming)
        f2 wn[QR_ROWS][2], wr[QR_ROWS][2];                        // staged rows of the strip, this lane's channels: normalised / raw
        int wc[QR_ROWS];
        auto fetch = [&](int i) {
            const int cd = code[i];
            wc[i] = cd;
            // rows outside a pair (separators, padding, rows past the ends) contribute zeros: multiply by the row's 0 / 1 flag
            // instead of branching (the staged bytes are finite: separator rows are zero rows, out-of-range rows are never read)
            f2 raw[2] = {0ull, 0ull};
            if (cd & 1) f2_lds(sx_u + (uint32_t)(i * C * 4 + c0 * 4), raw);
            wr[i][0] = raw[0]; wr[i][1] = raw[1];
            wn[i][0] = wn[i][1] = 0ull;
            if constexpr (ANY_PRE) {
                if (cd & 1) {
                    const float2 st = stats[i];
                    const f2 mm = f2_splat(st.x), rs = f2_splat(st.y);
#pragma unroll
                    for (int e = 0; e < 2; ++e) wn[i][e] = f2_fma(f2_mul(f2_sub(raw[e], mm), rs), pg[e], pb[e]);
                }
            }
        };
        TO* optr[NB];
#pragma unroll
        for (int b = 0; b < NB; ++b) optr[b] = (TO*)br.out[b] + (long long)r0 * br.ldo[b] + c0;
        fetch(0);
        fetch(1);
#pragma unroll
        for (int k = 0; k < QR_S / 2; ++k) {
            fetch(2 * k + 2);
            fetch(2 * k + 3);
            f2 y[2][NB][2];
            bool live[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                const int i = 2 * k + u;                           // window rows i, i + 1, i + 2; output row = staged row i + 1
                live[u] = (wc[i + 1] & 1) != 0;
                // tap +1 of the last row of a padded pair reads the first pad column, whose LN_pre value is the bias
                const bool hib = ANY_PRE && live[u] && (wc[i + 2] & 1) == 0 && (wc[i + 1] & 2) != 0;
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    const bool pre = (PREMASK >> b) & 1;
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const f2 a0 = pre ? wn[i][e] : wr[i][e];
                        const f2 a1 = pre ? wn[i + 1][e] : wr[i + 1][e];
                        f2 a2 = pre ? wn[i + 2][e] : wr[i + 2][e];
                        if (pre && hib) a2 = pb[e];
                        y[u][b][e] = f2_fma(a2, w2[b][e], f2_fma(a1, w1[b][e], f2_mul(a0, w0[b][e])));
                    }
                    s_scr[((u * NB + b) * 2) * 36 + lane] = f2_hsum(f2_add(y[u][b][0], y[u][b][1]));
                    s_scr[((u * NB + b) * 2 + 1) * 36 + lane] = f2_hsum(f2_fma(y[u][b][1], y[u][b][1], f2_mul(y[u][b][0], y[u][b][0])));
                }
            }
            __syncwarp();
            float* part = s_part + (k & 1) * 48;
            if (lane < 2 * NV) {                                   // lane = (value, half): 16 of the warp's 32 partial sums
                const float4* q = reinterpret_cast<const float4*>(s_scr + (lane >> 1) * 36 + (lane & 1) * 16);
                const float4 q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
                float t = (((q0.x + q0.y) + (q0.z + q0.w)) + ((q1.x + q1.y) + (q1.z + q1.w))) +
                          (((q2.x + q2.y) + (q2.z + q2.w)) + ((q3.x + q3.y) + (q3.z + q3.w)));
                t += __shfl_xor_sync((2 * NV >= 32) ? FULL_MASK : ((1u << (2 * NV)) - 1u), t, 1);
                if ((lane & 1) == 0) part[(lane >> 1) * 4 + wq] = t;
            }
            asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory");
            // lane j < 2 NB finishes (row u, branch b) = j: totals over the four quarters, mean and 1 / std (one-pass variance)
            float mj = 0.f, nj = 0.f;                              // 1 / std and -mean / std
            if (lane < 2 * NB) {
                const float4 s1 = *reinterpret_cast<const float4*>(part + (2 * lane) * 4);
                const float4 s2 = *reinterpret_cast<const float4*>(part + (2 * lane + 1) * 4);
                const float mean = ((s1.x + s1.y) + (s1.z + s1.w)) * (1.0f / C);
                const float var = fmaxf(((s2.x + s2.y) + (s2.z + s2.w)) * (1.0f / C) - mean * mean, 0.f);
                mj = rsqrtf(var + VRD_EPS);
                nj = -mean * mj;
            }
#pragma unroll
            for (int u = 0; u < 2; ++u) {
                f2 rs[NB], ms[NB];
#pragma unroll
                for (int b = 0; b < NB; ++b) {
                    rs[b] = f2_splat(__shfl_sync(FULL_MASK, mj, u * NB + b));
                    ms[b] = f2_splat(__shfl_sync(FULL_MASK, nj, u * NB + b));
                }
                if (live[u]) {                                     // warp-uniform
#pragma unroll
                    for (int b = 0; b < NB; ++b) {
                        f2 o[2];
#pragma unroll
                        for (int e = 0; e < 2; ++e) o[e] = f2_fma(f2_fma(y[u][b][e], rs[b], ms[b]), gm[b][e], bt[b][e]);
                        f2_stg(optr[b], o);
                    }
                } else {                                           // separator rows -> 0
                    const f2 z[2] = {0ull, 0ull};
#pragma unroll
                    for (int b = 0; b < NB; ++b) f2_stg(optr[b], z);
                }
#pragma unroll
                for (int b = 0; b < NB; ++b) optr[b] += br.ldo[b];
            }
        }
    }
}

template <typename TO>
static int dwconv_ln_tile(const float* x, Lay lay, const float* pre_g, const float* pre_b, const DwBranches& br, int streams,
                          cudaStream_t st) {
    const int num_sms = device_sm_count();
    int mask = 0;
    for (int b = 0; b < br.n; ++b) mask |= (br.use_pre[b] ? 1 : 0) << b;
    // 0: 8 warps x 32 rows, 1: 16 x 32, 2: 2 CTAs/SM of 8 x 16, 3: branch warps, 4: quarter-row warps, 5: quarter-row warps for the
    // launches with two or three branches (measured 9 % faster there, 5 % slower on the single-branch launch), tiles otherwise
    int cfg = vrd_options().dw_cfg;
    if (cfg == 5) cfg = br.n >= 2 ? 4 : 2;
    if (cfg == 4) {
        const int total = streams * lay.R;
        const int n_strips = total / QR_S;
#define LAUNCH_QR(NB, MASK, NG) do { \
        auto kern = dwconv_ln_qr_kernel<TO, NB, MASK, NG>; \
        constexpr int smem = qr_smem_bytes(NG); \
        static PerDeviceOnce once; \
        if (once.first() && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1; \
        const int need = (n_strips + NG - 1) / NG; \
        launch_k(kern, dim3(need < num_sms ? need : num_sms), dim3(NG * 128), smem, st, x, lay, pre_g, pre_b, br, total); } while (0)
        if (br.n == 3 && mask == 7) LAUNCH_QR(3, 7, QR_NG);
        else if (br.n == 3 && mask == 3) LAUNCH_QR(3, 3, QR_NG);
        else if (br.n == 2 && mask == 0) LAUNCH_QR(2, 0, QR_NG);
        else if (br.n == 1 && mask == 1) LAUNCH_QR(1, 1, QR_NG);
        else return 1;
#undef LAUNCH_QR
        return 0;
    }
    if (cfg == 3) {
        constexpr int TILE = 32;
        const int total = streams * lay.R;
        const int n_tiles = total / TILE;
        const int grid = n_tiles < num_sms ? n_tiles : num_sms;
#define LAUNCH_BW(NB, MASK, WPB) do { \
        auto kern = dwconv_ln_bw_kernel<TO, NB, MASK, WPB, TILE>; \
        constexpr int smem = dwt_smem_bytes(MASK != 0 && MASK != ((1 << NB) - 1), TILE); \
        static PerDeviceOnce once; \
        if (once.first() && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1; \
        launch_k(kern, dim3(grid), dim3(NB * WPB * 32), smem, st, x, lay, pre_g, pre_b, br, total); } while (0)
        if (br.n == 3 && mask == 7) LAUNCH_BW(3, 7, 4);
        else if (br.n == 3 && mask == 3) LAUNCH_BW(3, 3, 4);
        else if (br.n == 2 && mask == 0) LAUNCH_BW(2, 0, 8);
        else if (br.n == 1 && mask == 1) LAUNCH_BW(1, 1, 16);
        else return 1;
#undef LAUNCH_BW
        return 0;
    }
    const int tile_rows = cfg == 2 ? 16 : 32;
    const int total = streams * lay.R;
    const int n_tiles = total / tile_rows;
    const int max_ctas = num_sms * (32 / tile_rows);
    const int grid = n_tiles < max_ctas ? n_tiles : max_ctas;
#define LAUNCH_NW(NB, MASK, NW, TILE) do { \
        auto kern = dwconv_ln_tile_kernel<TO, NB, MASK, NW, TILE>; \
        constexpr int smem = dwt_smem_bytes(MASK != 0 && MASK != ((1 << NB) - 1), TILE); \
        static PerDeviceOnce once; \
        if (once.first() && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1; \
        launch_k(kern, dim3(grid), dim3(NW * 32), smem, st, x, lay, pre_g, pre_b, br, total); } while (0)
#define LAUNCH(NB, MASK) do { if (cfg == 0) LAUNCH_NW(NB, MASK, 8, 32); else if (cfg == 1) LAUNCH_NW(NB, MASK, 16, 32); \
                               else LAUNCH_NW(NB, MASK, 8, 16); } while (0)
    if (br.n == 3 && mask == 7) LAUNCH(3, 7);
    else if (br.n == 3 && mask == 3) LAUNCH(3, 3);
    else if (br.n == 2 && mask == 0) LAUNCH(2, 0);
    else if (br.n == 1 && mask == 1) LAUNCH(1, 1);
    else return 1;
#undef LAUNCH
#undef LAUNCH_NW
    return 0;
}

template <typename TI, typename TO, int NCH, int STRIDE>
static int dwconv_ln_mask(const void* x, long long ldx, Lay lin, Lay lout, const float* pre_g, const float* pre_b,
                          const DwBranches& br, int streams, cudaStream_t st) {
    int mask = 0;
    for (int b = 0; b < br.n; ++b) mask |= (br.use_pre[b] ? 1 : 0) << b;
    const int runs = streams * ((lout.R + DW_RUN - 1) / DW_RUN);
    const int grid = (runs + DW_WARPS - 1) / DW_WARPS;
#define LAUNCH(NB, MASK) \
    launch_k(dwconv_ln_kernel<TI, TO, NCH, STRIDE, NB, MASK>, dim3(grid), dim3(DW_WARPS * 32), 0, st, (const TI*)x, ldx, lin, lout, pre_g, pre_b, br, streams)
    if (br.n == 3 && mask == 7) LAUNCH(3, 7);
    else if (br.n == 3 && mask == 3) LAUNCH(3, 3);
    else if (br.n == 2 && mask == 0) LAUNCH(2, 0);
    else if (br.n == 1 && mask == 1) LAUNCH(1, 1);
    else return 1;
#undef LAUNCH
    return 0;
}

int dwconv_ln(const void* x, int xdt, long long ldx, Lay lin, Lay lout, int stride, const float* pre_g, const float* pre_b,
              const DwBranches& br, int odt, int C, int streams, cudaStream_t st) {
    if (xdt != VRD_F32) return 1;
    if (stride == 1 && C == 512 && ldx == C && lin.R == lout.R && lin.R % 128 == 0)
        return odt == VRD_BF16 ? dwconv_ln_tile<__nv_bfloat16>((const float*)x, lout, pre_g, pre_b, br, streams, st)
                               : dwconv_ln_tile<float>((const float*)x, lout, pre_g, pre_b, br, streams, st);
#define DISPATCH(TO, NCH) \
    (stride == 1 ? dwconv_ln_mask<float, TO, NCH, 1>(x, ldx, lin, lout, pre_g, pre_b, br, streams, st) \
                 : dwconv_ln_mask<float, TO, NCH, 2>(x, ldx, lin, lout, pre_g, pre_b, br, streams, st))
    if (C == 512) return odt == VRD_BF16 ? DISPATCH(__nv_bfloat16, 4) : DISPATCH(float, 4);
    if (C == 256) return odt == VRD_BF16 ? DISPATCH(__nv_bfloat16, 2) : DISPATCH(float, 2);
#undef DISPATCH
    return 1;
}

// ------------------------------------------------------------------------------------------------------------
// maxpool_skip: out[i] = max(x[2i-1], x[2i], x[2i+1]) (left edge ignored, first pad column counts as 0)
// ------------------------------------------------------------------------------------------------------------
template <int NCH>
__global__ void maxpool_skip_kernel(const float* __restrict__ x, long long ldx, Lay lin, Lay lout, float* __restrict__ out,
                                    long long ldo) {
    pdl_wait();
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (r >= lout.R) return;
    float* o = out + (long long)r * ldo;
    const int seq = lout.row_seq[r];
    if (seq < 0) { zero_row<float, NCH>(o, lane); return; }
    const int4 so = lout.seqinfo[seq], si = lin.seqinfo[seq];
    const int t = (r - so.x) * 2;
    const float* base = x + (long long)si.x * ldx;
    float m[NCH][4];
    load_row<float, NCH>(base + (long long)t * ldx, lane, m);
    if (t - 1 >= 0) {
        float v[NCH][4];
        load_row<float, NCH>(base + (long long)(t - 1) * ldx, lane, v);
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) m[j][i] = fmaxf(m[j][i], v[j][i]);
    }
    if (t + 1 < si.y) {
        float v[NCH][4];
        load_row<float, NCH>(base + (long long)(t + 1) * ldx, lane, v);
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) m[j][i] = fmaxf(m[j][i], v[j][i]);
    } else {   // t + 1 == len: a zero pad column
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int i = 0; i < 4; ++i) m[j][i] = fmaxf(m[j][i], 0.f);
    }
    store_row<float, NCH>(o, lane, m);
}

int maxpool_skip(const float* x, long long ldx, Lay lin, Lay lout, float* out, long long ldo, int C, cudaStream_t st) {
    if (C != 512) return 1;
    launch_k(maxpool_skip_kernel<4>, dim3((lout.R + WARPS - 1) / WARPS), dim3(WARPS * 32), 0, st, x, ldx, lin, lout, out, ldo);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// FPN kernels (fpn_dim = 256 = 2 chunks per lane; top level reads 512 input channels)
// ------------------------------------------------------------------------------------------------------------
// fpn_top: grouped k=3 conv, out[c] = sum_{j<2,k<3} W[c,j,k] * LN_pre(x)[t+k-1, 2c+j]; wt[k][e] = W[e/2, e%2, k].
__global__ void fpn_top_kernel(const float* __restrict__ x, long long ldx, Lay lay, const float* __restrict__ pre_g,
                               const float* __restrict__ pre_b, const float* __restrict__ wt, const float* __restrict__ g,
                               const float* __restrict__ b, float* __restrict__ out, long long ldo) {
    pdl_wait();
    constexpr int NCH = 4;
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (r >= lay.R) return;
    float* o = out + (long long)r * ldo;
    const int seq = lay.row_seq[r];
    if (seq < 0) {
        for (int j = 0; j < NCH; ++j) *reinterpret_cast<float2*>(o + (j * 32 + lane) * 2) = make_float2(0.f, 0.f);
        return;
    }
    const int4 si = lay.seqinfo[seq];
    const int t = r - si.x;
    float acc[NCH][2];
#pragma unroll
    for (int j = 0; j < NCH; ++j) acc[j][0] = acc[j][1] = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int tt = t + d - 1;
        float v[NCH][4];
        if (tt >= 0 && tt < si.y) {
            load_row<float, NCH>(x + (long long)(si.x + tt) * ldx, lane, v);
            row_normalize<NCH>(v, lane, pre_g, pre_b);
        } else if (tt == si.y && si.z != 0) {
#pragma unroll
            for (int j = 0; j < NCH; ++j) ld4(pre_b + (j * 32 + lane) * 4, v[j]);
        } else {
            continue;
        }
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            float w[4];
            ld4(wt + d * 512 + (j * 32 + lane) * 4, w);
            acc[j][0] += v[j][0] * w[0] + v[j][1] * w[1];
            acc[j][1] += v[j][2] * w[2] + v[j][3] * w[3];
        }
    }
    // LayerNorm over the 256 outputs: lane owns channels (j*32+lane)*2 + {0,1}
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) s += acc[j][0] + acc[j][1];
    const float mean = warp_sum(s) * (1.0f / 256);
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) { float d0 = acc[j][0] - mean, d1 = acc[j][1] - mean; q += d0 * d0 + d1 * d1; }
    const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / 256) + VRD_EPS);
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int c = (j * 32 + lane) * 2;
        const float2 gg = *reinterpret_cast<const float2*>(g + c), bb = *reinterpret_cast<const float2*>(b + c);
        *reinterpret_cast<float2*>(o + c) =
            make_float2((acc[j][0] - mean) * rstd * gg.x + bb.x, (acc[j][1] - mean) * rstd * gg.y + bb.y);
    }
}

int fpn_top(const float* x, long long ldx, Lay lay, const float* pre_g, const float* pre_b, const float* wt, const float* g,
            const float* b, float* out, long long ldo, cudaStream_t st) {
    launch_k(fpn_top_kernel, dim3((lay.R + WARPS - 1) / WARPS), dim3(WARPS * 32), 0, st, x, ldx, lay, pre_g, pre_b, wt, g, b, out, ldo);
    return 0;
}

// fpn_level: z[u] = LN_lat(cur[u]) + y_up[u/2]; out[t] = LN(w0 z[t-1] + w1 z[t] + w2 z[t+1]).
__global__ void fpn_level_kernel(const float* __restrict__ cur, long long ldc, const float* __restrict__ yup, long long ldu,
                                 Lay lay, Lay lup, const float* __restrict__ lat_g, const float* __restrict__ lat_b,
                                 const float* __restrict__ beta_up, const float* __restrict__ w, const float* __restrict__ g,
                                 const float* __restrict__ b, float* __restrict__ out, long long ldo) {
    pdl_wait();
    constexpr int NCH = 2;
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (r >= lay.R) return;
    float* o = out + (long long)r * ldo;
    const int seq = lay.row_seq[r];
    if (seq < 0) { zero_row<float, NCH>(o, lane); return; }
    const int4 si = lay.seqinfo[seq], su = lup.seqinfo[seq];
    const int t = r - si.x;
    float y[NCH][4];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) y[j][i] = 0.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int u = t + d - 1;
        float z[NCH][4];
        if (u >= 0 && u < si.y) {
            load_row<float, NCH>(cur + (long long)(si.x + u) * ldc, lane, z);
            row_normalize<NCH>(z, lane, lat_g, lat_b);
        } else if (u == si.y && si.z != 0) {
#pragma unroll
            for (int j = 0; j < NCH; ++j) ld4(lat_b + (j * 32 + lane) * 4, z[j]);
        } else {
            continue;
        }
        const int uu = u >> 1;
        float up[NCH][4];
        if (uu < su.y) load_row<float, NCH>(yup + (long long)(su.x + uu) * ldu, lane, up);
        else {
#pragma unroll
            for (int j = 0; j < NCH; ++j) ld4(beta_up + (j * 32 + lane) * 4, up[j]);
        }
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            float wv[4];
            ld4(w + d * 256 + (j * 32 + lane) * 4, wv);
#pragma unroll
            for (int i = 0; i < 4; ++i) y[j][i] += (z[j][i] + up[j][i]) * wv[i];
        }
    }
    row_normalize<NCH>(y, lane, g, b);
    store_row<float, NCH>(o, lane, y);
}

int fpn_level(const float* cur, long long ldc, const float* yup, long long ldu, Lay lay, Lay lup, const float* lat_g,
              const float* lat_b, const float* beta_up, const float* w, const float* g, const float* b, float* out,
              long long ldo, cudaStream_t st) {
    launch_k(fpn_level_kernel, dim3((lay.R + WARPS - 1) / WARPS), dim3(WARPS * 32), 0, st, cur, ldc, yup, ldu, lay, lup, lat_g, lat_b, beta_up, w,
                                                                           g, b, out, ldo);
    return 0;
}

// mask_features: depthwise k=3 + bias over the level-0 FPN output (whose first pad column is beta of its LayerNorm)
__global__ void mask_features_kernel(const float* __restrict__ y, long long ldy, Lay lay, const float* __restrict__ beta,
                                     const float* __restrict__ w, const float* __restrict__ bias, float* __restrict__ out,
                                     long long ldo) {
    pdl_wait();
    constexpr int NCH = 2;
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (r >= lay.R) return;
    float* o = out + (long long)r * ldo;
    const int seq = lay.row_seq[r];
    if (seq < 0) { zero_row<float, NCH>(o, lane); return; }
    const int4 si = lay.seqinfo[seq];
    const int t = r - si.x;
    float acc[NCH][4];
#pragma unroll
    for (int j = 0; j < NCH; ++j) ld4(bias + (j * 32 + lane) * 4, acc[j]);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const int u = t + d - 1;
        float z[NCH][4];
        if (u >= 0 && u < si.y) load_row<float, NCH>(y + (long long)(si.x + u) * ldy, lane, z);
        else if (u == si.y && si.z != 0) {
#pragma unroll
            for (int j = 0; j < NCH; ++j) ld4(beta + (j * 32 + lane) * 4, z[j]);
        } else continue;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            float wv[4];
            ld4(w + d * 256 + (j * 32 + lane) * 4, wv);
#pragma unroll
            for (int i = 0; i < 4; ++i) acc[j][i] += z[j][i] * wv[i];
        }
    }
    store_row<float, NCH>(o, lane, acc);
}

int mask_features(const float* y, long long ldy, Lay lay, const float* beta, const float* w, const float* bias, float* out,
                  long long ldo, cudaStream_t st) {
    launch_k(mask_features_kernel, dim3((lay.R + WARPS - 1) / WARPS), dim3(WARPS * 32), 0, st, y, ldy, lay, beta, w, bias, out, ldo);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// query_ln: out = LN2(dw * (LN(x) + pos[row % Q])) with each stage optional (256 channels); rows >= nrows -> 0
// ------------------------------------------------------------------------------------------------------------
template <typename TO>
__global__ void query_ln_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ g,
                                const float* __restrict__ b, const float* __restrict__ pos, int Q, int nrows, int total_rows,
                                const float* __restrict__ dw, const float* __restrict__ g2, const float* __restrict__ b2,
                                TO* __restrict__ out, long long ldo) {
    pdl_wait();
    constexpr int NCH = 2;
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (row >= total_rows) return;
    TO* o = out + (long long)row * ldo;
    if (row >= nrows) { zero_row<TO, NCH>(o, lane); return; }
    float v[NCH][4];
    load_row<float, NCH>(x + (long long)row * ldx, lane, v);
    if (g != nullptr) row_normalize<NCH>(v, lane, g, b);
    if (pos != nullptr) {
        const float* p = pos + (long long)(row % Q) * 256;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            float pv[4];
            ld4(p + (j * 32 + lane) * 4, pv);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[j][i] += pv[i];
        }
    }
    if (dw != nullptr) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            float wv[4];
            ld4(dw + (j * 32 + lane) * 4, wv);
#pragma unroll
            for (int i = 0; i < 4; ++i) v[j][i] *= wv[i];
        }
    }
    if (g2 != nullptr) row_normalize<NCH>(v, lane, g2, b2);
    store_row<TO, NCH>(o, lane, v);
}

int query_ln(const float* x, long long ldx, const float* g, const float* b, const float* pos, int Q, int nrows,
             int total_rows, const float* dw, const float* g2, const float* b2, void* out, int odt, long long ldo, int C,
             cudaStream_t st) {
    if (C != 256) return 1;
    const int grid = (total_rows + WARPS - 1) / WARPS;
    if (odt == VRD_BF16)
        launch_k(query_ln_kernel<__nv_bfloat16>, dim3(grid), dim3(WARPS * 32), 0, st, x, ldx, g, b, pos, Q, nrows, total_rows, dw, g2, b2, (__nv_bfloat16*)out, ldo);
    else
        launch_k(query_ln_kernel<float>, dim3(grid), dim3(WARPS * 32), 0, st, x, ldx, g, b, pos, Q, nrows, total_rows, dw, g2, b2, (float*)out, ldo);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// heads: mask logits + binarisation + first/last active frame; class softmax + top-k
// ------------------------------------------------------------------------------------------------------------
// One block per pair.  masks[row, q] = <me[pair*Q+q, :], mf[row, :]> (256 channels); a frame is active iff
// sigmoid(logit) > 0.5 evaluated in fp32 exactly as the reference does (maskvrd.py:287), NOT logit > 0.
__global__ void mask_logits_kernel(const float* __restrict__ me, long long ldm, const float* __restrict__ mf, long long ldf,
                                   Lay lay, int Q, float* __restrict__ masks, long long ldk, int* __restrict__ first_last) {
    pdl_wait();
    constexpr int NCH = 2;
    constexpr int MAXQ = 16;
    __shared__ int s_first[MAXQ], s_last[MAXQ];
    __shared__ float s_me[MAXQ * 256];
    const int pair = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int4 si = lay.seqinfo[pair];
    if (threadIdx.x < MAXQ) { s_first[threadIdx.x] = 0x7fffffff; s_last[threadIdx.x] = -1; }
    for (int i = threadIdx.x; i < Q * 256; i += blockDim.x)
        s_me[i] = me[(long long)(pair * Q + i / 256) * ldm + (i % 256)];
    __syncthreads();
    for (int t = warp; t < si.y; t += WARPS) {
        const long long row = si.x + t;
        float v[NCH][4];
        load_row<float, NCH>(mf + row * ldf, lane, v);
        for (int q = 0; q < Q; ++q) {
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                const float* m = s_me + q * 256 + (j * 32 + lane) * 4;
#pragma unroll
                for (int i = 0; i < 4; ++i) acc = fmaf(v[j][i], m[i], acc);
            }
            acc = warp_sum(acc);
            if (lane == 0) {
                if (masks != nullptr) masks[row * ldk + q] = acc;
                const float sg = 1.0f / (1.0f + expf(-acc));
                if (sg > 0.5f) { atomicMin(&s_first[q], t); atomicMax(&s_last[q], t); }
            }
        }
    }
    __syncthreads();
    if ((int)threadIdx.x < Q) {
        const int l = s_last[threadIdx.x];
        first_last[(pair * Q + threadIdx.x) * 2 + 0] = (l < 0) ? -1 : s_first[threadIdx.x];
        first_last[(pair * Q + threadIdx.x) * 2 + 1] = l;
    }
}

int mask_logits(const float* me, long long ldm, const float* mf, long long ldf, Lay lay, int Q, float* masks, long long ldk,
                int* first_last, cudaStream_t st) {
    if (Q > 16) return 1;
    launch_k(mask_logits_kernel, dim3(lay.B), dim3(WARPS * 32), 0, st, me, ldm, mf, ldf, lay, Q, masks, ldk, first_last);
    return 0;
}

// One warp per (pair, query): softmax over n_cls logits, then top-k of classes 1..n_cls-1 (ties -> lower class id).
__global__ void softmax_topk_kernel(const float* __restrict__ logits, long long ldl, int nrows, int n_cls, int topk,
                                    float* __restrict__ scores, int* __restrict__ ids) {
    pdl_wait();
    constexpr int PER = 8;   // up to 256 classes
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (row >= nrows) return;
    float p[PER];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = i * 32 + lane;
        p[i] = (c < n_cls) ? logits[(long long)row * ldl + c] : -INFINITY;
        mx = fmaxf(mx, p[i]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = i * 32 + lane;
        p[i] = (c < n_cls) ? expf(p[i] - mx) : 0.f;
        sum += p[i];
    }
    sum = warp_sum(sum);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
        const int c = i * 32 + lane;
        p[i] = (c < n_cls && c > 0) ? p[i] / sum : -1.f;   // background (class 0) and out-of-range are excluded
    }
    for (int k = 0; k < topk; ++k) {
        float best = -1.f;
        int bi = 0x7fffffff;
#pragma unroll
        for (int i = 0; i < PER; ++i) {
            const int c = i * 32 + lane;
            if (p[i] > best) { best = p[i]; bi = c; }   // ascending c within a lane: strict > keeps the lower id
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ob = __shfl_xor_sync(FULL_MASK, best, o);
            const int oi = __shfl_xor_sync(FULL_MASK, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (lane == 0) { scores[(long long)row * topk + k] = best; ids[(long long)row * topk + k] = bi; }
#pragma unroll
        for (int i = 0; i < PER; ++i)
            if (i * 32 + lane == bi) p[i] = -1.f;
    }
}

int softmax_topk(const float* logits, long long ldl, int nrows, int n_cls, int topk, float* scores, int* ids, cudaStream_t st) {
    if (n_cls > 256 || topk >= n_cls) return 1;
    launch_k(softmax_topk_kernel, dim3((nrows + WARPS - 1) / WARPS), dim3(WARPS * 32), 0, st, logits, ldl, nrows, n_cls, topk, scores, ids);
    return 0;
}

}  // namespace vrd
