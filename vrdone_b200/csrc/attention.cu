// Attention kernels over the varlen row layout.
// Reference semantics (under /root/reference/models/): windowed attention blocks.py:920-989 and
// local_transformer.py:553-623 (softmax over keys with |i-j| <= w inside the pair; the -1e4 / -inf masking of the
// padded batch reduces to "valid keys only" in the varlen layout); full attention local_transformer.py:144-187;
// query self/cross attention of the predictor local_transformer.py:33-67, 144-187.
// The 1/sqrt(head_dim) scale is folded into the query projection weights by the host.
// Kernels: window_attn_mma (bf16, head_dim 64: mma.sync, pair-aligned 16-row warp units) with window_attn_tma / window_attn as
// the fp32 and head_dim 128 paths; flash_attn_bf16 (mma.sync flash attention inside each pair) with full_attn as the fp32 path;
// query_self_attn / query_cross_attn for the Q queries of each pair.
#include <cstring>
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"

namespace vrd {

constexpr int WARPS = 8;

// ------------------------------------------------------------------------------------------------------------
// window_attn: one warp per token row, all heads.  LPH = lanes per head (head_dim / 4).
// ------------------------------------------------------------------------------------------------------------
template <typename T, int NCH, int LPH>
__global__ void window_attn_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                   T* __restrict__ out, long long ld, Lay lay, int w, int streams) {
    pdl_wait();
    constexpr int MAXK = 9;
    const int lane = threadIdx.x & 31;
    const int grow = blockIdx.x * WARPS + (threadIdx.x >> 5);
    if (grow >= streams * lay.R) return;
    const int s = grow / lay.R, r = grow - s * lay.R;
    T* o = out + (long long)grow * ld;
    const int seq = lay.row_seq[r];
    if (seq < 0) { zero_row<T, NCH>(o, lane); return; }
    const int4 si = lay.seqinfo[seq];
    const int t = r - si.x;
    const long long base = (long long)s * lay.R + si.x;
    float qv[NCH][4];
    load_row<T, NCH>(q + (long long)grow * ld, lane, qv);
    float sc[MAXK][NCH];
    float mx[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) mx[j] = -INFINITY;
#pragma unroll
    for (int kk = 0; kk < MAXK; ++kk) {
        const int tt = t - w + kk;
        const bool ok = (kk <= 2 * w) && tt >= 0 && tt < si.y;
        if (ok) {
            float kv[NCH][4];
            load_row<T, NCH>(k + (base + tt) * ld, lane, kv);
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                float d = qv[j][0] * kv[j][0] + qv[j][1] * kv[j][1] + qv[j][2] * kv[j][2] + qv[j][3] * kv[j][3];
#pragma unroll
                for (int off = LPH / 2; off > 0; off >>= 1) d += __shfl_xor_sync(FULL_MASK, d, off);
                sc[kk][j] = d;
                mx[j] = fmaxf(mx[j], d);
            }
        } else {
#pragma unroll
            for (int j = 0; j < NCH; ++j) sc[kk][j] = -INFINITY;
        }
    }
    float sum[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) sum[j] = 0.f;
#pragma unroll
    for (int kk = 0; kk < MAXK; ++kk)
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const float p = (sc[kk][j] == -INFINITY) ? 0.f : expf(sc[kk][j] - mx[j]);
            sc[kk][j] = p;
            sum[j] += p;
        }
    float acc[NCH][4];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] = 0.f;
#pragma unroll
    for (int kk = 0; kk < MAXK; ++kk) {
        const int tt = t - w + kk;
        const bool ok = (kk <= 2 * w) && tt >= 0 && tt < si.y;
        if (ok) {
            float vv[NCH][4];
            load_row<T, NCH>(v + (base + tt) * ld, lane, vv);
#pragma unroll
            for (int j = 0; j < NCH; ++j)
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[j][i] = fmaf(sc[kk][j], vv[j][i], acc[j][i]);
        }
    }
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const float inv = 1.0f / sum[j];
#pragma unroll
        for (int i = 0; i < 4; ++i) acc[j][i] *= inv;
    }
    store_row<T, NCH>(o, lane, acc);
}

// ------------------------------------------------------------------------------------------------------------
// window_attn_tma: the same arithmetic, shared-memory tiled.  A CTA owns TILE consecutive rows (of the stacked streams);
// one elected thread stages the K and V rows [r0 - 4, r0 + TILE + 4) with two TMA bulk copies (cp.async.bulk, completion
// on an mbarrier) while the warps fetch their query rows; a warp then handles one query row at a time with every key /
// value row read from shared memory (each is reused by up to 2w+1 queries), scores reduced with warp shuffles inside the
// LPH lanes that share a head.  A lane owns NV 16-byte vectors: vector j = channels j*32*VEC + lane*VEC .. +VEC-1, so
// global and shared accesses are contiguous across the warp (conflict-free LDS.128).  Requires ld == C.
// ------------------------------------------------------------------------------------------------------------
template <typename T> struct Vec16;
template <> struct Vec16<float> {
    static constexpr int N = 4;
    static __device__ __forceinline__ void ld(const float* p, float (&v)[4]) { ld4(p, v); }
    static __device__ __forceinline__ void st(float* p, const float (&v)[4]) { st4(p, v); }
};
template <> struct Vec16<__nv_bfloat16> {
    static constexpr int N = 8;
    static __device__ __forceinline__ void ld(const __nv_bfloat16* p, float (&v)[8]) {
        const uint4 t = *reinterpret_cast<const uint4*>(p);
        const uint32_t u[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&u[i]));
            v[2 * i] = f.x; v[2 * i + 1] = f.y;
        }
    }
    static __device__ __forceinline__ void st(__nv_bfloat16* p, const float (&v)[8]) {
        uint32_t u[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            u[i] = *reinterpret_cast<const uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p) = make_uint4(u[0], u[1], u[2], u[3]);
    }
};

constexpr int WIN_HALO = 4;   // maximum half window

template <typename T, int C, int HS, int TILE>
__global__ void __launch_bounds__(WARPS * 32) window_attn_tma_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                                     const T* __restrict__ v, T* __restrict__ out, Lay lay, int w,
                                                                     int total_rows) {
    pdl_wait();
    constexpr int VEC = Vec16<T>::N, NV = C / (32 * VEC), LPH = HS / VEC, MAXK = 2 * WIN_HALO + 1;
    constexpr int ROWS = TILE + 2 * WIN_HALO;
    extern __shared__ __align__(128) uint8_t win_smem[];
    T* sK = reinterpret_cast<T*>(win_smem);
    T* sV = sK + ROWS * C;
    uint64_t* bar = reinterpret_cast<uint64_t*>(sV + ROWS * C);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int r0 = blockIdx.x * TILE;
    const int lo = max(r0 - WIN_HALO, 0), hi = min(r0 + TILE + WIN_HALO, total_rows);
    const uint32_t bar_u = (uint32_t)__cvta_generic_to_shared(bar);
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_u));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = (uint32_t)(hi - lo) * C * sizeof(T);
        const long long goff = (long long)lo * C;
        const int soff = (lo - (r0 - WIN_HALO)) * C;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_u), "r"(2 * bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(sK + soff)), "l"(k + goff), "r"(bytes), "r"(bar_u) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"((uint32_t)__cvta_generic_to_shared(sV + soff)), "l"(v + goff), "r"(bytes), "r"(bar_u) : "memory");
    }
    auto wait_tiles = [&]() {
        uint32_t ok = 0;
        while (!ok) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.b32 %0, 1, 0, p;\n\t}"
                         : "=r"(ok) : "r"(bar_u), "r"(0u) : "memory");
        }
    };
    bool waited = false;
    for (int rr = warp; rr < TILE; rr += WARPS) {
        const int grow = r0 + rr;
        const int s = grow / lay.R, r = grow - s * lay.R;
        T* o = out + (long long)grow * C;
        const int seq = lay.row_seq[r];
        if (seq < 0) {
            const float z[VEC] = {};
#pragma unroll
            for (int j = 0; j < NV; ++j) Vec16<T>::st(o + (j * 32 + lane) * VEC, z);
            continue;
        }
        const int4 si = lay.seqinfo[seq];
        const int t = r - si.x;
        float qv[NV][VEC];
#pragma unroll
        for (int j = 0; j < NV; ++j) Vec16<T>::ld(q + (long long)grow * C + (j * 32 + lane) * VEC, qv[j]);
        if (!waited) { wait_tiles(); waited = true; }
        const T* kb = sK + (rr + WIN_HALO - w) * C;     // key kk of this query lives at tile row rr + HALO - w + kk
        const T* vb = sV + (rr + WIN_HALO - w) * C;
        float sc[MAXK][NV];
        float mx[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) mx[j] = -INFINITY;
#pragma unroll
        for (int kk = 0; kk < MAXK; ++kk) {
            const int tt = t - w + kk;
            const bool ok = (kk <= 2 * w) && tt >= 0 && tt < si.y;
            if (ok) {
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    float kv[VEC];
                    Vec16<T>::ld(kb + kk * C + (j * 32 + lane) * VEC, kv);
                    float d = 0.f;
#pragma unroll
                    for (int i = 0; i < VEC; ++i) d = fmaf(qv[j][i], kv[i], d);
#pragma unroll
                    for (int off = LPH / 2; off > 0; off >>= 1) d += __shfl_xor_sync(FULL_MASK, d, off);
                    sc[kk][j] = d;
                    mx[j] = fmaxf(mx[j], d);
                }
            } else {
#pragma unroll
                for (int j = 0; j < NV; ++j) sc[kk][j] = -INFINITY;
            }
        }
        float sum[NV];
#pragma unroll
        for (int j = 0; j < NV; ++j) sum[j] = 0.f;
#pragma unroll
        for (int kk = 0; kk < MAXK; ++kk)
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const float p = (sc[kk][j] == -INFINITY) ? 0.f : expf(sc[kk][j] - mx[j]);
                sc[kk][j] = p;
                sum[j] += p;
            }
        float acc[NV][VEC];
#pragma unroll
        for (int j = 0; j < NV; ++j)
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[j][i] = 0.f;
#pragma unroll
        for (int kk = 0; kk < MAXK; ++kk) {
            const int tt = t - w + kk;
            const bool ok = (kk <= 2 * w) && tt >= 0 && tt < si.y;
            if (ok) {
#pragma unroll
                for (int j = 0; j < NV; ++j) {
                    float vv[VEC];
                    Vec16<T>::ld(vb + kk * C + (j * 32 + lane) * VEC, vv);
#pragma unroll
                    for (int i = 0; i < VEC; ++i) acc[j][i] = fmaf(sc[kk][j], vv[i], acc[j][i]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const float inv = 1.0f / sum[j];
#pragma unroll
            for (int i = 0; i < VEC; ++i) acc[j][i] *= inv;
            Vec16<T>::st(o + (j * 32 + lane) * VEC, acc[j]);
        }
    }
    if (!waited) wait_tiles();    // never exit while the bulk copies into this CTA's shared memory are in flight
}

template <typename T, int HS, int TILE>
static int window_attn_tma_launch(const void* q, const void* k, const void* v, void* out, Lay lay, int w, int streams,
                                  cudaStream_t st) {
    constexpr int C = 512;
    constexpr int smem = 2 * (TILE + 2 * WIN_HALO) * C * (int)sizeof(T) + 16;
    static PerDeviceOnce once;
    auto kern = window_attn_tma_kernel<T, C, HS, TILE>;
    if (once.first() && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
    const int total = streams * lay.R;    // R % 128 == 0, so TILE divides it
    launch_k(kern, dim3(total / TILE), dim3(WARPS * 32), smem, st, (const T*)q, (const T*)k, (const T*)v, (T*)out, lay, w, total);
    return 0;
}

static int window_attn_simt(const void* q, const void* k, const void* v, void* out, int dt, long long ld, Lay lay, int n_head, int C,
                            int w, int streams, cudaStream_t st) {
    const int hs = C / n_head;
    if (ld == C && lay.R % 128 == 0 && (hs == 64 || hs == 128)) {
        if (dt == VRD_BF16)
            return hs == 64 ? window_attn_tma_launch<__nv_bfloat16, 64, 32>(q, k, v, out, lay, w, streams, st)
                            : window_attn_tma_launch<__nv_bfloat16, 128, 32>(q, k, v, out, lay, w, streams, st);
        return hs == 64 ? window_attn_tma_launch<float, 64, 16>(q, k, v, out, lay, w, streams, st)
                        : window_attn_tma_launch<float, 128, 16>(q, k, v, out, lay, w, streams, st);
    }
    const int grid = (streams * lay.R + WARPS - 1) / WARPS;
#define LAUNCH(T, LPH) \
    launch_k(window_attn_kernel<T, 4, LPH>, dim3(grid), dim3(WARPS * 32), 0, st, (const T*)q, (const T*)k, (const T*)v, (T*)out, ld, lay, w, streams)
    if (hs == 64) { if (dt == VRD_BF16) LAUNCH(__nv_bfloat16, 16); else LAUNCH(float, 16); }
    else if (hs == 128) { if (dt == VRD_BF16) LAUNCH(__nv_bfloat16, 32); else LAUNCH(float, 32); }
    else return 1;
#undef LAUNCH
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// full_attn: flash-style softmax(QK^T)V inside each pair, fp32 math on CUDA cores (round-1 baseline kernel).
// A group of TPQ = HS/32 threads owns one query; thread `part` holds dims  d*TPQ + part.
// grid = (query tiles, heads, pairs)
// ------------------------------------------------------------------------------------------------------------
template <typename T, int HS>
__global__ void __launch_bounds__(128) full_attn_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                        const T* __restrict__ v, T* __restrict__ out, long long ld, Lay lay) {
    pdl_wait();
    constexpr int TPQ = HS / 32;
    constexpr int QPB = 128 / TPQ;
    constexpr int KT = 32;
    __shared__ float Ks[KT][HS];
    __shared__ float Vs[KT][HS];
    const int seq = blockIdx.z, head = blockIdx.y;
    const int4 si = lay.seqinfo[seq];
    const int len = si.y;
    const int q0 = blockIdx.x * QPB;
    if (q0 >= len) return;
    const int qi = q0 + threadIdx.x / TPQ;
    const int part = threadIdx.x % TPQ;
    const bool qok = qi < len;
    const long long hoff = (long long)head * HS;
    float qv[32], acc[32];
    {
        const T* qr = q + (long long)(si.x + (qok ? qi : 0)) * ld + hoff;
#pragma unroll
        for (int d = 0; d < 32; ++d) { qv[d] = to_f(qr[d * TPQ + part]); acc[d] = 0.f; }
    }
    float m = -INFINITY, l = 0.f;
    for (int k0 = 0; k0 < len; k0 += KT) {
        const int nk = min(KT, len - k0);
        __syncthreads();
        for (int i = threadIdx.x; i < KT * HS; i += 128) {
            const int kk = i / HS, d = i % HS;
            float kvv = 0.f, vvv = 0.f;
            if (kk < nk) {
                const long long row = (long long)(si.x + k0 + kk) * ld + hoff + d;
                kvv = to_f(k[row]);
                vvv = to_f(v[row]);
            }
            Ks[kk][d] = kvv;
            Vs[kk][d] = vvv;
        }
        __syncthreads();
        for (int kk = 0; kk < nk; ++kk) {
            float sdot = 0.f;
#pragma unroll
            for (int d = 0; d < 32; ++d) sdot = fmaf(qv[d], Ks[kk][d * TPQ + part], sdot);
#pragma unroll
            for (int off = TPQ / 2; off > 0; off >>= 1) sdot += __shfl_xor_sync(FULL_MASK, sdot, off);
            const float mn = fmaxf(m, sdot);
            const float corr = expf(m - mn);      // exp(-inf) = 0 on the first key
            const float p = expf(sdot - mn);
            l = l * corr + p;
#pragma unroll
            for (int d = 0; d < 32; ++d) acc[d] = fmaf(p, Vs[kk][d * TPQ + part], acc[d] * corr);
            m = mn;
        }
    }
    if (qok) {
        const float inv = 1.0f / l;
        T* orow = out + (long long)(si.x + qi) * ld + hoff;
#pragma unroll
        for (int d = 0; d < 32; ++d) orow[d * TPQ + part] = from_f<T>(acc[d] * inv);
    }
}

// ------------------------------------------------------------------------------------------------------------
// flash_attn_bf16: tensor-core (mma.sync m16n8k16 bf16, fp32 accumulate) flash attention for the bf16 path.
// Block = 4 warps x 16 query rows; KV tiles of 64 keys staged in shared memory (rows padded by 16 B so that ldmatrix is
// conflict-free); online softmax in registers with exp2.  grid = (query tiles, heads, pairs).
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_trans(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// 2^x on the SFU without exp2f()'s range fix-ups (inputs here are <= 0 or -inf: ex2.approx(-inf) = +0, no overflow)
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, bool pred) {
    // 16-byte asynchronous global->shared copy; pred == false writes zeros (src-size 0), nothing is read
    const int sz = pred ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// NBUF = 2: K/V tiles double-buffered with cp.async (the next tile streams in while the tensor cores work on the current
// one; Q, K0 and V0 arrive together), NBUF = 1 for head_dim 128 (shared-memory budget).
template <int HS, int NBUF>
__global__ void __launch_bounds__(128, HS == 64 ? 4 : 1) flash_attn_bf16_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                              const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ out,
                                                              long long ld, Lay lay) {
    pdl_wait();
    constexpr int BM = 64, BN = 64, LDS = HS + 8, KSTEPS = HS / 16, DBLK = HS / 8, CPR = HS / 8;   // CPR: 16-byte chunks per row
    extern __shared__ __align__(16) uint8_t fa_smem[];
    __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(fa_smem);
    __nv_bfloat16* sK = sQ + BM * LDS;                 // [NBUF][BN * LDS]
    __nv_bfloat16* sV = sK + NBUF * BN * LDS;          // [NBUF][BN * LDS]
    const int head = blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t4 = lane & 3;
    const long long hoff = (long long)head * HS;
    const uint32_t sQ_u = (uint32_t)__cvta_generic_to_shared(sQ);
    const uint32_t sK_u = (uint32_t)__cvta_generic_to_shared(sK), sV_u = (uint32_t)__cvta_generic_to_shared(sV);
    // The CTA owns the query tiles (64 rows of one pair, aligned to the pair's first row) that START in rows
    // [64 b, 64 b + 64) of the layout: one for the inside of a long pair, one more for every pair that begins there.  (A grid
    // of (longest pair / 64, heads, pairs) launched 52 k CTAs per call of which two thirds exited at once.)
    __shared__ int s_starts[BM];
    __shared__ int s_nstart;
    if (tid == 0) s_nstart = 0;
    __syncthreads();
    if (tid < BM) {
        const int r = blockIdx.x * BM + tid;
        const int sq = r < lay.R ? lay.row_seq[r] : -1;
        if (sq >= 0 && ((r - lay.seqinfo[sq].x) & (BM - 1)) == 0) s_starts[atomicAdd(&s_nstart, 1)] = r;
    }
    __syncthreads();
    const int n_start = s_nstart;
  for (int ui = 0; ui < n_start; ++ui) {
    // order of the list does not matter: tiles are independent and each writes its own rows
    const int u_row = s_starts[ui];
    const int seq = lay.row_seq[u_row];
    const int4 si = lay.seqinfo[seq];
    const int len = si.y;
    const int q0 = u_row - si.x;

    // every thread copies the same 16-byte column chunk of rows lr, lr + RPP, ...: one pointer per operand, advanced by a
    // constant stride (the generic c / CPR, c % CPR form spent ~15 % of the kernel's instructions on 64-bit address math)
    constexpr int RPP = 128 / CPR;                     // rows covered by one pass of the 128 threads
    const int lr = tid / CPR, lc = tid % CPR;
    const long long row_off = (long long)(si.x + lr) * ld + hoff + lc * 8;
    const __nv_bfloat16* kp = k + row_off;
    const __nv_bfloat16* vp = v + row_off;
    const uint32_t sm_off = (uint32_t)((lr * LDS + lc * 8) * 2);
    auto load_kv = [&](int k0, int buf) {
        const uint32_t d0 = sm_off + (uint32_t)(buf * BN * LDS * 2);
#pragma unroll
        for (int i = 0; i < BN / RPP; ++i) {
            const int r = k0 + lr + i * RPP;
            const bool ok = r < len;
            const long long off = ok ? (long long)(k0 + i * RPP) * ld : 0;
            cp_async16(sK_u + d0 + i * RPP * LDS * 2, kp + off, ok);
            cp_async16(sV_u + d0 + i * RPP * LDS * 2, vp + off, ok);     // rows past the pair are zero-filled: P = 0 there, and 0 * garbage could be NaN
        }
    };
#pragma unroll
    for (int i = 0; i < BM / RPP; ++i) {
        const bool ok = q0 + lr + i * RPP < len;
        cp_async16(sQ_u + sm_off + i * RPP * LDS * 2, q + row_off + (ok ? (long long)(q0 + i * RPP) * ld : 0), ok);
    }
    load_kv(0, 0);
    cp_async_commit();

    uint32_t qf[KSTEPS][4];
    float o[DBLK][4];
#pragma unroll
    for (int d = 0; d < DBLK; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    constexpr float LOG2E = 1.4426950408889634f;

    for (int k0 = 0, kt = 0; k0 < len; k0 += BN, ++kt) {
        const int buf = (NBUF == 2) ? (kt & 1) : 0;
        if (NBUF == 2 && k0 + BN < len) {
            load_kv(k0 + BN, buf ^ 1);             // the buffer consumed in iteration kt - 1 (all warps passed its trailing sync)
            cp_async_commit();
            cp_async_wait<1>();
        } else {
            cp_async_wait<0>();
        }
        __syncthreads();
        if (kt == 0) {
#pragma unroll
            for (int ks = 0; ks < KSTEPS; ++ks) {
                const int row = warp * 16 + (lane & 15), col = ks * 16 + (lane >> 4) * 8;
                ldsm_x4(sQ_u + (row * LDS + col) * 2, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
            }
        }
        const uint32_t bK = sK_u + buf * BN * LDS * 2, bV = sV_u + buf * BN * LDS * 2;
        float sacc[BN / 8][4];
#pragma unroll
        for (int nb = 0; nb < BN / 8; ++nb) sacc[nb][0] = sacc[nb][1] = sacc[nb][2] = sacc[nb][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < KSTEPS; ++ks) {
#pragma unroll
            for (int nb = 0; nb < BN / 8; nb += 2) {
                // matrices: (keys nb*8.., dims ks*16..+7), (same keys, dims +8), (keys (nb+1)*8.., dims ..+7), (.., dims +8)
                const int row = nb * 8 + (lane & 7) + ((lane >> 4) << 3), col = ks * 16 + ((lane >> 3) & 1) * 8;
                uint32_t b0, b1, b2, b3;
                ldsm_x4(bK + (row * LDS + col) * 2, b0, b1, b2, b3);
                mma_bf16(sacc[nb], qf[ks], b0, b1);
                mma_bf16(sacc[nb + 1], qf[ks], b2, b3);
            }
        }
        if (k0 + BN > len) {   // mask the keys past the end of the pair
#pragma unroll
            for (int nb = 0; nb < BN / 8; ++nb) {
                const int c = k0 + nb * 8 + t4 * 2;
                if (c >= len) sacc[nb][0] = sacc[nb][2] = -INFINITY;
                if (c + 1 >= len) sacc[nb][1] = sacc[nb][3] = -INFINITY;
            }
        }
        float mx0 = m0, mx1 = m1;
#pragma unroll
        for (int nb = 0; nb < BN / 8; ++nb) {
            mx0 = fmaxf(mx0, fmaxf(sacc[nb][0], sacc[nb][1]));
            mx1 = fmaxf(mx1, fmaxf(sacc[nb][2], sacc[nb][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(FULL_MASK, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(FULL_MASK, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(FULL_MASK, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(FULL_MASK, mx1, 2));
        const float a0 = fast_exp2((m0 - mx0) * LOG2E), a1 = fast_exp2((m1 - mx1) * LOG2E);
        m0 = mx0; m1 = mx1;
        const float ms0 = mx0 * LOG2E, ms1 = mx1 * LOG2E;
        l0 *= a0; l1 *= a1;
#pragma unroll
        for (int d = 0; d < DBLK; ++d) { o[d][0] *= a0; o[d][1] *= a0; o[d][2] *= a1; o[d][3] *= a1; }
        uint32_t pf[BN / 16][4];
#pragma unroll
        for (int nb = 0; nb < BN / 8; ++nb) {
            const float p0 = fast_exp2(fmaf(sacc[nb][0], LOG2E, -ms0)), p1 = fast_exp2(fmaf(sacc[nb][1], LOG2E, -ms0));
            const float p2 = fast_exp2(fmaf(sacc[nb][2], LOG2E, -ms1)), p3 = fast_exp2(fmaf(sacc[nb][3], LOG2E, -ms1));
            l0 += p0 + p1; l1 += p2 + p3;
            pf[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16(p0, p1);
            pf[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16(p2, p3);
        }
#pragma unroll
        for (int ks = 0; ks < BN / 16; ++ks) {
#pragma unroll
            for (int d = 0; d < DBLK; d += 2) {
                // transposed 8x8 blocks: (keys ks*16..+7, dims d*8..), (keys +8, same dims), (keys .., dims (d+1)*8..), (keys +8, ..)
                const int row = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, col = d * 8 + (lane >> 4) * 8;
                uint32_t b0, b1, b2, b3;
                ldsm_x4_trans(bV + (row * LDS + col) * 2, b0, b1, b2, b3);
                mma_bf16(o[d], pf[ks], b0, b1);
                mma_bf16(o[d + 1], pf[ks], b2, b3);
            }
        }
        __syncthreads();       // every warp is done with this K/V buffer before it is refilled
        if (NBUF == 1 && k0 + BN < len) {
            load_kv(k0 + BN, 0);
            cp_async_commit();
        }
    }
    l0 += __shfl_xor_sync(FULL_MASK, l0, 1); l0 += __shfl_xor_sync(FULL_MASK, l0, 2);
    l1 += __shfl_xor_sync(FULL_MASK, l1, 1); l1 += __shfl_xor_sync(FULL_MASK, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
#pragma unroll
    for (int d = 0; d < DBLK; ++d) {
        const int c = d * 8 + t4 * 2;
        if (r0 < len) *reinterpret_cast<uint32_t*>(out + (long long)(si.x + r0) * ld + hoff + c) = pack_bf16(o[d][0] * i0, o[d][1] * i0);
        if (r1 < len) *reinterpret_cast<uint32_t*>(out + (long long)(si.x + r1) * ld + hoff + c) = pack_bf16(o[d][2] * i1, o[d][3] * i1);
    }
    __syncthreads();           // the next tile of this CTA refills sQ / sK / sV
  }
}

// Separator rows of the output are NOT written: the only consumer is the output-projection GEMM (taps == 1, with a layout),
// whose epilogue overwrites its own separator rows with zeros whatever the operand row holds.
template <int HS, int NBUF>
static int flash_launch(const void* q, const void* k, const void* v, void* out, long long ld, Lay lay, int n_head, int max_rows,
                        cudaStream_t st) {
    constexpr int smem = (1 + 2 * NBUF) * 64 * (HS + 8) * 2;
    static PerDeviceOnce once;
    auto kern = flash_attn_bf16_kernel<HS, NBUF>;
    if (once.first() && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
    (void)max_rows;
    const dim3 grid((lay.R + 63) / 64, n_head);
    launch_k(kern, dim3(grid), dim3(128), smem, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v, (__nv_bfloat16*)out, ld, lay);
    return 0;
}

int full_attn(const void* q, const void* k, const void* v, void* out, int dt, long long ld, Lay lay, int n_head, int C,
              int max_len, cudaStream_t st) {
    const int hs = C / n_head;
    if (lay.B > 65535 || max_len < 1) return 1;
    const int max_rows = max_len;   // longest pair of the level bounds the number of query tiles (blocks past a pair's end exit)
    if (dt == VRD_BF16) {
        // head_dim 64: the tcgen05 / TMEM kernel (attention_tc.cu); VRD_FLASH=mma keeps the mma.sync kernel for A/B measurements
        static const bool use_tc = !(getenv("VRD_FLASH") != nullptr && strcmp(getenv("VRD_FLASH"), "mma") == 0);
        if (hs == 64 && use_tc) {
            const int rc = full_attn_tcgen05(q, k, v, out, ld, lay, n_head, C, st);
            if (rc != 1) return rc;
        }
        if (hs == 64) return flash_launch<64, 2>(q, k, v, out, ld, lay, n_head, max_rows, st);
        if (hs == 128) return flash_launch<128, 1>(q, k, v, out, ld, lay, n_head, max_rows, st);
        return 1;
    }
#define LAUNCH(T, HS) \
    launch_k(full_attn_kernel<T, HS>, dim3(dim3((max_rows + (128 / (HS / 32)) - 1) / (128 / (HS / 32)), n_head, lay.B)), dim3(128), 0, st,  \
        (const T*)q, (const T*)k, (const T*)v, (T*)out, ld, lay)
    if (hs == 64) LAUNCH(float, 64);
    else if (hs == 128) LAUNCH(float, 128);
    else return 1;
#undef LAUNCH
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// window_attn_mma: the sliding-window attention (reference blocks.py:746-989, local_transformer.py:553-623) on tensor cores
// for the bf16 path, head_dim 64.  The unit of work is 16 consecutive query rows of ONE pair, aligned to the pair's first row
// (so a pair's result does not depend on where its rows sit in the layout), handled by one warp with no CTA-level barrier:
// per head the warp stages its 16 query rows and the 32 key / value rows [u0 - 8, u0 + 24) with cp.async (two stages: the next
// head streams in while this one is computed), multiplies Q K^T (mma.sync m16n8k16, 16 MMAs), masks everything outside the band
// |i - j| <= w and outside the pair, takes ONE softmax (no online rescaling: the band fits the key block) and multiplies with
// V (16 MMAs): ~25 instructions per row and head instead of the ~100 of the CUDA-core kernel (9 keys x (16-byte load, 8 FMAs,
// 3 shuffles) per 8 channels).  Warp b owns the units that START in rows [16 b, 16 b + 16) of the stacked layout: one for the
// inside of a long pair, one more for every pair that begins there.  The output tile goes through the warp's query rows in
// shared memory so that global stores are 16-byte vectors.  Separator rows are not written (as in flash_attn_bf16: the only
// consumer is the projection GEMM, whose epilogue zeroes its own separator rows whatever the operand holds).
// ------------------------------------------------------------------------------------------------------------
constexpr int WM_HALO = 8, WM_WARPS = 4;

template <int NH>
__global__ void __launch_bounds__(WM_WARPS * 32, 2) window_attn_mma_kernel(const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k,
                                                                           const __nv_bfloat16* __restrict__ v, __nv_bfloat16* __restrict__ out,
                                                                           Lay lay, int w, int total_rows) {
    pdl_wait();
    constexpr int HS = 64, BM = 16, KR = BM + 2 * WM_HALO, LDS = HS + 8, CPR = HS / 8, C = NH * HS;
    constexpr int STAGE = (BM + 2 * KR) * LDS;          // elements per pipeline stage: Q, K, V
    extern __shared__ __align__(16) uint8_t wm_smem[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane >> 2, t4 = lane & 3;
    __nv_bfloat16* smem_p = reinterpret_cast<__nv_bfloat16*>(wm_smem) + warp * 2 * STAGE;
    const uint32_t smem_u = (uint32_t)__cvta_generic_to_shared(smem_p);
    const int b0 = (blockIdx.x * WM_WARPS + warp) * BM;      // first row of this warp's block of the stacked layout
    if (b0 >= total_rows) return;
    constexpr float LOG2E = 1.4426950408889634f;

    // units that start in this block: row r of pair p with (r - first row of p) % 16 == 0
    int my_first = 0, my_last = -1;
    {
        const int i = b0 + (lane & 15);
        const int s = i / lay.R, r = i - s * lay.R;
        const int seq = lay.row_seq[r];
        if (seq >= 0) {
            const int4 si = lay.seqinfo[seq];
            my_first = s * lay.R + si.x;
            my_last = my_first + si.y - 1;
        }
        const bool starts = lane < 16 && my_last >= 0 && ((i - my_first) & 15) == 0;
        uint32_t units = __ballot_sync(FULL_MASK, starts);
        while (units != 0) {                              // warp-uniform
            const int ul = __ffs(units) - 1;
            units &= units - 1;
            const int u0 = b0 + ul;
            const int first = __shfl_sync(FULL_MASK, my_first, ul), last = __shfl_sync(FULL_MASK, my_last, ul);
            // band limits of this thread's two query rows (rows past the pair's end: a finite dummy softmax, never written)
            int lo[2], hi[2];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const int qi = u0 + g + 8 * h;
                lo[h] = hi[h] = qi;
                if (qi <= last) { lo[h] = max(qi - w, first); hi[h] = min(qi + w, last); }
            }
            const int kb = u0 - WM_HALO;                  // row of the first staged key
            auto load = [&](int head, int stage) {
                const uint32_t sQ = smem_u + (uint32_t)(stage * STAGE) * 2, sK = sQ + BM * LDS * 2, sV = sK + KR * LDS * 2;
                const long long hoff = (long long)head * HS;
#pragma unroll
                for (int c = lane; c < BM * CPR; c += 32) {
                    const int r = c / CPR, cc = c % CPR;
                    const bool ok = u0 + r < total_rows;
                    cp_async16(sQ + (uint32_t)((r * LDS + cc * 8) * 2), q + (long long)(ok ? u0 + r : 0) * C + hoff + cc * 8, ok);
                }
#pragma unroll
                for (int c = lane; c < KR * CPR; c += 32) {
                    const int r = c / CPR, cc = c % CPR;
                    const int pr = kb + r;
                    const bool ok = pr >= 0 && pr < total_rows;
                    const long long off = (long long)(ok ? pr : 0) * C + hoff + cc * 8;
                    const uint32_t d = (uint32_t)((r * LDS + cc * 8) * 2);
                    cp_async16(sK + d, k + off, ok);
                    cp_async16(sV + d, v + off, ok);
                }
            };
            load(0, 0);
            cp_async_commit();
            for (int head = 0; head < NH; ++head) {
                const int stage = head & 1;
                if (head + 1 < NH) {
                    load(head + 1, stage ^ 1);            // its last readers finished before the previous iteration's __syncwarp
                    cp_async_commit();
                    cp_async_wait<1>();
                } else {
                    cp_async_wait<0>();
                }
                __syncwarp();
                const uint32_t sQ = smem_u + (uint32_t)(stage * STAGE) * 2, sK = sQ + BM * LDS * 2, sV = sK + KR * LDS * 2;
                uint32_t qf[HS / 16][4];
#pragma unroll
                for (int ks = 0; ks < HS / 16; ++ks) {
                    const int row = lane & 15, col = ks * 16 + (lane >> 4) * 8;
                    ldsm_x4(sQ + (row * LDS + col) * 2, qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
                }
                float sacc[4][4];
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) sacc[nb][0] = sacc[nb][1] = sacc[nb][2] = sacc[nb][3] = 0.f;
#pragma unroll
                for (int ks = 0; ks < HS / 16; ++ks) {
#pragma unroll
                    for (int nb = 0; nb < 4; nb += 2) {
                        const int row = nb * 8 + (lane & 7) + ((lane >> 4) << 3), col = ks * 16 + ((lane >> 3) & 1) * 8;
                        uint32_t x0, x1, x2, x3;
                        ldsm_x4(sK + (row * LDS + col) * 2, x0, x1, x2, x3);
                        mma_bf16(sacc[nb], qf[ks], x0, x1);
                        mma_bf16(sacc[nb + 1], qf[ks], x2, x3);
                    }
                }
                float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    const int c = kb + nb * 8 + t4 * 2;
                    if (c < lo[0] || c > hi[0]) sacc[nb][0] = -INFINITY;
                    if (c + 1 < lo[0] || c + 1 > hi[0]) sacc[nb][1] = -INFINITY;
                    if (c < lo[1] || c > hi[1]) sacc[nb][2] = -INFINITY;
                    if (c + 1 < lo[1] || c + 1 > hi[1]) sacc[nb][3] = -INFINITY;
                    mx0 = fmaxf(mx0, fmaxf(sacc[nb][0], sacc[nb][1]));
                    mx1 = fmaxf(mx1, fmaxf(sacc[nb][2], sacc[nb][3]));
                }
                mx0 = fmaxf(mx0, __shfl_xor_sync(FULL_MASK, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(FULL_MASK, mx0, 2));
                mx1 = fmaxf(mx1, __shfl_xor_sync(FULL_MASK, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(FULL_MASK, mx1, 2));
                const float ms0 = mx0 * LOG2E, ms1 = mx1 * LOG2E;   // finite: the diagonal key is always inside the band
                float l0 = 0.f, l1 = 0.f;
                uint32_t pf[2][4];
#pragma unroll
                for (int nb = 0; nb < 4; ++nb) {
                    const float p0 = fast_exp2(fmaf(sacc[nb][0], LOG2E, -ms0)), p1 = fast_exp2(fmaf(sacc[nb][1], LOG2E, -ms0));
                    const float p2 = fast_exp2(fmaf(sacc[nb][2], LOG2E, -ms1)), p3 = fast_exp2(fmaf(sacc[nb][3], LOG2E, -ms1));
                    l0 += p0 + p1; l1 += p2 + p3;
                    pf[nb >> 1][(nb & 1) * 2 + 0] = pack_bf16(p0, p1);
                    pf[nb >> 1][(nb & 1) * 2 + 1] = pack_bf16(p2, p3);
                }
                float o[HS / 8][4];
#pragma unroll
                for (int d = 0; d < HS / 8; ++d) o[d][0] = o[d][1] = o[d][2] = o[d][3] = 0.f;
#pragma unroll
                for (int ks = 0; ks < 2; ++ks) {
#pragma unroll
                    for (int d = 0; d < HS / 8; d += 2) {
                        const int row = ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8, col = d * 8 + (lane >> 4) * 8;
                        uint32_t x0, x1, x2, x3;
                        ldsm_x4_trans(sV + (row * LDS + col) * 2, x0, x1, x2, x3);
                        mma_bf16(o[d], pf[ks], x0, x1);
                        mma_bf16(o[d + 1], pf[ks], x2, x3);
                    }
                }
                l0 += __shfl_xor_sync(FULL_MASK, l0, 1); l0 += __shfl_xor_sync(FULL_MASK, l0, 2);
                l1 += __shfl_xor_sync(FULL_MASK, l1, 1); l1 += __shfl_xor_sync(FULL_MASK, l1, 2);
                const float i0 = 1.0f / l0, i1 = 1.0f / l1;
                // the query rows (their fragments are in registers) stage the output tile
                __nv_bfloat16* sO = smem_p + stage * STAGE;
                __syncwarp();
#pragma unroll
                for (int d = 0; d < HS / 8; ++d) {
                    const int c = d * 8 + t4 * 2;
                    *reinterpret_cast<uint32_t*>(sO + g * LDS + c) = pack_bf16(o[d][0] * i0, o[d][1] * i0);
                    *reinterpret_cast<uint32_t*>(sO + (g + 8) * LDS + c) = pack_bf16(o[d][2] * i1, o[d][3] * i1);
                }
                __syncwarp();
#pragma unroll
                for (int c = lane; c < BM * CPR; c += 32) {
                    const int r = c / CPR, cc = c % CPR;
                    if (u0 + r <= last) {
                        const uint4 val = *reinterpret_cast<const uint4*>(sO + r * LDS + cc * 8);
                        *reinterpret_cast<uint4*>(out + (long long)(u0 + r) * C + head * HS + cc * 8) = val;
                    }
                }
                __syncwarp();      // this stage may be refilled by the next iteration's prefetch
            }
        }
    }
}

static int window_attn_mma_launch(const void* q, const void* k, const void* v, void* out, Lay lay, int w, int streams,
                                  cudaStream_t st) {
    constexpr int smem = WM_WARPS * 2 * (16 + 2 * (16 + 2 * WM_HALO)) * (64 + 8) * 2;
    static PerDeviceOnce once;
    auto kern = window_attn_mma_kernel<8>;
    if (once.first() && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 1;
    const int total = streams * lay.R;    // R % 128 == 0
    launch_k(kern, dim3(total / (16 * WM_WARPS)), dim3(WM_WARPS * 32), smem, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)v,
                                                             (__nv_bfloat16*)out, lay, w, total);
    return 0;
}

int window_attn(const void* q, const void* k, const void* v, void* out, int dt, long long ld, Lay lay, int n_head, int C, int w,
                int streams, cudaStream_t st) {
    if (C != 512 || w > 4 || w < 1) return 1;
    const int hs = C / n_head;
    static const bool use_mma = !(getenv("VRD_WINATTN") != nullptr && strcmp(getenv("VRD_WINATTN"), "simt") == 0);
    if (use_mma && dt == VRD_BF16 && hs == 64 && n_head == 8 && ld == C && lay.R % 128 == 0)
        return window_attn_mma_launch(q, k, v, out, lay, w, streams, st);
    return window_attn_simt(q, k, v, out, dt, ld, lay, n_head, C, w, streams, st);
}

// ------------------------------------------------------------------------------------------------------------
// query_self_attn: attention among the Q (<= 12) queries of each pair.  One block per pair, D = 256.
// ------------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) query_self_attn_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                              const T* __restrict__ v, T* __restrict__ out, long long ld,
                                                              int Q, int n_head) {
    pdl_wait();
    constexpr int D = 256, MAXQ = 12, MAXH = 8;
    // sk rows are padded by four words: in the score phase consecutive threads read the SAME column of consecutive key rows
    // (a 1 KB row pitch would put all of them on one bank)
    __shared__ float sq[MAXQ][D], sk[MAXQ][D + 4], sv[MAXQ][D];
    __shared__ float sp[MAXH][MAXQ][MAXQ];
    const int pair = blockIdx.x;
    const int hs = D / n_head;
    const long long r0 = (long long)pair * Q;
    for (int i = threadIdx.x; i < Q * D; i += 256) {
        const int qi = i / D, d = i % D;
        sq[qi][d] = to_f(q[(r0 + qi) * ld + d]);
        sk[qi][d] = to_f(k[(r0 + qi) * ld + d]);
        sv[qi][d] = to_f(v[(r0 + qi) * ld + d]);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_head * Q * Q; i += 256) {
        const int h = i / (Q * Q), qi = (i / Q) % Q, kj = i % Q;
        float s = 0.f;
        for (int d = 0; d < hs; ++d) s = fmaf(sq[qi][h * hs + d], sk[kj][h * hs + d], s);
        sp[h][qi][kj] = s;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n_head * Q; i += 256) {
        const int h = i / Q, qi = i % Q;
        float mx = -INFINITY;
        for (int j = 0; j < Q; ++j) mx = fmaxf(mx, sp[h][qi][j]);
        float sum = 0.f;
        for (int j = 0; j < Q; ++j) { const float p = expf(sp[h][qi][j] - mx); sp[h][qi][j] = p; sum += p; }
        const float inv = 1.0f / sum;
        for (int j = 0; j < Q; ++j) sp[h][qi][j] *= inv;
    }
    __syncthreads();
    const int d = threadIdx.x, h = d / hs;
    for (int qi = 0; qi < Q; ++qi) {
        float acc = 0.f;
        for (int j = 0; j < Q; ++j) acc = fmaf(sp[h][qi][j], sv[j][d], acc);
        out[(r0 + qi) * ld + d] = from_f<T>(acc);
    }
}

int query_self_attn(const void* q, const void* k, const void* v, void* out, int dt, long long ld, int B, int Q, int n_head, int C,
                    cudaStream_t st) {
    if (C != 256 || Q > 12 || n_head > 8 || B < 1) return 1;
    if (dt == VRD_BF16)
        launch_k(query_self_attn_kernel<__nv_bfloat16>, dim3(B), dim3(256), 0, st, (const __nv_bfloat16*)q, (const __nv_bfloat16*)k,
                                                                 (const __nv_bfloat16*)v, (__nv_bfloat16*)out, ld, Q, n_head);
    else
        launch_k(query_self_attn_kernel<float>, dim3(B), dim3(256), 0, st, (const float*)q, (const float*)k, (const float*)v, (float*)out, ld, Q, n_head);
    return 0;
}

// ------------------------------------------------------------------------------------------------------------
// query_cross_attn: Q queries of a pair attend to the pair's rows of the coarsest level.  One block per pair; a warp owns
// one HEAD and works on all Q queries of the pair at once: lanes = keys for the scores (a key's row is loaded once and used by
// every query: Q-way instruction-level parallelism instead of Q dependent passes over the same rows), lanes = head dims for
// P.V (a value row is loaded once per key, the Q probabilities arrive by shuffle).  Online softmax over tiles of 32 keys.
// ------------------------------------------------------------------------------------------------------------
template <typename T, int HS>
__global__ void __launch_bounds__(256) query_cross_attn_kernel(const T* __restrict__ q, const T* __restrict__ k,
                                                               const T* __restrict__ v, T* __restrict__ out, long long ld,
                                                               Lay lay, int Q, int n_head) {
    pdl_wait();
    constexpr int D = 256, MAXQ = 12, DPL = HS / 32;
    __shared__ __align__(16) float sq[MAXQ][D];
    const int pair = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int4 si = lay.seqinfo[pair];
    const long long r0 = (long long)pair * Q;
    for (int i = threadIdx.x; i < MAXQ * D; i += 256) sq[i / D][i % D] = (i / D < Q) ? to_f(q[(r0 + i / D) * ld + (i % D)]) : 0.f;
    __syncthreads();
    for (int h = warp; h < n_head; h += 8) {
        float m[MAXQ], l[MAXQ], acc[MAXQ][DPL];
#pragma unroll
        for (int qi = 0; qi < MAXQ; ++qi) {
            m[qi] = -INFINITY; l[qi] = 0.f;
#pragma unroll
            for (int i = 0; i < DPL; ++i) acc[qi][i] = 0.f;
        }
        for (int k0 = 0; k0 < si.y; k0 += 32) {
            const int kj = k0 + lane;
            const bool ok = kj < si.y;
            float s[MAXQ];
#pragma unroll
            for (int qi = 0; qi < MAXQ; ++qi) s[qi] = 0.f;
            if (ok) {
                const T* kr = k + (long long)(si.x + kj) * ld + h * HS;
#pragma unroll
                for (int c = 0; c < HS; c += 4) {
                    float kv[4];
                    ld4(kr + c, kv);
#pragma unroll
                    for (int qi = 0; qi < MAXQ; ++qi) {
                        const float4 qv = *reinterpret_cast<const float4*>(&sq[qi][h * HS + c]);   // same address for all lanes
                        s[qi] = fmaf(qv.x, kv[0], fmaf(qv.y, kv[1], fmaf(qv.z, kv[2], fmaf(qv.w, kv[3], s[qi]))));
                    }
                }
            }
            float p[MAXQ];
#pragma unroll
            for (int qi = 0; qi < MAXQ; ++qi) {
                const float sv = ok ? s[qi] : -INFINITY;
                const float mn = fmaxf(m[qi], warp_max(sv));
                const float corr = expf(m[qi] - mn);       // exp(-inf) = 0 on the first tile
                p[qi] = ok ? expf(sv - mn) : 0.f;
                l[qi] = l[qi] * corr + warp_sum(p[qi]);
#pragma unroll
                for (int i = 0; i < DPL; ++i) acc[qi][i] *= corr;
                m[qi] = mn;
            }
            const int nk = min(32, si.y - k0);
            for (int jj = 0; jj < nk; ++jj) {
                const T* vr = v + (long long)(si.x + k0 + jj) * ld + h * HS;
                float vv[DPL];
#pragma unroll
                for (int i = 0; i < DPL; ++i) vv[i] = to_f(vr[lane + 32 * i]);
#pragma unroll
                for (int qi = 0; qi < MAXQ; ++qi) {
                    const float pj = __shfl_sync(FULL_MASK, p[qi], jj);
#pragma unroll
                    for (int i = 0; i < DPL; ++i) acc[qi][i] = fmaf(pj, vv[i], acc[qi][i]);
                }
            }
        }
#pragma unroll
        for (int qi = 0; qi < MAXQ; ++qi) {
            if (qi < Q) {
                const float inv = 1.0f / l[qi];
#pragma unroll
                for (int i = 0; i < DPL; ++i) out[(r0 + qi) * ld + h * HS + lane + 32 * i] = from_f<T>(acc[qi][i] * inv);
            }
        }
    }
}

int query_cross_attn(const void* q, const void* k, const void* v, void* out, int dt, long long ld, Lay lay, int Q, int n_head,
                     int C, cudaStream_t st) {
    if (C != 256 || Q > 12) return 1;
    const int hs = C / n_head;
#define LAUNCH(T, HS) \
    launch_k(query_cross_attn_kernel<T, HS>, dim3(lay.B), dim3(256), 0, st, (const T*)q, (const T*)k, (const T*)v, (T*)out, ld, lay, Q, n_head)
    if (hs == 32) { if (dt == VRD_BF16) LAUNCH(__nv_bfloat16, 32); else LAUNCH(float, 32); }
    else if (hs == 64) { if (dt == VRD_BF16) LAUNCH(__nv_bfloat16, 64); else LAUNCH(float, 64); }
    else return 1;
#undef LAUNCH
    return 0;
}

}  // namespace vrd
