// Shared device helpers for the vrdone_b200 kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <cstdlib>
#include <utility>

#define VRD_F32 0
#define VRD_BF16 1

#define VRD_EPS 1e-5f
#define FULL_MASK 0xffffffffu

// Per-device "done once" flags for host-side launchers: function attributes (max dynamic shared memory) and the SM count
// belong to a DEVICE, not to the process -- a process that runs engines on two GPUs must set them on both.
struct PerDeviceOnce {
    bool done[64] = {false};
    // true exactly once per device ordinal (and always for ordinals past the table: setting an attribute twice is harmless)
    bool first() {
        int dev = 0;
        cudaGetDevice(&dev);
        if (dev < 0 || dev >= 64) return true;
        if (done[dev]) return false;
        done[dev] = true;
        return true;
    }
};
inline int device_sm_count() {
    static int sms[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) { int n = 0; cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); return n; }
    if (sms[dev] == 0) cudaDeviceGetAttribute(&sms[dev], cudaDevAttrMultiProcessorCount, dev);
    return sms[dev];
}

// ---- programmatic dependent launch (PDL).  Kernels launched through launch_k() carry the programmatic-stream-serialization
// attribute and start with pdl_wait(): the NEXT kernel of the stream is launched (its CTAs become resident as SM resources free
// up and run their prologue: barrier init, TMEM allocation, parameter loads) when every CTA of this one has exited or called
// pdl_launch_dependents(); pdl_wait() blocks until the PREVIOUS kernel has completed and its memory is visible, and must precede
// the first access to anything an earlier kernel wrote.  Measured on the ~2400 launches per step of the forward (network-only
// time of four cfg2 videos, tools/ab_switch.py): releasing the dependents EARLY -- griddepcontrol.launch_dependents at the top
// of every kernel -- was 2-3 % SLOWER than plain stream order, releasing them when a persistent GEMM CTA starts its last tile
// 1-2 % slower, and no explicit release at all (the implicit one at CTA exit: only launch latency and the prologue overlap the
// previous kernel's tail) ~1 % faster.  No kernel calls pdl_launch_dependents(); it stays for experiments.
// Without the launch attribute (VRD_PDL=0) both instructions are no-ops.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
// Experiment switches shared by the launchers (defined in cabi.cu): initialised from the environment on first use
// (VRD_PDL, VRD_DW_CFG), changeable at run time through vrd_set_option() so that one process can A/B them on the same inputs.
struct VrdOptions {
    int pdl;        // 1: kernels are launched with programmatic stream serialization (default), 0: plain stream order
    int dw_cfg;     // dwconv_ln_tile variant (see rows.cu)
    int gemm_spec;  // 1: specialised tcgen05 GEMM epilogues (default), 0: the generic run-time-flag epilogue for every launch
    int embed_ln;   // 1: LayerNorm + ReLU of the embedding convs as the GEMM's epilogue on the bf16 path (default), 0: separate launch
    int proj_ln;    // 1: encoder blocks: attention projection + residual + LayerNorm (MLP input) as one launch (default), 0: two
};
VrdOptions& vrd_options();
inline bool pdl_enabled() { return vrd_options().pdl != 0; }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
}

// One pyramid level of the varlen row layout (see vrdone_b200/layout.py).
// seqinfo[i] = (first row, valid rows, first-pad-column-exists, 0); row_seq[r] = owning pair or -1.
struct Lay {
    const int* __restrict__ row_seq;
    const int4* __restrict__ seqinfo;
    int R;   // rows per stream (multiple of 128)
    int B;   // pairs
    const int4* __restrict__ tiles = nullptr;   // level 0, optional: query tiles of the full-attention kernel (layout.py), longest first
    int n_tiles = 0;
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(FULL_MASK, v, o));
    return v;
}

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T> __device__ __forceinline__ T from_f(float v);
template <> __device__ __forceinline__ float from_f<float>(float v) { return v; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// ---- 4-element chunk loads/stores (a lane owns chunks  c = j*32 + lane,  elements 4c .. 4c+3) ----
__device__ __forceinline__ void ld4(const float* p, float (&v)[4]) {
    float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
__device__ __forceinline__ void ld4(const __nv_bfloat16* p, float (&v)[4]) {
    uint2 t = *reinterpret_cast<const uint2*>(p);
    __nv_bfloat162 a = *reinterpret_cast<__nv_bfloat162*>(&t.x);
    __nv_bfloat162 b = *reinterpret_cast<__nv_bfloat162*>(&t.y);
    float2 fa = __bfloat1622float2(a), fb = __bfloat1622float2(b);
    v[0] = fa.x; v[1] = fa.y; v[2] = fb.x; v[3] = fb.y;
}
__device__ __forceinline__ void st4(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
__device__ __forceinline__ void st4(__nv_bfloat16* p, const float (&v)[4]) {
    __nv_bfloat162 a = __floats2bfloat162_rn(v[0], v[1]);
    __nv_bfloat162 b = __floats2bfloat162_rn(v[2], v[3]);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&a);
    t.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(p) = t;
}

// A full row of C = 128*NCH channels distributed over a warp: v[j][i] = row[(j*32 + lane)*4 + i].
template <typename T, int NCH>
__device__ __forceinline__ void load_row(const T* row, int lane, float (&v)[NCH][4]) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) ld4(row + (j * 32 + lane) * 4, v[j]);
}
template <typename T, int NCH>
__device__ __forceinline__ void store_row(T* row, int lane, const float (&v)[NCH][4]) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) st4(row + (j * 32 + lane) * 4, v[j]);
}
template <typename T, int NCH>
__device__ __forceinline__ void zero_row(T* row, int lane) {
    const float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int j = 0; j < NCH; ++j) st4(row + (j * 32 + lane) * 4, z);
}

// Channel LayerNorm statistics of a warp-distributed row (biased variance, two-pass as the reference).
template <int NCH>
__device__ __forceinline__ void row_stats(const float (&v)[NCH][4], float& mean, float& rstd) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) s += (v[j][0] + v[j][1]) + (v[j][2] + v[j][3]);
    mean = warp_sum(s) * (1.0f / (NCH * 128));
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { float d = v[j][i] - mean; q += d * d; }
    }
    float var = warp_sum(q) * (1.0f / (NCH * 128));
    rstd = 1.0f / sqrtf(var + VRD_EPS);
}

// v <- (v - mean) * rstd * gamma + beta
template <int NCH>
__device__ __forceinline__ void row_normalize(float (&v)[NCH][4], int lane, const float* __restrict__ gamma,
                                              const float* __restrict__ beta) {
    float mean, rstd;
    row_stats<NCH>(v, mean, rstd);
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        float g[4], b[4];
        ld4(gamma + (j * 32 + lane) * 4, g);
        ld4(beta + (j * 32 + lane) * 4, b);
#pragma unroll
        for (int i = 0; i < 4; ++i) v[j][i] = (v[j][i] - mean) * rstd * g[i] + b[i];
    }
}

// ---- packed fp32 pairs (Blackwell FFMA2 / FADD2 / FMUL2: two IEEE fp32 operations per issued instruction; results are bit-identical to
// the scalar fma.rn / add.rn / mul.rn) for the issue-bound row kernels.  An f2 holds (lo, hi) = two consecutive channels. ----
typedef unsigned long long f2;
__device__ __forceinline__ f2 f2_mul(f2 a, f2 b) { f2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 f2_add(f2 a, f2 b) { f2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 f2_sub(f2 a, f2 b) { f2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f2 f2_fma(f2 a, f2 b, f2 c) { f2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f2 f2_splat(float x) { f2 d; asm("mov.b64 %0, {%1, %1};" : "=l"(d) : "f"(x)); return d; }
__device__ __forceinline__ void f2_unpack(f2 a, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a)); }
__device__ __forceinline__ float f2_hsum(f2 a) { float lo, hi; f2_unpack(a, lo, hi); return lo + hi; }
// four consecutive channels = two f2
__device__ __forceinline__ void f2_lds(uint32_t addr, f2 (&v)[2]) {
    asm volatile("ld.shared.v2.b64 {%0, %1}, [%2];" : "=l"(v[0]), "=l"(v[1]) : "r"(addr));
}
__device__ __forceinline__ void f2_sts(uint32_t addr, const f2 (&v)[2]) {
    asm volatile("st.shared.v2.b64 [%0], {%1, %2};" ::"r"(addr), "l"(v[0]), "l"(v[1]) : "memory");
}
__device__ __forceinline__ void f2_ldg(const float* p, f2 (&v)[2]) {
    const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(p);
    v[0] = t.x; v[1] = t.y;
}
__device__ __forceinline__ void f2_stg(float* p, const f2 (&v)[2]) { *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(v[0], v[1]); }
__device__ __forceinline__ void f2_stg(__nv_bfloat16* p, const f2 (&v)[2]) {
    float a, b, c, d;
    f2_unpack(v[0], a, b);
    f2_unpack(v[1], c, d);
    __nv_bfloat162 x = __floats2bfloat162_rn(a, b), y = __floats2bfloat162_rn(c, d);
    uint2 t;
    t.x = *reinterpret_cast<uint32_t*>(&x);
    t.y = *reinterpret_cast<uint32_t*>(&y);
    *reinterpret_cast<uint2*>(p) = t;
}

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }
