// bf16 GEMM on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands staged by TMA) with the
// fused epilogue -- the arithmetic of the "bf16 path" for every dense contraction of the model (k=1 convs, the
// conv-as-GEMM embeddings, MLPs, projections and heads; reference: nn.Conv1d call sites in models/blocks.py:46, 85,
// 728-737, 1054-1060, local_transformer.py:133-142, fpns.py:199, predictor.py:73-83).
//
//   D[M,N] = A[M, taps*K] * W[N, taps*K]^T        A, W bf16 (K contiguous), fp32 accumulation
//
// For taps == 3 the K-slab d of output row r is A[r + d - 1, :]: the k=3 convolution over the token-major varlen layout
// is three row-shifted TMA loads of the same matrix (out-of-range rows are zero-filled by TMA, sequence boundaries by
// the zero separator rows of layout.py).
//
// Kernel structure (one persistent CTA per SM, 384 threads):
//   warp 0     TMA producer          (4-stage ring of {A 128x64, W BNx64} bf16 tiles, 128B-swizzled)
//   warp 1     tcgen05.mma issuer    (UMMA 128 x BN x 16, kind::f16, two accumulator stages of 256 TMEM columns)
//   warp 2     TMEM allocator
//   warps 4-11 epilogue              (two warps per TMEM lane quarter, each owning half of the tile's columns:
//                                     software-pipelined tcgen05.ld 32x32b.x16 -> bias / pad correction from shared
//                                     memory / ReLU|GELU / prefetched residuals / separator zeroing -> fp32 or bf16
//                                     rows in HBM), overlapped with the next tile's MMAs through the second TMEM stage
#include <cuda.h>
#include <cstdio>
#include <cstring>
#include "common.cuh"
#include "kernels.h"

namespace vrd {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;            // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int STAGES = 4;
constexpr int ACC_STAGE_COLS = 256;    // TMEM columns per accumulator stage
constexpr int TMEM_COLS = 512;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 8;
constexpr int MAX_N = 2048;             // bias / correction vectors are staged in shared memory
constexpr unsigned SPIN_LIMIT = 1u << 24;

char g_err[256] = {0};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    unsigned spins = 0;
    while (true) {
        uint32_t ok;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(addr), "r"(parity)
            : "memory");
        if (ok) return;
        if (++spins > SPIN_LIMIT) __trap();   // a lost arrival must fail loudly, never hang the GPU
    }
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// GELU(x) = x * Phi(x) with erf from Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7, far below the bf16 rounding of
// this kernel's GELU outputs); ~3x cheaper than erff in an epilogue that evaluates it 32k times per tile.
__device__ __forceinline__ float gelu_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    const float t = __frcp_rn(fmaf(0.3275911f, z, 1.0f));
    float p = fmaf(1.061405429f, t, -1.453152027f);
    p = fmaf(p, t, 1.421413741f);
    p = fmaf(p, t, -0.284496736f);
    p = fmaf(p, t, 0.254829592f);
    const float erfz = 1.0f - p * t * __expf(-z * z);           // erf(|x|/sqrt2)
    return 0.5f * x * (1.0f + copysignf(erfz, x));
}

// K-major, 128-byte-swizzled shared-memory matrix descriptor (rows of 128 B, 8-row groups 1024 B apart).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);          // start address       bits [0,14)
    d |= (uint64_t)1 << 16;                                // leading byte offset bits [16,30) (unused for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                      // stride byte offset  bits [32,46)
    d |= (uint64_t)1 << 46;                                // descriptor version  bits [46,48)
    d |= (uint64_t)2 << 61;                                // SWIZZLE_128B        bits [61,64)
    return d;
}
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = 128.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

struct EpiArgs {
    const float* bias; void* out; int out_dtype; long long ldo;
    int M, N, act;
    const float* res1; long long ldr1; const float* res2; long long ldr2;
    const float* corr; const int* row_seq; const int4* seqinfo; int R;
};

__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w, EpiArgs e, int K,
                    int taps, int block_n) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: stages of A, stages of W, barriers
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int w_stage_bytes = block_n * BLOCK_K * 2;
    uint8_t* smem_a = smem;
    uint8_t* smem_w = smem + STAGES * A_STAGE_BYTES;
    uint64_t* bars = (uint64_t*)(smem_w + STAGES * w_stage_bytes);
    uint64_t* full = bars;                 // [STAGES]
    uint64_t* empty = bars + STAGES;       // [STAGES]
    uint64_t* tfull = bars + 2 * STAGES;   // [2]
    uint64_t* tempty = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = (uint32_t*)(bars + 2 * STAGES + 4);
    float* s_bias = (float*)(bars + 2 * STAGES + 8);   // [MAX_N]
    float* s_corr = s_bias + MAX_N;                     // [MAX_N]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles_n = e.N / block_n;
    const int n_tiles = (e.M / BLOCK_M) * n_tiles_n;
    const int kb_per_tap = K / BLOCK_K;
    const int num_kb = taps * kb_per_tap;

    for (int i = threadIdx.x; i < e.N; i += NUM_THREADS) {
        s_bias[i] = (e.bias != nullptr) ? e.bias[i] : 0.f;
        s_corr[i] = (e.corr != nullptr) ? e.corr[i] : 0.f;
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
                const int m0 = (tile / n_tiles_n) * BLOCK_M, n0 = (tile % n_tiles_n) * block_n;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    mbar_expect_tx(&full[stage], A_STAGE_BYTES + w_stage_bytes);
                    const int tap = kb / kb_per_tap;
                    const int ka = (kb - tap * kb_per_tap) * BLOCK_K;
                    const int row = m0 + (taps == 3 ? tap - 1 : 0);
                    tma_load_2d(smem_a + stage * A_STAGE_BYTES, &map_a, &full[stage], ka, row);
                    tma_load_2d(smem_w + stage * w_stage_bytes, &map_w, &full[stage], kb * BLOCK_K, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            const uint32_t idesc = make_idesc(block_n);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
                const int as = it & 1;
                const uint32_t aphase = (it >> 1) & 1;
                mbar_wait(&tempty[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * ACC_STAGE_COLS;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full[stage], phase);
                    tc_fence_after();
                    const uint64_t adesc = make_smem_desc(smem_u32(smem_a + stage * A_STAGE_BYTES));
                    const uint64_t bdesc = make_smem_desc(smem_u32(smem_w + stage * w_stage_bytes));
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        // advance 32 bytes (16 bf16) along K inside the swizzle atom: +2 in 16-byte units
                        umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty[stage]);          // frees the smem stage once these MMAs have read it
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                umma_commit(&tfull[as]);                 // accumulator stage complete
            }
        }
    } else if (warp >= EPI_WARP0) {
        const int wq = warp & 3;                         // TMEM lane quarter this warp may access
        const int half = (warp - EPI_WARP0) >> 2;        // which half of the tile's 16-column chunks
        const int n_chunks = block_n / 16;
        const int c_begin = half == 0 ? 0 : (n_chunks + 1) / 2;
        const int c_end = half == 0 ? (n_chunks + 1) / 2 : n_chunks;
        const bool has_res = e.res1 != nullptr;
        int it = 0;
        for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m0 = (tile / n_tiles_n) * BLOCK_M, n0 = (tile % n_tiles_n) * block_n;
            const int row = m0 + wq * 32 + lane;
            bool valid = true, add_corr = false;
            if (e.row_seq != nullptr) {
                const int rl = row % e.R;
                const int seq = e.row_seq[rl];
                valid = seq >= 0;
                if (valid && e.corr != nullptr) {
                    const int4 si = e.seqinfo[seq];
                    add_corr = (si.z != 0) && (rl - si.x == si.y - 1);
                }
            }
            const float* r1p = has_res ? e.res1 + (long long)row * e.ldr1 + n0 : nullptr;
            const float* r2p = e.res2 != nullptr ? e.res2 + (long long)row * e.ldr2 + n0 : nullptr;
            auto load_res = [&](int c, float4 (&dst)[4]) {
                const float4* p = reinterpret_cast<const float4*>(r1p + c * 16);
#pragma unroll
                for (int i = 0; i < 4; ++i) dst[i] = p[i];
                if (r2p != nullptr) {
                    const float4* p2 = reinterpret_cast<const float4*>(r2p + c * 16);
#pragma unroll
                    for (int i = 0; i < 4; ++i) { const float4 t = p2[i]; dst[i].x += t.x; dst[i].y += t.y; dst[i].z += t.z; dst[i].w += t.w; }
                }
            };
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + as * ACC_STAGE_COLS;
            // one pipeline step: wait for chunk c (already in flight into `cur`), launch chunk c+1 into `nxt`, finish chunk c
            auto step = [&](int c, uint32_t (&cur)[16], float4 (&rcur)[4], uint32_t (&nxt)[16], float4 (&rnxt)[4]) {
                tmem_ld_wait();
                if (c + 1 < c_end) {
                    tmem_ld16(taddr + (c + 1) * 16, nxt);
                    if (has_res && valid) load_res(c + 1, rnxt);
                }
                const int n = n0 + c * 16;
                float v[16];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 bi = *reinterpret_cast<const float4*>(s_bias + n + 4 * i);
                    v[4 * i + 0] = __uint_as_float(cur[4 * i + 0]) + bi.x;
                    v[4 * i + 1] = __uint_as_float(cur[4 * i + 1]) + bi.y;
                    v[4 * i + 2] = __uint_as_float(cur[4 * i + 2]) + bi.z;
                    v[4 * i + 3] = __uint_as_float(cur[4 * i + 3]) + bi.w;
                }
                if (add_corr) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += s_corr[n + i];
                }
                if (e.act == 1) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = fmaxf(v[i], 0.f);
                } else if (e.act == 2) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = gelu_fast(v[i]);
                }
                if (has_res && valid) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) { v[4 * i] += rcur[i].x; v[4 * i + 1] += rcur[i].y; v[4 * i + 2] += rcur[i].z; v[4 * i + 3] += rcur[i].w; }
                }
                if (!valid) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = 0.f;
                }
                if (e.out_dtype == VRD_BF16) {
                    __nv_bfloat16* op = (__nv_bfloat16*)e.out + (long long)row * e.ldo + n;
                    uint4 pk0, pk1;
                    pk0.x = pack2(v[0], v[1]); pk0.y = pack2(v[2], v[3]); pk0.z = pack2(v[4], v[5]); pk0.w = pack2(v[6], v[7]);
                    pk1.x = pack2(v[8], v[9]); pk1.y = pack2(v[10], v[11]); pk1.z = pack2(v[12], v[13]); pk1.w = pack2(v[14], v[15]);
                    reinterpret_cast<uint4*>(op)[0] = pk0;
                    reinterpret_cast<uint4*>(op)[1] = pk1;
                } else {
                    float4* op = reinterpret_cast<float4*>((float*)e.out + (long long)row * e.ldo + n);
#pragma unroll
                    for (int i = 0; i < 4; ++i) op[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                }
            };
            uint32_t acc0[16], acc1[16];
            float4 rs0[4], rs1[4];
            if (has_res && valid) load_res(c_begin, rs0);         // residual rows do not depend on the accumulator
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            tmem_ld16(taddr + c_begin * 16, acc0);
            for (int c = c_begin; c < c_end; c += 2) {
                step(c, acc0, rs0, acc1, rs1);
                if (c + 1 < c_end) step(c + 1, acc1, rs1, acc0, rs0);
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[as]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

// 2-D bf16 row-major matrix [rows, cols] with row pitch ld (elements); box = [box_rows, 64 cols], 128B swizzle.
bool make_map(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_rows) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) { snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled entry point not found"); return false; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled failed: %d", (int)r); return false; }
    return true;
}

}  // namespace

const char* gemm_tcgen05_error() { return g_err; }

int gemm_tcgen05_bf16(const GemmArgs& g, cudaStream_t st) {
    static int num_sms = 0;
    static bool attr_set = false;
    if (g.M % BLOCK_M != 0 || g.K % BLOCK_K != 0 || g.N % 16 != 0 || g.lda % 8 != 0 || ((uintptr_t)g.A & 15) != 0) {
        snprintf(g_err, sizeof g_err, "gemm_tcgen05: unsupported shape M=%d N=%d K=%d lda=%lld", g.M, g.N, g.K, g.lda);
        return 1;
    }
    int block_n;
    if (g.N % 256 == 0) block_n = 256;
    else if (g.N <= 256) block_n = g.N;
    else if (g.N % 128 == 0) block_n = 128;
    else { snprintf(g_err, sizeof g_err, "gemm_tcgen05: unsupported N=%d", g.N); return 1; }
    if (g.N > MAX_N) { snprintf(g_err, sizeof g_err, "gemm_tcgen05: N=%d exceeds %d", g.N, MAX_N); return 1; }
    if ((g.ldo % (g.out_dtype == VRD_BF16 ? 8 : 4)) != 0 || (g.res1 && g.ldr1 % 4) || (g.res2 && g.ldr2 % 4)) {
        snprintf(g_err, sizeof g_err, "gemm_tcgen05: output/residual pitch must be a multiple of 4");
        return 1;
    }
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
    }
    CUtensorMap map_a, map_w;
    if (!make_map(&map_a, g.A, g.M, g.K, g.lda, BLOCK_M)) return 1;
    if (!make_map(&map_w, g.W, g.N, (long long)g.taps * g.K, (long long)g.taps * g.K, block_n)) return 1;
    const int smem = 1024 + STAGES * (A_STAGE_BYTES + block_n * BLOCK_K * 2) + 256 + 2 * MAX_N * 4;
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024) != cudaSuccess) {
            snprintf(g_err, sizeof g_err, "cudaFuncSetAttribute(max dynamic smem) failed");
            return 1;
        }
        attr_set = true;
    }
    EpiArgs e{g.bias, g.out, g.out_dtype, g.ldo, g.M, g.N, g.act, g.res1, g.ldr1, g.res2, g.ldr2, g.corr, g.row_seq, g.seqinfo, g.R};
    const int n_tiles = (g.M / BLOCK_M) * (g.N / block_n);
    const int grid = n_tiles < num_sms ? n_tiles : num_sms;
    gemm_tcgen05_kernel<<<grid, NUM_THREADS, smem, st>>>(map_a, map_w, e, g.K, g.taps, block_n);
    return 0;
}

}  // namespace vrd
