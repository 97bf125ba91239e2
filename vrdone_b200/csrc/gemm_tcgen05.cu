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
//   warp 0     TMA producer          (3-4 stage ring of {A 128x64, W BNx64} bf16 tiles, 128B-swizzled)
//   warp 1     tcgen05.mma issuer    (UMMA 128 x BN x 16, kind::f16, two accumulator stages of 256 TMEM columns)
//   warp 2     TMEM allocator
//   warps 4-11 epilogue              (two warps per TMEM lane quarter, each owning half of the tile's columns).  Per chunk
//                                     of 32 columns: tcgen05.ld 32x32b.x32 -> bias / pad correction / ReLU|GELU / residual /
//                                     separator zeroing in registers (thread = row) -> swizzled shared-memory staging ->
//                                     one TMA store of the [32 rows x 32 cols] box ([32 x 64] for bf16 outputs without a
//                                     residual: two chunks per store).  The fp32 residual tile arrives the same
//                                     way (TMA load into the staging buffer, issued before the accumulator is waited for),
//                                     so the epilogue warps never issue row-strided global accesses (those cost one L1
//                                     wavefront per row and bounded the old epilogue).  Overlapped with the next tile's
//                                     MMAs through the second TMEM stage.
// Template parameters: CG (1 | 2 CTAs per tile), MODE (the epilogue's feature flags at run time -- EPI_GENERIC -- or fixed at
// compile time: EPI_BF16, EPI_BF16_GELU, EPI_F32_RES), BN (tile width when fixed).  EPI_LN is the odd one: a 512-column tile whose
// accumulator fills the tensor memory (one stage: main loop and epilogue alternate) so that the epilogue owns whole rows and applies
// the channel LayerNorm (+ ReLU) of the embedding convs; see the enum below.  Every kernel starts with griddepcontrol.wait after its
// prologue (programmatic dependent launch, common.cuh).
#include <cuda.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace vrd {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;            // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int MAX_STAGES = 4;
constexpr int ACC_STAGE_COLS = 256;    // TMEM columns per accumulator stage
constexpr int TMEM_COLS = 512;
constexpr int A_STAGE_BYTES = BLOCK_M * BLOCK_K * 2;
constexpr int NUM_THREADS = 384;
constexpr int EPI_WARP0 = 4;
constexpr int EPI_WARPS = 8;
constexpr int MAX_N = 2048;             // bias / correction vectors are staged in shared memory
constexpr int CHUNK = 32;               // accumulator columns per epilogue step = columns of one TMA store box
constexpr int STAGING_BYTES = 32 * CHUNK * 4;                  // one [32 rows x 32 cols] fp32 box
constexpr int STAGING_TOTAL = EPI_WARPS * 2 * STAGING_BYTES;   // two boxes per epilogue warp
constexpr int SMEM_LIMIT = 227 * 1024;
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
__device__ __forceinline__ void tma_load_2d_cg2(void* smem_dst, const CUtensorMap* map, uint32_t bar_cluster_addr, int c0, int c1) {
    // issued by both CTAs of a pair; the transaction bytes are credited to the barrier at `bar_cluster_addr` (the leader's)
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(smem_dst)), "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default semantics (release at CTA scope): the TMEM reads it orders were completed by tcgen05.wait::ld; a cluster-scope
    // release would also wait for this warp's outstanding shared-memory / TMA-store traffic (35 % of epilogue stall samples)
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void umma_commit_cg2(uint64_t* bar) {   // arrives on the barrier at this offset in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void umma_bf16_cg2(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
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
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]),
          "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]),
          "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
        "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
          "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
          "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
        : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t smem_src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(smem_src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// GELU(x) = x * Phi(x) with erf from Abramowitz-Stegun 7.1.26 (|erf error| <= 1.5e-7, far below the bf16 rounding of
// this kernel's GELU outputs).  Phi(x) = 1 - h for x >= 0 and h for x < 0 with h = 0.5 * P(t) * exp(-x^2 / 2),
// t = 1 / (1 + p |x| / sqrt2): two MUFU ops (rcp.approx, ex2.approx) and 13 FP32 ops per element, no slow paths.
__device__ __forceinline__ float gelu_fast(float x) {
    const float z = fabsf(x) * 0.70710678118654752440f;
    float t;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.0f)));
    float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
    p = fmaf(p, t, 0.5f * 1.421413741f);
    p = fmaf(p, t, 0.5f * -0.284496736f);
    p = fmaf(p, t, 0.5f * 0.254829592f);
    float e;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * (z * -1.4426950408889634f)));
    const float h = p * t * e;                                   // 0.5 * erfc(|x| / sqrt2)
    return x * (x >= 0.f ? 1.0f - h : h);
}

// GELU of two values on the packed-half pipe, returned as bf16x2 (lo = a, hi = b).  Used when the GEMM output is bf16 anyway:
// x * 0.5 * (1 + tanh(x * (c0 + c1 x^2))) with (c0, c1) fitted to the exact erf form (max |deviation| 2.7e-4), evaluated in
// f16x2 (cvt / 4 HFMA2-class ops / one MUFU.TANH per PAIR).  Its error (mean 2.8e-4 absolute on N(0, 1.5) inputs) is a third of
// the rounding error of the bf16 result it feeds (mean 8.4e-4), and it is 3x cheaper than the fp32 erf form, which made the
// 512 -> 2048 MLP GEMM epilogue-bound (0.80 PFLOP/s).
__device__ __forceinline__ uint32_t gelu_pair_bf16(float a, float b) {
    uint32_t xh, x2, p, u, t, hx, r;
    // upper half = b, lower half = a; saturating: |x| > 65504 becomes +-65504, not inf, so the last step never sees inf - inf
    // (x^2 may still overflow to inf: then tanh(+-inf) = +-1 and the result is x or 0, as it should be)
    asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(xh) : "f"(b), "f"(a));
    asm("mul.f16x2 %0, %1, %1;" : "=r"(x2) : "r"(xh));
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(x2), "r"(0x28712871u), "r"(0x3a673a67u));   // c1 = 0.034701, c0 = 0.800157
    asm("mul.f16x2 %0, %1, %2;" : "=r"(u) : "r"(xh), "r"(p));
    asm("tanh.approx.f16x2 %0, %1;" : "=r"(t) : "r"(u));
    asm("mul.f16x2 %0, %1, %2;" : "=r"(hx) : "r"(xh), "r"(0x38003800u));       // 0.5 x
    asm("fma.rn.f16x2 %0, %1, %2, %1;" : "=r"(r) : "r"(hx), "r"(t));
    float lo, hi;
    asm("{\n\t.reg .f16 l, h;\n\tmov.b32 {l, h}, %2;\n\tcvt.f32.f16 %0, l;\n\tcvt.f32.f16 %1, h;\n\t}" : "=f"(lo), "=f"(hi) : "r"(r));
    return pack2(lo, hi);
}

__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// read-only shared data (bias vector, written once before the kernel's first barrier): schedulable, no memory clobber
__device__ __forceinline__ float4 lds128_ro(uint32_t addr) {
    float4 r;
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr));
    return r;
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 r;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr) : "memory");
    return r;
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
// kind::f16 instruction descriptor: D fp32, A/B bf16, both K-major, M = m (128, or 256 for a CTA pair).
__host__ __device__ constexpr uint32_t make_idesc(int n, int m) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct EpiArgs {
    const float* bias; int out_dtype;
    int M, N, act, has_res;
    const float* res2; long long ldr2;      // second residual (only the SOS output projection uses it): boxes through map_res2, like res1
    const float* corr; const int* row_seq; const int4* seqinfo; int R;
    int wide;                               // bf16 output without residual: [32 x 64] store boxes (map in the map_res slot)
    int dbg;                                // timing experiments only (wrong results): 1 = no TMA stores, 2 = nothing after the TMEM load
    const float* ln_g; const float* ln_b; int ln_relu;   // EPI_LN: channel LayerNorm (+ ReLU) over the N = 512 outputs of a row
};

// CG = 1: one CTA per 128 x BN tile.  CG = 2: a CTA pair (cluster of two SMs of one TPC) shares a 256 x BN tile through
// tcgen05.mma.cta_group::2 -- each CTA stages its own 128 rows of A but only HALF of the W tile, and keeps its 128 accumulator
// rows in its own TMEM, so the shared-memory fill per FLOP (the L2 -> SM traffic that bounds the K = 512 GEMMs) drops by a third.
// Only the leader CTA (rank 0) issues MMAs; its commits arrive on the barriers of both CTAs.
// MODE: the epilogue's feature flags are run-time values in EPI_GENERIC and compile-time constants in the specialised modes (the
// same arithmetic on the same registers: results are bit-identical).  ncu on the generic epilogue of a K = 512 GEMM: 208 executed
// instructions per 32-column chunk, of which 32 FADD + 16 F2FP + 8 LDS + 4 STS + 1 LDTM do the work and ~120 are LDC / ISETP / BRA /
// BSSY chains re-deciding the flags (warp cycles per issued instruction 10-14: two epilogue warps per scheduler hide nothing), so
// the epilogue, not the tensor pipe, paced these GEMMs.  BN != 0 fixes the tile width (the chunk loop unrolls).
// EPI_LN (BN = 512): the tile spans the whole 512-channel row, so the epilogue owns complete rows and applies the channel LayerNorm
// (+ ReLU) of the embedding convs itself (reference blocks.py:143-158 after the k = 3 convs of backbones.py:184-197): the fp32
// conv output never reaches HBM and the standalone layernorm launch disappears.  The accumulator takes all 512 TMEM columns (two
// N = 256 instructions per K step), so main loop and epilogue of a CTA alternate instead of overlapping -- these GEMMs have K = 1536
// or 3072, the epilogue is ~5-10 % of a tile.
// EPI_LN_RES: the attention output projection of an encoder block (blocks.py:1070-1076): y = acc + bias + residual is the new
// residual stream (fp32, written through TMA boxes like the residual arrives) AND h = LayerNorm(y) (bf16) is the MLP's input --
// both from one pass over the accumulator row; y is parked in tensor memory between the statistics pass and the normalisation.
enum { EPI_GENERIC = 0, EPI_BF16 = 1, EPI_BF16_GELU = 2, EPI_F32_RES = 3, EPI_LN = 4, EPI_LN_RES = 5 };

template <int CG, int MODE, int BN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_w,
                    const __grid_constant__ CUtensorMap map_out, const __grid_constant__ CUtensorMap map_res,
                    const __grid_constant__ CUtensorMap map_res2, EpiArgs e, int K,
                    int taps, int block_n_rt, int stages, int split) {
    const int block_n = BN != 0 ? BN : block_n_rt;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    // carve: stages of A, stages of W, epilogue staging boxes, barriers, bias / correction vectors
    uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
    const int w_stage_bytes = (block_n / CG) * BLOCK_K * 2;       // each CTA of a pair holds block_n / CG rows of the W tile
    const uint32_t cta_rank = CG == 2 ? cluster_ctarank() : 0;
    const int tile0 = blockIdx.x / CG, tile_step = gridDim.x / CG;   // tiles are walked by clusters
    uint8_t* smem_a = smem;
    uint8_t* smem_w = smem + stages * A_STAGE_BYTES;
    uint8_t* staging = smem_w + stages * w_stage_bytes;          // 1024-aligned: every stage size is a multiple of 1024
    uint8_t* staging2 = staging + STAGING_TOTAL;                 // second-residual boxes (only carved when there is one)
    constexpr bool LNM = MODE == EPI_LN || MODE == EPI_LN_RES;   // the 512-column accumulator, one stage
    constexpr bool LNR = MODE == EPI_LN_RES;                     // ... with an fp32 residual in and the fp32 sum out
    static_assert(!LNM || BN == 512, "EPI_LN needs the full row in one tile");
    // EPI_LN: one staging box per epilogue warp; EPI_LN_RES: two (residual in / fp32 out, then the bf16 boxes)
    uint64_t* bars = (uint64_t*)(staging + (MODE == EPI_LN ? STAGING_TOTAL / 2 : (e.res2 != nullptr ? 2 : 1) * STAGING_TOTAL));
    uint64_t* full = bars;                           // [MAX_STAGES]
    uint64_t* empty = bars + MAX_STAGES;             // [MAX_STAGES]
    uint64_t* tfull = bars + 2 * MAX_STAGES;         // [2]
    uint64_t* tempty = bars + 2 * MAX_STAGES + 2;    // [2]
    uint64_t* resbar = bars + 2 * MAX_STAGES + 4;    // [EPI_WARPS][2]
    uint32_t* tmem_slot = (uint32_t*)(resbar + 2 * EPI_WARPS);
    float* s_bias = (float*)(tmem_slot + 4);         // [MAX_N]
    float* s_corr = s_bias + MAX_N;                  // [MAX_N]
    float* s_xch = s_corr + MAX_N / 2;               // EPI_LN (N = 512 <= MAX_N / 2): [2 tile parities][2 column halves][128 rows][sum, sum of squares]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n_tiles_n = e.N / block_n;
    const int n_tiles = ((e.M + BLOCK_M * CG - 1) / (BLOCK_M * CG)) * n_tiles_n;
    const int kb_per_tap = K / BLOCK_K;
    // split != 0: fp32 operands as bf16 pairs (hi, lo), A = [hi | lo] (2K columns), W = per tap [hi | lo]: three K-slabs per tap,
    // A_hi W_hi + A_lo W_hi + A_hi W_lo (the 3 x bf16 product with fp32 accumulation; the lo x lo term, 2^-16 relative, is dropped)
    const int num_kb = taps * kb_per_tap * (split ? 3 : 1);

    // weights only (bias, pad-column correction: constant during a forward), so this may run before pdl_wait()
    for (int i = threadIdx.x; i < e.N; i += NUM_THREADS) {
        s_bias[i] = (e.bias != nullptr) ? e.bias[i] : 0.f;
        s_corr[i] = (e.corr != nullptr) ? e.corr[i] : 0.f;
        if constexpr (LNM) {                                                                // N = 512: gamma, beta behind the bias
            s_bias[512 + i] = e.ln_g != nullptr ? e.ln_g[i] : 1.f;
            s_bias[1024 + i] = e.ln_g != nullptr ? e.ln_b[i] : 0.f;
        }
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < stages; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&tfull[i], 1); mbar_init(&tempty[i], EPI_WARPS * CG); }
        for (int i = 0; i < 2 * EPI_WARPS; ++i) mbar_init(&resbar[i], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    } else if (warp == 2) {
        if constexpr (CG == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // barriers of BOTH CTAs initialised before any remote arrive
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barrier init, TMEM allocation, bias staging) overlapped the previous kernel's tail; activations, residuals
    // and layout arrays are only touched from here on
    pdl_wait();

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0; uint32_t phase = 0;
            for (int tile = tile0; tile < n_tiles; tile += tile_step) {
                const int m0 = (tile / n_tiles_n) * (BLOCK_M * CG) + cta_rank * BLOCK_M;
                const int n0 = (tile % n_tiles_n) * block_n + cta_rank * (block_n / CG);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty[stage], phase ^ 1);
                    // the leader's barrier collects the bytes of both CTAs of a pair
                    if (cta_rank == 0) mbar_expect_tx(&full[stage], CG * (A_STAGE_BYTES + w_stage_bytes));
                    int tap, ka, wcol;
                    if (split) {
                        const int seg_all = kb / kb_per_tap;
                        tap = seg_all / 3;
                        const int seg = seg_all - 3 * tap;
                        const int kk = (kb - seg_all * kb_per_tap) * BLOCK_K;
                        ka = (seg == 1 ? K : 0) + kk;
                        wcol = tap * 2 * K + (seg == 2 ? K : 0) + kk;
                    } else {
                        tap = kb / kb_per_tap;
                        ka = (kb - tap * kb_per_tap) * BLOCK_K;
                        wcol = kb * BLOCK_K;
                    }
                    const int row = m0 + (taps == 3 ? tap - 1 : 0);
                    if constexpr (CG == 2) {
                        const uint32_t bar = mapa_shared(smem_u32(&full[stage]), 0);
                        tma_load_2d_cg2(smem_a + stage * A_STAGE_BYTES, &map_a, bar, ka, row);
                        if constexpr (LNM) {   // two boxes of 128 W rows: this CTA's share of output channels [0, 256) and of [256, 512)
                            tma_load_2d_cg2(smem_w + stage * w_stage_bytes, &map_w, bar, wcol, cta_rank * 128);
                            tma_load_2d_cg2(smem_w + stage * w_stage_bytes + w_stage_bytes / 2, &map_w, bar, wcol, 256 + cta_rank * 128);
                        } else {
                            tma_load_2d_cg2(smem_w + stage * w_stage_bytes, &map_w, bar, wcol, n0);
                        }
                    } else {
                        tma_load_2d(smem_a + stage * A_STAGE_BYTES, &map_a, &full[stage], ka, row);
                        if constexpr (LNM) {
                            tma_load_2d(smem_w + stage * w_stage_bytes, &map_w, &full[stage], wcol, 0);
                            tma_load_2d(smem_w + stage * w_stage_bytes + w_stage_bytes / 2, &map_w, &full[stage], wcol, 256);
                        } else {
                            tma_load_2d(smem_w + stage * w_stage_bytes, &map_w, &full[stage], wcol, n0);
                        }
                    }
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0 && cta_rank == 0) {
            const uint32_t idesc = make_idesc(LNM ? 256 : block_n, BLOCK_M * CG);
            int stage = 0; uint32_t phase = 0;
            int it = 0;
            for (int tile = tile0; tile < n_tiles; tile += tile_step, ++it) {
                const int as = LNM ? 0 : (it & 1);                       // EPI_LN: one accumulator stage of 512 columns
                const uint32_t aphase = LNM ? (it & 1) : ((it >> 1) & 1);
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
                        if constexpr (CG == 2) umma_bf16_cg2(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        else umma_bf16(tmem_d, adesc + 2 * k, bdesc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        if constexpr (LNM) {   // output channels [256, 512): the second half of the W stage, accumulator columns 256 ..
                            const uint64_t bdesc2 = make_smem_desc(smem_u32(smem_w + stage * w_stage_bytes + w_stage_bytes / 2));
                            if constexpr (CG == 2) umma_bf16_cg2(tmem_d + 256, adesc + 2 * k, bdesc2 + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                            else umma_bf16(tmem_d + 256, adesc + 2 * k, bdesc2 + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
                        }
                    }
                    // frees the smem stage (in both CTAs of a pair) once these MMAs have read it
                    if constexpr (CG == 2) umma_commit_cg2(&empty[stage]); else umma_commit(&empty[stage]);
                    if (++stage == stages) { stage = 0; phase ^= 1; }
                }
                if constexpr (CG == 2) umma_commit_cg2(&tfull[as]); else umma_commit(&tfull[as]);   // accumulator stage complete
            }
        }
    } else if (warp >= EPI_WARP0 && LNR) {
        // ---- EPI_LN_RES.  Pass 1 per 32-column chunk: residual box (TMA, two boxes per warp in flight) + accumulator + bias = y;
        // statistics; y goes back into tensor memory (tcgen05.st) and, through the same box, to HBM (fp32).  Pass 2: y from tensor
        // memory -> LayerNorm -> bf16 -> [32 x 64] boxes (the two boxes alternate).
        const int ew = warp - EPI_WARP0;
        const int wq = warp & 3;
        const int half = ew >> 2;
        const uint32_t s_bias_u = smem_u32(s_bias);
        uint8_t* my_stage = staging + ew * 2 * STAGING_BYTES;
        uint64_t* my_resbar = resbar + 2 * ew;
        uint32_t res_cnt[2] = {0u, 0u};                      // fills of each box so far: the barrier phases
        int it = 0;
        for (int tile = tile0; tile < n_tiles; tile += tile_step, ++it) {
            const int m0 = tile * (BLOCK_M * CG) + cta_rank * BLOCK_M;
            const int row0 = m0 + wq * 32, row = row0 + lane;
            const int col0 = half * 256;
            bool valid = row < e.M;
            if (e.row_seq != nullptr && valid) valid = e.row_seq[row % e.R] >= 0;
            if (lane == 0) {                                 // the residual boxes of chunks 0 and 1 travel while the MMAs of this tile run
                bulk_wait_read<0>();                         // the previous tile's bf16 stores have read both boxes
                for (int b = 0; b < 2; ++b) {
                    mbar_expect_tx(&my_resbar[b], STAGING_BYTES);
                    tma_load_2d(my_stage + b * STAGING_BYTES, &map_res, &my_resbar[b], col0 + b * CHUNK, row0);
                }
            }
            __syncwarp();
            mbar_wait(&tfull[0], it & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + half * 256;
            uint32_t acc[CHUNK];
            float s1 = 0.f, s2 = 0.f;
            tmem_ld32(taddr, acc);
#pragma unroll 2
            for (int c = 0; c < 8; ++c) {
                const int b = c & 1;
                const uint32_t box_u = smem_u32(my_stage + b * STAGING_BYTES);
                mbar_wait(&my_resbar[b], res_cnt[b] & 1);
                ++res_cnt[b];
                tmem_ld_wait();
                const int n = col0 + c * CHUNK;
                float v[CHUNK];
#pragma unroll
                for (int i = 0; i < CHUNK / 4; ++i) {
                    const float4 bi = lds128_ro(s_bias_u + (uint32_t)(n + 4 * i) * 4);
                    const float4 r = lds128(box_u + lane * 128 + ((i ^ (lane & 7)) << 4));
                    v[4 * i + 0] = (__uint_as_float(acc[4 * i + 0]) + bi.x) + r.x;
                    v[4 * i + 1] = (__uint_as_float(acc[4 * i + 1]) + bi.y) + r.y;
                    v[4 * i + 2] = (__uint_as_float(acc[4 * i + 2]) + bi.z) + r.z;
                    v[4 * i + 3] = (__uint_as_float(acc[4 * i + 3]) + bi.w) + r.w;
                }
                if (c + 1 < 8) tmem_ld32(taddr + (c + 1) * CHUNK, acc);
                if (!valid) {
#pragma unroll
                    for (int i = 0; i < CHUNK; ++i) v[i] = 0.f;
                }
                float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
                for (int i = 0; i < CHUNK; i += 2) {
                    a0 += v[i]; a1 += v[i + 1];
                    q0 = fmaf(v[i], v[i], q0); q1 = fmaf(v[i + 1], v[i + 1], q1);
                }
                s1 += a0 + a1;
                s2 += q0 + q1;
                {   // y back into the accumulator columns it came from (read again in pass 2)
                    uint32_t vb[CHUNK];
#pragma unroll
                    for (int i = 0; i < CHUNK; ++i) vb[i] = __float_as_uint(v[i]);
                    tmem_st32(taddr + c * CHUNK, vb);
                }
#pragma unroll
                for (int j = 0; j < CHUNK / 4; ++j)
                    sts128(box_u + lane * 128 + ((j ^ (lane & 7)) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                           __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
                fence_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&map_out, box_u, n, row0);
                    bulk_commit();
                    if (c + 2 < 8) {                         // refill this box with the residual of chunk c + 2
                        bulk_wait_read<0>();
                        mbar_expect_tx(&my_resbar[b], STAGING_BYTES);
                        tma_load_2d(my_stage + b * STAGING_BYTES, &map_res, &my_resbar[b], col0 + (c + 2) * CHUNK, row0);
                    }
                }
            }
            tmem_st_wait();
            float* xs = s_xch + (it & 1) * 512;
            *reinterpret_cast<float2*>(xs + (half * 128 + wq * 32 + lane) * 2) = make_float2(s1, s2);
            asm volatile("bar.sync %0, 64;" ::"r"(1 + wq) : "memory");
            const float2 xa = *reinterpret_cast<const float2*>(xs + (wq * 32 + lane) * 2);
            const float2 xb = *reinterpret_cast<const float2*>(xs + (128 + wq * 32 + lane) * 2);
            const float mean = (xa.x + xb.x) * (1.0f / 512.f);
            const float var = fmaxf((xa.y + xb.y) * (1.0f / 512.f) - mean * mean, 0.f);
            const float rstd = rsqrtf(var + VRD_EPS);
            const float nmr = -mean * rstd;
            tmem_ld32(taddr, acc);
#pragma unroll 2
            for (int c = 0; c < 8; ++c) {
                const uint32_t box_u = smem_u32(my_stage + ((c >> 1) & 1) * STAGING_BYTES);
                if ((c & 1) == 0) {
                    if (lane == 0) { if (c < 4) bulk_wait_read<0>(); else bulk_wait_read<1>(); }   // pass-1 stores / the store two boxes ago
                    __syncwarp();
                }
                tmem_ld_wait();
                const int n = col0 + c * CHUNK;
                float v[CHUNK];
#pragma unroll
                for (int i = 0; i < CHUNK; ++i) v[i] = __uint_as_float(acc[i]);
                if (c + 1 < 8) {
                    tmem_ld32(taddr + (c + 1) * CHUNK, acc);
                } else {
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[0]), 0));
                        else mbar_arrive(&tempty[0]);
                    }
                }
#pragma unroll
                for (int i = 0; i < CHUNK / 4; ++i) {
                    const float4 gi = lds128_ro(s_bias_u + (uint32_t)(512 + n + 4 * i) * 4);
                    const float4 be = lds128_ro(s_bias_u + (uint32_t)(1024 + n + 4 * i) * 4);
                    v[4 * i + 0] = fmaf(fmaf(v[4 * i + 0], rstd, nmr), gi.x, be.x);
                    v[4 * i + 1] = fmaf(fmaf(v[4 * i + 1], rstd, nmr), gi.y, be.y);
                    v[4 * i + 2] = fmaf(fmaf(v[4 * i + 2], rstd, nmr), gi.z, be.z);
                    v[4 * i + 3] = fmaf(fmaf(v[4 * i + 3], rstd, nmr), gi.w, be.w);
                }
                if (!valid) {
#pragma unroll
                    for (int i = 0; i < CHUNK; ++i) v[i] = 0.f;
                }
#pragma unroll
                for (int j = 0; j < CHUNK / 8; ++j)
                    sts128(box_u + lane * 128 + (((j + 4 * (c & 1)) ^ (lane & 7)) << 4), pack2(v[8 * j], v[8 * j + 1]),
                           pack2(v[8 * j + 2], v[8 * j + 3]), pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
                if ((c & 1) == 1) {
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&map_res2, box_u, n - CHUNK, row0);   // map_res2 carries the [32 x 64] bf16 box map of the LayerNorm output
                        bulk_commit();
                    }
                }
            }
        }
        if (lane == 0) bulk_wait_all();
    } else if (warp >= EPI_WARP0 && LNM) {
        // ---- EPI_LN: thread = row; the two warps of a TMEM lane quarter own 256 columns each.  Pass 1: sum and sum of squares of
        // v = acc + bias (+ pad correction) over the warp's columns, exchanged with the partner warp through shared memory behind a
        // 64-thread named barrier.  Pass 2 (the accumulator is read again, nothing is kept in registers): normalise, gamma / beta,
        // ReLU, bf16, [32 x 64] staging box, TMA store.
        const int ew = warp - EPI_WARP0;
        const int wq = warp & 3;
        const int half = ew >> 2;
        const uint32_t s_bias_u = smem_u32(s_bias);
        uint8_t* box = staging + ew * STAGING_BYTES;
        const uint32_t box_u = smem_u32(box);
        int it = 0;
        for (int tile = tile0; tile < n_tiles; tile += tile_step, ++it) {
            const int m0 = tile * (BLOCK_M * CG) + cta_rank * BLOCK_M;
            const int row0 = m0 + wq * 32, row = row0 + lane;
            bool valid = row < e.M, add_corr = false;
            if (e.row_seq != nullptr && valid) {
                const int rl = row % e.R;
                const int seq = e.row_seq[rl];
                valid = seq >= 0;
                if (valid && e.corr != nullptr) {
                    const int4 si = e.seqinfo[seq];
                    add_corr = (si.z != 0) && (rl - si.x == si.y - 1);
                }
            }
            mbar_wait(&tfull[0], it & 1);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + half * 256;
            uint32_t acc[CHUNK];
            float s1 = 0.f, s2 = 0.f;
            const bool do_ln = e.ln_g != nullptr;    // false: experiment, a plain bf16 GEMM on 512-column tiles (A is read once)
            tmem_ld32(taddr, acc);
#pragma unroll 2
            for (int c = 0; c < (do_ln ? 8 : 0); ++c) {
                tmem_ld_wait();
                const int n = half * 256 + c * CHUNK;
                float v[CHUNK];
#pragma unroll
                for (int i = 0; i < CHUNK / 4; ++i) {
                    const float4 bi = lds128_ro(s_bias_u + (uint32_t)(n + 4 * i) * 4);
                    v[4 * i + 0] = __uint_as_float(acc[4 * i + 0]) + bi.x;
                    v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + bi.y;
                    v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + bi.z;
                    v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + bi.w;
                }
                tmem_ld32(taddr + ((c + 1) & 7) * CHUNK, acc);      // the next chunk; after the last one: chunk 0 of pass 2
                if (add_corr) {
#pragma unroll
                    for (int i = 0; i < CHUNK; ++i) v[i] += s_corr[n + i];
                }
                float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll
                for (int i = 0; i < CHUNK; i += 2) {
                    a0 += v[i]; a1 += v[i + 1];
                    q0 = fmaf(v[i], v[i], q0); q1 = fmaf(v[i + 1], v[i + 1], q1);
                }
                s1 += a0 + a1;
                s2 += q0 + q1;
            }
            float rstd = 1.f, nmr = 0.f;
            if (do_ln) {
                float* xs = s_xch + (it & 1) * 512;
                *reinterpret_cast<float2*>(xs + (half * 128 + wq * 32 + lane) * 2) = make_float2(s1, s2);
                asm volatile("bar.sync %0, 64;" ::"r"(1 + wq) : "memory");
                const float2 xa = *reinterpret_cast<const float2*>(xs + (wq * 32 + lane) * 2);
                const float2 xb = *reinterpret_cast<const float2*>(xs + (128 + wq * 32 + lane) * 2);
                const float mean = (xa.x + xb.x) * (1.0f / 512.f);
                const float var = fmaxf((xa.y + xb.y) * (1.0f / 512.f) - mean * mean, 0.f);
                rstd = rsqrtf(var + VRD_EPS);
                nmr = -mean * rstd;
            }
#pragma unroll 2
            for (int c = 0; c < 8; ++c) {
                if ((c & 1) == 0) {
                    if (lane == 0) bulk_wait_read<0>();      // the previous store has read this warp's box
                    __syncwarp();
                }
                tmem_ld_wait();
                const int n = half * 256 + c * CHUNK;
                float v[CHUNK];
#pragma unroll
                for (int i = 0; i < CHUNK / 4; ++i) {
                    const float4 bi = lds128_ro(s_bias_u + (uint32_t)(n + 4 * i) * 4);
                    v[4 * i + 0] = __uint_as_float(acc[4 * i + 0]) + bi.x;
                    v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + bi.y;
                    v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + bi.z;
                    v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + bi.w;
                }
                if (c + 1 < 8) {
                    tmem_ld32(taddr + (c + 1) * CHUNK, acc);
                } else {                                     // accumulator read twice: hand the TMEM back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[0]), 0));
                        else mbar_arrive(&tempty[0]);
                    }
                }
                if (add_corr) {
#pragma unroll
                    for (int i = 0; i < CHUNK; ++i) v[i] += s_corr[n + i];
                }
                if (do_ln) {
#pragma unroll
                for (int i = 0; i < CHUNK / 4; ++i) {
                    const float4 gi = lds128_ro(s_bias_u + (uint32_t)(512 + n + 4 * i) * 4);
                    const float4 be = lds128_ro(s_bias_u + (uint32_t)(1024 + n + 4 * i) * 4);
                    v[4 * i + 0] = fmaf(fmaf(v[4 * i + 0], rstd, nmr), gi.x, be.x);
                    v[4 * i + 1] = fmaf(fmaf(v[4 * i + 1], rstd, nmr), gi.y, be.y);
                    v[4 * i + 2] = fmaf(fmaf(v[4 * i + 2], rstd, nmr), gi.z, be.z);
                    v[4 * i + 3] = fmaf(fmaf(v[4 * i + 3], rstd, nmr), gi.w, be.w);
                }
                }
                if (e.ln_relu) {
#pragma unroll
                    for (int i = 0; i < CHUNK; ++i) v[i] = fmaxf(v[i], 0.f);
                }
                if (!valid) {
#pragma unroll
                    for (int i = 0; i < CHUNK; ++i) v[i] = 0.f;
                }
                if (e.act == 2) {                            // GELU (plain wide-tile GEMMs only; gelu(0) = 0 keeps invalid rows zero)
#pragma unroll
                    for (int j = 0; j < CHUNK / 8; ++j)
                        sts128(box_u + lane * 128 + (((j + 4 * (c & 1)) ^ (lane & 7)) << 4), gelu_pair_bf16(v[8 * j], v[8 * j + 1]),
                               gelu_pair_bf16(v[8 * j + 2], v[8 * j + 3]), gelu_pair_bf16(v[8 * j + 4], v[8 * j + 5]),
                               gelu_pair_bf16(v[8 * j + 6], v[8 * j + 7]));
                } else {
#pragma unroll
                for (int j = 0; j < CHUNK / 8; ++j)          // 128-byte rows: chunk jj of row r stored at chunk jj ^ (r & 7) (TMA SWIZZLE_128B)
                    sts128(box_u + lane * 128 + (((j + 4 * (c & 1)) ^ (lane & 7)) << 4), pack2(v[8 * j], v[8 * j + 1]),
                           pack2(v[8 * j + 2], v[8 * j + 3]), pack2(v[8 * j + 4], v[8 * j + 5]), pack2(v[8 * j + 6], v[8 * j + 7]));
                }
                if ((c & 1) == 1) {
                    fence_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&map_res, box_u, n - CHUNK, row0);   // map_res carries the [32 x 64] bf16 box map
                        bulk_commit();
                    }
                }
            }
        }
        if (lane == 0) bulk_wait_all();
    } else if (warp >= EPI_WARP0) {
        constexpr bool GEN = MODE == EPI_GENERIC;
        const bool out_bf16 = GEN ? (e.out_dtype == VRD_BF16) : (MODE == EPI_BF16 || MODE == EPI_BF16_GELU);
        // bf16 outputs without a residual: two 32-column chunks share one [32 x 64] box (128-byte rows, same 4 KB as an fp32
        // [32 x 32] box) and ONE TMA store -- the TMA unit handles requests at a fixed rate, and with 32 small store boxes
        // per 128 x 256 tile on top of the 16 operand loads it, not the tensor pipe, set the pace of the K = 512 GEMMs
        const bool wide = GEN ? (e.wide != 0) : (MODE == EPI_BF16 || MODE == EPI_BF16_GELU);
        const bool has_res = GEN ? (e.has_res != 0) : (MODE == EPI_F32_RES);
        const int act = GEN ? e.act : (MODE == EPI_BF16_GELU ? 2 : 0);
        const bool has_res2 = GEN && e.res2 != nullptr;
        const bool has_corr = GEN && e.corr != nullptr;
        const int dbg = GEN ? e.dbg : 0;
        const int ew = warp - EPI_WARP0;
        const int wq = warp & 3;                         // TMEM lane quarter this warp may access
        const int half = ew >> 2;                        // which half of the tile's columns
        const int n_ch = block_n / (2 * CHUNK);          // chunks of 32 columns per warp and tile (<= 2 when has_res and K is short)
        const uint32_t s_bias_u = smem_u32(s_bias);
        uint8_t* my_stage = staging + ew * 2 * STAGING_BYTES;
        uint8_t* my_stage2 = staging2 + ew * 2 * STAGING_BYTES;
        uint64_t* my_resbar = resbar + 2 * ew;
        int it = 0;
        uint32_t box_cnt = 0;                            // boxes alternate across chunks AND tiles (no-residual path)
        for (int tile = tile0; tile < n_tiles; tile += tile_step, ++it) {
            const int as = it & 1;
            const uint32_t aphase = (it >> 1) & 1;
            const int m0 = (tile / n_tiles_n) * (BLOCK_M * CG) + cta_rank * BLOCK_M, n0 = (tile % n_tiles_n) * block_n;
            const int row0 = m0 + wq * 32, row = row0 + lane;
            const int col0 = n0 + half * (block_n / 2);
            bool valid = row < e.M, add_corr = false;     // a pair's second half can lie past M (loads zero-filled, stores clipped)
            if (e.row_seq != nullptr && valid) {
                const int rl = row % e.R;
                const int seq = e.row_seq[rl];
                valid = seq >= 0;
                if (valid && has_corr) {
                    const int4 si = e.seqinfo[seq];
                    add_corr = (si.z != 0) && (rl - si.x == si.y - 1);
                }
            }
            if (has_res) {
                // the residual boxes do not depend on the accumulator: fetch them while the MMAs of this tile still run
                if (lane == 0) {
                    bulk_wait_read<0>();                 // the previous tile's stores have read both staging boxes
                    for (int c = 0; c < min(n_ch, 2); ++c) {
                        mbar_expect_tx(&my_resbar[c], has_res2 ? 2 * STAGING_BYTES : STAGING_BYTES);
                        tma_load_2d(my_stage + c * STAGING_BYTES, &map_res, &my_resbar[c], col0 + c * CHUNK, row0);
                        if (has_res2) tma_load_2d(my_stage2 + c * STAGING_BYTES, &map_res2, &my_resbar[c], col0 + c * CHUNK, row0);
                    }
                }
                __syncwarp();
            }
            mbar_wait(&tfull[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(wq * 32) << 16) + as * ACC_STAGE_COLS + half * (block_n / 2);
            uint32_t acc[CHUNK];
            tmem_ld32(taddr, acc);                        // chunk c + 1 is always in flight while chunk c is processed
            auto chunk = [&](const int c) {
                const int box_i = has_res ? (c & 1) : (wide ? (int)((box_cnt >> 1) & 1) : (int)(box_cnt & 1));
                const bool box_first = !wide || (c & 1) == 0, box_last = !wide || (c & 1) == 1;
                ++box_cnt;
                uint8_t* box = my_stage + box_i * STAGING_BYTES;
                const uint32_t box_u = smem_u32(box);
                if (has_res) {
                    // box c & 1 is filled once (n_ch <= 2) or twice (n_ch == 4, long-K tiles) per tile
                    mbar_wait(&my_resbar[c & 1], n_ch <= 2 ? (it & 1) : ((c >> 1) & 1));
                } else if (box_first) {
                    if (lane == 0) bulk_wait_read<1>();  // the store issued from this box two boxes ago has read it
                    __syncwarp();
                }
                tmem_ld_wait();
                const int n = col0 + c * CHUNK;
                float v[CHUNK];
#pragma unroll
                for (int i = 0; i < CHUNK / 4; ++i) {
                    const float4 bi = lds128_ro(s_bias_u + (uint32_t)(n + 4 * i) * 4);
                    v[4 * i + 0] = __uint_as_float(acc[4 * i + 0]) + bi.x;
                    v[4 * i + 1] = __uint_as_float(acc[4 * i + 1]) + bi.y;
                    v[4 * i + 2] = __uint_as_float(acc[4 * i + 2]) + bi.z;
                    v[4 * i + 3] = __uint_as_float(acc[4 * i + 3]) + bi.w;
                }
                if (c + 1 < n_ch) {
                    tmem_ld32(taddr + (c + 1) * CHUNK, acc);
                } else {                                 // accumulator fully read: hand the TMEM stage back to the MMA warp
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        if constexpr (CG == 2) mbar_arrive_cluster(mapa_shared(smem_u32(&tempty[as]), 0));   // the leader issues the MMAs
                        else mbar_arrive(&tempty[as]);
                    }
                }
                if (dbg & 2) return;
                if (add_corr) {
#pragma unroll
                    for (int i = 0; i < CHUNK; ++i) v[i] += s_corr[n + i];
                }
                if (act == 1) {
#pragma unroll
                    for (int i = 0; i < CHUNK; ++i) v[i] = fmaxf(v[i], 0.f);
                } else if (act == 2 && !out_bf16) {
#pragma unroll
                    for (int i = 0; i < CHUNK; ++i) v[i] = gelu_fast(v[i]);
                }
                if (has_res) {
                    // residual box: 128-byte rows, 16-byte chunk j of row r stored at chunk j ^ (r & 7) (TMA SWIZZLE_128B)
#pragma unroll
                    for (int j = 0; j < CHUNK / 4; ++j) {
                        const float4 r = lds128(box_u + lane * 128 + ((j ^ (lane & 7)) << 4));
                        v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
                    }
                }
                if (has_res2) {
                    // second residual box, same layout.  (Row-strided global loads here -- one L1 wavefront per row -- made the SOS
                    // output projections 44 % slower than the single-residual ones: 247 vs 171 us at equal shapes.)
                    const uint32_t box2_u = smem_u32(my_stage2 + (c & 1) * STAGING_BYTES);
#pragma unroll
                    for (int j = 0; j < CHUNK / 4; ++j) {
                        const float4 r = lds128(box2_u + lane * 128 + ((j ^ (lane & 7)) << 4));
                        v[4 * j] += r.x; v[4 * j + 1] += r.y; v[4 * j + 2] += r.z; v[4 * j + 3] += r.w;
                    }
                }
                if (!valid) {
#pragma unroll
                    for (int i = 0; i < CHUNK; ++i) v[i] = 0.f;
                }
                if (out_bf16) {
                    // 64-byte rows, chunk j of row r stored at chunk j ^ ((r >> 1) & 3) (TMA SWIZZLE_64B)
#pragma unroll
                    for (int j = 0; j < CHUNK / 8; ++j) {
                        uint4 pk;
                        if (act == 2) {          // GELU GEMMs have no residual: the activation is the last step (gelu(0) = 0)
                            pk.x = gelu_pair_bf16(v[8 * j], v[8 * j + 1]); pk.y = gelu_pair_bf16(v[8 * j + 2], v[8 * j + 3]);
                            pk.z = gelu_pair_bf16(v[8 * j + 4], v[8 * j + 5]); pk.w = gelu_pair_bf16(v[8 * j + 6], v[8 * j + 7]);
                        } else {
                            pk.x = pack2(v[8 * j], v[8 * j + 1]); pk.y = pack2(v[8 * j + 2], v[8 * j + 3]);
                            pk.z = pack2(v[8 * j + 4], v[8 * j + 5]); pk.w = pack2(v[8 * j + 6], v[8 * j + 7]);
                        }
                        if (wide)   // 128-byte rows: chunk jj of row r stored at chunk jj ^ (r & 7) (TMA SWIZZLE_128B)
                            sts128(box_u + lane * 128 + (((j + 4 * (c & 1)) ^ (lane & 7)) << 4), pk.x, pk.y, pk.z, pk.w);
                        else
                            sts128(box_u + lane * 64 + ((j ^ ((lane >> 1) & 3)) << 4), pk.x, pk.y, pk.z, pk.w);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < CHUNK / 4; ++j)
                        sts128(box_u + lane * 128 + ((j ^ (lane & 7)) << 4), __float_as_uint(v[4 * j]), __float_as_uint(v[4 * j + 1]),
                               __float_as_uint(v[4 * j + 2]), __float_as_uint(v[4 * j + 3]));
                }
                if (!box_last) return;                   // the second chunk of a wide box completes it
                fence_async_smem();                      // make the generic-proxy writes visible to the TMA engine
                __syncwarp();
                if (lane == 0) {
                    if (!(dbg & 1)) {
                        if (wide) tma_store_2d(&map_res, box_u, n - CHUNK, row0);   // map_res carries the [32 x 64] bf16 box map
                        else tma_store_2d(&map_out, box_u, n, row0);
                    }
                    bulk_commit();
                    if (has_res && c + 2 < n_ch) {       // refill this box with the residual of chunk c + 2 (tiles with long K only:
                        bulk_wait_read<0>();             // the MMAs of the next tile hide this latency)
                        mbar_expect_tx(&my_resbar[c & 1], has_res2 ? 2 * STAGING_BYTES : STAGING_BYTES);
                        tma_load_2d(box, &map_res, &my_resbar[c & 1], col0 + (c + 2) * CHUNK, row0);
                        if (has_res2)
                            tma_load_2d(my_stage2 + (c & 1) * STAGING_BYTES, &map_res2, &my_resbar[c & 1], col0 + (c + 2) * CHUNK, row0);
                    }
                }
            };
            if constexpr (BN != 0) {
#pragma unroll
                for (int c = 0; c < BN / (2 * CHUNK); ++c) chunk(c);
            } else {
#pragma unroll 1
                for (int c = 0; c < n_ch; ++c) chunk(c);
            }
        }
        if (lane == 0) bulk_wait_all();                  // all output boxes written before the CTA retires
    }
    tc_fence_before();
    if constexpr (CG == 2) cluster_sync_all(); else __syncthreads();   // neither CTA of a pair retires while the other still uses it
    if (warp == 2) {
        tc_fence_after();
        if constexpr (CG == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS));
    }
}

typedef void (*GemmKernel)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, EpiArgs, int, int,
                           int, int, int);

GemmKernel pick_kernel(int cg, int mode, int bn) {
#define VRD_PICK(M, B) (cg == 1 ? (GemmKernel)gemm_tcgen05_kernel<1, M, B> : (GemmKernel)gemm_tcgen05_kernel<2, M, B>)
    if (mode == EPI_GENERIC && bn == 0) return VRD_PICK(EPI_GENERIC, 0);
    if (mode == EPI_BF16 && bn == 256) return VRD_PICK(EPI_BF16, 256);
    if (mode == EPI_BF16_GELU && bn == 256) return VRD_PICK(EPI_BF16_GELU, 256);
    if (mode == EPI_F32_RES && bn == 128) return VRD_PICK(EPI_F32_RES, 128);
    if (mode == EPI_F32_RES && bn == 256) return VRD_PICK(EPI_F32_RES, 256);
    if (mode == EPI_LN && bn == 512) return VRD_PICK(EPI_LN, 512);
    if (mode == EPI_LN_RES && bn == 512) return cg == 2 ? (GemmKernel)gemm_tcgen05_kernel<2, EPI_LN_RES, 512> : nullptr;
#undef VRD_PICK
    return nullptr;
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

// 2-D row-major matrix [rows, cols] with row pitch ld (elements); box = [box_rows, box_cols].
bool make_map(CUtensorMap* map, const void* ptr, CUtensorMapDataType dt, int esize, long long rows, long long cols, long long ld,
              int box_rows, int box_cols, CUtensorMapSwizzle swz) {
    EncodeTiledFn enc = get_encode();
    if (enc == nullptr) { snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled entry point not found"); return false; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * esize};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, dt, 2, const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { snprintf(g_err, sizeof g_err, "cuTensorMapEncodeTiled failed: %d", (int)r); return false; }
    return true;
}

}  // namespace

const char* gemm_tcgen05_error() { return g_err; }

bool make_tensor_map_2d(CUtensorMap* map, const void* ptr, CUtensorMapDataType dt, int esize, long long rows, long long cols, long long ld,
                        int box_rows, int box_cols, CUtensorMapSwizzle swz) {
    return make_map(map, ptr, dt, esize, rows, cols, ld, box_rows, box_cols, swz);
}

static int gemm_tcgen05_launch(const GemmArgs& g, cudaStream_t st, int split) {
    static PerDeviceOnce attr_once;      // the shared-memory attribute and the SM count belong to a device, not to the process
    const int num_sms = device_sm_count();
    static const bool wide_ok = !(getenv("VRD_GEMM_WIDE") != nullptr && atoi(getenv("VRD_GEMM_WIDE")) == 0);   // A/B switch
    static const int dbg = getenv("VRD_GEMM_DBG") ? atoi(getenv("VRD_GEMM_DBG")) : 0;
    const bool spec_ok = vrd_options().gemm_spec != 0;   // A/B switch: 0 = every launch through the generic epilogue
    static int force_cg = -1;
    if (force_cg < 0) { const char* v = getenv("VRD_GEMM_CG"); force_cg = v ? atoi(v) : 0; }
    if (g.M % BLOCK_M != 0 || g.K % BLOCK_K != 0 || g.N % 64 != 0 || g.lda % 8 != 0 || ((uintptr_t)g.A & 15) != 0) {
        snprintf(g_err, sizeof g_err, "gemm_tcgen05: unsupported shape M=%d N=%d K=%d lda=%lld (need M%%128, K%%64, N%%64 == 0)",
                 g.M, g.N, g.K, g.lda);
        return 1;
    }
    const bool has_res = g.res1 != nullptr;
    if (g.res2 != nullptr && !has_res) { snprintf(g_err, sizeof g_err, "gemm_tcgen05: res2 without res1"); return 1; }
    const bool ln = g.ln_gamma != nullptr;
    const bool ln_res = ln && has_res;      // y = acc + bias + res (fp32, g.out) and LayerNorm(y) (bf16, g.ln_out)
    if (ln && !ln_res && (g.N != 512 || g.out_dtype != VRD_BF16 || g.act != 0 || g.ln_beta == nullptr || split)) {
        snprintf(g_err, sizeof g_err, "gemm_tcgen05: the LayerNorm epilogue needs N = 512, a bf16 output and no activation");
        return 1;
    }
    if (ln_res && (g.N != 512 || g.out_dtype != VRD_F32 || g.act != 0 || g.ln_beta == nullptr || split || g.res2 != nullptr ||
                   g.corr != nullptr || g.ln_out == nullptr || (g.ld_ln * 2) % 16 != 0 || ((uintptr_t)g.ln_out & 15) != 0)) {
        snprintf(g_err, sizeof g_err, "gemm_tcgen05: the residual + LayerNorm epilogue needs N = 512, an fp32 output, one residual, "
                 "no activation and a 16-byte aligned bf16 LayerNorm output");
        return 1;
    }
    // experiment (gemm_spec = 3): bf16 (optionally GELU) GEMMs with N = 512 on 512-column tiles through the EPI_LN kernel without the
    // LayerNorm (A is read once).  Measured slower than the 256-column double-buffered tiles: 1007 vs 1029 TFLOP/s at K = 512, 1081 vs
    // 1136 at K = 1024 (GELU) -- only the K >= 1536 embedding convs, which also lose a whole LayerNorm launch, gain from the wide tile.
    const bool wide512 = !ln && vrd_options().gemm_spec == 3 && g.N == 512 && g.out_dtype == VRD_BF16 && !has_res &&
                         (g.act == 0 || g.act == 2) && g.corr == nullptr && !split && dbg == 0;
    int block_n;
    if (ln || wide512) {
        block_n = 512;          // (ln_res included)
    } else if (has_res) {   // two 32-column chunks per epilogue warp (both residual boxes prefetched), four when K is long enough to hide
        if (g.N % 256 == 0 && g.taps * g.K >= 1024 && g.res2 == nullptr) block_n = 256;   // a second residual doubles the staging boxes
        else if (g.N % 128 == 0) block_n = 128;
        else if (g.N <= 128) block_n = g.N;
        else { snprintf(g_err, sizeof g_err, "gemm_tcgen05: N=%d with a residual must be <= 128 or a multiple of 128", g.N); return 1; }
    } else {
        if (g.N % 256 == 0) block_n = 256;
        else if (g.N <= 256) block_n = g.N;
        else if (g.N % 128 == 0) block_n = 128;
        else { snprintf(g_err, sizeof g_err, "gemm_tcgen05: unsupported N=%d", g.N); return 1; }
    }
    if (g.N > MAX_N) { snprintf(g_err, sizeof g_err, "gemm_tcgen05: N=%d exceeds %d", g.N, MAX_N); return 1; }
    const int osize = g.out_dtype == VRD_BF16 ? 2 : 4;
    if ((g.ldo * osize) % 16 != 0 || ((uintptr_t)g.out & 15) != 0 || (has_res && ((g.ldr1 % 4) != 0 || ((uintptr_t)g.res1 & 15) != 0)) ||
        (g.res2 && ((g.ldr2 % 4) != 0 || ((uintptr_t)g.res2 & 15) != 0))) {
        snprintf(g_err, sizeof g_err, "gemm_tcgen05: output/residual base and pitch must be 16-byte aligned");
        return 1;
    }
    if (has_res && g.out_dtype != VRD_F32) { snprintf(g_err, sizeof g_err, "gemm_tcgen05: residual needs an fp32 output"); return 1; }
    // CTA pairs whenever there are enough rows to keep every pair busy (the small query-decoder GEMMs keep one CTA per tile):
    // measured +6..10 % on the long-K GEMMs, +4..9 % on the K = 512 ones, neutral on the HBM-bound residual projections
    // (the residual + LayerNorm kernel exists for CTA pairs only and is used for every M: results must not depend on the chunking)
    const int cg = ln_res ? 2 : (force_cg == 1 ? 1 : ((force_cg == 2 || g.M >= 128 * 2 * 64) ? 2 : 1));
    CUtensorMap map_a, map_w, map_out, map_res, map_res2;
    const long long kk = (long long)g.taps * g.K * (split ? 2 : 1);
    if (!make_map(&map_a, g.A, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.M, split ? 2 * g.K : g.K, g.lda, BLOCK_M, BLOCK_K, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    // EPI_LN: two boxes of block_n / (2 cg) W rows per stage (output channels [0, 256) and [256, 512))
    if (!make_map(&map_w, g.W, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.N, kk, kk, (ln || wide512) ? block_n / (2 * cg) : block_n / cg, BLOCK_K, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    if (g.out_dtype == VRD_BF16) {
        if (!make_map(&map_out, g.out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.M, g.N, g.ldo, 32, CHUNK, CU_TENSOR_MAP_SWIZZLE_64B)) return 1;
    } else {
        if (!make_map(&map_out, g.out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.M, g.N, g.ldo, 32, CHUNK, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    }
    if (has_res) {
        if (!make_map(&map_res, g.res1, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.M, g.N, g.ldr1, 32, CHUNK, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    } else if ((wide_ok || ln || wide512) && g.out_dtype == VRD_BF16 && (block_n / (2 * CHUNK)) % 2 == 0) {
        // no residual: the slot carries the [32 rows x 64 cols] box map of the wide bf16 epilogue
        if (!make_map(&map_res, g.out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.M, g.N, g.ldo, 32, 2 * CHUNK, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    } else {
        map_res = map_out;
    }
    if (ln_res) {   // the slot carries the [32 rows x 64 cols] bf16 box map of the LayerNorm output
        if (!make_map(&map_res2, g.ln_out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.M, g.N, g.ld_ln, 32, 2 * CHUNK, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    } else if (g.res2 != nullptr) {
        if (!make_map(&map_res2, g.res2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.M, g.N, g.ldr2, 32, CHUNK, CU_TENSOR_MAP_SWIZZLE_128B)) return 1;
    } else {
        map_res2 = map_res;
    }
    const int stage_bytes = A_STAGE_BYTES + (block_n / cg) * BLOCK_K * 2;
    // alignment slack, staging, barriers, bias + corr (EPI_LN: one box per epilogue warp; the statistics exchange buffer sits in the
    // unused half of the correction vector's slot)
    const int fixed = ln_res ? 1024 + STAGING_TOTAL + 1024 + 2 * MAX_N * 4
                     : (ln || wide512) ? 1024 + STAGING_TOTAL / 2 + 1024 + 2 * MAX_N * 4
                         : 1024 + (g.res2 != nullptr ? 2 : 1) * STAGING_TOTAL + 1024 + 2 * MAX_N * 4;
    int stages = (SMEM_LIMIT - fixed) / stage_bytes;
    if (stages > MAX_STAGES) stages = MAX_STAGES;
    if (stages < 2) { snprintf(g_err, sizeof g_err, "gemm_tcgen05: tile does not fit in shared memory (N=%d, residuals=%d)", g.N, has_res + (g.res2 != nullptr)); return 1; }
    const int smem = fixed + stages * stage_bytes;
    if (attr_once.first()) {
        for (int c = 1; c <= 2; ++c)
            for (int m = 0; m < 6; ++m)
                for (int bn = 0; bn <= 512; bn += 128) {
                    GemmKernel k = pick_kernel(c, m, bn);
                    if (k != nullptr && cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_LIMIT) != cudaSuccess) {
                        snprintf(g_err, sizeof g_err, "cudaFuncSetAttribute(max dynamic smem) failed");
                        return 1;
                    }
                }
    }
    const int wide = (!has_res && wide_ok && g.out_dtype == VRD_BF16 && (block_n / (2 * CHUNK)) % 2 == 0) ? 1 : 0;
    EpiArgs e{g.bias, g.out_dtype, g.M, g.N, g.act, has_res ? 1 : 0, g.res2, g.ldr2, g.corr, g.row_seq, g.seqinfo, g.R, wide, dbg,
              g.ln_gamma, g.ln_beta, g.ln_relu};
    // specialised epilogues for the shapes that carry the forward (q / k / v and other bf16 projections, the GELU MLP-up GEMMs, the
    // fp32 residual projections); everything else (ReLU, pad correction, second residual, narrow tiles, experiments) stays generic
    int mode = EPI_GENERIC, bn_ct = 0;
    if (ln_res) {
        mode = EPI_LN_RES; bn_ct = 512;
    } else if (ln || wide512) {
        mode = EPI_LN; bn_ct = 512;
    } else if (spec_ok && dbg == 0 && g.corr == nullptr && g.res2 == nullptr) {
        if (wide && block_n == 256 && g.act == 0) { mode = EPI_BF16; bn_ct = 256; }
        else if (wide && block_n == 256 && g.act == 2) { mode = EPI_BF16_GELU; bn_ct = 256; }
        // the fp32 residual projections are HBM-bound: the specialised epilogue measured -2 % on them (gemm_spec = 2 selects it)
        else if (vrd_options().gemm_spec == 2 && has_res && g.act == 0 && (block_n == 128 || block_n == 256)) { mode = EPI_F32_RES; bn_ct = block_n; }
    }
    GemmKernel kern = pick_kernel(cg, mode, bn_ct);
    if (kern == nullptr) { snprintf(g_err, sizeof g_err, "gemm_tcgen05: no kernel for cg=%d mode=%d bn=%d", cg, mode, bn_ct); return 1; }
    const int n_tiles = ((g.M + BLOCK_M * cg - 1) / (BLOCK_M * cg)) * (g.N / block_n);
    const int max_groups = num_sms / cg;
    const int grid = cg * (n_tiles < max_groups ? n_tiles : max_groups);
    if (cg == 1) {
        if (launch_k(kern, dim3(grid), dim3(NUM_THREADS), smem, st, map_a, map_w, map_out, map_res, map_res2, e, g.K, g.taps, block_n,
                     stages, split) != cudaSuccess) {
            snprintf(g_err, sizeof g_err, "launch of gemm_tcgen05_kernel<1> failed: %s", cudaGetErrorString(cudaGetLastError()));
            return 1;
        }
        return 0;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(NUM_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if (cudaLaunchKernelEx(&cfg, kern, map_a, map_w, map_out, map_res, map_res2, e, g.K, g.taps, block_n, stages, split) != cudaSuccess) {
        snprintf(g_err, sizeof g_err, "cluster launch of gemm_tcgen05_kernel<2> failed: %s", cudaGetErrorString(cudaGetLastError()));
        return 1;
    }
    return 0;
}

int gemm_tcgen05_bf16(const GemmArgs& g, cudaStream_t st) { return gemm_tcgen05_launch(g, st, 0); }

bool gemm_res_ln_fused_ok(const GemmArgs& g) {
    return vrd_options().proj_ln != 0 && g.N == 512 && g.M % BLOCK_M == 0 && g.K % BLOCK_K == 0 && g.taps == 1 && g.res1 != nullptr &&
           g.res2 == nullptr && g.corr == nullptr && g.act == 0;
}

// ---- fp32 operands on the tensor cores: 3 x bf16 split -------------------------------------------------------------------
namespace {

// x [rows, taps * K] fp32 (pitch ldx) -> out [rows, taps * 2K] bf16: per tap [hi(K) | lo(K)], hi = bf16(x), lo = bf16(x - hi)
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ x, long long ldx, long long rows, int K, int taps,
                                                         __nv_bfloat16* __restrict__ out) {
    pdl_wait();
    const int per_row = taps * K / 4;
    const long long total = rows * per_row;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / per_row;
        const int c = (int)(i - r * per_row) * 4;
        const int tap = c / K, k = c - tap * K;
        const float4 v = *reinterpret_cast<const float4*>(x + r * ldx + c);
        const __nv_bfloat162 h0 = __floats2bfloat162_rn(v.x, v.y), h1 = __floats2bfloat162_rn(v.z, v.w);
        const float2 f0 = __bfloat1622float2(h0), f1 = __bfloat1622float2(h1);
        const __nv_bfloat162 l0 = __floats2bfloat162_rn(v.x - f0.x, v.y - f0.y), l1 = __floats2bfloat162_rn(v.z - f1.x, v.w - f1.y);
        __nv_bfloat16* o = out + r * (2LL * taps * K) + (long long)tap * 2 * K + k;
        uint2 hi, lo;
        hi.x = *reinterpret_cast<const uint32_t*>(&h0); hi.y = *reinterpret_cast<const uint32_t*>(&h1);
        lo.x = *reinterpret_cast<const uint32_t*>(&l0); lo.y = *reinterpret_cast<const uint32_t*>(&l1);
        *reinterpret_cast<uint2*>(o) = hi;
        *reinterpret_cast<uint2*>(o + K) = lo;
    }
}

struct Scratch { void* p = nullptr; size_t bytes = 0; };

// grow-only device scratch; cudaFree synchronises the device, so kernels still reading the old block have finished
void* scratch_get(Scratch& s, size_t need) {
    if (s.bytes >= need) return s.p;
    if (s.p != nullptr) cudaFree(s.p);
    s.p = nullptr;
    s.bytes = 0;
    const size_t want = need + need / 4 + (1 << 20);
    if (cudaMalloc(&s.p, want) != cudaSuccess) { s.p = nullptr; return nullptr; }
    s.bytes = want;
    return s.p;
}

}  // namespace

// D = A W^T for fp32 A / W through the same tcgen05 kernel: both operands are split into bf16 (hi, lo) pairs by a streaming
// kernel (scratch: 4 bytes per operand element) and multiplied as A_hi W_hi + A_lo W_hi + A_hi W_lo with fp32 accumulation in
// TMEM -- relative error ~2^-16 per product instead of the 2^-9 of plain bf16 operands (SURVEY section 7 hard-part 3: tcgen05 has
// no fp32 MMA).  Returns 1 when the shape does not fit the tensor-core kernel (the caller falls back to the CUDA-core GEMM).
int gemm_tcgen05_f32split(const GemmArgs& g, cudaStream_t st) {
    static Scratch sa[64], sw[64];
    if (g.M % BLOCK_M != 0 || g.K % BLOCK_K != 0 || g.N % 64 != 0 || g.lda % 4 != 0 || ((uintptr_t)g.A & 15) != 0 || ((uintptr_t)g.W & 15) != 0)
        return 1;
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 1;
    const size_t a_bytes = (size_t)g.M * 2 * g.K * 2, w_bytes = (size_t)g.N * g.taps * 2 * g.K * 2;
    __nv_bfloat16* a2 = (__nv_bfloat16*)scratch_get(sa[dev], a_bytes);
    __nv_bfloat16* w2 = (__nv_bfloat16*)scratch_get(sw[dev], w_bytes);
    if (a2 == nullptr || w2 == nullptr) { snprintf(g_err, sizeof g_err, "gemm_tcgen05_f32split: scratch allocation failed"); return 2; }
    const int sms = device_sm_count();
    launch_k(split_bf16_kernel, dim3(sms * 8), dim3(256), 0, st, (const float*)g.A, g.lda, (long long)g.M, g.K, 1, a2);
    launch_k(split_bf16_kernel, dim3(sms * 2), dim3(256), 0, st, (const float*)g.W, (long long)g.taps * g.K, (long long)g.N, g.K, g.taps, w2);
    GemmArgs h = g;
    h.A = a2; h.lda = 2LL * g.K; h.W = w2;
    return gemm_tcgen05_launch(h, st, 1) == 0 ? 0 : 2;
}

}  // namespace vrd
