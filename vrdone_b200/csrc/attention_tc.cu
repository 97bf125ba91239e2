// Varlen full attention inside each pair on the 5th-generation tensor cores -- the Subject-Object-Synergy attention of the bf16
// path (reference models/local_transformer.py:170-183: softmax(Q K^T / sqrt(hs), key mask) @ V, scores never materialised in
// HBM).  head_dim 64; q, k, v, out are [R, C] bf16 matrices of the packed row layout (head h = columns [64 h, 64 h + 64)); the
// 1 / sqrt(hs) scale is folded into the query projection.
//
// Work item = (pair, head, query tile of 128 rows aligned to the pair's first row): results do not depend on where a pair's rows
// sit in the layout.  Keys / values stream in blocks of 64 rows.  Per block:
//     S[128, 64]  = Q K_j^T          tcgen05.mma (M 128, N 64, K 64: 4 instructions), accumulator in TMEM columns [0, 64) / [64, 128)
//     softmax     thread = query row: tcgen05.ld of its 64 scores, running max / sum in registers (no shuffles), P = exp2(.) as bf16
//                 into a 128-byte-swizzled shared-memory tile; when the running max of any row of the warp grew, the O
//                 accumulator is rescaled in TMEM (tcgen05.ld / .st) -- exact online softmax, no approximation threshold
//     O[128, 64] += P V_j            tcgen05.mma with V as an MN-major B operand straight from its TMA tile, accumulator in
//                 TMEM columns [128, 192)
// One CTA = 6 warps: 4 softmax warps (TMEM lane quarters 0-3), one TMA producer warp (two Q buffers alternating between items and
// a ring of three K/V slots: the next item's tiles arrive while this one is computed), one MMA issuer warp.  S is double
// buffered in TMEM, so Q K_{j+1}^T (and the first S of the next item) is issued while the softmax warps still work on block j.
// TWO co-resident CTAs per SM (97 KB of shared memory and 256 TMEM columns each) hide what is left of the serial
// QK -> softmax -> PV chain.  Persistent grid over (64-row block of the layout, head) units: a CTA handles the query tiles that
// START in its block (every role warp re-derives the same list from row_seq / seqinfo with a ballot).
// Separator rows of the output are not written (the consumer is a projection GEMM whose epilogue zeroes them).
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace vrd {

namespace {

using namespace tc;

constexpr int FA_BM = 128;                 // query rows per item (UMMA M)
constexpr int FA_BN = 64;                  // keys per block (UMMA N of S, K of P V)
constexpr int FA_HS = 64;                  // head dim
constexpr int FA_SLOTS = 3;                // K/V ring
constexpr int FA_SLOT_BYTES = FA_BM * FA_HS * 2;          // 16 KB: a Q tile, or K (8 KB) + V (8 KB) of one block
constexpr int FA_P_BYTES = FA_BM * FA_BN * 2;             // 16 KB
constexpr int FA_THREADS = 192;
constexpr int FA_TMEM_COLS = 256;          // S0 [0, 64), S1 [64, 128), O [128, 192)
constexpr int FA_SMEM = 1024 + (2 + FA_SLOTS) * FA_SLOT_BYTES + FA_P_BYTES + 256 + 192 * 12;
constexpr float FA_LOG2E = 1.4426950408889634f;

struct FaItem { int row0, off, len, q0; };   // first layout row of the tile, pair's first row, pair length, tile's first row in the pair

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int FA_LIST = 192;                // query tiles a CTA collects per round (a 64-row block starts at most 32 tiles)

// The query tiles that START in the 64-row block ``b``: appended to the CTA's shared-memory list by one warp.  A tile is
// (first layout row, pair's first row, pair length).
__device__ __forceinline__ void fa_scan_block(const Lay& lay, int b, int lane, int3* list, int* count) {
    const int r0 = b * 64 + lane, r1 = r0 + 32;
    const int p0 = lay.row_seq[r0], p1 = lay.row_seq[r1];
    int4 i0 = make_int4(0, 0, 0, 0), i1 = make_int4(0, 0, 0, 0);
    bool s0 = false, s1 = false;
    if (p0 >= 0) { i0 = lay.seqinfo[p0]; s0 = ((r0 - i0.x) & (FA_BM - 1)) == 0; }
    if (p1 >= 0) { i1 = lay.seqinfo[p1]; s1 = ((r1 - i1.x) & (FA_BM - 1)) == 0; }
    const unsigned m0 = __ballot_sync(FULL_MASK, s0), m1 = __ballot_sync(FULL_MASK, s1);
    const int n0 = __popc(m0), n = n0 + __popc(m1);
    int base = 0;
    if (lane == 0 && n > 0) base = atomicAdd(count, n);
    base = __shfl_sync(FULL_MASK, base, 0);
    if (s0) list[base + __popc(m0 & ((1u << lane) - 1))] = make_int3(r0, i0.x, i0.y);
    if (s1) list[base + n0 + __popc(m1 & ((1u << lane) - 1))] = make_int3(r1, i1.x, i1.y);
}

__global__ void __launch_bounds__(FA_THREADS, 2)
flash_attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, __nv_bfloat16* __restrict__ out, long long ld, Lay lay, int n_head) {
    extern __shared__ __align__(1024) uint8_t fa_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)fa_smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* q_tiles = smem;                                       // [2] Q tiles, alternating between items
    uint8_t* ring = smem + 2 * FA_SLOT_BYTES;                      // [FA_SLOTS] K | V blocks
    uint8_t* p_tile = ring + FA_SLOTS * FA_SLOT_BYTES;
    uint64_t* bars = (uint64_t*)(p_tile + FA_P_BYTES);
    uint64_t* full = bars;                     // [FA_SLOTS]  TMA -> MMA
    uint64_t* empty = bars + FA_SLOTS;         // [FA_SLOTS]  MMA (commit) -> TMA
    uint64_t* q_full = bars + 2 * FA_SLOTS;    // [2]
    uint64_t* q_empty = q_full + 2;            // [2]
    uint64_t* s_full = q_full + 4;             // [2] MMA (commit) -> softmax: S of block g in TMEM buffer g & 1
    uint64_t* p_full = q_full + 6;             // softmax -> MMA: P_g in shared memory, S_g consumed, O rescaled
    uint64_t* pv_done = q_full + 7;            // MMA (commit) -> softmax: O accumulated through block g, P tile free
    uint32_t* tmem_slot = (uint32_t*)(q_full + 8);
    int* list_count = (int*)(tmem_slot + 1);
    int3* list = (int3*)(tmem_slot + 4);                           // [FA_LIST] query tiles of the current round

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < FA_SLOTS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&s_full[i], 1); }
        mbar_init(p_full, 4);
        mbar_init(pv_done, 1);
        mbar_fence_init();
        *list_count = 0;
    }
    if (warp == 4) {
        tmem_alloc(tmem_slot, FA_TMEM_COLS);
        if (lane == 0) { prefetch_tensormap(&map_q); prefetch_tensormap(&map_k); prefetch_tensormap(&map_v); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + 2 * FA_BN;

    // per-role state (persists across rounds)
    int slot = 0; uint32_t phase = 0;       // K/V ring position (producer and MMA issuer each keep their own copy)
    uint32_t n_item = 0, g = 0;             // items / key blocks processed so far: barrier phase counters

    // ---------------- TMA producer (warp 4) ----------------
    auto producer_item = [&](const FaItem& it, int h) {
            if (lane == 0) {
                const int n_kv = (it.len + FA_BN - 1) / FA_BN;
                const int qb = n_item & 1;
                mbar_wait(&q_empty[qb], ((n_item >> 1) & 1) ^ 1);
                mbar_expect_tx(&q_full[qb], FA_SLOT_BYTES);
                tma_load_2d(q_tiles + qb * FA_SLOT_BYTES, &map_q, &q_full[qb], h * FA_HS, it.row0);
                for (int j = 0; j < n_kv; ++j) {
                    mbar_wait(&empty[slot], phase ^ 1);
                    mbar_expect_tx(&full[slot], FA_SLOT_BYTES);
                    uint8_t* dst = ring + slot * FA_SLOT_BYTES;
                    tma_load_2d(dst, &map_k, &full[slot], h * FA_HS, it.off + j * FA_BN);
                    tma_load_2d(dst + FA_SLOT_BYTES / 2, &map_v, &full[slot], h * FA_HS, it.off + j * FA_BN);
                    if (++slot == FA_SLOTS) { slot = 0; phase ^= 1; }
                }
            }
            ++n_item;
            __syncwarp();
    };
    // ---------------- MMA issuer (warp 5) ----------------
    constexpr uint32_t idesc_s = make_idesc_bf16(FA_BM, FA_BN, false);
    constexpr uint32_t idesc_o = make_idesc_bf16(FA_BM, FA_HS, true);      // V tile [keys, dims]: MN-major B
    const uint64_t pdesc = make_smem_desc_sw128(smem_u32(p_tile));
    auto mma_item = [&](const FaItem& it, int h) {
            if (lane == 0) {
                const int n_kv = (it.len + FA_BN - 1) / FA_BN;
                const int qb = n_item & 1;
                mbar_wait(&q_full[qb], (n_item >> 1) & 1);
                const uint64_t qdesc = make_smem_desc_sw128(smem_u32(q_tiles + qb * FA_SLOT_BYTES));
                // S_g = Q K^T into TMEM buffer g & 1.  That buffer was last read by the softmax of block g - 2, which arrived on
                // p_full before PV_{g-2} was issued by this thread: free by program order.
                auto issue_qk = [&](uint32_t gg, int kv, bool last) {
                    tc_fence_after();
                    const uint64_t kdesc = make_smem_desc_sw128(smem_u32(ring + kv * FA_SLOT_BYTES));
                    const uint32_t tmem_s = tmem_base + (gg & 1) * FA_BN;
#pragma unroll
                    for (int k = 0; k < FA_HS / 16; ++k) umma_f16_ss(tmem_s, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
                    umma_commit(&s_full[gg & 1]);
                    if (last) umma_commit(&q_empty[qb]);                        // the Q tile is free once the last S of the item is done
                };
                int kv_cur = slot;
                mbar_wait(&full[kv_cur], phase);
                if (++slot == FA_SLOTS) { slot = 0; phase ^= 1; }
                issue_qk(g, kv_cur, n_kv == 1);
                for (int j = 0; j < n_kv; ++j, ++g) {
                    int kv_next = -1;
                    if (j + 1 < n_kv) {                                         // next block's scores while the softmax warps work on this one
                        kv_next = slot;
                        mbar_wait(&full[kv_next], phase);
                        if (++slot == FA_SLOTS) { slot = 0; phase ^= 1; }
                        issue_qk(g + 1, kv_next, j + 2 == n_kv);
                    }
                    mbar_wait(p_full, g & 1);
                    tc_fence_after();
                    // O += P V : A = P [128 x 64 keys] K-major (+32 bytes per 16 keys), B = V [64 keys x 64 dims] MN-major
                    // (+16 key rows = 2048 bytes per step)
                    const uint64_t vdesc = make_smem_desc_sw128(smem_u32(ring + kv_cur * FA_SLOT_BYTES + FA_SLOT_BYTES / 2));
#pragma unroll
                    for (int k = 0; k < FA_BN / 16; ++k) umma_f16_ss(tmem_o, pdesc + 2 * k, vdesc + 128 * k, idesc_o, (j | k) != 0 ? 1u : 0u);
                    umma_commit(&empty[kv_cur]);
                    umma_commit(pv_done);
                    kv_cur = kv_next;
                }
            }
            ++n_item;
            __syncwarp();
    };
    // ---------------- softmax warps (0-3): thread = query row ----------------
    const int trow = (warp & 3) * 32 + lane;                                    // row of the tile = TMEM lane
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t p_row = smem_u32(p_tile) + (trow >> 3) * 1024 + (trow & 7) * 128;
    auto softmax_item = [&](const FaItem& it, int h) {
            const int n_kv = (it.len + FA_BN - 1) / FA_BN;
            float m = -INFINITY, l = 0.f;
            for (int j = 0; j < n_kv; ++j, ++g) {
                mbar_wait(&s_full[g & 1], (g >> 1) & 1);
                tc_fence_after();
                const uint32_t tmem_s = tmem_base + (g & 1) * FA_BN + lane_sel;
                uint32_t sr[FA_BN];
                tmem_ld32(tmem_s, sr);
                tmem_ld32(tmem_s + 32, sr + 32);
                tmem_ld_wait();
                const int nk = it.len - j * FA_BN;                              // valid keys of this block (>= 1)
                float mx = m;
                if (nk >= FA_BN) {
#pragma unroll
                    for (int i = 0; i < FA_BN; ++i) mx = fmaxf(mx, __uint_as_float(sr[i]));
                } else {
#pragma unroll
                    for (int i = 0; i < FA_BN; ++i) {
                        if (i >= nk) sr[i] = 0xff800000u;                       // -inf: keys past the end of the pair
                        mx = fmaxf(mx, __uint_as_float(sr[i]));
                    }
                }
                const float alpha = ex2f((m - mx) * FA_LOG2E);                  // 0 for the first block (m = -inf)
                if (j > 0) {
                    mbar_wait(pv_done, (g - 1) & 1);                            // O holds blocks < j; the P tile may be overwritten
                    tc_fence_after();
                    if (__any_sync(FULL_MASK, mx > m)) {                        // exact online softmax: rescale the accumulator rows
                        uint32_t orr[FA_HS];
                        tmem_ld32(tmem_o + lane_sel, orr);
                        tmem_ld32(tmem_o + lane_sel + 32, orr + 32);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < FA_HS; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * alpha);
                        tmem_st32(tmem_o + lane_sel, orr);
                        tmem_st32(tmem_o + lane_sel + 32, orr + 32);
                        tmem_st_wait();
                    }
                }
                const float ms = mx * FA_LOG2E;
                float sum = 0.f;
#pragma unroll
                for (int c = 0; c < FA_BN / 8; ++c) {
                    float p[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { p[i] = ex2f(fmaf(__uint_as_float(sr[8 * c + i]), FA_LOG2E, -ms)); sum += p[i]; }
                    // 16-byte chunk c of row r sits at chunk c ^ (r & 7) (128-byte swizzle, as TMA / the MMA descriptor expect)
                    const uint32_t addr = p_row + ((uint32_t)(c ^ (trow & 7)) << 4);
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(p[0], p[1])),
                                 "r"(pack_bf16x2(p[2], p[3])), "r"(pack_bf16x2(p[4], p[5])), "r"(pack_bf16x2(p[6], p[7])) : "memory");
                }
                l = l * alpha + sum;
                m = mx;
                fence_async_smem();          // P: generic-proxy writes -> visible to the tensor core's async proxy
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(p_full);
            }
            // epilogue: O / l -> bf16 -> global (valid query rows only)
            mbar_wait(pv_done, (g - 1) & 1);
            tc_fence_after();
            uint32_t orr[FA_HS];
            tmem_ld32(tmem_o + lane_sel, orr);
            tmem_ld32(tmem_o + lane_sel + 32, orr + 32);
            tmem_ld_wait();
            tc_fence_before();               // orders these TMEM reads before the next item's MMAs (released through p_full)
            const int qrow = it.q0 + trow;
            if (qrow < it.len) {
                const float inv = 1.0f / l;
                uint4* dst = reinterpret_cast<uint4*>(out + (long long)(it.row0 + trow) * ld + h * FA_HS);
#pragma unroll
                for (int c = 0; c < FA_HS / 8; ++c) {
                    uint4 v;
                    v.x = pack_bf16x2(__uint_as_float(orr[8 * c]) * inv, __uint_as_float(orr[8 * c + 1]) * inv);
                    v.y = pack_bf16x2(__uint_as_float(orr[8 * c + 2]) * inv, __uint_as_float(orr[8 * c + 3]) * inv);
                    v.z = pack_bf16x2(__uint_as_float(orr[8 * c + 4]) * inv, __uint_as_float(orr[8 * c + 5]) * inv);
                    v.w = pack_bf16x2(__uint_as_float(orr[8 * c + 6]) * inv, __uint_as_float(orr[8 * c + 7]) * inv);
                    dst[c] = v;
                }
            }
    };

    // Rounds: the six warps each scan one 64-row block of the layout for query-tile starts (one round of global-load latency for
    // up to six blocks instead of two dependent loads in front of every item), then every role walks the same list.
    const int n_blocks = lay.R / 64;
    for (int base = blockIdx.x; base < n_blocks; base += 6 * gridDim.x) {
        const int b = base + warp * gridDim.x;
        if (b < n_blocks) fa_scan_block(lay, b, lane, list, list_count);
        __syncthreads();
        const int n_tiles = *list_count;
        for (int t = 0; t < n_tiles; ++t) {
            const int3 tile = list[t];
            FaItem it;
            it.row0 = tile.x; it.off = tile.y; it.len = tile.z; it.q0 = tile.x - tile.y;
            for (int h = 0; h < n_head; ++h) {
                if (warp == 4) producer_item(it, h);
                else if (warp == 5) mma_item(it, h);
                else softmax_item(it, h);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) *list_count = 0;
        __syncthreads();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 4) {
        tc_fence_after();
        tmem_dealloc(tmem_base, FA_TMEM_COLS);
    }
}

}  // namespace

// Returns 0 on success, 1 when the shape is not supported (the caller falls back to the mma.sync kernel), 2 on a launch error.
int full_attn_tcgen05(const void* q, const void* k, const void* v, void* out, long long ld, Lay lay, int n_head, int C, cudaStream_t st) {
    if (C != n_head * FA_HS || ld % 8 != 0 || lay.R % 128 != 0) return 1;
    if ((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) != 0) return 1;
    CUtensorMap mq, mk, mv;
    if (!make_tensor_map_2d(&mq, q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, lay.R, C, ld, FA_BM, FA_HS, CU_TENSOR_MAP_SWIZZLE_128B)) return 2;
    if (!make_tensor_map_2d(&mk, k, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, lay.R, C, ld, FA_BN, FA_HS, CU_TENSOR_MAP_SWIZZLE_128B)) return 2;
    if (!make_tensor_map_2d(&mv, v, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, lay.R, C, ld, FA_BN, FA_HS, CU_TENSOR_MAP_SWIZZLE_128B)) return 2;
    static PerDeviceOnce once;
    if (once.first() && cudaFuncSetAttribute(flash_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM) != cudaSuccess) return 2;
    const int n_blocks = lay.R / 64;
    const int max_ctas = 2 * device_sm_count();
    const int grid = n_blocks < max_ctas ? n_blocks : max_ctas;
    flash_attn_tc_kernel<<<grid, FA_THREADS, FA_SMEM, st>>>(mq, mk, mv, (__nv_bfloat16*)out, ld, lay, n_head);
    return 0;
}

}  // namespace vrd
