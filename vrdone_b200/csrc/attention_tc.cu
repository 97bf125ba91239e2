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
// QK -> softmax -> PV chain.  Persistent grid: CTA c takes items c, c + grid, ... of the layout's query-tile list (layout.py:
// one entry per 128-row tile of every pair, longest pairs first) x heads.
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
constexpr int FA_SMEM = 1024 + (2 + FA_SLOTS) * FA_SLOT_BYTES + FA_P_BYTES + 256;
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

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < FA_SLOTS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&s_full[i], 1); }
        mbar_init(p_full, 4);
        mbar_init(pv_done, 1);
        mbar_fence_init();
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

    // Every role walks the same flat sequence of key blocks: item i = (query tile i / n_head, head i % n_head) of the layout's tile
    // list for i = blockIdx.x, blockIdx.x + gridDim.x, ...; per item its blocks of 64 keys.  The list is sorted by pair length
    // (longest first), so this static round-robin hands every CTA the same mix of long and short items (a long pair's tile
    // costs 10 x a short one's: any assignment by position in the layout leaves most CTAs idle behind a few).  g counts blocks,
    // n_item counts items: they carry the barrier phases.  The next item's tile is fetched one item ahead.
    struct Cursor {
        const int4* tiles; int n_items, n_head, stride, i, h, j, n_kv; int4 tile, pre;
        __device__ __forceinline__ void fetch() {
            const int nx = i + stride;
            if (nx < n_items) pre = __ldg(tiles + nx / n_head);
        }
        __device__ __forceinline__ void start(const int4* t, int n_tiles, int nh, int first, int step) {
            tiles = t; n_items = n_tiles * nh; n_head = nh; stride = step; i = first; j = 0;
            if (i < n_items) { tile = __ldg(t + i / nh); h = i % nh; n_kv = (tile.z + FA_BN - 1) / FA_BN; fetch(); }
        }
        __device__ __forceinline__ bool valid() const { return i < n_items; }
        __device__ __forceinline__ bool last_of_item() const { return j == n_kv - 1; }
        __device__ __forceinline__ void next() {
            if (++j == n_kv) {
                j = 0;
                i += stride;
                if (i < n_items) { tile = pre; h = i % n_head; n_kv = (tile.z + FA_BN - 1) / FA_BN; fetch(); }
            }
        }
    };
    uint32_t g = 0, n_item = 0;

    // ---------------- TMA producer (warp 4, one lane) ----------------
    auto produce = [&](Cursor c) {
        for (; c.valid(); c.next(), ++g) {
            if (c.j == 0) {
                const int qb = n_item & 1;
                mbar_wait(&q_empty[qb], ((n_item >> 1) & 1) ^ 1);
                mbar_expect_tx(&q_full[qb], FA_SLOT_BYTES);
                tma_load_2d(q_tiles + qb * FA_SLOT_BYTES, &map_q, &q_full[qb], c.h * FA_HS, c.tile.x);
                ++n_item;
            }
            const int slot = g % FA_SLOTS;
            mbar_wait(&empty[slot], ((g / FA_SLOTS) & 1) ^ 1);
            mbar_expect_tx(&full[slot], FA_SLOT_BYTES);
            uint8_t* dst = ring + slot * FA_SLOT_BYTES;
            tma_load_2d(dst, &map_k, &full[slot], c.h * FA_HS, c.tile.y + c.j * FA_BN);
            tma_load_2d(dst + FA_SLOT_BYTES / 2, &map_v, &full[slot], c.h * FA_HS, c.tile.y + c.j * FA_BN);
        }
    };

    // ---------------- MMA issuer (warp 5, one lane) ----------------
    // Software-pipelined over the flat block sequence: S of block g + 1 (also across an item boundary) is issued before the
    // wait for P of block g, so the softmax warps find their next scores ready.  S buffer (g + 1) & 1 was last read by the
    // softmax of block g - 1, which arrived on p_full before PV_{g-1} was issued by this thread: free by program order.
    constexpr uint32_t idesc_s = make_idesc_bf16(FA_BM, FA_BN, false);
    constexpr uint32_t idesc_o = make_idesc_bf16(FA_BM, FA_HS, true);      // V tile [keys, dims]: MN-major B
    const uint64_t pdesc = make_smem_desc_sw128(smem_u32(p_tile));
    uint32_t n_item_qk = 0;                                                 // items whose first S has been issued
    auto issue_qk = [&](const Cursor& c, uint32_t gg) {
        if (c.j == 0) {
            mbar_wait(&q_full[n_item_qk & 1], (n_item_qk >> 1) & 1);
            ++n_item_qk;
        }
        const int qb = (n_item_qk - 1) & 1;
        const int slot = gg % FA_SLOTS;
        mbar_wait(&full[slot], (gg / FA_SLOTS) & 1);
        tc_fence_after();
        const uint64_t qdesc = make_smem_desc_sw128(smem_u32(q_tiles + qb * FA_SLOT_BYTES));
        const uint64_t kdesc = make_smem_desc_sw128(smem_u32(ring + slot * FA_SLOT_BYTES));
        const uint32_t tmem_s = tmem_base + (gg & 1) * FA_BN;
#pragma unroll
        for (int k = 0; k < FA_HS / 16; ++k) umma_f16_ss(tmem_s, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[gg & 1]);
        if (c.last_of_item()) umma_commit(&q_empty[qb]);                    // the Q tile is free once the last S of the item is done
    };
    auto mma = [&](Cursor c) {
        if (!c.valid()) return;
        Cursor nx = c;
        issue_qk(c, g);
        nx.next();
        for (; c.valid(); ++g) {
            if (nx.valid()) issue_qk(nx, g + 1);
            mbar_wait(p_full, g & 1);
            tc_fence_after();
            // O += P V : A = P [128 x 64 keys] K-major (+32 bytes per 16 keys), B = V [64 keys x 64 dims] MN-major
            // (+16 key rows = 2048 bytes per step)
            const int slot = g % FA_SLOTS;
            const uint64_t vdesc = make_smem_desc_sw128(smem_u32(ring + slot * FA_SLOT_BYTES + FA_SLOT_BYTES / 2));
#pragma unroll
            for (int k = 0; k < FA_BN / 16; ++k) umma_f16_ss(tmem_o, pdesc + 2 * k, vdesc + 128 * k, idesc_o, (c.j | k) != 0 ? 1u : 0u);
            umma_commit(&empty[slot]);
            umma_commit(pv_done);
            c = nx;
            nx.next();
        }
    };

    // ---------------- softmax warps (0-3): thread = query row ----------------
    const int trow = (warp & 3) * 32 + lane;                                    // row of the tile = TMEM lane
    const uint32_t lane_sel = (uint32_t)((warp & 3) * 32) << 16;
    const uint32_t p_row = smem_u32(p_tile) + (trow >> 3) * 1024 + (trow & 7) * 128;
    // O / l -> bf16 -> global for the rows of a finished item (valid query rows only); the caller has waited for its last PV
    auto epilogue = [&](int row0, int len, int q0, int h, float l) {
        uint32_t orr[FA_HS];
        tmem_ld32(tmem_o + lane_sel, orr);
        tmem_ld32(tmem_o + lane_sel + 32, orr + 32);
        tmem_ld_wait();
        tc_fence_before();                   // orders these TMEM reads before the next item's first PV (released through p_full)
        if (q0 + trow < len) {
            const float inv = 1.0f / l;
            uint4* dst = reinterpret_cast<uint4*>(out + (long long)(row0 + trow) * ld + h * FA_HS);
#pragma unroll
            for (int c8 = 0; c8 < FA_HS / 8; ++c8) {
                uint4 v;
                v.x = pack_bf16x2(__uint_as_float(orr[8 * c8]) * inv, __uint_as_float(orr[8 * c8 + 1]) * inv);
                v.y = pack_bf16x2(__uint_as_float(orr[8 * c8 + 2]) * inv, __uint_as_float(orr[8 * c8 + 3]) * inv);
                v.z = pack_bf16x2(__uint_as_float(orr[8 * c8 + 4]) * inv, __uint_as_float(orr[8 * c8 + 5]) * inv);
                v.w = pack_bf16x2(__uint_as_float(orr[8 * c8 + 6]) * inv, __uint_as_float(orr[8 * c8 + 7]) * inv);
                dst[c8] = v;
            }
        }
    };
    auto softmax = [&](Cursor c) {
        float m = -INFINITY, l = 0.f;
        bool have_prev = false;
        int pv_row0 = 0, pv_len = 0, pv_q0 = 0, pv_h = 0;
        float pv_l = 1.f;
        for (; c.valid(); c.next(), ++g) {
            mbar_wait(&s_full[g & 1], (g >> 1) & 1);
            tc_fence_after();
            const uint32_t tmem_s = tmem_base + (g & 1) * FA_BN + lane_sel;
            uint32_t sr[FA_BN];
            tmem_ld32(tmem_s, sr);
            tmem_ld32(tmem_s + 32, sr + 32);
            if (c.j == 0) { m = -INFINITY; l = 0.f; }
            const int nk = c.tile.z - c.j * FA_BN;                              // valid keys of this block (>= 1)
            tmem_ld_wait();
            float mx = -INFINITY;
            if (nk >= FA_BN) {
#pragma unroll
                for (int i = 0; i < FA_BN; ++i) mx = fmaxf(mx, __uint_as_float(sr[i]));
            } else {
#pragma unroll
                for (int i = 0; i < FA_BN; ++i) {
                    if (i >= nk) sr[i] = 0xff800000u;                           // -inf: keys past the end of the pair
                    mx = fmaxf(mx, __uint_as_float(sr[i]));
                }
            }
            // Lazy rescaling (exact arithmetic, fewer TMEM round trips): the running maximum only moves when some row of the warp
            // outgrew it by more than 2^8 -- until then P = exp2(s - m_old) <= 256 and the sums stay well inside fp32 / bf16
            // range; the final division by l uses the same m, so the result is the softmax either way.
            const bool first = c.j == 0;
            const bool grow = first || __any_sync(FULL_MASK, (mx - m) * FA_LOG2E > 8.0f);
            float alpha = 1.0f;
            if (grow) {
                const float m_new = fmaxf(m, mx);
                alpha = ex2f((m - m_new) * FA_LOG2E);                           // 0 for the first block (m = -inf)
                m = m_new;
            }
            if (g > 0) {
                mbar_wait(pv_done, (g - 1) & 1);                                // O holds every earlier block; the P tile is free
                tc_fence_after();
            }
            if (first) {
                if (have_prev) epilogue(pv_row0, pv_len, pv_q0, pv_h, pv_l);    // deferred: its last PV ran behind this block's S phase
            } else if (grow) {                                                  // rescale the accumulator rows
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
            const float ms = m * FA_LOG2E;
            float sum = 0.f;
            const int n_chunks = (min(nk, FA_BN) + 7) >> 3;                     // chunks of 8 keys that hold a valid key (warp-uniform)
#pragma unroll
            for (int c8 = 0; c8 < FA_BN / 8; ++c8) {
                // 16-byte chunk c8 of row r sits at chunk c8 ^ (r & 7) (128-byte swizzle, as TMA / the MMA descriptor expect)
                const uint32_t addr = p_row + ((uint32_t)(c8 ^ (trow & 7)) << 4);
                if (c8 < n_chunks) {
                    float p[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) { p[i] = ex2f(fmaf(__uint_as_float(sr[8 * c8 + i]), FA_LOG2E, -ms)); sum += p[i]; }
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(pack_bf16x2(p[0], p[1])),
                                 "r"(pack_bf16x2(p[2], p[3])), "r"(pack_bf16x2(p[4], p[5])), "r"(pack_bf16x2(p[6], p[7])) : "memory");
                } else {                                                        // keys past the pair: P = 0 (V rows there belong to others)
                    asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(addr), "r"(0u) : "memory");
                }
            }
            l = l * alpha + sum;
            fence_async_smem();              // P: generic-proxy writes -> visible to the tensor core's async proxy
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
            if (c.last_of_item()) {
                have_prev = true;
                pv_row0 = c.tile.x; pv_len = c.tile.z; pv_q0 = c.tile.x - c.tile.y; pv_h = c.h; pv_l = l;
            }
        }
        if (have_prev) {                                                        // the round's last item
            mbar_wait(pv_done, (g - 1) & 1);
            tc_fence_after();
            epilogue(pv_row0, pv_len, pv_q0, pv_h, pv_l);
        }
    };

    Cursor c;
    c.start(lay.tiles, lay.n_tiles, n_head, blockIdx.x, gridDim.x);
    if (warp == 4) { if (lane == 0) produce(c); }
    else if (warp == 5) { if (lane == 0) mma(c); }
    else softmax(c);
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
    if (C != n_head * FA_HS || ld % 8 != 0 || lay.R % 128 != 0 || lay.tiles == nullptr || lay.n_tiles <= 0) return 1;
    if ((((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)out) & 15) != 0) return 1;
    CUtensorMap mq, mk, mv;
    if (!make_tensor_map_2d(&mq, q, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, lay.R, C, ld, FA_BM, FA_HS, CU_TENSOR_MAP_SWIZZLE_128B)) return 2;
    if (!make_tensor_map_2d(&mk, k, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, lay.R, C, ld, FA_BN, FA_HS, CU_TENSOR_MAP_SWIZZLE_128B)) return 2;
    if (!make_tensor_map_2d(&mv, v, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, lay.R, C, ld, FA_BN, FA_HS, CU_TENSOR_MAP_SWIZZLE_128B)) return 2;
    static PerDeviceOnce once;
    if (once.first() && cudaFuncSetAttribute(flash_attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FA_SMEM) != cudaSuccess) return 2;
    const int n_items = lay.n_tiles * n_head;
    const int max_ctas = 2 * device_sm_count();
    const int grid = n_items < max_ctas ? n_items : max_ctas;
    flash_attn_tc_kernel<<<grid, FA_THREADS, FA_SMEM, st>>>(mq, mk, mv, (__nv_bfloat16*)out, ld, lay, n_head);
    return 0;
}

}  // namespace vrd
