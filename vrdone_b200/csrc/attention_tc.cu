// Varlen full attention inside each pair on the 5th-generation tensor cores -- the Subject-Object-Synergy attention of the bf16
// path (reference models/local_transformer.py:170-183: softmax(Q K^T / sqrt(hs), key mask) @ V, scores never materialised in
// HBM).  head_dim 64; q, k, v, out are [R, C] bf16 matrices of the packed row layout (head h = columns [64 h, 64 h + 64)); the
// 1 / sqrt(hs) scale is folded into the query projection.
//
// Work item = (pair, head, query tile of 128 rows aligned to the pair's first row): results do not depend on where a pair's rows
// sit in the layout.  Keys / values stream in blocks of 64 rows.  Per block g:
//     S_g[128, 64]  = Q K^T          tcgen05.mma (SS: M 128, N 64, K 64 = 4 instructions), accumulator in TMEM buffer g & 1
//     softmax        thread = (query row, part of the 64 keys): tcgen05.ld of its scores, running max / sum in registers,
//                    P = exp2(.) rounded to bf16 and stored with tcgen05.st into TMEM buffer g & 1 (two keys per 32-bit column,
//                    the layout tcgen05.mma expects for an A operand in tensor memory)
//     O[128, 64]   += P V            tcgen05.mma (TS: A = P from TMEM, B = V as an MN-major operand straight from its TMA tile)
// The running maximum is exact but lazy: the accumulator rows are rescaled (tcgen05.ld / .st of O) only when a row outgrew its
// maximum by more than 2^8, so the softmax warps wait for the tensor core only then and at item boundaries.
//
// One CTA: softmax warps (NSPLIT per TMEM lane quarter, 64 / NSPLIT keys each) + one TMA producer warp (two Q buffers alternating
// between items, a ring of K/V slots: the next item's tiles arrive while this one is computed) + one MMA issuer warp, software-
// pipelined over the flat block sequence: S_{g+1} (also across an item boundary) is issued before the wait for P_g.  TMEM (256
// columns): S0 S1 (64 + 64), O (64), P0 P1 (32 + 32).  Two co-resident CTAs per SM.  Persistent grid: CTA c takes items c,
// c + grid, ... of the layout's query-tile list (layout.py: one entry per 128-row tile of every pair, LONGEST pairs first -- a
// long pair's tile costs 10 x a short one's, and any assignment by position leaves most CTAs idle behind a few) x heads.
// Separator rows of the output are not written (the consumer is a projection GEMM whose epilogue zeroes them).
#include <cstdlib>
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace vrd {

namespace {

using namespace tc;

constexpr int FA_BM = 128;                 // query rows per item (UMMA M)
constexpr int FA_BN = 64;                  // keys per block (UMMA N of S, K of P V)
constexpr int FA_HS = 64;                  // head dim
constexpr int FA_SLOT_BYTES = FA_BM * FA_HS * 2;          // 16 KB: a Q tile, or K (8 KB) + V (8 KB) of one block
// Two configurations (template parameter NCTA = co-resident CTAs per SM):
//   NCTA 2  TMEM 256 columns: S0 S1 | O | P0 P1, four K/V slots; S_{g+1} is computed while the softmax of block g runs.
//   NCTA 3  TMEM 128 columns: S | O with P ALIASED onto the first 32 columns of S (the scores are in registers by the time P is
//           written), two K/V slots; QK -> softmax -> PV of a CTA are strictly serial, and THREE such streams share the SM.  The
//           kernel is bound by the latency of one block's chain (ncu: issue slots 47 %, the softmax warps wait or do bookkeeping
//           70 % of their time), so a third independent stream pays more than overlapping QK inside a stream.
template <int NCTA> struct FaCfg;
template <> struct FaCfg<2> { static constexpr int SLOTS = 4, TMEM_COLS = 256, TM_S = 0, TM_O = 128, TM_P = 192, S_BUFS = 2; };
template <> struct FaCfg<3> { static constexpr int SLOTS = 2, TMEM_COLS = 128, TM_S = 0, TM_O = 64, TM_P = 0, S_BUFS = 1; };
template <int NCTA> constexpr int fa_smem() { return 1024 + (2 + FaCfg<NCTA>::SLOTS) * FA_SLOT_BYTES + 256 + 3 * 2 * FA_BM * 4; }
constexpr float FA_LOG2E = 1.4426950408889634f;

__device__ __forceinline__ float ex2f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
// D[tmem] (+)= A[tmem] * B[smem descriptor]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
        ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
          "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
template <int N> __device__ __forceinline__ void tmem_ld_n(uint32_t taddr, uint32_t* r) {
    if constexpr (N == 64) { tmem_ld32(taddr, r); tmem_ld32(taddr + 32, r + 32); }
    else if constexpr (N == 32) tmem_ld32(taddr, r);
    else tmem_ld16(taddr, r);
}
template <int N> __device__ __forceinline__ void tmem_st_n(uint32_t taddr, const uint32_t* r) {
    if constexpr (N == 64) { tmem_st32(taddr, r); tmem_st32(taddr + 32, r + 32); }
    else if constexpr (N == 32) tmem_st32(taddr, r);
    else tmem_st16(taddr, r);
}

// Every role walks the same flat sequence of key blocks: item i = (query tile i / n_head, head i % n_head) of the layout's tile
// list for i = blockIdx.x, blockIdx.x + gridDim.x, ...; per item its blocks of 64 keys.  The next item's tile is fetched one
// item ahead.
struct FaCursor {
    const int4* tiles; int n_items, n_head, nh_shift, stride, i, h, j, n_kv; int4 tile, pre;
    // i / n_head and i % n_head: a shift and a mask when the head count is a power of two (it is 8 in every config; the integer
    // division cost ~40 instructions per item on the softmax warps' critical path, and an item is only ~2 key blocks long)
    __device__ __forceinline__ int tile_of(int it) const { return nh_shift >= 0 ? (it >> nh_shift) : it / n_head; }
    __device__ __forceinline__ int head_of(int it) const { return nh_shift >= 0 ? (it & (n_head - 1)) : it % n_head; }
    // three scalar loads, not one 128-bit load: the vector load lands in four consecutive temporaries and the moves into `pre`
    // waited for it on the spot (2 % of all stall samples sat on that move), scalar loads go straight to their registers and the
    // latency stays hidden until the next item starts
    __device__ __forceinline__ void fetch() {
        const int nx = i + stride;
        if (nx < n_items) {
            const int* p = reinterpret_cast<const int*>(tiles + tile_of(nx));
            pre.x = __ldg(p); pre.y = __ldg(p + 1); pre.z = __ldg(p + 2);
        }
    }
    __device__ __forceinline__ void start(const int4* t, int n_tiles, int nh, int first, int step) {
        tiles = t; n_items = n_tiles * nh; n_head = nh; stride = step; i = first; j = 0; h = 0; n_kv = 1;
        nh_shift = (nh & (nh - 1)) == 0 ? 31 - __clz(nh) : -1;
        if (i < n_items) { tile = __ldg(t + tile_of(i)); h = head_of(i); n_kv = (tile.z + FA_BN - 1) / FA_BN; fetch(); }
    }
    __device__ __forceinline__ bool valid() const { return i < n_items; }
    __device__ __forceinline__ bool last_of_item() const { return j == n_kv - 1; }
    __device__ __forceinline__ void next() {
        if (++j == n_kv) {
            j = 0;
            i += stride;
            if (i < n_items) { tile = pre; h = head_of(i); n_kv = (tile.z + FA_BN - 1) / FA_BN; fetch(); }
        }
    }
};

template <int NSPLIT, int NCTA>
__global__ void __launch_bounds__((4 * NSPLIT + 2) * 32, NCTA)
flash_attn_tc_kernel(const __grid_constant__ CUtensorMap map_q, const __grid_constant__ CUtensorMap map_k,
                     const __grid_constant__ CUtensorMap map_v, __nv_bfloat16* __restrict__ out, long long ld, Lay lay, int n_head,
                     int fa_serial_safe) {
    using Cfg = FaCfg<NCTA>;
    constexpr int FA_SLOTS = Cfg::SLOTS, FA_TMEM_COLS = Cfg::TMEM_COLS, FA_TM_S = Cfg::TM_S, FA_TM_O = Cfg::TM_O, FA_TM_P = Cfg::TM_P;
    constexpr bool SERIAL = Cfg::S_BUFS == 1;                      // one S buffer: no QK lookahead, P aliased onto S
    constexpr int N_SOFTMAX = 4 * NSPLIT;                          // softmax warps; then the producer warp and the MMA warp
    constexpr int HC = FA_BN / NSPLIT;                             // keys (and output dims) per softmax thread
    extern __shared__ __align__(1024) uint8_t fa_smem_raw[];
    uint8_t* smem = (uint8_t*)(((uintptr_t)fa_smem_raw + 1023) & ~(uintptr_t)1023);
    uint8_t* q_tiles = smem;                                       // [2] Q tiles, alternating between items
    uint8_t* ring = smem + 2 * FA_SLOT_BYTES;                      // [FA_SLOTS] K | V blocks
    uint64_t* bars = (uint64_t*)(ring + FA_SLOTS * FA_SLOT_BYTES);
    uint64_t* full = bars;                     // [FA_SLOTS]  TMA -> MMA
    uint64_t* empty = bars + FA_SLOTS;         // [FA_SLOTS]  MMA (commit) -> TMA
    uint64_t* q_full = bars + 2 * FA_SLOTS;    // [2]
    uint64_t* q_empty = q_full + 2;            // [2]
    uint64_t* s_full = q_full + 4;             // [2] MMA (commit) -> softmax: S of block g in TMEM buffer g & 1
    uint64_t* p_full = q_full + 6;             // softmax -> MMA: P_g in TMEM buffer g & 1, S_g consumed, O rescaled if need be
    uint64_t* pv_done = q_full + 7;            // MMA (commit) -> softmax: O accumulated through block g, P buffer g & 1 free
    uint32_t* tmem_slot = (uint32_t*)(q_full + 8);
    float* xch = (float*)(tmem_slot + 4);      // [3][NSPLIT][FA_BM]: partial row maxima (two block parities) and row sums

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < FA_SLOTS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
        for (int i = 0; i < 2; ++i) { mbar_init(&q_full[i], 1); mbar_init(&q_empty[i], 1); mbar_init(&s_full[i], 1); }
        mbar_init(p_full, N_SOFTMAX);
        mbar_init(pv_done, 1);
        mbar_fence_init();
    }
    if (warp == N_SOFTMAX) {
        tmem_alloc(tmem_slot, FA_TMEM_COLS);
        if (lane == 0) { prefetch_tensormap(&map_q); prefetch_tensormap(&map_k); prefetch_tensormap(&map_v); }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t tmem_o = tmem_base + FA_TM_O;
    pdl_wait();                             // q / k / v and the tile list are the previous kernels' outputs: nothing of them was touched above

    uint32_t g = 0, n_item = 0;             // key blocks / items processed so far: they carry the barrier phases

    // ---------------- TMA producer (one lane) ----------------
    auto produce = [&](FaCursor c) {
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

    // ---------------- MMA issuer (one lane) ----------------
    // S buffer (g + 1) & 1 was last read by the softmax of block g - 1, which arrived on p_full before PV_{g-1} was issued by this
    // thread: free by program order.  The same holds for the P buffers (written by the softmax warps only after pv_done of the
    // block that last read them).
    constexpr uint32_t idesc_s = make_idesc_bf16(FA_BM, FA_BN, false);
    constexpr uint32_t idesc_o = make_idesc_bf16(FA_BM, FA_HS, true);      // V tile [keys, dims]: MN-major B
    uint32_t n_item_qk = 0;                                                 // items whose first S has been issued
    auto issue_qk = [&](const FaCursor& c, uint32_t gg) {
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
        const uint32_t tmem_s = tmem_base + FA_TM_S + (SERIAL ? 0 : (gg & 1) * FA_BN);
#pragma unroll
        for (int k = 0; k < FA_HS / 16; ++k) umma_f16_ss(tmem_s, qdesc + 2 * k, kdesc + 2 * k, idesc_s, k != 0 ? 1u : 0u);
        umma_commit(&s_full[SERIAL ? 0 : (gg & 1)]);
        if (c.last_of_item()) umma_commit(&q_empty[qb]);                    // the Q tile is free once the last S of the item is done
    };
    auto mma = [&](FaCursor c) {
        if (!c.valid()) return;
        FaCursor nx = c;
        issue_qk(c, g);
        nx.next();
        for (; c.valid(); ++g) {
            // one S buffer: S_{g+1} overwrites the columns P_g lives in, so it is issued behind PV_g (the tensor core executes a
            // CTA's MMAs in issue order); with two buffers it is issued now and runs beside the softmax of block g
            if (!SERIAL && nx.valid()) issue_qk(nx, g + 1);
            mbar_wait(p_full, g & 1);
            tc_fence_after();
            // O += P V : A = P_g from TMEM (16 keys = 8 columns per step), B = V [64 keys x 64 dims] MN-major (+16 key rows =
            // 2048 bytes per step)
            const int slot = g % FA_SLOTS;
            const uint64_t vdesc = make_smem_desc_sw128(smem_u32(ring + slot * FA_SLOT_BYTES + FA_SLOT_BYTES / 2));
            const uint32_t tmem_p = tmem_base + FA_TM_P + (SERIAL ? 0 : (g & 1) * (FA_BN / 2));
#pragma unroll
            for (int k = 0; k < FA_BN / 16; ++k) umma_f16_ts(tmem_o, tmem_p + 8 * k, vdesc + 128 * k, idesc_o, (c.j | k) != 0 ? 1u : 0u);
            umma_commit(&empty[slot]);
            umma_commit(pv_done);
            if (SERIAL && nx.valid()) {
                if (fa_serial_safe) { mbar_wait(pv_done, g & 1); tc_fence_after(); }   // experiment: do not rely on issue order
                issue_qk(nx, g + 1);
            }
            c = nx;
            nx.next();
        }
    };

    // ---------------- softmax warps ----------------
    // NSPLIT == 2: warps q and q + 4 share TMEM lane quarter q and split the block's keys (and the output dims); the halves of a
    // row agree on the running maximum through shared memory and a 64-thread named barrier per block, their partial row sums
    // meet once per item.
    const int quarter = warp & 3, part = (warp >> 2) % NSPLIT;
    const int trow = quarter * 32 + lane;                                       // row of the tile = TMEM lane
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const uint32_t col0 = part * HC;
    auto pair_sync = [&]() { if constexpr (NSPLIT > 1) asm volatile("bar.sync %0, %1;" ::"r"(1 + quarter), "n"(32 * NSPLIT) : "memory"); };
    // O / l -> bf16 -> global for the rows of a finished item (valid query rows only); the caller has waited for its last PV
    auto epilogue = [&](int row0, int len, int q0, int h, float l) {
        uint32_t orr[HC];
        tmem_ld_n<HC>(tmem_o + lane_sel + col0, orr);
        if constexpr (NSPLIT > 1) {
            float* xl = xch + 2 * NSPLIT * FA_BM;
            xl[part * FA_BM + trow] = l;
            pair_sync();
            l += xl[(part ^ 1) * FA_BM + trow];
        }
        tmem_ld_wait();
        tc_fence_before();                   // orders these TMEM reads before the next item's first PV (released through p_full)
        pair_sync();                         // the partner has read this thread's sum before the next item overwrites it
        if (q0 + trow < len) {
            const float inv = 1.0f / l;
            uint4* dst = reinterpret_cast<uint4*>(out + (long long)(row0 + trow) * ld + h * FA_HS + col0);
#pragma unroll
            for (int c8 = 0; c8 < HC / 8; ++c8) {
                uint4 v;
                v.x = pack_bf16x2(__uint_as_float(orr[8 * c8]) * inv, __uint_as_float(orr[8 * c8 + 1]) * inv);
                v.y = pack_bf16x2(__uint_as_float(orr[8 * c8 + 2]) * inv, __uint_as_float(orr[8 * c8 + 3]) * inv);
                v.z = pack_bf16x2(__uint_as_float(orr[8 * c8 + 4]) * inv, __uint_as_float(orr[8 * c8 + 5]) * inv);
                v.w = pack_bf16x2(__uint_as_float(orr[8 * c8 + 6]) * inv, __uint_as_float(orr[8 * c8 + 7]) * inv);
                dst[c8] = v;
            }
        }
    };
    // the same in two halves of 32 dims: the 3-CTA configuration lives in 96 registers
    auto epilogue_halves = [&](int row0, int len, int q0, int h, float l) {
        const float inv = 1.0f / l;
        const bool store = q0 + trow < len;
        uint4* dst = reinterpret_cast<uint4*>(out + (long long)(row0 + trow) * ld + h * FA_HS);
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            uint32_t orr[32];
            tmem_ld32(tmem_o + lane_sel + 32 * hf, orr);
            tmem_ld_wait();
            if (store) {
#pragma unroll
                for (int c8 = 0; c8 < 4; ++c8) {
                    uint4 v;
                    v.x = pack_bf16x2(__uint_as_float(orr[8 * c8]) * inv, __uint_as_float(orr[8 * c8 + 1]) * inv);
                    v.y = pack_bf16x2(__uint_as_float(orr[8 * c8 + 2]) * inv, __uint_as_float(orr[8 * c8 + 3]) * inv);
                    v.z = pack_bf16x2(__uint_as_float(orr[8 * c8 + 4]) * inv, __uint_as_float(orr[8 * c8 + 5]) * inv);
                    v.w = pack_bf16x2(__uint_as_float(orr[8 * c8 + 6]) * inv, __uint_as_float(orr[8 * c8 + 7]) * inv);
                    dst[4 * hf + c8] = v;
                }
            }
        }
        tc_fence_before();                   // orders these TMEM reads before the next item's first PV (released through p_full)
    };
    auto softmax = [&](FaCursor c) {
        float m = -INFINITY, l = 0.f;
        bool have_prev = false;
        int pv_row0 = 0, pv_len = 0, pv_q0 = 0, pv_h = 0;
        float pv_l = 1.f;
        uint32_t pv_waited = 0;              // pv_done phases [0, pv_waited) are known to be complete
        auto wait_pv = [&](uint32_t upto) {  // blocks < upto have been accumulated into O (and their P buffers are free)
            if (upto > pv_waited) {
                mbar_wait(pv_done, (upto - 1) & 1);
                tc_fence_after();
                pv_waited = upto;
            }
        };
        for (; c.valid(); c.next(), ++g) {
            if constexpr (SERIAL) mbar_wait(&s_full[0], g & 1); else mbar_wait(&s_full[g & 1], (g >> 1) & 1);
            tc_fence_after();
            if constexpr (SERIAL) {
                // S_g was issued behind the previous item's last PV: O is final.  Written out before the scores are loaded (registers).
                if (c.j == 0 && have_prev) { pv_waited = g; epilogue_halves(pv_row0, pv_len, pv_q0, pv_h, pv_l); }
            }
            uint32_t sr[HC];
            tmem_ld_n<HC>(tmem_base + FA_TM_S + (SERIAL ? 0 : (g & 1) * FA_BN) + lane_sel + col0, sr);
            if (c.j == 0) { m = -INFINITY; l = 0.f; }
            const int nk = c.tile.z - c.j * FA_BN - (int)col0;                  // valid keys among this thread's columns (may be <= 0)
            tmem_ld_wait();
            if (nk < HC) {
#pragma unroll
                for (int i = 0; i < HC; ++i)
                    if (i >= nk) sr[i] = 0xff800000u;                           // -inf: keys past the end of the pair
            }
            float mxa[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};        // four independent chains (a single one is 64 dependent ops)
#pragma unroll
            for (int i = 0; i < HC; i += 4) {
#pragma unroll
                for (int u = 0; u < 4; ++u) mxa[u] = fmaxf(mxa[u], __uint_as_float(sr[i + u]));
            }
            float mx = fmaxf(fmaxf(mxa[0], mxa[1]), fmaxf(mxa[2], mxa[3]));
            if constexpr (NSPLIT > 1) {
                float* xm = xch + (g & 1) * NSPLIT * FA_BM;
                xm[part * FA_BM + trow] = mx;
                pair_sync();
                mx = fmaxf(mx, xm[(part ^ 1) * FA_BM + trow]);                  // the row's maximum over all 64 keys (finite: >= 1 valid key)
            }
            // Lazy rescaling (exact arithmetic, fewer TMEM round trips): the running maximum only moves when some row of the warp
            // outgrew it by more than 2^8 -- until then P = exp2(s - m_old) <= 256 and the sums stay well inside fp32 / bf16
            // range; the final division by l uses the same m, so the result is the softmax either way.  All parts of a row see
            // the same mx and m, hence take the same decision.
            const bool first = c.j == 0;
            const bool grow = first || __any_sync(FULL_MASK, (mx - m) * FA_LOG2E > 8.0f);
            float alpha = 1.0f;
            if (grow) {
                const float m_new = fmaxf(m, mx);
                alpha = ex2f((m - m_new) * FA_LOG2E);                           // 0 for the first block (m = -inf)
                m = m_new;
            }
            if (first) {
                if (!SERIAL && have_prev) {                                     // deferred: its last PV ran behind this block's S phase
                    wait_pv(g);
                    epilogue(pv_row0, pv_len, pv_q0, pv_h, pv_l);
                }
            } else if (grow) {                                                  // rescale this thread's part of the accumulator row
                if constexpr (SERIAL) {
                    pv_waited = g;                                              // S_g complete => PV_{g-1} complete (issue order)
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        uint32_t orr[16];
                        tmem_ld16(tmem_o + lane_sel + 16 * q4, orr);
                        tmem_ld_wait();
#pragma unroll
                        for (int i = 0; i < 16; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * alpha);
                        tmem_st16(tmem_o + lane_sel + 16 * q4, orr);
                    }
                } else {
                    wait_pv(g);
                    uint32_t orr[HC];
                    tmem_ld_n<HC>(tmem_o + lane_sel + col0, orr);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < HC; ++i) orr[i] = __float_as_uint(__uint_as_float(orr[i]) * alpha);
                    tmem_st_n<HC>(tmem_o + lane_sel + col0, orr);
                }
            }
            const float ms = m * FA_LOG2E;
            float sum0 = 0.f, sum1 = 0.f;
            if constexpr (SERIAL) {
                // P_g replaces S_g in place (its columns 0 .. 31): every score of the row is in registers by now, and PV_{g-1} has
                // completed (S_g was issued behind it).  Two halves of 32 keys keep the live registers down.
                pv_waited = g;
#pragma unroll
                for (int hf = 0; hf < 2; ++hf) {
                    uint32_t pr[16];
#pragma unroll
                    for (int i = 0; i < 32; i += 4) {
                        const int k0 = 32 * hf + i;
                        const float p0 = ex2f(fmaf(__uint_as_float(sr[k0]), FA_LOG2E, -ms)), p1 = ex2f(fmaf(__uint_as_float(sr[k0 + 1]), FA_LOG2E, -ms));
                        const float p2 = ex2f(fmaf(__uint_as_float(sr[k0 + 2]), FA_LOG2E, -ms)), p3 = ex2f(fmaf(__uint_as_float(sr[k0 + 3]), FA_LOG2E, -ms));
                        sum0 += p0 + p1;
                        sum1 += p2 + p3;
                        pr[i / 2] = pack_bf16x2(p0, p1);
                        pr[i / 2 + 1] = pack_bf16x2(p2, p3);
                    }
                    tmem_st16(tmem_base + FA_TM_P + lane_sel + 16 * hf, pr);
                }
            } else {
            uint32_t pr[HC / 2];
#pragma unroll
            for (int i = 0; i < HC; i += 4) {                                   // masked keys: exp2(-inf) = 0 exactly
                const float p0 = ex2f(fmaf(__uint_as_float(sr[i]), FA_LOG2E, -ms)), p1 = ex2f(fmaf(__uint_as_float(sr[i + 1]), FA_LOG2E, -ms));
                const float p2 = ex2f(fmaf(__uint_as_float(sr[i + 2]), FA_LOG2E, -ms)), p3 = ex2f(fmaf(__uint_as_float(sr[i + 3]), FA_LOG2E, -ms));
                sum0 += p0 + p1;
                sum1 += p2 + p3;
                pr[i / 2] = pack_bf16x2(p0, p1);
                pr[i / 2 + 1] = pack_bf16x2(p2, p3);
            }
            // P buffer g & 1 was read by PV_{g-2}.  The wait names phase g - 1 (PV of the previous block, which ran behind the
            // exponentials above and has normally completed): a parity wait must never name a phase two behind the barrier --
            // with phases g - 2 and g - 1 both complete it would be taken for phase g, which needs this warp's own arrival.
            if (g >= 1) wait_pv(g);
            tmem_st_n<HC / 2>(tmem_base + FA_TM_P + (g & 1) * (FA_BN / 2) + lane_sel + col0 / 2, pr);
            }
            l = l * alpha + (sum0 + sum1);
            tmem_st_wait();                  // P (and a rescaled O) are in tensor memory
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(p_full);
            if (c.last_of_item()) {
                have_prev = true;
                pv_row0 = c.tile.x; pv_len = c.tile.z; pv_q0 = c.tile.x - c.tile.y; pv_h = c.h; pv_l = l;
            }
        }
        if (have_prev) {                                                        // the CTA's last item
            wait_pv(g);
            if constexpr (SERIAL) epilogue_halves(pv_row0, pv_len, pv_q0, pv_h, pv_l);
            else epilogue(pv_row0, pv_len, pv_q0, pv_h, pv_l);
        }
    };

    FaCursor c;
    c.start(lay.tiles, lay.n_tiles, n_head, blockIdx.x, gridDim.x);
    if (warp == N_SOFTMAX) { if (lane == 0) produce(c); }
    else if (warp == N_SOFTMAX + 1) { if (lane == 0) mma(c); }
    else softmax(c);
    tc_fence_before();
    __syncthreads();
    if (warp == N_SOFTMAX) {
        tc_fence_after();
        tmem_dealloc(tmem_base, FA_TMEM_COLS);
    }
}

template <int NSPLIT, int NCTA>
int fa_launch(const CUtensorMap& mq, const CUtensorMap& mk, const CUtensorMap& mv, void* out, long long ld, Lay lay, int n_head,
              cudaStream_t st) {
    static PerDeviceOnce once;
    auto kern = flash_attn_tc_kernel<NSPLIT, NCTA>;
    constexpr int smem = fa_smem<NCTA>();
    if (once.first() && cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return 2;
    const int n_items = lay.n_tiles * n_head;
    const int max_ctas = NCTA * device_sm_count();
    const int grid = n_items < max_ctas ? n_items : max_ctas;
    static const int safe = getenv("VRD_FA_SAFE") ? atoi(getenv("VRD_FA_SAFE")) : 0;
    launch_k(kern, dim3(grid), dim3((4 * NSPLIT + 2) * 32), smem, st, mq, mk, mv, (__nv_bfloat16*)out, ld, lay, n_head, safe);
    return 0;
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
    static const int nsplit = getenv("VRD_FA_SPLIT") ? atoi(getenv("VRD_FA_SPLIT")) : 1;      // A/B switch: softmax warps per lane quarter
    static const int ncta = getenv("VRD_FA_CTAS") ? atoi(getenv("VRD_FA_CTAS")) : 2;          // A/B switch: co-resident CTAs per SM (2 | 3)
    if (ncta == 3) return fa_launch<1, 3>(mq, mk, mv, out, ld, lay, n_head, st);
    return nsplit == 1 ? fa_launch<1, 2>(mq, mk, mv, out, ld, lay, n_head, st) : fa_launch<2, 2>(mq, mk, mv, out, ld, lay, n_head, st);
}

}  // namespace vrd
