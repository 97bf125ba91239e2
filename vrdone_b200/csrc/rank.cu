// Device-side triplet ranking: the candidate filter, the mean-score ranking and the top-n_max_pair selection of
// MaskVRD.forward_test (reference models/maskvrd.py:262-328: a Python loop over every (pair, query, top-k class) with one
// device sync per candidate, then torch.argsort over all candidates) on the compact per-(pair, query) outputs of the heads
// kernels.  SURVEY.md k10 "device-side compaction": only the <= n_max_pair reported candidates cross PCIe (24 bytes each)
// and the host builds the result lists for those alone.
//
//   candidate c = (pair p, query q, class rank j),  c = (p * Q + q) * topk + j
//   kept   iff  last >= 0  and  (last - first) * feat_stride + 1 >= pred_min_frames           (maskvrd.py:289-299)
//   score  =    ((cat_scores[sid] + topk_score) + cat_scores[oid]) / 3      in fp32, in this order (torch.tensor([s, p, o]).mean())
//   order  =    descending score, ties to the earlier candidate (a stable descending argsort; the reference's argsort
//               leaves tie order unspecified -- documented tie rule)
//
// Kernel 1 (grid-wide) writes one 64-bit key per candidate: (order-preserving bits of the score) << 32 | ~c, 0 when dropped;
// keys are unique, so "the n largest keys" is exactly the reported set in the reported order.  Kernel 2 (one CTA) finds the
// n-th largest key with an MSB-first radix select (8 passes over the L2-resident key array), gathers the keys >= it into
// shared memory, sorts them (bitonic) and writes the records.
#include "common.cuh"
#include "kernels.h"

namespace vrd {

namespace {

constexpr int RANK_THREADS = 1024;
constexpr int RANK_MAX_N = 1024;     // n_max_pair supported by the shared-memory sort

__device__ __forceinline__ unsigned int order_bits(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);     // larger float <=> larger unsigned
}

__global__ void __launch_bounds__(256) rank_keys_kernel(const float* __restrict__ topk_scores, const int* __restrict__ first_last,
                                                        const long long* __restrict__ sids, const long long* __restrict__ oids,
                                                        const float* __restrict__ cat_scores, const long long* __restrict__ durs,
                                                        const long long* __restrict__ so_offset, int B, int Q, int topk,
                                                        int feat_stride, int pred_min_frames, unsigned long long* __restrict__ keys,
                                                        int* __restrict__ header) {
    const long long M = (long long)B * Q * topk;
    const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M) return;
    const int pq = (int)(c / topk);
    const int p = pq / Q;
    const int first = first_last[2 * pq], last = first_last[2 * pq + 1];
    const bool keep = last >= 0 && (long long)(last - first) * feat_stride + 1 >= pred_min_frames;
    unsigned long long key = 0ull;
    if (keep) {
        const long long s = sids[p], o = oids[p];
        const float avg = ((cat_scores[s] + topk_scores[c]) + cat_scores[o]) / 3.0f;
        key = ((unsigned long long)order_bits(avg) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned int)c);
        if (c % topk == 0) {     // the reference asserts 0 <= start and end <= overlap length for every kept candidate (maskvrd.py:297)
            const long long lim = min(durs[2 * s + 1], durs[2 * o + 1]) - max(durs[2 * s], durs[2 * o]);
            const long long off = so_offset[p];
            const long long start = (long long)first * feat_stride + off, end = (long long)last * feat_stride + off + 1;
            if (start < 0 || end > lim) atomicOr(&header[1], 1);
        }
    }
    keys[c] = key;
}

__global__ void __launch_bounds__(RANK_THREADS) rank_select_kernel(const unsigned long long* __restrict__ keys, long long M,
                                                                   int n_max, const float* __restrict__ topk_scores,
                                                                   const int* __restrict__ topk_ids,
                                                                   const int* __restrict__ first_last, int topk,
                                                                   int* __restrict__ header, int* __restrict__ records) {
    __shared__ unsigned int hist[256];
    __shared__ unsigned long long s_prefix;
    __shared__ int s_remaining, s_count, s_all;
    __shared__ unsigned long long sel[RANK_MAX_N];
    const int tid = threadIdx.x;
    if (tid == 0) { s_prefix = 0ull; s_remaining = n_max; s_count = 0; s_all = 0; }
    __syncthreads();
    // MSB-first radix select of the n_max-th largest key among the non-zero keys
    for (int pass = 7; pass >= 0 && !s_all; --pass) {
        if (tid < 256) hist[tid] = 0u;
        __syncthreads();
        const unsigned long long himask = pass == 7 ? 0ull : (~0ull << (8 * (pass + 1)));
        const unsigned long long prefix = s_prefix;
        for (long long i = tid; i < M; i += RANK_THREADS) {
            const unsigned long long k = keys[i];
            if (k != 0ull && (k & himask) == prefix) atomicAdd(&hist[(unsigned int)(k >> (8 * pass)) & 255u], 1u);
        }
        __syncthreads();
        if (tid == 0) {
            int cum = 0, b = 255;
            for (; b >= 0; --b) {
                if (cum + (int)hist[b] >= s_remaining) break;
                cum += (int)hist[b];
            }
            if (b < 0) s_all = 1;            // fewer than n_max kept candidates in total: every non-zero key is reported
            else { s_remaining -= cum; s_prefix = prefix | ((unsigned long long)b << (8 * pass)); }
        }
        __syncthreads();
    }
    const unsigned long long thr = s_all ? 1ull : s_prefix;
    for (long long i = tid; i < M; i += RANK_THREADS) {
        const unsigned long long k = keys[i];
        if (k >= thr && k != 0ull) {
            const int slot = atomicAdd(&s_count, 1);
            if (slot < RANK_MAX_N) sel[slot] = k;
        }
    }
    __syncthreads();
    const int count = min(s_count, n_max);
    int n2 = 1;
    while (n2 < count) n2 <<= 1;
    for (int i = count + tid; i < n2; i += RANK_THREADS) sel[i] = 0ull;
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {                 // bitonic sort, descending
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = tid; i < n2; i += RANK_THREADS) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = sel[i], b = sel[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) { sel[i] = b; sel[ixj] = a; }
                }
            }
            __syncthreads();
        }
    }
    if (tid == 0) header[0] = count;
    for (int i = tid; i < count; i += RANK_THREADS) {
        const unsigned long long k = sel[i];
        const unsigned int c = 0xFFFFFFFFu - (unsigned int)(k & 0xFFFFFFFFull);
        const unsigned int ob = (unsigned int)(k >> 32);
        const unsigned int fb = (ob & 0x80000000u) ? (ob & 0x7FFFFFFFu) : ~ob;     // inverse of order_bits
        const int pq = (int)(c / (unsigned int)topk);
        int* r = records + 6 * i;
        r[0] = (int)c;
        r[1] = (int)fb;                                   // mean score (fp32 bits)
        r[2] = __float_as_int(topk_scores[c]);            // predicate score (fp32 bits)
        r[3] = topk_ids[c];                               // 1-based predicate id
        r[4] = first_last[2 * pq];
        r[5] = first_last[2 * pq + 1];
    }
}

}  // namespace

int rank_triplets(const float* topk_scores, const int* topk_ids, const int* first_last, const long long* sids, const long long* oids,
                  const float* cat_scores, const long long* durs, const long long* so_offset, int B, int Q, int topk, int feat_stride,
                  int pred_min_frames, int n_max, unsigned long long* keys, int* header, int* records, cudaStream_t st) {
    if (n_max < 1 || n_max > RANK_MAX_N || B < 1 || Q < 1 || topk < 1) return 1;
    const long long M = (long long)B * Q * topk;
    if (M >= 0xFFFFFFFFll) return 1;
    cudaMemsetAsync(header, 0, 4 * sizeof(int), st);
    rank_keys_kernel<<<(unsigned int)((M + 255) / 256), 256, 0, st>>>(topk_scores, first_last, sids, oids, cat_scores, durs, so_offset,
                                                                     B, Q, topk, feat_stride, pred_min_frames, keys, header);
    rank_select_kernel<<<1, RANK_THREADS, 0, st>>>(keys, M, n_max, topk_scores, topk_ids, first_last, topk, header, records);
    return 0;
}

}  // namespace vrd
