// Duplicate-tracklet vIoU filter on the device.
// Reference semantics: dataloaders/vidor.py:583-641 (and vidvrd.py, same code): for every tracklet pair base < ref of the same
// category with overlapping durations, the intersection / base / ref box volumes over the common frames decide whether ref
// (rule 1) or base (rule 2) is a duplicate; a greedy scan in tracklet order applies the decisions.  The reference does this in
// a Python double loop with O(frames) tensor ops per pair on the CPU; here one CTA reduces one (base, ref) pair over its
// frames (HBM/L2-bound: 32 B per frame pair) and one CTA replays the greedy scan on the N x N decision bytes in shared memory.
#include "common.cuh"
#include "kernels.h"

namespace vrd {

constexpr int VIOU_THREADS = 128;

// flags[b * N + r]: 0 = nothing, 1 = rule 1 (ref is dropped), 2 = rule 2 (base is dropped, its scan ends)
__global__ void __launch_bounds__(VIOU_THREADS) viou_pairs_kernel(const float4* __restrict__ boxes, const int* __restrict__ trk_base,
                                                                  const int* __restrict__ durs, const int* __restrict__ cat_ids,
                                                                  int N, float thr, double* __restrict__ sums,
                                                                  unsigned char* __restrict__ flags) {
    const int b = blockIdx.y, r = blockIdx.x;
    const long long cell = (long long)b * N + r;
    int b0 = 0, b1 = 0, r0 = 0, r1 = 0;
    bool active = r > b && cat_ids[b] == cat_ids[r];
    if (active) {
        b0 = durs[2 * b]; b1 = durs[2 * b + 1]; r0 = durs[2 * r]; r1 = durs[2 * r + 1];
        active = !(r0 >= b1 || r1 <= b0);
    }
    if (!active) {                                             // block-uniform
        if (threadIdx.x == 0) {
            flags[cell] = 0;
            if (sums != nullptr) { sums[3 * cell] = 0.0; sums[3 * cell + 1] = 0.0; sums[3 * cell + 2] = 0.0; }
        }
        return;
    }
    const int s = max(b0, r0), e = min(b1, r1);
    const float4* pb = boxes + trk_base[b] + (s - b0);
    const float4* pr = boxes + trk_base[r] + (s - r0);
    // per-frame terms in fp32 exactly as the reference computes them (TO_REMOVE = 1); the sums over frames in fp64
    double inter = 0.0, vol_b = 0.0, vol_r = 0.0;
    for (int t = threadIdx.x; t < e - s; t += VIOU_THREADS) {
        const float4 x = __ldg(pb + t), y = __ldg(pr + t);
        vol_b += (double)((x.z - x.x + 1.0f) * (x.w - x.y + 1.0f));
        vol_r += (double)((y.z - y.x + 1.0f) * (y.w - y.y + 1.0f));
        const float w = fmaxf(fminf(x.z, y.z) - fmaxf(x.x, y.x) + 1.0f, 0.0f);
        const float h = fmaxf(fminf(x.w, y.w) - fmaxf(x.y, y.y) + 1.0f, 0.0f);
        inter += (double)(w * h);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        inter += __shfl_xor_sync(FULL_MASK, inter, o);
        vol_b += __shfl_xor_sync(FULL_MASK, vol_b, o);
        vol_r += __shfl_xor_sync(FULL_MASK, vol_r, o);
    }
    __shared__ double part[VIOU_THREADS / 32][3];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { part[warp][0] = inter; part[warp][1] = vol_b; part[warp][2] = vol_r; }
    __syncthreads();
    if (threadIdx.x == 0) {
        inter = vol_b = vol_r = 0.0;
#pragma unroll
        for (int i = 0; i < VIOU_THREADS / 32; ++i) { inter += part[i][0]; vol_b += part[i][1]; vol_r += part[i][2]; }
        // the reference divides fp32 sums (vidor.py:633-634)
        const float viou_br = (float)inter / (float)vol_r, viou_rb = (float)inter / (float)vol_b;
        unsigned char f = 0;
        if (viou_br > thr && b0 <= r0 && b1 >= r1) f = 1;
        else if (viou_rb > thr && r0 <= b0 && r1 >= b1) f = 2;
        flags[cell] = f;
        if (sums != nullptr) { sums[3 * cell] = inter; sums[3 * cell + 1] = vol_b; sums[3 * cell + 2] = vol_r; }
    }
}

// The greedy scan (vidor.py:589-641): refs already dropped are skipped, a dropped base ends its own scan but a base dropped
// earlier still scans.  Sequential by nature; N is tens of tracklets, so one thread walks the decision bytes in shared memory.
__global__ void viou_resolve_kernel(const unsigned char* __restrict__ flags, int N, int* __restrict__ valid, int in_smem) {
    extern __shared__ unsigned char s_flags[];
    __shared__ int s_valid[1024];
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_valid[i] = 1;
    if (in_smem)
        for (int i = threadIdx.x; i < N * N; i += blockDim.x) s_flags[i] = flags[i];
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned char* f = in_smem ? s_flags : flags;
        for (int b = 0; b < N; ++b)
            for (int r = b + 1; r < N; ++r) {
                if (!s_valid[r]) continue;
                const unsigned char c = f[b * N + r];
                if (c == 1) s_valid[r] = 0;
                else if (c == 2) { s_valid[b] = 0; break; }
            }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) valid[i] = s_valid[i];
}

int viou_filter(const float* boxes, const int* trk_base, const int* durs, const int* cat_ids, int N, float thr, double* sums,
                unsigned char* flags, int* valid, cudaStream_t st) {
    if (N < 1 || N > 1024) return 1;
    viou_pairs_kernel<<<dim3(N, N), VIOU_THREADS, 0, st>>>(reinterpret_cast<const float4*>(boxes), trk_base, durs, cat_ids, N, thr,
                                                          sums, flags);
    const int in_smem = N * N <= 40 * 1024;
    viou_resolve_kernel<<<1, 256, in_smem ? N * N : 0, st>>>(flags, N, valid, in_smem);
    return 0;
}

}  // namespace vrd
