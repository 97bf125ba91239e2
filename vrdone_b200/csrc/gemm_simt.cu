// fp32 GEMM on CUDA cores with the fused epilogue -- the arithmetic of the "fp32 path" (parity anchor: the
// reference computes every conv in fp32, models/blocks.py:85-113, 46-61, 728-737).  D[M,N] = A[M,taps*K] * W[N,taps*K]^T
// where, for taps == 3, K-slab d of row r is A[r + d - 1, :] (k=3 convolution over the token-major layout; the zero
// separator rows of layout.py provide the zero padding).
#include <cstdlib>
#include <cstring>
#include "common.cuh"
#include "kernels.h"

namespace vrd {

constexpr int BM = 128, BN = 64, BK = 16;
constexpr int TM = 8, TN = 4;          // per-thread micro tile; 16 x 16 threads

template <typename TO>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs g) {
    __shared__ float As[BK][BM + 4];
    __shared__ float Ws[BK][BN + 4];
    const float* __restrict__ A = (const float*)g.A;
    const float* __restrict__ W = (const float*)g.W;
    const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
    const int tid = threadIdx.x;
    const int tx = tid % 16, ty = tid / 16;
    const int Ktot = g.taps * g.K;
    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    for (int k0 = 0; k0 < Ktot; k0 += BK) {
        const int tap = k0 / g.K;
        const int kk0 = k0 - tap * g.K;
        const int shift = (g.taps == 3) ? tap - 1 : 0;
        // A tile: 128 rows x 16 cols = 512 float4 -> 2 per thread
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int idx = tid + i * 256;
            const int r = idx / 4, c4 = (idx % 4) * 4;
            const int row = m0 + r + shift;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row >= 0 && row < g.M) v = *reinterpret_cast<const float4*>(A + (long long)row * g.lda + kk0 + c4);
            As[c4 + 0][r] = v.x; As[c4 + 1][r] = v.y; As[c4 + 2][r] = v.z; As[c4 + 3][r] = v.w;
        }
        // W tile: 64 rows x 16 cols = 256 float4 -> 1 per thread
        {
            const int r = tid / 4, c4 = (tid % 4) * 4;
            const int n = n0 + r;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (n < g.N) v = *reinterpret_cast<const float4*>(W + (long long)n * Ktot + k0 + c4);
            Ws[c4 + 0][r] = v.x; Ws[c4 + 1][r] = v.y; Ws[c4 + 2][r] = v.z; Ws[c4 + 3][r] = v.w;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int i = 0; i < TM; ++i) a[i] = As[kk][ty * TM + i];
#pragma unroll
            for (int j = 0; j < TN; ++j) b[j] = Ws[kk][tx * TN + j];
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    // epilogue
    TO* out = (TO*)g.out;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int row = m0 + ty * TM + i;
        if (row >= g.M) continue;
        bool valid = true, add_corr = false;
        if (g.row_seq != nullptr) {
            const int rl = row % g.R;
            const int seq = g.row_seq[rl];
            valid = seq >= 0;
            if (valid && g.corr != nullptr) {
                const int4 si = g.seqinfo[seq];
                add_corr = (si.z != 0) && (rl - si.x == si.y - 1);
            }
        }
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= g.N) continue;
            float v = 0.f;
            if (valid) {
                v = acc[i][j];
                if (g.bias != nullptr) v += g.bias[n];
                if (add_corr) v += g.corr[n];
                if (g.act == 1) v = fmaxf(v, 0.f);
                else if (g.act == 2) v = gelu_erf(v);
                if (g.res1 != nullptr) v += g.res1[(long long)row * g.ldr1 + n];
                if (g.res2 != nullptr) v += g.res2[(long long)row * g.ldr2 + n];
            }
            out[(long long)row * g.ldo + n] = from_f<TO>(v);
        }
    }
}

int gemm_simt_f32(const GemmArgs& g, cudaStream_t st) {
    if (g.K % BK != 0 || g.lda % 4 != 0) return 1;
    dim3 grid((g.M + BM - 1) / BM, (g.N + BN - 1) / BN);
    if (g.out_dtype == VRD_BF16) gemm_simt_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(g);
    else gemm_simt_kernel<float><<<grid, 256, 0, st>>>(g);
    return 0;
}

int gemm_f32(const GemmArgs& g, cudaStream_t st) {
    static const bool simt_only = getenv("VRD_FP32_GEMM") != nullptr && strcmp(getenv("VRD_FP32_GEMM"), "simt") == 0;
    if (!simt_only) {
        const int rc = gemm_tcgen05_f32split(g, st);
        if (rc != 1) return rc;
    }
    return gemm_simt_f32(g, st);
}

}  // namespace vrd
