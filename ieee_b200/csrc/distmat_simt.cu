// fp32 FMA contraction (no tensor cores): IEEE_PREC_FP32_SIMT.  Same epilogue as the tcgen05 kernel; exists to
// cross-check the tensor path on the device and for callers that want plain fp32 products
// (torchreid/metrics/distance.py:59-64, :77-80).
#include "common.cuh"

namespace ieee {

constexpr int ST = 64;   // tile edge
constexpr int SK = 16;   // k slab

__global__ void __launch_bounds__(256) distmat_simt_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                            const float* __restrict__ rq, const float* __restrict__ rg,
                                                            float alpha, float base0, int Q, int G, int Dp,
                                                            float* __restrict__ out, int64_t ldo) {
  __shared__ float sa[SK][ST + 1];
  __shared__ float sb[SK][ST + 1];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * ST, n0 = blockIdx.x * ST;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < Dp; k0 += SK) {
    for (int e = threadIdx.x; e < ST * SK; e += 256) {
      const int r = e / SK, c = e % SK;
      sa[c][r] = (m0 + r < Q) ? A[(int64_t)(m0 + r) * Dp + k0 + c] : 0.f;
      sb[c][r] = (n0 + r < G) ? B[(int64_t)(n0 + r) * Dp + k0 + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) { a[i] = sa[k][ty * 4 + i]; b[i] = sb[k][tx * 4 + i]; }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = __fmaf_rn(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int row = m0 + ty * 4 + i;
    if (row >= Q) continue;
    const float r1 = rq ? rq[row] : base0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int col = n0 + tx * 4 + j;
      if (col >= G) continue;
      const float r2 = rg ? rg[col] : 0.0f;
      out[(int64_t)row * ldo + col] = __fmaf_rn(alpha, acc[i][j], __fadd_rn(r1, r2));
    }
  }
}

int distmat_simt(const void* q_packed, int64_t Q, const void* g_packed, int64_t G, int64_t D, int metric, float* out,
                 int64_t ldo, cudaStream_t stream) {
  PackedLayout lq = packed_layout(Q, D, IEEE_PREC_FP32_SIMT), lg = packed_layout(G, D, IEEE_PREC_FP32_SIMT);
  const uint8_t* qb = static_cast<const uint8_t*>(q_packed);
  const uint8_t* gb = static_cast<const uint8_t*>(g_packed);
  const bool euclid = metric == IEEE_METRIC_EUCLIDEAN;
  dim3 grid((unsigned)((G + ST - 1) / ST), (unsigned)((Q + ST - 1) / ST));
  distmat_simt_kernel<<<grid, 256, 0, stream>>>(
      reinterpret_cast<const float*>(qb), reinterpret_cast<const float*>(gb),
      euclid ? reinterpret_cast<const float*>(qb + lq.norm_off) : nullptr,
      euclid ? reinterpret_cast<const float*>(gb + lg.norm_off) : nullptr, euclid ? -2.0f : -1.0f,
      metric == IEEE_METRIC_NEG_DOT ? 0.0f : 1.0f, (int)Q, (int)G,
      (int)lq.Dp, out, ldo); count_launch();
  IEEE_CUDA_CHECK(cudaGetLastError());
  return IEEE_OK;
}

}  // namespace ieee
