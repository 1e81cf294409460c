// Feature packing: fp32 / bf16 rows -> the operand planes the contraction kernels read.
//
// Covers torchreid/engine/engine.py:391-394 (optional F.normalize of both sets), the row-norm terms of
// torchreid/metrics/distance.py:59-61 (pow(2).sum(1)) and the F.normalize calls of distance.py:77-78
// (x / max(||x||_2, 1e-12)), fused with the fp32 -> fp16 hi/lo split (or bf16 rounding) the tensor-core kernel needs.
#include <cuda_fp16.h>

#include "common.cuh"

namespace ieee {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

template <typename T, int VEC>
struct RowLoader;
template <>
struct RowLoader<float, 4> {
  __device__ static void load(const float* p, int64_t i, float (&v)[4]) {
    float4 t = __ldg(reinterpret_cast<const float4*>(p + i));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  }
};
template <>
struct RowLoader<float, 1> {
  __device__ static void load(const float* p, int64_t i, float (&v)[1]) { v[0] = __ldg(p + i); }
};
template <>
struct RowLoader<__nv_bfloat16, 1> {
  __device__ static void load(const __nv_bfloat16* p, int64_t i, float (&v)[1]) { v[0] = __bfloat162float(p[i]); }
};

enum PackMode { PACK_BF16 = 0, PACK_F16_HILO = 1, PACK_F32 = 2 };

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// One warp per row.  n_norm = how many times the row is L2-normalised before the metric sees it: engine
// normalize_feature (engine.py:391-394) and/or cosine's own F.normalize (distance.py:77-78).  Normalising twice
// is not the identity in fp32, so the reference's sequence of divisions is reproduced, not collapsed.
//
// PACK_F16_HILO: y = x * 2^(14-e) with e chosen per row so that max|y| is in [2^13, 2^14) (exact scaling, below
// fp16's 65504), then hi = fp16(y), lo = fp16(y - hi): 22 mantissa bits survive, and lo (~2^-12 |y|) stays a
// NORMAL fp16 number for every element within 2^-16 of the row maximum.  The contraction's epilogue multiplies
// by 2^(e_q - 14) * 2^(e_g - 14).
//
// center (may be null): a D-vector c subtracted from every (normalised) row before it is split.  The squared
// euclidean distance is translation invariant, |q - g|^2 = |(q - c) - (g - c)|^2, but its fp32 evaluation
// |q|^2 + |g|^2 - 2 q.g is not: the cancellation error scales with |q|^2 + |g|^2, and post-ReLU features (all
// non-negative, ieee3modalPart.py:417) share a large common component.  Subtracting it shrinks the row terms to the
// spread of the data and makes the products mixed-sign, so the tensor core's truncating accumulator no longer
// drifts one way.  Zero padding beyond D stays zero.
template <typename T, int VEC>
__global__ void __launch_bounds__(256) pack_rows_kernel(const T* __restrict__ x, int64_t ld, int64_t rows, int D, int Dp,
                                                         int n_norm, int mode, const float* __restrict__ center,
                                                         uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                                         float* __restrict__ f32, float* __restrict__ norms,
                                                         float* __restrict__ row_scale, const ZeroJob zero) {
  zero.run();
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const T* xr = x + row * ld;
  float den[2] = {1.0f, 1.0f};
  float amax = 0.f;
  for (int pass = 0; pass < n_norm; ++pass) {   // the reference's F.normalize calls, one row reduction each
    float s = 0.f;
#pragma unroll 6
    for (int i = lane * VEC; i < D; i += 32 * VEC) {
      float v[VEC];
      RowLoader<T, VEC>::load(xr, i, v);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float t = v[j];
        if (pass == 1) t = __fdiv_rn(t, den[0]);
        s = __fmaf_rn(t, t, s);
      }
    }
    s = warp_sum(s);
    den[pass] = fmaxf(__fsqrt_rn(s), 1e-12f);  // F.normalize: clamp_min(norm, eps)
  }
  if (mode == PACK_F16_HILO) {                  // largest magnitude of the row as the contraction will see it
#pragma unroll 6
    for (int i = lane * VEC; i < D; i += 32 * VEC) {
      float v[VEC], c[VEC];
      RowLoader<T, VEC>::load(xr, i, v);
      if (center != nullptr) RowLoader<float, VEC>::load(center, i, c);
#pragma unroll
      for (int j = 0; j < VEC; ++j) {
        float t = v[j];
        if (n_norm >= 1) t = __fdiv_rn(t, den[0]);
        if (n_norm >= 2) t = __fdiv_rn(t, den[1]);
        if (center != nullptr) t = __fsub_rn(t, c[j]);
        amax = fmaxf(amax, fabsf(t));
      }
    }
    amax = warp_max(amax);
  }
  int e = 0;
  if (mode == PACK_F16_HILO && amax > 0.f && amax < 3.0e38f) {
    (void)frexpf(amax, &e);          // amax = m * 2^e, m in [0.5, 1)
    e = max(-100, min(100, e)) - 14;   // y = x * 2^-e has its maximum in [2^13, 2^14)
  }
  const float down = ldexpf(1.0f, -e);   // exact power of two
  float sq = 0.f;
#pragma unroll 6
  for (int i = lane * VEC; i < Dp; i += 32 * VEC) {
    float v[VEC], c[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) { v[j] = 0.f; c[j] = 0.f; }
    if (i < D) {
      RowLoader<T, VEC>::load(xr, i, v);  // D % VEC == 0 on the vector path
      if (center != nullptr) RowLoader<float, VEC>::load(center, i, c);
    }
    uint16_t h[VEC], l[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      float t = v[j];
      if (n_norm >= 1) t = __fdiv_rn(t, den[0]);
      if (n_norm >= 2) t = __fdiv_rn(t, den[1]);
      t = __fsub_rn(t, c[j]);
      float u = t;   // the value whose square enters the row term: what the contraction actually multiplies
      if (mode == PACK_BF16) {
        const __nv_bfloat16 b = __float2bfloat16_rn(t);
        h[j] = __bfloat16_as_ushort(b);
        l[j] = 0;
        u = __bfloat162float(b);
      } else if (mode == PACK_F16_HILO) {
        const float y = t * down;
        const __half hh = __float2half_rn(y);
        h[j] = __half_as_ushort(hh);
        l[j] = __half_as_ushort(__float2half_rn(y - __half2float(hh)));
      }
      sq = __fmaf_rn(u, u, sq);
      v[j] = t;
    }
    const int64_t o = row * (int64_t)Dp + i;
    if (mode == PACK_F32) {
#pragma unroll
      for (int j = 0; j < VEC; ++j) f32[o + j] = v[j];
    } else {
      if constexpr (VEC == 4) {
        *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<uint2*>(h);
        if (mode == PACK_F16_HILO) *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<uint2*>(l);
      } else {
        hi[o] = h[0];
        if (mode == PACK_F16_HILO) lo[o] = l[0];
      }
    }
  }
  sq = warp_sum(sq);
  if (lane == 0) {
    norms[row] = sq;
    row_scale[row] = ldexpf(1.0f, e);
  }
}

// Same arithmetic, one pass over HBM: the (normalised, centred) row is parked in shared memory between the reduction
// passes and the split, so global memory is read once per row; loops stay rolled (a fully unrolled register-resident
// variant stalled on instruction fetch: ncu, stall no_instruction 5.1 per issue).  fp32 rows with D % 4 == 0 and
// 16-byte alignment; identical results to pack_rows_kernel (same per-lane element order in every reduction).
__global__ void __launch_bounds__(256) pack_rows_smem_kernel(const float* __restrict__ x, int64_t ld, int64_t rows, int D, int Dp,
                                                              int n_norm, int mode, const float* __restrict__ center,
                                                              uint16_t* __restrict__ hi, uint16_t* __restrict__ lo,
                                                              float* __restrict__ f32, float* __restrict__ norms,
                                                              float* __restrict__ row_scale, const ZeroJob zero) {
  zero.run();
  extern __shared__ __align__(16) float rowbuf_all[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + w;
  if (row >= rows) return;
  float* buf = rowbuf_all + (size_t)w * Dp;
  const float* xr = x + row * ld;
  // pass 1 (the only read of x): copy the row into shared memory; sum of squares for the first normalisation
  float s = 0.f;
#pragma unroll 6
  for (int i = lane * 4; i < D; i += 128) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(xr + i));
    *reinterpret_cast<float4*>(buf + i) = t;
    s = __fmaf_rn(t.x, t.x, s); s = __fmaf_rn(t.y, t.y, s); s = __fmaf_rn(t.z, t.z, s); s = __fmaf_rn(t.w, t.w, s);
  }
  float den[2] = {1.0f, 1.0f};
  if (n_norm >= 1) den[0] = fmaxf(__fsqrt_rn(warp_sum(s)), 1e-12f);
  if (n_norm >= 2) {
    s = 0.f;
#pragma unroll 6
    for (int i = lane * 4; i < D; i += 128) {
      const float4 t = *reinterpret_cast<const float4*>(buf + i);
      float u = __fdiv_rn(t.x, den[0]); s = __fmaf_rn(u, u, s);
      u = __fdiv_rn(t.y, den[0]); s = __fmaf_rn(u, u, s);
      u = __fdiv_rn(t.z, den[0]); s = __fmaf_rn(u, u, s);
      u = __fdiv_rn(t.w, den[0]); s = __fmaf_rn(u, u, s);
    }
    den[1] = fmaxf(__fsqrt_rn(warp_sum(s)), 1e-12f);
  }
  // pass 2: the value the contraction multiplies (normalised, centred), back into the buffer; its largest magnitude
  float amax = 0.f;
#pragma unroll 4
  for (int i = lane * 4; i < D; i += 128) {
    const float4 t4 = *reinterpret_cast<const float4*>(buf + i);
    float t[4] = {t4.x, t4.y, t4.z, t4.w};
    float4 c4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (center != nullptr) c4 = __ldg(reinterpret_cast<const float4*>(center + i));
    const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (n_norm >= 1) t[j] = __fdiv_rn(t[j], den[0]);
      if (n_norm >= 2) t[j] = __fdiv_rn(t[j], den[1]);
      t[j] = __fsub_rn(t[j], cc[j]);
      amax = fmaxf(amax, fabsf(t[j]));
    }
    *reinterpret_cast<float4*>(buf + i) = make_float4(t[0], t[1], t[2], t[3]);
  }
  int e = 0;
  if (mode == PACK_F16_HILO) {
    amax = warp_max(amax);
    if (amax > 0.f && amax < 3.0e38f) {
      (void)frexpf(amax, &e);
      e = max(-100, min(100, e)) - 14;
    }
  }
  const float down = ldexpf(1.0f, -e);
  // pass 3: split and store (zero padding beyond D)
  float sq = 0.f;
#pragma unroll 4
  for (int i = lane * 4; i < Dp; i += 128) {
    float4 t4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (i < D) t4 = *reinterpret_cast<const float4*>(buf + i);
    const float t[4] = {t4.x, t4.y, t4.z, t4.w};
    uint16_t h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float u = t[j];
      if (mode == PACK_BF16) {
        const __nv_bfloat16 b = __float2bfloat16_rn(t[j]);
        h[j] = __bfloat16_as_ushort(b);
        l[j] = 0;
        u = __bfloat162float(b);
      } else if (mode == PACK_F16_HILO) {
        const float y = t[j] * down;
        const __half hh = __float2half_rn(y);
        h[j] = __half_as_ushort(hh);
        l[j] = __half_as_ushort(__float2half_rn(y - __half2float(hh)));
      }
      sq = __fmaf_rn(u, u, sq);
    }
    const int64_t o = row * (int64_t)Dp + i;
    if (mode == PACK_F32) {
      *reinterpret_cast<float4*>(f32 + o) = t4;
    } else {
      *reinterpret_cast<uint2*>(hi + o) = *reinterpret_cast<uint2*>(h);
      if (mode == PACK_F16_HILO) *reinterpret_cast<uint2*>(lo + o) = *reinterpret_cast<uint2*>(l);
    }
  }
  sq = warp_sum(sq);
  if (lane == 0) {
    norms[row] = sq;
    row_scale[row] = ldexpf(1.0f, e);
  }
}

int pack_features(const void* x, int dtype, int64_t ld, int64_t rows, int64_t D, int metric, int normalize, int precision,
                  const float* center, void* packed, cudaStream_t stream, const ZeroJob* zero) {
  const ZeroJob zj = zero ? *zero : ZeroJob{{nullptr, nullptr}, {0, 0}};
  IEEE_REQUIRE(x != nullptr && packed != nullptr, "pack_features: null pointer");
  IEEE_REQUIRE(rows >= 0 && D > 0 && ld >= D, "pack_features: bad shape rows=%lld D=%lld ld=%lld", (long long)rows,
               (long long)D, (long long)ld);
  IEEE_REQUIRE(D <= (1 << 24), "pack_features: feature dim too large");
  IEEE_REQUIRE(dtype == IEEE_DTYPE_F32 || dtype == IEEE_DTYPE_BF16, "pack_features: unknown dtype %d", dtype);
  IEEE_REQUIRE(metric >= IEEE_METRIC_EUCLIDEAN && metric <= IEEE_METRIC_NEG_DOT, "unknown metric %d", metric);
  IEEE_REQUIRE(precision >= IEEE_PREC_F16X3 && precision <= IEEE_PREC_FP32_SIMT, "unknown precision %d", precision);
  IEEE_REQUIRE(center == nullptr || metric == IEEE_METRIC_EUCLIDEAN,
               "pack_features: a centre only applies to the euclidean metric (cosine is not translation invariant)");
  IEEE_REQUIRE((reinterpret_cast<uintptr_t>(center) & 15) == 0, "pack_features: centre must be 16-byte aligned");
  if (rows == 0) {
    for (int i = 0; i < 2; ++i)
      if (zj.p[i] && zj.n[i] > 0) IEEE_CUDA_CHECK(cudaMemsetAsync(zj.p[i], 0, size_t(zj.n[i]) * 4, stream));
    return IEEE_OK;
  }
  PackedLayout L = packed_layout(rows, D, precision);
  uint8_t* base = static_cast<uint8_t*>(packed);
  IEEE_REQUIRE((reinterpret_cast<uintptr_t>(base) & 255) == 0, "pack_features: packed buffer must be 256-byte aligned");
  auto* hi = reinterpret_cast<uint16_t*>(base + L.hi_off);
  auto* lo = reinterpret_cast<uint16_t*>(base + L.lo_off);
  auto* f32 = reinterpret_cast<float*>(base);
  auto* norms = reinterpret_cast<float*>(base + L.norm_off);
  auto* rscale = reinterpret_cast<float*>(base + L.scale_off);
  const int n_norm = (normalize ? 1 : 0) + (metric == IEEE_METRIC_COSINE ? 1 : 0);
  const int mode = precision == IEEE_PREC_FP32_SIMT ? PACK_F32 : (precision == IEEE_PREC_F16X3 ? PACK_F16_HILO : PACK_BF16);
  const int warps = 8;
  dim3 grid((unsigned)((rows + warps - 1) / warps)), block(warps * 32);
  if (dtype == IEEE_DTYPE_F32) {
    const float* xf = static_cast<const float*>(x);
    const bool vec = (D % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(xf) & 15) == 0);
    const size_t row_smem = size_t(warps) * L.Dp * 4;       // one fp32 row per warp
    if (vec && row_smem <= 96 * 1024 && !(g_debug_flags & 64)) {
      IEEE_ENSURE_DYN_SMEM(pack_rows_smem_kernel, row_smem);
      pack_rows_smem_kernel<<<grid, block, row_smem, stream>>>(xf, ld, rows, (int)D, (int)L.Dp, n_norm, mode, center, hi, lo, f32, norms, rscale, zj);
    } else if (vec)
      pack_rows_kernel<float, 4><<<grid, block, 0, stream>>>(xf, ld, rows, (int)D, (int)L.Dp, n_norm, mode, center, hi, lo, f32, norms, rscale, zj);
    else
      pack_rows_kernel<float, 1><<<grid, block, 0, stream>>>(xf, ld, rows, (int)D, (int)L.Dp, n_norm, mode, center, hi, lo, f32, norms, rscale, zj);
  } else {
    pack_rows_kernel<__nv_bfloat16, 1><<<grid, block, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ld, rows, (int)D,
                                                                   (int)L.Dp, n_norm, mode, center, hi, lo, f32, norms, rscale, zj);
  }
  count_launch(1, "pack_rows kernel");
  IEEE_CUDA_CHECK(cudaGetLastError());
  return IEEE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// Centre of a feature set: the column mean over up to `max_rows` evenly strided sample rows (deterministic:
// fixed assignment of rows to warps, fixed summation order).  ANY vector is a valid centre -- it only has to be
// the same for both operands of a contraction -- so a sample is enough, and for normalised features the mean of
// the raw rows is simply scaled to unit length (m / max(|m|, 1e-12)) instead of normalising every sample row.
// ---------------------------------------------------------------------------------------------------------
// One launch: every CTA owns 128 columns, warp w adds sample rows w, w + 8, ...; the eight partial rows are summed
// in a fixed order.  With `unit` a second, single-CTA launch scales the mean to unit length.
template <typename T>
__global__ void __launch_bounds__(256) center_rows_kernel(const T* __restrict__ x, int64_t ld, int64_t stride, int n_s, int D,
                                                           float* __restrict__ center) {
  __shared__ float red[8][128];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int col0 = blockIdx.x * 128 + lane * 4;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int i = w; i < n_s; i += 8) {
    const T* xr = x + (int64_t)i * stride * ld;
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (col0 + j < D) acc[j] += static_cast<float>(xr[col0 + j]);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) red[w][lane * 4 + j] = acc[j];
  __syncthreads();
  if (threadIdx.x < 128) {
    float s = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
    const int col = blockIdx.x * 128 + threadIdx.x;
    if (col < D) center[col] = s * (1.0f / (float)n_s);
  }
}

__global__ void __launch_bounds__(1024) center_unit_kernel(int D, float* __restrict__ center) {
  __shared__ float red[32];
  __shared__ float scale_s;
  float sq = 0.f;
  for (int col = threadIdx.x; col < D; col += 1024) sq = __fmaf_rn(center[col], center[col], sq);
  sq = warp_sum(sq);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sq;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = warp_sum(red[threadIdx.x]);
    if (threadIdx.x == 0) scale_s = 1.0f / fmaxf(__fsqrt_rn(t), 1e-12f);
  }
  __syncthreads();
  const float sc = scale_s;
  for (int col = threadIdx.x; col < D; col += 1024) center[col] *= sc;
}

size_t feature_center_workspace_bytes(int64_t D) { return 256; }   // (kept in the interface; the kernels need none)

int feature_center(const void* x, int dtype, int64_t ld, int64_t rows, int64_t D, int normalize, int64_t max_rows,
                   float* center, void* workspace, cudaStream_t stream) {
  IEEE_REQUIRE(x && center, "feature_center: null pointer");
  IEEE_REQUIRE(rows > 0 && D > 0 && ld >= D && D <= (1 << 24), "feature_center: bad shape rows=%lld D=%lld ld=%lld",
               (long long)rows, (long long)D, (long long)ld);
  IEEE_REQUIRE(dtype == IEEE_DTYPE_F32 || dtype == IEEE_DTYPE_BF16, "feature_center: unknown dtype %d", dtype);
  (void)workspace;
  if (max_rows <= 0) max_rows = 64;
  const int n_s = (int)(rows < max_rows ? rows : max_rows);
  const int64_t stride = rows / n_s;
  const unsigned grid = (unsigned)((D + 127) / 128);
  if (dtype == IEEE_DTYPE_F32)
    center_rows_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(x), ld, stride, n_s, (int)D, center);
  else
    center_rows_kernel<__nv_bfloat16><<<grid, 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(x), ld, stride, n_s, (int)D, center);
  count_launch(1, "center_rows_kernel");
  if (normalize) {
    center_unit_kernel<<<1, 1024, 0, stream>>>((int)D, center);
    count_launch();
  }
  IEEE_CUDA_CHECK(cudaGetLastError());
  return IEEE_OK;
}

}  // namespace ieee
