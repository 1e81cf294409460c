// Ranking fused into the contraction: the kernels around distmat_umma_chunked_kernel<CG, true>.
//
// The staged path writes the Q x G distance block (4 bytes per pair), reads it back to gather the relevant pairs'
// distances, and reads it once more to count (rank.cu).  Positions only depend on how every output compares with the
// few distances of its row's same-identity gallery items (rank.py:117-160), so the contraction's epilogue can count
// while the tile is still in registers -- if it knows those distances BEFORE it runs.  It gets approximations:
//
//   fused_link_kernel      queries -> per-identity lists (the gallery's identity hash, rank.cu)
//   fused_prepass_kernel   d~(q, g) for every query and every gallery item of its identity, plain fp32 from the packed
//                          operands (~3 % of the pairs): thresholds thr[q][i] and a band half-width eps[q] that bounds
//                          |d~ - d| for both arithmetics
//   contraction epilogue   per thread = (row, 128 columns): if no output lies within eps of a threshold, every
//                          comparison against the approximation equals the comparison against the exact value -> add
//                          the counts; otherwise the 128 outputs are SPILLED (a few per cent of the block)
//   fused_extract_kernel   spilled spans: near-duplicate fix-up (as distmat_fixup_kernel), then the exact distance of
//                          every same-identity item is picked out of the span that holds it
//   fused_recount_kernel   spilled spans against the exact (distance, index) thresholds: exact counts, ties
//   fused_finalize_kernel  per query: junk correction, positions, AP / first hit / mINP term -- and the proof
//                          obligations: every threshold was found and lies inside its band, nothing overflowed.
//
// Whatever cannot be certified (a band violated, a list too long, a non-finite threshold, spill space exhausted) raises
// `fallback`, and the caller runs the staged path: the result is either bit-identical to the staged path or not
// produced at all.
#include <cuda_fp16.h>

#include "common.cuh"

namespace ieee {

constexpr float kFusedKappa = 3.814697265625e-6f;   // 2^-18 of |q|^2 + |g|^2: > 2x (contraction floor + pre-pass error)
constexpr float kFusedTau = 0.015625f;              // near-duplicate bound of the store epilogue (kFixTau)

struct FusedBuffers {
  int32_t* qhead;        // [T]      first query of the identity in gallery slot s (-1: none)
  int32_t* qnext;        // [Q]      next query of the same identity
  int32_t* qslot;        // [Q]      gallery slot of the query's identity (-1: not in the gallery)
  float* thr;            // [Q][LC]
  int32_t* tcol;         // [Q][LC]  gallery index of item i
  uint32_t* tjunk;       // [Q]      bit i: item i is junk (same camera, rank.py:136)
  int32_t* tn;           // [Q]
  float* eps;            // [Q]
  int32_t* cnt;          // [Q][LC]
  float* exact;          // [Q][LC]  exact distance of item i (contraction arithmetic, after the near-duplicate fix-up)
  uint32_t* tfound;      // [Q]      bit i: exact[q][i] is set
  unsigned int* spill_n;
  unsigned long long* spill_meta;
  float* spill_val;
  uint32_t spill_cap;
  unsigned long long* stats;   // [0] fallback reasons (bit mask), [1] tie pairs
  int32_t* active;       // [<= min(Q, T)] gallery slots that some query asks for
  unsigned int* n_active;
  int LC;
};

enum FusedFallback { FB_LIST_TOO_LONG = 1, FB_SPILL_FULL = 2, FB_NOT_FOUND = 4, FB_BAND_VIOLATED = 8, FB_NON_FINITE = 16 };

__global__ void fused_link_kernel(const int64_t* __restrict__ q_pids, int64_t Q, const long long* __restrict__ keys, int64_t T,
                                  FusedBuffers fb) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  const int s = group_find(keys, T, q_pids[q]);
  fb.qslot[q] = s;
  if (s >= 0) {
    const int32_t prev = atomicExch(fb.qhead + s, (int32_t)q);
    fb.qnext[q] = prev;
    if (prev < 0) fb.active[atomicAdd(fb.n_active, 1u)] = s;      // first query of this identity: the slot becomes active
  } else {   // identity not in the gallery: no thresholds (rank.py:142-144 skips the query)
    for (int m = 0; m < fb.LC; ++m) { fb.thr[(size_t)q * fb.LC + m] = INFINITY; fb.tcol[(size_t)q * fb.LC + m] = -1; }
    fb.tn[q] = 0;
    fb.tjunk[q] = 0;
    fb.eps[q] = 0.f;
  }
}

// One CTA per active gallery slot (identity), round robin.  The identity's gallery rows are staged in shared memory
// as fp32 (up to `chunk_rows` at a time: 22 rows of 2304 floats fill 200 KB), then each warp takes queries of that
// identity: its packed row streams through registers once and is multiplied with every staged row (conflict-free
// 16-byte shared loads).  Unsorted approximate distances go to fb.exact (scratch until the extract kernel owns it);
// a second sweep sorts each query's list by (value, item) and writes the tables the epilogue reads.
constexpr int kPreRows = 24;     // accumulators per lane = most rows a chunk may hold

__global__ void __launch_bounds__(256) fused_prepass_kernel(
    const __half* __restrict__ q_hi, const __half* __restrict__ q_lo, const __half* __restrict__ g_hi, const __half* __restrict__ g_lo,
    const float* __restrict__ sq, const float* __restrict__ sg, const float* __restrict__ rq, const float* __restrict__ rg, int Dp,
    float alpha, const int64_t* __restrict__ q_camids, const int64_t* __restrict__ g_camids, GroupTables gt, FusedBuffers fb,
    int chunk_rows) {
  extern __shared__ __align__(16) float gs[];                 // [chunk_rows][Dp]
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned int n_active = *fb.n_active;
  for (unsigned int a = blockIdx.x; a < n_active; a += gridDim.x) {
    const int slot = fb.active[a];
    const int q0 = fb.qhead[slot];
    const int n = gt.cnt[slot], off = gt.off[slot];
    if (n > fb.LC) {                                           // more same-identity items than the epilogue tables hold
      if (w == 0) {
        for (int q = q0; q >= 0; q = fb.qnext[q]) {
          for (int m = lane; m < fb.LC; m += 32) { fb.thr[(size_t)q * fb.LC + m] = INFINITY; fb.tcol[(size_t)q * fb.LC + m] = -1; }
          if (lane == 0) { fb.tn[q] = 0; fb.tjunk[q] = 0; fb.eps[q] = 0.f; }
        }
        if (lane == 0) atomicOr(fb.stats, (unsigned long long)FB_LIST_TOO_LONG);
      }
      continue;
    }
    for (int c0 = 0; c0 < n; c0 += chunk_rows) {
      const int R = min(chunk_rows, n - c0);
      __syncthreads();                                         // the previous chunk / slot is no longer being read
      const int per_row = Dp / 8, total = R * per_row;
      for (int idx0 = threadIdx.x; idx0 < total; idx0 += 4 * 256) {      // eight 16-byte loads in flight per thread
        uint4 y0[4], y1[4];
        int rr[4], ii[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int idx = idx0 + u * 256;
          rr[u] = -1;
          if (idx < total) {
            rr[u] = idx / per_row;
            ii[u] = idx - rr[u] * per_row;
            const int g = gt.members[off + c0 + rr[u]];
            y0[u] = reinterpret_cast<const uint4*>(g_hi + (size_t)g * Dp)[ii[u]];
            y1[u] = reinterpret_cast<const uint4*>(g_lo + (size_t)g * Dp)[ii[u]];
          }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          if (rr[u] < 0) continue;
          const __half2* yh = reinterpret_cast<const __half2*>(&y0[u]);
          const __half2* yl = reinterpret_cast<const __half2*>(&y1[u]);
          float v[8];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 h = __half22float2(yh[j]), l = __half22float2(yl[j]);
            v[2 * j] = __fadd_rn(h.x, l.x);
            v[2 * j + 1] = __fadd_rn(h.y, l.y);
          }
          float4* dst = reinterpret_cast<float4*>(gs + (size_t)rr[u] * Dp + 8 * ii[u]);
          dst[0] = make_float4(v[0], v[1], v[2], v[3]);
          dst[1] = make_float4(v[4], v[5], v[6], v[7]);
        }
      }
      __syncthreads();
      int q = q0;
      for (int i = 0; i < w && q >= 0; ++i) q = fb.qnext[q];   // warp w starts at the w-th query of the list
      while (q >= 0) {
        float acc[kPreRows];
#pragma unroll
        for (int m = 0; m < kPreRows; ++m) acc[m] = 0.f;
        const uint2* ah = reinterpret_cast<const uint2*>(q_hi + (size_t)q * Dp);
        const uint2* al = reinterpret_cast<const uint2*>(q_lo + (size_t)q * Dp);
        constexpr int kAhead = 6;                              // query loads in flight (2 x 8 bytes each)
        uint2 nx0[kAhead], nx1[kAhead];
        const int iters = Dp / 128;
#pragma unroll
        for (int u = 0; u < kAhead; ++u)
          if (u < iters) { nx0[u] = ah[u * 32 + lane]; nx1[u] = al[u * 32 + lane]; }
        for (int it = 0; it < iters; ++it) {                   // lane owns elements it * 128 + 4 * lane .. + 3
          const uint2 x0 = nx0[0], x1 = nx1[0];
#pragma unroll
          for (int u = 0; u + 1 < kAhead; ++u) { nx0[u] = nx0[u + 1]; nx1[u] = nx1[u + 1]; }
          if (it + kAhead < iters) { nx0[kAhead - 1] = ah[(it + kAhead) * 32 + lane]; nx1[kAhead - 1] = al[(it + kAhead) * 32 + lane]; }
          const __half2* xh = reinterpret_cast<const __half2*>(&x0);
          const __half2* xl = reinterpret_cast<const __half2*>(&x1);
          const float2 h0 = __half22float2(xh[0]), l0 = __half22float2(xl[0]), h1 = __half22float2(xh[1]), l1 = __half22float2(xl[1]);
          const float a0 = __fadd_rn(h0.x, l0.x), a1 = __fadd_rn(h0.y, l0.y), a2 = __fadd_rn(h1.x, l1.x), a3 = __fadd_rn(h1.y, l1.y);
          const float* col = gs + it * 128 + 4 * lane;
#pragma unroll
          for (int m = 0; m < kPreRows; ++m) {
            if (m < R) {
              const float4 b = *reinterpret_cast<const float4*>(col + (size_t)m * Dp);
              acc[m] = __fmaf_rn(a0, b.x, acc[m]);
              acc[m] = __fmaf_rn(a1, b.y, acc[m]);
              acc[m] = __fmaf_rn(a2, b.z, acc[m]);
              acc[m] = __fmaf_rn(a3, b.w, acc[m]);
            }
          }
        }
        for (int it4 = (Dp / 128) * 128 + 4 * lane; it4 < Dp; it4 += 128) {      // Dp is a multiple of 64: one half step left
          const uint2 x0 = *reinterpret_cast<const uint2*>(q_hi + (size_t)q * Dp + it4), x1 = *reinterpret_cast<const uint2*>(q_lo + (size_t)q * Dp + it4);
          const __half2* xh = reinterpret_cast<const __half2*>(&x0);
          const __half2* xl = reinterpret_cast<const __half2*>(&x1);
          const float2 h0 = __half22float2(xh[0]), l0 = __half22float2(xl[0]), h1 = __half22float2(xh[1]), l1 = __half22float2(xl[1]);
          const float a0 = __fadd_rn(h0.x, l0.x), a1 = __fadd_rn(h0.y, l0.y), a2 = __fadd_rn(h1.x, l1.x), a3 = __fadd_rn(h1.y, l1.y);
#pragma unroll
          for (int m = 0; m < kPreRows; ++m) {
            if (m < R) {
              const float4 b = *reinterpret_cast<const float4*>(gs + (size_t)m * Dp + it4);
              acc[m] = __fmaf_rn(a0, b.x, acc[m]);
              acc[m] = __fmaf_rn(a1, b.y, acc[m]);
              acc[m] = __fmaf_rn(a2, b.z, acc[m]);
              acc[m] = __fmaf_rn(a3, b.w, acc[m]);
            }
          }
        }
        const float a_s = sq ? sq[q] : 1.0f, a_n = rq ? rq[q] : 1.0f;
#pragma unroll
        for (int m = 0; m < kPreRows; ++m) {
          if (m < R) {
            float v = acc[m];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) {
              const int g = gt.members[off + c0 + m];
              const float b_n = rg ? rg[g] : 0.0f, b_s = sg ? sg[g] : 1.0f;
              // the contraction's epilogue formula: d = fma(alpha * s_q * s_g, dot, |q|^2 + |g|^2)
              fb.exact[(size_t)q * fb.LC + c0 + m] = __fmaf_rn(alpha * a_s * b_s, v, __fadd_rn(a_n, b_n));
            }
          }
        }
        for (int i = 0; i < 8 && q >= 0; ++i) q = fb.qnext[q];  // the warps take every 8th query of the list
      }
    }
    __syncthreads();                                           // all approximate distances of this identity are written
    // sort every query's list by (value, item) and emit the tables (lane m = item m)
    int q = q0;
    for (int i = 0; i < w && q >= 0; ++i) q = fb.qnext[q];
    while (q >= 0) {
      float v = INFINITY, b_n = 0.f;
      int g = -1;
      bool is_junk = false;
      if (lane < n) {
        g = gt.members[off + lane];
        v = fb.exact[(size_t)q * fb.LC + lane];
        b_n = rg ? rg[g] : 0.0f;
        is_junk = g_camids[g] == q_camids[q];
      }
      int rank = 0;
      for (int j = 0; j < n; ++j) {
        const float vj = __shfl_sync(0xffffffffu, v, j);
        rank += (vj < v) || (vj == v && j < lane);
      }
      if (!(v == v)) rank = lane;                              // NaN: keep every entry distinct; flagged by the finalize kernel
      float rg_max = b_n;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) rg_max = fmaxf(rg_max, __shfl_xor_sync(0xffffffffu, rg_max, o));
      const unsigned nan_any = __ballot_sync(0xffffffffu, lane < n && !(v == v));
      if (nan_any) { if (lane == 0) atomicOr(fb.stats, (unsigned long long)FB_NON_FINITE); rank = lane; }
      if (lane < n) { fb.thr[(size_t)q * fb.LC + rank] = v; fb.tcol[(size_t)q * fb.LC + rank] = g; }
      for (int m = n + lane; m < fb.LC; m += 32) { fb.thr[(size_t)q * fb.LC + m] = INFINITY; fb.tcol[(size_t)q * fb.LC + m] = -1; }
      // junk bits in sorted order: lane with rank r contributes bit r
      unsigned jb = (lane < n && is_junk) ? (1u << rank) : 0u;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) jb |= __shfl_xor_sync(0xffffffffu, jb, o);
      if (lane == 0) {
        fb.tn[q] = n;
        fb.tjunk[q] = jb;
        fb.eps[q] = kFusedKappa * __fadd_rn(rq ? rq[q] : 1.0f, rg_max);
      }
      for (int i = 0; i < 8 && q >= 0; ++i) q = fb.qnext[q];
    }
  }
}

// One warp per spilled span (lane l holds outputs 4l .. 4l+3).
__global__ void __launch_bounds__(256) fused_extract_kernel(
    const __half* __restrict__ q_hi, const __half* __restrict__ q_lo, const __half* __restrict__ g_hi, const __half* __restrict__ g_lo,
    const float* __restrict__ sq, const float* __restrict__ sg, const float* __restrict__ rq, const float* __restrict__ rg, int Dp, int G,
    float tau, FusedBuffers fb) {
  const unsigned int n_spill = min(*fb.spill_n, fb.spill_cap);
  const int lane = threadIdx.x & 31;
  const unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (unsigned int e = warp; e < n_spill; e += nwarps) {
    const unsigned long long rc = fb.spill_meta[e];
    const uint32_t row = (uint32_t)(rc >> 32), col0 = (uint32_t)rc;
    float* val = fb.spill_val + (size_t)e * 128;
    if (rq != nullptr && tau > 0.f) {
      // near-duplicate pairs of the span, recomputed in difference form (bit for bit what distmat_fixup_kernel writes)
      const float a_s = sq[row], a_n = rq[row];
      const uint4* ah = reinterpret_cast<const uint4*>(q_hi + (size_t)row * Dp);
      const uint4* al = reinterpret_cast<const uint4*>(q_lo + (size_t)row * Dp);
      for (int j0 = 0; j0 < 128; j0 += 32) {
        const int colj = (int)col0 + j0 + lane;
        bool hit = false;
        if (colj < G) hit = val[j0 + lane] < tau * __fadd_rn(a_n, rg[colj]);
        unsigned todo = __ballot_sync(0xffffffffu, hit);
        while (todo) {
          const int b = __ffs(todo) - 1;
          todo &= todo - 1;
          const uint32_t col = col0 + j0 + b;
          const float b_s = sg[col];
          const uint4* bh = reinterpret_cast<const uint4*>(g_hi + (size_t)col * Dp);
          const uint4* bl = reinterpret_cast<const uint4*>(g_lo + (size_t)col * Dp);
          float acc = 0.f;
          for (int i = lane; i < Dp / 8; i += 32) {
            const uint4 x0 = ah[i], x1 = al[i], y0 = bh[i], y1 = bl[i];
            const __half2* xh = reinterpret_cast<const __half2*>(&x0);
            const __half2* xl = reinterpret_cast<const __half2*>(&x1);
            const __half2* yh = reinterpret_cast<const __half2*>(&y0);
            const __half2* yl = reinterpret_cast<const __half2*>(&y1);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float2 a_hi = __half22float2(xh[j]), a_lo = __half22float2(xl[j]);
              const float2 b_hi = __half22float2(yh[j]), b_lo = __half22float2(yl[j]);
              const float d0 = __fsub_rn(__fadd_rn(a_hi.x, a_lo.x) * a_s, __fadd_rn(b_hi.x, b_lo.x) * b_s);
              const float d1 = __fsub_rn(__fadd_rn(a_hi.y, a_lo.y) * a_s, __fadd_rn(b_hi.y, b_lo.y) * b_s);
              acc = __fmaf_rn(d0, d0, acc);
              acc = __fmaf_rn(d1, d1, acc);
            }
          }
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
          if (lane == 0) val[j0 + b] = acc;
        }
        __syncwarp();
      }
    }
    // the row's same-identity items that live in this span: their exact distances
    const int n = fb.tn[row];
    if (lane < n) {
      const int c = fb.tcol[(size_t)row * fb.LC + lane];
      if (c >= (int)col0 && c < (int)col0 + 128) {
        fb.exact[(size_t)row * fb.LC + lane] = val[c - (int)col0];
        atomicOr(fb.tfound + row, 1u << lane);
      }
    }
  }
}

// One warp per spilled span.  The row's RELEVANT thresholds, as exact packed (distance key, gallery index) words, are
// sorted in shared memory (<= 32: a rank sort over lanes); each lane then places its four outputs among them with a
// binary search and bumps a per-position counter; the prefix sums are the counts.
__global__ void __launch_bounds__(256) fused_recount_kernel(int G, FusedBuffers fb) {
  __shared__ uint64_t Ts[8][kFusedLC];
  __shared__ int32_t item[8][kFusedLC];
  __shared__ int32_t bins[8][kFusedLC + 1];
  const unsigned int n_spill = min(*fb.spill_n, fb.spill_cap);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  long long ties = 0;
  for (unsigned int e = warp; e < n_spill; e += nwarps) {
    const unsigned long long rc = fb.spill_meta[e];
    const uint32_t row = (uint32_t)(rc >> 32), col0 = (uint32_t)rc;
    const float4 v = reinterpret_cast<const float4*>(fb.spill_val + (size_t)e * 128)[lane];
    const float d[4] = {v.x, v.y, v.z, v.w};
    const int n = fb.tn[row];
    const uint32_t junk = fb.tjunk[row], found = fb.tfound[row];
    // lane i = item i of the row: its exact key if it is a relevant item whose distance is known
    const bool use = lane < n && !((junk >> lane) & 1u) && ((found >> lane) & 1u);
    const uint64_t mykey = use ? pack_key(fb.exact[(size_t)row * fb.LC + lane], (uint32_t)fb.tcol[(size_t)row * fb.LC + lane]) : ~uint64_t(0);
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const uint64_t kj = __shfl_sync(0xffffffffu, mykey, j);
      rank += (kj < mykey) || (kj == mykey && j < lane);     // unused entries (all ones) sort to the end, in lane order
    }
    const int R = __popc(__ballot_sync(0xffffffffu, use));
    __syncwarp();
    if (lane < n) { Ts[w][rank] = mykey; item[w][rank] = lane; }
    bins[w][lane] = 0;
    if (lane == 0) bins[w][kFusedLC] = 0;
    __syncwarp();
    if (R > 0) {
      int same = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const uint32_t col = col0 + 4 * lane + j;
        if (col >= (uint32_t)G) continue;                    // beyond the gallery: after everything, never counted
        const uint64_t key = pack_key(d[j], col);
        int lo = 0, hi = R;                                  // pos = #{k < R : T_k <= key}
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (Ts[w][mid] <= key) lo = mid + 1; else hi = mid; }
        // an output below T_k is one with pos <= k, unless it IS T_k (pos = k + 1 then)
        atomicAdd(&bins[w][lo], 1);
        // bit-equal distances around the insertion point that are other entries: ties (rank.cu counts the same pairs)
        for (int k = lo - 1; k >= 0 && (uint32_t)(Ts[w][k] >> 32) == (uint32_t)(key >> 32); --k) same += Ts[w][k] != key;
        for (int k = lo; k < R && (uint32_t)(Ts[w][k] >> 32) == (uint32_t)(key >> 32); ++k) same += 1;
      }
      for (int o = 16; o > 0; o >>= 1) same += __shfl_xor_sync(0xffffffffu, same, o);
      if (lane == 0) ties += same;
      __syncwarp();
      // counts: threshold k (sorted) has sum_{b <= k} bins[b] outputs before it
      int c = lane < R ? bins[w][lane] : 0;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, c, o); if (lane >= o) c += t; }
      if (lane < R && c != 0) atomicAdd(fb.cnt + (size_t)row * fb.LC + item[w][lane], c);
    }
    __syncwarp();
  }
  if (lane == 0 && ties) atomicAdd(fb.stats + 1, (unsigned long long)ties);
}

// One warp per query: lane i = same-identity item i.
__global__ void __launch_bounds__(256) fused_finalize_kernel(int64_t Q, int64_t G_total, int max_rank, FusedBuffers fb,
                                                              double* __restrict__ ap, int32_t* __restrict__ first,
                                                              int32_t* __restrict__ is_short, double* __restrict__ inp) {
  __shared__ double terms[8][kFusedLC];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t q = (int64_t)blockIdx.x * 8 + w;
  if (q >= Q) return;
  const int n = fb.tn[q];
  const uint32_t junk = fb.tjunk[q], found = fb.tfound[q];
  const bool mine = lane < n;
  const bool is_junk = mine && ((junk >> lane) & 1u);
  const bool is_rel = mine && !is_junk;
  uint64_t key = ~uint64_t(0);
  unsigned fb_bits = 0;
  if (mine) {
    if (!((found >> lane) & 1u)) {
      fb_bits |= FB_NOT_FOUND;                              // its span was not spilled: the band did not hold
    } else {
      const float ex = fb.exact[(size_t)q * fb.LC + lane], th = fb.thr[(size_t)q * fb.LC + lane];
      if (!(fabsf(ex - th) < fb.eps[q])) fb_bits |= isfinite(ex) ? FB_BAND_VIOLATED : FB_NON_FINITE;
      key = pack_key(ex, (uint32_t)fb.tcol[(size_t)q * fb.LC + lane]);
    }
  }
  const unsigned any_fb = __reduce_or_sync(0xffffffffu, fb_bits);
  if (lane == 0 && any_fb) atomicOr(fb.stats, (unsigned long long)any_fb);
  // position of relevant item i among the kept ones: outputs before it minus junk items before it; its rank among the
  // relevant items orders the AP sum (rank.py:155-160)
  // (the recount counted every other entry of the span with a bit-equal distance as a tie of relevant item i; the staged
  // path does not count junk items, nor entries that are thresholds themselves: rank.cu, exact_bin / junk pass)
  int junk_before = 0, rel_before = 0, junk_ties = 0;
  for (int j = 0; j < n; ++j) {
    const uint64_t kj = __shfl_sync(0xffffffffu, key, j);
    const bool jj = (junk >> j) & 1u;
    const bool same = j != lane && (uint32_t)(kj >> 32) == (uint32_t)(key >> 32);
    if (jj) junk_before += kj < key; else rel_before += kj < key;
    junk_ties += same;
  }
  const int pos = is_rel ? fb.cnt[(size_t)q * fb.LC + lane] - junk_before : 0;
  const unsigned rel_mask = __ballot_sync(0xffffffffu, is_rel);
  const int R = __popc(rel_mask), nj = __popc(__ballot_sync(0xffffffffu, is_junk));
  int tie_fix = is_rel ? junk_ties : 0;                   // junk items were counted as "other kept item" ties by the recount
  for (int o = 16; o > 0; o >>= 1) tie_fix += __shfl_xor_sync(0xffffffffu, tie_fix, o);
  if (lane == 0 && tie_fix) atomicAdd(fb.stats + 1, (unsigned long long)(-(long long)tie_fix));
  if (is_rel) terms[w][rel_before] = (double)(rel_before + 1) / ((double)pos + 1.0);
  int first_pos = is_rel && rel_before == 0 ? pos : 0, last_pos = is_rel && rel_before == R - 1 ? pos : 0;
  for (int o = 16; o > 0; o >>= 1) { first_pos += __shfl_xor_sync(0xffffffffu, first_pos, o); last_pos += __shfl_xor_sync(0xffffffffu, last_pos, o); }
  __syncwarp();
  if (lane == 0) {
    if (R == 0) {
      ap[q] = 0.0; first[q] = -1; is_short[q] = 0;
      if (inp) inp[q] = 0.0;
    } else {
      double s = 0.0;
      for (int k = 0; k < R; ++k) s += terms[w][k];         // in rank order, as rank_query_kernel adds them
      ap[q] = s / (double)R;
      first[q] = first_pos;
      is_short[q] = (G_total - (int64_t)nj) < max_rank ? 1 : 0;
      if (inp) inp[q] = (double)R / ((double)last_pos + 1.0);
    }
  }
}

// ---- host side ------------------------------------------------------------------------------------------------
int distmat_umma_fused(const void* q_packed, int64_t Q, const void* g_packed, int64_t G, int64_t D, int metric,
                       const FusedCount& fc, cudaStream_t stream, int cta_group);
int rank_reduce(const double* ap, const int32_t* first, const int32_t* short_list, int64_t Q, int32_t max_rank,
                const unsigned long long* ties, float* cmc, ieee_eval_summary* summary, const double* inp,
                const int32_t* overflow, cudaStream_t stream, const PeerView* peers, long long* stats_out);

static uint32_t fused_spill_cap(int64_t Q, int64_t G) {
  const int64_t spans = Q * ((G + 127) / 128);
  int64_t cap = spans / 4 + 8192;
  if (cap > spans) cap = spans;
  return (uint32_t)(cap < 1 ? 1 : cap);
}

static size_t fused_carve(uint8_t* base, int64_t Q, int64_t G, int64_t T, FusedBuffers* fb) {
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align256(bytes); return base ? base + r : nullptr; };
  const int LC = kFusedLC;
  const uint32_t scap = fused_spill_cap(Q, G);
  FusedBuffers b;
  b.LC = LC;
  b.spill_cap = scap;
  // zeroed region first: cnt, tfound, spill_n, stats
  b.cnt = reinterpret_cast<int32_t*>(take(size_t(Q) * LC * 4));
  b.tfound = reinterpret_cast<uint32_t*>(take(size_t(Q) * 4));
  b.spill_n = reinterpret_cast<unsigned int*>(take(256));
  b.stats = reinterpret_cast<unsigned long long*>(reinterpret_cast<uint8_t*>(b.spill_n) + 64);
  b.n_active = reinterpret_cast<unsigned int*>(reinterpret_cast<uint8_t*>(b.spill_n) + 128);
  const size_t zero_bytes = o;
  b.qhead = reinterpret_cast<int32_t*>(take(size_t(T) * 4));
  b.active = reinterpret_cast<int32_t*>(take(size_t(Q < T ? Q : T) * 4));
  b.qnext = reinterpret_cast<int32_t*>(take(size_t(Q) * 4));
  b.qslot = reinterpret_cast<int32_t*>(take(size_t(Q) * 4));
  b.thr = reinterpret_cast<float*>(take(size_t(Q) * LC * 4));
  b.tcol = reinterpret_cast<int32_t*>(take(size_t(Q) * LC * 4));
  b.tjunk = reinterpret_cast<uint32_t*>(take(size_t(Q) * 4));
  b.tn = reinterpret_cast<int32_t*>(take(size_t(Q) * 4));
  b.eps = reinterpret_cast<float*>(take(size_t(Q) * 4));
  b.exact = reinterpret_cast<float*>(take(size_t(Q) * LC * 4));
  b.spill_meta = reinterpret_cast<unsigned long long*>(take(size_t(scap) * 8));
  b.spill_val = reinterpret_cast<float*>(take(size_t(scap) * 128 * 4));
  if (fb) { *fb = b; fb->spill_cap = scap; }
  (void)zero_bytes;
  return o;
}

static int64_t group_table_size(int64_t G) {
  int64_t T = 16;
  while (T < 2 * G) T <<= 1;
  return T;
}

size_t fused_workspace_bytes(int64_t Q, int64_t G) {
  return fused_carve(nullptr, Q, G, group_table_size(G), nullptr) + 2 * align256(size_t(Q) * 8) + 2 * align256(size_t(Q) * 4) + 256;
}

// Queue the fused evaluation of one query block against a prepared gallery.  stats_out (device, uint64[2]): [0] != 0
// means the result could not be certified (bit mask of FusedFallback) and must be discarded.
int fused_eval(const void* q_packed, int64_t Q, const void* g_packed, const void* group, int64_t G, int64_t D, int metric,
               const int64_t* q_pids, const int64_t* q_camids, const int64_t* g_camids, int32_t max_rank, float* cmc,
               ieee_eval_summary* summary, double* per_query_ap, int32_t* per_query_first, unsigned long long* stats_out,
               void* workspace, size_t workspace_bytes, cudaStream_t stream, int cta_group) {
  IEEE_REQUIRE(workspace_bytes >= fused_workspace_bytes(Q, G), "fused eval: workspace too small (%zu < %zu)", workspace_bytes,
               fused_workspace_bytes(Q, G));
  const int64_t T = group_table_size(G);
  FusedBuffers fb;
  uint8_t* base = static_cast<uint8_t*>(workspace);
  size_t used = fused_carve(base, Q, G, T, &fb);
  double* ap = per_query_ap ? per_query_ap : reinterpret_cast<double*>(base + used);
  used += align256(size_t(Q) * 8);
  double* inp = reinterpret_cast<double*>(base + used);
  used += align256(size_t(Q) * 8);
  int32_t* first = per_query_first ? per_query_first : reinterpret_cast<int32_t*>(base + used);
  used += align256(size_t(Q) * 4);
  int32_t* is_short = reinterpret_cast<int32_t*>(base + used);
  // zero: cnt, tfound, spill_n + stats (contiguous at the front of the carve); qhead = -1
  IEEE_CUDA_CHECK(cudaMemsetAsync(fb.cnt, 0, reinterpret_cast<uint8_t*>(fb.qhead) - reinterpret_cast<uint8_t*>(fb.cnt), stream));
  IEEE_CUDA_CHECK(cudaMemsetAsync(fb.qhead, 0xFF, size_t(T) * 4, stream));
  const GroupTables gt = group_tables(group, G);
  fused_link_kernel<<<(unsigned)((Q + 255) / 256), 256, 0, stream>>>(q_pids, Q, gt.keys, gt.T, fb);
  PackedLayout lq = packed_layout(Q, D, IEEE_PREC_F16X3), lg = packed_layout(G, D, IEEE_PREC_F16X3);
  const uint8_t* qb = static_cast<const uint8_t*>(q_packed);
  const uint8_t* gb = static_cast<const uint8_t*>(g_packed);
  const bool euclid = metric == IEEE_METRIC_EUCLIDEAN;
  const __half* qh = reinterpret_cast<const __half*>(qb + lq.hi_off);
  const __half* ql = reinterpret_cast<const __half*>(qb + lq.lo_off);
  const __half* gh = reinterpret_cast<const __half*>(gb + lg.hi_off);
  const __half* gl = reinterpret_cast<const __half*>(gb + lg.lo_off);
  const float* sq = reinterpret_cast<const float*>(qb + lq.scale_off);
  const float* sg = reinterpret_cast<const float*>(gb + lg.scale_off);
  const float* rq = euclid ? reinterpret_cast<const float*>(qb + lq.norm_off) : nullptr;
  const float* rg = euclid ? reinterpret_cast<const float*>(gb + lg.norm_off) : nullptr;
  int chunk_rows = (int)((200 * 1024) / (lq.Dp * 4));
  if (chunk_rows > kPreRows) chunk_rows = kPreRows;
  IEEE_REQUIRE(chunk_rows >= 1, "fused eval: feature dimension %lld too large for the pre-pass", (long long)D);
  const size_t smem = size_t(chunk_rows) * lq.Dp * 4;
  IEEE_ENSURE_DYN_SMEM(fused_prepass_kernel, smem);
  fused_prepass_kernel<<<sm_count(), 256, smem, stream>>>(qh, ql, gh, gl, sq, sg, rq, rg, (int)lq.Dp, euclid ? -2.0f : -1.0f,
                                                          q_camids, g_camids, gt, fb, chunk_rows);
  count_launch(2);
  IEEE_CUDA_CHECK(cudaGetLastError());
  FusedCount fc;
  fc.thr = fb.thr; fc.tn = fb.tn; fc.eps = fb.eps; fc.cnt = fb.cnt; fc.spill_n = fb.spill_n; fc.spill_meta = fb.spill_meta;
  fc.spill_val = fb.spill_val; fc.spill_cap = fb.spill_cap; fc.LC = fb.LC;
  int rc = distmat_umma_fused(q_packed, Q, g_packed, G, D, metric, fc, stream, cta_group);
  if (rc) return rc;
  const float tau = (euclid && !(g_debug_flags & 32)) ? kFusedTau : 0.f;
  const int grid = sm_count() * 4;
  fused_extract_kernel<<<grid, 256, 0, stream>>>(qh, ql, gh, gl, sq, sg, rq, rg, (int)lq.Dp, (int)G, tau, fb);
  fused_recount_kernel<<<grid, 256, 0, stream>>>((int)G, fb);
  if (max_rank > G) max_rank = (int32_t)G;
  fused_finalize_kernel<<<(unsigned)((Q + 7) / 8), 256, 0, stream>>>(Q, G, max_rank, fb, ap, first, is_short, inp);
  count_launch(3);
  IEEE_CUDA_CHECK(cudaGetLastError());
  if ((rc = rank_reduce(ap, first, is_short, Q, max_rank, fb.stats + 1, cmc, summary, inp, nullptr, stream, nullptr, nullptr))) return rc;
  // fallback word: the reasons the kernels raised, plus a full spill area
  IEEE_CUDA_CHECK(cudaMemcpyAsync(stats_out, fb.stats, 16, cudaMemcpyDeviceToDevice, stream));
  IEEE_CUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<uint8_t*>(stats_out) + 16, fb.spill_n, 4, cudaMemcpyDeviceToDevice, stream));
  return IEEE_OK;
}

uint32_t fused_spill_capacity(int64_t Q, int64_t G) { return fused_spill_cap(Q, G); }

}  // namespace ieee
