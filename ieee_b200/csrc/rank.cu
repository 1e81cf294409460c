// Ranking, CMC and mAP (Market-1501 protocol) without sorting the distance rows.
//
// Replaces torchreid/metrics/rank.py:103-171 (eval_market1501) and rank_cylib/rank_cy.pyx:156-243.
// The reference argsorts every row (rank.py:117), drops same-pid-same-camera gallery items (:136-140), then walks
// the kept list.  CMC and AP only depend on the POSITIONS of the relevant items among the kept ones,
//     pos(r) = #{ kept g : (d[q,g], g) <lex (d[q,r], r) },
// so one streaming pass over the row, binning every distance against the (few) sorted relevant distances, gives
// the same numbers: bit-exact CMC, fp64 AP, ties broken by gallery index, and integer partial counts that add up
// across gallery shards.  HBM-bound: 4 bytes per (query, gallery) pair, read once.
#include "common.cuh"

namespace ieee {

static constexpr uint64_t kPadKey = ~uint64_t(0);

// ---------------------------------------------------------------------------------------------------------
// CTA-wide bitonic sort of n (power of two) uint64 keys in shared memory, ascending.
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_bitonic_sort(uint64_t* s, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const uint64_t a = s[i], b = s[p];
          const bool up = (i & k) == 0;
          if ((a > b) == up) { s[i] = b; s[p] = a; }
        }
      }
    }
  }
  __syncthreads();
}
__host__ __device__ __forceinline__ int next_pow2(int x) {
  int p = 1;
  while (p < x) p <<= 1;
  return p;
}

// ---------------------------------------------------------------------------------------------------------
// Gallery grouping: sort (pid, local index) ascending.  Bitonic: chunks of kSortChunk in shared memory, wider
// compare-exchange distances in global memory.  Once per gallery (shard), reused by every query block.
// ---------------------------------------------------------------------------------------------------------
struct GroupView {
  int64_t* pids;   // [Gp]
  int32_t* idx;    // [Gp]
  int64_t Gp;
};
static inline int64_t group_padded(int64_t G) {
  int64_t p = 1;
  while (p < G) p <<= 1;
  return p < 2 ? 2 : p;
}
static inline GroupView group_view(const void* blob, int64_t G) {
  GroupView v;
  v.Gp = group_padded(G);
  uint8_t* b = static_cast<uint8_t*>(const_cast<void*>(blob));
  v.pids = reinterpret_cast<int64_t*>(b);
  v.idx = reinterpret_cast<int32_t*>(b + align256(size_t(v.Gp) * 8));
  return v;
}
size_t gallery_group_bytes(int64_t G) {
  const int64_t Gp = group_padded(G);
  return align256(size_t(Gp) * 8) + align256(size_t(Gp) * 4);
}

constexpr int kSortChunk = 8192;   // (pid, idx) pairs sorted per CTA in shared memory: 96 KB

__device__ __forceinline__ bool pair_greater(int64_t pa, int32_t ia, int64_t pb, int32_t ib) {
  return pa > pb || (pa == pb && ia > ib);
}

// k_start == 2 : load the raw ids (fused initialisation: entry i = (g_pids[i], i), padding sorts last) and fully
//                sort each chunk (all k <= chunk).
// k_start  > 2 : the j < chunk tail of merge step k = k_start on already initialised arrays.
__global__ void __launch_bounds__(1024) group_sort_local_kernel(const int64_t* __restrict__ g_pids, int64_t G, int64_t* pids,
                                                                 int32_t* idx, int64_t Gp, int64_t k_start) {
  extern __shared__ __align__(16) uint8_t gs_raw[];
  int64_t* sp = reinterpret_cast<int64_t*>(gs_raw);
  int32_t* si = reinterpret_cast<int32_t*>(sp + kSortChunk);
  const int64_t base = (int64_t)blockIdx.x * kSortChunk;
  const int n = (int)min((int64_t)kSortChunk, Gp - base);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const int64_t gi = base + i;
    if (k_start == 2) {
      sp[i] = gi < G ? g_pids[gi] : INT64_MAX;
      si[i] = gi < G ? (int32_t)gi : INT32_MAX;
    } else {
      sp[i] = pids[gi];
      si[i] = idx[gi];
    }
  }
  const int64_t k_end = (k_start == 2) ? n : k_start;
  for (int64_t k = k_start; k <= k_end; k <<= 1) {
    int j0 = (int)min(k >> 1, (int64_t)(n >> 1));
    for (int j = j0; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const bool up = ((base + i) & k) == 0;
          if (pair_greater(sp[i], si[i], sp[p], si[p]) == up) {
            int64_t tp = sp[i]; sp[i] = sp[p]; sp[p] = tp;
            int32_t ti = si[i]; si[i] = si[p]; si[p] = ti;
          }
        }
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) { pids[base + i] = sp[i]; idx[base + i] = si[i]; }
}

__global__ void group_sort_global_kernel(int64_t* pids, int32_t* idx, int64_t Gp, int64_t j, int64_t k) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Gp) return;
  const int64_t p = i ^ j;
  if (p > i) {
    const bool up = (i & k) == 0;
    const int64_t pa = pids[i], pb = pids[p];
    const int32_t ia = idx[i], ib = idx[p];
    if (pair_greater(pa, ia, pb, ib) == up) { pids[i] = pb; pids[p] = pa; idx[i] = ib; idx[p] = ia; }
  }
}

int gallery_group(const int64_t* g_pids, int64_t G, void* blob, cudaStream_t stream) {
  IEEE_REQUIRE(g_pids && blob && G > 0 && G < (int64_t(1) << 31), "gallery_group: bad arguments (G=%lld)", (long long)G);
  GroupView v = group_view(blob, G);
  const int threads = 256;
  const size_t smem = size_t(kSortChunk) * 12;
  static bool attr_set = false;
  if (!attr_set) {
    IEEE_CUDA_CHECK(cudaFuncSetAttribute(group_sort_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set = true;
  }
  const unsigned chunks = (unsigned)((v.Gp + kSortChunk - 1) / kSortChunk);
  group_sort_local_kernel<<<chunks, 1024, smem, stream>>>(g_pids, G, v.pids, v.idx, v.Gp, 2);
  count_launch();
  for (int64_t k = 2 * (int64_t)kSortChunk; k <= v.Gp; k <<= 1) {
    for (int64_t j = k >> 1; j >= kSortChunk; j >>= 1) {
      group_sort_global_kernel<<<(unsigned)((v.Gp + threads - 1) / threads), threads, 0, stream>>>(v.pids, v.idx, v.Gp, j, k);
      count_launch();
    }
    group_sort_local_kernel<<<chunks, 1024, smem, stream>>>(g_pids, G, v.pids, v.idx, v.Gp, k);
    count_launch();
  }
  IEEE_CUDA_CHECK(cudaGetLastError());
  return IEEE_OK;
}

// [lo, hi) of `pid` in the sorted id array.
__device__ __forceinline__ void pid_range(const int64_t* __restrict__ sp, int64_t n, int64_t pid, int64_t& lo, int64_t& hi) {
  int64_t a = 0, b = n;
  while (a < b) { int64_t m = (a + b) >> 1; if (sp[m] < pid) a = m + 1; else b = m; }
  lo = a;
  b = n;
  while (a < b) { int64_t m = (a + b) >> 1; if (sp[m] <= pid) a = m + 1; else b = m; }
  hi = a;
}

__global__ void list_cap_kernel(const int64_t* __restrict__ sp, int64_t Gp, const int64_t* __restrict__ q_pids, int64_t Q,
                                int32_t* cap_out) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int32_t n = 0;
  if (q < Q) {
    int64_t lo, hi;
    pid_range(sp, Gp, q_pids[q], lo, hi);
    n = (int32_t)(hi - lo);
  }
  for (int o = 16; o > 0; o >>= 1) n = max(n, __shfl_xor_sync(0xffffffffu, n, o));
  if ((threadIdx.x & 31) == 0 && n > 0) atomicMax(cap_out, n);
}

int rank_list_cap(const void* group, int64_t G, const int64_t* q_pids, int64_t Q, int32_t* cap_dev, cudaStream_t stream) {
  GroupView v = group_view(group, G);
  IEEE_CUDA_CHECK(cudaMemsetAsync(cap_dev, 0, 4, stream));
  if (Q > 0) {
    list_cap_kernel<<<(unsigned)((Q + 255) / 256), 256, 0, stream>>>(v.pids, v.Gp, q_pids, Q, cap_dev);
    count_launch();
  }
  IEEE_CUDA_CHECK(cudaGetLastError());
  return IEEE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// gather: one warp per query.  rank.py:136: junk = same pid AND same camera; relevant = same pid, other camera.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rank_gather_kernel(const float* __restrict__ distmat, int64_t ld, int64_t Q, int64_t G,
                                                           const int64_t* __restrict__ q_pids, const int64_t* __restrict__ q_camids,
                                                           const int64_t* __restrict__ g_camids, const int64_t* __restrict__ sp,
                                                           const int32_t* __restrict__ sidx, int64_t Gp, int64_t g_offset,
                                                           int32_t cap, uint64_t* __restrict__ rel, int32_t* __restrict__ n_rel,
                                                           uint64_t* __restrict__ junk, int32_t* __restrict__ n_junk,
                                                           int32_t* __restrict__ overflow) {
  const int lane = threadIdx.x & 31;
  const int64_t q = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= Q) return;
  const int64_t pid = q_pids[q], cam = q_camids[q];
  int64_t lo = 0, hi = 0;
  if (lane == 0) pid_range(sp, Gp, pid, lo, hi);
  lo = __shfl_sync(0xffffffffu, lo, 0);
  hi = __shfl_sync(0xffffffffu, hi, 0);
  int nr = 0, nj = 0;
  for (int64_t t0 = lo; t0 < hi; t0 += 32) {
    const int64_t t = t0 + lane;
    bool is_rel = false, is_junk = false;
    uint64_t key = 0;
    if (t < hi) {
      const int32_t gi = sidx[t];
      if (gi >= 0 && gi < G) {   // padding entries carry INT32_MAX
        key = pack_key(distmat[q * ld + gi], (uint32_t)(gi + g_offset));
        is_junk = g_camids[gi] == cam;
        is_rel = !is_junk;
      }
    }
    const unsigned mr = __ballot_sync(0xffffffffu, is_rel), mj = __ballot_sync(0xffffffffu, is_junk);
    const unsigned below = (1u << lane) - 1;
    if (is_rel) { const int s = nr + __popc(mr & below); if (s < cap) rel[q * cap + s] = key; }
    if (is_junk) { const int s = nj + __popc(mj & below); if (s < cap) junk[q * cap + s] = key; }
    nr += __popc(mr);
    nj += __popc(mj);
  }
  if (lane == 0) {
    if (nr > cap || nj > cap) { atomicMax(overflow, max(nr, nj)); nr = min(nr, cap); nj = min(nj, cap); }
    n_rel[q] = nr;
    n_junk[q] = nj;
  }
}

// ---------------------------------------------------------------------------------------------------------
// count: one CTA per query; single streaming pass over the local distance row.
//
// Every distance is binned against the query's sorted thresholds T_0 < ... < T_{R-1} (the relevant items' packed
// (distance key, global index)): b(e) = #{k : T_k <lex e}.  A 1024-cell table over [d(T_0), d(T_{R-1})] maps a
// distance to its bin with one multiply and one shared-memory load; only cells that contain a threshold need
// exact 64-bit compares.  Bin counters are PRIVATE per thread (16-bit, layout [bin][thread]: conflict-free plain
// read-modify-write, no atomics) and are summed once after the stream; queries with more relevant items than
// the private table can hold fall back to shared atomics.
// ---------------------------------------------------------------------------------------------------------
constexpr int kCountThreads = 256;
constexpr int kLutCells = 1024;
constexpr int kPrivateBinBudget = 96 * 1024;   // bytes of private counters per CTA ((R + 2) * 512 B)

__host__ __device__ inline size_t count_smem_bytes(int Rp, bool priv) {
  // T[Rp] u64 | hist[Rp + 2] i32 | cell[L] u32 | misc[64] i32 | priv[(Rp + 2) * threads] u16
  return size_t(Rp) * 8 + size_t(Rp + 2 + kLutCells + 64) * 4 + (priv ? size_t(Rp + 2) * kCountThreads * 2 : 0);
}

// b(e) = lo + #{k in [lo, lo+n) : T_k <lex e}; `same` = #{k : key(T_k) == key(e)}; is_thr = e is itself a threshold
__device__ __forceinline__ int exact_bin(const uint64_t* T, int lo, int n, uint64_t pe, int& same, bool& is_thr) {
  int b = lo;
  const uint32_t ke = (uint32_t)(pe >> 32);
  for (int j = lo; j < lo + n; ++j) {
    const uint64_t t = T[j];
    b += (t < pe);
    same += ((uint32_t)(t >> 32) == ke);
    is_thr |= (t == pe);
  }
  return b;
}

template <bool kPrivate>
__global__ void __launch_bounds__(kCountThreads)
rank_count_kernel(const float* __restrict__ distmat, int64_t ld, int64_t Q, int64_t G, int64_t g_offset, int shards, int cap,
                  int Rp, const uint64_t* __restrict__ rel_all, const int32_t* __restrict__ n_rel_all,
                  const uint64_t* __restrict__ junk, const int32_t* __restrict__ n_junk, int32_t* __restrict__ counts,
                  unsigned long long* __restrict__ ties_out) {
  extern __shared__ __align__(16) uint8_t cs_raw[];
  uint64_t* T = reinterpret_cast<uint64_t*>(cs_raw);
  int32_t* hist = reinterpret_cast<int32_t*>(cs_raw + size_t(Rp) * 8);
  uint32_t* cell = reinterpret_cast<uint32_t*>(hist + Rp + 2);     // first bin of the cell | (#thresholds in it) << 20
  int32_t* misc = reinterpret_cast<int32_t*>(cell + kLutCells);    // [0] R, [1] ties (signed), [2..] scan scratch
  uint16_t* priv = reinterpret_cast<uint16_t*>(misc + 64);         // [(R + 1)][kCountThreads]
  const int64_t q = blockIdx.x;
  const int tid = threadIdx.x;
  const int stride = shards * cap + 1;
  int32_t* out = counts + q * stride;

  // ---- thresholds: union of the shards' relevant lists -------------------------------------------------
  for (int i = tid; i < Rp; i += kCountThreads) T[i] = kPadKey;
  for (int i = tid; i < Rp + 2; i += kCountThreads) hist[i] = 0;
  for (int i = tid; i < kLutCells; i += kCountThreads) cell[i] = 0;
  if (tid == 0) { misc[0] = 0; misc[1] = 0; }
  __syncthreads();
  for (int s = 0; s < shards; ++s) {
    const int n = n_rel_all[(int64_t)s * Q + q];
    const uint64_t* src = rel_all + ((int64_t)s * Q + q) * cap;
    __shared__ int base_s;
    if (tid == 0) { base_s = misc[0]; misc[0] += n; }
    __syncthreads();
    for (int i = tid; i < n; i += kCountThreads) T[base_s + i] = src[i];
    __syncthreads();
  }
  const int R = misc[0];
  const int nj = n_junk[q];
  if (tid == 0) out[stride - 1] = nj;
  if (R == 0) return;   // invalid query (rank.py:142-144): nothing to rank against
  block_bitonic_sort(T, next_pow2(max(R, 2)));
  if constexpr (kPrivate) {
    uint32_t* pz = reinterpret_cast<uint32_t*>(priv);
    for (int i = tid; i < (R + 2) * kCountThreads / 2; i += kCountThreads) pz[i] = 0;
  }

  // ---- cell table over [lo, hi] of the threshold distances ----------------------------------------------
  const uint32_t kmin = (uint32_t)(T[0] >> 32), kmax = (uint32_t)(T[R - 1] >> 32);
  const float lo = key_to_float(kmin), hi = key_to_float(kmax);
  const float span = hi - lo;
  const bool use_lut = (kmax != 0xFFFFFFFFu) && isfinite(lo) && isfinite(hi) && span > 0.f &&
                       isfinite((float)kLutCells / span) && R < (1 << 11);
  const float scale = use_lut ? (float)kLutCells / span : 0.f;
  if (use_lut) {
    for (int k = tid; k < R; k += kCountThreads) {
      const float d = key_to_float((uint32_t)(T[k] >> 32));
      const int c = min((int)((d - lo) * scale), kLutCells - 1);
      atomicAdd(&cell[c], 1u << 20);
    }
    __syncthreads();
    // exclusive scan of the per-cell counts -> first bin of each cell (kLutCells == 4 * kCountThreads)
    {
      int v[4], sum = 0;
#pragma unroll
      for (int j = 0; j < 4; ++j) { v[j] = (int)(cell[tid * 4 + j] >> 20); sum += v[j]; }
      int incl = sum;
      const int lane = tid & 31, w = tid >> 5;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) { int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
      int32_t* wsum = misc + 2;
      if (lane == 31) wsum[w] = incl;
      __syncthreads();
      int woff = 0;
      for (int i = 0; i < w; ++i) woff += wsum[i];
      int run = woff + incl - sum;
#pragma unroll
      for (int j = 0; j < 4; ++j) { cell[tid * 4 + j] = (uint32_t)run | ((uint32_t)v[j] << 20); run += v[j]; }
    }
  }
  __syncthreads();

  // ---- stream the row ---------------------------------------------------------------------------------------
  const float* row = distmat + q * ld;
  int tie_local = 0;
  // 16-bit slot of this thread inside a bin's 256 counters: 32-bit word = lane + 32 * (warp & 3), half = warp >> 2,
  // so the 32 lanes of a warp always touch 32 different banks whatever their bins are
  const int priv_slot = (((tid & 31) + 32 * ((tid >> 5) & 3)) << 1) | (tid >> 7);
  // Branch-free per-element body (lanes must stay converged across the 16 elements in flight): every element
  // increments exactly one counter -- bin 0 if it precedes all thresholds, the trash bin R + 1 if it follows them
  // (or is NaN), else its bin from the cell table; only cells that hold a threshold take a (reconverging) branch.
  const int trash = R + 1;
  const bool may_wrap = (G + kCountThreads - 1) / kCountThreads >= 0xFFFF;   // uniform; false below 16.7 M columns
  auto bump = [&](int b) {
    if constexpr (kPrivate) {
      uint16_t* p16 = priv + b * kCountThreads + priv_slot;
      const uint16_t nv = (uint16_t)(*p16 + 1);
      *p16 = nv;
      if (may_wrap && nv == 0xFFFFu) { atomicAdd(&hist[b], 0xFFFF); *p16 = 0; }   // spill before the counter can wrap
    } else {
      atomicAdd(&hist[b], 1);
    }
  };
  auto visit = [&](float d, int64_t g) {
    int b;
    if (use_lut) {
      // thresholds are finite here, so float compares order exactly like the keys (NaN fails d <= hi: ranks last)
      const bool in = d <= hi;
      const bool before = d < lo;
      const int ci = max(0, min((int)((d - lo) * scale), kLutCells - 1));
      const uint32_t ce = cell[ci];
      b = (int)(ce & 0xFFFFFu);
      if (in && !before && (ce >> 20)) {           // the cell holds thresholds: exact (key, index) compares
        const uint64_t pe = pack_key(d, (uint32_t)(g + g_offset));
        int same = 0;
        bool is_thr = false;
        b = exact_bin(T, b, (int)(ce >> 20), pe, same, is_thr);
        if (!is_thr) tie_local += same;
      }
      b = before ? 0 : b;
      b = in ? b : trash;
    } else {                                       // non-finite / degenerate thresholds: search all of T
      const uint32_t ke = order_key(d);
      const uint64_t pe = (uint64_t(ke) << 32) | (uint32_t)(g + g_offset);
      int a = 0, e = R;
      while (a < e) { const int m = (a + e) >> 1; if (T[m] < pe) a = m + 1; else e = m; }
      int same = 0;
      bool is_thr = false;
      for (int j = a; j < R && (uint32_t)(T[j] >> 32) == ke; ++j) { same++; is_thr |= (T[j] == pe); }
      for (int j = a - 1; j >= 0 && (uint32_t)(T[j] >> 32) == ke; --j) same++;
      if (!is_thr) tie_local += same;
      b = (ke > kmax) ? trash : a;
    }
    bump(b);
  };
  {
    const uintptr_t addr = reinterpret_cast<uintptr_t>(row);
    int64_t head = ((16 - (addr & 15)) & 15) >> 2;
    if (head > G) head = G;
    for (int64_t g = tid; g < head; g += kCountThreads) visit(row[g], g);
    const int64_t nvec = (G - head) >> 2;
    const float4* rv = reinterpret_cast<const float4*>(row + head);
    int64_t i = tid;
    for (; i + 3 * kCountThreads < nvec; i += 4 * kCountThreads) {   // 4 independent 16-byte loads in flight
      float4 a0 = __ldcs(rv + i), a1 = __ldcs(rv + i + kCountThreads), a2 = __ldcs(rv + i + 2 * kCountThreads),
             a3 = __ldcs(rv + i + 3 * kCountThreads);
      int64_t g0 = head + 4 * i;
      visit(a0.x, g0); visit(a0.y, g0 + 1); visit(a0.z, g0 + 2); visit(a0.w, g0 + 3);
      g0 += 4 * kCountThreads;
      visit(a1.x, g0); visit(a1.y, g0 + 1); visit(a1.z, g0 + 2); visit(a1.w, g0 + 3);
      g0 += 4 * kCountThreads;
      visit(a2.x, g0); visit(a2.y, g0 + 1); visit(a2.z, g0 + 2); visit(a2.w, g0 + 3);
      g0 += 4 * kCountThreads;
      visit(a3.x, g0); visit(a3.y, g0 + 1); visit(a3.z, g0 + 2); visit(a3.w, g0 + 3);
    }
    for (; i < nvec; i += kCountThreads) {
      const float4 a = __ldcs(rv + i);
      const int64_t g0 = head + 4 * i;
      visit(a.x, g0); visit(a.y, g0 + 1); visit(a.z, g0 + 2); visit(a.w, g0 + 3);
    }
    for (int64_t g = head + 4 * nvec + tid; g < G; g += kCountThreads) visit(row[g], g);
  }
  __syncthreads();
  if constexpr (kPrivate) {   // fold the private counters: warp w sums bins w, w + 8, ...
    const int lane = tid & 31, w = tid >> 5;
    for (int b = w; b <= R; b += kCountThreads / 32) {
      int sum = 0;
#pragma unroll
      for (int j = 0; j < kCountThreads / 32; ++j) sum += priv[b * kCountThreads + j * 32 + lane];
      for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
      if (lane == 0 && sum) atomicAdd(&hist[b], sum);
    }
    __syncthreads();
  }
  // ---- junk items were streamed like everything else: take them out again (rank.py:136-140) -------------
  for (int i = tid; i < nj; i += kCountThreads) {
    const uint64_t pe = junk[q * cap + i];
    const uint32_t ke = (uint32_t)(pe >> 32);
    if (ke > kmax) continue;
    if (ke < kmin) { atomicSub(&hist[0], 1); continue; }
    int a = 0, e = R;
    while (a < e) { const int m = (a + e) >> 1; if (T[m] < pe) a = m + 1; else e = m; }
    for (int j = a; j < R && (uint32_t)(T[j] >> 32) == ke; ++j) tie_local--;
    for (int j = a - 1; j >= 0 && (uint32_t)(T[j] >> 32) == ke; --j) tie_local--;
    atomicSub(&hist[a], 1);
  }
  for (int o = 16; o > 0; o >>= 1) tie_local += __shfl_xor_sync(0xffffffffu, tie_local, o);
  if ((tid & 31) == 0 && tie_local) atomicAdd(&misc[1], tie_local);
  __syncthreads();
  // ---- counts[k] = sum_{b <= k} hist[b] - [T_k is a local row entry (it was binned at b == k)] -----------
  __shared__ int carry;
  if (tid == 0) carry = 0;
  __syncthreads();
  for (int base = 0; base < R; base += kCountThreads) {
    const int k = base + tid;
    int v = (k < R) ? hist[k] : 0;
    const int lane = tid & 31, w = tid >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
    int32_t* wsum = misc + 2;
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    int woff = carry;
    for (int i = 0; i < w; ++i) woff += wsum[i];
    if (k < R) {
      const int64_t gi = (int64_t)(uint32_t)T[k] - g_offset;
      out[k] = woff + incl - ((gi >= 0 && gi < G) ? 1 : 0);
    }
    __syncthreads();
    if (tid == kCountThreads - 1) carry = woff + incl;
    __syncthreads();
  }
  if (tid == 0 && ties_out != nullptr && misc[1] != 0) atomicAdd(ties_out, (unsigned long long)(long long)misc[1]);
}

static inline bool count_use_private(int Rp) { return size_t(Rp + 2) * kCountThreads * 2 <= size_t(kPrivateBinBudget); }
size_t rank_count_smem(int shards, int cap) {
  const int Rp = next_pow2(max(shards * cap, 2));
  return count_smem_bytes(Rp, count_use_private(Rp));
}

int rank_count(const float* distmat, int64_t ld, int64_t Q, int64_t G, int64_t g_offset, int shards, int cap,
               const uint64_t* rel_all, const int32_t* n_rel_all, const uint64_t* junk, const int32_t* n_junk,
               int32_t* counts, unsigned long long* ties, cudaStream_t stream) {
  IEEE_REQUIRE(distmat && rel_all && n_rel_all && junk && n_junk && counts, "rank_count: null pointer");
  IEEE_REQUIRE(Q >= 0 && G > 0 && ld >= G && shards >= 1 && cap >= 1, "rank_count: bad shape");
  if (Q == 0) return IEEE_OK;
  const int Rp = next_pow2(max(shards * cap, 2));
  const bool priv = count_use_private(Rp);
  const size_t smem = count_smem_bytes(Rp, priv);
  IEEE_REQUIRE(smem <= 200 * 1024, "rank_count: shards*cap=%d relevant items per query exceed the shared-memory budget",
               shards * cap);
  static size_t smem_set[2] = {0, 0};
  if (smem > 48 * 1024 && smem > smem_set[priv]) {
    if (priv)
      IEEE_CUDA_CHECK(cudaFuncSetAttribute(rank_count_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    else
      IEEE_CUDA_CHECK(cudaFuncSetAttribute(rank_count_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set[priv] = smem;
  }
  if (priv)
    rank_count_kernel<true><<<(unsigned)Q, kCountThreads, smem, stream>>>(distmat, ld, Q, G, g_offset, shards, cap, Rp, rel_all,
                                                                          n_rel_all, junk, n_junk, counts, ties);
  else
    rank_count_kernel<false><<<(unsigned)Q, kCountThreads, smem, stream>>>(distmat, ld, Q, G, g_offset, shards, cap, Rp, rel_all,
                                                                           n_rel_all, junk, n_junk, counts, ties);
  count_launch();
  IEEE_CUDA_CHECK(cudaGetLastError());
  return IEEE_OK;
}

int rank_gather(const float* distmat, int64_t ld, int64_t Q, int64_t G, const int64_t* q_pids, const int64_t* q_camids,
                const int64_t* g_camids, const void* group, int64_t g_offset, int32_t cap, uint64_t* rel, int32_t* n_rel,
                uint64_t* junk, int32_t* n_junk, int32_t* overflow, cudaStream_t stream) {
  IEEE_REQUIRE(distmat && q_pids && q_camids && g_camids && group && rel && n_rel && junk && n_junk && overflow,
               "rank_gather: null pointer");
  IEEE_REQUIRE(Q >= 0 && G > 0 && ld >= G && cap >= 1, "rank_gather: bad shape");
  IEEE_REQUIRE(g_offset >= 0 && g_offset + G <= (int64_t(1) << 32), "rank_gather: global gallery index must fit 32 bits");
  if (Q == 0) return IEEE_OK;
  GroupView v = group_view(group, G);
  rank_gather_kernel<<<(unsigned)((Q + 7) / 8), 256, 0, stream>>>(distmat, ld, Q, G, q_pids, q_camids, g_camids, v.pids, v.idx,
                                                                  v.Gp, g_offset, cap, rel, n_rel, junk, n_junk, overflow); count_launch();
  IEEE_CUDA_CHECK(cudaGetLastError());
  return IEEE_OK;
}

// ---------------------------------------------------------------------------------------------------------
// finalize: per-query AP / first hit, then one deterministic CTA-wide reduction.
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rank_query_kernel(const int32_t* __restrict__ counts, const int32_t* __restrict__ n_rel_all,
                                                          int64_t Q, int64_t G_total, int shards, int cap, int max_rank,
                                                          double* __restrict__ ap, int32_t* __restrict__ first,
                                                          int32_t* __restrict__ is_short) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= Q) return;
  int R = 0;
  for (int s = 0; s < shards; ++s) R += n_rel_all[(int64_t)s * Q + q];
  const int stride = shards * cap + 1;
  const int32_t* c = counts + q * stride;
  if (R == 0) { ap[q] = 0.0; first[q] = -1; is_short[q] = 0; return; }
  // rank.py:155-160: AP = (1/R) sum_k (k+1) / (pos_k + 1), float64
  double s = 0.0;
  for (int k = 0; k < R; ++k) s += (double)(k + 1) / ((double)c[k] + 1.0);
  ap[q] = s / (double)R;
  first[q] = c[0];
  const int64_t kept = G_total - (int64_t)c[stride - 1];
  is_short[q] = kept < max_rank ? 1 : 0;
}

__global__ void __launch_bounds__(1024) rank_reduce_kernel(const double* __restrict__ ap, const int32_t* __restrict__ first,
                                                            const int32_t* __restrict__ is_short, int64_t Q, int max_rank,
                                                            const unsigned long long* __restrict__ ties, float* __restrict__ cmc,
                                                            ieee_eval_summary* __restrict__ summary) {
  extern __shared__ __align__(16) uint8_t rs_raw[];
  double* sd = reinterpret_cast<double*>(rs_raw);                 // [1024]
  int32_t* hfirst = reinterpret_cast<int32_t*>(sd + 1024);        // [max_rank + 1]
  __shared__ long long s_valid, s_short;
  const int tid = threadIdx.x;
  for (int i = tid; i <= max_rank; i += 1024) hfirst[i] = 0;
  if (tid == 0) { s_valid = 0; s_short = 0; }
  __syncthreads();
  // fixed assignment of queries to threads + fixed tree => the fp64 sum does not depend on scheduling
  double acc = 0.0;
  long long nv = 0, ns = 0;
  for (int64_t q = tid; q < Q; q += 1024) {
    const int f = first[q];
    if (f >= 0) {
      acc += ap[q];
      ++nv;
      ns += is_short[q];
      atomicAdd(&hfirst[f < max_rank ? f : max_rank], 1);
    }
  }
  sd[tid] = acc;
  atomicAdd((unsigned long long*)&s_valid, (unsigned long long)nv);
  atomicAdd((unsigned long long*)&s_short, (unsigned long long)ns);
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (tid < o) sd[tid] += sd[tid + o];
    __syncthreads();
  }
  if (tid == 0) {
    const long long valid = s_valid;
    // rank.py:167-168: float32 sum of {0,1} rows (exact below 2^24) divided by the float count
    int run = 0;
    for (int j = 0; j < max_rank; ++j) {
      run += hfirst[j];
      cmc[j] = valid > 0 ? __fdiv_rn((float)run, (float)valid) : 0.f;
    }
    summary->sum_ap = sd[0];
    summary->mAP = valid > 0 ? sd[0] / (double)valid : 0.0;
    summary->num_valid = valid;
    summary->num_ties = ties ? (int64_t)*ties : 0;
    summary->num_short = s_short;
    summary->max_rank = max_rank;
    summary->status = valid == 0 ? IEEE_ERR_NO_VALID_QUERY : (s_short > 0 ? IEEE_ERR_SHORT_RANK_LIST : IEEE_OK);
    summary->reserved[0] = summary->reserved[1] = 0;
  }
}

size_t rank_finalize_workspace_bytes(int64_t Q) { return align256(size_t(Q) * 8) + 2 * align256(size_t(Q) * 4); }

int rank_query_metrics(const int32_t* counts, const int32_t* n_rel_all, int64_t Q, int64_t G_total, int32_t shards, int32_t cap,
                       int32_t max_rank, double* ap, int32_t* first, int32_t* short_list, cudaStream_t stream) {
  IEEE_REQUIRE(counts && n_rel_all && ap && first && short_list, "rank_query_metrics: null pointer");
  IEEE_REQUIRE(Q > 0 && G_total > 0 && max_rank >= 1 && shards >= 1 && cap >= 1, "rank_query_metrics: bad shape");
  if (max_rank > G_total) max_rank = (int32_t)G_total;   // rank.py:110-115
  rank_query_kernel<<<(unsigned)((Q + 255) / 256), 256, 0, stream>>>(counts, n_rel_all, Q, G_total, shards, cap, max_rank, ap,
                                                                     first, short_list); count_launch();
  IEEE_CUDA_CHECK(cudaGetLastError());
  return IEEE_OK;
}

int rank_reduce(const double* ap, const int32_t* first, const int32_t* short_list, int64_t Q, int32_t max_rank,
                const unsigned long long* ties, float* cmc, ieee_eval_summary* summary, cudaStream_t stream) {
  IEEE_REQUIRE(ap && first && short_list && cmc && summary, "rank_reduce: null pointer");
  IEEE_REQUIRE(Q > 0 && max_rank >= 1 && max_rank <= 8192, "rank_reduce: bad shape (max_rank=%d)", max_rank);
  const size_t smem = 1024 * 8 + size_t(max_rank + 1) * 4;
  rank_reduce_kernel<<<1, 1024, smem, stream>>>(ap, first, short_list, Q, max_rank, ties, cmc, summary); count_launch();
  IEEE_CUDA_CHECK(cudaGetLastError());
  return IEEE_OK;
}

int rank_finalize(const int32_t* counts, const int32_t* n_rel_all, int64_t Q, int64_t G_total, int32_t shards, int32_t cap,
                  int32_t max_rank, const unsigned long long* ties, float* cmc, ieee_eval_summary* summary,
                  double* per_query_ap, int32_t* per_query_first, void* workspace, cudaStream_t stream) {
  IEEE_REQUIRE(workspace != nullptr, "rank_finalize: null workspace");
  if (max_rank > G_total) max_rank = (int32_t)G_total;   // rank.py:110-115
  uint8_t* w = static_cast<uint8_t*>(workspace);
  double* ap = per_query_ap ? per_query_ap : reinterpret_cast<double*>(w);
  int32_t* first = per_query_first ? per_query_first : reinterpret_cast<int32_t*>(w + align256(size_t(Q) * 8));
  int32_t* is_short = reinterpret_cast<int32_t*>(w + align256(size_t(Q) * 8) + align256(size_t(Q) * 4));
  int rc = rank_query_metrics(counts, n_rel_all, Q, G_total, shards, cap, max_rank, ap, first, is_short, stream);
  if (rc) return rc;
  return rank_reduce(ap, first, is_short, Q, max_rank, ties, cmc, summary, stream);
}

// ---------------------------------------------------------------------------------------------------------
// Junk-masked top-k (ranked list): threshold filter into a shared candidate buffer, compacted by bitonic sort.
// ---------------------------------------------------------------------------------------------------------
constexpr int kTopkThreads = 256;
constexpr int kTopkBuf = 2048;   // candidates; compaction keeps k <= kTopkBuf / 2

__global__ void __launch_bounds__(kTopkThreads)
topk_kernel(const float* __restrict__ distmat, int64_t ld, int64_t Q, int64_t G, int64_t g_offset,
            const int64_t* __restrict__ q_pids, const int64_t* __restrict__ q_camids, const int64_t* __restrict__ g_pids,
            const int64_t* __restrict__ g_camids, int k, int32_t* __restrict__ idx_out, float* __restrict__ val_out) {
  __shared__ uint64_t cand[kTopkBuf];
  __shared__ int count;
  __shared__ uint64_t tau_s;
  const int64_t q = blockIdx.x;
  const int tid = threadIdx.x;
  const float* row = distmat + q * ld;
  const bool masked = q_pids != nullptr;
  const int64_t qp = masked ? q_pids[q] : 0, qc = masked ? q_camids[q] : 0;
  if (tid == 0) { count = 0; tau_s = kPadKey; }
  __syncthreads();
  constexpr int kTile = kTopkThreads * 4;
  for (int64_t g0 = 0; g0 < G; g0 += kTile) {
    const uint64_t tau = tau_s;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t g = g0 + j * kTopkThreads + tid;
      if (g < G) {
        const uint64_t pe = pack_key(__ldcs(row + g), (uint32_t)(g + g_offset));
        if (pe < tau) {
          const bool is_junk = masked && g_pids[g] == qp && g_camids[g] == qc;   // rank.py:136
          if (!is_junk) cand[atomicAdd(&count, 1)] = pe;
        }
      }
    }
    __syncthreads();
    const int filled = count;                // snapshot taken between two barriers: every thread sees the same value
    __syncthreads();                         // (without it fast threads would already be appending the next tile)
    if (filled > kTopkBuf - kTile) {         // next tile could overflow: keep the k best, tighten the threshold
      const int n = filled;
      for (int i = n + tid; i < kTopkBuf; i += kTopkThreads) cand[i] = kPadKey;
      block_bitonic_sort(cand, kTopkBuf);
      if (tid == 0) { count = min(n, k); if (n >= k) tau_s = cand[k - 1]; }
      __syncthreads();
    }
  }
  const int n = count;
  const int np = next_pow2(max(n, 2));
  for (int i = n + tid; i < np; i += kTopkThreads) cand[i] = kPadKey;
  block_bitonic_sort(cand, np);
  for (int i = tid; i < k; i += kTopkThreads) {
    if (i < n) {
      const uint32_t gi = (uint32_t)cand[i];
      idx_out[q * k + i] = (int32_t)gi;
      val_out[q * k + i] = row[(int64_t)gi - g_offset];
    } else {
      idx_out[q * k + i] = -1;
      val_out[q * k + i] = __int_as_float(0x7f800000);
    }
  }
}

int topk(const float* distmat, int64_t ld, int64_t Q, int64_t G, int64_t g_offset, const int64_t* q_pids,
         const int64_t* q_camids, const int64_t* g_pids, const int64_t* g_camids, int32_t k, int32_t* idx, float* val,
         cudaStream_t stream) {
  IEEE_REQUIRE(distmat && idx && val, "topk: null pointer");
  IEEE_REQUIRE(Q >= 0 && G > 0 && ld >= G && k >= 1 && k <= kTopkBuf / 2, "topk: bad shape (k=%d, max %d)", k, kTopkBuf / 2);
  IEEE_REQUIRE((q_pids == nullptr) == (q_camids == nullptr) && (q_pids == nullptr || (g_pids && g_camids)),
               "topk: junk masking needs all four label arrays");
  IEEE_REQUIRE(g_offset >= 0 && g_offset + G < (int64_t(1) << 31), "topk: global gallery index must fit int32");
  if (Q == 0) return IEEE_OK;
  topk_kernel<<<(unsigned)Q, kTopkThreads, 0, stream>>>(distmat, ld, Q, G, g_offset, q_pids, q_camids, g_pids, g_camids, k, idx, val); count_launch();
  IEEE_CUDA_CHECK(cudaGetLastError());
  return IEEE_OK;
}

// Merge per-shard lists: the global top-k is the k smallest (distance, index) among the shards' k best.
__global__ void __launch_bounds__(256) topk_merge_kernel(const int32_t* __restrict__ idx_all, const float* __restrict__ val_all,
                                                          int shards, int64_t Q, int k, int np, int32_t* __restrict__ idx,
                                                          float* __restrict__ val) {
  extern __shared__ __align__(16) uint8_t ms_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(ms_raw);   // [np]
  const int64_t q = blockIdx.x;
  const int total = shards * k;
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    uint64_t key = kPadKey;
    if (i < total) {
      const int s = i / k, j = i - s * k;
      const int32_t gi = idx_all[((int64_t)s * Q + q) * k + j];
      if (gi >= 0) key = pack_key(val_all[((int64_t)s * Q + q) * k + j], (uint32_t)gi);
    }
    keys[i] = key;
  }
  block_bitonic_sort(keys, np);
  for (int i = threadIdx.x; i < k; i += blockDim.x) {
    const uint64_t key = keys[i];
    if (key == kPadKey) { idx[q * k + i] = -1; val[q * k + i] = __int_as_float(0x7f800000); continue; }
    idx[q * k + i] = (int32_t)(uint32_t)key;
    // recover the exact distance bits from the shard list that held this entry
    float v = key_to_float((uint32_t)(key >> 32));
    for (int j = 0; j < total; ++j) {
      const int s = j / k, jj = j - s * k;
      if (idx_all[((int64_t)s * Q + q) * k + jj] == (int32_t)(uint32_t)key) { v = val_all[((int64_t)s * Q + q) * k + jj]; break; }
    }
    val[q * k + i] = v;
  }
}

int topk_merge(const int32_t* idx_all, const float* val_all, int32_t shards, int64_t Q, int32_t k, int32_t* idx, float* val,
               cudaStream_t stream) {
  IEEE_REQUIRE(idx_all && val_all && idx && val && shards >= 1 && k >= 1, "topk_merge: bad arguments");
  const int np = next_pow2(max(shards * k, 2));
  IEEE_REQUIRE(np <= 16384, "topk_merge: shards*k=%d too large", shards * k);
  if (Q == 0) return IEEE_OK;
  const size_t smem = size_t(np) * 8;
  static size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    IEEE_CUDA_CHECK(cudaFuncSetAttribute(topk_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  topk_merge_kernel<<<(unsigned)Q, 256, smem, stream>>>(idx_all, val_all, shards, Q, k, np, idx, val); count_launch();
  IEEE_CUDA_CHECK(cudaGetLastError());
  return IEEE_OK;
}

}  // namespace ieee
