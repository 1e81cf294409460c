// k-reciprocal re-ranking (Zhong et al., CVPR 2017) -- replaces torchreid/utils/rerank.py:31-113.
//
// Stages, N = Q + G:
//   K1  original_dist[i][j] = raw[j][i]^2 / max_k raw[k][i]^2 over the virtual block matrix
//       raw = [[qq, qg], [qg^T, gg]] (rerank.py:36-46: squared again, column-normalised, transposed)
//   K2  initial_rank = first k1+1 entries of each row's ascending order (rerank.py:48; only those prefixes are read)
//   K3  k-reciprocal sets + half-k expansion + Gaussian weights -> row-sparse V (rerank.py:54-82)
//   K4  query expansion: V_qe[i] = mean of the k2 rows V[initial_rank[i, :k2]] (rerank.py:84-89), still sparse
//   K5  Jaccard distance through the inverted index (rerank.py:91-106), accumulated in ascending column order
//   K6  final = jaccard * (1 - lambda) + original_dist * lambda, rows < Q, columns >= Q (rerank.py:108-113)
// V never exists densely: rows hold at most (k1+1)(round(k1/2)+2) entries, expanded rows k2 times that.
// Ties in K2 are broken by index (the reference's unstable argsort leaves them undefined).
#include "common.cuh"

namespace ieee {

int topk(const float* distmat, int64_t ld, int64_t Q, int64_t G, int64_t g_offset, const int64_t* q_pids,
         const int64_t* q_camids, const int64_t* g_pids, const int64_t* g_camids, int32_t k, int32_t* idx, float* val,
         cudaStream_t stream);

struct RawView {            // the virtual N x N matrix [[qq, qg], [qg^T, gg]]
  const float *qg, *qq, *gg;
  int64_t ld_qg, ld_qq, ld_gg;
  int Q, G;
  __device__ __forceinline__ float at(int r, int c) const {
    if (r < Q) return c < Q ? qq[(int64_t)r * ld_qq + c] : qg[(int64_t)r * ld_qg + (c - Q)];
    return c < Q ? qg[(int64_t)c * ld_qg + (r - Q)] : gg[(int64_t)(r - Q) * ld_gg + (c - Q)];
  }
};

// ---- K1a: column maxima of raw^2 ------------------------------------------------------------------------
// values are squares (>= 0), so the float order equals the order of their bit patterns: integer atomicMax.
__global__ void __launch_bounds__(256) rr_colmax_kernel(const float* __restrict__ m, int64_t ld, int rows, int cols,
                                                         int rows_per_block, int* __restrict__ colmax_bits) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  const int r0 = blockIdx.y * rows_per_block, r1 = min(rows, r0 + rows_per_block);
  float mx = 0.f;
  for (int r = r0; r < r1; ++r) {
    const float x = m[(int64_t)r * ld + c];
    mx = fmaxf(mx, __fmul_rn(x, x));
  }
  atomicMax(colmax_bits + c, __float_as_int(mx));
}
// maxima over the ROWS of qg: the qg^T block's contribution to columns < Q
__global__ void __launch_bounds__(256) rr_rowmax_kernel(const float* __restrict__ m, int64_t ld, int rows, int cols,
                                                         int* __restrict__ colmax_bits) {
  const int r = blockIdx.x;
  float mx = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) {
    const float x = m[(int64_t)r * ld + c];
    mx = fmaxf(mx, __fmul_rn(x, x));
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) atomicMax(colmax_bits + r, __float_as_int(mx));
}

// ---- K1b: orig[i][j] = raw[j][i]^2 / colmax[i]; 32 x 32 tiles transposed through shared memory -----------
__global__ void __launch_bounds__(256) rr_build_orig_kernel(RawView raw, int N, const float* __restrict__ colmax,
                                                             float* __restrict__ orig) {
  __shared__ float tile[32][33];
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  // raw[j][i] is contiguous in i except in the qg^T block (j >= Q, i < Q), which is contiguous in j
  const bool direct = (j0 >= raw.Q) && (i0 + 31 < raw.Q);
  if (direct) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int i = i0 + ty + 8 * k, j = j0 + tx;
      if (i < N && j < N) {
        const float x = raw.at(j, i);
        orig[(int64_t)i * N + j] = __fdiv_rn(__fmul_rn(x, x), colmax[i]);
      }
    }
    return;
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int j = j0 + ty + 8 * k, i = i0 + tx;
    tile[ty + 8 * k][tx] = (i < N && j < N) ? raw.at(j, i) : 0.f;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int i = i0 + ty + 8 * k, j = j0 + tx;
    if (i < N && j < N) {
      const float x = tile[tx][ty + 8 * k];
      orig[(int64_t)i * N + j] = __fdiv_rn(__fmul_rn(x, x), colmax[i]);
    }
  }
}

// ---- K3: k-reciprocal sets, expansion, weights; one warp per row ----------------------------------------
constexpr int kRowWarps = 4;

// members of `row`'s first kk neighbours whose own first kk neighbours contain `row` (rerank.py:56-59, :63-71),
// appended to `out` (shared memory) in neighbour order; returns the new count.
__device__ int reciprocal_append(const int32_t* __restrict__ rank, int ldr, int row, int kk, int lane, int* out, int n_out) {
  for (int t0 = 0; t0 < kk; t0 += 32) {
    const int t = t0 + lane;
    bool ok = false;
    int mine = -1;
    if (t < kk) {
      mine = rank[(int64_t)row * ldr + t];
      if (mine >= 0) {
        const int32_t* back = rank + (int64_t)mine * ldr;
        for (int u = 0; u < kk; ++u) ok |= (back[u] == row);
      }
    }
    const unsigned m = __ballot_sync(0xffffffffu, ok);
    if (ok) out[n_out + __popc(m & ((1u << lane) - 1))] = mine;
    n_out += __popc(m);
  }
  __syncwarp();
  return n_out;
}

__global__ void __launch_bounds__(32 * kRowWarps)
rr_krecip_kernel(const int32_t* __restrict__ rank, int ldr, const float* __restrict__ orig, int N, int k1p, int khp, int cap_v,
                 int cap_pow2, int32_t* __restrict__ v_col, float* __restrict__ v_val, int32_t* __restrict__ v_cnt) {
  extern __shared__ __align__(16) uint8_t kr_raw[];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  int* base_set = reinterpret_cast<int*>(kr_raw) + w * (k1p + khp + cap_pow2);   // R(i), up to k1p
  int* cand_set = base_set + k1p;                                                 // R_half(c), up to khp
  int* members = cand_set + khp;                                                  // expansion list, cap_pow2
  const int i = blockIdx.x * kRowWarps + w;
  if (i >= N) return;
  // R(i)
  const int nb = reciprocal_append(rank, ldr, i, k1p, lane, base_set, 0);
  for (int t = lane; t < nb; t += 32) members[t] = base_set[t];
  int nm = nb;
  __syncwarp();
  // expansion (rerank.py:61-78): candidates in R(i) order; the overlap test is against the UN-expanded R(i)
  for (int j = 0; j < nb; ++j) {
    const int c = base_set[j];
    const int nc = reciprocal_append(rank, ldr, c, khp, lane, cand_set, 0);
    int inter = 0;
    for (int t = lane; t < nc; t += 32) {
      const int x = cand_set[t];
      bool hit = false;
      for (int u = 0; u < nb; ++u) hit |= (base_set[u] == x);
      inter += hit;
    }
    for (int o = 16; o > 0; o >>= 1) inter += __shfl_xor_sync(0xffffffffu, inter, o);
    if (3 * inter > 2 * nc) {          // len(intersect) > 2/3 * len(candidate set), strict
      for (int t = lane; t < nc; t += 32) members[nm + t] = cand_set[t];
      nm += nc;
    }
    __syncwarp();
  }
  // np.unique: sort ascending, drop duplicates (warp bitonic sort in shared memory)
  for (int t = nm + lane; t < cap_pow2; t += 32) members[t] = INT32_MAX;
  __syncwarp();
  for (int k = 2; k <= cap_pow2; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = lane; t < cap_pow2; t += 32) {
        const int p = t ^ j;
        if (p > t) {
          const int a = members[t], b = members[p];
          if ((a > b) == ((t & k) == 0)) { members[t] = b; members[p] = a; }
        }
      }
      __syncwarp();
    }
  }
  int nu = 0;
  for (int t0 = 0; t0 < nm; t0 += 32) {
    const int t = t0 + lane;
    const bool keep = t < nm && (t == 0 || members[t] != members[t - 1]);
    const int val = t < nm ? members[t] : 0;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    __syncwarp();
    if (keep) v_col[(int64_t)i * cap_v + nu + __popc(m & ((1u << lane) - 1))] = val;
    nu += __popc(m);
  }
  __syncwarp();
  // weights (rerank.py:81-82): exp(-orig[i, members]) / sum, summed in ascending member order
  int32_t* cols = v_col + (int64_t)i * cap_v;
  float* vals = v_val + (int64_t)i * cap_v;
  for (int t = lane; t < nu; t += 32) vals[t] = expf(-orig[(int64_t)i * N + cols[t]]);
  __syncwarp();
  float total = 0.f;
  if (lane == 0)
    for (int t = 0; t < nu; ++t) total = __fadd_rn(total, vals[t]);
  total = __shfl_sync(0xffffffffu, total, 0);
  for (int t = lane; t < nu; t += 32) vals[t] = __fdiv_rn(vals[t], total);
  if (lane == 0) v_cnt[i] = nu;
}

// ---- K4: V_qe[i] = mean_j V[rank[i][j]], j < k2; one CTA per row ------------------------------------------
__device__ __forceinline__ void block_sort_u64(uint64_t* s, int n) {
  for (int k = 2; k <= n; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const int p = i ^ j;
        if (p > i) {
          const uint64_t a = s[i], b = s[p];
          if ((a > b) == ((i & k) == 0)) { s[i] = b; s[p] = a; }
        }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(256)
rr_expand_kernel(const int32_t* __restrict__ rank, int ldr, int N, int k2, int cap_v, const int32_t* __restrict__ v_col,
                 const float* __restrict__ v_val, const int32_t* __restrict__ v_cnt, int cap_e, int cap_e_pow2,
                 int32_t* __restrict__ e_col, float* __restrict__ e_val, int32_t* __restrict__ e_cnt) {
  extern __shared__ __align__(16) uint8_t ex_raw[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(ex_raw);                  // (col * k2 + j) << 32 | float bits
  __shared__ int n_items, n_out;
  const int i = blockIdx.x;
  if (threadIdx.x == 0) { n_items = 0; n_out = 0; }
  __syncthreads();
  for (int j = 0; j < k2; ++j) {
    const int r = rank[(int64_t)i * ldr + j];
    if (r < 0) continue;
    const int n = v_cnt[r];
    __shared__ int base;
    if (threadIdx.x == 0) { base = n_items; n_items += n; }
    __syncthreads();
    for (int t = threadIdx.x; t < n; t += blockDim.x) {
      const uint32_t c = (uint32_t)v_col[(int64_t)r * cap_v + t];
      keys[base + t] = (uint64_t(c * (uint32_t)k2 + (uint32_t)j) << 32) | __float_as_uint(v_val[(int64_t)r * cap_v + t]);
    }
    __syncthreads();
  }
  const int n = n_items;
  for (int t = n + threadIdx.x; t < cap_e_pow2; t += blockDim.x) keys[t] = ~uint64_t(0);
  block_sort_u64(keys, cap_e_pow2);
  // segment heads: first entry of each column; each head sums its column's entries in j order (np.add.reduce over
  // the k2 rows, rerank.py:87) and divides by k2 (np.mean)
  for (int t0 = 0; t0 < n; t0 += blockDim.x) {
    const int t = t0 + threadIdx.x;
    bool head = false;
    uint32_t col = 0;
    float sum = 0.f;
    if (t < n) {
      col = (uint32_t)(keys[t] >> 32) / (uint32_t)k2;
      head = (t == 0) || ((uint32_t)(keys[t - 1] >> 32) / (uint32_t)k2 != col);
      if (head) {
        for (int u = t; u < n && (uint32_t)(keys[u] >> 32) / (uint32_t)k2 == col; ++u)
          sum = __fadd_rn(sum, __uint_as_float((uint32_t)keys[u]));
        sum = __fdiv_rn(sum, (float)k2);
      }
    }
    // ordered compaction of the heads
    const unsigned m = __ballot_sync(0xffffffffu, head);
    __shared__ int warp_cnt[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_cnt[w] = __popc(m);
    __syncthreads();
    int off = n_out;
    for (int x = 0; x < w; ++x) off += warp_cnt[x];
    if (head) {
      const int pos = off + __popc(m & ((1u << lane) - 1));
      e_col[(int64_t)i * cap_e + pos] = (int32_t)col;
      e_val[(int64_t)i * cap_e + pos] = sum;
    }
    __syncthreads();
    if (threadIdx.x == 0) { int tot = 0; for (int x = 0; x < 8; ++x) tot += warp_cnt[x]; n_out += tot; }
    __syncthreads();
  }
  if (threadIdx.x == 0) e_cnt[i] = n_out;
}

// ---- K5a: inverted index (CSC of the row-sparse matrix), zero values excluded (rerank.py:93: V[:, i] != 0) ----
__global__ void rr_col_count_kernel(int N, int cap, const int32_t* __restrict__ col, const float* __restrict__ val,
                                    const int32_t* __restrict__ cnt, int32_t* __restrict__ col_count) {
  const int r = blockIdx.x;
  const int n = cnt[r];
  for (int t = threadIdx.x; t < n; t += blockDim.x)
    if (val[(int64_t)r * cap + t] != 0.f) atomicAdd(col_count + col[(int64_t)r * cap + t], 1);
}
// exclusive scan of col_count[N] -> col_ptr[N + 1]; single CTA
__global__ void __launch_bounds__(1024) rr_scan_kernel(const int32_t* __restrict__ in, int N, int32_t* __restrict__ out) {
  __shared__ int wsum[32];
  __shared__ int carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  for (int base = 0; base < N; base += 1024) {
    const int i = base + threadIdx.x;
    const int v = i < N ? in[i] : 0;
    int incl = v;
    for (int o = 1; o < 32; o <<= 1) { int n = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += n; }
    if (lane == 31) wsum[w] = incl;
    __syncthreads();
    int off = carry;
    for (int x = 0; x < w; ++x) off += wsum[x];
    if (i < N) out[i] = off + incl - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry = off + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[N] = carry;
}
__global__ void rr_col_fill_kernel(int N, int cap, const int32_t* __restrict__ col, const float* __restrict__ val,
                                   const int32_t* __restrict__ cnt, const int32_t* __restrict__ col_ptr,
                                   int32_t* __restrict__ cursor, int32_t* __restrict__ inv_row, float* __restrict__ inv_val) {
  const int r = blockIdx.x;
  const int n = cnt[r];
  for (int t = threadIdx.x; t < n; t += blockDim.x) {
    const float v = val[(int64_t)r * cap + t];
    if (v == 0.f) continue;
    const int c = col[(int64_t)r * cap + t];
    const int pos = col_ptr[c] + atomicAdd(cursor + c, 1);
    inv_row[pos] = r;
    inv_val[pos] = v;
  }
}

// ---- K5b + K6: Jaccard accumulation for one query per CTA, then the lambda blend ----------------------------
__global__ void __launch_bounds__(256)
rr_jaccard_kernel(int Q, int G, int N, int cap, const int32_t* __restrict__ col, const float* __restrict__ val,
                  const int32_t* __restrict__ cnt, const int32_t* __restrict__ col_ptr, const int32_t* __restrict__ inv_row,
                  const float* __restrict__ inv_val, const float* __restrict__ orig, float w_jaccard, float w_orig,
                  float* __restrict__ out, int64_t ldo) {
  const int i = blockIdx.x;
  float* acc = out + (int64_t)i * ldo;     // the output row doubles as the accumulator (columns >= Q only are kept)
  for (int g = threadIdx.x; g < G; g += blockDim.x) acc[g] = 0.f;
  __syncthreads();
  const int n = cnt[i];
  for (int t = 0; t < n; ++t) {            // ascending column order (rerank.py:99-105)
    const float vi = val[(int64_t)i * cap + t];
    if (vi == 0.f) continue;               // np.where(V[i, :] != 0)
    const int c = col[(int64_t)i * cap + t];
    const int p0 = col_ptr[c], p1 = col_ptr[c + 1];
    for (int p = p0 + threadIdx.x; p < p1; p += blockDim.x) {
      const int r = inv_row[p];
      if (r >= Q) acc[r - Q] = __fadd_rn(acc[r - Q], fminf(vi, inv_val[p]));   // rows are distinct inside a column
    }
    __syncthreads();
  }
  for (int g = threadIdx.x; g < G; g += blockDim.x) {
    const float tmin = acc[g];
    const float jac = __fsub_rn(1.0f, __fdiv_rn(tmin, __fsub_rn(2.0f, tmin)));          // rerank.py:106
    acc[g] = __fadd_rn(__fmul_rn(jac, w_jaccard), __fmul_rn(orig[(int64_t)i * N + Q + g], w_orig));   // :108
  }
}

// ---- host ----------------------------------------------------------------------------------------------------
static inline int pow2_at_least(int x) { int p = 1; while (p < x) p <<= 1; return p; }
static inline int half_k(int k1) {          // int(np.around(k1 / 2.)): round half to even
  const double h = k1 / 2.0;
  double r = floor(h + 0.5);
  if (h + 0.5 == r && (static_cast<long long>(r) & 1)) r -= 1.0;
  return (int)r;
}

struct RerankPlan {
  int N, k1p, khp, cap_v, cap_v_pow2, cap_e, cap_e_pow2;
  size_t off_orig, off_colmax, off_rank, off_rank_val, off_vcol, off_vval, off_vcnt, off_ecol, off_eval, off_ecnt, off_colcnt,
      off_colptr, off_cursor, off_invrow, off_invval, total;
};
static RerankPlan make_plan(int64_t Q, int64_t G, int k1, int k2) {
  RerankPlan p;
  p.N = (int)(Q + G);
  p.k1p = k1 + 1;
  p.khp = half_k(k1) + 1;
  p.cap_v = p.k1p * (p.khp + 1);
  p.cap_v_pow2 = pow2_at_least(p.cap_v);
  p.cap_e = (k2 > 1 ? k2 : 1) * p.cap_v;
  p.cap_e_pow2 = pow2_at_least(p.cap_e);
  size_t o = 0;
  auto take = [&](size_t bytes) { size_t r = o; o += align256(bytes); return r; };
  const size_t N = p.N;
  p.off_orig = take(N * N * 4);
  p.off_colmax = take(N * 4);
  p.off_rank = take(N * p.k1p * 4);
  p.off_rank_val = take(N * p.k1p * 4);
  p.off_vcol = take(N * p.cap_v * 4);
  p.off_vval = take(N * p.cap_v * 4);
  p.off_vcnt = take(N * 4);
  p.off_ecol = take(k2 > 1 ? N * p.cap_e * 4 : 0);
  p.off_eval = take(k2 > 1 ? N * p.cap_e * 4 : 0);
  p.off_ecnt = take(k2 > 1 ? N * 4 : 0);
  p.off_colcnt = take(N * 4);
  p.off_colptr = take((N + 1) * 4);
  p.off_cursor = take(N * 4);
  p.off_invrow = take(N * p.cap_e * 4);
  p.off_invval = take(N * p.cap_e * 4);
  p.total = o + 256;
  return p;
}

size_t rerank_workspace_bytes(int64_t Q, int64_t G, int32_t k1, int32_t k2) {
  if (Q <= 0 || G <= 0 || k1 < 1 || k2 < 1) return 0;
  return make_plan(Q, G, k1, k2).total;
}

int rerank(const float* q_g, int64_t ld_qg, const float* q_q, int64_t ld_qq, const float* g_g, int64_t ld_gg, int64_t Q,
           int64_t G, int32_t k1, int32_t k2, double lambda_value, float* out, int64_t ldo, void* workspace,
           size_t workspace_bytes, cudaStream_t stream) {
  IEEE_REQUIRE(q_g && q_q && g_g && out && workspace, "rerank: null pointer");
  IEEE_REQUIRE(Q > 0 && G > 0 && ld_qg >= G && ld_qq >= Q && ld_gg >= G && ldo >= G, "rerank: bad shape");
  IEEE_REQUIRE(Q + G < (int64_t(1) << 24), "rerank: N = Q + G too large");
  IEEE_REQUIRE(k1 >= 1 && k1 + 1 <= 512 && k2 >= 1 && k2 <= k1 + 1, "rerank: need 1 <= k2 <= k1 + 1 <= 512");
  IEEE_REQUIRE(k1 + 1 <= Q + G, "rerank: k1 + 1 exceeds the number of samples");
  IEEE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "rerank: workspace must be 256-byte aligned");
  const RerankPlan p = make_plan(Q, G, k1, k2);
  if (workspace_bytes < p.total) {
    set_error("rerank: workspace too small (%zu < %zu)", workspace_bytes, p.total);
    return IEEE_ERR_WORKSPACE;
  }
  IEEE_REQUIRE((size_t)p.cap_e_pow2 * 8 <= 160 * 1024, "rerank: k2 * (k1+1) * (k1/2+2) = %d entries per row exceed the shared-memory sort", p.cap_e);
  uint8_t* w = static_cast<uint8_t*>(workspace);
  const int N = p.N;
  float* orig = reinterpret_cast<float*>(w + p.off_orig);
  float* colmax = reinterpret_cast<float*>(w + p.off_colmax);
  int32_t* rank = reinterpret_cast<int32_t*>(w + p.off_rank);
  float* rank_val = reinterpret_cast<float*>(w + p.off_rank_val);
  int32_t* vcol = reinterpret_cast<int32_t*>(w + p.off_vcol);
  float* vval = reinterpret_cast<float*>(w + p.off_vval);
  int32_t* vcnt = reinterpret_cast<int32_t*>(w + p.off_vcnt);
  int32_t* ecol = reinterpret_cast<int32_t*>(w + p.off_ecol);
  float* eval = reinterpret_cast<float*>(w + p.off_eval);
  int32_t* ecnt = reinterpret_cast<int32_t*>(w + p.off_ecnt);
  int32_t* colcnt = reinterpret_cast<int32_t*>(w + p.off_colcnt);
  int32_t* colptr = reinterpret_cast<int32_t*>(w + p.off_colptr);
  int32_t* cursor = reinterpret_cast<int32_t*>(w + p.off_cursor);
  int32_t* invrow = reinterpret_cast<int32_t*>(w + p.off_invrow);
  float* invval = reinterpret_cast<float*>(w + p.off_invval);

  // K1
  IEEE_CUDA_CHECK(cudaMemsetAsync(colmax, 0, size_t(N) * 4, stream));
  int* cm_bits = reinterpret_cast<int*>(colmax);
  const int rpb = 256;
  auto colmax_of = [&](const float* m, int64_t ld, int rows, int cols, int* dst) {
    dim3 grid((cols + 255) / 256, (rows + rpb - 1) / rpb);
    rr_colmax_kernel<<<grid, 256, 0, stream>>>(m, ld, rows, cols, rpb, dst);
    count_launch();
  };
  colmax_of(q_q, ld_qq, (int)Q, (int)Q, cm_bits);
  colmax_of(q_g, ld_qg, (int)Q, (int)G, cm_bits + Q);
  colmax_of(g_g, ld_gg, (int)G, (int)G, cm_bits + Q);
  rr_rowmax_kernel<<<(unsigned)Q, 256, 0, stream>>>(q_g, ld_qg, (int)Q, (int)G, cm_bits);
  count_launch();
  RawView raw{q_g, q_q, g_g, ld_qg, ld_qq, ld_gg, (int)Q, (int)G};
  {
    dim3 grid((N + 31) / 32, (N + 31) / 32);
    rr_build_orig_kernel<<<grid, 256, 0, stream>>>(raw, N, colmax, orig);
    count_launch();
  }
  // K2: first k1+1 of each row's ascending (value, index) order
  int rc = topk(orig, N, N, N, 0, nullptr, nullptr, nullptr, nullptr, p.k1p, rank, rank_val, stream);
  if (rc) return rc;
  // K3
  {
    const size_t smem = size_t(kRowWarps) * (p.k1p + p.khp + p.cap_v_pow2) * 4;
    IEEE_ENSURE_DYN_SMEM(rr_krecip_kernel, smem);
    rr_krecip_kernel<<<(N + kRowWarps - 1) / kRowWarps, 32 * kRowWarps, smem, stream>>>(rank, p.k1p, orig, N, p.k1p, p.khp,
                                                                                         p.cap_v, p.cap_v_pow2, vcol, vval, vcnt);
    count_launch();
  }
  // K4
  const int32_t* fcol = vcol;
  const float* fval = vval;
  const int32_t* fcnt = vcnt;
  int fcap = p.cap_v;
  if (k2 != 1) {
    const size_t smem = size_t(p.cap_e_pow2) * 8;
    IEEE_ENSURE_DYN_SMEM(rr_expand_kernel, smem);
    rr_expand_kernel<<<N, 256, smem, stream>>>(rank, p.k1p, N, k2, p.cap_v, vcol, vval, vcnt, p.cap_e, p.cap_e_pow2, ecol, eval, ecnt);
    count_launch();
    fcol = ecol; fval = eval; fcnt = ecnt; fcap = p.cap_e;
  }
  // K5a
  IEEE_CUDA_CHECK(cudaMemsetAsync(colcnt, 0, size_t(N) * 4, stream));
  IEEE_CUDA_CHECK(cudaMemsetAsync(cursor, 0, size_t(N) * 4, stream));
  rr_col_count_kernel<<<N, 128, 0, stream>>>(N, fcap, fcol, fval, fcnt, colcnt);
  count_launch();
  rr_scan_kernel<<<1, 1024, 0, stream>>>(colcnt, N, colptr);
  count_launch();
  rr_col_fill_kernel<<<N, 128, 0, stream>>>(N, fcap, fcol, fval, fcnt, colptr, cursor, invrow, invval);
  count_launch();
  // K5b + K6.  (1 - lambda) is formed in double and rounded to float32 once, as NumPy does with the Python float.
  rr_jaccard_kernel<<<(unsigned)Q, 256, 0, stream>>>((int)Q, (int)G, N, fcap, fcol, fval, fcnt, colptr, invrow, invval, orig,
                                                     (float)(1.0 - lambda_value), (float)lambda_value, out, ldo);
  count_launch();
  IEEE_CUDA_CHECK(cudaGetLastError());
  return IEEE_OK;
}

// =========================================================================================================
// GNN re-ranking (torchreid/utils/GPU-Re-Ranking/gnn_reranking.py:27-59 and its two CUDA extensions), as an alternative
// `rerank` mode.  Input: the negated similarity matrix -(X_u X_u^T) over queries + gallery (our contraction with the
// NEG_DOT metric).  Steps, all on N x N float32 matrices (N = Q + G; 1.5 GB at Market scale, nothing on a 180 GB part):
//   top-k1 neighbours per row (ieee_topk on the negated scores: largest similarity first, ties by index)  :36-38
//   A[i, rank[i, j]] = 1                              build_adjacency_matrix_kernel.cu:10-17
//   S = S * S                                          :42
//   twice, if k2 != 1:  A = A + A^T;  A[i, :] = sum_{j < k2} S[i, j] * A[rank[i, j], :];  A[i, :] /= |A[i, :]|_2
//                                                      :46-53, gnn_propagate_kernel.cu:8-22
// The caller finishes with cosine = A[:Q] A[Q:]^T (:55) -- again our contraction.
// =========================================================================================================
__global__ void gnn_square_kernel(const float* __restrict__ val, float* __restrict__ S, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) S[i] = val[i] * val[i];                 // (-s)^2 == s^2: the negation of the scores drops out here
}

__global__ void gnn_adjacency_kernel(const int32_t* __restrict__ rank, int k1, int64_t N, float* __restrict__ A, int64_t ld) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= N * k1) return;
  const int64_t i = e / k1;
  const int32_t j = rank[e];
  if (j >= 0) A[i * ld + j] = 1.0f;
}

__global__ void __launch_bounds__(256) gnn_symmetrise_kernel(const float* __restrict__ A, float* __restrict__ B, int64_t N, int64_t ld) {
  __shared__ float t[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int r = ty; r < 32; r += 8) {
    const int64_t row = bx + r, col = by + tx;             // the transposed tile
    t[r][tx] = (row < N && col < N) ? A[row * ld + col] : 0.f;
  }
  __syncthreads();
  for (int r = ty; r < 32; r += 8) {
    const int64_t row = by + r, col = bx + tx;
    if (row < N && col < N) B[row * ld + col] = A[row * ld + col] + t[tx][r];
  }
}

// out[i, f] = sum_{j < k2} S[i, j] * A[rank[i, j], f], summed in j order like the reference kernel
__global__ void __launch_bounds__(256) gnn_propagate_kernel(const float* __restrict__ A, int64_t ld, const int32_t* __restrict__ rank,
                                                             const float* __restrict__ S, int k1, int k2, int64_t N,
                                                             float* __restrict__ out) {
  const int64_t i = blockIdx.y;
  const int64_t f = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (f >= N) return;
  float sum = 0.f;
  for (int j = 0; j < k2; ++j) {
    const int32_t nb = rank[i * k1 + j];
    if (nb >= 0) sum += A[(int64_t)nb * ld + f] * S[i * k1 + j];
  }
  out[i * ld + f] = sum;
}

__global__ void __launch_bounds__(256) gnn_normalise_kernel(float* __restrict__ A, int64_t ld, int64_t N) {
  __shared__ float red[8];
  __shared__ float inv_s;
  float* row = A + (int64_t)blockIdx.x * ld;
  float s = 0.f;
  for (int64_t f = threadIdx.x; f < N; f += 256) s = __fmaf_rn(row[f], row[f], s);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    inv_s = __fsqrt_rn(t);                               // torch.norm(A, p=2, dim=1): no epsilon (gnn_reranking.py:52-53)
  }
  __syncthreads();
  const float nrm = inv_s;
  for (int64_t f = threadIdx.x; f < N; f += 256) row[f] = __fdiv_rn(row[f], nrm);
}

size_t gnn_rerank_workspace_bytes(int64_t N, int32_t k1) {
  if (N <= 0 || k1 <= 0) return 0;
  const int64_t ld = round_up(N, 32);
  return 3 * align256(size_t(N) * k1 * 4) + align256(size_t(N) * ld * 4) + 256;
}

int gnn_rerank(const float* neg_score, int64_t lds, int64_t N, int32_t k1, int32_t k2, float* A, int64_t ldA, void* workspace,
               size_t workspace_bytes, cudaStream_t stream) {
  IEEE_REQUIRE(neg_score && A && workspace, "gnn_rerank: null pointer");
  IEEE_REQUIRE(N > 0 && lds >= N && ldA >= N && k1 >= 1 && k1 <= N && k1 <= 1024 && k2 >= 1 && k2 <= k1,
               "gnn_rerank: bad arguments N=%lld k1=%d k2=%d", (long long)N, k1, k2);
  IEEE_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "gnn_rerank: workspace must be 256-byte aligned");
  if (workspace_bytes < gnn_rerank_workspace_bytes(N, k1)) {
    set_error("gnn_rerank: workspace too small (%zu < %zu)", workspace_bytes, gnn_rerank_workspace_bytes(N, k1));
    return IEEE_ERR_WORKSPACE;
  }
  const int64_t ld = round_up(N, 32);
  IEEE_REQUIRE(ldA == ld, "gnn_rerank: the adjacency matrix must have a row pitch of %lld floats", (long long)ld);
  uint8_t* w = static_cast<uint8_t*>(workspace);
  int32_t* rank = reinterpret_cast<int32_t*>(w);
  w += align256(size_t(N) * k1 * 4);
  float* val = reinterpret_cast<float*>(w);
  w += align256(size_t(N) * k1 * 4);
  float* S = reinterpret_cast<float*>(w);
  w += align256(size_t(N) * k1 * 4);
  float* B = reinterpret_cast<float*>(w);
  int rc = topk(neg_score, lds, N, N, 0, nullptr, nullptr, nullptr, nullptr, k1, rank, val, stream);
  if (rc) return rc;
  const int64_t nk = N * k1;
  gnn_square_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, stream>>>(val, S, nk);
  IEEE_CUDA_CHECK(cudaMemsetAsync(A, 0, size_t(N) * ld * 4, stream));
  gnn_adjacency_kernel<<<(unsigned)((nk + 255) / 256), 256, 0, stream>>>(rank, k1, N, A, ld);
  count_launch(2);
  if (k2 != 1) {
    const dim3 tgrid((unsigned)((N + 31) / 32), (unsigned)((N + 31) / 32)), pgrid((unsigned)((N + 255) / 256), (unsigned)N);
    for (int it = 0; it < 2; ++it) {
      gnn_symmetrise_kernel<<<tgrid, 256, 0, stream>>>(A, B, N, ld);
      gnn_propagate_kernel<<<pgrid, 256, 0, stream>>>(B, ld, rank, S, k1, k2, N, A);
      gnn_normalise_kernel<<<(unsigned)N, 256, 0, stream>>>(A, ld, N);
      count_launch(3);
    }
  }
  IEEE_CUDA_CHECK(cudaGetLastError());
  return IEEE_OK;
}

}  // namespace ieee
