// k-reciprocal re-ranking (torchreid/utils/rerank.py:31-113) -- kernels land in a follow-up commit.
#include "common.cuh"

namespace ieee {

size_t rerank_workspace_bytes(int64_t, int64_t, int32_t, int32_t) { return 0; }

int rerank(const float*, int64_t, const float*, int64_t, const float*, int64_t, int64_t, int64_t, int32_t, int32_t, float,
           float*, int64_t, void*, size_t, cudaStream_t) {
  set_error("ieee_rerank: not built yet");
  return IEEE_ERR_INVALID;
}

}  // namespace ieee
