// Shared device helpers for the fused score + mask + top-K kernels.
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

namespace lgx {

constexpr float kMaskValue = -1024.0f;  // PT/Procedure.py:134: rating[exclude] = -(1<<10)

// Train-interaction lookup: is `item` (global id) in user u's row of the bipartite CSR?
// User rows hold columns n_users + item, ascending (the same CSR that drives propagation), so the
// reference's Python exclude lists (PT/Procedure.py:129-133) become a binary search.
struct TrainMask {
  const int64_t* indptr;   // NULL = no mask
  const int32_t* indices;
  int32_t n_users;
  int32_t m_items_hint;    // catalogue size when the caller knows it (0 = unknown); only used to skip range searches
  __device__ __forceinline__ bool contains(int64_t user, int64_t item) const {
    if (indptr == nullptr || user < 0) return false;
    int64_t lo = indptr[user], hi = indptr[user + 1];
    const int32_t key = (int32_t)(n_users + item);
    while (lo < hi) {
      const int64_t mid = (lo + hi) >> 1;
      const int32_t c = __ldg(indices + mid);
      if (c < key) lo = mid + 1; else hi = mid;
    }
    return lo < indptr[user + 1] && __ldg(indices + lo) == key;
  }
};

// Total order used everywhere: higher score first, ties by lower item id.
__device__ __forceinline__ bool better(float v, int32_t i, float v2, int32_t i2) {
  return v > v2 || (v == v2 && i < i2);
}

// Sorted (best first) K-entry list stored with a stride (column layout in shared memory).
// Inserts (v, i) if it beats the current last entry; returns the new K-th value (the threshold).
__device__ __forceinline__ float topk_insert(float* vals, int32_t* idxs, int K, int stride, float v, int32_t i) {
  int p = K - 1;
  if (!better(v, i, vals[p * stride], idxs[p * stride])) return vals[p * stride];
  while (p > 0 && better(v, i, vals[(p - 1) * stride], idxs[(p - 1) * stride])) {
    vals[p * stride] = vals[(p - 1) * stride];
    idxs[p * stride] = idxs[(p - 1) * stride];
    --p;
  }
  vals[p * stride] = v;
  idxs[p * stride] = i;
  return vals[(K - 1) * stride];
}

}  // namespace lgx
