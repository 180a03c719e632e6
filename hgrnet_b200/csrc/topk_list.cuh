// Register-resident sorted top-K list kept by one thread for one image row.
//
// Order: value descending; a new element is inserted only when strictly greater than an
// existing one, so among equal values the element seen first (the lower bank row, because
// every producer streams bank rows in ascending order) stays ahead.  This is the
// deterministic (value desc, index asc) order documented in hgr_b200.h; torch.topk's own
// tie order is unspecified (main.py:138).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace hgr {

template <int KL>
struct SortedList {
  float v[KL];
  int32_t i[KL];

  __device__ __forceinline__ void init() {
#pragma unroll
    for (int j = 0; j < KL; ++j) {
      v[j] = -INFINITY;
      i[j] = -1;
    }
  }
  __device__ __forceinline__ float thr() const { return v[KL - 1]; }

  // Insert into a list whose slots >= S are known to be empty (-inf): touches S slots only.  Used to seed the
  // list from the first columns of a segment at a fraction of the cost of full-depth inserts.
  template <int S>
  __device__ __forceinline__ void insert_prefix(float x, int32_t id) {
    static_assert(S >= 1 && S <= KL, "prefix length");
#pragma unroll
    for (int j = S - 1; j >= 1; --j) {
      const bool above = x > v[j - 1];
      const bool here = x > v[j];
      v[j] = above ? v[j - 1] : (here ? x : v[j]);
      i[j] = above ? i[j - 1] : (here ? id : i[j]);
    }
    if (x > v[0]) {
      v[0] = x;
      i[0] = id;
    }
  }

  // pre-condition: x > thr().  Fully unrolled, branch-free (predicated selects).
  __device__ __forceinline__ void insert(float x, int32_t id) {
#pragma unroll
    for (int j = KL - 1; j >= 1; --j) {
      const bool above = x > v[j - 1];  // x lands above slot j-1: slot j inherits slot j-1
      const bool here = x > v[j];       // x lands exactly at slot j
      v[j] = above ? v[j - 1] : (here ? x : v[j]);
      i[j] = above ? i[j - 1] : (here ? id : i[j]);
    }
    if (x > v[0]) {
      v[0] = x;
      i[0] = id;
    }
  }
};

}  // namespace hgr
