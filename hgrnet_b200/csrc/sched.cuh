// Static work decomposition of the fused logit + top-k head (shared by host and device).
//
// The class dimension is cut into UNITS of 16 bank rows (the N granularity of a
// tcgen05.mma with M = 128).  The (row-tile, unit) space is flattened row-tile-major and
// split into G contiguous, equally sized chunks, one per persistent CTA -- a stream-K-style
// split along N.  No fix-up reduction is needed because top-k lists merge associatively:
// every (row-tile, CTA) intersection ("segment") emits one partial top-K list per row and a
// small merge kernel combines the <= ceil(G/MT)+1 lists of a row.
#pragma once
#include <cstdint>

namespace hgr {

constexpr int kTileM = 128;   // image rows per tile = TMEM lanes
constexpr int kUnit = 16;     // bank rows per scheduling unit
constexpr int kSubN = 256;    // max bank rows per MMA sub-tile (one TMEM accumulator buffer)

struct Sched {
  int32_t G;    // persistent CTAs
  int32_t MT;   // row tiles = ceil(B / 128)
  int32_t U;    // units per row tile = ceil(C / 16)
  int32_t P;    // max partial lists per row = slots in the workspace
  int64_t T;    // MT * U

  __host__ __device__ int64_t unit_begin(int32_t cta) const { return (static_cast<int64_t>(cta) * T) / G; }
  // CTA whose chunk contains flattened unit x (chunks are non-empty because G <= T)
  __host__ __device__ int32_t owner(int64_t x) const {
    return static_cast<int32_t>(((x + 1) * G - 1) / T);
  }
  __host__ __device__ int32_t first_cta(int32_t mt) const { return owner(static_cast<int64_t>(mt) * U); }
  __host__ __device__ int32_t last_cta(int32_t mt) const { return owner(static_cast<int64_t>(mt + 1) * U - 1); }
  __host__ __device__ int32_t parts(int32_t mt) const { return last_cta(mt) - first_cta(mt) + 1; }
};

inline Sched make_sched(int64_t B, int64_t C, int num_ctas) {
  Sched s;
  s.MT = static_cast<int32_t>((B + kTileM - 1) / kTileM);
  s.U = static_cast<int32_t>((C + kUnit - 1) / kUnit);
  s.T = static_cast<int64_t>(s.MT) * s.U;
  s.G = static_cast<int32_t>(s.T < num_ctas ? s.T : num_ctas);
  if (s.G < 1) s.G = 1;
  s.P = 1;
  for (int32_t mt = 0; mt < s.MT; ++mt) {
    const int32_t p = s.parts(mt);
    if (p > s.P) s.P = p;
  }
  return s;
}

}  // namespace hgr
