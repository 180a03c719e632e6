// Merge P sorted partial top-K lists per image row into the final sorted top-K, map bank rows
// to node ids and count Hit@{1,2,5,10,20}.
//
// Reference: the tail of the eval reductions in main.test -- `model.test_index[pred]`,
// `pred.eq(targets)`, `correct[:k].sum()` (main.py:139-147) -- and, for the class-sharded
// multi-GPU head, the combination of the per-rank lists after the NCCL all-gather.
//
// One warp per row.  Lane l owns lists l, l+32, ...; K rounds of (lane-local best head,
// warp arg-max by (value desc, list asc), winner advances its head).  Lists are sorted by
// (value desc, position asc) and lists of one row cover ascending bank-row ranges, so the
// result is ordered by (value desc, bank row asc) -- deterministic.
#include "common.cuh"
#include "sched.cuh"

namespace hgr {
namespace {

constexpr int kMergeWarps = 4;
constexpr int kMaxListsPerLane = 4;  // P <= 128

struct MergeSched {
  int use;
  Sched s;
};

__global__ void __launch_bounds__(kMergeWarps * 32)
topk_merge_kernel(const float* __restrict__ part_val, const int32_t* __restrict__ part_idx, int P, int64_t B,
                  int K, int64_t pstride, MergeSched ms, const int32_t* __restrict__ col_id, int32_t id_base, float scale,
                  const int32_t* __restrict__ targets, float* __restrict__ topk_val,
                  int32_t* __restrict__ topk_idx, unsigned long long* __restrict__ hits) {
  __shared__ int s_hits[HGR_NUM_HITS];
  if (threadIdx.x < HGR_NUM_HITS) s_hits[threadIdx.x] = 0;
  __syncthreads();

  const int lane = threadIdx.x & 31;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kMergeWarps + (threadIdx.x >> 5);
  if (row < B) {
    int cnt = P;
    if (ms.use) cnt = ms.s.parts(static_cast<int32_t>(row / kTileM));
    const float* pv = part_val + row * K;
    const int32_t* pi = part_idx + row * K;

    int head[kMaxListsPerLane];
    float hv[kMaxListsPerLane];
#pragma unroll
    for (int q = 0; q < kMaxListsPerLane; ++q) {
      const int p = lane + 32 * q;
      head[q] = 0;
      hv[q] = (p < cnt && K > 0) ? pv[p * pstride] : -INFINITY;
      if (p < cnt && pi[p * pstride] < 0) hv[q] = -INFINITY;  // empty list
    }
    const int32_t target = targets ? targets[row] : -1;
    int hit_pos = 1 << 30;

    for (int r = 0; r < K; ++r) {
      // lane-local best head; strict > keeps the lower list index on ties
      float bv = hv[0];
      int bp = lane;
#pragma unroll
      for (int q = 1; q < kMaxListsPerLane; ++q) {
        if (hv[q] > bv) {
          bv = hv[q];
          bp = lane + 32 * q;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int op = __shfl_xor_sync(0xffffffffu, bp, o);
        if (ov > bv || (ov == bv && op < bp)) {
          bv = ov;
          bp = op;
        }
      }
      // every lane now agrees on (bv, bp)
      int32_t gi = -1;
      if (bv > -INFINITY) {
        const int owner_lane = bp & 31;
        int32_t li = 0;
        if (lane == owner_lane) {
#pragma unroll
          for (int q = 0; q < kMaxListsPerLane; ++q) {
            if (bp == lane + 32 * q) {
              li = pi[bp * pstride + head[q]];
              const int h = ++head[q];
              float nv = -INFINITY;
              if (h < K) {
                nv = pv[bp * pstride + h];
                if (pi[bp * pstride + h] < 0) nv = -INFINITY;
              }
              hv[q] = nv;
            }
          }
        }
        li = __shfl_sync(0xffffffffu, li, owner_lane);
        gi = col_id ? col_id[li] : id_base + li;
      }
      if (lane == 0) {
        topk_val[row * K + r] = bv > -INFINITY ? bv * scale : -INFINITY;
        topk_idx[row * K + r] = gi;
      }
      if (gi >= 0 && gi == target && hit_pos > r) hit_pos = r;
    }
    if (hits && lane == 0 && target >= 0) {
#pragma unroll
      for (int c = 0; c < HGR_NUM_HITS; ++c)
        if (hit_pos < hit_cut(c)) atomicAdd(&s_hits[c], 1);
    }
  }
  __syncthreads();
  if (hits && threadIdx.x < HGR_NUM_HITS && s_hits[threadIdx.x] != 0)
    atomicAdd(&hits[threadIdx.x], static_cast<unsigned long long>(s_hits[threadIdx.x]));
}

}  // namespace

int launch_topk_merge(const float* part_val, const int32_t* part_idx, int64_t P, int64_t B, int K,
                      int64_t part_stride, const Sched* sched, const int32_t* col_id, int32_t id_base, float scale,
                      const int32_t* targets, float* topk_val, int32_t* topk_idx, int64_t* hits,
                      cudaStream_t stream) {
  if (B == 0 || K == 0) return HGR_OK;
  if (P > 32 * kMaxListsPerLane)
    return set_error(HGR_ERR_UNSUPPORTED, "topk merge: P = %lld lists per row exceeds %d", (long long)P,
                     32 * kMaxListsPerLane);
  MergeSched ms;
  ms.use = sched != nullptr;
  if (sched) ms.s = *sched;
  else ms.s = Sched{1, 1, 1, 1, 1};
  const int blocks = static_cast<int>((B + kMergeWarps - 1) / kMergeWarps);
  topk_merge_kernel<<<blocks, kMergeWarps * 32, 0, stream>>>(
      part_val, part_idx, static_cast<int>(P), B, K, part_stride > 0 ? part_stride : B * K, ms, col_id, id_base, scale, targets, topk_val, topk_idx,
      reinterpret_cast<unsigned long long*>(hits));
  HGR_CHECK_LAUNCH();
  return HGR_OK;
}

}  // namespace hgr
