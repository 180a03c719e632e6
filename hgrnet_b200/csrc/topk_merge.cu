// Merge the partial top-K lists of one image row into the final sorted top-K, map bank rows to
// node ids and count Hit@{1,2,5,10,20}.
//
// Reference: the tail of the eval reductions in main.test -- `model.test_index[pred]`,
// `pred.eq(targets)`, `correct[:k].sum()` (main.py:139-147) -- and, for the class-sharded
// multi-GPU head, the combination of the per-rank lists after the NCCL all-gather.
//
// One warp per row, kMergeWarps rows per CTA.  The CTA first stages all lists of its rows in
// shared memory (coalesced), then each warp runs K rounds of a P-way merge: lane l owns lists
// l, l+32, ...; per round the lane-local best head, a warp arg-max by (value desc, list asc),
// and the winner advances its head (shared-memory read).  Heads compare by (value desc, item
// asc), every list is itself sorted that way, so the result is ordered (value desc, bank row asc)
// -- deterministic whatever way the producers interleave bank rows over lists.
//
// Speculative lists (KL < K, tcgen05 path): the row is CERTIFIED when every full list ends strictly
// below the merged K-th value -- anything such a list dropped is <= its last entry, hence cannot
// belong to (or tie with) the top-K.  An uncertified row is re-scanned exactly by its warp on the CUDA
// cores (simt_row.cuh).  With randomly ordered bank rows this is a ~1e-8-per-row event.
#include "common.cuh"
#include "simt_row.cuh"

namespace hgr {
namespace {

constexpr int kMergeWarps = 4;

// monotone map fp32 -> uint32 (larger float <=> larger key) and back
__device__ __forceinline__ uint32_t f32_order_key(float v) {
  const uint32_t b = __float_as_uint(v);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float f32_from_order_key(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// kMaxListsPerLane: 4 covers P <= 128 lists per row, 10 covers P <= 320 (one list per epilogue warp of
// every CTA when a single row tile is spread over all 148 SMs)
template <int kMaxListsPerLane>
__global__ void __launch_bounds__(kMergeWarps * 32)
topk_merge_kernel(const MergeArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ int s_hits[HGR_NUM_HITS];
  if (threadIdx.x < HGR_NUM_HITS) s_hits[threadIdx.x] = 0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KL = a.KL, K = a.K;
  const int64_t row0 = static_cast<int64_t>(blockIdx.x) * kMergeWarps;
  const int slots = static_cast<int>(a.P);                   // list slots per row in shared memory
  float* s_val = reinterpret_cast<float*>(smem_raw);         // [kMergeWarps][slots][KL]
  int32_t* s_idx = reinterpret_cast<int32_t*>(s_val + kMergeWarps * slots * KL);
  const int64_t pstride = a.part_stride > 0 ? a.part_stride : a.B * KL;

  // ---- stage: [list][row][KL] global -> [row][list][KL] shared
  for (int r = 0; r < kMergeWarps; ++r) {
    const int64_t row = row0 + r;
    if (row >= a.B) break;
    int cnt = slots;
    if (a.use_sched) cnt = a.sched.parts(static_cast<int32_t>(row / a.sched.rows)) * a.wpq;
    const int n = cnt * KL;
    for (int e = threadIdx.x; e < n; e += kMergeWarps * 32) {
      const int p = e / KL, k = e - p * KL;
      const int64_t g = p * pstride + row * KL + k;
      s_val[(r * slots + p) * KL + k] = a.part_val[g];
      s_idx[(r * slots + p) * KL + k] = a.part_idx[g];
    }
  }
  __syncthreads();

  const int64_t row = row0 + warp;
  if (row < a.B) {
    int cnt = slots;
    if (a.use_sched) cnt = a.sched.parts(static_cast<int32_t>(row / a.sched.rows)) * a.wpq;
    const float* lv = s_val + warp * slots * KL;
    const int32_t* li = s_idx + warp * slots * KL;

    int head[kMaxListsPerLane];
    float hv[kMaxListsPerLane];   // value of the current head of my q-th list (-inf: exhausted)
    int32_t hi[kMaxListsPerLane]; // its item (bank row / node id)
#pragma unroll
    for (int q = 0; q < kMaxListsPerLane; ++q) {
      const int p = lane + 32 * q;
      head[q] = 0;
      hi[q] = p < cnt ? li[p * KL] : -1;
      hv[q] = hi[q] >= 0 ? lv[p * KL] : -INFINITY;
    }
    float my_v = -INFINITY;  // lane r keeps rank r of the result
    int32_t my_i = -1;
    float kth = -INFINITY;
    for (int r = 0; r < K; ++r) {
      // lane-local best head, then warp arg-max; order: value desc, item asc
      float bv = hv[0];
      int32_t bi = hi[0];
#pragma unroll
      for (int q = 1; q < kMaxListsPerLane; ++q) {
        if (hv[q] > bv || (hv[q] == bv && hi[q] < bi && hi[q] >= 0)) {
          bv = hv[q];
          bi = hi[q];
        }
      }
      const float lbv = bv;
      const int32_t lbi = bi;
      // warp arg-max with two REDUX instructions: max of an order-preserving integer image of the value, then
      // the smallest item among the lanes that hold that value
      const uint32_t lkey = f32_order_key(lbv);
      const uint32_t mkey = __reduce_max_sync(0xffffffffu, lkey);
      const uint32_t mitem =
          __reduce_min_sync(0xffffffffu, (lkey == mkey && lbi >= 0) ? static_cast<uint32_t>(lbi) : 0xFFFFFFFFu);
      if (mitem == 0xFFFFFFFFu) {  // every list exhausted: fewer than K candidates, the rest stays (-inf, -1)
        kth = -INFINITY;
        break;
      }
      bv = f32_from_order_key(mkey);
      bi = static_cast<int32_t>(mitem);
      kth = bv;
      // the lowest lane holding the winner advances that list
      const unsigned holders = __ballot_sync(0xffffffffu, lbv == bv && lbi == bi);
      if (lane == __ffs(holders) - 1) {
        bool done = false;
#pragma unroll
        for (int q = 0; q < kMaxListsPerLane; ++q) {
          if (!done && hv[q] == bv && hi[q] == bi) {
            const int p = lane + 32 * q;
            const int h = ++head[q];
            hi[q] = h < KL ? li[p * KL + h] : -1;
            hv[q] = hi[q] >= 0 ? lv[p * KL + h] : -INFINITY;
            done = true;
          }
        }
      }
      if (lane == r) {
        my_v = bv;
        my_i = bi;
      }
    }

    if (KL < K) {
      // certificate for speculative (narrow) lists
      bool doubt = false;
#pragma unroll
      for (int q = 0; q < kMaxListsPerLane; ++q) {
        const int p = lane + 32 * q;
        if (p < cnt && li[p * KL + KL - 1] >= 0 && lv[p * KL + KL - 1] >= kth) doubt = true;
      }
      if (__any_sync(0xffffffffu, doubt)) {
        if (lane == 0 && a.rescan_count) atomicAdd(a.rescan_count, 1u);
        SortedList<HGR_TOPK_MAX> full;
        full.init();
        scan_row_range<HGR_TOPK_MAX>(reinterpret_cast<const uint4*>(a.X) + row * a.D8,
                                     reinterpret_cast<const uint4*>(a.bank), 0, a.C, a.D8, lane, full);
        my_v = -INFINITY;
        my_i = -1;
#pragma unroll
        for (int k = 0; k < HGR_TOPK_MAX; ++k) {
          if (lane == k) {
            my_v = full.v[k];
            my_i = full.i[k];
          }
        }
      }
    }

    int32_t gid = -1;
    if (lane < K && my_i >= 0) gid = a.col_id ? a.col_id[my_i] : a.id_base + my_i;
    if (lane < K) {
      a.topk_val[row * K + lane] = my_i >= 0 ? my_v * a.scale : -INFINITY;
      a.topk_idx[row * K + lane] = gid;
    }
    if (a.hits && a.targets) {
      const int32_t target = a.targets[row];
      const unsigned m = __ballot_sync(0xffffffffu, lane < K && gid >= 0 && gid == target);
      if (lane == 0 && m != 0 && target >= 0) {
        const int pos = __ffs(m) - 1;
#pragma unroll
        for (int c = 0; c < HGR_NUM_HITS; ++c)
          if (pos < hit_cut(c)) atomicAdd(&s_hits[c], 1);
      }
    }
  }
  __syncthreads();
  if (a.hits && threadIdx.x < HGR_NUM_HITS && s_hits[threadIdx.x] != 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(a.hits) + threadIdx.x,
              static_cast<unsigned long long>(s_hits[threadIdx.x]));
}

}  // namespace

int launch_topk_merge(const MergeArgs& args, cudaStream_t stream) {
  if (args.B == 0 || args.K == 0) return HGR_OK;
  if (args.P > 320)
    return set_error(HGR_ERR_UNSUPPORTED, "topk merge: %lld lists per row exceed 320", (long long)args.P);
  if (args.K > HGR_TOPK_MAX || args.KL < 1)
    return set_error(HGR_ERR_UNSUPPORTED, "topk merge: K = %d / KL = %d unsupported", args.K, args.KL);
  if (args.KL < args.K && (args.X == nullptr || args.bank == nullptr))
    return set_error(HGR_ERR_BAD_ARG, "topk merge: speculative lists need X / bank for the exact re-scan");
  const size_t smem = static_cast<size_t>(kMergeWarps) * args.P * args.KL * 8;
  if (smem > 200 * 1024) return set_error(HGR_ERR_UNSUPPORTED, "topk merge: %zu bytes of lists per CTA", smem);
  const int blocks = static_cast<int>((args.B + kMergeWarps - 1) / kMergeWarps);
  if (args.P <= 128) {
    if (smem > 48 * 1024)
      HGR_CHECK_CUDA(cudaFuncSetAttribute(topk_merge_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem)));
    topk_merge_kernel<4><<<blocks, kMergeWarps * 32, smem, stream>>>(args);
  } else {
    if (smem > 48 * 1024)
      HGR_CHECK_CUDA(cudaFuncSetAttribute(topk_merge_kernel<10>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(smem)));
    topk_merge_kernel<10><<<blocks, kMergeWarps * 32, smem, stream>>>(args);
  }
  HGR_CHECK_LAUNCH();
  return HGR_OK;
}

}  // namespace hgr
