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

constexpr int kDoubtWords = 10;  // bit per list, up to 320 lists per row

// Exact repair of a row whose speculative lists could not all be certified.  Only the column ranges of the
// DOUBTFUL lists are re-scanned on the CUDA cores (1 / lists-per-row of the bank each); the entries of the certified
// lists are final as they are (whatever such a list dropped is below the old K-th value, which can only rise) and are
// inserted as candidates.  Lists are visited in ascending column order and SortedList keeps the first of equal
// values, so the result has the documented (value desc, bank row asc) order.  The list is replicated in every lane.
// (Arguments by value: handing the kernel's whole MergeArgs to a non-inlined function would copy it to the stack of
// every thread at kernel entry.)
struct RepairArgs {
  const float* part_val;
  const int32_t* part_idx;
  int64_t pstride, C;
  const void* X;
  const void* bank;
  Sched sched;
  int KL, wpq, D8;
};
__device__ __noinline__ void repair_row(const RepairArgs a, int64_t row, int lane, int cnt, const uint32_t* dmask,
                                        SortedList<HGR_TOPK_MAX>& full) {
  full.init();
  const int KL = a.KL;
  const int64_t pstride = a.pstride;
  const int32_t mt = static_cast<int32_t>(row / a.sched.rows);
  const int32_t w0 = a.sched.first_cta(mt);
  const int workers = cnt / a.wpq;
  for (int wi = 0; wi < workers; ++wi) {
    bool doubtful = false;
    for (int m = 0; m < a.wpq; ++m) {
      const int p = wi * a.wpq + m;
      doubtful |= (dmask[p >> 5] >> (p & 31)) & 1u;
    }
    if (doubtful) {
      // the worker's chunk of flattened units, clipped to this row tile -> bank rows [c0, c1)
      const int64_t t0 = static_cast<int64_t>(mt) * a.sched.U;
      int64_t u0 = a.sched.unit_begin(w0 + wi) - t0, u1 = a.sched.unit_begin(w0 + wi + 1) - t0;
      u0 = u0 < 0 ? 0 : u0;
      u1 = u1 > a.sched.U ? a.sched.U : u1;
      int64_t c0 = u0 * kUnit, c1 = u1 * kUnit;
      c1 = c1 > a.C ? a.C : c1;
      if (c0 < c1)
        scan_row_range_lanes<HGR_TOPK_MAX>(reinterpret_cast<const uint4*>(a.X) + row * a.D8,
                                           reinterpret_cast<const uint4*>(a.bank), c0, c1, a.D8, lane, full);
    } else {
      for (int m = 0; m < a.wpq; ++m) {
        const int p = wi * a.wpq + m;
        for (int k = 0; k < KL; ++k) {
          const int64_t g = p * pstride + row * KL + k;
          const int32_t it = a.part_idx[g];
          const float v = a.part_val[g];
          if (it >= 0 && v > full.thr()) full.insert(v, it);
        }
      }
    }
  }
}

// Exact repair of a row of the class-sharded head whose GLOBAL certificate failed (hgr_topk_merge_certified): list p
// is shard p's final local top-K (node ids, scaled values).  A doubtful shard (bit p of `dm`: its bound reaches the
// merged K-th value) is re-scanned completely with the row's features -- over NVLink when the bank is a peer's --, its
// exact local top-K mapped to node ids and scaled like the producer did; the lists of the other shards are final as
// they are (what such a shard dropped is below the old K-th value, which can only rise).
__device__ __noinline__ void repair_row_global(const float* part_val, const int32_t* part_idx, int64_t pstride,
                                               int64_t row, int K, int P, unsigned dm, const uint4* xrow, int D8,
                                               const hgr_shard_t* shards, float scale, int lane,
                                               SortedList<HGR_TOPK_MAX>& full) {
  full.init();
  for (int p = 0; p < P; ++p) {
    if ((dm >> p) & 1u) {
      const hgr_shard_t sh = shards[p];
      SortedList<HGR_TOPK_MAX> sub;
      sub.init();
      if (sh.C > 0)
        scan_row_range_lanes<HGR_TOPK_MAX>(xrow, reinterpret_cast<const uint4*>(sh.bank), 0, sh.C, D8, lane, sub);
#pragma unroll
      for (int k = 0; k < HGR_TOPK_MAX; ++k) {
        const int32_t it = sub.i[k];
        const float v = sub.v[k] * scale;
        if (it >= 0 && v > full.thr()) full.insert(v, sh.col_id ? sh.col_id[it] : sh.id_base + it);
      }
    } else {
      for (int k = 0; k < K; ++k) {
        const int64_t g = p * pstride + row * K + k;
        const int32_t it = part_idx[g];
        const float v = part_val[g];
        if (it >= 0 && v > full.thr()) full.insert(v, it);
      }
    }
  }
}

// Everything the rare repair paths need lives in the frame of this one cold function (two 32-entry lists), so that the
// merge kernels themselves keep their register count and carry no stack.  `a` points at the kernel's __grid_constant__
// parameter.  Returns rank `lane` of the repaired row.
__device__ __noinline__ Cand repair_cold(const MergeArgs* a, int64_t row, int lane, int cnt, const uint32_t* dmask,
                                         unsigned dm_global) {
  SortedList<HGR_TOPK_MAX> full;
  if (dm_global != 0u) {
    if (lane == 0 && a->repair_count) atomicAdd(a->repair_count, 1u);
    repair_row_global(a->part_val, a->part_idx, a->part_stride > 0 ? a->part_stride : a->B * a->K, row, a->K,
                      static_cast<int>(a->P), dm_global, reinterpret_cast<const uint4*>(a->xrows) + row * a->xD8, a->xD8,
                      a->shards, a->shard_scale, lane, full);
  } else {
    if (lane == 0 && a->rescan_count) atomicAdd(a->rescan_count, 1u);
    RepairArgs ra;
    ra.part_val = a->part_val;
    ra.part_idx = a->part_idx;
    ra.pstride = a->part_stride > 0 ? a->part_stride : a->B * a->KL;
    ra.C = a->C;
    ra.X = a->X;
    ra.bank = a->bank;
    ra.sched = a->sched;
    ra.KL = a->KL;
    ra.wpq = a->wpq;
    ra.D8 = a->D8;
    repair_row(ra, row, lane, cnt, dmask, full);
  }
  Cand c;
  c.v = -INFINITY;
  c.i = -1;
#pragma unroll
  for (int k = 0; k < HGR_TOPK_MAX; ++k) {
    if (lane == k) {
      c.v = full.v[k];
      c.i = full.i[k];
    }
  }
  return c;
}

// Tail shared by both merge kernels.  Lane r holds rank r of the merged list (my_v, my_i = bank row or -1).
// `dmask` (per-warp shared memory, one bit per list): full speculative lists that end at or above the merged K-th
// value -> the row is repaired exactly (repair_row).  Then: bank row -> node id, scale, store (dense or row-block
// scatter), Hit@k.
// `tailmax`: lane-local largest last entry of a FULL narrow list (-inf: none) -- what the producer of a class shard
// reports as its bound when the certificate is left to the owner of the row (scatter.emit_bound).
__device__ __forceinline__ void finish_row(const MergeArgs& a, int64_t row, int lane, float my_v, int32_t my_i,
                                           bool doubt, const uint32_t* dmask, int cnt, int* s_hits,
                                           float tailmax = -INFINITY) {
  const int K = a.K;
  const bool emit_bound = a.scatter.n_blocks > 0 && a.scatter.emit_bound != 0;
  unsigned dm = 0u;
  if (a.part_bound != nullptr) {  // owner side of the global certificate: list p = shard p
    const float kth = __shfl_sync(0xffffffffu, my_i >= 0 ? my_v : -INFINITY, K - 1);
    const float bnd = lane < a.P ? a.part_bound[lane * a.bound_stride + row] : -INFINITY;
    dm = __ballot_sync(0xffffffffu, bnd > -INFINITY && bnd >= kth);
  }
  const bool local_doubt = a.KL < K && !emit_bound && __any_sync(0xffffffffu, doubt);
  if (dm != 0u || local_doubt) {
    __syncwarp();
    const Cand c = repair_cold(&a, row, lane, cnt, dmask, dm);
    my_v = c.v;
    my_i = c.i;
  }
  int32_t gid = -1;
  if (lane < K && my_i >= 0) gid = a.col_id ? a.col_id[my_i] : a.id_base + my_i;
  if (lane < K) {
    float* ov = a.topk_val;
    int32_t* oi = a.topk_idx;
    int64_t orow = row;
    if (a.scatter.n_blocks > 0) {  // row-block scatter (possibly into peer memory)
      const int64_t g = row / a.scatter.block_rows;
      ov = a.scatter.val[g];
      oi = a.scatter.idx[g];
      orow = row - g * a.scatter.block_rows;
    }
    ov[orow * K + lane] = my_i >= 0 ? my_v * a.scale : -INFINITY;
    oi[orow * K + lane] = gid;
  }
  if (emit_bound) {
    float b = a.KL < K ? tailmax : -INFINITY;   // K-entry lists drop nothing that could matter
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) b = fmaxf(b, __shfl_xor_sync(0xffffffffu, b, o));
    if (lane == 0) {
      const int64_t g = row / a.scatter.block_rows;
      a.scatter.bound[g][row - g * a.scatter.block_rows] = b * a.scale;
    }
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

// kMaxListsPerLane: 4 covers P <= 128 lists per row, 10 covers P <= 320 (one list per epilogue warp of
// every CTA when a single row tile is spread over all 148 SMs)
template <int kMaxListsPerLane>
__global__ void __launch_bounds__(kMergeWarps * 32, 7)  // (the cold repair functions must not set the register count)
topk_merge_kernel(const __grid_constant__ MergeArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ int s_hits[HGR_NUM_HITS];
  __shared__ uint32_t s_dmask[kMergeWarps][kDoubtWords];
  if (threadIdx.x < HGR_NUM_HITS) s_hits[threadIdx.x] = 0;
  if (threadIdx.x < kMergeWarps * kDoubtWords) (&s_dmask[0][0])[threadIdx.x] = 0;

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

    bool doubt = false;
    float tailmax = -INFINITY;
    if (KL < K) {
      // certificate for speculative (narrow) lists
#pragma unroll
      for (int q = 0; q < kMaxListsPerLane; ++q) {
        const int p = lane + 32 * q;
        if (p < cnt && li[p * KL + KL - 1] >= 0) {
          tailmax = fmaxf(tailmax, lv[p * KL + KL - 1]);
          if (lv[p * KL + KL - 1] >= kth) {
            doubt = true;
            atomicOr(&s_dmask[warp][p >> 5], 1u << (p & 31));
          }
        }
      }
    }
    finish_row(a, row, lane, my_v, my_i, doubt, s_dmask[warp], cnt, s_hits, tailmax);
  }
  __syncthreads();
  if (a.hits && threadIdx.x < HGR_NUM_HITS && s_hits[threadIdx.x] != 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(a.hits) + threadIdx.x,
              static_cast<unsigned long long>(s_hits[threadIdx.x]));
}

// ---- selection merge (few candidates per row) -----------------------------------------------------------------
// When a row has at most 32 * NPL candidates in all its lists, the K rounds of the P-way merge above (two warp
// reductions, a ballot and a dependent shared-memory read per round) are replaced by a SELECTION: every lane keeps
// NPL candidates in registers as order-preserving integer keys, the K-th largest key is found by bisection (one
// compare per candidate and one REDUX.ADD per step, ~24 steps for cosines), ties at the cut are resolved by
// ascending item, and the <= K winners are compacted and ranked by counting.  ~4x fewer instructions per row; the
// result is the same (value desc, item asc) order.
template <int NPL>
__global__ void __launch_bounds__(kMergeWarps * 32, 7)  // (the cold repair functions must not set the register count)
topk_select_kernel(const __grid_constant__ MergeArgs a) {
  __shared__ int s_hits[HGR_NUM_HITS];
  __shared__ unsigned long long s_win[kMergeWarps][32];
  __shared__ uint32_t s_dmask[kMergeWarps][kDoubtWords];
  if (threadIdx.x < HGR_NUM_HITS) s_hits[threadIdx.x] = 0;
  if (threadIdx.x < kMergeWarps * kDoubtWords) (&s_dmask[0][0])[threadIdx.x] = 0;
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KL = a.KL, K = a.K;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kMergeWarps + warp;
  const int64_t pstride = a.part_stride > 0 ? a.part_stride : a.B * KL;
  if (row < a.B) {
    int cnt = static_cast<int>(a.P);
    if (a.use_sched) cnt = a.sched.parts(static_cast<int32_t>(row / a.sched.rows)) * a.wpq;
    const int N = cnt * KL;

    uint32_t key[NPL];   // 0 = empty slot; larger key <=> larger value
    int32_t item[NPL];
    bool tail[NPL];      // last entry of a full list (certificate)
    float val[NPL];
    // all loads first, unconditionally (an out-of-range slot re-reads entry 0): 2 * NPL independent requests in
    // flight instead of a chain of dependent L2 round trips
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
      const int e = lane + 32 * s;
      const int p = e < N ? e / KL : 0, k = e < N ? e - p * KL : 0;
      const int64_t g = N > 0 ? p * pstride + row * KL + k : 0;
      item[s] = (N > 0) ? __ldg(a.part_idx + g) : -1;
      val[s] = (N > 0) ? __ldg(a.part_val + g) : 0.f;
      tail[s] = (k == KL - 1);
    }
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
      const bool ok = (lane + 32 * s < N) && item[s] >= 0;
      key[s] = ok ? f32_order_key(val[s]) : 0u;
      item[s] = ok ? item[s] : -1;
      tail[s] = ok && tail[s];
    }
    uint32_t kmax = 0, kmin = 0xFFFFFFFFu;
    int nvalid = 0;
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
      kmax = key[s] > kmax ? key[s] : kmax;
      kmin = (key[s] != 0 && key[s] < kmin) ? key[s] : kmin;
      nvalid += key[s] != 0;
    }
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    nvalid = __reduce_add_sync(0xffffffffu, nvalid);
    const int ksel = nvalid < K ? nvalid : K;

    float my_v = -INFINITY;
    int32_t my_i = -1;
    bool doubt = false;
    float tailmax = -INFINITY;
#pragma unroll
    for (int s = 0; s < NPL; ++s)
      if (tail[s]) tailmax = fmaxf(tailmax, val[s]);
    if (ksel > 0) {
      // largest T with #{key >= T} >= ksel
      uint32_t lo = kmin, hi = kmax;
      while (lo < hi) {
        const uint32_t mid = lo + ((hi - lo + 1) >> 1);
        int c = 0;
#pragma unroll
        for (int s = 0; s < NPL; ++s) c += key[s] >= mid;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= ksel) lo = mid;
        else hi = mid - 1;
      }
      const uint32_t T = lo;
      bool sel[NPL];
      int above = 0;
#pragma unroll
      for (int s = 0; s < NPL; ++s) {
        sel[s] = key[s] > T;
        above += sel[s];
        if (tail[s] && key[s] >= T) {
          doubt = true;
          const int p = (lane + 32 * s) / KL;
          atomicOr(&s_dmask[warp][p >> 5], 1u << (p & 31));
        }
      }
      above = __reduce_add_sync(0xffffffffu, above);
      for (int need = ksel - above; need > 0; --need) {  // ties at the cut: ascending item
        uint32_t best = 0xFFFFFFFFu;
#pragma unroll
        for (int s = 0; s < NPL; ++s)
          if (key[s] == T && !sel[s] && static_cast<uint32_t>(item[s]) < best) best = static_cast<uint32_t>(item[s]);
        best = __reduce_min_sync(0xffffffffu, best);
#pragma unroll
        for (int s = 0; s < NPL; ++s)
          if (key[s] == T && !sel[s] && static_cast<uint32_t>(item[s]) == best) sel[s] = true;
      }
      // compact the winners (one per lane), rank them by counting on (key desc, item asc)
      int base = 0;
#pragma unroll
      for (int s = 0; s < NPL; ++s) {
        const unsigned b = __ballot_sync(0xffffffffu, sel[s]);
        if (sel[s]) {
          const int pos = base + __popc(b & ((1u << lane) - 1u));
          if (pos < 32)
            s_win[warp][pos] = (static_cast<unsigned long long>(key[s]) << 32) | static_cast<uint32_t>(~item[s]);
        }
        base += __popc(b);
      }
      __syncwarp();
      unsigned long long mine = 0;
      int rank = 0;
      if (lane < ksel) {
        mine = s_win[warp][lane];
        for (int j = 0; j < ksel; ++j) rank += s_win[warp][j] > mine;
      }
      __syncwarp();
      if (lane < ksel) s_win[warp][rank] = mine;
      __syncwarp();
      if (lane < ksel) {
        const unsigned long long w = s_win[warp][lane];
        my_v = f32_from_order_key(static_cast<uint32_t>(w >> 32));
        my_i = static_cast<int32_t>(~static_cast<uint32_t>(w));
      }
    }
    finish_row(a, row, lane, my_v, my_i, doubt, s_dmask[warp], cnt, s_hits, tailmax);
  }
  __syncthreads();
  if (a.hits && threadIdx.x < HGR_NUM_HITS && s_hits[threadIdx.x] != 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(a.hits) + threadIdx.x,
              static_cast<unsigned long long>(s_hits[threadIdx.x]));
}

// ---- ranking merge (very few candidates per row) ------------------------------------------------------------------
// Up to 32 * NPL <= 96 candidates per row -- the class-sharded head at B = 4096: 4-6 lists of 10-16 entries for each of
// 4,096 rows, where the bisection of topk_select_kernel (~1,500 instructions per row, 66 % issue utilisation: it
// competes with the persistent GEMM of the next batch for SMs) is the expensive part.  Here every candidate is packed
// into one 64-bit word (order key << 32 | ~item: larger word = better candidate, ties by ascending item) and staged in
// shared memory list by list; the lists are SORTED, so the output position of a candidate is its position in its own
// list plus, for every other list, the number of entries that beat it -- a 5-step binary search each.
// ~350 instructions per row at 40 candidates.  Same result, same order as the other merge kernels.  Requires every
// list sorted by (value desc, item asc) -- true for the lists the scoring kernel writes.
template <int NPL>
__global__ void __launch_bounds__(kMergeWarps * 32, 7)
topk_rank_kernel(const __grid_constant__ MergeArgs a) {
  __shared__ int s_hits[HGR_NUM_HITS];
  __shared__ unsigned long long s_list[kMergeWarps][32 * NPL];
  __shared__ unsigned long long s_win[kMergeWarps][32];
  __shared__ uint32_t s_dmask[kMergeWarps][kDoubtWords];
  if (threadIdx.x < HGR_NUM_HITS) s_hits[threadIdx.x] = 0;
  if (threadIdx.x < kMergeWarps * kDoubtWords) (&s_dmask[0][0])[threadIdx.x] = 0;
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int KL = a.KL, K = a.K;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kMergeWarps + warp;
  const int64_t pstride = a.part_stride > 0 ? a.part_stride : a.B * KL;
  if (row < a.B) {
    int cnt = static_cast<int>(a.P);
    if (a.use_sched) cnt = a.sched.parts(static_cast<int32_t>(row / a.sched.rows)) * a.wpq;
    const int N = cnt * KL;

    unsigned long long w[NPL];   // 0 = empty slot
    int32_t item[NPL];
    float val[NPL];
    int lp[NPL], lk[NPL];        // list / position of my candidate
#pragma unroll
    for (int s = 0; s < NPL; ++s) {   // all loads first (an out-of-range slot re-reads entry 0)
      const int e = lane + 32 * s;
      lp[s] = e < N ? e / KL : 0;
      lk[s] = e < N ? e - lp[s] * KL : 0;
      const int64_t g = N > 0 ? lp[s] * pstride + row * KL + lk[s] : 0;
      item[s] = (N > 0) ? __ldg(a.part_idx + g) : -1;
      val[s] = (N > 0) ? __ldg(a.part_val + g) : 0.f;
    }
    int nvalid = 0;
    float tailmax = -INFINITY;
    bool tail[NPL];
#pragma unroll
    for (int s = 0; s < NPL; ++s) {
      const bool ok = (lane + 32 * s < N) && item[s] >= 0;
      w[s] = ok ? (static_cast<unsigned long long>(f32_order_key(val[s])) << 32) | static_cast<uint32_t>(~item[s]) : 0ull;
      tail[s] = ok && lk[s] == KL - 1;
      if (tail[s]) tailmax = fmaxf(tailmax, val[s]);
      nvalid += ok;
      s_list[warp][lane + 32 * s] = w[s];
    }
    __syncwarp();
    nvalid = __reduce_add_sync(0xffffffffu, nvalid);
    const int ksel = nvalid < K ? nvalid : K;

#pragma unroll
    for (int s = 0; s < NPL; ++s) {
      int rank = lk[s];
#pragma unroll 4
      for (int q = 0; q < cnt; ++q) {
        // entries of list q that come before mine: strictly larger words; an identical word (the same item reported
        // twice) counts when it sits in an earlier list
        const unsigned long long bar = w[s] - (q < lp[s] ? 1ull : 0ull);
        const unsigned long long* lq = &s_list[warp][q * KL];
        int lo = 0, hi = KL;
        while (lo < hi) {   // <= 6 steps, warp-uniform trip count would need KL a power of two: plain loop
          const int mid = (lo + hi) >> 1;
          if (lq[mid] > bar) lo = mid + 1;
          else hi = mid;
        }
        rank += q == lp[s] ? 0 : lo;
      }
      if (w[s] != 0ull && rank < 32) s_win[warp][rank] = w[s];
    }
    __syncwarp();
    float my_v = -INFINITY;
    int32_t my_i = -1;
    bool doubt = false;
    if (ksel > 0) {
      if (lane < ksel) {
        const unsigned long long x = s_win[warp][lane];
        my_v = f32_from_order_key(static_cast<uint32_t>(x >> 32));
        my_i = static_cast<int32_t>(~static_cast<uint32_t>(x));
      }
      const uint32_t T = static_cast<uint32_t>(s_win[warp][ksel - 1] >> 32);   // key of the K-th winner
#pragma unroll
      for (int s = 0; s < NPL; ++s) {
        if (tail[s] && static_cast<uint32_t>(w[s] >> 32) >= T) {
          doubt = true;
          atomicOr(&s_dmask[warp][lp[s] >> 5], 1u << (lp[s] & 31));
        }
      }
    }
    finish_row(a, row, lane, my_v, my_i, doubt, s_dmask[warp], cnt, s_hits, tailmax);
  }
  __syncthreads();
  if (a.hits && threadIdx.x < HGR_NUM_HITS && s_hits[threadIdx.x] != 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(a.hits) + threadIdx.x,
              static_cast<unsigned long long>(s_hits[threadIdx.x]));
}

// ---- candidate-list merge (floor-sketch epilogue) ---------------------------------------------------------------
// The lists of a row are short, UNSORTED and of different lengths (sk_cnt).  One warp per row: the lists are gathered
// into shared memory (lane = list, one 8-byte entry per step), the K-th largest order key is found by bisection over
// the gathered entries (stops early when a cut holds exactly K), ties at the cut resolve by ascending bank row, and the
// <= K winners are ranked by counting on (key desc, bank row asc) -- the same documented order as the other paths.
__global__ void __launch_bounds__(kMergeWarps * 32, 7)  // (the cold repair functions must not set the register count)
topk_merge_counts_kernel(const __grid_constant__ MergeArgs a) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ int s_hits[HGR_NUM_HITS];
  __shared__ unsigned long long s_win[kMergeWarps][32];
  if (threadIdx.x < HGR_NUM_HITS) s_hits[threadIdx.x] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = a.K, cap = a.sk_cap;
  const int64_t row = static_cast<int64_t>(blockIdx.x) * kMergeWarps + warp;
  uint2* cand = reinterpret_cast<uint2*>(smem_raw) + static_cast<size_t>(warp) * a.P * cap;
  if (row < a.B) {
    int lists = static_cast<int>(a.P);
    if (a.use_sched) lists = a.sched.parts(static_cast<int32_t>(row / a.sched.rows));
    int N = 0;
    for (int l0 = 0; l0 < lists; l0 += 32) {
      const int l = l0 + lane;
      int c = l < lists ? __ldg(a.sk_cnt + static_cast<int64_t>(l) * a.B + row) : 0;
      c = c < 0 ? 0 : (c > cap ? cap : c);
      int incl = c;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
      }
      const uint2* src = a.sk_part + (static_cast<int64_t>(l < lists ? l : 0) * a.B + row) * cap;
      uint2* dst = cand + N + incl - c;
      const int maxc = __reduce_max_sync(0xffffffffu, c);
      for (int k = 0; k < maxc; ++k)
        if (k < c) dst[k] = __ldg(src + k);
      N += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    // keys of this lane's entries: lane, lane + 32, ...
    uint32_t kmax = 0, kmin = 0xFFFFFFFFu;
    for (int e = lane; e < N; e += 32) {
      const uint32_t k = f32_order_key(__uint_as_float(cand[e].x));
      cand[e].x = k;   // keep the key: the value is recovered from it
      kmax = k > kmax ? k : kmax;
      kmin = k < kmin ? k : kmin;
    }
    __syncwarp();
    kmax = __reduce_max_sync(0xffffffffu, kmax);
    kmin = __reduce_min_sync(0xffffffffu, kmin);
    const int ksel = N < K ? N : K;
    float my_v = -INFINITY;
    int32_t my_i = -1;
    if (ksel > 0) {
      // largest T with #{key >= T} >= ksel; done as soon as a cut holds exactly ksel entries
      uint32_t lo = kmin, hi = kmax;
      int c_lo = N;
      while (lo < hi && c_lo != ksel) {
        const uint32_t mid = lo + ((hi - lo + 1u) >> 1);
        int c = 0;
        for (int e = lane; e < N; e += 32) c += cand[e].x >= mid;
        c = __reduce_add_sync(0xffffffffu, c);
        if (c >= ksel) {
          lo = mid;
          c_lo = c;
        } else {
          hi = mid - 1u;
        }
      }
      const uint32_t T = lo;
      int need = 0;   // tied entries at T still to be taken (ascending bank row)
      if (c_lo != ksel) {
        int above = 0;
        for (int e = lane; e < N; e += 32) above += cand[e].x > T;
        need = ksel - __reduce_add_sync(0xffffffffu, above);
      }
      uint32_t tie_floor = 0;   // tied bank rows below this one are already taken
      bool tie_any = false;
      int base = 0;
      // winners above the cut (or at it, when the cut is exact), compacted one per lane
      for (int e0 = 0; e0 < N; e0 += 32) {
        const int e = e0 + lane;
        const bool sel = e < N && (c_lo == ksel ? cand[e].x >= T : cand[e].x > T);
        const unsigned b = __ballot_sync(0xffffffffu, sel);
        if (sel) {
          const int pos = base + __popc(b & ((1u << lane) - 1u));
          if (pos < 32) s_win[warp][pos] = (static_cast<unsigned long long>(cand[e].x) << 32) | static_cast<uint32_t>(~cand[e].y);
        }
        base += __popc(b);
      }
      for (; need > 0; --need) {
        uint32_t best = 0xFFFFFFFFu;
        for (int e = lane; e < N; e += 32)
          if (cand[e].x == T && (!tie_any || cand[e].y > tie_floor) && cand[e].y < best) best = cand[e].y;
        best = __reduce_min_sync(0xffffffffu, best);
        if (lane == 0 && base < 32) s_win[warp][base] = (static_cast<unsigned long long>(T) << 32) | static_cast<uint32_t>(~best);
        ++base;
        tie_floor = best;
        tie_any = true;
      }
      __syncwarp();
      unsigned long long mine = 0;
      int rank = 0;
      if (lane < ksel) {
        mine = s_win[warp][lane];
        for (int j = 0; j < ksel; ++j) rank += s_win[warp][j] > mine;
      }
      __syncwarp();
      if (lane < ksel) s_win[warp][rank] = mine;
      __syncwarp();
      if (lane < ksel) {
        const unsigned long long w = s_win[warp][lane];
        my_v = f32_from_order_key(static_cast<uint32_t>(w >> 32));
        my_i = static_cast<int32_t>(~static_cast<uint32_t>(w));
      }
    }
    finish_row(a, row, lane, my_v, my_i, false, nullptr, 0, s_hits);
  }
  __syncthreads();
  if (a.hits && threadIdx.x < HGR_NUM_HITS && s_hits[threadIdx.x] != 0)
    atomicAdd(reinterpret_cast<unsigned long long*>(a.hits) + threadIdx.x,
              static_cast<unsigned long long>(s_hits[threadIdx.x]));
}

// ---- cross-GPU sequencing of the peer-memory exchange ---------------------------------------------------------
// Every rank keeps `n` flag words (one per producer rank) in its exchange buffer.  After the kernel that wrote its
// candidates into the peers' buffers, a rank launches peer_signal: thread g stores the rank's running sequence
// number into ITS flag word on rank g (system-scope release; the scatter kernel finished before, stream order).
// Before merging, a rank launches peer_wait: thread g spins (system-scope acquire) until producer g's flag has
// reached the consumer's own running sequence number.  Both counters live in device memory and advance by one per
// launch, so the pair can be captured in a CUDA graph and replayed.
struct PeerFlags {
  uint32_t* p[kMaxScatterBlocks];
};

__global__ void peer_signal_kernel(const PeerFlags flags, int n, uint32_t* seq) {
  __shared__ uint32_t s_v;
  if (threadIdx.x == 0) s_v = ++(*seq);
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < n) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flags.p[threadIdx.x]), "r"(s_v) : "memory");
  }
}

__global__ void peer_wait_kernel(const uint32_t* flags, int n, uint32_t* seq, unsigned long long timeout_ns) {
  __shared__ uint32_t s_v;
  if (threadIdx.x == 0) s_v = ++(*seq);
  __syncthreads();
  if (static_cast<int>(threadIdx.x) < n) {
    const uint32_t want = s_v;
    unsigned long long t0 = 0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
      uint32_t got;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(got) : "l"(flags + threadIdx.x) : "memory");
      if (static_cast<int32_t>(got - want) >= 0) break;
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > timeout_ns) {  // a peer died or the schedules diverged -- fail loudly, never hang the GPU
        printf("hgr peer_wait: producer %d stuck at %u, waiting for %u\n", static_cast<int>(threadIdx.x), got, want);
        __trap();
      }
      __nanosleep(200);
    }
  }
}

}  // namespace

int launch_peer_signal(uint32_t* const* flags, int n, uint32_t* seq, cudaStream_t stream) {
  if (n < 1 || n > kMaxScatterBlocks) return set_error(HGR_ERR_BAD_ARG, "peer signal: %d ranks outside [1, %d]", n, kMaxScatterBlocks);
  PeerFlags f{};
  for (int g = 0; g < n; ++g) f.p[g] = flags[g];
  peer_signal_kernel<<<1, 32, 0, stream>>>(f, n, seq);
  HGR_CHECK_LAUNCH();
  return HGR_OK;
}

int launch_peer_wait(const uint32_t* flags, int n, uint32_t* seq, cudaStream_t stream) {
  if (n < 1 || n > kMaxScatterBlocks) return set_error(HGR_ERR_BAD_ARG, "peer wait: %d ranks outside [1, %d]", n, kMaxScatterBlocks);
  // how long a consumer waits for the slowest producer before it traps: 10 s unless HGR_PEER_TIMEOUT_MS says otherwise
  // (a rank whose host stalls longer than this between two graph launches would take the job down)
  static const unsigned long long timeout_ns = [] {
    const char* e = getenv("HGR_PEER_TIMEOUT_MS");
    const long long ms = e ? atoll(e) : 10000;
    return static_cast<unsigned long long>(ms > 0 ? ms : 10000) * 1000000ull;
  }();
  peer_wait_kernel<<<1, 32, 0, stream>>>(flags, n, seq, timeout_ns);
  HGR_CHECK_LAUNCH();
  return HGR_OK;
}

int launch_topk_merge(const MergeArgs& args, cudaStream_t stream) {
  if (args.B == 0 || args.K == 0) return HGR_OK;
  if (args.sk_part != nullptr) {
    if (args.K > 32 || args.sk_cap < 1 || args.sk_cnt == nullptr)
      return set_error(HGR_ERR_BAD_ARG, "topk merge (candidate lists): K = %d / cap = %d", args.K, args.sk_cap);
    const int blocks = static_cast<int>((args.B + kMergeWarps - 1) / kMergeWarps);
    const size_t smem = static_cast<size_t>(kMergeWarps) * args.P * args.sk_cap * 8;
    if (smem > 200 * 1024) return set_error(HGR_ERR_UNSUPPORTED, "topk merge: %zu bytes of candidates per CTA", smem);
    HGR_CHECK_CUDA(cudaFuncSetAttribute(topk_merge_counts_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    topk_merge_counts_kernel<<<blocks, kMergeWarps * 32, smem, stream>>>(args);
    HGR_CHECK_LAUNCH();
    return HGR_OK;
  }
  if (args.P > 320)
    return set_error(HGR_ERR_UNSUPPORTED, "topk merge: %lld lists per row exceed 320", (long long)args.P);
  if (args.K > HGR_TOPK_MAX || args.KL < 1)
    return set_error(HGR_ERR_UNSUPPORTED, "topk merge: K = %d / KL = %d unsupported", args.K, args.KL);
  if (args.KL < args.K && (args.X == nullptr || args.bank == nullptr || !args.use_sched || args.wpq < 1))
    return set_error(HGR_ERR_BAD_ARG, "topk merge: speculative lists need X / bank / the schedule for the exact repair");
  const int blocks = static_cast<int>((args.B + kMergeWarps - 1) / kMergeWarps);
  const int64_t cand = args.P * args.KL;  // candidates per row (upper bound)
  // (only for the scoring kernel's own lists: bank rows, sorted (value desc, row asc) by construction -- lists handed
  //  to hgr_topk_merge carry node ids, whose order among equal values is the caller's)
#ifndef HGR_NO_RANK_KERNEL   // (kernel experiments: tools/build_variant.sh norank -DHGR_NO_RANK_KERNEL topk_merge.cu)
  if (cand <= 96 && args.use_sched) {
    if (cand <= 64) topk_rank_kernel<2><<<blocks, kMergeWarps * 32, 0, stream>>>(args);
    else topk_rank_kernel<3><<<blocks, kMergeWarps * 32, 0, stream>>>(args);
    HGR_CHECK_LAUNCH();
    return HGR_OK;
  }
#endif
  if (cand <= 320) {
    if (cand <= 128) topk_select_kernel<4><<<blocks, kMergeWarps * 32, 0, stream>>>(args);
    else if (cand <= 192) topk_select_kernel<6><<<blocks, kMergeWarps * 32, 0, stream>>>(args);
    else topk_select_kernel<10><<<blocks, kMergeWarps * 32, 0, stream>>>(args);
    HGR_CHECK_LAUNCH();
    return HGR_OK;
  }
  const size_t smem = static_cast<size_t>(kMergeWarps) * args.P * args.KL * 8;
  if (smem > 200 * 1024) return set_error(HGR_ERR_UNSUPPORTED, "topk merge: %zu bytes of lists per CTA", smem);
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
