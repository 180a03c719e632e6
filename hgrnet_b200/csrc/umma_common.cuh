// Device-side pieces shared by the single-CTA (score_umma.cu) and CTA-pair (score_umma2.cu) tcgen05 kernels:
// tile walker over the stream-K-style schedule, kernel parameters and the fused top-k epilogue.
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"
#include "topk_list.cuh"

namespace hgr {
namespace umma {

constexpr int kBlockK = 64;                         // bf16 per K block: 128 bytes = one swizzle row
constexpr int kUmmaK = 16;                          // K of one tcgen05.mma.kind::f16
constexpr int kABytes = kTileM * kBlockK * 2;       // 16 KB
constexpr int kBBoxRows = 64;                       // bank rows per TMA box
constexpr int kBBoxBytes = kBBoxRows * kBlockK * 2; // 8 KB
constexpr int kEpiWarp0 = 2;
constexpr int kTmemCols = 512;
constexpr int kChunk = 32;                          // accumulator columns per tcgen05.ld

enum EpiMode { kEpiDense = 0, kEpiTopkReload = 1, kEpiTopkQueue = 2, kEpiNull = 3 };

struct SubTile {
  int mt;      // row tile
  int col0;    // first bank row
  int n;       // MMA N (multiple of 16)
  int nvalid;  // bank rows < C inside the sub-tile
  bool first;  // first sub-tile of a (row tile, CTA) segment
  bool last;   // last sub-tile of the segment
};

struct TileWalker {
  int64_t u, u_end;
  int U;
  int64_t C;
  bool first;
  __device__ TileWalker(const Sched& s, int cta, int64_t C_)
      : u(s.unit_begin(cta)), u_end(s.unit_begin(cta + 1)), U(s.U), C(C_), first(true) {}
  __device__ bool next(SubTile& t) {
    if (u >= u_end) return false;
    const int mt = static_cast<int>(u / U);
    int uu = static_cast<int>(u - static_cast<int64_t>(mt) * U);
    int nu = kSubN / kUnit;
    if (U - uu < nu) nu = U - uu;
    if (u_end - u < nu) nu = static_cast<int>(u_end - u);
    t.mt = mt;
    t.col0 = uu * kUnit;
    t.n = nu * kUnit;
    const int64_t left = C - t.col0;
    t.nvalid = left < t.n ? static_cast<int>(left) : t.n;
    t.first = first;
    u += nu;
    uu += nu;
    t.last = (u >= u_end) || (uu == U);
    first = t.last;
    return true;
  }
};

struct Params {
  Sched sched;
  int64_t B, C;
  int num_k_blocks;
  int KL;              // entries written per list
  float scale;
  float* part_val;     // [slots][B][KL]
  int32_t* part_idx;   // [slots][B][KL] bank rows
  float* dense_out;    // [B][ldo]
  int64_t ldo;
  unsigned int* stats; // [0] = rows re-scanned by the merge kernel of this call (reset here)
};

template <int KL>
__device__ __forceinline__ void scan_chunk_reload(SortedList<KL>& list, const uint32_t (&r)[kChunk], int nv,
                                                  uint32_t taddr_chunk, int col_chunk) {
  const float thr = list.thr();
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < kChunk; ++j)
    if (__uint_as_float(r[j]) > thr) m |= (1u << j);
  if (nv < kChunk) m &= (1u << nv) - 1u;
  uint32_t wm = __reduce_or_sync(0xffffffffu, m);
  // visit the union of positions in ascending order; each lane re-reads its own value of that
  // column from TMEM (the address is warp-uniform) and inserts if it still qualifies
  while (wm) {
    const int j = __ffs(wm) - 1;
    wm &= wm - 1;
    const float x = __uint_as_float(ptx::tmem_ld_x1(taddr_chunk + j));
    ptx::tmem_ld_wait();
    if (((m >> j) & 1u) && x > list.thr()) list.insert(x, col_chunk + j);
  }
}

template <int KL, int J>
struct SeedPrefix {  // unconditional inserts of columns 0..J-1 into an empty list, slot depth growing with j
  static __device__ __forceinline__ void run(SortedList<KL>& list, const uint32_t (&r)[kChunk], int col_chunk) {
    SeedPrefix<KL, J - 1>::run(list, r, col_chunk);
    list.template insert_prefix<(J < KL ? J : KL)>(__uint_as_float(r[J - 1]), col_chunk + J - 1);
  }
};
template <int KL>
struct SeedPrefix<KL, 0> {
  static __device__ __forceinline__ void run(SortedList<KL>&, const uint32_t (&)[kChunk], int) {}
};

// qaddr: shared-space byte address of this thread's column of the [kChunk][epilogue threads] fp32 staging
// array; QSTRIDE_B: bytes between consecutive entries of one thread.  Columns < JSTART are skipped (already
// seeded into the list).
template <int KL, int QSTRIDE_B, int JSTART>
__device__ __forceinline__ void scan_chunk_queue(SortedList<KL>& list, const uint32_t (&r)[kChunk], int nv,
                                                 int col_chunk, uint32_t qaddr) {
  // 1) lane-private compaction of the values that beat the KL-th best at chunk entry:
  //    values go to the queue in column order, their positions into a bit mask
  const float thr = list.thr();
  uint32_t m = 0;
  uint32_t wr = qaddr;
#pragma unroll
  for (int j = JSTART; j < kChunk; ++j) {
    const float x = __uint_as_float(r[j]);
    if (x > thr) {
      ptx::st_shared_f32(wr, x);
      wr += QSTRIDE_B;
      m |= (1u << j);
    }
  }
  if (nv < kChunk) m &= (1u << nv) - 1u;  // ragged tail: columns >= C were zero-filled by TMA, drop them
  const int cnt = __popc(m);
  // 2) dense drain: lanes walk their own queues in lock-step, so the (long) insert body runs
  //    max_lane(cnt) times instead of once per column any lane hit
  const int maxc = __reduce_max_sync(0xffffffffu, cnt);
  uint32_t rd = qaddr;
  for (int e = 0; e < maxc; ++e) {
    if (e < cnt) {
      const float x = ptx::ld_shared_f32(rd);
      rd += QSTRIDE_B;
      const int j = __ffs(m) - 1;
      m &= m - 1;
      if (x > list.thr()) list.insert(x, col_chunk + j);
    }
  }
}

// CTA-pair kernel (score_umma2.cu)
constexpr int kPairWpq = 2;  // epilogue warps per TMEM lane quarter
int launch_pair_kernel(int epi, int KL, const CUtensorMap& mx, const CUtensorMap& mb, const Params& p,
                       cudaStream_t stream);

}  // namespace umma
}  // namespace hgr
