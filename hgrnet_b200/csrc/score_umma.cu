// Kernel (2): TMA-fed tcgen05/TMEM bf16 GEMM for image x class cosine logits with the per-row
// running top-K fused into the epilogue (HGR_IMPL_TCGEN05) -- the B x C logit matrix never
// reaches HBM.  The same main loop with a plain store epilogue backs hgr_logits_dense.
//
// Reference: `feats @ self.zsl_weights.T` (model/clip_tree.py:331), `logits[:, test_index]`
// + `.topk(20, 1, True, True)` (main.py:136-138); the id mapping / hit test (main.py:139-147)
// runs in the merge kernel (topk_merge.cu).
//
// Structure (one persistent CTA per SM, warp-specialised):
//   warp 0        TMA producer: X tile [128 x 64] and bank tile [<=256 x 64] bf16 per K block,
//                 128B-swizzled, 4-stage mbarrier ring;
//   warp 1        allocates TMEM (512 columns = two 128 x 256 fp32 accumulators), one lane issues
//                 tcgen05.mma (M = 128, N = 16..256, K = 16) and commits to mbarriers;
//   warps 2..     epilogue, WPQ warps per TMEM lane quarter: tcgen05.ld the accumulator of
//                 sub-tile t while the tensor core works on sub-tile t+1.  Thread = TMEM lane =
//                 image row; the WPQ warps of a quarter take the 32-column chunks of a sub-tile
//                 round-robin, and each thread keeps its row's sorted top-KL of the chunks it saw in
//                 registers across all sub-tiles of the CTA's class range.
//
// Epilogue cost model (what the design answers to): a warp owns an SM sub-partition's issue
// port, so (a) values that beat the row's current KL-th best are first COMPACTED per lane into
// a shared-memory queue (predicated stores, no divergence) and then drained in lock-step -- the
// long insert body runs max_lane(count) times per chunk instead of once per column that any
// lane hit; (b) several warps per quarter hide the dependent-select latency of the insert;
// (c) when a row is split over many lists (many CTAs per row tile) the lists are SPECULATIVE,
// KL < K entries, which divides the insert work by ~3; the merge kernel certifies each row and
// re-scans the (practically never occurring) uncertified ones exactly.
// Work split: sched.cuh (stream-K-style split of the class dimension in units of 16 rows).
#include <cuda.h>

#include <cmath>

#include "sketch_epi.cuh"
#include "umma_common.cuh"

namespace hgr {
using namespace umma;
namespace {

constexpr int kStages = 4;
constexpr int kBBytes = kSubN * kBlockK * 2;        // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;      // 48 KB

struct Ctl {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

template <int EPI, int KL, int WPQ>
__global__ void __launch_bounds__(64 + 128 * WPQ, 1)
score_umma_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_bank,
                  const Params p) {
  constexpr int kEpiThreads = 128 * WPQ;
  constexpr int kQueueBytes = EPI == kEpiTopkQueue ? kChunk * kEpiThreads * 4 : 0;
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles must start on 1024-byte boundaries
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  float* queue_base = reinterpret_cast<float*>(smem + kStages * kStageBytes);
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + kStages * kStageBytes + kQueueBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta = blockIdx.x;

  if (cta == 0 && threadIdx.x == 0 && p.stats != nullptr) p.stats[0] = 0;
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_x);
    ptx::prefetch_tensormap(&map_bank);
    for (int s = 0; s < kStages; ++s) {
      ptx::mbar_init(&ctl->full[s], 1);
      ptx::mbar_init(&ctl->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&ctl->tmem_full[b], 1);
      ptx::mbar_init(&ctl->tmem_empty[b], kEpiThreads / 32);
    }
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc(&ctl->tmem_base, kTmemCols);
    ptx::tmem_relinquish();
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const uint64_t pol = ptx::policy_evict_last();
      TileWalker walk(p.sched, cta, p.C, p.rem_first);
      SubTile t;
      int stage = 0;
      uint32_t phase = 0;
      while (walk.next(t)) {
        const int nbox = (t.n + kBBoxRows - 1) / kBBoxRows;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          ptx::mbar_wait(&ctl->empty[stage], phase ^ 1u);
          uint8_t* sa = stage_base + stage * kStageBytes;
          uint8_t* sb = sa + kABytes;
          ptx::mbar_arrive_expect_tx(&ctl->full[stage], kABytes + nbox * kBBoxBytes);
          ptx::tma_load_2d(sa, &map_x, &ctl->full[stage], kb * kBlockK, t.mt * kTileM, pol);
          for (int b = 0; b < nbox; ++b)
            ptx::tma_load_2d(sb + b * kBBoxBytes, &map_bank, &ctl->full[stage], kb * kBlockK,
                             t.col0 + b * kBBoxRows, pol);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      TileWalker walk(p.sched, cta, p.C, p.rem_first);
      SubTile t;
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      while (walk.next(t)) {
        const int buf = it & 1;
        ptx::mbar_wait(&ctl->tmem_empty[buf], ((it >> 1) & 1) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + buf * kSubN;
        const uint32_t idesc = ptx::umma_idesc_bf16(kTileM, t.n);
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          ptx::mbar_wait(&ctl->full[stage], phase);
          ptx::tc_fence_after();
          const uint32_t a_addr = ptx::smem_u32(stage_base + stage * kStageBytes);
          const uint32_t b_addr = a_addr + kABytes;
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t adesc = ptx::umma_smem_desc_sw128(a_addr + k * kUmmaK * 2);
            const uint64_t bdesc = ptx::umma_smem_desc_sw128(b_addr + k * kUmmaK * 2);
            ptx::umma_bf16(d_tmem, adesc, bdesc, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          ptx::umma_commit(&ctl->empty[stage]);  // smem slot reusable once these MMAs retire
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1u;
          }
        }
        ptx::umma_commit(&ctl->tmem_full[buf]);  // accumulator complete -> epilogue
        ++it;
      }
    }
  } else {
    // ===================== epilogue =====================
    const int quarter = warp & 3;                       // TMEM lane quarter this warp may access
    const int member = (warp - kEpiWarp0) >> 2;         // which of the WPQ warps of that quarter
    const int row_in_tile = quarter * 32 + lane;
    const int epi_tid = (warp - kEpiWarp0) * 32 + lane;
    const uint32_t qaddr = ptx::smem_u32(queue_base + epi_tid * kChunk);  // private 128-byte staging row
    const uint32_t qswz = static_cast<uint32_t>(epi_tid) & 7u;
    TileWalker walk(p.sched, cta, p.C, p.rem_first);
    SubTile t;
    SortedList<KL> list;
    list.init();
    float null_acc = -INFINITY;
    float floor_thr = -INFINITY;
    int it = 0;
    EpiClock ck(p.timeline != nullptr && epi_tid == 0);
    while (walk.next(t)) {
      const int buf = it & 1;
      ck.start();
      ptx::mbar_wait(&ctl->tmem_full[buf], (it >> 1) & 1);
      ptx::tc_fence_after();
      ck.lap(ck.wait);
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * kSubN;
      const int64_t row = static_cast<int64_t>(t.mt) * kTileM + row_in_tile;
      if (t.first) {
        list.init();
        null_acc = -INFINITY;
        floor_thr = -INFINITY;
      }
      if (EPI == kEpiTopkQueue && t.seq < 2) {  // warm-up floor from the first two sub-tiles of a segment
        floor_thr = fmaxf(floor_thr, warmup_floor<KL, WPQ>(taddr, member, t.nvalid));
        ck.lap(ck.warm);
      }
      for (int c0 = member * kChunk; c0 < t.nvalid; c0 += WPQ * kChunk) {
        uint32_t r[kChunk];
        ck.start();
        ptx::tmem_ld_x32(taddr + c0, r);
        ptx::tmem_ld_wait();
        ck.lap(ck.ld);
        const int nv = t.nvalid - c0;
        if (EPI == kEpiDense) {
          if (row < p.B) {
            float* o = p.dense_out + row * p.ldo + t.col0 + c0;
#pragma unroll
            for (int j = 0; j < kChunk; ++j)
              if (j < nv) o[j] = __uint_as_float(r[j]) * p.scale;
          }
        } else if (EPI == kEpiTopkReload) {
          scan_chunk_reload<KL>(list, r, nv, taddr + c0, t.col0 + c0);
        } else if (EPI == kEpiTopkQueue) {
          scan_chunk_queue<KL>(list, r, nv, t.col0 + c0, qaddr, qswz, floor_thr, ck);
        } else {
#pragma unroll
          for (int j = 0; j < kChunk; ++j) null_acc = fmaxf(null_acc, __uint_as_float(r[j]));
        }
      }
      // accumulator buffer drained: hand it back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&ctl->tmem_empty[buf]);
      if (EPI != kEpiDense && t.last && row < p.B) {
        const int slot = (cta - p.sched.first_cta(t.mt)) * WPQ + member;
        float* pv = p.part_val + (static_cast<int64_t>(slot) * p.B + row) * p.KL;
        int32_t* pi = p.part_idx + (static_cast<int64_t>(slot) * p.B + row) * p.KL;
        if (EPI == kEpiNull) {
          pv[0] = null_acc;
          pi[0] = -1;
        } else {
#pragma unroll
          for (int k = 0; k < KL; ++k) {
            if (k < p.KL) {
              pv[k] = list.v[k];
              pi[k] = list.i[k];
            }
          }
        }
      }
      ++it;
    }
  }

  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc(tmem_base, kTmemCols);
  }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// [rows, D] row-major bf16 matrix, box = [box_rows, 64] elements, 128B swizzle, zero OOB fill
int make_map(CUtensorMap* map, const void* base, int64_t rows, int64_t D, int box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return set_error(HGR_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(D), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(D) * 2};
  const cuuint32_t box[2] = {static_cast<cuuint32_t>(kBlockK), static_cast<cuuint32_t>(box_rows)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return set_error(HGR_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return HGR_OK;
}

template <int EPI, int KL, int WPQ>
int launch_kernel(const CUtensorMap& mx, const CUtensorMap& mb, const Params& p, cudaStream_t stream) {
  constexpr int threads = 64 + 128 * WPQ;
  const size_t smem = 1024 + static_cast<size_t>(kStages) * kStageBytes +
                      (EPI == kEpiTopkQueue ? static_cast<size_t>(kChunk) * 128 * WPQ * 4 : 0) + sizeof(Ctl);
  auto kern = score_umma_kernel<EPI, KL, WPQ>;
  HGR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<p.sched.G, threads, smem, stream>>>(mx, mb, p);
  HGR_CHECK_LAUNCH();
  return HGR_OK;
}

// SMs left free for concurrently running small kernels (NCCL, merge, normalise of neighbouring batches): the
// persistent GEMM CTAs take a whole SM each (shared memory), so anything else can only overlap on SMs it does
// not occupy.  HGR_RESERVE_SMS (default 0) trades that many SMs of GEMM throughput for the overlap.
int usable_sms() {
  static const int reserve = [] {
    const char* e = getenv("HGR_RESERVE_SMS");
    const int r = e ? atoi(e) : 0;
    return r < 0 ? 0 : r;
  }();
  const int n = num_sms() - reserve;
  return n < 2 ? 2 : n;
}

// Rough cost of the slowest worker of a schedule, in operand columns streamed through shared memory: every
// sub-tile re-streams the A operand (one row tile = 256 "columns" worth of bytes) next to its own bank rows, and
// every segment (row tile x worker intersection) starts a fresh list (floor pass, warm-up inserts, list write-out).
static double sched_cost(const Sched& s) {
  constexpr double kSegmentOverhead = 400.0;   // ~3-4 us of floor pass + warm-up inserts + list write-out
  double worst = 0.0;
  for (int32_t w = 0; w < s.G; ++w) {
    int64_t u = s.unit_begin(w);
    const int64_t u_end = s.unit_begin(w + 1);
    double cost = 0.0;
    while (u < u_end) {
      const int64_t tile_end = (u / s.U + 1) * s.U;
      const int64_t e = u_end < tile_end ? u_end : tile_end;
      const int64_t cols = (e - u) * kUnit;
      cost += static_cast<double>((cols + kSubN - 1) / kSubN) * 256.0 + static_cast<double>(cols) + kSegmentOverhead;
      u = e;
    }
    worst = cost > worst ? cost : worst;
  }
  return worst;
}

// Workers: all CTA pairs, unless a slightly smaller count that is a multiple of the row-tile count wins -- then no
// chunk straddles a row tile (one list per worker, no ragged sub-tiles at both ends; measured at B = 4096:
// 64 aligned pairs beat 74 for every bank size, e.g. 33.6 vs 39.7 us at C = 2,731).
Sched pick_sched(int64_t B, int64_t C, bool pair) {
  if (!pair) return make_sched(B, C, usable_sms(), kTileM);
  const int most = usable_sms() / 2;
  Sched best = make_sched(B, C, most, 2 * kTileM);
  static const bool no_align = getenv("HGR_NO_ALIGN") != nullptr;
  if (!no_align && best.MT > 1 && best.MT <= most && best.G == most) {
    const int aligned = most / best.MT * best.MT;
    if (aligned != most && aligned * 10 >= most * 8) {
      const Sched alt = make_sched(B, C, aligned, 2 * kTileM);
      // only for short streams (<= 3072 columns per worker): there the per-segment costs decide; on long streams the
      // fewer, larger lists of the aligned split sit closer to the speculation limit (an expected repair of a
      // 5,000-column range costs more than the alignment saves -- measured at cfg 5)
      const int64_t cols_per_worker = (alt.T + alt.G - 1) / alt.G * kUnit;
      if (cols_per_worker <= 3072 && sched_cost(alt) < sched_cost(best)) best = alt;
    }
  }
  return best;
}

// Schedule of the resident-A kernel: a segment switch inside a worker costs a pipeline bubble (the next row tile's
// A operand can only be loaded once the last MMA of the previous segment has retired) and a fresh list, so a worker
// count that is a multiple of the row-tile count (no chunk straddles a row tile) wins whenever the idle pairs cost
// less than that.  Costs in bank columns of the slowest worker.
static double resident_cost(const Sched& s) {
  constexpr double kSegment = 200.0;   // A reload bubble + list warm-up of one more segment
  double worst = 0.0;
  for (int32_t w = 0; w < s.G; ++w) {
    int64_t u = s.unit_begin(w);
    const int64_t u_end = s.unit_begin(w + 1);
    double cost = 0.0;
    while (u < u_end) {
      const int64_t tile_end = (u / s.U + 1) * s.U;
      const int64_t e = u_end < tile_end ? u_end : tile_end;
      cost += static_cast<double>((e - u) * kUnit) + kSegment;
      u = e;
    }
    worst = cost > worst ? cost : worst;
  }
  return worst;
}

Sched pick_sched_resident(int64_t B, int64_t C) {
  const int most = usable_sms() / 2;
  Sched best = make_sched(B, C, most, 2 * kTileM);
  if (best.MT > 1 && best.MT <= most && best.G == most) {
    const int aligned = most / best.MT * best.MT;
    if (aligned != most) {
      const Sched alt = make_sched(B, C, aligned, 2 * kTileM);
      if (resident_cost(alt) < resident_cost(best)) best = alt;
    }
  }
  return best;
}

int common_setup(const __nv_bfloat16* X, const __nv_bfloat16* bank, int64_t B, int64_t C, int64_t D, bool pair,
                 CUtensorMap* mx, CUtensorMap* mb, Params* p) {
  int rc = make_map(mx, X, B, D, kTileM);
  if (rc != HGR_OK) return rc;
  rc = make_map(mb, bank, C, D, kBBoxRows);
  if (rc != HGR_OK) return rc;
  p->sched = pick_sched(B, C, pair);
  p->B = B;
  p->C = C;
  p->num_k_blocks = static_cast<int>((D + kBlockK - 1) / kBlockK);
  static const int rem_first = [] {
    const char* e = getenv("HGR_REM_FIRST");
    return (e && e[0] == '1') ? 1 : 0;
  }();
  p->rem_first = rem_first;
  return HGR_OK;
}

// ---- list-width policy -------------------------------------------------------------------
// With bank rows in random order, each of a row's true top-K members falls into a given list with probability
// `share` = the fraction of the row's columns that list streams (the largest one: a worker's chunk of the schedule,
// halved when two epilogue warps alternate its chunks).  A list of KL < K entries cannot be certified when it
// receives KL or more of them: probability <= C(K, KL) * share^KL.  A repair re-scans that list's column range with
// ONE warp of the merge kernel: measured ~0.6 us per bank row of 1024 elements (a 590-column range of cfg 2:
// 0.36 ms; a 4,700-column range of cfg 5: 2.8 ms), and the whole call waits for it.
double overflow_bound(int K, int KL, double share) {
  double c = 1.0;
  for (int i = 0; i < KL; ++i) c = c * (K - i) / (i + 1);
  return c * std::pow(share, KL);
}

double max_list_share(const Sched& s, int wpq) {
  const double chunk = static_cast<double>((s.T + s.G - 1) / s.G);   // units per worker
  const double share = chunk / static_cast<double>(s.U);
  return (share > 1.0 ? 1.0 : share) / wpq;
}

// Speculate only while the EXPECTED repair time per call stays below 1.5 % of the call's estimated duration (main
// loop at 60 % of the bf16 peak).  cfg 2 (37 lists of 590 columns per row, KL = 8): 7e-4 repairs x 0.36 ms = 0.25 us
// of 24 us -- accepted, and worth it (exact 20-entry lists cost 55 us there).  B = 4096 (4-6 lists per row): 16-entry
// lists would save ~6 us per call but cost 1-8 us in expected repairs -- rejected, exact lists run in the same
// deferred-insert mode.
int pick_list_len(int K, int64_t B, int64_t C, int64_t D, int lists_per_row, double share, bool allow_speculation) {
  const int exact = K <= 8 ? 8 : (K <= 12 ? 12 : (K <= 20 ? 20 : 32));
  if (!allow_speculation) return exact;
  const int cand[4] = {8, 10, 12, 16};
  for (int i = 0; i < 4; ++i) {
    const int kl = cand[i];
    if (kl >= K) break;
    const double repairs = static_cast<double>(B) * lists_per_row * overflow_bound(K, kl, share);
    const double t_repair = share * static_cast<double>(C) * (static_cast<double>(D) / 1024.0) * 0.6e-6;
    const double t_call = 2.0 * static_cast<double>(B) * static_cast<double>(C) * static_cast<double>(D) / (0.6 * 1.6e15);
    static const double max_cost = getenv("HGR_SPEC_COST") ? atof(getenv("HGR_SPEC_COST")) : 0.015;
    if (repairs * t_repair < max_cost * t_call) return kl;
  }
  return exact;
}

constexpr int kWpq = 2;  // epilogue warps per TMEM lane quarter of the production kernel

}  // namespace

// floor-sketch epilogue: [B][20] floor words, [P][B] list lengths, [P][B][kSkCap] (value, bank row) entries
static size_t sketch_workspace_bytes(int64_t B, int P) {
  return static_cast<size_t>(B) * kSkSlots * 8 + static_cast<size_t>(P) * B * 4 + static_cast<size_t>(P) * B * kSkCap * 8 + 64;
}

bool umma_supported(int64_t B, int64_t C, int64_t D, int K) {
  return B >= 1 && C >= 1 && D >= 8 && D % 8 == 0 && K >= 1 && K <= HGR_TOPK_MAX &&
         C < (int64_t(1) << 31) - 512 && B < (int64_t(1) << 31) - 512;
}

size_t umma_score_workspace_bytes(int64_t B, int64_t C, int K) {
  // worst case over the kernel variants (single CTA / CTA pair) and the list-width policy (exact lists)
  const int p1 = pick_sched(B, C, false).P * kWpq, p2 = pick_sched(B, C, true).P * 2, p3 = pick_sched_resident(B, C).P;
  const int kl = K <= 8 ? 8 : (K <= 12 ? 12 : (K <= 20 ? 20 : 32));
  const int pm = p1 > p2 ? (p1 > p3 ? p1 : p3) : (p2 > p3 ? p2 : p3);
  const size_t lists = static_cast<size_t>(pm) * B * kl * (sizeof(float) + sizeof(int32_t));
  const int p4 = make_sched(B, C, usable_sms() / 2, 2 * kTileM).P;
  const size_t sketch = sketch_workspace_bytes(B, p4 > p2 / 2 ? p4 : p2 / 2);
  return (lists > sketch ? lists : sketch) + kWsHeaderBytes;
}

// The decisions launch_score_topk_umma takes for the production variant, without launching anything.
void umma_plan(int64_t B, int64_t C, int64_t D, int K, int32_t* plan) {
  if (resident_supported(D)) {
    const Sched s = pick_sched_resident(B, C);
    const ResGeom g = resident_geom(D, kEpiTopkDefer);
    plan[0] = s.G;
    plan[1] = s.MT;
    plan[2] = s.U;
    plan[3] = s.P;
    plan[4] = pick_list_len(K, B, C, D, s.P, max_list_share(s, 1), true);
    plan[5] = 1;
    plan[6] = g.stages;
    plan[7] = static_cast<int32_t>((s.T + s.G - 1) / s.G * kUnit);
    return;
  }
  const Sched s = pick_sched(B, C, true);
  const int kl1 = pick_list_len(K, B, C, D, s.P, max_list_share(s, 1), true);
  const int wpq = pair_wpq(kl1);
  const int KL = wpq == 1 ? kl1 : pick_list_len(K, B, C, D, s.P * 2, max_list_share(s, 2), true);
  plan[0] = s.G;
  plan[1] = s.MT;
  plan[2] = s.U;
  plan[3] = s.P * wpq;
  plan[4] = KL;
  plan[5] = wpq;
  plan[6] = pair_ring_depth(B);
  plan[7] = static_cast<int32_t>((s.T + s.G - 1) / s.G * kUnit);
}

int launch_score_topk_umma(const __nv_bfloat16* X, const __nv_bfloat16* bank, const int32_t* col_id,
                           int32_t id_base, const int32_t* targets, int64_t B, int64_t C, int64_t D, float scale,
                           int K, void* ws, size_t ws_bytes, float* topk_val, int32_t* topk_idx, int64_t* hits,
                           int variant, bool skip_merge, cudaStream_t stream, const OutScatter* scatter) {
  // variants: 0 = production: resident-A CTA-pair kernel, speculative lists when provably safe
  //           2 = production kernel with exact lists; 3 = production main loop, null epilogue (ceiling; NOT a top-k)
  // variants: 0 = production: CTA-pair kernel, queue epilogue, speculative lists when provably safe
  //           1 = single-CTA kernel, reload epilogue, 1 warp/quarter, exact lists (cross-check)
  //           2 = production kernel with exact lists (no speculation)
  //           3 = production main loop with a null epilogue (ceiling; NOT a top-k) -- diagnostics only
  //           4 = single-CTA kernel, queue epilogue, speculative lists (the pre-pair production kernel)
  //           5 = single-CTA main loop with a null epilogue -- diagnostics only
  //           8 = streaming CTA-pair kernel (A re-streamed per sub-tile; the round-1 production kernel), 9 = its null epilogue
  if (ws == nullptr || ws_bytes < umma_score_workspace_bytes(B, C, K))
    return set_error(HGR_ERR_WORKSPACE, "hgr_score_topk(tcgen05): workspace %zu < %zu bytes", ws_bytes,
                     umma_score_workspace_bytes(B, C, K));
  if ((variant == 0 || variant == 2 || variant == 3) && resident_supported(D) && num_sms() >= 2) {
    const int epi = variant == 3 ? kEpiNull : kEpiTopkDefer;
    const ResGeom g = resident_geom(D, epi);
    CUtensorMap mx, mb;
    Params p{};
    int rc = make_map(&mx, X, B, D, kTileM);
    if (rc != HGR_OK) return rc;
    rc = make_map(&mb, bank, C, D, g.sub_n / 2);
    if (rc != HGR_OK) return rc;
    p.sched = pick_sched_resident(B, C);
    p.B = B;
    p.C = C;
    p.num_k_blocks = g.nkb;
    p.kb_tmem = g.kb_tmem;
    p.a_slots = g.a_slots;
    p.sub_n = g.sub_n;
    p.b_stage_bytes = g.b_stage_bytes;
    p.stage_kb = g.stage_kb;
    p.stages = g.stages;
    static const int no_prefetch = getenv("HGR_NO_PREFETCH") != nullptr;
    p.prefetch = no_prefetch ? 0 : 1;
    const int KL = pick_list_len(K, B, C, D, p.sched.P, max_list_share(p.sched, 1), variant == 0);
    p.KL = KL;
    p.scale = scale;
    p.stats = static_cast<unsigned int*>(ws);
    static const bool want_tl = getenv("HGR_TIMELINE") != nullptr;
    p.timeline = want_tl ? reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(ws) + 64) : nullptr;
    p.part_val = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + kWsHeaderBytes);
    p.part_idx = reinterpret_cast<int32_t*>(p.part_val + static_cast<size_t>(p.sched.P) * B * KL);
    rc = launch_resident_kernel(epi, KL, mx, mb, p, g, stream);
    if (rc != HGR_OK || skip_merge || variant == 3) return rc;
    MergeArgs m{};
    m.part_val = p.part_val;
    m.part_idx = p.part_idx;
    m.P = p.sched.P;
    m.B = B;
    m.KL = KL;
    m.K = K;
    m.use_sched = 1;
    m.wpq = 1;
    m.sched = p.sched;
    m.col_id = col_id;
    m.id_base = id_base;
    m.scale = scale;
    m.targets = targets;
    m.topk_val = topk_val;
    m.topk_idx = topk_idx;
    m.hits = hits;
    m.X = X;
    m.bank = bank;
    m.C = C;
    m.D8 = static_cast<int>(D / 8);
    m.rescan_count = p.stats;
    if (scatter) m.scatter = *scatter;
    return launch_topk_merge(m, stream);
  }
  if (variant == 10) {
    // streaming CTA-pair main loop + floor-sketch epilogue (sketch_epi.cuh): exact lists of <= kSkCap unsorted entries
    if (K > kSkKeep) return set_error(HGR_ERR_UNSUPPORTED, "hgr_score_topk(tcgen05 sketch): K = %d > %d", K, kSkKeep);
    CUtensorMap mx, mb;
    Params p{};
    int rc = common_setup(X, bank, B, C, D, true, &mx, &mb, &p);
    if (rc != HGR_OK) return rc;
    static const bool no_align = getenv("HGR_SK_ALIGN") == nullptr;
    if (no_align) p.sched = make_sched(B, C, usable_sms() / 2, 2 * kTileM);   // lists are cheap: every pair works
    p.scale = scale;
    p.stats = static_cast<unsigned int*>(ws);
    static const bool want_tl = getenv("HGR_TIMELINE") != nullptr;
    p.timeline = want_tl ? reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(ws) + 64) : nullptr;
    uint8_t* w = static_cast<uint8_t*>(ws) + kWsHeaderBytes;
    p.sk_floors = reinterpret_cast<unsigned long long*>(w);
    w += static_cast<size_t>(B) * kSkSlots * 8;
    p.sk_part = reinterpret_cast<uint2*>(w);
    w += static_cast<size_t>(p.sched.P) * B * kSkCap * 8;
    p.sk_cnt = reinterpret_cast<int32_t*>(w);
    rc = launch_pair_kernel(kEpiSketch, 8, 1, mx, mb, p, stream);
    if (rc != HGR_OK || skip_merge) return rc;
    MergeArgs m{};
    m.sk_part = p.sk_part;
    m.sk_cnt = p.sk_cnt;
    m.sk_cap = kSkCap;
    m.P = p.sched.P;
    m.B = B;
    m.KL = K;
    m.K = K;
    m.use_sched = 1;
    m.wpq = 1;
    m.sched = p.sched;
    m.col_id = col_id;
    m.id_base = id_base;
    m.scale = scale;
    m.targets = targets;
    m.topk_val = topk_val;
    m.topk_idx = topk_idx;
    m.hits = hits;
    m.C = C;
    if (scatter) m.scatter = *scatter;
    return launch_topk_merge(m, stream);
  }
  if (variant == 8) variant = 0;
  if (variant == 9) variant = 3;
  const bool pair = (variant == 0 || variant == 2 || variant == 3) && num_sms() >= 2;
  CUtensorMap mx, mb;
  Params p{};
  int rc = common_setup(X, bank, B, C, D, pair, &mx, &mb, &p);
  if (rc != HGR_OK) return rc;
  // the pair kernel picks its epilogue arrangement from the list length, which depends on lists/row = P * wpq
  int wpq = variant == 1 ? 1 : kWpq;
  int KL = pick_list_len(K, B, C, D, p.sched.P * wpq, max_list_share(p.sched, wpq), variant == 0 || variant == 4);
  if (pair) {
    const int kl1 = pick_list_len(K, B, C, D, p.sched.P, max_list_share(p.sched, 1), variant == 0);   // one list per (row, pair)
    wpq = pair_wpq(kl1);
    KL = wpq == 1 ? kl1 : pick_list_len(K, B, C, D, p.sched.P * 2, max_list_share(p.sched, 2), variant == 0);
  }
  const int lists = p.sched.P * wpq;
  p.KL = KL;
  p.scale = scale;
  p.stats = static_cast<unsigned int*>(ws);                              // header: statistics, timeline stamps
  static const bool want_timeline = getenv("HGR_TIMELINE") != nullptr;
  p.timeline = want_timeline ? reinterpret_cast<unsigned long long*>(static_cast<uint8_t*>(ws) + 64) : nullptr;
  p.part_val = reinterpret_cast<float*>(static_cast<uint8_t*>(ws) + kWsHeaderBytes);
  p.part_idx = reinterpret_cast<int32_t*>(p.part_val + static_cast<size_t>(lists) * B * KL);
  if (variant == 3) return launch_pair_kernel(kEpiNull, 8, wpq, mx, mb, p, stream);
  if (variant == 5) return launch_kernel<kEpiNull, 8, kWpq>(mx, mb, p, stream);
  if (pair) {
    rc = launch_pair_kernel(kEpiTopkQueue, KL, wpq, mx, mb, p, stream);
  } else {
#define HGR_UMMA_CASE(KLV)                                                                            \
  case KLV:                                                                                           \
    rc = variant == 1 ? launch_kernel<kEpiTopkReload, KLV, 1>(mx, mb, p, stream)                      \
                      : launch_kernel<kEpiTopkQueue, KLV, kWpq>(mx, mb, p, stream);                   \
    break
    switch (KL) {
      HGR_UMMA_CASE(8);
      HGR_UMMA_CASE(10);
      HGR_UMMA_CASE(12);
      HGR_UMMA_CASE(16);
      HGR_UMMA_CASE(20);
      HGR_UMMA_CASE(32);
      default:
        return set_error(HGR_ERR_UNSUPPORTED, "hgr_score_topk(tcgen05): list length %d", KL);
    }
#undef HGR_UMMA_CASE
  }
  if (rc != HGR_OK || skip_merge) return rc;
  MergeArgs m{};
  m.part_val = p.part_val;
  m.part_idx = p.part_idx;
  m.P = lists;
  m.B = B;
  m.KL = KL;
  m.K = K;
  m.use_sched = 1;
  m.wpq = wpq;
  m.sched = p.sched;
  m.col_id = col_id;
  m.id_base = id_base;
  m.scale = scale;
  m.targets = targets;
  m.topk_val = topk_val;
  m.topk_idx = topk_idx;
  m.hits = hits;
  m.X = X;
  m.bank = bank;
  m.C = C;
  m.D8 = static_cast<int>(D / 8);
  m.rescan_count = p.stats;
  if (scatter) m.scatter = *scatter;
  return launch_topk_merge(m, stream);
}

int launch_logits_umma(const __nv_bfloat16* X, const __nv_bfloat16* bank, int64_t B, int64_t C, int64_t D,
                       float scale, float* out, int64_t ldo, cudaStream_t stream) {
  if (resident_supported(D) && num_sms() >= 2 && getenv("HGR_DENSE_STREAM") == nullptr) {
    const ResGeom g = resident_geom(D, kEpiDense);
    CUtensorMap mx, mb;
    Params p{};
    int rc = make_map(&mx, X, B, D, kTileM);
    if (rc != HGR_OK) return rc;
    rc = make_map(&mb, bank, C, D, g.sub_n / 2);
    if (rc != HGR_OK) return rc;
    p.sched = pick_sched_resident(B, C);
    p.B = B;
    p.C = C;
    p.num_k_blocks = g.nkb;
    p.kb_tmem = g.kb_tmem;
    p.a_slots = g.a_slots;
    p.sub_n = g.sub_n;
    p.b_stage_bytes = g.b_stage_bytes;
    p.stage_kb = g.stage_kb;
    p.stages = g.stages;
    p.prefetch = 1;
    p.scale = scale;
    p.dense_out = out;
    p.ldo = ldo;
    return launch_resident_kernel(kEpiDense, 8, mx, mb, p, g, stream);
  }
  const bool pair = num_sms() >= 2 && getenv("HGR_DENSE_1CTA") == nullptr;
  CUtensorMap mx, mb;
  Params p{};
  int rc = common_setup(X, bank, B, C, D, pair, &mx, &mb, &p);
  if (rc != HGR_OK) return rc;
  p.KL = 0;
  p.scale = scale;
  p.dense_out = out;
  p.ldo = ldo;
  if (pair) return launch_pair_kernel(kEpiDense, 8, 2, mx, mb, p, stream);
  return launch_kernel<kEpiDense, 8, kWpq>(mx, mb, p, stream);
}

}  // namespace hgr
