// Kernel (2): TMA-fed tcgen05/TMEM bf16 GEMM for image x class cosine logits with the per-row
// running top-K fused into the epilogue (HGR_IMPL_TCGEN05) -- the B x C logit matrix never
// reaches HBM.  The same main loop with a plain store epilogue backs hgr_logits_dense.
//
// Reference: `feats @ self.zsl_weights.T` (model/clip_tree.py:331), `logits[:, test_index]`
// + `.topk(20, 1, True, True)` (main.py:136-138); the id mapping / hit test (main.py:139-147)
// runs in the merge kernel (topk_merge.cu).
//
// Structure (one persistent CTA per SM, 192 threads, warp-specialised):
//   warp 0      TMA producer: X tile [128 x 64] and bank tile [<=256 x 64] bf16 per K block,
//               128B-swizzled, 4-stage mbarrier ring;
//   warp 1      allocates TMEM (512 columns = two 128 x 256 fp32 accumulators), one lane issues
//               tcgen05.mma (M = 128, N = 16..256, K = 16) and commits to mbarriers;
//   warps 2-5   epilogue: tcgen05.ld the accumulator of sub-tile t while the tensor core
//               works on sub-tile t+1; thread = TMEM lane = image row keeps that row's sorted
//               top-K in registers across all sub-tiles of the CTA's class range.
// Work split: sched.cuh (stream-K-style split of the class dimension in units of 16 rows).
#include <cuda.h>

#include "common.cuh"
#include "ptx.cuh"
#include "sched.cuh"
#include "topk_list.cuh"

namespace hgr {
namespace {

constexpr int kBlockK = 64;                         // bf16 per K block: 128 bytes = one swizzle row
constexpr int kUmmaK = 16;                          // K of one tcgen05.mma.kind::f16
constexpr int kStages = 4;
constexpr int kABytes = kTileM * kBlockK * 2;       // 16 KB
constexpr int kBBytes = kSubN * kBlockK * 2;        // 32 KB
constexpr int kStageBytes = kABytes + kBBytes;      // 48 KB
constexpr int kBBoxRows = 64;                       // bank rows per TMA box
constexpr int kBBoxBytes = kBBoxRows * kBlockK * 2; // 8 KB
constexpr int kThreads = 192;
constexpr int kEpiWarp0 = 2;
constexpr int kEpiThreads = 128;
constexpr int kTmemCols = 512;
constexpr int kQueueDepth = 32;                     // candidate slots per epilogue thread (one chunk)
constexpr int kQueueBytes = kQueueDepth * kEpiThreads * 8;

enum EpiMode { kEpiDense = 0, kEpiTopkReload = 1, kEpiTopkQueue = 2 };

struct Ctl {
  uint64_t full[kStages];
  uint64_t empty[kStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

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
  int K;
  float scale;
  float* part_val;     // [P][B][K]
  int32_t* part_idx;   // [P][B][K] bank rows
  float* dense_out;    // [B][ldo]
  int64_t ldo;
};

template <int KL>
__device__ __forceinline__ void scan_chunk_reload(SortedList<KL>& list, const uint32_t (&r)[32], int nv,
                                                  uint32_t taddr_chunk, int col_chunk) {
  // which of my 32 values beat my current K-th best?
  const float thr = list.thr();
  uint32_t m = 0;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (__uint_as_float(r[j]) > thr) m |= (1u << j);
  if (nv < 32) m &= (1u << nv) - 1u;
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

template <int KL>
__device__ __forceinline__ void scan_chunk_queue(SortedList<KL>& list, const uint32_t (&r)[32], int nv,
                                                 int col_chunk, uint2* queue /* + epilogue thread id */) {
  // 1) lane-private compaction of the values that beat the K-th best at chunk entry
  const float thr = list.thr();
  int cnt = 0;
  if (nv >= 32) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (__uint_as_float(r[j]) > thr) {
        queue[cnt * kEpiThreads] = make_uint2(r[j], static_cast<uint32_t>(col_chunk + j));
        ++cnt;
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      if (j < nv && __uint_as_float(r[j]) > thr) {
        queue[cnt * kEpiThreads] = make_uint2(r[j], static_cast<uint32_t>(col_chunk + j));
        ++cnt;
      }
    }
  }
  // 2) dense drain: lanes walk their own queues in lock-step, so the (long) insert body is
  //    executed max_lane(cnt) times instead of once per column any lane hit
  const int maxc = __reduce_max_sync(0xffffffffu, cnt);
  for (int e = 0; e < maxc; ++e) {
    if (e < cnt) {
      const uint2 c = queue[e * kEpiThreads];
      const float x = __uint_as_float(c.x);
      if (x > list.thr()) list.insert(x, static_cast<int32_t>(c.y));
    }
  }
}

template <int EPI, int KL>
__global__ void __launch_bounds__(kThreads, 1)
score_umma_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_bank,
                  const Params p) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles must start on 1024-byte boundaries
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_base = smem;
  uint2* queue_base = reinterpret_cast<uint2*>(smem + kStages * kStageBytes);
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + kStages * kStageBytes + (EPI == kEpiTopkQueue ? kQueueBytes : 0));

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta = blockIdx.x;

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
      TileWalker walk(p.sched, cta, p.C);
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
      TileWalker walk(p.sched, cta, p.C);
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
    const int row_in_tile = quarter * 32 + lane;
    const int epi_tid = (warp - kEpiWarp0) * 32 + lane;
    uint2* queue = queue_base + epi_tid;
    TileWalker walk(p.sched, cta, p.C);
    SubTile t;
    SortedList<KL> list;
    list.init();
    int it = 0;
    while (walk.next(t)) {
      const int buf = it & 1;
      ptx::mbar_wait(&ctl->tmem_full[buf], (it >> 1) & 1);
      ptx::tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + buf * kSubN;
      const int64_t row = static_cast<int64_t>(t.mt) * kTileM + row_in_tile;
      if (EPI != kEpiDense && t.first) list.init();
      for (int c0 = 0; c0 < t.nvalid; c0 += 32) {
        uint32_t r[32];
        ptx::tmem_ld_x32(taddr + c0, r);
        ptx::tmem_ld_wait();
        const int nv = t.nvalid - c0;
        if (EPI == kEpiDense) {
          if (row < p.B) {
            float* o = p.dense_out + row * p.ldo + t.col0 + c0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (j < nv) o[j] = __uint_as_float(r[j]) * p.scale;
          }
        } else if (EPI == kEpiTopkReload) {
          scan_chunk_reload<KL>(list, r, nv, taddr + c0, t.col0 + c0);
        } else {
          scan_chunk_queue<KL>(list, r, nv, t.col0 + c0, queue);
        }
      }
      // accumulator buffer drained: hand it back to the MMA warp
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) ptx::mbar_arrive(&ctl->tmem_empty[buf]);
      if (EPI != kEpiDense && t.last && row < p.B) {
        const int slot = cta - p.sched.first_cta(t.mt);
        float* pv = p.part_val + (static_cast<int64_t>(slot) * p.B + row) * p.K;
        int32_t* pi = p.part_idx + (static_cast<int64_t>(slot) * p.B + row) * p.K;
#pragma unroll
        for (int k = 0; k < KL; ++k) {
          if (k < p.K) {
            pv[k] = list.v[k];
            pi[k] = list.i[k];
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

template <int EPI, int KL>
int launch_kernel(const CUtensorMap& mx, const CUtensorMap& mb, const Params& p, cudaStream_t stream) {
  const size_t smem = 1024 + static_cast<size_t>(kStages) * kStageBytes + (EPI == kEpiTopkQueue ? kQueueBytes : 0) +
                      sizeof(Ctl);
  auto kern = score_umma_kernel<EPI, KL>;
  HGR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
  kern<<<p.sched.G, kThreads, smem, stream>>>(mx, mb, p);
  HGR_CHECK_LAUNCH();
  return HGR_OK;
}

int common_setup(const __nv_bfloat16* X, const __nv_bfloat16* bank, int64_t B, int64_t C, int64_t D, CUtensorMap* mx,
                 CUtensorMap* mb, Params* p) {
  int rc = make_map(mx, X, B, D, kTileM);
  if (rc != HGR_OK) return rc;
  rc = make_map(mb, bank, C, D, kBBoxRows);
  if (rc != HGR_OK) return rc;
  p->sched = make_sched(B, C, num_sms());
  p->B = B;
  p->C = C;
  p->num_k_blocks = static_cast<int>((D + kBlockK - 1) / kBlockK);
  return HGR_OK;
}

}  // namespace

bool umma_supported(int64_t B, int64_t C, int64_t D, int K) {
  return B >= 1 && C >= 1 && D >= 8 && D % 8 == 0 && K >= 1 && K <= HGR_TOPK_MAX &&
         C < (int64_t(1) << 31) - 512 && B < (int64_t(1) << 31) - 512;
}

size_t umma_score_workspace_bytes(int64_t B, int64_t C, int K) {
  const Sched s = make_sched(B, C, num_sms());
  return static_cast<size_t>(s.P) * B * K * (sizeof(float) + sizeof(int32_t));
}

int launch_score_topk_umma(const __nv_bfloat16* X, const __nv_bfloat16* bank, const int32_t* col_id,
                           int32_t id_base, const int32_t* targets, int64_t B, int64_t C, int64_t D, float scale,
                           int K, void* ws, size_t ws_bytes, float* topk_val, int32_t* topk_idx, int64_t* hits,
                           bool reload_epilogue, bool skip_merge, cudaStream_t stream) {
  CUtensorMap mx, mb;
  Params p{};
  int rc = common_setup(X, bank, B, C, D, &mx, &mb, &p);
  if (rc != HGR_OK) return rc;
  const size_t need = static_cast<size_t>(p.sched.P) * B * K * (sizeof(float) + sizeof(int32_t));
  if (ws == nullptr || ws_bytes < need)
    return set_error(HGR_ERR_WORKSPACE, "hgr_score_topk(tcgen05): workspace %zu < %zu bytes", ws_bytes, need);
  p.K = K;
  p.scale = scale;
  p.part_val = static_cast<float*>(ws);
  p.part_idx = reinterpret_cast<int32_t*>(p.part_val + static_cast<size_t>(p.sched.P) * B * K);
  const int mode = reload_epilogue ? kEpiTopkReload : kEpiTopkQueue;
#define HGR_UMMA_LAUNCH(KL)                                                                  \
  rc = mode == kEpiTopkReload ? launch_kernel<kEpiTopkReload, KL>(mx, mb, p, stream)         \
                              : launch_kernel<kEpiTopkQueue, KL>(mx, mb, p, stream)
  if (K <= 8) HGR_UMMA_LAUNCH(8);
  else if (K <= 20) HGR_UMMA_LAUNCH(20);
  else HGR_UMMA_LAUNCH(32);
#undef HGR_UMMA_LAUNCH
  if (rc != HGR_OK || skip_merge) return rc;
  return launch_topk_merge(p.part_val, p.part_idx, p.sched.P, B, K, 0, &p.sched, col_id, id_base, scale, targets,
                           topk_val, topk_idx, hits, stream);
}

int launch_logits_umma(const __nv_bfloat16* X, const __nv_bfloat16* bank, int64_t B, int64_t C, int64_t D,
                       float scale, float* out, int64_t ldo, cudaStream_t stream) {
  CUtensorMap mx, mb;
  Params p{};
  int rc = common_setup(X, bank, B, C, D, &mx, &mb, &p);
  if (rc != HGR_OK) return rc;
  p.K = 0;
  p.scale = scale;
  p.dense_out = out;
  p.ldo = ldo;
  return launch_kernel<kEpiDense, 8>(mx, mb, p, stream);
}

}  // namespace hgr
