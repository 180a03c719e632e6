// extern "C" surface of libhgr_b200.so (declared in include/hgr_b200.h): argument checks,
// implementation selection and error reporting.  No device memory is allocated here.
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "sched.cuh"

namespace hgr {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int num_sms() {
  static int cached = [] {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      return kNumSMsB200;  // no device visible (build box): size queries assume a B200
    }
    return n;
  }();
  return cached;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

static int pick_impl(int impl, int64_t B, int64_t C, int64_t D, int K) {
  if (impl == HGR_IMPL_AUTO) {
    const char* e = getenv("HGR_IMPL");
    if (e && !strcmp(e, "simt")) return HGR_IMPL_SIMT;
    return umma_supported(B, C, D, K) ? HGR_IMPL_TCGEN05 : HGR_IMPL_SIMT;
  }
  return impl;
}

}  // namespace hgr

using namespace hgr;

extern "C" {

int hgr_version(void) { return HGR_ABI_VERSION; }

const char* hgr_last_error(void) { return g_err; }

int64_t hgr_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }

int hgr_aggregate_normalize(const void* E, int e_dtype, int64_t n_src, int64_t D, const int32_t* rowptr,
                            const int32_t* col, const float* w, int64_t n_rows, const int32_t* row_map,
                            int64_t n_out, void* out, int out_dtype, float* out_norm, void* stream) {
  HGR_CHECK_ARG(n_src >= 0 && n_rows >= 0 && n_out >= 0, "hgr_aggregate_normalize: negative size");
  HGR_CHECK_ARG(D > 0 && D % 8 == 0, "hgr_aggregate_normalize: D = %lld must be a positive multiple of 8", (long long)D);
  if (n_out == 0) return HGR_OK;
  HGR_CHECK_ARG(E && out, "hgr_aggregate_normalize: null E/out");
  HGR_CHECK_ARG(aligned16(E) && aligned16(out), "hgr_aggregate_normalize: E/out must be 16-byte aligned");
  HGR_CHECK_ARG(rowptr == nullptr || col != nullptr, "hgr_aggregate_normalize: rowptr given without col");
  HGR_CHECK_ARG(rowptr != nullptr || n_rows == n_src, "hgr_aggregate_normalize: identity CSR needs n_rows == n_src");
  HGR_CHECK_ARG(row_map != nullptr || n_out == n_rows, "hgr_aggregate_normalize: n_out != n_rows without row_map");
  HGR_CHECK_ARG(n_src < (int64_t(1) << 31) && n_rows < (int64_t(1) << 31), "hgr_aggregate_normalize: > 2^31 rows");
  return launch_aggregate_normalize(E, e_dtype, n_src, D, rowptr, col, w, row_map, n_out, out, out_dtype, out_norm,
                                    static_cast<cudaStream_t>(stream));
}

int hgr_normalize_rows_dual(const void* E, int e_dtype, int64_t n_rows, int64_t D, void* out, const int32_t* dst_map,
                            void* out2, void* stream) {
  HGR_CHECK_ARG(n_rows >= 0, "hgr_normalize_rows_dual: negative size");
  HGR_CHECK_ARG(D > 0 && D % 8 == 0, "hgr_normalize_rows_dual: D = %lld must be a positive multiple of 8", (long long)D);
  if (n_rows == 0) return HGR_OK;
  HGR_CHECK_ARG(E && out, "hgr_normalize_rows_dual: null E/out");
  HGR_CHECK_ARG((dst_map == nullptr) == (out2 == nullptr), "hgr_normalize_rows_dual: dst_map and out2 go together");
  HGR_CHECK_ARG(aligned16(E) && aligned16(out) && aligned16(out2), "hgr_normalize_rows_dual: E/out/out2 must be 16-byte aligned");
  return launch_normalize_dual(E, e_dtype, n_rows, D, out, dst_map, out2, static_cast<cudaStream_t>(stream));
}

size_t hgr_score_topk_workspace_bytes(int64_t B, int64_t C, int64_t D, int K) {
  if (B <= 0 || C <= 0 || K <= 0) return 16;
  size_t a = simt_score_workspace_bytes(B, C, K);
  size_t b = umma_supported(B, C, D, K) ? umma_score_workspace_bytes(B, C, K) : 0;
  return (a > b ? a : b) + 16;
}

static int score_topk_impl(const void* X, const void* bank, const int32_t* col_id, int32_t id_base, const int32_t* targets,
                           int64_t B, int64_t C, int64_t D, float scale, int K, void* workspace, size_t workspace_bytes,
                           float* topk_val, int32_t* topk_idx, int64_t* hits, int impl, void* stream,
                           const OutScatter* scatter, int64_t C_total = 0) {
  HGR_CHECK_ARG(B >= 0 && C >= 0, "hgr_score_topk: negative size");
  HGR_CHECK_ARG(D > 0 && D % 8 == 0, "hgr_score_topk: D = %lld must be a positive multiple of 8", (long long)D);
  HGR_CHECK_ARG(K >= 1 && K <= HGR_TOPK_MAX, "hgr_score_topk: K = %d outside [1, %d]", K, HGR_TOPK_MAX);
  HGR_CHECK_ARG(scale > 0.f, "hgr_score_topk: scale must be > 0 (top-k order is taken on unscaled cosines)");
  if (B == 0) return HGR_OK;
  HGR_CHECK_ARG(scatter || (topk_val && topk_idx), "hgr_score_topk: null output");
  HGR_CHECK_ARG(C == 0 || (X && bank), "hgr_score_topk: null X/bank");
  HGR_CHECK_ARG(aligned16(X) && aligned16(bank) && aligned16(workspace), "hgr_score_topk: X/bank/workspace must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (C == 0) {
    // empty class set: every list is (-inf, -1); the merge of zero lists writes exactly that
    MergeArgs m{};
    m.B = B;
    m.KL = K;
    m.K = K;
    m.scale = 1.f;
    m.topk_val = topk_val;
    m.topk_idx = topk_idx;
    if (scatter) m.scatter = *scatter;
    return launch_topk_merge(m, s);
  }
  const bool skip_merge = (impl & HGR_IMPL_FLAG_NO_MERGE) != 0;
  impl &= ~HGR_IMPL_FLAG_NO_MERGE;
  const int which = pick_impl(impl, B, C, D, K);
  if (which == HGR_IMPL_TCGEN05 || which == HGR_IMPL_TCGEN05_EXACT || which == HGR_IMPL_TCGEN05_NULL ||
      which == HGR_IMPL_TCGEN05_SKETCH) {
    if (!umma_supported(B, C, D, K)) return set_error(HGR_ERR_UNSUPPORTED, "hgr_score_topk: shape not supported by the tcgen05 kernel");
    const int variant = which - HGR_IMPL_TCGEN05;  // umma::Variant
    return launch_score_topk_umma(static_cast<const __nv_bfloat16*>(X), static_cast<const __nv_bfloat16*>(bank), col_id,
                                  id_base, targets, B, C, D, scale, K, workspace, workspace_bytes, topk_val, topk_idx,
                                  hits, variant, skip_merge || which == HGR_IMPL_TCGEN05_NULL, s, scatter, C_total);
  }
  if (which == HGR_IMPL_SIMT)
    return launch_score_topk_simt(static_cast<const __nv_bfloat16*>(X), static_cast<const __nv_bfloat16*>(bank), col_id,
                                  id_base, targets, B, C, D, scale, K, workspace, workspace_bytes, topk_val, topk_idx,
                                  hits, s, scatter);
  return set_error(HGR_ERR_BAD_ARG, "hgr_score_topk: unknown impl %d", impl);
}

int hgr_score_topk_plan(int64_t B, int64_t C, int64_t D, int K, int32_t* plan) {
  HGR_CHECK_ARG(plan != nullptr, "hgr_score_topk_plan: null plan");
  HGR_CHECK_ARG(K >= 1 && K <= HGR_TOPK_MAX, "hgr_score_topk_plan: K = %d outside [1, %d]", K, HGR_TOPK_MAX);
  if (!umma_supported(B, C, D, K)) return set_error(HGR_ERR_UNSUPPORTED, "hgr_score_topk_plan: shape not supported by the tcgen05 kernel");
  umma_plan(B, C, D, K, plan);
  return HGR_OK;
}

int hgr_score_topk(const void* X, const void* bank, const int32_t* col_id, int32_t id_base, const int32_t* targets,
                   int64_t B, int64_t C, int64_t D, float scale, int K, void* workspace, size_t workspace_bytes,
                   float* topk_val, int32_t* topk_idx, int64_t* hits, int impl, void* stream) {
  return score_topk_impl(X, bank, col_id, id_base, targets, B, C, D, scale, K, workspace, workspace_bytes, topk_val,
                         topk_idx, hits, impl, stream, nullptr);
}

static int score_topk_scatter_impl(const void* X, const void* bank, const int32_t* col_id, int32_t id_base, int64_t B,
                                   int64_t C, int64_t D, float scale, int K, void* workspace, size_t workspace_bytes,
                                   int64_t block_rows, int n_blocks, float* const* val_blocks, int32_t* const* idx_blocks,
                                   float* const* bound_blocks, int64_t C_total, int impl, void* stream) {
  HGR_CHECK_ARG(n_blocks >= 1 && n_blocks <= kMaxScatterBlocks, "hgr_score_topk_scatter: n_blocks = %d outside [1, %d]",
                n_blocks, kMaxScatterBlocks);
  HGR_CHECK_ARG(block_rows >= 1 && block_rows * n_blocks >= B, "hgr_score_topk_scatter: %d blocks of %lld rows do not cover B = %lld",
                n_blocks, (long long)block_rows, (long long)B);
  HGR_CHECK_ARG(val_blocks && idx_blocks, "hgr_score_topk_scatter: null block tables");
  // (NO_MERGE with bounds: diagnostics -- the scoring kernel alone with the list length of the global certificate;
  //  nothing is scattered, the block tables are not dereferenced)
  HGR_CHECK_ARG((impl & HGR_IMPL_FLAG_NO_MERGE) == 0 || bound_blocks != nullptr, "hgr_score_topk_scatter: NO_MERGE makes no sense here");
  OutScatter sc;
  sc.n_blocks = n_blocks;
  sc.block_rows = block_rows;
  sc.emit_bound = bound_blocks != nullptr;
  for (int g = 0; g < n_blocks; ++g) {
    HGR_CHECK_ARG(val_blocks[g] && idx_blocks[g], "hgr_score_topk_scatter: null block %d", g);
    HGR_CHECK_ARG(bound_blocks == nullptr || bound_blocks[g], "hgr_score_topk_scatter_bounded: null bound block %d", g);
    sc.val[g] = val_blocks[g];
    sc.idx[g] = idx_blocks[g];
    sc.bound[g] = bound_blocks ? bound_blocks[g] : nullptr;
  }
  return score_topk_impl(X, bank, col_id, id_base, nullptr, B, C, D, scale, K, workspace, workspace_bytes, nullptr, nullptr,
                         nullptr, impl, stream, &sc, C_total);
}

int hgr_score_topk_scatter(const void* X, const void* bank, const int32_t* col_id, int32_t id_base, int64_t B, int64_t C,
                           int64_t D, float scale, int K, void* workspace, size_t workspace_bytes, int64_t block_rows,
                           int n_blocks, float* const* val_blocks, int32_t* const* idx_blocks, int impl, void* stream) {
  return score_topk_scatter_impl(X, bank, col_id, id_base, B, C, D, scale, K, workspace, workspace_bytes, block_rows,
                                 n_blocks, val_blocks, idx_blocks, nullptr, 0, impl, stream);
}

int hgr_score_topk_scatter_bounded(const void* X, const void* bank, const int32_t* col_id, int32_t id_base, int64_t B,
                                   int64_t C, int64_t D, float scale, int K, void* workspace, size_t workspace_bytes,
                                   int64_t block_rows, int n_blocks, float* const* val_blocks, int32_t* const* idx_blocks,
                                   float* const* bound_blocks, int64_t C_total, int impl, void* stream) {
  HGR_CHECK_ARG(bound_blocks != nullptr, "hgr_score_topk_scatter_bounded: null bound table");
  HGR_CHECK_ARG(C_total >= C, "hgr_score_topk_scatter_bounded: C_total = %lld < C = %lld", (long long)C_total, (long long)C);
  return score_topk_scatter_impl(X, bank, col_id, id_base, B, C, D, scale, K, workspace, workspace_bytes, block_rows,
                                 n_blocks, val_blocks, idx_blocks, bound_blocks, C_total, impl, stream);
}

int hgr_score_topk_global_list_len(int64_t B, int64_t C, int64_t D, int K, int64_t C_total) {
  HGR_CHECK_ARG(K >= 1 && K <= HGR_TOPK_MAX, "hgr_score_topk_global_list_len: K = %d outside [1, %d]", K, HGR_TOPK_MAX);
  if (!umma_supported(B, C, D, K)) return set_error(HGR_ERR_UNSUPPORTED, "hgr_score_topk_global_list_len: shape not supported by the tcgen05 kernel");
  return umma_global_list_len(B, C, D, K, C_total);
}

int hgr_peer_alloc(size_t bytes, void** ptr, unsigned char* handle) {
  HGR_CHECK_ARG(ptr && handle && bytes > 0, "hgr_peer_alloc: bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == HGR_IPC_HANDLE_BYTES, "IPC handle size");
  void* p = nullptr;
  HGR_CHECK_CUDA(cudaMalloc(&p, bytes));
  HGR_CHECK_CUDA(cudaMemset(p, 0, bytes));
  HGR_CHECK_CUDA(cudaDeviceSynchronize());
  cudaIpcMemHandle_t h;
  cudaError_t e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    return set_error(HGR_ERR_CUDA, "hgr_peer_alloc: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  }
  memcpy(handle, &h, sizeof(h));
  *ptr = p;
  return HGR_OK;
}

int hgr_peer_open(const unsigned char* handle, void** ptr) {
  HGR_CHECK_ARG(ptr && handle, "hgr_peer_open: bad argument");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  HGR_CHECK_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return HGR_OK;
}

int hgr_peer_close(void* ptr) {
  if (ptr) HGR_CHECK_CUDA(cudaIpcCloseMemHandle(ptr));
  return HGR_OK;
}

int hgr_peer_free(void* ptr) {
  if (ptr) HGR_CHECK_CUDA(cudaFree(ptr));
  return HGR_OK;
}

int hgr_normalize_rows_bcast(const void* E, int e_dtype, int64_t n_rows, int64_t D, int64_t row0, int n_dst,
                             void* const* dst, void* stream) {
  HGR_CHECK_ARG(n_rows >= 0 && row0 >= 0, "hgr_normalize_rows_bcast: negative size");
  HGR_CHECK_ARG(D > 0 && D % 8 == 0, "hgr_normalize_rows_bcast: D = %lld must be a positive multiple of 8", (long long)D);
  HGR_CHECK_ARG(n_dst >= 1 && n_dst <= kMaxScatterBlocks, "hgr_normalize_rows_bcast: n_dst = %d outside [1, %d]", n_dst,
                kMaxScatterBlocks);
  if (n_rows == 0) return HGR_OK;
  HGR_CHECK_ARG(E && dst, "hgr_normalize_rows_bcast: null pointer");
  for (int g = 0; g < n_dst; ++g)
    HGR_CHECK_ARG(dst[g] && aligned16(dst[g]), "hgr_normalize_rows_bcast: destination %d null or not 16-byte aligned", g);
  HGR_CHECK_ARG(aligned16(E), "hgr_normalize_rows_bcast: E must be 16-byte aligned");
  return launch_normalize_bcast(E, e_dtype, n_rows, D, row0, n_dst, dst, static_cast<cudaStream_t>(stream));
}

int hgr_peer_signal(uint32_t* const* flags, int n, uint32_t* seq, void* stream) {
  HGR_CHECK_ARG(flags && seq, "hgr_peer_signal: null pointer");
  return launch_peer_signal(flags, n, seq, static_cast<cudaStream_t>(stream));
}

int hgr_peer_wait(const uint32_t* flags, int n, uint32_t* seq, void* stream) {
  HGR_CHECK_ARG(flags && seq, "hgr_peer_wait: null pointer");
  return launch_peer_wait(flags, n, seq, static_cast<cudaStream_t>(stream));
}

int hgr_topk_merge(const float* part_val, const int32_t* part_idx, int64_t P, int64_t B, int K, int64_t part_stride,
                   const int32_t* targets, float* topk_val, int32_t* topk_idx, int64_t* hits, void* stream) {
  HGR_CHECK_ARG(P >= 0 && B >= 0, "hgr_topk_merge: negative size");
  HGR_CHECK_ARG(K >= 1 && K <= HGR_TOPK_MAX, "hgr_topk_merge: K = %d outside [1, %d]", K, HGR_TOPK_MAX);
  if (B == 0) return HGR_OK;
  HGR_CHECK_ARG(topk_val && topk_idx, "hgr_topk_merge: null output");
  HGR_CHECK_ARG(P == 0 || (part_val && part_idx), "hgr_topk_merge: null parts");
  HGR_CHECK_ARG(part_stride == 0 || part_stride >= B * K, "hgr_topk_merge: part_stride < B*K");
  MergeArgs m{};
  m.part_val = part_val;
  m.part_idx = part_idx;
  m.P = P;
  m.B = B;
  m.KL = K;
  m.K = K;
  m.part_stride = part_stride;
  m.scale = 1.f;
  m.targets = targets;
  m.topk_val = topk_val;
  m.topk_idx = topk_idx;
  m.hits = hits;
  return launch_topk_merge(m, static_cast<cudaStream_t>(stream));
}

int hgr_topk_merge_certified(const float* part_val, const int32_t* part_idx, const float* part_bound, int64_t P, int64_t B,
                             int K, int64_t part_stride, int64_t bound_stride, const int32_t* targets, float* topk_val,
                             int32_t* topk_idx, int64_t* hits, const void* X, int64_t D, const hgr_shard_t* shards,
                             float scale, unsigned int* repair_count, void* stream) {
  HGR_CHECK_ARG(P >= 1 && P <= kMaxScatterBlocks && B >= 0, "hgr_topk_merge_certified: P = %lld outside [1, %d]", (long long)P,
                kMaxScatterBlocks);
  HGR_CHECK_ARG(K >= 1 && K <= HGR_TOPK_MAX, "hgr_topk_merge_certified: K = %d outside [1, %d]", K, HGR_TOPK_MAX);
  if (B == 0) return HGR_OK;
  HGR_CHECK_ARG(topk_val && topk_idx && part_val && part_idx && part_bound, "hgr_topk_merge_certified: null lists / bounds / output");
  HGR_CHECK_ARG(X && shards && D > 0 && D % 8 == 0 && aligned16(X), "hgr_topk_merge_certified: the repair needs X (16-byte aligned, D %% 8 == 0) and the shard table");
  HGR_CHECK_ARG(part_stride == 0 || part_stride >= B * K, "hgr_topk_merge_certified: part_stride < B*K");
  HGR_CHECK_ARG(bound_stride >= B, "hgr_topk_merge_certified: bound_stride < B");
  HGR_CHECK_ARG(scale > 0.f, "hgr_topk_merge_certified: scale must be > 0");
  MergeArgs m{};
  m.part_val = part_val;
  m.part_idx = part_idx;
  m.P = P;
  m.B = B;
  m.KL = K;
  m.K = K;
  m.part_stride = part_stride;
  m.scale = 1.f;
  m.targets = targets;
  m.topk_val = topk_val;
  m.topk_idx = topk_idx;
  m.hits = hits;
  m.part_bound = part_bound;
  m.bound_stride = bound_stride;
  m.shards = shards;
  m.xrows = X;
  m.xD8 = static_cast<int>(D / 8);
  m.shard_scale = scale;
  m.repair_count = repair_count;
  return launch_topk_merge(m, static_cast<cudaStream_t>(stream));
}

int64_t hgr_sample_replay(const uint32_t* words, int64_t n_words, int64_t n, int64_t k, int64_t setsize, int32_t* out_pos,
                          int32_t* scratch) {
  if (!words || !out_pos || !scratch || n < 0 || k < 0 || k > n || n >= (int64_t(1) << 31)) return -2;
  auto bit_length = [](uint32_t v) { return v == 0 ? 0 : 32 - __builtin_clz(v); };
  int64_t p = 0;
  if (n <= setsize) {
    for (int64_t i = 0; i < n; ++i) scratch[i] = static_cast<int32_t>(i);
    for (int64_t i = 0; i < k; ++i) {
      const uint32_t left = static_cast<uint32_t>(n - i);
      const int sh = 32 - bit_length(left);
      uint32_t r;
      do {
        if (p >= n_words) return -1;
        r = words[p++] >> sh;
      } while (r >= left);
      out_pos[i] = scratch[r];
      scratch[r] = scratch[left - 1];
    }
    return p;
  }
  for (int64_t i = 0; i < n; ++i) scratch[i] = 0;
  const int sh = 32 - bit_length(static_cast<uint32_t>(n));
  for (int64_t i = 0; i < k; ++i) {
    uint32_t r;
    do {
      do {
        if (p >= n_words) return -1;
        r = words[p++] >> sh;
      } while (r >= static_cast<uint32_t>(n));
    } while (scratch[r]);
    scratch[r] = 1;
    out_pos[i] = static_cast<int32_t>(r);
  }
  return p;
}

int64_t hgr_sample_replay_many(const uint32_t* words, int64_t n_words, int64_t count, const int64_t* n, const int64_t* k,
                               int32_t* out_pos, int32_t* scratch) {
  if (!n || !k || count < 0) return -2;
  int64_t used = 0, off = 0;
  for (int64_t c = 0; c < count; ++c) {
    int64_t setsize = 21;                       // CPython: 21, + 4 ** ceil(log(3 k, 4)) when k > 5
    if (k[c] > 5) {
      int64_t p4 = 1;
      while (p4 < 3 * k[c]) p4 *= 4;
      setsize += p4;
    }
    const int64_t r = hgr_sample_replay(words + used, n_words - used, n[c], k[c], setsize, out_pos + off, scratch);
    if (r < 0) return r;
    used += r;
    off += k[c];
  }
  return used;
}

int64_t hgr_om_plan(const uint32_t* words, int64_t n_words, int64_t T, const int64_t* const* cand, const int64_t* n_cand,
                    const int64_t* anchor, int64_t num_compare, int64_t n_nodes, int32_t* set_ptr, int32_t* set_col,
                    int32_t* label_pos, int32_t* union_ids, int64_t* counts, int32_t* scratch) {
  if (!cand || !n_cand || !anchor || !set_ptr || !set_col || !label_pos || !union_ids || !counts || !scratch || T < 0 ||
      num_compare < 0 || n_nodes <= 0 || n_nodes >= (int64_t(1) << 31))
    return -2;
  int64_t max_n = 0;
  for (int64_t t = 0; t < T; ++t) {
    if (n_cand[t] < 0 || (n_cand[t] > 0 && !cand[t]) || anchor[t] < 0 || anchor[t] >= n_nodes) return -2;
    max_n = n_cand[t] > max_n ? n_cand[t] : max_n;
  }
  int32_t* mark = scratch;                 // [n_nodes]: 0, then union position + 1
  int32_t* pool = scratch + n_nodes;       // [max_n]
  int32_t* pos = pool + max_n;             // [num_compare]
  for (int64_t i = 0; i < n_nodes; ++i) mark[i] = 0;
  int64_t setsize = 21;                    // CPython: 21, + 4 ** ceil(log(3 k, 4)) when k > 5
  if (num_compare > 5) {
    int64_t p4 = 1;
    while (p4 < 3 * num_compare) p4 *= 4;
    setsize += p4;
  }
  int64_t used = 0;
  set_ptr[0] = 0;
  for (int64_t t = 0; t < T; ++t) {
    int32_t* out = set_col + set_ptr[t];
    int64_t m = 0;
    if (n_cand[t] > num_compare) {
      if (words == nullptr) return -2;
      const int64_t r = hgr_sample_replay(words + used, n_words - used, n_cand[t], num_compare, setsize, pos, pool);
      if (r < 0) return r;
      used += r;
      for (int64_t i = 0; i < num_compare; ++i) out[i] = static_cast<int32_t>(cand[t][pos[i]]);
      m = num_compare;
    } else {
      for (int64_t i = 0; i < n_cand[t]; ++i) out[i] = static_cast<int32_t>(cand[t][i]);
      m = n_cand[t];
    }
    int64_t lp = -1;
    for (int64_t i = 0; i < m && lp < 0; ++i)
      if (out[i] == anchor[t]) lp = i;
    if (lp < 0) {
      out[m] = static_cast<int32_t>(anchor[t]);
      lp = m++;
    }
    label_pos[t] = static_cast<int32_t>(lp);
    for (int64_t i = 0; i < m; ++i) {
      if (out[i] < 0 || out[i] >= n_nodes) return -2;
      mark[out[i]] = 1;
    }
    set_ptr[t + 1] = set_ptr[t] + static_cast<int32_t>(m);
  }
  int64_t nu = 0;
  for (int64_t i = 0; i < n_nodes; ++i)
    if (mark[i]) {
      union_ids[nu] = static_cast<int32_t>(i);
      mark[i] = static_cast<int32_t>(++nu);
    }
  const int64_t n_col = set_ptr[T];
  for (int64_t e = 0; e < n_col; ++e) set_col[e] = mark[set_col[e]] - 1;
  counts[0] = n_col;
  counts[1] = nu;
  return used;
}

int hgr_logits_dense(const void* X, const void* bank, int64_t B, int64_t C, int64_t D, float scale, float* out,
                     int64_t ldo, int impl, void* stream) {
  HGR_CHECK_ARG(B >= 0 && C >= 0, "hgr_logits_dense: negative size");
  HGR_CHECK_ARG(D > 0 && D % 8 == 0, "hgr_logits_dense: D = %lld must be a positive multiple of 8", (long long)D);
  if (B == 0 || C == 0) return HGR_OK;
  HGR_CHECK_ARG(X && bank && out, "hgr_logits_dense: null pointer");
  HGR_CHECK_ARG(ldo >= C, "hgr_logits_dense: ldo < C");
  HGR_CHECK_ARG(aligned16(X) && aligned16(bank), "hgr_logits_dense: X/bank must be 16-byte aligned");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int which = pick_impl(impl, B, C, D, 1);
  if (which == HGR_IMPL_TCGEN05)
    return launch_logits_umma(static_cast<const __nv_bfloat16*>(X), static_cast<const __nv_bfloat16*>(bank), B, C, D,
                              scale, out, ldo, s);
  if (which == HGR_IMPL_SIMT)
    return launch_logits_simt(static_cast<const __nv_bfloat16*>(X), static_cast<const __nv_bfloat16*>(bank), B, C, D,
                              scale, out, ldo, s);
  return set_error(HGR_ERR_BAD_ARG, "hgr_logits_dense: unknown impl %d", impl);
}

int hgr_hier_metrics(const float* logits, int64_t ldl, int64_t B, int64_t N, const int32_t* cols, int64_t M,
                     const int8_t* level, int n_levels, const int32_t* first_out, const int32_t* chain,
                     const int32_t* chain_level, int L, int32_t* lvl_idx, int32_t* top1, int64_t* counts, void* stream) {
  HGR_CHECK_ARG(B >= 0 && N > 0 && M > 0 && M <= (int64_t(1) << 31) - 1, "hgr_hier_metrics: bad sizes B=%lld N=%lld M=%lld",
                (long long)B, (long long)N, (long long)M);
  HGR_CHECK_ARG(ldl >= N, "hgr_hier_metrics: ldl < N");
  HGR_CHECK_ARG(L >= 1 && L <= 64, "hgr_hier_metrics: chain length %d outside [1, 64]", L);
  if (B == 0) return HGR_OK;
  HGR_CHECK_ARG(logits && level && first_out && chain && chain_level && counts, "hgr_hier_metrics: null pointer");
  HGR_CHECK_ARG(cols != nullptr || M == N, "hgr_hier_metrics: cols == NULL needs M == N");
  return launch_hier_metrics(logits, ldl, B, cols, M, level, n_levels, first_out, chain, chain_level, L, lvl_idx, top1,
                             counts, static_cast<cudaStream_t>(stream));
}

int hgr_hier_metrics_fused(const void* X, const void* bank_sorted, int64_t B, int64_t M, int64_t D, const int32_t* level_end,
                           int n_levels, const int32_t* sorted_to_pos, const int32_t* first_out, const int32_t* chain,
                           const int32_t* chain_level, int L, void* workspace, size_t workspace_bytes, int32_t* lvl_idx,
                           int32_t* top1, int64_t* counts, void* stream) {
  HGR_CHECK_ARG(B >= 0 && M > 0 && M <= (int64_t(1) << 31) - 512, "hgr_hier_metrics_fused: bad sizes B=%lld M=%lld", (long long)B,
                (long long)M);
  HGR_CHECK_ARG(D > 0 && D % 8 == 0, "hgr_hier_metrics_fused: D = %lld must be a positive multiple of 8", (long long)D);
  HGR_CHECK_ARG(L >= 1 && L <= 64, "hgr_hier_metrics_fused: chain length %d outside [1, 64]", L);
  HGR_CHECK_ARG(n_levels >= 1 && n_levels <= 32, "hgr_hier_metrics_fused: %d levels outside [1, 32]", n_levels);
  if (B == 0) return HGR_OK;
  HGR_CHECK_ARG(X && bank_sorted && level_end && sorted_to_pos && first_out && chain && chain_level && counts && workspace,
                "hgr_hier_metrics_fused: null pointer");
  HGR_CHECK_ARG(aligned16(X) && aligned16(bank_sorted) && aligned16(workspace), "hgr_hier_metrics_fused: X / bank / workspace must be 16-byte aligned");
  HGR_CHECK_ARG(workspace_bytes >= static_cast<size_t>(B) * n_levels * 8, "hgr_hier_metrics_fused: workspace %zu < %zu bytes",
                workspace_bytes, static_cast<size_t>(B) * n_levels * 8);
  if (!umma_supported(B, M, D, 1)) return set_error(HGR_ERR_UNSUPPORTED, "hgr_hier_metrics_fused: shape not supported by the tcgen05 kernel");
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  unsigned long long* best = static_cast<unsigned long long*>(workspace);
  const int rc = launch_level_argmax_umma(static_cast<const __nv_bfloat16*>(X), static_cast<const __nv_bfloat16*>(bank_sorted), B, M,
                                          D, level_end, n_levels, best, s);
  if (rc != HGR_OK) return rc;
  return launch_hier_finish(best, B, M, n_levels, sorted_to_pos, first_out, chain, chain_level, L, lvl_idx, top1, counts, s);
}

size_t hgr_masked_ce_workspace_bytes(int64_t B, int64_t U, int64_t T) { return masked_ce_workspace_bytes(B, U, T) + 16; }

int hgr_masked_ce(const float* logits, int64_t ldl, int64_t B, int64_t U, const int32_t* set_ptr, const int32_t* set_col,
                  const int32_t* label_pos, const float* weight, int64_t T, float* loss, float* dlogits, void* workspace,
                  size_t workspace_bytes, void* stream) {
  HGR_CHECK_ARG(B > 0 && U > 0 && T >= 0, "hgr_masked_ce: bad sizes B=%lld U=%lld T=%lld", (long long)B, (long long)U, (long long)T);
  HGR_CHECK_ARG(ldl >= U, "hgr_masked_ce: ldl < U");
  if (T == 0) return HGR_OK;
  HGR_CHECK_ARG(logits && set_ptr && set_col && label_pos && weight && loss, "hgr_masked_ce: null pointer");
  return launch_masked_ce(logits, ldl, B, U, set_ptr, set_col, label_pos, weight, T, loss, dlogits, workspace,
                          workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t hgr_om_backward_workspace_bytes(int64_t B, int64_t U, int64_t D) { return om_backward_workspace_bytes(B, U, D) + 16; }

int hgr_om_backward(const float* dlogits, const float* logits, int64_t ldl, int64_t B, int64_t U, int64_t D, const void* x,
                    const float* x_norm, const void* tn, const float* t_norm, float scale, float* d_img, float* d_text,
                    float* d_log_scale, void* workspace, size_t workspace_bytes, void* stream) {
  HGR_CHECK_ARG(B > 0 && U > 0 && D > 0, "hgr_om_backward: bad sizes B=%lld U=%lld D=%lld", (long long)B, (long long)U, (long long)D);
  HGR_CHECK_ARG(ldl >= U, "hgr_om_backward: ldl < U");
  HGR_CHECK_ARG(dlogits && logits && x && x_norm && tn && t_norm && d_img && d_text, "hgr_om_backward: null pointer");
  HGR_CHECK_ARG(aligned16(x) && aligned16(tn) && aligned16(d_img) && aligned16(d_text) && aligned16(workspace),
                "hgr_om_backward: x / tn / d_img / d_text / workspace must be 16-byte aligned");
  return launch_om_backward(dlogits, logits, ldl, B, U, D, static_cast<const __nv_bfloat16*>(x), x_norm,
                            static_cast<const __nv_bfloat16*>(tn), t_norm, scale, d_img, d_text, d_log_scale, workspace,
                            workspace_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
