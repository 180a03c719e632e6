// TOR / POR hierarchical metrics of one single-label batch in ONE pass over the dense logits.
//
// Reference: main.py:143,152-191.  Per batch the reference takes the arg-max over `train_index` (TOR, :155-160)
// and, for every node p of the label's ancestor chain, clones the [B, N] logits, fills every column that is not at
// p's depth with -1, gathers `train_index` and takes another arg-max (:163-176) -- L + 1 passes over B x N plus a
// Python B x L loop with device->host copies (:177-191).  Here one CTA per image row walks the train columns once,
// keeps a running (value, position) arg-max PER DEPTH LEVEL (the depth of a column is one int8 look-up), and then
// counts chain matches itself:
//   counts[0] += #{k : top1 == chain[k]}                                   -> hits_all          (main.py:158-160)
//   counts[1] += #{k : lvl[depth(chain[k])] == chain[k]}                   -> point             (main.py:182-189)
//   counts[2] += #{k < L-1 : match[k] and match[k+1]}  (L == 1: match[0])  -> edge / single-node path (main.py:179-185)
// HBM-bound: B * M * 4 bytes of logits read once (+ M * 5 bytes of column ids / levels from L2).
//
// Tie rule = first position (what an arg-max over the masked array returns on the CPU): among equal values the
// smallest position j in `cols` wins; a masked column holds exactly -1.0f, so when no in-level value exceeds -1 the
// first position holding -1 wins -- `first_out[l]` is the first position whose column is NOT at level l.
#include "common.cuh"

namespace hgr {
namespace {

constexpr int kHierUnroll = 8;  // independent (column id -> level, logit) load chains in flight per thread
constexpr int kMaxLevels = 32;   // <= 16 levels: 256 threads per row; 17..32 levels: 128 threads (same 32 KB of slots)

struct Best {
  float v;
  int32_t j;
};
__device__ __forceinline__ bool better(float v, int32_t j, float bv, int32_t bj) {
  return v > bv || (v == bv && j < bj);
}

template <int kLevels, int kHierThreads>
__global__ void __launch_bounds__(kHierThreads)
hier_metrics_kernel(const float* __restrict__ logits, int64_t ldl, const int32_t* __restrict__ cols, int64_t M,
                    const int8_t* __restrict__ level, int n_levels, const int32_t* __restrict__ first_out,
                    const int32_t* __restrict__ chain, const int32_t* __restrict__ chain_level, int L,
                    int32_t* __restrict__ lvl_idx, int32_t* __restrict__ top1, unsigned long long* counts) {
  __shared__ float s_v[kLevels][kHierThreads];
  __shared__ int32_t s_j[kLevels][kHierThreads];
  __shared__ float s_wv[kLevels][kHierThreads / 32];
  __shared__ int32_t s_wj[kLevels][kHierThreads / 32];
  __shared__ int32_t s_node[kLevels];      // winner node per level (after the -1 rule)
  __shared__ float s_tv[kLevels];          // in-level maxima (before the -1 rule): their best is the TOR top-1
  __shared__ int32_t s_tj[kLevels];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t row = blockIdx.x;
  const float* lr = logits + row * ldl;
  for (int l = 0; l < n_levels; ++l) {
    s_v[l][tid] = -INFINITY;
    s_j[l][tid] = 0x7FFFFFFF;
  }
  // one pass: positions tid, tid + 128, ... (ascending per thread, so `>` keeps the first of equal values)
  for (int64_t j0 = tid; j0 < M; j0 += kHierUnroll * kHierThreads) {
    int32_t c[kHierUnroll];
    int l[kHierUnroll];
    float v[kHierUnroll];
#pragma unroll
    for (int u = 0; u < kHierUnroll; ++u) {
      const int64_t j = j0 + u * kHierThreads;
      c[u] = j < M ? (cols ? __ldg(cols + j) : static_cast<int32_t>(j)) : -1;
    }
#pragma unroll
    for (int u = 0; u < kHierUnroll; ++u) {
      l[u] = c[u] >= 0 ? __ldg(level + c[u]) : -1;
      v[u] = c[u] >= 0 ? __ldg(lr + c[u]) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < kHierUnroll; ++u) {
      if (l[u] >= 0 && l[u] < n_levels && v[u] > s_v[l[u]][tid]) {
        s_v[l[u]][tid] = v[u];
        s_j[l[u]][tid] = static_cast<int32_t>(j0 + u * kHierThreads);
      }
    }
  }
  // per level: block arg-max by (value desc, position asc)
  for (int l = 0; l < n_levels; ++l) {
    float v = s_v[l][tid];
    int32_t j = s_j[l][tid];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, v, o);
      const int32_t oj = __shfl_xor_sync(0xffffffffu, j, o);
      if (better(ov, oj, v, j)) {
        v = ov;
        j = oj;
      }
    }
    if (lane == 0) {
      s_wv[l][warp] = v;
      s_wj[l][warp] = j;
    }
  }
  __syncthreads();
  if (tid <= n_levels) {
    if (tid < n_levels) {
      float v = s_wv[tid][0];
      int32_t j = s_wj[tid][0];
      for (int w = 1; w < kHierThreads / 32; ++w)
        if (better(s_wv[tid][w], s_wj[tid][w], v, j)) {
          v = s_wv[tid][w];
          j = s_wj[tid][w];
        }
      s_tv[tid] = v;
      s_tj[tid] = j;
      // the masked array holds -1 at every out-of-level position (main.py:171)
      const int32_t fo = first_out[tid];
      if (fo >= 0 && fo < M && better(-1.0f, fo, v, j)) {
        v = -1.0f;
        j = fo;
      }
      if (j == 0x7FFFFFFF) j = 0;  // M == 0 cannot happen (checked on the host); defensive
      const int32_t node = cols ? cols[j] : j;
      s_node[tid] = node;
      if (lvl_idx) lvl_idx[row * n_levels + tid] = node;
    }
  }
  __syncthreads();
  if (tid == 0) {
    // overall top-1 over the train columns (main.py:155-156) = best of the per-level in-level maxima
    float bv = -INFINITY;
    int32_t bj = 0x7FFFFFFF;
    for (int l = 0; l < n_levels; ++l)
      if (s_tj[l] != 0x7FFFFFFF && better(s_tv[l], s_tj[l], bv, bj)) {
        bv = s_tv[l];
        bj = s_tj[l];
      }
    const int32_t t1 = bj == 0x7FFFFFFF ? -1 : (cols ? cols[bj] : bj);
    if (top1) top1[row] = t1;
    unsigned long long tor = 0, point = 0, edge = 0;
    bool prev = false;
    for (int k = 0; k < L; ++k) {
      const int32_t p = chain[k];
      tor += (t1 == p);
      const int cl = chain_level[k];
      const bool m = cl >= 0 && cl < n_levels && s_node[cl] == p;
      point += m;
      if (k > 0) edge += (prev && m);
      prev = m;
    }
    if (L == 1) edge = prev ? 1 : 0;
    if (tor) atomicAdd(counts + 0, tor);
    if (point) atomicAdd(counts + 1, point);
    if (edge) atomicAdd(counts + 2, edge);
  }
}

// Second half of the FUSED path (hgr_hier_metrics_fused): the per-level in-level maxima come from the GEMM epilogue
// (score_pair.cu, kEpiLevel) as (order key << 32 | ~sorted bank row) words; one thread per image row decodes them, maps
// the sorted bank row back to its position in `train_index`, and applies the same rules as the kernel above: the -1
// fill (first_out), first position among equal values, chain matching and counting.
__global__ void __launch_bounds__(128)
hier_finish_kernel(unsigned long long* __restrict__ lvl_best, int64_t B, int64_t M, int n_levels,
                   const int32_t* __restrict__ sorted_to_pos, const int32_t* __restrict__ first_out,
                   const int32_t* __restrict__ chain, const int32_t* __restrict__ chain_level, int L,
                   int32_t* __restrict__ lvl_idx, int32_t* __restrict__ top1, unsigned long long* counts) {
  // one WARP per image row, lane = level (n_levels <= 32): every load of a row is issued at once
  const int lane = threadIdx.x & 31;
  const int64_t row = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= B) return;
  float v = -INFINITY;          // in-level maximum of my level (before the -1 rule)
  int32_t j = 0x7FFFFFFF;       // its position in train_index
  int32_t node = 0;             // winner of my level after the -1 rule
  if (lane < n_levels) {
    const unsigned long long w = lvl_best[row * n_levels + lane];
    lvl_best[row * n_levels + lane] = 0ull;     // the workspace is handed back zeroed: the next call needs no memset
    const int32_t fo = first_out[lane];
    if (w != 0ull) {
      const uint32_t k = static_cast<uint32_t>(w >> 32);
      v = __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
      j = sorted_to_pos[~static_cast<uint32_t>(w)];
    }
    node = j;
    if (fo >= 0 && fo < M && better(-1.0f, fo, v, j)) node = fo;   // the masked array holds -1 at every out-of-level position
    if (node == 0x7FFFFFFF) node = 0;
    if (lvl_idx) lvl_idx[row * n_levels + lane] = node;
  }
  // overall top-1 over the train columns = best of the in-level maxima (value desc, position asc)
  float bv = v;
  int32_t bj = j;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int32_t oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (better(ov, oj, bv, bj)) {
      bv = ov;
      bj = oj;
    }
  }
  const int32_t t1 = bj == 0x7FFFFFFF ? -1 : bj;
  if (top1 && lane == 0) top1[row] = t1;
  // chain matching: lane k takes chain entries k, k + 32 (L <= 64); match[k] needs the winner of level chain_level[k]
  unsigned long long tor = 0, point = 0, edge = 0;
  bool carry = false;           // match[k0 - 1] of the previous group of 32
  for (int k0 = 0; k0 < L; k0 += 32) {
    const int k = k0 + lane;
    const int32_t p = k < L ? chain[k] : -2;
    const int cl = k < L ? chain_level[k] : -1;
    const int32_t win = __shfl_sync(0xffffffffu, node, cl >= 0 && cl < 32 ? cl : 0);
    const bool m = k < L && cl >= 0 && cl < n_levels && win == p;
    const unsigned mm = __ballot_sync(0xffffffffu, m);
    tor += __popc(__ballot_sync(0xffffffffu, k < L && t1 == p));
    point += __popc(mm);
    edge += __popc(mm & ((mm << 1) | (carry ? 1u : 0u)));   // match[k] and match[k - 1]
    carry = (mm >> 31) & 1u;
  }
  if (L == 1) edge = point;
  if (lane == 0) {
    if (tor) atomicAdd(counts + 0, tor);
    if (point) atomicAdd(counts + 1, point);
    if (edge) atomicAdd(counts + 2, edge);
  }
}

}  // namespace

int launch_hier_finish(unsigned long long* lvl_best, int64_t B, int64_t M, int n_levels, const int32_t* sorted_to_pos,
                       const int32_t* first_out, const int32_t* chain, const int32_t* chain_level, int L, int32_t* lvl_idx,
                       int32_t* top1, int64_t* counts, cudaStream_t stream) {
  if (n_levels < 1 || n_levels > kMaxLevels)
    return set_error(HGR_ERR_UNSUPPORTED, "hgr_hier_metrics_fused: %d levels outside [1, %d]", n_levels, kMaxLevels);
  const int blocks = static_cast<int>((B + 3) / 4);     // one warp per image row
  hier_finish_kernel<<<blocks, 128, 0, stream>>>(lvl_best, B, M, n_levels, sorted_to_pos, first_out, chain, chain_level, L,
                                                 lvl_idx, top1, reinterpret_cast<unsigned long long*>(counts));
  HGR_CHECK_LAUNCH();
  return HGR_OK;
}

int launch_hier_metrics(const float* logits, int64_t ldl, int64_t B, const int32_t* cols, int64_t M,
                        const int8_t* level, int n_levels, const int32_t* first_out, const int32_t* chain,
                        const int32_t* chain_level, int L, int32_t* lvl_idx, int32_t* top1, int64_t* counts,
                        cudaStream_t stream) {
  if (n_levels < 1 || n_levels > kMaxLevels)
    return set_error(HGR_ERR_UNSUPPORTED, "hgr_hier_metrics: %d levels outside [1, %d]", n_levels, kMaxLevels);
  unsigned long long* cnt = reinterpret_cast<unsigned long long*>(counts);
  if (n_levels <= 16)
    hier_metrics_kernel<16, 256><<<static_cast<unsigned>(B), 256, 0, stream>>>(
        logits, ldl, cols, M, level, n_levels, first_out, chain, chain_level, L, lvl_idx, top1, cnt);
  else
    hier_metrics_kernel<32, 128><<<static_cast<unsigned>(B), 128, 0, stream>>>(
        logits, ldl, cols, M, level, n_levels, first_out, chain, chain_level, L, lvl_idx, top1, cnt);
  HGR_CHECK_LAUNCH();
  return HGR_OK;
}

}  // namespace hgr
