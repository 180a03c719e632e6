// Kernel (3): fused masked cross-entropy (forward + backward) of the OM training step.
//
// Reference: the (k_loop, m_loop) double loop of tree_model.train_batch
// (model/clip_tree.py:241-277): per iteration t a class set `compare_idx` is sampled
// (:257), logits = (img @ text[compare_idx]^T) * logit_scale.exp() (:261-263),
// loss_t = CrossEntropyLoss(logits, labels) * w_in[m] * w_out[k] (:275), loss_t.backward()
// (:276), loss_t.item() (:277).  Here all T iterations run in ONE launch over the logits of
// the UNION of the sampled classes: iteration t is a column mask (its set) over that matrix.
// No host synchronisation, no per-iteration autograd graph.
//
// One CTA per image row b:  the row's dlogits accumulate in shared memory across the T sets
// (deterministic: sets are processed in order, columns inside a set are distinct), the
// per-(t, b) losses go to the workspace and a second tiny kernel reduces them over b in a
// fixed order.  HBM traffic: read B*U logits once (set columns re-read from L1/L2), write B*U
// dlogits once -- algorithmic bytes 8*B*U (+ sets).
#include "common.cuh"

namespace hgr {
namespace {

constexpr int kCeThreads = 256;

__device__ __forceinline__ float block_reduce(float v, float* s_red, bool is_max) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float t = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, t) : v + t;
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();  // s_red free to overwrite
  if (lane == 0) s_red[warp] = v;
  __syncthreads();
  float r = s_red[0];
#pragma unroll
  for (int i = 1; i < kCeThreads / 32; ++i) r = is_max ? fmaxf(r, s_red[i]) : r + s_red[i];
  return r;
}

__global__ void __launch_bounds__(kCeThreads)
masked_ce_kernel(const float* __restrict__ logits, int64_t ldl, int B, int U, const int32_t* __restrict__ set_ptr,
                 const int32_t* __restrict__ set_col, const int32_t* __restrict__ label_pos,
                 const float* __restrict__ weight, int T, float* __restrict__ part_loss,
                 float* __restrict__ dlogits) {
  extern __shared__ float s_dl[];  // [U] when dlogits != nullptr
  __shared__ float s_red[kCeThreads / 32];
  const int b = blockIdx.x;
  const float* row = logits + static_cast<int64_t>(b) * ldl;
  if (dlogits)
    for (int u = threadIdx.x; u < U; u += kCeThreads) s_dl[u] = 0.f;
  __syncthreads();

  for (int t = 0; t < T; ++t) {
    const int beg = set_ptr[t], n = set_ptr[t + 1] - beg;
    const int32_t* cols = set_col + beg;
    float m = -INFINITY;
    for (int j = threadIdx.x; j < n; j += kCeThreads) m = fmaxf(m, row[cols[j]]);
    m = block_reduce(m, s_red, true);
    float s = 0.f;
    for (int j = threadIdx.x; j < n; j += kCeThreads) s += expf(row[cols[j]] - m);
    s = block_reduce(s, s_red, false);
    const float lse = m + logf(s);
    const int lp = label_pos[t];
    if (threadIdx.x == 0) part_loss[static_cast<int64_t>(t) * B + b] = lse - row[cols[lp]];
    if (dlogits) {
      const float coef = weight[t] / static_cast<float>(B);  // mean over the batch (:49) times the level weights
      for (int j = threadIdx.x; j < n; j += kCeThreads) {
        const int c = cols[j];
        s_dl[c] += coef * (expf(row[c] - lse) - (j == lp ? 1.f : 0.f));
      }
      __syncthreads();
    }
  }
  if (dlogits) {
    float* drow = dlogits + static_cast<int64_t>(b) * ldl;
    for (int u = threadIdx.x; u < U; u += kCeThreads) drow[u] = s_dl[u];
  }
}

// loss[t] = weight[t] * mean_b part_loss[t][b]; one warp per t, fixed summation order.
__global__ void __launch_bounds__(32)
masked_ce_reduce_kernel(const float* __restrict__ part_loss, const float* __restrict__ weight, int B,
                        float* __restrict__ loss) {
  const int t = blockIdx.x, lane = threadIdx.x;
  float s = 0.f;
  for (int b = lane; b < B; b += 32) s += part_loss[static_cast<int64_t>(t) * B + b];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) loss[t] = weight[t] * (s / static_cast<float>(B));
}

}  // namespace

size_t masked_ce_workspace_bytes(int64_t B, int64_t U, int64_t T) {
  (void)U;
  return static_cast<size_t>(B) * static_cast<size_t>(T > 0 ? T : 1) * sizeof(float);
}

int launch_masked_ce(const float* logits, int64_t ldl, int64_t B, int64_t U, const int32_t* set_ptr,
                     const int32_t* set_col, const int32_t* label_pos, const float* weight, int64_t T, float* loss,
                     float* dlogits, void* ws, size_t ws_bytes, cudaStream_t stream) {
  const size_t need = masked_ce_workspace_bytes(B, U, T);
  if (ws == nullptr || ws_bytes < need)
    return set_error(HGR_ERR_WORKSPACE, "hgr_masked_ce: workspace %zu < %zu bytes", ws_bytes, need);
  const size_t smem = dlogits ? static_cast<size_t>(U) * sizeof(float) : 0;
  if (smem > 200 * 1024)
    return set_error(HGR_ERR_UNSUPPORTED, "hgr_masked_ce: union of %lld classes exceeds the shared-memory row buffer",
                     (long long)U);
  if (smem > 48 * 1024)
    HGR_CHECK_CUDA(cudaFuncSetAttribute(masked_ce_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        static_cast<int>(smem)));
  float* part = static_cast<float*>(ws);
  masked_ce_kernel<<<static_cast<unsigned>(B), kCeThreads, smem, stream>>>(
      logits, ldl, static_cast<int>(B), static_cast<int>(U), set_ptr, set_col, label_pos, weight,
      static_cast<int>(T), part, dlogits);
  HGR_CHECK_LAUNCH();
  masked_ce_reduce_kernel<<<static_cast<unsigned>(T), 32, 0, stream>>>(part, weight, static_cast<int>(B), loss);
  HGR_CHECK_LAUNCH();
  return HGR_OK;
}

}  // namespace hgr
