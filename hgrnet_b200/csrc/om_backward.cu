// Backward of the OM step's logits (kernel 3's other half): the two gradient GEMMs on the tcgen05 loop and the backward
// of both row normalisations, without leaving the device or the library.
//
// Reference: `loss_j.backward()` of every (k, m) iteration (model/clip_tree.py:276) through
// `logits = (img_feats_ @ text_feats.t()) * logit_scale.exp()` (:263) and the two `x / x.norm(dim=-1, keepdim=True)`
// (:225, :262), then `img_feats.backward(img_feats_.grad)` (:280).  The masked CE kernel (masked_ce.cu) has already
// summed the T iterations into ONE dlogits [B, U] over the union of the sampled classes, so the T autograd passes of
// the reference collapse into
//     d_x  = scale * dlogits   . tn       [B, D]      d_tn = scale * dlogits^T . x        [U, D]
//     d_log_scale = sum(dlogits * logits)             d_raw = (d - y (y . d)) / |raw|     (rows of x and of tn)
// Both GEMMs run on the dense tcgen05 kernel of the scoring head (score_pair.cu, store epilogue), which wants both
// operands K-major in bf16: a pack kernel writes dlogits as bf16 in both layouts ([B, U] and [U, B], zero padded to a
// multiple of 8 columns) and accumulates d_log_scale on the way; the unit vectors are transposed once.
#include "common.cuh"

namespace hgr {
namespace {

constexpr int kTile = 32;

// dlogits fp32 [R, C] (ld) -> a16 [R, Cp] bf16 and aT16 [C, Rp] bf16 (padding columns zero); optional sum(d * l).
__global__ void __launch_bounds__(kTile * 8)
om_pack_kernel(const float* __restrict__ d, const float* __restrict__ l, int64_t ld, int R, int C, int Rp, int Cp,
               __nv_bfloat16* __restrict__ a16, __nv_bfloat16* __restrict__ aT16, float* __restrict__ dscale) {
  __shared__ float tile[kTile][kTile + 1];
  __shared__ float s_part[8];
  const int c0 = blockIdx.x * kTile, r0 = blockIdx.y * kTile;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < kTile; i += 8) {
    const int r = r0 + ty + i, c = c0 + tx;
    float v = 0.f;
    if (r < R && c < C) {
      v = d[static_cast<int64_t>(r) * ld + c];
      if (dscale) acc = fmaf(v, l[static_cast<int64_t>(r) * ld + c], acc);
    }
    tile[ty + i][tx] = v;
    if (r < R && c < Cp) a16[static_cast<int64_t>(r) * Cp + c] = __float2bfloat16(v);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kTile; i += 8) {
    const int c = c0 + ty + i, r = r0 + tx;   // transposed: row c of aT16, column r
    if (c < C && r < Rp) aT16[static_cast<int64_t>(c) * Rp + r] = __float2bfloat16(tile[tx][ty + i]);
  }
  if (dscale) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (tx == 0) s_part[ty] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += s_part[i];
      atomicAdd(dscale, s);
    }
  }
}

// y bf16 [R, C] -> yT bf16 [C, Rp] (padding columns zero)
__global__ void __launch_bounds__(kTile * 8)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ y, int R, int C, int Rp, __nv_bfloat16* __restrict__ yT) {
  __shared__ __nv_bfloat16 tile[kTile][kTile + 2];
  const int c0 = blockIdx.x * kTile, r0 = blockIdx.y * kTile;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kTile; i += 8) {
    const int r = r0 + ty + i, c = c0 + tx;
    tile[ty + i][tx] = (r < R && c < C) ? y[static_cast<int64_t>(r) * C + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < kTile; i += 8) {
    const int c = c0 + ty + i, r = r0 + tx;
    if (c < C && r < Rp) yT[static_cast<int64_t>(c) * Rp + r] = tile[tx][ty + i];
  }
}

// Backward of y = raw / |raw| for every row, in place on the fp32 gradient: g <- (g - y (y . g)) / |raw|.  One warp per
// row, 8 elements per lane and step.
__global__ void __launch_bounds__(256)
normalize_backward_kernel(float* __restrict__ g, const __nv_bfloat16* __restrict__ y, const float* __restrict__ norm,
                          int64_t R, int D) {
  const int lane = threadIdx.x & 31;
  const int64_t warp0 = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t r = warp0; r < R; r += nwarps) {
    float* gr = g + r * D;
    const __nv_bfloat16* yr = y + r * D;
    float dot = 0.f;
    for (int c = lane * 4; c < D; c += 128) {
      const float4 gv = *reinterpret_cast<const float4*>(gr + c);
      const uint2 yb = *reinterpret_cast<const uint2*>(yr + c);
      const float2 y01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yb.x));
      const float2 y23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yb.y));
      dot = fmaf(gv.x, y01.x, fmaf(gv.y, y01.y, fmaf(gv.z, y23.x, fmaf(gv.w, y23.y, dot))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
    const float inv = 1.f / norm[r];
    for (int c = lane * 4; c < D; c += 128) {
      float4 gv = *reinterpret_cast<const float4*>(gr + c);
      const uint2 yb = *reinterpret_cast<const uint2*>(yr + c);
      const float2 y01 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yb.x));
      const float2 y23 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yb.y));
      gv.x = (gv.x - y01.x * dot) * inv;
      gv.y = (gv.y - y01.y * dot) * inv;
      gv.z = (gv.z - y23.x * dot) * inv;
      gv.w = (gv.w - y23.y * dot) * inv;
      *reinterpret_cast<float4*>(gr + c) = gv;
    }
  }
}

inline int64_t pad8(int64_t v) { return (v + 7) / 8 * 8; }
inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

}  // namespace

size_t om_backward_workspace_bytes(int64_t B, int64_t U, int64_t D) {
  const int64_t Bp = pad8(B), Up = pad8(U);
  return align256(static_cast<size_t>(B) * Up * 2) + align256(static_cast<size_t>(U) * Bp * 2) +
         align256(static_cast<size_t>(D) * Up * 2) + align256(static_cast<size_t>(D) * Bp * 2) + 256;
}

int launch_om_backward(const float* dlogits, const float* logits, int64_t ldl, int64_t B, int64_t U, int64_t D,
                       const __nv_bfloat16* x, const float* x_norm, const __nv_bfloat16* tn, const float* t_norm,
                       float scale, float* d_img, float* d_text, float* d_log_scale, void* ws, size_t ws_bytes,
                       cudaStream_t stream) {
  if (ws == nullptr || ws_bytes < om_backward_workspace_bytes(B, U, D))
    return set_error(HGR_ERR_WORKSPACE, "hgr_om_backward: workspace %zu < %zu bytes", ws_bytes,
                     om_backward_workspace_bytes(B, U, D));
  if (D % 8 != 0 || D < 8) return set_error(HGR_ERR_BAD_ARG, "hgr_om_backward: D = %lld must be a multiple of 8", (long long)D);
  const int64_t Bp = pad8(B), Up = pad8(U);
  uint8_t* w = static_cast<uint8_t*>(ws);
  __nv_bfloat16* dl16 = reinterpret_cast<__nv_bfloat16*>(w);
  w += align256(static_cast<size_t>(B) * Up * 2);
  __nv_bfloat16* dlT16 = reinterpret_cast<__nv_bfloat16*>(w);
  w += align256(static_cast<size_t>(U) * Bp * 2);
  __nv_bfloat16* tnT = reinterpret_cast<__nv_bfloat16*>(w);
  w += align256(static_cast<size_t>(D) * Up * 2);
  __nv_bfloat16* xT = reinterpret_cast<__nv_bfloat16*>(w);
  const dim3 thr(kTile * 8);
  // the pack grid covers the PADDED extents so that the padding columns are written (zero)
  om_pack_kernel<<<dim3(static_cast<unsigned>((Up + kTile - 1) / kTile), static_cast<unsigned>((Bp + kTile - 1) / kTile)),
                   thr, 0, stream>>>(dlogits, logits, ldl, static_cast<int>(B), static_cast<int>(U), static_cast<int>(Bp),
                                     static_cast<int>(Up), dl16, dlT16, d_log_scale);
  HGR_CHECK_LAUNCH();
  transpose_bf16_kernel<<<dim3(static_cast<unsigned>((D + kTile - 1) / kTile), static_cast<unsigned>((Up + kTile - 1) / kTile)),
                          thr, 0, stream>>>(tn, static_cast<int>(U), static_cast<int>(D), static_cast<int>(Up), tnT);
  HGR_CHECK_LAUNCH();
  transpose_bf16_kernel<<<dim3(static_cast<unsigned>((D + kTile - 1) / kTile), static_cast<unsigned>((Bp + kTile - 1) / kTile)),
                          thr, 0, stream>>>(x, static_cast<int>(B), static_cast<int>(D), static_cast<int>(Bp), xT);
  HGR_CHECK_LAUNCH();
  // d_x = scale * dlogits . tn: rows B, K = Up, output columns = the D rows of tn^T
  int rc = launch_logits_umma(dl16, tnT, B, D, Up, scale, d_img, D, stream);
  if (rc != HGR_OK) return rc;
  // d_tn = scale * dlogits^T . x: rows U, K = Bp, output columns = the D rows of x^T
  rc = launch_logits_umma(dlT16, xT, U, D, Bp, scale, d_text, D, stream);
  if (rc != HGR_OK) return rc;
  const int blocks_i = static_cast<int>((B + 7) / 8 < 4 * num_sms() ? (B + 7) / 8 : 4 * num_sms());
  normalize_backward_kernel<<<blocks_i < 1 ? 1 : blocks_i, 256, 0, stream>>>(d_img, x, x_norm, B, static_cast<int>(D));
  HGR_CHECK_LAUNCH();
  const int blocks_t = static_cast<int>((U + 7) / 8 < 4 * num_sms() ? (U + 7) / 8 : 4 * num_sms());
  normalize_backward_kernel<<<blocks_t < 1 ? 1 : blocks_t, 256, 0, stream>>>(d_text, tn, t_norm, U, static_cast<int>(D));
  HGR_CHECK_LAUNCH();
  return HGR_OK;
}

}  // namespace hgr
