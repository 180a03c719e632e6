// Micro-benchmark of the tcgen05.mma issue / execution rate on sm_100a (diagnostic, not product code):
// one CTA pair per cluster issues long runs of cta_group::2 MMAs (M = 256, K = 16) of a given N with the A operand
// from shared memory (SS) or tensor memory (TS) and reports SM cycles per MMA.  Operands are whatever bits happen to
// be in memory -- only the timing matters.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/mma_probe tools/mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "../hgrnet_b200/csrc/ptx.cuh"

using namespace hgr;

struct Ctl {
  uint64_t done;
  uint64_t sink;      // commit target of the periodic commits (nobody waits on it)
  uint64_t tma[4];    // free-running TMA traffic ring
  uint32_t tmem_base;
  volatile int stop;
};

// mode: 0 = SS, 1 = TS, 2 = alternate SS / TS per group of 4
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(96, 1)
probe(int n, int mode, int groups, int a_stride, long long* out, int commit_every, int tma_kb, int extras,
      const __grid_constant__ CUtensorMap map) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  Ctl* ctl = reinterpret_cast<Ctl*>(smem + 192 * 1024);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  if (warp == 0 && lane == 0) {
    ptx::mbar_init(&ctl->done, 1);
    ptx::mbar_init(&ctl->sink, 1 << 20);
    for (int i = 0; i < 4; ++i) ptx::mbar_init(&ctl->tma[i], 1);
    ctl->stop = 0;
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_cg2(&ctl->tmem_base, 512);
    ptx::tmem_relinquish_cg2();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  if (warp == 1 && rank == 0) {
    const uint32_t a_smem = ptx::smem_u32(smem);                 // 8 x 16 KB A blocks
    const uint32_t b_smem = ptx::smem_u32(smem + 128 * 1024);    // 4 x 16 KB B stages
    const uint32_t idesc = ptx::umma_idesc_bf16(256, n);
    const uint32_t done = ptx::smem_u32(&ctl->done);
    unsigned long long g0, g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    const long long t0 = clock64();
    int since = 0;
    // extras bit 0: tcgen05.fence::after_thread_sync per group; bit 1: try_wait on a completed barrier per group;
    // bit 2: test_wait on a completed barrier per group; bit 3: fence::before_thread_sync per group
    if (lane == 0) ptx::mbar_arrive(&ctl->tma[3]);   // completes phase 0 of a barrier nobody else uses when tma_kb == 0
    __syncwarp();
    for (int g = 0; g < groups; ++g) {
      if (extras & 2) ptx::mbar_wait(&ctl->tma[3], 0);
      if (extras & 4) (void)ptx::mbar_test(&ctl->tma[3], 0);
      if (extras & 1) ptx::tc_fence_after();
      if (extras & 8) ptx::tc_fence_before();
      const uint32_t a_lo = ptx::desc_lo_sw128(a_smem + ((g * a_stride) & 7) * 16384);
      const uint32_t b_lo = ptx::desc_lo_sw128(b_smem + (g & 3) * 16384);
      const uint32_t a_tm = tmem_base + 256 + (g & 7) * 32;
      const bool ts = mode == 1 || (mode == 2 && (g & 1));
      if (ptx::elect_one()) {
        if (ts) {
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_bf16_cg2_ts_lo(tmem_base, a_tm + 8 * k, b_lo + 2 * k, idesc, 1u);
        } else {
#pragma unroll
          for (int k = 0; k < 4; ++k) ptx::umma_bf16_cg2_lo(tmem_base, a_lo + 2 * k, b_lo + 2 * k, idesc, 1u);
        }
      }
      __syncwarp();
      since += 4;
      if (since == commit_every) {
        since = 0;
        if (ptx::elect_one()) ptx::umma_commit_cg2_mc_addr(ptx::smem_u32(&ctl->sink), 0x1);
        __syncwarp();
      }
    }
    const long long t_issue = clock64();
    if (ptx::elect_one()) ptx::umma_commit_cg2_mc_addr(done, 0x1);
    __syncwarp();
    ptx::mbar_wait(&ctl->done, 0);
    const long long t1 = clock64();
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    ctl->stop = 1;
    if (lane == 0) {
      out[(blockIdx.x >> 1) * 4 + 0] = t1 - t0;
      out[(blockIdx.x >> 1) * 4 + 1] = t_issue - t0;
      out[(blockIdx.x >> 1) * 4 + 2] = static_cast<long long>(g1 - g0);
    }
  }
  if (warp == 2 && tma_kb > 0 && lane == 0) {
    // free-running TMA traffic into the upper half of the B region: `tma_kb` KB per request, 4 in flight
    const uint64_t pol = ptx::policy_evict_last();
    uint32_t ph = 0;
    int i = 0;
    long long bytes = 0;
    // (the leader's flag is only visible in the leader CTA; the peer streams a fixed amount)
    const long long limit = static_cast<long long>(groups) * 4 * n * 16;   // ~ the bytes a real B stream would carry
    while (!ctl->stop && bytes < limit) {
      if (i >= 4) ptx::mbar_wait(&ctl->tma[i & 3], ph);
      ptx::mbar_arrive_expect_tx(&ctl->tma[i & 3], tma_kb * 1024);
      for (int q = 0; q < tma_kb / 8; ++q)
        ptx::tma_load_2d(smem + 128 * 1024 + (i & 3) * 16384 + q * 8192, &map, &ctl->tma[i & 3], (q & 15) * 64,
                         ((blockIdx.x * 977 + i * 64) % 20000), pol);
      bytes += tma_kb * 1024;
      ++i;
      if ((i & 3) == 0 && i > 4) ph ^= 1u;
    }
    // drain what is in flight before the CTA may exit
    for (int k = (i > 4 ? i - 4 : 0); k < i; ++k) ptx::mbar_wait(&ctl->tma[k & 3], ((k >> 2) & 1));
    out[(blockIdx.x >> 1) * 4 + 3] = bytes;
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_cg2(tmem_base, 512);
  }
}

#include <cuda.h>
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
int main() {
  const int smem = 1024 + 192 * 1024 + 128;
  void* bank;
  cudaMalloc(&bank, 21841ull * 1024 * 2);
  cudaMemset(bank, 0, 21841ull * 1024 * 2);
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q);
  CUtensorMap map;
  {
    const cuuint64_t dims[2] = {1024, 21841};
    const cuuint64_t strides[1] = {2048};
    const cuuint32_t box[2] = {64, 64};
    const cuuint32_t estr[2] = {1, 1};
    reinterpret_cast<EncodeTiledFn>(f)(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, bank, dims, strides, box, estr,
                                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                                       CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  }
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  long long* out;
  cudaMalloc(&out, 74 * 4 * sizeof(long long));
  long long h[74 * 4];
  const int groups = 2000;
  const char* names[3] = {"SS", "TS", "SS/TS alternating"};
  for (int pairs : {74}) {
    for (int mode = 2; mode < 3; ++mode) {
      for (int n : {64, 128, 256}) {
        for (int rep = 0; rep < 2; ++rep) {
          probe<<<2 * pairs, 96, smem>>>(n, mode, groups, 1, out, 0, 0, 0, map);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("launch failed: %s\n", cudaGetErrorString(e));
            return 1;
          }
        }
        cudaMemcpy(h, out, pairs * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
        double cyc = 0, iss = 0, ns = 0;
        for (int p = 0; p < pairs; ++p) cyc += h[p * 4], iss += h[p * 4 + 1], ns += h[p * 4 + 2];
        cyc /= pairs, iss /= pairs, ns /= pairs;
        printf("pairs %2d  %-18s N=%3d : %.1f cycles/MMA (issue %.1f), %.1f ns/MMA, SM clock %.2f GHz, %.0f%% of 4096 MAC/clk/SM\n",
               pairs, names[mode], n, cyc / (groups * 4), iss / (groups * 4), ns / (groups * 4), cyc / ns,
               100.0 * (128.0 * n * 16) / (cyc / (groups * 4)) / 4096.0);
      }
    }
  }
  // same A block every time (a_stride 0) vs walking 8 blocks: does the A read matter?
  for (int n : {64, 128, 256}) {
    probe<<<2, 96, smem>>>(n, 0, groups, 0, out, 0, 0, 0, map);
    cudaDeviceSynchronize();
    cudaMemcpy(h, out, 4 * sizeof(long long), cudaMemcpyDeviceToHost);
    printf("same-A SS N=%3d : %.1f cycles/MMA\n", n, double(h[0]) / (groups * 4));
  }
  // issue-loop constructs of the real kernel, one at a time (all 74 pairs, N = 128, SS/TS alternating, commit every 8)
  const char* ex_names[] = {"none", "fence::after", "try_wait(done barrier)", "fence::after + try_wait", "test_wait",
                            "fence + test_wait", "try+test", "all three", "fence::before"};
  const int ex_codes[] = {0, 1, 2, 3, 4, 5, 6, 7, 8};
  for (int n : {128, 256}) {
    for (int e = 0; e < 9; ++e) {
      probe<<<148, 96, smem>>>(n, n == 128 ? 2 : 0, groups, 1, out, 8, 0, ex_codes[e], map);
      cudaError_t err = cudaDeviceSynchronize();
      if (err != cudaSuccess) {
        printf("launch failed: %s\n", cudaGetErrorString(err));
        return 1;
      }
      cudaMemcpy(h, out, 74 * 4 * sizeof(long long), cudaMemcpyDeviceToHost);
      double cyc = 0;
      for (int p = 0; p < 74; ++p) cyc += h[p * 4];
      printf("N=%3d per group of 4 MMAs + %-26s: %6.1f cycles/MMA\n", n, ex_names[e], cyc / 74 / (groups * 4));
    }
  }
  return 0;
}
