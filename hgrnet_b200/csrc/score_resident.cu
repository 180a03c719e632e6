// Kernel (2): image x class cosine logits on tcgen05 with the per-row running top-K fused into the epilogue --
// the B x C logit matrix never reaches HBM.  The same main loop with a store epilogue backs hgr_logits_dense.
//
// Reference: `feats @ self.zsl_weights.T` (model/clip_tree.py:331), `logits[:, test_index]` +
// `.topk(20, 1, True, True)` (main.py:136-138); the id mapping / hit test (main.py:139-147) runs in the merge
// kernel (topk_merge.cu).
//
// RESIDENT-A main loop.  A worker (a CTA pair sharing tcgen05.mma.cta_group::2, M = 256) walks the bank columns
// of ONE row tile at a time (sched.cuh), so the image operand A -- 128 rows x D per CTA -- is the same for every
// sub-tile of a segment.  Streaming it again per sub-tile is what held the previous kernel at the L2 -> SM limit
// (64 B/clk/SM asked of ~43; profiles/r01c_ncu_summary.md: 213 MB of TMA reads for 45.8 MB of unique data).  Here
// A is loaded ONCE per segment and kept on chip:
//   * up to 8 K blocks (128 rows x 64 bf16 = 16 KB each, 128B-swizzled) stay in shared memory -> SS MMAs;
//   * the remaining K blocks (D = 1024: the first 8) are staged in those slots by TMA, copied into TENSOR MEMORY
//     by the MMA thread (tcgen05.cp 128x256b: row = TMEM lane, two bf16 per 32-bit column; the copies run in the
//     tensor pipe, in order, ahead of the MMAs that read them) and feed the MMAs from there
//     (tcgen05.mma ... [d], [a_tmem], b_desc); the slot is then refilled with its resident block.  TMEM: 32 columns per K block for A, the rest split into two accumulator buffers
//     (D = 1024: 256 + 2 x 128; D <= 512: 0 + 2 x 256).
// Only the bank streams: each CTA feeds HALF of a sub-tile's bank rows per K block (8 KB at D = 1024) through an
// mbarrier ring, with an L2 prefetch of the next sub-tile's boxes one sub-tile ahead (the ring only has to cover
// the L2 latency, HBM latency is taken by the prefetch).  Operand traffic per CTA and 128 x 128 x 64 MACs: 8 KB
// per 256 tensor cycles = 32 B/clk/SM.
//
// Roles per CTA: warp 0 TMA producer, warp 1 TMEM allocation + (leader CTA) MMA issue, warps 2-5 epilogue (one
// warp per TMEM lane quarter; thread = image row).  Epilogue: umma_common.cuh (deferred inserts).
#include <cuda.h>

#include <cmath>
#include <cstdlib>

#include "umma_common.cuh"

namespace hgr {
namespace umma {
namespace {

constexpr int kASlotBytes = kABytes;   // 16 KB
constexpr int kMaxASlots = 8;          // K blocks of A resident in shared memory
constexpr int kMaxATmem = 8;           // K blocks of A resident in tensor memory
constexpr int kMaxBStages = 12;
constexpr int kEpiThreads = 128;
constexpr int kQDepth = 48;            // deferred-insert queue entries per row (+ 1 dud slot)
constexpr int kQueueBytesTopk = (kQDepth + 1) * kEpiThreads * 8;
constexpr int kDenseTile = 32 * 33;    // dense epilogue: one padded 32 x 32 transpose tile per warp
constexpr int kQueueBytesDense = 4 * kDenseTile * 4;
constexpr int kSmemLimit = 232448;     // 227 KB opt-in limit per CTA

struct ResCtl {
  uint64_t full[kMaxBStages];          // leader's copy in use: bank stage landed in both CTAs
  uint64_t empty[kMaxBStages];         // per CTA, multicast commit of the leader's MMA thread
  uint64_t tmem_full[2];               // per CTA, multicast commit
  uint64_t tmem_empty[2];              // leader's copy: 4 epilogue warps of each CTA
  uint64_t a_stage_full[kMaxATmem];    // leader's copy: TMEM-bound A block landed in its staging slot in both CTAs
  uint64_t a_stage_free[kMaxATmem];    // per CTA, multicast commit: the block has been copied to tensor memory
  uint64_t a_res_full[kMaxASlots];     // leader's copy: resident A block landed in both CTAs
  uint64_t seg_done;                   // per CTA, multicast commit: every MMA of the segment retired
  uint32_t tmem_base;
};

inline int queue_bytes(int epi) {
  return epi == kEpiTopkDefer ? kQueueBytesTopk : epi == kEpiDense ? kQueueBytesDense : 0;
}

template <int EPI, int KL>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + kEpiThreads, 1)
score_resident_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_bank,
                      const Params p) {
  constexpr int kQueueBytes = EPI == kEpiTopkDefer ? kQueueBytesTopk : EPI == kEpiDense ? kQueueBytesDense : 0;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int kb_t = p.kb_tmem, n_slots = p.a_slots, nkb = p.num_k_blocks, n_stages = p.stages;
  const int skb = p.stage_kb;                      // K blocks per bank stage (one barrier hand-off per stage)
  const int box_bytes = p.b_stage_bytes;           // one K block of this CTA's half sub-tile
  const int stage_bytes = skb * box_bytes;
  uint8_t* a_base = smem;                                        // [n_slots][16 KB]
  uint8_t* b_base = smem + n_slots * kASlotBytes;                // [n_stages][stage_bytes]
  float* queue_base = reinterpret_cast<float*>(b_base + n_stages * stage_bytes);
  ResCtl* ctl = reinterpret_cast<ResCtl*>(reinterpret_cast<uint8_t*>(queue_base) + kQueueBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();  // 0 = leader (issues the MMAs), 1 = peer
  const int pair = blockIdx.x >> 1;
  const uint32_t a_col0 = 2u * static_cast<uint32_t>(p.sub_n);   // TMEM column of A block 0 (behind the accumulators)

  if (blockIdx.x == 0 && threadIdx.x == 0 && p.stats != nullptr) p.stats[0] = 0;
  if (threadIdx.x == 0) stamp(p, 0);
  if (warp == 0 && lane == 0) {
    ptx::prefetch_tensormap(&map_x);
    ptx::prefetch_tensormap(&map_bank);
    for (int s = 0; s < n_stages; ++s) {
      ptx::mbar_init(&ctl->full[s], 1);
      ptx::mbar_init(&ctl->empty[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      ptx::mbar_init(&ctl->tmem_full[b], 1);
      ptx::mbar_init(&ctl->tmem_empty[b], 2 * 4);
    }
    for (int i = 0; i < kMaxATmem; ++i) {
      ptx::mbar_init(&ctl->a_stage_full[i], 1);
      ptx::mbar_init(&ctl->a_stage_free[i], 1);
    }
    for (int i = 0; i < kMaxASlots; ++i) ptx::mbar_init(&ctl->a_res_full[i], 1);
    ptx::mbar_init(&ctl->seg_done, 1);
    ptx::fence_mbar_init();
  }
  if (warp == 1) {
    ptx::tmem_alloc_cg2(&ctl->tmem_base, kTmemCols);
    ptx::tmem_relinquish_cg2();
  }
  ptx::tc_fence_before();
  ptx::cluster_sync_all();  // barriers of both CTAs initialised, TMEM allocated
  ptx::tc_fence_after();
  const uint32_t tmem_base = ctl->tmem_base;
  if (threadIdx.x == 0) stamp(p, 1);

  if (warp == 0) {
    // ===================== TMA producer (both CTAs) =====================
    // The whole warp walks the schedule and waits on the barriers (warp-uniform control flow: TMA operands stay in
    // uniform registers, see ptx.cuh); one elected lane issues.
    const uint64_t pol_a = ptx::policy_evict_last();
    const uint64_t pol_b = ptx::policy_evict_last();
    TileWalker walk(p.sched, pair, p.C, 0, p.sub_n);
    TileWalker ahead(p.sched, pair, p.C, 0, p.sub_n);   // one sub-tile ahead: L2 prefetch of its bank boxes
    SubTile t, nx;
    bool has_nx = ahead.next(nx);
    int stage = 0, seg = -1;
    uint32_t phase = 0;
    while (walk.next(t)) {
      has_nx = ahead.next(nx);
      const int half = t.n >> 1;                                   // bank rows this CTA feeds
      const int row0 = t.mt * (2 * kTileM) + static_cast<int>(rank) * kTileM;
      const int bcol0 = t.col0 + static_cast<int>(rank) * half;
      const int nx_bcol0 = has_nx ? nx.col0 + static_cast<int>(rank) * (nx.n >> 1) : 0;
      int j_next = n_slots;          // next resident A slot to load (first sub-tile of a segment only)
      uint32_t sp = 0;
      if (t.first) {
        ++seg;
        sp = static_cast<uint32_t>(seg) & 1u;
        // the A operand of the previous segment is still being read until its last MMA retires
        if (seg > 0) ptx::mbar_wait(&ctl->seg_done, static_cast<uint32_t>(seg - 1) & 1u);
        // every TMEM-bound A block of the segment at once: K block i is staged in slot i, the leader's MMA thread
        // copies it to tensor memory (tcgen05.cp) and hands the slot back
        if (ptx::elect_one()) {
          for (int i = 0; i < kb_t; ++i) {
            if (rank == 0) ptx::mbar_arrive_expect_tx(&ctl->a_stage_full[i], 2 * kASlotBytes);
            ptx::tma_load_2d_cg2(a_base + i * kASlotBytes, &map_x,
                                 ptx::mapa_shared(ptx::smem_u32(&ctl->a_stage_full[i]), 0), i * kBlockK, row0, pol_a);
          }
        }
        __syncwarp();
        j_next = 0;
      }
      // resident slot j holds K block kb_t + j; slots < kb_t first have to be released by the copy
      auto load_resident = [&](int j) {
        if (ptx::elect_one()) {
          if (rank == 0) ptx::mbar_arrive_expect_tx(&ctl->a_res_full[j], 2 * kASlotBytes);
          ptx::tma_load_2d_cg2(a_base + j * kASlotBytes, &map_x,
                               ptx::mapa_shared(ptx::smem_u32(&ctl->a_res_full[j]), 0), (kb_t + j) * kBlockK, row0, pol_a);
        }
        __syncwarp();
      };
      for (int kb0 = 0; kb0 < nkb; kb0 += skb) {
        const int nk = nkb - kb0 < skb ? nkb - kb0 : skb;   // K blocks of this stage
        if (j_next < n_slots) {
          // whatever has been released goes out at once; what this stage's MMAs need must be out before the ring
          // wait below can depend on them (the wait needs MMA progress, the MMAs need these blocks)
          while (j_next < n_slots && (j_next >= kb_t || ptx::mbar_test(&ctl->a_stage_free[j_next], sp))) load_resident(j_next++);
          const int need = kb0 + nk - kb_t < n_slots ? kb0 + nk - kb_t : n_slots;
          while (j_next < need) {
            if (j_next < kb_t) ptx::mbar_wait(&ctl->a_stage_free[j_next], sp);
            load_resident(j_next++);
          }
        }
        ptx::mbar_wait(&ctl->empty[stage], phase ^ 1u);
        if (ptx::elect_one()) {
          const uint32_t full_leader = ptx::mapa_shared(ptx::smem_u32(&ctl->full[stage]), 0);
          if (rank == 0) ptx::mbar_arrive_expect_tx(&ctl->full[stage], 2 * nk * box_bytes);
          for (int q = 0; q < nk; ++q)
            ptx::tma_load_2d_cg2(b_base + stage * stage_bytes + q * box_bytes, &map_bank, full_leader,
                                 (kb0 + q) * kBlockK, bcol0, pol_b);
          if (has_nx && p.prefetch)
            for (int q = 0; q < nk; ++q) ptx::tma_prefetch_2d(&map_bank, (kb0 + q) * kBlockK, nx_bcol0);
        }
        __syncwarp();
        if (++stage == n_stages) {
          stage = 0;
          phase ^= 1u;
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only; the whole warp stays convergent) =====================
    // This loop's INSTRUCTION COUNT is the tensor pipe's feed rate: one warp retires an instruction every ~6-8 cycles
    // here, an N = 128 MMA executes in 64 -- so a ring stage (8 MMAs, 512 tensor cycles) may cost ~70 instructions
    // in all (measured with tools/mma_probe.cu: 64 cycles/MMA with a 35-instruction group of four, 114 with 60).
    // Hence: waits are single asm statements (no visible data-dependent branch: cursors stay in uniform registers),
    // descriptors advance by additions, and nothing else lives in the loop.
    if (rank == 0) {
      TileWalker walk(p.sched, pair, p.C, 0, p.sub_n);
      SubTile t;
      int stage = 0, it = 0, seg = -1;
      uint32_t phase = 0;
      const uint32_t a_lo0 = ptx::desc_lo_sw128(ptx::smem_u32(a_base));   // slots / stages are 1024-byte aligned: the
      const uint32_t b_lo0 = ptx::desc_lo_sw128(ptx::smem_u32(b_base));   // address field never carries
      const uint32_t a_tmem0 = tmem_base + a_col0;
      const uint32_t full0 = ptx::smem_u32(&ctl->full[0]);
      const uint32_t empty0 = ptx::smem_u32(&ctl->empty[0]);
      const uint32_t res_full0 = ptx::smem_u32(&ctl->a_res_full[0]);
      const uint32_t box16 = static_cast<uint32_t>(box_bytes >> 4), stage16 = static_cast<uint32_t>(stage_bytes >> 4);
      const bool prof = p.timeline != nullptr;
      long long c_a = 0;
      const long long c_begin = prof ? clock64() : 0;
      while (walk.next(t)) {
        const int buf = it & 1;
        if (t.first) {
          // this segment's TMEM-bound A blocks: smem slot i -> tensor memory, four 128 x 16 slices per block; the
          // copies run in the tensor pipe ahead of the MMAs that read them, the commit hands slot i to the producers
          ++seg;
          const long long c_t = prof ? clock64() : 0;
          for (int i = 0; i < kb_t; ++i) {
            ptx::mbar_wait_spin(ptx::smem_u32(&ctl->a_stage_full[i]), static_cast<uint32_t>(seg) & 1u);
            ptx::tc_fence_after();
            if (ptx::elect_one()) {
              const uint32_t src = a_lo0 + static_cast<uint32_t>((i * kASlotBytes) >> 4);
#pragma unroll
              for (int k = 0; k < kBlockK / kUmmaK; ++k)
                ptx::tmem_cp_128x256b_cg2_lo(a_tmem0 + static_cast<uint32_t>(i * (kBlockK / 2) + k * (kUmmaK / 2)),
                                             src + k * (kUmmaK * 2 / 16));
              ptx::umma_commit_cg2_mc(&ctl->a_stage_free[i], 0x3);
            }
            __syncwarp();
            if (prof && seg == 0 && lane == 0 && (i == 0 || i == kb_t - 1)) stamp(p, i == 0 ? 26 : 23);
          }
          if (prof) c_a += clock64() - c_t;
        }
        const uint32_t sp = static_cast<uint32_t>(seg) & 1u;
        ptx::mbar_wait_spin(ptx::smem_u32(&ctl->tmem_empty[buf]), ((it >> 1) & 1) ^ 1u);
        ptx::tc_fence_after();
        const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(buf * p.sub_n);
        const uint32_t idesc = ptx::umma_idesc_bf16(2 * kTileM, t.n);
        uint32_t b_lo = b_lo0 + static_cast<uint32_t>(stage) * stage16;
        const bool first = t.first;
        for (int kb0 = 0; kb0 < nkb; kb0 += skb) {
          const int nk = nkb - kb0 < skb ? nkb - kb0 : skb;
          if (first) {   // the resident A blocks arrive during the first sub-tile of a segment
            for (int kb = kb0 > kb_t ? kb0 : kb_t; kb < kb0 + nk; ++kb) ptx::mbar_wait_spin(res_full0 + (kb - kb_t) * 8, sp);
          }
          ptx::mbar_wait_spin(full0 + stage * 8, phase);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            uint32_t b = b_lo;
            for (int kb = kb0; kb < kb0 + nk; ++kb, b += box16) {
              if (kb < kb_t)   // K block kb lives in tensor memory: 32 columns per block
                ptx::umma_kblock_cg2_ts(d_tmem, a_tmem0 + (static_cast<uint32_t>(kb) << 5), b, idesc, kb != 0 ? 1u : 0u);
              else             // ... in shared-memory slot kb - kb_t: 16 KB = 1024 descriptor units per slot
                ptx::umma_kblock_cg2_ss(d_tmem, a_lo0 + (static_cast<uint32_t>(kb - kb_t) << 10), b, idesc,
                                        kb != 0 ? 1u : 0u);
            }
            ptx::umma_commit_cg2_mc_addr(empty0 + stage * 8, 0x3);  // frees this bank stage in both CTAs
          }
          b_lo += stage16;
          if (++stage == n_stages) {
            stage = 0;
            phase ^= 1u;
            b_lo = b_lo0;
          }
        }
        if (ptx::elect_one()) {
          ptx::umma_commit_cg2_mc(&ctl->tmem_full[buf], 0x3);  // accumulators of both CTAs complete
          if (t.last) ptx::umma_commit_cg2_mc(&ctl->seg_done, 0x3);  // A operand of this segment no longer read
        }
        __syncwarp();
        ++it;
      }
      if (lane == 0) stamp(p, 3);  // last MMA issued
      if (prof && lane == 0 && blockIdx.x < 256) {
        unsigned long long* tl = p.timeline + blockIdx.x * kTimelineSlots;
        tl[21] = c_a, tl[22] = clock64() - c_begin;
      }
    }
  } else {
    // ===================== epilogue (both CTAs, own 128 rows each) =====================
    const int quarter = warp & 3;
    const int row_local = quarter * 32 + lane;
    const int row_in_tile = static_cast<int>(rank) * kTileM + row_local;
    const int epi_tid = (warp - kEpiWarp0) * 32 + lane;
    const uint32_t lane_taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16);
    TileWalker walk(p.sched, pair, p.C, 0, p.sub_n);
    SubTile t;
    SortedList<KL> list;
    list.init();
    float null_acc = -INFINITY;
    float floor_thr = -INFINITY;
    CandQueue<kEpiThreads, kQDepth> cq;
    cq.init(ptx::smem_u32(reinterpret_cast<uint2*>(queue_base) + epi_tid));
    int it = 0, seg = -1;
    EpiClock ck(p.timeline != nullptr && epi_tid == 0);
    while (walk.next(t)) {
      const int buf = it & 1;
      const int64_t row = static_cast<int64_t>(t.mt) * (2 * kTileM) + row_in_tile;
      if (t.first) {
        ++seg;
        list.init();
        null_acc = -INFINITY;
        floor_thr = -INFINITY;
      }
      ck.start();
      ptx::mbar_wait(&ctl->tmem_full[buf], (it >> 1) & 1);
      ptx::tc_fence_after();
      ck.lap(ck.wait);
      if (epi_tid == 0 && it < 4) stamp(p, 4 + it);  // accumulator of sub-tile `it` ready
      const uint32_t taddr = lane_taddr + static_cast<uint32_t>(buf * p.sub_n);
      float sub_thr = -INFINITY;
      if (EPI == kEpiTopkDefer) {
        if (t.first) floor_thr = warmup_floor_pairs<KL>(taddr, t.nvalid);
        ck.lap(ck.warm);
        sub_thr = fmaxf(floor_thr, list.thr());   // fixed for the whole sub-tile
      }
      for (int c0 = 0; c0 < t.nvalid; c0 += kChunk) {
        uint32_t r[kChunk];
        ck.start();
        ptx::tmem_ld_x32(taddr + c0, r);
        ptx::tmem_ld_wait();
        ck.lap(ck.ld);
        const int nv = t.nvalid - c0;
        if (EPI == kEpiDense) {
          // thread = row holds 32 consecutive columns; transpose the warp's 32x32 block through a padded smem
          // tile so that every store instruction writes 128 contiguous bytes of ONE row
          float* tile = queue_base + (warp - kEpiWarp0) * kDenseTile;
#pragma unroll
          for (int j = 0; j < kChunk; ++j) tile[lane * 33 + j] = __uint_as_float(r[j]) * p.scale;
          __syncwarp();
          const int64_t row_base = row - lane;
          float* o = p.dense_out + row_base * p.ldo + t.col0 + c0 + lane;
          const int nrows = p.B - row_base < 32 ? static_cast<int>(p.B - row_base) : 32;
          if (lane < nv) {
#pragma unroll 8
            for (int rr = 0; rr < nrows; ++rr) o[static_cast<int64_t>(rr) * p.ldo] = tile[rr * 33 + lane];
          }
          __syncwarp();
        } else if (EPI == kEpiTopkDefer) {
          const bool roomy = __reduce_max_sync(0xffffffffu, cq.count()) <= kQDepth - kChunk;
          // Long lists (many survivors per sub-tile) make room up front; short lists rarely pass the roomy mark, and
          // when a lane does they take the saturating append below, so that only a queue that really fills up makes
          // the warp drain while it still holds the TMEM buffer.
          if (!roomy && KL > 10) {
            ck.lap(ck.scan);
            cand_drain<KL>(list, cq);
            sub_thr = fmaxf(sub_thr, list.thr());
            ck.lap(ck.drain);
          }
          if (roomy || KL > 10) {
            if (nv >= kChunk) cand_append_chunk_roomy<true>(cq, r, nv, t.col0 + c0, sub_thr);
            else cand_append_chunk_roomy<false>(cq, r, nv, t.col0 + c0, sub_thr);
            ck.lap(ck.scan);
            continue;
          }
          const uint32_t wr0 = cq.wr;
          if (!cand_append_chunk(cq, r, nv, t.col0 + c0, sub_thr)) {  // a lane ran out of slots (rare):
            cq.wr = wr0;                                              // rewind, insert what is queued,
            ck.lap(ck.scan);
            cand_drain<KL>(list, cq);                                 // tighten the threshold and redo the chunk
            sub_thr = fmaxf(sub_thr, list.thr());
            ck.lap(ck.drain);
            if (!cand_append_chunk(cq, r, nv, t.col0 + c0, sub_thr)) {
              // more survivors in ONE chunk than the queue holds (list still warming up): insert straight from
              // TMEM, column by column
              cq.wr = cq.base;
              scan_chunk_reload<KL>(list, r, nv, taddr + c0, t.col0 + c0);
              sub_thr = fmaxf(sub_thr, list.thr());
            }
          }
          ck.lap(ck.scan);
        } else {
#pragma unroll
          for (int j = 0; j < kChunk; ++j) null_acc = fmaxf(null_acc, __uint_as_float(r[j]));
        }
      }
      // this CTA's half of the accumulator buffer is drained: tell the leader's MMA thread
      ptx::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) ptx::mbar_arrive(&ctl->tmem_empty[buf]);
        else ptx::mbar_arrive_cluster(ptx::mapa_shared(ptx::smem_u32(&ctl->tmem_empty[buf]), 0));
      }
      if (EPI == kEpiTopkDefer) {  // the buffer is already back with the tensor core: now pay for the inserts
        ck.start();
        cand_drain<KL>(list, cq);
        ck.lap(ck.drain);
      }
      if (epi_tid == 0 && it < 4) stamp(p, 8 + it);  // this warp is done with sub-tile `it`
      if (EPI != kEpiDense && t.last && row < p.B) {
        const int slot = pair - p.sched.first_cta(t.mt);
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
    if (epi_tid == 0) stamp(p, 12);  // epilogue done
    if (ck.on && blockIdx.x < 256) {
      unsigned long long* tl = p.timeline + blockIdx.x * kTimelineSlots;
      tl[16] = ck.wait, tl[17] = ck.warm, tl[18] = ck.ld, tl[19] = ck.scan, tl[20] = ck.drain;
    }
  }

  // no CTA may exit (or free TMEM) while its partner can still signal its barriers or read its memory
  ptx::tc_fence_before();
  ptx::cluster_sync_all();
  if (warp == 1) {
    ptx::tc_fence_after();
    ptx::tmem_dealloc_cg2(tmem_base, kTmemCols);
  }
  if (threadIdx.x == 32) stamp(p, 13);  // exit
}

template <int EPI, int KL>
int launch_one(const CUtensorMap& mx, const CUtensorMap& mb, const Params& p, size_t smem, cudaStream_t stream) {
  auto kern = score_resident_kernel<EPI, KL>;
  // the opt-in limit is a per-function attribute: always ask for the maximum so that concurrent callers with
  // different geometries cannot lower it under each other
  HGR_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemLimit));
  kern<<<2 * p.sched.G, 64 + kEpiThreads, smem, stream>>>(mx, mb, p);
  HGR_CHECK_LAUNCH();
  return HGR_OK;
}

}  // namespace

bool resident_supported(int64_t D) { return (D + kBlockK - 1) / kBlockK <= kMaxASlots + kMaxATmem; }

ResGeom resident_geom(int64_t D, int epi) {
  ResGeom g{};
  g.nkb = static_cast<int>((D + kBlockK - 1) / kBlockK);
  g.a_slots = g.nkb < kMaxASlots ? g.nkb : kMaxASlots;
  g.kb_tmem = g.nkb - g.a_slots;
  static const int force_kbt = getenv("HGR_RES_KBT") ? atoi(getenv("HGR_RES_KBT")) : -1;   // experiments
  static const int force_subn = getenv("HGR_RES_SUBN") ? atoi(getenv("HGR_RES_SUBN")) : 0;
  if (force_kbt >= 0 && force_kbt <= kMaxATmem && g.nkb - force_kbt >= force_kbt && g.nkb - force_kbt <= kMaxASlots &&
      g.nkb - force_kbt >= 1) {
    g.kb_tmem = force_kbt;
    g.a_slots = g.nkb - force_kbt;
  }
  g.sub_n = g.kb_tmem == 0 ? kSubN : ((kTmemCols - (kBlockK / 2) * g.kb_tmem) / 2) / kUnit * kUnit;
  if (force_subn >= 16 && force_subn <= g.sub_n && force_subn % 16 == 0) g.sub_n = force_subn;
  g.b_stage_bytes = (g.sub_n / 2) * kBlockK * 2;
  // K blocks per ring stage: every barrier hand-off costs the issuing threads a fixed ~250 cycles (wait, fence,
  // commit, descriptor set-up), so a stage carries at least 512 tensor cycles of MMAs (N = 128: two K blocks)
  static const int force_skb = getenv("HGR_RES_SKB") ? atoi(getenv("HGR_RES_SKB")) : 0;
  g.stage_kb = g.sub_n <= 128 ? 2 : 1;
  if (force_skb >= 1 && force_skb <= 4) g.stage_kb = force_skb;
  if (g.stage_kb > g.nkb) g.stage_kb = g.nkb;
  const int stage_bytes = g.stage_kb * g.b_stage_bytes;
  const int fixed = 1024 + static_cast<int>(sizeof(ResCtl)) + g.a_slots * kASlotBytes + queue_bytes(epi);
  int stages = (kSmemLimit - fixed) / stage_bytes;
  static const int forced = [] {
    const char* e = getenv("HGR_STAGES");
    return e ? atoi(e) : 0;
  }();
  if (forced >= 2 && forced < stages) stages = forced;
  g.stages = stages > kMaxBStages ? kMaxBStages : stages;
  g.smem = static_cast<size_t>(fixed) + static_cast<size_t>(g.stages) * stage_bytes;
  return g;
}

int launch_resident_kernel(int epi, int KL, const CUtensorMap& mx, const CUtensorMap& mb, const Params& p,
                           const ResGeom& g, cudaStream_t stream) {
  if (g.stages < 2) return set_error(HGR_ERR_UNSUPPORTED, "hgr_score_topk(tcgen05): no room for the bank ring");
  if (epi == kEpiDense) return launch_one<kEpiDense, 8>(mx, mb, p, g.smem, stream);
  if (epi == kEpiNull) return launch_one<kEpiNull, 8>(mx, mb, p, g.smem, stream);
  switch (KL) {
    case 8: return launch_one<kEpiTopkDefer, 8>(mx, mb, p, g.smem, stream);
    case 10: return launch_one<kEpiTopkDefer, 10>(mx, mb, p, g.smem, stream);
    case 12: return launch_one<kEpiTopkDefer, 12>(mx, mb, p, g.smem, stream);
    case 16: return launch_one<kEpiTopkDefer, 16>(mx, mb, p, g.smem, stream);
    case 20: return launch_one<kEpiTopkDefer, 20>(mx, mb, p, g.smem, stream);
    case 32: return launch_one<kEpiTopkDefer, 32>(mx, mb, p, g.smem, stream);
  }
  return set_error(HGR_ERR_UNSUPPORTED, "hgr_score_topk(tcgen05): list length %d", KL);
}

}  // namespace umma
}  // namespace hgr
