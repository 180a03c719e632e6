#!/usr/bin/env python
"""Benchmark of the HGR-Net scoring head on B200 -- the driver contract.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg4|cfg5]

Metric (BASELINE.json): scored images/s at 21,841 classes.  A STEP is one pass of the hot path
over one batch of synthetic image features: row-normalise the batch (kernel 1), fused
logits + top-20 + Hit@k against the class bank (kernel 2 + merge).  N=1 runs BASELINE cfg 2
(B=512, C=21,841, D=1024); N>1 runs the class-sharded sweep of cfg 5 (B=4096, bank row-sharded
over the ranks) -- ``config.workload`` names which.  Because the two differ, the N=1 line also
carries ``cfg5_single_gpu`` (cfg 5 on this one GPU) and every N>1 line ``single_gpu_same_workload``
(cfg 5 with the whole bank on rank 0's GPU, measured in the same run) and ``speedup_vs_1gpu``.

``value``   device-timed throughput with inputs resident in HBM (CUDA events, max over ranks).
``e2e``     the same metric through the public API (``tree_model.score_topk``) with pinned HOST
            feature batches: H2D copy in, hit counters read back D2H, every step, inside the
            timed region.
``roofline`` the dominant kernel (tcgen05 GEMM + fused top-k) timed alone, FLOPs = 2*B*C*D.
``cpu_baseline`` / ``--impl reference``: the reference's own torch ops for this path
            (oracle port, fp32) on the box's host cores.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    "cfg2": dict(B=512, C=21841, D=1024, name="ImageNet-21K zero-shot head: 21,841 classes, 12-level synthetic hierarchy, RN50 dim 1024, batch 512"),
    "cfg4": dict(B=1024, C=10450, D=512, name="ImageNet-21K-P split: 10,450 classes, ViT-B/32 dim 512, batch 1024"),
    "cfg5": dict(B=4096, C=21841, D=1024, name="class-sharded sweep: 21,841 classes x batch 4096, dim 1024"),
}
K = 20
METRIC = "scored images/sec @21,841 classes"


def _config(wl, gpus):
    """`config` of both arms (identical keys and values, so that the driver can match them); implementation details of
    the GPU arm live in the separate `impl` object."""
    return {"workload": wl["name"], "B": wl["B"], "C": wl["C"], "D": wl["D"], "K": K,
            "l2": "GPU arm: inputs larger than L2 -- %d bank copies and 8 feature batches rotated, never the same "
                  "operands in consecutive steps" % _n_bank(wl, gpus)}


def _n_bank(wl, gpus):
    if gpus > 1:
        return 2
    return max(2, -(-int(1.6 * 126e6) // (wl["C"] * wl["D"] * 2)))


def _graph_steps(steps):
    """Batches per CUDA graph of the class-sharded evaluator: a divisor of `steps`, so that exactly `steps` are timed.
    16-20 batches on 8 exchange channels measured best at N = 8 (45.5 us per step against 49.5 with 10 batches on 4)."""
    for d in (16, 20, 18, 12, 14, 10, 8, 9, 7, 6, 5, 4, 3, 2):
        if steps % d == 0:
            return d
    return steps if 2 <= steps <= 32 else 8


CHANNELS = 8   # independent exchange channels of the class-sharded evaluator (at most half the batches of a graph)


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops", 1590.0), d.get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json, burst)"
    return 1590.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.rows = []
        self.proc = None
        self.gpu = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for n, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------ CPU reference arm
def _cpu_steps(wl, steps, warmup, seed=0):
    """The reference's torch ops for this path on the host: row-normalise + `feats @ bank.T`
    (model/clip_tree.py:330-331), column select + topk(20) + id map + eq + per-k sums (main.py:136-147)."""
    import torch
    from oracle import hgr_oracle as orc
    from hgrnet_b200.synthetic import synthetic_embeddings
    # all the host threads the box offers: torchrun exports OMP_NUM_THREADS=1 to every rank, which would turn the
    # CPU arm into a single-core run
    try:
        ncpu = len(os.sched_getaffinity(0))
    except AttributeError:
        ncpu = os.cpu_count() or 1
    if torch.get_num_threads() < ncpu:
        torch.set_num_threads(ncpu)
    B, C, D = wl["B"], wl["C"], wl["D"]
    bank = orc.normalize_rows(synthetic_embeddings(C, D, seed + 1))
    test_index = torch.arange(C)
    feats = [synthetic_embeddings(B, D, seed + 10 + i, normalize=False) for i in range(2)]
    targets = torch.randint(0, C, (1,)).expand(B).contiguous()
    hits_tot = {k: 0 for k in orc.TOPK}

    def one(i):
        logits = orc.forward_logits(feats[i % 2], bank)
        _, _, h = orc.eval_hits(logits, test_index, targets)
        for k in h:
            hits_tot[k] += h[k]

    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    for i in range(steps):
        one(i)
    dt = time.perf_counter() - t0
    return dt / steps, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    wl = WORKLOADS[args.workload or ("cfg2" if args.gpus == 1 else "cfg5")]
    steps = max(1, args.steps)
    sec, cores = _cpu_steps(wl, steps, max(1, args.warmup))
    value = wl["B"] / sec
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
            "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": _config(wl, args.gpus),
            "cpu_baseline": {"value": value, "unit": "images/s", "cores": cores, "kind": "port",
                             "sample": "%d batches of %d images (one batch per step), torch CPU fp32" % (steps, wl["B"])},
            "e2e": {"value": value, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch
    import torch.distributed as dist
    from hgrnet_b200 import _cabi, ops
    from hgrnet_b200.dist import ShardedScorer, shard_bounds
    from hgrnet_b200.head import tree_model
    from hgrnet_b200.flags import parse_args
    from hgrnet_b200.hierarchy import scaled_levels, synthetic_hierarchy
    from hgrnet_b200.synthetic import TableEncoder, node_id_tokens, synthetic_embeddings

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if args.gpus > 1 and world == 1:
            raise SystemExit("bench.py --gpus %d must be launched with torch.distributed.run (one rank per GPU)" % args.gpus)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    wl_key = args.workload or ("cfg2" if world == 1 else "cfg5")
    wl = WORKLOADS[wl_key]
    B, C, D = wl["B"], wl["C"], wl["D"]
    steps, warmup = args.steps, max(3, args.warmup)
    tf_peak, hbm_peak, peak_src = _peaks()

    # ---- synthetic hierarchy + class bank through the public module surface (kernel 1)
    hier = synthetic_hierarchy(scaled_levels(C), seed=1)
    table = synthetic_embeddings(C, D, 1, normalize=False)
    opts = parse_args([])
    opts.device, opts.folder, opts.weights = local_rank, "/tmp/hgr_bench_out_%d" % rank, "equal"
    model = tree_model(opts, hier.nodes, hier.nodes, clip_model=TableEncoder(table).to(dev), hierarchy=hier,
                       node_tokens=node_id_tokens(C))
    model.eval()
    model.update_classifier()
    torch.cuda.synchronize()

    # inputs larger than L2 (126 MB): rotate over several bank copies and feature batches
    n_bank = _n_bank(wl, world)
    lo, hi = shard_bounds(C, world)[rank]
    shard = model.bank_test[lo:hi]
    shard_ids = model._test_index_i32[lo:hi].contiguous()          # node id of every bank row of this shard
    banks = [shard.clone() for _ in range(n_bank)]
    n_feat = 8
    feats_host = [synthetic_embeddings(B, D, 100 + i, normalize=False).pin_memory() for i in range(n_feat)]
    # resident inputs are kept as bf16: the synthetic embeddings are bf16-valued, so this is the
    # same data in half the bytes -- what the normalise kernel (clip_tree.py:330) then has to read
    feats_dev = [f.to(dev).to(torch.bfloat16) for f in feats_host]
    g = torch.Generator().manual_seed(7)
    labels_host = [torch.full((B,), int(torch.randint(0, C, (1,), generator=g)), dtype=torch.long).pin_memory()
                   for _ in range(n_feat)]
    labels_dev = [l.to(dev).to(torch.int32) for l in labels_host]
    hits = ops.new_hits(dev)
    scorers = [ShardedScorer(b, shard_ids, id_base=lo, K=K) for b in banks] if world > 1 else None

    pending = []

    def step_eager(i):
        x = ops.normalize_rows(feats_dev[i % n_feat])
        if world == 1:
            ops.score_topk(x, banks[i % n_bank], col_id=shard_ids, targets=labels_dev[i % n_feat], K=K, hits=hits)
        else:
            sc = scorers[i % n_bank]
            sc.submit(x, labels_dev[i % n_feat])
            pending.append(sc)
            if len(pending) > 1:           # merge batch i-1 while batch i's GEMM / gather are in flight
                pending.pop(0).collect(hits)

    def drain():
        while pending:
            pending.pop(0).collect(hits)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n, flush=None, begin=None):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if begin:
            begin()                                 # worker streams start after e0
        for i in range(n):
            fn(i)
        if flush:
            flush()                                 # ... and are joined into this stream before e1
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t)
        return ms

    def timed_blocks(fn, n, flush=None, begin=None, min_ms=60.0, max_blocks=400):
        """`value` protocol: the block of exactly `n` steps is timed (device events, max over ranks) repeatedly until
        >= min_ms of device time has been covered; the MEDIAN block decides.  A single 20-step block of this head is
        under a millisecond -- too short for a stable clock."""
        first = timed(fn, n, flush, begin)
        blocks = [first]
        reps = int(min(max_blocks, max(2, math.ceil(min_ms / max(first, 1e-3)))))
        if world > 1:
            t = torch.tensor([reps], device=dev)
            dist.broadcast(t, 0)
            reps = int(t)
        for _ in range(reps - 1):
            blocks.append(timed(fn, n, flush, begin))
        blocks.sort()
        return blocks[len(blocks) // 2], len(blocks), blocks[0], blocks[-1]

    def graphed(fn, n_variants, n_streams=1):
        """One CUDA graph per rotation index, replayed round-robin on `n_streams` streams (a step is a single
        cudaGraphLaunch, no interpreter on the path).  Returns (step, begin, end)."""
        sts = [torch.cuda.Stream() for _ in range(n_streams)]
        cur = torch.cuda.current_stream()
        graphs = []
        for st in sts:
            st.wait_stream(cur)
        for i in range(n_variants):
            with torch.cuda.stream(sts[i % n_streams]):
                fn(i)                                   # warm-up outside capture (per-stream workspace, attributes)
        torch.cuda.synchronize()
        for i in range(n_variants):
            st = sts[i % n_streams]
            with torch.cuda.stream(st):
                gr = torch.cuda.CUDAGraph()
                with torch.cuda.graph(gr, stream=st):
                    fn(i)
            graphs.append(gr)
        torch.cuda.synchronize()

        def step(i):
            with torch.cuda.stream(sts[(i % n_variants) % n_streams]):
                graphs[i % n_variants].replay()

        def begin():
            for st in sts:
                st.wait_stream(torch.cuda.current_stream())

        def end():
            for st in sts:
                torch.cuda.current_stream().wait_stream(st)
        return step, begin, end

    cycle = n_bank * n_feat // math.gcd(n_bank, n_feat)
    n_streams = args.streams
    l0 = _cabi.launch_count()
    step_eager(0)
    drain()
    kernels_per_step = _cabi.launch_count() - l0
    multi_graph = False

    # ---- device-resident throughput: the public streaming evaluator without host I/O
    if world == 1:
        es_res = model.make_eval_stream(batch=B, slots=cycle, streams=n_streams, banks=banks, host_io=False,
                                        feat_dtype=torch.bfloat16)
        for s_ in range(cycle):
            es_res.dev_feats[s_].copy_(feats_dev[s_ % n_feat])
            es_res.dev_labels[s_].copy_(labels_dev[s_ % n_feat])
        step_resident = lambda i: es_res.step(i % cycle)
        res_begin, res_end, hits_src = es_res.begin, es_res.end, es_res.hits
    else:
        from hgrnet_b200.dist import ShardedEvalStream
        G_STEPS = _graph_steps(steps)
        from hgrnet_b200.dist import PeerMemoryUnavailable
        try:
            ses = ShardedEvalStream(banks[0], lo, batch=B, K=K, steps=G_STEPS, banks=banks, exchange=args.exchange,
                                    col_id=shard_ids, channels=args.channels, feat_dtype=torch.bfloat16)
        except PeerMemoryUnavailable as e:      # raised on every rank alike: fall back together
            if rank == 0:
                print("bench: %s -- falling back to the NCCL exchange" % (e,), file=sys.stderr)
            args.exchange = "nccl"
            ses = ShardedEvalStream(banks[0], lo, batch=B, K=K, steps=G_STEPS, banks=banks, exchange="nccl", col_id=shard_ids,
                                    feat_dtype=torch.bfloat16)
        for s_ in range(G_STEPS):
            ses.dev_feats[s_].copy_(feats_dev[s_ % n_feat])
            ses.dev_labels[s_].copy_(labels_dev[s_ % n_feat])
        multi_graph = ses.graph is not None
        # our kernels per batch on this path (one eager issue of the 8-batch pipeline, every rank alike)
        l0 = _cabi.launch_count()
        with torch.cuda.stream(ses.stream):
            ses._issue()
        torch.cuda.synchronize()
        kernels_per_step = (_cabi.launch_count() - l0) // G_STEPS

        def step_resident(i):                    # one replay = G_STEPS batches; issue it on every G_STEPS-th step
            if i % G_STEPS == 0:
                ses.run()
        steps = max(G_STEPS, steps // G_STEPS * G_STEPS)            # == args.steps whenever it has a divisor <= 10
        warmup_run = -(-warmup // G_STEPS) * G_STEPS                 # at least `warmup` steps, whole replays
        res_begin, res_end, hits_src = (lambda: None), (lambda: None), ses.hits
    for i in range(warmup_run if world > 1 else warmup):
        step_resident(i)
    res_end()
    torch.cuda.synchronize()
    hits_src.zero_()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, n_blocks, ms_best, ms_worst = timed_blocks(step_resident, steps, res_end, res_begin)
    launches = kernels_per_step * steps
    ms_per_step = ms / steps
    value = B / (ms_per_step * 1e-3)
    hits_resident = (ses.all_reduce_hits() if world > 1 else hits_src).tolist()   # p2p: per-rank row blocks, summed once
    hits_resident = [h // n_blocks for h in hits_resident]                        # per block of `steps` batches

    # ---- dominant kernel alone (GEMM + fused top-k, no merge), one stream: the roofline figure
    Cs = hi - lo
    nomerge = ops.HGR_IMPL_TCGEN05 | _cabi.HGR_IMPL_FLAG_NO_MERGE
    xs = [ops.normalize_rows(f) for f in feats_dev]

    global_cert = world > 1 and getattr(ses, "certify", "local") == "global"
    list_len = ops.score_topk_plan(B, Cs, D, K)["list_len"] if Cs > 0 else K
    if global_cert:
        # the kernel a rank really runs: lists sized for the GLOBAL certificate (hgr_score_topk_scatter_bounded);
        # NO_MERGE: the scoring kernel alone, nothing is scattered (the block tables are not dereferenced)
        list_len = ops.global_list_len(B, Cs, D, K, C)
        dummy = [xs[0].data_ptr()] * world
        blk = (B + world - 1) // world

    def kern_eager(i):
        if global_cert:
            ops.score_topk_scatter(xs[i % n_feat], ses.banks[i % n_bank], dummy, dummy, blk, K=K, impl=nomerge,
                                   bound_block_ptrs=dummy, C_total=C)
        else:
            ops.score_topk(xs[i % n_feat], banks[i % n_bank], K=K, impl=nomerge)

    # ONE graph holding `cycle` back-to-back launches (rotating inputs): a launch's duration, not a graph launch's
    kst = torch.cuda.Stream()
    kst.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(kst):
        kern_eager(0)
        kst.synchronize()
        kgraph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(kgraph, stream=kst):
            for i in range(cycle):
                kern_eager(i)
    torch.cuda.current_stream().wait_stream(kst)
    n_rep = max(3, steps // cycle)
    for i in range(max(1, warmup // cycle)):
        kgraph.replay()
    kms = timed(lambda i: kgraph.replay(), n_rep) / (n_rep * cycle)
    flops = 2.0 * B * Cs * D
    achieved = flops / (kms * 1e-3) / 1e12

    # ---- sustained loop (>= ~1.5 s) so that the clock / throttle samples mean something
    n_sus = int(min(2_000_000, max(steps, 1.5 / (ms_per_step * 1e-3))))
    sus_ms = timed(step_resident, n_sus, res_end, res_begin) / n_sus
    clocks = sampler.stop() if rank == 0 else None

    # ---- local (single-GPU, no collective) timing helpers
    def time_graph_us(fn, variants, min_ms=40.0):
        """`variants` back-to-back calls of fn(i) in ONE CUDA graph on a side stream, replayed for >= min_ms: us per call."""
        st = torch.cuda.Stream()
        st.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(st):
            fn(0)
            st.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr, stream=st):
                for i in range(variants):
                    fn(i)
        torch.cuda.current_stream().wait_stream(st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            gr.replay()
        e0.record()
        gr.replay()
        e1.record()
        torch.cuda.synchronize()
        reps = int(max(3, min_ms / max(e0.elapsed_time(e1), 1e-3)))
        e0.record()
        for _ in range(reps):
            gr.replay()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / (reps * variants) * 1e3

    def single_gpu_point(Bx, full_bank, ids):
        """The same head on THIS GPU alone with the whole bank: streaming evaluator (device-resident inputs) and the
        dominant kernel alone.  No collective inside: the other ranks wait at the next barrier."""
        nb = _n_bank(dict(C=full_bank.shape[0], D=D), 1)
        fb = [full_bank] + [full_bank.clone() for _ in range(nb - 1)]
        from hgrnet_b200.stream import EvalStream
        slots = nb * n_feat // math.gcd(nb, n_feat)
        es1 = EvalStream(full_bank, col_id=ids, batch=Bx, K=K, slots=slots, streams=n_streams, banks=fb, host_io=False,
                         feat_dtype=torch.bfloat16)       # same resident dtype as the sharded run it is compared with
        gx = torch.Generator().manual_seed(5)
        fx = [torch.randn(Bx, D, generator=gx).to(dev).to(torch.bfloat16) for _ in range(n_feat)]
        for s_ in range(slots):
            es1.dev_feats[s_].copy_(fx[s_ % n_feat])
            es1.dev_labels[s_].fill_(int(ids[s_ % ids.numel()]))
        for i in range(2 * slots):
            es1.step(i % slots)
        es1.end()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        blocks = []
        n_it = max(steps, slots)
        for _ in range(12):
            e0.record()
            es1.begin()
            for i in range(n_it):
                es1.step(i % slots)
            es1.end()
            e1.record()
            torch.cuda.synchronize()
            blocks.append(e0.elapsed_time(e1) / n_it)
        blocks.sort()
        msx = blocks[len(blocks) // 2]
        xn = [ops.normalize_rows(f) for f in fx]
        kus = time_graph_us(lambda i: ops.score_topk(xn[i % n_feat], fb[i % nb], K=K, impl=nomerge), slots)
        fl = 2.0 * Bx * full_bank.shape[0] * D
        del es1, fb
        return {"value": Bx / (msx * 1e-3), "unit": "images/s", "ms_per_step": msx, "B": Bx, "C": int(full_bank.shape[0]),
                "kernel_us": kus, "kernel_frac_of_peak": fl / (kus * 1e-6) / 1e12 / tf_peak}

    same_wl = None
    if rank == 0 and (world > 1 or wl_key == "cfg2"):
        # cfg 5 (B = 4096, the whole 21,841-class bank) on one GPU: the 1-GPU point of the scaling sweep
        same_wl = single_gpu_point(WORKLOADS["cfg5"]["B"], model.bank_test, model._test_index_i32)
    if world > 1:
        dist.barrier()

    # ---- bank row order (N = 1): the scoring kernel + merge on an i.i.d. bank, on a hierarchy-ordered (clustered)
    # bank as the reference's `nodes` order would give it, and on the same bank in the pseudo-random row order that
    # update_classifier applies; plus the floor-sketch kernel (exact by construction) on the un-permuted bank
    bank_order = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        from hgrnet_b200.synthetic import clustered_bank, near_leaf_features
        cb = clustered_bank(C, D, 3)
        cbd = cb.to(dev).to(torch.bfloat16)
        perm = torch.randperm(C, generator=torch.Generator().manual_seed(9)).to(dev)
        cbp = cbd[perm].contiguous()
        qs = [ops.normalize_rows(near_leaf_features(cb, B, 20 + i).to(dev)) for i in range(4)]
        iid = banks[0]

        def one(bank_, impl_):
            us_ = time_graph_us(lambda i: ops.score_topk(qs[i % 4], bank_, K=K, impl=impl_), 8, min_ms=20.0)
            ops.score_topk(qs[0], bank_, K=K, impl=impl_)
            torch.cuda.synchronize()
            return us_, ops.last_rescan_count()
        a_us, a_rs = one(iid, ops.HGR_IMPL_TCGEN05)
        b_us, b_rs = one(cbp, ops.HGR_IMPL_TCGEN05)
        c_us, c_rs = one(cbd, ops.HGR_IMPL_TCGEN05)
        d_us, d_rs = one(cbd, ops.HGR_IMPL_TCGEN05_SKETCH)
        bank_order = {"unit": "us per call (scoring kernel + merge, one stream), near-leaf queries",
                      "iid_bank": a_us, "clustered_bank_permuted_rows": b_us, "clustered_bank_tree_order": c_us,
                      "clustered_bank_tree_order_sketch_kernel": d_us,
                      "rows_repaired": {"iid": a_rs, "permuted": b_rs, "tree_order": c_rs},
                      "permuted_over_iid": b_us / a_us,
                      "note": "tree_model.update_classifier stores the test bank in the permuted order (head.py)"}
        del cbd, cbp

    # ---- BASELINE cfg 3 (N = 1): one OM training step of the head -- batch 256, --sample_strategy topk, out 0.25 / in 0.5,
    # adaptive weights, deepest-level target of the same 12-level hierarchy (T = 17 iterations) -- through
    # tree_model.train_batch (kernel 1, tcgen05 logits, fused masked CE, tcgen05 backward), with the reference's loop
    # (oracle port, fp32) timed on the host beside it
    om = None
    if world == 1 and rank == 0 and not args.no_cpu_baseline and wl_key == "cfg2":
        import random as _random
        import numpy as _np
        from oracle import hgr_oracle as orc
        from hgrnet_b200.levels import layer_weight_init
        o3 = parse_args([])
        o3.device, o3.folder = local_rank, "/tmp/hgr_bench_out_om"
        o3.weights, o3.out_ratio, o3.in_ratio, o3.k, o3.num_compare, o3.weighting, o3.scale = "adaptive", 0.25, 0.5, 1, 256, "both", 1.0
        table3 = (table * 0.05).to(torch.bfloat16).float()
        m3 = tree_model(o3, hier.nodes, hier.nodes, clip_model=TableEncoder(table3).to(dev), hierarchy=hier,
                        node_tokens=node_id_tokens(C)).to(dev)
        B3 = 256
        img3 = synthetic_embeddings(B3, D, 82, normalize=False)
        tgt = max(range(C), key=lambda i: len(hier.c2p[i]))
        x3 = img3.to(dev).requires_grad_(True)
        t3 = torch.full((B3,), tgt, dtype=torch.long, device=dev)
        _random.seed(17)
        for _ in range(3):
            loss3 = m3.train_batch(x3, t3, "OM", "topk")
        torch.cuda.synchronize()
        ts = []
        for _ in range(15):
            t0 = time.perf_counter()
            loss3 = m3.train_batch(x3, t3, "OM", "topk")
            torch.cuda.synchronize()
            ts.append(time.perf_counter() - t0)
        ts.sort()
        _random.seed(17)
        t0 = time.perf_counter()
        n_cpu = 3
        for _ in range(n_cpu):
            ref3 = orc.om_step(img3, table3, torch.tensor(float(_np.log(1 / 0.07))), hier.c2p, hier.d2n, tgt, out_ratio=0.25,
                               in_ratio=0.5, weights="adaptive", weighting="both", k=1, num_compare=256,
                               layer_weight=layer_weight_init(hier.d2n, 1.0))
        cpu_s = (time.perf_counter() - t0) / n_cpu
        # the head's own device work of such a step (two normalisations, tcgen05 logits, fused masked CE, tcgen05
        # backward) as one CUDA graph: T = 17 sets of 257 classes over a union of 3,000
        rs = _np.random.RandomState(0)
        U3 = 3000
        sets3 = [rs.permutation(U3)[:257] for _ in range(17)]
        sp3 = torch.tensor(_np.concatenate([[0], _np.cumsum([len(s_) for s_ in sets3])]).astype(_np.int32), device=dev)
        sc3 = torch.tensor(_np.concatenate(sets3).astype(_np.int32), device=dev)
        lp3 = torch.tensor(rs.randint(0, 257, 17).astype(_np.int32), device=dev)
        w3 = torch.rand(17, device=dev)
        xr3, tr3 = torch.randn(B3, D, device=dev), torch.randn(U3, D, device=dev)

        def head3(i):
            xn_, xnorm_ = ops.normalize_rows(xr3, return_norm=True)
            tn_, tnorm_ = ops.normalize_rows(tr3, return_norm=True)
            lg_ = ops.logits_dense(xn_, tn_, scale=14.2857)
            _, dl_ = ops.masked_ce(lg_, sp3, sc3, lp3, w3)
            ops.om_backward(dl_, lg_, xn_, xnorm_, tn_, tnorm_, 14.2857)
        head_us = time_graph_us(head3, 4, min_ms=20.0)
        om = {"workload": "OM training step: --sample_strategy topk, in_ratio 0.5, out_ratio 0.25, adaptive weights, batch 256, "
                          "RN50 dim 1024, T = %d iterations" % len(m3.last_losses),
              "ms_per_step": ts[len(ts) // 2] * 1e3, "steps_per_s": 1.0 / ts[len(ts) // 2], "loss": loss3,
              "head_kernels_us_per_step": head_us,
              "cpu_baseline": {"ms_per_step": cpu_s * 1e3, "kind": "port", "cores": torch.get_num_threads(),
                               "sample": "%d steps of the reference loop (oracle port), torch CPU fp32" % n_cpu,
                               "loss": float(ref3["loss"])},
              "note": "ms_per_step = wall clock of tree_model.train_batch incl. the host-side sampling (random.sample draws "
                      "identical to the reference's), three host syncs and the stand-in encoder's autograd; "
                      "head_kernels_us_per_step = the head's own kernels of such a step, one CUDA graph replay"}
        del m3

    # ---- end to end through the public API with pinned HOST buffers
    # Host features are bf16: the synthetic embeddings ARE bf16-valued (the reference's GPU
    # encoder emits 2-byte features too, clip/model.py:371-392), so nothing is lost and half the PCIe bytes of an fp32
    # copy are saved -- PCIe is what bounds this number.  The fp32 variant is measured next to it (`e2e_fp32_features`).
    def run_e2e(feat_dtype):
        esz = torch.empty((), dtype=feat_dtype).element_size()
        if world == 1:
            es = model.make_eval_stream(batch=B, slots=cycle, streams=n_streams, banks=banks, feat_dtype=feat_dtype)
            for s_ in range(cycle):
                es.host_feats[s_].copy_(feats_host[s_ % n_feat].to(feat_dtype))
                es.host_labels[s_].copy_(labels_host[s_ % n_feat].to(torch.int32))
            for i in range(warmup):
                es.step(i % cycle)   # graph: H2D feats+labels -> normalise -> score/top-20/Hit@k -> D2H hit counters
            es.end()
            ms_ = timed_blocks(lambda i: es.step(i % cycle), steps, es.end, es.begin)[0] / steps
            h2d, d2h = B * D * esz + B * 4, 5 * 8
        elif args.exchange == "p2p":
            # host-fed sharded evaluator: per batch every rank copies ITS block of image rows (+ labels) from pinned
            # host memory, NVLink replicates the normalised rows, Hit@k counters are read back after every batch
            ses_h = ShardedEvalStream(banks[0], lo, batch=B, K=K, steps=G_STEPS, banks=banks, host_io=True,
                                      col_id=shard_ids, channels=args.channels, feat_dtype=feat_dtype)
            for s_ in range(G_STEPS):
                ses_h.host_feats[s_].copy_(feats_host[s_ % n_feat][ses_h.row_lo:ses_h.row_hi].to(feat_dtype))
                ses_h.host_labels[s_].copy_(labels_host[s_ % n_feat][ses_h.row_lo:ses_h.row_hi].to(torch.int32))

            def step_h(i):
                if i % G_STEPS == 0:
                    ses_h.run()
            for i in range(warmup_run):
                step_h(i)
            ms_ = timed_blocks(step_h, steps)[0] / steps
            ses_h.close()
            # whole job: every image row crosses PCIe once (on the rank that owns it)
            h2d, d2h = B * D * esz + B * 4, 5 * 8 * world
        else:
            e2e_out = [None]

            def step_n(i):
                f = feats_host[i % n_feat].to(dev, non_blocking=True)
                t = labels_host[i % n_feat].to(dev, non_blocking=True)
                x = model.encode_image_normalized(f)
                scorers[i % n_bank].score(x, t, hits)
                e2e_out[0] = hits.cpu()                      # D2H read of the step's result (Hit@k counters)
            for i in range(warmup_run):
                step_n(i)
            ms_ = timed_blocks(step_n, steps)[0] / steps
            h2d, d2h = world * (B * D * 4 + B * 8), 5 * 8 * world   # NCCL exchange: every rank copies the whole batch
        return {"value": B / (ms_ * 1e-3), "unit": "images/s", "ms_per_step": ms_, "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h, "host_feature_dtype": str(feat_dtype).replace("torch.", "")}

    if world > 1 and args.exchange != "p2p":
        e2e, e2e32 = run_e2e(torch.float32), None
    else:
        e2e, e2e32 = run_e2e(torch.bfloat16), run_e2e(torch.float32)

    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sec, cores = _cpu_steps(wl, 40 if wl_key == "cfg2" else 8, 3)
            cpu = {"value": wl["B"] / sec, "unit": "images/s", "cores": cores, "kind": "port",
                   "sample": "%d batches of %d images, torch CPU fp32 (reference ops of clip_tree.py:330-331 + main.py:136-147)"
                             % (40 if wl_key == "cfg2" else 8, wl["B"])}
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak" if world == 1 else "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": _config(wl, world),
            "impl": {"sharding": "none" if world == 1 else "class dimension row-sharded over %d ranks" % world,
                     "streams": n_streams if world == 1 else 1,
                     "pipeline": ("EvalStream: 1 CUDA graph per batch on %d round-robin streams" % n_streams) if world == 1
                                 else ("ShardedEvalStream: %d batches per %s; %s" % (
                                     G_STEPS, "CUDA graph" if multi_graph else "eager issue",
                                     "peer-memory exchange: final lists stored into the row owner's buffer over NVLink, "
                                     "flag-ordered, owner merges its rows (no collective on the data path)"
                                     if args.exchange == "p2p" else
                                     "NCCL all-gather of batch i overlaps the GEMM of batch i+1")),
                     "exchange": None if world == 1 else args.exchange,
                     "resident_features": "bfloat16 [B, D], unnormalised (the synthetic embeddings are bf16-valued)",
                     "lists": ("%d-entry lists per (row, worker), sized for the row's GLOBAL stream and certified by the row "
                               "owner against the global K-th value (hgr_score_topk_scatter_bounded / "
                               "hgr_topk_merge_certified); rows repaired on rank 0 in this run: %d"
                               % (list_len, int(ses.repairs.item()))) if global_cert
                              else "%d-entry lists per (row, worker), certificate + exact repair inside the call" % list_len,
                     "bank_rows": "pseudo-random order (tree_model.update_classifier), col_id maps back to node ids",
                     "timing": "value = median of %d blocks of %d steps (fastest %.4f ms, slowest %.4f ms per block)"
                               % (n_blocks, steps, ms_best, ms_worst),
                     "warmup_steps_run": warmup_run if world > 1 else warmup},
            "e2e": e2e,
            "e2e_fp32_features": e2e32,
            "gpu_launches": int(launches),
            "sustained": {"value": B / (sus_ms * 1e-3), "unit": "images/s", "steps": n_sus, "ms_per_step": sus_ms},
            "hits": hits_resident,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s",
                         "frac": achieved / tf_peak,
                         # DRAM bytes of one launch from the committed `ncu --set full` capture of this workload
                         # (profiles/r01c_ncu_summary.md); compulsory bytes = bank + features + lists
                         "traffic": _traffic(B, Cs, D),
                         "traffic_unit": "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, offline capture "
                                         "under profiles/ -- a profiler cannot run inside the timed region)",
                         "compulsory_bytes": Cs * D * 2 + B * D * 2,
                         "kernel": "score_umma_pair_kernel (TMA + tcgen05 GEMM + fused top-20)",
                         "kernel_ms": kms, "flops_per_launch": flops, "peak_source": peak_src},
            "cpu_baseline": cpu,
            "clocks": clocks,
        }
        if same_wl is not None and world > 1:
            line["single_gpu_same_workload"] = same_wl
            line["speedup_vs_1gpu"] = value / same_wl["value"]
        elif same_wl is not None:
            line["cfg5_single_gpu"] = same_wl
        if bank_order is not None:
            line["bank_order"] = bank_order
        if om is not None:
            line["om_step_cfg3"] = om
        print(json.dumps(line))
    if world > 1:
        # CUDA graphs hold NCCL kernels: tear down without destroy_process_group (which can block on them)
        sys.stdout.flush()
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)
    return 0


# dram__bytes_read.sum + dram__bytes_write.sum of ONE launch of the dominant kernel, keyed by (B, C_local, D); from the
# `ncu --set full` captures summarised in profiles/r02_ncu_summary.md and, for the shard shapes in the list lengths of
# the global certificate, profiles/r02b_ncu_certified.md (a profiler cannot run inside the timed region).
# Shards of one bank differ by a row or two: _traffic() matches C within 2 rows.
NCU_DRAM_BYTES = {(512, 21841, 1024): 45837056 + 22272, (4096, 21841, 1024): 53206784 + 1389312,
                  (4096, 10921, 1024): 30800640 + 1536, (4096, 5461, 1024): 19621376 + 5888, (4096, 2731, 1024): 14041344,
                  (1024, 10450, 512): 11798016}


def _traffic(B, Cs, D):
    for (b, c, d), v in NCU_DRAM_BYTES.items():
        if b == B and d == D and abs(c - Cs) <= 2:
            return v
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2000)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=None, choices=[None, "cfg2", "cfg4", "cfg5"])
    ap.add_argument("--no-cpu-baseline", dest="no_cpu_baseline", action="store_true")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N > 1: peer-memory exchange (default) or NCCL all-gather of the per-rank candidate lists")
    ap.add_argument("--channels", type=int, default=CHANNELS,
                    help="N > 1: independent exchange channels of the class-sharded evaluator (measured 2/4/6/8: 74.8 / 82.8 / 88.1 / 90.0 M images/s at N = 8 in round 2a)")
    ap.add_argument("--streams", type=int, default=6, help="round-robin CUDA streams of the streaming evaluator (measured 2/3/4/6: 14.2 / 15.6 / 16.6 / 17.4 M images/s at cfg 2)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
