#!/usr/bin/env python
"""bench.py -- the full-catalog scoring hot path on N B200s of one node.

A "step" is one pass of the hot path over one batch of synthetic input (BASELINE.json configs[2],
the configuration the metric's target is quoted on: 1M-item catalog per GPU, d=128, 4096 query
rows, bf16 operands / fp32 accumulate):

    item gather (4096x50 ids) -> fused full-catalog CE forward + dU (one sweep, two MMAs per tile)
    -> CE backward dW (second sweep) -> gather's scatter-add backward
    -> masked top-K evaluation of 4096 rows (K=50; one candidate sweep seeded by a 2.5 % prefix sweep).

metric = full-catalog scored user-item pairs / s = (train rows + eval rows) x catalog size / time.
With N GPUs the item table is row-sharded (1M rows per GPU => weak scaling; queries replicated);
the exchange steps are one all-reduce (the gathered input rows: global ids over the sharded table), one all-gather
(CE stats), one all-reduce (dU) and one all-gather (top-K).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

`--impl reference` times the reference's own CPU implementation of the path (the oracle port of
the reference lines; freerec is not installable here) on the host cores, on a bounded sample.
"""
from __future__ import annotations

import argparse
import gc
import json
import os
import statistics
import subprocess
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_ITEMS_PER_GPU = 1_000_000
ROWS = 4096
D = 128
SEQ = 50
TOPK = 50
METRIC = "full-catalog scored user-item pairs/sec (CE train + top-K eval)"
# dram__bytes_read.sum + dram__bytes_write.sum of one pair_kernel<PASS_DW> launch at this workload (ncu --set full)
PROFILED_TRAFFIC_BYTES = 261_439_488 + 208_969_472
PROFILED_TRAFFIC_SOURCE = "profiles/r2_ncu_summary.md, final capture (ncu --set full, one launch of pair_kernel<PASS_DW>, N=1M shard: dram read + write)"
UNIT = "pairs/s"
WORKLOAD = ("configs[2]: SASRec bf16 full-softmax CE train + masked top-50 eval, "
            "1M-item catalog per GPU (row-sharded), d=128, 4096 query rows, gather 4096x50")


# ------------------------------------------------------------------------------ CPU arm
def cpu_sample(rows: int, n_items: int, d: int, reps: int, seed: int = 2026):
    """The reference lines (oracle port) on the host cores: CE fwd+bwd of `rows` queries and a
    masked top-K/metrics batch of `rows` queries against `n_items` items, fp32."""
    import torch
    from oracle import reference_path as orc

    torch.set_num_threads(os.cpu_count() or 1)
    orc.TOPK_IMPL = "torch"  # time what the reference runs: one torch.topk per metric@k
    g = torch.Generator().manual_seed(seed)
    U = torch.randn(rows, d, generator=g) / d ** 0.25
    W = torch.randn(n_items, d, generator=g) / d ** 0.25
    labels = torch.randint(0, n_items, (rows,), generator=g)
    seen = [torch.randint(0, n_items, (25,), generator=g).tolist() for _ in range(rows)]
    crow, col = orc.lists_to_csr(seen)
    tcrow, tcol = orc.lists_to_csr([[int(x)] for x in labels])
    mons = ["HITRATE@1", "HITRATE@5", "HITRATE@10", "HITRATE@20", "HITRATE@50", "NDCG@5", "NDCG@10", "NDCG@20", "NDCG@50"]
    times = []
    for _ in range(reps):
        t0 = time.perf_counter()
        orc.ce_fwd_bwd(U, W, labels)
        orc.evaluate_batch(orc.score_dense(U, W), crow, col, tcrow, tcol, mons)
        times.append(time.perf_counter() - t0)
    return times, 2 * rows * n_items


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    rows, n_items = 2048, 200_000
    for _ in range(min(args.warmup, 2)):
        cpu_sample(rows, n_items, D, 1)
    times, pairs = cpu_sample(rows, n_items, D, args.steps)
    total = sum(times)
    value = pairs * args.steps / total
    cores = os.cpu_count() or 1
    sample = f"{rows} train rows + {rows} eval rows x {n_items} items, d={D}, fp32, per step"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows": ROWS, "n_items_total": N_ITEMS_PER_GPU * max(args.gpus, 1), "d": D, "topk": TOPK,
                   "note": "reference arm = oracle port of the reference's PyTorch lines (fp32) on the host CPU; every step is "
                           "a bounded sample of this workload: " + sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.proc, self.path = index, None, f"/tmp/rb_clocks_{os.getpid()}.csv"

    def start(self):
        """Start nvidia-smi and wait for its first sample: its start-up (NVML initialisation) can stall kernel
        launches for tens of milliseconds, so it must be over before anything is timed."""
        try:
            self.f = open(self.path, "w")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=self.f, stderr=subprocess.DEVNULL)
            t0 = time.time()
            while time.time() - t0 < 5.0 and os.path.getsize(self.path) == 0:
                time.sleep(0.02)
        except Exception:
            self.proc = None

    def mark(self):
        """Samples before this point (idle GPU) are dropped by stop()."""
        try:
            self.skip = sum(1 for _ in open(self.path))
        except Exception:
            self.skip = 0

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.f.close()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for i, ln in enumerate(open(self.path)):
            if i < getattr(self, "skip", 0):
                continue
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0])); mx.append(float(p[1])); power.append(float(p[2]))
            except ValueError:
                continue
            for n, v in zip(names, p[3:7]):
                if v == "Active":
                    reasons.add(n)
        try:
            os.unlink(self.path)
        except OSError:
            pass
        if sm:
            out = {"sm_mhz": statistics.median(sm), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm), "power_w_max": max(power)}
        return out


def run_ours(args):
    # rank 0 prints ONE JSON line on stdout.  Libraries that log to file descriptor 1 (NCCL prints its version banner
    # there at NCCL_DEBUG=VERSION and above) must not get in front of it: fd 1 is pointed at stderr for the whole run
    # and the line goes to a duplicate of the original stdout.
    sys.stdout.flush()
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    import torch
    import torch.distributed as dist

    from recboard_b200 import _lib as L
    from recboard_b200 import metrics as MX
    from recboard_b200 import ops, sharded, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device: recboard_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # rank 0 prints ONE JSON line on stdout: whatever NCCL logs (its version banner at NCCL_DEBUG=VERSION and above)
        # goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    L.lib()

    n_total = N_ITEMS_PER_GPU * world
    row_start, row_end = sharded.shard_bounds(n_total, world, rank)
    n_shard = row_end - row_start

    # ---- parameters (resident): this rank's shard of the item table as ONE bf16 parameter (n_shard + 1 rows, row 0 =
    #      padding, SASRec/main.py:70-77) with its gradient buffer kept allocated, as a training loop that calls
    #      optimizer.zero_grad(set_to_none=False) has it
    gw = torch.Generator(device=dev).manual_seed(1000 + rank)
    table = torch.cat([torch.zeros(1, D, dtype=torch.bfloat16, device=dev),
                       synth.embeddings(n_shard, D, gw, dev, torch.bfloat16, gain=1.5)]).requires_grad_(True)
    table.grad = torch.zeros_like(table)
    W = table.detach()[1:]                                            # the scored view weight[NUM_PADS:] (:193)
    peers = None
    if world > 1 and os.environ.get("RB_BENCH_GATHER", "peer") == "peer":
        try:
            peers = sharded.PeerTable(table, row_start, n_skip=1)
        except RuntimeError as e:   # raised on ALL ranks together; the collective-only gather (also CUDA + NCCL) takes over
            if rank == 0:
                print(f"bench: peer-memory gather unavailable, using the all-reduce variant: {e}", file=sys.stderr)
    # ---- one batch of inputs (identical on every rank: queries are replicated)
    g = torch.Generator(device=dev).manual_seed(2026 + 3)
    U_train = synth.embeddings(ROWS, D, g, dev, torch.bfloat16, gain=1.5)
    labels = synth.zipf_ids(ROWS, n_total, g, dev)
    seqs = synth.sequences(ROWS, SEQ, n_total, g, dev)               # GLOBAL item ids + 1, 0 = padding (lpad_, SASRec/main.py:150-154)
    gather_grad = synth.embeddings(ROWS * SEQ, D, g, dev, torch.bfloat16, gain=0.01).view(ROWS, SEQ, D)
    U_eval = synth.embeddings(ROWS, D, g, dev, torch.bfloat16, gain=1.5)
    seen_crow, seen_col = synth.seen_csr(ROWS, n_total, g, dev)
    tgt = synth.targets(ROWS, n_total, g, dev, (seen_crow, seen_col))
    tgt_crow = torch.arange(ROWS + 1, device=dev, dtype=torch.int64)
    monitors = ["HITRATE@1", "HITRATE@5", "HITRATE@10", "HITRATE@20", "HITRATE@50", "NDCG@5", "NDCG@10", "NDCG@20", "NDCG@50"]
    # Random embeddings never rank a random held-out item among the top 50 of a million, so the reported metrics would
    # all be 0: every fourth row's target becomes the item this very model ranks 1st / 8th / 30th (cycling) -- the metric
    # values in the line then have a known answer (HITRATE@1 = 1/12, @10 = 2/12, @50 = 3/12 of the rows).
    with torch.no_grad():
        if world > 1:
            _, ids0 = sharded.sharded_topk(U_eval, W, TOPK, row_start, seen_crow, seen_col)
        else:
            _, ids0 = ops.topk_eval(U_eval, W, TOPK, seen_crow, seen_col)
    pick = torch.arange(0, ROWS, 4, device=dev)
    rank_of = torch.tensor([0, 7, 29], device=dev)[(pick // 4) % 3]
    tgt[pick] = ids0[pick, rank_of].long()
    n_r = [int((rank_of == v).sum()) for v in (0, 7, 29)]
    metrics_expected = {"HITRATE@1": n_r[0] / ROWS, "HITRATE@10": (n_r[0] + n_r[1]) / ROWS, "HITRATE@50": sum(n_r) / ROWS}

    def hot_path(U_tr, lab, sq, U_ev, s_crow, s_col):
        """One pass of the path through the public API (recboard_b200.ops / .sharded), gradients through autograd:
        the table's gradient -- the gather's scatter-add rows plus the scoring head's dW -- ends up in table.grad."""
        table.grad = None                                                         # optimizer.zero_grad() (set_to_none=True, torch's default)
        if world > 1:   # a2 over the row-sharded table, GLOBAL ids: rows are read where they live (own shard, or a peer's
            #              over NVLink through the IPC mapping); fence() is what follows an optimizer step in a real loop.
            #              RB_BENCH_GATHER=allreduce: owners gather + one all-reduce of the replicated rows instead.
            if peers is not None:
                peers.fence()
                emb = peers.gather(sq - 1, padding_idx=-1, accumulate=True)
            else:
                emb = sharded.sharded_gather_rows(table, sq - 1, row_start - 1, padding_idx=-1, accumulate=True)
        else:
            emb = ops.gather_rows(table, sq, padding_idx=0, accumulate=True)      # a2
        Uq = U_tr.detach().requires_grad_(True)
        if world > 1:
            loss = sharded.sharded_fused_ce(Uq, table, lab, row_start, n_skip=1, accumulate=True)   # a5+a6 (+ all-gather)
        else:
            loss = ops.fused_ce(Uq, table, lab, n_skip=1, accumulate=True)
        # a7 (+ all-reduce of dU) and a3: gather_grad stands in for what the encoder's backward hands to the gathered rows
        torch.autograd.backward([loss, emb], [None, gather_grad])
        if world > 1:
            vals, ids = sharded.sharded_topk(U_ev, W, TOPK, row_start, s_crow, s_col)  # a8-a10
        else:
            vals, ids = ops.topk_eval(U_ev, W, TOPK, s_crow, s_col)
        return loss, ids, emb

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- clock sampler first (its start-up must not overlap the timed region), then warm-up
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    # A GPU that has idled through process start-up and input generation boosts for a few hundred milliseconds
    # before the 1 kW power cap pulls the SM clock down (5.2-5.4 ms/step cool against 5.7-5.9 sustained): run the
    # step untimed ~2.3 s first (400 times), then the W warm-up steps, so that the timed region is the sustained
    # state.  (The 7-9 ms/step first regions earlier builds saw were not clocks but the allocator stall below.)
    # The warm-up loops keep the previous step's results alive while the next step is enqueued, exactly as the
    # timed loop does: otherwise the timed loop's second step needs fresh blocks from the caching allocator, and
    # that cudaMalloc blocked the launching thread for 2-80 ms (GPU idle meanwhile: 9 ms/step outliers).
    n_pre = int(os.environ.get("RB_BENCH_PREWARM", "400"))   # a fixed count (~2.3 s): every rank must issue the same
    # collectives.  RB_BENCH_PREWARM=0 is for the ncu launch-list pass of this command (tools/gpu_round.sh), where
    # every launch costs milliseconds of profiler overhead; never for a reported number.
    for _ in range(n_pre):
        loss, ids, _ = hot_path(U_train, labels, seqs, U_eval, seen_crow, seen_col)
    torch.cuda.synchronize()
    for i in range(max(args.warmup, 3)):
        loss, ids, _ = hot_path(U_train, labels, seqs, U_eval, seen_crow, seen_col)
        if i == 0:
            barrier()
            if rank == 0:
                sampler.mark()   # clocks are sampled under load from here on (warm-up + timed steps)
    barrier()

    # ---- device-resident timing of exactly K steps (inputs > L2: the 256 MB table shard is streamed
    #      from HBM several times per step, so no L2 flush is needed between iterations)
    launches0 = L.launch_count()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]   # one per step boundary
    barrier()
    gc.collect()
    gc.disable()   # no collector pause on the launching thread inside the timed regions
    mallocs0 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    host_ms = []
    barrier()
    evs[0].record()
    for i in range(args.steps):
        th = time.perf_counter()
        loss, ids, _ = hot_path(U_train, labels, seqs, U_eval, seen_crow, seen_col)
        evs[i + 1].record()
        host_ms.append(1e3 * (time.perf_counter() - th))
    barrier()
    if os.environ.get("RB_BENCH_DEBUG") and rank == 0:
        print("host enqueue ms:", " ".join(f"{x:.2f}" for x in host_ms), file=sys.stderr)
    ms = evs[0].elapsed_time(evs[-1])   # the K steps, first launch to last completion
    per_step = [evs[i].elapsed_time(evs[i + 1]) for i in range(args.steps)]
    if os.environ.get("RB_BENCH_DEBUG") and rank == 0:
        print("per-step ms:", " ".join(f"{x:.2f}" for x in per_step), file=sys.stderr)
    per_step.sort()
    step_ms = {"min": per_step[0], "median": statistics.median(per_step), "max": per_step[-1],
               "host_enqueue_ms_max": max(host_ms),
               "cuda_mallocs_in_region": torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - mallocs0}
    launches = L.launch_count() - launches0
    t = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t)
    pairs_per_step = 2 * ROWS * n_total
    value = pairs_per_step * args.steps / (ms * 1e-3)

    # ---- end-to-end: host buffers in, host results out, every step.  Written the way a training / evaluation
    #      loop is: the pinned inputs of step i+1 travel on a copy stream while step i computes; the step's results
    #      -- the loss value and the batch means of every configured metric@k, reduced on the device from the ranked
    #      ids (rb_topk_metrics) -- are copied out and consumed on the host inside the timed region, every step.
    pin = lambda x: x.cpu().pin_memory()
    hosts = [pin(x) for x in (U_train, labels, seqs, U_eval, seen_crow, seen_col)]
    h2d = sum(x.numel() * x.element_size() for x in hosts)
    d2h = 4 + len(monitors) * 4
    copy_stream = torch.cuda.Stream(device=dev)
    h_loss = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
    h_mets = [torch.empty(len(monitors), dtype=torch.float32).pin_memory() for _ in range(2)]
    done = [torch.cuda.Event() for _ in range(2)]

    # two sets of device input buffers: no allocation inside the loop (a cudaMalloc next to NCCL costs tens of ms)
    dev_in = [[torch.empty_like(h, device=dev) for h in hosts] for _ in range(2)]
    in_ready = [torch.cuda.Event() for _ in range(2)]
    step_done = [torch.cuda.Event() for _ in range(2)]

    def stage_inputs(i):
        j = i & 1
        with torch.cuda.stream(copy_stream):
            if i >= 2:
                copy_stream.wait_event(step_done[j])   # step i-2 has finished reading this buffer set
            for dst, src in zip(dev_in[j], hosts):
                dst.copy_(src, non_blocking=True)
            in_ready[j].record(copy_stream)

    def consume(slot):
        done[slot].synchronize()
        return float(h_loss[slot][0]), {m: float(v) for m, v in zip(monitors, h_mets[slot])}

    def e2e_loop(n):
        out = None
        dbg = os.environ.get("RB_E2E_DEBUG") and rank == 0
        tl = time.perf_counter()
        stage_inputs(0)
        for i in range(n):
            if dbg:
                now = time.perf_counter(); print(f"e2e iter {i}: {1e3 * (now - tl):.2f} ms since previous", file=sys.stderr); tl = now
            main = torch.cuda.current_stream()
            main.wait_event(in_ready[i & 1])
            if i + 1 < n:
                stage_inputs(i + 1)
            loss, ids, _ = hot_path(*dev_in[i & 1])
            step_done[i & 1].record(main)
            mets = MX.batch_metrics_device(ids, tgt_crow, tgt, monitors)   # a10: every metric@k of the batch in one pass
            h_loss[i & 1].copy_(loss.detach().reshape(1), non_blocking=True)
            h_mets[i & 1].copy_(mets, non_blocking=True)
            done[i & 1].record()
            if i > 0:
                out = consume((i - 1) & 1)
        return consume((n - 1) & 1) if n > 0 else out

    e2e_loop(4)
    barrier()
    mallocs1 = torch.cuda.memory_stats(dev).get("num_device_alloc", 0)
    t0 = time.perf_counter()
    loss_host, res = e2e_loop(args.steps)
    if os.environ.get("RB_E2E_DEBUG"):
        print(f"rank {rank}: before barrier {1e3 * (time.perf_counter() - t0):.2f} ms", file=sys.stderr)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_mallocs = torch.cuda.memory_stats(dev).get("num_device_alloc", 0) - mallocs1
    if os.environ.get("RB_E2E_DEBUG"):
        print(f"rank {rank}: e2e loop {1e3 * e2e_s:.2f} ms for {args.steps} steps", file=sys.stderr)
    clocks = sampler.stop() if rank == 0 else None   # sampled over both timed regions (device-timed and end-to-end)
    gc.enable()
    t = torch.tensor([e2e_s], device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = pairs_per_step * args.steps / float(t)

    # ---- per-kernel timing of the path's sweeps (CUDA events on the launching stream)
    def time_op(fn, reps=5):
        fn(); torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(reps):
            fn()
        b.record(); torch.cuda.synchronize()
        return a.elapsed_time(b) / reps

    breakdown = None
    roofline = None
    if rank == 0:
        Wd = W
        table_grad = torch.zeros(n_shard + 1, D, dtype=torch.bfloat16, device=dev)
        gather_flat = gather_grad.view(-1, D)
        with torch.no_grad():
            m_, l_, ll_ = ops.ce_rowstats(U_train, Wd, labels, label_base=row_start)
            lse = m_ + torch.log(l_)
            t_stats = time_op(lambda: ops.ce_rowstats(U_train, Wd, labels, label_base=row_start))
            t_fwd = time_op(lambda: ops.ce_rowstats(U_train, Wd, labels, label_base=row_start, want_dU=True))
            t_dW = time_op(lambda: ops.ce_backward(U_train, Wd, labels, lse, 1.0 / ROWS, label_base=row_start, need_dU=False,
                                                   need_dW=True, dw_dtype=torch.bfloat16))   # what the step runs: bf16 gradient rows
            t_topk = time_op(lambda: ops.topk_eval(U_eval, Wd, TOPK, seen_crow, seen_col, id_base=row_start))
            t_gather = time_op(lambda: ops.gather_rows_raw(table.detach(), seqs))
            t_scatter = time_op(lambda: ops.scatter_add_rows_(table_grad, gather_flat, seqs.view(-1), padding_idx=0))
        flop_tile = 2.0 * ROWS * n_shard * D   # one (rows x items x d) contraction
        breakdown = {
            "ce_fwd_dU_ms": t_fwd, "ce_bwd_dW_ms": t_dW, "ce_stats_only_ms": t_stats, "topk_ms": t_topk, "gather_ms": t_gather,
            "scatter_add_ms": t_scatter,
            "ce_train_algorithmic_tflops": 3 * flop_tile / ((t_fwd + t_dW) * 1e-3) / 1e12,
            "ce_train_executed_tflops": 4 * flop_tile / ((t_fwd + t_dW) * 1e-3) / 1e12,
            "topk_algorithmic_tflops": flop_tile / (t_topk * 1e-3) / 1e12,
            # one candidate sweep over the catalog + the seeding sweep over a prefix of max(4K, 2.5 %) of its tiles
            "topk_executed_tflops": (1 + min(1.0, max(4 * TOPK, -(-(n_shard // 128) // 40)) / (n_shard / 128))) * flop_tile / (t_topk * 1e-3) / 1e12,
            "gather_gbs": (ROWS * SEQ * (8 + 2 * D * 2)) / (t_gather * 1e-3) / 1e9,
        }
        peaks = {}
        try:
            peaks = json.load(open(ROOT / "MEASURED_PEAKS.json"))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops", 1590.0)
        # Dominant kernel: pair_kernel<PASS_DW> (items stationary: dW = G^T U, the longest launch of the step).
        # Algorithmic work of that launch = the dW GEMM, 2*M*N*d flop (6d flop per pair per train
        # step = 2d scores + 2d dU in the forward launch + 2d dW here); the launch also re-executes the score
        # GEMM (another 2*M*N*d, "executed_tflops"), which is not counted.  The timed launch sequence includes
        # the small finishing kernels of rb_ce_bwd (lse2, label scatter); they are < 3 % of it.
        ach = flop_tile / (t_dW * 1e-3) / 1e12
        roofline = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "traffic": PROFILED_TRAFFIC_BYTES, "kernel": "pair_kernel<PASS_DW> (dW = (softmax - onehot)^T U)",
                    "executed_tflops": 2 * ach,
                    "traffic_source": PROFILED_TRAFFIC_SOURCE,
                    "algorithmic_bytes": n_shard * D * 2 + n_shard * D * 2 + ROWS * D * 2,   # W in, bf16 dW out, U
                    "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst: the kernel is timed alone)" if peaks else "fallback 1.59 PFLOP/s"}
        if peaks.get("bf16_tflops_sustained"):   # for reference: the same launches against the sustained cuBLAS figure
            roofline["peak_sustained"] = peaks["bf16_tflops_sustained"]
            roofline["frac_of_sustained"] = ach / peaks["bf16_tflops_sustained"]

    # ---- fixed-size sharded workloads (north_star's scaling target, BASELINE configs[3] and [4]) and, on one GPU,
    #      the PyTorch-eager bar of the reference lines on this very B200 (bench_extra.py).  RB_BENCH_EXTRAS=0 skips them.
    strong = c5 = eager = None
    if os.environ.get("RB_BENCH_EXTRAS", "1") != "0":
        import bench_extra as BX
        table = W = Wd = table_grad = None   # free the headline workload's tensors
        gc.collect()
        torch.cuda.empty_cache()
        strong = BX.strong_scaling(dev, world, rank)
        c5 = BX.config5(dev, world, rank)
        if rank == 0 and world == 1:
            eager = BX.eager_baseline(dev)

    cpu_baseline = None
    if rank == 0 and world == 1:
        rows_s, n_s = 2048, 200_000   # ~10 s of host work in total
        cpu_sample(rows_s, n_s, D, 1)
        times, pairs = cpu_sample(rows_s, n_s, D, 4)
        cpu_baseline = {"value": pairs / statistics.median(times), "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                        "sample": f"{rows_s} train + {rows_s} eval rows x {n_s} items, d={D}, fp32 (oracle port of the reference lines), median of 4"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16", "data": "synthetic", "step_ms": step_ms,
            "config": {"workload": WORKLOAD,
                       "rows": ROWS, "n_items_total": n_total, "d": D, "topk": TOPK, "parallelism": f"row-sharded table x{world}",
                       "input_gather": ("local" if world == 1 else ("global ids, rows read from the owning GPU over NVLink (CUDA IPC peer memory)"
                                                                    if peers is not None else "global ids, owners gather + all-reduce")),
                       "l2": "inputs larger than L2 (256 MB table shard streamed per sweep); no flush",
                       "pre_warm": f"{n_pre} untimed steps (~2.3 s) before the {max(args.warmup, 3)} warm-up steps"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "loss": loss_host, "metrics": res, "metrics_expected": metrics_expected, "cuda_mallocs_in_region": e2e_mallocs,
                    "how": "pinned host inputs copied in, loss + the 9 metric@k batch means (computed on the device from the ranked "
                           "ids) copied out and consumed every step; the copy of step i+1 overlaps step i"},
            "gpu_launches": launches,
            "roofline": roofline, "breakdown": breakdown, "cpu_baseline": cpu_baseline,
            "strong_10m": strong, "config5_50m": c5, "gpu_eager_baseline": eager,
        }
        real_stdout.write(json.dumps(line) + "\n")
        real_stdout.flush()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={args.gpus}",
               "--master-addr", "127.0.0.1", "--master-port", "29511", __file__, "--gpus", str(args.gpus),
               "--steps", str(args.steps), "--warmup", str(args.warmup)]
        return subprocess.call(cmd)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
