#!/usr/bin/env python
"""bench.py — kNN queries/s on the north-star workload (BASELINE.json):
knn_predict on an 811,457 x 512 bank, k=200, t=0.1, 9 classes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--mode bf16|tf32x3|exact]
  python bench.py --impl reference ...      # the reference algorithm on the host CPU cores

A step = one knn_predict call over one batch of Q synthetic queries.
  value : queries/s with queries, bank (prepared, cached) and labels resident in HBM
  e2e   : the same through the public API with HOST (pinned) queries: H2D copy of the batch,
          knn_predict, D2H of the predicted class column, all inside the timed region
  N > 1 : the bank is row-sharded over the ranks (north-star mode 4); every rank holds the
          whole query batch; one all-gather of (Q,k) candidate keys + merge per step.
          Fixed total work as N grows -> "scaling": "strong".
One JSON line on rank 0 (contract in the task statement).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "self-supervised-wafermaps_b200"), os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

N_BANK, DIM, KNN_K, KNN_T, N_CLASSES = 811457, 512, 200, 0.1, 9
METRIC = "kNN queries/s @811k×512 bank, k=200"
WAVE = 148 * 128  # queries one resident wave of CTAs covers


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm=p["hbm_gbs"], bf16=p["bf16_tflops_sustained"], bf16_burst=p["bf16_tflops"], src="measured")
    return dict(hbm=6650.0, bf16=1400.0, bf16_burst=1590.0, src="fallback")


def ncu_traffic(mode, Q, N, world):
    """DRAM bytes (read+write) per launch of the dominant kernel from the committed
    `ncu --set full` capture of this exact configuration (profiles/ncu_traffic.json), else None."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        for e in json.load(open(path)):
            if e["mode"] == mode and e["Q"] == Q and e["N"] == N and e["n_gpus"] == world:
                return e["dram_bytes_per_launch"]
    except Exception:
        pass
    return None


def make_inputs(device, n_bank, n_query, dim, seed, row_range=None):
    """clustered synthetic embeddings generated on the device under test, in chunks of 65,536
    rows with one generator per chunk, so a rank can generate only the bank rows it owns
    (row_range) and results do not depend on the total size.  Labels are
    generated for all rows (they are replicated)."""
    import torch

    g0 = torch.Generator(device=device).manual_seed(seed)
    prior = torch.tensor([859, 111, 1037, 1936, 719, 30, 173, 239, 7345], dtype=torch.float64, device=device)
    if N_CLASSES != 9:  # MixedWM38 (config c3): flat prior
        prior = torch.ones(N_CLASSES, dtype=torch.float64, device=device)
    cent = torch.nn.functional.normalize(torch.randn(N_CLASSES, dim, generator=g0, device=device), dim=1)
    lo_own, hi_own = row_range if row_range is not None else (0, n_bank)

    def rows(n, stream_id, lo_want, hi_want):
        out = torch.empty(max(0, hi_want - lo_want), dim, device=device)
        lab = torch.empty(n, dtype=torch.int64, device=device)
        for ci, lo in enumerate(range(0, n, 65536)):
            hi = min(n, lo + 65536)
            g = torch.Generator(device=device).manual_seed(seed * 1000003 + stream_id * 100003 + ci)
            l = torch.multinomial(prior, hi - lo, replacement=True, generator=g)
            lab[lo:hi] = l
            a, b = max(lo, lo_want), min(hi, hi_want)
            if a < b:
                x = cent[l] + 1.4 * torch.randn(hi - lo, dim, generator=g, device=device) / dim ** 0.5
                out[a - lo_want:b - lo_want] = torch.nn.functional.normalize(x, dim=1)[a - lo:b - lo]
        return out, lab

    bank_nd, labels = rows(n_bank, 1, lo_own, hi_own)
    q, _ = rows(n_query, 2, 0, n_query)
    return bank_nd, labels, q


class ClockSampler:
    """Samples SM clocks and throttle reasons with one streaming `nvidia-smi -lms 50` process
    while the timed region runs (a fresh nvidia-smi per sample is too slow for sub-second runs)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.rows, self.index, self.proc = [], index, None
        self.th = threading.Thread(target=self.run, daemon=True)

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                parts = [x.strip() for x in line.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
        except Exception:
            pass

    def __enter__(self):
        self.th.start()
        time.sleep(0.15)  # let the first sample land before the timed region starts
        return self

    def __exit__(self, *a):
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        self.th.join(timeout=3)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_reference_rate(batch, n_calls, threads=None):
    """The reference algorithm (oracle R32: lightly's knn_predict restated op for op) on torch
    CPU fp32 with all host threads, on a bounded sample of the SAME workload: `n_calls` calls
    of `batch` queries against the full 811,457 x 512 bank (after one untimed warm-up call)."""
    import torch

    from oracle import knn_oracle as O

    threads = threads or os.cpu_count() or 1
    torch.set_num_threads(threads)
    g = torch.Generator().manual_seed(811)
    bank = torch.nn.functional.normalize(torch.randn(N_BANK, DIM, generator=g), dim=1).t().contiguous()
    labels = torch.randint(0, N_CLASSES, (N_BANK,), generator=g)
    q = torch.nn.functional.normalize(torch.randn(batch, DIM, generator=g), dim=1)
    O.knn_predict_r32(q, bank, labels, N_CLASSES, KNN_K, KNN_T)
    times = []
    for _ in range(n_calls):
        t0 = time.perf_counter()
        O.knn_predict_r32(q, bank, labels, N_CLASSES, KNN_K, KNN_T)
        times.append(time.perf_counter() - t0)
    return batch * n_calls / sum(times), times, threads


def cpu_name():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = 1024
    rate, times, threads = cpu_reference_rate(batch, args.warmup + args.steps)
    times = times[args.warmup:]
    rate = batch * len(times) / sum(times)
    sample = f"{len(times)} calls of B={batch} queries against the full {N_BANK}x{DIM} bank, k={KNN_K}"
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"knn_predict N={N_BANK} D={DIM} k={KNN_K} t={KNN_T} C={N_CLASSES}; B={batch} per step",
                   "reference": "lightly knn_predict restated (oracle R32), torch CPU fp32 (MKL), all host threads",
                   "cpu": cpu_name()},
        "cpu_baseline": {"value": rate, "unit": "queries/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    global DIM, KNN_K, N_BANK, N_CLASSES
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--mode", default=os.environ.get("B200KNN_BENCH_MODE", "bf16"))
    ap.add_argument("--queries", type=int, default=4 * WAVE, help="queries per step (default 75,776 = 4 waves)")
    ap.add_argument("--bank", type=int, default=N_BANK)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fp32-line", action="store_true", help="skip the secondary measurement of the bit-exact mode")
    ap.add_argument("--dim", type=int, default=DIM, help="vector dimension (default 512; 768 = config c5)")
    ap.add_argument("--k", type=int, default=KNN_K)
    ap.add_argument("--classes", type=int, default=N_CLASSES, help="9 = WM-811K priors; 38 = MixedWM38 (config c3)")
    args = ap.parse_args()
    DIM, KNN_K, N_BANK, N_CLASSES = args.dim, args.k, args.bank, args.classes
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist

    import b200knn
    from b200knn import _lib
    from b200knn import knn as K

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    Q, N, mode = args.queries, args.bank, args.mode
    b200knn.set_default_mode(mode)

    lo, hi = b200knn.shard_bounds(N, world, rank)
    bank_nd, labels, q = make_inputs(dev, N, Q, DIM, seed=811, row_range=(lo, hi) if world > 1 else None)
    if world > 1:
        shard = bank_nd.t().contiguous()  # this rank's (D, rows) slice, reference layout
        del bank_nd
        sb = b200knn.ShardedBank(shard, labels, N, mode=mode)
        bank = shard
        n_local = hi - lo

        def predict(qd):
            return sb.knn_predict(qd, N_CLASSES, KNN_K, KNN_T)
    else:
        bank = bank_nd.t().contiguous()  # (D,N) contiguous as knn.py:80 builds it
        del bank_nd
        n_local = N

        def predict(qd):
            return b200knn.knn_predict(qd, bank, labels, N_CLASSES, KNN_K, KNN_T)

    # bank preparation (cached per bank tensor; once per validation epoch in the reference) — timed apart
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    if mode != "exact":
        pbank = K.bank_cache.get(bank, K.RESCORED_MODES[mode]["cand"] if mode in K.RESCORED_MODES else mode)
        if mode in K.RESCORED_MODES:
            pbank.max_norm()
    e1.record()
    torch.cuda.synchronize()
    prepare_ms = e0.elapsed_time(e1)

    q_host = q.cpu().pin_memory()
    top1_host = torch.empty(Q, dtype=torch.int64).pin_memory()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        barrier()
        ms = torch.tensor([s.elapsed_time(e)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def step_resident():
        K.query_cache.clear()  # a new batch every step in real use: re-prepare (cast) the queries
        return predict(q)

    # e2e: every step uploads its own batch from pinned host memory and reads its predictions
    # back.  The upload of step i+1 is enqueued on a copy stream before step i's compute, so
    # it overlaps the kernels (double-buffered device staging); knn_predict itself synchronises
    # once per call (status word), and the D2H of the result closes the step.
    copy_stream = torch.cuda.Stream(device=dev)
    staging = [torch.empty_like(q), torch.empty_like(q)]
    uploaded = [torch.cuda.Event(), torch.cuda.Event()]
    consumed = [torch.cuda.Event(), torch.cuda.Event()]
    e2e_state = {"i": 0, "primed": False}

    # N > 1: every rank uploads only its 1/N slice of the batch over its own PCIe link and the
    # full batch is assembled with an all-gather over NVLink (every rank needs all queries)
    q_per = (Q + world - 1) // world
    q_lo, q_hi = min(Q, rank * q_per), min(Q, (rank + 1) * q_per)
    if world > 1:
        slice_host = torch.zeros((q_per, DIM), dtype=q.dtype).pin_memory()
        slice_host[: q_hi - q_lo] = q_host[q_lo:q_hi]
        slice_dev = [torch.empty((q_per, DIM), dtype=q.dtype, device=dev) for _ in range(2)]
        gathered = [torch.empty((q_per * world, DIM), dtype=q.dtype, device=dev) for _ in range(2)]
        staging = [g[:Q] for g in gathered]

    def upload(slot):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[slot])
            if world > 1:
                slice_dev[slot].copy_(slice_host, non_blocking=True)
                dist.all_gather_into_tensor(gathered[slot], slice_dev[slot])
            else:
                staging[slot].copy_(q_host, non_blocking=True)
            uploaded[slot].record(copy_stream)

    def step_e2e():
        i = e2e_state["i"]
        slot = i & 1
        if not e2e_state["primed"]:
            consumed[0].record()
            consumed[1].record()
            upload(slot)
            e2e_state["primed"] = True
        upload(slot ^ 1)  # next step's batch, overlapped with this step's kernels
        cur = torch.cuda.current_stream()
        cur.wait_event(uploaded[slot])
        K.query_cache.clear()
        pred = predict(staging[slot])
        consumed[slot].record(cur)
        top1_host.copy_(pred[:, 0], non_blocking=True)
        cur.synchronize()
        e2e_state["i"] = i + 1

    for _ in range(args.warmup):
        step_resident()
    K.profile_events = []
    _lib.launch_counter["kernels"] = 0
    with ClockSampler(local) as clocks:
        total_ms = timed(step_resident, args.steps)
    n_abi_kernels = _lib.launch_counter["kernels"]
    events = K.profile_events
    K.profile_events = None
    kern_ms = [a.elapsed_time(b) for a, b in events]
    for _ in range(max(1, args.warmup // 2)):
        step_e2e()
    e2e_ms = timed(step_e2e, args.steps)

    # result quality, outside the timed regions: this mode's neighbours and predictions for a
    # sample of the batch against the on-device exact mode (sequential-fma fp32, bit-checked
    # against the oracle by tests/): recall@k, and bitwise key equality for the fp32-matching modes
    quality = None
    rescore_stats, prepass_stats = dict(K.last_rescore_stats), dict(K.last_prepass_stats)
    if world == 1 and mode != "exact":
        nq = min(Q, 256)
        qs = q[:nq].contiguous()
        ek = b200knn.topk_keys(qs, bank, KNN_K, mode="exact")
        tk = b200knn.topk_keys(qs, bank, KNN_K, mode=mode)
        ei = b200knn.decode_keys(ek)[1].cpu().numpy()
        ti = b200knn.decode_keys(tk)[1].cpu().numpy()
        recall = sum(len(set(a.tolist()) & set(b.tolist())) for a, b in zip(ei, ti)) / float(nq * KNN_K)
        b200knn.set_default_mode("exact")
        pe = b200knn.knn_predict(qs, bank, labels, N_CLASSES, KNN_K, KNN_T)
        b200knn.set_default_mode(mode)
        pm = b200knn.knn_predict(qs, bank, labels, N_CLASSES, KNN_K, KNN_T)
        quality = {"sample_queries": nq, "recall_at_k": recall, "keys_bitwise_equal_exact_mode": bool(torch.equal(ek, tk)),
                   "top1_agreement_with_exact_mode": float((pe[:, 0] == pm[:, 0]).float().mean().item()),
                   "class_ranking_equal_exact_mode": bool(torch.equal(pe, pm))}

    # the fp32-matching ("bit-exact") mode on the same workload, reported beside the headline mode
    fp32_line = None
    if world == 1 and mode == "bf16" and not args.no_fp32_line:
        b200knn.set_default_mode("fp32")
        for _ in range(2):
            step_resident()
        n_fp = max(2, min(5, args.steps))
        fp_ms = timed(step_resident, n_fp) / n_fp
        st = dict(K.last_rescore_stats)
        ek = b200knn.topk_keys(q[:256].contiguous(), bank, KNN_K, mode="exact")
        fk = b200knn.topk_keys(q[:256].contiguous(), bank, KNN_K, mode="fp32")
        fp32_line = {"mode": "fp32", "dtype": "f16+f32", "value": Q / (fp_ms * 1e-3), "unit": "queries/s",
                     "ms_per_step": fp_ms, "steps": n_fp, "uncertified_rows_last_step": st["uncertified"],
                     "keys_bitwise_equal_exact_mode": bool(torch.equal(ek, fk)),
                     "note": "tensor-core candidates (fp16, 1 MMA per k-step; rows it cannot certify: fp16 x split-fp16, "
                             "2 MMAs) + exact sequential-fma re-scoring + per-row certificate: neighbours, similarities "
                             "and class ranking bit for bit those of the exact mode / oracle"}
        b200knn.set_default_mode(mode)

    phases = None
    if world > 1:  # where a sharded step spends its time (one extra untimed step, rank 0's view)
        sb.phase_log = []
        step_resident()
        torch.cuda.synchronize()
        log, sb.phase_log = sb.phase_log, None
        phases = {}
        for i in range(1, len(log)):  # same-named phases (cascade levels) are summed
            if log[i][0] != "start":
                phases[log[i][0]] = round(phases.get(log[i][0], 0.0) + log[i - 1][1].elapsed_time(log[i][1]), 3)

    if rank == 0:
        pk = peaks()
        ms_per_step = total_ms / args.steps
        value = Q / (ms_per_step * 1e-3)
        # time of the similarity + top-k calls of one step (main pass; in the fp32 cascade also the
        # small second-level pass over the rows the first level could not certify)
        kern = sum(kern_ms) / max(1, args.steps)
        flops = 2.0 * Q * n_local * DIM  # algorithmic: 2*N*D per query, this rank's rows
        achieved = flops / (kern * 1e-3) / 1e12
        cand_mode = K.RESCORED_MODES[mode]["cand"] if mode in K.RESCORED_MODES else mode
        k_plan = KNN_K + (K.RESCORED_MODES[mode]["margin"] if mode in K.RESCORED_MODES else 0)
        plan = b200knn.plan_info(Q, n_local, DIM, k_plan, cand_mode)
        if cand_mode == "bf16":
            roof = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16"], "unit": "TFLOP/s",
                    "frac": achieved / pk["bf16"], "traffic": None,
                    "kernel": "tc_topk_kernel<BF16,256,cta_group::2> main pass (one launch per step)",
                    "peak_source": f"{pk['src']} bf16 sustained (MEASURED_PEAKS.json)", "kernel_ms": kern}
        elif cand_mode == "bf16x3":
            roof = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16"], "unit": "TFLOP/s",
                    "frac": achieved / pk["bf16"], "executed_frac": 3 * achieved / pk["bf16"],
                    "traffic": None, "kernel": "tc_topk_kernel<BF16X3,128> candidates (3 bf16 MMAs per k-step)"
                    + (" + rescore_dot_kernel" if mode in K.RESCORED_MODES else ""),
                    "peak_source": f"{pk['src']} bf16 sustained (MEASURED_PEAKS.json); frac counts the useful "
                                   "2*N*D flops per query, executed_frac the 3 MMAs actually issued", "kernel_ms": kern}
        elif cand_mode == "f16":
            roof = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16"], "unit": "TFLOP/s",
                    "frac": achieved / pk["bf16"], "traffic": None,
                    "kernel": "tc_topk_kernel<F16,256,cta_group::2> candidates (1 fp16 MMA per k-step)"
                    + (" + rescore_dot_kernel" if mode in K.RESCORED_MODES else ""),
                    "peak_source": f"{pk['src']} bf16 sustained (MEASURED_PEAKS.json; fp16 runs at the bf16 rate)",
                    "kernel_ms": kern}
        elif cand_mode == "f16x2":
            roof = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16"], "unit": "TFLOP/s",
                    "frac": achieved / pk["bf16"], "executed_frac": 2 * achieved / pk["bf16"],
                    "traffic": None, "kernel": "tc_topk_kernel<F16X2,256,cta_group::2> candidates (2 fp16 MMAs per k-step)"
                    + (" + rescore_dot_kernel" if mode in K.RESCORED_MODES else ""),
                    "peak_source": f"{pk['src']} bf16 sustained (MEASURED_PEAKS.json; fp16 runs at the bf16 rate); frac "
                                   "counts the useful 2*N*D flops per query, executed_frac the 2 MMAs actually issued",
                    "kernel_ms": kern}
        elif cand_mode == "tf32x3":
            roof = {"bound": "tensor", "achieved": achieved, "peak": pk["bf16"] / 2, "unit": "TFLOP/s",
                    "frac": achieved / (pk["bf16"] / 2), "executed_frac": 3 * achieved / (pk["bf16"] / 2),
                    "traffic": None, "kernel": "tc_topk_kernel<TF32X3,128>" + (" (candidates; + rescore_kernel)" if mode in K.RESCORED_MODES else ""),
                    "peak_source": f"{pk['src']} bf16 sustained / 2 (tf32 runs at half the bf16 rate)",
                    "kernel_ms": kern}
        else:
            fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12
            roof = {"bound": "fp32-cuda-core", "achieved": achieved, "peak": fp32_peak, "unit": "TFLOP/s",
                    "frac": achieved / fp32_peak, "traffic": None, "kernel": "exact_topk_kernel",
                    "peak_source": "nominal 148 SM x 128 FMA/clk x 1.965 GHz", "kernel_ms": kern}
        # kernels launched through the C ABI inside the timed region (counted by the loader
        # proxy) + the split merges b200knn_topk* adds internally when its plan splits the bank
        splits_extra = 0
        if plan["splits"] > 1:
            splits_extra += args.steps
        if K.prepass_stride(N, k_plan) and mode != "exact":
            sp = b200knn.plan_info(Q, max(1, n_local // K.prepass_stride(N, k_plan)), DIM, K.PREPASS["r"], cand_mode)
            splits_extra += args.steps if sp["splits"] > 1 else 0
        gpu_launches = n_abi_kernels + splits_extra
        roof["traffic"] = ncu_traffic(mode, Q, N, world)
        line = {
            "metric": METRIC if (N, DIM, KNN_K) == (811457, 512, 200) else f"kNN queries/s @{N}x{DIM} bank, k={KNN_K}",
            "value": value, "unit": "queries/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "tf32x3": "tf32x3", "bf16x3": "bf16x3", "f16x2": "f16x2", "f16": "f16", "exact": "f32",
                      "fp32": "f16+f32", "fp32_f16": "f16+f32", "fp32_f16x2": "f16x2+f32", "fp32_bf16x3": "bf16x3+f32",
                      "fp32_tf32": "tf32x3+f32", "fp32_bf16": "bf16+f32"}[mode], "data": "synthetic",
            "config": {"workload": f"knn_predict N={N} D={DIM} k={KNN_K} t={KNN_T} C={N_CLASSES}; "
                                   f"Q={Q} queries per step (clustered synthetic, WM-811K class priors)",
                       "mode": mode, "bank_sharding": f"row-sharded over {world} GPU(s)" if world > 1 else "none",
                       "l2": "inputs larger than L2 (prepared bank %.0f MB)" % (n_local * DIM * {"bf16": 2, "f16": 2, "bf16x3": 4, "f16x2": 4, "tf32x3": 8, "exact": 4}[cand_mode] / 1e6),
                       "plan": plan, "bank_prepare_ms_excluded": prepare_ms},
            "roofline": roof,
            "e2e": {"value": Q / (e2e_ms / args.steps * 1e-3), "unit": "queries/s",
                    "h2d_bytes_per_step": Q * DIM * 4, "d2h_bytes_per_step": Q * 8 * world,
                    "note": "whole-job bytes; N>1: each rank uploads 1/N of the batch, NVLink all-gather assembles it",
                    "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": gpu_launches,
            "clocks": clocks.summary(),
        }
        if quality is not None:
            line["quality"] = quality
        if fp32_line is not None:
            line["fp32_mode"] = fp32_line
        if phases is not None:
            line["config"]["phases_ms"] = phases
            line["config"]["exchange"] = "all-to-all of (B,k) keys by query slice (fused into the top-k kernel over NVLink peer memory when available) + all-gather of (B,C) rankings (NCCL)" + (
                "; fp32: candidates routed to the shard owning each bank row for exact re-scoring and back (2 all-to-alls)" if mode in K.RESCORED_MODES else "")
        if mode in K.RESCORED_MODES:
            line["config"]["uncertified_rows_last_step"] = rescore_stats["uncertified"] if world == 1 else sb.last_uncertified
        line["config"]["prepass"] = {"stride": K.prepass_stride(N, k_plan), "r": K.PREPASS["r"],
                                     "repaired_rows_last_step": prepass_stats["repaired"]}
        if world == 1 and not args.no_cpu_baseline:
            rate, times, threads = cpu_reference_rate(1024, 12)
            line["cpu_baseline"] = {"value": rate, "unit": "queries/s", "cores": threads, "kind": "port",
                                    "cpu": cpu_name(),
                                    "sample": f"12 calls of B=1024 queries against the full {N_BANK}x{DIM} bank "
                                              f"(oracle R32 = lightly knn_predict restated, torch CPU fp32, "
                                              f"{sum(times):.1f} s of CPU work)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
