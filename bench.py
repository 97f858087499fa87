#!/usr/bin/env python3
"""Benchmark of the node-depth hot path (`fgfa depth -d`: depth + depth.uniq).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C|B|E] [--impl ours|reference]

One process per GPU (under torchrun for N > 1).  A "step" is one full pass of the hot
path over the synthetic graph: zero the outputs, the engine the plan chose -- window (pre-pass
S1-S3, kernel W, kernel B2) or stream (kernel A, kernel B) -- and, for N > 1, the exchange of
[depth | uniq].  The workload is
BASELINE.json configs[2] (the 400M-step graph the metric is quoted on); at N > 1 the SAME
graph is sharded by whole paths (configs[3]), so scaling is "strong".

`value`     steps/s with the shard resident in HBM (CUDA events, max over ranks).
`e2e`       the same metric through the host-buffer entry point: pinned host steps ->
            H2D -> kernels (-> allreduce) -> D2H of depth/uniq, every step.
`roofline`  algorithmic bytes of the dominant kernel (W or A) / its CUDA-event duration inside
            the timed region, against MEASURED_PEAKS.json's HBM copy bandwidth.
`parity`    after the timed region rank 0 recomputes the WHOLE graph with the oracle and compares
            every depth and uniq value with what the GPUs hold; a mismatch exits non-zero.
`extra_configs`  BASELINE.json configs[1] (B) and configs[4] (E) measured outside the main timed
            region: B and E at N = 1, E again (one path per GPU) at N = 8.
`cpu_baseline`  the oracle (C port of depth.rs:15-39, single thread like the reference); its
                `path_parallel_variant` is a multi-threaded CPU figure that is NOT the reference's algorithm
            timed on this box, rank 0, N = 1 only.
`--impl reference` times that CPU port alone (the Rust reference cannot be built here).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "path steps/sec for fgfa depth (depth + depth.uniq)"
UNIT = "steps/s"
FALLBACK_HBM_GBS = 6650.0   # B200_PROFILING.md fallback, used only if MEASURED_PEAKS.json is absent


def peak_hbm():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_evt = index, [], threading.Event()

    def run(self):
        while not self._stop_evt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                for line in out.strip().split("\n"):
                    f = [x.strip() for x in line.split(",")]
                    if len(f) >= 8:
                        self.rows.append(f)
            except Exception:
                pass
            self._stop_evt.wait(0.1)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=6)
        sm = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def time_oracle(cfg, steps, start, end, reps):
    """Oracle (kind "port") on the full workload: returns best steps/s over `reps`."""
    import oracle_lib as O
    best = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        rc, d, u = O.depth_with_uniq(steps, start, end, cfg.n_segs)
        dt = time.perf_counter() - t0
        assert rc == 0
        best = max(best, cfg.n_steps / dt)
    return best


def time_parallel_variant(cfg, steps, start, end, reps=2):
    """The path-parallel CPU variant (oracle_seg_depth_with_uniq_parallel), reported BESIDE the
    baseline: it is not the reference's algorithm (depth.rs:15-39 is single-threaded), SURVEY 8d
    asks for it as an optional, fairer CPU figure.  Never raises: the baseline proper does not depend on it."""
    try:
        import oracle_lib as O
        threads = max(1, min(os.cpu_count() or 1, 64, cfg.n_paths))
        best = 0.0
        for _ in range(reps):
            t0 = time.perf_counter()
            rc, d, u = O.depth_with_uniq_parallel(steps, start, end, cfg.n_segs, threads)
            dt = time.perf_counter() - t0
            if rc != 0:
                return None
            best = max(best, cfg.n_steps / dt)
        return {"value": best, "unit": UNIT, "threads": threads,
                "note": "whole paths per thread, private counters + reduction; NOT the reference's algorithm"}
    except Exception as exc:   # noqa: BLE001
        return {"unavailable": type(exc).__name__}


def run_reference(args, cfg, rank):
    """--impl reference: the reference's algorithm on the host CPU.  The Rust crate cannot
    be built here (no cargo, crates not vendored), so this is the oracle port of
    flatgfa/src/ops/depth.rs:15-39; like the reference loop it is single-threaded."""
    if rank != 0:
        return
    from pollen_b200 import synth
    steps, start, end = synth.make_graph(cfg)
    import oracle_lib as O
    times = []
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        rc, d, u = O.depth_with_uniq(steps, start, end, cfg.n_segs)
        dt = time.perf_counter() - t0
        assert rc == 0
        if i >= args.warmup:
            times.append(dt)
    total = sum(times)
    v = cfg.n_steps * len(times) / total
    sample = f"full config {cfg.name} ({cfg.n_steps} steps) per step, {len(times)} timed passes"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u64",
        "data": "synthetic", "config": workload_config(cfg, args.gpus),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count(),
                         "path_parallel_variant": time_parallel_variant(cfg, steps, start, end)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def workload_config(cfg, n_gpus):
    """Identical for the product arm and the reference arm (the driver compares them)."""
    return {
        "workload": f"{cfg.name}: {cfg.description}",
        "n_segs": cfg.n_segs, "n_paths": cfg.n_paths, "n_steps": cfg.n_steps,
        "generator": "haplotype walk (SURVEY.md 8d), seed 0xB1011054" if cfg.kind == 0 else "see pollen_b200/csrc/synth.cpp",
        "sharding": "single GPU" if n_gpus == 1 else f"whole paths LPT-partitioned over {n_gpus} GPUs + allreduce of [depth|uniq]",
        "l2_policy": "inputs larger than L2 (per-GPU steps shard >= 200 MB vs 126 MB L2); no explicit flush",
    }


def measure_extra(cfg, world, rank, dev, steps, peak):
    """One of the other BASELINE.json configs on this job's GPUs: steps/s, fraction of the HBM
    roofline for the whole step, engine, and an oracle check of the result (rank 0)."""
    import torch
    import torch.distributed as dist
    from pollen_b200 import sharding, synth

    start, end = synth.make_spans(cfg.n_paths, cfg.n_steps, cfg.jitter_pct)
    parts = sharding.lpt_partition(end - start, world)
    if world == 1:
        h, ls, le = synth.make_graph(cfg)
    else:
        h, ls, le = synth.make_graph(cfg, path_subset=parts[rank])
    d_steps = torch.from_numpy(h.view(np.int32)).to(dev)
    eng = sharding.ShardedDepth(ls, le, cfg.n_segs, dev, n_paths_global=cfg.n_paths if world > 1 else None)
    stream = torch.cuda.current_stream(dev)
    engine = eng.plan.autotune(d_steps, stream.cuda_stream)
    for _ in range(3):
        eng.run(d_steps, stream)
    eng.status()
    torch.cuda.synchronize(dev)
    if world > 1:
        dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(steps):
        eng.run(d_steps, stream)
    b.record(stream)
    torch.cuda.synchronize(dev)
    t = torch.tensor([a.elapsed_time(b) / steps], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    eng.status()
    gd, gu = eng.results()
    ok = None
    if rank == 0:
        import oracle_lib as O
        f_steps, f_start, f_end = (h, ls, le) if world == 1 else synth.make_graph(cfg)
        rc, od, ou = O.depth_with_uniq(f_steps, f_start, f_end, cfg.n_segs)
        ok = rc == 0 and bool((gd == od).all()) and bool((gu == ou).all())
    alg = 4.0 * cfg.n_steps + 8.0 * cfg.n_paths + 8.0 * cfg.n_segs
    out = {"workload": f"{cfg.name}: {cfg.description}", "n_gpus": world, "steps": steps, "ms_per_step": ms,
           "value": cfg.n_steps / (ms * 1e-3), "unit": UNIT, "engine": engine,
           "whole_step_frac": alg / world / (ms * 1e-3) / 1e9 / peak, "parity_vs_oracle": ok,
           "exchange": "none" if world == 1 else "nccl all_reduce of [depth u32 | uniq u8]"}
    del eng, d_steps
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="C")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--e2e-steps", type=int, default=0, help="0 = --steps")
    ap.add_argument("--no-extra", action="store_true", help="skip extra_configs (B, E)")
    args = ap.parse_args()

    from pollen_b200 import synth
    cfg = synth.CONFIGS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    import torch
    import torch.distributed as dist
    import pollen_b200 as pb
    from pollen_b200 import sharding

    if not torch.cuda.is_available() or pb.device_count() == 0:
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cpu_group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        cpu_group = dist.new_group(backend="gloo")     # a barrier that does not spin on the GPUs
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"

    # ---- this rank's shard of the graph ------------------------------------------------
    start, end = synth.make_spans(cfg.n_paths, cfg.n_steps, cfg.jitter_pct)
    parts = sharding.lpt_partition(end - start, world)
    my_paths = parts[rank]
    if world == 1:
        h_steps_np, ls, le = synth.make_graph(cfg)
    else:
        h_steps_np, ls, le = synth.make_graph(cfg, path_subset=my_paths)
    n_local = int(h_steps_np.size)
    h_steps = torch.from_numpy(h_steps_np.view(np.int32)).pin_memory()      # pinned host copy for e2e
    d_steps = torch.empty(n_local, dtype=torch.int32, device=dev)
    d_steps.copy_(h_steps)
    eng = sharding.ShardedDepth(ls, le, cfg.n_segs, dev, n_paths_global=cfg.n_paths if world > 1 else None)
    engine_name = eng.plan.autotune(d_steps, torch.cuda.current_stream(dev).cuda_stream)
    exchange = "none (single GPU)" if world == 1 else "nccl all_reduce of [depth u32 | uniq u8]" if eng.compact else "nccl all_reduce of [depth u32 | uniq u32]"
    if world > 1 and cfg.n_paths <= 255 and os.environ.get("FGFA_EXCHANGE", "fused") == "fused":
        # Preferred exchange: popcount fused with a reduce-scatter/all-gather over NVLink peer
        # memory (kernel X).  It is checked against the NCCL engine once, here, and dropped if
        # symmetric memory is unavailable on this box or the results differ.
        # Candidates, each checked against the NCCL engine once, here, and dropped if symmetric memory is
        # unavailable on this box or the results differ; the fastest by a start-up probe on the real shard runs:
        #   pull  kernel X: popcount fused with a reduce-scatter / all-gather by peer LOADS + multicast stores
        #   push  kernels P + R: partial depth and u8 uniq counts travel as peer STORES, owners reduce locally
        st0 = torch.cuda.current_stream(dev)

        def probe(e, reps=8):
            for _ in range(3):
                e.run(d_steps, st0)
            torch.cuda.synchronize(dev)
            dist.barrier()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(st0)
            for _ in range(reps):
                e.run(d_steps, st0)
            b.record(st0)
            torch.cuda.synchronize(dev)
            return a.elapsed_time(b) / reps

        names = {"pull": "fused popcount + reduce-scatter/all-gather over NVLink peer memory (kernel X: peer loads%s)",
                 "push": "fused popcount + push of the partials into the slice owners + local reduce + all-gather over NVLink peer memory (kernels P + R: peer stores%s)",
                 "push-window": "window engine per rank, then push of [depth u32 | uniq u8] into the slice owners + local reduce + all-gather over NVLink peer memory (kernels P + R: peer stores%s)"}
        forms = [f for f in os.environ.get("FGFA_FUSED_FORMS", "pull,push,push-window").split(",") if f in names]
        if engine_name != "window":
            forms = [f for f in forms if f != "push-window"]
        try:
            eng.run(d_steps, st0)
            torch.cuda.synchronize(dev)
            cands, bad_any = {}, 0.0
            for form in forms:
                f = sharding.FusedShardedDepth(ls, le, cfg.n_segs, dev, [len(p) for p in parts], form=form.split("-")[0],
                                               local_engine="window" if form == "push-window" else "stream")
                f.run(d_steps, st0)
                torch.cuda.synchronize(dev)
                same = torch.equal(eng.depth, f.depth) and torch.equal(eng.uniq, f.uniq)
                cands[form] = (f, same)
            t = torch.tensor([probe(eng)] + [probe(cands[f][0]) for f in forms] + [0.0 if cands[f][1] else 1.0 for f in forms],
                             dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            vals = [float(x) for x in t.tolist()]
            t_nccl, t_forms, bad = vals[0], vals[1:1 + len(forms)], vals[1 + len(forms):]
            ok = [(tf, f) for tf, f, bd in zip(t_forms, forms, bad) if bd == 0.0]
            probe_txt = "; probe: nccl %.3f ms" % t_nccl + "".join(
                ", %s %.3f ms%s" % (f, tf, "" if bd == 0.0 else " (MISMATCH)") for f, tf, bd in zip(forms, t_forms, bad))
            if ok and min(ok)[0] < t_nccl:             # every rank takes the same decision
                best = min(ok)[1]
                eng = cands[best][0]
                exchange = names[best] % (", multicast stores" if eng.mc_ptr else "") + probe_txt
            else:
                exchange += probe_txt
            for f in forms:
                if cands[f][0] is not eng:
                    cands[f] = None
        except Exception as exc:  # symmetric memory not available: keep NCCL
            exchange += f" (fused exchange unavailable: {type(exc).__name__}: {exc})"[:300]
    stream = torch.cuda.current_stream(dev)
    if isinstance(eng, sharding.ShardedDepth):
        launches_per_step = eng.plan.launches(True)                    # S1-S3 + W + B2, or A + B
    elif eng.form == "push":
        launches_per_step = (eng.plan.launches(True) if eng.local_engine == "window" else 1) + 2    # ... + P + R
    else:
        launches_per_step = 2                                          # A + X
    engine_name = eng.plan.engine

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    # ---- device-resident throughput ----------------------------------------------------
    for _ in range(args.warmup):
        eng.run(d_steps, stream)
    eng.status()
    probes = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for a, b in probes:            # torch creates the CUDA event lazily, on first record
        a.record(stream)
        b.record(stream)
    # clocks are sampled from here to the end of the e2e measurement (the timed region alone
    # lasts ~15 ms, shorter than one nvidia-smi poll)
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    barrier()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev0.record(stream)
    for k in range(args.steps):
        eng.plan.set_probe(probes[k][0].cuda_event, probes[k][1].cuda_event)
        eng.run(d_steps, stream)
        marks[k].record(stream)
    ev1.record(stream)
    barrier()
    eng.status()
    ms_total = torch.tensor([ev0.elapsed_time(ev1)], dtype=torch.float64, device=dev)
    ms_kernel = torch.tensor([statistics.mean(a.elapsed_time(b) for a, b in probes)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ms_total, op=dist.ReduceOp.MAX)
        dist.all_reduce(ms_kernel, op=dist.ReduceOp.MAX)
    ms_per_step = float(ms_total.item()) / args.steps
    value = cfg.n_steps / (ms_per_step * 1e-3)
    per_step = [(ev0 if k == 0 else marks[k - 1]).elapsed_time(marks[k]) for k in range(args.steps)]   # this rank
    step_stats = {"median_ms": statistics.median(per_step), "best_ms": min(per_step), "worst_ms": max(per_step)}

    # checksum of the result (sum of depth must equal the number of steps)
    depth_sum = int(eng.depth.to(torch.int64).bitwise_and(0xFFFFFFFF).sum().item())
    assert depth_sum == cfg.n_steps, (depth_sum, cfg.n_steps)

    # ---- roofline of the dominant kernel (kernel W of the window engine / kernel A of the stream engine) ----
    peak, peak_src = peak_hbm()
    alg_bytes = 4.0 * n_local + 8.0 * len(my_paths) + 4.0 * cfg.n_segs   # per launch, this rank (slowest rank's time)
    k_ms = float(ms_kernel.item())
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": None, "kernel": "k_window_count" if engine_name == "window" else "k_step_stream_merged",
                "engine": engine_name, "kernel_ms": k_ms,
                "algorithmic_bytes_per_launch": alg_bytes, "peak_source": peak_src,
                "whole_step_frac": (4.0 * cfg.n_steps + 8.0 * cfg.n_paths + 8.0 * cfg.n_segs) / world / (ms_per_step * 1e-3) / 1e9 / peak}

    roofline["frac_of_nominal_8tbs"] = achieved / 8000.0
    try:   # DRAM traffic of one kernel-A launch from the committed `ncu --set full` capture (N=1, config C)
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f).get(f"{cfg.name}:{world}:{engine_name}")
        if t:
            roofline["traffic"] = t["dram_bytes_read"] + t["dram_bytes_write"]
            roofline["traffic_source"] = t["source"]
    except Exception:
        pass

    # ---- split: depth only (seg_depth, depth.rs:45-56) and the collective alone ---------
    def timed(fn, reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()
        barrier()
        a.record(stream)
        for _ in range(reps):
            fn()
        b.record(stream)
        barrier()
        t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    scratch_depth = torch.empty(cfg.n_segs, dtype=torch.int32, device=dev)
    depth_only_ms = timed(lambda: eng.plan.run(d_steps, scratch_depth, None, stream.cuda_stream), args.steps)
    if world > 1 and isinstance(eng, sharding.ShardedDepth):
        allreduce_ms = timed(lambda: sharding.allreduce_counts(eng.out), args.steps)
    elif world > 1 and getattr(eng, "local_engine", "stream") == "window":
        # push after the window engine: everything that is not the rank's own S1-S3 + W + B2 (u8 uniq) is exchange
        scratch_u8 = torch.empty((cfg.n_segs + 31) // 32 * 32, dtype=torch.uint8, device=dev)
        local_ms = timed(lambda: eng.plan.run(d_steps, scratch_depth, scratch_u8, stream.cuda_stream), args.steps)
        allreduce_ms = max(0.0, ms_per_step - local_ms)  # kernels P + R + two barriers
    elif world > 1:
        allreduce_ms = max(0.0, ms_per_step - k_ms)      # barriers + kernel X (or P + R) + bitmap reset
    else:
        allreduce_ms = 0.0
    eng.run(d_steps, stream)          # leave a valid result behind
    eng.status()
    split = {"depth_only_ms_per_step": depth_only_ms, "depth_only_steps_per_s": cfg.n_steps / (depth_only_ms * 1e-3),
             "exchange_ms": allreduce_ms, "exchange_bytes_per_rank": eng.exchange_bytes if world > 1 else 0,
             "stream_kernel_ms": k_ms, "per_step_rank0": step_stats}

    # ---- end to end: host buffers in, host results out ---------------------------------
    e2e_steps = max(1, args.e2e_steps or args.steps)
    if world == 1:
        # the drop-in C-ABI call on host buffers (allocates, uploads in pipelined groups,
        # runs, downloads, widens to u64) -- what a caller of the reference's op would use
        lib = pb.lib()
        d64 = np.empty(cfg.n_segs, np.uint64)
        u64 = np.empty(cfg.n_segs, np.uint64)
        s32, e32 = np.ascontiguousarray(ls, np.uint32), np.ascontiguousarray(le, np.uint32)

        def e2e_once():
            rc = lib.fgfa_seg_depth_with_uniq_steps(h_steps.data_ptr(), n_local, s32.ctypes.data, e32.ctypes.data,
                                                    len(s32), cfg.n_segs, d64.ctypes.data, u64.ctypes.data)
            assert rc == 0, rc
        e2e_once()
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_once()
        e2e_s = (time.perf_counter() - t0) / e2e_steps
        assert int(d64.sum()) == cfg.n_steps
        e2e_api = "fgfa_seg_depth_with_uniq_steps (C ABI, pinned host steps)"
    else:
        res_dev = (eng.out,) if isinstance(eng, sharding.ShardedDepth) else (eng.depth, eng.uniq)
        res_host = [torch.empty(t.shape, dtype=t.dtype).pin_memory() for t in res_dev]

        def e2e_once():
            d_steps.copy_(h_steps, non_blocking=True)
            eng.run(d_steps, stream)
            if rank == 0:                      # the table is printed by one process
                for h, t in zip(res_host, res_dev):
                    h.copy_(t, non_blocking=True)
            stream.synchronize()
        e2e_once()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_once()
        barrier()
        e2e_t = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
        e2e_s = float(e2e_t.item())
        e2e_api = "%s.run on pinned host shards (H2D + kernels + exchange + D2H on rank 0)" % type(eng).__name__
    e2e = {"value": cfg.n_steps / e2e_s, "unit": UNIT, "h2d_bytes_per_step": 4 * n_local + 8 * len(my_paths),
           "d2h_bytes_per_step": 5 * cfg.n_segs if (world > 1 and eng.compact) else 8 * cfg.n_segs, "ms_per_step": e2e_s * 1e3, "api": e2e_api, "steps": e2e_steps}

    # ---- end to end for the reference's normal caller: an mmapped .flatgfa file (memfile.rs:7-10,
    # file.rs:185-213) handed to the image entry point; the steps are PAGEABLE memory and go through
    # the library's pinned ring (host threads fill a slot while the previous one crosses PCIe)
    if world == 1 and not args.no_extra:
        try:
            import tempfile
            from pollen_b200 import flatgfa_io
            tmpdir = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else tempfile.gettempdir()
            path = os.path.join(tmpdir, f"fgfa_bench_{os.getpid()}.flatgfa")
            flatgfa_io.write_flatgfa(path, h_steps_np, ls, le, cfg.n_segs)
            try:
                img = np.memmap(path, dtype=np.uint8, mode="r")
                lib = pb.lib()

                def image_once():
                    rc = lib.fgfa_seg_depth_with_uniq(img.ctypes.data, img.size, d64.ctypes.data, u64.ctypes.data)
                    assert rc == 0, rc
                image_once()                      # faults the mapping in, sizes the ring
                reps = max(2, min(e2e_steps, 5))
                t0 = time.perf_counter()
                for _ in range(reps):
                    image_once()
                dt = (time.perf_counter() - t0) / reps
                assert int(d64.sum()) == cfg.n_steps
                e2e["pageable"] = {"value": cfg.n_steps / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": reps,
                                   "api": "fgfa_seg_depth_with_uniq on an mmapped .flatgfa image (C ABI, pageable host memory, warm page cache)",
                                   "h2d_bytes_per_step": 4 * n_local + 8 * len(my_paths), "d2h_bytes_per_step": 8 * cfg.n_segs}
                del img
            finally:
                os.unlink(path)
        except Exception as exc:   # noqa: BLE001  -- a full /dev/shm must not cost the headline line
            e2e["pageable"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}

    clocks = sampler.stop() if sampler else None

    # ---- parity against the oracle, in the driver-visible run: the whole graph, every segment ----
    eng.run(d_steps, stream)
    eng.status()
    g_depth, g_uniq = eng.results()
    parity = None
    if rank == 0:
        import oracle_lib as O
        if world == 1:
            f_steps, f_start, f_end = h_steps_np, ls, le
        else:
            f_steps, f_start, f_end = synth.make_graph(cfg)
        rc, od, ou = O.depth_with_uniq(f_steps, f_start, f_end, cfg.n_segs)
        ok = rc == 0 and bool((g_depth == od).all()) and bool((g_uniq == ou).all())
        parity = {"vs_oracle": ok, "n_gpus": world, "engine": engine_name,
                  "checked": f"depth and depth.uniq of all {cfg.n_segs} segments against the C oracle over the whole graph"}
    flag = torch.tensor([0 if (parity is None or parity["vs_oracle"]) else 1], dtype=torch.int32, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
    if int(flag.item()):
        if rank == 0:
            print(json.dumps({"metric": METRIC, "error": "PARITY MISMATCH against the oracle", "parity": parity}))
        raise SystemExit(3)

    # ---- end to end through the C ABI at N > 1: ONE process (rank 0) drives all N devices with
    # fgfa_depth_multi_run_host (host steps in, u64 tables out), the form a compiled `fgfa --gpus N`
    # uses; the other ranks wait on a CPU barrier so that no NCCL kernel spins on their GPUs
    if world > 1:
        torch.cuda.synchronize(dev)
        dist.barrier(group=cpu_group)
        if rank == 0:
            try:
                hp = torch.from_numpy(f_steps.view(np.int32)).pin_memory()
                md = pb.MultiDepth(list(range(world)), f_start, f_end, cfg.n_segs, cfg.n_steps, exchange="nccl")
                d64 = np.empty(cfg.n_segs, np.uint64)
                u64 = np.empty(cfg.n_segs, np.uint64)
                md.run_host_ptr(hp.data_ptr(), d64.ctypes.data, u64.ctypes.data)
                same = bool((d64 == od).all()) and bool((u64 == ou).all())
                t0 = time.perf_counter()
                for _ in range(e2e_steps):
                    md.run_host_ptr(hp.data_ptr(), d64.ctypes.data, u64.ctypes.data)
                dt = (time.perf_counter() - t0) / e2e_steps
                md.close()
                del hp
                e2e["per_rank_processes"] = {k: e2e[k] for k in ("value", "ms_per_step", "api", "h2d_bytes_per_step", "d2h_bytes_per_step")}
                e2e.update({"value": cfg.n_steps / dt, "ms_per_step": dt * 1e3, "steps": e2e_steps, "parity_vs_oracle": same,
                            "api": f"fgfa_depth_multi_run_host (C ABI, one process driving {world} GPUs, NCCL all-reduce, pinned host steps)",
                            "h2d_bytes_per_step": 4 * cfg.n_steps, "d2h_bytes_per_step": 5 * cfg.n_segs})
                if not same:
                    parity["vs_oracle"] = False
            except Exception as exc:   # noqa: BLE001 -- keep the per-rank number as the headline
                e2e["c_abi"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:300]}
            del f_steps
        dist.barrier(group=cpu_group)

    # ---- the other BASELINE configs, outside the main timed region ----------------------
    extra = None
    if not args.no_extra and cfg.name == "C":
        extra = {}
        del d_steps, h_steps
        torch.cuda.empty_cache()
        names = ["B", "E"] if world == 1 else (["E"] if world == 8 else [])
        for name in names:
            extra[f"{name}@{world}"] = measure_extra(synth.CONFIGS[name], world, rank, dev, max(5, args.steps), peak)

    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------
    cpu = None
    if rank == 0 and world == 1:
        v = time_oracle(cfg, h_steps_np, ls, le, reps=3)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"full config {cfg.name} ({cfg.n_steps} steps), best of 3 passes of the single-threaded oracle",
               "host_cores": os.cpu_count(),
               "path_parallel_variant": time_parallel_variant(cfg, h_steps_np, ls, le)}

    if rank == 0:
        print(json.dumps({
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(cfg, world), "exchange": exchange, "engine": engine_name, "roofline": roofline, "split": split, "cpu_baseline": cpu, "e2e": e2e, "parity": parity, "extra_configs": extra,
            "gpu_launches": launches_per_step * args.steps, "clocks": clocks,
        }))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
