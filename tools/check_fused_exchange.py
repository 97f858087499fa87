#!/usr/bin/env python3
"""Multi-GPU check + timing of the fused popcount/exchange against the NCCL allreduce path.
  torchrun --nproc-per-node N tools/check_fused_exchange.py [config]"""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pollen_b200 import sharding, synth  # noqa: E402


def main():
    cfg = synth.CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "C"]
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    start, end = synth.make_spans(cfg.n_paths, cfg.n_steps, cfg.jitter_pct)
    parts = sharding.lpt_partition(end - start, world)
    steps, ls, le = synth.make_graph(cfg, path_subset=parts[rank])
    d_steps = torch.from_numpy(steps.view(np.int32)).to(dev)
    ref = sharding.ShardedDepth(ls, le, cfg.n_segs, dev, n_paths_global=cfg.n_paths)
    fused = sharding.FusedShardedDepth(ls, le, cfg.n_segs, dev, [len(p) for p in parts])
    push = sharding.FusedShardedDepth(ls, le, cfg.n_segs, dev, [len(p) for p in parts], form="push")
    st = torch.cuda.current_stream(dev)
    ok = True
    for it in range(3):
        ref.run(d_steps, st); ref.status()
        fused.run(d_steps, st); fused.status()
        push.run(d_steps, st); push.status()
        torch.cuda.synchronize(dev)
        rd, ru = ref.results()
        fd, fu = fused.results()
        pd, pu = push.results()
        ok = ok and bool((rd == fd).all() and (ru == fu).all()) and bool((rd == pd).all() and (ru == pu).all())
    if rank == 0 and cfg.n_steps <= 400_000_000:
        import oracle_lib as O
        full_steps, s, e = synth.make_graph(cfg)
        rc, od, ou = O.depth_with_uniq(full_steps, s, e, cfg.n_segs)
        ok = ok and rc == 0 and bool((od == fd).all() and (ou == fu).all()) and bool((od == rd).all() and (ou == ru).all())

    def timed(fn, reps=20):
        for _ in range(3):
            fn()
        torch.cuda.synchronize(dev); dist.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st)
        for _ in range(reps):
            fn()
        b.record(st)
        torch.cuda.synchronize(dev); dist.barrier()
        t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    t_ref = timed(lambda: ref.run(d_steps, st))
    t_fused = timed(lambda: fused.run(d_steps, st))
    t_push = timed(lambda: push.run(d_steps, st))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(json.dumps({"config": cfg.name, "n_gpus": world, "parity_fused_vs_nccl_vs_oracle": bool(flag.item()),
                          "engine_nccl_form": ref.plan.engine, "engine_fused_form": fused.plan.engine,
                          "nccl_step_ms": t_ref, "fused_step_ms": t_fused, "push_step_ms": t_push,
                          "multicast": bool(push.mc_ptr),
                          "nccl_steps_per_s": cfg.n_steps / (t_ref * 1e-3), "fused_steps_per_s": cfg.n_steps / (t_fused * 1e-3)}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
