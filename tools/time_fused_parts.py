#!/usr/bin/env python3
"""Times the pieces of the fused exchange: barrier, kernel X, kernel A.  torchrun, N GPUs."""
import json, os, sys
import numpy as np, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pollen_b200 import sharding, synth
from pollen_b200.binding import exchange_uniq_depth

cfg = synth.CONFIGS["C"]
rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr); dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
start, end = synth.make_spans(cfg.n_paths, cfg.n_steps, cfg.jitter_pct)
parts = sharding.lpt_partition(end - start, world)
steps, ls, le = synth.make_graph(cfg, path_subset=parts[rank])
d_steps = torch.from_numpy(steps.view(np.int32)).to(dev)
f = sharding.FusedShardedDepth(ls, le, cfg.n_segs, dev, [len(p) for p in parts])
st = torch.cuda.current_stream(dev)

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize(dev); dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(st)
    for _ in range(reps): fn()
    b.record(st)
    torch.cuda.synchronize(dev); dist.barrier()
    t = torch.tensor([a.elapsed_time(b) / reps], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

def x_only():
    exchange_uniq_depth(f.world, f.rank, [p + f.off_bitmap for p in f.ptrs], f.rows, [p + f.off_partial for p in f.ptrs],
                        [p + f.off_final_depth for p in f.ptrs], [p + f.off_final_uniq for p in f.ptrs], f.n_segs, st.cuda_stream,
                        multicast_base=f.mc_ptr, off_partial=f.off_partial, off_final_depth=f.off_final_depth, off_final_uniq=f.off_final_uniq)
res = {
 "barrier_ms": timed(lambda: f.hdl.barrier(channel=0)),
 "two_barriers_ms": timed(lambda: (f.hdl.barrier(channel=0), f.hdl.barrier(channel=1))),
 "kernel_x_ms": timed(x_only),
 "stream_only_ms": timed(lambda: f.plan.run_stream_only(d_steps, f.ptrs[f.rank] + f.off_partial, st.cuda_stream)),
 "bitmap_zero_ms": timed(lambda: f.bitmap_view.zero_()),
 "full_ms": timed(lambda: f.run(d_steps, st)),
 "nccl_barrier_ms": timed(lambda: dist.barrier()),
}
if rank == 0: print(json.dumps({"n_gpus": world, "multicast": bool(f.mc_ptr), **res}))
dist.destroy_process_group()
