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

# ---- NVLink byte counters around a loop of kernel X alone (NVML field values, KiB, all links summed) ----
def nvlink_counters():
    try:
        import pynvml
        pynvml.nvmlInit()
        out = []
        for i in range(world):
            h = pynvml.nvmlDeviceGetHandleByIndex(i)
            vals = pynvml.nvmlDeviceGetFieldValues(h, [pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_TX, pynvml.NVML_FI_DEV_NVLINK_THROUGHPUT_DATA_RX])
            out.append([int(v.value.ullVal) if v.nvmlReturn == 0 else None for v in vals])
        return out
    except Exception as exc:   # noqa: BLE001
        return f"unavailable: {type(exc).__name__}: {exc}"[:200]

def counted(fn, reps=200):
    for _ in range(3): fn()
    torch.cuda.synchronize(dev); dist.barrier()
    before = nvlink_counters() if rank == 0 else None
    dist.barrier()
    for _ in range(reps): fn()
    torch.cuda.synchronize(dev); dist.barrier()
    after = nvlink_counters() if rank == 0 else None
    dist.barrier()
    if rank != 0 or isinstance(before, str) or isinstance(after, str):
        return before if isinstance(before, str) else after if isinstance(after, str) else None
    per_gpu = []
    for b, a in zip(before, after):
        per_gpu.append({"tx_bytes_per_launch": None if None in (a[0], b[0]) else (a[0] - b[0]) * 1024 / reps,
                        "rx_bytes_per_launch": None if None in (a[1], b[1]) else (a[1] - b[1]) * 1024 / reps})
    return per_gpu

def x_with_barriers():
    f.hdl.barrier(channel=0); x_only(); f.hdl.barrier(channel=1)
res["nvlink_kernel_x"] = counted(x_with_barriers)
res["nvlink_full_step"] = counted(lambda: f.run(d_steps, st))
n, w = cfg.n_segs, world
res["model_bytes_per_rank"] = {"partial_depth_in": (w - 1) * (n // w) * 4, "bitmap_rows_in": (cfg.n_paths - len(parts[rank])) * (n // w) // 8,
                               "result_in": (w - 1) * (n // w) * 5, "result_out_multicast": (n // w) * 5}
if rank == 0: print(json.dumps({"n_gpus": world, "multicast": bool(f.mc_ptr), **res}))
dist.destroy_process_group()
