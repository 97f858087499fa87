"""Multi-GPU node depth: whole paths are partitioned across ranks by step count and the
per-rank partial ``[depth | uniq]`` arrays are combined with one allreduce-sum.

``depth`` is a sum over steps and ``uniq`` a sum over *paths* of per-path indicator
vectors (reference loop: flatgfa/src/ops/depth.rs:25-35), so both shard exactly as long
as a path never straddles two ranks.  One process per GPU; ``torch.distributed`` is the
plumbing (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def lpt_partition(lengths: Sequence[int], n_parts: int) -> List[List[int]]:
    """Longest-processing-time-first: paths by decreasing step count (``Path::step_count``,
    flatgfa/src/flatgfa.rs:114-118) onto the least-loaded part; ties by lower index so the
    result is deterministic on every rank.  Each part's path list is returned ascending."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    loads = [0] * n_parts
    parts: List[List[int]] = [[] for _ in range(n_parts)]
    for i in order:
        k = min(range(n_parts), key=lambda j: (loads[j], j))
        parts[k].append(i)
        loads[k] += int(lengths[i])
    return [sorted(p) for p in parts]


def pack_shard(steps: np.ndarray, span_start, span_end, paths: Sequence[int]):
    """Concatenate the selected paths' steps into one local pool + local span table."""
    span_start = np.asarray(span_start, dtype=np.int64)
    span_end = np.asarray(span_end, dtype=np.int64)
    lens = np.array([span_end[p] - span_start[p] for p in paths], dtype=np.int64)
    local_end = np.cumsum(lens)
    local_start = local_end - lens
    out = np.empty(int(lens.sum()), dtype=np.uint32)
    for k, p in enumerate(paths):
        out[local_start[k]:local_end[k]] = steps[span_start[p]:span_end[p]]
    return out, local_start.astype(np.uint32), local_end.astype(np.uint32)


def allreduce_counts(buf, group=None) -> None:
    """Sum the concatenated ``[depth | uniq]`` u32 buffer over all ranks, in place.
    The tensor is int32-typed for the collective: two's-complement addition is the same
    bit pattern as u32 addition, and no true sum exceeds n_steps < 2^32 (pool.rs:9-11)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(buf, op=dist.ReduceOp.SUM, group=group)


class ShardedDepth:
    """Per-rank engine: a DepthPlan over this rank's paths + the allreduce.

    ``local_start/local_end`` index the rank's packed local steps pool (see
    ``pack_shard`` / ``synth.make_graph(path_subset=...)``).  ``n_paths_global`` is the
    number of paths of the whole graph: when it is <= 255 the exchange uses u8 uniq
    counters packed four to a word behind the u32 depths, so ONE allreduce moves
    ``4*n_segs + n_segs`` bytes instead of ``8*n_segs`` (no byte can overflow because
    ``uniq <= n_paths_global``)."""

    def __init__(self, local_start, local_end, n_segs: int, device, n_paths_global=None):
        import torch

        from .binding import DepthPlan

        self.torch = torch
        self.device = device
        self.n_segs = int(n_segs)
        self.n_local_steps = int(local_end[-1]) if len(local_end) else 0
        self.plan = DepthPlan(local_start, local_end, n_segs, self.n_local_steps)
        self.compact = n_paths_global is not None and int(n_paths_global) <= 255
        if self.compact:
            self.plan.set_uniq_width(1)
            self.out = torch.zeros(self.n_segs + (self.n_segs + 3) // 4, dtype=torch.int32, device=device)
            self._uniq = self.out[self.n_segs:].view(torch.uint8)[: self.n_segs]
        else:
            # one buffer so that a single collective moves both arrays
            self.out = torch.zeros(2 * self.n_segs, dtype=torch.int32, device=device)
            self._uniq = self.out[self.n_segs:]

    @property
    def depth(self):
        return self.out[: self.n_segs]

    @property
    def uniq(self):
        """uint8 tensor in compact mode, int32 (u32 bit pattern) otherwise."""
        return self._uniq

    @property
    def exchange_bytes(self) -> int:
        return int(self.out.numel()) * 4

    def results(self):
        """(depth, uniq) as uint64 numpy arrays (downloads)."""
        import numpy as np

        d = self.depth.cpu().numpy().view(np.uint32).astype(np.uint64)
        u = self.uniq.cpu().numpy()
        u = (u if self.compact else u.view(np.uint32)).astype(np.uint64)
        return d, u

    def run(self, d_steps, stream=None) -> None:
        """Enqueue: zero + kernels on this rank's shard, then the allreduce."""
        torch = self.torch
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        self.plan.run(d_steps, self.depth, self.uniq, st.cuda_stream)
        allreduce_counts(self.out)

    def status(self) -> None:
        self.plan.status(self.torch.cuda.current_stream(self.device).cuda_stream)


class FusedShardedDepth:
    """Per-rank engine with the exchange fused into the popcount kernel (kernel X,
    ``k_uniq_exchange``): partial depth and the seen-bitmap live in symmetric (peer-mapped)
    memory, every rank reduces its slice of the segment axis straight from its peers over
    NVLink and stores the final slice into every rank's result buffer.  Needs <= 255 paths
    in the whole graph (u8 uniq) and ``torch.distributed._symmetric_memory``.

    ``rows_per_rank``: number of paths each rank holds (same list on every rank)."""

    def __init__(self, local_start, local_end, n_segs: int, device, rows_per_rank, group=None, use_multicast=True,
                 form: str = "pull", local_engine: str = "stream"):
        """``form``: "pull" = kernel X (every rank reads its slice of all peers' partials and bitmaps),
        "push" = kernels P + R (every rank stores its u8 uniq counts and partial depth into the slice owners'
        receive slots, the owners reduce locally and multicast the result).  ``local_engine`` (push only):
        "stream" = kernel A + the popcount inside kernel P, "window" = the window engine with u8 uniq counts
        (kernels S1-S3, W, B2), which kernel P then forwards."""
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        from .binding import DepthPlan

        self.torch, self.dist = torch, dist
        self.device = device
        self.n_segs = int(n_segs)
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        self.rows = [int(r) for r in rows_per_rank]
        assert len(self.rows) == self.world and sum(self.rows) <= 255
        assert len(local_start) == self.rows[self.rank]
        self.n_local_steps = int(local_end[-1]) if len(local_end) else 0
        self.plan = DepthPlan(local_start, local_end, n_segs, self.n_local_steps)
        row_bytes = self.plan.bitmap_row_bytes
        al = lambda x: (x + 255) // 256 * 256
        self.off_partial = 0
        self.off_bitmap = al(self.n_segs * 4)
        self.off_final_depth = self.off_bitmap + al(row_bytes * max(1, max(self.rows)))
        self.off_final_uniq = self.off_final_depth + al(self.n_segs * 4)
        self.off_puniq = self.off_final_uniq + al(self.n_segs)
        self.off_recv = self.off_puniq + al((self.n_segs + 31) // 32 * 32)
        assert form in ("pull", "push") and local_engine in ("stream", "window") and (form == "push" or local_engine == "stream")
        self.form, self.local_engine = form, local_engine
        from .binding import exchange_recv_bytes
        self.recv_bytes = exchange_recv_bytes(self.world, self.n_segs) if form == "push" else 0
        total = self.off_recv + al(self.recv_bytes)
        self.buf = symm_mem.empty(total, dtype=torch.uint8, device=device)
        self.buf.zero_()
        self.hdl = symm_mem.rendezvous(self.buf, group if group is not None else dist.group.WORLD)
        self.ptrs = [int(p) for p in self.hdl.buffer_ptrs]
        # NVLS multicast mapping of the same buffer, when the fabric offers one
        import os

        if os.environ.get("FGFA_MULTICAST", "1") == "0":
            use_multicast = False
        self.mc_ptr = int(getattr(self.hdl, "multicast_ptr", 0) or 0) if use_multicast else 0
        if local_engine == "window":
            self.plan.set_uniq_width(1)
            self.plan.set_engine("window")
        else:
            self.plan.use_bitmap(self.ptrs[self.rank] + self.off_bitmap, row_bytes * max(1, self.rows[self.rank]))
        self.bitmap_view = self.buf[self.off_bitmap: self.off_bitmap + row_bytes * max(1, self.rows[self.rank])]
        self.depth = self.buf[self.off_final_depth: self.off_final_depth + 4 * self.n_segs].view(torch.int32)
        self.uniq = self.buf[self.off_final_uniq: self.off_final_uniq + self.n_segs]
        self.compact = True
        torch.cuda.synchronize(device)
        dist.barrier(group)

    @property
    def exchange_bytes(self) -> int:
        """NVLink bytes this rank moves per step (in + out)."""
        n, w = self.n_segs, self.world
        slice_segs = -(-n // w)
        if self.form == "push":                              # partials out as stores, result slices in
            return (w - 1) * slice_segs * 5 + (w - 1) * slice_segs * 5
        rows_remote = sum(self.rows) - self.rows[self.rank]
        return (w - 1) * slice_segs * 4 + rows_remote * slice_segs // 8 + (w - 1) * slice_segs * 5

    def results(self):
        import numpy as np

        d = self.depth.cpu().numpy().view(np.uint32).astype(np.uint64)
        u = self.uniq.cpu().numpy().astype(np.uint64)
        return d, u

    def run(self, d_steps, stream=None) -> None:
        from .binding import exchange_push, exchange_reduce, exchange_uniq_depth

        torch = self.torch
        st = stream if stream is not None else torch.cuda.current_stream(self.device)
        if self.form == "push":
            with torch.cuda.stream(st):
                me = self.ptrs[self.rank]
                # (every peer finished reducing the previous run before it passed that run's second barrier)
                if self.local_engine == "window":
                    self.plan.run(d_steps, me + self.off_partial, me + self.off_puniq, st.cuda_stream)
                    exchange_push(self.world, self.rank, 0, 0, me + self.off_partial,
                                  [p + self.off_recv for p in self.ptrs], self.n_segs, st.cuda_stream,
                                  partial_uniq=me + self.off_puniq)
                else:
                    self.plan.run_stream_only(d_steps, me + self.off_partial, st.cuda_stream)
                    exchange_push(self.world, self.rank, me + self.off_bitmap, self.rows[self.rank], me + self.off_partial,
                                  [p + self.off_recv for p in self.ptrs], self.n_segs, st.cuda_stream)
                self.hdl.barrier(channel=0)                  # every rank's slots of my slice have landed
                exchange_reduce(self.world, self.rank, me + self.off_recv,
                                [p + self.off_final_depth for p in self.ptrs],
                                [p + self.off_final_uniq for p in self.ptrs], self.n_segs, st.cuda_stream,
                                multicast_base=self.mc_ptr, off_final_depth=self.off_final_depth,
                                off_final_uniq=self.off_final_uniq)
                self.hdl.barrier(channel=1)                  # every rank's result slices have landed
            return
        with torch.cuda.stream(st):
            self.plan.run_stream_only(d_steps, self.ptrs[self.rank] + self.off_partial, st.cuda_stream)
            self.hdl.barrier(channel=0)                      # every rank's partials are complete
            exchange_uniq_depth(
                self.world, self.rank,
                [p + self.off_bitmap for p in self.ptrs], self.rows,
                [p + self.off_partial for p in self.ptrs],
                [p + self.off_final_depth for p in self.ptrs],
                [p + self.off_final_uniq for p in self.ptrs],
                self.n_segs, st.cuda_stream,
                multicast_base=self.mc_ptr, off_partial=self.off_partial,
                off_final_depth=self.off_final_depth, off_final_uniq=self.off_final_uniq)
            self.hdl.barrier(channel=1)                      # every rank's result slices have landed
            self.bitmap_view.zero_()                         # clean seen-bits for the next run

    def status(self) -> None:
        self.plan.status(self.torch.cuda.current_stream(self.device).cuda_stream)
