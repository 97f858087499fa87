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
