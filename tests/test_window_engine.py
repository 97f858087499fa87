"""The window engine (segment-major kernels S1-S3, W, B2 of pollen_b200/csrc/window_kernels.cuh)
against the oracle and against the stream engine, through the C ABI.  Plans pick the window engine
only for pools of >= 64 Mi steps, so these tests force it on small inputs (`set_engine`,
FGFA_ENGINE) to reach the edge cases in seconds; the full-size configs in test_gpu_parity.py run
it by default.  Bit-exact: integer work (reference loop: flatgfa/src/ops/depth.rs:15-56)."""
import numpy as np
import pytest

import oracle_lib as O
import pollen_b200 as pb
from pollen_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    torch.cuda.set_device(0)
    return torch


def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()


def _run_plan(torch, steps, s, e, n_segs, engine, budget=0, uniq=True, feeds=None, uniq_width=4, offset=0):
    """One run of a DepthPlan on `engine`; returns (depth, uniq) as uint32 arrays."""
    steps = np.ascontiguousarray(steps, np.uint32)
    plan = pb.DepthPlan(s, e, n_segs, int(steps.size), bitmap_budget_bytes=budget)
    plan.set_engine(engine)
    assert plan.engine == engine
    buf = torch.zeros(steps.size + offset + 8, dtype=torch.int32, device="cuda")
    buf[offset: offset + steps.size] = _dev(torch, steps) if steps.size else buf[:0]
    d_steps = buf[offset:]
    st = torch.cuda.current_stream().cuda_stream
    depth = torch.full((max(n_segs, 1),), -1, dtype=torch.int32, device="cuda")[:n_segs]
    if uniq_width == 1:
        plan.set_uniq_width(1)
        u = torch.full((max(n_segs, 1),), 255, dtype=torch.uint8, device="cuda")[:n_segs]
    else:
        u = torch.full((max(n_segs, 1),), -1, dtype=torch.int32, device="cuda")[:n_segs]
    out = None
    for _ in range(2):                      # the second run must find clean scratch
        if feeds is None:
            plan.run(d_steps, depth, u if uniq else None, st)
        else:
            plan.begin(depth, st)
            for lo, hi in feeds:
                plan.feed(d_steps, lo, hi, depth, u if uniq else None, st)
            plan.finish(u if uniq else None, st)
        plan.status(st)
        d = depth.cpu().numpy().view(np.uint32).copy()
        g = (u.cpu().numpy().astype(np.uint32) if uniq_width == 1 else u.cpu().numpy().view(np.uint32).copy()) if uniq else None
        if out is not None:
            assert (out[0] == d).all() and (not uniq or (out[1] == g).all())
        out = (d, g)
    plan.close()
    return out


def _check(torch, steps, s, e, n_segs, **kw):
    s, e = np.asarray(s, np.uint32), np.asarray(e, np.uint32)
    rc, od, ou = O.depth_with_uniq(np.ascontiguousarray(steps, np.uint32), s, e, n_segs)
    assert rc == 0
    d, u = _run_plan(torch, steps, s, e, n_segs, "window", **kw)
    assert (d == od).all()
    if u is not None:
        assert (u == ou).all()
    return od, ou


@pytest.mark.parametrize("name", ["tiny", "tinyE", "B"])
def test_window_engine_vs_oracle_and_stream_engine(torch_cuda, name):
    cfg = synth.CONFIGS[name]
    steps, s, e = synth.make_graph(cfg)
    od, ou = _check(torch_cuda, steps, s, e, cfg.n_segs)
    d2, u2 = _run_plan(torch_cuda, steps, s, e, cfg.n_segs, "stream")
    assert (d2 == od).all() and (u2 == ou).all()
    assert int(od.sum()) == cfg.n_steps
    # seg_depth (depth.rs:45-56): the depth-only kernel has its own window geometry
    _check(torch_cuda, steps, s, e, cfg.n_segs, uniq=False)


def test_goldens_with_the_window_engine(torch_cuda, golden):
    for c in golden:
        with open(c["dir"] + "/" + c["gfa"], encoding="utf-8") as f:
            names, steps, start, end, pnames = O.read_gfa(f.read())
        if len(start) == 0 or len(names) == 0:
            continue
        _check(torch_cuda, steps, start, end, len(names))


def test_many_paths_span_several_mask_planes_and_passes(torch_cuda):
    """100 paths = 4 planes of 32 paths; a budget of one plane forces four passes, of two planes two."""
    rng = np.random.default_rng(11)
    n_segs, n_paths = 40_000, 100
    lens = rng.integers(0, 3000, n_paths)
    lens[[3, 50, 99]] = 0                                   # empty paths (depth.rs:25 skips nothing else)
    e = np.cumsum(lens).astype(np.uint32)
    s = (e - lens).astype(np.uint32)
    cur = rng.integers(0, n_segs, n_paths)
    steps = np.empty(int(e[-1]), np.uint32)
    for p in range(n_paths):
        walk = (cur[p] + np.cumsum(rng.integers(-3, 5, lens[p]))) % n_segs
        steps[s[p]:e[p]] = (walk.astype(np.uint32) << 1) | rng.integers(0, 2, lens[p]).astype(np.uint32)
    plane = ((n_segs + 31) // 32 * 32) * 4
    for budget in (0, plane, 2 * plane + 4):
        _check(torch_cuda, steps, s, e, n_segs, budget=budget)
    # piecewise feed across plane and pass boundaries
    _check(torch_cuda, steps, s, e, n_segs, budget=2 * plane, feeds=((0, 1), (1, 1), (1, 40), (40, 64), (64, 65), (65, 100)))
    _check(torch_cuda, steps, s, e, n_segs, uniq_width=1)


def test_ragged_spans_every_alignment_and_boundary_length(torch_cuda):
    """Span starts at every offset modulo 32 (sub-chunks start on 128-byte boundaries) and lengths
    around the 256-step sub-chunk size; unordered and overlapping spans are honoured (pool.rs:341-347)."""
    rng = np.random.default_rng(5)
    n_segs = 3000
    steps = ((rng.integers(0, n_segs, 20_000).astype(np.uint32)) << 1) | rng.integers(0, 2, 20_000).astype(np.uint32)
    s, e = [], []
    pos = 0
    for k, ln in enumerate([0, 1, 31, 32, 33, 255, 256, 257, 511, 512, 513, 700, 1, 2, 3] * 2):
        pos += k % 7
        s.append(pos)
        e.append(pos + ln)
        pos += ln
    _check(torch_cuda, steps, s, e, n_segs)
    _check(torch_cuda, steps, [100, 0, 50, 4000, 100], [900, 300, 50, 4100, 900], n_segs)     # overlapping, unordered, empty, duplicate
    for off in (1, 2, 3):                                    # device pointer not 16-byte aligned
        _check(torch_cuda, steps, s, e, n_segs, offset=off)


def test_far_jumps_take_the_scattered_route(torch_cuda):
    """Sub-chunks whose steps jump across the whole id space (uniform ids) and windows whose halo is
    exceeded by single outliers: both leave shared memory for the L2 route; results do not change."""
    cfg = synth.Config("u", 700_001, 40, 1_000_003, synth.KIND_UNIFORM, 40, "")
    steps, s, e = synth.make_graph(cfg)
    _check(torch_cuda, steps, s, e, cfg.n_segs)
    rng = np.random.default_rng(9)
    n_segs, n = 2_000_000, 300_000
    walk = (np.arange(n) * 3 % n_segs).astype(np.uint32)
    out = rng.random(n) < 0.02
    walk[out] = rng.integers(0, n_segs, int(out.sum()))      # 2 % outliers inside otherwise local sub-chunks
    steps = walk << 1
    _check(torch_cuda, steps, [0, n // 3, n // 3 * 2], [n // 3, n // 3 * 2, n], n_segs)


def test_hot_segments_and_graph_edges(torch_cuda):
    """One segment taking every step (same-address shared-memory adds), the last segment of the
    graph (windows are clipped at n_segs), n_segs not a multiple of 32, a single-segment graph."""
    n = 100_000
    _check(torch_cuda, np.full(n, 7 << 1, np.uint32), [0, n // 2], [n // 2, n], 33)
    _check(torch_cuda, np.full(n, (70_000 - 1) << 1 | 1, np.uint32), [0], [n], 70_000)
    _check(torch_cuda, np.zeros(n, np.uint32), [0, 10], [n, 20], 1)
    seg = np.arange(n, dtype=np.uint32) % 50_001
    _check(torch_cuda, seg << 1, [0], [n], 50_001)


def test_out_of_range_segment_id_is_reported(torch_cuda):
    torch = torch_cuda
    steps = (np.arange(5000, dtype=np.uint32) % 1000) << 1
    steps[1234] = 1000 << 1                                   # depth.rs:29 would panic: index out of bounds
    for uniq in (True, False):
        plan = pb.DepthPlan(np.array([0], np.uint32), np.array([5000], np.uint32), 1000, 5000)
        plan.set_engine("window")
        st = torch.cuda.current_stream().cuda_stream
        depth = torch.zeros(1000, dtype=torch.int32, device="cuda")
        u = torch.zeros(1000, dtype=torch.int32, device="cuda")
        plan.run(_dev(torch, steps), depth, u if uniq else None, st)
        with pytest.raises(pb.DepthError) as ei:
            plan.status(st)
        assert ei.value.code == pb.binding.FGFA_ERR_SEG_OOB
        plan.close()


def test_engine_selection(torch_cuda, monkeypatch):
    """Small pools default to the stream engine; FGFA_ENGINE forces either (also through the
    host-buffer entry point, whose cached plan must not outlive the setting)."""
    torch = torch_cuda
    cfg = synth.CONFIGS["tinyE"]
    steps, s, e = synth.make_graph(cfg)
    rc, od, ou = O.depth_with_uniq(steps, s, e, cfg.n_segs)
    plan = pb.DepthPlan(s, e, cfg.n_segs, cfg.n_steps)
    assert plan.engine == "stream" and plan.launches(True) == 2
    plan.set_engine("window")
    assert plan.launches(True) == 5 and plan.launches(False) == 4
    assert plan.autotune(_dev(torch, steps), torch.cuda.current_stream().cuda_stream) == "stream"   # below the size threshold
    plan.close()
    for eng in ("window", "stream", "window"):
        monkeypatch.setenv("FGFA_ENGINE", eng)
        assert pb.DepthPlan(s, e, cfg.n_segs, cfg.n_steps).engine == eng
        d, u = pb.seg_depth_with_uniq_steps(steps, s, e, cfg.n_segs)
        assert (d == od).all() and (u == ou).all()
        assert (pb.seg_depth_steps(steps, s, e, cfg.n_segs) == od).all()
