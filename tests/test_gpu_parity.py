"""Parity tests proper: the CUDA path, called through the C ABI (libflatgfa.so), against
the golden vectors and the oracle.  Bit-exact: this is integer work."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import pollen_b200 as pb
from pollen_b200 import binding, flatgfa_io, sharding, synth

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def torch_cuda():
    import torch

    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    torch.cuda.set_device(0)
    return torch


def _read(case, key):
    with open(os.path.join(case["dir"], case[key]), "rb") as f:
        return f.read()


def _check_vs_oracle(steps, start, end, n_segs):
    rc, od, ou = O.depth_with_uniq(steps, start, end, n_segs)
    assert rc == 0
    gd, gu = pb.seg_depth_with_uniq_steps(steps, start, end, n_segs)
    assert (gd == od).all() and (gu == ou).all()
    assert (pb.seg_depth_steps(steps, start, end, n_segs) == od).all()
    return od, ou


# ---------------------------------------------------------------- golden vectors ----
def test_goldens_through_flatgfa_c_abi(golden):
    """flatgfa_parse -> flatgfa_seg_depth -> flatgfa_format_seg_depth == slow_odgi output."""
    for c in golden:
        with pb.FlatGFA.parse(os.path.join(c["dir"], c["gfa"])) as g:
            d, u = pb.seg_depth_with_uniq(g)
            assert pb.SegDepth(g, d, u).emit() == _read(c, "depth"), c["name"]
            assert (pb.seg_depth(g) == d).all()


def test_goldens_through_cli(golden, fgfa_bin, tmp_path):
    """`fgfa -I x.gfa depth -d` (tests/turnt.toml:179-181), stdin, and `-i x.flatgfa`."""
    for c in golden:
        src = os.path.join(c["dir"], c["gfa"])
        want = _read(c, "depth")
        assert subprocess.run([fgfa_bin, "-I", src, "depth", "-d"], capture_output=True, check=True).stdout == want
        with open(src, "rb") as f:
            assert subprocess.run([fgfa_bin, "depth", "--graph-depth-table"], stdin=f, capture_output=True, check=True).stdout == want
        flat = tmp_path / (c["name"] + ".flatgfa")
        subprocess.run([fgfa_bin, "-I", src, "-o", str(flat)], check=True)
        assert subprocess.run([fgfa_bin, "-i", str(flat), "depth", "-d"], capture_output=True, check=True).stdout == want


def test_golden_path_subsets_via_span_table(golden):
    """odgi/slow_odgi's `-s/--paths` subset == running the op on a subset of spans."""
    for c in golden:
        if "subset_paths" not in c:
            continue
        names, steps, start, end, pnames = O.read_gfa(_read(c, "gfa").decode())
        keep = [i for i, n in enumerate(pnames) if n in c["subset_paths"]]
        d, u = pb.seg_depth_with_uniq_steps(steps, start[keep], end[keep], len(names))
        assert O.emit(names, d, u) == _read(c, "subset_depth"), c["name"]


# ------------------------------------------------------------- oracle, seeded -------
@pytest.mark.parametrize("name", ["tiny", "tinyE", "B"])
def test_synthetic_vs_oracle(name):
    cfg = synth.CONFIGS[name]
    steps, s, e = synth.make_graph(cfg)
    d, u = _check_vs_oracle(steps, s, e, cfg.n_segs)
    assert int(d.sum()) == cfg.n_steps


def test_uniform_random_ids_vs_oracle():
    cfg = synth.Config("u", 70_001, 11, 2_000_003, synth.KIND_UNIFORM, 40, "")
    steps, s, e = synth.make_graph(cfg)
    _check_vs_oracle(steps, s, e, cfg.n_segs)


def test_image_entry_points_and_misaligned_steps_pool(tmp_path):
    """.flatgfa images whose steps pool starts at byte offsets 0..3 mod 4 (odd header
    lengths) and with capacity > len (file.rs:163-167)."""
    cfg = synth.CONFIGS["tiny"]
    steps, s, e = synth.make_graph(cfg)
    rc, od, ou = O.depth_with_uniq(steps, s, e, cfg.n_segs)
    lib = pb.lib()
    for hdr, slack in ((b"VN:Z:1.0", 0), (b"VN:Z:1.0x", 0), (b"VN:Z:1.0xy", 3), (b"VN:Z:1.0xyz", 1), (b"", 2)):
        img = flatgfa_io.build_image(steps, s, e, cfg.n_segs, header=hdr, slack=slack)
        d = np.empty(cfg.n_segs, np.uint64)
        u = np.empty(cfg.n_segs, np.uint64)
        assert lib.fgfa_seg_depth_with_uniq(img.ctypes.data, img.size, d.ctypes.data, u.ctypes.data) == 0
        assert (d == od).all() and (u == ou).all()
        d2 = np.empty(cfg.n_segs, np.uint64)
        assert lib.fgfa_seg_depth(img.ctypes.data, img.size, d2.ctypes.data) == 0 and (d2 == od).all()
        rc2, names, fd, fu = O.file_depth(img.tobytes())
        assert rc2 == 0 and (fd == od).all() and (fu == ou).all()
    p = tmp_path / "t.flatgfa"
    img.tofile(str(p))
    with pb.FlatGFA.load(str(p)) as g:
        d, u = g.seg_depth_with_uniq()
        assert (d == od).all() and (u == ou).all()
        assert pb.SegDepth(g, d, u).emit() == O.emit(names, od, ou)


# ------------------------------------------------------------------- edge cases -----
def test_empty_and_degenerate_graphs():
    z = np.zeros(0, np.uint32)
    d, u = pb.seg_depth_with_uniq_steps(z, [], [], 0)
    assert d.size == 0 and u.size == 0
    d, u = pb.seg_depth_with_uniq_steps(z, [], [], 37)            # segments but no paths
    assert not d.any() and not u.any()
    d, u = pb.seg_depth_with_uniq_steps(z, [0, 0, 0], [0, 0, 0], 5)  # only empty paths
    assert not d.any() and not u.any()
    one = np.array([(4 << 1) | 1], np.uint32)
    d, u = pb.seg_depth_with_uniq_steps(one, [0], [1], 5)
    assert d.tolist() == [0, 0, 0, 0, 1] and u.tolist() == [0, 0, 0, 0, 1]
    loop = np.full(10_000, 2 << 1, np.uint32)                     # one path, one segment, 10k times
    d, u = pb.seg_depth_with_uniq_steps(loop, [0], [10_000], 3)
    assert d.tolist() == [0, 0, 10_000] and u.tolist() == [0, 0, 1]


def test_ragged_spans_every_alignment_and_length():
    """Path starts at every offset mod 4, lengths around the 4096-step chunk size, empty
    paths in between, n_segs not a multiple of 32."""
    rng = np.random.default_rng(11)
    lens = [0, 1, 2, 3, 4, 5, 4095, 4096, 4097, 0, 8191, 8193, 12288, 7, 0, 4093, 1, 16385]
    end = np.cumsum(lens).astype(np.uint32)
    start = (end - np.array(lens)).astype(np.uint32)
    n_segs = 1021
    steps = rng.integers(0, 2 * n_segs, int(end[-1]), dtype=np.uint32)
    _check_vs_oracle(steps, start, end, n_segs)
    # same spans, shifted by 1..3 elements inside a larger pool
    for shift in (1, 2, 3):
        pool = np.concatenate([rng.integers(0, 2 * n_segs, shift, dtype=np.uint32), steps,
                               rng.integers(0, 2 * n_segs, 5, dtype=np.uint32)])
        _check_vs_oracle(pool, start + shift, end + shift, n_segs)


def test_overlapping_and_unordered_spans():
    """The reference honours arbitrary spans (pool.rs:341-347); so must the kernels."""
    rng = np.random.default_rng(5)
    n_segs = 333
    steps = rng.integers(0, 2 * n_segs, 50_000, dtype=np.uint32)
    start = np.array([40_000, 0, 10_000, 10_000, 49_999, 123], np.uint32)
    end = np.array([50_000, 30_000, 20_000, 10_000, 50_000, 4567], np.uint32)
    _check_vs_oracle(steps, start, end, n_segs)


def test_one_giant_path_and_many_tiny_ones():
    """SURVEY §8(d) adversarial shape: one path holds almost every step, hundreds of paths of
    1-40 steps follow (more paths than a u8 counter could hold, several bitmap batches)."""
    rng = np.random.default_rng(17)
    n_segs = 40_003
    lens = np.concatenate([[1_500_000], rng.integers(1, 41, 700)])
    end = np.cumsum(lens).astype(np.uint32)
    start = (end - lens).astype(np.uint32)
    walk = (np.cumsum(rng.integers(0, 3, int(end[-1]))) % n_segs).astype(np.uint32)
    walk[lens[0]:] = rng.integers(0, 20, walk.size - lens[0])      # all the tiny paths share 20 segments
    steps = (walk << 1) | rng.integers(0, 2, walk.size, dtype=np.uint32)
    od, ou = _check_vs_oracle(steps, start, end, n_segs)
    assert int(ou.max()) > 255
    # the same graph through a plan whose bitmap holds only 64 paths at a time
    import torch
    row_bytes = ((n_segs + 31) // 32 + 31) // 32 * 32 * 4
    plan = pb.DepthPlan(start, end, n_segs, int(end[-1]), bitmap_budget_bytes=64 * row_bytes)
    out = torch.empty(2 * n_segs, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    plan.run(_dev(torch, steps), out[:n_segs], out[n_segs:], st)
    plan.status(st)
    g = out.cpu().numpy().view(np.uint32)
    assert plan.launches(True) == 2 * 11 and (g[:n_segs] == od).all() and (g[n_segs:] == ou).all()
    # LPT keeps the giant alone
    parts = sharding.lpt_partition(lens, 4)
    assert [0] in parts


def test_errors_where_the_reference_panics():
    steps = np.array([0, 2, 8, 6], np.uint32)
    with pytest.raises(pb.DepthError) as e:
        pb.seg_depth_with_uniq_steps(steps, [0], [4], 4)          # segment 4 >= n_segs (depth.rs:29)
    assert e.value.code == binding.FGFA_ERR_SEG_OOB
    with pytest.raises(pb.DepthError) as e:
        pb.seg_depth_steps(steps, [0], [4], 4)
    assert e.value.code == binding.FGFA_ERR_SEG_OOB
    with pytest.raises(pb.DepthError) as e:
        pb.seg_depth_with_uniq_steps(steps, [0], [5], 9)          # span past the pool (pool.rs:341-347)
    assert e.value.code == binding.FGFA_ERR_SPAN_OOB
    with pytest.raises(pb.DepthError) as e:
        pb.seg_depth_with_uniq_steps(steps, [3], [2], 9)
    assert e.value.code == binding.FGFA_ERR_SPAN_OOB
    # a bad step buried in a full interior chunk
    big = np.zeros(50_000, np.uint32)
    big[20_000] = 2 * 7
    with pytest.raises(pb.DepthError) as e:
        pb.seg_depth_with_uniq_steps(big, [0], [50_000], 7)
    assert e.value.code == binding.FGFA_ERR_SEG_OOB
    # and the engine is still usable afterwards
    d, u = pb.seg_depth_with_uniq_steps(steps, [0], [4], 5)
    assert d.tolist() == [1, 1, 0, 1, 1]


# --------------------------------------------------------- device-resident plan -----
def _dev(torch, a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()


def test_plan_reuse_batches_and_feed_pipeline(torch_cuda):
    torch = torch_cuda
    cfg = synth.CONFIGS["tinyE"]
    steps, s, e = synth.make_graph(cfg)
    rc, od, ou = O.depth_with_uniq(steps, s, e, cfg.n_segs)
    d_steps = _dev(torch, steps)
    st = torch.cuda.current_stream().cuda_stream
    row_bytes = ((cfg.n_segs + 31) // 32 + 31) // 32 * 32 * 4
    for budget, batches in ((0, 1), (row_bytes, 5), (2 * row_bytes + 8, 3)):
        plan = pb.DepthPlan(s, e, cfg.n_segs, cfg.n_steps, bitmap_budget_bytes=budget)
        depth = torch.full((cfg.n_segs,), -1, dtype=torch.int32, device="cuda")
        uniq = torch.full((cfg.n_segs,), -1, dtype=torch.int32, device="cuda")
        for _ in range(3):                                        # re-runs must see a clean bitmap
            plan.run(d_steps, depth, uniq, st)
            plan.status(st)
            assert (depth.cpu().numpy().view(np.uint32) == od).all()
            assert (uniq.cpu().numpy().view(np.uint32) == ou).all()
        # piecewise feed
        depth.fill_(-1)
        uniq.fill_(-1)
        plan.begin(depth, st)
        for lo, hi in ((0, 1), (1, 1), (1, 4), (4, 5)):
            plan.feed(d_steps, lo, hi, depth, uniq, st)
        plan.finish(uniq, st)
        plan.status(st)
        assert (depth.cpu().numpy().view(np.uint32) == od).all()
        assert (uniq.cpu().numpy().view(np.uint32) == ou).all()
        # depth only
        depth.fill_(-1)
        plan.run(d_steps, depth, None, st)
        plan.status(st)
        assert (depth.cpu().numpy().view(np.uint32) == od).all()
        assert plan.launches(True) == 2 * batches and plan.launches(False) == batches
        plan.close()


@pytest.mark.parametrize("mode", ["direct", "deferred", "window"])
def test_both_seen_bit_strategies(torch_cuda, mode, monkeypatch):
    """Kernel A's ways of recording seen-bits (straight RED.OR per run / runs parked and
    issued by ordinal / shared-memory window flushed per bitmap sector) give identical
    results; the plan picks one by path length, FGFA_SEEN_MODE forces it."""
    torch = torch_cuda
    monkeypatch.setenv("FGFA_SEEN_MODE", mode)
    for name in ("tiny", "tinyE", "B"):
        cfg = synth.CONFIGS[name]
        steps, s, e = synth.make_graph(cfg)
        rc, od, ou = O.depth_with_uniq(steps, s, e, cfg.n_segs)
        plan = pb.DepthPlan(s, e, cfg.n_segs, cfg.n_steps)
        d_steps = _dev(torch, steps)
        out = torch.empty(2 * cfg.n_segs, dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(2):
            plan.run(d_steps, out[: cfg.n_segs], out[cfg.n_segs:], st)
            plan.status(st)
            g = out.cpu().numpy().view(np.uint32)
            assert (g[: cfg.n_segs] == od).all() and (g[cfg.n_segs:] == ou).all(), (mode, name)
        plan.close()
    # a window-hostile graph: consecutive steps alternate between far-apart regions, so
    # window slots collide and the fallback path is exercised
    rng = np.random.default_rng(3)
    n_segs = 3_000_000
    a = rng.integers(0, 1000, 40_000, dtype=np.uint32)
    segs = np.where(np.arange(40_000) % 2 == 0, a, a + 2_097_152 + 131_072 * (np.arange(40_000) % 5)).astype(np.uint32)
    steps = (segs << 1).astype(np.uint32)
    _check_vs_oracle(steps, np.array([0, 20_000], np.uint32), np.array([20_000, 40_000], np.uint32), n_segs)


def test_workspace_lifecycle_and_feed_contract(torch_cuda, monkeypatch):
    """The host-buffer entry points cache their staging between calls; releasing it, or
    disabling the cache with FGFA_WORKSPACE=0, must not change results.  A begin..finish run
    takes d_uniq on every feed or on none."""
    torch = torch_cuda
    lib = pb.lib()
    for name in ("tiny", "tinyE", "tiny"):          # shrink, grow, same plan key again
        cfg = synth.CONFIGS[name]
        steps, s, e = synth.make_graph(cfg)
        _check_vs_oracle(steps, s, e, cfg.n_segs)
        lib.fgfa_release_workspace()
        _check_vs_oracle(steps, s, e, cfg.n_segs)
    monkeypatch.setenv("FGFA_WORKSPACE", "0")
    _check_vs_oracle(steps, s, e, cfg.n_segs)
    monkeypatch.delenv("FGFA_WORKSPACE")
    plan = pb.DepthPlan(s, e, cfg.n_segs, cfg.n_steps)
    d_steps = _dev(torch, steps)
    out = torch.empty(2 * cfg.n_segs, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    plan.begin(out[: cfg.n_segs], st)
    plan.feed(d_steps, 0, 2, out[: cfg.n_segs], out[cfg.n_segs:], st)
    with pytest.raises(pb.DepthError):
        plan.feed(d_steps, 2, cfg.n_paths, out[: cfg.n_segs], None, st)        # uniq dropped mid-run
    with pytest.raises(pb.DepthError):
        plan.feed(d_steps, 3, cfg.n_paths, out[: cfg.n_segs], out[cfg.n_segs:], st)   # gap in the path order
    with pytest.raises(pb.DepthError):
        plan.finish(out[cfg.n_segs:], st)                                       # not all paths fed
    # a fresh run resets the cursor and must not see the seen-bits the abandoned run left
    # behind: run it on DIFFERENT steps (same spans) to prove that
    steps2 = np.roll(steps, 12345) ^ np.uint32(2)
    steps2 = np.minimum(steps2, np.uint32(2 * cfg.n_segs - 1))
    plan.run(_dev(torch, steps2), out[: cfg.n_segs], out[cfg.n_segs:], st)
    plan.status(st)
    rc, od, ou = O.depth_with_uniq(steps2, s, e, cfg.n_segs)
    g = out.cpu().numpy().view(np.uint32)
    assert rc == 0 and (g[: cfg.n_segs] == od).all() and (g[cfg.n_segs:] == ou).all()


def test_misaligned_device_pointer_and_one_shot_device_abi(torch_cuda):
    torch = torch_cuda
    cfg = synth.CONFIGS["tiny"]
    steps, s, e = synth.make_graph(cfg)
    rc, od, ou = O.depth_with_uniq(steps, s, e, cfg.n_segs)
    lib = pb.lib()
    for shift in (0, 1, 2, 3):
        padded = np.concatenate([np.full(shift, 0xFFFFFFFF, np.uint32), steps])
        d_all = _dev(torch, padded)
        d_steps = d_all[shift:]
        assert (d_steps.data_ptr() // 4) % 4 == shift
        depth = torch.empty(cfg.n_segs, dtype=torch.int32, device="cuda")
        uniq = torch.empty(cfg.n_segs, dtype=torch.int32, device="cuda")
        d_s, d_e = _dev(torch, s), _dev(torch, e)
        torch.cuda.synchronize()
        rc = lib.fgfa_depth_device(d_steps.data_ptr(), cfg.n_steps, d_s.data_ptr(), d_e.data_ptr(),
                                   cfg.n_paths, cfg.n_segs, depth.data_ptr(), uniq.data_ptr(), None)
        assert rc == 0
        assert (depth.cpu().numpy().view(np.uint32) == od).all()
        assert (uniq.cpu().numpy().view(np.uint32) == ou).all()


def test_sharded_partials_sum_to_the_whole(torch_cuda):
    """Whole-path shards are exactly additive (depth.rs:25-35): what the allreduce relies on."""
    torch = torch_cuda
    cfg = synth.CONFIGS["B"]
    steps, s, e = synth.make_graph(cfg)
    rc, od, ou = O.depth_with_uniq(steps, s, e, cfg.n_segs)
    parts = sharding.lpt_partition(e - s, 4)
    for n_global in (None, cfg.n_paths):        # u32 exchange buffer / compact [depth u32 | uniq u8]
        tot_d = np.zeros(cfg.n_segs, np.uint64)
        tot_u = np.zeros(cfg.n_segs, np.uint64)
        words = None
        for part in parts:
            ls_steps, ls, le = synth.make_graph(cfg, path_subset=part)
            eng = sharding.ShardedDepth(ls, le, cfg.n_segs, torch.device("cuda:0"), n_paths_global=n_global)
            assert eng.compact == (n_global is not None)
            eng.run(_dev(torch, ls_steps))
            eng.status()
            d, u = eng.results()
            tot_d += d
            tot_u += u
            w = eng.out.cpu().numpy().view(np.uint32).astype(np.uint64)
            words = w if words is None else words + w     # what the allreduce would compute
        assert (tot_d == od).all() and (tot_u == ou).all()
        assert (words[: cfg.n_segs] == od).all()
        if n_global is not None:                            # packed bytes never carry
            packed = words[cfg.n_segs:].astype(np.uint32).view(np.uint8)[: cfg.n_segs]
            assert (packed == ou).all()


def test_fused_exchange_kernel_on_one_device(torch_cuda):
    """Kernel X (popcount fused with the cross-rank reduce-scatter/all-gather) with all
    "ranks" living on one GPU: each rank's slice launch reads every rank's partial depth and
    bitmap rows and writes every rank's result buffers."""
    torch = torch_cuda
    from pollen_b200.binding import exchange_uniq_depth
    for name, world in (("tinyE", 3), ("B", 4), ("tiny", 7)):
        cfg = synth.CONFIGS[name]
        steps, s, e = synth.make_graph(cfg)
        rc, od, ou = O.depth_with_uniq(steps, s, e, cfg.n_segs)
        parts = sharding.lpt_partition(e - s, world)
        st = torch.cuda.current_stream().cuda_stream
        plans, partial, bitmaps, fin_d, fin_u, keep = [], [], [], [], [], []
        for part in parts:
            ls_steps, ls, le = synth.make_graph(cfg, path_subset=part)
            plan = pb.DepthPlan(ls, le, cfg.n_segs, int(le[-1]) if len(le) else 0)
            bm = torch.zeros(plan.bitmap_row_bytes * max(1, len(part)), dtype=torch.uint8, device="cuda")
            plan.use_bitmap(bm.data_ptr(), bm.numel())
            pd = torch.empty(cfg.n_segs, dtype=torch.int32, device="cuda")
            d_steps = _dev(torch, ls_steps) if len(ls_steps) else torch.zeros(4, dtype=torch.int32, device="cuda")
            plan.run_stream_only(d_steps, pd.data_ptr(), st)
            plan.status(st)
            plans.append(plan); partial.append(pd); bitmaps.append(bm); keep.append(d_steps)
            fin_d.append(torch.full((cfg.n_segs,), -1, dtype=torch.int32, device="cuda"))
            fin_u.append(torch.full((cfg.n_segs,), 255, dtype=torch.uint8, device="cuda"))
        for r in range(world):
            exchange_uniq_depth(world, r, [b.data_ptr() for b in bitmaps], [len(p) for p in parts],
                                [p.data_ptr() for p in partial], [t.data_ptr() for t in fin_d],
                                [t.data_ptr() for t in fin_u], cfg.n_segs, st)
        torch.cuda.synchronize()
        for r in range(world):
            assert (fin_d[r].cpu().numpy().view(np.uint32) == od).all(), (name, r)
            assert (fin_u[r].cpu().numpy() == ou).all(), (name, r)
        for p in plans:
            p.close()


def test_push_exchange_kernels_on_one_device(torch_cuda):
    """Kernels P + R (push form of the exchange) with all "ranks" on one GPU: every rank pushes its u8 uniq
    counts and partial depth into the slice owners' receive slots, then every owner reduces its slice and
    writes every rank's result buffers.  Sizes with and without a ragged last word / slice."""
    torch = torch_cuda
    from pollen_b200.binding import exchange_push, exchange_recv_bytes, exchange_reduce
    for name, world in (("tinyE", 3), ("B", 4), ("tiny", 7), ("tiny", 1)):
        cfg = synth.CONFIGS[name]
        steps, s, e = synth.make_graph(cfg)
        rc, od, ou = O.depth_with_uniq(steps, s, e, cfg.n_segs)
        parts = sharding.lpt_partition(e - s, world)
        st = torch.cuda.current_stream().cuda_stream
        nbytes = exchange_recv_bytes(world, cfg.n_segs)
        assert nbytes >= world * 5 * ((cfg.n_segs + world - 1) // world)
        plans, partial, bitmaps, recv, fin_d, fin_u, keep = [], [], [], [], [], [], []
        for part in parts:
            ls_steps, ls, le = synth.make_graph(cfg, path_subset=part)
            plan = pb.DepthPlan(ls, le, cfg.n_segs, int(le[-1]) if len(le) else 0)
            bm = torch.zeros(plan.bitmap_row_bytes * max(1, len(part)), dtype=torch.uint8, device="cuda")
            plan.use_bitmap(bm.data_ptr(), bm.numel())
            pd = torch.empty(cfg.n_segs, dtype=torch.int32, device="cuda")
            d_steps = _dev(torch, ls_steps) if len(ls_steps) else torch.zeros(4, dtype=torch.int32, device="cuda")
            plan.run_stream_only(d_steps, pd.data_ptr(), st)
            plan.status(st)
            plans.append(plan); partial.append(pd); bitmaps.append(bm); keep.append(d_steps)
            recv.append(torch.full((nbytes,), 0x5A, dtype=torch.uint8, device="cuda"))    # every byte read must have been pushed
            fin_d.append(torch.full((cfg.n_segs,), -1, dtype=torch.int32, device="cuda"))
            fin_u.append(torch.full((cfg.n_segs,), 255, dtype=torch.uint8, device="cuda"))
        for _ in range(2):                                  # twice: P leaves the bitmaps clean, so run the stream again
            for r in range(world):
                exchange_push(world, r, bitmaps[r].data_ptr(), len(parts[r]), partial[r].data_ptr(),
                              [t.data_ptr() for t in recv], cfg.n_segs, st)
            for r in range(world):
                exchange_reduce(world, r, recv[r].data_ptr(), [t.data_ptr() for t in fin_d],
                                [t.data_ptr() for t in fin_u], cfg.n_segs, st)
            torch.cuda.synchronize()
            for r in range(world):
                assert (fin_d[r].cpu().numpy().view(np.uint32) == od).all(), (name, r)
                assert (fin_u[r].cpu().numpy() == ou).all(), (name, r)
                assert int(bitmaps[r].max()) == 0, (name, r)
                fin_d[r].fill_(-1); fin_u[r].fill_(255)
            for r in range(world):
                plans[r].run_stream_only(keep[r], partial[r].data_ptr(), st)
        # the same exchange fed with u8 counts that a plan has already popcounted (what the window engine hands over)
        pu8 = []
        for r, part in enumerate(parts):
            ls_steps, ls, le = synth.make_graph(cfg, path_subset=part)
            q = pb.DepthPlan(ls, le, cfg.n_segs, int(le[-1]) if len(le) else 0)
            q.set_uniq_width(1)
            u8 = torch.full(((cfg.n_segs + 31) // 32 * 32,), 0xEE, dtype=torch.uint8, device="cuda")
            q.run(keep[r], partial[r], u8, st)
            q.status(st)
            pu8.append(u8)
            q.close()
        for r in range(world):
            exchange_push(world, r, 0, 0, partial[r].data_ptr(), [t.data_ptr() for t in recv], cfg.n_segs, st,
                          partial_uniq=pu8[r].data_ptr())
        for r in range(world):
            exchange_reduce(world, r, recv[r].data_ptr(), [t.data_ptr() for t in fin_d],
                            [t.data_ptr() for t in fin_u], cfg.n_segs, st)
        torch.cuda.synchronize()
        for r in range(world):
            assert (fin_d[r].cpu().numpy().view(np.uint32) == od).all(), (name, r)
            assert (fin_u[r].cpu().numpy() == ou).all(), (name, r)
        for p in plans:
            p.close()


# ------------------------------------------------------------------- full size ------
def test_full_size_config_C_vs_oracle_and_properties(torch_cuda):
    """BASELINE.json configs[2]: 5M segments, 90 paths, 400M steps (device-resident)."""
    torch = torch_cuda
    cfg = synth.CONFIGS["C"]
    steps, s, e = synth.make_graph(cfg)
    d_steps = _dev(torch, steps)
    plan = pb.DepthPlan(s, e, cfg.n_segs, cfg.n_steps)
    depth = torch.empty(cfg.n_segs, dtype=torch.int32, device="cuda")
    uniq = torch.empty(cfg.n_segs, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    plan.run(d_steps, depth, uniq, st)
    plan.status(st)
    gd = depth.cpu().numpy().view(np.uint32)
    gu = uniq.cpu().numpy().view(np.uint32)
    # size-independent properties
    assert int(gd.sum(dtype=np.uint64)) == cfg.n_steps
    assert (gu <= np.minimum(gd, cfg.n_paths)).all() and ((gu > 0) == (gd > 0)).all()
    # path-additivity: first 30 paths + remaining 60 paths == all 90
    acc_d = np.zeros(cfg.n_segs, np.uint64)
    acc_u = np.zeros(cfg.n_segs, np.uint64)
    for lo, hi in ((0, 30), (30, 90)):
        sub = pb.DepthPlan(s[lo:hi], e[lo:hi], cfg.n_segs, cfg.n_steps)
        sub.run(d_steps, depth, uniq, st)
        sub.status(st)
        acc_d += depth.cpu().numpy().view(np.uint32)
        acc_u += uniq.cpu().numpy().view(np.uint32)
        sub.close()
    assert (acc_d == gd).all() and (acc_u == gu).all()
    # and the oracle itself (about 1.5 s of CPU at this size)
    rc, od, ou = O.depth_with_uniq(steps, s, e, cfg.n_segs)
    assert rc == 0 and (gd == od).all() and (gu == ou).all()


def test_full_size_config_E_skewed_vs_oracle(torch_cuda):
    """BASELINE.json configs[4] shape: 8 long looping paths with hot segments."""
    torch = torch_cuda
    cfg = synth.CONFIGS["E"]
    steps, s, e = synth.make_graph(cfg)
    d_steps = _dev(torch, steps)
    plan = pb.DepthPlan(s, e, cfg.n_segs, cfg.n_steps)
    out = torch.empty(2 * cfg.n_segs, dtype=torch.int32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    plan.run(d_steps, out[: cfg.n_segs], out[cfg.n_segs:], st)
    plan.status(st)
    g = out.cpu().numpy().view(np.uint32)
    rc, od, ou = O.depth_with_uniq(steps, s, e, cfg.n_segs)
    assert rc == 0 and (g[: cfg.n_segs] == od).all() and (g[cfg.n_segs:] == ou).all()
    assert int(od.max()) > 10_000                                  # hot segments are hot
