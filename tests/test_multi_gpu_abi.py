"""fgfa_depth_multi_* (include/fgfa_depth.h): the multi-GPU form of the node-depth path behind
the C ABI -- one process, N devices, whole paths per device (flatgfa/src/ops/depth.rs:25-35 is why
a path is never split), partial [depth | uniq] combined by NCCL or by kernel X over peer memory."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import pollen_b200 as pb
from pollen_b200 import binding, sharding, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _graph(seed=3, n_segs=20_000, n_paths=37, max_len=4000):
    rng = np.random.default_rng(seed)
    lens = rng.integers(0, max_len, n_paths)
    lens[4] = 0
    e = np.cumsum(lens).astype(np.uint32)
    s = (e - lens).astype(np.uint32)
    steps = np.empty(int(e[-1]), np.uint32)
    for p in range(n_paths):
        walk = (rng.integers(0, n_segs) + np.cumsum(rng.integers(-2, 4, lens[p]))) % n_segs
        steps[s[p]:e[p]] = (walk.astype(np.uint32) << 1) | rng.integers(0, 2, lens[p]).astype(np.uint32)
    return steps, s, e, n_segs


def test_c_partition_equals_the_python_partition():
    """Both front ends must shard a graph identically (LPT, ties by lower index)."""
    rng = np.random.default_rng(1)
    for n in (1, 2, 3, 4, 8):
        lens = rng.integers(0, 1000, 57)
        lens[5] = lens[9] = lens[30]                       # ties
        e = np.cumsum(lens).astype(np.uint32)
        s = (e - lens).astype(np.uint32)
        own = binding.lpt_partition_c(s, e, n)
        parts = sharding.lpt_partition(e - s, n)
        for k in range(n):
            assert sorted(np.nonzero(own == k)[0].tolist()) == parts[k]


def test_multi_without_a_device_fails_loudly():
    if pb.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pb.DepthError) as e:
        pb.MultiDepth([0], [0], [2], 2, 2)
    assert e.value.code == binding.FGFA_ERR_NO_DEVICE


@pytest.mark.gpu
@pytest.mark.parametrize("n_shards", [1, 2, 3, 5])
def test_peer_exchange_with_all_shards_on_one_device(n_shards):
    """kernel X + the event ordering + partition + packing, with every shard on device 0 (a box
    with one GPU can run it): identical to the oracle, twice (the bitmaps must be clean again)."""
    steps, s, e, n_segs = _graph()
    rc, od, ou = O.depth_with_uniq(steps, s, e, n_segs)
    m = pb.MultiDepth([0] * n_shards, s, e, n_segs, steps.size, exchange="peer")
    owner, dev_steps = m.partition()
    assert int(dev_steps.sum()) == steps.size and owner.max() < n_shards
    for _ in range(2):
        d, u = m.run_host(steps)
        assert (d == od).all() and (u == ou).all()
    m.close()


@pytest.mark.gpu
def test_nccl_form_on_every_visible_device():
    """One ncclAllReduce per device (north_star form).  With one GPU this is the degenerate
    single-shard case; the 2/4/8-GPU runs are in profiles/."""
    n = min(pb.device_count(), 8)
    steps, s, e, n_segs = _graph(seed=7, n_paths=300)      # > 255 paths: uniq travels as u32
    rc, od, ou = O.depth_with_uniq(steps, s, e, n_segs)
    for paths in (300, 37):
        m = pb.MultiDepth(list(range(n)), s[:paths], e[:paths], n_segs, steps.size, exchange="nccl")
        rc, od, ou = O.depth_with_uniq(steps, s[:paths], e[:paths], n_segs)
        for _ in range(2):
            d, u = m.run_host(steps)
            assert (d == od).all() and (u == ou).all()
        m.upload(steps)
        m.run(with_uniq=False)                             # seg_depth (depth.rs:45-56)
        d, _ = m.download(with_uniq=False)
        assert (d == od).all()
        m.close()


@pytest.mark.gpu
def test_multi_reports_out_of_range_segments_and_bad_arguments():
    steps, s, e, n_segs = _graph()
    bad = steps.copy()
    bad[100] = n_segs << 1
    m = pb.MultiDepth([0, 0], s, e, n_segs, steps.size, exchange="peer")
    with pytest.raises(pb.DepthError) as ei:
        m.run_host(bad)
    assert ei.value.code == binding.FGFA_ERR_SEG_OOB
    d, u = m.run_host(steps)                               # the handle stays usable
    rc, od, ou = O.depth_with_uniq(steps, s, e, n_segs)
    assert (d == od).all() and (u == ou).all()
    m.close()
    with pytest.raises(pb.DepthError):
        pb.MultiDepth([0, 0], s, e, n_segs, steps.size, exchange="nccl")      # duplicate device
    with pytest.raises(pb.DepthError):
        pb.MultiDepth([99], s, e, n_segs, steps.size)
    big_e = e.copy()
    big_e[-1] += 5
    with pytest.raises(pb.DepthError) as ei:
        pb.MultiDepth([0], s, big_e, n_segs, steps.size)
    assert ei.value.code == binding.FGFA_ERR_SPAN_OOB


@pytest.mark.gpu
def test_cli_gpus_flag_prints_the_same_table(golden, fgfa_bin):
    """`fgfa --gpus N ... depth -d`: byte-identical to the single-GPU table (N is clamped to the box)."""
    for c in golden[:6]:
        src = os.path.join(c["dir"], c["gfa"])
        with open(os.path.join(c["dir"], c["depth"]), "rb") as f:
            want = f.read()
        for n in ("1", "2", "8"):
            got = subprocess.run([fgfa_bin, "--gpus", n, "-I", src, "depth", "-d"], capture_output=True, check=True).stdout
            assert got == want, (c["name"], n)
    assert subprocess.run([fgfa_bin, "--gpus", "0", "depth", "-d"], capture_output=True, stdin=subprocess.DEVNULL).returncode == 1
    assert subprocess.run([fgfa_bin, "-p", "x1", "depth", "-d"], capture_output=True, stdin=subprocess.DEVNULL).returncode == 1
