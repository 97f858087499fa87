"""path-depth mode (`fgfa depth` without -d; depth.rs:88-131,136-160,192-197).
The oracle for this row is unpinned (no runnable reference golden exists offline, see
oracle/depth_oracle.c), so the CPU tests fix it by hand-computed cases and numpy, and the GPU
tests require bit-exact u64 sums and identical f64 means / formatted tables."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import pollen_b200 as pb
from pollen_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _gfa_arrays(path):
    text = open(path, encoding="utf-8").read()
    names, steps, start, end, pnames = O.read_gfa(text)
    seg_len = np.array([len(l.split("\t")[2]) for l in text.split("\n") if l.startswith("S\t")], np.uint32)
    return names, steps, start, end, pnames, seg_len


def test_format_float_matches_rust_semantics():
    cases = {2.0: "2", 1.9: "1.9", 2.1666666: "2.17", 2.25: "2.25", 0.0: "0", 10.0: "10", 100.5: "100.5",
             0.125: "0.12", 0.375: "0.38", 1234567.891: "1234567.89", 0.004: "0", 0.005: "0.01", 1e-9: "0"}
    for x, want in cases.items():
        assert O.format_float(x) == want, (x, O.format_float(x))
    assert O.format_float(float("nan")) == "NaN"        # empty path: 0/0 (depth.rs:129)
    assert O.format_float(20.0) == "20" and O.format_float(0.10) == "0.1"


def test_oracle_path_depth_hand_computed_ex2():
    """tests/depth/basic/ex2.gfa: all sequences have length 2; node depths 2,3,2,1,2."""
    names, steps, start, end, pnames, seg_len = _gfa_arrays(os.path.join(GOLD, "ref_ex2.gfa"))
    rc, lengths, means = O.path_depth(steps, start, end, seg_len)
    assert rc == 0
    assert lengths.tolist() == [12, 8]
    assert means.tolist() == [26 / 12, 18 / 8]
    assert O.emit_path_depth(pnames, lengths, means) == b"#path\tstart\tend\tmean.depth\npath0\t0\t12\t2.17\npath1\t0\t8\t2.25\n"
    rc, lengths, means = O.path_depth(steps, start, end, seg_len, [1])
    assert lengths.tolist() == [8] and means.tolist() == [2.25]


def test_oracle_path_depth_vs_numpy():
    cfg = synth.CONFIGS["tiny"]
    steps, s, e = synth.make_graph(cfg)
    seg_len = np.random.default_rng(1).integers(0, 50, cfg.n_segs).astype(np.uint32)
    rc, lengths, means = O.path_depth(steps, s, e, seg_len)
    assert rc == 0
    segs = steps >> 1
    depth = np.bincount(segs, minlength=cfg.n_segs).astype(np.uint64)
    for p in range(cfg.n_paths):
        sp = segs[s[p]:e[p]]
        ln = seg_len[sp].astype(np.uint64)
        assert int(lengths[p]) == int(ln.sum())
        assert means[p] == float(int((depth[sp] * ln).sum())) / float(int(ln.sum()))


@pytest.mark.gpu
def test_gpu_path_depth_goldens_api_and_cli(golden, fgfa_bin):
    for c in golden:
        src = os.path.join(c["dir"], c["gfa"])
        names, steps, start, end, pnames, seg_len = _gfa_arrays(src)
        rc, ol, om = O.path_depth(steps, start, end, seg_len)
        assert rc == 0
        want = O.emit_path_depth(pnames, ol, om)
        with pb.FlatGFA.parse(src) as g:
            lengths, means = g.path_depth()
            assert (lengths == ol).all()
            assert means.tobytes() == om.tobytes()            # same f64 bit patterns (NaN included)
            assert g.format_path_depth(lengths, means) == want
        assert subprocess.run([fgfa_bin, "-I", src, "depth"], capture_output=True, check=True).stdout == want
        if pnames:
            sel = [len(pnames) - 1, 0]
            rc, sl, sm = O.path_depth(steps, start, end, seg_len, sel)
            args = [fgfa_bin, "-I", src, "depth"]
            for i in sel:
                args += ["-r", pnames[i]]
            args += ["-r", "no-such-path"]                    # cmds.rs:271-275: filter_map drops it
            got = subprocess.run(args, capture_output=True, check=True).stdout
            assert got == O.emit_path_depth([pnames[i] for i in sel], sl, sm)


@pytest.mark.gpu
def test_gpu_path_depth_synthetic_exact_sums():
    for name in ("tinyE", "B"):
        cfg = synth.CONFIGS[name]
        steps, s, e = synth.make_graph(cfg)
        seg_len = np.random.default_rng(2).integers(0, 5000, cfg.n_segs).astype(np.uint32)
        rc, ol, om = O.path_depth(steps, s, e, seg_len)
        assert rc == 0
        lengths, weighted, means = pb.path_depth_steps(steps, s, e, seg_len)
        assert (lengths == ol).all() and means.tobytes() == om.tobytes()
        segs = steps >> 1
        depth = np.bincount(segs, minlength=cfg.n_segs).astype(np.uint64)
        for p in (0, cfg.n_paths - 1):
            sp = segs[s[p]:e[p]]
            assert int(weighted[p]) == int((depth[sp] * seg_len[sp].astype(np.uint64)).sum())
        ids = [cfg.n_paths - 1, 0, 0]
        l2, w2, m2 = pb.path_depth_steps(steps, s, e, seg_len, ids)
        assert (l2 == ol[ids]).all() and m2.tobytes() == om[ids].tobytes()


def test_c_example_fails_loudly_without_a_gpu():
    exe = os.path.join(ROOT, "build", "depth_example")
    if pb.device_count() > 0:
        pytest.skip("a CUDA device is present")
    r = subprocess.run([exe, os.path.join(GOLD, "ref_ex2.gfa")], capture_output=True)
    assert r.returncode == 1 and b"no CUDA device" in r.stderr and r.stdout == b""


@pytest.mark.gpu
def test_gpu_c_example_prints_the_three_tables():
    """examples/depth.c (plain C over include/flatgfa.h): node table, path table, window table."""
    exe = os.path.join(ROOT, "build", "depth_example")
    src = os.path.join(GOLD, "ref_ex2.gfa")
    with pb.FlatGFA.parse(src) as g:
        d, u = g.seg_depth_with_uniq()
        lengths, means = g.path_depth()
        want = g.format_seg_depth(d, u) + g.format_path_depth(lengths, means) + g.window_depth("path0", 4)
    got = subprocess.run([exe, src, "path0", "4"], capture_output=True, check=True).stdout
    assert got == want
    assert got.startswith(b"#node.id\tdepth\tdepth.uniq\n") and b"#path\tstart\tend\tmean.depth\n" in got
