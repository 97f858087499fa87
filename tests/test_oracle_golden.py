"""The oracle (oracle/depth_oracle.c) against the golden vectors recorded from the
reference's own `slow_odgi depth` (tests/golden/, made by oracle/make_golden.py) and
against the tables the reference documents."""
import os

import numpy as np
import pytest

import oracle_lib as O


def _read(case, key):
    with open(os.path.join(case["dir"], case[key]), "rb") as f:
        return f.read()


def test_manifest_complete(golden):
    assert len(golden) >= 28
    assert {c["name"] for c in golden} >= {"ref_ex1", "ref_ex2", "ref_tiny", "ref_readme"}


def test_oracle_matches_slow_odgi_all_paths(golden):
    for c in golden:
        names, steps, start, end, _ = O.read_gfa(_read(c, "gfa").decode())
        rc, d, u = O.depth_with_uniq(steps, start, end, len(names))
        assert rc == 0
        assert O.emit(names, d, u) == _read(c, "depth"), c["name"]


def test_oracle_matches_slow_odgi_path_subsets(golden):
    seen = 0
    for c in golden:
        if "subset_paths" not in c:
            continue
        names, steps, start, end, pnames = O.read_gfa(_read(c, "gfa").decode())
        keep = [i for i, n in enumerate(pnames) if n in c["subset_paths"]]
        rc, d, u = O.depth_with_uniq(steps, start[keep], end[keep], len(names))
        assert rc == 0
        assert O.emit(names, d, u) == _read(c, "subset_depth"), c["name"]
        seen += 1
    assert seen >= 20


def test_documented_tables():
    """Rows the reference documents: slow_odgi/README.md:163-174 (subset x,y) and the
    tables SURVEY.md §4 derived for the in-repo fixtures."""
    here = os.path.join(os.path.dirname(__file__), "golden")
    want = {
        "ref_ex1.depth": b"#node.id\tdepth\tdepth.uniq\n1\t1\t1\n2\t2\t1\n",
        "ref_ex2.depth": b"#node.id\tdepth\tdepth.uniq\n1\t2\t2\n2\t3\t2\n3\t2\t2\n4\t1\t1\n5\t2\t2\n",
        "ref_ex2.subset.depth": b"#node.id\tdepth\tdepth.uniq\n1\t1\t1\n2\t2\t1\n3\t1\t1\n4\t1\t1\n5\t1\t1\n",
        "ref_tiny.depth": b"#node.id\tdepth\tdepth.uniq\n1\t2\t2\n2\t2\t2\n3\t1\t1\n4\t2\t2\n",
        "ref_readme.depth": b"#node.id\tdepth\tdepth.uniq\n1\t2\t2\n2\t0\t0\n3\t4\t3\n4\t2\t2\n",
        "ref_readme.subset.depth": b"#node.id\tdepth\tdepth.uniq\n1\t2\t2\n2\t0\t0\n3\t3\t2\n4\t1\t1\n",
    }
    for name, content in want.items():
        with open(os.path.join(here, name), "rb") as f:
            assert f.read() == content, name


def test_oracle_vs_numpy_bincount():
    """Independent restatement in numpy on a mid-sized random graph."""
    rng = np.random.default_rng(7)
    n_segs, n_paths = 3001, 9
    lens = rng.integers(0, 5000, n_paths)
    end = np.cumsum(lens).astype(np.uint32)
    start = (end - lens).astype(np.uint32)
    segs = rng.integers(0, n_segs, int(end[-1]), dtype=np.uint32)
    steps = (segs << 1) | rng.integers(0, 2, segs.size, dtype=np.uint32)
    rc, d, u = O.depth_with_uniq(steps, start, end, n_segs)
    assert rc == 0
    assert (d == np.bincount(segs, minlength=n_segs)).all()
    want_u = np.zeros(n_segs, dtype=np.uint64)
    for p in range(n_paths):
        want_u[np.unique(segs[start[p]:end[p]])] += 1
    assert (u == want_u).all()
    rc, d2 = O.depth_only(steps, start, end, n_segs)
    assert rc == 0 and (d2 == d).all()


def test_oracle_reports_what_the_reference_panics_on():
    steps = np.array([0, 2, 8], dtype=np.uint32)
    rc, _, _ = O.depth_with_uniq(steps, [0], [3], 4)     # seg 4 >= n_segs (depth.rs:29)
    assert rc == -1
    rc, _, _ = O.depth_with_uniq(steps, [0], [4], 5)     # span past the pool (pool.rs:341-347)
    assert rc == -1
    rc, _, _ = O.depth_with_uniq(steps, [2], [1], 5)     # start > end
    assert rc == -1


def test_oracle_empty_inputs():
    rc, d, u = O.depth_with_uniq(np.zeros(0, np.uint32), [], [], 0)
    assert rc == 0 and d.size == 0
    rc, d, u = O.depth_with_uniq(np.zeros(0, np.uint32), [0, 0], [0, 0], 3)
    assert rc == 0 and not d.any() and not u.any()


def test_path_parallel_variant_equals_the_oracle():
    """The multi-threaded CPU variant bench.py reports beside the baseline gives the oracle's results."""
    from pollen_b200 import synth
    for name in ("tiny", "tinyE"):
        cfg = synth.CONFIGS[name]
        steps, s, e = synth.make_graph(cfg)
        rc, d, u = O.depth_with_uniq(steps, s, e, cfg.n_segs)
        for threads in (1, 3, 8):
            rc2, d2, u2 = O.depth_with_uniq_parallel(steps, s, e, cfg.n_segs, threads)
            assert rc == 0 and rc2 == 0 and (d == d2).all() and (u == u2).all()
    rc, _, _ = O.depth_with_uniq_parallel(np.array([0, 9], np.uint32), [0], [2], 3, 4)
    assert rc == -1                                        # segment id out of range
    rc, d, u = O.depth_with_uniq_parallel(np.zeros(0, np.uint32), [], [], 5, 4)
    assert rc == 0 and not d.any() and not u.any()
