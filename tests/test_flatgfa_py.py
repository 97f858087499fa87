"""the flatgfa-py-compatible front end (pollen_b200/flatgfa_py.py).
The CPU tests follow the reference binding's own test file (flatgfa-py/test/test_flatgfa.py) on
the same fixture (tests/golden/ref_tiny.gfa is flatgfa-py/test/tiny.gfa); the GPU test checks the
depth methods the reference binding lacks against the committed goldens and against the
Python-loop version the reference ships as an example (flatgfa-py/examples/depth.py)."""
import os
from collections import Counter

import numpy as np
import pytest

from pollen_b200 import flatgfa_py as flatgfa

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TEST_GFA = os.path.join(ROOT, "tests", "golden", "ref_tiny.gfa")


@pytest.fixture
def gfa():
    with open(TEST_GFA, "rb") as f:
        return flatgfa.parse_bytes(f.read())


def test_segs(gfa):
    assert len(gfa.segments) == 4
    seg = gfa.segments[0]
    assert seg.name == 1
    assert seg.sequence() == b"CAAATAAG"
    assert len(seg) == 8
    seg = list(gfa.segments)[2]
    assert seg.name == 3
    assert str(seg) == "S\t3\tTTG"


def test_segs_find(gfa):
    seg = gfa.segments.find(3)
    assert seg.id == 2
    assert seg.sequence() == b"TTG"
    assert gfa.segments.find(99) is None


def test_paths(gfa):
    assert len(gfa.paths) == 2
    assert len(list(gfa.paths)) == 2
    path = gfa.paths[0]
    assert path.name == "one"
    assert str(path) == "P\tone\t1+,2+,4-\t*"


def test_paths_find(gfa):
    path = gfa.paths.find("two")
    assert path.id == 1
    assert path.name == "two"
    assert gfa.paths.find("three") is None


def test_path_steps(gfa):
    path = gfa.paths[1]
    assert len(path) == 4
    assert len(list(path)) == 4
    step = path[0]
    assert step.segment.name == 1
    assert step.is_forward
    assert str(step) == "1+"
    assert repr(path[3]) == "<Handle 3->" and str(path[3]) == "4-"


def test_links(gfa):
    assert len(gfa.links) == 4
    assert len(list(gfa.links)) == 4
    link = gfa.links[1]
    assert link.from_.segment.name == 2
    assert link.from_.is_forward
    assert link.to.segment.name == 4
    assert not link.to.is_forward
    assert str(link) == "L\t2\t+\t4\t-\t0M"


def test_gfa_str(gfa):
    with open(TEST_GFA, "r", encoding="utf-8") as f:
        assert str(gfa) == f.read()


def test_read_write_gfa(gfa, tmp_path):
    gfa_path = str(tmp_path / "tiny.gfa")
    gfa.write_gfa(gfa_path)
    with open(TEST_GFA, "rb") as orig_f, open(gfa_path, "rb") as written_f:
        assert orig_f.read() == written_f.read()
    new_gfa = flatgfa.parse(gfa_path)
    assert len(new_gfa.segments) == len(gfa.segments)


def test_read_write_flatgfa(gfa, tmp_path):
    flatgfa_path = str(tmp_path / "tiny.flatgfa")
    gfa.write_flatgfa(flatgfa_path)
    new_gfa = flatgfa.load(flatgfa_path)
    assert len(new_gfa.segments) == len(gfa.segments)
    assert str(new_gfa) == str(gfa)
    assert new_gfa.size == gfa.size == os.path.getsize(flatgfa_path)


def test_eq(gfa):
    assert gfa.segments[0] == gfa.segments[0]
    assert gfa.segments[0] != gfa.segments[1]
    assert gfa.paths[0] == gfa.paths[0]
    assert gfa.paths[0] != gfa.paths[1]
    assert gfa.links[0] == gfa.links[0]
    assert gfa.links[0] != gfa.links[1]
    assert gfa.links[1].from_ == gfa.links[2].from_
    assert gfa.links[1].from_ != gfa.links[1].to


def test_hash(gfa):
    d = {gfa.segments[0]: "foo", gfa.paths[0]: "bar", gfa.links[0]: "baz", gfa.links[1].from_: "qux"}
    assert d[gfa.segments[0]] == "foo"
    assert d[gfa.paths[0]] == "bar"
    assert d[gfa.links[0]] == "baz"
    assert d[gfa.links[1].from_] == "qux"


def test_slice(gfa):
    assert len(gfa.segments[1:3]) == 2
    assert len(gfa.segments[2:]) == len(gfa.segments) - 2
    assert gfa.segments[1:3][0].name == gfa.segments[1].name
    assert len(gfa.paths[1:]) == 1
    assert len(gfa.links[2:100]) == 2
    assert len(list(gfa.paths[:1])) == 1
    path = gfa.paths[0]
    assert len(path[2:]) == len(path) - 2
    assert path[2:][0] == path[2]
    assert len(list(path[2:])) == len(path) - 2


def test_graph_without_recorded_lines_prints_normalized(tmp_path):
    """print.rs:129-153: a graph whose line_order is empty (what the reference's extract / chop
    produce) is printed header, segments, paths, links -- str() and write_gfa() must not come out empty."""
    from pollen_b200 import flatgfa_io

    steps = np.array([0, 2, 5, 2], np.uint32)
    for record in (True, False):
        p = tmp_path / f"g{int(record)}.flatgfa"
        flatgfa_io.write_flatgfa(str(p), steps, [0, 2], [2, 4], 3, record_lines=record)
    a, b = flatgfa.load(str(tmp_path / "g1.flatgfa")), flatgfa.load(str(tmp_path / "g0.flatgfa"))
    want = "H\tVN:Z:1.0\nS\t1\tA\nS\t2\tC\nS\t3\tG\nP\tp0\t1+,2+\t*\nP\tp1\t3-,2+\t*\n"
    assert str(a) == want and str(b) == want
    b.write_gfa(str(tmp_path / "out.gfa"))
    assert (tmp_path / "out.gfa").read_text() == want


def test_round_trips_every_golden_graph(golden):
    """str(graph) reproduces the GFA text for every fixture whose lines our writer covers
    (H/S/P/L, original order: print.rs:100-127)."""
    for c in golden:
        src = os.path.join(c["dir"], c["gfa"])
        text = open(src, "rb").read()
        if not text.endswith(b"\n"):
            continue
        g = flatgfa.parse(src)
        assert str(g).encode() == text, c["name"]


@pytest.mark.gpu
def test_gpu_depth_methods_match_goldens_and_python_loop(golden):
    for c in golden:
        src = os.path.join(c["dir"], c["gfa"])
        graph = flatgfa.parse(src)
        want = open(os.path.join(c["dir"], c["depth"]), "rb").read()
        assert graph.depth_table() == want, c["name"]
        # flatgfa-py/examples/depth.py, verbatim in spirit: count every step in Python
        depths = Counter()
        for path in graph.paths:
            for step in path:
                depths[step.segment.id] += 1
        d, u = graph.depth()
        assert [int(x) for x in d] == [depths[seg.id] for seg in graph.segments]
        uniq = Counter()
        for path in graph.paths:
            for sid in {step.seg_id for step in path}:
                uniq[sid] += 1
        assert [int(x) for x in u] == [uniq[seg.id] for seg in graph.segments]
        if len(graph.paths):
            lengths, means = graph.path_depth([graph.paths[0]])
            assert int(lengths[0]) == sum(len(step.segment) for step in graph.paths[0])
