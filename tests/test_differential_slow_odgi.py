"""Differential test against the reference's own Python implementation, when it is there.
In the build container /root/reference is mounted, so `slow_odgi depth` can be imported
unmodified and run on freshly generated graphs (hypothesis); on the GPU box it is absent and
these tests skip (the committed tests/golden/ vectors cover that side)."""
import contextlib
import io
import os
import subprocess
import sys

import numpy as np
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

import oracle_lib as O

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "slow_odgi")), reason="reference checkout not present")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def slow_odgi_depth(gfa_text, paths=None):
    for p in (os.path.join(REF, "mygfa"), os.path.join(REF, "slow_odgi")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import mygfa
    from slow_odgi import depth as so_depth

    graph = mygfa.Graph.parse(io.StringIO(gfa_text))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        so_depth.depth(graph, paths)
    return buf.getvalue().encode()


@st.composite
def gfa_graphs(draw):
    n_segs = draw(st.integers(1, 30))
    names = draw(st.lists(st.integers(1, 10_000), min_size=n_segs, max_size=n_segs, unique=True))
    if draw(st.booleans()):
        names = list(range(1, n_segs + 1))          # the NameMap's sequential fast path
    lines = ["H\tVN:Z:1.0"]
    for nm in names:
        seq = draw(st.text(alphabet="ACGTN", min_size=1, max_size=5))
        lines.append(f"S\t{nm}\t{seq}")
    n_paths = draw(st.integers(0, 5))
    for p in range(n_paths):
        steps = draw(st.lists(st.tuples(st.sampled_from(names), st.sampled_from("+-")), min_size=1, max_size=40))
        lines.append(f"P\tp{p}\t" + ",".join(f"{n}{o}" for n, o in steps) + "\t*")
    for _ in range(draw(st.integers(0, 4))):
        a, b = draw(st.sampled_from(names)), draw(st.sampled_from(names))
        lines.append(f"L\t{a}\t+\t{b}\t-\t{draw(st.integers(0, 200))}M")
    order = draw(st.permutations(range(1, len(lines))))      # any line order after the header
    body = [lines[i] for i in order]
    # a P or L line may precede the S lines it names: the parser defers them (parse.rs:83-91)
    return "\n".join([lines[0]] + body) + "\n"


@settings(max_examples=150, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
@given(gfa_graphs())
def test_oracle_and_product_parser_match_slow_odgi(tmp_path_factory, text):
    want = slow_odgi_depth(text)
    # (1) the oracle over an independent reader
    names, steps, start, end, _ = O.read_gfa(text)
    rc, d, u = O.depth_with_uniq(steps, start, end, len(names))
    assert rc == 0 and O.emit(names, d, u) == want
    # (2) the product's parser + writer, then the oracle over the .flatgfa image
    d_ = tmp_path_factory.mktemp("g")
    src, flat = d_ / "g.gfa", d_ / "g.flatgfa"
    src.write_text(text)
    subprocess.run([os.path.join(ROOT, "bin", "fgfa"), "-I", str(src), "-o", str(flat)], check=True)
    rc, fnames, fd, fu = O.file_depth(flat.read_bytes())
    assert rc == 0 and O.emit(fnames, fd, fu) == want


def test_every_reference_gfa_fixture_round_trips(tmp_path):
    """The reference's own round-trip harness (tests/turnt.toml:162-172, envs flatgfa_mem and
    flatgfa_file) over every GFA file in the reference checkout."""
    import glob

    fgfa = os.path.join(ROOT, "bin", "fgfa")
    files = sorted(glob.glob(os.path.join(REF, "tests", "**", "*.gfa"), recursive=True)) + \
        sorted(glob.glob(os.path.join(REF, "flatgfa-py", "test", "*.gfa")))
    assert len(files) >= 8
    for f in files:
        text = open(f, "rb").read()
        want = text if text.endswith(b"\n") else text[: text.rfind(b"\n") + 1]   # parse_mem drops an unterminated last line
        got = subprocess.run([fgfa, "-I", f], capture_output=True)
        assert got.returncode == 0, (f, got.stderr)
        assert got.stdout == want, f
        flat = str(tmp_path / "t.flatgfa")
        subprocess.run([fgfa, "-I", f, "-o", flat], check=True)
        assert subprocess.run([fgfa, "-i", flat], capture_output=True, check=True).stdout == want, f
        subprocess.run([fgfa, "-m", "-p", "128", "-o", flat, "-I", f], check=True)          # env flatgfa_file_inplace
        assert subprocess.run([fgfa, "-m", "-i", flat], capture_output=True, check=True).stdout == want, f
