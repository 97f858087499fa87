"""interval / window depth along a path (flatgfa/src/ops/window_depth.rs,
flatgfa/src/flatbed.rs; CLI `fgfa depth -b BED`, `fgfa window-depth PATH SIZE`).

The C oracle for this row is unpinned (no runnable reference golden exists offline), so the CPU
tests fix it three ways: an independent pure-Python restatement of the Rust loop (Python floats
are the same IEEE doubles), hand-computed cases, and the reference's documented table
(flatgfa-sh/README.md:282-294) reproduced on a hand-made graph of the same shape.  The GPU tests
require bit-identical f64 values and byte-identical tables."""
import os
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import pollen_b200 as pb
from pollen_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


# ---- independent restatement (pure Python, straight from window_depth.rs:84-153) ---------------
def py_interval_depth(steps, start, end, seg_len, path, windows):
    n_segs = len(seg_len)
    depth = [0] * n_segs
    for p in range(len(start)):                                   # depth.rs:45-56
        for h in steps[start[p]:end[p]]:
            depth[int(h) >> 1] += 1
    out = [0.0] * len(windows)
    cur, pos = 0, 0
    for h in steps[start[path]:end[path]]:
        seg = int(h) >> 1
        ln = int(seg_len[seg])
        r0, r1 = pos, pos + ln
        pos = r1
        total = float((depth[seg] * ln) & 0xFFFFFFFFFFFFFFFF)     # `as f64`
        while cur < len(windows):
            w0, w1 = windows[cur]
            a, b = max(w0, r0), min(w1, r1)
            if b > a:
                amt = float(b - a) / float(r1 - r0)
                out[cur] += (total * amt) / float(w1 - w0)
            if w1 > r1:
                break
            cur += 1
    return np.array(out, dtype=np.float64)


def _random_case(rng, n_segs, n_paths, max_steps, max_len):
    lens = rng.integers(0, max_steps + 1, n_paths)
    start = np.concatenate([[0], np.cumsum(lens)[:-1]]).astype(np.uint32)
    end = (start + lens).astype(np.uint32)
    steps = ((rng.integers(0, n_segs, int(lens.sum())).astype(np.uint32) << 1) | rng.integers(0, 2, int(lens.sum())).astype(np.uint32)).astype(np.uint32)
    seg_len = rng.integers(0, max_len + 1, n_segs).astype(np.uint32)
    return steps, start, end, seg_len


def _window_sets(rng, total):
    """Interval lists along a path of `total` base pairs: sorted tilings, ragged, unsorted,
    overlapping, empty, past-the-end."""
    sets = []
    for size in (1, 3, 7, 64, max(1, total // 3), total + 5):
        ws, we = O.windows(0, total, size)
        sets.append(list(zip(ws.tolist(), we.tolist())))
    cuts = np.unique(rng.integers(0, total + 1, 12))
    sets.append([(int(a), int(b)) for a, b in zip(cuts[:-1], cuts[1:])])                 # sorted, ragged
    sets.append([(int(a), int(a)) for a in cuts])                                        # empty intervals
    sets.append([(int(rng.integers(0, total + 1)), int(rng.integers(0, total + 20))) for _ in range(15)])   # anything, incl. start > end
    sets.append([(0, total), (0, total), (1, max(2, total - 1))])                        # overlapping
    sets.append([(total + 3, total + 9), (0, 4)])                                        # first one never finishes
    sets.append([])
    return sets


def test_oracle_interval_depth_vs_python_restatement():
    rng = np.random.default_rng(7)
    for trial in range(40):
        steps, start, end, seg_len = _random_case(rng, int(rng.integers(1, 30)), int(rng.integers(1, 5)), 60, int(rng.integers(1, 9)))
        for path in range(len(start)):
            total = O.path_length(steps, start, end, seg_len, path)
            assert total == int(seg_len[steps[start[path]:end[path]] >> 1].astype(np.uint64).sum())
            for wins in _window_sets(rng, total):
                ws = [w[0] for w in wins]
                we = [w[1] for w in wins]
                rc, got = O.interval_depth(steps, start, end, seg_len, path, ws, we)
                assert rc == 0
                want = py_interval_depth(steps, start, end, seg_len, path, wins)
                assert got.tobytes() == want.tobytes(), (trial, path, wins)


def test_oracle_interval_depth_hand_computed():
    # one path over segments of length 4,2,6 with node depths 1,3,2 (a second path adds the extra visits)
    seg_len = [4, 2, 6]
    steps = np.array([0, 2, 4, 2, 2, 4], dtype=np.uint32)          # path0 = s0 s1 s2 ; path1 = s1 s1 s2
    start, end = [0, 3], [3, 6]
    rc, d = O.depth_only(steps, start, end, 3)
    assert d.tolist() == [1, 3, 2]
    rc, got = O.interval_depth(steps, start, end, seg_len, 0, [0, 4, 6, 3], [4, 6, 12, 9])
    assert rc == 0
    # [0,4) lies in s0: 1.  [4,6) = s1: 3.  [6,12) = s2: 2.  [3,9) comes after the cursor passed s0 and s1: only s2's
    # overlap [6,9) is seen -> (2*6 * 3/6) / 6 = 1.
    assert got.tolist() == [1.0, 3.0, 2.0, 1.0]
    rc, got = O.interval_depth(steps, start, end, seg_len, 0, [0, 3], [12, 9])
    # [0,12): (1*4*1)/12 + (3*2*1)/12 + (2*6*1)/12, left to right; it is finished by s2, and the step that finishes an
    # interval is offered to the next one too (window_depth.rs:140-146): [3,9) sees s2's overlap [6,9) -> 1
    assert got.tolist() == [(4.0 / 12.0 + 6.0 / 12.0) + 12.0 / 12.0, 1.0]
    assert O.format_float(got[0], 4) == "1.8333"


def test_windows_match_reference_semantics():
    ws, we = O.windows(0, 13, 4)                                    # flatgfa-sh/README.md:284-288
    assert list(zip(ws.tolist(), we.tolist())) == [(0, 4), (4, 8), (8, 12), (12, 13)]
    assert O.windows(0, 0, 5)[0].size == 0
    assert O.windows(0, 12, 4)[1].tolist() == [4, 8, 12]
    assert O.windows(5, 7, 100)[0].tolist() == [5] and O.windows(5, 7, 100)[1].tolist() == [7]


def test_oracle_bed_parser_quirks():
    assert O.parse_bed(b"x\t0\t4\nx\t4\t8\n") == [(b"x", 0, 4), (b"x", 4, 8)]
    assert O.parse_bed(b"#path\tstart\tend\nx\t0\t4\textra\tcols\n") == [(b"x", 0, 4)]      # comments; trailing columns ignored
    assert O.parse_bed(b"x\t0\t4\nx\t4\t8") == [(b"x", 0, 4)]                                # unterminated last line dropped
    assert O.parse_bed(b"a b\t10 20\n") == [(b"a b", 10, 20)]                                # any one byte separates start and end
    assert O.parse_bed(b"x\t007\t08z\n") == [(b"x", 7, 8)]
    assert O.parse_bed(b"") == [] and O.parse_bed(b"# only\n") == []
    for bad in (b"x\n", b"\n", b"x\t\t4\n", b"x\t4\n", b"x\t4\t\n", b"x\t-1\t4\n", b"x\t4\t+8\n", b"x 0 4\n"):
        assert O.parse_bed(bad) is None, bad


def test_product_bed_parser_and_windows_match_oracle():
    """Host logic of the product (no GPU needed): BEDParser and Windows::as_bed through the C ABI."""
    cases = [b"x\t0\t4\nx\t4\t8\n", b"#path\tstart\tend\nx\t0\t4\textra\tcols\n", b"x\t0\t4\nx\t4\t8", b"a b\t10 20\n",
             b"x\t007\t08z\n", b"", b"# only\n", b"\t1\t2\n", b"chr1\t18446744073709551615\t18446744073709551616\n",
             b"x\n", b"\n", b"x\t\t4\n", b"x\t4\n", b"x\t4\t\n", b"x\t-1\t4\n", b"x\t4\t+8\n", b"x 0 4\n", b"x\t1\t2\n\nx\t3\t4\n"]
    for text in cases:
        want = O.parse_bed(text)
        if want is None:
            with pytest.raises(pb.DepthError):
                pb.FlatBED.parse(text)
        else:
            assert pb.FlatBED.parse(text).entries() == want, text
    for start, end, size in ((0, 13, 4), (0, 0, 5), (0, 12, 4), (5, 7, 100), (3, 1000, 1)):
        ws, we = O.windows(start, end, size)
        assert pb.FlatBED.windows(b"p", start, end, size).entries() == [(b"p", int(a), int(b)) for a, b in zip(ws, we)]
    with pytest.raises(pb.DepthError):
        pb.FlatBED.windows(b"p", 0, 10, 0)


def test_documented_window_table_shape():
    """flatgfa-sh/README.md:282-294: path `5` of note5.gfa is 13 bp long and every window of 4 has depth 2.
    note5.gfa itself is not in the reference tree; a hand-made graph with a 13 bp path whose segments all
    have depth 2 must give the same rows."""
    text = "S\t1\tACGTA\nS\t2\tCCCC\nS\t3\tGGGG\nP\t5\t1+,2+,3+\t*\nP\t6\t3+,2-,1-\t*\n"
    names, steps, start, end, pnames = O.read_gfa(text)
    seg_len = [5, 4, 4]
    ws, we = O.windows(0, O.path_length(steps, start, end, seg_len, 0), 4)
    rc, d = O.interval_depth(steps, start, end, seg_len, 0, ws, we)
    assert rc == 0
    got = O.emit_interval_depth([(b"5", int(a), int(b)) for a, b in zip(ws, we)], d)
    assert got == b"5\t0\t4\t2\n5\t4\t8\t2\n5\t8\t12\t2\n5\t12\t13\t2\n"


# ---- GPU -----------------------------------------------------------------------------------------
def _check_sets(steps, start, end, seg_len, path, sets):
    for wins in sets:
        ws = [w[0] for w in wins]
        we = [w[1] for w in wins]
        rc, want = O.interval_depth(steps, start, end, seg_len, path, ws, we)
        assert rc == 0
        got = pb.interval_depth_steps(steps, start, end, seg_len, path, ws, we)
        assert got.tobytes() == want.tobytes(), (path, wins[:8])


@pytest.mark.gpu
def test_gpu_interval_depth_random_cases_bit_exact():
    rng = np.random.default_rng(11)
    for trial in range(12):
        steps, start, end, seg_len = _random_case(rng, int(rng.integers(1, 40)), int(rng.integers(1, 5)), 200, int(rng.integers(1, 9)))
        for path in range(len(start)):
            total = O.path_length(steps, start, end, seg_len, path)
            _check_sets(steps, start, end, seg_len, path, _window_sets(rng, total))


@pytest.mark.gpu
def test_gpu_window_depth_synthetic_bit_exact():
    """Paths longer than one scan tile, more windows than one scan tile, short and long intervals."""
    cfg = synth.CONFIGS["tinyE"]
    steps, s, e = synth.make_graph(cfg)
    rng = np.random.default_rng(3)
    seg_len = rng.integers(0, 40, cfg.n_segs).astype(np.uint32)
    for path in (0, cfg.n_paths - 1):
        total = O.path_length(steps, s, e, seg_len, path)
        for size in (1, 5, 97, 1000, 50_000, total, total + 1):
            ws, we, got, length = pb.window_depth_steps(steps, s, e, seg_len, path, size)
            ows, owe = O.windows(0, total, size)
            assert length == total and (ws == ows).all() and (we == owe).all()
            rc, want = O.interval_depth(steps, s, e, seg_len, path, ows, owe)
            assert rc == 0
            assert got.tobytes() == want.tobytes(), (path, size)
        _check_sets(steps, s, e, seg_len, path, _window_sets(rng, total)[6:])


@pytest.mark.gpu
def test_gpu_window_depth_config_b_bit_exact_and_conservation():
    """Config B (20 M steps): bit-exact against the oracle, and the size-independent property
    sum_w depth_w * len_w == sum over the path of depth[seg] * len(seg) (up to f64 rounding)."""
    cfg = synth.CONFIGS["B"]
    steps, s, e = synth.make_graph(cfg)
    seg_len = np.random.default_rng(5).integers(1, 200, cfg.n_segs).astype(np.uint32)
    path = 3
    total = O.path_length(steps, s, e, seg_len, path)
    for size in (1000, 1_000_003):
        ws, we, got, length = pb.window_depth_steps(steps, s, e, seg_len, path, size)
        assert length == total
        rc, want = O.interval_depth(steps, s, e, seg_len, path, ws, we)
        assert rc == 0 and got.tobytes() == want.tobytes()
        segs = steps >> 1
        depth = np.bincount(segs, minlength=cfg.n_segs).astype(np.uint64)
        sp = segs[s[path]:e[path]]
        weighted = float(int((depth[sp] * seg_len[sp].astype(np.uint64)).sum()))
        assert abs(float((got * (we - ws).astype(np.float64)).sum()) - weighted) <= 1e-9 * weighted


@pytest.mark.gpu
def test_gpu_window_and_bed_tables_api_and_cli(golden, fgfa_bin, tmp_path):
    cli_budget = 3                                                # every CLI run pays ~1 s of CUDA start-up
    for c in golden:
        src = os.path.join(c["dir"], c["gfa"])
        text = open(src, encoding="utf-8").read()
        names, steps, start, end, pnames = O.read_gfa(text)
        if not pnames or any("\t" in p or " " in p for p in pnames):
            continue
        seg_len = np.array([len(l.split("\t")[2]) for l in text.split("\n") if l.startswith("S\t")], np.uint32)
        with pb.FlatGFA.parse(src) as g:
            for p in sorted({0, len(pnames) - 1}):
                if pnames.index(pnames[p]) != p:
                    continue                                      # find_path returns the first of duplicate names
                total = O.path_length(steps, start, end, seg_len, p)
                for size in (1, 4, 1000):
                    ws, we = O.windows(0, total, size)
                    rc, d = O.interval_depth(steps, start, end, seg_len, p, ws, we)
                    assert rc == 0
                    entries = [(pnames[p].encode(), int(a), int(b)) for a, b in zip(ws, we)]
                    want = O.emit_interval_depth(entries, d)
                    assert g.window_depth(pnames[p], size) == want
                    use_cli = cli_budget > 0 and size == 4 and total > 4
                    if use_cli:
                        cli_budget -= 1
                        cli = subprocess.run([fgfa_bin, "-I", src, "window-depth", pnames[p], str(size)], capture_output=True, check=True)
                        assert cli.stdout == want
                    # the same windows as a BED file, with a header comment and an unterminated last line that is dropped
                    bed = b"#path\tstart\tend\n" + b"".join(b"%s\t%d\t%d\n" % e for e in entries) + b"ignored\t0\t1"
                    assert g.bed_depth(bed) == want
                    assert g.interval_depth(pb.FlatBED.parse(bed)).tobytes() == d.tobytes()
                    if use_cli:
                        f = tmp_path / "w.bed"
                        f.write_bytes(bed)
                        cli = subprocess.run([fgfa_bin, "-I", src, "depth", "-b", str(f)], capture_output=True, check=True)
                        assert cli.stdout == want


@pytest.mark.gpu
def test_gpu_window_depth_errors(fgfa_bin, tmp_path):
    src = os.path.join(GOLD, "ref_ex2.gfa")
    with pb.FlatGFA.parse(src) as g:
        with pytest.raises(pb.DepthError):
            g.window_depth("no-such-path", 4)                      # cmds.rs:489 expect("path not found")
        with pytest.raises(pb.DepthError):
            g.window_depth("path0", 0)
        with pytest.raises(pb.DepthError):
            g.bed_depth(b"")                                       # window_depth.rs:207 entries[0] on an empty list
        with pytest.raises(pb.DepthError):
            g.bed_depth(b"nope\t0\t4\n")                           # :208 expect("path not found in graph")
        with pytest.raises(pb.DepthError):
            g.bed_depth(b"path0\t0\n")                             # flatbed.rs:149
    assert subprocess.run([fgfa_bin, "-I", src, "window-depth", "nope", "4"], capture_output=True).returncode != 0
    steps = np.array([0, 2, 4], np.uint32)
    with pytest.raises(pb.DepthError):
        pb.window_depth_steps(steps, [0], [3], [1, 1, 1], 1, 4)    # path index out of range
    assert pb.interval_depth_steps(steps, [0], [3], [1, 1, 1], 0, [], []).size == 0


def test_bed_parsers_agree_on_arbitrary_bytes():
    """Property test: the product's BEDParser (C++) and the oracle's restatement (C) accept and
    reject the same inputs and produce the same entries, whatever the bytes (flatbed.rs:126-152)."""
    from hypothesis import given, settings
    from hypothesis import strategies as st

    line = st.text(alphabet="\t 0123456789xy#-+.", min_size=0, max_size=14).map(lambda s: s.encode())

    @settings(max_examples=400, deadline=None, derandomize=True)
    @given(st.lists(line, min_size=0, max_size=6), st.booleans())
    def check(lines, terminated):
        text = b"\n".join(lines) + (b"\n" if terminated and lines else b"")
        want = O.parse_bed(text)
        if want is None:
            with pytest.raises(pb.DepthError):
                pb.FlatBED.parse(text)
        else:
            assert pb.FlatBED.parse(text).entries() == want

    check()
