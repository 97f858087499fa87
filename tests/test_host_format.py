"""Host-side logic that needs no GPU: the .flatgfa reader/writer, the GFA text parser,
the flatgfa-c accessors, the CLI's conversion mode, and the C ABI's export list."""
import ctypes as C
import os
import re
import struct
import subprocess

import numpy as np
import pytest

import oracle_lib as O
import pollen_b200 as pb
from pollen_b200 import binding, flatgfa_io, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def test_every_declared_symbol_is_exported():
    """Every function declared in include/*.h is exported by libflatgfa.so and bound."""
    lib = pb.lib()
    declared = set()
    for h in ("fgfa_depth.h", "flatgfa.h"):
        text = open(os.path.join(ROOT, "include", h), encoding="utf-8").read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        declared |= set(re.findall(r"\b((?:fgfa|flatgfa|flatbed)_[a-z0-9_]+)\s*\(", text))
    assert len(declared) >= 30
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/ but not exported"
    assert declared == set(binding.EXPORTS), declared ^ set(binding.EXPORTS)


def test_no_gpu_means_loud_failure():
    if pb.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(pb.DepthError) as e:
        pb.seg_depth_with_uniq_steps(np.array([0, 2], np.uint32), [0], [2], 2)
    assert e.value.code == binding.FGFA_ERR_NO_DEVICE


def test_ex1_image_matches_hand_derived_layout(tmp_path, fgfa_bin):
    """byte layout of tests/depth/basic/ex1.gfa as a .flatgfa file."""
    out = tmp_path / "ex1.flatgfa"
    subprocess.run([fgfa_bin, "-I", os.path.join(GOLD, "ref_ex1.gfa"), "-o", str(out)], check=True)
    img = out.read_bytes()
    assert len(img) == 329
    assert struct.unpack_from("<Q", img, 0)[0] == 0xB1011054
    sizes = struct.unpack_from("<22Q", img, 8)
    assert sizes == (8, 8, 2, 2, 1, 1, 2, 2, 3, 3, 2, 2, 0, 0, 2, 2, 5, 5, 0, 0, 6, 6)
    assert img[184:192] == b"VN:Z:1.0"
    assert struct.unpack_from("<QIIII", img, 192) == (1, 0, 1, 0, 0)
    assert struct.unpack_from("<QIIII", img, 216) == (2, 1, 2, 0, 0)
    assert struct.unpack_from("<6I", img, 240) == (0, 5, 0, 3, 0, 0)
    assert struct.unpack_from("<4I", img, 264) == (0, 2, 0, 1)
    assert struct.unpack_from("<4I", img, 280) == (2, 2, 1, 2)
    assert img[296:308] == bytes.fromhex("000000000200000002000000")
    assert img[308:310] == b"AC"
    assert img[310:318] == bytes(8)
    assert img[318:323] == b"path1"
    assert img[323:329] == bytes([0, 1, 3, 1, 3, 2])


def test_stdin_and_file_parsers_agree(tmp_path, fgfa_bin, golden):
    for c in golden[:8]:
        src = os.path.join(c["dir"], c["gfa"])
        a, b = tmp_path / "a.flatgfa", tmp_path / "b.flatgfa"
        subprocess.run([fgfa_bin, "-I", src, "-o", str(a)], check=True)
        with open(src, "rb") as f:
            subprocess.run([fgfa_bin, "-o", str(b)], stdin=f, check=True)
        ia, ib = a.read_bytes(), b.read_bytes()
        # parse_stream unwinds links before paths, parse_mem in file order (parse.rs:61-70
        # vs 110-123): the depth-relevant pools (segs, paths' steps spans, steps) are identical.
        ra, rb = O.file_depth(ia), O.file_depth(ib)
        assert ra[0] == 0 and rb[0] == 0
        for x, y in zip(ra[1:], rb[1:]):
            assert (x == y).all()


def test_product_parser_plus_oracle_reproduce_goldens(tmp_path, fgfa_bin, golden):
    """GFA text -> (product parser, product writer) -> .flatgfa -> oracle == slow_odgi."""
    for c in golden:
        out = tmp_path / (c["name"] + ".flatgfa")
        subprocess.run([fgfa_bin, "-I", os.path.join(c["dir"], c["gfa"]), "-o", str(out)], check=True)
        rc, names, d, u = O.file_depth(out.read_bytes())
        assert rc == 0
        with open(os.path.join(c["dir"], c["depth"]), "rb") as f:
            assert O.emit(names, d, u) == f.read(), c["name"]


def test_parse_mem_drops_unterminated_last_line(tmp_path):
    """memfile.rs:54-61: MemchrSplit yields nothing after the last newline."""
    p = tmp_path / "g.gfa"
    p.write_bytes(b"S\t1\tA\nS\t2\tC\nP\tx\t1+,2+\t*\nP\ty\t2+,2-\t*")   # no trailing newline
    with pb.FlatGFA.parse(str(p)) as g:
        assert (g.segment_count, g.path_count) == (2, 1)


def test_parse_stream_keeps_unterminated_last_line(tmp_path, fgfa_bin):
    out = tmp_path / "g.flatgfa"
    subprocess.run([fgfa_bin, "-o", str(out)], input=b"S\t1\tA\nS\t2\tC\nP\tx\t1+,2+\t*\nP\ty\t2+,2-\t*", check=True)
    with pb.FlatGFA.load(str(out)) as g:
        assert (g.segment_count, g.path_count) == (2, 2)
        assert g.step(1, 1) == (1, False)


def test_accessors_match_flatgfa_c_semantics():
    """flatgfa-c/src/lib.rs:79-172 on the reference's tiny.gfa (steps incl. `4-`)."""
    with pb.FlatGFA.parse(os.path.join(GOLD, "ref_tiny.gfa")) as g:
        assert g.segment_count == 4 and g.path_count == 2
        assert g.path_name(0) == b"one" and g.path_name(1) == b"two" and g.path_name(2) is None
        assert g.path_step_count(0) == 3 and g.path_step_count(1) == 4
        assert g.path_step_count(2) == 0xFFFFFFFF                      # lib.rs:133-134
        assert [g.step(0, i) for i in range(4)] == [(0, True), (1, True), (3, False), None]
        assert g.step(7, 0) is None
        assert g.seq(0) == b"CAAATAAG" and g.seq(3) == b"CCAACTCTCTG" and g.seq(4) is None


def test_non_sequential_names_use_the_name_table(tmp_path):
    p = tmp_path / "g.gfa"
    p.write_bytes(b"S\t7\tA\nS\t3\tC\nS\t1\tG\nP\tx\t1+,7-,3+\t*\n")
    with pb.FlatGFA.parse(str(p)) as g:
        assert [g.step(0, i) for i in range(3)] == [(2, True), (0, False), (1, True)]


def test_malformed_inputs_return_null_not_abort(tmp_path):
    bad = {
        "unknown_seg": b"S\t1\tA\nP\tx\t1+,9+\t*\n",
        "no_tab": b"S 1 A\n",
        "bad_kind": b"X\t1\n",
        "empty_line": b"S\t1\tA\n\nS\t2\tC\n",
        "bad_steps": b"S\t1\tA\nP\tx\t1+;1+\t*\n",
        "long_cigar": b"S\t1\tA\nL\t1\t+\t1\t+\t300M\n",
    }
    for name, data in bad.items():
        p = tmp_path / (name + ".gfa")
        p.write_bytes(data)
        with pytest.raises(pb.DepthError):
            pb.FlatGFA.parse(str(p))
    with pytest.raises(pb.DepthError):
        pb.FlatGFA.parse(str(tmp_path / "does_not_exist.gfa"))


def test_steps_parser_quirks_match_gfaline(tmp_path):
    """gfaline.rs:230-263: a trailing segment number without orientation is dropped; one
    stray byte after an orientation is swallowed (the state machine consumes it)."""
    p = tmp_path / "g.gfa"
    p.write_bytes(b"S\t1\tA\nS\t2\tC\nP\tx\t1+,2\t*\nP\ty\t2-x\t*\n")
    with pb.FlatGFA.parse(str(p)) as g:
        assert g.path_step_count(0) == 1 and g.path_step_count(1) == 1
        assert g.step(1, 0) == (1, False)


def test_view_errors(tmp_path):
    img = flatgfa_io.build_image(np.array([0, 2, 2], np.uint32), [0], [3], 2)
    n = C.c_uint64()
    lib = pb.lib()
    assert lib.fgfa_flatgfa_counts(img.ctypes.data, img.size, C.byref(n), None, None) == 0 and n.value == 2
    bad = img.copy()
    bad[0] ^= 0xFF
    assert lib.fgfa_flatgfa_counts(bad.ctypes.data, bad.size, None, None, None) == binding.FGFA_ERR_BAD_MAGIC
    assert lib.fgfa_flatgfa_counts(img.ctypes.data, img.size - 1, None, None, None) == binding.FGFA_ERR_TRUNCATED
    assert lib.fgfa_flatgfa_counts(img.ctypes.data, 100, None, None, None) == binding.FGFA_ERR_TRUNCATED
    for path, code in ((tmp_path / "bad.flatgfa", bad), ):
        code.tofile(str(path))
        with pytest.raises(pb.DepthError):
            pb.FlatGFA.load(str(path))


def test_capacity_slack_is_skipped(tmp_path):
    """file.rs:163-167: each pool occupies capacity (not len) elements."""
    cfg = synth.CONFIGS["tiny"]
    steps, s, e = synth.make_graph(cfg)
    tight = flatgfa_io.build_image(steps, s, e, cfg.n_segs)
    slack = flatgfa_io.build_image(steps, s, e, cfg.n_segs, slack=5)
    assert slack.size > tight.size
    a, b = O.file_depth(tight.tobytes()), O.file_depth(slack.tobytes())
    assert a[0] == 0 and b[0] == 0
    for x, y in zip(a[1:], b[1:]):
        assert (x == y).all()
    p = tmp_path / "slack.flatgfa"
    slack.tofile(str(p))
    with pb.FlatGFA.load(str(p)) as g:
        assert g.segment_count == cfg.n_segs and g.path_count == cfg.n_paths
        assert g.path_step_count(3) == int(e[3] - s[3])
        assert g.step(2, 5) == (int(steps[s[2] + 5] >> 1), not bool(steps[s[2] + 5] & 1))


def test_dump_round_trip(tmp_path):
    src = os.path.join(GOLD, "rand_05.gfa")
    with pb.FlatGFA.parse(src) as g:
        g.dump(str(tmp_path / "a.flatgfa"))
        with pb.FlatGFA.load(str(tmp_path / "a.flatgfa")) as h:
            assert (h.segment_count, h.path_count) == (g.segment_count, g.path_count)
            for p in range(g.path_count):
                assert h.path_name(p) == g.path_name(p)
                n = g.path_step_count(p)
                assert h.path_step_count(p) == n
                assert [h.step(p, i) for i in range(n)] == [g.step(p, i) for i in range(n)]
            h.dump(str(tmp_path / "b.flatgfa"))
    assert (tmp_path / "a.flatgfa").read_bytes() == (tmp_path / "b.flatgfa").read_bytes()


def test_cli_rejects_out_of_scope(fgfa_bin):
    r = subprocess.run([fgfa_bin, "-I", os.path.join(GOLD, "ref_ex1.gfa"), "chop", "-c", "3"], capture_output=True)
    assert r.returncode != 0 and b"scope" in r.stderr


def test_synth_is_deterministic_and_shaped():
    cfg = synth.CONFIGS["tinyE"]
    a = synth.make_graph(cfg)
    b = synth.make_graph(cfg)
    assert all((x == y).all() for x, y in zip(a, b))
    steps, s, e = a
    assert int(e[-1]) == cfg.n_steps and int(s[0]) == 0 and (s[1:] == e[:-1]).all()
    assert int((steps >> 1).max()) < cfg.n_segs
    rc, d, u = O.depth_with_uniq(steps, s, e, cfg.n_segs)
    assert rc == 0 and int(d.sum()) == cfg.n_steps
    assert int(d.max()) > 50            # the hot windows really are hot
    assert (u <= np.minimum(d, cfg.n_paths)).all()


def test_reference_c_example_builds_and_runs_unchanged(tmp_path):
    """flatgfa-c/example/example.c (includes "../include/flatgfa.h") against this library.
    Needs the reference checkout, which only the build container has."""
    src = "/root/reference/flatgfa-c/example/example.c"
    if not os.path.exists(src):
        pytest.skip("reference checkout not present")
    (tmp_path / "include").mkdir()
    (tmp_path / "example").mkdir()
    import shutil
    shutil.copy(os.path.join(ROOT, "include", "flatgfa.h"), tmp_path / "include" / "flatgfa.h")
    shutil.copy(src, tmp_path / "example" / "example.c")
    libdir = os.path.join(ROOT, "pollen_b200", "lib")
    subprocess.run(["gcc", "-w", "example.c", "-L" + libdir, "-lflatgfa", "-Wl,-rpath," + libdir, "-o", "example"],
                   cwd=tmp_path / "example", check=True)
    out = subprocess.run(["./example", os.path.join(GOLD, "ref_tiny.gfa")], cwd=tmp_path / "example",
                         capture_output=True, check=True).stdout
    assert out == b"one:\n  + CAAATAAG\n  + AAATTTTCTGGAGTTCTAT\ntwo:\n  + CAAATAAG\n  + AAATTTTCTGGAGTTCTAT\n"


def _write_gfa(path, steps, start, end, n_segs, bad_path=None):
    with open(path, "wb") as f:
        f.write(b"H\tVN:Z:1.0\n")
        f.write(("\n".join(f"S\t{i}\tA" for i in range(1, n_segs + 1)) + "\n").encode())
        for p in range(len(start)):
            h = steps[start[p]:end[p]]
            toks = np.char.add(((h >> 1) + 1).astype(str), np.where(h & 1, "-", "+"))
            body = ",".join(toks.tolist())
            if p == bad_path:
                body = body[: len(body) // 2] + ";" + body[len(body) // 2:]
            f.write(b"P\tp%d\t" % p + body.encode() + b"\t*\n")


def test_threaded_step_list_tokenisation_is_exact(tmp_path, fgfa_bin):
    """Large P lines are tokenised by worker threads (same pools, same order as the serial
    reference parser, parse.rs:110-123); a malformed line still fails the whole parse."""
    cfg = synth.Config("big", 20_000, 6, 1_200_000, synth.KIND_WALK, 25, "")
    steps, s, e = synth.make_graph(cfg)
    src, out = tmp_path / "big.gfa", tmp_path / "big.flatgfa"
    _write_gfa(src, steps, s, e, cfg.n_segs)
    assert src.stat().st_size > (4 << 20)
    subprocess.run([fgfa_bin, "-I", str(src), "-o", str(out)], check=True)
    img = out.read_bytes()
    want = flatgfa_io.build_image(steps, s, e, cfg.n_segs)
    rc, names, d, u = O.file_depth(img)
    rc2, names2, d2, u2 = O.file_depth(want.tobytes())
    assert rc == 0 and rc2 == 0 and (d == d2).all() and (u == u2).all()
    with pb.FlatGFA.load(str(out)) as g:
        assert g.path_count == cfg.n_paths
        for p in (0, 3, 5):
            assert g.path_name(p) == b"p%d" % p and g.path_step_count(p) == int(e[p] - s[p])
            for i in (0, 1, int(e[p] - s[p]) - 1):
                h = int(steps[s[p] + i])
                assert g.step(p, i) == (h >> 1, not (h & 1))
    _write_gfa(src, steps, s, e, cfg.n_segs, bad_path=4)
    with pytest.raises(pb.DepthError):
        pb.FlatGFA.parse(str(src))


HANDMADE = (
    "H\tVN:Z:1.0\n"
    "S\t1\tCAAATAAG\tLN:i:8\n"
    "S\t7\tA\n"
    "L\t1\t+\t7\t-\t4M\n"
    "S\t3\tTTG\tRC:i:5\txx:Z:y\n"
    "P\tx\t1+,7-,3+\t4M,2M3N1M\n"
    "L\t7\t-\t3\t+\t0M\n"
    "P\ty-rev\t3-,1-\t*\n"
)


def test_gfa_text_round_trips_like_the_reference_harness(golden, fgfa_bin, tmp_path):
    """tests/turnt.toml:162-172 (`flatgfa_mem`: `fgfa < x.gfa`; `flatgfa_file`: `fgfa -o x.flatgfa < x.gfa;
    fgfa -i x.flatgfa`): both must reproduce the input text.  No GPU involved."""
    cases = [(c["name"], open(os.path.join(c["dir"], c["gfa"]), "rb").read()) for c in golden]
    cases.append(("handmade", HANDMADE.encode()))
    for name, text in cases:
        if not text.endswith(b"\n"):
            continue
        got = subprocess.run([fgfa_bin], input=text, capture_output=True, check=True).stdout
        assert got == text, name
        flat = str(tmp_path / "t.flatgfa")
        subprocess.run([fgfa_bin, "-o", flat], input=text, capture_output=True, check=True)
        assert subprocess.run([fgfa_bin, "-i", flat], capture_output=True, check=True).stdout == text, name
    src = tmp_path / "h.gfa"
    src.write_bytes(HANDMADE.encode())
    out = tmp_path / "h.out.gfa"
    subprocess.run([fgfa_bin, "-I", str(src), "-O", str(out)], check=True)
    assert out.read_bytes() == HANDMADE.encode()


def test_gfa_text_two_implementations_agree(golden):
    """The C++ printer (print.cpp) and the Python front end's str() (flatgfa_py.py) are written
    independently from print.rs; they must agree, also for the normalized order of a graph
    without a recorded line order."""
    from pollen_b200 import flatgfa_py
    texts = [open(os.path.join(c["dir"], c["gfa"]), "rb").read() for c in golden] + [HANDMADE.encode()]
    for text in texts:
        g = flatgfa_py.parse_bytes(text)
        assert g._h.format_gfa().decode() == str(g)
    g = flatgfa_py.parse_bytes(HANDMADE.encode())
    assert [str(s) for s in g.segments] == ["S\t1\tCAAATAAG\tLN:i:8", "S\t7\tA", "S\t3\tTTG\tRC:i:5\txx:Z:y"]
    assert str(g.paths[0]) == "P\tx\t1+,7-,3+\t4M,2M3N1M" and str(g.links[0]) == "L\t1\t+\t7\t-\t4M"
    # Reference quirk kept on purpose: the parser reads 'D' as Deletion and 'I' as Insertion
    # (gfaline.rs:176-181) but the printer writes Insertion as "D" and Deletion as "I"
    # (print.rs:13-22), so the two letters swap on a round trip.
    q = flatgfa_py.parse_bytes(b"S\t1\tA\nS\t2\tC\nL\t1\t+\t2\t+\t2M1D3I\n")
    assert str(q.links[0]) == "L\t1\t+\t2\t+\t2M1I3D" and q._h.format_gfa().endswith(b"2M1I3D\n")


def test_long_step_lists_are_cut_into_pieces_exactly(tmp_path, fgfa_bin):
    """One path of several MiB of step text is tokenised in pieces by the worker threads; the
    result (and every failure) must equal the sequential stream parser's (`fgfa < file`), which
    keeps the reference's one-pass StepsParser (gfaline.rs:201-263)."""
    rng = np.random.default_rng(5)
    n_segs = 3000
    segs = rng.integers(1, n_segs + 1, 1_400_000)
    toks = [b"%d%s" % (int(x), b"+" if o else b"-") for x, o in zip(segs.tolist(), rng.integers(0, 2, segs.size).tolist())]
    head = b"H\tVN:Z:1.0\n" + b"".join(b"S\t%d\tAC\n" % i for i in range(1, n_segs + 1))

    def gfa(step_field: bytes, extra=b"") -> bytes:
        return head + b"L\t1\t+\t2\t-\t3M\n" + b"P\tshort\t1+,2-\t*\n" + b"P\tgiant\t" + step_field + b"\t*\n" + extra

    def both(text: bytes):
        src = tmp_path / "g.gfa"
        src.write_bytes(text)
        a = subprocess.run([fgfa_bin, "-I", str(src), "-o", str(tmp_path / "a.flatgfa")], capture_output=True)
        b = subprocess.run([fgfa_bin, "-o", str(tmp_path / "b.flatgfa")], input=text, capture_output=True)
        return a, b

    field = b",".join(toks)
    assert len(field) > 5 * (1 << 20)                   # at least four pieces
    a, b = both(gfa(field, b"P\tafter\t3+\t*\n"))
    assert a.returncode == 0 and b.returncode == 0
    img = (tmp_path / "a.flatgfa").read_bytes()
    assert img == (tmp_path / "b.flatgfa").read_bytes()
    with pb.FlatGFA.load(str(tmp_path / "a.flatgfa")) as g:
        assert g.path_count == 3 and g.path_step_count(1) == segs.size and g.path_step_count(2) == 1
        for i in (0, 1, 700_000, segs.size - 1):
            assert g.step(1, i) == (int(segs[i]) - 1, toks[i].endswith(b"+"))
    # failures and quirks anywhere in the list behave like the one-pass parser
    mid = len(toks) // 2
    variants = {
        "stray byte mid-list": b",".join(toks[:mid]) + b"x" + b",".join(toks[mid:]),
        "double sign mid-list": b",".join(toks[:mid]) + b"-," + b",".join(toks[mid:]),
        "double comma mid-list": b",".join(toks[:mid]) + b",," + b",".join(toks[mid:]),
        "unknown segment late": b",".join(toks[:-5] + [b"999999+"] + toks[-5:]),
        "unknown segment before a stray byte": b",".join(toks[:100] + [b"999999+"] + toks[100:mid]) + b"x" + b",".join(toks[mid:]),
        "trailing number is dropped": field + b",77",
        "trailing stray byte is swallowed": field + b"x",
        "trailing comma": field + b",",
    }
    for name, f in variants.items():
        a, b = both(gfa(f))
        assert a.returncode == b.returncode, name
        if a.returncode == 0:
            assert (tmp_path / "a.flatgfa").read_bytes() == (tmp_path / "b.flatgfa").read_bytes(), name
        else:
            assert a.stderr == b.stderr and a.stderr, name


def test_large_node_depth_table_is_formatted_in_blocks_exactly(tmp_path):
    """Above 2^18 rows SegDepth::emit formats blocks of rows on several threads; the text must
    still be the oracle's byte for byte (depth.rs:61-82), including 20-digit counters and names
    that the `as u32` cast truncates."""
    n = (1 << 18) + 12_345
    rng = np.random.default_rng(3)
    names = rng.integers(1, 1 << 40, n).astype(np.uint64)
    img = flatgfa_io.build_image(np.zeros(1, np.uint32), [0], [1], n, seg_names=names)
    f = tmp_path / "wide.flatgfa"
    img.tofile(str(f))
    d = rng.integers(0, 1 << 63, n).astype(np.uint64) * np.uint64(2) + np.uint64(1)
    u = rng.integers(0, 1000, n).astype(np.uint64)
    with pb.FlatGFA.load(str(f)) as g:
        assert g.format_seg_depth(d, u) == O.emit(names, d, u)


def test_random_gfa_round_trips_through_both_writers():
    """Property test (hypothesis): any well-formed GFA of H/S/L/P lines -- arbitrary line order,
    non-sequential names, optional fields, CIGAR overlaps on links and paths -- survives
    text -> FlatGFA -> text and text -> .flatgfa image -> text unchanged (tests/turnt.toml:162-172)."""
    from hypothesis import HealthCheck, given, settings
    from hypothesis import strategies as st

    from pollen_b200 import flatgfa_py

    # CIGAR lengths above 255 are rejected by the reference itself (flatgfa.rs:231 "length too large")
    cigar = st.lists(st.tuples(st.integers(0, 255), st.sampled_from("MN")), min_size=1, max_size=3).map(
        lambda ops: "".join(f"{n}{c}" for n, c in ops))

    @st.composite
    def graphs(draw):
        n = draw(st.integers(1, 12))
        names = draw(st.lists(st.integers(1, 5000), min_size=n, max_size=n, unique=True))
        lines = []
        for nm in names:
            seq = draw(st.text("ACGTN", min_size=1, max_size=12))
            opt = draw(st.sampled_from(["", "\tLN:i:%d" % len(seq), "\tRC:i:7\txx:Z:a b"]))
            lines.append(("S", f"S\t{nm}\t{seq}{opt}"))
        handle = st.tuples(st.sampled_from(names), st.sampled_from("+-"))
        for _ in range(draw(st.integers(0, 8))):
            (a, ao), (b, bo) = draw(handle), draw(handle)
            lines.append(("L", f"L\t{a}\t{ao}\t{b}\t{bo}\t{draw(cigar)}"))
        for k in range(draw(st.integers(0, 5))):
            steps = draw(st.lists(handle, min_size=1, max_size=15))
            ov = draw(st.one_of(st.just("*"), st.lists(cigar, min_size=1, max_size=3).map(",".join)))
            lines.append(("P", f"P\tp{k}#x\t" + ",".join(f"{s}{o}" for s, o in steps) + f"\t{ov}"))
        # links and paths may come before the segments they name (parse.rs:83-91 defers them)
        order = draw(st.permutations(range(len(lines))))
        body = [lines[i][1] for i in order]
        if draw(st.booleans()):
            body.insert(draw(st.integers(0, len(body))), "H\tVN:Z:1.0")
        return "\n".join(body) + "\n"

    @settings(max_examples=60, deadline=None, derandomize=True, suppress_health_check=[HealthCheck.too_slow])
    @given(graphs())
    def check(text):
        g = flatgfa_py.parse_bytes(text.encode())
        assert g._h.format_gfa().decode() == text          # C++ printer
        assert str(g) == text                              # Python front end
        img = g._h.image()
        assert str(flatgfa_py.FlatGFA(g._h, img)) == text  # through the .flatgfa image

    check()
    with pytest.raises(pb.DepthError):                     # the limit itself
        pb.FlatGFA.parse_bytes(b"S\t1\tA\nL\t1\t+\t1\t+\t256M\n")
    assert pb.FlatGFA.parse_bytes(b"S\t1\tA\nL\t1\t+\t1\t+\t255M\n").format_gfa().endswith(b"255M\n")


def test_preallocated_translation_like_the_reference_harness(golden, fgfa_bin, tmp_path):
    """tests/turnt.toml `flatgfa_file_inplace`: `fgfa -m -p 128 -o x.inplace.flatgfa -I x.gfa ; fgfa -m -i
    x.inplace.flatgfa` must print the input text.  The file's pools carry the capacities of
    `Toc::estimate` (file.rs:136-158, from parse.rs:176-216's scan); with stdin input those of
    `Toc::guess(p)` (file.rs:117-132); a pool that does not fit fails like the reference's fixed store."""
    flat = str(tmp_path / "x.inplace.flatgfa")
    for c in golden:
        src = os.path.join(c["dir"], c["gfa"])
        text = open(src, "rb").read()
        if not text.endswith(b"\n"):
            continue
        subprocess.run([fgfa_bin, "-m", "-p", "128", "-o", flat, "-I", src], check=True)
        assert subprocess.run([fgfa_bin, "-m", "-i", flat], capture_output=True, check=True).stdout == text, c["name"]
        toc = np.fromfile(flat, dtype="<u8", count=23)
        sizes = toc[1:].reshape(11, 2)
        lines = text.split(b"\n")[:-1]
        segs = sum(l.startswith(b"S") for l in lines)
        links = sum(l.startswith(b"L") for l in lines)
        paths = sum(l.startswith(b"P") for l in lines)
        nbytes = lambda k: sum(len(l) for l in lines if l.startswith(k))
        want_caps = [nbytes(b"H"), segs, paths, links, nbytes(b"P") // 3, nbytes(b"S"), (links + paths) * 2,
                     links * 2 + paths * 4, paths * 512, links * 16, segs + links + paths + 8]
        assert sizes[:, 1].tolist() == want_caps, c["name"]
        assert (sizes[:, 0] <= sizes[:, 1]).all()
        assert os.path.getsize(flat) == 184 + int((sizes[:, 1] * np.array([1, 24, 24, 16, 4, 1, 8, 4, 1, 1, 1], np.uint64)).sum())
    # stdin: capacities are guessed from -p
    text = open(os.path.join(GOLD, "ref_tiny.gfa"), "rb").read()
    subprocess.run([fgfa_bin, "-m", "-p", "2", "-o", flat], input=text, check=True)
    toc = np.fromfile(flat, dtype="<u8", count=23)
    assert toc[1:].reshape(11, 2)[:, 1].tolist() == [128, 128, 2, 128, 4096, 2048, 512, 256, 128, 2048, 256]
    assert subprocess.run([fgfa_bin, "-i", flat], capture_output=True, check=True).stdout == text
    # too small a guess: the reference's fixed-size store panics, this build reports it
    r = subprocess.run([fgfa_bin, "-m", "-p", "1", "-o", flat], input=text, capture_output=True)
    assert r.returncode != 0 and b"capacity overflow" in r.stderr      # 2 paths do not fit 1 slot
    r = subprocess.run([fgfa_bin, "-m", "-o", flat, "-I", os.path.join(GOLD, "ref_tiny.gfa"), "depth"], capture_output=True)
    assert r.returncode != 0 or r.stdout                              # with a command -m is the ordinary flow
