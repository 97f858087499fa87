"""the GPU step-list tokenizer (`fgfa_tokenizer_*`) against the host parser
semantics (gfaline.rs:201-263, namemap.rs:8-33, flatgfa.rs:192-198) on the golden graphs, on a
large synthetic GFA, and on inputs outside the strict grammar (which must be refused, not
mis-parsed)."""
import os

import numpy as np
import pytest

import oracle_lib as O
import pollen_b200 as pb
from pollen_b200 import binding, synth

pytestmark = pytest.mark.gpu


def p_fields(text: bytes):
    """(offset, length) of the steps field of every P line, in file order."""
    out, pos = [], 0
    for line in text.split(b"\n"):
        if line.startswith(b"P\t"):
            f = line.split(b"\t")
            off = pos + 2 + len(f[1]) + 1
            out.append((off, len(f[2])))
        pos += len(line) + 1
    return out


def name_map(text: bytes):
    """NameMap contents after the S lines (namemap.rs:17-25): (sequential_max, others)."""
    seq_max, others, idx = 0, {}, 0
    for line in text.split(b"\n"):
        if line.startswith(b"S\t"):
            name = int(line.split(b"\t")[1])
            if name - 1 == seq_max and name - 1 == idx:
                seq_max += 1
            else:
                others[name] = idx
            idx += 1
    return seq_max, others


def test_goldens_match_the_host_parser(golden):
    for c in golden:
        text = open(os.path.join(c["dir"], c["gfa"]), "rb").read()
        names, steps, start, end, _ = O.read_gfa(text.decode())
        seq_max, others = name_map(text)
        got, gs, ge = pb.tokenize_steps(text, p_fields(text), seq_max, others)
        assert (got == steps).all() and (gs == start).all() and (ge == end).all(), c["name"]


def test_large_text_every_tile_boundary():
    cfg = synth.Config("tok", 3_000_000, 5, 700_000, synth.KIND_UNIFORM, 30, "")   # 1-7 digit names: ragged tokens
    steps, s, e = synth.make_graph(cfg)
    parts = [b"H\tVN:Z:1.0\n"]
    for p in range(cfg.n_paths):
        h = steps[s[p]:e[p]]
        toks = np.char.add(((h >> 1) + 1).astype(str), np.where(h & 1, "-", "+"))
        parts.append(b"P\tp%d\t" % p + ",".join(toks.tolist()).encode() + b"\t*\n")
    text = b"".join(parts)
    fields = p_fields(text)
    assert sum(f[1] for f in fields) > 4_000_000
    got, gs, ge = pb.tokenize_steps(text, fields, cfg.n_segs, None)
    assert (gs == s).all() and (ge == e).all() and (got == steps).all()
    # the same through a name table instead of the sequential fast path
    perm = {int(n) + 1: int(n) for n in np.unique(steps >> 1)}
    got2, _, _ = pb.tokenize_steps(text, fields, 0, perm)
    assert (got2 == steps).all()


def test_empty_fields_and_single_tokens():
    text = b"P\ta\t\t*\nP\tb\t7+\t*\nP\tc\t1-,2+\t*\n"
    got, s, e = pb.tokenize_steps(text, p_fields(text), 10, None)
    assert s.tolist() == [0, 0, 1] and e.tolist() == [0, 1, 3]
    assert got.tolist() == [(6 << 1), (0 << 1) | 1, (1 << 1)]
    got, s, e = pb.tokenize_steps(b"", [], 0, None)
    assert got.size == 0 and s.size == 0


@pytest.mark.parametrize("field", [b"1+,2", b"1+,,2+", b"1+,2+,", b",1+", b"1+2+", b"1*", b"+", b"1+,x-", b"12345678901234567890123+",
                                   b"1+ ,2+", b"99+"])
def test_inputs_outside_the_strict_grammar_are_refused(field):
    text = b"P\tp\t" + field + b"\t*\n"
    with pytest.raises(pb.DepthError) as ei:
        pb.tokenize_steps(text, p_fields(text), 5, None)
    assert ei.value.code == binding.FGFA_ERR_PARSE


def test_parser_gpu_route_builds_the_identical_store(golden, fgfa_bin, tmp_path):
    """`fgfa -I x.gfa -o x.flatgfa` with the step lists tokenised on the GPU (forced with
    FGFA_GPU_PARSE=1) and on the host (=0) must write byte-identical files; a graph the
    tokenizer refuses (quirky step list) silently takes the host route."""
    import subprocess
    for c in golden:
        src = os.path.join(c["dir"], c["gfa"])
        a, b = tmp_path / "gpu.flatgfa", tmp_path / "host.flatgfa"
        subprocess.run([fgfa_bin, "-I", src, "-o", str(a)], check=True, env={**os.environ, "FGFA_GPU_PARSE": "1"})
        subprocess.run([fgfa_bin, "-I", src, "-o", str(b)], check=True, env={**os.environ, "FGFA_GPU_PARSE": "0"})
        assert a.read_bytes() == b.read_bytes(), c["name"]
        want = open(os.path.join(c["dir"], c["depth"]), "rb").read()
        got = subprocess.run([fgfa_bin, "-I", src, "depth", "-d"], capture_output=True, check=True,
                             env={**os.environ, "FGFA_GPU_PARSE": "1"}).stdout
        assert got == want, c["name"]
    quirky = tmp_path / "q.gfa"
    quirky.write_bytes(b"S\t1\tA\nS\t2\tC\nP\tx\t1+,2\t*\nP\ty\t2-x\t*\n")      # gfaline.rs quirks
    for flag in ("1", "0"):
        out = tmp_path / f"q{flag}.flatgfa"
        subprocess.run([fgfa_bin, "-I", str(quirky), "-o", str(out)], check=True, env={**os.environ, "FGFA_GPU_PARSE": flag})
    assert (tmp_path / "q1.flatgfa").read_bytes() == (tmp_path / "q0.flatgfa").read_bytes()
