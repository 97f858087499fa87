"""Static checks on the SASS of the shipped library (CPU only: cuobjdump reads the sm_100a cubin).
They pin the claims DESIGN.md makes about the hand-written kernels: the step stream is staged with
cp.async (LDGSTS), the counters are updated with fire-and-forget L2 reductions (REDG, never a
returning ATOMG), the exchange kernel uses multicast stores, and no hot kernel spills to local memory."""
import collections
import os
import re
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pollen_b200", "lib", "libflatgfa.so")


@pytest.fixture(scope="module")
def sass():
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    assert "sm_100a" in out
    hist, cur = collections.defaultdict(collections.Counter), None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            hist[cur][m.group(1)] += 1
    return hist


def _kernels(hist, needle):
    ks = {k: v for k, v in hist.items() if needle in k}
    assert ks, needle
    return ks


def test_only_sm_100a_code_is_shipped():
    out = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout if shutil.which("cuobjdump") else ""
    if not out:
        pytest.skip("cuobjdump not available")
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_step_stream_kernels(sass):
    for name, ops in _kernels(sass, "k_step_stream_merged").items():
        assert ops["LDGSTS"] > 0, name            # cp.async staging
        assert ops["REDG"] > 0, name              # RED.ADD / RED.OR
        assert ops["ATOMG"] == 0, name            # no returning global atomics on the hot path
        assert ops["STL"] == 0 and ops["LDL"] == 0, name   # no local-memory spills


def test_popcount_measure_and_interval_kernels_do_not_spill(sass):
    for needle in ("k_uniq_popcount", "k_path_measure", "k_tile_scan", "k_tile_reduce", "k_interval_lower_bound",
                   "k_steps_count", "k_steps_parse"):
        for name, ops in _kernels(sass, needle).items():
            assert ops["STL"] == 0 and ops["LDL"] == 0, name
    for name, ops in _kernels(sass, "k_interval_accumulate").items():
        assert ops["DADD"] > 0 and ops["DFMA"] >= 0, name
    # the ordered f64 sum must stay a chain of plain adds: no fused multiply-add may absorb the
    # product or the quotient of a term (DFMA only appears inside the division sequences)
    for name, ops in _kernels(sass, "k_interval_accumulate_long").items():
        assert ops["DADD"] >= 32, name            # the 32 unrolled in-order additions of a batch


def test_exchange_kernel_uses_multicast(sass):
    """kernel X: multimem.ld_reduce compiles to LDGMC, multimem.st to system-scope 128-bit stores."""
    for name, ops in _kernels(sass, "k_uniq_exchange").items():
        assert ops["LDGMC"] > 0, name
        assert ops["STG"] > 0 and ops["ATOMG"] == 0, name


def test_window_kernel_counts_in_shared_memory_and_overlaps_its_loads():
    """kernel W (the shipped OVL instantiations): shared-memory ATOMS for counters and path masks, descriptors
    by shuffle, no spills -- and, in program order, every burst of ATOMS is preceded by the step loads of the
    NEXT sub-chunk (the order DESIGN.md 4.1 relies on: ptxas tracks all step loads on one scoreboard, so
    'count, then refill' would expose the DRAM latency in every iteration)."""
    if shutil.which("cuobjdump") is None:
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    funcs, cur = {}, None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            funcs[cur].append(m.group(1))
    ws = {k: v for k, v in funcs.items() if "k_window_count" in k}
    assert len(ws) == 2, list(ws)                       # with and without path masks
    for name, ops in ws.items():
        base = [o.split(".")[0] for o in ops]
        assert "STL" not in base and "LDL" not in base, name
        assert base.count("ATOMS") >= 8 and "SHFL" in base and "REDG" in base, name
        # first ATOMS burst of the main loop: the 8 streaming step loads (LDG.E.NA...) just before it
        first = base.index("ATOMS")
        window = ops[max(0, first - 400):first]
        assert sum(o.startswith("LDG.E.NA") for o in window) >= 8, name
        # ... and none between the bursts' ATOMS (the refill is not interleaved after the counting starts)
        last = len(base) - 1 - base[::-1].index("ATOMS")
        assert base[first:last].count("ATOMS") >= 16, name
