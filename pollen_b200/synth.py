"""Synthetic pangenome generators for the BASELINE.json configurations (bench/test
infrastructure; see pollen_b200/csrc/synth.cpp and SURVEY.md §8(d) for the definitions)."""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "lib", "libfgfa_synth.so")
_lib = None

SEED = 0xB1011054
KIND_WALK, KIND_SKEWED, KIND_UNIFORM, KIND_SORTED = 0, 1, 2, 3


@dataclass(frozen=True)
class Config:
    name: str
    n_segs: int
    n_paths: int
    n_steps: int
    kind: int
    jitter_pct: int
    description: str


CONFIGS = {
    # BASELINE.json configs[1]
    "B": Config("B", 1_000_000, 16, 20_000_000, KIND_WALK, 0,
                "synthetic graph 1M segments, 16 paths, 20M steps"),
    # BASELINE.json configs[2] (the one the metric is quoted on) and [3] (sharded)
    "C": Config("C", 5_000_000, 90, 400_000_000, KIND_WALK, 20,
                "HPRC-chromosome-scale synthetic: 5M segments, 90 paths, 400M steps"),
    # BASELINE.json configs[4]
    "E": Config("E", 5_000_000, 8, 400_000_000, KIND_SKEWED, 0,
                "skewed synthetic: 8 long looping paths, hot 4096-segment windows"),
    # adversarial extra: uniform-random segment ids
    "U": Config("U", 5_000_000, 90, 400_000_000, KIND_UNIFORM, 20, "uniform-random segment ids"),
    # not in BASELINE.json: dense, id-sorted haplotype walks (what `odgi sort`ed HPRC graphs look like)
    "R": Config("R", 5_000_000, 90, 400_000_000, KIND_SORTED, 20, "sorted haplotype walks, 5M segments, 90 paths, 400M steps"),
    # small shapes for tests
    "tiny": Config("tiny", 5_000, 7, 100_003, KIND_WALK, 30, "test-sized haplotype walk"),
    "tinyE": Config("tinyE", 20_000, 5, 300_007, KIND_SKEWED, 10, "test-sized skewed loops"),
}


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} is missing: build it with `make`")
        h = C.CDLL(_LIB_PATH)
        h.fgfa_synth_spans.restype = C.c_int
        h.fgfa_synth_spans.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p]
        h.fgfa_synth_steps.restype = C.c_int
        h.fgfa_synth_steps.argtypes = [C.c_int, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_int]
        h.fgfa_synth_path.restype = C.c_int
        h.fgfa_synth_path.argtypes = [C.c_int, C.c_uint32, C.c_uint64, C.c_uint64, C.c_uint32, C.c_void_p]
        _lib = h
    return _lib


def make_spans(n_paths: int, n_steps: int, jitter_pct: int = 0, seed: int = SEED):
    start = np.empty(n_paths, dtype=np.uint32)
    end = np.empty(n_paths, dtype=np.uint32)
    if lib().fgfa_synth_spans(n_paths, n_steps, jitter_pct, seed, start.ctypes.data, end.ctypes.data):
        raise ValueError("bad span request")
    return start, end


def fill_steps(kind: int, n_segs: int, span_start, span_end, out: np.ndarray, seed: int = SEED, threads: int = 0) -> None:
    """Generate the steps of every path into ``out`` (a uint32 array covering the pool)."""
    assert out.dtype == np.uint32 and out.flags.c_contiguous
    threads = threads or min(64, os.cpu_count() or 1)
    if lib().fgfa_synth_steps(kind, n_segs, len(span_start), span_start.ctypes.data, span_end.ctypes.data, seed, out.ctypes.data, threads):
        raise ValueError("bad synth request")


def make_graph(cfg: Config, seed: int = SEED, out: np.ndarray | None = None, path_subset=None):
    """Returns (steps, span_start, span_end) for a configuration.  With ``path_subset``
    (indices) only those paths are generated, packed back to back, bit-identical to the
    same paths of the full graph (paths are seeded individually) -- this is how a rank
    materialises just its shard."""
    start, end = make_spans(cfg.n_paths, cfg.n_steps, cfg.jitter_pct, seed)
    if path_subset is None:
        steps = out if out is not None else np.empty(cfg.n_steps, dtype=np.uint32)
        fill_steps(cfg.kind, cfg.n_segs, start, end, steps[: cfg.n_steps], seed)
        return steps, start, end
    idx = np.asarray(path_subset, dtype=np.int64)
    lens = (end[idx] - start[idx]).astype(np.int64)
    total = int(lens.sum())
    steps = out if out is not None else np.empty(total, dtype=np.uint32)
    new_end = np.cumsum(lens).astype(np.uint32)
    new_start = (new_end - lens).astype(np.uint32)
    # paths are seeded individually (seed + original path index): generate one at a time
    h = lib()
    threads = min(64, os.cpu_count() or 1)
    import concurrent.futures as cf

    def one(k):
        lo, hi = int(new_start[k]), int(new_end[k])
        view = steps[lo:hi]
        return h.fgfa_synth_path(cfg.kind, cfg.n_segs, hi - lo, seed, int(idx[k]), view.ctypes.data if hi > lo else None)

    with cf.ThreadPoolExecutor(threads) as ex:
        for rc in ex.map(one, range(len(idx))):
            if rc:
                raise ValueError("bad synth request")
    return steps, new_start, new_end
