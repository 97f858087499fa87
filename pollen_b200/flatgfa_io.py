"""Write synthetic graphs as real ``.flatgfa`` images (the layout of the reference's
``file::dump``, flatgfa/src/file.rs:290-307, with the record layouts of
flatgfa/src/flatgfa.rs:71-133).  Test/bench infrastructure: lets the CLI and the
host-buffer C ABI be exercised on generated graphs."""
from __future__ import annotations

import numpy as np

MAGIC = 0xB1011054


def build_image(steps: np.ndarray, span_start, span_end, n_segs: int, header: bytes = b"VN:Z:1.0",
                seg_names=None, slack: int = 0, record_lines: bool = True) -> np.ndarray:
    """Returns the file image as a uint8 array.  Segments are named 1..n_segs (or
    ``seg_names``) with 1-byte sequences; paths are named p0, p1, ...; no links.
    ``slack`` extra capacity slots are left in every pool (capacity > len).  ``record_lines=False``
    leaves line_order empty, like the graphs the reference's extract / chop build (they never call
    record_line): printing such a graph takes print.rs's normalised branch."""
    steps = np.ascontiguousarray(steps, dtype=np.uint32)
    span_start = np.asarray(span_start, dtype=np.uint32)
    span_end = np.asarray(span_end, dtype=np.uint32)
    n_paths = len(span_start)
    names = np.arange(1, n_segs + 1, dtype=np.uint64) if seg_names is None else np.asarray(seg_names, dtype=np.uint64)

    segs = np.zeros(n_segs, dtype=np.dtype([("name", "<u8"), ("s0", "<u4"), ("s1", "<u4"), ("o0", "<u4"), ("o1", "<u4")]))
    segs["name"] = names
    segs["s0"] = np.arange(n_segs, dtype=np.uint32)
    segs["s1"] = np.arange(1, n_segs + 1, dtype=np.uint32)
    seq_data = np.frombuffer(b"ACGT", dtype=np.uint8)[np.arange(n_segs) % 4]

    path_names = [b"p%d" % i for i in range(n_paths)]
    name_data = np.frombuffer(b"".join(path_names), dtype=np.uint8)
    name_end = np.cumsum([len(x) for x in path_names], dtype=np.uint64).astype(np.uint32) if n_paths else np.zeros(0, np.uint32)
    paths = np.zeros(n_paths, dtype=np.dtype([("n0", "<u4"), ("n1", "<u4"), ("st0", "<u4"), ("st1", "<u4"), ("ov0", "<u4"), ("ov1", "<u4")]))
    if n_paths:
        paths["n1"] = name_end
        paths["n0"] = np.concatenate([[0], name_end[:-1]])
        paths["st0"] = span_start
        paths["st1"] = span_end
    line_order = np.concatenate([
        np.full(1 if header else 0, 0, np.uint8), np.full(n_segs, 1, np.uint8), np.full(n_paths, 2, np.uint8)])
    if not record_lines:
        line_order = np.zeros(0, np.uint8)

    pools = [
        (np.frombuffer(header, dtype=np.uint8), 1),
        (segs.view(np.uint8), 24),
        (paths.view(np.uint8), 24),
        (np.zeros(0, np.uint8), 16),          # links
        (steps.view(np.uint8), 4),
        (seq_data, 1),
        (np.zeros(0, np.uint8), 8),           # overlaps
        (np.zeros(0, np.uint8), 4),           # alignment
        (name_data, 1),
        (np.zeros(0, np.uint8), 1),           # optional_data
        (line_order, 1),
    ]
    toc = np.zeros(1 + 22, dtype="<u8")
    toc[0] = MAGIC
    total = 184
    for i, (buf, esz) in enumerate(pools):
        n = buf.size // esz
        toc[1 + 2 * i] = n
        toc[2 + 2 * i] = n + slack
        total += (n + slack) * esz
    img = np.zeros(total, dtype=np.uint8)
    img[:184] = toc.view(np.uint8)
    off = 184
    for buf, esz in pools:
        img[off:off + buf.size] = buf
        off += buf.size + slack * esz
    return img


def write_flatgfa(path: str, *args, **kw) -> None:
    build_image(*args, **kw).tofile(path)
