"""A ``flatgfa``-module-compatible Python front end with the depth operators attached.

Mirrors the reference's Python binding (flatgfa-py/src/lib.rs, flatgfa-py/flatgfa.pyi):
``parse`` / ``parse_bytes`` / ``load``, ``FlatGFA.segments`` / ``.paths`` / ``.links`` as
list-like views with slicing and ``find``, ``Segment`` / ``Path`` / ``Handle`` / ``Link``
value objects with the reference's ``str()`` (GFA text, flatgfa/src/print.rs), equality and
hashing, ``write_flatgfa`` / ``write_gfa``.  The reference's binding has no depth method
(flatgfa-py/examples/depth.py loops over every step in Python); here the graph carries

    FlatGFA.depth()          -> (depth, uniq)      ops::depth::seg_depth_with_uniq  (depth.rs:15-39)
    FlatGFA.depth_table()    -> bytes              SegDepth::emit                   (depth.rs:61-82)
    FlatGFA.path_depth(...)  -> (lengths, means)   ops::depth::path_depth           (depth.rs:88-113)
    FlatGFA.window_depth(...)                      ops::window_depth                (window_depth.rs:183-197)

which run on the GPU through libflatgfa.so.  The list views read the
graph's .flatgfa image through numpy; nothing here computes depth on the CPU.  GAF and
pangenotype methods of the reference binding are outside this repository's scope.
"""
from __future__ import annotations

from typing import Iterator, Optional

import numpy as np

from . import binding

_MAGIC = 0xB1011054
_SEG = np.dtype([("name", "<u8"), ("s0", "<u4"), ("s1", "<u4"), ("o0", "<u4"), ("o1", "<u4")])
_PATH = np.dtype([("n0", "<u4"), ("n1", "<u4"), ("st0", "<u4"), ("st1", "<u4"), ("ov0", "<u4"), ("ov1", "<u4")])
_LINK = np.dtype([("from", "<u4"), ("to", "<u4"), ("a0", "<u4"), ("a1", "<u4")])
_SPAN = np.dtype([("a", "<u4"), ("b", "<u4")])
# Toc order (file.rs:16-26) and element sizes
_POOLS = [("header", 1), ("segs", 24), ("paths", 24), ("links", 16), ("steps", 4), ("seq_data", 1),
          ("overlaps", 8), ("alignment", 4), ("name_data", 1), ("optional_data", 1), ("line_order", 1)]
_OPCODES = "MNDI"   # print.rs:13-22: Match, Gap, Insertion -> "D", Deletion -> "I"


class _View:
    """The eleven pools of a .flatgfa image (file.rs:185-213) as numpy arrays (no copies)."""

    def __init__(self, image: np.ndarray):
        image = np.ascontiguousarray(image, dtype=np.uint8)
        toc = np.frombuffer(image[:184].tobytes(), dtype="<u8")
        if image.size < 184 or int(toc[0]) != _MAGIC:
            raise ValueError("not a .flatgfa image")
        off = 184
        raw = {}
        for i, (name, esz) in enumerate(_POOLS):
            n, cap = int(toc[1 + 2 * i]), int(toc[2 + 2 * i])
            raw[name] = image[off:off + n * esz]
            off += cap * esz
        self.image = image
        self.header = raw["header"]
        self.segs = raw["segs"].view(_SEG)
        self.paths = raw["paths"].view(_PATH)
        self.links = raw["links"].view(_LINK)
        self.steps = raw["steps"].view("<u4")
        self.seq_data = raw["seq_data"]
        self.overlaps = raw["overlaps"].view(_SPAN)
        self.alignment = raw["alignment"].view("<u4")
        self.name_data = raw["name_data"]
        self.optional_data = raw["optional_data"]
        self.line_order = raw["line_order"]

    def alignment_str(self, a: int, b: int) -> str:          # print.rs:24-34
        if a == b:
            return "0M"
        return "".join(f"{int(op) >> 8}{_OPCODES[int(op) & 0xFF]}" for op in self.alignment[a:b])

    def handle_str(self, h: int) -> str:                      # print.rs:39-45
        return f"{int(self.segs['name'][h >> 1])}{'-' if h & 1 else '+'}"


class Handle:
    """A segment and an orientation (flatgfa-py/src/lib.rs:691-744)."""

    __slots__ = ("_g", "_bits")

    def __init__(self, g: "FlatGFA", bits: int):
        self._g, self._bits = g, int(bits)

    @property
    def seg_id(self) -> int:
        return self._bits >> 1

    @property
    def is_forward(self) -> bool:
        return (self._bits & 1) == 0

    @property
    def segment(self) -> "Segment":
        return Segment(self._g, self._bits >> 1)

    def __repr__(self) -> str:
        return f"<Handle {self.seg_id}{'+' if self.is_forward else '-'}>"

    def __str__(self) -> str:
        return self._g._v.handle_str(self._bits)

    def __eq__(self, other) -> bool:
        return isinstance(other, Handle) and other._g is self._g and other._bits == self._bits

    def __hash__(self) -> int:
        return self.seg_id ^ ((0 if self.is_forward else 1) << 16)


class _Entity:
    __slots__ = ("_g", "id")
    _kind = "Entity"

    def __init__(self, g: "FlatGFA", index: int):
        self._g, self.id = g, int(index)

    def __repr__(self) -> str:                                # lib.rs:261-263
        return f"<{self._kind} {self.id}>"

    def __eq__(self, other) -> bool:
        return type(other) is type(self) and other._g is self._g and other.id == self.id

    def __hash__(self) -> int:
        return self.id


class Segment(_Entity):
    """flatgfa-py/src/lib.rs:341-399."""
    _kind = "Segment"

    @property
    def name(self) -> int:
        return int(self._g._v.segs["name"][self.id])

    def sequence(self) -> bytes:
        r = self._g._v.segs[self.id]
        return self._g._v.seq_data[int(r["s0"]):int(r["s1"])].tobytes()

    def __len__(self) -> int:
        r = self._g._v.segs[self.id]
        return int(r["s1"]) - int(r["s0"])

    def __str__(self) -> str:                                 # print.rs:89-98
        v, r = self._g._v, self._g._v.segs[self.id]
        s = f"S\t{int(r['name'])}\t{self.sequence().decode()}"
        if r["o0"] != r["o1"]:
            s += "\t" + v.optional_data[int(r["o0"]):int(r["o1"])].tobytes().decode()
        return s


class StepList:
    """A (slice of a) path's steps (flatgfa-py/src/lib.rs:755-797)."""

    def __init__(self, g: "FlatGFA", start: int, end: int):
        self._g, self._start, self._end = g, start, end

    def __len__(self) -> int:
        return self._end - self._start

    def __iter__(self) -> Iterator[Handle]:
        g = self._g
        for bits in g._v.steps[self._start:self._end]:
            yield Handle(g, bits)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            a, b, step = idx.indices(len(self))
            if step != 1:
                raise ValueError("only unit-stride slices are supported")
            return StepList(self._g, self._start + a, self._start + max(a, b))
        n = len(self)
        if idx < 0:
            idx += n
        if not 0 <= idx < n:
            raise IndexError("step index out of range")
        return Handle(self._g, self._g._v.steps[self._start + idx])


class Path(_Entity):
    """flatgfa-py/src/lib.rs:435-505: a path acts as a list of steps."""
    _kind = "Path"

    @property
    def name(self) -> str:
        r = self._g._v.paths[self.id]
        return self._g._v.name_data[int(r["n0"]):int(r["n1"])].tobytes().decode()

    @property
    def steps(self) -> StepList:
        r = self._g._v.paths[self.id]
        return StepList(self._g, int(r["st0"]), int(r["st1"]))

    def __iter__(self) -> Iterator[Handle]:
        return iter(self.steps)

    def __getitem__(self, idx):
        return self.steps[idx]

    def __len__(self) -> int:
        return len(self.steps)

    def __str__(self) -> str:                                 # print.rs:47-66
        v, r = self._g._v, self._g._v.paths[self.id]
        steps = ",".join(v.handle_str(int(h)) for h in v.steps[int(r["st0"]):int(r["st1"])])
        ov = v.overlaps[int(r["ov0"]):int(r["ov1"])]
        ovs = "*" if ov.size == 0 else ",".join(v.alignment_str(int(o["a"]), int(o["b"])) for o in ov)
        return f"P\t{self.name}\t{steps}\t{ovs}"


class Link(_Entity):
    """flatgfa-py/src/lib.rs:832-886."""
    _kind = "Link"

    @property
    def from_(self) -> Handle:
        return Handle(self._g, self._g._v.links["from"][self.id])

    @property
    def to(self) -> Handle:
        return Handle(self._g, self._g._v.links["to"][self.id])

    def __str__(self) -> str:                                 # print.rs:68-87
        v, r = self._g._v, self._g._v.links[self.id]
        f, t = int(r["from"]), int(r["to"])
        return (f"L\t{int(v.segs['name'][f >> 1])}\t{'-' if f & 1 else '+'}\t{int(v.segs['name'][t >> 1])}\t"
                f"{'-' if t & 1 else '+'}\t{v.alignment_str(int(r['a0']), int(r['a1']))}")


class _List:
    """List-like view with slicing (flatgfa-py/src/lib.rs:180-336)."""
    _item = _Entity

    def __init__(self, g: "FlatGFA", start: int, end: int):
        self._g, self._start, self._end = g, start, end

    def __len__(self) -> int:
        return self._end - self._start

    def __iter__(self):
        for i in range(self._start, self._end):
            yield self._item(self._g, i)

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            a, b, step = idx.indices(len(self))
            if step != 1:
                raise ValueError("only unit-stride slices are supported")
            return type(self)(self._g, self._start + a, self._start + max(a, b))
        n = len(self)
        if idx < 0:
            idx += n
        if not 0 <= idx < n:
            raise IndexError("index out of range")
        return self._item(self._g, self._start + idx)


class SegmentList(_List):
    _item = Segment

    def find(self, name: int) -> Optional[Segment]:           # lib.rs:415-428 (linear search over the slice)
        names = self._g._v.segs["name"][self._start:self._end]
        hit = np.nonzero(names == np.uint64(name))[0]
        return Segment(self._g, self._start + int(hit[0])) if hit.size else None


class PathList(_List):
    _item = Path

    def find(self, name) -> Optional[Path]:                   # lib.rs:677-687
        want = name.encode() if isinstance(name, str) else bytes(name)
        v = self._g._v
        for i in range(self._start, self._end):
            r = v.paths[i]
            if v.name_data[int(r["n0"]):int(r["n1"])].tobytes() == want:
                return Path(self._g, i)
        return None


class LinkList(_List):
    _item = Link


class FlatGFA:
    """An efficient representation of a pangenome graph (flatgfa-py/src/lib.rs:57-178), with the
    B200 depth operators attached."""

    def __init__(self, handle: binding.FlatGFA, image: Optional[np.ndarray] = None):
        self._h = handle
        self._v = _View(handle.image() if image is None else image)

    # ---- the reference binding's surface ----
    @property
    def segments(self) -> SegmentList:
        return SegmentList(self, 0, self._v.segs.size)

    @property
    def paths(self) -> PathList:
        return PathList(self, 0, self._v.paths.size)

    @property
    def links(self) -> LinkList:
        return LinkList(self, 0, self._v.links.size)

    @property
    def size(self) -> int:                                     # lib.rs:148-151: bytes of the flat image
        return int(self._v.image.size)

    def __str__(self) -> str:                                  # print.rs:100-127 (original line order)
        v = self._v
        if len(v.line_order) == 0:
            # print.rs:129-153 `write_normalized`: graphs built by ops that never call record_line
            # (extract, chop ...) have an empty line_order and are printed header, segments, paths, links
            return self._h.format_gfa().decode()
        out = []
        it = {1: iter(self.segments), 2: iter(self.paths), 3: iter(self.links)}
        for kind in v.line_order:
            k = int(kind)
            if k == 0:
                out.append("H\t" + v.header.tobytes().decode())
            else:
                out.append(str(next(it[k])))
        return "".join(line + "\n" for line in out)

    def write_gfa(self, filename: str) -> None:                # lib.rs:120-126
        with open(filename, "w", encoding="utf-8", newline="") as f:
            f.write(str(self))

    def write_flatgfa(self, filename: str) -> None:            # lib.rs:129-133
        self._h.dump(filename)

    def all_reads(self, gaf: str):                              # lib.rs:136-146
        raise NotImplementedError("GAF parsing is outside the scope of this package (node-depth path only)")

    def print_gaf_lookup(self, gaf: str) -> None:               # lib.rs:154-169
        raise NotImplementedError("GAF lookup is outside the scope of this package (node-depth path only)")

    def make_pangenotype_matrix(self, gaf_files):               # lib.rs:172-177
        raise NotImplementedError("the pangenotype matrix is outside the scope of this package (node-depth path only)")

    # ---- depth operators (GPU) ----
    def depth(self):
        """(depth, uniq): per segment, how many path steps cross it and how many distinct paths do."""
        return self._h.seg_depth_with_uniq()

    def depth_table(self) -> bytes:
        """The odgi-style ``#node.id depth depth.uniq`` table (``fgfa depth -d``)."""
        d, u = self._h.seg_depth_with_uniq()
        return self._h.format_seg_depth(d, u)

    def path_depth(self, paths=None):
        """(lengths, mean depths) of the given paths (Path objects, ids; default all)."""
        ids = None if paths is None else [p.id if isinstance(p, Path) else int(p) for p in paths]
        return self._h.path_depth(ids)

    def window_depth(self, path, window_size: int) -> bytes:
        """``fgfa window-depth PATH SIZE``: name/start/end/mean-depth rows for equal windows along a path."""
        name = path.name if isinstance(path, Path) else path
        return self._h.window_depth(name, window_size)

    def bed_depth(self, bed_text: bytes) -> bytes:
        """``fgfa depth -b BED``."""
        return self._h.bed_depth(bed_text)

    def close(self) -> None:
        self._h.close()


def parse(filename: str) -> FlatGFA:
    """Parse a GFA text file (lib.rs:63-65)."""
    return FlatGFA(binding.FlatGFA.parse(filename))


def parse_bytes(gfa: bytes) -> FlatGFA:
    """Parse GFA text held in a bytes object (lib.rs:69-71)."""
    return FlatGFA(binding.FlatGFA.parse_bytes(bytes(gfa)))


def load(filename: str) -> FlatGFA:
    """Map a binary .flatgfa file (lib.rs:79-81)."""
    return FlatGFA(binding.FlatGFA.load(filename), np.memmap(filename, dtype=np.uint8, mode="r"))
