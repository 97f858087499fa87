"""`flash`: the FlatGFA fake shell, with the depth instructions running on the GPU.

Mirrors the reference's flatgfa-sh crate: a shell-syntax front end whose
`odgi depth ...` / `bedtools makewindows ...` commands are translated to a small IR, optionally
optimised, and evaluated in-process.  Same structure, same names:

    parse_sh / sh_to_ir      flatgfa-sh/src/parse.rs    shell text -> IR (one instruction per action)
    Builder                  flatgfa-sh/src/builder.rs  resources, load_gfa / load_bed / maybe_decompress
    optimize                 flatgfa-sh/src/opt.rs      the six rewrites behind `-O`
    Program.__str__          flatgfa-sh/src/pretty.rs   what `-p` (pretend mode) prints
    run                      flatgfa-sh/src/eval/*.rs   the evaluator

The evaluator's node-depth, path-depth, path-length and interval-depth instructions call the
same entry points as everything else in this package (libflatgfa.so: kernels A, B, C, W1-W3);
nothing is computed on the CPU here beyond text plumbing.  Pipes are in-memory byte buffers
rather than OS pipes (the reference's sequential evaluator would block on a full OS pipe).

    python -m pollen_b200.flash [-p] [-O] (-c COMMAND | SCRIPT)
"""
from __future__ import annotations

import gzip
import os
import subprocess
import sys
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

# ---- IR (flatgfa-sh/src/ir.rs) -------------------------------------------------------------

FILE, STDIN, STDOUT, PIPE, GFA_STORE, MMAP, BED_STORE = "file", "stdin", "stdout", "pipe", "gfa-store", "mmap", "bed-store"
BYTE_STREAMS = (FILE, MMAP, PIPE, STDIN, STDOUT)


@dataclass(frozen=True)
class Rsrc:
    """ir.rs:10-15 ResourceRef: per-kind index spaces; `gz` marks an encoded byte stream."""
    kind: str
    index: int = 0
    gz: bool = False

    def encoded(self) -> "Rsrc":                      # ir.rs:98-107
        assert self.kind in BYTE_STREAMS
        return Rsrc(self.kind, self.index, True)


@dataclass
class Instr:
    """ir.rs:41-46.  `op` is the instruction name as pretty.rs prints it; `arg` its parameter:
    path-depth: Optional[str] path; path-length: str path; make-windows: int size;
    shell: (command, [args])."""
    inputs: List[Rsrc]
    output: Rsrc
    op: str
    arg: object = None


def _rust_debug(s: str) -> str:
    """`{:?}` of a Rust String."""
    out = ['"']
    for ch in s:
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\t":
            out.append("\\t")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\0":
            out.append("\\0")
        elif ord(ch) < 0x20 or ord(ch) == 0x7F:
            out.append("\\u{%x}" % ord(ch))
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


@dataclass
class Program:
    instrs: List[Instr]
    file_names: List[str]
    counts: Dict[str, int]

    def _rsrc(self, r: Rsrc) -> str:                  # pretty.rs:75-91
        pre = "gz " if r.gz else ""
        if r.kind == FILE:
            return f'{pre}"{self.file_names[r.index]}"'
        if r.kind in (STDIN, STDOUT):
            return pre + r.kind
        return f"{pre}{r.kind}-{r.index}"

    def _instr(self, i: Instr) -> str:                # pretty.rs:22-73
        if i.op == "shell":
            command, args = i.arg
            return "shell(%s, [%s], input=%s) -> %s" % (
                _rust_debug(command), ", ".join(_rust_debug(a) for a in args), self._rsrc(i.inputs[0]), self._rsrc(i.output))
        s = f"{i.op}({self._rsrc(i.inputs[0])}"
        for r in i.inputs[1:]:
            s += ", " + self._rsrc(r)
        if i.op == "path-depth" and i.arg is not None:
            s += ", path=" + _rust_debug(i.arg)
        elif i.op == "path-length":
            s += ", path=" + _rust_debug(i.arg)
        elif i.op == "make-windows":
            s += f", size={i.arg}"
        return s + f") -> {self._rsrc(i.output)}"

    def __str__(self) -> str:                         # pretty.rs:93-107
        return "".join(self._instr(i) + "\n" for i in self.instrs)


# ---- builder (flatgfa-sh/src/builder.rs) ----------------------------------------------------

class Builder:
    def __init__(self, prog: Optional[Program] = None):
        self.instrs: List[Instr] = prog.instrs if prog else []
        self.file_names: List[str] = prog.file_names if prog else []
        self.files: Dict[str, int] = {n: i for i, n in enumerate(self.file_names)}
        self.counts: Dict[str, int] = prog.counts if prog else {}

    def instr(self, inputs, output, op, arg=None) -> None:
        self.instrs.append(Instr(list(inputs), output, op, arg))

    def file(self, name: str) -> Rsrc:                # builder.rs:53-63
        if name not in self.files:
            self.files[name] = len(self.files)
            self.file_names.append(name)
        return Rsrc(FILE, self.files[name])

    def file_name(self, r: Rsrc) -> str:
        assert r.kind == FILE
        return self.file_names[r.index]

    def rsrc(self, kind: str) -> Rsrc:                # builder.rs:71-75
        i = self.counts.get(kind, 0)
        self.counts[kind] = i + 1
        return Rsrc(kind, i)

    def load_gfa(self, inp: Rsrc) -> Rsrc:            # builder.rs:83-110
        if inp.kind == FILE and self.file_name(inp).endswith(".flatgfa"):
            out = self.rsrc(MMAP)
            self.instr([inp], out, "map-file")
            return out
        if inp.kind == FILE and self.file_name(inp).endswith(".og"):
            pipe = self.rsrc(PIPE)
            self.instr([inp], pipe, "odgi-view")
            return self.load_gfa(pipe)
        if inp.kind in (PIPE, STDIN, FILE):
            inp = self.maybe_decompress(inp)
            out = self.rsrc(GFA_STORE)
            self.instr([inp], out, "parse-gfa")
            return out
        raise ValueError("cannot parse this resource as GFA text")

    def load_bed(self, inp: Rsrc) -> Rsrc:            # builder.rs:114-124
        if inp.kind in (PIPE, STDIN, FILE):
            inp = self.maybe_decompress(inp)
            out = self.rsrc(BED_STORE)
            self.instr([inp], out, "parse-bed")
            return out
        raise ValueError("cannot parse this resource as BED text")

    def maybe_decompress(self, inp: Rsrc) -> Rsrc:    # builder.rs:131-140
        if inp.kind == FILE and self.file_name(inp).endswith(".gz"):
            pipe = self.rsrc(PIPE)
            self.instr([inp], pipe, "gzip-decompress")
            return pipe
        return inp

    def replace_rsrc(self, old: Rsrc, new: Rsrc) -> None:   # builder.rs:143-154
        for i in self.instrs:
            i.inputs = [new if r == old else r for r in i.inputs]
            if i.output == old:
                i.output = new

    def build(self) -> Program:
        return Program(self.instrs, self.file_names, self.counts)


# ---- shell syntax (flatgfa-sh/src/parse.rs over brush-parser) -------------------------------

class Unsupported(NotImplementedError):
    """Syntax the reference leaves `unimplemented!()`."""


@dataclass
class SimpleCommand:
    words: List[str] = field(default_factory=list)
    redirects: List[Tuple[str, str]] = field(default_factory=list)   # ("<" | ">", filename)


def parse_sh(text: str) -> List[List[SimpleCommand]]:
    """Shell text -> list of pipelines (each a list of simple commands).  The subset the
    reference handles: words with quotes and backslash escapes (`word_str`, parse.rs:222-254),
    `|`, `;` / newlines, `<` and `>` redirections, `#` comments."""
    pipelines: List[List[SimpleCommand]] = []
    pipeline: List[SimpleCommand] = []
    cmd = SimpleCommand()
    pending: Optional[str] = None       # a redirection operator waiting for its filename
    i, n = 0, len(text)

    def end_command():
        nonlocal cmd
        if pending is not None:
            raise Unsupported("redirection without a filename")
        if cmd.words or cmd.redirects:
            if not cmd.words:
                raise Unsupported("command name")           # `simple.word_or_name.expect("command name")`
            pipeline.append(cmd)
        cmd = SimpleCommand()

    def end_pipeline():
        nonlocal pipeline
        end_command()
        if pipeline:
            pipelines.append(pipeline)
        pipeline = []

    while i < n:
        c = text[i]
        if c in " \t":
            i += 1
        elif c == "\n" or c == ";":
            end_pipeline()
            i += 1
        elif c == "#":
            while i < n and text[i] != "\n":
                i += 1
        elif c == "|":
            if text[i:i + 2] == "||":
                raise Unsupported("&& and || not supported")
            end_command()
            if not pipeline:
                raise Unsupported("pipe without a command")
            i += 1
        elif c == "&":
            raise Unsupported("&& and || not supported" if text[i:i + 2] == "&&" else "async commands not supported")
        elif c in "<>":
            if text[i:i + 2] in (">>", "<<", "<&", ">&", "<>", ">|"):
                raise Unsupported("only < and > redirections are supported")
            if pending is not None:
                raise Unsupported("redirection without a filename")
            pending = c
            i += 1
        else:
            word = []
            while i < n and text[i] not in " \t\n;|&<>":
                c = text[i]
                if c == "\\":
                    if i + 1 >= n:
                        raise Unsupported("trailing backslash")
                    if text[i + 1] != "\n":                  # backslash-newline is a line continuation
                        word.append(text[i + 1])
                    i += 2
                elif c == "'":
                    j = text.find("'", i + 1)
                    if j < 0:
                        raise Unsupported("unterminated single quote")
                    word.append(text[i + 1:j])
                    i = j + 1
                elif c == '"':
                    i += 1
                    while i < n and text[i] != '"':
                        if text[i] == "\\" and i + 1 < n and text[i + 1] in '"\\$`\n':
                            if text[i + 1] != "\n":
                                word.append(text[i + 1])
                            i += 2
                        elif text[i] in "$`":
                            raise Unsupported("expansions are not supported")
                        else:
                            word.append(text[i])
                            i += 1
                    if i >= n:
                        raise Unsupported("unterminated double quote")
                    i += 1
                elif c in "$`":
                    raise Unsupported("expansions are not supported")
                else:
                    word.append(c)
                    i += 1
            w = "".join(word)
            if pending is not None:
                cmd.redirects.append((pending, w))
                pending = None
            else:
                cmd.words.append(w)
    end_pipeline()
    return pipelines


class _Args:
    """The pico-args calls parse.rs makes: options are taken out of the list wherever they are."""

    def __init__(self, args: List[str]):
        self.args = list(args)

    def opt_value(self, *keys: str) -> Optional[str]:
        for k, a in enumerate(self.args):
            for key in keys:
                if a == key:
                    if k + 1 >= len(self.args):
                        raise Unsupported(f"the '{key}' option doesn't have an associated value")
                    v = self.args[k + 1]
                    del self.args[k:k + 2]
                    return v
                if a.startswith(key + "="):
                    del self.args[k]
                    return a[len(key) + 1:]
        return None

    def contains(self, key: str) -> bool:
        if key in self.args:
            self.args.remove(key)
            return True
        return False

    def subcommand(self) -> Optional[str]:
        if self.args and not self.args[0].startswith("-"):
            return self.args.pop(0)
        return None


def _translate_odgi(b: Builder, args: List[str], inp: Rsrc, out: Rsrc) -> None:   # parse.rs:73-120
    argp = _Args(args)
    filename = argp.opt_value("-i", "--input")
    if filename is not None:
        inp = b.file(filename)
    gfa = b.load_gfa(inp)
    if argp.subcommand() != "depth":
        raise Unsupported("unsupported odgi subcommand")
    if argp.contains("-d"):
        b.instr([gfa], out, "node-depth")
        return
    bed_file = argp.opt_value("-b")
    if bed_file is not None:
        bed = b.load_bed(b.file(bed_file))
        b.instr([gfa, bed], out, "interval-depth")
        return
    b.instr([gfa], out, "path-depth", argp.opt_value("-r"))


def _translate_bedtools(b: Builder, args: List[str], inp: Rsrc, out: Rsrc) -> None:   # parse.rs:123-154
    argp = _Args(args)
    if argp.subcommand() != "makewindows":
        raise Unsupported("unsupported bedtools subcommand")
    filename = argp.opt_value("-b")
    if filename is None:
        raise Unsupported("missing option '-b'")
    if filename != "/dev/stdin":
        inp = b.file(filename)
    bed = b.load_bed(inp)
    size = argp.opt_value("-w")
    if size is None:
        raise Unsupported("missing option '-w'")
    b.instr([bed], out, "make-windows", int(size))


def _translate_command(b: Builder, c: SimpleCommand, inp: Rsrc, out: Rsrc) -> None:   # parse.rs:24-70
    for op, filename in c.redirects:
        if op == "<":
            inp = b.file(filename)
        else:
            out = b.file(filename)
    name, args = c.words[0], c.words[1:]
    if name == "odgi":
        _translate_odgi(b, args, inp, out)
    elif name == "bedtools":
        _translate_bedtools(b, args, inp, out)
    elif name == "gunzip":
        if args:
            raise Unsupported("no gunzip arguments are supported")
        b.instr([inp], out, "gzip-decompress")
    else:
        b.instr([inp], out, "shell", (name, args))


def sh_to_ir(pipelines: List[List[SimpleCommand]]) -> Program:   # parse.rs:177-219
    b = Builder()
    for pipeline in pipelines:
        inp = Rsrc(STDIN)
        for k, c in enumerate(pipeline):
            out = Rsrc(STDOUT) if k == len(pipeline) - 1 else b.rsrc(PIPE)
            _translate_command(b, c, inp, out)
            inp = out
    return b.build()


# ---- optimizer (flatgfa-sh/src/opt.rs) ------------------------------------------------------

def _def_use(instrs: List[Instr]):                    # opt.rs:398-430
    defs, uses, last = [], [[] for _ in instrs], {}
    for idx, i in enumerate(instrs):
        defs.append([last.get(r) for r in i.inputs])
        for r in i.inputs:
            if r in last:
                uses[last[r]].append(idx)
        last[i.output] = idx
    return defs, uses


def _drop(b: Builder, indices) -> None:
    dead = set(indices)
    b.instrs[:] = [i for k, i in enumerate(b.instrs) if k not in dead]


def _replace_with_flat(b: Builder, stem: str, idx: int) -> bool:   # opt.rs:352-381
    flat = stem + ".flatgfa"
    if not os.path.exists(flat):
        return False
    new = b.rsrc(MMAP)
    old = b.instrs[idx].output
    b.instrs[idx] = Instr([b.file(flat)], new, "map-file")
    b.replace_rsrc(old, new)
    return True


def _strip_suffix(name: str, suffix: str, what: str) -> str:
    if not name.endswith(suffix):
        raise ValueError(f"{what} inputs must end in {suffix}")
    return name[: -len(suffix)]


def _opt_gfa_parse(b: Builder) -> None:               # opt.rs:91-128
    for idx in [k for k, i in enumerate(b.instrs) if i.op == "parse-gfa" and i.inputs[0].kind == FILE]:
        _replace_with_flat(b, _strip_suffix(b.file_name(b.instrs[idx].inputs[0]), ".gfa", "parse-gfa"), idx)


def _opt_og_parse(b: Builder) -> None:                # opt.rs:36-88
    defs, _ = _def_use(b.instrs)
    pairs = [(defs[k][0], k) for k, i in enumerate(b.instrs)
             if i.op == "parse-gfa" and defs[k][0] is not None and b.instrs[defs[k][0]].op == "odgi-view"]
    dead = []
    for view_idx, parse_idx in pairs:
        stem = _strip_suffix(b.file_name(b.instrs[view_idx].inputs[0]), ".og", "odgi-view")
        if _replace_with_flat(b, stem, parse_idx):
            dead.append(view_idx)
        elif os.path.exists(stem + ".gfa"):
            b.instrs[parse_idx].inputs[0] = b.file(stem + ".gfa")
            dead.append(view_idx)
    _drop(b, dead)


def _skip_bed_files(b: Builder) -> None:              # opt.rs:140-186
    defs, uses = _def_use(b.instrs)
    dead = []
    for k, i in enumerate(b.instrs):
        if i.op != "parse-bed" or defs[k][0] is None:
            continue
        d = defs[k][0]
        if len(uses[d]) == 1 and b.instrs[d].op in ("make-windows", "path-depth"):
            b.instrs[d].output = i.output
            dead.append(k)
    _drop(b, dead)


def _simplify_depth_to_length(b: Builder) -> None:    # opt.rs:203-238
    defs, uses = _def_use(b.instrs)
    for k, i in enumerate(b.instrs):
        if i.op != "make-windows" or defs[k][0] is None:
            continue
        d = b.instrs[defs[k][0]]
        if len(uses[defs[k][0]]) == 1 and d.op == "path-depth" and d.arg is not None:
            d.op = "path-length"


def _dedup_files(b: Builder) -> None:                 # opt.rs:253-293
    seen: Dict[Rsrc, Rsrc] = {}
    redundant = []
    for k, i in enumerate(b.instrs):
        if i.op == "map-file":
            if i.inputs[0] in seen:
                redundant.append(k)
            else:
                seen[i.inputs[0]] = i.output
        if i.output.kind == FILE:
            seen.pop(i.output, None)
    for k in redundant:
        if b.instrs[k].inputs[0] not in seen:
            raise ValueError("original file not found")
        b.replace_rsrc(b.instrs[k].output, seen[b.instrs[k].inputs[0]])
    _drop(b, redundant)


def _opt_decompress(b: Builder) -> None:              # opt.rs:310-349
    _, uses = _def_use(b.instrs)
    decomp = [k for k, i in enumerate(b.instrs)
              if i.op == "gzip-decompress" and all(b.instrs[u].op == "parse-gfa" for u in uses[k])]
    for k in decomp:
        b.replace_rsrc(b.instrs[k].output, b.instrs[k].inputs[0].encoded())
    _drop(b, decomp)


def optimize(prog: Program) -> Program:               # opt.rs:8-22
    b = Builder(prog)
    _opt_gfa_parse(b)
    _opt_og_parse(b)
    _skip_bed_files(b)
    _simplify_depth_to_length(b)
    _dedup_files(b)
    _opt_decompress(b)
    return b.build()


# ---- evaluator (flatgfa-sh/src/eval) --------------------------------------------------------

class _Env:
    def __init__(self, prog: Program, stdin: Optional[bytes], stdout):
        self.prog = prog
        self.pipes: Dict[int, bytes] = {}
        self.graphs: Dict[Tuple[str, int], object] = {}
        self.beds: Dict[int, List[Tuple[bytes, int, int]]] = {}
        self._stdin = stdin
        self.stdout = stdout

    def name(self, r: Rsrc) -> str:
        return self.prog.file_names[r.index]

    def read(self, r: Rsrc) -> Tuple[bytes, bool]:
        """Bytes of a byte-stream resource and whether it is a stream (stdin / pipe) rather than a file."""
        if r.kind == FILE:
            with open(self.name(r), "rb") as f:
                data, stream = f.read(), False
        elif r.kind == PIPE:
            data, stream = self.pipes.pop(r.index), True
        elif r.kind == STDIN:
            if self._stdin is None:
                self._stdin = sys.stdin.buffer.read()
            data, stream = self._stdin, True
            self._stdin = b""
        else:
            raise ValueError("text input")
        if r.gz:
            data = gzip.decompress(data)
        return data, stream

    def write(self, r: Rsrc, data: bytes) -> None:
        if r.kind == FILE:
            with open(self.name(r), "wb") as f:
                f.write(data)
        elif r.kind == PIPE:
            self.pipes[r.index] = data
        elif r.kind == STDOUT:
            self.stdout.write(data)
            self.stdout.flush()
        else:
            raise ValueError("bytes output")

    def graph(self, r: Rsrc):
        return self.graphs[(r.kind, r.index)]


def _terminated(data: bytes) -> bytes:
    """The reference parses streams with `parse_stream` (keeps an unterminated last line) and files
    with `parse_mem` (drops it); the library's parse_mem applied to newline-terminated text is both."""
    return data if not data or data.endswith(b"\n") else data + b"\n"


def _bed_text(entries) -> bytes:
    return b"".join(b"%s\t%d\t%d\n" % e for e in entries)


def _find_path(g, name: str) -> int:
    want = name.encode()
    for p in range(g.path_count):
        if g.path_name(p) == want:
            return p
    raise ValueError("no such path found")             # eval/instr.rs:47, :89


def _eval(env: _Env, i: Instr) -> None:               # eval/instr.rs:11-24
    from . import binding

    if i.op == "shell" or i.op == "odgi-view":
        command, args = i.arg if i.op == "shell" else ("odgi", ["view", "-g", "-i", env.name(i.inputs[0])])
        src = i.inputs[0] if i.op == "shell" else Rsrc(STDIN)
        data = None
        if src.kind != STDIN or env._stdin is not None:
            data, _ = env.read(src)
        capture = i.output.kind != STDOUT or env.stdout is not sys.stdout.buffer
        r = subprocess.run([command] + list(args), input=data, stdout=subprocess.PIPE if capture else None)
        if capture:
            env.write(i.output, r.stdout)
    elif i.op == "gzip-decompress":
        data, _ = env.read(i.inputs[0])
        env.write(i.output, gzip.decompress(data))
    elif i.op == "parse-gfa":
        data, stream = env.read(i.inputs[0])
        env.graphs[(i.output.kind, i.output.index)] = binding.FlatGFA.parse_bytes(_terminated(data) if stream or i.inputs[0].gz else data)
    elif i.op == "map-file":
        env.graphs[(i.output.kind, i.output.index)] = binding.FlatGFA.load(env.name(i.inputs[0]))
    elif i.op == "parse-bed":
        data, stream = env.read(i.inputs[0])
        env.beds[i.output.index] = binding.FlatBED.parse(_terminated(data) if stream else data).entries()
    elif i.op == "node-depth":
        g = env.graph(i.inputs[0])
        d, u = g.seg_depth_with_uniq()
        env.write(i.output, g.format_seg_depth(d, u))
    elif i.op == "path-depth":
        g = env.graph(i.inputs[0])
        ids = None if i.arg is None else [_find_path(g, i.arg)]
        lengths, means = g.path_depth(ids)
        if i.output.kind == BED_STORE:                 # PathDepth::as_bed, depth.rs:173-184
            order = range(g.path_count) if ids is None else ids
            env.beds[i.output.index] = [(g.path_name(p), 0, int(ln)) for p, ln in zip(order, lengths)]
        else:
            env.write(i.output, g.format_path_depth(lengths, means, ids))
    elif i.op == "path-length":
        g = env.graph(i.inputs[0])
        lengths, _ = g.path_depth([_find_path(g, i.arg)])
        env.beds[i.output.index] = [(i.arg.encode(), 0, int(lengths[0]))]
    elif i.op == "make-windows":
        windows = []
        for name, start, end in env.beds.pop(i.inputs[0].index):
            windows += binding.FlatBED.windows(name, start, end, i.arg).entries()   # Windows, window_depth.rs:20-52
        if i.output.kind == BED_STORE:
            env.beds[i.output.index] = windows
        else:
            env.write(i.output, _bed_text(windows))
    elif i.op == "interval-depth":
        g = env.graph(i.inputs[0])
        entries = env.beds.pop(i.inputs[1].index)
        env.write(i.output, b"#path\tstart\tend\tmean.depth\n" + g.bed_depth(_bed_text(entries)))   # eval/instr.rs:212-217
    else:
        raise ValueError(f"unknown instruction {i.op}")


def run(prog: Program, stdin: Optional[bytes] = None, stdout=None) -> None:   # eval/mod.rs:225-230
    env = _Env(prog, stdin, stdout if stdout is not None else sys.stdout.buffer)
    for i in prog.instrs:
        _eval(env, i)


def run_shell(line: str, pretend: bool = False, optimize_ir: bool = False, stdin: Optional[bytes] = None, stdout=None) -> Optional[str]:
    """main.rs:11-20.  In pretend mode returns the printed program instead of running it."""
    prog = sh_to_ir(parse_sh(line))
    if optimize_ir:
        prog = optimize(prog)
    if pretend:
        return str(prog)
    run(prog, stdin, stdout)
    return None


def main(argv: Optional[List[str]] = None) -> int:    # main.rs:38-54
    args = list(sys.argv[1:] if argv is None else argv)
    pretend = optimize_ir = False
    cmd = script = None
    k = 0
    while k < len(args):
        a = args[k]
        if a in ("-p", "--pretend"):
            pretend = True
        elif a in ("-O", "--optimize"):
            optimize_ir = True
        elif a == "-c":
            k += 1
            if k >= len(args):
                print("flash: -c needs a command", file=sys.stderr)
                return 2
            cmd = args[k]
        elif script is None:
            script = a
        k += 1
    if cmd is None and script is not None:
        with open(script, encoding="utf-8") as f:
            cmd = f.read()
    def once(text: str) -> int:
        try:
            out = run_shell(text, pretend, optimize_ir)
        except (Unsupported, ValueError, OSError, RuntimeError) as exc:   # the reference panics here
            print(f"flash: {type(exc).__name__}: {exc}", file=sys.stderr)
            return 1
        if out is not None:
            sys.stdout.write(out)
        return 0

    if cmd is not None:
        return once(cmd)
    while True:                                        # the interactive prompt, main.rs:22-36
        try:
            line = input("$ ")
        except (EOFError, KeyboardInterrupt):
            return 0
        once(line)


if __name__ == "__main__":
    sys.exit(main())
