"""`flash`, the fake shell front end (pollen_b200/flash.py, mirroring
flatgfa-sh).  The pretend-mode (`-p`) transcripts of flatgfa-sh/README.md are the reference's own
golden vectors for the parser, the IR printer and the optimizer (the crate runs them with trycmd,
flatgfa-sh/src/main.rs:56-62); they are restated here verbatim.  Transcripts whose optimisation
depends on files existing (`-O` with note5.flatgfa) are reproduced inside a temp directory."""
import io
import os
import subprocess
import sys

import pytest

from pollen_b200 import flash

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")

# (flags, command, expected stdout) -- flatgfa-sh/README.md:66-102, 112-121, 153-221, 231-253
PRETEND = [
    ("", "odgi depth", "parse-gfa(stdin) -> gfa-store-0\npath-depth(gfa-store-0) -> stdout\n"),
    ("", "odgi depth -d", "parse-gfa(stdin) -> gfa-store-0\nnode-depth(gfa-store-0) -> stdout\n"),
    ("", "odgi depth -i chr8.gfa", 'parse-gfa("chr8.gfa") -> gfa-store-0\npath-depth(gfa-store-0) -> stdout\n'),
    ("", 'odgi depth -i chr8.gfa -r "chm13#chr8"',
     'parse-gfa("chr8.gfa") -> gfa-store-0\npath-depth(gfa-store-0, path="chm13#chr8") -> stdout\n'),
    ("", "odgi depth < chr8.gfa > depth.tsv",
     'parse-gfa("chr8.gfa") -> gfa-store-0\npath-depth(gfa-store-0) -> "depth.tsv"\n'),
    ("", "foo | bar | baz",
     'shell("foo", [], input=stdin) -> pipe-0\nshell("bar", [], input=pipe-0) -> pipe-1\nshell("baz", [], input=pipe-1) -> stdout\n'),
    ("", "foo | bar | baz > qux",
     'shell("foo", [], input=stdin) -> pipe-0\nshell("bar", [], input=pipe-0) -> pipe-1\nshell("baz", [], input=pipe-1) -> "qux"\n'),
    ("", "odgi depth -i chr8.flatgfa", 'map-file("chr8.flatgfa") -> mmap-0\npath-depth(mmap-0) -> stdout\n'),
    ("", "odgi depth -i chr8.og",
     'odgi-view("chr8.og") -> pipe-0\nparse-gfa(pipe-0) -> gfa-store-0\npath-depth(gfa-store-0) -> stdout\n'),
    ("", "bedtools makewindows -b in.bed -w 16 > intermediate.bed ; odgi depth -i g.gfa -b intermediate.bed",
     'parse-bed("in.bed") -> bed-store-0\nmake-windows(bed-store-0, size=16) -> "intermediate.bed"\n'
     'parse-gfa("g.gfa") -> gfa-store-0\nparse-bed("intermediate.bed") -> bed-store-1\n'
     "interval-depth(gfa-store-0, bed-store-1) -> stdout\n"),
    ("-O", "bedtools makewindows -b in.bed -w 16 > intermediate.bed ; odgi depth -i g.gfa -b intermediate.bed",
     'parse-bed("in.bed") -> bed-store-0\nmake-windows(bed-store-0, size=16) -> bed-store-1\n'
     'parse-gfa("g.gfa") -> gfa-store-0\ninterval-depth(gfa-store-0, bed-store-1) -> stdout\n'),
    ("", "odgi depth -i g.flatgfa -r foo ; odgi depth -i g.flatgfa -r bar",
     'map-file("g.flatgfa") -> mmap-0\npath-depth(mmap-0, path="foo") -> stdout\n'
     'map-file("g.flatgfa") -> mmap-1\npath-depth(mmap-1, path="bar") -> stdout\n'),
    ("-O", "odgi depth -i g.flatgfa -r foo ; odgi depth -i g.flatgfa -r bar",
     'map-file("g.flatgfa") -> mmap-0\npath-depth(mmap-0, path="foo") -> stdout\npath-depth(mmap-0, path="bar") -> stdout\n'),
    ("", "odgi depth -i g.gfa -r foo | bedtools makewindows -b /dev/stdin -w 4",
     'parse-gfa("g.gfa") -> gfa-store-0\npath-depth(gfa-store-0, path="foo") -> pipe-0\n'
     "parse-bed(pipe-0) -> bed-store-0\nmake-windows(bed-store-0, size=4) -> stdout\n"),
    ("-O", "odgi depth -i g.gfa -r foo | bedtools makewindows -b /dev/stdin -w 4",
     'parse-gfa("g.gfa") -> gfa-store-0\npath-length(gfa-store-0, path="foo") -> bed-store-0\n'
     "make-windows(bed-store-0, size=4) -> stdout\n"),
    ("", "gunzip < foo.gfa.gz | odgi depth",
     'gzip-decompress("foo.gfa.gz") -> pipe-0\nparse-gfa(pipe-0) -> gfa-store-0\npath-depth(gfa-store-0) -> stdout\n'),
    ("", "odgi depth -i foo.gfa.gz",
     'gzip-decompress("foo.gfa.gz") -> pipe-0\nparse-gfa(pipe-0) -> gfa-store-0\npath-depth(gfa-store-0) -> stdout\n'),
    ("-O", "odgi depth -i foo.gfa.gz", 'parse-gfa(gz "foo.gfa.gz") -> gfa-store-0\npath-depth(gfa-store-0) -> stdout\n'),
    ("-O", "odgi depth -i foo.gfa.gz", 'parse-gfa(gz "foo.gfa.gz") -> gfa-store-0\npath-depth(gfa-store-0) -> stdout\n'),
]


@pytest.mark.parametrize("flags,command,want", PRETEND)
def test_readme_pretend_transcripts(flags, command, want, tmp_path, monkeypatch):
    monkeypatch.chdir(tmp_path)                        # no *.flatgfa / *.gfa files around
    assert flash.run_shell(command, pretend=True, optimize_ir=flags == "-O") == want


def test_readme_transcripts_that_depend_on_existing_files(tmp_path, monkeypatch):
    """flatgfa-sh/README.md:127-151: with -O a text GFA (or an .og file) is swapped for the
    .flatgfa file next to it when that exists."""
    monkeypatch.chdir(tmp_path)
    os.makedirs("tests")
    os.chdir("tests")
    os.makedirs("../flatgfa-sh")
    os.chdir("../flatgfa-sh")
    open("../tests/note5.flatgfa", "wb").close()
    assert flash.run_shell("odgi depth -i ../tests/note5.gfa", True, False) == \
        'parse-gfa("../tests/note5.gfa") -> gfa-store-0\npath-depth(gfa-store-0) -> stdout\n'
    assert flash.run_shell("odgi depth -i ../tests/note5.gfa", True, True) == \
        'map-file("../tests/note5.flatgfa") -> mmap-0\npath-depth(mmap-0) -> stdout\n'
    assert flash.run_shell("odgi depth -i ../tests/note5.og", True, False) == \
        'odgi-view("../tests/note5.og") -> pipe-0\nparse-gfa(pipe-0) -> gfa-store-0\npath-depth(gfa-store-0) -> stdout\n'
    assert flash.run_shell("odgi depth -i ../tests/note5.og", True, True) == \
        'map-file("../tests/note5.flatgfa") -> mmap-0\npath-depth(mmap-0) -> stdout\n'
    os.remove("../tests/note5.flatgfa")                # only the text file is there: opt.rs:79-85
    open("../tests/note5.gfa", "wb").close()
    assert flash.run_shell("odgi depth -i ../tests/note5.og", True, True) == \
        'parse-gfa("../tests/note5.gfa") -> gfa-store-0\npath-depth(gfa-store-0) -> stdout\n'


def test_shell_words_and_syntax_limits():
    p = flash.run_shell("echo 'a b' \"c d\" e\\ f \"g\\\"h\" | cat > out\\ file", True)
    assert p == 'shell("echo", ["a b", "c d", "e f", "g\\"h"], input=stdin) -> pipe-0\nshell("cat", [], input=pipe-0) -> "out file"\n'
    assert flash.run_shell("#!/usr/bin/env flash\nodgi depth -d -i x.gfa   # comment\n\nodgi depth -i x.gfa\n", True) == \
        'parse-gfa("x.gfa") -> gfa-store-0\nnode-depth(gfa-store-0) -> stdout\nparse-gfa("x.gfa") -> gfa-store-1\npath-depth(gfa-store-1) -> stdout\n'
    assert flash.run_shell("odgi depth --input=x.gfa -r p", True) == 'parse-gfa("x.gfa") -> gfa-store-0\npath-depth(gfa-store-0, path="p") -> stdout\n'
    for bad in ("a && b", "a || b", "a &", "echo $HOME", "odgi view -i x.gfa", "bedtools intersect -a x", "gunzip -k < x.gz", "a >> b"):
        with pytest.raises((flash.Unsupported, ValueError)):
            flash.run_shell(bad, True)


def test_passthrough_commands_really_run(tmp_path):
    """flatgfa-sh/README.md:12-25: ordinary commands are executed; pipes and redirections connect them."""
    src = tmp_path / "in.txt"
    src.write_text("The FlatGFA Fake Shell\nsecond line\n")
    out = io.BytesIO()
    flash.run_shell(f"head -n1 < {src} | rev", stdout=out)
    assert out.getvalue() == b"llehS ekaF AFGtalF ehT\n"
    dst = tmp_path / "out.txt"
    flash.run_shell(f"cat < {src} | tail -n1 > {dst}", stdout=io.BytesIO())
    assert dst.read_text() == "second line\n"
    out = io.BytesIO()
    flash.run_shell("tr a-z A-Z", stdin=b"abc\n", stdout=out)
    assert out.getvalue() == b"ABC\n"
    r = subprocess.run([sys.executable, "-m", "pollen_b200.flash", "-p", "-c", "odgi depth -d"], capture_output=True, cwd=ROOT, check=True)
    assert r.stdout == b"parse-gfa(stdin) -> gfa-store-0\nnode-depth(gfa-store-0) -> stdout\n"


@pytest.mark.gpu
def test_gpu_flash_evaluates_depth_pipelines(tmp_path, fgfa_bin):
    """The evaluator against the CLI: node depth, path depth, the windows pipeline of
    flatgfa-sh/windows.sh (plain and -O, which must print the same, README.md:281-296), gzip input,
    and the .flatgfa shortcut."""
    import gzip
    import shutil

    src = str(tmp_path / "g.gfa")
    shutil.copy(os.path.join(GOLD, "ref_ex2.gfa"), src)
    node = subprocess.run([fgfa_bin, "-I", src, "depth", "-d"], capture_output=True, check=True).stdout
    path = subprocess.run([fgfa_bin, "-I", src, "depth"], capture_output=True, check=True).stdout

    def sh(cmd, opt=False, stdin=None):
        out = io.BytesIO()
        flash.run_shell(cmd, optimize_ir=opt, stdin=stdin, stdout=out)
        return out.getvalue()

    assert sh(f"odgi depth -d -i {src}") == node
    assert sh(f"odgi depth -i {src}") == path
    assert sh("odgi depth -d", stdin=open(src, "rb").read()) == node
    assert sh(f"odgi depth -d -i {src} | tail -n1") == node.splitlines(keepends=True)[-1]
    assert sh(f"odgi depth -i {src} -r path1") == path.splitlines(keepends=True)[0] + path.splitlines(keepends=True)[2]
    bed = str(tmp_path / "w4.bed")
    script = f"odgi depth -i {src} -r path0 | bedtools makewindows -b /dev/stdin -w 4 > {bed}\nodgi depth -i {src} -b {bed}\n"
    want = b"#path\tstart\tend\tmean.depth\n" + subprocess.run([fgfa_bin, "-I", src, "window-depth", "path0", "4"], capture_output=True, check=True).stdout
    assert sh(script) == want
    assert open(bed, "rb").read() == b"path0\t0\t4\npath0\t4\t8\npath0\t8\t12\n"
    os.remove(bed)
    assert sh(script, opt=True) == want
    assert not os.path.exists(bed)                     # the intermediate BED file is optimised away
    with open(src + ".gz", "wb") as f:
        f.write(gzip.compress(open(src, "rb").read()))
    assert sh(f"odgi depth -i {src}.gz") == path and sh(f"odgi depth -i {src}.gz", opt=True) == path
    assert sh(f"gunzip < {src}.gz | odgi depth -d") == node
    subprocess.run([fgfa_bin, "-I", src, "-o", str(tmp_path / "g.flatgfa")], check=True)
    assert "map-file" in flash.run_shell(f"odgi depth -i {src}", True, True)
    assert sh(f"odgi depth -d -i {src}", opt=True) == node
