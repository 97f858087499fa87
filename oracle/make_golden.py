#!/usr/bin/env python3
"""Generate tests/golden/: golden input/output vectors for the node-depth path.

Run in the BUILD container, where the reference checkout is mounted at /root/reference:
it imports the reference's own Python implementation (`slow_odgi depth`,
slow_odgi/slow_odgi/depth.py:6-16 over mygfa/mygfa/preprocess.py:5-18) unmodified and
records its output for
  * the reference's in-repo fixtures for this path (tests/depth/basic/ex1.gfa, ex2.gfa,
    tests/subset-paths/ex{1,2}.paths, flatgfa-py/test/tiny.gfa, the worked example in
    slow_odgi/README.md:147-175), and
  * seeded random graphs (sequential and non-sequential segment names, reverse steps,
    revisits, empty and unused segments, path subsets drawn the way the reference's
    depth_setup does, slow_odgi/slow_odgi/somepaths.py:10-15).
The vectors travel with the repo; the GPU box has no reference checkout.  This is test
infrastructure (see oracle/depth_oracle.c header).

usage: python oracle/make_golden.py [--ref /root/reference] [--out tests/golden]
"""
import argparse
import contextlib
import io
import json
import os
import random
import re
import sys


def slow_odgi_depth(ref, gfa_text, paths=None):
    sys.path[:0] = [os.path.join(ref, "mygfa"), os.path.join(ref, "slow_odgi")]
    import mygfa  # noqa: E402
    from slow_odgi import depth as so_depth  # noqa: E402

    graph = mygfa.Graph.parse(io.StringIO(gfa_text))
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        so_depth.depth(graph, paths)
    return buf.getvalue()


def readme_graph(ref):
    """The worked example of slow_odgi/README.md (depth section).  Its `P y`/`P z`
    lines carry double tabs, which mygfa accepts and the tab-exact FlatGFA parser
    rejects (gfaline.rs:88-100), so tabs are normalised to single ones."""
    lines = open(os.path.join(ref, "slow_odgi", "README.md"), encoding="utf-8").read().split("\n")
    start = next(i for i, l in enumerate(lines) if l.strip() == "#### `depth`")
    fence = [i for i in range(start, len(lines)) if lines[i].startswith("```")]
    body = lines[fence[0] + 1:fence[1]]
    subset = [l for l in lines[fence[2] + 1:fence[3]] if l.strip()]
    expected = [l for l in lines[fence[4] + 1:fence[5]] if l.strip()]
    gfa = "\n".join(re.sub(r"\t+", "\t", l) for l in body) + "\n"
    return gfa, subset, expected


def random_gfa(rng, sequential):
    n_segs = rng.randint(1, 40)
    if sequential:
        names = list(range(1, n_segs + 1))
    else:
        names = rng.sample(range(1, 5000), n_segs)
    out = ["H\tVN:Z:1.0"]
    for nm in names:
        seq = "".join(rng.choice("ACGT") for _ in range(rng.randint(1, 6)))
        out.append(f"S\t{nm}\t{seq}")
    for _ in range(rng.randint(0, n_segs)):
        a, b = rng.choice(names), rng.choice(names)
        out.append(f"L\t{a}\t{rng.choice('+-')}\t{b}\t{rng.choice('+-')}\t{rng.randint(0, 9)}M")
    n_paths = rng.randint(0, 7)
    used = names[: max(1, (3 * n_segs) // 4)]          # leave some segments untouched
    for p in range(n_paths):
        n_steps = rng.choice([1, 2, 3, 5, 17, 64, 130])
        steps, cur = [], rng.randrange(len(used))
        for _ in range(n_steps):
            steps.append(f"{used[cur]}{rng.choice('++++-')}")
            r = rng.random()
            if r < 0.6:
                cur = (cur + 1) % len(used)
            elif r < 0.8:
                pass                                     # self loop: in-path revisit
            else:
                cur = rng.randrange(len(used))
        out.append(f"P\tpath{p}\t{','.join(steps)}\t*")
    return "\n".join(out) + "\n"


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref", default="/root/reference")
    ap.add_argument("--out", default=os.path.join(os.path.dirname(__file__), "..", "tests", "golden"))
    a = ap.parse_args()
    os.makedirs(a.out, exist_ok=True)
    cases = []

    def add(name, gfa_text, paths=None, source=""):
        gfa_file = f"{name}.gfa"
        with open(os.path.join(a.out, gfa_file), "w", encoding="utf-8") as f:
            f.write(gfa_text)
        all_out = slow_odgi_depth(a.ref, gfa_text, None)
        with open(os.path.join(a.out, f"{name}.depth"), "w", encoding="utf-8") as f:
            f.write(all_out)
        case = {"name": name, "gfa": gfa_file, "depth": f"{name}.depth", "source": source}
        if paths is not None:
            sub_out = slow_odgi_depth(a.ref, gfa_text, paths)
            with open(os.path.join(a.out, f"{name}.subset.depth"), "w", encoding="utf-8") as f:
                f.write(sub_out)
            case["subset_paths"] = paths
            case["subset_depth"] = f"{name}.subset.depth"
        cases.append(case)

    def rd(rel):
        return open(os.path.join(a.ref, rel), encoding="utf-8").read()

    add("ref_ex1", rd("tests/depth/basic/ex1.gfa"), rd("tests/subset-paths/ex1.paths").split(),
        "tests/depth/basic/ex1.gfa + tests/subset-paths/ex1.paths")
    add("ref_ex2", rd("tests/depth/basic/ex2.gfa"), rd("tests/subset-paths/ex2.paths").split(),
        "tests/depth/basic/ex2.gfa + tests/subset-paths/ex2.paths")
    add("ref_tiny", rd("flatgfa-py/test/tiny.gfa"), None, "flatgfa-py/test/tiny.gfa")
    gfa, subset, expected = readme_graph(a.ref)
    add("ref_readme", gfa, subset, "slow_odgi/README.md:147-175")
    # the README prints its expected table: pin slow_odgi's output to the document itself
    got = slow_odgi_depth(a.ref, gfa, subset).split("\n")[1:]
    assert [l for l in got if l] == expected, (got, expected)

    rng = random.Random(20261017)
    for i in range(24):
        text = random_gfa(rng, sequential=(i % 3 != 0))
        names = [l.split("\t")[1] for l in text.split("\n") if l.startswith("P\t")]
        subset = None
        if names:
            random.seed(4)                                # somepaths.py:12
            subset = random.sample(names, int(0.5 * len(names)))
        add(f"rand_{i:02d}", text, subset, "seeded random graph (oracle/make_golden.py)")

    with open(os.path.join(a.out, "manifest.json"), "w", encoding="utf-8") as f:
        json.dump({"generator": "oracle/make_golden.py", "reference_impl": "slow_odgi depth (slow_odgi/slow_odgi/depth.py:6-16)",
                   "cases": cases}, f, indent=1)
    print(f"wrote {len(cases)} cases to {a.out}")


if __name__ == "__main__":
    main()
