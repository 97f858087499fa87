#!/usr/bin/env python3
"""Times path-depth mode on one config (depth-only pass + kernel C), device-resident, no CPU legs.
usage: [FGFA_MEASURE_BLOCKS=n] python tools/bench_path_mode.py [C]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import pollen_b200 as pb  # noqa: E402
from pollen_b200 import synth  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "C"
    cfg = synth.CONFIGS[name]
    steps, s, e = synth.make_graph(cfg)
    torch.cuda.set_device(0)
    d_steps = torch.from_numpy(steps.view(np.int32)).cuda()
    plan = pb.DepthPlan(s, e, cfg.n_segs, cfg.n_steps)
    out = torch.empty(2 * cfg.n_segs, dtype=torch.int32, device="cuda")
    seg_len = np.random.default_rng(0).integers(1, 200, cfg.n_segs).astype(np.uint32)
    d_len = torch.from_numpy(seg_len.view(np.int32)).cuda()
    scratch = torch.empty(2 * cfg.n_segs, dtype=torch.int32, device="cuda")
    sums = torch.empty(2 * cfg.n_paths, dtype=torch.int64, device="cuda")
    lib = pb.lib()
    st = torch.cuda.current_stream()
    plan.run(d_steps, out[: cfg.n_segs], None, st.cuda_stream)

    def kernel_c():
        rc = lib.fgfa_depth_plan_path_sums(plan._h, d_steps.data_ptr(), out.data_ptr(), d_len.data_ptr(),
                                           scratch.data_ptr(), sums.data_ptr(), st.cuda_stream)
        assert rc == 0
    for _ in range(3):
        kernel_c()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(st)
    for _ in range(20):
        kernel_c()
    b.record(st)
    torch.cuda.synchronize()
    plan.status(st.cuda_stream)
    got = sums.cpu().numpy().view(np.uint64)
    segs = steps >> 1
    p = cfg.n_paths - 1
    assert int(got[2 * p + 1]) == int(seg_len[segs[s[p]:e[p]]].astype(np.uint64).sum())
    print(json.dumps({"config": name, "measure_blocks": os.environ.get("FGFA_MEASURE_BLOCKS", "default"),
                      "path_measure_ms": a.elapsed_time(b) / 20}))


if __name__ == "__main__":
    main()
