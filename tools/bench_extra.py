#!/usr/bin/env python3
"""Secondary measurements that bench.py's one-line contract has no room for: every
BASELINE.json config on one GPU (whole step, depth only) and the path-depth mode.
Writes one JSON object per line.   usage: python tools/bench_extra.py [B C E U]"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import torch  # noqa: E402

import oracle_lib as O  # noqa: E402
import pollen_b200 as pb  # noqa: E402
from pollen_b200 import synth  # noqa: E402


def timed(fn, stream, reps=10, warm=3):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps


def main():
    torch.cuda.set_device(0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    for name in (sys.argv[1:] or ["B", "C", "E", "U"]):
        cfg = synth.CONFIGS[name]
        steps, s, e = synth.make_graph(cfg)
        d_steps = torch.from_numpy(steps.view(np.int32)).cuda()
        plan = pb.DepthPlan(s, e, cfg.n_segs, cfg.n_steps)
        out = torch.empty(2 * cfg.n_segs, dtype=torch.int32, device="cuda")
        st = torch.cuda.current_stream()
        full = timed(lambda: plan.run(d_steps, out[: cfg.n_segs], out[cfg.n_segs:], st.cuda_stream), st)
        donly = timed(lambda: plan.run(d_steps, out[: cfg.n_segs], None, st.cuda_stream), st)
        plan.status(st.cuda_stream)
        alg = 4.0 * cfg.n_steps + 8.0 * cfg.n_paths + 8.0 * cfg.n_segs
        # path-depth mode: depth-only run + kernel C
        seg_len = np.random.default_rng(0).integers(1, 200, cfg.n_segs).astype(np.uint32)
        d_len = torch.from_numpy(seg_len.view(np.int32)).cuda()
        scratch = torch.empty(2 * cfg.n_segs, dtype=torch.int32, device="cuda")
        sums = torch.empty(2 * cfg.n_paths, dtype=torch.int64, device="cuda")
        lib = pb.lib()

        def path_mode():
            plan.run(d_steps, out[: cfg.n_segs], None, st.cuda_stream)
            rc = lib.fgfa_depth_plan_path_sums(plan._h, d_steps.data_ptr(), out.data_ptr(), d_len.data_ptr(),
                                               scratch.data_ptr(), sums.data_ptr(), st.cuda_stream)
            assert rc == 0
        pmode = timed(path_mode, st)
        plan.status(st.cuda_stream)
        t0 = time.perf_counter()
        rc, ol, om = O.path_depth(steps, s, e, seg_len)
        cpu_path = time.perf_counter() - t0
        got = sums.cpu().numpy().view(np.uint64)
        assert rc == 0 and (got[1::2] == ol).all()
        # window-depth mode (window_depth.rs:183-197) along path 0, 1000 bp windows: depth-only run + W1..W3
        wpath, wsize = 0, 1000
        n_path = int(e[wpath] - s[wpath])
        total = O.path_length(steps, s, e, seg_len, wpath)
        n_win = (total + wsize - 1) // wsize
        seg_end = torch.empty(max(n_path, 1), dtype=torch.int64, device="cuda")
        wins = torch.empty(3 * max(n_win, 1), dtype=torch.int64, device="cuda")
        d_win = torch.empty(max(n_win, 1), dtype=torch.float64, device="cuda")
        scr_bytes = lib.fgfa_interval_scratch_bytes(n_path, n_win)
        scr = torch.empty(scr_bytes, dtype=torch.uint8, device="cuda")
        d_path = d_steps.data_ptr() + 4 * int(s[wpath])
        ws_ptr, we_ptr = wins.data_ptr(), wins.data_ptr() + 8 * max(n_win, 1)

        def interval_kernels():
            rc = lib.fgfa_path_offsets_device(d_path, n_path, d_len.data_ptr(), cfg.n_segs, seg_end.data_ptr(),
                                              scr.data_ptr(), scr_bytes, st.cuda_stream)
            rc = rc or lib.fgfa_make_windows_device(0, total, wsize, n_win, ws_ptr, we_ptr, st.cuda_stream)
            rc = rc or lib.fgfa_interval_depth_device(d_path, n_path, out.data_ptr(), d_len.data_ptr(), cfg.n_segs,
                                                      seg_end.data_ptr(), ws_ptr, we_ptr, n_win, d_win.data_ptr(),
                                                      scr.data_ptr(), scr_bytes, st.cuda_stream)
            assert rc == 0
        plan.run(d_steps, out[: cfg.n_segs], None, st.cuda_stream)
        wkern = timed(interval_kernels, st)
        assert lib.fgfa_interval_status(scr.data_ptr(), st.cuda_stream) == 0
        ows, owe = O.windows(0, total, wsize)
        t0 = time.perf_counter()
        rc, owd = O.interval_depth(steps, s, e, seg_len, wpath, ows, owe)
        cpu_window = time.perf_counter() - t0
        assert rc == 0 and d_win.cpu().numpy()[:n_win].tobytes() == owd.tobytes()
        t0 = time.perf_counter()
        rc, od, ou = O.depth_with_uniq(steps, s, e, cfg.n_segs)
        cpu_node = time.perf_counter() - t0
        g = out.cpu().numpy().view(np.uint32)
        plan.run(d_steps, out[: cfg.n_segs], out[cfg.n_segs:], st.cuda_stream)
        plan.status(st.cuda_stream)
        g = out.cpu().numpy().view(np.uint32)
        assert (g[: cfg.n_segs] == od).all() and (g[cfg.n_segs:] == ou).all()
        print(json.dumps({
            "config": name, "n_segs": cfg.n_segs, "n_paths": cfg.n_paths, "n_steps": cfg.n_steps,
            "node_depth_ms": full, "node_depth_steps_per_s": cfg.n_steps / (full * 1e-3),
            "node_depth_frac_of_measured_hbm": alg / (full * 1e-3) / 1e9 / peak,
            "depth_only_ms": donly, "depth_only_frac_of_measured_hbm": (alg - 4.0 * cfg.n_segs) / (donly * 1e-3) / 1e9 / peak,
            "path_depth_ms": pmode, "path_measure_kernel_ms": pmode - donly,
            "window_depth_ms": donly + wkern, "interval_kernels_ms": wkern, "window_path_steps": n_path, "windows": int(n_win),
            "cpu_window_depth_s_1core": cpu_window,
            "cpu_node_depth_s_1core": cpu_node, "cpu_path_depth_s_1core": cpu_path, "parity": "bit-exact"}))
        plan.close()
        del d_steps, out, scratch


if __name__ == "__main__":
    main()
