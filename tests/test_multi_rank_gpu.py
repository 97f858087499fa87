"""Oracle parity at N > 1 on hardware: two torchrun ranks (one process per GPU, NCCL), both
exchanges -- the NCCL all-reduce of [depth | uniq] and kernel X over symmetric memory -- and both
engines, each compared with the C oracle over the whole graph.  Skipped on a box with one GPU
(the round's 2/4/8-GPU outputs of the same tool are under profiles/)."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    import torch

    return torch.cuda.device_count()


@pytest.mark.parametrize("engine", ["stream", "window"])
def test_two_ranks_match_the_oracle(engine):
    if _device_count() < 2:
        pytest.skip("needs two GPUs")
    env = {**os.environ, "FGFA_ENGINE": engine}
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(ROOT, "tools", "check_fused_exchange.py"), "B"],
                       capture_output=True, text=True, cwd=ROOT, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    d = json.loads(line)
    assert d["n_gpus"] == 2 and d["parity_fused_vs_nccl_vs_oracle"] is True
    assert d["engine_nccl_form"] == engine and d["engine_fused_form"] == "stream"
