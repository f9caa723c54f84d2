"""-m gpu, needs >= 2 GPUs (skipped on a one-GPU box; run with `gpurun --gpus 2`): the peer-memory gradient exchange under
torchrun -- `multimem.ld_reduce/st` (NVLS), peer loads and the NCCL fallback -- against the dense all-reduce of the ranks'
flat gradient buffers, bit-identical across ranks (scripts/check_exchange_multi.py)."""
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _ngpu():
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("path", ["auto", "pull", "nccl"])     # auto at 4 ranks: NVLS, tails pushed by multicast stores; pull: by peer loads
@pytest.mark.parametrize("world", [2, 4])
def test_exchange_under_torchrun(world, path):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs, this box has {_ngpu()}")
    env = dict(os.environ)
    env.pop("SPV_EXCHANGE", None); env.pop("SPV_EXCHANGE_PULL", None)
    if path == "nccl":
        env["SPV_EXCHANGE"] = "nccl"
    elif path == "pull":
        if world < 4:
            pytest.skip("two ranks already exchange by peer loads (one-shot mode)")
        env["SPV_EXCHANGE_PULL"] = "1"
    port = 29600 + world * 3 + {"auto": 0, "nccl": 1, "pull": 2}[path]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                        "--master-port", str(port), os.path.join(ROOT, "scripts", "check_exchange_multi.py")],
                       capture_output=True, text=True, timeout=420, env=env, cwd=ROOT)
    assert r.returncode == 0, (r.stdout + r.stderr)[-3000:]
    assert "exchange check ok" in r.stdout, (r.stdout + r.stderr)[-3000:]
    if path == "nccl":
        assert "path=nccl" in r.stdout
