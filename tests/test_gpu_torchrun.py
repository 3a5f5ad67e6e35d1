"""GPU, >= 2 devices: the REAL transports of the domain decomposition (CUDA-IPC peer memory, ncclSend/ncclRecv) under torchrun.
Collects tests/run_nccl_check.py (fp64 runs with 1 and 2 force evaluations per step, a quench + gravity schedule, and an fp32
tile-list pass with bit-exact lists) into pytest; skipped on single-GPU boxes, where tests/test_gpu_multirank.py covers the same
kernels through the in-process transport.  Logs of the 2- / 4- / 8-GPU runs of r02 are kept under profiles/r02/."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _device_count():
    try:
        import torch

        return torch.cuda.device_count()
    except Exception:
        return 0


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_real_transports_under_torchrun(transport):
    n = _device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    if os.environ.get("MDSEA_TEST_TORCHRUN", "1") == "0":
        pytest.skip("switched off (MDSEA_TEST_TORCHRUN=0)")
    world = 8 if n >= 8 else (4 if n >= 4 else 2)
    env = dict(os.environ)
    env.pop("MDSEA_HALO", None)
    if transport == "nccl":
        env["MDSEA_HALO"] = "nccl"
    port = 29600 + (os.getpid() % 300) + (0 if transport == "p2p" else 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "run_nccl_check.py")]
    out = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=600)
    ok = [ln for ln in out.stdout.splitlines() if "check OK" in ln]
    assert out.returncode == 0 and len(ok) == 4, out.stdout[-2000:] + out.stderr[-4000:]
    assert f"transport={1 if transport == 'p2p' else 2}" in ok[-1], ok[-1]
