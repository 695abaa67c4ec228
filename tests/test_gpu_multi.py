"""Multi-GPU parity of slab mode (BASELINE configs[4]): needs >= 2 CUDA devices, skipped otherwise.
One process per GPU under torch.distributed.run; the worker (tests/slab_multi_worker.py) compares the
N-rank dataflow kernel and the barrier-per-sweep kernel with the plain-C oracle (1e-10, identical sweep
counts) and, bitwise, with the 1-GPU cooperative-grid kernels."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _n_devices():
    try:
        import torch
        return torch.cuda.device_count() if torch.cuda.is_available() else 0
    except Exception:
        return 0


@pytest.mark.timeout(600)
@pytest.mark.parametrize("ranks", [2, 4, 8])
def test_slab_mode_on_several_gpus_against_the_c_oracle(ranks):
    if _n_devices() < ranks:
        pytest.skip("needs %d CUDA devices, %d visible" % (ranks, _n_devices()))
    env = dict(os.environ)
    env.pop("IRLB200_SLAB_FLOW", None)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(ranks),
           "--master-addr", "127.0.0.1", "--master-port", str(29600 + ranks),
           os.path.join(ROOT, "tests", "slab_multi_worker.py")]
    res = subprocess.run(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=540)
    assert res.returncode == 0 and "SLAB_MULTI_OK" in res.stdout, res.stdout[-4000:]
