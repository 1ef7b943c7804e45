"""Multi-rank parity (needs at least two GPUs on the box; skipped otherwise): the decomposed runs of tools/multi_gpu_check.py
(LJ, 12 steps with migration, peer-memory halo refresh and its NCCL fallback) and tools/multi_gpu_sph.py --check (SPH
density -> halo column refresh -> hydro force) against the single-GPU run of the whole periodic system."""
import ctypes
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _num_gpus():
    try:
        n = ctypes.c_int(0)
        return n.value if ctypes.CDLL("libcudart.so").cudaGetDeviceCount(ctypes.byref(n)) == 0 else 0
    except OSError:
        return 0


def _torchrun(script, ranks, port, extra_args=(), env=None):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={ranks}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", script), *extra_args]
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=e, cwd=ROOT)


@pytest.mark.skipif(_num_gpus() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("no_p2p", ["0", "1"])
def test_two_rank_lj_run_matches_single_gpu(no_p2p):
    """APB_NO_P2P_HALO=1 keeps the NCCL send / recv refresh; the default is the peer-memory path (CUDA IPC arenas)."""
    r = _torchrun("multi_gpu_check.py", 2, 29531 + int(no_p2p), env={"APB_NO_P2P_HALO": no_p2p} if no_p2p == "1" else None)
    assert r.returncode == 0 and "MULTI_GPU_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]


@pytest.mark.skipif(_num_gpus() < 2, reason="needs two GPUs")
def test_two_rank_sph_pass_matches_single_gpu():
    r = _torchrun("multi_gpu_sph.py", 2, 29533, extra_args=("--check",))
    assert r.returncode == 0 and "MULTI_GPU_SPH_CHECK PASS" in r.stdout, r.stdout[-2000:] + r.stderr[-3000:]
