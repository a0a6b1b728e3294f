"""First hardware run of the GPU tests that were written after the round's GPU budget was spent.

Those tests are skipped unless V3D_RUN_UNVALIDATED=1.  This file - collected last - runs them that way in CHILD
processes under hard timeouts, so that a device fault in not-yet-seen code cannot poison the CUDA context of the
validated suite; a failing child is reported as an expected failure (xfail, non-strict) with its output saved under
gpurun_out/, a passing one as XPASS.  Once a group has passed on hardware its skip marker is removed and it leaves
this list.  The CTA-pair GEMM test is not run from here (a cluster-barrier bug could hang the device): it stays a
manual `V3D_RUN_UNVALIDATED=1 pytest -k cta_pair` under `timeout`.
"""
import os
import signal
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu

TESTS = Path(__file__).resolve().parent
OUT = TESTS.parent / "gpurun_out"

# group -> (pytest selection, timeout in seconds).  Kept light: the whole file adds a few minutes to the GPU suite; the
# 4-process composite plan and the NCCL variants are run by hand (tools/gpu_round.sh).
GROUPS = {
    "gemm_tma_staged_residual": ([str(TESTS / "test_kernels_gpu.py"), "-k", "tma_staged_residual"], 600),
    "gemm_cta_pair": ([str(TESTS / "test_kernels_gpu.py"), "-k", "cta_pair"], 600),
}


@pytest.mark.parametrize("group", list(GROUPS))
@pytest.mark.xfail(os.environ.get("V3D_RUN_UNVALIDATED") != "1", strict=False,
                   reason="first hardware run of code written without GPU access")
def test_first_hardware_run(group):
    env = dict(os.environ, V3D_RUN_UNVALIDATED="1")
    selection, limit = GROUPS[group]
    cmd = [sys.executable, "-m", "pytest", "-q", "-m", "gpu", "-p", "no:cacheprovider", *selection]
    # own session: on a timeout the whole process group goes (pytest child and any workers it spawned), so nothing
    # is left holding the GPU for the steps that follow this suite
    proc = subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True,
                            start_new_session=True)
    try:
        out, _ = proc.communicate(timeout=limit)
        text, rc = out[-8000:], proc.returncode
    except subprocess.TimeoutExpired:
        os.killpg(proc.pid, signal.SIGKILL)
        out, _ = proc.communicate()
        text, rc = f"TIMEOUT after {limit}s\n{(out or '')[-4000:]}", 124
    try:
        OUT.mkdir(exist_ok=True)
        (OUT / f"first_run_{group}.log").write_text(f"exit {rc}\n{text}")
    except OSError:
        pass
    print(text)
    assert rc == 0, f"{group}: exit {rc}\n{text}"
