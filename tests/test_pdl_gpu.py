"""Programmatic dependent launch (rx_kernels.cu: pdl_enter / pdl_launch) must not change a pixel: the same frames with the
attribute on (default) and off (RXC_PDL=0), each in its own process."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "pdl_worker.py")


def _digest(pdl):
    env = dict(os.environ, RXC_PDL=str(pdl))
    r = subprocess.run([sys.executable, WORKER], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("PDL_WORKER_OK")]
    assert line, r.stdout[-3000:] + r.stderr[-3000:]
    return line[-1].split()[1]


@pytest.mark.gpu
def test_frames_identical_with_and_without_dependent_launch():
    assert _digest(1) == _digest(0)
