"""-m gpu: tools/sanity_small.py (every hot kernel at small sizes against dense numpy, no torch in the process) as a test."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_small_sweep_of_every_kernel():
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "sanity_small.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "sanity_small ok" in r.stdout
