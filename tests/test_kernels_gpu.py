"""GPU: every C-ABI kernel against plain torch on identical bf16-rounded operands
(tests/kernel_checks.py; one subprocess per check so a trap is contained and reported)."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import kernel_checks  # noqa: E402


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(kernel_checks.CHECKS))
def test_kernel(name):
    r = subprocess.run([sys.executable, os.path.join(HERE, "kernel_checks.py"), name], capture_output=True,
                       text=True, timeout=300)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, f"{name} failed:\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}"
