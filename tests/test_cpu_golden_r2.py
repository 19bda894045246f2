"""CPU: host logic of the mirror against outputs of the UNMODIFIED reference (tests/golden/reference_vectors_r2.npz).

* pure host functions (get_correlation_length_matrix, meanify) are checked in process;
* everything that crosses the C ABI is checked in a SUBPROCESS in which the device operators are stood in for by
  the oracle (tests/refsuite/cpu_backend.py) -- there is no GPU here and the product has no CPU path.  That pins the
  control flow around the operators (optimiser loops, bootstrap stream and bookkeeping, masks, covariance,
  chi-square algebra, E/B combination); the same checks run on the real CUDA path in tests/test_gpu_golden_r2.py."""
import os
import subprocess
import sys

import pytest

import golden_r2_checks as chk

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("name", [n for n, _ in chk.HOST_CHECKS])
def test_host_functions_match_reference_vectors(name):
    dict(chk.HOST_CHECKS)[name](chk.load_golden())


def test_host_logic_around_the_operators_matches_reference_vectors():
    env = dict(os.environ)
    env["OMP_NUM_THREADS"] = str(os.cpu_count())
    proc = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "golden_r2_checks.py")], cwd=ROOT, env=env,
                          capture_output=True, text=True, timeout=900)
    tail = (proc.stdout + proc.stderr)[-3000:]
    assert proc.returncode == 0, tail
    for name, _ in chk.DEVICE_CHECKS + chk.HOST_CHECKS:
        assert "ok " + name in proc.stdout, tail
