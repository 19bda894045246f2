"""CPU, build container only: run the REFERENCE's own test files, unchanged, against the host mirror.

`import treegp` resolves to treegp_b200 (tests/refsuite/shims), the third-party packages that are absent from
this image are stood in for (fitsio: imported only; matplotlib: headless no-op), and -- because there is no GPU
here and the product has no CPU path -- the device operators are replaced by the oracle
(tests/refsuite/cpu_backend.py).  What this pins is the drop-in API and all host-side logic: constructor
signatures and defaults, attribute names the tests read, the optimisers (L-BFGS-B loop, chi-square fits, the
MIGRAD stand-in with its restart grid), bootstrap bookkeeping, meanify, E/B utilities, plot_fitted_kernel.
The CUDA arithmetic behind the same operators is pinned by the `-m gpu` parity tests.

Skipped where /root/reference does not exist (the GPU box).  The test files are copied to a temporary
directory at run time (the reference tree is read-only and the tests write under ./outputs); nothing of the
reference is stored in this repository.
"""
import os
import shutil
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_TESTS = "/root/reference/tests"

CONFTEST = '''
import os, sys
sys.path.insert(0, %r)
sys.path.insert(0, os.path.join(%r, "tests"))
from refsuite import cpu_backend
cpu_backend.install()
'''


@pytest.mark.skipif(not os.path.isdir(REF_TESTS), reason="reference tree not available on this machine")
def test_reference_test_suite_passes_unchanged(tmp_path):
    work = tmp_path / "reference_tests"
    shutil.copytree(REF_TESTS, work)
    os.makedirs(work / "outputs", exist_ok=True)
    (work / "conftest.py").write_text(CONFTEST % (ROOT, ROOT))
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([os.path.join(ROOT, "tests", "refsuite", "shims"), ROOT])
    env["OMP_NUM_THREADS"] = str(os.cpu_count())
    proc = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider", "."], cwd=str(work),
                          env=env, capture_output=True, text=True, timeout=1500)
    tail = (proc.stdout + proc.stderr)[-3000:]
    assert proc.returncode == 0, tail
    assert " passed" in proc.stdout and "failed" not in proc.stdout, tail
