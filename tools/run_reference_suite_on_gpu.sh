#!/bin/bash
# Run the REFERENCE's own, unmodified test files on a B200 through the real CUDA path.
#
# The GPU box has no /root/reference and nothing of the reference may live in this repository, so the
# test files (python sources only, ~6 KB compressed) travel inside the gpurun command line and are unpacked
# under /tmp on the box.  `import treegp` resolves to treegp_b200 via tests/refsuite/shims (which also stands
# in for the fitsio / matplotlib packages that are absent from the image).  tests/inputs/mean_gp_stat_mean.fits
# is regenerated on the box by the reference's own test_meanify (it is a stored copy of that test's output).
# Output: gpurun_out/reference_suite_on_b200.log
set -e
PACK=$(mktemp -d)
cp /root/reference/tests/*.py "$PACK"/
B64=$(tar czf - -C "$PACK" . | base64 -w0)
/usr/local/graft/bin/gpurun --timeout 1200 -- "mkdir -p /tmp/reft/inputs /tmp/reft/outputs && echo $B64 | base64 -d | tar xz -C /tmp/reft && cd /tmp/reft && export PYTHONPATH=\$GRAFT_REPO_ROOT/tests/refsuite/shims:\$GRAFT_REPO_ROOT && python -m pytest -x -q -p no:cacheprovider test_meanify.py::test_meanify 2>&1 | tail -3 && cp outputs/mean_gp_stat_mean.fits inputs/ && python -m pytest -x -q -p no:cacheprovider --durations=12 . 2>&1 | tail -25 > \$GRAFT_REPO_ROOT/gpurun_out/reference_suite_on_b200.log; cat \$GRAFT_REPO_ROOT/gpurun_out/reference_suite_on_b200.log"
