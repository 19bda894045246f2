import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from treegp_b200 import backend
print("plain chain: %.2f us per launch" % backend.microbench_fp64(2, 2000))
print("PDL chain:   %.2f us per launch" % backend.microbench_fp64(3, 2000))
