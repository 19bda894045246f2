"""`import treegp` -> the B200 host mirror, so that the reference's own test files run unchanged."""
from treegp_b200 import *  # noqa: F401,F403
from treegp_b200 import __version__, kernels, two_pcf as _tp  # noqa: F401
