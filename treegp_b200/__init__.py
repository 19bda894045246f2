"""treegp_b200 -- B200-native (sm_100a) implementation of treegp's Gaussian-process hot path.

Public names mirror /root/reference/treegp/__init__.py:23-36: ``GPInterpolation``, ``two_pcf`` and
``log_likelihood`` (the classes, as in the reference), the three treegp kernels and ``eval_kernel``.
All O(N^2)/O(N^3) arithmetic runs in hand-written CUDA behind the C ABI of include/treegp_b200.h;
there is no CPU fallback.
"""
from .kernels import AnisotropicRBF, AnisotropicVonKarman, VonKarman, eval_kernel
from .log_likelihood import log_likelihood
from .two_pcf import two_pcf
from .gp_interp import GPInterpolation
from .meanify import meanify
from .utils import comp_eb, comp_eb_treecorr

__version__ = "0.1.0"

__all__ = [
    "__version__",
    "GPInterpolation",
    "two_pcf",
    "log_likelihood",
    "AnisotropicRBF",
    "VonKarman",
    "AnisotropicVonKarman",
    "eval_kernel",
    "meanify",
    "comp_eb",
    "comp_eb_treecorr",
]
