"""treegp_b200 -- B200-native (sm_100a) implementation of treegp's Gaussian-process hot path.

Public names mirror /root/reference/treegp/__init__.py:23-36.
"""
from .kernels import AnisotropicRBF, AnisotropicVonKarman, VonKarman, eval_kernel

__version__ = "0.1.0"

__all__ = ["AnisotropicRBF", "VonKarman", "AnisotropicVonKarman", "eval_kernel"]
