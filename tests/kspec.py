"""Parse the kernel strings stored in the golden file into oracle arguments (family, amp, metric)
without importing the reference or the product: a tiny stand-in namespace for eval()."""
import numpy as np


class Spec:
    def __init__(self, family, invLam=None, length_scale=None, amp=1.0):
        self.family, self.invLam, self.length_scale, self.amp = family, invLam, length_scale, amp

    def __rmul__(self, c):
        return Spec(self.family, self.invLam, self.length_scale, self.amp * float(c))

    __mul__ = __rmul__

    def oracle_args(self):
        d = dict(family=self.family, amp=self.amp)
        if self.invLam is not None:
            d["invLam"] = self.invLam
        else:
            d["length_scale"] = self.length_scale
        return d


def _matern(length_scale=1.0, nu=1.5):
    return Spec({0.5: "matern12", 1.5: "matern32", 2.5: "matern52"}[nu], length_scale=length_scale)


NAMESPACE = {
    "array": np.array,
    "AnisotropicRBF": lambda invLam: Spec("rbf", invLam=np.asarray(invLam, dtype=float)),
    "AnisotropicVonKarman": lambda invLam: Spec("vonkarman", invLam=np.asarray(invLam, dtype=float)),
    "RBF": lambda length_scale=1.0: Spec("rbf", length_scale=length_scale),
    "VonKarman": lambda length_scale=1.0: Spec("vonkarman", length_scale=length_scale),
    "Matern": _matern,
}


def parse(kernel_string):
    return eval(str(kernel_string), dict(NAMESPACE))


def vk_sym_quirk(K, X, family):
    """Reference K(X,X) for von Karman leaves coincident OFF-diagonal pairs at 0 (kernels.py:253-262,
    :360-367) -- the oracle applies the same rule."""
    if family != "vonkarman":
        return K
    K = K.copy()
    X = np.atleast_2d(X)
    same = (np.abs(X[:, None, :] - X[None, :, :]).sum(-1) == 0) & ~np.eye(len(X), dtype=bool)
    K[same] = 0.0
    return K
