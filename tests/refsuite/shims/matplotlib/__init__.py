"""Headless no-op stand-in for matplotlib (absent from this image): enough for
GPInterpolation.plot_fitted_kernel, which the reference's test only requires to run (test_hyp_search.py:128-132)."""


def use(_backend):
    return None
