"""Mean-function ("meanify") support for GPInterpolation.

SURVEY.md section 8f-1 ranks this path NEXT after the GP hot path; this round provides the consumer side
that GPInterpolation needs (gp_interp.py:97-107, :229-243): reading the spatial-average table and the
k-nearest-neighbour lookup.  The neighbour search itself is not on the O(N^2)/O(N^3) path (a few
thousand grid points) and is still sklearn's KD-tree on the host, exactly as in the reference.
"""
import numpy as np

from . import fitstable


def read_average(path):
    """COORDS0 (n, 2) and PARAMS0 (n,) of the 'average_solution' table (gp_interp.py:100-102)."""
    tab = fitstable.read_table(path, ext=1)
    return tab["COORDS0"][0], tab["PARAMS0"][0]


def knn_average(X0, y0, X, n_neighbors):
    """Uniform mean of the `n_neighbors` nearest mean-grid values (gp_interp.py:236-238)."""
    from sklearn.neighbors import KNeighborsRegressor

    neigh = KNeighborsRegressor(n_neighbors=n_neighbors)
    neigh.fit(X0, y0)
    return neigh.predict(X)
