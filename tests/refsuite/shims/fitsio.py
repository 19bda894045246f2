"""Stand-in for the fitsio package (imported, but not called, by the reference's tests/test_meanify.py:7);
the package itself reads and writes its FITS tables with treegp_b200.fitstable."""
from treegp_b200.fitstable import read_table, write_table  # noqa: F401


def read(path, ext=1):
    return read_table(path, ext=ext)
