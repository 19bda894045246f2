"""Minimal FITS binary-table I/O for the mean-function files.

The reference reads/writes its spatial-average tables with fitsio (cfitsio), a third-party native
library that is not part of this image (/root/reference/treegp/gp_interp.py:97-102,
meanify.py:139-165).  The files are one-row BINTABLEs of big-endian float64 vectors with a TDIMn
shape, which is all this module understands -- enough to read files written by the reference
(tests/inputs/mean_gp_stat_mean.fits) and to write files the reference (fitsio/astropy) can read.
"""
import re

import numpy as np

_BLOCK = 2880
_FORMATS = {"D": ">f8", "E": ">f4", "K": ">i8", "J": ">i4", "I": ">i2", "B": "u1", "L": "u1"}


def _cards(block_bytes):
    for i in range(0, len(block_bytes), 80):
        yield block_bytes[i:i + 80].decode("ascii", "replace")


def _read_header(buf, pos):
    header = {}
    while True:
        block = buf[pos:pos + _BLOCK]
        if len(block) < _BLOCK:
            raise ValueError("truncated FITS header")
        pos += _BLOCK
        done = False
        for card in _cards(block):
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] != "= ":
                continue
            val = card[10:].split(" /")[0].strip() if not card[10:].lstrip().startswith("'") else card[10:]
            if val.lstrip().startswith("'"):
                m = re.match(r"\s*'((?:[^']|'')*)'", val)
                header[key] = m.group(1).rstrip() if m else val.strip()
            elif val in ("T", "F"):
                header[key] = val == "T"
            else:
                try:
                    header[key] = int(val)
                except ValueError:
                    try:
                        header[key] = float(val)
                    except ValueError:
                        header[key] = val
        if done:
            return header, pos


def read_table(path, ext=1):
    """Return {column name: array with leading row axis} for BINTABLE extension number `ext`."""
    with open(path, "rb") as fh:
        buf = fh.read()
    pos = 0
    header, pos = _read_header(buf, pos)
    hdu = 0
    while True:
        naxis = header.get("NAXIS", 0)
        size = 0
        if naxis:
            size = abs(header["BITPIX"]) // 8
            for a in range(1, naxis + 1):
                size *= header["NAXIS%d" % a]
            size += header.get("PCOUNT", 0)
        if hdu == ext:
            break
        pos += (size + _BLOCK - 1) // _BLOCK * _BLOCK
        header, pos = _read_header(buf, pos)
        hdu += 1
    if header.get("XTENSION") != "BINTABLE":
        raise ValueError("extension %d of %s is not a BINTABLE" % (ext, path))
    nrows, rowbytes = header["NAXIS2"], header["NAXIS1"]
    out = {}
    off = 0
    for c in range(1, header["TFIELDS"] + 1):
        m = re.match(r"(\d*)([A-Z])", header["TFORM%d" % c])
        rep = int(m.group(1)) if m.group(1) else 1
        dt = np.dtype(_FORMATS[m.group(2)])
        shape = (rep,)
        tdim = header.get("TDIM%d" % c)
        if tdim:
            dims = [int(v) for v in tdim.strip("() ").split(",")]
            shape = tuple(reversed(dims))  # FITS is Fortran-ordered
        col = np.empty((nrows,) + shape, dtype=dt.newbyteorder("="))
        for r in range(nrows):
            start = pos + r * rowbytes + off
            col[r] = np.frombuffer(buf, dtype=dt, count=rep, offset=start).reshape(shape)
        out[header["TTYPE%d" % c]] = col
        off += rep * dt.itemsize
    return out


def _card(key, value, comment=""):
    if isinstance(value, bool):
        v = "%20s" % ("T" if value else "F")
    elif isinstance(value, (int, np.integer)):
        v = "%20d" % value
    else:
        v = "%-20s" % ("'%-8s'" % value)
    s = "%-8s= %s" % (key, v)
    if comment:
        s += " / " + comment
    return s[:80].ljust(80)


def _pad(b, fill):
    return b + fill * ((-len(b)) % _BLOCK)


def write_table(path, columns, extname="average_solution"):
    """Write a one-row BINTABLE; `columns` is an ordered {name: float64 ndarray}."""
    primary = [_card("SIMPLE", True, "file does conform to FITS standard"), _card("BITPIX", 16),
               _card("NAXIS", 0), _card("EXTEND", True), "END".ljust(80)]
    data = b""
    cards = []
    for i, (name, arr) in enumerate(columns.items(), start=1):
        a = np.ascontiguousarray(arr, dtype=np.float64)
        cards.append(_card("TTYPE%d" % i, name))
        cards.append(_card("TFORM%d" % i, "%dD" % a.size))
        if a.ndim > 1:
            cards.append(_card("TDIM%d" % i, "(" + ",".join(str(d) for d in reversed(a.shape)) + ")"))
        data += a.astype(">f8").tobytes()
    head = [_card("XTENSION", "BINTABLE", "binary table extension"), _card("BITPIX", 8), _card("NAXIS", 2),
            _card("NAXIS1", len(data), "width of table in bytes"), _card("NAXIS2", 1), _card("PCOUNT", 0),
            _card("GCOUNT", 1), _card("TFIELDS", len(columns))] + cards + [_card("EXTNAME", extname), "END".ljust(80)]
    with open(path, "wb") as fh:
        fh.write(_pad("".join(primary).encode("ascii"), b" "))
        fh.write(_pad("".join(head).encode("ascii"), b" "))
        fh.write(_pad(data, b"\0"))
