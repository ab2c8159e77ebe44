"""Column-oriented FITS tables (one-row BINTABLE, every column a vector with TDIM).

This is the on-disk format of every table egg reads or writes: the filter
curves (``LAM``/``RES``), the SED libraries and the output catalog
(reference: doc/doc-egg-generic.tex:67-69; columns written at
src/egg-gencat.cpp:2253-2286, libraries read at src/egg-gencat.cpp:401-458).
Python side of the format, used by the tests, the data packer and bench.py;
the C++ twin used by the product lives in ``csrc/fitscol.cpp``.
"""
from __future__ import annotations

import re
from collections import OrderedDict

import numpy as np

_BLOCK = 2880

# FITS TFORM letter -> numpy big-endian dtype
_TFORM = {
    "L": np.dtype("S1"), "B": np.dtype("u1"), "I": np.dtype(">i2"), "J": np.dtype(">i4"),
    "K": np.dtype(">i8"), "E": np.dtype(">f4"), "D": np.dtype(">f8"), "A": np.dtype("S1"),
}


def _parse_header(buf: bytes, off: int):
    """Parse one header unit starting at ``off``; returns (cards, offset after header)."""
    cards = OrderedDict()
    while True:
        block = buf[off:off + _BLOCK]
        if len(block) < _BLOCK:
            raise ValueError("truncated FITS header")
        off += _BLOCK
        done = False
        for i in range(0, _BLOCK, 80):
            card = block[i:i + 80].decode("ascii", "replace")
            key = card[:8].strip()
            if key == "END":
                done = True
                break
            if card[8:10] != "= ":
                continue
            val = card[10:]
            m = re.match(r"\s*'((?:[^']|'')*)'", val)
            if m:
                cards[key] = m.group(1).replace("''", "'").rstrip()
            else:
                v = val.split("/")[0].strip()
                if v in ("T", "F"):
                    cards[key] = v == "T"
                else:
                    try:
                        cards[key] = int(v)
                    except ValueError:
                        try:
                            cards[key] = float(v.replace("D", "E"))
                        except ValueError:
                            cards[key] = v
        if done:
            return cards, off


def read_table(path: str) -> "OrderedDict[str, np.ndarray]":
    """Read the first BINTABLE of a column-oriented FITS file.

    Returns name -> array in native byte order, shaped per TDIM with the FITS
    (fastest-first) axis order reversed into C order.  Logical columns come
    back as bool, character columns as arrays of python strings.
    """
    with open(path, "rb") as fh:
        buf = fh.read()
    hdr, off = _parse_header(buf, 0)
    # skip primary data (normally NAXIS=0)
    naxis = hdr.get("NAXIS", 0)
    if naxis:
        n = abs(hdr["BITPIX"]) // 8
        for k in range(1, naxis + 1):
            n *= hdr[f"NAXIS{k}"]
        off += (n + _BLOCK - 1) // _BLOCK * _BLOCK
    hdr, off = _parse_header(buf, off)
    if hdr.get("XTENSION", "").strip() != "BINTABLE":
        raise ValueError(f"{path}: first extension is not a BINTABLE")
    if hdr["NAXIS2"] != 1:
        raise ValueError(f"{path}: not a column-oriented table (NAXIS2={hdr['NAXIS2']})")
    out = OrderedDict()
    pos = off
    for c in range(1, hdr["TFIELDS"] + 1):
        tform = hdr[f"TFORM{c}"].strip()
        m = re.match(r"(\d*)([LBIJKEDA])", tform)
        if not m:
            raise ValueError(f"{path}: unsupported TFORM '{tform}'")
        count = int(m.group(1)) if m.group(1) else 1
        dt = _TFORM[m.group(2)]
        raw = np.frombuffer(buf, dtype=dt, count=count, offset=pos)
        pos += count * dt.itemsize
        dims = None
        tdim = hdr.get(f"TDIM{c}")
        if tdim:
            dims = [int(x) for x in tdim.strip("() ").split(",") if x.strip()][::-1]
        if m.group(2) == "A":
            if dims and len(dims) >= 1:
                width = dims[-1]
                arr = raw.view(f"S{width}").reshape(dims[:-1] if len(dims) > 1 else ())
                arr = np.char.rstrip(np.char.decode(np.atleast_1d(arr), "ascii"))
                if len(dims) == 1:
                    arr = arr.reshape(())
            else:
                arr = np.array(raw.tobytes().decode("ascii").rstrip())
        elif m.group(2) == "L":
            arr = raw == b"T"
            if dims:
                arr = arr.reshape(dims)
        else:
            arr = raw.astype(dt.newbyteorder("="))
            if dims:
                arr = arr.reshape(dims)
        out[hdr[f"TTYPE{c}"].strip().upper()] = arr
    return out


def _card(key: str, value, comment: str = "") -> bytes:
    if isinstance(value, bool):
        v = f"{'T' if value else 'F':>20}"
    elif isinstance(value, int):
        v = f"{value:>20d}"
    elif isinstance(value, str):
        v = "'" + value.replace("'", "''").ljust(8) + "'"
        v = f"{v:<20}"
    else:
        v = f"{value:>20}"
    card = f"{key:<8}= {v}"
    if comment:
        card += f" / {comment}"
    return card[:80].ljust(80).encode("ascii")


def _pad(b: bytes, fill: bytes) -> bytes:
    r = (-len(b)) % _BLOCK
    return b + fill * r


def write_table(path: str, columns: "dict[str, np.ndarray]") -> None:
    """Write a one-row BINTABLE with one vector column per entry (TDIM reversed C shape)."""
    cards, payload = [], []
    idx = 0
    for name, val in columns.items():
        idx += 1
        if isinstance(val, str):
            val = np.array(val)
        a = np.asarray(val)
        if a.dtype.kind in ("U", "S"):
            s = np.char.encode(np.atleast_1d(a), "ascii") if a.dtype.kind == "U" else np.atleast_1d(a)
            width = max(1, max(len(x) for x in s.ravel()))
            data = b"".join(x.ljust(width) for x in s.ravel())
            shape = list(a.shape) + [width]
            letter = "A"
        elif a.dtype == np.bool_:
            data = np.where(a.ravel(), b"T", b"F").astype("S1").tobytes()
            shape, letter = list(a.shape), "L"
        else:
            conv = {"f4": ("E", ">f4"), "f8": ("D", ">f8"), "i8": ("K", ">i8"), "u8": ("K", ">i8"),
                    "i4": ("J", ">i4"), "u4": ("J", ">i4"), "i2": ("I", ">i2"), "u1": ("B", "u1")}
            key = a.dtype.kind + str(a.dtype.itemsize)
            letter, be = conv[key]
            data = np.ascontiguousarray(a).astype(be).tobytes()
            shape = list(a.shape)
        count = len(data) // _TFORM[letter].itemsize
        cards.append(_card(f"TTYPE{idx}", name.upper()))
        cards.append(_card(f"TFORM{idx}", f"{count}{letter}"))
        if len(shape) > 1 or letter == "A":
            cards.append(_card(f"TDIM{idx}", "(" + ",".join(str(d) for d in shape[::-1]) + ")"))
        payload.append(data)
    body = b"".join(payload)
    prim = _pad(b"".join([_card("SIMPLE", True), _card("BITPIX", 8), _card("NAXIS", 0),
                          _card("EXTEND", True), b"END".ljust(80)]), b" ")
    ext = [_card("XTENSION", "BINTABLE"), _card("BITPIX", 8), _card("NAXIS", 2),
           _card("NAXIS1", len(body)), _card("NAXIS2", 1), _card("PCOUNT", 0), _card("GCOUNT", 1),
           _card("TFIELDS", idx)] + cards + [b"END".ljust(80)]
    with open(path, "wb") as fh:
        fh.write(prim)
        fh.write(_pad(b"".join(ext), b" "))
        fh.write(_pad(body, b"\0"))


def read_filter_db(db_file: str) -> "OrderedDict[str, str]":
    """``name=relative/path.fits`` lines -> name -> absolute path (reference: gencat:317)."""
    import os
    base = os.path.dirname(os.path.abspath(db_file))
    out = OrderedDict()
    with open(db_file) as fh:
        for line in fh:
            line = line.strip()
            if not line or line.startswith("#") or "=" not in line:
                continue
            k, v = line.split("=", 1)
            out[k.strip()] = os.path.join(base, v.strip())
    return out
