"""Tranche files: the reference's CSV (stage_3_synthetic_data_generation.py:42,49-50) and a binary
row-major format that can be read straight into pinned host memory.

CSV parsing runs at ~1 M rows/s/core -- four orders of magnitude below the fit -- so large
tranches are stored as ``regression-dataset-YYYY-MM-DD.b2t``:

    offset  size  field
    0       8     magic  b"B2TRNCH1"
    8       8     n_rows (uint64 LE)
    16      4     d      (uint32)
    20      4     x_kind (uint32: 0 = fp32, 1 = bf16 bit patterns)
    24      10    date   "YYYY-MM-DD" ASCII
    34      30    zero padding (header is 64 bytes)
    64      ...   X row-major [n_rows][d]  then  y fp32 [n_rows]   (X padded to a 64-byte boundary)

The same key scheme and date regex as the reference (stage_1_train_model.py:47) select and order tranches.
"""
from __future__ import annotations

import os
import re
import struct
from datetime import date, datetime
from typing import List, Tuple

import numpy as np

MAGIC = b"B2TRNCH1"
HEADER_BYTES = 64
_DATE_RE = re.compile("20[2-9][0-9]-[0-1][0-9]-[0-3][0-9]")


def _pad64(n: int) -> int:
    return (n + 63) // 64 * 64


def write_tranche(path: str, X: np.ndarray, y: np.ndarray, day: date, x_kind: str = "f32") -> None:
    X = np.ascontiguousarray(X, dtype=np.float32 if x_kind == "f32" else np.uint16)
    y = np.ascontiguousarray(y, dtype=np.float32)
    if X.ndim != 2 or y.shape != (X.shape[0],):
        raise ValueError("X must be (n, d) and y (n,)")
    header = MAGIC + struct.pack("<QII", X.shape[0], X.shape[1], 0 if x_kind == "f32" else 1) + \
        str(day).encode("ascii")
    header += b"\0" * (HEADER_BYTES - len(header))
    with open(path, "wb") as fh:
        fh.write(header)
        fh.write(X.tobytes())
        fh.write(b"\0" * (_pad64(X.nbytes) - X.nbytes))
        fh.write(y.tobytes())


def read_header(path: str) -> Tuple[int, int, str, date]:
    with open(path, "rb") as fh:
        h = fh.read(HEADER_BYTES)
    if len(h) != HEADER_BYTES or h[:8] != MAGIC:
        raise RuntimeError(f"{path} is not a b2 tranche file")
    n, d, kind = struct.unpack("<QII", h[8:24])
    day = datetime.strptime(h[24:34].decode("ascii"), "%Y-%m-%d").date()
    return int(n), int(d), ("f32" if kind == 0 else "bf16"), day


def read_tranche(path: str, out_X: np.ndarray = None, out_y: np.ndarray = None):
    """Read a tranche; with ``out_X`` / ``out_y`` (e.g. views of pinned buffers) the bytes land there directly."""
    n, d, kind, day = read_header(path)
    xdt = np.float32 if kind == "f32" else np.uint16
    xbytes = n * d * np.dtype(xdt).itemsize
    if out_X is None:
        out_X = np.empty((n, d), dtype=xdt)
    if out_y is None:
        out_y = np.empty(n, dtype=np.float32)
    if out_X.shape != (n, d) or out_X.dtype != xdt or out_y.shape != (n,) or not out_X.flags.c_contiguous:
        raise ValueError("destination buffers do not match the tranche")
    with open(path, "rb") as fh:
        fh.seek(HEADER_BYTES)
        got = fh.readinto(memoryview(out_X).cast("B"))
        fh.seek(HEADER_BYTES + _pad64(xbytes))
        got_y = fh.readinto(memoryview(out_y).cast("B"))
    if got != xbytes or got_y != n * 4:
        raise RuntimeError(f"{path} is truncated")
    return out_X, out_y, kind, day


def list_tranches(folder: str) -> List[Tuple[str, date]]:
    """(path, date) of every tranche file (CSV or binary), oldest first -- stage_1_train_model.py:62-67."""
    out = []
    for key in sorted(os.listdir(folder)):
        m = _DATE_RE.findall(key)
        if m and key.endswith((".csv", ".b2t")):
            out.append((os.path.join(folder, key), datetime.strptime(m[0], "%Y-%m-%d").date()))
    return sorted(out, key=lambda e: e[1])


def load_all(folder: str, ctx=None):
    """Concatenate every tranche of a folder into (X, y, newest date).  Binary tranches are read directly into one
    pinned buffer when a context is given (ready for B2_MEM_HOST streaming); CSV tranches go through pandas."""
    import pandas as pd
    items = list_tranches(folder)
    if not items:
        raise RuntimeError(f"no tranche files under {folder}")
    sizes, d = [], None
    for path, _ in items:
        if path.endswith(".b2t"):
            n, dd, kind, _ = read_header(path)
            if kind != "f32":
                raise RuntimeError("load_all concatenates fp32 tranches only")
        else:
            df = pd.read_csv(path, usecols=lambda c: c != "date")
            n, dd = len(df), len([c for c in df.columns if c != "y"])
        if d is not None and dd != d:
            raise RuntimeError("tranches disagree on the feature count")
        d = dd
        sizes.append(n)
    total = int(sum(sizes))
    if ctx is not None:
        Xp, yp = ctx.pinned((total, d), np.float32), ctx.pinned((total,), np.float32)
        X, y, keep = Xp.array, yp.array, (Xp, yp)
    else:
        X, y, keep = np.empty((total, d), np.float32), np.empty(total, np.float32), None
    lo = 0
    for (path, _), n in zip(items, sizes):
        if path.endswith(".b2t"):
            read_tranche(path, X[lo:lo + n], y[lo:lo + n])
        else:
            df = pd.read_csv(path)
            cols = ["X"] if "X" in df.columns else sorted((c for c in df.columns if re.fullmatch(r"X\d+", c)),
                                                          key=lambda c: int(c[1:]))
            X[lo:lo + n] = df[cols].to_numpy(dtype=np.float32)
            y[lo:lo + n] = df["y"].to_numpy(dtype=np.float32)
        lo += n
    return X, y, items[-1][1], keep
