"""Reader / writer for the "JPPB0001" container: a flat list of named n-d arrays.

Staged lattice batches (the pre-state of ``computeScores``) and the expected
post-state travel in this format between the reference harness, the CPU oracle,
the tests and ``bench.py``.  Layout (little endian):

    "JPPB0001" | u32 n_arrays | n_arrays x {
        u32 name_len | name | u32 dtype | u32 ndim | u64 dims[ndim] | u64 nbytes |
        pad to 8 | data }
"""
from __future__ import annotations

import struct
from pathlib import Path

import numpy as np

_DTYPES = {0: np.uint8, 1: np.int32, 2: np.uint32, 3: np.float32, 4: np.uint64, 5: np.uint16}
_CODES = {np.dtype(v): k for k, v in _DTYPES.items()}
MAGIC = b"JPPB0001"


def load(path: str | Path, mmap: bool = False) -> dict[str, np.ndarray]:
    path = str(path)
    if path.endswith(".gz"):
        import gzip
        with gzip.open(path, "rb") as f:
            return loads(f.read())
    raw = np.memmap(path, dtype=np.uint8, mode="r") if mmap else np.fromfile(path, dtype=np.uint8)
    return _parse(raw, path)


def loads(data: bytes) -> dict[str, np.ndarray]:
    return _parse(np.frombuffer(data, dtype=np.uint8).copy(), "<bytes>")


def _parse(raw: np.ndarray, path: str) -> dict[str, np.ndarray]:
    if bytes(raw[:8]) != MAGIC:
        raise ValueError(f"{path}: not a JPPB0001 file")
    pos = 8
    (n,) = struct.unpack_from("<I", raw, pos)
    pos += 4
    out: dict[str, np.ndarray] = {}
    for _ in range(n):
        (nl,) = struct.unpack_from("<I", raw, pos)
        pos += 4
        name = bytes(raw[pos:pos + nl]).decode()
        pos += nl
        dtype, nd = struct.unpack_from("<II", raw, pos)
        pos += 8
        dims = struct.unpack_from(f"<{nd}Q", raw, pos)
        pos += 8 * nd
        (nbytes,) = struct.unpack_from("<Q", raw, pos)
        pos += 8
        pos += (8 - pos % 8) % 8
        arr = raw[pos:pos + nbytes].view(_DTYPES[dtype]).reshape(dims)
        out[name] = arr
        pos += nbytes
    return out


def save(path: str | Path, arrays: dict[str, np.ndarray]) -> None:
    with open(path, "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<I", len(arrays)))
        for name, a in arrays.items():
            a = np.ascontiguousarray(a)
            nb = name.encode()
            f.write(struct.pack("<I", len(nb)))
            f.write(nb)
            f.write(struct.pack("<II", _CODES[a.dtype], a.ndim))
            f.write(struct.pack(f"<{a.ndim}Q", *a.shape))
            f.write(struct.pack("<Q", a.nbytes))
            pos = f.tell()
            f.write(b"\0" * ((8 - pos % 8) % 8))
            f.write(a.tobytes())
