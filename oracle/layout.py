"""Python-int <-> b200zk.h limb-array converters for the tests (TEST INFRASTRUCTURE ONLY)."""
from __future__ import annotations

import numpy as np

from . import bn254 as o


def _limbs(x: int):
    return [(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def fr_to_arr(vals) -> np.ndarray:
    """canonical ints -> (n,4) uint64 Montgomery limbs"""
    return np.array([_limbs(o.fr_mont(v % o.R)) for v in vals], dtype=np.uint64).reshape(-1, 4)


def arr_to_fr(arr) -> list:
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [o.fr_unmont(int.from_bytes(row.tobytes(), "little")) for row in arr]


def fq_to_limbs(v: int):
    return _limbs(o.fq_mont(v % o.P))


def g1_to_arr(pts) -> np.ndarray:
    rows = []
    for p in pts:
        rows.append([0] * 8 if p is None else fq_to_limbs(p[0]) + fq_to_limbs(p[1]))
    return np.array(rows, dtype=np.uint64).reshape(-1, 8)


def g2_to_arr(pts) -> np.ndarray:
    rows = []
    for p in pts:
        if p is None:
            rows.append([0] * 16)
        else:
            rows.append(fq_to_limbs(p[0][0]) + fq_to_limbs(p[0][1]) + fq_to_limbs(p[1][0]) + fq_to_limbs(p[1][1]))
    return np.array(rows, dtype=np.uint64).reshape(-1, 16)


def _fq(row) -> int:
    return o.fq_unmont(int.from_bytes(np.asarray(row, dtype=np.uint64).tobytes(), "little"))


def arr_to_g1(arr) -> list:
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 8)
    out = []
    for row in arr:
        out.append(None if not row.any() else (_fq(row[:4]), _fq(row[4:])))
    return out


def arr_to_g2(arr) -> list:
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 16)
    out = []
    for row in arr:
        out.append(None if not row.any() else ((_fq(row[:4]), _fq(row[4:8])), (_fq(row[8:12]), _fq(row[12:]))))
    return out
