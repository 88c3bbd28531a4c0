"""snarkjs / circom binary formats -> limb arrays.

Mirrors /root/reference/ark-circom/src/zkey.rs:53-387 (`read_zkey`: header, IC, coefficient section, the five
query sections), ark-circom/src/circom/r1cs_reader.rs:54-249 and the .wtns layout the reference consumes
through its witness calculator.  Pure byte shuffling with numpy -- no field arithmetic happens on the host:

  * curve points are stored by snarkjs in Montgomery form already (zkey.rs:340-345) and are returned as the
    (n, 8) / (n, 16) u64 limb arrays the C ABI takes; (0, 0) stays the infinity encoding (zkey.rs:353-373);
  * matrix coefficients are stored multiplied by R^2 (zkey.rs:333-338) and witness / r1cs values are
    canonical; they are returned raw together with the number of Montgomery reductions / conversions the
    device must apply (`Net.fr_convert`), which is how `groth16.qap.qap_from_zkey` feeds them to the GPU.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

FQ_MODULUS = 21888242871839275222246405745257275088696311157297823662689037894645226208583
FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617


class FormatError(ValueError):
    pass


def _sections(buf: bytes, magic: bytes):
    if buf[:4] != magic:
        raise FormatError("bad magic %r (expected %r)" % (buf[:4], magic))
    _version, nsec = struct.unpack_from("<II", buf, 4)
    off = 12
    secs = {}
    for _ in range(nsec):
        sid, ln = struct.unpack_from("<IQ", buf, off)
        off += 12
        secs.setdefault(sid, []).append((off, ln))
        off += ln
    return secs


def _limbs(buf: bytes, off: int, count: int, width: int) -> np.ndarray:
    """count records of width u64 limbs starting at byte offset off."""
    return np.frombuffer(buf, dtype="<u8", count=count * width, offset=off).reshape(count, width).copy()


@dataclass
class ZKey:
    n_vars: int
    n_public: int
    domain_size: int
    alpha_g1: np.ndarray
    beta_g1: np.ndarray
    beta_g2: np.ndarray
    gamma_g2: np.ndarray
    delta_g1: np.ndarray
    delta_g2: np.ndarray
    ic: np.ndarray
    a_query: np.ndarray
    b_g1_query: np.ndarray
    b_g2_query: np.ndarray
    l_query: np.ndarray
    h_query: np.ndarray
    # coefficient section as COO triplets; values are value * R^2 mod r (raw file words)
    coef_matrix: np.ndarray
    coef_row: np.ndarray
    coef_col: np.ndarray
    coef_val_r2: np.ndarray

    @property
    def n_inputs(self) -> int:                 # num_instance_variables = n_public + 1 (zkey.rs:178)
        return self.n_public + 1

    @property
    def num_constraints(self) -> int:
        """zkey.rs:171: max constraint index - n_public (the appended public-input rows are dropped)."""
        return int(self.coef_row.max()) - self.n_public if self.coef_row.size else 0

    def vk_points(self) -> np.ndarray:
        """alpha_g1 beta_g1 delta_g1 beta_g2 delta_g2 -- the 56 limbs b200zk_pk_upload takes."""
        return np.concatenate([self.alpha_g1, self.beta_g1, self.delta_g1, self.beta_g2, self.delta_g2]).astype(np.uint64)


def read_zkey(buf: bytes) -> ZKey:
    secs = _sections(buf, b"zkey")
    for sid in (2, 3, 4, 5, 6, 7, 8, 9):
        if sid not in secs:
            raise FormatError("zkey section %d missing" % sid)
    off, _ = secs[2][0]
    n8q = struct.unpack_from("<I", buf, off)[0]
    off += 4
    q = int.from_bytes(buf[off:off + n8q], "little")
    off += n8q
    n8r = struct.unpack_from("<I", buf, off)[0]
    off += 4
    r = int.from_bytes(buf[off:off + n8r], "little")
    off += n8r
    if q != FQ_MODULUS or r != FR_MODULUS or n8q != 32 or n8r != 32:
        raise FormatError("zkey is not over BN254")
    n_vars, n_public, domain_size = struct.unpack_from("<III", buf, off)
    off += 12
    alpha_g1 = _limbs(buf, off, 1, 8)[0]; off += 64
    beta_g1 = _limbs(buf, off, 1, 8)[0]; off += 64
    beta_g2 = _limbs(buf, off, 1, 16)[0]; off += 128
    gamma_g2 = _limbs(buf, off, 1, 16)[0]; off += 128
    delta_g1 = _limbs(buf, off, 1, 8)[0]; off += 64
    delta_g2 = _limbs(buf, off, 1, 16)[0]; off += 128

    def sec(sid, count, width):
        o, ln = secs[sid][0]
        if ln < count * width * 8:
            raise FormatError("zkey section %d too short" % sid)
        return _limbs(buf, o, count, width)

    if n_vars == 0 or n_public + 1 > n_vars:
        raise FormatError("zkey header: n_public + 1 = %d variables are public but n_vars = %d" % (n_public + 1, n_vars))
    if domain_size == 0 or domain_size & (domain_size - 1):
        raise FormatError("zkey header: domain size %d is not a power of two" % domain_size)
    o4, l4 = secs[4][0]
    ncoef = struct.unpack_from("<I", buf, o4)[0]
    if l4 < 4 + ncoef * 44:
        raise FormatError("zkey coefficient section too short for %d records" % ncoef)
    rec = np.frombuffer(buf, dtype=np.dtype([("m", "<u4"), ("c", "<u4"), ("s", "<u4"), ("v", "<u8", (4,))]), count=ncoef,
                        offset=o4 + 4)
    # the device kernels index z[col] and the QAP rows with these values unchecked (csrc/qap.cu): reject what the reference
    # would panic on (index out of bounds) here, on the host
    if ncoef and (int(rec["m"].max()) > 1 or int(rec["s"].max()) >= n_vars or int(rec["c"].max()) >= domain_size):
        raise FormatError("zkey coefficient section: matrix index > 1, signal index >= n_vars or constraint index >= domain size")
    return ZKey(n_vars=n_vars, n_public=n_public, domain_size=domain_size, alpha_g1=alpha_g1, beta_g1=beta_g1,
                beta_g2=beta_g2, gamma_g2=gamma_g2, delta_g1=delta_g1, delta_g2=delta_g2,
                ic=sec(3, n_public + 1, 8), a_query=sec(5, n_vars, 8), b_g1_query=sec(6, n_vars, 8),
                b_g2_query=sec(7, n_vars, 16), l_query=sec(8, n_vars - n_public - 1, 8), h_query=sec(9, domain_size, 8),
                coef_matrix=rec["m"].copy(), coef_row=rec["c"].copy(), coef_col=rec["s"].copy(),
                coef_val_r2=rec["v"].copy())


def read_wtns(buf: bytes) -> np.ndarray:
    """(n, 4) u64 canonical (non-Montgomery) witness values."""
    secs = _sections(buf, b"wtns")
    off, _ = secs[1][0]
    n8 = struct.unpack_from("<I", buf, off)[0]
    prime = int.from_bytes(buf[off + 4:off + 4 + n8], "little")
    if n8 != 32 or prime != FR_MODULUS:
        raise FormatError("wtns is not over BN254 Fr")
    n = struct.unpack_from("<I", buf, off + 4 + n8)[0]
    o2, ln = secs[2][0]
    if ln < n * 32:
        raise FormatError("wtns data section too short")
    return _limbs(buf, o2, n, 4)


@dataclass
class R1CS:
    n_wires: int
    n_pub_out: int
    n_pub_in: int
    n_prv_in: int
    n_constraints: int
    # COO per matrix (A, B, C): rows, cols, canonical coefficient limbs
    rows: list
    cols: list
    vals: list


def read_r1cs(buf: bytes) -> R1CS:
    secs = _sections(buf, b"r1cs")
    off, _ = secs[1][0]
    fs = struct.unpack_from("<I", buf, off)[0]
    prime = int.from_bytes(buf[off + 4:off + 4 + fs], "little")
    if fs != 32 or prime != FR_MODULUS:                       # r1cs_reader.rs:180-188
        raise FormatError("r1cs is not over BN254 Fr")
    off += 4 + fs
    n_wires, n_pub_out, n_pub_in, n_prv_in = struct.unpack_from("<IIII", buf, off)
    off += 16 + 8
    n_constraints = struct.unpack_from("<I", buf, off)[0]
    off, l2 = secs[2][0]
    end2 = off + l2
    rows = [[], [], []]
    cols = [[], [], []]
    vals = [[], [], []]
    term = np.dtype([("w", "<u4"), ("v", "<u8", (4,))])
    for i in range(n_constraints):
        for k in range(3):
            if off + 4 > end2:
                raise FormatError("r1cs constraint section ends inside constraint %d" % i)
            nterm = struct.unpack_from("<I", buf, off)[0]
            off += 4
            if off + nterm * 36 > end2:
                raise FormatError("r1cs constraint section ends inside constraint %d" % i)
            t = np.frombuffer(buf, dtype=term, count=nterm, offset=off)
            off += nterm * 36
            if nterm and int(t["w"].max()) >= n_wires:
                raise FormatError("r1cs constraint %d names wire %d of %d" % (i, int(t["w"].max()), n_wires))
            rows[k].append(np.full(nterm, i, dtype=np.uint32))
            cols[k].append(t["w"].astype(np.uint32))
            vals[k].append(t["v"].astype(np.uint64).reshape(nterm, 4))
    cat = lambda parts, shape: np.concatenate(parts) if parts else np.zeros(shape, dtype=np.uint32)
    return R1CS(n_wires, n_pub_out, n_pub_in, n_prv_in, n_constraints,
                [cat(r, (0,)) for r in rows], [cat(c, (0,)) for c in cols],
                [np.concatenate(v) if v else np.zeros((0, 4), dtype=np.uint64) for v in vals])


def coo_to_csr(rows: np.ndarray, cols: np.ndarray, vals: np.ndarray, n_rows: int):
    """Stable sort by row -> (row_ptr[n_rows+1] u32, col u32, val (nnz,4) u64). Index-only host work."""
    keep = rows < n_rows
    rows, cols, vals = rows[keep], cols[keep], vals[keep]
    order = np.argsort(rows, kind="stable")
    counts = np.bincount(rows, minlength=n_rows).astype(np.uint64)
    row_ptr = np.zeros(n_rows + 1, dtype=np.uint32)
    row_ptr[1:] = np.cumsum(counts).astype(np.uint32)
    return row_ptr, np.ascontiguousarray(cols[order], dtype=np.uint32), np.ascontiguousarray(vals[order], dtype=np.uint64)
