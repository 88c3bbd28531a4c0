"""ark-serialize `Compress::Yes` forms of the Groth16 key / proof types, with the point codec on the GPU.

The reference serialises everything it stores or ships with `CanonicalSerialize::serialize_compressed`
(common/src/utils/serializer.rs:20-49): `proving_key.bin` / the verifying key (mpc-api/src/main.rs:161-165) and the
128-byte proof (zk-cli/src/main.rs:130-136).  Layout = the derive order of ark-groth16 0.4's structs (third-party, not
in /root/reference; its `Proof` layout is pinned by zk-cli/test-circuits/sha256/proof.bin, the key layouts follow the
same rules: fields in declaration order, `Vec<T>` = u64-LE length + items):

    VerifyingKey { alpha_g1: G1, beta_g2: G2, gamma_g2: G2, delta_g2: G2, gamma_abc_g1: Vec<G1> }
    ProvingKey   { vk, beta_g1: G1, delta_g1: G1, a_query: Vec<G1>, b_g1_query: Vec<G1>, b_g2_query: Vec<G2>,
                   h_query: Vec<G1>, l_query: Vec<G1> }
    Proof        { a: G1, b: G2, c: G1 }

Points are host u64 arrays of Montgomery limbs (n, 8) / (n, 16), infinity all-zero -- the layout of the rest of the package.
Decompressing a key costs one square root per point: that runs as `b200zk_points_decompress_dev` (one thread per point).
"""
from __future__ import annotations

import struct
from dataclasses import dataclass

import numpy as np

from .context import Net

VK_FIELDS = (("alpha_g1", False), ("beta_g2", True), ("gamma_g2", True), ("delta_g2", True))
PK_POINTS = (("beta_g1", False), ("delta_g1", False))
PK_VECS = (("a_query", False), ("b_g1_query", False), ("b_g2_query", True), ("h_query", False), ("l_query", False))


@dataclass
class ArkVerifyingKey:
    alpha_g1: np.ndarray
    beta_g2: np.ndarray
    gamma_g2: np.ndarray
    delta_g2: np.ndarray
    gamma_abc_g1: np.ndarray


@dataclass
class ArkProvingKey:
    vk: ArkVerifyingKey
    beta_g1: np.ndarray
    delta_g1: np.ndarray
    a_query: np.ndarray
    b_g1_query: np.ndarray
    b_g2_query: np.ndarray
    h_query: np.ndarray
    l_query: np.ndarray


def _enc(net: Net, pts, g2: bool) -> bytes:
    arr = np.ascontiguousarray(pts, dtype=np.uint64).reshape(-1, 16 if g2 else 8)
    if arr.shape[0] == 0:
        return b""
    return net.points_compress(arr, g2=g2).cpu().numpy().tobytes()


def _dec(net: Net, buf, n: int, g2: bool, check_subgroup: bool) -> np.ndarray:
    if n == 0:
        return np.zeros((0, 16 if g2 else 8), dtype=np.uint64)
    return net.points_decompress(bytes(buf), g2=g2, check_subgroup=check_subgroup).cpu().numpy().view(np.uint64)


class _Reader:
    def __init__(self, buf: bytes):
        self.buf, self.off = memoryview(bytes(buf)), 0

    def take(self, n: int):
        if self.off + n > len(self.buf):
            raise ValueError("ark-serialize: unexpected end of input")
        v = self.buf[self.off:self.off + n]
        self.off += n
        return v

    def u64(self) -> int:
        return struct.unpack("<Q", self.take(8))[0]


def serialize_verifying_key(net: Net, vk: ArkVerifyingKey) -> bytes:
    out = b"".join(_enc(net, getattr(vk, name), g2) for name, g2 in VK_FIELDS)
    abc = np.ascontiguousarray(vk.gamma_abc_g1, dtype=np.uint64).reshape(-1, 8)
    return out + struct.pack("<Q", abc.shape[0]) + _enc(net, abc, False)


def _read_vk(net: Net, rd: _Reader, check_subgroup: bool) -> ArkVerifyingKey:
    vals = {name: _dec(net, rd.take(64 if g2 else 32), 1, g2, check_subgroup)[0] for name, g2 in VK_FIELDS}
    n = rd.u64()
    vals["gamma_abc_g1"] = _dec(net, rd.take(32 * n), n, False, check_subgroup)
    return ArkVerifyingKey(**vals)


def deserialize_verifying_key(net: Net, buf: bytes, check_subgroup: bool = True) -> ArkVerifyingKey:
    rd = _Reader(buf)
    vk = _read_vk(net, rd, check_subgroup)
    if rd.off != len(rd.buf):
        raise ValueError("ark-serialize: trailing bytes")
    return vk


def serialize_proving_key(net: Net, pk: ArkProvingKey) -> bytes:
    parts = [serialize_verifying_key(net, pk.vk)]
    parts += [_enc(net, getattr(pk, name), g2) for name, g2 in PK_POINTS]
    for name, g2 in PK_VECS:
        arr = np.ascontiguousarray(getattr(pk, name), dtype=np.uint64).reshape(-1, 16 if g2 else 8)
        parts.append(struct.pack("<Q", arr.shape[0]))
        parts.append(_enc(net, arr, g2))
    return b"".join(parts)


def deserialize_proving_key(net: Net, buf: bytes, check_subgroup: bool = False) -> ArkProvingKey:
    """check_subgroup=True is arkworks' `Validate::Yes` (every G2 point times r); False = `deserialize_compressed_unchecked`
    plus the on-curve check the square root gives for free."""
    rd = _Reader(buf)
    vals = {"vk": _read_vk(net, rd, check_subgroup)}
    for name, g2 in PK_POINTS:
        vals[name] = _dec(net, rd.take(64 if g2 else 32), 1, g2, check_subgroup)[0]
    for name, g2 in PK_VECS:
        n = rd.u64()
        vals[name] = _dec(net, rd.take((64 if g2 else 32) * n), n, g2, check_subgroup)
    if rd.off != len(rd.buf):
        raise ValueError("ark-serialize: trailing bytes")
    return ArkProvingKey(**vals)


def serialize_proof(net: Net, a, b, c) -> bytes:
    return _enc(net, a, False) + _enc(net, b, True) + _enc(net, c, False)


def deserialize_proof(net: Net, buf: bytes, check_subgroup: bool = True):
    if len(buf) != 128:
        raise ValueError("a compressed Proof<Bn254> is 128 bytes")
    return (_dec(net, buf[:32], 1, False, check_subgroup)[0], _dec(net, buf[32:96], 1, True, check_subgroup)[0],
            _dec(net, buf[96:], 1, False, check_subgroup)[0])
