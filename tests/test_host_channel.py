"""dist_primitives/channel.py: the reference's king / client wire format (channel/mod.rs:8-56, mpc-net/src/multi.rs:26-33,
prod.rs:126-131) at the byte level, on a CPU stand-in for `Net` whose conversions come from the oracle codec."""
import struct

import numpy as np
import pytest

from distributed_groth16_b200 import MpcNetError
from distributed_groth16_b200.dist_primitives import channel as ch


class _T:
    """numpy array with the two tensor methods channel.py touches"""

    def __init__(self, a):
        self.a = np.ascontiguousarray(a)
        self.shape = self.a.shape

    def cpu(self):
        return self

    def numpy(self):
        return self.a

    def contiguous(self):
        return self

    def data_ptr(self):          # marks it as "already on the device" for serialize_fr_vec
        return 0


class OracleCodecNet:
    def to_device(self, arr):
        return _T(np.ascontiguousarray(arr, dtype=np.uint64).view(np.int64))

    def fr_convert(self, x, to_mont, times=1):
        from oracle import bn254 as o, layout
        vals = layout.arr_to_fr_raw(x.a.view(np.uint64).reshape(-1, 4)) if hasattr(layout, "arr_to_fr_raw") else \
            [int.from_bytes(row.tobytes(), "little") for row in x.a.view(np.uint64).reshape(-1, 4)]
        R = (1 << 256) % o.R
        conv = (lambda v: v * R % o.R) if to_mont else (lambda v: v * pow(R, -1, o.R) % o.R)
        out = np.array([[(conv(v) >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for v in vals], dtype=np.uint64)
        return _T(out.view(np.int64))

    def points_compress(self, pts, g2=False):
        from oracle import bn254 as o, layout
        w = 16 if g2 else 8
        p = (layout.arr_to_g2 if g2 else layout.arr_to_g1)(np.asarray(pts, dtype=np.uint64).reshape(-1, w))
        enc = b"".join((o.g2_compress if g2 else o.g1_compress)(q) for q in p)
        return _T(np.frombuffer(enc, dtype=np.uint8).reshape(len(p), -1))

    def points_decompress(self, data, g2=False, check_subgroup=False):
        from oracle import bn254 as o, layout
        w = 64 if g2 else 32
        pts = [(o.g2_decompress if g2 else o.g1_decompress)(bytes(data[i:i + w])) for i in range(0, len(data), w)]
        return _T((layout.g2_to_arr if g2 else layout.g1_to_arr)(pts).view(np.int64))


def test_frames_and_prod_packets_are_the_documented_bytes():
    assert ch.frame(b"abc") == bytes([0, 0, 0, 3]) + b"abc"                      # LengthDelimitedCodec, big-endian u32
    assert ch.prod_packet(b"abc") == bytes([2, 0, 0, 0]) + struct.pack("<Q", 3) + b"abc"      # bincode2 enum variant 2 + Vec<u8>
    msg = ch.client_message(b"hello", prod=True)
    assert msg[:4] == struct.pack(">I", 4 + 8 + 5)
    stream = msg + ch.client_message(b"second", prod=True)[:7]                  # one whole frame + a partial one
    payload, rest = ch.read_message(stream, prod=True)
    assert payload == b"hello"
    assert ch.read_message(rest, prod=True) == (None, rest)                      # incomplete frame: nothing consumed
    assert ch.parse_prod_packet(struct.pack("<I", 0)) == (ch.PACKET_SYN, None)
    with pytest.raises(MpcNetError):
        ch.parse_prod_packet(struct.pack("<IQ", 2, 9) + b"short")
    own, out = ch.king_scatter([b"aa", b"bb", b"cc"])
    assert own == b"aa" and out == {1: ch.frame(b"bb"), 2: ch.frame(b"cc")}
    with pytest.raises(MpcNetError) as e:
        ch.king_scatter([b"aa", b"b"])
    assert e.value.kind == "Protocol"                                             # "Peer 1 sent wrong number of bytes"


def test_fr_vec_payload_is_ark_serialize_compressed(cref):
    from oracle import bn254 as o, layout
    net = OracleCodecNet()
    x = cref.fr_generate(5, 9)
    vals = layout.arr_to_fr(x)
    want = struct.pack("<Q", 9) + b"".join(v.to_bytes(32, "little") for v in vals)   # u64-LE length + canonical LE integers
    got = ch.serialize_fr_vec(net, x)
    assert got == want
    assert (ch.deserialize_fr_vec(net, got) == x).all()
    assert ch.serialize_fr_vec(net, x[:0]) == struct.pack("<Q", 0)
    bad = struct.pack("<Q", 1) + o.R.to_bytes(32, "little")                       # r itself is not canonical
    with pytest.raises(MpcNetError):
        ch.deserialize_fr_vec(net, bad)
    with pytest.raises(MpcNetError):
        ch.deserialize_fr_vec(net, got + b"\0")
    with pytest.raises(MpcNetError):
        ch.deserialize_fr_vec(net, got[:-1])


def test_group_element_payload_is_the_compressed_point(cref):
    from oracle import bn254 as o, layout
    net = OracleCodecNet()
    p1, p2 = cref.g1_generate(3, 2), cref.g2_generate(4, 1)
    enc = ch.serialize_point(net, p1[0])
    assert enc == o.g1_compress(layout.arr_to_g1(p1[:1])[0]) and len(enc) == 32
    assert (ch.deserialize_point(net, enc) == p1[0]).all()
    enc2 = ch.serialize_point(net, p2[0], g2=True)
    assert enc2 == o.g2_compress(layout.arr_to_g2(p2)[0]) and len(enc2) == 64
    assert ch.serialize_point(net, np.zeros(8, dtype=np.uint64)) == bytes(31) + bytes([0x40])     # infinity flag
    with pytest.raises(MpcNetError):
        ch.deserialize_point(net, enc[:-1])
