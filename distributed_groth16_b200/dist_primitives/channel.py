"""The king / client wire format of the reference, so that a GPU node can sit in its real network as one party.

What travels (all of it mirrored here, byte for byte):
  * payload  -- `MpcSerNet::{send_to_king, recv_from_king}` (/root/reference/dist-primitives/src/channel/mod.rs:8-56):
                `T::serialize_compressed`, T = `G` for d_msm (dmsm/mod.rs:84,94: one compressed group element, 32 B G1 / 64 B G2)
                and T = `Vec<F>` for d_fft / d_ifft (dfft/mod.rs:185-256: u64-LE length, then every share as its canonical
                -- not Montgomery -- 32-byte little-endian integer);
  * frame    -- tokio's `LengthDelimitedCodec`, big-endian u32 length prefix (mpc-net/src/multi.rs:26-33), one framed stream
                per `MultiplexedStreamID` (three per peer, multi.rs:62-99);
  * packet   -- `ProdNet` wraps every payload as bincode2 `ProtocolPacket::Packet(Vec<u8>)` (mpc-net/src/prod.rs:126-131,
                352-378): u32-LE variant index 2, u64-LE byte count, bytes; `LocalTestNet` sends the bare payload
                (multi.rs:371-398).
The Montgomery conversions and the point (de)compression run on the GPU (`Net.fr_convert`, `Net.points_compress`,
`Net.points_decompress`: one square root per point); framing is byte shuffling.  The sockets, TLS and the smux multiplexer
themselves stay out of scope: these functions produce / consume exactly the bytes those layers carry."""
from __future__ import annotations

import struct

import numpy as np

from ..context import MpcNetError

FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617
PACKET_SYN, PACKET_SYNACK, PACKET_PACKET = 0, 1, 2          # enum ProtocolPacket, prod.rs:126-131
MULTIPLEXED_STREAMS = 3                                     # multi.rs:63


# ---- frames (LengthDelimitedCodec, big-endian u32) ---------------------------------------------------------------------------
def frame(payload: bytes) -> bytes:
    if len(payload) >= 1 << 32:
        raise MpcNetError("Generic", "frame too large")
    return struct.pack(">I", len(payload)) + bytes(payload)


def unframe(buf: bytes):
    """-> (payload, rest of the stream), or (None, buf) while the frame is still incomplete (the codec's behaviour)."""
    if len(buf) < 4:
        return None, buf
    (n,) = struct.unpack(">I", buf[:4])
    if len(buf) < 4 + n:
        return None, buf
    return bytes(buf[4:4 + n]), bytes(buf[4 + n:])


# ---- ProdNet packets (bincode2) ------------------------------------------------------------------------------------------------
def prod_packet(payload: bytes) -> bytes:
    """bincode2::serialize(&ProtocolPacket::Packet(payload))"""
    return struct.pack("<IQ", PACKET_PACKET, len(payload)) + bytes(payload)


def parse_prod_packet(buf: bytes):
    """-> (variant, payload or None)"""
    if len(buf) < 4:
        raise MpcNetError("Generic", "io error: unexpected end of file")
    (variant,) = struct.unpack("<I", buf[:4])
    if variant in (PACKET_SYN, PACKET_SYNACK):
        if len(buf) != 4:
            raise MpcNetError("Generic", "trailing bytes after a unit variant")
        return variant, None
    if variant != PACKET_PACKET or len(buf) < 12:
        raise MpcNetError("Generic", "invalid ProtocolPacket")
    (n,) = struct.unpack("<Q", buf[4:12])
    if len(buf) != 12 + n:
        raise MpcNetError("Generic", "ProtocolPacket length mismatch")
    return variant, bytes(buf[12:])


# ---- payloads (ark-serialize, Compress::Yes) ------------------------------------------------------------------------------------
def serialize_fr_vec(net, shares) -> bytes:
    """`Vec<Fr>::serialize_compressed`: shares (n, 4) Montgomery limbs (host u64 array or CUDA int64 tensor)."""
    d = shares if hasattr(shares, "data_ptr") else net.to_device(np.ascontiguousarray(shares, dtype=np.uint64).reshape(-1, 4))
    n = int(d.shape[0])
    body = net.fr_convert(d.contiguous(), to_mont=False).cpu().numpy().tobytes() if n else b""
    return struct.pack("<Q", n) + body


def deserialize_fr_vec(net, buf: bytes) -> np.ndarray:
    """-> (n, 4) Montgomery limbs; rejects trailing bytes and non-canonical elements like arkworks (`InvalidData`)."""
    if len(buf) < 8:
        raise MpcNetError("Generic", "io error: unexpected end of file")
    (n,) = struct.unpack("<Q", buf[:8])
    if len(buf) != 8 + 32 * n:
        raise MpcNetError("Generic", "io error: unexpected end of file" if len(buf) < 8 + 32 * n else "trailing bytes")
    if n == 0:
        return np.zeros((0, 4), dtype=np.uint64)
    raw = np.frombuffer(buf, dtype="<u8", offset=8).reshape(n, 4)
    # canonical range check: compare the limbs with r from the top
    r = np.array([(FR_MODULUS >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
    lt = np.zeros(n, dtype=bool)
    eq = np.ones(n, dtype=bool)
    for i in (3, 2, 1, 0):
        lt |= eq & (raw[:, i] < r[i])
        eq &= raw[:, i] == r[i]
    if not lt.all():
        raise MpcNetError("Generic", "the input buffer contained invalid data")
    return net.fr_convert(net.to_device(raw.astype(np.uint64)), to_mont=True).cpu().numpy().view(np.uint64)


def serialize_point(net, limbs, g2: bool = False) -> bytes:
    """`G::serialize_compressed` of one group element given as affine Montgomery limbs (all-zero = infinity)."""
    arr = np.ascontiguousarray(limbs, dtype=np.uint64).reshape(1, 16 if g2 else 8)
    return net.points_compress(arr, g2=g2).cpu().numpy().tobytes()


def deserialize_point(net, buf: bytes, g2: bool = False) -> np.ndarray:
    if len(buf) != (64 if g2 else 32):
        raise MpcNetError("Generic", "io error: unexpected end of file")
    try:
        return net.points_decompress(bytes(buf), g2=g2, check_subgroup=True).cpu().numpy().view(np.uint64)[0]
    except Exception as e:                                      # arkworks: SerializationError::InvalidData -> Generic(err.to_string())
        raise MpcNetError("Generic", "the input buffer contained invalid data") from e


# ---- what one call puts on / takes off the wire ----------------------------------------------------------------------------------
def client_message(payload: bytes, prod: bool = False) -> bytes:
    """the bytes a client writes on stream `sid` towards the king for one `send_to_king` (ProdNet: wrapped in a packet)"""
    return frame(prod_packet(payload) if prod else payload)


def read_message(stream: bytes, prod: bool = False):
    """inverse of client_message on a byte stream: -> (payload, rest) or (None, stream) when incomplete"""
    body, rest = unframe(stream)
    if body is None:
        return None, stream
    if prod:
        variant, body = parse_prod_packet(body)
        if variant != PACKET_PACKET:
            raise MpcNetError("Protocol", "unexpected handshake packet")
    return body, rest


def king_scatter(payloads, prod: bool = False):
    """`client_receive_or_king_send` (mpc-net/src/lib.rs:106-139): one equally long payload per party; -> (king's own payload,
    {party id: framed bytes}) -- the reference rejects unequal lengths with MpcNetError::Protocol."""
    m = len(payloads[0])
    for pid, p in enumerate(payloads):
        if len(p) != m:
            raise MpcNetError("Protocol", "Peer %d sent wrong number of bytes" % pid)
    return payloads[0], {pid: client_message(p, prod) for pid, p in enumerate(payloads) if pid != 0}
