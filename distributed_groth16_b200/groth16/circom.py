"""Circom / snarkjs artefacts -> device-resident prover inputs, and the one-call prover on top of them
(the flow of /root/reference/groth16/examples/sha256.rs:127-169 and mpc-api/src/main.rs:282-421 minus the
WASM witness calculator: the witness comes from a .wtns file or from the caller)."""
from __future__ import annotations

import numpy as np

from .. import formats
from ..context import Net
from . import prove
from .proving_key import ProvingKey
from .qap import ConstraintMatrices, qap


def load_zkey(net: Net, zkey_bytes: bytes):
    """-> (ProvingKey resident in HBM, ConstraintMatrices, formats.ZKey)   (ark-circom/src/zkey.rs:53-60)."""
    zk = formats.read_zkey(zkey_bytes)
    pk = ProvingKey(net, zk.a_query, zk.b_g1_query, zk.b_g2_query, zk.l_query, zk.h_query, zk.n_inputs, zk.alpha_g1,
                    zk.beta_g1, zk.delta_g1, zk.beta_g2, zk.delta_g2)
    coo = []
    for mi in (0, 1):
        sel = zk.coef_matrix == mi
        coo.append((zk.coef_row[sel], zk.coef_col[sel], zk.coef_val_r2[sel]))
    mats = ConstraintMatrices(net, zk.n_inputs, zk.num_constraints, coo[0], coo[1], values_montgomery_depth=1)
    return pk, mats, zk


def load_witness(net: Net, wtns_bytes: bytes):
    """.wtns -> full assignment z on the device in Montgomery form (wire index == witness index: the reference
    disables the wire mapping, ark-circom/src/circom/builder.rs:63-64)."""
    w = formats.read_wtns(wtns_bytes)
    return net.fr_convert(net.to_device(w), to_mont=True)


def witness_from_ints(net: Net, values):
    """canonical Python ints -> Montgomery device tensor (conversion on the GPU)."""
    arr = np.array([[(int(v) >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for v in values], dtype=np.uint64)
    return net.fr_convert(net.to_device(arr), to_mont=True)


def prove_from_matrices(pk: ProvingKey, matrices: ConstraintMatrices, z, r=None, s=None, mirror_reference_bg1=False) -> bytes:
    """qap() -> h -> MSMs -> 128 proof bytes, everything resident on the GPU
    (== Groth16::create_proof_with_reduction_and_matrices + serialize, sha256.rs:159-168)."""
    q = qap(matrices, z, pk.net)
    return prove.create_proof_dev(pk, z, q.a, q.b, q.c, r, s, mirror_reference_bg1)


def prove_zkey_wtns(net: Net, zkey_bytes: bytes, wtns_bytes: bytes, r=None, s=None):
    """-> (proof bytes, public inputs as canonical limbs (n_public, 4))."""
    pk, mats, zk = load_zkey(net, zkey_bytes)
    w = formats.read_wtns(wtns_bytes)
    if w.shape[0] != zk.n_vars:
        raise formats.FormatError("witness has %d entries, the key expects %d" % (w.shape[0], zk.n_vars))
    z = net.fr_convert(net.to_device(w), to_mont=True)
    proof = prove_from_matrices(pk, mats, z, r, s)
    pk.free()
    return proof, w[1:1 + zk.n_public].copy()
