"""Groth16 verification on the GPU (csrc/verify.cu, csrc/pairing.cuh).

Mirror of the call the reference makes after every proof, `Groth16::<Bn254>::verify_with_processed_vk(&pvk, &inputs, &proof)`
(groth16/examples/sha256.rs:229-254, mpc-api/src/main.rs:187-247): True iff
e(A, B) = e(alpha, beta) e(IC_0 + sum_i x_i IC_{i+1}, gamma) e(C, delta)."""
from __future__ import annotations

import ctypes

import numpy as np

from ..ark_serialize import ArkVerifyingKey, deserialize_proof, serialize_proof
from ..context import Net, _as_u64, _ptr


def verify_proof(net: Net, vk: ArkVerifyingKey, public_inputs, proof) -> bool:
    """vk: ark_serialize.ArkVerifyingKey (Montgomery limb arrays); public_inputs: (n, 4) u64 Montgomery Fr elements
    (the witness entries z[1..n_public]); proof: the 128 compressed bytes, or (A, B, C) limb arrays.
    Raises B200zkError when the proof bytes are not curve points (arkworks: deserialisation error, not `false`)."""
    if isinstance(proof, (bytes, bytearray, memoryview)):
        a, b, c = deserialize_proof(net, bytes(proof), check_subgroup=True)
    else:
        # raw limb tuples get the validation the byte form gets for free: a round trip through the compressed encoding
        # rebuilds y from x (on-curve) and runs the G2 subgroup check; anything that does not come back unchanged is rejected
        a, b, c = proof
        a, c = _as_u64(a, 8).reshape(-1), _as_u64(c, 8).reshape(-1)
        b = _as_u64(b, 16).reshape(-1)
        a2, b2, c2 = deserialize_proof(net, serialize_proof(net, a, b, c), check_subgroup=True)
        if not ((a2 == a).all() and (b2 == b).all() and (c2 == c).all()):
            raise ValueError("proof points are not valid curve points")
    a, c = _as_u64(a, 8).reshape(-1), _as_u64(c, 8).reshape(-1)
    b = _as_u64(b, 16).reshape(-1)
    ic = _as_u64(vk.gamma_abc_g1, 8)
    x = _as_u64(public_inputs, 4) if len(public_inputs) else np.zeros((0, 4), dtype=np.uint64)
    if ic.shape[0] != x.shape[0] + 1:
        raise ValueError("verifying key expects %d public inputs, got %d" % (ic.shape[0] - 1, x.shape[0]))
    ok = ctypes.c_int(0)
    net.check(net._lib.b200zk_groth16_verify(net._h, _ptr(_as_u64(vk.alpha_g1, 8)), _ptr(_as_u64(vk.beta_g2, 16)),
                                             _ptr(_as_u64(vk.gamma_g2, 16)), _ptr(_as_u64(vk.delta_g2, 16)), _ptr(ic),
                                             x.shape[0], _ptr(x) if x.shape[0] else None, _ptr(a), _ptr(b), _ptr(c),
                                             ctypes.byref(ok)))
    return bool(ok.value)
