"""Groth16 verification on the GPU (csrc/verify.cu) against the reference's snarkjs fixtures and the oracle's pairing.

The reference verifies every proof it makes (`verify_with_processed_vk`, groth16/examples/sha256.rs:229-254); its tree
holds snarkjs-made (vk, public, proof) triples (fixtures/million/*.json -> tests/golden/reference_goldens.json)."""
import json
import os

import numpy as np
import pytest

from distributed_groth16_b200 import B200zkError, ark_serialize as ark
from distributed_groth16_b200.groth16 import verify

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _snarkjs_million():
    from oracle import layout
    s = json.load(open(os.path.join(G, "reference_goldens.json")))["snarkjs_million"]
    g1 = lambda v: (int(v[0]), int(v[1]))
    g2 = lambda v: ((int(v[0][0]), int(v[0][1])), (int(v[1][0]), int(v[1][1])))
    vk, pr = s["vk"], s["proof"]
    avk = ark.ArkVerifyingKey(layout.g1_to_arr([g1(vk["vk_alpha_1"])])[0], layout.g2_to_arr([g2(vk["vk_beta_2"])])[0],
                              layout.g2_to_arr([g2(vk["vk_gamma_2"])])[0], layout.g2_to_arr([g2(vk["vk_delta_2"])])[0],
                              layout.g1_to_arr([g1(x) for x in vk["IC"]]))
    proof = (layout.g1_to_arr([g1(pr["pi_a"])])[0], layout.g2_to_arr([g2(pr["pi_b"])])[0], layout.g1_to_arr([g1(pr["pi_c"])])[0])
    return avk, [int(x) for x in s["public"]], proof


def test_reference_snarkjs_fixture_verifies_and_tampering_is_rejected(net):
    from oracle import bn254 as o, layout
    vk, pub, proof = _snarkjs_million()
    assert verify.verify_proof(net, vk, layout.fr_to_arr(pub), proof)
    assert not verify.verify_proof(net, vk, layout.fr_to_arr([pub[0] + 1]), proof)
    a, b, c = proof
    assert not verify.verify_proof(net, vk, layout.fr_to_arr(pub), (c, b, a))
    neg_a = layout.g1_to_arr([o.G1.neg(layout.arr_to_g1(a.reshape(1, -1))[0])])[0]
    assert not verify.verify_proof(net, vk, layout.fr_to_arr(pub), (neg_a, b, c))
    # same proof as 128 compressed bytes (what zk-cli writes), decoded by the GPU codec
    blob = ark.serialize_proof(net, a, b, c)
    assert len(blob) == 128 and verify.verify_proof(net, vk, layout.fr_to_arr(pub), blob)
    with pytest.raises(ValueError):
        verify.verify_proof(net, vk, layout.fr_to_arr(pub + [1]), proof)
    with pytest.raises(B200zkError):
        verify.verify_proof(net, vk, layout.fr_to_arr(pub), (4).to_bytes(32, "little") + blob[32:])   # A.x = 4 is off the curve


def test_golden_proofs_of_the_f1_zkey_verify_for_r_s_zero_and_nonzero(net):
    """Proof bytes committed under tests/golden (GPU == oracle == these) against the zkey's own verifying key."""
    from oracle import bn254 as o, layout
    d = np.load(os.path.join(G, "complex_circuit.zkey.pk.npz"))
    exp = json.load(open(os.path.join(G, "complex_circuit_proof.json")))
    n_vars = int(d["dims"][0])
    z = 3
    for _ in range(3, n_vars):
        z = z * z % o.R
    pub = z * z % o.R                                       # z[1] = z[n_vars-1]^2 (see test_gpu_prove.py)
    vk = ark.ArkVerifyingKey(d["vk_g1"][0], d["vk_g2"][0], d["vk_g2"][2], d["vk_g2"][1], d["ic"])
    for key in ("r0s0", "r_s"):
        blob = bytes.fromhex(exp[key]["proof_hex"])
        assert verify.verify_proof(net, vk, layout.fr_to_arr([pub]), blob), key
        assert not verify.verify_proof(net, vk, layout.fr_to_arr([pub + 1]), blob)


def test_degenerate_inputs(net):
    """Identity points make the corresponding pairing 1 (arkworks' behaviour); zero public inputs use IC_0 alone."""
    from oracle import layout
    vk, pub, (a, b, c) = _snarkjs_million()
    zero1, zero2 = np.zeros(8, dtype=np.uint64), np.zeros(16, dtype=np.uint64)
    assert not verify.verify_proof(net, vk, layout.fr_to_arr(pub), (zero1, b, c))
    assert not verify.verify_proof(net, vk, layout.fr_to_arr(pub), (a, zero2, zero1))
    vk0 = ark.ArkVerifyingKey(vk.alpha_g1, vk.beta_g2, vk.gamma_g2, vk.delta_g2, vk.gamma_abc_g1[:1])
    assert not verify.verify_proof(net, vk0, [], (a, b, c))


def test_raw_limb_proofs_are_validated_like_the_byte_form(net):
    """(A, B, C) given as limb arrays: an off-curve point is rejected (the byte form rejects it at decompression)."""
    from oracle import layout
    vk, pub, (a, b, c) = _snarkjs_million()
    assert verify.verify_proof(net, vk, layout.fr_to_arr(pub), (a, b, c))
    bad = np.array(a, dtype=np.uint64, copy=True).reshape(-1)
    bad[4] ^= np.uint64(1)                                   # y changed: no longer on y^2 = x^3 + 3
    with pytest.raises(ValueError):
        verify.verify_proof(net, vk, layout.fr_to_arr(pub), (bad, b, c))
