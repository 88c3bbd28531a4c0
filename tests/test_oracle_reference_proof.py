"""BASELINE config 4, absolutely pinned: the reference's committed proof `zk-cli/test-circuits/sha256/proof.bin`
(tests/golden/sha256_proof.bin) is regenerated bit for bit on the CPU from nothing but the seed `[42u8; 32]`
(mpc-api/src/main.rs:148-152), the reference's sha256 circuit (fixtures/sha256/sha256.r1cs -> tests/golden/sha256_circuit.npz)
and the witness of {a: 1, b: 2} -- oracle/ark_rand.py restates rand's StdRng and arkworks' sampling / setup order, the
C++ twin builds the proving key and proves.  Everything the prover path computes (QAP, h, the five MSMs, assembly,
compression) is thereby checked against an output of the reference itself."""
import os
import struct

import numpy as np

from oracle import ark_rand as ar, bn254 as o, layout

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SEED = bytes([42] * 32)


def test_chacha_block_function_matches_rfc7539_zero_key_vector():
    blk = struct.pack("<16I", *ar.chacha_block((0,) * 8, 0, 0, 20)).hex()
    assert blk == ("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
                   "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")


def reference_sha256_instance(cref):
    """-> (pk arrays, vk points dict, z, a, b, c as Montgomery limb arrays, dims) of the reference's seeded setup."""
    from oracle import reference_instance
    return reference_instance.sha256_instance(cref, G, SEED)


def test_reference_proof_bin_is_reproduced_from_the_seed(cref):
    gold = open(os.path.join(G, "sha256_proof.bin"), "rb").read()
    pk, vk, z, a, b, c, (n_vars, n_inputs, m) = reference_sha256_instance(cref)
    zero = np.zeros(4, dtype=np.uint64)
    vk_pts = np.concatenate([vk["alpha_g1"], vk["beta_g1"], vk["delta_g1"], vk["beta_g2"], vk["delta_g2"]])
    got = cref.groth16_prove(pk["a_query"], pk["b_g1_query"], pk["b_g2_query"], pk["l_query"], pk["h_query"], vk_pts, n_inputs,
                             z, cref.h_circom(a, b, c), zero, zero, mirror_bg1=False)
    assert got == gold
    # and it verifies under the regenerated verifying key (zk-cli/README.md:82 reports is_valid: true)
    A, B, C = o.proof_decompress(gold)
    g1p, g2p = (lambda x: layout.arr_to_g1(x.reshape(1, -1))[0]), (lambda x: layout.arr_to_g2(x.reshape(1, -1))[0])
    pub = int.from_bytes(bytes(np.load(os.path.join(G, "sha256_circuit.npz"))["witness"][1].tobytes()), "little")
    assert o.groth16_verify(g1p(vk["alpha_g1"]), g2p(vk["beta_g2"]), g2p(vk["gamma_g2"]), g2p(vk["delta_g2"]),
                            layout.arr_to_g1(vk["gamma_abc_g1"]), [pub], A, B, C)
