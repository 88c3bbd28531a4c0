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
    d = np.load(os.path.join(G, "sha256_circuit.npz"))
    n_vars, n_pub, n_cons = (int(x) for x in d["dims"])
    n_inputs = n_pub + 1
    m = 1
    while m < n_cons + n_inputs:
        m <<= 1
    ints = lambda arr: [int.from_bytes(arr[i].tobytes(), "little") for i in range(arr.shape[0])]
    z = ints(d["witness"])
    coo = lambda k: (d[k + "_rows"], d[k + "_cols"], ints(d[k + "_vals"]))
    tw = ar.groth16_toxic_waste(SEED, m)
    qs = ar.query_scalars(tw, coo("a"), coo("b"), coo("c"), n_vars, n_inputs, n_cons, m)
    g1, g2 = layout.g1_to_arr([tw["g1"]])[0], layout.g2_to_arr([tw["g2"]])[0]
    fb = lambda sc, g2_=False: cref.fixed_base_mul(g2 if g2_ else g1, layout.fr_to_arr(sc), g2=g2_)
    pk = dict(a_query=fb(qs["a"]), b_g1_query=fb(qs["b"]), b_g2_query=fb(qs["b"], True), l_query=fb(qs["l"]), h_query=fb(qs["h"]))
    vk1 = fb([tw["alpha"], tw["beta"], tw["delta"]])                       # alpha_g1, beta_g1, delta_g1
    vk2 = fb([tw["beta"], tw["delta"], tw["gamma"]], True)                 # beta_g2, delta_g2, gamma_g2
    vk = dict(alpha_g1=vk1[0], beta_g1=vk1[1], delta_g1=vk1[2], beta_g2=vk2[0], delta_g2=vk2[1], gamma_g2=vk2[2],
              gamma_abc_g1=fb(qs["gamma_abc"]))
    # QAP evaluation vectors (groth16/src/qap.rs:44-91): a = A z with z_i on the input rows, b = B z, c = a . b
    a, b = [0] * m, [0] * m
    for vec, (rows, cols, vals) in ((a, coo("a")), (b, coo("b"))):
        for r_, c_, v in zip(rows, cols, vals):
            vec[int(r_)] = (vec[int(r_)] + v * z[int(c_)]) % o.R
    for i in range(n_inputs):
        a[n_cons + i] = z[i]
    c = [x * y % o.R for x, y in zip(a, b)]
    return pk, vk, layout.fr_to_arr(z), layout.fr_to_arr(a), layout.fr_to_arr(b), layout.fr_to_arr(c), (n_vars, n_inputs, m)


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
