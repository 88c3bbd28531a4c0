"""Pins the oracle on every golden vector the reference holds for the hot path."""
import json
import os

import numpy as np

from oracle import bn254 as o, layout

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
gold = json.load(open(os.path.join(G, "reference_goldens.json")))


def test_bn_parameters_and_fixture_primes():
    assert o.P == 36 * o.BN_U**4 + 36 * o.BN_U**3 + 24 * o.BN_U**2 + 6 * o.BN_U + 1
    assert o.R == o.P + 1 - (6 * o.BN_U**2 + 1)                       # #E(Fp) = p + 1 - t
    assert int(gold["sha256_r1cs_header"]["prime"]) == o.R            # r1cs_reader.rs:180-188
    assert gold["sha256_r1cs_header"]["n_wires"] == 29823
    assert pow(o.FR_GENERATOR, (o.R - 1) // 2, o.R) == o.R - 1        # 5 is a non-residue => generator-compatible
    w = o.fr_root_of_unity(1 << 28)
    assert pow(w, 1 << 27, o.R) == o.R - 1 and (o.R - 1) % (1 << 28) == 0 and (o.R - 1) % (1 << 29) != 0
    assert w == 19103219067921713944291392827692070036145651957329286315305642004821462161904


def test_montgomery_byte_goldens_from_zkey_rs():
    """ark-circom/src/zkey.rs:417-455: snarkjs toRprLEM of F.one, G1.one, G2.one."""
    assert bytes(gold["fq_one_mont_bytes"]) == o.fq_mont(1).to_bytes(32, "little")
    g1 = bytes(gold["g1_gen_mont_bytes"])
    assert g1 == layout.g1_to_arr([o.G1_GEN]).tobytes()
    g2 = bytes(gold["g2_gen_mont_bytes"])
    assert g2 == layout.g2_to_arr([o.G2_GEN]).tobytes()
    assert o.G1.is_on_curve(o.G1_GEN) and o.G2.is_on_curve(o.G2_GEN)
    assert o.G2.from_jac(o.G2.jac_mul(o.G2.to_jac(o.G2_GEN), o.R)) is None


def test_sha256_proof_bin_decodes_to_printed_coordinates():
    """zk-cli/test-circuits/sha256/proof.bin vs the decimal coordinates in zk-cli/README.md:82:
    pins x endianness, the y-sign flag for Fq, and the (c1, c0) ordering rule for Fq2."""
    pb = open(os.path.join(G, "sha256_proof.bin"), "rb").read()
    A, B, C = o.proof_decompress(pb)
    c = gold["sha256_proof_coords"]
    assert A == (int(c["a"][0]), int(c["a"][1]))
    assert B == ((int(c["b"][0][0]), int(c["b"][0][1])), (int(c["b"][1][0]), int(c["b"][1][1])))
    assert C == (int(c["c"][0]), int(c["c"][1]))
    assert o.proof_compress(A, B, C) == pb
    assert o.G1.is_on_curve(A) and o.G2.is_on_curve(B) and o.G1.is_on_curve(C)
    assert int(gold["sha256_public_input"]) == 72587776472194017031617589674261467945970986113287823188107011979


def test_snarkjs_proof_verifies_under_oracle_pairing():
    s = gold["snarkjs_million"]
    g1 = lambda v: (int(v[0]), int(v[1]))
    g2 = lambda v: ((int(v[0][0]), int(v[0][1])), (int(v[1][0]), int(v[1][1])))
    vk, pr = s["vk"], s["proof"]
    args = (g1(vk["vk_alpha_1"]), g2(vk["vk_beta_2"]), g2(vk["vk_gamma_2"]), g2(vk["vk_delta_2"]),
            [g1(x) for x in vk["IC"]])
    pub = [int(x) for x in s["public"]]
    assert o.groth16_verify(*args, pub, g1(pr["pi_a"]), g2(pr["pi_b"]), g1(pr["pi_c"]))
    assert not o.groth16_verify(*args, [pub[0] + 1], g1(pr["pi_a"]), g2(pr["pi_b"]), g1(pr["pi_c"]))
    e = o.pairing(o.G1_GEN, o.G2_GEN)
    assert o.fq12_pow(e, 35) == o.pairing(o.G1.mul(o.G1_GEN, 5), o.G2.mul(o.G2_GEN, 7)) and e != o.FQ12_ONE


def _load_f1():
    d = np.load(os.path.join(G, "complex_circuit.zkey.pk.npz"))
    n_vars, n_public, m, nc = (int(x) for x in d["dims"])
    z = [0] * n_vars
    z[0], z[2] = 1, 3
    for i in range(3, n_vars):
        z[i] = z[i - 1] * z[i - 1] % o.R
    z[1] = z[n_vars - 1] ** 2 % o.R

    def rows(r, c, v):
        vals = layout.arr_to_fr(v)
        out = [[] for _ in range(nc)]
        for i, w, x in zip(r, c, vals):
            out[int(i)].append((x, int(w)))
        return out

    ma, mb = rows(d["a_rows"], d["a_cols"], d["a_vals"]), rows(d["b_rows"], d["b_cols"], d["b_vals"])
    return d, n_vars, n_public, m, nc, z, ma, mb


def test_f1_real_zkey_proof_reproduced_by_cpu_twin_and_verifies(cref):
    """Fixture F1: pk from the snarkjs-made complex-circuit-10000-10000.zkey, witness a = 3.
    The committed proofs were produced by oracle/bn254.py and verified with its pairing
    (tests/golden/make_golden.py); here the C++ twin must reproduce the same 128 bytes and the
    r = s = 0 proof is re-verified against the zkey's vk."""
    d, n_vars, n_public, m, nc, z, ma, mb = _load_f1()
    exp = json.load(open(os.path.join(G, "complex_circuit_proof.json")))
    assert int(exp["public_input"]) == z[1]
    qa, qb, qc = o.qap(ma, mb, n_public + 1, nc, z)
    a, b, c = layout.fr_to_arr(qa), layout.fr_to_arr(qb), layout.fr_to_arr(qc)
    h = cref.h_circom(a, b, c)
    assert layout.arr_to_fr(h[:64]) == o.h_circom(qa, qb, qc)[:64]
    vk = np.concatenate([d["vk_g1"].reshape(-1), d["vk_g2"][:2].reshape(-1)])
    zz = layout.fr_to_arr(z)
    for key in ("r0s0", "r_s"):
        r, s = layout.fr_to_arr([exp[key]["r"]])[0], layout.fr_to_arr([exp[key]["s"]])[0]
        got = cref.groth16_prove(d["a_query"], d["b_g1_query"], d["b_g2_query"], d["l_query"], d["h_query"], vk,
                                 n_public + 1, zz, h, r, s)
        assert got.hex() == exp[key]["proof_hex"]
    A, B, C = o.proof_decompress(bytes.fromhex(exp["r0s0"]["proof_hex"]))
    vk1, vk2 = layout.arr_to_g1(d["vk_g1"]), layout.arr_to_g2(d["vk_g2"])
    assert o.groth16_verify(vk1[0], vk2[0], vk2[2], vk2[1], layout.arr_to_g1(d["ic"]), [z[1]], A, B, C)


def test_f2_sha256_witness_fixture_satisfies_the_r1cs_and_the_reference_kat():
    """Fixture F2: witness of fixtures/sha256 for {a: 1, b: 2} (computed by oracle/wasm_witness.py from the reference's
    own sha256.wasm).  witness[1] must equal the public output asserted in groth16/examples/sha256.rs:231-233 and every
    constraint of sha256.r1cs must hold."""
    d = np.load(os.path.join(G, "sha256_circuit.npz"))
    n_wires, n_pub, n_cons = (int(x) for x in d["dims"])
    assert (n_wires, n_pub, n_cons) == (29823, 1, 30134)                      # SURVEY 8a sizes
    w = [int.from_bytes(r.tobytes(), "little") for r in d["witness"]]
    assert w[0] == 1 and w[1] == int(gold["sha256_public_input"])
    acc = {}
    for k in "abc":
        vals = [int.from_bytes(r.tobytes(), "little") for r in d[k + "_vals"]]
        v = [0] * n_cons
        for r_, c_, x in zip(d[k + "_rows"], d[k + "_cols"], vals):
            v[int(r_)] = (v[int(r_)] + x * w[int(c_)]) % o.R
        acc[k] = v
    assert all(x * y % o.R == z for x, y, z in zip(acc["a"], acc["b"], acc["c"]))


def test_pairing_value_matches_the_snarkjs_vk_alphabeta_12():
    """Absolute known answer for the pairing itself: snarkjs stores e(alpha, beta) in the verification key
    (fixtures/million/verification_key.json, `vk_alphabeta_12`, Fq12 as the 2 x 3 x 2 tower over w^2 = v, v^3 = 9 + u).
    ffjavascript's final exponentiation uses the Fuentes-Castaneda hard part, which raises to m * (p^4 - p^2 + 1)/r with
    m = 2u(6u^2 + 3u + 1), so the stored value is e(alpha, beta)^m for the plain (p^12 - 1)/r pairing of the oracle."""
    vk = gold["snarkjs_million"]["vk"]
    g1 = lambda v: (int(v[0]), int(v[1]))
    g2 = lambda v: ((int(v[0][0]), int(v[0][1])), (int(v[1][0]), int(v[1][1])))
    e = o.pairing(g1(vk["vk_alpha_1"]), g2(vk["vk_beta_2"]))
    ab = vk["vk_alphabeta_12"]
    want = [0] * 12
    for (h, k), i in zip(((0, 0), (0, 1), (0, 2), (1, 0), (1, 1), (1, 2)), (0, 2, 4, 1, 3, 5)):   # coefficient of w^i
        x, y = int(ab[h][k][0]), int(ab[h][k][1])                                               # (x + y u) w^i, u = w^6 - 9
        want[i] = (want[i] + x - 9 * y) % o.P
        want[i + 6] = (want[i + 6] + y) % o.P
    u = o.BN_U
    assert o.fq12_pow(e, 2 * u * (6 * u * u + 3 * u + 1)) == tuple(want)
    assert o.fq12_pow(tuple(want), o.R) == o.FQ12_ONE and tuple(want) != o.FQ12_ONE
