"""ext_wit::h and the full prover vs the oracle (dummy CRS, as the reference's own bench builds it:
groth16/examples/local_groth_bench.rs:21-52, groth16/src/proving_key.rs:112-155)."""
import numpy as np
import pytest

from distributed_groth16_b200.groth16 import ProvingKey, ext_wit, prove
from distributed_groth16_b200.groth16.qap import PackedQAPShare, Radix2Domain

pytestmark = pytest.mark.gpu


def _dummy_instance(cref, m, n_vars, n_inputs, seed):
    aq = cref.g1_generate(seed + 1, n_vars)
    b1 = cref.g1_generate(seed + 2, n_vars)
    b2 = cref.g2_generate(seed + 3, n_vars)
    lq = cref.g1_generate(seed + 4, n_vars - n_inputs)
    hq = cref.g1_generate(seed + 5, m)
    vk1 = cref.g1_generate(seed + 6, 3)
    vk2 = cref.g2_generate(seed + 7, 2)
    aq[3] = 0; b1[0] = 0; b2[5] = 0                      # real pks contain points at infinity (zkey.rs:633-674)
    z = cref.fr_generate(seed + 8, n_vars)
    from oracle import layout
    z[0] = layout.fr_to_arr([1])[0]
    a = cref.fr_generate(seed + 9, m)
    b = cref.fr_generate(seed + 10, m)
    c = cref.fr_generate(seed + 11, m)
    return aq, b1, b2, lq, hq, vk1, vk2, z, a, b, c


@pytest.mark.parametrize("log_m", [0, 1, 3, 8, 12, 15])
def test_h_matches_oracle_and_literal_reference_flow(net, cref, log_m):
    m = 1 << log_m
    a, b, c = (cref.fr_generate(s, m) for s in (1, 2, 3))
    exp = cref.h_circom(a, b, c)
    share = PackedQAPShare(1, m - 1, a, b, c, Radix2Domain(m), rearranged=False)
    assert (ext_wit.h(share, None, net) == exp).all()
    if log_m <= 12:
        assert (ext_wit.h_via_d_fft(share, None, net) == exp).all()


@pytest.mark.parametrize("rs", [(0, 0), (12345, 67890)])
@pytest.mark.parametrize("mirror", [False, True])
def test_prove_bytes_match_oracle(net, cref, rs, mirror):
    from oracle import layout
    m, n_vars, n_inputs = 1 << 10, 900, 2
    aq, b1, b2, lq, hq, vk1, vk2, z, a, b, c = _dummy_instance(cref, m, n_vars, n_inputs, 100)
    r, s = layout.fr_to_arr([rs[0]])[0], layout.fr_to_arr([rs[1]])[0]
    vk = np.concatenate([vk1.reshape(-1), vk2.reshape(-1)])
    h = cref.h_circom(a, b, c)
    exp = cref.groth16_prove(aq, b1, b2, lq, hq, vk, n_inputs, z, h, r, s, mirror_bg1=mirror)
    pk = ProvingKey(net, aq, b1, b2, lq, hq, n_inputs, vk1[0], vk1[1], vk1[2], vk2[0], vk2[1])
    got = prove.create_proof(pk, z, a, b, c, r, s, mirror_reference_bg1=mirror)
    assert got == exp
    pk.free()


def test_prove_bytes_do_not_depend_on_the_fixed_base_tables(net, cref):
    """The key's window tables (b200zk_pk_precompute) are an HBM-for-time trade: same 128 bytes with automatic
    windows, a forced small window, and no tables at all."""
    from oracle import layout
    m, n_vars, n_inputs = 1 << 11, 2500, 3
    aq, b1, b2, lq, hq, vk1, vk2, z, a, b, c = _dummy_instance(cref, m, n_vars, n_inputs, 300)
    r, s = layout.fr_to_arr([777])[0], layout.fr_to_arr([999])[0]
    vk = np.concatenate([vk1.reshape(-1), vk2.reshape(-1)])
    exp = cref.groth16_prove(aq, b1, b2, lq, hq, vk, n_inputs, z, cref.h_circom(a, b, c), r, s, mirror_bg1=False)
    pk = ProvingKey(net, aq, b1, b2, lq, hq, n_inputs, vk1[0], vk1[1], vk1[2], vk2[0], vk2[1])
    assert pk.table_bytes > 0                              # built by the upload
    assert prove.create_proof(pk, z, a, b, c, r, s) == exp
    assert pk.precompute(7) > 0
    assert prove.create_proof(pk, z, a, b, c, r, s) == exp
    assert pk.precompute(None) == 0
    assert prove.create_proof(pk, z, a, b, c, r, s) == exp
    pk.free()


def test_prove_structs_mirror_reference_call_pattern(net, cref):
    """groth16/examples/sha256.rs:45-88 + :208-212 with r = s = 0, assembled from A/B/C structs."""
    from oracle import layout
    m, n_vars, n_inputs = 1 << 8, 200, 2
    aq, b1, b2, lq, hq, vk1, vk2, z, a, b, c = _dummy_instance(cref, m, n_vars, n_inputs, 500)
    zero = np.zeros(4, dtype=np.uint64)
    h = ext_wit.h(PackedQAPShare(n_inputs, m - n_inputs, a, b, c, Radix2Domain(m)), None, net)
    # L/Z = vk.alpha + a_query[0] etc. are added by the driver afterwards (sha256.rs:208-212); fold via extra adds
    A = prove.A(L=aq[0], N=vk1[2], r=zero, pp=None, S=aq[1:], a=z[1:]).compute(net)
    B = prove.B(Z=b2[0], K=vk2[1], s=zero, pp=None, V=b2[1:], a=z[1:]).compute(net)
    C = prove.C(A=A, M=vk1[2], s=zero, r=zero, pp=None, W=lq, U=hq, H=b1[1:], a=z[1:], ax=z[n_inputs:], h=h).compute(net)
    one = layout.fr_to_arr([1])[0]
    # a += alpha ; b += beta2   (sha256.rs:208-212)
    A2, _ = cref.msm_g1(np.stack([A.limbs, vk1[0]]), np.stack([one, one]))
    B2, _ = cref.msm_g2(np.stack([B.limbs, vk2[0]]), np.stack([one, one]))
    vk = np.concatenate([vk1.reshape(-1), vk2.reshape(-1)])
    exp = cref.groth16_prove(aq, b1, b2, lq, hq, vk, n_inputs, z, cref.h_circom(a, b, c), zero, zero)
    from oracle import bn254 as o
    got = o.proof_compress(layout.arr_to_g1(A2)[0], layout.arr_to_g2(B2)[0], layout.arr_to_g1(C.limbs)[0])
    assert got == exp


def test_prove_structs_with_nonzero_r_and_s(net, cref):
    """prove.rs:36-44,75-83,128-134 with r, s != 0: A = L + r N + MSM, B = Z + s K + MSM, C = w + u + s A + r M + r MSM(H, a)
    (M = beta_g1 + b_g1_query[0], A = the finished A) reproduces the randomised proof of the fused prover / the oracle."""
    from oracle import layout, bn254 as o
    m, n_vars, n_inputs = 1 << 8, 200, 2
    aq, b1, b2, lq, hq, vk1, vk2, z, a, b, c = _dummy_instance(cref, m, n_vars, n_inputs, 510)
    r, s = layout.fr_to_arr([123456789])[0], layout.fr_to_arr([987654321])[0]
    one = layout.fr_to_arr([1])[0]
    h = ext_wit.h(PackedQAPShare(n_inputs, m - n_inputs, a, b, c, Radix2Domain(m)), None, net)
    A = prove.A(L=aq[0], N=vk1[2], r=r, pp=None, S=aq[1:], a=z[1:]).compute(net)
    B = prove.B(Z=b2[0], K=vk2[1], s=s, pp=None, V=b2[1:], a=z[1:]).compute(net)
    A2, _ = cref.msm_g1(np.stack([A.limbs, vk1[0]]), np.stack([one, one]))          # a += alpha  (sha256.rs:208-212)
    B2, _ = cref.msm_g2(np.stack([B.limbs, vk2[0]]), np.stack([one, one]))          # b += beta_g2
    M, _ = cref.msm_g1(np.stack([vk1[1], b1[0]]), np.stack([one, one]))             # beta_g1 + b_g1_query[0]
    C = prove.C(A=A2, M=M, s=s, r=r, pp=None, W=lq, U=hq, H=b1[1:], a=z[1:], ax=z[n_inputs:], h=h).compute(net)
    vk = np.concatenate([vk1.reshape(-1), vk2.reshape(-1)])
    exp = cref.groth16_prove(aq, b1, b2, lq, hq, vk, n_inputs, z, cref.h_circom(a, b, c), r, s)
    got = o.proof_compress(layout.arr_to_g1(A2)[0], layout.arr_to_g2(B2)[0], layout.arr_to_g1(C.limbs)[0])
    assert got == exp


def test_f1_real_zkey_prove_on_gpu_matches_golden_and_verifies(net):
    """Fixture F1: proving key of the snarkjs-made complex-circuit-10000-10000.zkey, witness a = 3.
    The GPU proof must equal the committed golden bytes (oracle-produced, pairing-verified against the zkey's vk)
    for r = s = 0 and for r, s != 0; the r = s = 0 proof is verified again here."""
    import json
    import os
    from oracle import bn254 as o, layout
    G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    d = np.load(os.path.join(G, "complex_circuit.zkey.pk.npz"))
    exp = json.load(open(os.path.join(G, "complex_circuit_proof.json")))
    n_vars, n_public, m, nc = (int(x) for x in d["dims"])
    z = [0] * n_vars
    z[0], z[2] = 1, 3
    for i in range(3, n_vars):
        z[i] = z[i - 1] * z[i - 1] % o.R
    z[1] = z[n_vars - 1] ** 2 % o.R
    # QAP rows of this chain circuit: A_i = -w_{i+2}, B_i = w_{i+2} (read from the zkey's coefficient section)
    vals_a, vals_b = layout.arr_to_fr(d["a_vals"]), layout.arr_to_fr(d["b_vals"])
    a = [0] * m
    b = [0] * m
    for r_, c_, v in zip(d["a_rows"], d["a_cols"], vals_a):
        a[int(r_)] = (a[int(r_)] + v * z[int(c_)]) % o.R
    for r_, c_, v in zip(d["b_rows"], d["b_cols"], vals_b):
        b[int(r_)] = (b[int(r_)] + v * z[int(c_)]) % o.R
    for j in range(n_public + 1):
        a[nc + j] = z[j]
    c = [x * y % o.R for x, y in zip(a, b)]
    pk = ProvingKey(net, d["a_query"], d["b_g1_query"], d["b_g2_query"], d["l_query"], d["h_query"], n_public + 1,
                    d["vk_g1"][0], d["vk_g1"][1], d["vk_g1"][2], d["vk_g2"][0], d["vk_g2"][1])
    zz, aa, bb, cc = (layout.fr_to_arr(v) for v in (z, a, b, c))
    for key in ("r0s0", "r_s"):
        r, s = layout.fr_to_arr([exp[key]["r"]])[0], layout.fr_to_arr([exp[key]["s"]])[0]
        got = prove.create_proof(pk, zz, aa, bb, cc, r, s)
        assert got.hex() == exp[key]["proof_hex"], key
    A, B, C = o.proof_decompress(prove.create_proof(pk, zz, aa, bb, cc))
    vk1, vk2 = layout.arr_to_g1(d["vk_g1"]), layout.arr_to_g2(d["vk_g2"])
    assert o.groth16_verify(vk1[0], vk2[0], vk2[2], vk2[1], layout.arr_to_g1(d["ic"]), [z[1]], A, B, C)
    pk.free()


def test_qap_matvec_and_prove_from_zkey_wtns_bytes(net):
    """SURVEY f1 + f2: zkey / wtns bytes -> product readers -> GPU qap() -> prove; QAP vectors vs the oracle's
    qap(), proof bytes vs the committed (pairing-verified) golden."""
    import json
    import os
    import sys
    HERE = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, HERE)
    import artefact_writer as aw
    from oracle import bn254 as o, layout
    from distributed_groth16_b200.groth16 import circom, qap as qapmod
    d = np.load(os.path.join(HERE, "golden", "complex_circuit.zkey.pk.npz"))
    exp = json.load(open(os.path.join(HERE, "golden", "complex_circuit_proof.json")))
    n_vars, n_public, m, nc = (int(x) for x in d["dims"])
    zkey_bytes = aw.write_zkey(d)
    z_int = aw.f1_witness(n_vars)
    wtns_bytes = aw.write_wtns(z_int)
    pk, mats, zk = circom.load_zkey(net, zkey_bytes)
    z = circom.load_witness(net, wtns_bytes)
    assert (z.cpu().numpy().view(np.uint64) == layout.fr_to_arr(z_int)).all()
    q = qapmod.qap(mats, z, net)
    vals_a, vals_b = layout.arr_to_fr(d["a_vals"]), layout.arr_to_fr(d["b_vals"])
    ma = [[] for _ in range(nc)]
    mb = [[] for _ in range(nc)]
    for r_, c_, v in zip(d["a_rows"], d["a_cols"], vals_a):
        ma[int(r_)].append((v, int(c_)))
    for r_, c_, v in zip(d["b_rows"], d["b_cols"], vals_b):
        mb[int(r_)].append((v, int(c_)))
    ea, eb, ec = o.qap(ma, mb, n_public + 1, nc, z_int)
    assert q.domain.size() == m == len(ea)
    for got, want in ((q.a, ea), (q.b, eb), (q.c, ec)):
        assert (got.cpu().numpy().view(np.uint64) == layout.fr_to_arr(want)).all()
    assert circom.prove_from_matrices(pk, mats, z).hex() == exp["r0s0"]["proof_hex"]
    proof, public = circom.prove_zkey_wtns(net, zkey_bytes, wtns_bytes)
    assert proof.hex() == exp["r0s0"]["proof_hex"]
    assert int.from_bytes(public[0].tobytes(), "little") == int(exp["public_input"])
    pk.free()


def test_config4_sha256_circuit_shape_qap_h_and_prove(net, cref):
    """BASELINE config 4 (fixture F2): the reference's sha256 circuit (m = 2^15, MSM sizes 29 822 / 29 821 / 32 768) with
    its real witness (29 821 of 29 823 entries are 0 or 1 -- the giant-bucket case).  qap() and h are checked against the
    oracle on the real matrices; the proving key of the reference run is not reproducible here (StdRng([42;32]) setup,
    SURVEY 8c), so the prover runs on a dummy CRS of exactly these sizes and must match the CPU twin byte for byte."""
    import os
    from oracle import bn254 as o, layout
    from distributed_groth16_b200.groth16 import circom, qap as qapmod
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sha256_circuit.npz"))
    n_wires, n_pub, n_cons = (int(x) for x in d["dims"])
    n_inputs = n_pub + 1
    z = net.fr_convert(net.to_device(d["witness"]), to_mont=True)
    mats = qapmod.ConstraintMatrices(net, n_inputs, n_cons, (d["a_rows"], d["a_cols"], d["a_vals"]),
                                     (d["b_rows"], d["b_cols"], d["b_vals"]), values_montgomery_depth=-1)
    q = qapmod.qap(mats, z, net)
    m = q.domain.size()
    assert m == 1 << 15
    w = [int.from_bytes(r.tobytes(), "little") for r in d["witness"]]

    def rows(k):
        vals = [int.from_bytes(r.tobytes(), "little") for r in d[k + "_vals"]]
        out = [[] for _ in range(n_cons)]
        for r_, c_, x in zip(d[k + "_rows"], d[k + "_cols"], vals):
            out[int(r_)].append((x, int(c_)))
        return out

    ea, eb, ec = o.qap(rows("a"), rows("b"), n_inputs, n_cons, w)
    a_h, b_h, c_h = (t.cpu().numpy().view(np.uint64) for t in (q.a, q.b, q.c))
    assert (a_h == layout.fr_to_arr(ea)).all() and (b_h == layout.fr_to_arr(eb)).all() and (c_h == layout.fr_to_arr(ec)).all()
    hh = cref.h_circom(a_h, b_h, c_h)
    assert (net.h_circom(a_h, b_h, c_h) == hh).all()
    aq, b1, lq, hq = cref.g1_generate(41, n_wires), cref.g1_generate(42, n_wires), cref.g1_generate(44, n_wires - n_inputs), cref.g1_generate(45, m)
    b2 = cref.g2_generate(43, n_wires)
    vk1, vk2 = cref.g1_generate(46, 3), cref.g2_generate(47, 2)
    vk = np.concatenate([vk1.reshape(-1), vk2.reshape(-1)])
    zero = np.zeros(4, dtype=np.uint64)
    z_h = z.cpu().numpy().view(np.uint64)
    exp = cref.groth16_prove(aq, b1, b2, lq, hq, vk, n_inputs, z_h, hh, zero, zero, mirror_bg1=True)
    pk = ProvingKey(net, aq, b1, b2, lq, hq, n_inputs, vk1[0], vk1[1], vk1[2], vk2[0], vk2[1])
    got = circom.prove_from_matrices(pk, mats, z, mirror_reference_bg1=True)
    assert got == exp
    pk.free()


def test_f3_gpu_setup_then_prove_sha256_and_verify_with_the_oracle_pairing(net):
    """SURVEY f3 + config 4 closed end to end: circuit-specific setup ON THE GPU for the reference's sha256 circuit (known
    trapdoor, like the reference's fixed-seed setup), GPU prove with the real witness, and the proof VERIFIES under the
    oracle's pairing for the public input asserted in groth16/examples/sha256.rs:231-233 -- and fails for another input."""
    import os
    from oracle import bn254 as o, layout
    from distributed_groth16_b200.groth16 import circom, setup
    d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sha256_circuit.npz"))
    n_wires, n_pub, n_cons = (int(x) for x in d["dims"])
    n_inputs = n_pub + 1
    toxic = (0x1234567890ABCDEF1234567890ABCDEF % o.R, 11111111111111111111, 22222222222222222223, 33333333333333333337,
             44444444444444444447)
    coo = lambda k: (d[k + "_rows"], d[k + "_cols"], d[k + "_vals"])
    pk, vk, mats = setup.circuit_specific_setup(net, n_wires, n_inputs, n_cons, coo("a"), coo("b"), coo("c"), toxic)
    z = net.fr_convert(net.to_device(d["witness"]), to_mont=True)
    for r, s in ((0, 0), (987654321, 123456789)):
        proof = circom.prove_from_matrices(pk, mats, z, layout.fr_to_arr([r])[0], layout.fr_to_arr([s])[0])
        A, B, C = o.proof_decompress(proof)
        g1 = lambda a: layout.arr_to_g1(a)[0]
        g2 = lambda a: layout.arr_to_g2(a)[0]
        ic = layout.arr_to_g1(vk.gamma_abc_g1)
        pub = int.from_bytes(d["witness"][1].tobytes(), "little")
        assert pub == 72587776472194017031617589674261467945970986113287823188107011979
        args = (g1(vk.alpha_g1), g2(vk.beta_g2), g2(vk.gamma_g2), g2(vk.delta_g2), ic)
        assert o.groth16_verify(*args, [pub], A, B, C), (r, s)
        assert not o.groth16_verify(*args, [pub + 1], A, B, C)
        from distributed_groth16_b200.groth16 import verify                      # and by the GPU verifier (csrc/verify.cu)
        assert verify.verify_proof(net, vk, layout.fr_to_arr([pub]), proof)
        assert not verify.verify_proof(net, vk, layout.fr_to_arr([pub + 1]), proof)
    pk.free()
