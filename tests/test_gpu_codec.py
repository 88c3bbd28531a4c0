"""ark-serialize Compress::Yes point codec on the GPU (csrc/codec.cu) vs the oracle's big-int rules, pinned on the
reference's own golden proof (zk-cli/test-circuits/sha256/proof.bin, coordinates printed in zk-cli/README.md:82)."""
import os

import numpy as np
import pytest

from distributed_groth16_b200 import B200zkError, ark_serialize as ark

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("g2", [False, True])
def test_compress_matches_oracle_and_round_trips(net, cref, g2):
    from oracle import bn254 as o, layout
    n = 300
    pts = (cref.g2_generate if g2 else cref.g1_generate)(0xC0DEC, n)
    pts[5] = 0                                             # infinity
    to_pts = layout.arr_to_g2 if g2 else layout.arr_to_g1
    from_pts = layout.g2_to_arr if g2 else layout.g1_to_arr
    neg = (o.G2 if g2 else o.G1).neg(to_pts(pts[6:7])[0])
    pts[7] = from_pts([neg])[0]                            # both signs of one x
    enc = net.points_compress(pts, g2=g2).cpu().numpy()
    comp = o.g2_compress if g2 else o.g1_compress
    exp = b"".join(comp(p) for p in to_pts(pts))
    assert enc.tobytes() == exp
    assert enc[6].tobytes()[:-1] == enc[7].tobytes()[:-1] and (enc[6][-1] ^ enc[7][-1]) == 0x80
    dec = net.points_decompress(enc.tobytes(), g2=g2, check_subgroup=True).cpu().numpy().view(np.uint64)
    assert (dec == pts).all()


def test_reference_golden_proof_decodes_to_the_readme_coordinates(net):
    from oracle import bn254 as o, layout
    buf = open(os.path.join(GOLD, "sha256_proof.bin"), "rb").read()
    a, b, c = ark.deserialize_proof(net, buf)
    ea, eb, ec = o.proof_decompress(buf)
    assert (a == layout.g1_to_arr([ea])[0]).all()
    assert (b == layout.g2_to_arr([eb])[0]).all()
    assert (c == layout.g1_to_arr([ec])[0]).all()
    # zk-cli/README.md:82 prints A.x of this proof in decimal
    assert ea[0] == int.from_bytes(bytes(buf[:31]) + bytes([buf[31] & 0x3F]), "little")
    assert ark.serialize_proof(net, a, b, c) == buf


def test_invalid_encodings_are_rejected(net):
    from oracle import bn254 as o
    x = 1
    while pow((x ** 3 + 3) % o.P, (o.P - 1) // 2, o.P) == 1:   # smallest x with x^3 + 3 a non-residue
        x += 1
    good = o.g1_compress(o.G1_GEN if hasattr(o, "G1_GEN") else (1, 2))
    bad_curve = x.to_bytes(32, "little")
    bad_range = (o.P + 1).to_bytes(32, "little")              # x >= p (fits below the flag bits: p < 2^254)
    bad_inf = bytes(31) + bytes([0xC0])                       # infinity + sign flag
    for blob, count in ((bad_curve, 1), (bad_range, 1), (bad_inf, 1), (good + bad_curve + bad_range, 2)):
        with pytest.raises(B200zkError) as ei:
            net.points_decompress(blob)
        assert ("%d of" % count) in str(ei.value)
    assert net.points_decompress(good).shape[0] == 1
    with pytest.raises(B200zkError):
        net.points_decompress(good[:-1])


def test_g2_subgroup_check(net):
    """A point of the twist outside the order-r subgroup (the twist's cofactor is ~2^254) passes the curve equation and
    fails Validate::Yes."""
    from oracle import bn254 as o
    x0 = 1
    while True:
        x = (x0, 1)
        y2 = o.fq2_add(o.fq2_mul(o.fq2_sqr(x), x), o.B_G2)
        y = o.fq2_sqrt(y2)
        if y is not None and o.G2.from_jac(o.G2.jac_mul(o.G2.to_jac((x, y)), o.R)) is not None:   # (Curve.mul reduces k mod r)
            break
        x0 += 1
    blob = o.g2_compress((x, y))
    got = net.points_decompress(blob, g2=True, check_subgroup=False).cpu().numpy().view(np.uint64)
    from oracle import layout
    assert (got == layout.g2_to_arr([o.g2_decompress(blob)])).all()
    with pytest.raises(B200zkError):
        net.points_decompress(blob, g2=True, check_subgroup=True)


def test_proving_key_round_trip_and_layout(net, cref):
    n_vars, n_inputs, m = 700, 3, 1024
    g1, g2 = cref.g1_generate, cref.g2_generate
    vk = ark.ArkVerifyingKey(g1(1, 1)[0], g2(2, 1)[0], g2(3, 1)[0], g2(4, 1)[0], g1(5, n_inputs))
    pk = ark.ArkProvingKey(vk, g1(6, 1)[0], g1(7, 1)[0], g1(8, n_vars), g1(9, n_vars), g2(10, n_vars), g1(11, m - 1),
                           g1(12, n_vars - n_inputs))
    pk.a_query[3] = 0
    pk.b_g2_query[0] = 0
    buf = ark.serialize_proving_key(net, pk)
    vk_len = 32 + 3 * 64 + 8 + 32 * n_inputs
    assert len(buf) == vk_len + 64 + 8 * 5 + 32 * (2 * n_vars + (m - 1) + (n_vars - n_inputs)) + 64 * n_vars
    assert int.from_bytes(buf[32 + 192:32 + 200], "little") == n_inputs          # Vec length prefix of gamma_abc_g1
    assert int.from_bytes(buf[vk_len + 64:vk_len + 72], "little") == n_vars      # a_query
    back = ark.deserialize_proving_key(net, buf, check_subgroup=True)
    for name in ("beta_g1", "delta_g1", "a_query", "b_g1_query", "b_g2_query", "h_query", "l_query"):
        assert (getattr(back, name) == getattr(pk, name)).all(), name
    for name in ("alpha_g1", "beta_g2", "gamma_g2", "delta_g2", "gamma_abc_g1"):
        assert (getattr(back.vk, name) == getattr(vk, name)).all(), name
    assert ark.serialize_verifying_key(net, back.vk) == buf[:vk_len]
    assert (ark.deserialize_verifying_key(net, buf[:vk_len]).gamma_abc_g1 == vk.gamma_abc_g1).all()
    with pytest.raises(ValueError):
        ark.deserialize_proving_key(net, buf[:-1])


def test_decompress_2_20_points_round_trip(net):
    """Key-sized batch: 2^20 G1 points through compress -> decompress on the device."""
    n = 1 << 20
    pts = net.generate_g1(0xC0DE, n)
    enc = net.points_compress(pts)
    dec = net.points_decompress(enc)
    assert bool((dec == pts).all())
