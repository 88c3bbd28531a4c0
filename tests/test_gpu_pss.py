"""The reference's own secret-sharing / dmsm unit tests (secret-sharing/src/pss.rs:164-241,
dist-primitives/src/dmsm/mod.rs:127-193, examples/dmsm_test.rs) re-stated on the GPU kernels, BN254 instead of BLS12-377."""
import numpy as np
import pytest

from distributed_groth16_b200.dist_primitives import d_msm, d_msm_mpc, packexp_from_public, unpackexp
from distributed_groth16_b200.secret_sharing import PackedSharingParams

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("l", [2, 4])
def test_pack_unpack_and_degree2_unpack(net, cref, l):
    """pss.rs test_initialize / test_pack_from_public / test_multiplication."""
    pp = PackedSharingParams(l, net)
    assert (pp.n, pp.t) == (4 * l, l - 1)
    secrets = cref.fr_generate(17 + l, l)
    shares = pp.pack_from_public(secrets)
    assert shares.shape[0] == pp.n
    assert (pp.unpack(shares) == secrets).all()
    sq = net.field_op(1, 0, shares, shares)                     # share-wise product has degree 2(t + l)
    exp = net.field_op(1, 0, secrets, secrets)
    assert (pp.unpack2(sq) == exp).all()


def test_packexp_unpackexp_roundtrip(net, cref):
    """dmsm/mod.rs pack_unpack_test (l = 2)."""
    pp = PackedSharingParams(2, net)
    secrets = cref.g1_generate(5, pp.l)
    shares = packexp_from_public(secrets, pp, net)
    assert (unpackexp(shares, False, pp, net) == secrets).all()


def test_d_msm_protocol_equals_plain_msm(net, cref):
    """dmsm/mod.rs pack_unpack2_test + examples/dmsm_test.rs:49-64: the n-party protocol on packed shares of bases and
    scalars, degree-2 unpacking in the exponent, equals G::msm on the public vectors."""
    l, M = 2, 64
    pp = PackedSharingParams(l, net)
    bases = cref.g1_generate(31, M)
    scalars = cref.fr_generate(32, M)
    b_shares = np.zeros((pp.n, M // l, 8), dtype=np.uint64)
    s_shares = np.zeros((pp.n, M // l, 4), dtype=np.uint64)
    for i in range(M // l):                                     # chunk i holds the l secrets i*l .. i*l + l - 1
        b_shares[:, i] = packexp_from_public(bases[i * l:(i + 1) * l], pp, net)
        s_shares[:, i] = pp.pack_from_public(scalars[i * l:(i + 1) * l])
    got = d_msm_mpc(list(b_shares), list(s_shares), pp, net)
    exp = d_msm(bases, scalars, None, net)
    ref, inf = cref.msm_g1(bases, scalars)
    assert not inf and (exp.limbs == ref).all()
    assert got == exp
