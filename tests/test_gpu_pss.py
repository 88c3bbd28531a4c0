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


@pytest.mark.parametrize("g2", [False, True])
def test_batched_packexp_equals_the_chunk_by_chunk_reference_flow(net, cref, g2):
    """b200zk_points_matmul_dev with the sharing's pack / unpack matrices == packexp_from_public / unpackexp applied
    chunk by chunk (the latter already checked above against the reference's own round-trip tests)."""
    from distributed_groth16_b200.dist_primitives.dmsm import packexp_from_public_batch, unpackexp_batch
    l = 2
    pp = PackedSharingParams(l, net)
    k = 11                                                   # odd: the last chunk is padded with the identity
    pts = (cref.g2_generate if g2 else cref.g1_generate)(77, k)
    pts[4] = 0                                               # an identity inside a chunk
    shares = packexp_from_public_batch(net.to_device(pts), pp, net, g2=g2)
    chunks = -(-k // l)
    assert tuple(shares.shape) == (chunks, pp.n, 16 if g2 else 8)
    sh = shares.cpu().numpy().view(np.uint64)
    for i in (0, 2, chunks - 1):
        assert (sh[i] == packexp_from_public(pts[i * l:(i + 1) * l], pp, net, g2=g2)).all(), i
    back = unpackexp_batch(shares, False, pp, net, g2=g2).cpu().numpy().view(np.uint64).reshape(chunks * l, -1)
    assert (back[:k] == pts).all() and not back[k:].any()


def test_qap_pss_and_packed_proving_key_shares(net, cref):
    """QAP::pss (qap.rs:143-187) and PackedProvingKeyShare::pack_from_arkworks_proving_key (proving_key.rs:35-110):
    layout (party p gets share p of every chunk; chunk i of a QAP vector = x_rev[i], x_rev[i + m/l], ...) and content
    (unpacking every party's shares gives the public vectors back)."""
    import torch
    from distributed_groth16_b200.dist_primitives import fft_in_place_rearrange
    from distributed_groth16_b200.dist_primitives.dmsm import unpackexp_batch
    from distributed_groth16_b200.groth16 import PackedProvingKeyShare
    from distributed_groth16_b200.groth16.qap import QAP, Radix2Domain, qap_pss
    l, m = 2, 64
    pp = PackedSharingParams(l, net)
    a, b, c = (cref.fr_generate(s, m) for s in (1, 2, 3))
    q = QAP(3, m - 3, net.to_device(a), net.to_device(b), net.to_device(c), Radix2Domain(m))
    shares = qap_pss(q, pp)
    assert len(shares) == pp.n and all(tuple(s.a.shape) == (m // l, 4) and s.rearranged for s in shares)
    for name, vec in (("a", a), ("b", b), ("c", c)):
        xr = fft_in_place_rearrange(vec)
        for i in (0, 5, m // l - 1):
            chunk_shares = np.stack([getattr(s, name)[i].cpu().numpy().view(np.uint64) for s in shares])
            assert (pp.unpack(chunk_shares) == xr[i::m // l]).all(), (name, i)
            assert (chunk_shares == pp.pack_from_public(xr[i::m // l])).all()
    # proving key
    n_vars, n_inputs = 21, 2
    aq, b1, lq, hq = cref.g1_generate(4, n_vars), cref.g1_generate(5, n_vars), cref.g1_generate(6, n_vars - n_inputs), cref.g1_generate(7, m)
    b2 = cref.g2_generate(8, n_vars)
    parties = PackedProvingKeyShare.pack_from_arkworks_proving_key(net, aq, b1, b2, lq, hq, pp)
    assert len(parties) == pp.n
    for field, src, g2 in (("s", aq[1:], False), ("u", hq, False), ("w", lq, False), ("h", b1[1:], False), ("v", b2[1:], True)):
        stacked = torch.stack([getattr(p, field) for p in parties], dim=1).contiguous()      # (chunks, n, w)
        assert stacked.shape[0] == -(-src.shape[0] // l)
        back = unpackexp_batch(stacked, False, pp, net, g2=g2).cpu().numpy().view(np.uint64).reshape(-1, src.shape[1])
        assert (back[: src.shape[0]] == src).all(), field
