"""The reference's n-party flows on the GPU `Net`: `d_fft` / `d_ifft` on packed shares
(dist-primitives/src/dfft/mod.rs:285-556 are the reference's own checks), `ext_wit::h` on shares (ext_wit.rs:118-190) and the
whole `dsha256` prover (groth16/examples/sha256.rs:26-95) with all parties simulated in this process -- every field / group
operation runs in libb200zk.so -- plus the parties' messages pushed through the king / client wire codec
(dist_primitives/channel.py: ark-serialize payload + u32-BE frame + ProdNet packet) on the way."""
import numpy as np
import pytest

from distributed_groth16_b200.dist_primitives import channel as ch, d_fft_mpc, d_ifft_mpc, fft_in_place_rearrange, packexp_from_public
from distributed_groth16_b200.groth16 import PackedProvingKeyShare, mpc
from distributed_groth16_b200.groth16.qap import PackedQAPShare, Radix2Domain
from distributed_groth16_b200.secret_sharing import PackedSharingParams
from test_host_dfft_mpc import _open, _share
from test_host_mpc_prover import _pack_points

pytestmark = pytest.mark.gpu


def _over_the_wire(net, shares, prod):
    """every party's share vector as the king would receive it: serialize -> packet -> frame -> stream -> back"""
    out = []
    for s in shares:
        stream = ch.client_message(ch.serialize_fr_vec(net, s), prod=prod)
        payload, rest = ch.read_message(stream, prod=prod)
        assert rest == b"" and payload is not None
        out.append(ch.deserialize_fr_vec(net, payload))
    return out


@pytest.mark.parametrize("l,m", [(2, 64), (4, 256), (2, 1024)])
def test_d_fft_and_d_ifft_protocols_on_the_gpu_net(net, cref, l, m):
    pp = PackedSharingParams(l, net)
    x = cref.fr_generate(500 + m + l, m)
    shares = _share(fft_in_place_rearrange(x), pp)
    wired = _over_the_wire(net, shares, prod=(l == 4))
    assert all((a == b).all() for a, b in zip(wired, shares))                       # the wire format is lossless
    got = _open(d_fft_mpc(wired, False, 1, False, m, pp, net), pp)
    assert (got == cref.ntt(x)).all()                                              # d_fft_works
    got = _open(d_ifft_mpc(wired, False, 1, False, m, pp, net), pp)
    assert (got == cref.ntt(x, inverse=True)).all()                                # d_ifft_works
    # rearrange + pad chaining as ext_wit::h uses it (ext_wit.rs:34-52)
    mid = d_ifft_mpc(shares, True, 2, False, m, pp, net)
    got = _open(d_fft_mpc(mid, False, 1, False, 2 * m, pp, net), pp)
    coeffs = np.concatenate([cref.ntt(x, inverse=True), np.zeros((m, 4), dtype=np.uint64)])
    assert (got == cref.ntt(coeffs)).all()


def test_wire_payloads_match_the_oracle_codec_on_the_gpu(net, cref):
    """channel.py's payloads produced with the GPU conversions / point codec == the oracle's ark-serialize bytes."""
    import struct
    from oracle import bn254 as o, layout
    x = cref.fr_generate(9, 300)
    want = struct.pack("<Q", 300) + b"".join(v.to_bytes(32, "little") for v in layout.arr_to_fr(x))
    got = ch.serialize_fr_vec(net, x)
    assert got == want and (ch.deserialize_fr_vec(net, got) == x).all()
    p1, p2 = cref.g1_generate(3, 4), cref.g2_generate(4, 2)
    for i in range(4):
        enc = ch.serialize_point(net, p1[i])
        assert enc == o.g1_compress(layout.arr_to_g1(p1[i:i + 1])[0]) and (ch.deserialize_point(net, enc) == p1[i]).all()
    for i in range(2):
        enc = ch.serialize_point(net, p2[i], g2=True)
        assert enc == o.g2_compress(layout.arr_to_g2(p2[i:i + 1])[0]) and (ch.deserialize_point(net, enc, g2=True) == p2[i]).all()
    from distributed_groth16_b200 import MpcNetError
    with pytest.raises(MpcNetError):
        ch.deserialize_point(net, (4).to_bytes(32, "little"))                      # x^3 + 3 is a non-residue: InvalidData


def test_mpc_prover_flow_on_the_gpu_net_gives_the_single_node_proof(net, cref):
    """`dsha256` party by party on the GPU: h on shares, prove::{A,B,C} over packed CRS / witness shares, client finish ==
    the single-node proof (with r = s = 0 the same group elements, sha256.rs:240-254)."""
    from oracle import bn254 as o, layout
    l, m, n_vars, n_inputs = 2, 64, 50, 2
    pp = PackedSharingParams(l, net)
    aq, b1, lq, hq = (cref.g1_generate(s, k) for s, k in ((61, n_vars), (62, n_vars), (64, n_vars - n_inputs), (65, m)))
    b2, vk1, vk2 = cref.g2_generate(63, n_vars), cref.g1_generate(66, 3), cref.g2_generate(67, 2)
    hq[m - 1] = 0                                            # arkworks' h_query has m - 1 entries
    z = cref.fr_generate(68, n_vars)
    z[0] = layout.fr_to_arr([1])[0]
    a, b, c = (cref.fr_generate(s, m) for s in (69, 70, 71))
    zero = np.zeros(4, dtype=np.uint64)
    vk = np.concatenate([vk1.reshape(-1), vk2.reshape(-1)])
    want = cref.groth16_prove(aq, b1, b2, lq, hq, vk, n_inputs, z, cref.h_circom(a, b, c), zero, zero, mirror_bg1=False)
    wa, wb, wc = o.proof_decompress(want)
    dom = Radix2Domain(m)
    sa, sb, sc = (_share(fft_in_place_rearrange(v), pp) for v in (a, b, c))                      # QAP::pss
    qap_shares = [PackedQAPShare(n_inputs, m - n_inputs, sa[p], sb[p], sc[p], dom, rearranged=True) for p in range(pp.n)]
    s_sh, u_sh, w_sh, h_sh = (_pack_points(v, pp, net) for v in (aq[1:], hq, lq, b1[1:]))
    v_sh = _pack_points(b2[1:], pp, net, g2=True)
    crs_shares = [PackedProvingKeyShare(s_sh[p], u_sh[p], v_sh[p], w_sh[p], h_sh[p]) for p in range(pp.n)]
    a_shares = mpc.pack_from_witness(pp, z[1:])
    ax_shares = mpc.pack_from_witness(pp, z[n_inputs:])
    h_shares = mpc.h_mpc(qap_shares, pp, net)
    opened = np.concatenate([pp.unpack(np.stack([hs[i] for hs in h_shares])) for i in range(m // l)])
    assert (opened == cref.h_circom(a, b, c)).all()                               # ext_wit.rs:137-187's assertion
    ga, gb, gc = mpc.prove_mpc(net, pp, crs_shares, qap_shares, a_shares, ax_shares)
    # what each d_msm puts on the wire: one compressed group element per party (dmsm/mod.rs:84)
    assert (ch.deserialize_point(net, ch.serialize_point(net, gc.limbs)) == gc.limbs).all()
    ga, gb = mpc.client_finish(net, ga, gb, aq[0], vk1[0], b2[0], vk2[0])
    assert layout.arr_to_g1(ga.limbs.reshape(1, -1))[0] == wa
    assert layout.arr_to_g2(gb.limbs.reshape(1, -1))[0] == wb
    assert layout.arr_to_g1(gc.limbs.reshape(1, -1))[0] == wc
    assert o.proof_compress(wa, wb, wc) == want
