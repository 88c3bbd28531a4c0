"""The reference's n-party prover flow (`dsha256`: ext_wit::h on shares, prove::{A,B,C} over packed CRS / witness shares,
client finish -- groth16/examples/sha256.rs:26-95,170-212) mirrored in groth16/mpc.py, run on a CPU stand-in for `Net`
(oracle arithmetic) and compared with the single-node prover: the reference's own acceptance criterion is that the MPC
proof is a valid proof of the same statement; with r = s = 0 it is the *same* proof, which is what is asserted here."""
import numpy as np

from distributed_groth16_b200.dist_primitives import fft_in_place_rearrange, packexp_from_public
from distributed_groth16_b200.groth16 import PackedProvingKeyShare, mpc
from distributed_groth16_b200.groth16.qap import PackedQAPShare, Radix2Domain
from distributed_groth16_b200.secret_sharing import PackedSharingParams
from test_host_dfft_mpc import OracleNet, _share


class OracleMsmNet(OracleNet):
    def msm(self, bases, scalars, g2=False, sid=0):
        from oracle import bn254 as o, layout
        w = 16 if g2 else 8
        pts = (layout.arr_to_g2 if g2 else layout.arr_to_g1)(np.asarray(bases, dtype=np.uint64).reshape(-1, w))
        sc = layout.arr_to_fr(np.asarray(scalars, dtype=np.uint64).reshape(-1, 4))
        assert len(pts) == len(sc)
        res = (o.G2 if g2 else o.G1).msm_naive(pts, sc)
        arr = (layout.g2_to_arr if g2 else layout.g1_to_arr)([res])[0]
        return arr, res is None


def _pack_points(points, pp, net, g2=False):
    """proving_key.rs:66-80 on the host path: chunk, pad with the identity, packexp; party p gets share p of every chunk."""
    w = points.shape[1]
    chunks = -(-points.shape[0] // pp.l)
    padded = np.zeros((chunks * pp.l, w), dtype=np.uint64)
    padded[: points.shape[0]] = points
    packed = [packexp_from_public(padded[i * pp.l:(i + 1) * pp.l], pp, net, g2=g2) for i in range(chunks)]
    return [np.stack([packed[i][p] for i in range(chunks)]) for p in range(pp.n)]


def test_mpc_prover_flow_gives_the_single_node_proof(cref):
    from oracle import bn254 as o, layout
    net = OracleMsmNet()
    l, m, n_vars, n_inputs = 2, 8, 7, 2
    pp = PackedSharingParams(l, net)
    aq, b1, lq, hq = (cref.g1_generate(s, k) for s, k in ((1, n_vars), (2, n_vars), (4, n_vars - n_inputs), (5, m)))
    b2, vk1, vk2 = cref.g2_generate(3, n_vars), cref.g1_generate(6, 3), cref.g2_generate(7, 2)
    hq[m - 1] = 0                                            # arkworks' h_query has m - 1 entries
    z = cref.fr_generate(8, n_vars)
    z[0] = layout.fr_to_arr([1])[0]
    a, b, c = (cref.fr_generate(s, m) for s in (9, 10, 11))
    zero = np.zeros(4, dtype=np.uint64)
    vk = np.concatenate([vk1.reshape(-1), vk2.reshape(-1)])
    want = cref.groth16_prove(aq, b1, b2, lq, hq, vk, n_inputs, z, cref.h_circom(a, b, c), zero, zero, mirror_bg1=False)
    wa, wb, wc = o.proof_decompress(want)

    # what the reference's driver prepares (sha256.rs:170-189)
    dom = Radix2Domain(m)
    sa, sb, sc = (_share(fft_in_place_rearrange(v), pp) for v in (a, b, c))                      # QAP::pss
    qap_shares = [PackedQAPShare(n_inputs, m - n_inputs, sa[p], sb[p], sc[p], dom, rearranged=True) for p in range(pp.n)]
    s_sh, u_sh, w_sh, h_sh = (_pack_points(v, pp, net) for v in (aq[1:], hq, lq, b1[1:]))
    v_sh = _pack_points(b2[1:], pp, net, g2=True)
    crs_shares = [PackedProvingKeyShare(s_sh[p], u_sh[p], v_sh[p], w_sh[p], h_sh[p]) for p in range(pp.n)]
    a_shares = mpc.pack_from_witness(pp, z[1:])
    ax_shares = mpc.pack_from_witness(pp, z[n_inputs:])

    # h alone: the opened shares are the CircomReduction h of the single-node prover
    h_shares = mpc.h_mpc(qap_shares, pp, net)
    opened = np.concatenate([pp.unpack(np.stack([hs[i] for hs in h_shares])) for i in range(m // l)])
    assert (opened == cref.h_circom(a, b, c)).all()

    ga, gb, gc = mpc.prove_mpc(net, pp, crs_shares, qap_shares, a_shares, ax_shares)
    ga, gb = mpc.client_finish(net, ga, gb, aq[0], vk1[0], b2[0], vk2[0])
    assert layout.arr_to_g1(ga.limbs.reshape(1, -1))[0] == wa
    assert layout.arr_to_g2(gb.limbs.reshape(1, -1))[0] == wb
    assert layout.arr_to_g1(gc.limbs.reshape(1, -1))[0] == wc
