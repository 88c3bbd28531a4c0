"""The reference's n-party prover flow (`dsha256`, groth16/examples/sha256.rs:26-95, == mpc-api/src/main.rs:688-757),
with all pp.n parties simulated in this process the way its `LocalTestNet` does (compatibility layer).

Every arithmetic step runs on the CUDA library through `Net` (`field_op`, `ntt`, `msm`); what is mirrored here is the
protocol: which vectors are packed-secret-shared, what each party computes on its shares, what the king opens and
re-shares.  The single-box prover (`groth16.prove`, `b200zk_groth16_prove`) computes the same proof without any of it."""
from __future__ import annotations

import numpy as np

from ..context import Net
from ..dist_primitives.dfft import _mont_limbs, d_fft_mpc, d_ifft_mpc
from ..dist_primitives.dmsm import GroupElement, d_msm_mpc


def pack_from_witness(pp, assignment) -> list:
    """`pack_from_witness` (sha256.rs:97-121): consecutive chunks of l scalars (the last one zero-padded), packed; party p
    gets the p-th share of every chunk."""
    a = np.ascontiguousarray(assignment, dtype=np.uint64).reshape(-1, 4)
    chunks = -(-a.shape[0] // pp.l)
    padded = np.zeros((chunks * pp.l, 4), dtype=np.uint64)
    padded[: a.shape[0]] = a
    packed = [pp.pack_from_public(padded[i * pp.l:(i + 1) * pp.l]) for i in range(chunks)]
    return [np.stack([packed[i][p] for i in range(chunks)]) for p in range(pp.n)]


def h_mpc(qap_shares, pp, net: Net) -> list:
    """`ext_wit::h` on packed shares (groth16/src/ext_wit.rs:16-101): three d_ifft (rearrange, pad 2) and three d_fft over
    the doubled domain per party; the king opens p, q, w, keeps the odd-coset evaluations (the `swap(i, i*l + t)` loop),
    forms h = p q - w and deals packed shares of it.  Returns the pp.n share vectors of h (m / l entries each)."""
    m = qap_shares[0].domain.size()
    opened = []
    for name in ("a", "b", "c"):
        shares = [np.ascontiguousarray(_host(getattr(q, name)), dtype=np.uint64).reshape(-1, 4) for q in qap_shares]
        coeff = d_ifft_mpc(shares, True, 2, False, m, pp, net)
        evals = d_fft_mpc(coeff, False, 1, False, 2 * m, pp, net)
        s1 = np.concatenate([pp.unpack(np.stack([e[i] for e in evals])) for i in range(evals[0].shape[0])])
        idx = list(range(s1.shape[0]))
        for i in range(m):                                   # for i in 0..m { s1.swap(i, i * pp.l + pp.t) }
            j = i * pp.l + pp.t
            idx[i], idx[j] = idx[j], idx[i]
        opened.append(s1[np.array(idx[:m])])
    p, q, w = opened
    h = net.field_op(1, 2, net.field_op(1, 0, p, q), w)
    packed = [pp.pack_from_public(h[i * pp.l:(i + 1) * pp.l]) for i in range(m // pp.l)]          # pack_vec
    return [np.stack([packed[i][party] for i in range(m // pp.l)]) for party in range(pp.n)]


def _host(x):
    return x.cpu().numpy().view(np.uint64) if hasattr(x, "cpu") else x


def _sum(net: Net, elems, g2: bool) -> GroupElement:
    pts = [e.limbs for e in elems if not e.infinity]
    if not pts:
        return GroupElement(np.zeros(16 if g2 else 8, dtype=np.uint64), True, g2)
    limbs, inf = net.msm(np.stack(pts), _mont_limbs([1] * len(pts)), g2=g2)
    return GroupElement(limbs, inf, g2)


def _scale(net: Net, e: GroupElement, k_mont) -> GroupElement:
    if e.infinity:
        return e
    limbs, inf = net.msm(e.limbs.reshape(1, -1), np.ascontiguousarray(k_mont, dtype=np.uint64).reshape(1, 4), g2=e.g2)
    return GroupElement(limbs, inf, e.g2)


def prove_mpc(net: Net, pp, crs_shares, qap_shares, a_shares, ax_shares, r=None, s=None):
    """`dsha256` for every party at once: h, then prove::{A, B, C}::compute with L = N = M = Z = K = identity as the
    example passes them (sha256.rs:45-88).  crs_shares: pp.n PackedProvingKeyShare; qap_shares: pp.n PackedQAPShare
    (`qap_pss`); a_shares / ax_shares: `pack_from_witness` of z[1..] and z[num_inputs..].  r, s: Montgomery limbs
    (the reference passes zero).  Returns (A, B, C) GroupElements = result[0] of the simulated round; the client still
    adds a_query[0] + alpha_g1 to A and b_g2_query[0] + beta_g2 to B (sha256.rs:208-212, `client_finish`)."""
    zero = np.zeros(4, dtype=np.uint64)
    r = zero if r is None else np.ascontiguousarray(r, dtype=np.uint64)
    s = zero if s is None else np.ascontiguousarray(s, dtype=np.uint64)
    h_shares = h_mpc(qap_shares, pp, net)
    crs = lambda f: [_host(getattr(c, f)) for c in crs_shares]
    a = d_msm_mpc(crs("s"), a_shares, pp, net)                              # A = L + r N + MSM(S, a)
    b = d_msm_mpc(crs("v"), a_shares, pp, net, g2=True)                     # B = Z + s K + MSM(V, a)
    w = d_msm_mpc(crs("w"), ax_shares, pp, net)
    u = d_msm_mpc(crs("u"), h_shares, pp, net)
    terms = [w, u]
    if s.any():
        terms.append(_scale(net, a, s))                                     # A^s
    if r.any():
        terms.append(_scale(net, d_msm_mpc(crs("h"), a_shares, pp, net), r))   # (MSM(H, a))^r ; M^r with M = identity
    return a, b, _sum(net, terms, False)


def client_finish(net: Net, a: GroupElement, b: GroupElement, a_query0, alpha_g1, b_g2_query0, beta_g2):
    """`a += pk.a_query[0] + vk.alpha_g1; b += pk.b_g2_query[0] + vk.beta_g2` (sha256.rs:208-212)."""
    g = lambda limbs, g2: GroupElement(np.ascontiguousarray(limbs, dtype=np.uint64).reshape(-1), not np.any(limbs), g2)
    return (_sum(net, [a, g(a_query0, False), g(alpha_g1, False)], False),
            _sum(net, [b, g(b_g2_query0, True), g(beta_g2, True)], True))
