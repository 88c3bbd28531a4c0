"""`PackedSharingParams` -- /root/reference/secret-sharing/src/pss.rs:14-149 (compatibility layer).

n = 4l parties, threshold t = l - 1; three tiny radix-2 domains: `share` (size n), `secret` (size l + t + 1 = 2l, coset
of the generator) and `secret2` (size 4l, coset).  pack = iFFT over `secret` then FFT over `share` (pss.rs:87-93);
unpack = iFFT over `share`, FFT over `secret`, keep the first l (pss.rs:110-128); unpack2 = iFFT over `share`, FFT over
`secret2`, keep every other one of the first 2l (pss.rs:131-149).  All transforms run through `b200zk_ntt_fr`
(arkworks semantics: the vector is zero-padded / truncated to the domain size before each transform).  The single-box
prover does not need secret sharing; this mirror exists so that the GPU kernels can be dropped under the reference's
actual MPC protocol (see dist_primitives/dmsm.py::{packexp_from_public, unpackexp, d_msm_mpc})."""
from __future__ import annotations

import numpy as np

from ..context import Net


def _resize(v: np.ndarray, n: int) -> np.ndarray:
    v = np.ascontiguousarray(v, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros((n, 4), dtype=np.uint64)
    k = min(n, v.shape[0])
    out[:k] = v[:k]
    return out


class PackedSharingParams:
    def __init__(self, l: int, net: Net):
        assert l >= 1 and (l & (l - 1)) == 0, "the reference only instantiates powers of two (l = 2)"
        self.l, self.t, self.n = l, l - 1, 4 * l
        assert self.n == 2 * (self.t + self.l + 1)               # pss.rs:37
        self.share_size, self.secret_size, self.secret2_size = self.n, 2 * l, 4 * l
        self.net = net

    # field-element versions -----------------------------------------------------------------------
    def pack_from_public(self, secrets) -> np.ndarray:
        s = np.ascontiguousarray(secrets, dtype=np.uint64).reshape(-1, 4)
        assert s.shape[0] == self.l, "Secrets length mismatch"
        coeffs = self.net.ntt(_resize(s, self.secret_size), inverse=True, coset=True)       # secret.ifft_in_place
        return self.net.ntt(_resize(coeffs, self.share_size))                               # share.fft_in_place

    def unpack(self, shares) -> np.ndarray:
        c = self.net.ntt(_resize(shares, self.share_size), inverse=True)
        return self.net.ntt(_resize(c, self.secret_size), coset=True)[: self.l]

    def unpack2(self, shares) -> np.ndarray:
        c = self.net.ntt(_resize(shares, self.share_size), inverse=True)
        return self.net.ntt(_resize(c, self.secret2_size), coset=True)[: 2 * self.l: 2]

    # whole vectors at once ------------------------------------------------------------------------
    def pack_from_public_batch(self, secrets):
        """secrets: CUDA int64 tensor (chunks, l, 4) -> shares (chunks, n, 4): pack_from_public of every chunk, as two
        batched device transforms (what QAP::pss does chunk by chunk, groth16/src/qap.rs:151-166)."""
        import torch
        chunks = int(secrets.shape[0])
        assert tuple(secrets.shape[1:]) == (self.l, 4)
        x = torch.zeros((chunks, self.secret_size, 4), dtype=torch.int64, device=secrets.device)
        x[:, : self.l] = secrets
        coeffs = self.net.ntt_dev(x, inverse=True, coset=True, batch=chunks)
        y = torch.zeros((chunks, self.share_size, 4), dtype=torch.int64, device=secrets.device)
        y[:, : self.secret_size] = coeffs
        return self.net.ntt_dev(y, batch=chunks)

    def pack_matrix(self) -> np.ndarray:
        """(n, l) matrix M with shares = M * secrets (pack is linear): column i = pack_from_public(e_i)."""
        if getattr(self, "_pack_m", None) is None:
            from .._constants import FR_ONE_MONT
            cols = []
            for i in range(self.l):
                e = np.zeros((self.l, 4), dtype=np.uint64)
                e[i] = np.array(FR_ONE_MONT, dtype=np.uint64)
                cols.append(self.pack_from_public(e))
            self._pack_m = np.ascontiguousarray(np.stack(cols, axis=1))        # (n, l, 4)
        return self._pack_m

    def unpack_matrix(self, degree2: bool = False) -> np.ndarray:
        """(l, n) matrix U with secrets = U * shares: column j = unpack(e_j) (unpack2 when degree2)."""
        key = "_unpack2_m" if degree2 else "_unpack_m"
        if getattr(self, key, None) is None:
            from .._constants import FR_ONE_MONT
            cols = []
            for j in range(self.n):
                e = np.zeros((self.n, 4), dtype=np.uint64)
                e[j] = np.array(FR_ONE_MONT, dtype=np.uint64)
                cols.append(self.unpack2(e) if degree2 else self.unpack(e))
            setattr(self, key, np.ascontiguousarray(np.stack(cols, axis=1)))     # (l, n, 4)
        return getattr(self, key)
