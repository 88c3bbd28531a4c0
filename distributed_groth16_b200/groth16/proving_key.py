"""Device-resident proving key -- the vectors `PackedProvingKeyShare{s,u,v,w,h}` carries
(/root/reference/groth16/src/proving_key.rs:19-25; mapping :48-65: s = a_query[1..], u = h_query,
w = l_query, h = b_g1_query[1..], v = b_g2_query[1..]) plus the vk points the assembly needs."""
from __future__ import annotations

import ctypes

import numpy as np

from .. import _native
from ..context import Net, _as_u64, _ptr


class ProvingKey:
    def __init__(self, net: Net, a_query, b_g1_query, b_g2_query, l_query, h_query, n_inputs: int, alpha_g1, beta_g1,
                 delta_g1, beta_g2, delta_g2):
        self.net = net
        aq, b1, lq, hq = (_as_u64(v, 8) for v in (a_query, b_g1_query, l_query, h_query))
        b2 = _as_u64(b_g2_query, 16)
        self.n_vars = aq.shape[0]
        self.n_inputs = int(n_inputs)
        self.m = hq.shape[0]
        assert b1.shape[0] == self.n_vars and b2.shape[0] == self.n_vars
        assert lq.size == 0 or lq.shape[0] == self.n_vars - self.n_inputs
        vk = np.concatenate([np.asarray(v, dtype=np.uint64).reshape(-1) for v in
                             (alpha_g1, beta_g1, delta_g1, beta_g2, delta_g2)])
        assert vk.size == 56
        self.host = dict(a_query=aq, b_g1_query=b1, b_g2_query=b2, l_query=lq, h_query=hq, vk=vk)
        h = _native.c_vp()
        lq_ptr = _ptr(lq) if lq.size else None
        net.check(net._lib.b200zk_pk_upload(net._h, _ptr(aq), _ptr(b1), _ptr(b2), lq_ptr, _ptr(hq), self.n_vars,
                                            self.n_inputs, self.m, _ptr(vk), ctypes.byref(h)))
        self._h = h

    @classmethod
    def from_device(cls, net: Net, a_query, b_g1_query, b_g2_query, l_query, h_query, n_inputs: int, vk_points):
        """Device-resident CUDA int64 tensors (e.g. a dummy CRS made by net.generate_g1/g2); vk_points: 56 host limbs."""
        self = cls.__new__(cls)
        self.net = net
        self.n_vars = int(a_query.shape[0])
        self.n_inputs = int(n_inputs)
        self.m = int(h_query.shape[0])
        self.host = None
        vk = np.ascontiguousarray(vk_points, dtype=np.uint64).reshape(-1)
        assert vk.size == 56
        h = _native.c_vp()
        lq_ptr = _native.c_vp(l_query.data_ptr()) if l_query.numel() else None
        net.check(net._lib.b200zk_pk_upload_dev(net._h, _native.c_vp(a_query.data_ptr()), _native.c_vp(b_g1_query.data_ptr()),
                                                _native.c_vp(b_g2_query.data_ptr()), lq_ptr,
                                                _native.c_vp(h_query.data_ptr()), self.n_vars, self.n_inputs, self.m,
                                                _ptr(vk), ctypes.byref(h)))
        self._h = h
        return self

    def precompute(self, c: int = 0):
        """(Re)build the fixed-base window tables of the five query vectors (b200zk_pk_precompute): c = 0 -> automatic
        window per query; c = None -> drop the tables, proving then runs the generic MSM on the queries.  The upload
        already did precompute(0) unless B200ZK_PK_TABLES=0.  Proof bytes are identical either way."""
        self.net.check(self.net._lib.b200zk_pk_precompute(self.net._h, self._h, 0xFFFFFFFF if c is None else int(c)))
        return self.table_bytes

    @property
    def table_bytes(self) -> int:
        return int(self.net._lib.b200zk_pk_table_bytes(self._h))

    def free(self):
        if getattr(self, "_h", None) and self.net._h:
            self.net._lib.b200zk_pk_free(self.net._h, self._h)
        self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class PackedProvingKeyShare:
    """`PackedProvingKeyShare` (groth16/src/proving_key.rs:16-27): one party's packed shares of the five query vectors,
    s = a_query[1..], u = h_query, w = l_query, h = b_g1_query[1..] (G1) and v = b_g2_query[1..] (G2); one share per
    l-point chunk.  Arrays are CUDA int64 tensors (chunks, 8 | 16) of affine Montgomery limbs."""

    def __init__(self, s, u, v, w, h):
        self.s, self.u, self.v, self.w, self.h = s, u, v, w, h

    @staticmethod
    def pack_from_arkworks_proving_key(net: Net, a_query, b_g1_query, b_g2_query, l_query, h_query, pp) -> list:
        """proving_key.rs:35-110: every chunk of every query goes through `packexp_from_public`; here one
        `b200zk_points_matmul_dev` launch per query.  Returns pp.n PackedProvingKeyShare objects."""
        from ..dist_primitives.dmsm import packexp_from_public_batch
        dev = lambda a, w: a if hasattr(a, "data_ptr") else net.to_device(_as_u64(a, w))
        packed = {
            "s": packexp_from_public_batch(dev(a_query, 8)[1:].contiguous(), pp, net),
            "u": packexp_from_public_batch(dev(h_query, 8), pp, net),
            "w": packexp_from_public_batch(dev(l_query, 8), pp, net),
            "h": packexp_from_public_batch(dev(b_g1_query, 8)[1:].contiguous(), pp, net),
            "v": packexp_from_public_batch(dev(b_g2_query, 16)[1:].contiguous(), pp, net, g2=True),
        }
        return [PackedProvingKeyShare(**{k: t[:, p].contiguous() for k, t in packed.items()}) for p in range(pp.n)]
