"""Several GPUs of one box behind single calls: the Python face of `b200zk_group_*` (csrc/group.cu).

The reference reaches its parties through `Net: MpcSerNet` (dist-primitives/src/channel/mod.rs:8-56); `Group` is that handle
for "all the GPUs of this box, one host process".  The sharded MSM / four-step NTT / h pipeline / Groth16 prover run inside
the library (peer stores over NVLink instead of the king/client star), so the Rust prover flow of
groth16/src/prove.rs:106-136 can reach BASELINE config 5 through FFI without torch.distributed.  `parallel.py` keeps the
one-process-per-GPU (torchrun + NCCL) orchestration that bench.py's scaling runs use."""
from __future__ import annotations

import ctypes

import numpy as np

from . import _native
from ._native import B200zkError, c_vp
from .context import MpcNetError, _as_u64, _ptr
from .dist_primitives.dmsm import GroupElement

ZERO_FR = np.zeros(4, dtype=np.uint64)


class Group:
    def __init__(self, devices):
        self._lib = _native.lib()
        devs = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        h = c_vp()
        rc = self._lib.b200zk_group_create(devs, len(devices), ctypes.byref(h))
        if rc != 0:
            raise B200zkError(rc, "cannot create a GPU group on devices %r (1, 2, 4 or 8 peer-connected GPUs; no CPU fallback)"
                              % (list(devices),))
        self._h = h
        self.devices = [int(d) for d in devices]

    def size(self) -> int:
        return int(self._lib.b200zk_group_size(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200zk_group_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != 0:
            msg = self._lib.b200zk_group_last_error(self._h).decode()
            if rc == _native.ERR_LENGTH:
                raise MpcNetError("Generic", msg)
            raise B200zkError(rc, msg)

    # -- d_msm -----------------------------------------------------------------------------------------------------------------
    def d_msm(self, bases, scalars, g2: bool | None = None) -> GroupElement:
        width = int(np.asarray(bases).shape[-1]) if np.asarray(bases).ndim == 2 else None
        if g2 is None:
            g2 = width == 16
        w = 16 if g2 else 8
        b, s = _as_u64(bases, w), _as_u64(scalars, 4)
        out = np.zeros(w, dtype=np.uint64)
        inf = ctypes.c_int(0)
        fn = self._lib.b200zk_group_msm_g2 if g2 else self._lib.b200zk_group_msm_g1
        self.check(fn(self._h, _ptr(b), b.shape[0] if b.size else 0, _ptr(s), s.shape[0] if s.size else 0, _ptr(out), ctypes.byref(inf)))
        return GroupElement(out, bool(inf.value), g2)

    # -- d_fft / d_ifft ----------------------------------------------------------------------------------------------------------
    def d_fft(self, x, inverse: bool = False) -> np.ndarray:
        buf = np.array(_as_u64(x, 4), dtype=np.uint64, copy=True)
        n = buf.shape[0]
        log_n = n.bit_length() - 1
        if n == 0 or (1 << log_n) != n:
            raise B200zkError(_native.ERR_DOMAIN, "length must be a power of two")
        self.check(self._lib.b200zk_group_ntt_fr(self._h, _ptr(buf), log_n, int(inverse)))
        return buf

    def d_ifft(self, x) -> np.ndarray:
        return self.d_fft(x, inverse=True)

    # -- ext_wit::h --------------------------------------------------------------------------------------------------------------
    def h(self, a, b, c) -> np.ndarray:
        a, b, c = (_as_u64(v, 4) for v in (a, b, c))
        m = a.shape[0]
        log_m = m.bit_length() - 1
        if m == 0 or (1 << log_m) != m or b.shape[0] != m or c.shape[0] != m:
            raise B200zkError(_native.ERR_DOMAIN, "a, b, c must share one power-of-two length")
        out = np.zeros((m, 4), dtype=np.uint64)
        self.check(self._lib.b200zk_group_h_circom(self._h, _ptr(a), _ptr(b), _ptr(c), log_m, _ptr(out)))
        return out


class GroupProvingKey:
    """A proving key sharded over the GPUs of a Group (arguments as groth16.ProvingKey)."""

    def __init__(self, group: Group, a_query, b_g1_query, b_g2_query, l_query, h_query, n_inputs: int, alpha_g1, beta_g1, delta_g1,
                 beta_g2, delta_g2):
        aq, b1, b2, hq = _as_u64(a_query, 8), _as_u64(b_g1_query, 8), _as_u64(b_g2_query, 16), _as_u64(h_query, 8)
        lq = _as_u64(l_query, 8)
        self.group = group
        self.n_vars, self.m, self.n_inputs = aq.shape[0], hq.shape[0], int(n_inputs)
        vk = np.concatenate([np.asarray(p, dtype=np.uint64).reshape(-1) for p in (alpha_g1, beta_g1, delta_g1, beta_g2, delta_g2)])
        assert vk.size == 56
        h = c_vp()
        group.check(group._lib.b200zk_group_pk_upload(group._h, _ptr(aq), _ptr(b1), _ptr(b2), _ptr(lq) if lq.size else None, _ptr(hq),
                                                      self.n_vars, self.n_inputs, self.m, _ptr(vk), ctypes.byref(h)))
        self._h = h

    @property
    def table_bytes(self) -> int:
        return int(self.group._lib.b200zk_group_pk_table_bytes(self._h))

    def free(self):
        if getattr(self, "_h", None):
            self.group._lib.b200zk_group_pk_free(self.group._h, self._h)
            self._h = None

    def create_proof(self, z, a, b, c, r=None, s=None) -> bytes:
        """128 compressed proof bytes, identical to groth16.prove.create_proof on one GPU."""
        z, a, b, c = (_as_u64(v, 4) for v in (z, a, b, c))
        assert z.shape[0] == self.n_vars and a.shape[0] == self.m and b.shape[0] == self.m and c.shape[0] == self.m
        r = ZERO_FR if r is None else np.ascontiguousarray(r, dtype=np.uint64)
        s = ZERO_FR if s is None else np.ascontiguousarray(s, dtype=np.uint64)
        out = (ctypes.c_uint8 * 128)()
        self.group.check(self.group._lib.b200zk_group_groth16_prove(self.group._h, self._h, _ptr(z), _ptr(a), _ptr(b), _ptr(c), _ptr(r),
                                                                    _ptr(s), out))
        return bytes(out)
