"""Device context: the role the reference's `Net: MpcSerNet` handle plays for every `d_*` primitive
(/root/reference/dist-primitives/src/channel/mod.rs:8-56, mpc-net/src/lib.rs:37-156) -- it names
"where the parties are".  Here a party is one B200: `Net` wraps one `b200zk_ctx` (one process per GPU)
and, when torch.distributed is initialised, the NCCL world that replaces the king/client star."""
from __future__ import annotations

import ctypes
import json
from enum import IntEnum

import numpy as np

from . import _native
from ._native import B200zkError, c_vp


class MultiplexedStreamID(IntEnum):
    """mpc-net/src/lib.rs:29-33 -- three logical channels; here three CUDA stream slots."""
    Zero = 0
    One = 1
    Two = 2


class MpcNetError(Exception):
    """mpc-net/src/lib.rs:15-26.  `Generic(String)` is what `?` produces from arkworks' msm
    `Err(min_len)` at dist-primitives/src/dmsm/mod.rs:82."""

    def __init__(self, kind: str, message: str):
        super().__init__("%s(%s)" % (kind, message))
        self.kind = kind
        self.message = message


def _as_u64(a, width: int) -> np.ndarray:
    arr = np.ascontiguousarray(a, dtype=np.uint64)
    if arr.ndim == 1 and width and arr.size % width == 0:
        arr = arr.reshape(-1, width)
    return arr


def _ptr(arr: np.ndarray):
    return ctypes.c_void_p(arr.ctypes.data)


class _StreamOrderedLib:
    """The ctypes library as `Net` and the operator modules call it (`net._lib.b200zk_*`), with one rule enforced in one
    place: every entry point that takes DEVICE pointers (`*_dev`) first makes slot 0 the caller's current torch stream
    (`Net._t0`), so tensors produced by torch ops / NCCL and consumed by library kernels (and vice versa) are ordered on
    one stream.  Host-buffer entry points synchronise internally and pass through untouched."""

    def __init__(self, lib, net):
        object.__setattr__(self, "_lib", lib)
        object.__setattr__(self, "_net", net)
        object.__setattr__(self, "_cache", {})

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        if not name.endswith("_dev"):
            return fn
        w = self._cache.get(name)
        if w is None:
            net = self._net

            def w(*args, _fn=fn):
                net._t0()
                return _fn(*args)
            self._cache[name] = w
        return w


class Net:
    """One GPU party.  `Net()` picks cuda:LOCAL_RANK (or cuda:0)."""

    def __init__(self, device: int | None = None):
        import os
        self._bound0 = None
        self._h = None
        self._lib = _StreamOrderedLib(_native.lib(), self)
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0"))
        h = c_vp()
        rc = self._lib.b200zk_ctx_create(int(device), ctypes.byref(h))
        if rc != 0:
            raise B200zkError(rc, "cannot create a CUDA context on device %d (no CPU fallback exists)" % device)
        self._h = h
        self.device = int(device)
        # Stream contract of every tensor-taking method: slot 0 IS torch's current stream, so torch ops, NCCL collectives
        # and library kernels on the same tensors are ordered without the caller having to remember anything
        # (`_t0()` re-binds when the caller switches streams).  Slots 1 / 2 are independent streams for callers that
        # overlap several calls and synchronise explicitly (`sync(sid)`), like the reference's mux streams.
        self._t0()

    # -- mpc-net::MpcNet surface that still makes sense -------------------------------------
    def party_id(self) -> int:
        try:
            import torch.distributed as dist
            return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        except Exception:
            return 0

    def n_parties(self) -> int:
        try:
            import torch.distributed as dist
            return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        except Exception:
            return 1

    def is_king(self) -> bool:
        return self.party_id() == 0

    # -- plumbing ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.b200zk_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc: int):
        if rc != 0:
            msg = self._lib.b200zk_last_error(self._h).decode()
            if rc == _native.ERR_LENGTH:
                raise MpcNetError("Generic", msg)
            raise B200zkError(rc, msg)

    def use_torch_stream(self, sid: int = 0):
        import torch
        h = torch.cuda.current_stream(self.device).cuda_stream
        self.check(self._lib.b200zk_ctx_set_stream(self._h, int(sid), c_vp(h)))
        if int(sid) == 0:
            self._bound0 = h

    def _t0(self):
        """Slot 0 follows torch's current stream on this device (no-op without torch / CUDA)."""
        try:
            import torch
            if not torch.cuda.is_available():
                return
            h = torch.cuda.current_stream(self.device).cuda_stream
        except Exception:
            return
        if h != self._bound0 and self._h:
            self.check(self._lib.b200zk_ctx_set_stream(self._h, 0, c_vp(h)))
            self._bound0 = h

    def sync(self, sid: int = 0):
        self.check(self._lib.b200zk_ctx_sync(self._h, int(sid)))

    def profile(self, on: bool = True):
        self.check(self._lib.b200zk_profile_enable(self._h, int(on)))

    def profile_reset(self):
        self.check(self._lib.b200zk_profile_reset(self._h))

    def profile_report(self) -> dict:
        buf = ctypes.create_string_buffer(1 << 16)
        self.check(self._lib.b200zk_profile_json(self._h, buf, len(buf)))
        return json.loads(buf.value.decode())

    def launch_count(self) -> int:
        return int(self._lib.b200zk_launch_count(self._h))

    # -- raw calls (host numpy buffers) -----------------------------------------------------------
    def msm(self, bases, scalars, g2: bool = False, sid: int = 0):
        w = 16 if g2 else 8
        b = _as_u64(bases, w)
        s = _as_u64(scalars, 4)
        out = np.zeros(w, dtype=np.uint64)
        inf = ctypes.c_int(0)
        fn = self._lib.b200zk_msm_g2 if g2 else self._lib.b200zk_msm_g1
        self.check(fn(self._h, int(sid), _ptr(b), b.shape[0] if b.size else 0, _ptr(s), s.shape[0] if s.size else 0,
                      _ptr(out), ctypes.byref(inf)))
        return out, bool(inf.value)

    def ntt(self, data, inverse=False, coset=False, bitrev_in=False, bitrev_out=False, pad: int = 1, sid: int = 0):
        x = _as_u64(data, 4)
        n = x.shape[0]
        log_n = n.bit_length() - 1
        if n == 0 or (1 << log_n) != n:
            raise B200zkError(_native.ERR_DOMAIN, "length must be a power of two")
        buf = np.zeros((n * pad, 4), dtype=np.uint64)
        buf[:n] = x
        self.check(self._lib.b200zk_ntt_fr(self._h, int(sid), _ptr(buf), log_n, int(inverse), int(coset), int(bitrev_in),
                                           int(bitrev_out), int(pad)))
        return buf

    def h_circom(self, a, b, c):
        a, b, c = (_as_u64(v, 4) for v in (a, b, c))
        m = a.shape[0]
        log_m = m.bit_length() - 1
        if (1 << log_m) != m or b.shape[0] != m or c.shape[0] != m:
            raise B200zkError(_native.ERR_DOMAIN, "a, b, c must share a power-of-two length")
        out = np.zeros((m, 4), dtype=np.uint64)
        self.check(self._lib.b200zk_h_circom(self._h, _ptr(a), _ptr(b), _ptr(c), log_m, _ptr(out)))
        return out

    def field_op(self, field: int, op: int, a, b):
        a, b = _as_u64(a, 4), _as_u64(b, 4)
        out = np.zeros_like(a)
        if field == 1:        # Fr: the documented share-arithmetic entry point; Fq exists for the self-tests only
            self.check(self._lib.b200zk_fr_op(self._h, op, _ptr(a), _ptr(b), _ptr(out), a.shape[0]))
        else:
            self.check(self._lib.b200zk_test_field_op(self._h, field, op, _ptr(a), _ptr(b), _ptr(out), a.shape[0]))
        return out

    def fr_powers(self, base: int, scale: int, n: int) -> np.ndarray:
        """scale * base^i, i < n, as (n, 4) Montgomery limbs -- computed on the device (b200zk_fr_powers_dev): the twiddle
        columns of the n-party fft1 / fft2 stages (dfft/mod.rs:124-134 walks `factor *= factor_stride` serially)."""
        import torch
        R = 21888242871839275222246405745257275088548364400416034343698204186575808495617
        mont = lambda v: np.array([(((int(v) % R) << 256) % R >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)
        out = torch.empty((n, 4), dtype=torch.int64, device=self._dev())
        b, s = mont(base), mont(scale)
        self.check(self._lib.b200zk_fr_powers_dev(self._h, c_vp(b.ctypes.data), c_vp(s.ctypes.data), n, c_vp(out.data_ptr())))
        return out.cpu().numpy().view(np.uint64)

    # -- device-resident (torch tensors, int64 view of the u64 limbs) -----------------------------
    def _dev(self):
        import torch
        return torch.device("cuda", self.device)

    def generate_g1(self, seed: int, n: int):
        import torch
        t = torch.empty((n, 8), dtype=torch.int64, device=self._dev())
        self.check(self._lib.b200zk_g1_generate_dev(self._h, ctypes.c_uint64(seed), n, c_vp(t.data_ptr())))
        self.sync(0)
        return t

    def generate_g2(self, seed: int, n: int):
        import torch
        t = torch.empty((n, 16), dtype=torch.int64, device=self._dev())
        self.check(self._lib.b200zk_g2_generate_dev(self._h, ctypes.c_uint64(seed), n, c_vp(t.data_ptr())))
        self.sync(0)
        return t

    def generate_fr(self, seed: int, n: int):
        import torch
        t = torch.empty((n, 4), dtype=torch.int64, device=self._dev())
        self.check(self._lib.b200zk_fr_generate_dev(self._h, ctypes.c_uint64(seed), n, c_vp(t.data_ptr())))
        self.sync(0)
        return t

    def msm_dev(self, bases, scalars, out_xyzz=None, g2: bool = False, sid: int = 0):
        """bases/scalars: CUDA int64 tensors; returns a CUDA tensor holding the XYZZ partial."""
        import torch
        n = int(bases.shape[0])
        if int(scalars.shape[0]) != n:
            raise MpcNetError("Generic", str(min(n, int(scalars.shape[0]))))
        if out_xyzz is None:
            out_xyzz = torch.empty(32 if g2 else 16, dtype=torch.int64, device=bases.device)
        fn = self._lib.b200zk_msm_g2_dev if g2 else self._lib.b200zk_msm_g1_dev
        self.check(fn(self._h, int(sid), c_vp(bases.data_ptr()), c_vp(scalars.data_ptr()), n, c_vp(out_xyzz.data_ptr())))
        return out_xyzz

    def msm_staged(self, bases, scalars, out_xyzz=None, g2: bool = False, sid: int = 0):
        """Host numpy buffers in (pinned for full PCIe speed), this GPU's XYZZ partial out as a CUDA tensor: the transfer runs in
        parts behind the bucket kernels (b200zk_msm_staged_dev)."""
        import torch
        w = 16 if g2 else 8
        b, s = _as_u64(bases, w), _as_u64(scalars, 4)
        if out_xyzz is None:
            out_xyzz = torch.empty(32 if g2 else 16, dtype=torch.int64, device=self._dev())
        self.check(self._lib.b200zk_msm_staged_dev(self._h, int(sid), 1 if g2 else 0, _ptr(b), b.shape[0] if b.size else 0, _ptr(s),
                                                   s.shape[0] if s.size else 0, c_vp(out_xyzz.data_ptr())))
        return out_xyzz

    def msm_table_windows(self, c: int) -> int:
        return int(self._lib.b200zk_msm_table_windows(int(c)))

    def msm_table_auto_window(self, n: int) -> int:
        return int(self._lib.b200zk_msm_table_auto_window(int(n)))

    def msm_table_build(self, bases, c: int, g2: bool = False, sid: int = 0):
        """Fixed-base window table of `bases` (CUDA int64, n x 8 / n x 16): (windows * n) points, table[w*n+i] =
        2^{c w} * bases[i].  For bases that stay resident across calls (a proving key's query vectors)."""
        import torch
        n = int(bases.shape[0])
        w = self.msm_table_windows(c)
        table = torch.empty((w * n, 16 if g2 else 8), dtype=torch.int64, device=bases.device)
        self.check(self._lib.b200zk_msm_table_build_dev(self._h, int(sid), 1 if g2 else 0, c_vp(bases.data_ptr()), n, int(c),
                                                        c_vp(table.data_ptr())))
        self.sync(sid)                 # one-off preprocessing: hand back a finished table whatever stream the caller uses
        return table

    def msm_table_dev(self, table, scalars, c: int, out_xyzz=None, g2: bool = False, sid: int = 0):
        """Same sum as msm_dev(bases, scalars) from the table msm_table_build(bases, c) made."""
        import torch
        n = int(scalars.shape[0])
        if int(table.shape[0]) != n * self.msm_table_windows(c):
            raise MpcNetError("Generic", str(min(n, int(table.shape[0]) // max(self.msm_table_windows(c), 1))))
        if out_xyzz is None:
            out_xyzz = torch.empty(32 if g2 else 16, dtype=torch.int64, device=table.device)
        self.check(self._lib.b200zk_msm_table_dev(self._h, int(sid), 1 if g2 else 0, c_vp(table.data_ptr()),
                                                  c_vp(scalars.data_ptr()), n, int(c), c_vp(out_xyzz.data_ptr())))
        return out_xyzz

    # -- ark-serialize Compress::Yes point codec (csrc/codec.cu) -------------------------------------
    def points_compress(self, points, g2: bool = False, sid: int = 0):
        """points: CUDA int64 tensor (n, 8 | 16) or host u64 array -> CUDA uint8 tensor (n, 32 | 64)."""
        import torch
        if not isinstance(points, torch.Tensor):
            points = torch.from_numpy(_as_u64(points, 16 if g2 else 8).view(np.int64)).to(self._dev())
        points = points.contiguous()
        n = int(points.shape[0])
        out = torch.empty((n, 64 if g2 else 32), dtype=torch.uint8, device=points.device)
        self.check(self._lib.b200zk_points_compress_dev(self._h, int(sid), 1 if g2 else 0, c_vp(points.data_ptr()), n,
                                                        c_vp(out.data_ptr())))
        self.sync(sid)
        return out

    def points_decompress(self, data, g2: bool = False, check_subgroup: bool = False, sid: int = 0):
        """data: bytes / uint8 array / CUDA uint8 tensor of n encodings -> CUDA int64 tensor (n, 8 | 16) of affine points.
        Raises B200zkError when an encoding is not a curve point (arkworks: SerializationError::InvalidData)."""
        import torch
        w = 64 if g2 else 32
        if not isinstance(data, torch.Tensor):
            raw = np.frombuffer(bytes(data), dtype=np.uint8) if isinstance(data, (bytes, bytearray, memoryview)) else np.ascontiguousarray(data, dtype=np.uint8)
            data = torch.from_numpy(raw.reshape(-1).copy()).to(self._dev())
        data = data.contiguous()
        if data.numel() % w:
            raise B200zkError(_native.ERR_ARG, "encoding length is not a multiple of %d" % w)
        n = data.numel() // w
        out = torch.empty((n, 16 if g2 else 8), dtype=torch.int64, device=data.device)
        bad = ctypes.c_size_t(0)
        self.check(self._lib.b200zk_points_decompress_dev(self._h, int(sid), 1 if g2 else 0, c_vp(data.data_ptr()), n,
                                                          1 if check_subgroup else 0, c_vp(out.data_ptr()), ctypes.byref(bad)))
        return out

    def sum_points_dev(self, xyzz, count: int, g2: bool = False, sid: int = 0):
        w = 16 if g2 else 8
        out = np.zeros(w, dtype=np.uint64)
        inf = ctypes.c_int(0)
        fn = self._lib.b200zk_g2_sum_dev if g2 else self._lib.b200zk_g1_sum_dev
        self.check(fn(self._h, int(sid), c_vp(xyzz.data_ptr()), int(count), _ptr(out), ctypes.byref(inf)))
        return out, bool(inf.value)

    def ntt_dev(self, x, out=None, inverse=False, coset=False, batch: int = 1, sid: int = 0):
        import torch
        n = int(x.shape[-2]) if x.dim() >= 2 else 0
        log_n = n.bit_length() - 1
        if n == 0 or (1 << log_n) != n:
            raise B200zkError(_native.ERR_DOMAIN, "length must be a power of two")
        if out is None:
            out = torch.empty_like(x)
        self.check(self._lib.b200zk_ntt_fr_dev(self._h, int(sid), c_vp(x.data_ptr()), c_vp(out.data_ptr()), log_n,
                                               int(inverse), int(coset), int(batch)))
        return out

    def fr_convert(self, x, to_mont: bool, times: int = 1, sid: int = 0):
        """Montgomery <-> canonical conversion of a CUDA int64 (n, 4) tensor, on the device."""
        import torch
        out = torch.empty_like(x)
        self.check(self._lib.b200zk_fr_convert_dev(self._h, int(sid), c_vp(x.data_ptr()), c_vp(out.data_ptr()),
                                                   x.numel() // 4, int(to_mont), int(times)))
        return out

    def to_device(self, arr):
        """host (n, w) u64 array -> CUDA int64 tensor on this party's GPU."""
        import torch
        a = np.ascontiguousarray(arr)
        if a.dtype == np.uint64:
            a = a.view(np.int64)
        return torch.from_numpy(a).to(self._dev())

    def h_circom_dev(self, a, b, c, out=None):
        import torch
        m = int(a.shape[0])
        log_m = m.bit_length() - 1
        if out is None:
            out = torch.empty_like(a)
        self.check(self._lib.b200zk_h_circom_dev(self._h, c_vp(a.data_ptr()), c_vp(b.data_ptr()), c_vp(c.data_ptr()),
                                                 log_m, c_vp(out.data_ptr())))
        return out
