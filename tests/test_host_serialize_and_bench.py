"""CPU-only coverage of host logic that normally sits on top of GPU calls:
* `ark_serialize` container layout (field order, Vec length prefixes, truncation / trailing bytes) with the point codec
  replaced by the oracle's big-int codec (a stand-in `Net`; the GPU codec itself is covered by tests/test_gpu_codec.py);
* `bench.py --impl reference`: exactly one JSON line on stdout, with the keys the bench contract names."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class _Bytes:
    def __init__(self, arr):
        self._a = arr

    def cpu(self):
        return self

    def numpy(self):
        return self._a


class OracleCodecNet:
    """points_compress / points_decompress with the signatures `ark_serialize` uses, computed by oracle/bn254.py."""

    def points_compress(self, points, g2=False, sid=0):
        from oracle import bn254 as o, layout
        pts = (layout.arr_to_g2 if g2 else layout.arr_to_g1)(np.ascontiguousarray(points, dtype=np.uint64).reshape(-1, 16 if g2 else 8))
        blob = b"".join((o.g2_compress if g2 else o.g1_compress)(p) for p in pts)
        return _Bytes(np.frombuffer(blob, dtype=np.uint8).reshape(len(pts), -1))

    def points_decompress(self, data, g2=False, check_subgroup=False, sid=0):
        from oracle import bn254 as o, layout
        w = 64 if g2 else 32
        data = bytes(data)
        pts = [(o.g2_decompress if g2 else o.g1_decompress)(data[i:i + w]) for i in range(0, len(data), w)]
        return _Bytes((layout.g2_to_arr if g2 else layout.g1_to_arr)(pts).view(np.int64))


def test_ark_serialize_container_layout_with_the_oracle_codec(cref):
    from distributed_groth16_b200 import ark_serialize as ark
    from oracle import bn254 as o, layout
    net = OracleCodecNet()
    n_vars, n_inputs, m = 9, 2, 8
    g1, g2 = cref.g1_generate, cref.g2_generate
    vk = ark.ArkVerifyingKey(g1(1, 1)[0], g2(2, 1)[0], g2(3, 1)[0], g2(4, 1)[0], g1(5, n_inputs + 1))
    pk = ark.ArkProvingKey(vk, g1(6, 1)[0], g1(7, 1)[0], g1(8, n_vars), g1(9, n_vars), g2(10, n_vars), g1(11, m - 1),
                           g1(12, n_vars - n_inputs))
    pk.l_query[2] = 0
    buf = ark.serialize_proving_key(net, pk)
    # hand-assembled expectation: declaration order of ark-groth16's structs, u64-LE Vec prefixes
    c1 = lambda a: b"".join(o.g1_compress(p) for p in layout.arr_to_g1(np.asarray(a).reshape(-1, 8)))
    c2 = lambda a: b"".join(o.g2_compress(p) for p in layout.arr_to_g2(np.asarray(a).reshape(-1, 16)))
    u64 = lambda v: int(v).to_bytes(8, "little")
    exp = (c1(vk.alpha_g1) + c2(vk.beta_g2) + c2(vk.gamma_g2) + c2(vk.delta_g2) + u64(n_inputs + 1) + c1(vk.gamma_abc_g1)
           + c1(pk.beta_g1) + c1(pk.delta_g1) + u64(n_vars) + c1(pk.a_query) + u64(n_vars) + c1(pk.b_g1_query)
           + u64(n_vars) + c2(pk.b_g2_query) + u64(m - 1) + c1(pk.h_query) + u64(n_vars - n_inputs) + c1(pk.l_query))
    assert buf == exp
    back = ark.deserialize_proving_key(net, buf)
    for name in ("beta_g1", "delta_g1", "a_query", "b_g1_query", "b_g2_query", "h_query", "l_query"):
        assert (getattr(back, name) == getattr(pk, name)).all(), name
    assert (back.vk.gamma_abc_g1 == vk.gamma_abc_g1).all() and (back.vk.delta_g2 == vk.delta_g2).all()
    for bad in (buf[:-1], buf + b"\0", buf[:40]):
        with pytest.raises(ValueError):
            ark.deserialize_proving_key(net, bad)
    # the reference's golden proof through the same container code
    blob = open(os.path.join(ROOT, "tests", "golden", "sha256_proof.bin"), "rb").read()
    a, b, c = ark.deserialize_proof(net, blob)
    assert ark.serialize_proof(net, a, b, c) == blob
    with pytest.raises(ValueError):
        ark.deserialize_proof(net, blob[:-1])


def test_bench_reference_arm_prints_exactly_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, out.stdout[:500]
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Mpairs/s" and d["higher_is_better"] is True
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for key in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert key in d, key
