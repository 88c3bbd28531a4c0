"""Device field arithmetic (csrc/fp.cuh) vs the oracle, through the C ABI self-test hook."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _edge_and_random(cref, field, n):
    c = cref.constants()
    mod = c["q"] if field == 0 else c["r"]
    rng = np.random.default_rng(1234 + field)
    vals = [0, 1, 2, mod - 1, mod - 2, (mod - 1) // 2, (mod + 1) // 2, (1 << 253) % mod, (1 << 128) - 1, (1 << 64)]
    while len(vals) < n:
        vals.append(int.from_bytes(rng.bytes(32), "little") % mod)
    arr = np.array([[(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)] for v in vals], dtype=np.uint64)
    return arr


@pytest.mark.parametrize("field", [0, 1])
@pytest.mark.parametrize("op", [0, 1, 2])
def test_field_ops_bit_exact(net, cref, field, op):
    n = 4096
    a = _edge_and_random(cref, field, n)
    b = np.roll(_edge_and_random(cref, field, n), 7, axis=0).copy()
    got = net.field_op(field, op, a, b)
    exp = np.stack([cref.field_op(field, op, a[i], b[i]) for i in range(n)])
    assert (got == exp).all()
