"""Test-side writers for snarkjs artefacts (zkey / wtns) so that the product readers can be exercised on the GPU box,
where /root/reference does not exist.  The byte layout follows ark-circom/src/zkey.rs:53-387 (sections 1-9)."""
import struct

Q = 21888242871839275222246405745257275088696311157297823662689037894645226208583
R = 21888242871839275222246405745257275088548364400416034343698204186575808495617


def _sec(sid, payload):
    return struct.pack("<IQ", sid, len(payload)) + payload


def write_zkey(d) -> bytes:
    """d: the npz produced by tests/golden/make_golden.py (limb arrays; coefficient values in Montgomery form)."""
    from oracle import bn254 as o, layout
    n_vars, n_public, m, nc = (int(x) for x in d["dims"])
    hdr = struct.pack("<I", 32) + Q.to_bytes(32, "little") + struct.pack("<I", 32) + R.to_bytes(32, "little")
    hdr += struct.pack("<III", n_vars, n_public, m)
    vk1, vk2 = d["vk_g1"], d["vk_g2"]           # vk_g1 = alpha, beta1, delta1 ; vk_g2 = beta2, delta2, gamma2
    hdr += vk1[0].tobytes() + vk1[1].tobytes() + vk2[0].tobytes() + vk2[2].tobytes() + vk1[2].tobytes() + vk2[1].tobytes()
    # coefficients: file stores value * R^2; the npz holds Montgomery form (value * R) -> multiply by R once more
    coefs = []
    for mi, key in ((0, "a"), (1, "b")):
        vals = layout.arr_to_fr(d[key + "_vals"])
        for r_, c_, v in zip(d[key + "_rows"], d[key + "_cols"], vals):
            coefs.append(struct.pack("<III", mi, int(r_), int(c_)) + (v * o.MONT_R * o.MONT_R % o.R).to_bytes(32, "little"))
    # snarkjs appends the public-input rows (constraint index nc + j): they fix max_constraint_index (zkey.rs:171)
    for j in range(n_public + 1):
        coefs.append(struct.pack("<III", 0, nc + j, j) + (o.MONT_R * o.MONT_R % o.R).to_bytes(32, "little"))
    sec4 = struct.pack("<I", len(coefs)) + b"".join(coefs)
    body = [_sec(1, struct.pack("<I", 1)), _sec(2, hdr), _sec(3, d["ic"].tobytes()), _sec(4, sec4),
            _sec(5, d["a_query"].tobytes()), _sec(6, d["b_g1_query"].tobytes()), _sec(7, d["b_g2_query"].tobytes()),
            _sec(8, d["l_query"].tobytes()), _sec(9, d["h_query"].tobytes())]
    return b"zkey" + struct.pack("<II", 1, len(body)) + b"".join(body)


def write_wtns(values) -> bytes:
    s1 = struct.pack("<I", 32) + R.to_bytes(32, "little") + struct.pack("<I", len(values))
    s2 = b"".join(int(v).to_bytes(32, "little") for v in values)
    return b"wtns" + struct.pack("<II", 2, 2) + _sec(1, s1) + _sec(2, s2)


def f1_witness(n_vars):
    z = [0] * n_vars
    z[0], z[2] = 1, 3
    for i in range(3, n_vars):
        z[i] = z[i - 1] * z[i - 1] % R
    z[1] = z[n_vars - 1] ** 2 % R
    return z
