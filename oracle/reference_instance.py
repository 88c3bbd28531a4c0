"""TEST INFRASTRUCTURE (oracle): BASELINE config 4's instance rebuilt on the CPU from the reference's seed.

`sha256_instance` turns the seed `[42u8; 32]` of /root/reference/mpc-api/src/main.rs:148-152, the reference's sha256 circuit
(fixtures/sha256/sha256.r1cs, committed as tests/golden/sha256_circuit.npz) and the witness of {a: 1, b: 2} into the proving /
verifying key of the reference's setup (oracle/ark_rand.py: rand's StdRng + arkworks' sampling order; C++ twin: fixed-base
multiplications) and the QAP evaluation vectors (groth16/src/qap.rs:44-91).  Used by tests/ and by bench.py's `cpu_baseline`
leg (the CPU restatement timed beside the GPU prover on config 4); never by the product path."""
import os

import numpy as np

from . import ark_rand as ar, bn254 as o, layout

SEED = bytes([42] * 32)


def sha256_instance(cref, golden_dir, seed=SEED):
    """-> (pk arrays, vk points dict, z, a, b, c as Montgomery limb arrays, dims) of the reference's seeded setup."""
    d = np.load(os.path.join(golden_dir, "sha256_circuit.npz"))
    n_vars, n_pub, n_cons = (int(x) for x in d["dims"])
    n_inputs = n_pub + 1
    m = 1
    while m < n_cons + n_inputs:
        m <<= 1
    ints = lambda arr: [int.from_bytes(arr[i].tobytes(), "little") for i in range(arr.shape[0])]
    z = ints(d["witness"])
    coo = lambda k: (d[k + "_rows"], d[k + "_cols"], ints(d[k + "_vals"]))
    tw = ar.groth16_toxic_waste(seed, m)
    qs = ar.query_scalars(tw, coo("a"), coo("b"), coo("c"), n_vars, n_inputs, n_cons, m)
    g1, g2 = layout.g1_to_arr([tw["g1"]])[0], layout.g2_to_arr([tw["g2"]])[0]
    fb = lambda sc, g2_=False: cref.fixed_base_mul(g2 if g2_ else g1, layout.fr_to_arr(sc), g2=g2_)
    pk = dict(a_query=fb(qs["a"]), b_g1_query=fb(qs["b"]), b_g2_query=fb(qs["b"], True), l_query=fb(qs["l"]), h_query=fb(qs["h"]))
    vk1 = fb([tw["alpha"], tw["beta"], tw["delta"]])                       # alpha_g1, beta_g1, delta_g1
    vk2 = fb([tw["beta"], tw["delta"], tw["gamma"]], True)                 # beta_g2, delta_g2, gamma_g2
    vk = dict(alpha_g1=vk1[0], beta_g1=vk1[1], delta_g1=vk1[2], beta_g2=vk2[0], delta_g2=vk2[1], gamma_g2=vk2[2],
              gamma_abc_g1=fb(qs["gamma_abc"]))
    # QAP evaluation vectors (groth16/src/qap.rs:44-91): a = A z with z_i on the input rows, b = B z, c = a . b
    a, b = [0] * m, [0] * m
    for vec, (rows, cols, vals) in ((a, coo("a")), (b, coo("b"))):
        for r_, c_, v in zip(rows, cols, vals):
            vec[int(r_)] = (vec[int(r_)] + v * z[int(c_)]) % o.R
    for i in range(n_inputs):
        a[n_cons + i] = z[i]
    c = [x * y % o.R for x, y in zip(a, b)]
    return pk, vk, layout.fr_to_arr(z), layout.fr_to_arr(a), layout.fr_to_arr(b), layout.fr_to_arr(c), (n_vars, n_inputs, m)
