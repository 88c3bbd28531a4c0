"""`b200zk_group_*` (csrc/group.cu): the sharded hot path behind single C calls, vs the oracle and vs the single-GPU calls.

World 1 runs on any box (the peer stores degenerate to local stores, every layout map and event barrier is still
exercised); on a multi-GPU box (`gpurun --gpus N`) the same assertions run over all GPUs -- BASELINE config 5's path
(MSM split + four-step NTT exchange + sharded prove) reached from ONE call, as a Rust binding at
groth16/src/prove.rs:106-136 would."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worlds():
    import torch
    n = torch.cuda.device_count()
    out = [1]
    w = 2
    while w <= min(n, 8):
        out.append(w)
        w *= 2
    return out


@pytest.fixture(scope="module", params=["world1", "all"])
def group(request):
    from distributed_groth16_b200.group import Group
    worlds = _worlds()
    if request.param == "all" and worlds[-1] == 1:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus N)")
    w = 1 if request.param == "world1" else worlds[-1]
    g = Group(list(range(w)))
    yield g
    g.close()


def test_group_d_msm_matches_oracle(group, cref):
    for n in (0, 1, 5, 1000, 1 << 14):
        if n == 0:
            got = group.d_msm(np.zeros((0, 8), dtype=np.uint64), np.zeros((0, 4), dtype=np.uint64), g2=False)
            assert got.infinity
            continue
        bases, scalars = cref.g1_generate(0xB2000001, n), cref.fr_generate(0xB2000001, n)
        if n >= 5:
            scalars[0] = 0                       # zero scalar
            bases[1] = 0                         # infinity base
            bases[3] = bases[2]                  # duplicate base
        got = group.d_msm(bases, scalars)
        exp, inf = cref.msm_g1(bases, scalars)
        assert got.infinity == bool(inf) and (got.infinity or (got.limbs == exp).all()), n
    b2, s2 = cref.g2_generate(7, 3000), cref.fr_generate(8, 3000)
    got = group.d_msm(b2, s2)
    exp, inf = cref.msm_g2(b2, s2)
    assert not got.infinity and (got.limbs == exp).all()


def test_group_d_msm_length_mismatch_is_the_reference_error(group, cref):
    from distributed_groth16_b200 import MpcNetError
    with pytest.raises(MpcNetError) as e:
        group.d_msm(cref.g1_generate(1, 7), cref.fr_generate(1, 5))
    assert e.value.kind == "Generic" and e.value.message == "5"          # arkworks Err(min_len) through `?`, dmsm/mod.rs:82


def test_group_d_fft_and_d_ifft_match_oracle(group, cref):
    for log_n in (6, 10, 13, 16):
        if (1 << (log_n // 2)) < group.size():
            continue
        x = cref.fr_generate(0xB2000003 + log_n, 1 << log_n)
        y = group.d_fft(x)
        assert (y == cref.ntt(x)).all(), log_n
        assert (group.d_ifft(y) == x).all(), log_n


def test_group_h_matches_oracle(group, cref):
    for log_m in (6, 11, 15):
        if (1 << (log_m // 2)) < group.size():
            continue
        a, b, c = (cref.fr_generate(s + log_m, 1 << log_m) for s in (21, 22, 23))
        assert (group.h(a, b, c) == cref.h_circom(a, b, c)).all(), log_m


@pytest.mark.parametrize("rs", [(0, 0), (777, 999)])
def test_group_prove_equals_single_gpu_prover_and_oracle(group, cref, net, rs):
    from oracle import layout
    from distributed_groth16_b200.group import GroupProvingKey
    from distributed_groth16_b200.groth16 import ProvingKey, prove
    from distributed_groth16_b200._constants import FR_ONE_MONT
    m, n_vars, n_inputs = 1 << 12, 3001, 3
    aq, b1, lq, hq = (cref.g1_generate(s, k) for s, k in ((31, n_vars), (32, n_vars), (34, n_vars - n_inputs), (35, m)))
    b2, vk1, vk2 = cref.g2_generate(33, n_vars), cref.g1_generate(36, 3), cref.g2_generate(37, 2)
    z = cref.fr_generate(38, n_vars)
    z[0] = np.array(FR_ONE_MONT, dtype=np.uint64)
    a, b, c = (cref.fr_generate(s, m) for s in (39, 40, 41))
    r, s = (layout.fr_to_arr([v])[0] for v in rs)
    vk = np.concatenate([vk1.reshape(-1), vk2.reshape(-1)])
    want = cref.groth16_prove(aq, b1, b2, lq, hq, vk, n_inputs, z, cref.h_circom(a, b, c), r, s, mirror_bg1=False)
    gpk = GroupProvingKey(group, aq, b1, b2, lq, hq, n_inputs, vk1[0], vk1[1], vk1[2], vk2[0], vk2[1])
    assert gpk.table_bytes > 0
    got = gpk.create_proof(z, a, b, c, r, s)
    gpk.free()
    assert got == want
    pk = ProvingKey(net, aq, b1, b2, lq, hq, n_inputs, vk1[0], vk1[1], vk1[2], vk2[0], vk2[1])
    assert prove.create_proof(pk, z, a, b, c, r, s) == got
    pk.free()
