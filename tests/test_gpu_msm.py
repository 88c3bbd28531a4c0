"""d_msm parity: CUDA Pippenger (csrc/msm.cu) vs the oracle, through the C ABI.

Mirrors dist-primitives/examples/dmsm_test.rs:49-64 (d_msm == G::msm on the public vectors) and
dmsm/mod.rs:147-193 (all scalars = 1), instantiated for BN254 as BASELINE config 1 asks."""
import numpy as np
import pytest

from distributed_groth16_b200 import MpcNetError
from distributed_groth16_b200.dist_primitives import d_msm

pytestmark = pytest.mark.gpu


def _limbs(v):
    return [(v >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]


def _check(net, cref, bases, scalars, g2):
    got = d_msm(bases, scalars, None, net)
    exp, inf = (cref.msm_g2 if g2 else cref.msm_g1)(bases, scalars)
    assert got.infinity == inf
    assert (got.limbs == exp).all()
    return got


@pytest.mark.parametrize("g2", [False, True])
@pytest.mark.parametrize("n", [1, 2, 3, 17, 100, 1000])
def test_small_sizes(net, cref, n, g2):
    bases = (cref.g2_generate if g2 else cref.g1_generate)(0xB2000001, n)
    scalars = cref.fr_generate(42 + n, n)
    _check(net, cref, bases, scalars, g2)


@pytest.mark.parametrize("g2", [False, True])
def test_config1_4096_with_edge_cases(net, cref, g2):
    """BASELINE config 1: 2^12 pairs + forced edge cases (0, 1, r-1, 2^253, duplicates, P and -P, infinity base)."""
    from oracle import layout, bn254 as o
    n = 4096
    bases = (cref.g2_generate if g2 else cref.g1_generate)(0xB2000001, n)
    scalars = cref.fr_generate(0xB2000001, n)
    sc = layout.fr_to_arr([0, 1, o.R - 1, 1 << 253, 5, 5, 7, 7, 9])
    scalars[:9] = sc
    w = bases.shape[1]
    bases[4] = bases[5]                                   # duplicate point, same scalar -> P + P inside a bucket
    pts = (layout.arr_to_g2 if g2 else layout.arr_to_g1)(bases[6:7])
    neg = (o.G2 if g2 else o.G1).neg(pts[0])
    bases[7] = (layout.g2_to_arr if g2 else layout.g1_to_arr)([neg])[0]   # P and -P with the same scalar -> cancels
    bases[8] = np.zeros(w, dtype=np.uint64)               # infinity base (zkey convention)
    _check(net, cref, bases, scalars, g2)


@pytest.mark.parametrize("g2", [False, True])
def test_all_scalars_one_and_all_zero(net, cref, g2):
    """dmsm/mod.rs:173-193 uses scalars = 1 for every base."""
    from oracle import layout
    n = 256
    bases = (cref.g2_generate if g2 else cref.g1_generate)(3, n)
    _check(net, cref, bases, layout.fr_to_arr([1] * n), g2)
    got = _check(net, cref, bases, layout.fr_to_arr([0] * n), g2)
    assert got.infinity


def test_skewed_witness_like_scalars(net, cref):
    """Real witnesses are mostly 0/1/small: giant buckets must still give the right point."""
    from oracle import layout
    n = 5000
    rng = np.random.default_rng(5)
    vals = [int(v) for v in rng.choice([0, 1, 1, 1, 2, 3, 255, 65535], size=n)]
    bases = cref.g1_generate(9, n)
    _check(net, cref, bases, layout.fr_to_arr(vals), False)


def test_empty_and_length_mismatch(net, cref):
    bases = cref.g1_generate(1, 8)
    scalars = cref.fr_generate(1, 8)
    got = d_msm(bases[:0], scalars[:0], None, net, g2=False)
    assert got.infinity
    with pytest.raises(MpcNetError) as ei:
        d_msm(bases, scalars[:5], None, net)
    assert ei.value.kind == "Generic" and ei.value.message == "5"     # arkworks Err(min_len) -> to_string()


def test_g1_2_16_vs_oracle(net, cref):
    n = 1 << 16
    bases = cref.g1_generate(0xB2000002, n)
    scalars = cref.fr_generate(0xB2000002, n)
    _check(net, cref, bases, scalars, False)


def test_generator_kernels_match_oracle(net, cref):
    g1 = net.generate_g1(123, 300).cpu().numpy().view(np.uint64)
    g2 = net.generate_g2(123, 100).cpu().numpy().view(np.uint64)
    fr = net.generate_fr(123, 1000).cpu().numpy().view(np.uint64)
    assert (g1 == cref.g1_generate(123, 300)).all()
    assert (g2 == cref.g2_generate(123, 100)).all()
    assert (fr == cref.fr_generate(123, 1000)).all()


def test_config2_g1_2_20_device_resident(net, cref):
    """BASELINE config 2: 2^20 pairs, inputs generated and kept in HBM; oracle on the same inputs."""
    import torch
    n = 1 << 20
    bases = net.generate_g1(0xB2000002, n)
    scalars = net.generate_fr(0xB2000002, n)
    got = d_msm(bases, scalars, None, net)
    bh = bases.cpu().numpy().view(np.uint64)
    sh = scalars.cpu().numpy().view(np.uint64)
    assert cref.g1_on_curve(bh[:4096])
    exp, inf = cref.msm_g1(bh, sh)
    assert not inf and (got.limbs == exp).all()
    # linearity: MSM(P, s) + MSM(P, t) == MSM(P, s + t)
    t = net.generate_fr(77, n)
    st = torch.from_numpy(net.field_op(1, 1, sh, t.cpu().numpy().view(np.uint64)).view(np.int64)).to(bases.device)
    lhs = d_msm(bases, st, None, net)
    part = d_msm(bases, t, None, net)
    both = np.stack([got.limbs, part.limbs])
    from oracle import layout
    exp_sum, _ = cref.msm_g1(both, layout.fr_to_arr([1, 1]))
    assert (lhs.limbs == exp_sum).all()


def test_larger_sizes_2_22_vs_oracle_and_2_24_linearity(net, cref):
    """Sizes above the benchmark point (BASELINE config 2 lists 2^22..2^26): 2^22 against the CPU twin; 2^24 through a
    size-independent property: MSM(P, s) + MSM(P, t) == MSM(P, s + t)."""
    import torch
    from oracle import layout
    n = 1 << 22
    bases = net.generate_g1(0xB2000022, n)
    scalars = net.generate_fr(0xB2000022, n)
    got = d_msm(bases, scalars, None, net)
    exp, inf = cref.msm_g1(bases.cpu().numpy().view(np.uint64), scalars.cpu().numpy().view(np.uint64))
    assert not inf and (got.limbs == exp).all()
    del bases, scalars
    n = 1 << 24
    bases = net.generate_g1(0xB2000024, n)
    s = net.generate_fr(1, n)
    t = net.generate_fr(2, n)
    st = torch.from_numpy(net.field_op(1, 1, s.cpu().numpy().view(np.uint64), t.cpu().numpy().view(np.uint64)).view(np.int64)).to(bases.device)
    a, b, c = d_msm(bases, s, None, net), d_msm(bases, t, None, net), d_msm(bases, st, None, net)
    both = np.stack([a.limbs, b.limbs])
    exp_sum, _ = cref.msm_g1(both, layout.fr_to_arr([1, 1]))
    assert (c.limbs == exp_sum).all() and not c.infinity


@pytest.mark.parametrize("g2", [False, True])
@pytest.mark.parametrize("n,c", [(1, 9), (37, 4), (3000, 8), (3000, 13), (20000, 15)])
def test_fixed_base_tables_give_the_same_point(net, cref, n, c, g2):
    """b200zk_msm_table_* (one bucket set over 2^{cw} P_i tables) vs the oracle, edge cases included."""
    import torch
    from oracle import layout, bn254 as o
    bases = (cref.g2_generate if g2 else cref.g1_generate)(0xB2000007 + c, n)
    scalars = cref.fr_generate(0xB2000008 + n, n)
    if n >= 37:
        scalars[:6] = layout.fr_to_arr([0, 1, o.R - 1, 1 << 253, (1 << 253) + (1 << 252) - 1, 5])
        bases[7] = 0                                      # infinity base stays infinity in every window
        bases[9] = bases[10]
        scalars[9] = scalars[10]                          # equal points meet in every bucket they share
    db = torch.from_numpy(bases.view(np.int64)).cuda()
    ds = torch.from_numpy(scalars.view(np.int64)).cuda()
    table = net.msm_table_build(db, c, g2=g2)
    w = net.msm_table_windows(c)
    assert w == -(-255 // c) and table.shape[0] == w * n
    assert (table[:n].cpu().numpy().view(np.uint64) == bases).all()
    if not g2 and n >= 37:                                # window 1 of point 0 = 2^c * P_0
        exp1, _ = cref.msm_g1(bases[:1], layout.fr_to_arr([1 << c]))
        assert (table[n].cpu().numpy().view(np.uint64) == exp1).all()
        assert not table[n + 7].any() and not table[(w - 1) * n + 7].any()
    part = net.msm_table_dev(table, ds, c, g2=g2)
    got, inf = net.sum_points_dev(part, 1, g2=g2)
    exp, einf = (cref.msm_g2 if g2 else cref.msm_g1)(bases, scalars)
    assert inf == einf and (got == exp).all()
    with pytest.raises(MpcNetError):
        net.msm_table_dev(table, ds[: n - 1] if n > 1 else ds[:0], c, g2=g2)


def test_fixed_base_tables_2_20_match_generic_path(net):
    """BASELINE config 2 size: table path == generic path (both normalised), c = 20 as the proving key picks."""
    n = 1 << 20
    bases = net.generate_g1(0xB2000002, n)
    scalars = net.generate_fr(0xB2000002, n)
    a, ainf = net.sum_points_dev(net.msm_dev(bases, scalars), 1)
    table = net.msm_table_build(bases, 20)
    b, binf = net.sum_points_dev(net.msm_table_dev(table, scalars, 20), 1)
    assert not ainf and not binf and (a == b).all()


def test_g2_2_16_and_2_18_vs_oracle(net, cref):
    """Standalone G2 MSM against the CPU twin above the prover-sized cases (no G2 oracle test >= 2^16)."""
    for log_n in (16, 18):
        n = 1 << log_n
        bases = net.generate_g2(0xB2000016 + log_n, n)
        scalars = net.generate_fr(0xB2000016 + log_n, n)
        got = d_msm(bases, scalars, None, net, g2=True)
        exp, inf = cref.msm_g2(bases.cpu().numpy().view(np.uint64), scalars.cpu().numpy().view(np.uint64))
        assert not inf and not got.infinity and (got.limbs == exp).all(), log_n


def _fr_add_dev(net, a, b):
    """a + b on the device (slabs through the C ABI's element-wise hook on host buffers would move gigabytes: use the
    pointwise kernel a * 1 - (-b) = a + b instead)."""
    import torch
    from oracle import layout
    from distributed_groth16_b200._native import c_vp
    from distributed_groth16_b200._constants import FR_ONE_MONT
    n = int(a.shape[0])
    one = torch.from_numpy(np.tile(np.array(FR_ONE_MONT, dtype=np.uint64), (n, 1)).view(np.int64)).to(a.device)
    zero = torch.zeros_like(a)
    negb = torch.empty_like(a)
    out = torch.empty_like(a)
    lib, h = net._lib, net._h
    net.check(lib.b200zk_fr_mul_sub_dev(h, 0, c_vp(zero.data_ptr()), c_vp(one.data_ptr()), c_vp(b.data_ptr()), c_vp(negb.data_ptr()), n))   # 0 * 1 - b
    net.check(lib.b200zk_fr_mul_sub_dev(h, 0, c_vp(a.data_ptr()), c_vp(one.data_ptr()), c_vp(negb.data_ptr()), c_vp(out.data_ptr()), n))  # a * 1 - (-b)
    return out


def test_2_26_linearity_generic_and_window_groups(net, cref):
    """north_star's largest size, 2^26 pairs (4 GiB of points): MSM(P, s) + MSM(P, t) == MSM(P, s + t), everything resident.
    At this size the MSM runs as a 4-group window pipeline (csrc/msm.cu): the property exercises every group's tables."""
    from oracle import layout
    n = 1 << 26
    bases = net.generate_g1(0xB2000026, n)
    s, t = net.generate_fr(11, n), net.generate_fr(12, n)
    st = _fr_add_dev(net, s, t)
    # spot-check the device addition against the oracle's field addition on a slice
    sl = slice(12345, 12345 + 257)
    from oracle import bn254 as o
    want = layout.fr_to_arr([(x + y) % o.R for x, y in zip(layout.arr_to_fr(s[sl].cpu().numpy().view(np.uint64)),
                                                          layout.arr_to_fr(t[sl].cpu().numpy().view(np.uint64)))])
    assert (st[sl].cpu().numpy().view(np.uint64) == want).all()
    a, b, c = d_msm(bases, s, None, net), d_msm(bases, t, None, net), d_msm(bases, st, None, net)
    exp_sum, _ = cref.msm_g1(np.stack([a.limbs, b.limbs]), layout.fr_to_arr([1, 1]))
    assert not c.infinity and (c.limbs == exp_sum).all()


def test_host_staged_parts_share_one_bucket_set(net, cref):
    """b200zk_msm_g1 on host buffers >= 2^18 pairs travels in 4 parts whose bucket kernels add into ONE bucket set
    (csrc/msm.cu, read-modify-write of the XYZZ buckets): random scalars, then a 0/1-heavy witness-like vector (giant
    buckets: the multi-task merges of every part add into the shared buckets), equal points and opposite points that only
    meet ACROSS parts (P + P and P + (-P) through the read-modify-write path), an infinity base and zero scalars."""
    from oracle import layout, bn254 as o
    n = (1 << 18) + 3
    bases = cref.g1_generate(0xB2000018, n)
    scalars = cref.fr_generate(0xB2000018, n)
    q = n // 4
    bases[q + 11] = bases[11]                              # same point in parts 0 and 1 ...
    scalars[q + 11] = scalars[11]                          # ... with the same scalar: doubling inside a shared bucket
    neg = layout.g1_to_arr([o.G1.neg(layout.arr_to_g1(bases[12:13])[0])])[0]
    bases[3 * q + 12] = neg                                # P in part 0, -P in part 3, same scalar: the bucket empties again
    scalars[3 * q + 12] = scalars[12]
    bases[2 * q + 5] = 0                                   # infinity base
    scalars[7] = 0
    got = d_msm(bases, scalars, None, net)
    exp, inf = cref.msm_g1(bases, scalars)
    assert got.infinity == inf and (got.limbs == exp).all()
    rng = np.random.default_rng(11)
    small = layout.fr_to_arr([int(v) for v in rng.choice([0, 1, 1, 1, 1, 2, 3, 65535], size=n)])
    got = d_msm(bases, small, None, net)
    exp, inf = cref.msm_g1(bases, small)
    assert got.infinity == inf and (got.limbs == exp).all()
    b2, s2 = cref.g2_generate(0xB2000019, n), cref.fr_generate(0xB2000019, n)
    got = d_msm(b2, s2, None, net, g2=True)
    exp, inf = cref.msm_g2(b2, s2)
    assert got.infinity == inf and (got.limbs == exp).all()
