"""The C++ CPU twin (oracle/bn254_ref.cpp) against the Python big-int oracle (oracle/bn254.py)."""
import random

from oracle import bn254 as o, layout as L


def test_constants(cref):
    c = cref.constants()
    assert c["q"] == o.P and c["r"] == o.R
    assert c["inv_q"] == o.mont_inv64(o.P) == 0x87d20782e4866389
    assert c["inv_r"] == o.mont_inv64(o.R) == 0xc2e1f593efffffff
    assert c["r1_q"] == o.MONT_R % o.P and c["r1_r"] == o.MONT_R % o.R
    assert c["r2_q"] == o.MONT_R**2 % o.P and c["r2_r"] == o.MONT_R**2 % o.R


def test_field_ops(cref):
    rnd = random.Random(1)
    for field, mod, enc, dec in ((1, o.R, L.fr_to_arr, L.arr_to_fr),):
        for _ in range(300):
            a, b = rnd.randrange(mod), rnd.randrange(mod)
            A, B = enc([a])[0], enc([b])[0]
            assert dec(cref.field_op(field, 0, A, B)) == [a * b % mod]
            assert dec(cref.field_op(field, 1, A, B)) == [(a + b) % mod]
            assert dec(cref.field_op(field, 2, A, B)) == [(a - b) % mod]
        assert dec(cref.field_op(field, 3, A)) == [pow(a, -1, mod)]


def test_generated_points_and_msm(cref):
    rnd = random.Random(2)
    g1 = cref.g1_generate(7, 60)
    g2 = cref.g2_generate(7, 24)
    assert cref.g1_on_curve(g1) and cref.g2_on_curve(g2)
    p1, p2 = L.arr_to_g1(g1), L.arr_to_g2(g2)
    assert all(o.G1.is_on_curve(p) for p in p1) and all(o.G2.is_on_curve(p) for p in p2)
    sc = [rnd.randrange(o.R) for _ in range(60)]
    sc[3], sc[4], sc[5] = 0, 1, o.R - 1
    S = L.fr_to_arr(sc)
    r, inf = cref.msm_g1(g1, S)
    assert not inf and L.arr_to_g1(r)[0] == o.G1.msm_naive(p1, sc)
    assert (cref.msm_g1_naive(g1, S)[0] == r).all()
    r2, _ = cref.msm_g2(g2, S[:24])
    assert L.arr_to_g2(r2)[0] == o.G2.msm_naive(p2, sc[:24])
    # pippenger path (n >= 32 switches the window rule) vs double-and-add on 3000 points
    g = cref.g1_generate(9, 3000)
    s = cref.fr_generate(5, 3000)
    assert (cref.msm_g1(g, s)[0] == cref.msm_g1_naive(g, s)[0]).all()
    # infinity bases and cancelling pairs
    g[10] = 0
    neg = L.g1_to_arr([o.G1.neg(L.arr_to_g1(g[20:21])[0])])[0]
    g[21] = neg
    s[21] = s[20]
    assert (cref.msm_g1(g, s)[0] == cref.msm_g1_naive(g, s)[0]).all()


def test_ntt_and_h(cref):
    rnd = random.Random(3)
    for lg in (0, 1, 2, 5, 9):
        v = [rnd.randrange(o.R) for _ in range(1 << lg)]
        V = L.fr_to_arr(v)
        for inv in (False, True):
            for cos in (False, True):
                exp = (o.intt if inv else o.ntt)(v, coset=cos)
                assert L.arr_to_fr(cref.ntt(V, inv, cos)) == exp, (lg, inv, cos)
        assert L.arr_to_fr(cref.bitrev(V)) == o.bit_reverse_permute(v)
    v = [rnd.randrange(o.R) for _ in range(16)]
    assert o.ntt(v) == o.dft_naive(v) and o.intt(v) == o.dft_naive(v, True)
    a = [rnd.randrange(o.R) for _ in range(64)]
    b = [rnd.randrange(o.R) for _ in range(64)]
    c = [x * y % o.R for x, y in zip(a, b)]
    assert L.arr_to_fr(cref.h_circom(L.fr_to_arr(a), L.fr_to_arr(b), L.fr_to_arr(c))) == o.h_circom(a, b, c)


def test_reference_local_dfft_flow_equals_plain_dft():
    """Transliteration check of dist-primitives/examples/local_dfft_test.rs:26-83 (fft1 + fft2 with the
    off-by-one twiddles and rotate_right(1)) == dom.fft(x), for (m, l) = (8, 2), (32, 2), (64, 4)."""
    rnd = random.Random(4)
    for m, l in ((8, 2), (32, 2), (64, 4)):
        x = [rnd.randrange(o.R) for _ in range(m)]
        gen = o.fr_root_of_unity(m)
        xb = o.bit_reverse_permute(x)
        # stride packing without PSS: party vector px[i] holds group i = xb[i], xb[i + m/l], ... (dfft/mod.rs:307-318)
        px = [[xb[i + k * (m // l)] for k in range(l)] for i in range(m // l)]
        lg_m, lg_l = m.bit_length() - 1, l.bit_length() - 1
        # fft1 on each of the l "columns" (dfft/mod.rs:122-135)
        for col in range(l):
            v = [px[i][col] for i in range(m // l)]
            for i in range(lg_m, lg_l, -1):
                poly = m // (1 << i)
                fs = pow(gen, 1 << (i - 1), o.R)
                f = fs
                for k in range(poly):
                    for j in range((1 << (i - 1)) // l):
                        xx = v[(2 * j) * poly + k]
                        yy = v[(2 * j + 1) * poly + k] * f % o.R
                        v[j * 2 * poly + k] = (xx + yy) % o.R
                        v[j * 2 * poly + k + poly] = (xx - yy) % o.R
                    f = f * fs % o.R
            for i in range(m // l):
                px[i][col] = v[i]
        s1 = [px[i][j] for i in range(m // l) for j in range(l)]              # dfft/mod.rs:216-218
        for i in range(lg_l, 0, -1):                                           # fft2 (dfft/mod.rs:161-175)
            poly = m // (1 << i)
            fs = pow(gen, 1 << (i - 1), o.R)
            f = fs
            s2 = [0] * m
            for k in range(poly):
                for j in range(1 << (i - 1)):
                    xx = s1[k * (1 << i) + 2 * j]
                    yy = s1[k * (1 << i) + 2 * j + 1] * f % o.R
                    s2[k * (1 << (i - 1)) + j] = (xx + yy) % o.R
                    s2[(k + poly) * (1 << (i - 1)) + j] = (xx - yy) % o.R
                f = f * fs % o.R
            s1 = s2
        s1 = s1[-1:] + s1[:-1]                                                 # rotate_right(1) (:177)
        assert s1 == o.ntt(x), (m, l)
