"""Re-entrancy of the C ABI: the reference keeps three d_msm / d_fft calls in flight at once, told apart by their
MultiplexedStreamID (groth16/src/prove.rs:119-125 `join!` of three d_msm; ext_wit.rs:34-49 three d_ifft / d_fft).  Three host
threads issue calls on the three stream slots concurrently (ctypes releases the GIL for the duration of a call); every result
must equal the oracle's, run after run."""
import threading

import numpy as np
import pytest

from distributed_groth16_b200.dist_primitives import d_fft, d_ifft, d_msm, fft_in_place_rearrange

pytestmark = pytest.mark.gpu


def test_three_streams_from_three_host_threads(net, cref):
    jobs = []
    for sid in range(3):
        n = 3000 + 517 * sid
        g1b, g1s = cref.g1_generate(900 + sid, n), cref.fr_generate(910 + sid, n)
        g2b, g2s = cref.g2_generate(920 + sid, 700), cref.fr_generate(930 + sid, 700)
        x = cref.fr_generate(940 + sid, 1 << (11 + sid))
        jobs.append(dict(sid=sid, g1=(g1b, g1s), g2=(g2b, g2s), x=fft_in_place_rearrange(x),      # the protocol's input order
                         exp_g1=cref.msm_g1(g1b, g1s), exp_g2=cref.msm_g2(g2b, g2s),
                         exp_f=cref.ntt(x), exp_i=cref.ntt(x, inverse=True)))
    errors = []
    start = threading.Barrier(3)

    def worker(j):
        try:
            start.wait()
            for rep in range(6):
                got = d_msm(j["g1"][0], j["g1"][1], None, net, sid=j["sid"])
                assert got.infinity == j["exp_g1"][1] and (got.limbs == j["exp_g1"][0]).all(), ("g1", j["sid"], rep)
                out = d_fft(j["x"], False, 1, False, j["x"].shape[0], None, net, sid=j["sid"])
                assert (out == j["exp_f"]).all(), ("fft", j["sid"], rep)
                got = d_msm(j["g2"][0], j["g2"][1], None, net, sid=j["sid"], g2=True)
                assert got.infinity == j["exp_g2"][1] and (got.limbs == j["exp_g2"][0]).all(), ("g2", j["sid"], rep)
                out = d_ifft(j["x"], False, 1, False, j["x"].shape[0], None, net, sid=j["sid"])
                assert (out == j["exp_i"]).all(), ("ifft", j["sid"], rep)
        except BaseException as e:      # noqa: BLE001 -- reported by the main thread
            errors.append(repr(e))

    threads = [threading.Thread(target=worker, args=(j,)) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join(timeout=300)
    assert not any(t.is_alive() for t in threads), "a worker thread hung"
    assert not errors, errors
