"""Pins of the draw contract on the CPU oracle: Philox4x32-10 against the Random123 known-answer
vectors, the contract logarithm against libm, the uniform maps."""
import math

import numpy as np


def test_philox_known_answers(oracle):
    # Random123 kat_vectors: philox4x32 10 <ctr x4> <key x2> -> <out x4>
    kat = [
        ([0, 0, 0, 0], [0, 0], "6627e8d5 e169c58d bc57ac4c 9b00dbd8"),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, "408f276d 41c83b0e a20bc7c6 6d5451fd"),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         "d16cfe09 94fdcceb 5001e420 24126ea1"),
    ]
    for ctr, key, expect in kat:
        out = oracle.philox(ctr, key)
        assert " ".join(f"{x:08x}" for x in out) == expect


def test_philox_matches_independent_python_model(oracle):
    def py_philox(c, k):
        c = list(c); k0, k1 = k
        for r in range(10):
            if r:
                k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF
                k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF
            p0, p1 = 0xD2511F53 * c[0], 0xCD9E8D57 * c[2]
            c = [(p1 >> 32) ^ c[1] ^ k0, p1 & 0xFFFFFFFF, (p0 >> 32) ^ c[3] ^ k1, p0 & 0xFFFFFFFF]
        return c
    rng = np.random.default_rng(7)
    for _ in range(200):
        c = [int(x) for x in rng.integers(0, 2**32, 4)]
        k = [int(x) for x in rng.integers(0, 2**32, 2)]
        assert list(oracle.philox(c, k)) == py_philox(c, k)


def test_contract_log_within_one_ulp_of_libm(oracle):
    rng = np.random.default_rng(1)
    xs = np.concatenate([rng.random(20000), rng.random(500) * 1e-12, 1 - rng.random(500) * 1e-9,
                         [2.0 ** -53, 0.5, 0.7071067811865476, 0.70710678118, 1 - 2.0 ** -53, 1.5 * 2.0 ** -53]])
    worst = 0.0
    for x in xs:
        a, b = oracle.contract_log(float(x)), math.log(float(x))
        if b != 0.0:
            worst = max(worst, abs(a - b) / math.ulp(b))
    assert worst <= 1.0


def test_exponential_draws_have_unit_mean(oracle):
    # D1: e = -log(((w >> 12) + 0.5) * 2^-52); over Philox words the sample mean/variance are 1
    es = []
    for it in range(1, 4001):
        o = oracle.philox([it, 0, 5, 0], [1234, 0])
        w = int(o[0]) | (int(o[1]) << 32)
        u = ((w >> 12) + 0.5) * 2.0 ** -52
        assert 0.0 < u < 1.0
        es.append(-oracle.contract_log(u))
    es = np.array(es)
    assert abs(es.mean() - 1.0) < 0.06 and abs(es.var() - 1.0) < 0.15
