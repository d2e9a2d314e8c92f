"""Known-answer tests for the oracle's generators."""
import oracle_py as O


def test_philox4x32_10_known_answers():
    # Random123 kat_vectors (philox4x32 10 rounds)
    assert O.philox([0, 0, 0, 0], [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert O.philox([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert O.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_park_miller_ten_thousandth():
    # Park & Miller 1988: seed 1 -> after 10000 steps the state is 1043618065
    import ctypes as C

    class P(C.Structure):
        _fields_ = [("seed", C.c_int)]
    L = O.lib()
    L.orc_park_uniform.restype = C.c_double
    p = P(1)
    for _ in range(10000):
        L.orc_park_uniform(C.byref(p))
    assert p.seed == 1043618065


def test_injection_identity():
    # 16807^-1 mod (2^31-1) = 1407677000 (used by oracle/ref/ref_driver.cpp to inject uniforms)
    assert (16807 * 1407677000) % O.IM == 1


def test_accept_threshold_monotone():
    t = [O.accept_threshold(0.7, d) for d in range(1, 27)]
    assert all(a >= b for a, b in zip(t, t[1:])) and t[0] > t[5] > 0 and t[0] < 2 ** 32
