"""CPU tests of the oracle against the reference's own pins:
golden integer vectors, doctest known answers, analytic frequency responses (tests/test_flow.rs:18-80,
tolerance 2e-4), allpass |H| = 1 (test_flow.rs:252-283), tick == process within 1e-4
(tests/test_basic.rs:21-47), structural equivalences and pseudorandom-phase divergence
(test_basic.rs:392-406, 520-612)."""
import json
import math
import os

import ctypes as C

import numpy as np
import pytest

from fundsp_b200.graph import ArityError
from fundsp_b200.prelude import *  # noqa: F401,F403
from oracle import OracleUnit, lib, oracle_bank_render

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "int_vectors.json")))
FP = C.POINTER(C.c_float)
L = lib()


# ---------------------------------------------------------------- integer paths: bit exact
def test_golden_integer_paths():
    for x, bits in GOLD["rnd1_bits"]:
        assert L.fo_rnd1(x) == bits * 2.0 ** -53
    for x, h in GOLD["hash1"]:
        assert L.fo_hash1(x) == h
    for s, d, h in GOLD["attohash"]:
        assert L.fo_attohash(s, d) == h
    for x, h in GOLD["hash32x"]:
        assert L.fo_hash32x(x) == h


def test_golden_leaf_hashes():
    f, m = 440.0, 2.0
    graphs = {
        "sine_hz>>lowpass_hz": sine_hz(440.0) >> lowpass_hz(1000.0, 1.0),
        "saw_hz>>lowpass_hz": saw_hz(110.0) >> lowpass_hz(1000.0, 1.0),
        "white>>lowpass_hz": white() >> lowpass_hz(1000.0, 1.0),
        "noise|noise": noise() | noise(),
        "fm": sine_hz(f) * f * m + f >> sine(),
    }
    for k, g in graphs.items():
        assert OracleUnit(g).leaf_hashes() == GOLD["leaf_hashes"][k], k


def test_known_initial_sine_phase():
    # Sine in `sine_hz(440) >> lowpass_hz(1000, 1)` gets hash 0x81b253a0d4def3fc -> phase 0.6899407
    h = OracleUnit(sine_hz(440.0) >> lowpass_hz(1000.0, 1.0)).leaf_hashes()[1]
    assert h == 0x81B253A0D4DEF3FC
    assert np.float32(L.fo_rnd1(h)) == np.float32(0.6899407)


# ---------------------------------------------------------------- doctest known answers
def test_doctest_known_answers():
    assert L.fo_lerpd(0.0, 5.0, 0.5) == 2.5 and L.fo_lerpd(0.0, 5.0, 1.0) == 5.0  # math.rs:184-188
    assert L.fo_delerpd(2.0, 4.0, 3.0) == 0.5  # math.rs:212-215
    assert 1.4125 < L.fo_db_amp(3.0) < 1.4126  # math.rs:285-287
    assert OracleUnit(pass_()).tick([2.0])[0] == 2.0  # audionode.rs:77
    assert OracleUnit(dc(2.0)).tick()[0] == 2.0  # audionode.rs:225
    assert list(OracleUnit(dc((5.0, 6.0))).tick()) == [5.0, 6.0]  # audionode.rs:248
    assert OracleUnit(add(1.0)).tick([1.0])[0] == 2.0  # audionode.rs:266
    assert list(OracleUnit(add((2.0, 3.0))).tick([4.0, 5.0])) == [6.0, 8.0]  # audionode.rs:281
    # net.rs:361-371
    net = L.fo_net_new(1, 1)
    L.fo_net_chain(net, add(1.0).lower(__import__("oracle").OracleBackend()))
    L.fo_net_chain(net, add(2.0).lower(__import__("oracle").OracleBackend()))
    u = OracleUnit(net)
    assert L.fo_net_size(net) == 2 and u.tick([1.0])[0] == 4.0


def test_arity_table():  # tests/test_basic.rs:616-658
    io = lambda g: (g.inputs(), g.outputs())  # noqa: E731
    assert io(pass_() ^ pass_()) == (1, 2)
    assert io(mul(0.5) + mul(0.5)) == (2, 1)
    assert io(sink() | zero()) == (1, 1)
    assert io(sink() | zero() | pass_()) == (2, 2)
    assert io(mul((0.0, 1.0))) == (2, 2)
    assert io(~butterpass() >> ~butterpass() >> butterpass()) == (2, 1)
    assert io(~resonator() >> resonator()) == (3, 1)
    assert io(sine_hz(2.0) * 2.0 * 1.0 + 2.0 >> sine()) == (0, 1)
    assert io((pass_() ^ mul(2.0)) >> sine() + sine()) == (1, 1)
    assert io(sine() & mul(2.0) >> sine()) == (1, 1)
    assert io(feedback(delay(0.5) * 0.5)) == (1, 1)
    assert io(~zero()) == (0, 0)
    assert io(-(-sink()) - 42.0 ^ sink() & -(-(-sink())) * 3.15) == (1, 0)
    with pytest.raises(ArityError):
        pass_() >> (pass_() | pass_())
    for g in (pass_() ^ pass_(), mul(0.5) + mul(0.5), ~resonator() >> resonator(), sine() & mul(2.0) >> sine()):
        u = OracleUnit(g)
        assert (u.inputs(), u.outputs()) == io(g)


# ---------------------------------------------------------------- analytic frequency responses
def svf_response(mode, sr, fc, q, gain, f):
    """Closed forms from src/svf.rs:315-322 ... :720-741 (f64)."""
    g = math.tan(math.pi * fc / sr)
    k = 1.0 / q
    z = np.exp(1j * f * 2 * math.pi / sr)
    den = (z - 1) ** 2 + g * g * (1 + z) ** 2 + g * k * (z * z - 1)
    if mode == LOWPASS:
        return g * g * (1 + z) ** 2 / den
    if mode == HIGHPASS:
        return (z - 1) ** 2 / den
    if mode == BANDPASS:
        return g * (z * z - 1) / den
    if mode == NOTCH:
        return ((z - 1) ** 2 + g * g * (1 + z) ** 2) / den
    if mode == PEAK:
        return -((1 + g + (g - 1) * z) * (-1 + g + z + g * z)) / den
    if mode == ALLPASS:
        return ((z - 1) ** 2 + g * g * (1 + z) ** 2 + g * (k - k * z * z)) / den
    a = math.sqrt(gain)
    if mode == BELL:
        return (g * k * (z * z - 1) + a * (g * (1 + z) * ((a * a - 1) * k / a * (z - 1)) + ((z - 1) ** 2 + g * g * (1 + z) ** 2))) / (
            g * k * (z * z - 1) + a * ((z - 1) ** 2 + g * g * (z + 1) ** 2))
    sa = math.sqrt(a)
    if mode == LOWSHELF:
        return (a * (z - 1) ** 2 + g * g * a * a * (z + 1) ** 2 + sa * g * a * k * (z * z - 1)) / (
            a * (z - 1) ** 2 + g * g * (1 + z) ** 2 + sa * g * k * (z * z - 1))
    return (sa * g * (1 + z) * (-(a - 1) * a * k * (z - 1) + sa * g * (1 - a * a) * (1 + z))
            + a * a * ((z - 1) ** 2 + a * g * g * (1 + z) ** 2 + sa * g * k * (z * z - 1))) / (
        (z - 1) ** 2 + a * g * g * (1 + z) ** 2 + sa * g * k * (z * z - 1))


def biquad_response(c, sr, f):  # src/biquad.rs:119-128
    a1, a2, b0, b1, b2 = c
    z1 = np.exp(-1j * 2 * math.pi * f / sr)
    return (b0 + b1 * z1 + b2 * z1 * z1) / (1 + a1 * z1 + a2 * z1 * z1)


def measure_response(unit, sr=44100.0, length=0x8000):
    """tests/test_flow.rs:25-49: warm up with zeros, feed an impulse, FFT."""
    unit.set_sample_rate(sr)
    x = np.zeros((1, length // 2 + length), np.float32)
    x[0, length // 2] = 1.0
    y = unit.process_many(x.shape[1], x)[0, length // 2:]
    return np.fft.rfft(y.astype(np.float64))


def check_response(unit, analytic, sr=44100.0, length=0x8000):
    spec = measure_response(unit, sr, length)
    f = 10.0
    while f <= 22000.0:
        i = int(round(f * length / sr))
        if i >= len(spec):
            break
        fi = i / length * sr
        x, y = analytic(fi), spec[i]
        tol = 2.0e-4 * max(1.0, abs(x), abs(y))  # test_flow.rs:18-23 is_equal_response
        assert abs(x - y) <= tol, (fi, x, y)
        f += 10.0 if f < 1000.0 else 100.0


SR = 44100.0


@pytest.mark.parametrize("mode,fc,q,gain", [
    (BELL, 500.0, 1.0, 2.0), (LOWSHELF, 2000.0, 10.0, 5.0), (HIGHSHELF, 2000.0, 10.0, 5.0), (PEAK, 5000.0, 1.0, 1.0),
    (ALLPASS, 500.0, 5.0, 1.0), (NOTCH, 1000.0, 1.0, 1.0), (LOWPASS, 50.0, 1.0, 1.0), (HIGHPASS, 5000.0, 1.0, 1.0),
    (BANDPASS, 100.0, 1.0, 1.0)])
def test_svf_responses(mode, fc, q, gain):  # test_flow.rs:85-94
    from fundsp_b200.prelude import _svf_hz
    check_response(OracleUnit(_svf_hz(mode, fc, q, gain)), lambda f: svf_response(mode, SR, fc, q, gain, f))


def test_biquad_family_responses():  # test_flow.rs:103-110,158-166
    c = np.zeros(5, np.float32)
    cp = c.ctypes.data_as(__import__("ctypes").POINTER(__import__("ctypes").c_float))
    L.fo_biquad_coefs(1, SR, 300.0, 20.0, 1.0, cp)
    check_response(OracleUnit(resonator_hz(300.0, 20.0)), lambda f, c=c.copy(): biquad_response(c, SR, f))
    for fc in (200.0, 1000.0):
        L.fo_biquad_coefs(0, SR, fc, 1.0, 1.0, cp)
        check_response(OracleUnit(butterpass_hz(fc)), lambda f, c=c.copy(): biquad_response(c, SR, f))
    check_response(OracleUnit(biquad(0.1, 0.2, 0.3, 0.4, 0.5)), lambda f: biquad_response((0.1, 0.2, 0.3, 0.4, 0.5), SR, f))
    # fir: test_flow.rs:158-160
    w = (0.5, 0.3, 0.2)
    check_response(OracleUnit(fir(w)), lambda f: __import__("builtins").sum(w[2 - i] * np.exp(-1j * 2 * math.pi * f / SR) ** i for i in range(3)))
    # delays: test_flow.rs:98-100
    check_response(OracleUnit(delay(0.0001) >> delay(0.0002)),
                   lambda f: np.exp(-1j * 2 * math.pi * (round(0.0001 * SR) + round(0.0002 * SR)) * f / SR))
    check_response(OracleUnit(pass_() & tick()), lambda f: 1 + np.exp(-1j * 2 * math.pi * f / SR))
    # bus of SVFs: test_flow.rs:94
    check_response(OracleUnit(highpass_hz(500.0, 1.0) & bandpass_hz(500.0, 2.0)),
                   lambda f: svf_response(HIGHPASS, SR, 500.0, 1.0, 1.0, f) + svf_response(BANDPASS, SR, 500.0, 2.0, 1.0, f))


def test_onepole_family_responses():  # test_flow.rs:95,101-105,114-115 against the closed forms of src/filter.rs `route`
    z1 = lambda f: np.exp(-1j * 2 * math.pi * f / SR)
    lowc = lambda fc: math.exp(-2 * math.pi * fc / SR)
    for fc in (1000.0, 10000.0):
        check_response(OracleUnit(lowpole_hz(fc)), lambda f, c=lowc(fc): (1 - c) / (1 - c * z1(f)))
    check_response(OracleUnit(split(2) >> (lowpole_hz(100.0) + lowpole_hz(190.0))),
                   lambda f: (1 - lowc(100.0)) / (1 - lowc(100.0) * z1(f)) + (1 - lowc(190.0)) / (1 - lowc(190.0) * z1(f)))
    hp = lambda fc, f: lowc(fc) * (1 - z1(f)) / (1 - lowc(fc) * z1(f))
    check_response(OracleUnit(highpole_hz(5000.0) & highpole_hz(500.0) & highpole_hz(2000.0)), lambda f: hp(5000.0, f) + hp(500.0, f) + hp(2000.0, f))
    ap = lambda d, f: ((1 - d) / (1 + d) + z1(f)) / (1 + (1 - d) / (1 + d) * z1(f))
    check_response(OracleUnit(allpole_delay(0.5) & allpole_delay(1.3) & allpole_delay(0.1)), lambda f: ap(0.5, f) + ap(1.3, f) + ap(0.1, f))
    dcc = lambda fc: 1 - 2 * math.pi / SR * fc
    check_response(OracleUnit(dcblock()), lambda f: (1 - z1(f)) / (1 - dcc(10.0) * z1(f)))
    check_response(OracleUnit(dcblock_hz(100.0)), lambda f: (1 - z1(f)) / (1 - dcc(100.0) * z1(f)))

    def pink_h(f):
        z = z1(f)
        return (0.0555179 / (1 - 0.99886 * z) + 0.0750759 / (1 - 0.99332 * z) + 0.1538520 / (1 - 0.96900 * z) + 0.3104856 / (1 - 0.86650 * z)
                + 0.5329522 / (1 - 0.55000 * z) - 0.016898 / (1 + 0.7616 * z) + 0.115926 * z + 0.5362) * 0.115830421
    check_response(OracleUnit(pinkpass() * dc(2.0)), lambda f: 2.0 * pink_h(f))
    # tick == process: tests/test_basic.rs:172,198,202,335-336
    L.fo_set_denormal_emulation(0)
    check_wave(pink().seed(2) & noise() | sine_hz(440.0) & -noise())
    check_wave((noise() | dc(440.0)) >> pipei(3, lambda i: ~lowpole()) >> lowpole())
    check_wave((brown().seed(2) | dc(440.0)) >> pipei(4, lambda i: ~peak_q(1.0)) >> bell_q(1.0, 2.0))
    check_wave(noise() >> (butterpass_hz(1000.0) ^ lowpole_hz(100.0)) | noise() >> (allpole_delay(0.5) ^ highpole_hz(500.0)))
    check_wave((noise() | sine_hz(2.0) * 300.0 + 500.0) >> highpole() | (noise() | sine_hz(1.0) * 0.4 + 0.6) >> allpole())
    L.fo_restore_denormals()


def test_composite_graph_responses():  # tests/test_flow.rs:107-155: combinators x filters against composed closed forms
    import ctypes
    z1 = lambda f: np.exp(-1j * 2 * math.pi * f / SR)

    def bq(kind, fc, q=1.0):
        c = np.zeros(5, np.float32)
        L.fo_biquad_coefs(kind, SR, fc, q, 1.0, c.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        return lambda f, c=c.copy(): biquad_response(c, SR, f)
    butter = lambda fc: bq(0, fc)
    reson = lambda fc, q: bq(1, fc, q)
    svf = lambda mode, fc, q, gain=1.0: (lambda f: svf_response(mode, SR, fc, q, gain, f))
    lowp = lambda fc: (lambda f, c=math.exp(-2 * math.pi * fc / SR): (1 - c) / (1 - c * z1(f)))
    dly = lambda t: (lambda f, n=round(t * SR): z1(f) ** n)
    check_response(OracleUnit(butterpass_hz(500.0) & bell_hz(2000.0, 10.0, 5.0)), lambda f: butter(500.0)(f) + svf(BELL, 2000.0, 10.0, 5.0)(f))          # :107
    check_response(OracleUnit(butterpass_hz(6000.0) >> lowpass_hz(500.0, 3.0)), lambda f: butter(6000.0)(f) * svf(LOWPASS, 500.0, 3.0)(f))                 # :108
    check_response(OracleUnit(pass_() * 0.25 & tick() * 0.5 & tick() >> tick() * 0.25), lambda f: 0.25 + 0.5 * z1(f) + 0.25 * z1(f) ** 2)                  # :110
    check_response(OracleUnit(tick() & lowshelf_hz(500.0, 2.0, 0.1)), lambda f: z1(f) + svf(LOWSHELF, 500.0, 2.0, 0.1)(f))                                 # :111
    check_response(OracleUnit((delay(0.001) ^ delay(0.002)) >> reverse(2) >> (delay(0.003) | delay(0.007)) >> join(2)),
                   lambda f: 0.5 * (dly(0.002)(f) * dly(0.003)(f) + dly(0.001)(f) * dly(0.007)(f)))                                                         # :114-116
    check_response(OracleUnit((butterpass_hz(15000.0) ^ allpass_hz(10000.0, 10.0)) >> lowpole_hz(500.0) + pass_()),
                   lambda f: butter(15000.0)(f) * lowp(500.0)(f) + svf(ALLPASS, 10000.0, 10.0)(f))                                                         # :117-119
    check_response(OracleUnit((resonator_hz(12000.0, 500.0) ^ lowpass_hz(3000.0, 0.5)) >> pass_() + highshelf_hz(3000.0, 0.5, 4.0)),
                   lambda f: reson(12000.0, 500.0)(f) + svf(LOWPASS, 3000.0, 0.5)(f) * svf(HIGHSHELF, 3000.0, 0.5, 4.0)(f))                                 # :120-123
    check_response(OracleUnit(split(32) >> multipass(32) >> join(32)), lambda f: 1.0 + 0.0 * f)                                                            # :124
    check_response(OracleUnit(split(8) >> stacki(8, lambda i: resonator_hz(1000.0 + 1000.0 * i, 100.0 + 100.0 * i)) >> join(8)),
                   lambda f: __import__("builtins").sum(reson(1000.0 + 1000.0 * i, 100.0 + 100.0 * i)(f) for i in range(8)) / 8.0)                       # :125-131
    check_response(OracleUnit(pipei(4, lambda i: bell_hz(1000.0 + 1000.0 * i, float(i + 1), db_amp(float(i + 6))))),
                   lambda f: np.prod([svf(BELL, 1000.0 + 1000.0 * i, float(i + 1), float(np.float32(db_amp(float(i + 6)))))(f) for i in range(4)], axis=0))   # :133-139
    check_response(OracleUnit(split(5) >> stacki(5, lambda i: lowpole_hz(1000.0 + 1000.0 + i)) >> join(5)),
                   lambda f: __import__("builtins").sum(lowp(2000.0 + i)(f) for i in range(5)) / 5.0)                                                       # :140-142
    lw, rw = math.cos((0.5 + 1.0) * math.pi / 4), math.sin((0.5 + 1.0) * math.pi / 4)
    check_response(OracleUnit(0.5 * pan(0.0) >> join(2)), lambda f: 0.5 * (math.cos(math.pi / 4) + math.sin(math.pi / 4)) / 2 + 0.0 * f)                 # :153
    check_response(OracleUnit(pan(-1.0) * 0.5 >> multijoin(1, 2)), lambda f: 0.5 * (1.0 + 0.0) / 2 + 0.0 * f)                                               # :155
    check_response(OracleUnit(fir((0.4, 0.3, 0.2, 0.1))), lambda f: 0.1 + 0.2 * z1(f) + 0.3 * z1(f) ** 2 + 0.4 * z1(f) ** 3)                               # :159
    check_response(OracleUnit(morph_hz(1000.0, 1.0, 0.5)), lambda f: (svf(PEAK, 1000.0, 1.0)(f) + 0.5) * 0.5)                                              # :160
    check_response(OracleUnit(morph_hz(2000.0, 2.0, -0.5)), lambda f: (svf(PEAK, 2000.0, 2.0)(f) - 0.5) * 0.5)                                             # :161
    check_response(OracleUnit((pass_() | dc((500.0, 2.0, -1.0))) >> morph()), lambda f: (svf(PEAK, 500.0, 2.0)(f) - 1.0) * 0.5)                            # :163
    check_response(OracleUnit(biquad(0.0, 0.17149, 0.29287, 0.58574, 0.29287)), lambda f: biquad_response((0.0, 0.17149, 0.29287, 0.58574, 0.29287), SR, f))  # :164
    check_response(OracleUnit(biquad(0.033717, 0.171773, 1.059253, -0.035714, 0.181952)), lambda f: biquad_response((0.033717, 0.171773, 1.059253, -0.035714, 0.181952), SR, f))
    from fundsp_b200.net import Net
    net1 = Net(1, 1); net1.chain(lowpole_hz(1500.0))
    check_response(OracleUnit(net1), lowp(1500.0))                                                                                                           # :176-178


def test_biquad_bank_lane_response():  # test_flow.rs:171-177
    import ctypes
    c = np.zeros(5, np.float32)
    L.fo_biquad_coefs(2, SR, 1000.0, 2.0, 1.0, c.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    g = (pass_() | multizero(7)) >> reverse(8) >> biquad_bank().set(6, c.tolist(), address=[(1, 7)]) >> (multisink(7) | pass_())
    check_response(OracleUnit(g), lambda f: biquad_response(c, SR, f))


def test_allpass_magnitude():  # test_flow.rs:252-283
    for g in (allpass_hz(1000.0, 1.0), allpass_hz(100.0, 10.0), allpass_hz(10000.0, 0.5), tick(), delay(0.001), pass_()):
        spec = measure_response(OracleUnit(g))
        mag = np.abs(spec[1:-1])
        assert np.all(np.abs(mag - 1.0) < 1.0e-5 * 20), float(np.abs(mag - 1).max())


# ---------------------------------------------------------------- tick == process (test_basic.rs:21-47)
def check_wave(g, n=441, sr=44100.0):
    a = OracleUnit(g)
    wave = a.render(sr, n / sr)
    a.reset()
    ticks = np.stack([a.tick() for _ in range(n)], axis=1)
    assert wave.shape == ticks.shape
    assert np.abs(wave - ticks).max() <= 1.0e-4, float(np.abs(wave - ticks).max())


def test_tick_equals_process():
    L.fo_set_denormal_emulation(0)
    check_wave(noise().seed(1) * noise() | noise() + noise())
    check_wave(noise() | sine_hz(440.0) & -noise())
    check_wave(dc((110.0, 220.0)) >> multipass(2) >> -stackf(2, lambda f: (f - 0.5) * sine()))
    check_wave(dc((110.0, 220.0, 440.0, 880.0)) >> multipass(4) >> (sink() | -sine().phase(0.0) | sink() | sine()))
    check_wave(dc((880.0, 440.0)) >> pass_() - pass_() >> branchf(2, lambda f: (f - 0.5) * triangle()))
    check_wave(dc((440.0, 880.0)) >> multisplit(2, 5) >> sumi(10, lambda i: saw() * 0.1) | saw_hz(220.0).phase(0.5) * 0.1)
    check_wave(dc((440.0, 880.0)) >> multisplit(2, 3) >> multijoin(2, 3) >> (sine() | sine()))
    check_wave((noise() >> split(16) >> join(16)) | (noise() >> split(11) >> join(11)))
    check_wave((square_hz(110.0).phase(0.25) | dc(440.0)) >> pipei(4, lambda i: ~lowpass_q(1.0)) >> highpass_q(1.0)
               | ((noise() | dc(880.0)) >> ~bandpass_q(1.0) >> notch_q(2.0)))
    check_wave(noise() >> moog_hz(1500.0, 0.8) | noise() >> moog_hz(500.0, 0.4))
    check_wave((noise() | dc((1000.0, 0.5))) >> moog() | (noise() | dc(800.0)) >> moog_q(0.3))
    bq = biquad_bank().set(6, (0.0, 0.0, 0.2, 0.2, 0.2), address=[(1, 0)]).set(6, (0.2, 0.2, 0.1, 0.3, 0.5), address=[(1, 1)])
    check_wave((noise() | noise() | multizero(6)) >> bq >> (pass_() | pass_() | multisink(6)))
    check_wave((noise() | noise()) >> reverb_stereo(10.0, 5.0, 0.5))
    check_wave(dc(1.0) >> adsr_live(0.001, 0.002, 0.5, 0.003) | dc(0.0) >> adsr_live(0.001, 0.002, 0.5, 0.003))
    check_wave(organ_hz(330.0) | hammond_hz(220.0) * 0.5 & soft_saw_hz(55.0))
    check_wave(noise() >> pan(0.3))
    check_wave(noise().seed(1) * noise() | busi(4, lambda i: mls_bits(10 + i)))                        # test_basic.rs:171
    check_wave(dc(440.0) >> ramp() | ramp_hz(-220.0).phase(0.5))                                       # test_basic.rs:235
    check_wave(impulse(2))                                                                              # test_basic.rs:307
    check_wave(poly_saw_hz(440.0) | poly_square_hz(4400.0))                                             # test_basic.rs:308
    check_wave(poly_saw_hz(550.0).phase(0.75) | poly_square_hz(5500.0).phase(0.5))                      # test_basic.rs:309
    check_wave(dc((660.0, 0.1)) >> poly_pulse().phase(0.75) | poly_pulse_hz(6600.0, 0.9).phase(0.9))    # test_basic.rs:311
    lfo2 = (sine_hz(0.7) * 0.4 + 0.5) | (sine_hz(1.3) * 0.3 + 0.35)
    check_wave((noise() | lfo2) >> multitap(2, 0.0, 1.0))                                               # test_basic.rs:347-349
    check_wave((noise() | lfo2) >> multitap_linear(2, 0.0, 1.0))                                        # test_basic.rs:350-352
    check_wave((mls() | dc(880.0)) >> ~butterpass() >> butterpass())                                    # test_basic.rs:199
    check_wave((noise() | dc((440.0, 110.0))) >> resonator())
    check_wave(noise() >> feedback2(delay(0.002) * 0.5, lowpass_hz(2000.0, 1.0)) | (noise() | noise()) >> fdn2(stacki(2, lambda i: delay(0.001 + 0.0005 * i) * 0.4), stacki(2, lambda i: fir3(0.5))))
    L.fo_restore_denormals()


def test_restated_libm_accuracy():
    """The musl/FreeBSD-lineage scalar functions restated in oracle/fo_libm.h (and, identically, in csrc/dsp/libm.cuh) against
    float64 numpy: a wrong constant or branch shows up as errors of many ulp."""
    rng = np.random.default_rng(11)
    n = 400_000

    def ulps(got, ref):
        ref = ref.astype(np.float64)
        e = np.floor(np.log2(np.maximum(np.abs(ref), 1e-300)))
        u = np.exp2(np.maximum(e, -126.0) - 23.0)
        ok = np.isfinite(ref) & (np.abs(ref) < 3.0e38) & np.isfinite(got)
        return float((np.abs(got.astype(np.float64) - ref) / u)[ok].max())

    def run(fn, x, y=None):
        x = np.ascontiguousarray(x, np.float32)
        y = np.ascontiguousarray(x if y is None else y, np.float32)
        out = np.zeros_like(x)
        L.fo_libm_eval(fn, x.ctypes.data_as(FP), y.ctypes.data_as(FP), out.ctypes.data_as(FP), x.size)
        return out

    xs = np.concatenate([rng.uniform(-20.0, 20.0, n), rng.uniform(-1e4, 1e4, n // 4), rng.uniform(-1e-3, 1e-3, n // 4)]).astype(np.float32)
    x64 = xs.astype(np.float64)
    assert ulps(run(0, xs), np.sin(x64)) <= 1.0
    assert ulps(run(1, xs), np.cos(x64)) <= 1.0
    tx = xs[np.abs(np.cos(x64)) > 1e-3]
    assert ulps(run(2, tx), np.tan(tx.astype(np.float64))) <= 1.5
    assert ulps(run(3, xs), np.tanh(x64)) <= 2.5
    ex = rng.uniform(-80.0, 80.0, n).astype(np.float32)
    assert ulps(run(4, ex), np.expm1(ex.astype(np.float64))) <= 1.0
    assert ulps(run(5, ex), np.exp(ex.astype(np.float64))) <= 1.0
    px = np.concatenate([rng.uniform(1e-4, 0.9999, n), rng.uniform(0.0, 50.0, n)]).astype(np.float32)
    py = np.concatenate([np.floor(rng.uniform(1.0, 1200.0, n)), rng.uniform(-8.0, 8.0, n)]).astype(np.float32)
    with np.errstate(over="ignore", under="ignore"):
        ref = np.power(px.astype(np.float64), py.astype(np.float64))
    keep = ref > 1.2e-38                                   # normal results (the subnormal tail rounds on a coarser grid)
    assert ulps(run(6, px[keep], py[keep]), ref[keep]) <= 1.0
    sp = run(6, np.float32([2, -2, -2, 0, np.inf, 0.5, 1, np.nan]), np.float32([10, 3, 2, -1, -1, np.inf, np.nan, 0]))
    assert np.array_equal(sp, np.float32([1024, -8, 4, np.inf, 0, 0, 1, 1]))


def test_dsf_spectrum_known_answer():  # src/oscillator.rs:104-112: sum over i of r**i * sin(f + i*d)
    sr, f0 = 44100.0, 441.0
    for mk, spacing in ((dsf_saw_r, 1), (dsf_square_r, 2)):
        w = OracleUnit(dc(f0) >> mk(0.5).phase(0.0)).render(sr, 1.0)[0]
        X = np.abs(np.fft.rfft(w)) / (len(w) / 2)
        got = [float(X[int(round((1 + i * spacing) * f0))]) for i in range(6)]
        assert np.allclose(got, [0.5 ** i for i in range(6)], atol=2e-4), got   # partial i sits at (1 + i*spacing)*f with amplitude r**i
    check_wave(dc(330.0) >> dsf_saw_r(0.8) | (dc(220.0) | sine_hz(0.5) * 0.3 + 0.5) >> dsf_square())


def test_declick():  # src/dynamics.rs:245-315
    check_wave(noise() >> declick() | dc(1.0) >> declick_s(0.002))
    y = OracleUnit(dc(1.0) >> declick_s(0.005)).render(44100.0, 0.02)[0]
    n = int(round(0.005 * 44100.0))
    assert y[0] == 0.0 and np.all(np.diff(y[: n + 2]) >= -1e-6) and np.all(y[n + 2:] == 1.0) and abs(y[n // 2] - 0.5) < 0.03


def test_chaotic_oscillators():  # src/oscillator.rs:318-438, tests/test_basic.rs (check_wave of lorenz / rossler)
    check_wave(dc(220.0) >> lorenz() | dc(110.0) >> rossler())
    for g in (dc(440.0) >> lorenz(), dc(440.0) >> rossler()):
        y = OracleUnit(g).render(44100.0, 1.0)[0]
        assert np.isfinite(y).all() and 0.2 < np.abs(y).max() < 1.5 and np.abs(y[-20000:]).std() > 0.05   # bounded, keeps moving
    a, b = OracleUnit(dc(440.0) >> lorenz() | dc(440.0) >> lorenz()).render(44100.0, 0.5)
    assert not np.array_equal(a, b)      # initial x comes from the node's own location hash


def test_rez_and_morph():  # tests/test_basic.rs:338-341 (check_wave_filter), src/rez.rs, src/svf.rs:1034-1111
    L.fo_set_denormal_emulation(0)
    check_wave((noise() | noise()) >> ((pass_() | dc((2000.0, 5.0, 0.8))) >> morph() | morph_hz(440.0, 1.0, 0.0)))
    check_wave((noise() | noise()) >> (lowrez_hz(440.0, 0.5) | bandrez_hz(440.0, 0.5)))
    check_wave((noise() | sine_hz(2.0) * 300.0 + 800.0) >> lowrez_q(0.3) | (noise() | dc((1200.0, 0.7))) >> bandrez())
    L.fo_restore_denormals()
    # morph = -1 / 0 / +1 turns the peak SVF (lowpass - highpass) into lowpass / peak / highpass: (peak + m * x) / 2 with x = lp + bp/q.. + hp
    sr = 44100.0
    for m, dc_gain, ny_gain in ((-1.0, 1.0, 0.0), (1.0, 0.0, 1.0)):
        u = OracleUnit(morph_hz(1000.0, 1.0, m))
        spec = measure_response(u, sr, 0x4000)
        assert abs(abs(spec[1]) - dc_gain) < 2e-2 and abs(abs(spec[len(spec) - 2]) - ny_gain) < 2e-2, (m, abs(spec[1]), abs(spec[-2]))
    # rez: the lowpass output passes DC with unit gain (buf1 follows the input), the bandpass output blocks it
    ylo = OracleUnit(lowrez_hz(1000.0, 0.2)).filter(sr, np.ones((1, 4000), np.float32))[0]
    ybp = OracleUnit(bandrez_hz(1000.0, 0.2)).filter(sr, np.ones((1, 4000), np.float32))[0]
    assert abs(ylo[-1] - 1.0) < 1e-3 and abs(ybp[-1]) < 1e-3


def test_follow_filters():  # test_flow.rs:96-97,102 and src/follow.rs
    z1 = lambda f: np.exp(-1j * 2 * math.pi * f / SR)

    def hc(samples):
        r0 = math.log(max(1.0, samples)) - 0.861624594696583
        r1 = 1.0 / (1.0 + math.exp(-r0))
        return 1.0 - min(0.9999999, r1 * 1.13228543863477 - 0.1322853859)
    for t in (0.0002, 0.001, 0.01):
        c = 1.0 - float(np.float32(hc(float(np.float32(t) * np.float32(SR)))))
        check_response(OracleUnit(follow(t)), lambda f, c=c: ((1 - c) / (1 - c * z1(f))) ** 3)
    c = 1.0 - float(np.float32(hc(float(np.float32(0.001) * np.float32(SR)))))
    check_response(OracleUnit(dcblock_hz(100.0) & follow(0.001)), lambda f: (1 - z1(f)) / (1 - (1 - 2 * math.pi / SR * 100.0) * z1(f)) + ((1 - c) / (1 - c * z1(f))) ** 3)
    # halfway response: a step reaches 1/2 after `response_time` (0.5 % accuracy of the approximation, src/follow.rs:18)
    step = np.concatenate([np.zeros(100, np.float32), np.ones(2000, np.float32)])[None, :]   # (the first sample after a reset is copied through)
    y = OracleUnit(follow(0.01)).filter(SR, step)[0]
    assert abs(int(np.argmax(y >= 0.5)) - 100 - 441) <= 4
    # asymmetric: fast attack, slow release
    x = np.concatenate([np.zeros(100, np.float32), np.ones(2000, np.float32), np.zeros(2000, np.float32)])[None, :]
    ya = OracleUnit(afollow(0.001, 0.02)).filter(SR, x)[0]
    assert abs(int(np.argmax(ya >= 0.5)) - 100 - 44) <= 2 and abs(int(np.argmax(ya[2100:] <= 0.5)) - 882) <= 8
    check_wave(noise() >> follow(0.002) | noise().seed(2) >> afollow(0.001, 0.01))


def test_nonlinear_biquads():  # tests/test_basic.rs:218-233 (the lines whose shapes do not need atan), src/biquad.rs:494-920
    import ctypes
    L.fo_set_denormal_emulation(0)
    check_wave(noise() >> dbell_hz(Tanh(1.0), 1000.0, 10.0, 2.0) | noise().seed(2) >> dhighpass_hz(Softsign(1.0), 2000.0, 2.0))     # :218-221
    check_wave(noise() >> dresonator_hz(Tanh(0.5), 1000.0, 10.0) | noise().seed(2) >> dlowpass_hz(Softsign(0.5), 2000.0, 2.0))     # :222-225
    check_wave(noise() >> fbell_hz(Tanh(1.0), 500.0, 50.0, 0.5) | noise().seed(2) >> flowpass_hz(Clip(1.0), 2000.0, 2.0))           # :226-229 (Tanh for Atan)
    check_wave(noise() >> fresonator_hz(Tanh(0.5), 500.0, 50.0) | noise().seed(2) >> fhighpass_hz(Softsign(0.2), 2000.0, 2.0))     # :230-233
    check_wave((noise() | sine_hz(1.0) * 500.0 + 1500.0 | dc(2.0)) >> dlowpass(Tanh(1.0)) | (noise().seed(3) | dc((800.0, 3.0, 2.0))) >> fbell(Softsign(1.0)))
    L.fo_restore_denormals()
    # with a clipper that never engages the structure is a plain transposed-direct-form-II biquad: pins coefficients + recurrence
    for mode, mk_d, mk_f, args in ((2, dlowpass_hz, flowpass_hz, (2000.0, 2.0)), (3, dhighpass_hz, fhighpass_hz, (3000.0, 1.0)), (1, dresonator_hz, fresonator_hz, (1200.0, 5.0))):
        c = np.zeros(5, np.float32)
        L.fo_biquad_coefs(mode, SR, args[0], args[1], 1.0, c.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
        for mk in (mk_d, mk_f):
            check_response(OracleUnit(mk(ClipTo(-100.0, 100.0), *args)), lambda f, c=c.copy(): biquad_response(c, SR, f))
    c = np.zeros(5, np.float32)
    L.fo_biquad_coefs(4, SR, 1000.0, 2.0, 3.0, c.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    check_response(OracleUnit(dbell_hz(ClipTo(-100.0, 100.0), 1000.0, 2.0, 3.0)), lambda f: biquad_response(c, SR, f))


def test_shapers():  # src/shape.rs: Shape::shape (tick) and Shape::simd (block path)
    x = np.float32([[-2.0, -0.75, -0.26, 0.0, 0.1, 0.26, 0.5, 0.75, 1.5, 3.0]])
    f = lambda g: OracleUnit(g).filter(SR, x)[0]     # 10 samples: one SIMD group of 8 + 2 tail samples through `shape`
    assert np.array_equal(f(clip()), np.clip(x[0], -1, 1))
    assert np.array_equal(f(clip_to(-0.5, 0.25)), np.clip(x[0], -0.5, 0.25))
    assert np.allclose(f(shape(Tanh(1.5))), np.tanh(1.5 * x[0].astype(np.float64)), atol=2e-7)
    assert np.allclose(f(shape(Softsign(2.0))), 2 * x[0] / (1 + np.abs(2 * x[0])), atol=1e-7)
    # Crush(2): block path rounds half to even (wide), the tail rounds half away from zero (f32::round)
    y = f(shape(Crush(2.0)))
    assert np.array_equal(y[:8], np.float32([-2.0, -1.0, -0.5, 0.0, 0.0, 0.5, 0.5, 1.0])) and np.array_equal(y[8:], np.float32([1.5, 3.0]))
    ticks = OracleUnit(shape(Crush(2.0)))
    assert [float(ticks.tick([v])[0]) for v in (-0.75, 0.75, 0.25)] == [-1.0, 1.0, 0.5]
    sc = f(shape(SoftCrush(4.0)))
    assert np.all(np.abs(sc - x[0]) <= 0.125 + 1e-6) and np.all(np.diff(sc) >= 0)     # monotone staircase with soft steps
    L.fo_set_denormal_emulation(0)
    check_wave(noise() >> shape(Tanh(2.0)) | noise().seed(3) >> shape(Softsign(3.0)) | noise().seed(4) >> shape(SoftCrush(7.0)) | noise().seed(5) >> clip_to(-0.3, 0.6))
    L.fo_restore_denormals()


def test_convolver():  # tests/test_basic.rs:698-711 (the reference's own pin, tolerance 1e-4) and :329-330 (check_wave)
    u = OracleUnit(convolve([1.00, 0.75, 0.50, 0.25]))
    got = [float(u.tick([x])[0]) for x in (0.0, 1.0, 0.0, 0.0, 0.0)]
    assert np.allclose(got, [0.00, 1.00, 0.75, 0.50, 0.25], atol=1e-4)
    check_wave(noise() >> convolve([1.0, 0.9, 0.8]) | noise().seed(5) >> convolve([0.5, 0.4, 0.3]))
    rng = np.random.default_rng(3)
    h = rng.uniform(-1, 1, 300).astype(np.float32)
    x = rng.uniform(-1, 1, (1, 2000)).astype(np.float32)
    y = OracleUnit(convolve(h)).filter(44100.0, x)[0]
    ref = np.convolve(x[0].astype(np.float64), h.astype(np.float64))[:2000]
    assert np.abs(y - ref).max() <= 1e-6 * np.abs(ref).max()


def test_feedback_unit():  # tests/test_basic.rs:243-253 (FeedbackUnit inside check_wave), src/feedback.rs:316-481
    mk = lambda: (noise() >> feedback_unit(0.01, 0.5 * lowpass_hz(1000.0, 1.0))) | (noise() >> feedback_unit(0.001, 0.5 * highpass_hz(1000.0, 1.0)))
    check_wave(mk())
    # with a pure gain inside, y[t] = x[t] * g + g * y[t - d]: an impulse comes back every d samples scaled by g
    u = OracleUnit(impulse(1) >> feedback_unit(100.0 / 44100.0, 0.5 * pass_()))
    y = u.process_many(450)[0]
    assert np.array_equal(np.nonzero(y)[0], [0, 100, 200, 300, 400]) and np.allclose(y[[0, 100, 200, 300, 400]], [0.5, 0.25, 0.125, 0.0625, 0.03125])
    # the minimum delay is one sample, also for delay = 0
    y0 = OracleUnit(impulse(1) >> feedback_unit(0.0, 0.5 * pass_())).process_many(5)[0]
    assert np.allclose(y0, [0.5, 0.25, 0.125, 0.0625, 0.03125])


def test_reverb3_and_var():  # src/reverb.rs:139-279, src/prelude.rs:1856 doc example, src/shared.rs:84-131
    mk = lambda: (noise().seed(1) | noise().seed(2)) >> (multipass(2) & 0.25 * reverb3_stereo(2.0, 0.5, lowpass_hz(8000.0, 0.7)))
    # tick == process on two FRESH units: Reverb::reset leaves the four pre-delay allpasses alone (src/reverb.rs:215-228), so the
    # reset-then-tick pattern of check_wave would compare different states
    blocks, ticks = OracleUnit(mk()), OracleUnit(mk())
    wave = blocks.process_many(700)
    tk = np.stack([ticks.tick() for _ in range(700)], axis=1)
    assert np.abs(wave - tk).max() <= 1.0e-4 and np.abs(wave).max() > 0.1
    # with an identity loop filter the structure is allpasses + delays + one gain a < 1: stable, and silent until the first tap
    r = OracleUnit(impulse(2) >> reverb3_stereo(1.0, 0.5, pass_())).render(44100.0, 2.0)
    first = int(np.argmax(np.abs(r).max(axis=0) > 0))
    assert first == 0 and np.isfinite(r).all()     # every Schroeder allpass passes eta * x straight through
    en = lambda a, b: 10.0 * np.log10(float((r[:, a:b] ** 2).mean()) + 1e-30)
    assert en(2000, 6000) > en(42100, 46100) + 15.0 > en(84000, 88200) + 30.0     # monotone decay of the tail
    # longer time = slower decay (a = db_amp(-60) ** (0.035 / time))
    r4 = OracleUnit(impulse(2) >> reverb3_stereo(4.0, 0.5, pass_())).render(44100.0, 2.0)
    assert (r4[:, 80000:] ** 2).mean() > 10.0 * (r[:, 80000:] ** 2).mean()
    v = OracleUnit(var(0.25) * dc(2.0))
    assert np.array_equal(v.process_many(70)[0], np.full(70, 0.5, np.float32))
    v.L.fo_set(v.h, 4, (C.c_float * 1)(0.75), 1, 0, (C.c_int64 * 2)(1, 0), 1)   # Setting::value(0.75).left()
    assert np.array_equal(v.process_many(70)[0], np.full(70, 1.5, np.float32))


def test_more_reference_check_wave_lines():  # tests/test_basic.rs:170,187-190,199-209 restated with the wider node set
    L.fo_set_denormal_emulation(0)
    check_wave(noise() >> declick() | noise() + noise())                                                                 # :170
    check_wave(dc((110.0, 220.0)) >> declick_s(0.1) + pass_() >> (saw() ^ dsf_square_r(0.9)))                             # :187
    check_wave(dc((20.0, 40.0)) >> reverse(2) >> pass_() * pass_() >> (dsf_saw_r(0.999) ^ square() * 0.1))                # :188-190
    check_wave((brown().seed(2) | dc(440.0)) >> pipei(4, lambda i: ~peak_q(1.0)) >> bell_q(1.0, 2.0)
               | ((mls() | dc(880.0)) >> ~lowshelf_q(1.0, 0.5) >> highshelf_q(2.0, 2.0)))                                # :201-204
    check_wave((square_hz(110.0).phase(0.25) | dc(440.0)) >> pipei(4, lambda i: ~lowpass_q(1.0)) >> highpass_q(1.0)
               | ((mls() | dc(880.0)) >> ~bandpass_q(1.0) >> notch_q(2.0)))                                             # :205-210
    L.fo_restore_denormals()


def test_pulse_rotate_mixer_reverb4():  # src/wavetable.rs:361-491, src/pan.rs:95-160, src/prelude.rs:1873-1946,2876
    L.fo_set_denormal_emulation(0)
    # tests/test_basic.rs:237: tick == process
    check_wave(dc((110.0, 0.5)) >> pulse() * 0.2 >> delay(0.1))           # (check_wave_big: 441 samples)
    check_wave(dc((110.0, 0.5)) >> pulse() * 0.2)
    check_wave((sine_hz(3.0) * 30.0 + 220.0 | sine_hz(0.7) * 0.4 + 0.5) >> pulse())
    assert outputs_diverge(dc((220.0, 0.3, 220.0, 0.3)) >> (pulse() | pulse()))   # test_basic.rs:606: two pulses draw different initial phases
    # a pulse is the difference of two band-limited saws `width` apart: two levels 2A apart (A = the saws' slope
    # amplitude), the duty cycle is the width, zero mean
    sr, f = 44100.0, 110.25                                       # 400 samples per cycle
    saw = OracleUnit(saw_hz(f).phase(0.0)).render(sr, 0.1)[0]
    A = 2.0 * abs(float(saw[100 + 400]))                          # slope amplitude of the table (its peak 1.0 is the Gibbs overshoot)
    assert 0.8 < A < 0.9
    for w in (0.5, 0.25, 0.8):
        y = OracleUnit(dc((f, w)) >> pulse()).render(sr, 1.0)[0]
        a, b = np.median(y[y > np.median(y) + 0.5] if (y > np.median(y) + 0.5).any() else y), np.median(y[y < np.median(y) - 0.5] if (y < np.median(y) - 0.5).any() else y)
        levels = sorted([float(a), float(b)])
        assert abs((levels[1] - levels[0]) - 2.0 * A) < 0.05 and abs(y.mean()) < 0.01, (w, levels, y.mean())
        frac_hi = float((y > 0.5 * (levels[0] + levels[1])).mean())
        assert min(abs(frac_hi - w), abs(frac_hi - (1.0 - w))) < 0.01, (w, frac_hi)
        assert abs(levels[1] - 2.0 * A * (1.0 - frac_hi)) < 0.05 and abs(levels[0] + 2.0 * A * frac_hi) < 0.05   # zero mean fixes both levels
    # PhaseSynth driven by a ramp reproduces the free-running oscillator of the same table (past the first, Nyquist-band, sample)
    a = OracleUnit(ramp_hz(441.0).phase(0.0) >> phase_synth(TRIANGLE)).render(sr, 2000 / sr)[0]
    b = OracleUnit(triangle_hz(441.0).phase(0.0)).render(sr, 2000 / sr)[0]
    assert min(np.abs(a[2:-2] - b[k:k + len(a) - 4]).max() for k in (1, 2, 3)) < 2e-3
    # rotate (test_flow.rs:168-169): [[c g, -s g], [s g, c g]] with libm cos / sin in f32
    x = np.random.default_rng(3).uniform(-1, 1, (2, 500)).astype(np.float32)
    for ang, g in ((0.5, 1.0), (-0.1, 0.5)):
        c, sn, g32 = np.float32(math.cos(np.float32(ang))), np.float32(math.sin(np.float32(ang))), np.float32(g)
        y = OracleUnit(rotate(ang, g)).filter(sr, x)
        want = np.stack([x[0] * (c * g32) + x[1] * (-sn * g32), x[0] * (sn * g32) + x[1] * (c * g32)])
        assert np.abs(y - want).max() < 1e-6
        u = OracleUnit(rotate(ang, g))
        assert np.array_equal(y[:, :50], np.stack([u.tick(x[:, i]) for i in range(50)], axis=1))
    y = OracleUnit(mixer([[0.5, -0.25, 1.0], [0.0, 2.0, 0.0]])).filter(sr, np.stack([x[0], x[1], x[0]]))
    assert np.array_equal(y[1], x[1] * np.float32(2.0)) and np.abs(y[0] - (1.5 * x[0] - 0.25 * x[1])).max() < 1e-6
    # reverb4_stereo: silent until the shortest path (two lines of about 31.5 ms x 1.5 at the 15 m floor) has passed, then an exponential tail
    imp = np.zeros((2, int(sr * 4.6)), np.float32); imp[:, 0] = 1.0
    y = OracleUnit(reverb4_stereo(10.0, 1.0)).filter(sr, imp)
    first = int((0.031615064 + 0.031507637) * 1.5 * sr)          # the shortest line of each of the two FDNs in series
    assert np.abs(y[:, :first - 3]).max() == 0.0 and np.abs(y[:, first - 3:first + 4]).max() > 0.0
    # every trip round a line is scaled by a = 10**(-3 * 0.03 / time) (the loop FIR's gain at DC): with lines of 1.5 x 51.5 ms on
    # average that is 60 * 0.03 / 0.0773 = 23 dB per second for each FDN; the cascade's t**2 envelope takes ~2.5 dB/s off between 3 and 4 s
    e = (y.astype(np.float64) ** 2).sum(0)
    mean_line = 1.5 * float(np.mean(REVERB4_DELAYS))
    rate = 10.0 * np.log10(e[int(3.0 * sr):int(3.5 * sr)].sum() / e[int(4.0 * sr):int(4.5 * sr)].sum())
    assert abs(rate - (60.0 * 0.03 / mean_line - 2.5)) < 4.0, (rate, 60.0 * 0.03 / mean_line)
    assert np.array_equal(y[:, :3000], OracleUnit(reverb4_stereo(15.0, 1.0)).filter(sr, imp[:, :3000]))   # rooms below 15 m clamp
    check_wave((noise() | noise().seed(3)) >> reverb4_stereo(20.0, 2.0), n=2205)
    L.fo_restore_denormals()


def test_meter_playwave_resample():  # src/dynamics.rs:316-437, src/wave.rs:739-797, src/resample.rs:210-300
    sr = 44100.0
    x = np.random.default_rng(5).uniform(-1, 1, (1, 3000)).astype(np.float32)
    # Meter::Sample passes the signal through (test_dynamics.rs:53-60 compares it with the monitored value)
    assert np.array_equal(OracleUnit(meter(Meter.Sample)).filter(sr, x), x)
    # Meter::Peak: max(state * smoothing, |x|) with smoothing = 0.5 ** (1 / (timescale * sr)): a unit impulse halves in `timescale` seconds
    imp = np.zeros((1, 9000), np.float32); imp[0, 10] = -1.0
    y = OracleUnit(meter(Meter.Peak(0.1))).filter(sr, imp)[0]
    assert y[9] == 0.0 and y[10] == 1.0 and abs(y[10 + 4410] - 0.5) < 1e-3 and np.all(np.diff(y[10:]) < 0)
    yp = OracleUnit(meter(Meter.Peak(0.01))).filter(sr, x)[0]
    assert np.all(yp >= np.abs(x[0])) and np.all(yp[1:] <= np.maximum(yp[:-1], np.abs(x[0, 1:])))
    # Meter::Rms of a full-scale sine settles at 1 / sqrt(2)
    yr = OracleUnit(sine_hz(1000.0) >> meter(Meter.Rms(0.01))).render(sr, 0.5)[0]
    assert abs(yr[-2000:].mean() - 2.0 ** -0.5) < 2e-3
    # WavePlayer: exact samples, silence after the end, loop jumps, play regions
    w = np.random.default_rng(6).uniform(-1, 1, (2, 300)).astype(np.float32)
    y = OracleUnit(playwave(w, 1)).render(sr, 500 / sr)[0]
    assert np.array_equal(y[:300], w[1]) and not y[300:].any()
    y = OracleUnit(playwave(w, 0, 100)).render(sr, 1000 / sr)[0]
    assert np.array_equal(y, np.concatenate([w[0], np.tile(w[0, 100:], 4)])[:1000])
    y = OracleUnit(playwave_at(w, 0, 50, 120, 60)).render(sr, 400 / sr)[0]
    assert np.array_equal(y, np.concatenate([w[0, 50:120], np.tile(w[0, 60:120], 6)])[:400])
    check_wave(playwave(w, 0, 0) | playwave_at(w, 1, 10, 40))
    # Resample at speed 1: the read head starts one sample in and moves before it reads, so output n = inner sample n + 2 exactly
    # (Catmull-Rom passes through its knots); at speed 1/2 the knots interleave with the cubic's midpoints (-a + 9b + 9c - d) / 16
    src = OracleUnit(playwave(w, 0)).render(sr, 300 / sr)[0]
    y = OracleUnit(dc(1.0) >> resample(playwave(w, 0))).render(sr, 200 / sr)[0]
    assert np.array_equal(y, src[2:202])
    y = OracleUnit(dc(0.5) >> resample(playwave(w, 0))).render(sr, 200 / sr)[0]
    assert np.array_equal(y[1::2][:90], src[2:92])
    mid = (-src[0:90] + 9.0 * src[1:91] + 9.0 * src[2:92] - src[3:93]) / 16.0
    assert np.abs(y[0::2][:90] - mid).max() < 1e-6
    # speed 2 reads every other sample; a negative speed is clamped to 0 and holds the value
    y = OracleUnit(dc(2.0) >> resample(playwave(w, 0))).render(sr, 100 / sr)[0]
    assert np.array_equal(y, src[3:203:2])
    y = OracleUnit(dc(-1.0) >> resample(playwave(w, 0))).render(sr, 50 / sr)[0]
    assert np.all(y == src[1])
    # pitched sine: resampling a 440 Hz sine at speed 1.5 gives 660 Hz
    y = OracleUnit(dc(1.5) >> resample(sine_hz(440.0))).render(sr, 1.0)[0]
    spec = np.abs(np.fft.rfft(y * np.hanning(len(y))))
    assert abs(int(np.argmax(spec)) - 660) <= 1
    check_wave((sine_hz(2.0) * 0.5 + 1.0) >> resample(playwave(w, 0, 0)) | dc(0.37) >> resample(noise()))
    assert outputs_diverge(dc(1.0) >> resample(noise()) | dc(1.0) >> resample(noise()))


def test_limiter():  # tests/test_dynamics.rs:30-49 restated (seeded sizes instead of funutd's Rnd), src/dynamics.rs:56-243
    rng = np.random.default_rng(11)
    sr = 48000.0
    for samples in [2, 3, 17, 100] + [int(round(2.0 * (20000.0 / 2.0) ** rng.uniform())) for _ in range(6)]:
        u = OracleUnit(limiter(samples / sr, samples / sr))
        u.set_sample_rate(sr)
        edge = np.float32(math.exp((100.0 / 20.0) * math.log(10.0)))                     # a +100 dB edge: the hardest case
        x = np.concatenate([np.zeros(samples, np.float32), np.full(samples + 1, edge, np.float32)])[None, :]
        y = u.process_many(x.shape[1], x)[0]
        assert np.all(y[samples:2 * samples] <= 1.0), samples
        assert 0.90 <= y[2 * samples] <= 1.00, (samples, y[2 * samples])                   # headroom, but sufficient range
    # below the threshold the limiter is a pure delay of round(sr * attack) samples with unit gain ... / 1.0 (the follower sits at 1)
    x = np.random.default_rng(12).uniform(-0.5, 0.5, (1, 2000)).astype(np.float32)
    y = OracleUnit(limiter(0.005, 0.05)).filter(44100.0, x)[0]
    L = round(44100.0 * 0.005)
    assert not y[:L].any() and np.array_equal(y[L:], x[0, :-L])
    # stereo: both channels share the gain computed from the louder one
    a = np.random.default_rng(13).uniform(-4, 4, (1, 4000)).astype(np.float32)
    y2 = OracleUnit(limiter_stereo(0.002, 0.02)).filter(44100.0, np.concatenate([a, 0.25 * a]))
    y1 = OracleUnit(limiter(0.002, 0.02)).filter(44100.0, a)
    assert np.array_equal(y2[0], y1[0]) and np.abs(y2[1] - np.float32(0.25) * y1[0]).max() < 1e-6 and np.abs(y1).max() <= 1.0
    # tick == process on fresh units. (After `reset()` the reference's limiter is NOT in its constructed state: Limiter::reset
    # re-derives the follower's coefficients but neither clears its three poles nor re-arms its first-sample coefficient of 1
    # (src/dynamics.rs:181-195, src/follow.rs:177-190), so the reference's own check_wave would not hold for it either.)
    g = lambda: noise() * 3.0 >> limiter(0.002, 0.01) | (noise() | sine_hz(300.0) * 4.0) >> limiter_stereo(0.0005, 0.003) >> join(2)
    wave = OracleUnit(g()).render(44100.0, 441 / 44100.0)
    t = OracleUnit(g()); t.set_sample_rate(44100.0)
    assert np.abs(wave - np.stack([t.tick() for _ in range(441)], axis=1)).max() <= 1e-4
    u = OracleUnit(g()); first = u.render(44100.0, 0.01); u.reset()
    assert np.abs(first - np.stack([u.tick() for _ in range(441)], axis=1)).max() > 1e-2


def _seq_four():   # tests/test_basic.rs:255-273
    from fundsp_b200.sequencer import Sequencer, Fade, ReplayMode
    q = Sequencer(0, 2, ReplayMode.All)
    q.push(0.1, 0.2, Fade.Smooth, 0.01, 0.0, noise() | sine_hz(220.0))
    q.push(0.3, 0.4, Fade.Smooth, 0.09, 0.08, sine_hz(110.0) | noise())
    q.push(0.25, 0.5, Fade.Power, 0.0, 0.01, mls() | noise())
    q.push(0.6, 0.7, Fade.Power, 0.02, 0.03, noise() | mls())
    return q


def test_sequencer():  # src/sequencer.rs; tests/test_basic.rs:192-193,255-274,713-765; src/sequencer.rs:918-935 (module tests)
    from fundsp_b200.sequencer import Sequencer, Fade, ReplayMode
    L.fo_set_denormal_emulation(0)
    sr = 44100.0
    # an empty sequencer inside a graph (test_basic.rs:192-193) and the four-event sequence (:255-274): tick == process
    check_wave((noise() | noise()) >> Sequencer(2, 2, ReplayMode.None_).node())
    check_wave(_seq_four().node())
    u = OracleUnit(_seq_four().node())
    wave = u.render(sr, 0.75)
    u.reset()                                                      # ReplayMode::All: reset replays every event
    ticks = np.stack([u.tick() for _ in range(wave.shape[1])], axis=1)
    assert np.abs(wave - ticks).max() <= 1e-4 and np.abs(wave).max() > 0.5
    # silence outside the events, sound inside; event 0 starts at sample 4410 exactly
    assert not wave[:, :4410].any() and wave[1, 4411] != 0.0 and not wave[:, int(0.2 * sr) + 1:int(0.25 * sr) - 1].any() and not wave[:, int(0.7 * sr) + 1:].any()
    # test_sequencer_passthrough (:713-728): events with inputs, tick path
    q = Sequencer(1, 1, ReplayMode.None_)
    q.push(0.0, 1.0, Fade.Smooth, 0.0, 0.0, pass_())
    q.push(1.0 / 44100.0, 1.0, Fade.Smooth, 0.0, 0.0, mul(2.0))
    u = OracleUnit(q.node())
    assert [float(u.tick([x])[0]) for x in (1.0, 2.0, 0.5)] == [1.0, 6.0, 1.5]
    # test_sequencer_loop (:730-765)
    q = Sequencer(0, 1, ReplayMode.Loop(79.0 / 44100.0))
    q.push(12.0 / 44100.0, 89.0 / 44100.0, Fade.Smooth, 0.0, 0.0, dc(1.0))
    u = OracleUnit(q.node())
    got = [float(u.tick()[0]) for _ in range(12 + 77 + 2 + 77 + 2)]
    assert got == [0.0] * 12 + [1.0] * 77 + [0.0] * 2 + [1.0] * 77 + [0.0] * 2
    # reset_replays_events (src/sequencer.rs:918-935)
    q = Sequencer(0, 1, ReplayMode.All); q.push(0.0, 1.0, Fade.Smooth, 0.0, 0.0, sine_hz(440.0))
    u = OracleUnit(q.node()); first = u.tick(); u.reset()
    assert np.array_equal(first, u.tick())
    # known answers on the block path: dc events make the envelope itself visible.
    q = Sequencer(0, 1, ReplayMode.None_)
    q.push(100.0 / sr, 1100.0 / sr, Fade.Smooth, 200.0 / sr, 400.0 / sr, dc(1.0))
    y = OracleUnit(q.node()).render(sr, 1300 / sr)[0]
    assert not y[:100].any() and not y[1100:].any() and np.all(y[300:700] == 1.0)
    s5 = lambda x: ((x * 6.0 - 15.0) * x + 10.0) * x ** 3
    assert np.abs(y[100:300] - s5(np.arange(200) / 200.0)).max() < 1e-4                       # fade in: smooth5 over 200 samples
    assert np.abs(y[700:1100] - s5(1.0 - np.arange(400) / 400.0)).max() < 1e-4               # fade out
    q = Sequencer(0, 1, ReplayMode.None_)
    q.push(37.0 / sr, 937.0 / sr, Fade.Power, 300.0 / sr, 300.0 / sr, dc(1.0))
    y = OracleUnit(q.node()).render(sr, 1000 / sr)[0]
    ph = np.arange(300) / 300.0
    assert np.abs(y[37:337] - np.sin(ph * np.pi / 2)).max() < 2e-3 and np.abs(y[637:937] - np.cos(ph * np.pi / 2)).max() < 2e-3   # Bhaskara sine
    # equal-power crossfade of two dc events: squares sum to one across the overlap
    q = Sequencer(0, 2, ReplayMode.None_)
    q.push(0.0, 1000.0 / sr, Fade.Power, 0.0, 500.0 / sr, dc((1.0, 0.0)))
    q.push(500.0 / sr, 1500.0 / sr, Fade.Power, 500.0 / sr, 0.0, dc((0.0, 1.0)))
    y = OracleUnit(q.node()).render(sr, 1500 / sr)
    assert np.abs(y[0, 500:1000] ** 2 + y[1, 500:1000] ** 2 - 1.0).max() < 5e-3
    # the sum of overlapping events equals the events rendered alone (same seeds: the same graph pushed at the same time)
    mk = lambda: noise().seed(5) >> lowpass_hz(1200.0, 1.0)
    q1 = Sequencer(0, 1, ReplayMode.None_); q1.push(0.01, 0.03, Fade.Smooth, 0.002, 0.004, mk())
    q2 = Sequencer(0, 1, ReplayMode.None_); q2.push(0.02, 0.05, Fade.Smooth, 0.001, 0.001, sine_hz(300.0).phase(0.0))
    q3 = Sequencer(0, 1, ReplayMode.None_); q3.push(0.01, 0.03, Fade.Smooth, 0.002, 0.004, mk()); q3.push(0.02, 0.05, Fade.Smooth, 0.001, 0.001, sine_hz(300.0).phase(0.0))
    a, b, c = (OracleUnit(q.node()).render(sr, 0.06) for q in (q1, q2, q3))
    assert np.array_equal(a + b, c)
    # edit: shortening an event before it starts, and while it plays (the fade-out moves with the new end)
    q = Sequencer(0, 1, ReplayMode.None_)
    e = q.push(100.0 / sr, 2000.0 / sr, Fade.Smooth, 0.0, 0.0, dc(1.0))
    q.edit(e, 600.0 / sr, 100.0 / sr)
    y = OracleUnit(q.node()).render(sr, 800 / sr)[0]
    assert np.all(y[100:500] == 1.0) and not y[600:].any() and np.abs(y[500:600] - s5(1.0 - np.arange(100) / 100.0)).max() < 1e-4
    # a push that lands in the past of a running sequencer starts at once (:347-353)
    h = L.fo_sequencer(0, 1, 1, 0.0); u = OracleUnit(h)
    u.process_many(128)
    L.fo_sequencer_push(h, 0.0, 1.0, 1, 0.0, 0.0, OracleUnit(dc(0.5)).take())
    assert np.all(u.process_many(64)[0] == 0.5) and abs(L.fo_sequencer_time(h) - 192 / 44100.0) < 1e-12
    L.fo_restore_denormals()


def test_envelope_lfo():  # src/envelope.rs:14-183; tests/test_basic.rs:173-176,238,645
    xerp = lambda a, b, t: math.exp(math.log(a) * (1.0 - t) + math.log(b) * t)
    c01 = lambda t: min(1.0, max(0.0, t))
    check_wave(lfo(lambda t: xerp(110.0, 220.0, c01(t))) >> sine() | (envelope(lambda t: xerp(220.0, 440.0, c01(t))) >> pass_() >> sine()) & mls())   # :173-176
    check_wave(envelope(lambda t: math.exp(-t * 10.0)))                                                                  # :238
    g = envelope(lambda t: math.exp(-t)) * noise()
    assert (g.inputs(), g.outputs()) == (0, 1)                                                                           # :645
    sr = 44100.0
    # a linear closure is reproduced by linear interpolation: y[n] = n / sr to rounding, on both paths, f32 and f64 time
    for t64 in (False, True):
        y = OracleUnit(envelope(lambda t: t, time64=t64)).render(sr, 0.5)[0]
        assert np.abs(y - np.arange(len(y)) / sr).max() < 2e-6
    # the sample points are 0.75 .. 1.25 intervals apart: the knots of a piecewise-linear rendering of t**2 show them
    y = OracleUnit(envelope(lambda t: t * t, interval=0.01)).render(sr, 1.0)[0].astype(np.float64)
    dd = np.abs(np.diff(y, 2)) * sr * sr
    knots = np.flatnonzero(dd > 100.0)                                    # slope changes by 2 * interval * (gap) at a knot
    knots = knots[np.insert(np.diff(knots) > 3, 0, True)]
    gaps = np.diff(knots) / sr
    assert len(knots) > 80 and gaps.min() >= 0.0075 - 2 / sr and gaps.max() <= 0.0125 + 2 / sr and gaps.std() > 0.0005
    # two envelopes in one graph draw different jitter (their hashes differ); the same graph built twice is identical
    two = lambda: envelope(lambda t: t * t, interval=0.01) | envelope(lambda t: t * t, interval=0.01)
    a, b = OracleUnit(two()).render(sr, 0.2), OracleUnit(two()).render(sr, 0.2)
    assert np.array_equal(a, b) and not np.array_equal(a[0], a[1]) and np.abs(a[0] - a[1]).max() < 1e-4
    # an lfo drives a parameter: vibrato depth shows up as the spread of instantaneous frequency
    y = OracleUnit(lfo(lambda t: 440.0 + 40.0 * math.sin(2.0 * math.pi * 5.0 * t)) >> sine()).render(sr, 1.0)[0]
    zc = np.flatnonzero((y[:-1] < 0) & (y[1:] >= 0))
    f = sr / np.diff(zc)
    assert 395.0 < f.min() < 410.0 and 470.0 < f.max() < 485.0


def test_monitor_and_unit_are_transparent():  # src/dynamics.rs:441-520 (test_flow.rs:159), src/audiounit.rs:430-484
    x = np.random.default_rng(8).uniform(-1, 1, (1, 500)).astype(np.float32)
    assert np.array_equal(OracleUnit(monitor()).filter(44100.0, x), x)
    assert np.array_equal(OracleUnit(unit(monitor() >> mul(2.0))).filter(44100.0, x), x * np.float32(2.0))
    # a Monitor is not a Pass to the phase hashes of the graph around it (ID 56 vs 48; a constructor's probe ping sees every node)
    assert OracleUnit(monitor() >> sine()).leaf_hashes() != OracleUnit(pass_() >> sine()).leaf_hashes()


def test_oversample():  # src/oversample.rs: 2x oversampling between 43-tap minimum-phase halfbands
    sr = 44100.0
    # transparent in the pass band: DC settles at 1, a 1 kHz and a 15 kHz sine keep their amplitude
    y = OracleUnit(dc(1.0) >> oversample(pass_())).render(sr, 0.02)[0]
    assert abs(y[-1] - 1.0) < 2e-3 and np.abs(y[200:] - 1.0).max() < 2e-3
    for f, tol in ((1000.0, 0.01), (15000.0, 0.05)):
        y = OracleUnit(sine_hz(f) >> oversample(pass_())).render(sr, 0.1)[0]
        assert abs(np.abs(y[1000:]).max() - 1.0) < tol, (f, np.abs(y[1000:]).max())
    # what it is for: the third harmonic of a hard-driven 8.5 kHz sine (25.5 kHz) folds to 18.6 kHz at 1x and is filtered out at 2x
    n = 1 << 15
    def level(g, f):
        y = OracleUnit(g).render(sr, n / sr)[0].astype(np.float64)
        sp = np.abs(np.fft.rfft(y * np.hanning(n)))
        k = int(round(f * n / sr))
        return 20.0 * np.log10(sp[k - 3:k + 4].max() / sp.max())
    drive = lambda: sine_hz(8500.0) * 4.0
    plain, over = level(drive() >> shape(Tanh(1.0)), 18600.0), level(drive() >> oversample(shape(Tanh(1.0))), 18600.0)
    assert plain > -25.0 and over < plain - 25.0, (plain, over)
    # the inner node runs at twice the rate: a one-pole at the same cutoff behaves the same inside and outside
    a = OracleUnit(noise().seed(3) >> oversample(lowpole_hz(500.0))).render(sr, 0.2)[0]
    b = OracleUnit(noise().seed(3) >> lowpole_hz(500.0)).render(sr, 0.2)[0]
    assert abs(a[2000:].std() / b[2000:].std() - 1.0) < 0.05
    # tick == process over whole even blocks (an odd block leaves its last sample unwritten in the reference: see the oracle header)
    g = lambda: (noise().seed(1) | noise().seed(2)) >> oversample(lowpass_hz(3000.0, 1.0) | shape(Tanh(2.0)))
    u = OracleUnit(g()); blocks = u.process_many(64 * 7)
    t = OracleUnit(g()); ticks = np.stack([t.tick() for _ in range(64 * 7)], axis=1)
    assert np.abs(blocks - ticks).max() <= 1e-4 and np.abs(blocks).max() > 0.1


def test_flanger_phaser():  # src/prelude.rs:2719-2753: compositions of lfo, tap, feedback2 / feedback, allpole
    import math as m
    check_wave(noise() >> flanger(0.5, 0.005, 0.010, lambda t: 0.0075 + 0.0025 * m.sin(2.0 * t)) | noise() >> phaser(0.5, lambda t: 0.5 + 0.5 * m.sin(3.0 * t)))
    sr = 44100.0
    # a flanger with a constant delay d and no feedback is a comb: dry + x[n - d], notches at odd multiples of 1 / (2 d)
    d = 100.0 / sr
    x = np.random.default_rng(4).uniform(-1, 1, (1, 1 << 15)).astype(np.float32)
    y = OracleUnit(flanger(0.0, d, d, lambda t: d)).filter(sr, x)[0].astype(np.float64)
    H = np.abs(np.fft.rfft(y * np.hanning(len(y)))[1:] / np.fft.rfft(x[0].astype(np.float64) * np.hanning(len(y)))[1:])
    f = np.fft.rfftfreq(len(y), 1 / sr)[1:]
    peak_bins = np.abs((f * d) % 1.0 - 0.0) < 0.02
    notch_bins = np.abs((f * d) % 1.0 - 0.5) < 0.02
    assert np.median(H[peak_bins]) > 1.8 and np.median(H[notch_bins]) < 0.25
    # the phaser's ten first-order allpasses leave the magnitude of the wet path flat: with feedback 0 the output is the dry signal
    y0 = OracleUnit(phaser(0.0, lambda t: 0.3)).filter(sr, x[:, :2000])
    assert np.array_equal(y0, x[:, :2000])


def test_mls_is_maximum_length():  # src/noise.rs:11-98: the sequence of an n-bit MLS repeats after exactly 2**n - 1 steps
    for n in range(2, 15):
        u = OracleUnit(mls_bits(n))
        period = (1 << n) - 1
        x = u.process_many(2 * period, None)[0]
        assert set(np.unique(x)) == {-1.0, 1.0}
        assert np.array_equal(x[:period], x[period:])
        assert int((x[:period] > 0).sum()) == 1 << (n - 1)          # balance property: one more 1 than 0
        ac = [float(np.dot(x[:period], np.roll(x[:period], k))) for k in (1, 2, 5)]
        assert ac == [-1.0, -1.0, -1.0], (n, ac)                    # two-valued autocorrelation <=> maximal period


def test_polyblep_oscillators_known_values():  # src/oscillator.rs:510-760
    sr, f = 44100.0, 441.0
    u = OracleUnit(poly_saw_hz(f).phase(0.0) | ramp_hz(f).phase(0.0) | poly_square_hz(f).phase(0.0) | poly_pulse_hz(f, 0.25).phase(0.0))
    w = u.render(sr, 400 / sr)
    ph = (np.arange(400) * 0.01) % 1.0
    d = np.abs(w[1] - ph)
    assert np.minimum(d, 1.0 - d).max() < 1e-4                       # Ramp outputs the phase itself (f32 accumulation)
    mid = (ph > 0.02) & (ph < 0.97)
    assert np.abs(w[0][mid] - (2.0 * ph[mid] - 1.0)).max() < 1e-4    # saw is naive away from the step
    assert w[0][0] == 0.0                                            # and the BLEP residual centres the step
    sq_mid = mid & (np.abs(ph - 0.5) > 0.02)
    assert np.array_equal(w[2][sq_mid], np.where(ph[sq_mid] < 0.5, 1.0, -1.0).astype(np.float32))
    pu_mid = mid & (np.abs(ph - 0.25) > 0.02)
    assert np.array_equal(w[3][pu_mid], np.where(ph[pu_mid] < 0.25, 1.0, -1.0).astype(np.float32))
    imp = OracleUnit(impulse(1)).process_many(70, None)[0]
    assert imp[0] == 1.0 and not imp[1:].any()


def test_tap_reads_exact_delay():  # src/delay.rs:141-286: an integer tap delay reproduces the input shifted by that many samples
    sr = 44100.0
    for mk in (tap, tap_linear):
        u = OracleUnit((pass_() | dc(100.0 / sr)) >> mk(0.0, 0.01))
        rng = np.random.default_rng(5)
        x = rng.uniform(-1, 1, (1, 700)).astype(np.float32)
        y = u.filter(sr, x)[0]
        assert np.abs(y[100:] - x[0, :600]).max() < 1e-6 and not y[:98].any()


# ---------------------------------------------------------------- equivalences (test_basic.rs:392-406,520-529)
def is_equal(a, b, n=200, seed=1):
    ua, ub = OracleUnit(a), OracleUnit(b)
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1, 1, (max(1, ua.inputs()), n)).astype(np.float32)
    return np.array_equal(ua.process_many(n, x), ub.process_many(n, x))


def test_structural_equivalences():
    v, w, x, y, z = 1.0, -2.0, 3.0, -4.0, 5.0
    assert is_equal((pass_() ^ mul(y)) >> add(z) + sub(x), add(z) & mul(y) >> sub(x))
    assert is_equal((pass_() ^ mul(y) ^ add(w)) >> add(z) + sub(x) + mul(y), add(z) & mul(y) >> sub(x) & add(w) >> mul(y))
    assert is_equal(tick() >> tick() >> tick(), delay(3.0 / 44100.0))
    assert is_equal(tick() >> tick() >> tick() >> tick() >> tick(), delay(5.0 / 44100.0))
    assert is_equal((pass_() ^ mul(y) ^ add(w) ^ sub(x)) >> add(z) + sub(x) + mul(y) + add(z), add(z) & mul(y) >> sub(x) & add(w) >> mul(y) & sub(x) >> add(z))   # :402-406
    # multichannel constants vs. stacked constants (test_basic.rs:480-505)
    assert is_equal(dc(w) | dc(x), constant((w, x)))
    assert is_equal(dc(x) | dc(y) | dc(z), constant((x, y, z)))
    assert is_equal(dc(x) | dc(y) | dc(z) | dc(w), constant((x, y, z, w)))
    assert is_equal(dc(w) | dc(v) | dc(x) | dc(y) | dc(z), constant((w, v, x, y, z)))
    assert is_equal(dc((w, x)) | dc((y, z, w)), constant((w, x, y, z, w)))
    # sinks and zeros (test_basic.rs:507-517)
    assert is_equal(sink() | sink() | zero() | zero(), zero() | zero() | sink() | sink())
    assert is_equal(sink() | zero() | sink() | zero() | zero() | sink() | zero(), zero() | zero() | zero() | sink() | sink() | zero() | sink())


def outputs_diverge(g, n=64):
    y = OracleUnit(g).render(44100.0, n / 44100.0)
    for i in range(y.shape[0]):
        for j in range(i + 1, y.shape[0]):
            if np.array_equal(y[i], y[j]):
                return False
    return True


def test_nodes_vs_networks():  # tests/test_basic.rs:409-466 (the parts that do not remove vertices)
    from fundsp_b200.net import Net
    pt = Net(2, 2); pt.pass_through(0, 0); pt.pass_through(1, 1)
    assert is_equal(pass_() | pass_(), pt)
    sw = Net(2, 2); sw.pass_through(0, 1); sw.pass_through(1, 0)
    assert is_equal(reverse(2), sw)
    mn = Net(2, 2)
    id0 = mn.push(mul(2.0)); mn.push(sink()); mn.push(sine()); id1 = mn.push(mul(3.0))      # two idle vertices stay in the net
    mn.connect_input(0, id0, 0); mn.connect_input(1, id1, 0); mn.connect_output(id0, 0, 0); mn.connect_output(id1, 0, 1)
    assert is_equal(mul(2.0) | mul(3.0), mn)
    an = Net(2, 2)
    a0 = an.push(add((2.0, 3.0))); a1 = an.push(multipass(2))
    an.pipe_input(a0); an.pipe_all(a0, a1); an.pipe_output(a1)
    assert is_equal(add((2.0, 3.0)), an)
    # Net operators against the static combinators (test_basic.rs:468-518 style)
    x, y = lowpass_hz(800.0, 1.0), highpass_hz(300.0, 2.0)
    assert is_equal(x >> y, Net.wrap(x) >> Net.wrap(y))
    assert is_equal(x & y, Net.wrap(x) & Net.wrap(y))
    assert is_equal(x ^ y, Net.wrap(x) ^ Net.wrap(y))
    assert is_equal(x | y, Net.wrap(x) | Net.wrap(y))
    assert is_equal(x + y, Net.wrap(x) + Net.wrap(y)) and is_equal(x * y, Net.wrap(x) * Net.wrap(y)) and is_equal(x - y, Net.wrap(x) - Net.wrap(y))


def test_pseudorandom_phase_divergence():  # test_basic.rs:532-612
    assert outputs_diverge(noise() | (~zero() >> noise()) | noise() | (~zero() >> noise()) | noise() | noise() | noise())
    assert outputs_diverge(noise() ^ noise() ^ noise() & zero() ^ noise() ^ (noise() >> pass_()) ^ noise() ^ noise())
    assert outputs_diverge((sine_hz(1.0) >> pass_()) | sine_hz(1.0) | (sine_hz(1.0) >> pass_() >> pass_()) | sine_hz(1.0) | sine_hz(1.0))
    assert outputs_diverge(sine_hz(1.0) ^ sine_hz(1.0) ^ sine_hz(1.0) | sine_hz(1.0) | sine_hz(1.0))
    assert outputs_diverge(sine_hz(1.0) - zero() | sine_hz(1.0) - zero())
    assert outputs_diverge(noise() | noise())
    assert outputs_diverge((dc(110.0) >> saw()) | (dc(110.0) >> saw()))
    assert outputs_diverge((dc(110.0) >> square()) | (dc(110.0) >> triangle()) | (dc(110.0) >> square()))
    # two structurally identical voices built separately are identical
    a = OracleUnit(sine_hz(1.0)).render(44100.0, 0.01)
    b = OracleUnit(sine_hz(1.0)).render(44100.0, 0.01)
    assert np.array_equal(a, b)


# ---------------------------------------------------------------- composites, tables, block quirks
def test_reverb_composite_matches_python_prelude():
    L.fo_set_denormal_emulation(0)
    a = OracleUnit(L.fo_reverb_stereo(10.0, 2.0, 0.5))
    b = OracleUnit(reverb_stereo(10.0, 2.0, 0.5))
    assert a.leaf_hashes() == b.leaf_hashes()
    x = np.random.default_rng(0).uniform(-1, 1, (2, 4000)).astype(np.float32)
    ya, yb = a.filter(48000.0, x), b.filter(48000.0, x)
    assert np.array_equal(ya, yb) and np.abs(ya).max() > 1e-3
    L.fo_restore_denormals()


def test_wavetables():
    n = L.fo_wavetable_count(0)
    assert n == 40  # 20 * 2^(k/4) <= 20 kHz
    total, peak = 0, 0.0
    for i in range(n):
        ln = L.fo_wavetable_len(0, i)
        assert ln & (ln - 1) == 0 and 32 <= ln <= 8192
        t = np.ctypeslib.as_array(L.fo_wavetable_data(0, i), (ln,))
        total += ln
        peak = max(peak, float(np.abs(t).max()))
        assert abs(float(t.astype(np.float64).mean())) < 1e-6
    assert total == 41024 and abs(peak - 1.0) < 1e-6
    assert abs(L.fo_wavetable_pitch(0, 4) - 40.0) < 1e-4
    # a 110 Hz saw has the harmonic series 1/k up to the band limit
    y = OracleUnit(saw_hz(110.0)).render(44100.0, 1.0)[0].astype(np.float64)
    spec = np.abs(np.fft.rfft(y))
    assert all(abs(spec[110] / spec[110 * k] - k) < 1e-3 * k for k in (2, 3, 4, 5))


def test_wide_sin_and_floor():
    x = np.linspace(-200.0, 200.0, 20001).astype(np.float32)
    y = np.array([L.fo_wide_sinf(float(v)) for v in x])
    assert np.abs(y - np.sin(x.astype(np.float64))).max() < 4e-7
    for v in (0.0, 0.25, 0.9999999, 1.0, 1.5, 7.99999, 27.3):
        assert L.fo_wide_floorf(v) == math.floor(v)


def test_block_path_quirks():
    # Sine: block path keeps the phase unwrapped inside a block, tail samples go through tick (oscillator.rs:74-86)
    a = OracleUnit(sine_hz(20000.0))
    a.set_sample_rate(48000.0)
    y = np.concatenate([a.process(61)[0], a.process(64)[0]])
    b = OracleUnit(sine_hz(20000.0))
    b.set_sample_rate(48000.0)
    t = np.array([b.tick()[0] for _ in range(125)])
    assert np.abs(y - t).max() < 1e-4 and not np.array_equal(y, t)
    # size 0 is a no-op
    assert a.process(0).shape == (1, 0)
    # Join scales then adds in process (audionode.rs:642-659)
    g = OracleUnit((dc(0.1) | dc(0.2) | dc(0.7)) >> join(3))
    z = np.float32(1.0) / np.float32(3.0)
    assert g.process(8)[0, 0] == (np.float32(0.1) * z + np.float32(0.2) * z) + np.float32(0.7) * z


def test_net_equals_static_and_bus_tree():
    import oracle as O
    be = O.OracleBackend()
    v = [sine_hz(110.0 * (i + 1)) >> lowpass_hz(1000.0, 1.0) >> pan(0.0) for i in range(4)]
    nets = [L.fo_net_wrap(g.lower(be)) for g in v]
    top = L.fo_net_combine(0, L.fo_net_combine(0, nets[0], nets[1]), L.fo_net_combine(0, nets[2], nets[3]))
    u = OracleUnit(top)
    assert L.fo_net_size(top) == 4 + 6 and L.fo_net_has_cycle(top) == 0
    y = u.render(48000.0, 0.01)
    assert y.shape == (2, 480) and np.abs(y).max() > 0.1
    # cycle detection (test_basic.rs:355-362)
    c = L.fo_net_new(2, 1)
    i1 = L.fo_net_chain(c, join(2).lower(be))
    i2 = L.fo_net_chain(c, pass_().lower(be))
    assert L.fo_net_has_cycle(c) == 0
    L.fo_net_connect(c, i2, 0, i1, 1)
    assert L.fo_net_has_cycle(c) == 1
    L.fo_free(c)
    # Net == static graph for a chain (test_basic.rs:409-466)
    n = L.fo_net_new(0, 2)
    L.fo_net_chain(n, (noise() | noise()).lower(be))
    L.fo_net_chain(n, (moog_hz(1500.0, 0.5) | moog_hz(1000.0, 0.6)).lower(be))
    y1 = OracleUnit(n).render(44100.0, 0.01)
    assert y1.shape == (2, 441) and np.isfinite(y1).all()


def test_reference_net_check_wave_lines():  # tests/test_basic.rs:275-305,319-326: the Net scheduler itself, tick == process
    from fundsp_b200.net import Net
    L.fo_set_denormal_emulation(0)
    net = Net(0, 2)
    i = net.push(noise() >> moog_hz(1500.0, 0.8) | noise() >> moog_hz(500.0, 0.4))
    net.connect_output(i, 0, 0); net.connect_output(i, 1, 1)
    check_wave(net)                                                                    # :275-282
    net = Net(0, 2)
    net.chain(noise() | noise()); net.chain(moog_hz(1500.0, 0.5) | moog_hz(1000.0, 0.6)); net.chain(lowpole_hz(1000.0) | lowpole_hz(500.0))
    check_wave(net)                                                                    # :284-289
    static = (noise() | noise()) >> (moog_hz(1500.0, 0.5) | moog_hz(1000.0, 0.6)) >> (lowpole_hz(1000.0) | lowpole_hz(500.0))
    assert OracleUnit(net).render(44100.0, 0.01).shape == OracleUnit(static).render(44100.0, 0.01).shape
    net = Net(0, 2)
    net.chain(noise()); net.chain(lowpole_hz(1000.0) ^ lowpole_hz(500.0)); net.chain(lowpole_hz(1000.0) | lowpole_hz(500.0))
    check_wave(net)                                                                    # :291-296
    net = Net.wrap(sine_hz(42.0))
    net = net._copy() | net
    net.chain(Net.wrap(reverb_stereo(10.0, 5.0, 0.5)))
    check_wave(net)                                                                    # :298-303
    dc42 = Net.wrap(dc(42.0))
    dcs = dc42._copy() | dc42
    filt = Net.wrap(lowpass_hz(1729.0, 1.0))
    check_wave(dcs >> Net.wrap(reverb_stereo(40.0, 5.0, 1.0)) >> (filt._copy() | filt))  # :319-326
    L.fo_restore_denormals()


def test_bank_render_mix_is_index_order_sum():
    ex = [white().seed(i) >> lowpass_hz(500.0 + 100.0 * i, 1.0) for i in range(5)]
    out, mix = oracle_bank_render(ex, 48000.0, 200, per_voice=True, mix=True, threads=1)
    acc = out[0, 0].copy()
    for i in range(1, 5):
        acc = acc + out[i, 0]
    assert np.array_equal(acc, mix[0])
    out2, _ = oracle_bank_render(ex, 48000.0, 200, threads=3)
    assert np.array_equal(out, out2)
