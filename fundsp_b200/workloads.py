"""The synthetic voice banks of BASELINE.json `configs` (concretised in SURVEY.md §8d).

Per-voice parameters are drawn with the reference's own `rnd1` (src/math.rs:569-576): u_k(i) = rnd1(4*i + k),
so every backend (GPU bank, CPU oracle, a future Rust host) builds bit-identical voices.
"""
from __future__ import annotations

import math

from .graph import f32
from .prelude import (adsr_live, bandpass_hz, biquad_bank, dc, highpass_hz, lowpass_hz, moog, moog_hz, multipass, pan, reverb_stereo,
                      saw, saw_hz, sine, sine_hz, stacki, white)

M64 = (1 << 64) - 1
SR = 48000.0


def rnd1(x: int) -> float:
    x = (x ^ 0x5555555555555555) & M64
    x = (x * 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    x ^= x >> 31
    return (x >> 11) * (1.0 / (1 << 53))


def u(i, k):
    return rnd1(4 * i + k)


def lerp(a, b, t):
    return a * (1.0 - t) + b * t


def xerp(a, b, t):
    return math.exp(lerp(math.log(a), math.log(b), t))


# ---- config 1: plumbing
def plumbing():
    return sine_hz(440.0) >> lowpass_hz(1000.0, 1.0)


# ---- config 2: FM bank, voice i = sine_hz(f) * f * m + f >> sine()  (README.md:102)
def fm_voice(i):
    f = f32(xerp(55.0, 1760.0, u(i, 0)))
    m = f32(lerp(0.5, 8.0, u(i, 1)))
    return sine_hz(f).phase(u(i, 2)) * f * m + f >> sine().phase(u(i, 3))


# ---- config 3a: white().seed(i) >> lowpass_hz(fc, q)
def noise_svf_voice(i):
    fc = xerp(100.0, 12000.0, u(i, 0))
    q = lerp(0.5, 10.0, u(i, 1))
    return white().seed(i) >> lowpass_hz(fc, q)


# ---- north-star headline: saw_hz(f) >> lowpass_hz(fc, q)
def saw_svf_voice(i):
    fc = xerp(100.0, 12000.0, u(i, 0))
    q = lerp(0.5, 10.0, u(i, 1))
    f = xerp(55.0, 1760.0, u(i, 2))
    return saw_hz(f).phase(u(i, 3)) >> lowpass_hz(fc, q)


# ---- config 3b: one biquad_bank() = 8 voices; lane coefficients BiquadCoefs::lowpass (src/biquad.rs:62-74) via Setting::biquad(..).index(l)
def biquad_lowpass_coefs(sr, cutoff, q):
    import numpy as np
    F = np.float32
    omega = F(6.2831855) * F(cutoff) / F(sr)
    alpha = np.sin(omega, dtype=F) / (F(2.0) * F(q))
    beta = np.cos(omega, dtype=F)
    a0r = F(1.0) / (F(1.0) + alpha)
    a1 = F(-2.0) * beta * a0r
    a2 = (F(1.0) - alpha) * a0r
    b1 = (F(1.0) - beta) * a0r
    b0 = b1 * F(0.5)
    return (float(a1), float(a2), float(b0), float(b1), float(b0))


def biquad_bank_unit(b):
    node = biquad_bank()
    for lane in range(8):
        i = 8 * b + lane
        fc = xerp(100.0, 12000.0, u(i, 0))
        q = lerp(0.5, 10.0, u(i, 1))
        node = node.set(6, biquad_lowpass_coefs(SR, fc, q), address=[(1, lane)])
    return stacki(8, lambda lane: white().seed(8 * b + lane)) >> node


# ---- config 4: subtractive voice with feedback reverb; 1 input (gate), 2 outputs
def subtractive_dry_voice(i):
    f = xerp(55.0, 880.0, u(i, 0))
    fc = xerp(200.0, 8000.0, u(i, 1))
    q = lerp(0.1, 0.9, u(i, 2))
    p = lerp(-1.0, 1.0, u(i, 3))
    return (((dc(f) >> saw()) | dc((fc, q))) >> moog()) * adsr_live(0.01, 0.1, 0.6, 0.3) >> pan(p)


def subtractive_voice(i):
    return subtractive_dry_voice(i) >> (multipass(2) & 0.2 * reverb_stereo(10.0, 2.0, 0.5))


def gate_signal(n, sr=SR):
    """Gate for config 4: low for 10 ms (adsr_live arms on a low->high edge, src/adsr.rs:34-41), high until 0.5 s, then low."""
    import numpy as np
    g = np.zeros((1, n), np.float32)
    g[0, int(0.01 * sr): int(0.5 * sr)] = 1.0
    return g


# ---- config 5: dynamic Net, 4 voice classes round-robin, each >> pan(p)
def net_voice(i):
    p = lerp(-1.0, 1.0, u(i, 3))
    k = i & 3
    if k == 0:
        g = sine_hz(xerp(55.0, 1760.0, u(i, 0))) >> lowpass_hz(xerp(100.0, 12000.0, u(i, 1)), lerp(0.5, 10.0, u(i, 2)))
    elif k == 1:
        g = saw_hz(xerp(55.0, 1760.0, u(i, 0))) >> moog_hz(xerp(200.0, 8000.0, u(i, 1)), lerp(0.1, 0.9, u(i, 2)))
    elif k == 2:
        g = white().seed(i) >> bandpass_hz(xerp(100.0, 12000.0, u(i, 1)), lerp(0.5, 10.0, u(i, 2)))
    else:
        f = f32(xerp(55.0, 1760.0, u(i, 0)))
        m = f32(lerp(0.5, 8.0, u(i, 1)))
        g = (sine_hz(f) * f * m + f >> sine()) >> highpass_hz(xerp(100.0, 12000.0, u(i, 1)), lerp(0.5, 10.0, u(i, 2)))
    return g >> pan(p)


# ---- the headline voices as sequencer events (src/sequencer.rs): notes that start within the first half second and are held, so the
# steps after the first measure the steady state of Event<X> (the block plan + X's own group form) against the plain `saw_svf` bank
def saw_svf_event_voice(i):
    from .sequencer import event, Fade
    start = 0.5 * u(i, 4)
    return event(saw_svf_voice(i), start, start + 1.0e6, Fade.Smooth, 0.005, 0.0)


# ---- the dense tap contraction (north-star: tensor cores "only on the dense FIR/convolve tap contraction"): every voice convolves its own
# noise with ONE shared K-tap response (class-uniform data), src/convolve.rs:9-59
def conv_response(K=1000):
    import numpy as np
    k = np.arange(K)
    h = np.array([2.0 * rnd1(1_000_003 + int(j)) - 1.0 for j in k]) * np.exp(-k / (K / 4.0))
    return (h / np.abs(h).max()).astype(np.float32)


_CONV_H = {}


def conv_voice(i, K=1000):
    from .prelude import convolve
    if K not in _CONV_H:
        _CONV_H[K] = conv_response(K)
    return white().seed(i) * f32(lerp(0.25, 1.0, u(i, 0))) >> convolve(_CONV_H[K])


WORKLOADS = {
    # name: (voice builder, default voices, inputs)
    "fm": (fm_voice, 4096),
    "noise_svf": (noise_svf_voice, 16384),
    "saw_svf": (saw_svf_voice, 16384),
    "biquad_bank": (biquad_bank_unit, 2048),
    "subtractive_dry": (subtractive_dry_voice, 1024),
    "subtractive": (subtractive_voice, 1024),
    "net": (net_voice, 65536),
    "saw_svf_events": (saw_svf_event_voice, 16384),
    "conv": (conv_voice, 16384),
}


def build(name, voices=None, first=0):
    fn, default = WORKLOADS[name]
    n = default if voices is None else voices
    return [fn(first + i) for i in range(n)]
