"""Generates tests/golden/int_vectors.json: known answers for the reference's integer paths, computed by an
independent pure-Python (arbitrary precision int) restatement of
  src/math.rs:569-576 (rnd1), :589-597 (hash1), :649-658 (AttoHash::hash), src/noise.rs:150-157 (hash32x),
and the ping-derived node hashes of a few graphs (src/audionode.rs:156-161 + combinator ping rules).
The reference cannot be executed here (no Rust toolchain), so these pin the C++ oracle and the product's
host code against implementation slips, not against a misreading of the reference.
Run: python tests/golden/gen_int_vectors.py
"""
import json
import os

M = (1 << 64) - 1


def rnd1_bits(x):
    x ^= 0x5555555555555555
    x = (x * 0x9E3779B97F4A7C15) & M
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M
    x ^= x >> 31
    return x >> 11  # rnd1 = bits * 2^-53


def hash1(x):
    x ^= 0x5555555555555555
    x = (x * 0x517CC1B727220A95) & M
    x = ((x ^ (x >> 32)) * 0xD6E8FEB86659FD93) & M
    x = ((x ^ (x >> 32)) * 0xD6E8FEB86659FD93) & M
    return x ^ (x >> 32)


def atto(state, data):
    r = ((state << 5) | (state >> 59)) & M
    return ((r ^ data) * 0x517CC1B727220A95) & M


def hash32x(x):
    m = 0x45D9F3B
    for _ in range(3):
        x = ((x ^ (x >> 16)) * m) & 0xFFFFFFFF
    return x


# -- ping over a tiny structural description: ("leaf", id) | ("node", id, [children])
def ping(tree, probe, h, out):
    if tree[0] == "leaf":
        if not probe:
            out.append(h)
        return atto(h, tree[1])
    h = atto(h, tree[1])
    for c in tree[2]:
        h = ping(c, probe, h, out)
    return h


def leaf_hashes(tree):
    h = ping(tree, True, tree[1], [])
    out = []
    ping(tree, False, h, out)
    return out


C, SINE, SVF, NOISE, WAVE = ("leaf", 2), ("leaf", 21), ("leaf", 43), ("leaf", 20), ("leaf", 34)


def pipe(a, b):
    return ("node", 6, [a, b])


graphs = {
    "sine_hz>>lowpass_hz": pipe(pipe(C, SINE), SVF),
    "saw_hz>>lowpass_hz": pipe(pipe(C, WAVE), SVF),
    "white>>lowpass_hz": pipe(NOISE, SVF),
    "noise|noise": ("node", 7, [NOISE, NOISE]),
    "fm": pipe(("node", 4, [("node", 4, [("node", 4, [pipe(C, SINE)])])]), SINE),
}

xs = [0, 1, 2, 3, 0xDEADBEEF, 0x0123456789ABCDEF, M, 1 << 63, 12345678901234567]
vec = {
    "rnd1_bits": [[x, rnd1_bits(x)] for x in xs],
    "hash1": [[x, hash1(x)] for x in xs],
    "attohash": [[s, d, atto(s, d)] for s in xs[:5] for d in (0, 6, 21, 63)],
    "hash32x": [[x & 0xFFFFFFFF, hash32x(x & 0xFFFFFFFF)] for x in xs],
    "leaf_hashes": {k: leaf_hashes(v) for k, v in graphs.items()},
}
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "int_vectors.json"), "w") as f:
    json.dump(vec, f, indent=1)
print("ok")
