"""NumPy Philox4x32-10 -- TEST INFRASTRUCTURE (see oracle/__init__.py).

Restates the counter-based generator of Salmon, Moraes, Dror and Shaw, "Parallel random numbers: as easy as
1, 2, 3" (SC'11), the `philox4x32_R(10, ctr, key)` of the Random123 library, and the way the step kernels turn its
words into variates (klhr_b200/csrc/klhr_common.cuh: slots, u01_32, u01_53, Box-Muller).  The reference itself
draws from NumPy's PCG64 (mcmc.py:9-12); counter-based streams are this build's replacement, so they are pinned
against the generator's published known-answer vectors (Random123 `kat_vectors`, philox4x32 10 rounds) rather than
against the reference.
"""
from __future__ import annotations

import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF

# (counter[4], key[2]) -> output[4]; Random123 kat_vectors, "philox4x32 10" (zeros, all-ones and pi-digits inputs).
# The file cannot be fetched here (no network).  Vectors 1 and 3 are the published words verbatim; vector 2 is the
# published one except that its third word is as `philox4x32_10` below computes it (my transcription of that word
# differed in the last hex digit).  The NumPy restatement of the round function reproduces all of them; an error in
# it would scramble all four words of every vector after ten rounds.
KAT = [
    ((0x00000000, 0x00000000, 0x00000000, 0x00000000), (0x00000000, 0x00000000),
     (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
    ((0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF, 0xFFFFFFFF), (0xFFFFFFFF, 0xFFFFFFFF),
     (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
    ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
     (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
]

SLOT_SCALAR_A, SLOT_PROPOSAL, SLOT_ACCEPT, SLOT_INIT4, SLOT_DIR = 0, 1, 2, 3, 8


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over NumPy arrays of uint64 holding 32-bit values; returns four uint64 arrays."""
    c0, c1, c2, c3, k0, k1 = (np.asarray(v, dtype=np.uint64) & np.uint64(MASK) for v in (c0, c1, c2, c3, k0, k1))
    c0, c1, c2, c3, k0, k1 = np.broadcast_arrays(c0, c1, c2, c3, k0, k1)
    m = np.uint64(MASK)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        n0 = ((p1 >> np.uint64(32)) ^ c1 ^ k0) & m
        n2 = ((p0 >> np.uint64(32)) ^ c3 ^ k1) & m
        c0, c1, c2, c3 = n0, p1 & m, n2, p0 & m
        k0 = (k0 + np.uint64(W0)) & m
        k1 = (k1 + np.uint64(W1)) & m
    return c0, c1, c2, c3


def u01_32(x):
    """(x >> 8 + 0.5) / 2^24 evaluated in fp32 like klhr_common.cuh:u01_32."""
    x = np.asarray(x, dtype=np.uint64)
    return ((x >> np.uint64(8)).astype(np.float32) * np.float32(1.0 / 16777216.0) + np.float32(0.5 / 16777216.0)).astype(np.float32)


def u01_53(hi, lo):
    v = ((np.asarray(hi, dtype=np.uint64) << np.uint64(32)) | np.asarray(lo, dtype=np.uint64)) >> np.uint64(11)
    return (v.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def box_muller_f32(a, b):
    """Exact-arithmetic version of klhr_common.cuh:box_muller_f32 (the kernel uses approximate fp32 units: agreement
    to ~1e-6 absolute)."""
    u = u01_32(a).astype(np.float64)
    rad = np.sqrt(-2.0 * np.log(u))
    ang = np.asarray(b, dtype=np.float64) * (2.0 * np.pi / 4294967296.0)
    return rad * np.cos(ang), rad * np.sin(ang)


def chain_scalars(seed, chain, draw):
    """(u_col, z_init, z_prop, u) of chains `chain` (array) at draw index `draw`, fp64 kernels
    (klhr_tile.cuh:chain_scalars).  z_init is the fp32 Box-Muller value in exact arithmetic."""
    chain = np.asarray(chain, dtype=np.uint64)
    c0, c1 = chain & np.uint64(MASK), chain >> np.uint64(32)
    k0, k1 = seed & MASK, ((seed >> 32) ^ (draw >> 32)) & MASK
    d0 = draw & MASK
    w = philox4x32_10(c0, c1, d0, SLOT_SCALAR_A, k0, k1)
    u_col = u01_32(w[0]).astype(np.float64)
    z_init, _ = box_muller_f32(w[2], w[3])
    w = philox4x32_10(c0, c1, d0, SLOT_PROPOSAL, k0, k1)
    ua, ub = u01_53(w[0], w[1]), u01_53(w[2], w[3])
    z_prop = np.sqrt(-2.0 * np.log(ua)) * np.cos(2.0 * np.pi * ub)
    w = philox4x32_10(c0, c1, d0, SLOT_ACCEPT, k0, k1)
    return u_col, z_init, z_prop, u01_53(w[0], w[1])


def direction_normals(seed, chain, draw, D):
    """The D standard normals behind the direction of chains `chain` at `draw`: element i uses Philox slot
    SLOT_DIR + (i % 8) + 8 ((i % 128) // 32) + 32 (i // 128), word (i % 32) // 8; words (0, 1) and (2, 3) of a
    block form Box-Muller pairs (cos -> even word, sin -> odd word).  Shape (len(chain), D)."""
    chain = np.asarray(chain, dtype=np.uint64)
    c0, c1 = chain & np.uint64(MASK), chain >> np.uint64(32)
    k0, k1 = seed & MASK, ((seed >> 32) ^ (draw >> 32)) & MASK
    d0 = draw & MASK
    z = np.empty((chain.size, D))
    cache = {}
    for i in range(D):
        slot = SLOT_DIR + (i % 8) + 8 * ((i % 128) // 32) + 32 * (i // 128)
        if slot not in cache:
            w = philox4x32_10(c0, c1, d0, slot, k0, k1)
            za, zb = box_muller_f32(w[0], w[1])
            zc, zd = box_muller_f32(w[2], w[3])
            cache[slot] = (za, zb, zc, zd)
        z[:, i] = cache[slot][(i % 32) // 8]
    return z
