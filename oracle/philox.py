"""Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11) and the unit normals the fused
tick derives from it (csrc/rtd3_tick.cu: philox_normal2).  TEST INFRASTRUCTURE.

There is no counterpart in the reference: its exploration noise is np.random.normal on the global MT19937 stream (robot.py:640),
which the exact mode reproduces bit for bit (oracle/mt19937.py).  The throughput mode of the batched loop draws its normals from
this counter-based generator instead; this restatement pins the device code to the published algorithm through the Random123
known-answer vectors (tests/test_oracle_philox.py).
"""
import numpy as np

M0, M1 = 0xD2511F53, 0xCD9E8D57
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """ctr: 4 uint32, key: 2 uint32 -> 4 uint32 (ten rounds)."""
    c0, c1, c2, c3 = (int(v) & MASK for v in ctr)
    k0, k1 = (int(v) & MASK for v in key)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & MASK, p1 & MASK, ((p0 >> 32) ^ c3 ^ k1) & MASK, p0 & MASK
        k0, k1 = (k0 + W0) & MASK, (k1 + W1) & MASK
    return c0, c1, c2, c3


def normal2(seed, tick, env):
    """The two unit normals of (env, tick): counter (env lo, env hi, tick lo, tick hi), key (seed lo, seed hi); two 53-bit uniforms
    built like numpy's random_double, Box-Muller with 1 - u1 in (0, 1]."""
    r = philox4x32_10((env & MASK, env >> 32, tick & MASK, tick >> 32), (seed & MASK, seed >> 32))
    u1 = ((r[0] >> 5) * 67108864.0 + (r[1] >> 6)) / 9007199254740992.0
    u2 = ((r[2] >> 5) * 67108864.0 + (r[3] >> 6)) / 9007199254740992.0
    rad = np.sqrt(-2.0 * np.log(1.0 - u1))
    return rad * np.cos(2.0 * np.pi * u2), rad * np.sin(2.0 * np.pi * u2)
