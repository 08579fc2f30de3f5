"""ORACLE (test infrastructure): NumPy restatement of the dropout bit source of the CUDA path.

tf.nn.dropout (vlmap/modules.py:82, vqa/model_vlmap_answer.py:180) draws from TF's stateful RNG, which no
other implementation can reproduce; the CUDA path draws keep bits from Philox4x32-10 (Salmon et al., SC'11)
instead. This file pins that generator twice: against the published Random123 known-answer vectors
(tests/test_golden.py) and, on the GPU, bit-for-bit against vqa_dropout_masks().

Mask definition (csrc/philox.cuh): elements are taken in groups of 8 consecutive flat indices; group g draws
philox4x32_10(counter = (g_lo, g_hi, site, step_lo), key = (seed_lo, seed_hi ^ step_hi)); element j of the
group keeps iff the j-th 16-bit field (little end first) is < floor(keep * 65536).
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK32 = np.uint64(0xFFFFFFFF)
SITE_ATT, SITE_JOINT = 1, 2


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised over the counter words (uint64 arrays holding 32-bit values); keys are Python ints."""
    c0, c1, c2, c3 = (np.asarray(c, dtype=np.uint64) & MASK32 for c in (c0, c1, c2, c3))
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK32
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK32
        n0 = hi1 ^ c1 ^ np.uint64(k0)
        n2 = hi0 ^ c3 ^ np.uint64(k1)
        c0, c1, c2, c3 = n0, lo1, n2, lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def keep_mask(n, keep, seed, step, site):
    """0/1 uint8 array of n elements (n % 8 == 0): the bits vqa_dropout_masks() materialises."""
    assert n % 8 == 0
    thr = int(min(max(np.float32(keep) * np.float32(65536.0), 0.0), 65536.0))
    if thr >= 65536:
        return np.ones(n, np.uint8)
    g = np.arange(n // 8, dtype=np.uint64)
    k0 = seed & 0xFFFFFFFF
    k1 = ((seed >> 32) ^ (step >> 32)) & 0xFFFFFFFF
    w = philox4x32_10(g & MASK32, g >> np.uint64(32), np.full_like(g, site), np.full_like(g, step & 0xFFFFFFFF),
                      k0, k1)
    out = np.empty((n // 8, 8), np.uint8)
    for j in range(8):
        u = (w[j >> 1] >> np.uint64((j & 1) * 16)) & np.uint64(0xFFFF)
        out[:, j] = u < np.uint64(thr)
    return out.reshape(n)


SITE_NOISE = 4


def normal_draw(n, seed, step):
    """The N(0, 1) draw of the 'full' variant's reparameterisation (csrc/variants.cu normal4; replaces
    tf.random_normal(seed=123) of vqa/model_vlmap_answer_full.py:133, which only TF can reproduce): elements in
    groups of 4; group g draws philox4x32_10 at site 4; u_i = ((w_i >> 8) + 0.5) / 2^24; Box-Muller on (u0, u1) and
    (u2, u3): sqrt(-2 ln u0) * (cos, sin)(2 pi u1), sqrt(-2 ln u2) * (cos, sin)(2 pi u3). float64 here; the device
    evaluates the same formula in fp32."""
    assert n % 4 == 0
    g = np.arange(n // 4, dtype=np.uint64)
    k0 = seed & 0xFFFFFFFF
    k1 = ((seed >> 32) ^ (step >> 32)) & 0xFFFFFFFF
    w = philox4x32_10(g & MASK32, g >> np.uint64(32), np.full_like(g, SITE_NOISE), np.full_like(g, step & 0xFFFFFFFF),
                      k0, k1)
    u = [((x >> np.uint64(8)).astype(np.float64) + 0.5) / 16777216.0 for x in w]
    r0, r1 = np.sqrt(-2.0 * np.log(u[0])), np.sqrt(-2.0 * np.log(u[2]))
    out = np.stack([r0 * np.cos(2 * np.pi * u[1]), r0 * np.sin(2 * np.pi * u[1]),
                    r1 * np.cos(2 * np.pi * u[3]), r1 * np.sin(2 * np.pi * u[3])], axis=1)
    return out.reshape(n)
