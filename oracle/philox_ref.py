"""TEST INFRASTRUCTURE ONLY — numpy restatement of the in-kernel noise generator.

The reference draws its perturbations with numpy's global legacy MT19937 stream inside a user
closure (e.g. examples/pendulum/pendulum_zero_order.py:38-43) and never seeds it, so there is no
reference bit stream to match.  The CUDA fast path instead defines its own counter-based stream
(Philox4x32 with STREAM_ROUNDS = 7 rounds, Salmon et al. SC'11; same round function as Random123 /
cuRAND, pinned on the 10-round known-answer vectors) whose *integer*
bookkeeping is checked bit-for-bit against this file:

  counter = (sample index i, timestep t, (iter << 8) | word-block j, instance id)
  key     = (seed & 0xffffffff, seed >> 32)
  words   -> w[0..3];  normals e[4j+0..4j+3] from Box-Muller on (w0,w1) and (w2,w3):
      f(w)  = float32 with bit pattern (w >> 9) | 0x3f800000          in [1, 2)
      r     = sqrt(-2 ln(2 - f(wa)))                                  (2 - f in (0, 1])
      theta = 2 pi (f(wb) - 1.5)                                      in [-pi, pi)
      e_a, e_b = r cos(theta), r sin(theta)
  delta z[i, c] = sigma[c] * e[c]   for c < d  (dx in the first n columns, du in the last m)

Antithetic stream (GaussianSampling(antithetic=True), kernel flag IRS_ANTITHETIC): samples come in
pairs that share one counter — sample index i uses counter word 0 = i >> 1 and
  delta z[i, c] = (+1 if i even else -1) * sigma[c] * e[c].
"""
import numpy as np

M0 = np.uint64(0xD2511F53)
M1 = np.uint64(0xCD9E8D57)
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)
STREAM_ROUNDS = 7     # must equal kPhiloxRounds in irs_mpc_b200/csrc/common.cuh


def philox4x32(counter, key, rounds=10):
    """counter: uint32 array [..., 4]; key: uint32 array [..., 2] (broadcastable). Returns [..., 4]."""
    c = [np.asarray(counter[..., i], dtype=np.uint64) for i in range(4)]
    k0 = np.asarray(key[..., 0], dtype=np.uint64)
    k1 = np.asarray(key[..., 1], dtype=np.uint64)
    for r in range(rounds):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(W0)) & MASK
        k1 = (k1 + np.uint64(W1)) & MASK
    return np.stack(c, axis=-1).astype(np.uint32)


def words_for(T, N, d, seed, it, instance=0, t0=0, i0=0, antithetic=False):
    """uint32 words [T, N, ceil(d/4), 4] exactly as the kernel draws them."""
    nblk = (d + 3) // 4
    i = (np.arange(N, dtype=np.uint64) + np.uint64(i0))[None, :, None]
    if antithetic:
        i = i >> np.uint64(1)
    t = (np.arange(T, dtype=np.uint64) + np.uint64(t0))[:, None, None]
    j = np.arange(nblk, dtype=np.uint64)[None, None, :]
    ctr = np.zeros((T, N, nblk, 4), dtype=np.uint32)
    ctr[..., 0] = (i + 0 * t + 0 * j).astype(np.uint32)
    ctr[..., 1] = (t + 0 * i + 0 * j).astype(np.uint32)
    ctr[..., 2] = ((np.uint64(it) << np.uint64(8)) | (j + 0 * i + 0 * t)).astype(np.uint32)
    ctr[..., 3] = np.uint32(instance)
    key = np.array([seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF], dtype=np.uint32)
    return philox4x32(ctr, key, rounds=STREAM_ROUNDS)


def unit_float(w):
    """float32 in [1,2) built from the top 23 bits of w (exact, same as the kernel)."""
    bits = (np.asarray(w, dtype=np.uint32) >> np.uint32(9)) | np.uint32(0x3F800000)
    return bits.view(np.float32)


def box_muller(wa, wb):
    fa = unit_float(wa).astype(np.float64)
    fb = unit_float(wb).astype(np.float64)
    r = np.sqrt(-2.0 * np.log(2.0 - fa))
    th = 2.0 * np.pi * (fb - 1.5)
    return r * np.cos(th), r * np.sin(th)


def standard_normals(T, N, d, seed, it, instance=0, t0=0, i0=0, antithetic=False):
    """float64 normals [T, N, d] (the kernel computes the same values in float32)."""
    w = words_for(T, N, d, seed, it, instance, t0, i0, antithetic)
    e0, e1 = box_muller(w[..., 0], w[..., 1])
    e2, e3 = box_muller(w[..., 2], w[..., 3])
    e = np.stack((e0, e1, e2, e3), axis=-1).reshape(T, N, -1)
    if antithetic:
        sign = np.where((np.arange(N, dtype=np.uint64) + np.uint64(i0)) & np.uint64(1), -1.0, 1.0)
        e = e * sign[None, :, None]
    return e[..., :d]


def deltas(T, N, sigma, seed, it, instance=0, t0=0, i0=0, antithetic=False):
    sigma = np.asarray(sigma, dtype=np.float32).astype(np.float64)
    return standard_normals(T, N, sigma.shape[0], seed, it, instance, t0, i0, antithetic) * sigma
