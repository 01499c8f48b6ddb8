"""Known-answer and statistical checks of the numpy Philox restatement (oracle/philox_ref.py)."""
import numpy as np

from oracle import philox_ref


def test_philox4x32_10_random123_known_answers():
    # Random123 kat_vectors, philox4x32 with 10 rounds
    kats = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
            ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
            ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
             (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, expect in kats:
        out = philox_ref.philox4x32(np.array(ctr, dtype=np.uint32), np.array(key, dtype=np.uint32), rounds=10)
        assert tuple(int(v) for v in out) == expect


def test_stream_statistics():
    e = philox_ref.standard_normals(2, 100000, 16, 0x1255, 1)
    assert abs(e.mean()) < 2e-3 and abs(e.std() - 1.0) < 2e-3
    c = np.corrcoef(e[0].T)
    assert np.max(np.abs(c - np.eye(16))) < 0.02
    # different timesteps / iterations / streams give different noise
    a = philox_ref.words_for(1, 8, 4, 1, 1)
    assert not np.array_equal(a, philox_ref.words_for(1, 8, 4, 1, 2))
    assert not np.array_equal(a, philox_ref.words_for(1, 8, 4, 1, 1, instance=1))
    assert not np.array_equal(a, philox_ref.words_for(1, 8, 4, 1, 1, t0=1))
    # global indexing: an offset range equals the tail of the full range
    full = philox_ref.words_for(2, 10, 7, 9, 3)
    np.testing.assert_array_equal(full[1:, 4:], philox_ref.words_for(1, 6, 7, 9, 3, t0=1, i0=4))


def test_antithetic_stream_bookkeeping():
    """Antithetic stream: samples 2q and 2q + 1 share counter q and differ by the sign; any window [i0, i0+N)
    of the stream is the same numbers (sharding over ranks by i0)."""
    w = philox_ref.words_for(2, 11, 7, 9, 3, antithetic=True)
    np.testing.assert_array_equal(w[:, 0:10:2], w[:, 1:11:2])
    np.testing.assert_array_equal(w[:, 0:10:2], philox_ref.words_for(2, 5, 7, 9, 3))
    z = philox_ref.deltas(2, 11, np.ones(7), 9, 3, antithetic=True)
    np.testing.assert_array_equal(z[:, 0:10:2], -z[:, 1:11:2])
    np.testing.assert_array_equal(z[:, 0:10:2], philox_ref.deltas(2, 5, np.ones(7), 9, 3))
    np.testing.assert_array_equal(z[:, 3:], philox_ref.deltas(2, 8, np.ones(7), 9, 3, i0=3, antithetic=True))
    e = philox_ref.standard_normals(1, 200000, 8, 0x1255, 1, antithetic=True)[0]
    assert abs(e.mean()) < 1e-12                        # exactly symmetric
    assert abs(e.std() - 1.0) < 5e-3


def test_seven_round_stream_statistics():
    """The sample stream uses Philox4x32 with 7 rounds (the smallest count Salmon et al. report as
    Crush-resistant; 10 is the default safety margin).  Beyond the known-answer vectors of the round
    function: uniformity of the raw words (chi-square over 256 byte bins, every byte lane), no linear
    correlation between neighbouring counters, samples, timesteps and word lanes, and the first four moments
    and the tail mass of the normals built from them."""
    T, N, d = 4, 250000, 16
    w = philox_ref.words_for(T, N, d, 0x1255, 1)                          # [T, N, 4, 4] uint32
    flat = w.reshape(-1)
    for shift in (0, 8, 16, 24):
        counts = np.bincount((flat >> np.uint32(shift)) & np.uint32(0xFF), minlength=256)
        expect = flat.size / 256.0
        chi2 = float(np.sum((counts - expect) ** 2 / expect))
        assert 160.0 < chi2 < 370.0, (shift, chi2)                        # chi-square(255): mean 255, sd 22.6
    u = (w.astype(np.float64) + 0.5) / 2.0 ** 32
    c = u - 0.5

    def corr(a, b):
        return float(np.mean(a * b) / (1.0 / 12.0))
    lim = 5.0 / np.sqrt(T * (N - 1) * 16)                                 # 5 sigma of a sample correlation
    assert abs(corr(c[:, 1:], c[:, :-1])) < lim                           # neighbouring samples (counter + 1)
    assert abs(corr(c[1:], c[:-1])) < 5.0 / np.sqrt((T - 1) * N * 16)    # neighbouring timesteps
    assert abs(corr(c[:, :, 1:], c[:, :, :-1])) < 5.0 / np.sqrt(T * N * 12)   # neighbouring counter blocks
    assert abs(corr(c[..., 1:], c[..., :-1])) < 5.0 / np.sqrt(T * N * 12)     # neighbouring words of a block
    e = philox_ref.standard_normals(T, N, d, 0x1255, 1).reshape(-1)
    n = e.size
    assert abs(e.mean()) < 5.0 / np.sqrt(n)
    assert abs(e.var() - 1.0) < 5.0 * np.sqrt(2.0 / n)
    assert abs(np.mean(e ** 3)) < 5.0 * np.sqrt(15.0 / n)
    assert abs(np.mean(e ** 4) - 3.0) < 5.0 * np.sqrt(96.0 / n)
    tail = np.mean(np.abs(e) > 3.0)
    assert abs(tail - 0.0026998) < 5.0 * np.sqrt(0.0027 / n)
    # cross-coordinate independence of the normals of one sample (what the Gram fit relies on)
    z = philox_ref.standard_normals(1, 400000, d, 0x1255, 2)[0]
    C = (z.T @ z) / z.shape[0]
    assert float(np.max(np.abs(C - np.eye(d)))) < 5.0 / np.sqrt(z.shape[0]) * 1.5
