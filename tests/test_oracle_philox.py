"""oracle/philox.py against the Random123 known-answer vectors, and the law of the normals it derives (CPU)."""
import numpy as np

from oracle import philox as P


def test_philox4x32_10_known_answers():
    for ctr, key, out in P.KAT:
        got = tuple(int(x) for x in P.philox4x32_10(*ctr, *key))
        assert got == out, [hex(g) for g in got]


def test_direction_normals_have_gaussian_moments_and_tails():
    """Box-Muller on 24-bit uniforms: radius capped at sqrt(2 * 25 * ln 2) = 5.9 sigma; the tail frequencies up to
    there must be those of N(0, 1) (|z| > 3: 2.70e-3, |z| > 4: 6.33e-5)."""
    z = P.direction_normals(seed=20261018, chain=np.arange(20000), draw=3, D=128).ravel()      # 2.56e6 normals
    n = z.size
    assert abs(z.mean()) < 4 / np.sqrt(n) and abs(z.var() - 1) < 4 * np.sqrt(2 / n)
    assert abs(np.mean(z ** 4) - 3) < 4 * np.sqrt(96 / n)
    for cut, p in ((3.0, 2.6998e-3), (4.0, 6.3342e-5)):
        k = int((np.abs(z) > cut).sum())
        assert abs(k - n * p) < 4.5 * np.sqrt(n * p), (cut, k, n * p)
    assert np.abs(z).max() < 5.9
    u_col, z_init, z_prop, u = P.chain_scalars(7, np.arange(100000), 11)
    assert 0 < u.min() and u.max() < 1 and abs(u.mean() - 0.5) < 4 / np.sqrt(12 * u.size)
    assert abs(z_prop.mean()) < 4 / np.sqrt(z_prop.size) and abs(z_prop.var() - 1) < 4 * np.sqrt(2 / z_prop.size)
