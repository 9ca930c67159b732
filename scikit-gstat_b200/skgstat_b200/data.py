"""Seeded synthetic inputs for bench.py and examples (no gstools in the image).

gaussian_field: stationary Gaussian random field by the randomisation method
(sum of random cosine modes; Gaussian spectrum -> Gaussian covariance with the
given correlation length) + white noise, the generator SURVEY.md section 8(d)
specifies for BASELINE.json's synthetic configurations.
"""
import numpy as np


def gaussian_field(coords, corr_len, seed=0, modes=256, noise=0.05, mean=10.0, std=2.0):
    rng = np.random.default_rng(seed + 7919)
    c = np.asarray(coords, dtype=float)
    dim = c.shape[1]
    k = rng.normal(0.0, 1.0 / corr_len, size=(modes, dim))
    phase = rng.uniform(0, 2 * np.pi, size=modes)
    out = np.zeros(len(c))
    step = 65536
    for i in range(0, len(c), step):
        out[i:i + step] = np.cos(c[i:i + step] @ k.T + phase).sum(axis=1)
    out *= np.sqrt(2.0 / modes)
    out = out + noise * rng.normal(size=len(c))
    return mean + std * out


def random_points(n, extent=1000.0, dim=2, seed=0):
    rng = np.random.default_rng(seed)
    return rng.random((n, dim)) * extent
