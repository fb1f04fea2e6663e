"""Seeded synthetic cross-modal data of the shapes BASELINE.json names (SURVEY.md 8(d) generator).

No network in this environment, so t2i / LAION / WebVid are replaced by clustered base vectors and
out-of-distribution queries with a modality gap.  Works on numpy (host) or torch (device) so 10M x 200
can be generated directly in HBM on the GPU box.
"""
from __future__ import annotations

import numpy as np


def make_numpy(nb: int, nq_train: int, nq_test: int, dim: int, seed: int = 1234, latent: int = 64):
    rng = np.random.default_rng(seed)
    nclus = max(16, nb // 1000)
    A = (rng.standard_normal((latent, dim)) / np.sqrt(latent)).astype(np.float32)
    cent = rng.standard_normal((nclus, latent)).astype(np.float32)
    c = rng.integers(0, nclus, nb)
    base = (cent[c] + 1.0 * rng.standard_normal((nb, latent)).astype(np.float32)) @ A
    base = base + 0.3 * rng.standard_normal((nb, dim)).astype(np.float32)
    g = (0.5 * rng.standard_normal(dim)).astype(np.float32)
    B = A + (0.15 * rng.standard_normal((latent, dim)) / np.sqrt(latent)).astype(np.float32)

    def queries(n):
        cq = rng.integers(0, nclus, n)
        q = (cent[cq] + 1.2 * rng.standard_normal((n, latent)).astype(np.float32)) @ B
        return (q + 0.3 * rng.standard_normal((n, dim)).astype(np.float32) + g).astype(np.float32)

    train = queries(nq_train)
    test = queries(nq_test)
    return base.astype(np.float32), train, test


def make_torch(nb: int, nq_train: int, nq_test: int, dim: int, device, seed: int = 1234, latent: int = 64,
               chunk: int = 1 << 20):
    """Same distribution family generated on `device` with torch's generator (not bit-identical to make_numpy)."""
    import torch

    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    nclus = max(16, nb // 1000)
    rn = lambda *s: torch.randn(*s, generator=gen, device=device, dtype=torch.float32)
    A = rn(latent, dim) / latent ** 0.5
    cent = rn(nclus, latent)
    base = torch.empty(nb, dim, device=device, dtype=torch.float32)
    for s in range(0, nb, chunk):
        e = min(nb, s + chunk)
        c = torch.randint(0, nclus, (e - s,), generator=gen, device=device)
        base[s:e] = (cent[c] + rn(e - s, latent)) @ A + 0.3 * rn(e - s, dim)
    g = 0.5 * rn(dim)
    B = A + 0.15 * rn(latent, dim) / latent ** 0.5

    def queries(n):
        out = torch.empty(n, dim, device=device, dtype=torch.float32)
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            c = torch.randint(0, nclus, (e - s,), generator=gen, device=device)
            out[s:e] = (cent[c] + 1.2 * rn(e - s, latent)) @ B + 0.3 * rn(e - s, dim) + g
        return out

    train = queries(nq_train)
    test = queries(nq_test)
    return base, train, test
