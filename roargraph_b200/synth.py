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


class TorchSynth:
    """The generator's parameters on `device` (same distribution family as make_numpy, torch's generator): base rows once,
    then any number of query sets from the SAME modality-gap distribution — training queries, test queries, and the
    10 M-query batch stream of BASELINE.json configs[3]."""

    def __init__(self, nb: int, dim: int, device, seed: int = 1234, latent: int = 64, chunk: int = 1 << 20):
        import torch

        self.nb, self.dim, self.device, self.latent, self.chunk = nb, dim, device, latent, chunk
        self.gen = torch.Generator(device=device)
        self.gen.manual_seed(seed)
        self.nclus = max(16, nb // 1000)
        self.A = self._rn(latent, dim) / latent ** 0.5
        self.cent = self._rn(self.nclus, latent)
        self._have_base = False

    def _rn(self, *s):
        import torch
        return torch.randn(*s, generator=self.gen, device=self.device, dtype=torch.float32)

    def base(self):
        import torch

        assert not self._have_base, "the base is drawn once, before any query set (the generator is sequential)"
        out = torch.empty(self.nb, self.dim, device=self.device, dtype=torch.float32)
        for s in range(0, self.nb, self.chunk):
            e = min(self.nb, s + self.chunk)
            c = torch.randint(0, self.nclus, (e - s,), generator=self.gen, device=self.device)
            out[s:e] = (self.cent[c] + self._rn(e - s, self.latent)) @ self.A + 0.3 * self._rn(e - s, self.dim)
        self.g = 0.5 * self._rn(self.dim)
        self.B = self.A + 0.15 * self._rn(self.latent, self.dim) / self.latent ** 0.5
        self._have_base = True
        return out

    def queries(self, n: int, out=None):
        import torch

        assert self._have_base
        if out is None:
            out = torch.empty(n, self.dim, device=self.device, dtype=torch.float32)
        for s in range(0, n, self.chunk):
            e = min(n, s + self.chunk)
            c = torch.randint(0, self.nclus, (e - s,), generator=self.gen, device=self.device)
            out[s:e] = (self.cent[c] + 1.2 * self._rn(e - s, self.latent)) @ self.B + 0.3 * self._rn(e - s, self.dim) + self.g
        return out


def make_torch(nb: int, nq_train: int, nq_test: int, dim: int, device, seed: int = 1234, latent: int = 64,
               chunk: int = 1 << 20, return_gen: bool = False):
    """Same distribution family generated on `device` with torch's generator (not bit-identical to make_numpy)."""
    gen = TorchSynth(nb, dim, device, seed=seed, latent=latent, chunk=chunk)
    base = gen.base()
    train = gen.queries(nq_train)
    test = gen.queries(nq_test)
    if return_gen:
        return base, train, test, gen
    return base, train, test
