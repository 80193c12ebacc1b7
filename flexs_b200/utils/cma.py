"""A small (mu/mu_w, lambda)-CMA-ES with the ask/tell interface the CMAES explorer needs.

The reference drives the third-party ``cma`` package (``cma.CMAEvolutionStrategy(x0, sigma0,
{"popsize": n})``, cmaes.py:94-96; docs pin cma==3.0.3), which is not installed here.  This is the
published algorithm (Hansen, "The CMA Evolution Strategy: A Tutorial", 2016) with the default strategy
parameters; like pycma it MINIMISES the values passed to ``tell``.  Above ``full_cov_max_dim``
dimensions (AAV 735 x 20 = 14 700) the full covariance matrix would need 14 700^2 entries, so the
separable variant (diagonal C, learning rates scaled as in Ros & Hansen 2008) is used instead.
"""
from typing import List, Optional

import numpy as np


class CMAEvolutionStrategy:
    def __init__(self, x0, sigma0: float, opts: Optional[dict] = None, full_cov_max_dim: int = 512):
        opts = dict(opts or {})
        self.mean = np.array(x0, dtype=np.float64).reshape(-1)
        self.N = n = self.mean.size
        self.sigma = float(sigma0)
        self.popsize = int(opts.get("popsize", 4 + int(3 * np.log(n))))
        self.rng = np.random.default_rng(opts.get("seed"))
        self.mu = self.popsize // 2
        w = np.log(self.mu + 0.5) - np.log(np.arange(1, self.mu + 1))
        self.weights = w / w.sum()
        self.mueff = 1.0 / np.sum(self.weights ** 2)
        self.separable = n > full_cov_max_dim
        self.cc = (4 + self.mueff / n) / (n + 4 + 2 * self.mueff / n)
        self.cs = (self.mueff + 2) / (n + self.mueff + 5)
        self.c1 = 2 / ((n + 1.3) ** 2 + self.mueff)
        self.cmu = min(1 - self.c1, 2 * (self.mueff - 2 + 1 / self.mueff) / ((n + 2) ** 2 + self.mueff))
        if self.separable:  # sep-CMA-ES: the diagonal model can learn (n + 2) / 3 times faster
            scale = (n + 2) / 3.0
            self.c1 = min(1.0, self.c1 * scale)
            self.cmu = min(1 - self.c1, self.cmu * scale)
        self.damps = 1 + 2 * max(0.0, np.sqrt((self.mueff - 1) / (n + 1)) - 1) + self.cs
        self.chiN = np.sqrt(n) * (1 - 1 / (4 * n) + 1 / (21 * n * n))
        self.pc = np.zeros(n)
        self.ps = np.zeros(n)
        if self.separable:
            self.diagC = np.ones(n)
        else:
            self.C = np.eye(n)
            self.B = np.eye(n)
            self.D = np.ones(n)
            self._eigen_age = 0
        self.countiter = 0
        self._last_z = None

    # -- sampling -----------------------------------------------------------------------------
    def _refresh_eigen(self):
        if self.separable:
            return
        if self._eigen_age > max(1, int(1 / ((self.c1 + self.cmu) * self.N * 10))):
            self.C = np.triu(self.C) + np.triu(self.C, 1).T
            d2, self.B = np.linalg.eigh(self.C)
            self.D = np.sqrt(np.maximum(d2, 1e-20))
            self._eigen_age = 0

    def ask(self, number: Optional[int] = None) -> List[np.ndarray]:
        lam = number or self.popsize
        z = self.rng.standard_normal((lam, self.N))
        if self.separable:
            y = z * np.sqrt(self.diagC)
        else:
            self._refresh_eigen()
            y = (z * self.D) @ self.B.T
        self._last_y = y
        return [self.mean + self.sigma * yi for yi in y]

    def ask_and_eval(self, func):
        """``cma``'s convenience used at cmaes.py:108: sample a population and evaluate it one by one."""
        xs = self.ask()
        return xs, [func(x) for x in xs]

    # -- update -------------------------------------------------------------------------------
    def tell(self, solutions, function_values) -> None:
        xs = np.asarray(solutions, dtype=np.float64)
        f = np.asarray(function_values, dtype=np.float64)
        n, mu, w = self.N, self.mu, self.weights
        order = np.argsort(f, kind="stable")  # minimisation
        y = (xs[order[:mu]] - self.mean) / self.sigma
        yw = w @ y
        self.mean = self.mean + self.sigma * yw
        self.countiter += 1
        if self.separable:
            invsqrt_yw = yw / np.sqrt(self.diagC)
        else:
            invsqrt_yw = self.B @ ((self.B.T @ yw) / self.D)
        self.ps = (1 - self.cs) * self.ps + np.sqrt(self.cs * (2 - self.cs) * self.mueff) * invsqrt_yw
        hsig = (np.linalg.norm(self.ps) / np.sqrt(1 - (1 - self.cs) ** (2 * self.countiter)) / self.chiN) < (1.4 + 2 / (n + 1))
        self.pc = (1 - self.cc) * self.pc + hsig * np.sqrt(self.cc * (2 - self.cc) * self.mueff) * yw
        delta = (1 - hsig) * self.cc * (2 - self.cc)
        if self.separable:
            self.diagC = ((1 - self.c1 - self.cmu) * self.diagC + self.c1 * (self.pc ** 2 + delta * self.diagC)
                          + self.cmu * (w @ (y ** 2)))
        else:
            rank_mu = (y.T * w) @ y
            self.C = ((1 - self.c1 - self.cmu) * self.C + self.c1 * (np.outer(self.pc, self.pc) + delta * self.C)
                      + self.cmu * rank_mu)
            self._eigen_age += 1
        self.sigma *= float(np.exp((self.cs / self.damps) * (np.linalg.norm(self.ps) / self.chiN - 1)))
