"""Separable CMA-ES on the GPU: ``ask`` / ``tell`` over torch CUDA tensors.

The reference drives pycma (``cma.CMAEvolutionStrategy(x0, sigma0, {"popsize": n})``, cmaes.py:94-96); pycma is not
installed, and at the dimension of the relaxed one-hot of a protein (AAV 735 x 20 = 14 700) a full covariance matrix is
out of reach anyway.  This is the published separable variant (Ros & Hansen 2008; Hansen's tutorial for everything
else), the same update as the host sampler ``flexs_b200.utils.cma`` (the CPU test checks the two against each other),
with the population as ONE ``float32[popsize, N]`` tensor that never leaves the device: sampling, the weighted
recombination and the diagonal update are a handful of elementwise / reduction launches, so a population of tens of
thousands of relaxed sequences costs what its HBM traffic costs.  Minimises the values passed to ``tell`` (like pycma).
"""
import math
from typing import Optional


class SepCMA:
    def __init__(self, x0, sigma0: float, popsize: int, seed: Optional[int] = None):
        import torch

        self.torch = torch
        self.mean = x0.detach().to(torch.float32).reshape(-1).clone()
        self.device = self.mean.device
        self.N = n = int(self.mean.numel())
        self.sigma = float(sigma0)
        self.popsize = int(popsize)
        self.gen = torch.Generator(device=self.device)
        self.gen.manual_seed(int(seed) if seed is not None else torch.seed() & 0x7FFFFFFF)
        self.mu = self.popsize // 2
        w = math.log(self.mu + 0.5) - torch.log(torch.arange(1, self.mu + 1, dtype=torch.float64))
        w = w / w.sum()
        self.mueff = float(1.0 / (w ** 2).sum())
        self.weights = w.to(torch.float32).to(self.device)
        self.cc = (4 + self.mueff / n) / (n + 4 + 2 * self.mueff / n)
        self.cs = (self.mueff + 2) / (n + self.mueff + 5)
        c1 = 2 / ((n + 1.3) ** 2 + self.mueff)
        cmu = min(1 - c1, 2 * (self.mueff - 2 + 1 / self.mueff) / ((n + 2) ** 2 + self.mueff))
        scale = (n + 2) / 3.0  # sep-CMA-ES: the diagonal model can learn (n + 2) / 3 times faster
        self.c1 = min(1.0, c1 * scale)
        self.cmu = min(1 - self.c1, cmu * scale)
        self.damps = 1 + 2 * max(0.0, math.sqrt((self.mueff - 1) / (n + 1)) - 1) + self.cs
        self.chiN = math.sqrt(n) * (1 - 1 / (4 * n) + 1 / (21 * n * n))
        self.pc = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.ps = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.diagC = torch.ones(n, dtype=torch.float32, device=self.device)
        self.countiter = 0

    def ask(self):
        """``float32[popsize, N]`` candidate solutions (a fresh tensor every call)."""
        torch = self.torch
        z = torch.randn((self.popsize, self.N), dtype=torch.float32, device=self.device, generator=self.gen)
        z.mul_(torch.sqrt(self.diagC)).mul_(self.sigma).add_(self.mean)
        return z

    def tell(self, solutions, function_values) -> None:
        torch = self.torch
        f = function_values.to(torch.float64).reshape(-1)
        order = torch.argsort(f, stable=True)[: self.mu]     # minimisation; ties by position like np.argsort(kind="stable")
        y = (solutions[order] - self.mean) / self.sigma
        yw = self.weights @ y
        self.mean = self.mean + self.sigma * yw
        self.countiter += 1
        self.ps = (1 - self.cs) * self.ps + math.sqrt(self.cs * (2 - self.cs) * self.mueff) * (yw / torch.sqrt(self.diagC))
        ps_norm = float(torch.linalg.vector_norm(self.ps))
        hsig = float((ps_norm / math.sqrt(1 - (1 - self.cs) ** (2 * self.countiter)) / self.chiN) < (1.4 + 2 / (self.N + 1)))
        self.pc = (1 - self.cc) * self.pc + hsig * math.sqrt(self.cc * (2 - self.cc) * self.mueff) * yw
        delta = (1 - hsig) * self.cc * (2 - self.cc)
        self.diagC = ((1 - self.c1 - self.cmu) * self.diagC + self.c1 * (self.pc ** 2 + delta * self.diagC)
                      + self.cmu * (self.weights @ (y ** 2)))
        self.sigma *= math.exp((self.cs / self.damps) * (ps_norm / self.chiN - 1))
