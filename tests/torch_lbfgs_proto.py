"""TEST INFRASTRUCTURE ONLY: an eager-PyTorch projected L-BFGS over a batched evaluator, the first prototype of the
device-resident optimiser.  The product's optimisers are the CUDA ones behind tpl_optimize_batch
(treepl_b200/csrc/tpl_lbfgs.cuh, tpl_tn.cuh); this class stays as an independent cross-check that drives the ORACLE on
the CPU (tests/test_batchopt_cpu.py)."""
from __future__ import annotations

import numpy as np
import torch

LARGE = 1e15


class BatchedLBFGS:
    def __init__(self, plp, P, B, n_rates, lower, upper, mode=1, history=24, device=None, date_scale=1.0):
        """plp: a callable  X[P,B] (torch) -> (F[B], G[P,B])  (the tests pass the CPU oracle).
        n_rates: the first n_rates parameters are rates (optimised in log space).
        lower/upper: numpy [P] bounds in the ORIGINAL parameters.
        date_scale: dates are optimised as h / date_scale; with date_scale ~ a typical branch duration the date
        curvature c/d^2 * scale^2 ~ c is of the same order as the log-rate curvature r d ~ c."""
        self.plp, self.P, self.B, self.nr, self.mode, self.m = plp, P, B, n_rates, mode, history
        assert callable(plp)
        self.evaluator = plp
        if device is None:
            device = torch.device("cpu")
        self.dev = device
        self.ds = float(date_scale)
        lo = np.array(lower, dtype=np.float64).copy()
        hi = np.array(upper, dtype=np.float64).copy()
        lo[:n_rates] = np.log(np.maximum(lo[:n_rates], 1e-300))
        hi[:n_rates] = np.log(hi[:n_rates])
        # Date bounds are kept OPEN by a relative margin: exactly ON a penalised bound the reference's
        # objectives drop the (infinite) barrier term (calc_penalty resets, the AD paths skip it,
        # SURVEY 8 a13), so a step clamped onto the bound would land in a spurious discontinuous minimum —
        # the reference's own TNC runs end there (examples/test.cppr8s: final PL 234.23 vs the interior
        # optimum 232.31).  Just inside the bound the barrier is large and the line search backs off.
        dlo, dhi = lo[n_rates:], hi[n_rates:]
        margin = 1e-9 * np.maximum(1.0, np.abs(np.where(dhi < 1e19, dhi, dlo)))
        dlo += margin
        dhi -= np.where(dhi < 1e19, margin, 0.0)
        lo[n_rates:] /= self.ds
        hi[n_rates:] /= self.ds
        self.lo = torch.from_numpy(lo).to(self.dev)[:, None]
        self.hi = torch.from_numpy(hi).to(self.dev)[:, None]
        self.X = torch.empty((P, B), dtype=torch.float64, device=self.dev)     # original parameters (kernel input)
        self.G = torch.empty((P, B), dtype=torch.float64, device=self.dev)
        self.F = torch.empty(B, dtype=torch.float64, device=self.dev)
        self.nfev = 0

    # ---- objective in the transformed variables ----
    def _eval(self, Z):
        """Z [P,B] transformed -> (F [B], dF/dZ [P,B]).  One launch pair for all B vectors."""
        nr = self.nr
        torch.exp(Z[:nr], out=self.X[:nr])
        torch.mul(Z[nr:], self.ds, out=self.X[nr:])
        f, g = self.evaluator(self.X)
        self.F.copy_(f)
        self.G.copy_(g)
        self.nfev += 1
        G = self.G.clone()
        G[:nr] *= self.X[:nr]                        # chain rule d/du = r d/dr
        G[nr:] *= self.ds
        F = self.F.clone()
        bad = ~torch.isfinite(F) | (F >= LARGE) | ~torch.isfinite(G).all(0)
        F = torch.where(bad, torch.full_like(F, LARGE), F)
        G = torch.where(bad[None, :], torch.zeros_like(G), G)
        return F, G, bad

    def minimize(self, X0, max_iter=2000, ftol=1e-12, gtol=1e-8, patience=5, verbose=False):
        """X0: torch/numpy [P,B] in ORIGINAL parameters (feasible).  Returns (X [P,B] numpy, F [B] numpy, info)."""
        nr, m, B, P = self.nr, self.m, self.B, self.P
        X0 = torch.as_tensor(X0, dtype=torch.float64, device=self.dev)
        Z = X0.clone()
        Z[:nr] = torch.log(X0[:nr])
        Z[nr:] = X0[nr:] / self.ds
        Z = torch.minimum(torch.maximum(Z, self.lo), self.hi)
        F, G, bad = self._eval(Z)
        if bad.any():
            raise ValueError("start point infeasible for vectors %s" % torch.nonzero(bad).flatten().tolist()[:8])
        S = torch.zeros((m, P, B), dtype=torch.float64, device=self.dev)
        Y = torch.zeros_like(S)
        rho = torch.zeros((m, B), dtype=torch.float64, device=self.dev)
        valid = torch.zeros((m, B), dtype=torch.bool, device=self.dev)
        gamma = torch.ones(B, dtype=torch.float64, device=self.dev)
        done = torch.zeros(B, dtype=torch.bool, device=self.dev)
        stall = torch.zeros(B, dtype=torch.int32, device=self.dev)
        head = 0
        first = True
        it = 0
        for it in range(1, max_iter + 1):
            # projected gradient: variables sitting on a bound with the gradient pushing outward are frozen
            at_lo = (Z <= self.lo) & (G > 0)
            at_hi = (Z >= self.hi) & (G < 0)
            free = ~(at_lo | at_hi)
            Gp = torch.where(free, G, torch.zeros_like(G))
            gnorm = Gp.abs().amax(0)
            done = done | (gnorm <= gtol * torch.clamp(F.abs(), min=1.0))
            if bool(done.all()):
                break
            # two-loop recursion, all vectors at once
            q = Gp.clone()
            alphas = []
            order = [(head - 1 - k) % m for k in range(m)]
            for k in order:
                a = torch.where(valid[k], rho[k] * (S[k] * q).sum(0), torch.zeros_like(gamma))
                q -= a[None, :] * Y[k]
                alphas.append(a)
            q *= gamma[None, :]
            for k, a in zip(reversed(order), reversed(alphas)):
                b = torch.where(valid[k], rho[k] * (Y[k] * q).sum(0), torch.zeros_like(gamma))
                q += (a - b)[None, :] * S[k]
            D = -torch.where(free, q, torch.zeros_like(q))
            gd = (Gp * D).sum(0)
            # not a descent direction (or first step): steepest descent, scaled
            use_sd = (gd >= 0) | ~torch.isfinite(gd)
            if first:
                use_sd = torch.ones_like(use_sd)
            sd_scale = 1.0 / torch.clamp(gnorm, min=1e-300)
            D = torch.where(use_sd[None, :], -Gp * sd_scale[None, :] * 0.1, D)
            gd = (Gp * D).sum(0)
            # backtracking Armijo line search with per-vector step
            t = torch.ones(B, dtype=torch.float64, device=self.dev)
            t = torch.where(done, torch.zeros_like(t), t)
            Zn, Fn, Gn = Z, F, G
            accepted = done.clone()
            Zbest = Z.clone()
            Fbest = F.clone()
            Gbest = G.clone()
            for ls in range(40):
                Zt = torch.minimum(torch.maximum(Z + t[None, :] * D, self.lo), self.hi)
                Ft, Gt, badt = self._eval(Zt)
                ok = (~badt) & (Ft <= F + 1e-4 * t * gd) & ~accepted
                Zbest = torch.where(ok[None, :], Zt, Zbest)
                Fbest = torch.where(ok, Ft, Fbest)
                Gbest = torch.where(ok[None, :], Gt, Gbest)
                accepted = accepted | ok
                if bool(accepted.all()):
                    break
                t = torch.where(accepted, t, t * 0.5)
            failed = ~accepted
            # curvature pairs
            s = Zbest - Z
            y = Gbest - G
            sy = (s * y).sum(0)
            yy = (y * y).sum(0)
            good = (sy > 1e-10 * torch.sqrt((s * s).sum(0) * yy + 1e-300)) & ~done & ~failed
            S[head] = s
            Y[head] = y
            rho[head] = torch.where(good, 1.0 / torch.where(good, sy, torch.ones_like(sy)), torch.zeros_like(sy))
            valid[head] = good
            gamma = torch.where(good, sy / torch.where(good, yy, torch.ones_like(yy)), gamma)
            head = (head + 1) % m
            # convergence bookkeeping
            rel = (F - Fbest).abs() / torch.clamp(F.abs(), min=1.0)
            small = (rel <= ftol) | failed
            stall = torch.where(small, stall + 1, torch.zeros_like(stall))
            # a failed line search drops the history of that vector (restart from steepest descent)
            valid = valid & ~failed[None, :]
            done = done | (stall >= patience)
            Z, F, G = Zbest, Fbest, Gbest
            first = False
            if verbose and it % 50 == 0:
                print("it %d  fev %d  done %d/%d  F[0]=%.10g  |g|max=%.3g" % (it, self.nfev, int(done.sum()), B, float(F[0]), float(gnorm.max())))
        Xout = Z.clone()
        Xout[:nr] = torch.exp(Z[:nr])
        Xout[nr:] = Z[nr:] * self.ds
        if self.dev.type == "cuda":
            torch.cuda.synchronize()
        return Xout.cpu().numpy(), F.cpu().numpy(), {"iterations": it, "nfev": self.nfev, "converged": done.cpu().numpy()}
