"""Wavelength-sharded single column (BASELINE config 5, SURVEY.md 8e: an experiment, not the production path).

One column is split over ranks by wavelength: rank r owns the contiguous range [lo_r, hi_r) of the spectrum and
runs the ordinary formal solution on the sub-problem made of those wavelengths (every transition clipped to the
range, its wavelength weights taken from the full grid).  J and I stay sharded; the rate matrices are summed over
ranks -- Gamma is linear in the per-wavelength integrands, and so is its diagonal fix-up
(Gamma[i,i] = -sum_l Gamma[l,i], rh_method.py:698-703), so with the collisional rates C kept on rank 0 only

    Gamma = allreduce_sum(Gamma_r),      dJ = allreduce_max(dJ_r)

and the statistical-equilibrium solve is replicated.  The summation order over wavelengths changes, so Gamma agrees
with the unsharded solve to rounding (~1e-13 relative), not bit for bit; J and I(lambda, mu) of a shard are
bit-identical to the unsharded ones as long as the populations are.
"""
import numpy as np
import torch
import torch.distributed as dist

from .tables import ModelTables, transition_offsets


def lambda_ranges(Nspect, world, align=1, cost=None):
    """[(lo, hi)] per rank: contiguous, covering [0, Nspect), boundaries on multiples of `align` (a tile width).
    cost: optional per-wavelength work estimate; the boundaries then equalise the ranks' summed cost instead of their
    wavelength counts (the formal solution of a wavelength costs about 1.1 + its number of active transitions, SURVEY.md
    8d -- with equal counts the rank holding the line cores was measured 1.5x slower than the mean, and the others
    spent that time waiting in the exchange)."""
    nblk = (Nspect + align - 1) // align
    out = []
    if cost is None:
        for r in range(world):
            b0 = (nblk * r) // world
            b1 = (nblk * (r + 1)) // world
            out.append((min(b0 * align, Nspect), min(b1 * align, Nspect)))
        return out
    c = np.zeros(nblk * align)
    c[:Nspect] = np.asarray(cost, dtype=np.float64)[:Nspect]
    cum = np.concatenate([[0.0], np.cumsum(c.reshape(nblk, align).sum(axis=1))])      # cost of the first b blocks
    b_prev = 0
    for r in range(world):
        if r == world - 1:
            b1 = nblk
        else:
            b1 = int(np.searchsorted(cum, cum[-1] * (r + 1) / world))
            b1 = min(max(b1, b_prev + 1), nblk - (world - 1 - r))     # at least one block for every rank
        out.append((min(b_prev * align, Nspect), min(b1 * align, Nspect)))
        b_prev = b1
    return out


def wavelength_cost(p):
    """Work estimate per wavelength of problem `p`: 1.1 + number of transitions active there (31 + 28 A)."""
    trans = np.asarray(p['trans'], dtype=np.int64).reshape(-1, 6)
    cost = np.full(int(p['Nspect']), 1.1)
    for (_, _, _, _, Nblue, Nlam) in trans:
        cost[int(Nblue):int(Nblue + Nlam)] += 1.0
    return cost


def lambda_shard_problem(p, lo, hi, keep_C):
    """Sub-problem holding wavelengths [lo, hi) of problem `p` (same atoms, levels, depth grid and populations)."""
    mt = ModelTables(p)                     # full-grid wavelength weights
    trans = np.asarray(p['trans'], dtype=np.int32).reshape(-1, 6)
    N, R = int(p['Nspace']), int(p['Nrays'])
    toff = transition_offsets(trans)
    phi = np.asarray(p['phi'], dtype=np.float64)
    phioff = np.asarray(p['phioff'], dtype=np.int64)
    q = dict(p)
    new_trans, new_alpha, new_w, new_phi, new_phioff, keep = [], [], [], [], [], []
    po = 0
    for t, (atom, i, j, isLine, Nblue, Nlam) in enumerate(trans):
        a, b = max(int(Nblue), lo), min(int(Nblue + Nlam), hi)
        if b <= a:
            continue
        keep.append(t)
        l0, l1 = a - int(Nblue), b - int(Nblue)
        new_trans.append([atom, i, j, isLine, a - lo, b - a])
        new_alpha.append(np.asarray(p['alpha'], dtype=np.float64)[toff[t] + l0:toff[t] + l1])
        new_w.append(mt.wlambda[toff[t] + l0:toff[t] + l1])
        if isLine:
            blk = phi[phioff[t]:phioff[t] + int(Nlam) * R * 2 * N].reshape(int(Nlam), R * 2 * N)[l0:l1]
            new_phi.append(blk.reshape(-1))
            new_phioff.append(po)
            po += blk.size
        else:
            new_phioff.append(0)
    if not keep:
        raise ValueError('no transition overlaps wavelengths [%d, %d)' % (lo, hi))
    keep = np.asarray(keep)
    q['Nspect'] = hi - lo
    q['wavelength'] = np.asarray(p['wavelength'], dtype=np.float64)[lo:hi]
    q['trans'] = np.asarray(new_trans, dtype=np.int32)
    q['linepar'] = np.asarray(p['linepar'], dtype=np.float64).reshape(-1, 4)[keep]
    q['alpha'] = np.concatenate(new_alpha)
    q['wlambda_table'] = np.concatenate(new_w)
    q['phi'] = np.concatenate(new_phi) if new_phi else np.zeros(0)
    q['phioff'] = np.asarray(new_phioff, dtype=np.int64)
    q['wphi'] = np.asarray(p['wphi'], dtype=np.float64)[keep]
    for k in ('aDamp',):
        if k in p:
            q[k] = np.asarray(p[k])[keep]
    for k in ('bg_chi', 'bg_eta', 'bg_sca'):
        q[k] = np.ascontiguousarray(np.asarray(p[k], dtype=np.float64)[lo:hi])
    q['C'] = np.asarray(p['C'], dtype=np.float64) if keep_C else np.zeros_like(np.asarray(p['C'], dtype=np.float64))
    return q


class GammaExchange:
    """The path's only exchange as ONE collective: every rank contributes [Gamma | dJ] (147 KB for the stress column);
    after the all-gather each rank adds the contributions in ascending rank order and takes the maximum of the dJ's --
    a fixed order, so every rank holds the same bits (the replicated statistical equilibrium stays in lockstep) and the
    result does not depend on the library's reduction tree.  One latency-bound call instead of two (sum, max)."""

    def __init__(self, gamma, dJ, group=None):
        self.gamma, self.dJ, self.group = gamma.view(-1), dJ.view(-1)[:1], group
        self.world = dist.get_world_size(group)
        n = self.gamma.numel()
        self.send = torch.empty(n + 1, dtype=gamma.dtype, device=gamma.device)
        self.recv = torch.empty(self.world * (n + 1), dtype=gamma.dtype, device=gamma.device)

    def __call__(self):
        n = self.gamma.numel()
        self.send[:n].copy_(self.gamma)
        self.send[n:].copy_(self.dJ)
        dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        R = self.recv.view(self.world, n + 1)
        torch.add(R[0, :n], R[1, :n], out=self.gamma)
        for r in range(2, self.world):
            self.gamma.add_(R[r, :n])
        torch.amax(R[:, n:], dim=0, out=self.dJ)


class LambdaShardedColumn:
    """One column, wavelength-sharded over the ranks of `group` (one process per GPU, NCCL over NVLink)."""

    def __init__(self, p, device=None, group=None, specialize=False):
        from .engine import MaliEngine
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        Lw = max(1, 32 // int(p['Nrays']))
        self.ranges = lambda_ranges(int(p['Nspect']), self.world, align=Lw, cost=wavelength_cost(p))
        self.lo, self.hi = self.ranges[self.rank]
        self.sub = lambda_shard_problem(p, self.lo, self.hi, keep_C=(self.rank == 0))
        self.eng = MaliEngine(self.sub, 1, device=device, specialize=(p if specialize else False))
        self.eng.upload([self.sub])
        self.exchange = GammaExchange(self.eng.t_Gamma, self.eng.t_dJ, group) if self.world > 1 else (lambda: None)

    def formal_sol_gamma_matrices(self):
        """FS on this rank's wavelengths, then the only exchange of the path: Gamma summed, dJ maximised."""
        self.eng.formal_sol_gamma_async()
        self.exchange()
        return float(self.eng.t_dJ[0].item())

    def stat_equil(self):
        """Replicated on every rank (same Gamma, same populations -> same result)."""
        return float(self.eng.stat_equil()[0])

    def n(self):
        return self.eng.n(0)

    def close(self):
        self.eng.close()
