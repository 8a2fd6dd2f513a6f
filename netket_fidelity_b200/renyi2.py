"""Renyi2EntanglementEntropy -- drop-in for the observable the reference re-exports at
netket_fidelity/renyi2.py:1 (netket.experimental.observable.Renyi2EntanglementEntropy).

Swap estimator on an MCState (NetKet 3.10 semantics; parity unpinned): the flattened
samples are split in two halves (sigma, sigma'); with sigma~ = (sigma'_A, sigma_B) and
sigma~' = (sigma_A, sigma'_B),
    q = psi(sigma~) psi(sigma~') / (psi(sigma) psi(sigma')),   S2 = -log2 Re mean(q),
error propagated as sigma_q / (mean ln 2).  The four amplitude evaluations and the reduction run
in CUDA (nkfb_renyi2)."""

from __future__ import annotations

import math

import numpy as np
import torch


class Renyi2EntanglementEntropy:
    def __init__(self, hilbert, partition):
        self._hilbert = hilbert
        part = np.unique(np.asarray(list(partition), dtype=np.int64))
        if part.size and (part.min() < 0 or part.max() >= hilbert.size):
            raise ValueError("Invalid partition: sites outside the Hilbert space")
        self._partition = part

    hilbert = property(lambda self: self._hilbert)
    partition = property(lambda self: self._partition)

    def mask_words(self, device):
        N = self._hilbert.size
        W32 = (N + 31) // 32
        words = np.zeros(W32, dtype=np.uint32)
        for i in self._partition:
            words[i >> 5] |= np.uint32(1 << (i & 31))
        return torch.tensor(words.view(np.int32), device=device)

    def __repr__(self):
        return f"Renyi2EntanglementEntropy(hilbert={self.hilbert}, partition={self.partition})"


def renyi2_expect(vstate, op):
    from .nk.stats import Stats
    if op.hilbert != vstate.hilbert:
        raise TypeError("Hilbert spaces should match")
    if len(op.partition) in (0, vstate.hilbert.size):
        return Stats(mean=0.0, error_of_mean=0.0, variance=0.0)
    smp = vstate.samples_packed()
    W32 = smp.shape[-1]
    flat = smp.view(-1, W32)
    S = flat.shape[0]
    if S % 2:
        flat = flat[:S - 1]
    q, sums = vstate._h.renyi2(vstate._spec(), flat.contiguous(), op.mask_words(vstate.device), vstate._ls)
    d = torch.distributed
    if d.is_available() and d.is_initialized() and d.get_world_size() > 1:
        d.all_reduce(sums, op=d.ReduceOp.SUM)
    s = sums.cpu().numpy()
    n = s[3]
    mean_q = s[0] / n
    var_q = max(0.0, s[1] / n - mean_q * mean_q)
    n_rows = max(1, min(vstate._n_chains_local, q.numel()))
    while q.numel() % n_rows:
        n_rows -= 1
    qs = vstate._stats_of(q.real.contiguous(), n_rows)
    err_q = qs.error_of_mean if not math.isnan(qs.error_of_mean) else math.sqrt(var_q / n)
    S2 = -math.log2(mean_q) if mean_q > 0 else math.nan
    return Stats(mean=S2, error_of_mean=err_q / (abs(mean_q) * math.log(2)),
                 variance=var_q / (mean_q * math.log(2)) ** 2, tau_corr=qs.tau_corr, R_hat=qs.R_hat)
