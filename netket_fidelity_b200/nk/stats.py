"""Stats (NetKet `nk.stats.Stats` / `nk.stats.statistics`): mean, variance over all
samples; error_of_mean from chain (batch) or block means; split-chain R_hat; tau_corr.
The O(S) reductions run on the GPU (nkfb_chain_stats); only O(rows) partial sums reach the host."""
import math

import numpy as np


class Stats:
    def __init__(self, mean=math.nan, error_of_mean=math.nan, variance=math.nan, tau_corr=math.nan,
                 R_hat=math.nan, tau_corr_max=math.nan):
        self.mean, self.error_of_mean, self.variance = mean, error_of_mean, variance
        self.tau_corr, self.R_hat, self.tau_corr_max = tau_corr, R_hat, tau_corr_max
        self.n_nonfinite = 0     # local estimators that were NaN / infinite (counted by the estimator kernel, sums[5])

    def replace(self, **kw):
        d = dict(mean=self.mean, error_of_mean=self.error_of_mean, variance=self.variance, tau_corr=self.tau_corr,
                 R_hat=self.R_hat, tau_corr_max=self.tau_corr_max)
        d.update(kw)
        out = Stats(**d)
        out.n_nonfinite = self.n_nonfinite
        return out

    def to_dict(self):
        return {"Mean": _f(self.mean), "Variance": _f(self.variance), "Sigma": _f(self.error_of_mean),
                "R_hat": _f(self.R_hat), "TauCorr": _f(self.tau_corr)}

    def __repr__(self):
        m = self.mean
        ms = f"{m.real:.4e}{m.imag:+.4e}j" if isinstance(m, complex) else f"{m:.4e}"
        return (f"{ms} ± {self.error_of_mean:.1e} [σ²={self.variance:.1e}, R̂={self.R_hat:.4f}, "
                f"τ={self.tau_corr:.1f}]")


def _f(v):
    if isinstance(v, complex):
        return {"real": v.real, "imag": v.imag}
    return float(v)


def stats_from_partials(part, rows, cols, l_block, n_blocks, sum_sq_centered, batch_size=32, world_sums=None):
    """part: (rows, 3 + n_blocks) doubles from nkfb_chain_stats over this rank's data viewed as
    (rows, cols); sum_sq_centered: sum (x - mean)^2 over all data.  Follows netket/stats/mc_stats.py
    `_statistics` (NetKet 3.10; parity unpinned -- the reference's tests read only mean, variance and
    error_of_mean)."""
    part = np.asarray(part)
    n = rows * cols
    mean = part[:, 0].sum() / n
    variance = sum_sq_centered / n
    row_means = part[:, 0] / cols
    n_batches = rows
    batch_var = np.mean((row_means - mean) ** 2)
    if n_blocks > 0:
        block_means = part[:, 3:].reshape(-1) / l_block
        n_blk = block_means.size
        block_var = np.mean((block_means - block_means.mean()) ** 2)
    else:
        n_blk, block_var = 0, math.nan
    with np.errstate(divide="ignore", invalid="ignore"):
        tau_batch = ((n / n_batches) * batch_var / variance - 1) * 0.5
        tau_block = ((n / n_blk) * block_var / variance - 1) * 0.5 if n_blk else math.nan
    batch_good = (tau_batch < 6 * cols) and (n_batches >= batch_size)
    block_good = n_blk and (tau_block < 6 * l_block) and (n_blk >= batch_size)
    if batch_good:
        err, tau = math.sqrt(batch_var / n_batches), max(0.0, tau_batch)
    elif block_good:
        err, tau = math.sqrt(block_var / n_blk), max(0.0, tau_block)
    else:
        err, tau = math.nan, math.nan
    R_hat = math.nan
    if n_batches > 1 and cols >= 2:
        half = cols // 2
        hm = np.concatenate([part[:, 1], part[:, 2]]) / half
        hv = np.mean((hm - hm.mean()) ** 2)
        W = variance
        if W > 0:
            R_hat = math.sqrt((cols - 1) / cols + hv / W)
    return Stats(mean=float(mean), error_of_mean=err, variance=float(variance), tau_corr=tau, R_hat=R_hat)
