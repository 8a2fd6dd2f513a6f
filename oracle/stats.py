"""Oracle: NetKet `nk.stats.statistics` (netket/stats/mc_stats.py, NetKet 3.10).

TEST INFRASTRUCTURE ONLY.  Parity unpinned for tau_corr / R_hat (the reference's tests read only
mean, variance, error_of_mean: test/test_infidelity_operator.py:117-125)."""
import numpy as np


def _var(x):
    return np.mean(np.abs(x - np.mean(x)) ** 2)


def statistics(data, batch_size=32):
    data = np.atleast_1d(np.asarray(data))
    if data.ndim == 1:
        data = data.reshape(1, -1)
    mean = data.mean()
    variance = _var(data)
    ts = data.size
    n_batches, n_s = data.shape
    batch_var = _var(data.mean(axis=1))
    l_block = max(1, n_s // batch_size)
    nb = n_s // l_block
    blocks = data[:, :nb * l_block].reshape(n_batches, nb, l_block).mean(axis=2).reshape(-1)
    block_var, n_blocks = _var(blocks), blocks.size
    with np.errstate(divide="ignore", invalid="ignore"):
        tau_batch = ((ts / n_batches) * batch_var / variance - 1) * 0.5
        tau_block = ((ts / n_blocks) * block_var / variance - 1) * 0.5
    batch_good = (tau_batch < 6 * n_s) and (n_batches >= batch_size)
    block_good = (tau_block < 6 * l_block) and (n_blocks >= batch_size)
    if batch_good:
        err, tau = np.sqrt(batch_var / n_batches), max(0.0, tau_batch)
    elif block_good:
        err, tau = np.sqrt(block_var / n_blocks), max(0.0, tau_block)
    else:
        err, tau = np.nan, np.nan
    R_hat = np.nan
    if n_batches > 1 and n_s >= 2:
        half = n_s // 2
        split = data[:, :2 * half].reshape(n_batches, 2, half).mean(axis=2)
        hm = np.concatenate([split[:, 0], split[:, 1]])
        R_hat = np.sqrt((n_s - 1) / n_s + _var(hm) / variance) if variance > 0 else np.nan
    return dict(mean=mean, variance=variance, error_of_mean=err, tau_corr=tau, R_hat=R_hat)
