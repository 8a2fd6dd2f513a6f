"""MetropolisLocal (NetKet `nk.sampler.MetropolisLocal`): single-spin-flip Metropolis,
`sweep_size` steps per emitted sample.  The transition kernel runs on the GPU
(csrc/sampler.cu); this object only carries the configuration."""


class MetropolisLocal:
    def __init__(self, hilbert, n_chains=None, n_chains_per_rank=None, sweep_size=None, reset_chains=False,
                 machine_pow=2, dtype=float, site_mode="shared"):
        """site_mode: "shared" -- all chains of a call propose the same (random) site sequence, which lets
        a warp load one row of W for all its chains (each chain is still an exact random-scan
        Metropolis chain); "per_chain" -- every chain draws its own site, as NetKet does."""
        import torch.distributed as dist
        self.hilbert = hilbert
        world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if n_chains is not None and n_chains_per_rank is not None:
            raise ValueError("Cannot specify both `n_chains` and `n_chains_per_rank`")
        if n_chains is not None:
            n_chains_per_rank = max(1, -(-int(n_chains) // world))
        if n_chains_per_rank is None:
            n_chains_per_rank = 16
        self.n_chains_per_rank = int(n_chains_per_rank)
        self.n_chains = self.n_chains_per_rank * world
        self.sweep_size = int(sweep_size) if sweep_size is not None else hilbert.size
        self.reset_chains = reset_chains
        if machine_pow != 2:
            raise NotImplementedError("machine_pow != 2")
        self.machine_pow = 2
        self.dtype = dtype
        if site_mode not in ("shared", "per_chain"):
            raise ValueError(site_mode)
        self.site_mode = site_mode

    def __repr__(self):
        return f"MetropolisLocal(n_chains={self.n_chains}, sweep_size={self.sweep_size}, sites={self.site_mode})"
