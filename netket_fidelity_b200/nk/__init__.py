"""Minimal stand-ins for the NetKet objects that the reference's public surface is written
against (hilbert spaces, graphs, RBM, MetropolisLocal, MCState, Stats, optimisers, callbacks,
loggers).  Only what the infidelity hot path needs; same names and argument meaning as
NetKet 3.10 so that reference scripts port by replacing `import netket as nk` with
`from netket_fidelity_b200 import nk`."""
from . import hilbert, graph, models, sampler, stats, optimizer, callbacks, logging, vqs, operator  # noqa: F401
