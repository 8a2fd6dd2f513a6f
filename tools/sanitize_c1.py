"""The smallest configuration (BASELINE C1: README example) through the public API, a handful of optimisation steps:
the program compute-sanitizer runs (tools/final_measure.sh); also touches the lean sampler and the tcgen05 kernels once
at a reduced shape so that memcheck sees them."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import netket_fidelity_b200 as nkf  # noqa: E402
from netket_fidelity_b200 import nk  # noqa: E402

hi = nk.hilbert.Spin(0.5, 4)
sampler = nk.sampler.MetropolisLocal(hilbert=hi, n_chains_per_rank=16)
model = nk.models.RBM(alpha=1, param_dtype=complex, use_visible_bias=False)
phi = nk.vqs.MCState(sampler=sampler, model=model, n_samples=100, seed=1)
psi = nk.vqs.MCState(sampler=sampler, model=model, n_samples=100, seed=2)
U = nkf.operator.Hadamard(hi, 0)
te = nkf.driver.InfidelityOptimizer(phi, nk.optimizer.Adam(learning_rate=0.01), U=U, U_dagger=U, variational_state=psi,
                                    is_unitary=True, cv_coeff=-1 / 2)
te.run(n_iter=5, show_progress=False)
print("C1:", te.infidelity)

# north-star kernels at a reduced size: one wave of the lean sampler (4736 chains x 1 sample) + tcgen05 estimator / gradient
from netket_fidelity_b200 import _native as nv  # noqa: E402
from netket_fidelity_b200.engine import InfidelityStep, StepConfig  # noqa: E402
from netket_fidelity_b200.workloads import gate_spec, random_rbm  # noqa: E402

dev = torch.device("cuda", 0)
N, M = 100, 400
specs = [nv.RbmSpec(*[t.to(dev) for t in random_rbm(N, M, s, torch.complex64, True)]) for s in (1234, 5678)]
eng = InfidelityStep(N, M, StepConfig(4736, 1, 1, N, seed=3, cv_coeff=-0.5), "upsi", gate_spec("rx", 0, 0.1),
                     gate_spec("rx", 0, -0.1), device=dev)
sums, L, grad = eng.step(*specs)
torch.cuda.synchronize()
print("C4 kernels (reduced):", float(1 - sums[0] / sums[4]), float(grad.abs().max()))
