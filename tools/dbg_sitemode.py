import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import netket_fidelity_b200 as nkf
from netket_fidelity_b200 import nk
N = 20
hi = nk.hilbert.Spin(0.5, N)
for pd in (complex, "complex64"):
    for sm in ("shared", "per_chain"):
        sa = nk.sampler.MetropolisLocal(hi, n_chains=256, sweep_size=N, site_mode=sm)
        ma = nk.models.RBM(alpha=2, param_dtype=pd, stddev=0.2)
        for r in range(3):
            vs = nk.vqs.MCState(sa, ma, n_samples=256 * 64, n_discard_per_chain=5, seed=7, sampler_seed=500 + r)
            tg = nk.vqs.MCState(sa, ma, n_samples=256 * 64, n_discard_per_chain=5, seed=8, sampler_seed=900 + r)
            op = nkf.InfidelityOperator(tg, U=nkf.operator.Rx(hi, 3, 0.3), U_dagger=nkf.operator.Rx(hi, 3, -0.3),
                                        cv_coeff=-0.5, is_unitary=True)
            st = vs.expect(op)
            print(pd, sm, r, st.mean, st.variance, st.error_of_mean, vs.acceptance, float(abs(vs._flat).mean()), float(abs(tg._flat - vs._flat).mean()))
