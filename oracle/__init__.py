"""CPU oracle for the infidelity-gradient hot path of netket_fidelity.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`netket_fidelity_b200/`) may import, call or link anything in here; only
`tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` / `--impl reference`
legs of `bench.py` do, and only as the checker / reported baseline.

The oracle is a from-scratch NumPy complex128 restatement of the mathematics of

* `netket_fidelity/operator/singlequbit_gates.py:89-105,184-201,270-288`
  (connected elements of Rx / Ry / Hadamard),
* `netket_fidelity/utils/sampling_Ustate.py:34-47` (log(U phi) via logsumexp),
* `netket_fidelity/infidelity/overlap/expect.py:58-130` and
  `netket_fidelity/infidelity/overlap_U/expect.py:70-161` (local estimator,
  control variates, score-function gradient),
* `netket_fidelity/infidelity/overlap/exact.py:53-78`,
  `netket_fidelity/infidelity/overlap_U/exact.py:62-89`,
  `test/_infidelity_exact.py:4-20`, `test/_finite_diff.py:8-31` (ground truth),

plus the NetKet 3.10 semantics those files delegate to (RBM, log_cosh,
MetropolisLocal, nkjax.expect / vjp, IsingJax, Renyi2EntanglementEntropy,
Stats).  NetKet/JAX are NOT installable in this image (no wheels, no network),
so the NetKet-side pieces are restated from the published algorithm.

PINNING STATUS
--------------
* gates: pinned by the reference's own golden vectors
  (`test/test_operator.py:66-96`, `test/test_inverserot.py:15-25`) ->
  tests/test_oracle_gates.py.
* estimator + gradient: pinned by the reference's own test method
  (`test/test_infidelity_operator.py:95-127`): exact full-summation infidelity
  and its central-difference gradient, and additionally by an independent
  torch-autograd restatement of `nkjax.expect` + `nkjax.vjp(conjugate=True)`.
* IsingJax connected elements, Renyi2, Stats(R_hat, tau_corr): **parity
  unpinned** (the reference's tests hold no vector for them); they are
  validated only through physics invariants (dense-matrix reconstruction,
  explicit reduced density matrix).
"""
