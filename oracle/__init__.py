"""CPU oracle for the ESN hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in the product package (`xesn_b200/`) may import from here.  The only
legitimate users are `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline`
/ `--impl reference` legs of `bench.py`, and there only as the checker or as
the timed CPU baseline, never as the thing shipped.

Parity status: PINNED for the hot path (a1-a6); `nrmse_cost`, `psd_1d`, `psd_nrmse_fg` (the SURVEY section 8f cost terms) are UNPINNED: the
reference's versions need xarray, which this image lacks.

Hot path:  `oracle/make_golden.py` imports the reference's own
source files (through a stub shim for the missing xarray/dask packages) in the
build container, runs them on seeded inputs, and commits the input/output pairs
under `tests/golden/`.  `tests/test_oracle_golden.py` checks every function of
`oracle/esn_oracle.py` against those vectors.
"""
