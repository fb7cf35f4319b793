"""CPU oracle for the cslam loop-closure hot path.

TEST INFRASTRUCTURE ONLY.  This package restates, in numpy/scipy, the
arithmetic of the reference (lajoiepy/cslam) functions that libcslam_b200
replaces, each function citing the reference file:line it follows.  Only
`tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline /
`--impl reference` legs may import it; the product package `cslam_b200`
never does (tests/test_abi.py::test_product_does_not_import_oracle checks).

Pinning: the reference ships no golden vectors for this path,
so every restatement is pinned against outputs of the reference's own Python
modules executed in the build container on seeded inputs; the generating
script is `oracle/make_golden.py` and the vectors live in `tests/golden/`.
"""
