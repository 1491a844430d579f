"""CPU oracle for the dpgo RBCD hot path -- TEST INFRASTRUCTURE ONLY.

This package restates, in numpy/scipy, the algorithm of the reference's per-agent
Riemannian block-coordinate-descent local solve (QuadraticProblem / QuadraticOptimizer /
PoseGraph data matrices / the PGOAgent iterate loop).  It is the checker the CUDA path is
compared with.  Only `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` may import it; the product (`dpgo_b200/`) never does.

Parity status
-------------
* Data matrices (Q, G, preconditioner matrix), cost, Euclidean gradient, Hessian-vector
  product, the g2o parser and chordal initialization follow reference source that IS present
  under /root/reference (citations on each function).  They are pinned by the reference's own
  known-answer tests (triangle graph `tests/testTriangleGraph.cpp`, prior test
  `tests/testPGO.cpp:131-190`) -- see tests/test_oracle.py.
* The Riemannian trust-region / truncated-CG control flow, the Stiefel tangent projection, the
  Riemannian Hessian correction and the QF retraction live in ROPTLIB
  (github.com/yuluntian/ROPTLIB, branch feature/cmake, NO commit pin, fetched at configure
  time by cmake/roptlib.cmake:7-8) which is absent from /root/reference.  They are restated
  from the published algorithm (Absil, Baker, Gallivan, "Trust-region methods on Riemannian
  manifolds", 2007; Absil/Mahony/Sepulchre, "Optimization Algorithms on Matrix Manifolds",
  Alg. 10/11) with ROPTLIB's default constants as used at the reference call sites
  (src/QuadraticOptimizer.cpp:61-100).  Per-iteration traces (f, |grad|, tCG counts, radii)
  are therefore **parity unpinned**; converged solutions are pinned by the tests above.
"""
