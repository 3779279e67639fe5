"""CPU oracle for the ITensorsGPU.jl hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT.

This package is a NumPy/SciPy (float64 / complex128, OpenBLAS + LAPACK) restatement of
the CPU algorithm the reference's GPU overrides are checked against: dense ITensor
contraction, truncation, svd/eigen/qr, ProjMPO (H_eff), Lanczos/Davidson, two-site
DMRG and gate application (TEBD).  Only ``tests/``, ``__graft_entry__.smoke()`` and
the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The
product (``itensorsgpu.jl_b200``) never imports it and has no CPU fallback.

Where the algorithm lives
-------------------------
The reference (``/root/reference``, 1,472 lines of Julia) only *overrides* primitives on
``CuDense`` storage; each function here cites the reference file:line whose semantics it
follows.  The algorithms that *sequence* those primitives (``dmrg``, ``ProjMPO``,
``factorize``, ``truncate!`` on CPU, ``apply``, KrylovKit's Lanczos) live in un-vendored
dependencies -- ITensors.jl 0.2 with its bundled NDTensors (``Project.toml:13,21``;
no Manifest, so no exact pin exists) and KrylovKit.jl (transitive).  Those are restated
from their published algorithms and marked ``[EXT]`` in the docstrings; parity is
anchored on the reference's own call sites and tests.

Pinning status
--------------
Julia is not installed in the build container, so the reference cannot be executed.
The oracle is pinned against every known-answer / assertion the reference's tests hold
for this path (see ``tests/test_oracle_pins.py``):

* ``test/test_cutruncate.jl:9-17``  three truncate! known-answer vectors (KAT1-3) --
  reproduced by ``truncate.truncate_gpu_reference`` (a transliteration of
  ``src/tensor/cutruncate.jl:1-93``); the CPU rule ``truncate.truncate`` agrees on
  KAT1-2 and deliberately differs on KAT3.
* ``test/dmrg.jl:28``  S=1 Heisenberg N=10 energy < -12.0, and additionally exact
  diagonalisation to 1e-8.
* ``test/dmrg.jl:79-80``  TFIM N=32 closed form within 1e-2.
* ``test/test_cuitensor.jl:105-112,125-130``  svd/qr reconstruction + isometry 1e-14.
* ``test/test_cuiterativesolvers.jl:22,26``  davidson eigen-residual.
* ``test/test_cumps.jl:208-226``  orthogonality to 1e-12.
* BASELINE.json configs[0] (ITensors.jl's stock ``examples/dmrg.jl`` schedule, N=100 S=1 chain): the converged
  energy that example publishes, -138.940086 (quoted from memory of the ITensors.jl README / SURVEY.md 8c: there
  is no network here to re-fetch it) -- the one absolute number from outside this repository that exercises
  the whole [EXT] restatement at once (MPO, environments, Lanczos, factorize rule, truncation, sweep order).
* ``tests/golden/hotpath_small.npz``: the outputs of this package on seeded inputs, one case per entry point, committed
  with their generator; ``tests/test_golden_cpu.py`` fails if a change here moves any of them.

The reference's contraction/permute/add tests are GPU-vs-CPU *relational* with no stored
arrays, and the [EXT] CPU path cannot be run here: **at the [EXT] boundary parity is
unpinned** (no outputs of the real ITensors.jl are available); it is constrained only
by the independent truths above (ED, closed form, algebraic invariants).
"""
from . import tensor, truncate, linalg, models, mps, dmrg, tebd  # noqa: F401
