"""CPU oracle for the BarBay ADVI hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package (``barbay.jl_b200``)
may import, call, link or execute anything under ``oracle/``.  The only allowed
callers are ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py``.

PARITY UNPINNED: the reference (mrazomej/BarBay.jl) is pure Julia on top of
un-vendored third-party packages (Turing 0.36, AdvancedVI 0.2, DynamicPPL 0.32,
Bijectors 0.15, Distributions 0.25 -- Project.toml:26-43, compat ranges only, no
Manifest) and Julia is not installed in this image, so the reference cannot be
executed here.  Its own tests (test/vi_tests.jl:23-236) hold no golden vectors
for the log-joint, the gradient, the ELBO or the posterior.  The oracle is
therefore a line-by-line restatement of ``src/model_*.jl`` plus the published
AdvancedVI 0.2 semantics, cross-checked by independent means (scipy.stats
densities, autograd vs analytic gradient, finite differences, simulator
ground-truth recovery) but not pinned against reference outputs.
"""
