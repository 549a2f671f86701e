"""CPU oracle for the v2 latent-DDPM sampling hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package may import this;
only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs use it, and only as the checker or the timed CPU baseline.

Parity pinning: the reference ships no tests, golden vectors or checkpoints
(SURVEY.md section 4), so the restatement in oracle/restate.py is pinned
against outputs of the reference itself, imported live in the build container
by oracle/ref_loader.py; the vectors are committed under tests/golden/ by
oracle/make_golden.py.
"""
