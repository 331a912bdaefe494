"""CPU oracle for the bellman Groth16 prover hot path (TEST INFRASTRUCTURE ONLY).

This package is a plain restatement (Python big-ints + a C port under
``oracle/csrc/cref.cpp``) of the reference's CPU algorithms for
``bellman::multiexp``, ``bellman::domain::EvaluationDomain`` and
``groth16::create_proof``.  It exists to *check* the CUDA product path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs may import it.  Nothing under
``zcash-gpu-thesis_b200/`` imports, links or executes anything from here: the
product path fails loudly when ``libb200zk.so`` is missing.

Parity pinning: the oracle is checked (tests/test_oracle_*.py) against the
reference's own golden vectors -- field KATs (fr.rs:1240, fq.rs:2558 ...),
curve KATs (ec.rs:1060-1262), the four ``tests/*.dat`` byte-vector files
(sha256 + leading entries committed under tests/golden/) and the
``test_xordemo`` pipeline KAT (groth16/tests/mod.rs:98-400).
"""
