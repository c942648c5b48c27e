"""CPU oracle for the EmailVerifier witness + Groth16 path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` may be imported, linked or
executed by the product path (``zk-email-verify_b200/``); only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs use it, and only as the checker / the timed CPU baseline.

Parity status (see DESIGN.md "Oracle"):
  * Groth16 *verifier*  : PINNED  - accepts the reference's only Groth16 known-answer fixture
                          (/root/reference/packages/rust-verifier/tests/data/proof_of_twitter/*,
                          asserted true at rust-verifier/tests/verifier_utils.rs:11-18); copies of
                          the three JSON files live in tests/golden/proof_of_twitter/.
  * circuit templates    : PINNED by the reference's circuit unit-test vectors
                          (packages/circuits/tests/*.test.ts, see tests/test_templates_*.py).
  * Groth16 *prover*     : parity unpinned - the reference holds no proof-producing test, no zkey /
                          r1cs / wtns fixture and snarkjs draws random (r, s); the prover oracle is a
                          restatement of the snarkjs 0.5.0 algorithm (un-vendored dependency,
                          packages/helpers/package.json:26) anchored on the pinned verifier.
"""
