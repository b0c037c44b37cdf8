"""CPU oracle for the zk-fhe `prove` hot path.  TEST INFRASTRUCTURE ONLY.

This package is a plain-Python (and, under ``oracle/c``, plain-C) restatement
of the reference algorithms.  It is the *checker*: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import or execute it.  The product path
(``zk-fhe_b200/``) never imports it and fails loudly when the CUDA library is
missing.

Parity status:
  * stage (1) witness arithmetic  -- PINNED by /root/reference/data/bfv/bfv.in
    (c0, c1 are known answers) and by configs/bfv.json (column counts and
    break points); fixtures committed under tests/golden/.
  * BN254 field / curve constants -- PINNED by SURVEY.md App. A (recomputed).
  * MSM / NTT values              -- mathematically unique; checked against
    big-int definitions here.  The upstream crates (halo2-axiom, halo2curves,
    snark-verifier, poseidon; all un-vendored, versions unpinned) are absent,
    so everything past "advice table filled" is **parity unpinned**.
"""
