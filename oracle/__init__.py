"""oracle/ -- CPU restatement of the reference's hot-path algorithms.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import
this package, and only as the checker (or as the timed CPU baseline) -- never as a product path.
The product (deep-gan-encoders_b200/) must not import it and has no CPU fallback.

Every function is a plain PyTorch fp32 (CPU-capable) restatement of one reference function and cites
the reference file:line it follows (paths relative to the reference repository root,
disanda/Deep-GAN-Encoders @ 0c2655f).  The reference ships no tests or golden vectors (SURVEY.md
section 4), so the oracle is pinned against outputs of the reference itself:

  * tests/golden/*.pt were generated HERE by importing the unmodified reference from
    /root/reference (tests/golden/make_golden.py, committed) on seeded random-init weights;
    tests/test_oracle_golden.py checks every oracle function against them (fp32, <= 2e-5 relative).
  * the SURVEY Appendix-C known-answer fingerprints are checked where the construction is reproducible.

Third-party arithmetic not in the reference tree: `lpips` (PyPI, unpinned) -- parity unpinned, see DESIGN.md.
"""
