"""CPU oracle for the CRAFT -> TrOCR hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this
package; the product (marie-icr_b200/) never does.  Every function cites the reference file:line it restates.

Pinning status:
  * CRAFT half (imgproc, CRAFT.forward, getDetBoxes, crop loop, line merge): PINNED — the restatements here are
    checked in tests/test_oracle_vs_reference.py against the reference's own modules imported by path from
    /root/reference (oracle/ref_loader.py), and golden vectors generated from those modules are committed
    under tests/golden/ (tools/make_golden.py).
  * TrOCR half (timm ViT + fairseq decoder + fairseq beam search): the arithmetic lives in fairseq (unpinned
    git HEAD) and timm==0.6.12, neither present in /root/reference nor installed; the reference holds no golden
    vectors for it.  The restatement is cross-checked against HuggingFace transformers' port of the same
    checkpoints (tests/test_oracle_trocr.py) — "parity unpinned" against the reference itself.
"""
