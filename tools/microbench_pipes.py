#!/usr/bin/env python3
"""Is there a second multiply pipe worth feeding?  Montgomery products per second on a full B200 with
  kind 0  the prover's product (IMAD.WIDE, the fmaheavy pipe),
  kind 6  the instruction mix of the same product on the FP64 pipe (52-bit limbs, fma_rz splitting; an experiment),
  kind 7  both at once: even warps kind 0, odd warps kind 6,
and point additions per second (points gathered from the window table) with
  kind 8  mixed XYZZ additions, one accumulator per thread (what k_msm_accumulate does),
  kind 9  affine additions in batches of 16 pairs per thread sharing one inversion (an experiment).
Usage: python tools/microbench_pipes.py   (needs a GPU)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_fhe_b200  # noqa: E402

ctx = zk_fhe_b200.Context(0)
ctx.srs_setup(13, zk_fhe_b200.reference_test_tau())          # kinds 8 / 9 gather from the k = 13 window table (10.5 MB, L2-resident)
for rep in range(2):
    for kind, name in ((8, "mixed XYZZ additions (10 products each)"), (9, "batched-affine additions, 16 pairs per inversion")):
        ms, ops = ctx.microbench(kind, 512)
        print(f"{name:52s} {ops / ms / 1e6:8.2f} G additions/s   ({ms:.2f} ms)")
    for kind, name in ((0, "IMAD product, every warp"), (6, "FP64 (DFMA) product mix, every warp"), (7, "even warps IMAD, odd warps DFMA")):
        ms, ops = ctx.microbench(kind, 2000)
        print(f"{name:40s} {ops / ms / 1e6:8.2f} G products/s   ({ms:.2f} ms)")
ctx.close()
