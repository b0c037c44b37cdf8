#!/usr/bin/env python3
"""Is there a second multiply pipe worth feeding?  Montgomery products per second on a full B200 with
  kind 0  the prover's product (IMAD.WIDE, the fmaheavy pipe),
  kind 6  the instruction mix of the same product on the FP64 pipe (52-bit limbs, fma_rz splitting; an experiment),
  kind 7  both at once: even warps kind 0, odd warps kind 6.
Usage: python tools/microbench_pipes.py   (needs a GPU)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import zk_fhe_b200  # noqa: E402

ctx = zk_fhe_b200.Context(0)
for rep in range(2):
    for kind, name in ((0, "IMAD product, every warp"), (6, "FP64 (DFMA) product mix, every warp"), (7, "even warps IMAD, odd warps DFMA")):
        ms, ops = ctx.microbench(kind, 2000)
        print(f"{name:40s} {ops / ms / 1e6:8.2f} G products/s   ({ms:.2f} ms)")
ctx.close()
