#!/usr/bin/env python3
"""Host-side timing of the phase-0 steps of one proof (upload, Poly::mul / reduce / divide_by_cyclo, from_poly), each
followed by a stream synchronise: where the 1.2 ms of `phase0 issue` goes.   python tools/phase0_profile.py"""
import os, sys, time
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
import zk_fhe_b200
from zk_fhe_b200 import bfv
from zk_fhe_b200.poly import Poly
from zk_fhe_b200.poly_chip import PolyChip, CTX_PHASE0
ctx = zk_fhe_b200.Context(0)
inp = bfv.load_input(os.path.join(os.environ.get("GRAFT_REPO_ROOT", "/root/repo"), "tests/golden/bfv.in"))
circ = bfv.BfvCircuit(ctx)
for it in range(3):
    circ.wit.reset(); ctx.sync()
    t = [time.perf_counter()]
    un = circ.upload(inp); ctx.sync(); t.append(time.perf_counter())
    P = {}
    for name, key in (("pk0","pk0"),("pk1","pk1"),("m","m"),("u","u"),("e0","e0"),("e1","e1"),("c0","c0"),("c1","c1"),("cyclo","cyclo")):
        P[name] = PolyChip.from_poly(un[key], circ.wit, CTX_PHASE0)
    ctx.sync(); t.append(time.perf_counter())
    a = un["pk0"].mul(un["u"]); b = un["pk1"].mul(un["u"]); ctx.sync(); t.append(time.perf_counter())
    ar = a.reduce_by_modulus(536870909); br = b.reduce_by_modulus(536870909); ctx.sync(); t.append(time.perf_counter())
    q0, r0 = ar.divide_by_cyclo(un["cyclo"], 536870909, check=False); q1, r1 = br.divide_by_cyclo(un["cyclo"], 536870909, check=False); ctx.sync(); t.append(time.perf_counter())
    q0c = q0.mul(un["cyclo"]); q1c = q1.mul(un["cyclo"]); ctx.sync(); t.append(time.perf_counter())
    for p in (a, b, q0, q1, q0c, q1c, r0, r1):
        PolyChip.from_poly(p, circ.wit, CTX_PHASE0)
    ctx.sync(); t.append(time.perf_counter())
    ctx.status(); t.append(time.perf_counter())
    names = ["upload9", "from_poly9", "mul2", "reduce2", "divide2", "mul2b", "from_poly8", "status"]
    print(it, " ".join(f"{n}={1e3*(t[i+1]-t[i]):.3f}" for i, n in enumerate(names)), f"total={1e3*(t[-1]-t[0]):.3f} ms")
