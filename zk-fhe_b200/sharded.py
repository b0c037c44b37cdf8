"""ONE proof over N GPUs: measurement / self-check helper shared by bench.py and tools/sharded_prove.py.

Every rank loads the same SRS and proving key, proves the same input with the same seed once on its own (the
single-GPU proof) and then, after binding an NCCL communicator to its context (dist.bind_sharded_prover), as one shard
of the sharded proof (csrc/comm.cu, csrc/prover.cu: commitment phases by column, quotient by coset).  Checks that every
rank's sharded proof equals the single-GPU bytes and reports the single-proof latency of both (host wall clock around
the whole prove, max over ranks) with the device time spent in the collectives.
"""
import time

import numpy as np

TAU = None          # run(): the trapdoor of the reference's own fallback SRS (capi.reference_test_tau)


def run(k, proofs, transcript, dist, rank, world, local_rank):
    import torch
    from . import capi
    torch.cuda.set_device(local_rank)          # the caller may be a worker thread: the CUDA device is per thread
    from . import bfv, bfv_py, prover
    from . import dist as zd
    if k == 13:
        params = bfv.BfvParams(N=1024, Q=536870909, T=7, B=19)
    elif k == 16:
        params = bfv.BfvParams(N=4096, Q=(1 << 61) - 1, T=65537, B=19)
    else:
        raise SystemExit("--k must be 13 (config 1) or 16 (the shape of configs 3/4)")
    global TAU
    TAU = capi.reference_test_tau()
    ctx = capi.Context(local_rank)
    ctx.set_blocking_sync(False)               # single-proof latency: spinning waits
    ctx.srs_setup(k, TAU)
    zeros = {key: ["0"] * (params.N + 1 if key == "cyclo" else params.N) for key in bfv.INPUT_KEYS}
    kg = bfv.BfvCircuit(ctx, params, record=True)
    kg.phase0(zeros).phase1(3)
    pk = prover.keygen(kg.wit, k, 109)
    del kg
    inp = bfv_py.keygen_and_encrypt(ctx, params, np.random.default_rng(1234), with_secret_key=False)   # same seed on every rank
    circ = bfv.BfvCircuit(ctx, params)
    pr = prover.Prover(pk, bytes(32), transcript)

    rounds = []

    def prove_once(seed):
        circ.wit.reset()
        circ.phase0(inp)
        pr.reset(seed)
        gamma = pr.phase0(circ.wit)
        circ.phase1(gamma)
        return pr.finish(circ.wit)

    def timed(n):
        out, ms = None, []
        rounds.clear()
        for i in range(n):
            ctx.sync()
            if world > 1:
                dist.barrier()
            t0 = time.perf_counter()
            out = prove_once(bytes(32))
            ctx.sync()
            ms.append(zd.max_over_ranks(1e3 * (time.perf_counter() - t0), device=torch.device("cuda", local_rank) if world > 1 else "cpu"))
            rounds.append(pr.round_ms())
        return out, ms

    def median_rounds():
        return {key: round(float(np.median([r[key] for r in rounds])), 3) for key in rounds[0]}

    prove_once(bytes(32))                                   # warm-up: one-time allocations
    single, ms_single = timed(proofs)
    rounds_single = median_rounds()
    cats = {}
    if world > 1:
        zd.bind_sharded_prover(ctx, device=torch.device("cuda", local_rank))
    prove_once(bytes(32))                                   # warm-up: NCCL channels, new workspaces
    ctx.timing_reset()
    sharded, ms_sharded = timed(proofs)
    names = {0: "msm_accumulate", 1: "ntt", 2: "msm_sort", 3: "msm_fold", 4: "msm_final", 7: "collectives"}
    cats = {names[c]: round(ctx.timing(c)[0] / proofs, 3) for c in names}
    comm_calls = ctx.timing(7)[1] // max(proofs, 1)
    rounds_sharded = median_rounds()
    same = torch.tensor([1 if sharded == single else 0], device=torch.device("cuda", local_rank))
    if world > 1:
        dist.all_reduce(same, op=dist.ReduceOp.MIN)
    res = {"k": k, "N": params.N, "n_ranks": world, "transcript": "poseidon" if transcript == 1 else "blake2b",
           "identical_to_single_gpu": bool(same.item()), "proof_bytes": len(sharded),
           "single_gpu_latency_ms": round(float(np.median(ms_single)), 3), "sharded_latency_ms": round(float(np.median(ms_sharded)), 3),
           "speedup": round(float(np.median(ms_single)) / float(np.median(ms_sharded)), 3),
           "rank0_round_ms_single": rounds_single, "rank0_round_ms_sharded": rounds_sharded,
           "rank0_device_ms_per_sharded_proof": cats, "collectives_per_proof": int(comm_calls),
           "limiters": "replicated on every rank: stage (1) witness kernels, column fills, lookup permutations, the chain of the "
                       "permutation products, the extended iNTT of the quotient, the six coset NTTs and two quotients of SHPLONK, "
                       "the host transcript (sequential sponge), and the latency of the four 1-3 column commits; sharded: "
                       "grand products and MSMs by column, lagrange->coeff->extended NTTs and the quotient identities by "
                       "expression, evaluations and opening sums by column"}
    ctx.comm_destroy()
    ctx.close()
    return res
