#!/usr/bin/env python3
"""bench.py -- BFV `prove` at config 1 (N=1024, Q=536870909, T=7, B=19, k=13), full proofs.

  python bench.py --gpus N --steps K --warmup W          (N>1: launched under torchrun)
  python bench.py --impl reference ...                   (CPU arm: whole CPU passes of the oracle's C restatement)

A step is ONE complete proof of the BFV encryption circuit in the reference's shape
(configs/bfv.json: 3+153 gate, 5 RLC, 36 lookup advice columns, 5121 instances): witness
generation (stage 1), ~411 KZG commitments (stage 2), the coset NTTs / quotient (stage 3),
evaluations and the SHPLONK opening -- everything `prove` does after the input is parsed
and the proving key is loaded, which is what the reference's "Proving time" brackets.
  value  proofs/s with the nine input polynomials already resident in HBM
  e2e    proofs/s from host decimal strings (the bfv.in format) to proof bytes on the host
Multi-GPU: proofs are independent units, so each rank proves its own witnesses (weak
scaling, no data-path collective); value is the whole-job proofs/s.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

K = 13
N_ROWS = 1 << K
K_EXT = 15
N_POLY, Q_MOD, T_MOD, B_ERR = 1024, 536870909, 7, 19
TAU = None          # set in main(): the trapdoor of the reference's own fallback SRS (zkfhe_reference_test_tau), a test setup
# column counts of one proof (reference circuit shape; permutation chunks of 2 columns at degree 4)
C_ADVICE = 3 + 153 + 5 + 36
C_LOOKUP_PERM = 2 * 36
C_PERM_Z = 100
C_LOOKUP_Z = 36
C_MSM = C_ADVICE + C_LOOKUP_PERM + C_PERM_Z + C_LOOKUP_Z + 1 + 3 + 2      # 411
C_NTT = C_ADVICE + C_LOOKUP_PERM + C_PERM_Z + C_LOOKUP_Z + 1              # 406
MSM_BYTES_PER_PAIR = 96          # 32 B scalar + 64 B affine base
NTT_BYTES_PER_ELEM = 64          # 32 B read + 32 B write


def _traffic():
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this
    command (profiles/r02_traffic.json, written by tools/ncu_traffic.py; the round-1 file as a fallback); {} if absent."""
    for name in ("r02_traffic.json", "r01_traffic.json"):
        p = os.path.join(ROOT, "profiles", name)
        if os.path.exists(p):
            return json.load(open(p))
    return {}


TRAFFIC = _traffic()


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)).get("hbm_gbs", 6650.0), "measured"
    return 6650.0, "fallback"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-i", str(self.idx), "-lms", "100"], stdout=self.f, stderr=subprocess.DEVNULL)
        except FileNotFoundError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        rows = [ln.split(", ") for ln in open(self.f.name).read().strip().splitlines() if ln.strip()]
        os.unlink(self.f.name)
        sm = sorted(float(r[1]) for r in rows if len(r) >= 9)
        reasons = set()
        for r in rows:
            if len(r) < 9:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.strip() == "Active":
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None,
                "sm_max_mhz": float(rows[0][2]) if rows and len(rows[0]) >= 3 else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def host_threads():
    """Host threads the CPU arm may use: the cores this process is allowed to run on.  NOT omp_get_max_threads():
    torchrun exports OMP_NUM_THREADS=1 to its workers, which would pin the reference arm to one core."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def cpu_synth_input(rng):
    """(CPU arm) one synthetic BFV encryption as a bfv.in dict, SURVEY.md section 8(d): pk0, pk1 uniform, u uniform on
    {0, 1, Q-1}, e0 / e1 rounded N(0, 3.2^2) clipped to +-B, m uniform on [-T/2, T/2]; c0, c1 by the oracle's C
    restatement of the ring arithmetic (schoolbook product, reduction mod x^N + 1 and mod Q)."""
    from oracle import cbind
    N, Q, T, B = N_POLY, Q_MOD, T_MOD, B_ERR
    uni = lambda: [int(x) for x in rng.integers(0, Q, N, dtype=np.uint64)]
    u = [Q - 1 if x == 2 else int(x) for x in rng.integers(0, 3, N)]
    err = lambda: [int(x) % Q for x in np.clip(np.rint(rng.normal(0, 3.2, N)), -B, B).astype(np.int64)]
    m = [int(x) % Q for x in rng.integers(-(T // 2), T // 2 + 1, N)]
    pk0, pk1, e0, e1 = uni(), uni(), err(), err()

    def ring_mul(a, b):                       # a * b mod (x^N + 1, Q); big-endian coefficient order
        p = cbind.poly_reduce(cbind.poly_mul(a, b), Q)          # 2N - 1 coefficients, index 0 = x^(2N-2)
        return [(p[N - 1 + i] - (p[i - 1] if i >= 1 else 0)) % Q for i in range(N)]

    delta = Q // T
    c0 = [(x + delta * mm + e) % Q for x, mm, e in zip(ring_mul(pk0, u), m, e0)]
    c1 = [(x + e) % Q for x, e in zip(ring_mul(pk1, u), e1)]
    return {"pk0": pk0, "pk1": pk1, "m": m, "u": u, "e0": e0, "e1": e1, "c0": c0, "c1": c1, "cyclo": [1] + [0] * (N - 1) + [1]}


class CpuProver:
    """The reference's CPU work for one proof, restated in C (oracle/c) and run on the host cores: a MEASURED whole pass,
    not an extrapolated sample.  Per step:
      stage (1)  schoolbook Poly::mul (src/poly.rs:86-90), literal long division (:133-142), reduce_by_modulus, and
                 every cell of the halo2-base gates behind src/poly_chip.rs (1,288,314 advice cells + 286,756 lookup
                 cells at config 1), single-threaded as in the reference;
      stage (2)  all 411 commitments with halo2's best_multiexp shape (columns one after the other, points split
                 over the threads, serial Pippenger per chunk);
      stage (3)  406 lagrange_to_coeff iNTTs (2^13), 407 coeff_to_extended coset NTTs (2^15, the instance column
                 included), 1 extended_to_coeff iNTT (2^15), halo2's best_fft shape.
    The 197 advice and 72 permuted-lookup columns are the real witness of the step's input, cut at the reference's
    break points (configs/bfv.json); the 142 grand-product / random / quotient / opening columns are uniform field
    elements (which is what they are in a real proof).  NOT included: grand products, quotient evaluation, the ~900
    evaluations at x, SHPLONK linear combinations, the transcript -- so `value` is an upper bound on CPU proofs/s."""

    def __init__(self, threads):
        from oracle import cbind
        self.cb = cbind
        self.threads = threads
        self.n = N_ROWS
        self.g, self.gl = cbind.srs(K, 0x5EED5EED5EED, threads=threads)
        pin = json.load(open(os.path.join(ROOT, "tests", "golden", "bfv_pinning.json")))
        self.breaks = [pin["break_points"]["gate"][0], pin["break_points"]["gate"][1], pin["break_points"]["rlc"]]
        self.unusable = pin["params"]["unusable_rows"]
        self.max_rows = self.n - self.unusable
        self.usable = self.n - 6 - 1                    # halo2 blinding_factors() = 6 for this constraint system
        self.gamma = 0x0123456789ABCDEF0123456789ABCDEF0123456789ABCDEF0123456789ABCDEF % ((1 << 254) - 1)

    def _cut(self, flat, breaks):
        """halo2-base assign_all in witness-gen mode: a column ends at its break point; the break cell is repeated at
        row 0 of the next column."""
        cols, start = [], 0
        for bp in breaks:
            cols.append(flat[start:start + bp + 1])
            start += bp
        cols.append(flat[start:])
        return cols

    def _lookup_permute(self, col_small):
        """halo2 permute_expression_pair for table {0..255} (+ zeros): A' sorted; S' = A' where a run starts, the
        other rows take the unused table values in ascending order."""
        u = self.usable
        a = np.sort(col_small[:u])
        first = np.ones(u, bool)
        first[1:] = a[1:] != a[:-1]
        used = np.zeros(256, bool)
        used[a[first]] = True
        zeros_left = (u - 255) - (1 if used[0] else 0)
        fill = np.concatenate([np.zeros(max(zeros_left, 0), np.uint64), np.nonzero(~used)[0].astype(np.uint64)[1 if not used[0] else 0:]])
        s = a.copy()
        s[~first] = fill[:int((~first).sum())]
        return a, s

    def step(self, inp, rng):
        cb, n, th = self.cb, self.n, self.threads
        t = {}
        t0 = time.perf_counter()
        adv0, adv1, adv2, lk = cb.bfv_witness(inp, N_POLY, Q_MOD, T_MOD, B_ERR, self.gamma)
        t["stage1"] = time.perf_counter() - t0
        # ---- assemble the Lagrange columns (part of create_proof's synthesis; timed under "assemble") ----
        t0 = time.perf_counter()
        cols = self._cut(adv0, self.breaks[0]) + self._cut(adv1, self.breaks[1]) + self._cut(adv2, self.breaks[2])
        cols += [lk[i:i + self.max_rows] for i in range(0, lk.shape[0], self.max_rows)]
        assert len(cols) == C_ADVICE, len(cols)
        P = np.zeros((C_NTT, n, 4), np.uint64)
        blind = rng.integers(0, 1 << 62, size=(C_NTT, n - self.usable, 4), dtype=np.uint64)
        for j, c in enumerate(cols):
            P[j, :c.shape[0]] = c
        P[:, self.usable:] = blind
        lk_small = cb.from_mont_array(np.ascontiguousarray(lk))[:, 0]
        perm = np.zeros((C_LOOKUP_PERM, n, 4), np.uint64)
        for l in range(C_LOOKUP_PERM // 2):
            col = np.zeros(n, np.uint64)
            seg = lk_small[l * self.max_rows:(l + 1) * self.max_rows]
            col[:seg.shape[0]] = seg
            a, s = self._lookup_permute(col)
            perm[2 * l, :self.usable, 0] = a
            perm[2 * l + 1, :self.usable, 0] = s
        flat = perm.reshape(-1, 4)
        cb.lib().orc_to_mont_array(0, flat.ctypes.data, flat.ctypes.data, flat.shape[0])
        P[C_ADVICE:C_ADVICE + C_LOOKUP_PERM] = perm
        P[C_ADVICE:C_ADVICE + C_LOOKUP_PERM, self.usable:] = blind[C_ADVICE:C_ADVICE + C_LOOKUP_PERM]
        full = rng.integers(0, 1 << 62, size=(C_NTT - C_ADVICE - C_LOOKUP_PERM, n, 4), dtype=np.uint64)
        P[C_ADVICE + C_LOOKUP_PERM:] = full             # Z_perm (100), Z_lookup (36), R: uniform field elements
        inst = np.zeros((n, 4), np.uint64)
        inst[:5121] = adv0[:5121]                       # stand-in with the instance column's value mix
        t["assemble"] = time.perf_counter() - t0
        # ---- stage (2): commitments, phase by phase as create_proof makes them ----
        t0 = time.perf_counter()
        flatP = P.reshape(-1, 4)
        for lo, hi in ((0, 3), (3, C_ADVICE), (C_ADVICE, C_ADVICE + C_LOOKUP_PERM), (C_ADVICE + C_LOOKUP_PERM, C_NTT)):
            cb.msm(flatP[lo * n:hi * n], self.gl, n, hi - lo, threads=th)
        t["msm_lagrange"] = time.perf_counter() - t0
        # ---- stage (3): lagrange_to_coeff, coeff_to_extended ----
        t0 = time.perf_counter()
        cb.ntt(flatP, K, C_NTT, inverse=True, threads=th)
        cb.ntt(inst, K, 1, inverse=True, threads=th)
        t["intt"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        n4 = 1 << K_EXT
        ext = np.zeros((8 * n4, 4), np.uint64)
        for lo in range(0, C_NTT + 1, 8):
            hi = min(lo + 8, C_NTT + 1)
            ext[:] = 0
            for j in range(lo, hi):
                ext[(j - lo) * n4:(j - lo) * n4 + n] = P[j] if j < C_NTT else inst
            cb.ntt(ext[:(hi - lo) * n4], K_EXT, hi - lo, coset=True, threads=th)
        t["coset_ntt"] = time.perf_counter() - t0
        # ---- quotient: extended_to_coeff, 3 pieces committed; SHPLONK: 2 commitments (coefficient basis) ----
        t0 = time.perf_counter()
        h = rng.integers(0, 1 << 62, size=(n4, 4), dtype=np.uint64)
        cb.ntt(h, K_EXT, 1, inverse=True, coset=True, threads=th)
        t["ext_intt"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        cb.msm(np.ascontiguousarray(h[:3 * n]), self.g, n, 3, threads=th)
        cb.msm(np.ascontiguousarray(full[:2].reshape(-1, 4)), self.g, n, 2, threads=th)
        t["msm_coeff"] = time.perf_counter() - t0
        return t


def cpu_reference_arm(steps, warmup, threads=0, budget_s=240.0):
    """`steps` timed whole CPU passes after `warmup` untimed ones (see CpuProver); if the projected run exceeds
    `budget_s` the counts are cut and the TRUE counts are what is returned and printed."""
    threads = threads or host_threads()
    rng = np.random.default_rng(1)
    cp = CpuProver(threads)
    inputs = [cpu_synth_input(rng) for _ in range(2)]
    t0 = time.perf_counter()
    first = cp.step(inputs[0], rng)                  # counts as the first warm-up pass
    one = time.perf_counter() - t0
    warm_done = 1
    if one * (steps + warmup) > budget_s:
        steps = max(1, int(budget_s / one) - 1)
        warmup = 1
    while warm_done < warmup:
        cp.step(inputs[warm_done % 2], rng)
        warm_done += 1
    times, parts = [], []
    for it in range(steps):
        t0 = time.perf_counter()
        parts.append(cp.step(inputs[it % 2], rng))
        times.append(time.perf_counter() - t0)
    sec = float(np.mean(times))
    split = {k: round(float(np.mean([p[k] for p in parts])), 4) for k in first}
    return {"value": 1.0 / sec, "unit": "proofs/s", "cores": int(threads), "kind": "port",
            "sample": f"{steps} whole CPU passes (after {warm_done} warm-up), each ONE proof (a bounded sample of the GPU arm's step, "
                      f"which is a batch of proofs; the unit, proofs/s, is the same) = stage (1) in C (schoolbook products, long division, "
                      f"all {23558 + 1231992 + 32764} advice + 286756 lookup cells) + all {C_MSM} MSMs + {C_NTT} iNTT(2^13) + "
                      f"{C_NTT + 1} coset NTT(2^15) + 1 extended iNTT, on the step's real advice / permuted-lookup columns (the 142 "
                      f"grand-product / random / quotient / opening columns are uniform field elements, as in a real proof); grand "
                      f"products, quotient evaluation, evaluations, transcript excluded, so an upper bound on CPU proofs/s; C "
                      f"restatement (oracle/c) of src/poly.rs, src/poly_chip.rs over halo2-base, halo2 best_multiexp / best_fft -- "
                      f"not the reference binary (no Rust toolchain); the reference README quotes 10.2 s per proof on an 8-core M2",
            "sec_per_proof": sec, "steps_run": int(steps), "warmup_run": int(warm_done), "seconds_by_part": split}


def synth_inputs(ctx, rng, count):
    """SURVEY.md §8(d) synthetic BFV witnesses for config 1: u uniform on {0,1,Q-1}, e0/e1 rounded
    N(0,3.2^2) clipped to +-B, m uniform on [-T/2,T/2], pk0/pk1 uniform; c0, c1 computed with the
    library's own device-resident Poly arithmetic (the package's witness front-end, bfv_py)."""
    from zk_fhe_b200 import bfv, bfv_py
    params = bfv.BfvParams(N=N_POLY, Q=Q_MOD, T=T_MOD, B=B_ERR)
    return [bfv_py.keygen_and_encrypt(ctx, params, rng, with_secret_key=False) for _ in range(count)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--transcript", type=int, default=1, help="1 Poseidon (default: the reference's transcript), 0 BLAKE2b")
    ap.add_argument("--spin", action="store_true", help="spinning host waits (cudaStreamSynchronize) instead of blocking events")
    ap.add_argument("--no-sharded", action="store_true", help="N > 1: skip the one-proof-over-N-GPUs latency measurement")
    ap.add_argument("--streams", type=int, default=16, help="proofs in flight per GPU (one CUDA stream + host thread each)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    config = {"workload": "bfv_prove_N1024_Q29bit_T7_B19_k13", "N": N_POLY, "Q": Q_MOD, "k": K, "advice_columns": C_ADVICE,
              "instances": 5121, "commitments_per_proof": C_MSM, "ext_k": K_EXT,
              "transcript": "blake2b" if args.transcript == 0 else "poseidon",
              "proofs_per_step": max(1, args.streams),
              "step": f"one batch of {max(1, args.streams)} proofs per GPU (one synthetic witness per proof stream); the timed region is "
                      f"steps x {max(1, args.streams)} proofs per GPU, handed to the streams as they become free",
              "parallelism": f"{args.streams} proofs in flight per GPU (one stream + host thread each) x {max(world, args.gpus)} GPU(s); "
                             f"proofs are independent units, no data-path collective",
              "cache": "per-proof working set (prover polynomials 104 MB + extended 416 MB + fixed extended 383 MB + MSM "
                       "workspace) exceeds the 126 MB L2; 4 distinct witnesses rotate between steps",
              "reference_readme": "10.2 s per proof, M2 MacBook Air 8 cores (README.md:58)"}

    if args.impl == "reference":
        if rank != 0:
            return 0
        cb = cpu_reference_arm(max(1, args.steps), max(1, args.warmup))
        line = {"impl": "reference", "metric": "bfv_prove_proofs_per_s", "value": cb["value"], "unit": "proofs/s",
                "n_gpus": 0, "steps": cb["steps_run"], "warmup": cb["warmup_run"], "ms_per_step": 1e3 * cb["sec_per_proof"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u256 (BN254 Fr/Fq, 4xu64 Montgomery)",
                "data": "synthetic", "config": config, "cpu_baseline": cb,
                "e2e": {"value": cb["value"], "unit": "proofs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return 0

    import torch
    import torch.distributed as dist

    import zk_fhe_b200
    from zk_fhe_b200 import bfv, prover
    global TAU
    TAU = zk_fhe_b200.reference_test_tau()

    if world > 1:
        # NCCL prints its version banner on stdout at communicator creation; the contract is ONE JSON
        # line on stdout, so stdout points at stderr until the first collective has run
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            torch.cuda.set_device(local_rank)
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    ctx = zk_fhe_b200.Context(local_rank)
    # every proof stream keeps its context's own non-blocking CUDA stream; torch's stream only carries
    # the two timing events, which are recorded after a device-wide synchronise on both sides
    stream = torch.cuda.current_stream()

    # ---- setup (untimed): SRS, keygen on the all-zero input, synthetic witnesses ---------------------
    ctx.srs_setup(K, TAU)
    params = bfv.BfvParams(N=N_POLY, Q=Q_MOD, T=T_MOD, B=B_ERR)
    zeros = {key: ["0"] * (N_POLY + 1 if key == "cyclo" else N_POLY) for key in bfv.INPUT_KEYS}
    kg = bfv.BfvCircuit(ctx, params, record=True)
    kg.phase0(zeros).phase1(1)
    pk = prover.keygen(kg.wit, K, 109)
    del kg
    rng = np.random.default_rng(20261017 + rank)
    inputs = synth_inputs(ctx, rng, 4)
    import threading

    class ProofStream:
        """One in-flight proof: its own context (CUDA stream), witness buffers and prover buffers;
        the proving key and the commitment-key tables are shared."""

        def __init__(self, index):
            self.ctx = ctx if index == 0 else zk_fhe_b200.Context(local_rank)
            if index:
                self.ctx.share_srs(ctx)
            self.ctx.set_blocking_sync(not args.spin)
            self.circ = bfv.BfvCircuit(self.ctx, params)
            self.resident = [self.circ.upload(inp) for inp in inputs]
            self.pr = prover.Prover(pk, bytes(32), args.transcript, ctx=self.ctx)
            self.ctx.sync()
            self.proof_len = 0

        def prove(self, i, from_host):
            self.circ.wit.reset()
            if from_host:
                self.circ.phase0(inputs[i % len(inputs)])
            else:
                self.circ.phase0(None, resident=self.resident[i % len(self.resident)])
            self.pr.reset(i.to_bytes(32, "little"))
            gamma = self.pr.phase0(self.circ.wit)
            self.circ.phase1(gamma)
            self.proof_len = len(self.pr.finish(self.circ.wit))

    streams = [ProofStream(i) for i in range(max(1, args.streams))]
    counter = [0]
    lock = threading.Lock()

    def run_steps(steps, from_host, workers):
        """Exactly `steps` proofs, handed out dynamically to the proof streams."""
        steps = int(steps)
        start = counter[0]
        counter[0] += steps
        nxt = [start]
        errors = []

        def work(ps):
            try:
                while True:
                    with lock:
                        i = nxt[0]
                        if i >= start + steps:
                            return
                        nxt[0] += 1
                    ps.prove(i, from_host)
            except Exception as e:       # surface worker failures in the main thread
                errors.append(e)

        if len(workers) == 1:
            work(workers[0])
        else:
            th = [threading.Thread(target=work, args=(ps,)) for ps in workers]
            for t in th:
                t.start()
            for t in th:
                t.join()
        if errors:
            raise errors[0]

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(steps, from_host, workers):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record(stream)
        run_steps(steps, from_host, workers)
        torch.cuda.synchronize()         # all proof streams have drained before the closing event
        e1.record(stream)
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    # a step is one batch of len(streams) proofs: with one proof per step the timed region of a 20-step run would be a
    # single generation of 16 in-flight proofs plus a drain (0.25 s); steps x batch proofs time the steady state
    batch = len(streams)
    run_steps(max(args.warmup, 1) * batch, False, streams)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    launches0 = sum(ps.ctx.launch_count() for ps in streams)
    total_ms = timed(args.steps * batch, False, streams)
    launches = sum(ps.ctx.launch_count() for ps in streams) - launches0
    run_steps(min(args.warmup, 2) * batch, True, streams)
    e2e_ms = timed(args.steps * batch, True, streams)
    clocks = sampler.stop() if rank == 0 else None
    # single-proof latency and per-kernel device times: one stream alone, CUDA events inside the library
    lat_steps = 5
    ctx.set_blocking_sync(False)         # a lone proof waits spinning: sleeping costs ~0.3 ms per wake-up, ~4 ms per proof
    ctx.timing_reset()
    lat_ms = timed(lat_steps, True, streams[:1]) / lat_steps
    ctx.set_blocking_sync(not args.spin)
    acc_ms, acc_spans, acc_pairs = ctx.timing(0)
    ntt_ms, ntt_spans, ntt_elems = ctx.timing(1)
    red_ms = ctx.timing(2)[0] + ctx.timing(3)[0] + ctx.timing(4)[0]
    msm_adds = ctx.timing(5)[2]                 # mixed point additions issued by k_msm_accumulate (10 field products each)
    ntt_products = ctx.timing(6)[2]
    proof_len = [streams[0].proof_len]
    # the arithmetic ceiling of this GPU, measured live: Montgomery products/s with every SM full
    mb_ms, mb_ops = ctx.microbench(0, 2000)
    peak_products = mb_ops / (mb_ms * 1e-3)

    if rank == 0:
        peak, peak_kind = peaks()
        ms_per_step = total_ms / args.steps
        value = world * batch * 1e3 / ms_per_step
        config["ms_per_proof"] = ms_per_step / batch
        achieved = MSM_BYTES_PER_PAIR * acc_pairs / (acc_ms * 1e-3) / 1e9
        ntt_ach = NTT_BYTES_PER_ELEM * ntt_elems / (ntt_ms * 1e-3) / 1e9
        in_bytes = sum(len(v) for v in inputs[0].values()) * 8
        line = {
            "metric": "bfv_prove_proofs_per_s", "value": value, "unit": "proofs/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u256 (BN254 Fr/Fq, 8xu32 Montgomery)",
            "data": "synthetic", "config": config, "clocks": clocks,
            "e2e": {"value": world * batch * 1e3 / (e2e_ms / args.steps), "unit": "proofs/s",
                    "h2d_bytes_per_step": int(in_bytes) * batch, "d2h_bytes_per_step": int(proof_len[0]) * batch,
                    "prove_latency_s_single_stream": lat_ms / 1e3},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "k_msm_accumulate", "achieved": achieved, "peak": peak,
                         "peak_source": f"MEASURED_PEAKS.json hbm_gbs ({peak_kind})", "unit": "GB/s",
                         "frac": achieved / peak, "traffic": TRAFFIC.get("k_msm_accumulate_dram_bytes_per_launch"),
                         "algorithmic_bytes_per_launch": MSM_BYTES_PER_PAIR * acc_pairs / max(acc_spans, 1),
                         "kernel_ms_per_launch": acc_ms / max(acc_spans, 1), "launches_per_proof": acc_spans / lat_steps,
                         "share_of_single_stream_proof": acc_ms / (lat_ms * lat_steps),
                         "measured": "one proof stream alone, CUDA events around every launch on its stream",
                         "traffic_source": TRAFFIC.get("source"),
                         "imad": {"bound": "int32 multiply pipe (fmaheavy)", "unit": "G Montgomery products/s",
                                  "achieved": 10 * msm_adds / (acc_ms * 1e-3) / 1e9, "peak": peak_products / 1e9,
                                  "frac": 10 * msm_adds / (acc_ms * 1e-3) / peak_products,
                                  "point_additions_per_proof": msm_adds / lat_steps,
                                  "peak_source": "zkfhe_microbench kind 0, measured in this run (two independent product "
                                                 "chains per thread, 8 CTAs x 256 threads per SM)"},
                         "note": "256-bit modular arithmetic: every kernel of the proof is bound by the INT32 multiply pipe "
                                 "(ncu sm__pipe_fmaheavy_cycles_active 87% for this kernel, profiles/), not by HBM; the "
                                 "hbm fraction is reported because the metric names it, the imad fraction is the one that "
                                 "says how close the kernel is to the chip's ceiling"},
            "roofline_ntt": {"bound": "hbm", "kernel": "k_ntt_pass", "achieved": ntt_ach, "peak": peak, "unit": "GB/s",
                             "frac": ntt_ach / peak, "share_of_single_stream_proof": ntt_ms / (lat_ms * lat_steps),
                             "kernel_ms_per_proof": ntt_ms / lat_steps,
                             "imad": {"unit": "G Montgomery products/s", "achieved": ntt_products / (ntt_ms * 1e-3) / 1e9,
                                      "peak": peak_products / 1e9, "frac": ntt_products / (ntt_ms * 1e-3) / peak_products}},
            "single_stream": {"prove_latency_ms": lat_ms, "msm_accumulate_ms": acc_ms / lat_steps,
                              "msm_sort_reduce_ms": red_ms / lat_steps, "ntt_ms": ntt_ms / lat_steps,
                              "other_ms": lat_ms - (acc_ms + ntt_ms + red_ms) / lat_steps},
        }
        if not args.no_cpu_baseline and world == 1:          # rank 0 at N = 1 only
            line["cpu_baseline"] = cpu_reference_arm(2, 1, budget_s=30.0)
    # ---- N > 1: ONE proof over all N GPUs (strong scaling of the single-proof latency) ----
    # The throughput number above keeps whole proofs per GPU (no data-path collective).  This is the other split: the
    # library shards the commitment phases by column and the quotient by coset behind zkfhe_prove_*; measured after the
    # timed region, on fresh contexts, and checked to give the single-GPU proof bytes on every rank.  It runs under a
    # watchdog so that a stuck collective can never cost the headline line.
    hung = False
    if world > 1 and not args.no_sharded:
        from zk_fhe_b200 import sharded
        for ps in streams:
            ps.ctx.sync()
        sharded_res = {}

        def measure():
            for label, kk, tk in (("k13_poseidon", 13, 1), ("k13_blake2b", 13, 0), ("k16_poseidon", 16, 1), ("k16_blake2b", 16, 0)):
                try:
                    sharded_res[label] = sharded.run(kk, 5 if kk == 13 else 3, tk, dist, rank, world, local_rank)
                except Exception as e:
                    sharded_res[label] = {"error": repr(e)[:300]}
                    return

        th = threading.Thread(target=measure, daemon=True)
        th.start()
        th.join(timeout=300.0)
        hung = th.is_alive()
        if rank == 0:
            line["sharded_single_proof"] = dict(sharded_res, **({"error": "timed out after 300 s"} if hung else {}))
    if rank == 0:
        print(json.dumps(line), flush=True)
    if hung:
        os._exit(0)                      # do not wait for a communicator that will never come back
    if world > 1:
        dist.destroy_process_group()
    ctx.close()
    return 0


if __name__ == "__main__":
    sys.exit(main())
