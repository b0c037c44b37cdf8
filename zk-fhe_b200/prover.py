"""Host-side handles of keygen / prove (C ABI stage: keygen + prover), for the CLI mirror and
the tests.  `keygen(circuit)` is the reference's `keygen` subcommand; the pinning it returns has
the schema of the reference's configs/bfv.json.
"""
import ctypes
import json
import os

import numpy as np

from .capi import _addr

INFO_FIELDS = ("k", "n_gate0", "n_gate1", "n_rlc", "n_lookup", "n_advice", "n_perm", "n_fixed", "n_chunks",
               "usable_rows", "max_rows", "lookup_bits", "instances", "fx_sigma", "fx_const", "fx_table")


class ProvingKey:
    def __init__(self, ctx, handle):
        self.ctx = ctx
        self.h = handle
        info = (ctypes.c_uint32 * 16)()
        ctx._check(ctx.lib.zkfhe_pk_info(handle, info))
        self.info = dict(zip(INFO_FIELDS, (int(x) for x in info)))

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.zkfhe_pk_free(self.h)
        except Exception:
            pass
        self.h = None

    def pinning(self):
        """dict in the schema of the reference's configs/<name>.json."""
        need = ctypes.c_size_t()
        self.ctx._check(self.ctx.lib.zkfhe_pk_pinning_json(self.h, None, 0, ctypes.byref(need)))
        buf = ctypes.create_string_buffer(need.value)
        self.ctx._check(self.ctx.lib.zkfhe_pk_pinning_json(self.h, buf, need.value, None))
        return json.loads(buf.value.decode())

    def vk_bytes(self):
        """The verifying key as bytes (the reference's data/<name>.vk)."""
        need = ctypes.c_size_t()
        self.ctx._check(self.ctx.lib.zkfhe_vk_export(self.h, None, 0, ctypes.byref(need)))
        buf = bytearray(need.value)
        self.ctx._check(self.ctx.lib.zkfhe_vk_export(self.h, _addr(buf), need.value, None))
        return bytes(buf)

    def export_bytes(self):
        """The proving key as bytes (the reference's data/<name>.pk); `import_key` reads it back."""
        need = ctypes.c_size_t()
        self.ctx._check(self.ctx.lib.zkfhe_pk_export(self.h, None, 0, ctypes.byref(need)))
        buf = bytearray(need.value)
        self.ctx._check(self.ctx.lib.zkfhe_pk_export(self.h, _addr(buf), need.value, None))
        return bytes(buf)

    def fixed(self, index, form=0):
        """(rows, 4) uint64 Montgomery; form 0 Lagrange, 1 coefficients, 2 extended coset."""
        n = (1 << self.info["k"]) * (4 if form == 2 else 1)
        out = np.zeros((n, 4), dtype=np.uint64)
        self.ctx._check(self.ctx.lib.zkfhe_pk_download_fixed(self.h, index, form, _addr(out)))
        return out

    def fixed_commitments(self):
        out = np.zeros((self.info["n_fixed"], 8), dtype=np.uint64)
        self.ctx._check(self.ctx.lib.zkfhe_pk_fixed_commitments(self.h, _addr(out)))
        return out


TRANSCRIPT_BLAKE2B, TRANSCRIPT_POSEIDON = 0, 1


class Prover:
    """One proof: phase0(witness) -> gamma; (caller runs the phase-1 chip calls); finish(witness) -> bytes."""

    def __init__(self, pk, seed=None, transcript=TRANSCRIPT_POSEIDON, ctx=None):
        """`ctx`: the context (stream) this prover runs on; defaults to the key's own context.
        `seed`: 32-byte ChaCha20 key of the blinding factors.  None (the default) draws it from the OS, as the
        reference's `StdRng::from_entropy()` does; a fixed seed gives publicly known blinding factors and is for
        tests / reproducible benchmarks only."""
        self.pk = pk
        self.ctx = ctx or pk.ctx
        if seed is None:
            seed = os.urandom(32)
        assert len(seed) == 32
        self._seed = bytearray(seed)
        h = ctypes.c_void_p()
        self.ctx._check(self.ctx.lib.zkfhe_prove_begin(self.ctx.h, pk.h, _addr(self._seed), transcript, ctypes.byref(h)))
        self.h = h

    def __del__(self):
        try:
            if self.h and self.ctx.h:
                self.ctx.lib.zkfhe_prover_free(self.h)
        except Exception:
            pass
        self.h = None

    ROUNDS = ("phase0_commit", "phase1_commit", "lookup_permute", "grand_products", "quotient", "evaluations",
              "shplonk_quotient", "shplonk_opening")

    def round_ms(self):
        """Wall-clock of each round of the last proof (ms)."""
        out = (ctypes.c_double * 8)()
        self.ctx._check(self.ctx.lib.zkfhe_prover_round_ms(self.h, out))
        return dict(zip(self.ROUNDS, (round(float(x), 3) for x in out)))

    def reset(self, seed):
        """Next proof with the same device buffers."""
        assert len(seed) == 32
        self._seed = bytearray(seed)
        self.ctx._check(self.ctx.lib.zkfhe_prove_reset(self.h, _addr(self._seed)))

    def phase0(self, witness):
        """Commit the phase-0 advice; returns the challenge gamma as a canonical int."""
        out = bytearray(32)
        self.ctx._check(self.ctx.lib.zkfhe_prove_phase0(self.h, witness.h, _addr(out)))
        from .capi import FR_MODULUS
        return int.from_bytes(out, "little") * pow(1 << 256, -1, FR_MODULUS) % FR_MODULUS

    def finish(self, witness):
        p, n = ctypes.c_void_p(), ctypes.c_size_t()
        self.ctx._check(self.ctx.lib.zkfhe_prove_finish(self.h, witness.h, ctypes.byref(p), ctypes.byref(n)))
        proof = ctypes.string_at(p.value, n.value)
        self.ctx.lib.zkfhe_proof_free(p)
        return proof


def prove(pk, circuit_factory, inp, seed=None, transcript=TRANSCRIPT_POSEIDON):
    """The reference's `prove` subcommand for one input: returns (proof bytes, circuit).  seed=None: OS entropy."""
    circ = circuit_factory()
    circ.phase0(inp)
    pr = Prover(pk, seed, transcript)
    gamma = pr.phase0(circ.wit)
    circ.phase1(gamma)
    return pr.finish(circ.wit), circ


def verify(ctx, vk_bytes, instances, proof, s_g2, transcript=TRANSCRIPT_POSEIDON):
    """The reference's `verify` subcommand: True / False.  `instances`: canonical ints; `s_g2`: [tau]_2
    (Context.srs_g2 for the test SRS).  ctx.last_rejection holds the reason of a rejection."""
    inst = b"".join(int(v).to_bytes(32, "little") for v in instances)
    ok = ctypes.c_int(0)
    ctx._check(ctx.lib.zkfhe_verify(ctx.h, _addr(vk_bytes), len(vk_bytes), _addr(inst) if inst else None, len(instances),
                                    _addr(proof), len(proof), _addr(s_g2), transcript, ctypes.byref(ok)))
    ctx.last_rejection = None if ok.value else ctx.lib.zkfhe_last_error(ctx.h).decode()
    return bool(ok.value)


def import_key(ctx, blob):
    """A proving key from the bytes `ProvingKey.export_bytes` wrote; the SRS for its k must be loaded on `ctx`."""
    h = ctypes.c_void_p()
    ctx._check(ctx.lib.zkfhe_pk_import(ctx.h, _addr(blob), len(blob), ctypes.byref(h)))
    return ProvingKey(ctx, h)


def keygen(witness, k, unusable_rows=109):
    """`witness`: a Witness built with record=True on the keygen input (both phases run)."""
    h = ctypes.c_void_p()
    witness.ctx._check(witness.ctx.lib.zkfhe_keygen(witness.h, k, unusable_rows, ctypes.byref(h)))
    return ProvingKey(witness.ctx, h)
