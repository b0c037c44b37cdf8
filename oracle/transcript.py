"""Fiat-Shamir transcripts (oracle; test-only): independent restatement of the two hashes in
zk-fhe_b200/csrc/host_ff.h so that the verifier replays exactly what the prover hashed.

  Blake2bTranscript   halo2 `Blake2bWrite`/`Challenge255` shape [UPSTREAM-RECALL]: running
                      BLAKE2b-512 with one-byte tags; a challenge is the digest of the state so
                      far reduced from 512 bits (halo2curves `from_uniform_bytes`).
  PoseidonTranscript  snark-verifier's PoseidonTranscript<NativeLoader> over the PSE `poseidon` crate
                      (t=5, rate 4, R_F=8, R_P=60) [UPSTREAM-RECALL]: state[0] starts at 2^64, scalars
                      are absorbed natively, points as (x mod r, y mod r), a squeeze pads with one 1.
                      Constants come from the published Grain-LFSR procedure; the procedure and the
                      permutation below are PINNED by the published t=3 / t=2 vectors
                      (tests/test_oracle_transcript.py); the t=5 tables themselves have no published
                      vector offline (the upstream crates are un-vendored).
"""
import hashlib

from .field import R_MOD

TAG = b"zkfhe-b200-transcript-v1"


class Blake2bTranscript:
    def __init__(self):
        self.h = hashlib.blake2b(digest_size=64)
        self.h.update(TAG)

    def common_scalar(self, x):
        self.h.update(b"\x02" + int(x).to_bytes(32, "little"))

    def common_point(self, pt):
        x, y = (0, 0) if pt is None else pt
        self.h.update(b"\x01" + x.to_bytes(32, "little") + y.to_bytes(32, "little"))

    def squeeze(self):
        self.h.update(b"\x00")
        return int.from_bytes(self.h.copy().digest(), "little") % R_MOD


# ---- Poseidon --------------------------------------------------------------------------------
T, R_F, R_P, N_BITS = 5, 8, 60, 254


def _grain_bits(T=T, R_F=R_F, R_P=R_P):
    state = []
    for val, width in ((1, 2), (0, 4), (N_BITS, 12), (T, 12), (R_F, 10), (R_P, 10)):
        state += [(val >> (width - 1 - i)) & 1 for i in range(width)]
    state += [1] * 30

    def step():
        b = state[62] ^ state[51] ^ state[38] ^ state[23] ^ state[13] ^ state[0]
        state.pop(0)
        state.append(b)
        return b

    for _ in range(160):
        step()
    while True:
        b = step()
        while b == 0:
            step()
            b = step()
        yield step()


def poseidon_params(T=T, R_F=R_F, R_P=R_P):
    """(round constants, MDS matrix) of Poseidon-x^5 over BN254 Fr for the given width / round numbers."""
    g = _grain_bits(T, R_F, R_P)

    def bits(n):
        v = 0
        for _ in range(n):
            v = (v << 1) | next(g)
        return v

    rc = []
    while len(rc) < (R_F + R_P) * T:
        v = bits(N_BITS)
        if v < R_MOD:
            rc.append(v)
    while True:
        xy = [bits(N_BITS) % R_MOD for _ in range(2 * T)]
        if len(set(xy)) != 2 * T:
            continue
        xs, ys = xy[:T], xy[T:]
        if any((a + b) % R_MOD == 0 for a in xs for b in ys):
            continue
        return rc, [[pow((a + b) % R_MOD, -1, R_MOD) for b in ys] for a in xs]


_PARAMS = None


def poseidon_permute(s, params=None, R_F=R_F, R_P=R_P):
    """The plain (textbook) permutation; `params` defaults to the transcript's t=5 tables."""
    global _PARAMS
    if params is None:
        if _PARAMS is None:
            _PARAMS = poseidon_params()
        params = _PARAMS
    rc, mds = params
    T = len(mds)
    half = R_F // 2
    for r in range(R_F + R_P):
        s = [(s[i] + rc[r * T + i]) % R_MOD for i in range(T)]
        if r < half or r >= half + R_P:
            s = [pow(v, 5, R_MOD) for v in s]
        else:
            s[0] = pow(s[0], 5, R_MOD)
        s = [sum(mds[i][j] * s[j] for j in range(T)) % R_MOD for i in range(T)]
    return s


class PoseidonTranscript:
    def __init__(self):
        self.state = [1 << 64] + [0] * (T - 1)
        self.buf = []

    def common_scalar(self, x):
        self.buf.append(int(x) % R_MOD)

    def common_point(self, pt):
        x, y = (0, 0) if pt is None else pt
        self.buf += [x % R_MOD, y % R_MOD]

    def squeeze(self):
        self.buf.append(1)
        while len(self.buf) % 4:
            self.buf.append(0)
        for i in range(0, len(self.buf), 4):
            for j in range(4):
                self.state[1 + j] = (self.state[1 + j] + self.buf[i + j]) % R_MOD
            self.state = poseidon_permute(self.state)
        self.buf = []
        return self.state[1]


def make(kind):
    return PoseidonTranscript() if kind == 1 else Blake2bTranscript()


# ---- the reference's fallback SRS trapdoor ------------------------------------------------------------------------
# halo2-scaffold `gen_srs(k)` without a params file: `ParamsKZG::setup(k, ChaCha20Rng::from_seed([0; 32]))`; the first
# draw is tau = `Fr::random(rng)`: eight `next_u64` = the first 64 keystream bytes, little-endian, as one 512-bit integer
# reduced mod r [UPSTREAM-RECALL, SURVEY.md App. C.1].  Zero key / counter / nonce is RFC 7539 A.1 test vector #1.
def chacha20_block(key_words, counter=0, stream=0):
    def rotl(v, n):
        return ((v << n) | (v >> (32 - n))) & 0xFFFFFFFF

    init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574, *key_words, counter & 0xFFFFFFFF, counter >> 32,
            stream & 0xFFFFFFFF, stream >> 32]
    x = list(init)

    def qr(a, b, c, d):
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = rotl(x[d] ^ x[a], 16)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = rotl(x[b] ^ x[c], 12)
        x[a] = (x[a] + x[b]) & 0xFFFFFFFF; x[d] = rotl(x[d] ^ x[a], 8)
        x[c] = (x[c] + x[d]) & 0xFFFFFFFF; x[b] = rotl(x[b] ^ x[c], 7)

    for _ in range(10):
        qr(0, 4, 8, 12); qr(1, 5, 9, 13); qr(2, 6, 10, 14); qr(3, 7, 11, 15)
        qr(0, 5, 10, 15); qr(1, 6, 11, 12); qr(2, 7, 8, 13); qr(3, 4, 9, 14)
    return b"".join(((x[i] + init[i]) & 0xFFFFFFFF).to_bytes(4, "little") for i in range(16))


def reference_test_tau():
    return int.from_bytes(chacha20_block([0] * 8), "little") % R_MOD
