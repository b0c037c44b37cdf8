"""GPU parity tests: the CUDA path, called through the C ABI, against the oracle.

Bit-exact: everything here is integer / finite-field work.  Sizes the oracle
finishes in seconds are compared value by value; BASELINE.json's full sizes are
covered by size-independent properties (round trips, linearity).
"""
import random

import numpy as np
import pytest

from oracle import cbind, curve, field, ntt as ontt
from tests.util import fr_to_mont_array, mont_array_to_fr, random_fr_mont, toy_srs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    import zk_fhe_b200
    c = zk_fhe_b200.Context(0)
    yield c
    c.close()


def test_field_arithmetic_selftest_on_device(ctx):
    """Generated PTX Montgomery ops vs an independent plain-C product + identities, on the chip."""
    assert ctx.selftest(1 << 18, seed=7) == 0
    assert ctx.selftest(1 << 12, seed=123456789) == 0


# ---------------------------------------------------------------- NTT ------
def test_ntt_golden_vectors(ctx, digests):
    for key, v in digests["ntt"].items():
        k = int(key[1:])
        a = [int(x, 16) for x in v["in"]]
        want = [int(x, 16) for x in v["out"]]
        data = fr_to_mont_array(a)
        ctx.ntt_fr(data, k, 1)
        assert mont_array_to_fr(data) == want
        ctx.ntt_fr(data, k, 1, inverse=True)
        assert mont_array_to_fr(data) == a


@pytest.mark.parametrize("k", [1, 2, 5, 9, 11, 12, 13, 15, 16])
@pytest.mark.parametrize("inverse,coset", [(False, False), (True, False), (False, True), (True, True)])
def test_ntt_matches_oracle(ctx, k, inverse, coset):
    rng = np.random.default_rng(1000 + k)
    batch = 3 if k <= 13 else 2
    data = random_fr_mont(rng, batch << k)
    want = data.copy()
    cbind.ntt(want, k, batch, inverse=inverse, coset=coset)
    ctx.ntt_fr(data, k, batch, inverse=inverse, coset=coset)
    assert np.array_equal(data, want)


@pytest.mark.parametrize("k,batch", [(8, 1), (8, 32), (10, 2), (10, 8), (10, 9), (11, 1), (11, 4), (11, 5), (7, 3), (12, 1), (13, 1), (13, 6), (13, 9),
                                     (15, 1), (15, 2), (15, 3), (16, 1), (16, 2), (17, 1)])
@pytest.mark.parametrize("inverse,coset", [(False, False), (True, True)])
def test_small_ntt_both_paths(ctx, k, batch, inverse, coset):
    """Transforms of 2^8..2^11 points take the tiled two-pass path while batch * n <= 8192 (the lone transforms of
    stage (1)) and the one-CTA path above that; larger ones use 256-element tiles while batch * n <= 65536 and
    1024-element tiles above: all against the oracle, either side of each switch."""
    rng = np.random.default_rng(7000 + 31 * k + batch)
    data = random_fr_mont(rng, batch << k)
    want = data.copy()
    cbind.ntt(want, k, batch, inverse=inverse, coset=coset)
    ctx.ntt_fr(data, k, batch, inverse=inverse, coset=coset)
    assert np.array_equal(data, want)


def test_ntt_edge_inputs(ctx):
    k = 13
    n = 1 << k
    zero = np.zeros((n, 4), np.uint64)
    ctx.ntt_fr(zero, k, 1)
    assert not zero.any()
    # delta at 0 -> all ones; constant -> n * delta
    one = fr_to_mont_array([1] + [0] * (n - 1))
    ctx.ntt_fr(one, k, 1)
    assert mont_array_to_fr(one) == [1] * n
    ctx.ntt_fr(one, k, 1)
    assert mont_array_to_fr(one) == [n] + [0] * (n - 1)
    # maximal values
    mx = fr_to_mont_array([field.R_MOD - 1] * n)
    ctx.ntt_fr(mx, k, 1)
    assert mont_array_to_fr(mx) == [(field.R_MOD - n) % field.R_MOD] + [0] * (n - 1)
    ctx.ntt_fr(np.zeros((0, 4), np.uint64), k, 0)          # empty batch is a no-op


def test_ntt_2_pow_21_matches_oracle(ctx):
    """The extended domain of config 5 (k = 19 -> 2^21 points): the largest transform the prover issues."""
    k = 21
    rng = np.random.default_rng(21)
    data = random_fr_mont(rng, 1 << k)
    want = data.copy()
    cbind.ntt(want, k, 1, coset=True)
    ctx.ntt_fr(data, k, 1, coset=True)
    assert np.array_equal(data, want)
    ctx.ntt_fr(data, k, 1, inverse=True, coset=True)
    cbind.ntt(want, k, 1, inverse=True, coset=True)
    assert np.array_equal(data, want)


@pytest.mark.parametrize("k", [13, 16, 19])
def test_ntt_full_size_roundtrip_and_linearity(ctx, k):
    """BASELINE.json sizes (k = 13, 16, 19): iNTT(NTT(x)) == x, coset round trip, linearity."""
    rng = np.random.default_rng(k)
    n = 1 << k
    x = random_fr_mont(rng, 2 * n)
    orig = x.copy()
    ctx.ntt_fr(x, k, 2)
    fx = x.copy()
    ctx.ntt_fr(x, k, 2, inverse=True)
    assert np.array_equal(x, orig)
    ctx.ntt_fr(x, k, 2, coset=True)
    ctx.ntt_fr(x, k, 2, inverse=True, coset=True)
    assert np.array_equal(x, orig)
    # NTT(a + b) == NTT(a) + NTT(b) on a few sampled positions
    a = cbind.u64x4_to_ints(orig[:n])
    b = cbind.u64x4_to_ints(orig[n:2 * n])
    s = cbind.ints_to_u64x4([(u + v) % field.R_MOD for u, v in zip(a, b)])
    ctx.ntt_fr(s, k, 1)
    fa = cbind.u64x4_to_ints(fx[:n])
    fb = cbind.u64x4_to_ints(fx[n:2 * n])
    fs = cbind.u64x4_to_ints(s)
    for i in (0, 1, n // 2, n - 1, 12345 % n):
        assert fs[i] == (fa[i] + fb[i]) % field.R_MOD


def test_coeff_to_extended_matches_oracle(ctx):
    import torch
    k, k_ext, batch = 11, 13, 2
    rng = np.random.default_rng(5)
    coeffs = random_fr_mont(rng, batch << k)
    d_in = torch.from_numpy(coeffs.view(np.int64)).cuda()
    d_out = torch.empty((batch << k_ext, 4), dtype=torch.int64, device="cuda")
    ctx.coeff_to_extended_dev(d_in.data_ptr(), k, d_out.data_ptr(), k_ext, batch)
    ctx.sync()
    got = d_out.cpu().numpy().view(np.uint64)
    want = np.zeros((batch << k_ext, 4), np.uint64)
    for b in range(batch):
        want[b << k_ext:(b << k_ext) + (1 << k)] = coeffs[b << k:(b + 1) << k]
    cbind.ntt(want, k_ext, batch, coset=True)
    assert np.array_equal(got, want)
    # spot check one value against the definition: p(zeta * w_ext^j)
    c0 = mont_array_to_fr(coeffs[:1 << k])
    j = 77
    xj = field.FR_ZETA * pow(field.omega(k_ext), j, field.R_MOD) % field.R_MOD
    assert mont_array_to_fr(got[j:j + 1])[0] == ontt.poly_eval(c0, xj)


# ---------------------------------------------------------------- MSM ------
def _pt(row):
    return curve.g1_from_mont_bytes(row.tobytes())


def test_msm_golden_vector(ctx, digests):
    m = digests["msm"]
    pts = [(int(x, 16), int(y, 16)) for x, y in m["points"]]
    sc = [int(s, 16) for s in m["scalars"]]
    bases = np.frombuffer(b"".join(curve.g1_to_mont_bytes(p) for p in pts), np.uint64).reshape(16, 8).copy()
    ctx.load_srs(4, g=bases, g_lagrange=bases)
    out = ctx.msm_g1(fr_to_mont_array(sc), 1, basis=0)
    assert curve.g1_from_mont_bytes(out) == (int(m["result"][0], 16), int(m["result"][1], 16))


@pytest.mark.parametrize("k", [3, 6, 10])
def test_msm_matches_oracle_random_and_edge_scalars(ctx, k):
    n = 1 << k
    g, gl = toy_srs(k)
    ctx.load_srs(k, g=g, g_lagrange=gl)
    rng = np.random.default_rng(k)
    pyr = random.Random(k)
    cols = [random_fr_mont(rng, n)]
    edge = [0, 1, field.R_MOD - 1, 2, 255, 256, field.R_MOD - 2, (1 << 253) % field.R_MOD]
    cols.append(fr_to_mont_array([edge[i % len(edge)] for i in range(n)]))
    cols.append(fr_to_mont_array([0] * n))                                   # all-zero column -> identity
    cols.append(fr_to_mont_array([1] * n))                                   # one hot bucket (skew)
    cols.append(fr_to_mont_array([pyr.randrange(256) for _ in range(n)]))    # lookup-like 8-bit cells
    cols.append(fr_to_mont_array([(-(1 << 32)) % field.R_MOD if i % 7 == 0 else pyr.randrange(2) for i in range(n)]))
    scal = np.ascontiguousarray(np.concatenate(cols))
    for basis, bases in ((0, g), (1, gl)):
        want = cbind.msm(scal, bases, n, len(cols))
        got = np.frombuffer(ctx.msm_g1(scal, len(cols), basis=basis), np.uint64).reshape(len(cols), 8)
        assert np.array_equal(got, want), f"basis {basis}"
    assert not got[2].any()                                                  # identity is 64 zero bytes


def test_msm_k13_witness_like_columns_match_oracle(ctx):
    """Config-1 size (n = 2^13) with the value mix of real advice columns."""
    k, n = 13, 1 << 13
    g, gl = toy_srs(k)
    ctx.load_srs(k, g=None, g_lagrange=gl)
    rng = np.random.default_rng(99)
    pyr = random.Random(99)
    cols = [random_fr_mont(rng, n), random_fr_mont(rng, n)]
    cols.append(fr_to_mont_array([pyr.randrange(256) for _ in range(n)]))
    cols.append(fr_to_mont_array([pyr.randrange(536870909) if i % 3 else (-(1 << 40)) % field.R_MOD for i in range(n)]))
    scal = np.ascontiguousarray(np.concatenate(cols))
    want = cbind.msm(scal, gl, n, len(cols))
    got = np.frombuffer(ctx.msm_g1(scal, len(cols), basis=1), np.uint64).reshape(len(cols), 8)
    assert np.array_equal(got, want)
    # linearity at full size: msm(a) + msm(b) == msm(a + b)
    a = cbind.u64x4_to_ints(cols[0])
    b = cbind.u64x4_to_ints(cols[1])
    s = cbind.ints_to_u64x4([(u + v) % field.R_MOD for u, v in zip(a, b)])   # Montgomery form is linear
    ps = _pt(np.frombuffer(ctx.msm_g1(s, 1, basis=1), np.uint64))
    assert ps == curve.g1_add(_pt(got[0]), _pt(got[1]))


@pytest.mark.parametrize("batch", [1, 2, 3, 31, 33])
def test_msm_k13_uniform_columns_few_and_many(ctx, batch):
    """Uniform scalars at config-1 size on both sides of the few-column switch (batch < 32: cluster sort, 16-reference
    slices, warp fold; batch >= 32: one CTA per column, 64-reference slices), all ending
    in the CTA-wide final reduction: against the C oracle, column by column."""
    k, n = 13, 1 << 13
    g, gl = toy_srs(k)
    ctx.load_srs(k, g=g, g_lagrange=None)
    rng = np.random.default_rng(4000 + batch)
    scal = random_fr_mont(rng, batch * n)
    want = cbind.msm(scal, g, n, batch)
    got = np.frombuffer(ctx.msm_g1(scal, batch, basis=0), np.uint64).reshape(batch, 8)
    assert np.array_equal(got, want)


def test_msm_small_value_hint_gives_identical_points(ctx):
    """zkfhe_msm_g1_dev_ex(small_values=1) (narrow-window table) == the plain MSM == the oracle, on
    witness-like, full-size, all-zero and single-hot-bucket columns."""
    import torch
    k, n = 13, 1 << 13
    _, gl = toy_srs(k)
    ctx.load_srs(k, g=None, g_lagrange=gl)
    rng = np.random.default_rng(5)
    pyr = random.Random(5)
    cols = [random_fr_mont(rng, n),
            fr_to_mont_array([pyr.randrange(536870909) if i % 5 else (field.R_MOD - 1 - i) for i in range(n)]),
            fr_to_mont_array([pyr.randrange(4096) for _ in range(n)]),
            fr_to_mont_array([0] * n),
            fr_to_mont_array([7] * n)]
    scal = np.ascontiguousarray(np.concatenate(cols))
    want = cbind.msm(scal, gl, n, len(cols))
    d = torch.from_numpy(scal.view(np.int64).reshape(-1)).cuda()
    out = torch.zeros(len(cols) * 8, dtype=torch.int64, device="cuda")
    for small in (True, False):
        out.zero_()
        ctx.msm_g1_dev(d.data_ptr(), len(cols), 1, out.data_ptr(), small_values=small)
        ctx.sync()
        got = out.cpu().numpy().view(np.uint64).reshape(len(cols), 8)
        assert np.array_equal(got, want), f"small_values={small}"


def test_msm_runs_of_equal_scalars_go_through_the_prefix_table(ctx):
    """Runs of equal scalars are committed as prefix[b+1] - prefix[a] of the basis (two references per window
    instead of one per row): runs at both ends of the column, runs of length 2, full-size values repeated, runs of
    zeros, a column that is one single run, runs touching each other, with and without the small-value hint."""
    import torch
    k, n = 13, 1 << 13
    _, gl = toy_srs(k)
    ctx.load_srs(k, g=None, g_lagrange=gl)
    pyr = random.Random(77)
    big = [pyr.randrange(field.R_MOD) for _ in range(8)]

    def runs(lengths, values):
        out = []
        for i in range(10 ** 9):
            if len(out) >= n:
                return out[:n]
            out += [values[i % len(values)]] * lengths[i % len(lengths)]

    cols = [runs([3000, 1, 2, 5000, 7], big),                      # the shape of a grand-product column
            runs([2], big[:3]),                                    # every run has length 2
            runs([1, 2, 1, 3], [0, 5, 0, 0, 9, 1]),                # runs of zeros between short runs
            runs([n], [big[0]]),                                   # one run
            runs([1], big),                                        # no runs at all
            runs([n - 1, 1], [big[1], big[2]]),                    # run up to the last row but one
            runs([1, n - 1], [big[3], big[4]]),                    # run from row 1 to the end
            runs([17, 4000], [1, 2, field.R_MOD - 1])]
    scal = np.ascontiguousarray(np.concatenate([fr_to_mont_array(c) for c in cols]))
    want = cbind.msm(scal, gl, n, len(cols))
    d = torch.from_numpy(scal.view(np.int64).reshape(-1)).cuda()
    out = torch.zeros(len(cols) * 8, dtype=torch.int64, device="cuda")
    for small in (False, True):
        out.zero_()
        ctx.msm_g1_dev(d.data_ptr(), len(cols), 1, out.data_ptr(), small_values=small)
        ctx.sync()
        assert np.array_equal(out.cpu().numpy().view(np.uint64).reshape(len(cols), 8), want), f"small_values={small}"
    # one column alone and three columns: the cluster sort path of the few-column commits
    for lo, cnt in ((0, 1), (3, 3)):
        out.zero_()
        ctx.msm_g1_dev(d.data_ptr() + lo * n * 32, cnt, 1, out.data_ptr())
        ctx.sync()
        assert np.array_equal(out.cpu().numpy().view(np.uint64).reshape(len(cols), 8)[:cnt], want[lo:lo + cnt])


def test_msm_k16_long_skewed_columns_match_oracle(ctx):
    """2^16-scalar columns (config 3/4 size): the cluster counting sort and the multi-level combine of hot buckets
    (0/1-valued and constant columns put tens of thousands of references into one bucket), with and without the
    small-value hint, against the C oracle."""
    import torch
    k, n = 16, 1 << 16
    _, gl = toy_srs(k)
    ctx.load_srs(k, g=None, g_lagrange=gl)
    rng = np.random.default_rng(16)
    pyr = random.Random(16)
    cols = [random_fr_mont(rng, n),
            fr_to_mont_array([pyr.randrange(2) for _ in range(n)]),                       # bits: one hot bucket
            fr_to_mont_array([3] * n),                                                    # a constant column
            fr_to_mont_array([pyr.randrange(1 << 61) if i % 9 else field.R_MOD - 1 - (i % 5) for i in range(n)])]
    scal = np.ascontiguousarray(np.concatenate(cols))
    want = cbind.msm(scal, gl, n, len(cols))
    d = torch.from_numpy(scal.view(np.int64).reshape(-1)).cuda()
    out = torch.zeros(len(cols) * 8, dtype=torch.int64, device="cuda")
    for small in (False, True):
        out.zero_()
        ctx.msm_g1_dev(d.data_ptr(), len(cols), 1, out.data_ptr(), small_values=small)
        ctx.sync()
        assert np.array_equal(out.cpu().numpy().view(np.uint64).reshape(len(cols), 8), want), f"small_values={small}"
    # the same columns ten times over (40 columns): the thread-fold path of the big commits
    reps = 10
    d2 = d.repeat(reps)
    out2 = torch.zeros(reps * len(cols) * 8, dtype=torch.int64, device="cuda")
    ctx.msm_g1_dev(d2.data_ptr(), reps * len(cols), 1, out2.data_ptr())
    ctx.sync()
    got2 = out2.cpu().numpy().view(np.uint64).reshape(reps, len(cols), 8)
    for r in range(reps):
        assert np.array_equal(got2[r], want)


def test_srs_setup_on_gpu_matches_oracle(ctx):
    """zkfhe_srs_setup (ParamsKZG::setup shape) vs the oracle's g[i] = tau^i G, g_lagrange[i] = l_i(tau) G."""
    import zk_fhe_b200
    k, tau = 7, 0x1F2E3D4C5B6A79880123456789
    c2 = zk_fhe_b200.Context(0)
    g, gl = c2.srs_setup(k, tau, want_host_copy=True)
    og, ogl = cbind.srs(k, tau)
    assert g == og.tobytes() and gl == ogl.tobytes()
    rng = np.random.default_rng(3)
    sc = random_fr_mont(rng, 1 << k)
    for basis, bases in ((0, og), (1, ogl)):
        assert c2.msm_g1(sc, 1, basis=basis) == cbind.msm(sc, bases, 1 << k, 1).tobytes()
    c2.close()


def test_fr_convert_roundtrip(ctx):
    import torch
    vals = [0, 1, 2, field.R_MOD - 1, 1 << 200, 536870909]
    canon = cbind.ints_to_u64x4(vals)
    d = torch.from_numpy(canon.view(np.int64)).cuda()
    ctx.fr_convert_dev(d.data_ptr(), len(vals), True)
    ctx.sync()
    assert mont_array_to_fr(d.cpu().numpy().view(np.uint64)) == vals
    ctx.fr_convert_dev(d.data_ptr(), len(vals), False)
    ctx.sync()
    assert np.array_equal(d.cpu().numpy().view(np.uint64), canon)


def test_msm_errors(ctx):
    import zk_fhe_b200
    c2 = zk_fhe_b200.Context(0)
    with pytest.raises(zk_fhe_b200.ZkfheError) as e:
        c2.msm_g1(np.zeros((8, 4), np.uint64), 1)
    assert e.value.code == -3        # ZKFHE_ERR_STATE: MSM before load_srs
    with pytest.raises(zk_fhe_b200.ZkfheError):
        c2.ntt_fr(np.zeros((8, 4), np.uint64), 23, 1)
    c2.close()
