"""The `bfv` entrypoint (C++ host mirror of examples/bfv.rs + halo2-scaffold's CLI) through its
command line, as the reference's README drives it: mock, keygen, prove, verify."""
import json
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "zk-fhe_b200", "bin", "bfv")


@pytest.fixture()
def workdir(tmp_path, golden_dir):
    (tmp_path / "data" / "bfv").mkdir(parents=True)
    (tmp_path / "configs").mkdir()
    (tmp_path / "params").mkdir()
    for name in ("bfv.in", "bfv_empty.in"):
        shutil.copy(os.path.join(golden_dir, name), tmp_path / "data" / "bfv" / name)
    return tmp_path


def _run(workdir, *args):
    return subprocess.run([BIN, "--name", "bfv", "-k", "13", *args], cwd=workdir, capture_output=True, text=True, timeout=600)


def test_cli_mock_keygen_prove(workdir, golden_dir):
    import zk_fhe_b200
    zk_fhe_b200.load_library()           # builds the library and the binary if they are missing
    r = _run(workdir, "--input", "bfv/bfv.in", "mock")
    assert r.returncode == 0 and "all constraints satisfied" in r.stdout, r.stderr
    # no params file and no explicit opt-in to the public test trapdoor: refuse (the reference falls back silently)
    r = _run(workdir, "--input", "bfv/bfv_empty.in", "keygen")
    assert r.returncode == 1 and "setup" in r.stderr and "--insecure-test-srs" in r.stderr
    r = _run(workdir, "setup")                                  # params/kzg_bn254_13.srs from an OS-entropy trapdoor
    assert r.returncode == 0, r.stderr
    assert os.path.getsize(workdir / "params" / "kzg_bn254_13.srs") == 16 + 128 * 8192 + 128
    r = _run(workdir, "--input", "bfv/bfv_empty.in", "keygen")
    assert r.returncode == 0, r.stderr
    assert os.path.getsize(workdir / "data" / "bfv.pk") > 365 * 8192 * 32      # README.md:38: keygen writes the .pk
    got = json.load(open(workdir / "configs" / "bfv.json"))
    assert got == json.load(open(os.path.join(golden_dir, "bfv_pinning.json")))     # the reference's configs/bfv.json
    r = _run(workdir, "--input", "bfv/bfv.in", "prove")         # reads data/bfv.pk: no keygen inside prove
    assert r.returncode == 0 and "Proving time" in r.stdout and "Proving key loaded" in r.stdout, r.stderr
    assert os.path.getsize(workdir / "data" / "bfv.snark") > 50_000
    # README.md:48-54: verify reads data/bfv.vk (keygen) and data/bfv.snark (prove)
    assert os.path.getsize(workdir / "data" / "bfv.vk") == 72 + 64 * 365
    r = _run(workdir, "--input", "bfv/bfv.in", "verify")
    assert r.returncode == 0 and "Snark verified successfully" in r.stdout and "Verification time" in r.stdout, r.stdout + r.stderr
    snark = bytearray(open(workdir / "data" / "bfv.snark", "rb").read())
    snark[16 + 32 * 5121 + 32 * 10 + 3] ^= 1                    # one bit of an advice commitment (32 bytes per point)
    open(workdir / "data" / "bfv.snark", "wb").write(snark)
    r = _run(workdir, "--input", "bfv/bfv.in", "verify")
    assert r.returncode == 1 and "REJECTED" in r.stdout
    snark[16 + 32 * 5121 + 32 * 10 + 3] ^= 1
    snark[16 + 32 * 7] ^= 2                                      # a public input (a pk0 coefficient)
    open(workdir / "data" / "bfv.snark", "wb").write(snark)
    r = _run(workdir, "--input", "bfv/bfv.in", "verify")
    assert r.returncode == 1 and "REJECTED" in r.stdout


def test_cli_insecure_test_srs_is_opt_in_and_loud(workdir):
    """The fixed public trapdoor only behind --insecure-test-srs, with a warning; BLAKE2b transcript selectable."""
    import zk_fhe_b200
    zk_fhe_b200.load_library()
    flags = ("--insecure-test-srs", "--transcript", "blake2b")
    r = _run(workdir, *flags, "--input", "bfv/bfv_empty.in", "keygen")
    assert r.returncode == 0 and "WARNING" in r.stderr and "forged" in r.stderr, r.stderr
    r = _run(workdir, *flags, "--input", "bfv/bfv.in", "prove")
    assert r.returncode == 0 and "WARNING" in r.stderr, r.stderr
    r = _run(workdir, *flags, "verify")
    assert r.returncode == 0 and "Snark verified successfully" in r.stdout, r.stdout + r.stderr
    # serving shape from C++ host threads: 12 more proofs on 4 proof streams, then the usual .snark (which still verifies)
    r = _run(workdir, *flags, "--input", "bfv/bfv.in", "--repeat", "12", "--streams", "4", "prove")
    assert r.returncode == 0 and "Throughput:" in r.stdout and "4 proof streams" in r.stdout, r.stdout + r.stderr
    r = _run(workdir, *flags, "verify")
    assert r.returncode == 0 and "Snark verified successfully" in r.stdout, r.stdout + r.stderr
    r = _run(workdir, "verify")                                  # the params file is missing: verify refuses too
    assert r.returncode == 1 and "setup" in r.stderr
    os.remove(workdir / "data" / "bfv.pk")
    r = _run(workdir, *flags, "--input", "bfv/bfv.in", "prove")
    assert r.returncode == 1 and "run keygen first" in r.stderr


def test_cli_mock_rejects_a_tampered_input(workdir):
    d = json.load(open(workdir / "data" / "bfv" / "bfv.in"))
    d["c0"][0] = str((int(d["c0"][0]) + 1) % 536870909)
    json.dump(d, open(workdir / "data" / "bfv" / "bad.in", "w"))
    r = _run(workdir, "--input", "bfv/bad.in", "mock")
    assert r.returncode == 1 and "constraint violations" in r.stderr
    d["pk0"] = d["pk0"][:-1]                                   # wrong degree: examples/bfv.rs:82 assert
    json.dump(d, open(workdir / "data" / "bfv" / "short.in", "w"))
    r = _run(workdir, "--input", "bfv/short.in", "mock")
    assert r.returncode == 1 and "bfv.rs:82" in r.stderr
