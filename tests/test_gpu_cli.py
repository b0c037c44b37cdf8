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
    r = _run(workdir, "--input", "bfv/bfv_empty.in", "keygen")
    assert r.returncode == 0, r.stderr
    got = json.load(open(workdir / "configs" / "bfv.json"))
    assert got == json.load(open(os.path.join(golden_dir, "bfv_pinning.json")))     # the reference's configs/bfv.json
    r = _run(workdir, "--input", "bfv/bfv.in", "prove")
    assert r.returncode == 0 and "Proving time" in r.stdout, r.stderr
    assert os.path.getsize(workdir / "data" / "bfv.snark") > 50_000
    # README.md:48-54: verify reads data/bfv.vk (keygen) and data/bfv.snark (prove)
    assert os.path.getsize(workdir / "data" / "bfv.vk") == 72 + 64 * 365
    r = _run(workdir, "--input", "bfv/bfv.in", "verify")
    assert r.returncode == 0 and "Snark verified successfully" in r.stdout and "Verification time" in r.stdout, r.stdout + r.stderr
    snark = bytearray(open(workdir / "data" / "bfv.snark", "rb").read())
    snark[16 + 32 * 5121 + 64 * 10 + 3] ^= 1                    # one bit of an advice commitment
    open(workdir / "data" / "bfv.snark", "wb").write(snark)
    r = _run(workdir, "--input", "bfv/bfv.in", "verify")
    assert r.returncode == 1 and "REJECTED" in r.stdout
    snark[16 + 32 * 5121 + 64 * 10 + 3] ^= 1
    snark[16 + 32 * 7] ^= 2                                      # a public input (a pk0 coefficient)
    open(workdir / "data" / "bfv.snark", "wb").write(snark)
    r = _run(workdir, "--input", "bfv/bfv.in", "verify")
    assert r.returncode == 1 and "REJECTED" in r.stdout


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
