// The `bfv` example entrypoint (reference examples/bfv.rs:306-312 + halo2-scaffold's `run_eth`):
//
//   bfv --name bfv -k 13 --input bfv/bfv.in {mock|keygen|prove|verify}
//       [--config-path configs] [--data-path data] [--unusable-rows 109]
//
// Same CLI and file contract as the reference's README.md:14-54: the input is read from
// <data-path>/<input> (nine arrays of decimal strings, examples/bfv.rs:50-61), keygen writes the
// pinning to <config-path>/<name>.json (schema of configs/bfv.json), prove writes
// <data-path>/<name>.snark and prints the proving time.
// <data-path>/<name>.snark (public instances + proof) and prints the proving time, keygen also writes
// <data-path>/<name>.vk, and verify reads the two files back and prints the verification time.
// keygen also writes <data-path>/<name>.pk, which prove reads back (README.md:38: keygen once, prove many).
//
// Commitment key: `--srs <file>` (default params/kzg_bn254_<k>.srs, the path halo2-scaffold's gen_srs reads) holds
// g, g_lagrange and [tau]_2; `bfv -k <k> setup` writes one from a trapdoor drawn from the OS and then forgotten.
// Without a file the tool REFUSES to run unless `--insecure-test-srs` is given: that flag uses the public trapdoor of the
// reference's own fallback (`ParamsKZG::setup` from `ChaCha20Rng` seed 0, zkfhe_reference_test_tau)
// (anyone can forge proofs against it) and exists for tests and benchmarks only.  The reference falls back to such a
// test setup silently; this tool makes the caller say so.
// The proof / pk / vk / snark / srs formats are this implementation's own.
#include <atomic>
#include <chrono>
#include <memory>
#include <thread>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>

#include "zk_fhe.hpp"
#include "../csrc/host_ff.h"

using namespace zkfhe;

static CircuitInput parse_input(const std::string& path) {
    std::ifstream f(path);
    if (!f) throw Error(ZKFHE_ERR_ARG, "cannot open input file " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    const std::string s = ss.str();
    CircuitInput in;
    size_t i = 0;
    auto skip = [&] { while (i < s.size() && (s[i] == ' ' || s[i] == '\n' || s[i] == '\r' || s[i] == '\t' || s[i] == ',')) i++; };
    auto str = [&]() {
        if (s[i] != '"') throw Error(ZKFHE_ERR_ARG, "input JSON: expected a string at offset " + std::to_string(i));
        size_t j = s.find('"', i + 1);
        if (j == std::string::npos) throw Error(ZKFHE_ERR_ARG, "input JSON: unterminated string");
        std::string out = s.substr(i + 1, j - i - 1);
        i = j + 1;
        return out;
    };
    skip();
    if (i >= s.size() || s[i] != '{') throw Error(ZKFHE_ERR_ARG, "input JSON: expected an object");
    i++;
    for (;;) {
        skip();
        if (i >= s.size()) throw Error(ZKFHE_ERR_ARG, "input JSON: unexpected end");
        if (s[i] == '}') break;
        std::string key = str();
        skip();
        if (s[i] != ':') throw Error(ZKFHE_ERR_ARG, "input JSON: expected ':'");
        i++;
        skip();
        if (s[i] != '[') throw Error(ZKFHE_ERR_ARG, "input JSON: field `" + key + "` is not an array");
        i++;
        std::vector<std::string> vals;
        for (;;) {
            skip();
            if (s[i] == ']') { i++; break; }
            vals.push_back(str());
        }
        in[key] = std::move(vals);
    }
    return in;
}

static void fr_mont_from_u64(uint64_t v, uint8_t out[32]) {
    // v * 2^256 mod r by 256 modular doublings (host, once per run)
    static const uint64_t R[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    uint64_t a[4] = {v, 0, 0, 0};
    for (int k = 0; k < 256; k++) {
        uint64_t c = 0;
        for (int i = 0; i < 4; i++) { uint64_t t = (a[i] << 1) | c; c = a[i] >> 63; a[i] = t; }
        bool ge = c;
        if (!ge) {
            ge = true;
            for (int i = 3; i >= 0; i--) { if (a[i] != R[i]) { ge = a[i] > R[i]; break; } }
        }
        if (ge) {
            unsigned __int128 b = 0;
            for (int i = 0; i < 4; i++) { unsigned __int128 d = (unsigned __int128)a[i] - R[i] - b; a[i] = (uint64_t)d; b = (d >> 64) & 1; }
        }
    }
    memcpy(out, a, 32);
}

static std::string slurp(const std::string& path, const char* hint) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error(ZKFHE_ERR_ARG, "cannot open " + path + hint);
    std::stringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// params file: "ZKFHESRS" | u32 k | u32 reserved | g (n x 64) | g_lagrange (n x 64) | [tau]_2 (128); Montgomery coordinates
struct Srs { uint32_t k = 0; std::string g, gl; uint8_t s_g2[128]; };
static Srs read_srs(const std::string& path, uint32_t k) {
    const std::string raw = slurp(path, " (run `bfv -k <k> setup` first, or pass --srs <file> / --insecure-test-srs)");
    const size_t n = (size_t)1 << k;
    if (raw.size() != 16 + 128 * n + 128 || memcmp(raw.data(), "ZKFHESRS", 8)) throw Error(ZKFHE_ERR_ARG, path + " is not a zkfhe params file for this k");
    Srs s;
    memcpy(&s.k, raw.data() + 8, 4);
    if (s.k != k) throw Error(ZKFHE_ERR_ARG, path + " holds parameters for another k");
    s.g = raw.substr(16, 64 * n);
    s.gl = raw.substr(16 + 64 * n, 64 * n);
    memcpy(s.s_g2, raw.data() + 16 + 128 * n, 128);
    return s;
}

static void os_random(uint8_t* out, size_t len) {
    std::ifstream ur("/dev/urandom", std::ios::binary);
    ur.read((char*)out, (std::streamsize)len);
    if (!ur.good() || (size_t)ur.gcount() != len) throw Error(ZKFHE_ERR_STATE, "cannot read " + std::to_string(len) + " bytes from /dev/urandom");
}

int main(int argc, char** argv) {
    std::string name = "bfv", input, config_path = "configs", data_path = "data", cmd, srs_path;
    uint32_t k = 13, unusable = 109;
    int transcript = host::TRANSCRIPT_POSEIDON;
    bool insecure_srs = false;
    uint32_t repeat = 0, n_streams = 8;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--name") name = next();
        else if (a == "-k" || a == "--degree") k = (uint32_t)std::stoul(next());
        else if (a == "--input") input = next();
        else if (a == "--config-path") config_path = next();
        else if (a == "--data-path") data_path = next();
        else if (a == "--unusable-rows") unusable = (uint32_t)std::stoul(next());
        else if (a == "--srs") srs_path = next();
        else if (a == "--insecure-test-srs") insecure_srs = true;
        else if (a == "--repeat") repeat = (uint32_t)std::stoul(next());
        else if (a == "--streams") n_streams = (uint32_t)std::stoul(next());
        else if (a == "--transcript") { std::string t = next(); transcript = t == "blake2b" ? host::TRANSCRIPT_BLAKE2B : host::TRANSCRIPT_POSEIDON; }
        else if (a == "mock" || a == "keygen" || a == "prove" || a == "verify" || a == "setup") cmd = a;
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (cmd.empty() || (input.empty() && cmd != "setup" && cmd != "verify")) {
        fprintf(stderr, "usage: bfv --name <name> -k <degree> --input <file under data/> {mock|keygen|prove|verify}\n"
                        "       bfv -k <degree> [--srs <file>] setup\n"
                        "       options: --srs <file> | --insecure-test-srs, --config-path, --data-path, --unusable-rows, --transcript poseidon|blake2b\n"
                        "       prove --repeat <R> [--streams <S>]: prove the input R more times on S proof streams (host threads) and print proofs/s\n");
        return 2;
    }
    if (srs_path.empty()) srs_path = "params/kzg_bn254_" + std::to_string(k) + ".srs";
    try {
        uint8_t s_g2[128];
        // the commitment key for keygen / prove (loaded on the device) and [tau]_2 for verify
        auto load_srs = [&](Device* dev) {
            if (insecure_srs) {
                fprintf(stderr, "WARNING: --insecure-test-srs: the KZG trapdoor is a PUBLIC constant (the reference's own fallback, "
                                "ParamsKZG::setup from ChaCha20Rng seed 0); proofs against this key can be forged by anyone.  "
                                "Tests and benchmarks only.\n");
                uint8_t tau[32];
                if (zkfhe_reference_test_tau(tau, nullptr) != ZKFHE_OK) throw Error(ZKFHE_ERR_ARG, "reference_test_tau failed");
                if (dev) dev->check(zkfhe_srs_setup(dev->raw(), k, tau, nullptr, nullptr));
                if (zkfhe_srs_g2(tau, s_g2) != ZKFHE_OK) throw Error(ZKFHE_ERR_ARG, "srs_g2 failed");
                return;
            }
            Srs srs = read_srs(srs_path, k);
            if (dev) dev->check(zkfhe_load_srs(dev->raw(), k, (const uint8_t*)srs.g.data(), (const uint8_t*)srs.gl.data()));
            memcpy(s_g2, srs.s_g2, 128);
        };
        if (cmd == "setup") {
            // halo2 `ParamsKZG::setup` shape from a trapdoor that never leaves this function.  Still a single-party
            // setup: whoever runs it could have kept tau.  Use a ceremony transcript for anything real.
            uint8_t seed[64], tau[32];
            if (insecure_srs) { if (zkfhe_reference_test_tau(tau, nullptr) != ZKFHE_OK) throw Error(ZKFHE_ERR_ARG, "reference_test_tau failed"); }
            else {
                os_random(seed, 64);
                host::Fr t = host::from_uniform_bytes(seed);
                memcpy(tau, t.l, 32);
            }
            Device dev(0);
            const size_t n = (size_t)1 << k;
            std::string g(64 * n, '\0'), gl(64 * n, '\0');
            dev.check(zkfhe_srs_setup(dev.raw(), k, tau, (uint8_t*)&g[0], (uint8_t*)&gl[0]));
            if (zkfhe_srs_g2(tau, s_g2) != ZKFHE_OK) throw Error(ZKFHE_ERR_ARG, "srs_g2 failed");
            memset(tau, 0, sizeof tau);
            memset(seed, 0, sizeof seed);
            std::ofstream out(srs_path, std::ios::binary);
            if (!out) throw Error(ZKFHE_ERR_ARG, "cannot write " + srs_path + " (does the directory exist?)");
            const uint32_t head[2] = {k, 0};
            out.write("ZKFHESRS", 8);
            out.write((const char*)head, 8);
            out.write(g.data(), (std::streamsize)g.size());
            out.write(gl.data(), (std::streamsize)gl.size());
            out.write((const char*)s_g2, 128);
            printf("setup: wrote %s (k = %u, %zu bytes)\n", srs_path.c_str(), k, (size_t)(16 + 128 * n + 128));
            return 0;
        }
        if (cmd == "verify") {
            // reference README.md:48-54: reads the .vk written by keygen and the .snark written by prove
            const std::string vk = slurp(data_path + "/" + name + ".vk", " (run keygen first)");
            const std::string snark = slurp(data_path + "/" + name + ".snark", " (run prove first)");
            if (snark.size() < 16 || memcmp(snark.data(), "ZKFHESN1", 8)) throw Error(ZKFHE_ERR_ARG, "not a zkfhe snark file");
            uint32_t n_inst, kind;
            memcpy(&n_inst, snark.data() + 8, 4);
            memcpy(&kind, snark.data() + 12, 4);
            if (snark.size() < 16 + 32 * (size_t)n_inst) throw Error(ZKFHE_ERR_ARG, "snark file truncated");
            Device dev(0);
            load_srs(nullptr);
            int ok = 0;
            auto t0 = std::chrono::steady_clock::now();
            dev.check(zkfhe_verify(dev.raw(), (const uint8_t*)vk.data(), vk.size(), (const uint8_t*)snark.data() + 16, n_inst,
                                   (const uint8_t*)snark.data() + 16 + 32 * (size_t)n_inst, snark.size() - 16 - 32 * (size_t)n_inst,
                                   s_g2, (int)kind, &ok));
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            printf("Verification time: %.3f ms\n", ms);
            if (!ok) {
                printf("Snark REJECTED: %s\n", zkfhe_last_error(dev.raw()));
                return 1;
            }
            printf("Snark verified successfully\n");
            return 0;
        }
        CircuitInput in = parse_input(data_path + "/" + input);
        Device dev(0);
        uint8_t gamma[32];
        if (cmd == "mock") {
            BfvCircuit circ(dev, BfvParams(), 8, /*record=*/true);
            circ.phase0(in);
            fr_mont_from_u64(0x9E3779B97F4A7C15ULL, gamma);        // any fixed non-degenerate challenge
            circ.phase1(gamma);
            circ.builder().mock();                                 // throws on any violated constraint
            printf("Mock prover: all constraints satisfied\n");
            return 0;
        }
        load_srs(&dev);
        const std::string pk_path = data_path + "/" + name + ".pk";
        if (cmd == "keygen") {
            // README.md:28-38: the circuit is synthesised on the given input (bfv_empty.in) in recording mode
            BfvCircuit kg(dev, BfvParams(), 8, /*record=*/true);
            kg.phase0(in);
            fr_mont_from_u64(1, gamma);
            kg.phase1(gamma);
            zkfhe_pk* pk = nullptr;
            dev.check(zkfhe_keygen(kg.builder().raw(), k, unusable, &pk));
            size_t need = 0;
            zkfhe_pk_pinning_json(pk, nullptr, 0, &need);
            std::string pin(need, '\0');
            zkfhe_pk_pinning_json(pk, &pin[0], need, nullptr);
            pin.resize(need - 1);
            std::ofstream(config_path + "/" + name + ".json") << pin << "\n";
            size_t vk_len = 0, pk_len = 0;
            dev.check(zkfhe_vk_export(pk, nullptr, 0, &vk_len));
            std::string vk(vk_len, '\0');
            dev.check(zkfhe_vk_export(pk, (uint8_t*)&vk[0], vk_len, nullptr));
            std::ofstream(data_path + "/" + name + ".vk", std::ios::binary).write(vk.data(), (std::streamsize)vk.size());
            dev.check(zkfhe_pk_export(pk, nullptr, 0, &pk_len));
            std::string pkb(pk_len, '\0');
            dev.check(zkfhe_pk_export(pk, (uint8_t*)&pkb[0], pk_len, nullptr));
            std::ofstream(pk_path, std::ios::binary).write(pkb.data(), (std::streamsize)pkb.size());
            printf("keygen: wrote %s/%s.json, %s/%s.vk and %s (%zu bytes)\n", config_path.c_str(), name.c_str(), data_path.c_str(),
                   name.c_str(), pk_path.c_str(), pk_len);
            zkfhe_pk_free(pk);
            return 0;
        }
        // prove: reads the proving key keygen wrote (README.md:38)
        zkfhe_pk* pk = nullptr;
        {
            const std::string pkb = slurp(pk_path, " (run keygen first)");
            auto t0 = std::chrono::steady_clock::now();
            dev.check(zkfhe_pk_import(dev.raw(), (const uint8_t*)pkb.data(), pkb.size(), &pk));
            printf("Proving key loaded in %.1f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count());
        }
        // twice: the first pass also pays one-time device allocations; both times are printed
        BfvCircuit circ(dev);
        zkfhe_prover* pr = nullptr;
        uint8_t seed[32];
        os_random(seed, 32);                                       // the reference seeds its RNG from OS entropy too
        dev.check(zkfhe_prove_begin(dev.raw(), pk, seed, transcript, &pr));
        uint8_t* proof = nullptr;
        size_t len = 0;
        for (int pass = 0; pass < 2; pass++) {
            if (proof) { zkfhe_proof_free(proof); proof = nullptr; }
            circ.builder().reset();
            os_random(seed, 32);
            dev.check(zkfhe_prove_reset(pr, seed));
            auto t0 = std::chrono::steady_clock::now();
            circ.phase0(in);
            dev.check(zkfhe_prove_phase0(pr, circ.builder().raw(), gamma));
            circ.phase1(gamma);
            dev.check(zkfhe_prove_finish(pr, circ.builder().raw(), &proof, &len));
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            printf("Proving time%s: %.3f ms (%zu proof bytes)\n", pass ? "" : " (cold, incl. one-time allocations)", ms, len);
        }
        if (repeat) {
            // Serving shape: S proofs in flight on one GPU, one host thread + CUDA stream + witness / prover buffers each,
            // the commitment-key tables and the proving key shared (what bench.py does from Python, here from C++ threads).
            if (n_streams < 1) n_streams = 1;
            std::vector<std::unique_ptr<Device>> devs;
            std::vector<std::unique_ptr<BfvCircuit>> circs;
            std::vector<zkfhe_prover*> provers(n_streams, nullptr);
            for (uint32_t i = 0; i < n_streams; i++) {
                devs.emplace_back(new Device(0));
                devs[i]->check(zkfhe_share_srs(devs[i]->raw(), dev.raw()));
                devs[i]->check(zkfhe_set_blocking_sync(devs[i]->raw(), 1));
                circs.emplace_back(new BfvCircuit(*devs[i]));
                os_random(seed, 32);
                devs[i]->check(zkfhe_prove_begin(devs[i]->raw(), pk, seed, transcript, &provers[i]));
            }
            std::atomic<uint32_t> next{0};
            std::atomic<int> failed{0};
            auto worker = [&](uint32_t i, uint32_t total) {
                try {
                    uint8_t g[32], sd[32];
                    while (next.fetch_add(1) < total) {
                        circs[i]->builder().reset();
                        os_random(sd, 32);
                        devs[i]->check(zkfhe_prove_reset(provers[i], sd));
                        circs[i]->phase0(in);
                        devs[i]->check(zkfhe_prove_phase0(provers[i], circs[i]->builder().raw(), g));
                        circs[i]->phase1(g);
                        uint8_t* pf = nullptr;
                        size_t pl = 0;
                        devs[i]->check(zkfhe_prove_finish(provers[i], circs[i]->builder().raw(), &pf, &pl));
                        zkfhe_proof_free(pf);
                    }
                } catch (const Error& e) {
                    fprintf(stderr, "stream %u: error (%d): %s\n", i, e.code, e.what());
                    failed = 1;
                }
            };
            auto run = [&](uint32_t total) {
                next = 0;
                std::vector<std::thread> th;
                for (uint32_t i = 0; i < n_streams; i++) th.emplace_back(worker, i, total);
                for (auto& t : th) t.join();
            };
            run(2 * n_streams);                                    // warm-up: one-time allocations of every stream
            auto t0 = std::chrono::steady_clock::now();
            run(repeat);
            const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            printf("Throughput: %.2f proofs/s (%u proofs from host strings to proof bytes, %u proof streams, %.3f s)\n", repeat / sec, repeat,
                   n_streams, sec);
            for (uint32_t i = 0; i < n_streams; i++) zkfhe_prover_free(provers[i]);
            circs.clear();
            devs.clear();
            if (failed) return 1;
        }
        // .snark = "ZKFHESN1" | u32 instances | u32 transcript kind | instances (canonical 32-byte LE) | proof
        uint32_t info[16];
        dev.check(zkfhe_pk_info(pk, info));
        const uint32_t n_inst = info[12], kind = (uint32_t)transcript;
        std::vector<host::Fr> inst(n_inst);
        if (n_inst) dev.check(zkfhe_witness_download(circ.builder().raw(), 4, (uint8_t*)inst.data()));
        for (auto& v : inst) v = host::from_mont(v);
        std::ofstream out(data_path + "/" + name + ".snark", std::ios::binary);
        out.write("ZKFHESN1", 8);
        out.write((const char*)&n_inst, 4);
        out.write((const char*)&kind, 4);
        out.write((const char*)inst.data(), (std::streamsize)(32 * (size_t)n_inst));
        out.write((const char*)proof, (std::streamsize)len);
        zkfhe_proof_free(proof);
        zkfhe_prover_free(pr);
        zkfhe_pk_free(pk);
        return 0;
    } catch (const Error& e) {
        fprintf(stderr, "error (%d): %s\n", e.code, e.what());
        return 1;
    }
}
