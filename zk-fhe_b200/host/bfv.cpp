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
// Differences, stated plainly: the proving key is rebuilt in-process (no .pk file); the SRS is
// the deterministic *test* setup every time (halo2-scaffold's gen_srs falls back to one too); the
// proof / vk / snark formats are this implementation's own.
#include <chrono>
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

int main(int argc, char** argv) {
    std::string name = "bfv", input, config_path = "configs", data_path = "data", cmd;
    uint32_t k = 13, unusable = 109;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() -> std::string { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "--name") name = next();
        else if (a == "-k" || a == "--degree") k = (uint32_t)std::stoul(next());
        else if (a == "--input") input = next();
        else if (a == "--config-path") config_path = next();
        else if (a == "--data-path") data_path = next();
        else if (a == "--unusable-rows") unusable = (uint32_t)std::stoul(next());
        else if (a == "mock" || a == "keygen" || a == "prove" || a == "verify") cmd = a;
        else { fprintf(stderr, "unknown argument %s\n", a.c_str()); return 2; }
    }
    if (cmd.empty() || input.empty()) {
        fprintf(stderr, "usage: bfv --name <name> -k <degree> --input <file under data/> {mock|keygen|prove|verify}\n");
        return 2;
    }
    try {
        uint8_t tau[32];
        fr_mont_from_u64(0x5EED5EED5EED5EEDULL, tau);              // deterministic TEST setup
        if (cmd == "verify") {
            // reference README.md:48-54: reads the .vk written by keygen and the .snark written by prove
            auto slurp = [](const std::string& path) {
                std::ifstream f(path, std::ios::binary);
                if (!f) throw Error(ZKFHE_ERR_ARG, "cannot open " + path + " (run keygen / prove first)");
                std::stringstream ss;
                ss << f.rdbuf();
                return ss.str();
            };
            const std::string vk = slurp(data_path + "/" + name + ".vk"), snark = slurp(data_path + "/" + name + ".snark");
            if (snark.size() < 16 || memcmp(snark.data(), "ZKFHESN1", 8)) throw Error(ZKFHE_ERR_ARG, "not a zkfhe snark file");
            uint32_t n_inst, kind;
            memcpy(&n_inst, snark.data() + 8, 4);
            memcpy(&kind, snark.data() + 12, 4);
            if (snark.size() < 16 + 32 * (size_t)n_inst) throw Error(ZKFHE_ERR_ARG, "snark file truncated");
            Device dev(0);
            uint8_t s_g2[128];
            dev.check(zkfhe_srs_g2(tau, s_g2));
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
        dev.check(zkfhe_srs_setup(dev.raw(), k, tau, nullptr, nullptr));
        // keygen always runs on an input of the same shape with all-zero values (README.md:31-36)
        CircuitInput zeros;
        for (auto& kv : in) zeros[kv.first] = std::vector<std::string>(kv.second.size(), "0");
        BfvCircuit kg(dev, BfvParams(), 8, /*record=*/true);
        kg.phase0(cmd == "keygen" ? in : zeros);
        fr_mont_from_u64(1, gamma);
        kg.phase1(gamma);
        zkfhe_pk* pk = nullptr;
        dev.check(zkfhe_keygen(kg.builder().raw(), k, unusable, &pk));
        size_t need = 0;
        zkfhe_pk_pinning_json(pk, nullptr, 0, &need);
        std::string pin(need, '\0');
        zkfhe_pk_pinning_json(pk, &pin[0], need, nullptr);
        pin.resize(need - 1);
        if (cmd == "keygen") {
            std::ofstream(config_path + "/" + name + ".json") << pin << "\n";
            size_t vk_len = 0;
            dev.check(zkfhe_vk_export(pk, nullptr, 0, &vk_len));
            std::string vk(vk_len, '\0');
            dev.check(zkfhe_vk_export(pk, (uint8_t*)&vk[0], vk_len, nullptr));
            std::ofstream(data_path + "/" + name + ".vk", std::ios::binary).write(vk.data(), (std::streamsize)vk.size());
            printf("keygen: wrote %s/%s.json and %s/%s.vk\n", config_path.c_str(), name.c_str(), data_path.c_str(), name.c_str());
            zkfhe_pk_free(pk);
            return 0;
        }
        // prove (twice: the first pass also pays one-time device allocations; both times are printed)
        BfvCircuit circ(dev);
        zkfhe_prover* pr = nullptr;
        uint8_t seed[32];
        {
            std::ifstream ur("/dev/urandom", std::ios::binary);   // the reference seeds its RNG from OS entropy too
            ur.read((char*)seed, 32);
        }
        dev.check(zkfhe_prove_begin(dev.raw(), pk, seed, 0, &pr));
        uint8_t* proof = nullptr;
        size_t len = 0;
        for (int pass = 0; pass < 2; pass++) {
            if (proof) { zkfhe_proof_free(proof); proof = nullptr; }
            circ.builder().reset();
            dev.check(zkfhe_prove_reset(pr, seed));
            auto t0 = std::chrono::steady_clock::now();
            circ.phase0(in);
            dev.check(zkfhe_prove_phase0(pr, circ.builder().raw(), gamma));
            circ.phase1(gamma);
            dev.check(zkfhe_prove_finish(pr, circ.builder().raw(), &proof, &len));
            double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            printf("Proving time%s: %.3f ms (%zu proof bytes)\n", pass ? "" : " (cold, incl. one-time allocations)", ms, len);
        }
        // .snark = "ZKFHESN1" | u32 instances | u32 transcript kind | instances (canonical 32-byte LE) | proof
        uint32_t info[16];
        dev.check(zkfhe_pk_info(pk, info));
        const uint32_t n_inst = info[12], kind = 0;
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
