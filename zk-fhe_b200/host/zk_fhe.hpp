// Host-side C++ mirror of the reference's Rust surface for the prove path, over the C ABI
// (include/zkfhe_b200.h).  The reference is compiled Rust and no Rust toolchain exists in this
// image, so the host side above the ABI is C++ with the reference's names, argument meaning and
// error behaviour:
//
//   zk_fhe::poly::Poly            src/poly.rs:9-13        -> zkfhe::Poly
//   zk_fhe::poly_chip::PolyChip   src/poly_chip.rs:19-23  -> zkfhe::PolyChip
//   halo2_base::Context (x3)      [upstream]              -> zkfhe::Builder (contexts 0,1,2 in HBM)
//   bfv_encryption_circuit        examples/bfv.rs:63-304  -> zkfhe::BfvCircuit::{phase0, phase1}
//
// Reference `assert!`/panic conditions become zkfhe::Error (never an abort across the ABI).
// All arithmetic runs on the GPU; nothing here touches coefficients except decimal parsing.
#pragma once
#include <cstdint>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/zkfhe_b200.h"

namespace zkfhe {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

class Device {   // one GPU context (stream); owns the SRS tables
public:
    explicit Device(int index = 0) {
        int rc = zkfhe_init(index, &ctx_);
        if (rc != ZKFHE_OK) throw Error(rc, "zkfhe_init failed: no usable CUDA device (there is no CPU fallback)");
    }
    ~Device() { zkfhe_destroy(ctx_); }
    Device(const Device&) = delete;
    Device& operator=(const Device&) = delete;
    zkfhe_ctx* raw() const { return ctx_; }
    void check(int rc) const {
        if (rc != ZKFHE_OK) throw Error(rc, zkfhe_last_error(ctx_));
    }
    void status() const { check(zkfhe_status(ctx_)); }

private:
    zkfhe_ctx* ctx_ = nullptr;
};

// ---- src/poly.rs --------------------------------------------------------------------------------------
class Poly {
public:
    // Poly::from_string (poly.rs:21-40): decimal strings, asserts coeff <= modulus
    static Poly from_string(const Device& dev, const std::vector<std::string>& coefficients, uint64_t modulus) {
        std::vector<uint64_t> v;
        v.reserve(coefficients.size());
        for (const auto& s : coefficients) {
            if (s.empty() || s[0] == '-') throw Error(ZKFHE_ERR_ARG, "negative or empty coefficient (the reference modulus is u64)");
            unsigned __int128 acc = 0;
            for (char ch : s) {
                if (ch < '0' || ch > '9') throw Error(ZKFHE_ERR_ARG, "called `Option::unwrap()` on a `None` value: parse_bytes (src/poly.rs:27)");
                acc = acc * 10 + (unsigned)(ch - '0');
                if (acc >> 64) throw Error(ZKFHE_ERR_ASSERT, "assertion failed: coeff <= modulus_bigint (src/poly.rs:28)");
            }
            v.push_back((uint64_t)acc);
        }
        zkfhe_poly* h = nullptr;
        dev.check(zkfhe_poly_from_u64(dev.raw(), v.data(), (uint32_t)v.size(), modulus, &h));
        return Poly(dev, h);
    }
    Poly(Poly&& o) noexcept : dev_(o.dev_), h_(o.h_) { o.h_ = nullptr; }
    Poly(const Poly&) = delete;
    ~Poly() { zkfhe_poly_free(h_); }

    size_t deg() const { return zkfhe_poly_len(h_) - 1; }
    uint64_t max_bits() const { return zkfhe_poly_max_bits(h_); }
    zkfhe_poly* raw() const { return h_; }

    Poly mul(const Poly& other) const {                      // poly.rs:75-103
        zkfhe_poly* h = nullptr;
        dev_->check(zkfhe_poly_mul(dev_->raw(), h_, other.h_, &h));
        return Poly(*dev_, h);
    }
    Poly reduce_by_modulus(uint64_t modulus) const {         // poly.rs:180-191
        zkfhe_poly* h = nullptr;
        dev_->check(zkfhe_poly_reduce_by_modulus(dev_->raw(), h_, modulus, &h));
        return Poly(*dev_, h);
    }
    std::pair<Poly, Poly> divide_by_cyclo(const Poly& cyclo, uint64_t modulus) const {   // poly.rs:113-177
        zkfhe_poly *q = nullptr, *r = nullptr;
        dev_->check(zkfhe_poly_divide_by_cyclo(dev_->raw(), h_, cyclo.h_, modulus, &q, &r));
        return {Poly(*dev_, q), Poly(*dev_, r)};
    }
    std::vector<uint64_t> coefficients_u256() const {        // 4 x u64 limbs per coefficient
        std::vector<uint64_t> out((size_t)zkfhe_poly_len(h_) * 4);
        dev_->check(zkfhe_poly_download(dev_->raw(), h_, out.data()));
        return out;
    }

private:
    Poly(const Device& d, zkfhe_poly* h) : dev_(&d), h_(h) {}
    const Device* dev_;
    zkfhe_poly* h_;
};

// ---- halo2-base contexts + src/poly_chip.rs ---------------------------------------------------------------
enum : uint32_t { CTX_PHASE0 = 0, CTX_GATE = 1, CTX_RLC = 2 };

class Builder {   // the three Contexts of the two-phase circuit, resident in HBM
public:
    Builder(const Device& dev, uint32_t lookup_bits = 8, bool record = false) : dev_(&dev) {
        dev.check(zkfhe_witness_new(dev.raw(), lookup_bits, &w_));
        if (record) dev.check(zkfhe_witness_set_recording(w_, 1));
    }
    ~Builder() { zkfhe_witness_free(w_); }
    Builder(const Builder&) = delete;
    zkfhe_witness* raw() const { return w_; }
    const Device& dev() const { return *dev_; }
    void reset() { dev_->check(zkfhe_witness_reset(w_)); }
    void set_challenge(const uint8_t gamma_fr[32]) { dev_->check(zkfhe_chip_set_challenge(w_, gamma_fr)); }
    zkfhe_cell load_constant(uint32_t ctx_id, uint64_t value) {
        zkfhe_cell c;
        dev_->check(zkfhe_chip_load_constant(w_, ctx_id, value, &c));
        return c;
    }
    uint64_t mock() {                                         // `mock` subcommand: throws Error(ZKFHE_ERR_UNSATISFIED)
        uint64_t n = 0, first = 0;
        dev_->check(zkfhe_witness_mock(w_, &n, &first));
        return n;
    }

private:
    const Device* dev_;
    zkfhe_witness* w_ = nullptr;
};

class PolyChip {
public:
    PolyChip() : b_(nullptr), ap_{} {}
    // PolyChip::from_poly (poly_chip.rs:27-42)
    static PolyChip from_poly(const Poly& poly, Builder& b, uint32_t ctx_id = CTX_PHASE0) {
        PolyChip p(b);
        b.dev().check(zkfhe_chip_from_poly(b.raw(), ctx_id, poly.raw(), &p.ap_));
        return p;
    }
    uint64_t max_num_bits() const { return ap_.max_num_bits; }
    size_t degree() const { return ap_.len - 1; }
    const zkfhe_assigned_poly& assigned() const { return ap_; }

    void to_public() const { chk(zkfhe_chip_to_public(b_->raw(), &ap_)); }                                  // :58-62
    void constrain_mul(const PolyChip& b, const PolyChip& c, uint32_t ctx_gate = CTX_GATE, uint32_t ctx_rlc = CTX_RLC) const {
        chk(zkfhe_chip_constrain_mul(b_->raw(), ctx_gate, ctx_rlc, &ap_, &b.ap_, &c.ap_));                  // :81-116
    }
    PolyChip add(const PolyChip& other, uint32_t ctx_gate = CTX_GATE) const {                               // :122-144
        PolyChip o(*b_);
        chk(zkfhe_chip_add(b_->raw(), ctx_gate, &ap_, &other.ap_, &o.ap_));
        return o;
    }
    PolyChip scalar_mul(const zkfhe_cell& scalar, uint64_t scalar_value, uint32_t ctx_gate = CTX_GATE) const {   // :150-174
        PolyChip o(*b_);
        chk(zkfhe_chip_scalar_mul(b_->raw(), ctx_gate, &ap_, &scalar, scalar_value, &o.ap_));
        return o;
    }
    PolyChip reduce_by_cyclo(const PolyChip& cyclo, const PolyChip& quotient, const PolyChip& quotient_times_cyclo,
                             const PolyChip& remainder, uint64_t modulus, uint32_t ctx_gate = CTX_GATE,
                             uint32_t ctx_rlc = CTX_RLC) const {                                                // :183-223
        PolyChip o(*b_);
        chk(zkfhe_chip_reduce_by_cyclo(b_->raw(), ctx_gate, ctx_rlc, &ap_, &cyclo.ap_, &quotient.ap_, &quotient_times_cyclo.ap_,
                                       &remainder.ap_, modulus, &o.ap_));
        return o;
    }
    PolyChip reduce_by_modulo(uint64_t modulus, uint32_t ctx_gate = CTX_GATE) const {                        // :226-252
        PolyChip o(*b_);
        chk(zkfhe_chip_reduce_by_modulo(b_->raw(), ctx_gate, &ap_, modulus, &o.ap_));
        return o;
    }
    void constrain_equality(const PolyChip& other, uint32_t ctx_gate = CTX_GATE) const {                     // :255-264
        chk(zkfhe_chip_constrain_equality(b_->raw(), ctx_gate, &ap_, &other.ap_));
    }
    void constrain_coefficients_in_range(uint64_t z, uint64_t y, uint32_t ctx_gate = CTX_GATE) const {       // :270-317
        chk(zkfhe_chip_constrain_coefficients_in_range(b_->raw(), ctx_gate, &ap_, z, y));
    }
    void constrain_from_distribution_chi_key(uint64_t z, uint32_t ctx_gate = CTX_GATE) const {               // :320-354
        chk(zkfhe_chip_constrain_from_distribution_chi_key(b_->raw(), ctx_gate, &ap_, z));
    }
    void constrain_coefficients_in_modulus_field(uint64_t modulus, uint32_t ctx_gate = CTX_GATE) const {     // :357-366
        chk(zkfhe_chip_constrain_coefficients_in_modulus_field(b_->raw(), ctx_gate, &ap_, modulus));
    }

private:
    explicit PolyChip(Builder& b) : b_(&b), ap_{} {}
    void chk(int rc) const { b_->dev().check(rc); }
    Builder* b_;
    zkfhe_assigned_poly ap_;
};

// ---- examples/bfv.rs --------------------------------------------------------------------------------------------
struct BfvParams {          // bfv.rs:27-30 (compile-time consts there)
    size_t N = 1024;
    uint64_t Q = 536870909, T = 7, B = 19;
    // RNS: one circuit per limb prime proves c0 = pk0*u + delta_i*m + e0 (mod q_i), delta_i = (Q_total / T) mod q_i
    uint64_t delta_override = 0;
    bool has_delta_override = false;
    uint64_t delta() const { return has_delta_override ? delta_override : Q / T; }     // bfv.rs:112
};

using CircuitInput = std::map<std::string, std::vector<std::string>>;   // bfv.rs:50-61: nine arrays of decimal strings

class BfvCircuit {
public:
    BfvCircuit(const Device& dev, BfvParams p = BfvParams(), uint32_t lookup_bits = 8, bool record = false)
        : dev_(dev), params_(p), builder_(dev, lookup_bits, record) {}
    Builder& builder() { return builder_; }

    // Phase 0 (bfv.rs:70-165): assignments + the off-circuit precomputation
    void phase0(const CircuitInput& in) {
        const uint64_t Q = params_.Q;
        const size_t N = params_.N;
        static const char* keys[9] = {"pk0", "pk1", "m", "u", "e0", "e1", "c0", "c1", "cyclo"};
        std::map<std::string, std::unique_ptr<Poly>> un;
        for (auto k : keys) {
            auto it = in.find(k);
            if (it == in.end()) throw Error(ZKFHE_ERR_ARG, std::string("input misses field `") + k + "`");
            un[k].reset(new Poly(Poly::from_string(dev_, it->second, Q)));                                   // :71-79
        }
        for (int i = 0; i < 8; i++)
            if (un[keys[i]]->deg() != N - 1) throw Error(ZKFHE_ERR_ASSERT, std::string("assertion failed: deg(") + keys[i] + ") == N - 1 (examples/bfv.rs:82-89)");
        if (un["cyclo"]->deg() != N) throw Error(ZKFHE_ERR_ASSERT, "assertion failed: deg(cyclo) == N (examples/bfv.rs:90)");
        auto assign = [&](const char* name, const Poly& p) { P_[name] = PolyChip::from_poly(p, builder_, CTX_PHASE0); };
        assign("pk0", *un["pk0"]); assign("pk1", *un["pk1"]); assign("m", *un["m"]); assign("u", *un["u"]);   // :101-109
        assign("e0", *un["e0"]); assign("e1", *un["e1"]); assign("expected_c0", *un["c0"]);
        assign("expected_c1", *un["c1"]); assign("cyclo", *un["cyclo"]);
        delta_ = builder_.load_constant(CTX_PHASE0, params_.delta());                                         // :115
        for (auto name : {"pk0", "pk1", "expected_c0", "expected_c1", "cyclo"}) P_[name].to_public();         // :118-122
        Poly pk0_u = un["pk0"]->mul(*un["u"]), pk1_u = un["pk1"]->mul(*un["u"]);                               // :131-132
        assign("pk0_u", pk0_u); assign("pk1_u", pk1_u);                                                        // :135-136
        Poly r0 = pk0_u.reduce_by_modulus(Q), r1 = pk1_u.reduce_by_modulus(Q);                                 // :139-140
        auto d0 = r0.divide_by_cyclo(*un["cyclo"], Q);                                                         // :143-146
        auto d1 = r1.divide_by_cyclo(*un["cyclo"], Q);
        Poly q0c = d0.first.mul(*un["cyclo"]), q1c = d1.first.mul(*un["cyclo"]);                               // :149-150
        assign("quotient_0", d0.first); assign("quotient_1", d1.first);                                        // :156-157
        assign("quotient_0_times_cyclo", q0c); assign("quotient_1_times_cyclo", q1c);                          // :160-161
        assign("remainder_0", d0.second); assign("remainder_1", d1.second);                                    // :164-165
        dev_.status();     // data-dependent reference asserts of phase 0 (one synchronisation)
    }

    // Phase 1, the callback (bfv.rs:172-301); gamma = the phase-0 challenge (Fr, Montgomery bytes)
    void phase1(const uint8_t gamma_fr[32]) {
        const uint64_t Q = params_.Q, T = params_.T, B = params_.B;
        builder_.set_challenge(gamma_fr);
        P_["e0"].constrain_coefficients_in_range(B, Q);                                                       // :189
        P_["e1"].constrain_coefficients_in_range(B, Q);                                                       // :190
        P_["u"].constrain_from_distribution_chi_key(Q - 1);                                                   // :201
        P_["m"].constrain_coefficients_in_range(T / 2, Q);                                                    // :210
        auto half = [&](const char* pk, const char* pk_u, const char* quot, const char* qtc, const char* rem) {
            P_[pk].constrain_mul(P_["u"], P_[pk_u]);                                                          // :215 / :264
            PolyChip red = P_[pk_u].reduce_by_modulo(Q);                                                      // :219 / :268
            P_[quot].constrain_coefficients_in_modulus_field(Q);                                              // :225 / :274
            P_[rem].constrain_coefficients_in_modulus_field(Q);                                               // :226 / :275
            return red.reduce_by_cyclo(P_["cyclo"], P_[quot], P_[qtc], P_[rem], Q);                           // :228 / :277
        };
        PolyChip pk0_u = half("pk0", "pk0_u", "quotient_0", "quotient_0_times_cyclo", "remainder_0");
        PolyChip m_delta = P_["m"].scalar_mul(delta_, params_.delta());                                       // :243
        PolyChip c0 = pk0_u.add(m_delta).add(P_["e0"]).reduce_by_modulo(Q);                                   // :247-255
        c0.constrain_equality(P_["expected_c0"]);                                                             // :259
        PolyChip pk1_u = half("pk1", "pk1_u", "quotient_1", "quotient_1_times_cyclo", "remainder_1");
        PolyChip c1 = pk1_u.add(P_["e1"]).reduce_by_modulo(Q);                                                // :292-296
        c1.constrain_equality(P_["expected_c1"]);                                                             // :300
        dev_.status();     // data-dependent asserts recorded by the phase-1 kernels (one synchronisation)
    }

private:
    const Device& dev_;
    BfvParams params_;
    Builder builder_;
    std::map<std::string, PolyChip> P_;
    zkfhe_cell delta_{};
};

}  // namespace zkfhe
