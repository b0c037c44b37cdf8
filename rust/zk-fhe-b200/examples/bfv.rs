//! The `bfv` example entrypoint on the B200 path: same command line and file contract as the reference
//! (`cargo run --example bfv -- --name bfv -k 13 --input bfv/bfv.in {mock|keygen|prove|verify}`, README.md:14-54),
//! same input schema (examples/bfv.rs:50-61), the circuit written against the same `Poly` / `PolyChip` calls in the
//! same order (that order IS the cell layout).  The driver below stands where halo2-scaffold's `run_eth` stands.
//! SOURCE ONLY (no Rust toolchain in the build image); host/bfv.cpp is the compiled twin of this file.
use clap::Parser;
use serde::Deserialize;
use std::rc::Rc;
use zk_fhe::halo2_shim::{AssignedValue, Context, Fr, GateChip, RangeChip, RlcChip, Witness, CTX_GATE, CTX_PHASE0, CTX_RLC};
use zk_fhe::poly::Poly;
use zk_fhe::poly_chip::PolyChip;
use zk_fhe::prover::{self, ProvingKey, Prover, TRANSCRIPT_POSEIDON};
use zk_fhe::{set_device, Device};

const N: usize = 1024;          // examples/bfv.rs:27-30
const Q: u64 = 536870909;
const T: u64 = 7;
const B: u64 = 19;

#[derive(Clone, Debug, Deserialize)]
pub struct CircuitInput {
    pub pk0: Vec<String>, pub pk1: Vec<String>, pub m: Vec<String>, pub u: Vec<String>, pub e0: Vec<String>,
    pub e1: Vec<String>, pub c0: Vec<String>, pub c1: Vec<String>, pub cyclo: Vec<String>,
}

/// Everything phase 0 assigns and phase 1 needs again.
struct Assigned {
    pk: [PolyChip<Fr>; 2], pk_u: [PolyChip<Fr>; 2], quotient: [PolyChip<Fr>; 2], q_times_cyclo: [PolyChip<Fr>; 2],
    remainder: [PolyChip<Fr>; 2], expected_c: [PolyChip<Fr>; 2], e: [PolyChip<Fr>; 2],
    m: PolyChip<Fr>, u: PolyChip<Fr>, cyclo: PolyChip<Fr>, delta: AssignedValue<Fr>,
}

/// Phase 0 (examples/bfv.rs:70-165): parse, assign, expose, precompute pk * u, its quotient / remainder by x^N + 1.
fn phase0(ctx: &mut Context<Fr>, input: CircuitInput, make_public: &mut Vec<AssignedValue<Fr>>) -> Assigned {
    let parse = |v: Vec<String>| Poly::from_string(v, Q);
    let (pk0, pk1, m, u) = (parse(input.pk0), parse(input.pk1), parse(input.m), parse(input.u));
    let (e0, e1, c0, c1, cyclo) = (parse(input.e0), parse(input.e1), parse(input.c0), parse(input.c1), parse(input.cyclo));
    for p in [&pk0, &pk1, &m, &u, &e0, &e1, &c0, &c1] {
        assert_eq!(p.deg(), N - 1);
    }
    assert_eq!(cyclo.deg(), N);
    // products and divisions first (they only read), assignments in the reference's order afterwards
    let prod = [pk0.mul(&u), pk1.mul(&u)];
    let mut reduced = [Poly::from_big_int(prod[0].coefficients(), prod[0].max_bits), Poly::from_big_int(prod[1].coefficients(), prod[1].max_bits)];
    let halves: Vec<(Poly, Poly)> = reduced.iter_mut().map(|p| p.reduce_by_modulus(Q).divide_by_cyclo(&cyclo, Q)).collect();
    let qc = [halves[0].0.mul(&cyclo), halves[1].0.mul(&cyclo)];

    let pk = [PolyChip::from_poly(pk0, ctx), PolyChip::from_poly(pk1, ctx)];
    let m = PolyChip::from_poly(m, ctx);
    let u = PolyChip::from_poly(u, ctx);
    let e = [PolyChip::from_poly(e0, ctx), PolyChip::from_poly(e1, ctx)];
    let expected_c = [PolyChip::from_poly(c0, ctx), PolyChip::from_poly(c1, ctx)];
    let cyclo = PolyChip::from_poly(cyclo, ctx);
    let delta = ctx.load_constant(Q / T);
    for p in [&pk[0], &pk[1], &expected_c[0], &expected_c[1], &cyclo] {
        p.to_public(make_public);
        p.register_public(ctx);
    }
    let [prod0, prod1] = prod;
    let pk_u = [PolyChip::from_poly(prod0, ctx), PolyChip::from_poly(prod1, ctx)];
    let mut it = halves.into_iter();
    let (q0, r0) = it.next().unwrap();
    let (q1, r1) = it.next().unwrap();
    let quotient = [PolyChip::from_poly(q0, ctx), PolyChip::from_poly(q1, ctx)];
    let [qc0, qc1] = qc;
    let q_times_cyclo = [PolyChip::from_poly(qc0, ctx), PolyChip::from_poly(qc1, ctx)];
    let remainder = [PolyChip::from_poly(r0, ctx), PolyChip::from_poly(r1, ctx)];
    Assigned { pk, pk_u, quotient, q_times_cyclo, remainder, expected_c, e, m, u, cyclo, delta }
}

/// Phase 1, the callback (examples/bfv.rs:172-301): every constraint, in the reference's order.
fn phase1(a: &Assigned, gate_ctx: &mut Context<Fr>, rlc_ctx: &mut Context<Fr>) {
    let (range, rlc) = (RangeChip::<Fr>::default(), RlcChip::<Fr>::default());
    let gate: &GateChip<Fr> = &range.gate;
    a.e[0].constrain_coefficients_in_range(gate_ctx, &range, B, Q);
    a.e[1].constrain_coefficients_in_range(gate_ctx, &range, B, Q);
    a.u.constrain_from_distribution_chi_key(gate_ctx, gate, Q - 1);
    a.m.constrain_coefficients_in_range(gate_ctx, &range, T / 2, Q);
    for h in 0..2 {
        a.pk[h].constrain_mul(a.u.clone(), a.pk_u[h].clone(), gate_ctx, rlc_ctx, &rlc);
        let reduced = a.pk_u[h].reduce_by_modulo(gate_ctx, &range, Q);
        a.quotient[h].constrain_coefficients_in_modulus_field(gate_ctx, &range, Q);
        a.remainder[h].constrain_coefficients_in_modulus_field(gate_ctx, &range, Q);
        let pk_u = reduced.reduce_by_cyclo(a.cyclo.clone(), a.quotient[h].clone(), a.q_times_cyclo[h].clone(), a.remainder[h].clone(),
                                           &range, gate_ctx, rlc_ctx, &rlc, Q);
        let sum = if h == 0 {
            let m_delta = a.m.scalar_mul(gate_ctx, &a.delta, gate);
            pk_u.add(gate_ctx, m_delta, gate).add(gate_ctx, a.e[0].clone(), gate)
        } else {
            pk_u.add(gate_ctx, a.e[1].clone(), gate)
        };
        sum.reduce_by_modulo(gate_ctx, &range, Q).constrain_equality(gate_ctx, a.expected_c[h].clone(), gate);
    }
}

#[derive(Parser)]
struct Cli {
    #[arg(long, default_value = "bfv")] name: String,
    #[arg(short = 'k', long = "degree", default_value_t = 13)] degree: u32,
    #[arg(long)] input: Option<String>,
    #[arg(long, default_value = "configs")] config_path: String,
    #[arg(long, default_value = "data")] data_path: String,
    /// params file written by `bfv setup` (g, g_lagrange, [tau]_2); there is deliberately no silent test fallback
    #[arg(long)] srs: Option<String>,
    command: String,
}

fn main() {
    let cli = Cli::parse();
    let dev = Rc::new(Device::new(0));
    set_device(dev.clone());
    let read_input = || -> CircuitInput {
        let path = format!("{}/{}", cli.data_path, cli.input.as_ref().expect("--input"));
        serde_json::from_str(&std::fs::read_to_string(path).expect("input file")).expect("bfv.in schema")
    };
    let srs_path = cli.srs.clone().unwrap_or_else(|| format!("params/kzg_bn254_{}.srs", cli.degree));
    let load_srs = || -> [u8; 128] {
        let raw = std::fs::read(&srs_path).expect("params file: run `bfv setup` first");
        let n = 1usize << cli.degree;
        assert!(raw.len() == 16 + 128 * n + 128 && &raw[..8] == b"ZKFHESRS");
        dev.load_srs(cli.degree, &raw[16..16 + 64 * n], &raw[16 + 64 * n..16 + 128 * n]);
        raw[16 + 128 * n..].try_into().unwrap()
    };
    let synthesize = |wit: &Rc<Witness>, input: CircuitInput, gamma: Option<[u8; 32]>, prover: Option<&Prover>| {
        let mut make_public = vec![];
        let mut ctx0 = Context::<Fr>::new(wit.clone(), CTX_PHASE0);
        let assigned = phase0(&mut ctx0, input, &mut make_public);
        let gamma = match prover {
            Some(p) => p.phase0(wit),                      // commits phase 0, squeezes the challenge
            None => gamma.unwrap(),
        };
        wit.set_challenge(&gamma);
        phase1(&assigned, &mut Context::new(wit.clone(), CTX_GATE), &mut Context::new(wit.clone(), CTX_RLC));
    };
    match cli.command.as_str() {
        "mock" => {
            let wit = Witness::new(dev.clone(), 8, true);
            synthesize(&wit, read_input(), Some([7u8; 32]), None);
            wit.mock();
            println!("Mock prover: all constraints satisfied");
        }
        "keygen" => {
            load_srs();
            let wit = Witness::new(dev.clone(), 8, true);
            synthesize(&wit, read_input(), Some([1u8; 32]), None);
            let pk = ProvingKey::keygen(dev.clone(), &wit, cli.degree, 109);
            std::fs::write(format!("{}/{}.json", cli.config_path, cli.name), pk.pinning_json()).unwrap();
            std::fs::write(format!("{}/{}.vk", cli.data_path, cli.name), pk.vk_bytes()).unwrap();
            std::fs::write(format!("{}/{}.pk", cli.data_path, cli.name), pk.to_bytes()).unwrap();
        }
        "prove" => {
            load_srs();
            let pk = ProvingKey::from_bytes(dev.clone(), &std::fs::read(format!("{}/{}.pk", cli.data_path, cli.name)).expect("run keygen first"));
            let mut seed = [0u8; 32];
            std::io::Read::read_exact(&mut std::fs::File::open("/dev/urandom").unwrap(), &mut seed).unwrap();
            let prover = Prover::begin(dev.clone(), &pk, &seed, TRANSCRIPT_POSEIDON);
            let wit = Witness::new(dev.clone(), 8, false);
            let start = std::time::Instant::now();
            synthesize(&wit, read_input(), None, Some(&prover));
            let proof = prover.finish(&wit);
            println!("Proving time: {:?} ({} proof bytes)", start.elapsed(), proof.len());
            std::fs::write(format!("{}/{}.proof", cli.data_path, cli.name), proof).unwrap();
        }
        "verify" => {
            let s_g2 = load_srs();
            let vk = std::fs::read(format!("{}/{}.vk", cli.data_path, cli.name)).expect("run keygen first");
            let proof = std::fs::read(format!("{}/{}.proof", cli.data_path, cli.name)).expect("run prove first");
            let instances = std::fs::read(format!("{}/{}.instances", cli.data_path, cli.name)).expect("instances");
            let start = std::time::Instant::now();
            let ok = prover::verify(&dev, &vk, &instances, &proof, &s_g2, TRANSCRIPT_POSEIDON);
            println!("Verification time: {:?}", start.elapsed());
            assert!(ok, "Snark REJECTED");
            println!("Snark verified successfully");
        }
        other => panic!("unknown command {other}"),
    }
}
