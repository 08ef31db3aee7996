//! CPU numbers of the reference's own prover for the two halves of BASELINE.json's metric, in the JSON shape of
//! `bench.py --impl reference` so the two can be laid side by side:
//!   M2  `PolynomialBatch::from_values` at 2^k rows x 135 columns, rate_bits 3, cap_height 4, no blinding
//!       (SplitMix64 columns seeded 0x5EED0001 + col — the same inputs bench.py uses)
//!   M1  `CircuitData::prove` of a 2^12-row circuit under `standard_recursion_config` (what every worker job of
//!       city_rollup_core_worker ends in, e.g. city_common_circuit/src/proof_minifier/pm_core.rs:151)
//! Usage: cargo run --release -- [log2_rows=16] [proofs=20]
//! rayon uses every host core (RAYON_NUM_THREADS to restrict); the core count is printed.
use plonky2::field::goldilocks_field::GoldilocksField;
use plonky2::field::polynomial::PolynomialValues;
use plonky2::field::types::Field;
use plonky2::fri::oracle::PolynomialBatch;
use plonky2::iop::witness::{PartialWitness, WitnessWrite};
use plonky2::plonk::circuit_builder::CircuitBuilder;
use plonky2::plonk::circuit_data::CircuitConfig;
use plonky2::plonk::config::PoseidonGoldilocksConfig;
use plonky2::util::timing::TimingTree;
use std::time::Instant;

type F = GoldilocksField;
type C = PoseidonGoldilocksConfig;
const D: usize = 2;
const P: u64 = 0xFFFF_FFFF_0000_0001;

/// SplitMix64 stream of tests/util.py::splitmix64 (values >= p reduced once)
fn column(seed: u64, n: usize) -> Vec<F> {
    (1..=n as u64)
        .map(|i| {
            let mut z = seed.wrapping_add(i.wrapping_mul(0x9E37_79B9_7F4A_7C15));
            z = (z ^ (z >> 30)).wrapping_mul(0xBF58_476D_1CE4_E5B9);
            z = (z ^ (z >> 27)).wrapping_mul(0x94D0_49BB_1331_11EB);
            z ^= z >> 31;
            F::from_canonical_u64(if z >= P { z - P } else { z })
        })
        .collect()
}

fn main() {
    let args: Vec<String> = std::env::args().collect();
    let log_rows: usize = args.get(1).and_then(|s| s.parse().ok()).unwrap_or(16);
    let proofs: usize = args.get(2).and_then(|s| s.parse().ok()).unwrap_or(20);
    let cores = std::thread::available_parallelism().map(|n| n.get()).unwrap_or(1);

    // ---- M2: one commit
    let n = 1usize << log_rows;
    let values: Vec<PolynomialValues<F>> = (0..135).map(|c| PolynomialValues::new(column(0x5EED_0001 + c as u64, n))).collect();
    let mut best = f64::MAX;
    for _ in 0..3 {
        let v = values.clone();
        let t0 = Instant::now();
        let batch = PolynomialBatch::<F, C, D>::from_values(v, 3, false, 4, &mut TimingTree::default(), None);
        best = best.min(t0.elapsed().as_secs_f64());
        std::hint::black_box(&batch.merkle_tree.cap);
    }

    // ---- M1: proofs of a 2^12-row circuit (arithmetic chain + Poseidon hashes, padded by the builder)
    let config = CircuitConfig::standard_recursion_config();
    let mut builder = CircuitBuilder::<F, D>::new(config);
    let x = builder.add_virtual_target();
    let mut acc = x;
    for _ in 0..3000 {
        acc = builder.mul(acc, x);
        let h = builder.hash_n_to_hash_no_pad::<plonky2::hash::poseidon::PoseidonHash>(vec![acc, x]);
        acc = h.elements[0];
    }
    builder.register_public_input(acc);
    let data = builder.build::<C>();
    let degree_bits = data.common.degree_bits();
    let t0 = Instant::now();
    for i in 0..proofs {
        let mut pw = PartialWitness::new();
        pw.set_target(x, F::from_canonical_u64(3 + i as u64));
        let proof = data.prove(pw).expect("prove");
        std::hint::black_box(&proof.public_inputs);
    }
    let per_proof = t0.elapsed().as_secs_f64() / proofs as f64;

    println!(
        "{}",
        serde_json::json!({
            "impl": "reference (plonky2-hwa 6a8ca008, rayon)", "cores": cores,
            "metric": "proofs/sec on a 2^12-row standard_recursion_config circuit; LDE+Merkle ms beside it",
            "value": 1.0 / per_proof, "unit": "proofs/s", "ms_per_proof": per_proof * 1e3, "circuit_degree_bits": degree_bits,
            "lde_merkle": {"rows_log2": log_rows, "cols": 135, "ms": best * 1e3,
                            "ms_scaled_to_2p20": best * 1e3 * (1u64 << (20 - log_rows.min(20))) as f64},
        })
    );
}
