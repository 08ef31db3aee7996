"""city_rollup_b200 — B200-native (sm_100a CUDA) Plonky2 proving hot path behind City Rollup's
worker jobs: PolynomialBatch commitment (iNTT, rate-8 coset LDE, Poseidon Merkle tree), Challenger,
the PLONK stages between the commitments (Z / partial products, quotient polynomials) and the FRI commit phase, behind the C ABI in include/p2b.h.  See DESIGN.md."""
from ._lib import SO_PATH, build, load  # noqa: F401
from .plonky2 import (Challenger, CircuitData, Context, circuit_fingerprint, FriParams, MerkleTree, P2BError,  # noqa: F401
                      PolynomialBatch, all_wires_permutation_partial_products, compute_quotient_polys,
                      fri_committed_trees, fri_proof_of_work, proof_from_bincode, proof_shape, proof_to_bincode,
                      prove, prove_collect, prove_native, prove_native_device, prove_openings, prove_poll,
                      prove_submit, prove_upload_poll)
