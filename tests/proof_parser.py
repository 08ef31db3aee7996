"""bincode parser for plonky2 ProofWithPublicInputs<GoldilocksField, PoseidonGoldilocksConfig, 2>.

Test helper.  Layout per SURVEY.md A.10 (bincode 1.3.3 default: little-endian, u64 lengths); the
blobs come from qbench_data/example.bin, stored by the reference through
city_rollup_common/src/qworker/memory_proof_store/mod.rs:31-46 (bincode::serialize).
"""
import struct

import numpy as np


class _R:
    def __init__(self, b):
        self.b, self.o = b, 0

    def u64(self):
        (v,) = struct.unpack_from("<Q", self.b, self.o)
        self.o += 8
        return v

    def felts(self, n):
        a = np.frombuffer(self.b, dtype="<u8", count=n, offset=self.o).astype(np.uint64)
        self.o += 8 * n
        return a

    def cap(self):
        n = self.u64()
        return self.felts(4 * n).reshape(n, 4)

    def ext_vec(self):
        n = self.u64()
        return self.felts(2 * n).reshape(n, 2)

    def merkle_proof(self):
        n = self.u64()
        return self.felts(4 * n).reshape(n, 4)


def parse_proof(blob):
    r = _R(blob)
    p = {}
    p["wires_cap"] = r.cap()
    p["zs_pp_cap"] = r.cap()
    p["quotient_cap"] = r.cap()
    op = {}
    for k in ("constants", "plonk_sigmas", "wires", "plonk_zs", "plonk_zs_next", "partial_products",
              "quotient_polys", "lookup_zs", "lookup_zs_next"):
        op[k] = r.ext_vec()
    p["openings"] = op
    ncaps = r.u64()
    p["commit_phase_merkle_caps"] = [r.cap() for _ in range(ncaps)]
    nq = r.u64()
    rounds = []
    for _ in range(nq):
        ninit = r.u64()
        init = []
        for _ in range(ninit):
            nl = r.u64()
            leaf = r.felts(nl)
            sib = r.merkle_proof()
            init.append((leaf, sib))
        nsteps = r.u64()
        steps = []
        for _ in range(nsteps):
            evals = r.ext_vec()
            sib = r.merkle_proof()
            steps.append((evals, sib))
        rounds.append({"initial": init, "steps": steps})
    p["query_rounds"] = rounds
    p["final_poly"] = r.ext_vec()
    p["pow_witness"] = r.u64()
    npi = r.u64()
    p["public_inputs"] = r.felts(npi)
    assert r.o == len(blob), (r.o, len(blob))
    return p
