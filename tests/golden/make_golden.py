#!/usr/bin/env python3
"""Regenerate tests/golden/* from the (read-only) reference checkout.

This script is the ONLY thing in the repo that reads /root/reference, and it is
never run by tests, smoke() or bench.py (the GPU box has no /root/reference).
It extracts, verbatim, the known-answer data the reference holds for the
Plonky2 hot path:

  K1/K2  city_crypto/src/hash/cached_zero_hashes.rs:10-1036, :1039-2066
         -> zero_hashes.json  {"zero": [[4 u64] x128], "marked": [[4 u64] x128]}
  K3     qbench_data/example.bin (10 stored plonky2 proofs, bincode)
         -> example_proofs.bin (concatenated proof blobs) + example_proofs.json (index)
  K4/K5  city_common_circuit/src/circuits/zk_signature2/mod.rs:31-145
         -> circuit_params.json (FRI/Plonk parameters + the 80 k_is)
  K6     city_rollup_common/src/config/sighash_wrapper_config.rs:16-23 (whitelist root), :24-1900 (1 875 circuit fingerprints)
         -> sighash_whitelist.json {"root": [4 u64], "fingerprints": [[4 u64] x1875]}
  DAG    qbench_data/example.bin's job ids, level counters, goals and next-job lists
         -> example_dag.bin (the dump with witness / proof payloads stripped; read by tools/qbench_replay.cpp)

Usage:  python tests/golden/make_golden.py [/root/reference]
"""
import json
import os
import re
import struct
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
OUT = os.path.dirname(os.path.abspath(__file__))


def zero_hash_tables():
    src = open(os.path.join(REF, "city_crypto/src/hash/cached_zero_hashes.rs")).read()
    # four consts in the file; the first two are the HashOut tables, the last two are
    # QHashOut copies.  Split on the const declarations.
    parts = re.split(r"const CACHED_(?:MARKED_LEAF_)?ZERO_HASHES", src)[1:]
    tables = []
    for p in parts:
        nums = [int(x) for x in re.findall(r"GoldilocksField\((\d+)\)", p)]
        assert len(nums) == 128 * 4, len(nums)
        tables.append([nums[i : i + 4] for i in range(0, 512, 4)])
    assert tables[0] == tables[2] and tables[1] == tables[3]
    return {"zero": tables[0], "marked": tables[1],
            "source": "city_crypto/src/hash/cached_zero_hashes.rs:10-1036,1039-2066"}


def example_proofs():
    d = open(os.path.join(REF, "qbench_data/example.bin"), "rb").read()
    off = 8 + 4 + 48  # checkpoint_id u64, rpc_node_id u32, CityOpJobConfig 6xu64
    (n,) = struct.unpack_from("<Q", d, off)
    off += 8
    blobs, index = [], []
    pos = 0
    for _ in range(n):
        kid = d[off : off + 24]
        off += 24
        (l,) = struct.unpack_from("<Q", d, off)
        off += 8
        if l > 100000:  # the ten stored ProofWithPublicInputs blobs
            blobs.append(d[off : off + l])
            index.append({"job_id": kid.hex(), "circuit_type": kid[9], "offset": pos, "len": l})
            pos += l
        off += l
    assert len(blobs) == 10
    return b"".join(blobs), {"proofs": index, "source": "qbench_data/example.bin"}


def example_dag():
    """The job DAG of the dumped block, in the dump's own container format (bincode BlockProofStoreDump,
    city_rollup_core_worker_qbench/src/dump.rs:16-27: DumpProofStoreConfig, then SimpleProofStoreMemory {proofs, counters}):
    every key of the store is kept; the payload is kept only for the Counter entries (value / goal / next-jobs list,
    city_rollup_common/src/qworker/proof_store.rs:60-89) and stripped (length 0) for witnesses and proofs, which the
    native replay does not read.  1.4 MB -> a few KB."""
    d = open(os.path.join(REF, "qbench_data/example.bin"), "rb").read()
    off = 8 + 4 + 48
    out = bytearray(d[:off])
    (n,) = struct.unpack_from("<Q", d, off)
    off += 8
    out += struct.pack("<Q", n)
    kept = 0
    for _ in range(n):
        kid = d[off : off + 24]
        off += 24
        (l,) = struct.unpack_from("<Q", d, off)
        off += 8
        payload = d[off : off + l] if kid[22] == 16 else b""  # ProvingJobDataType::Counter = 16
        kept += 1 if payload else 0
        out += kid + struct.pack("<Q", len(payload)) + payload
        off += l
    (nc,) = struct.unpack_from("<Q", d, off)
    assert nc == 0 and off + 8 == len(d)
    out += struct.pack("<Q", 0)
    return bytes(out), kept


def circuit_params():
    src = open(os.path.join(REF, "city_common_circuit/src/circuits/zk_signature2/mod.rs")).read()
    body = src[src.index("pub fn get_verifier_template_zk_signature") :]
    body = body[: body.index("\n}\n")]
    kis = [int(x) for x in re.findall(r"^\s+(\d+),\s*$", body[body.index("k_is") :], re.M)]
    assert len(kis) == 80, len(kis)

    def field(name):
        return int(re.search(name + r":\s*(\d+)", body).group(1))

    arity = re.search(r"ConstantArityBits\((\d+),\s*(\d+)\)", body)
    return {
        "source": "city_common_circuit/src/circuits/zk_signature2/mod.rs:31-145",
        "rate_bits": field("rate_bits"),
        "cap_height": field("cap_height"),
        "proof_of_work_bits": field("proof_of_work_bits"),
        "constant_arity_bits": [int(arity.group(1)), int(arity.group(2))],
        "num_query_rounds": field("num_query_rounds"),
        "degree_bits": field("degree_bits"),
        "reduction_arity_bits": [int(x) for x in re.search(r"reduction_arity_bits:\s*vec!\[([^\]]*)\]", body).group(1).split(",")],
        "num_leaves_per_oracle": [int(x) for x in re.search(r"num_leaves_per_oracle:\s*vec!\[([^\]]*)\]", body).group(1).split(",")],
        "num_challenges": field("num_challenges"),
        "num_constants": field("num_constants"),
        "num_routed_wires": field("num_routed_wires"),
        "num_wires": field("num_wires"),
        "num_quotient_polys": field("num_quotient_polys"),
        "quotient_degree_factor": field("quotient_degree_factor"),
        "num_gate_constraints": field("num_gate_constraints"),
        "num_partial_products": field("num_partial_products"),
        "total_partial_products": field("total_partial_products"),
        "k_is": kis,
    }


def sighash_whitelist():
    """The root of the sighash-circuit whitelist tree and the 1 875 circuit fingerprints it is built from
    (city_store/src/store/sighash/mod.rs:44-75: leaf i of a height-16 zero-hash Merkle tree = the fingerprint of the i-th
    gadget id in SORTED order; the tests rebuild the order from the id generator,
    city_rollup_common/src/introspection/rollup/introspection.rs:402-431, and derive(Ord) on the id's fields, :156-163)."""
    src = open(os.path.join(REF, "city_rollup_common/src/config/sighash_wrapper_config.rs")).read()
    a = src.index("SIGHASH_WHITELIST_TREE_ROOT")
    b = src.index("SIGHASH_CIRCUIT_FINGERPRINTS")
    e = src.index("];", b)
    root = [int(x) for x in re.findall(r"GoldilocksField\((\d+)\)", src[a:b])]
    nums = [int(x) for x in re.findall(r"GoldilocksField\((\d+)\)", src[b:e])]
    assert len(root) == 4 and len(nums) == 4 * 1875, (len(root), len(nums))
    height = int(re.search(r"SIGHASH_CIRCUIT_WHITELIST_TREE_HEIGHT: u8 = (\d+)", src).group(1))
    md = int(re.search(r"^pub const SIGHASH_CIRCUIT_MAX_DEPOSITS: usize = (\d+)", src, re.M).group(1))
    mw = int(re.search(r"^pub const SIGHASH_CIRCUIT_MAX_WITHDRAWALS: usize = (\d+)", src, re.M).group(1))
    return {"source": "city_rollup_common/src/config/sighash_wrapper_config.rs:7,14-23,24-1900", "tree_height": height,
            "max_deposits": md, "max_withdrawals": mw, "root": root,
            "fingerprints": [nums[i : i + 4] for i in range(0, len(nums), 4)]}


if __name__ == "__main__":
    json.dump(sighash_whitelist(), open(os.path.join(OUT, "sighash_whitelist.json"), "w"))
    json.dump(zero_hash_tables(), open(os.path.join(OUT, "zero_hashes.json"), "w"))
    blob, idx = example_proofs()
    open(os.path.join(OUT, "example_proofs.bin"), "wb").write(blob)
    json.dump(idx, open(os.path.join(OUT, "example_proofs.json"), "w"), indent=1)
    json.dump(circuit_params(), open(os.path.join(OUT, "circuit_params.json"), "w"), indent=1)
    dag, kept = example_dag()
    open(os.path.join(OUT, "example_dag.bin"), "wb").write(dag)
    print("example_dag.bin:", len(dag), "bytes,", kept, "counter entries kept")
    print("golden fixtures written to", OUT)
