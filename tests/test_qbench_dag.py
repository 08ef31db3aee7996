"""CPU: the native replay tool reads the job DAG of the reference's dumped block (bincode BlockProofStoreDump,
city_rollup_core_worker_qbench/src/dump.rs:16-27) — the level counters, goals and next-job lists the job planner wrote
into the store (city_rollup_common/src/qworker/proof_store.rs:60-89) — and, replaying it without a GPU (--plan-only),
finds what SURVEY.md Appendix B decoded by hand: 43 plonky2 jobs / 67 proofs, 3 Groth16 wrappers, and reaches
NotifyOrchestratorComplete."""
import json
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def exe(tmp_path_factory):
    import city_rollup_b200 as m

    if not shutil.which("g++"):
        pytest.skip("no g++")
    m.build()
    out = tmp_path_factory.mktemp("qb") / "qbench_replay"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-Wall", "-I", ROOT, os.path.join(ROOT, "tools", "qbench_replay.cpp"), "-L",
                           os.path.dirname(m.SO_PATH), "-lp2b", "-lpthread", "-o", str(out)])
    env = dict(os.environ, LD_LIBRARY_PATH=os.path.dirname(m.SO_PATH) + ":" + os.environ.get("LD_LIBRARY_PATH", ""))
    return str(out), env


def plan(exe, *args):
    path, env = exe
    res = subprocess.run([path, "--plan-only"] + list(args), env=env, capture_output=True, text=True, timeout=60)
    assert res.returncode == 0, res.stderr + res.stdout
    return json.loads(res.stdout.strip().splitlines()[-1])


def test_dumped_block_dag(exe):
    d = plan(exe, "-d", os.path.join(ROOT, "tests", "golden", "example_dag.bin"))
    assert d["checkpoint_id"] == 4
    assert d["plonky2_jobs"] == 43 and d["plonky2_proofs"] == 67 and d["groth16_jobs"] == 3
    assert d["notify_orchestrator_complete"] == 1 and d["processed"] == d["jobs_in_store"] == 60
    assert d["entry_jobs"] == 23  # 20 op leaves + 3 sighash introspections
    # CityOpJobConfig {register 4, claim 2, transfer 4, add_withdrawal 4, process_withdrawal 4, add_deposit 2} and their
    # binary aggregation trees (n - 1 aggregates), the two block aggregators, state transition, 3 + 3 sighash jobs
    assert d["jobs_per_circuit"] == {"0": 4, "1": 3, "2": 2, "3": 1, "4": 2, "5": 1, "6": 4, "7": 3, "8": 4, "9": 3, "10": 4,
                                     "11": 3, "32": 1, "33": 3, "34": 3, "40": 1, "41": 1}


def test_fixture_equals_the_reference_dump(exe):
    ref = "/root/reference/qbench_data/example.bin"
    if not os.path.exists(ref):
        pytest.skip("reference checkout not present (GPU box)")
    a = plan(exe, "-d", os.path.join(ROOT, "tests", "golden", "example_dag.bin"))
    b = plan(exe, "-d", ref)
    a.pop("source"), b.pop("source")
    assert a == b


def test_built_in_plans_complete(exe):
    d = plan(exe)
    assert d["plonky2_jobs"] == 43 and d["plonky2_proofs"] == 67 and d["notify_orchestrator_complete"] == 1
    t = plan(exe, "--agg-tree", "6")
    assert t["plonky2_jobs"] == 127 and t["notify_orchestrator_complete"] == 1 and t["entry_jobs"] == 64


def test_truncated_dump_is_rejected(exe, tmp_path):
    path, env = exe
    blob = open(os.path.join(ROOT, "tests", "golden", "example_dag.bin"), "rb").read()
    bad = tmp_path / "bad.bin"
    bad.write_bytes(blob[: len(blob) // 2])
    res = subprocess.run([path, "--plan-only", "-d", str(bad)], env=env, capture_output=True, text=True, timeout=60)
    assert res.returncode != 0 and "dump" in res.stderr


def test_worker_pool_keeps_the_workers_busy_on_the_dumped_dag(exe):
    """the store protocol + ready queue + worker pool with a sleeping prover (--fake-ms): every job of 8 blocks in flight is
    processed, every proving job recorded, and 8 workers are busy most of the time (the DAG itself allows 99 %: what is
    measured on the GPU is then the prover, not the scheduler)"""
    path, env = exe
    res = subprocess.run([path, "--fake-ms", "4", "-d", os.path.join(ROOT, "tests", "golden", "example_dag.bin"), "-n", "8",
                          "--contexts", "8"], env=env, capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stderr + res.stdout
    d = json.loads(res.stdout.strip().splitlines()[-1])
    assert d["jobs"] == 8 * 60 and d["proving_jobs"] == d["jobs_recorded"] == d["stored_proofs"] == 8 * 43 and d["proofs"] == 8 * 67
    assert d["worker_busy_fraction"] > 0.8, d
    assert d["wall_s"] <= d["wall_incl_teardown_s"]
