#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200-native Plonky2 hot path, on BASELINE.json's metric:
"proofs/sec on qbench worker jobs at 1/2/4/8 B200; LDE+Merkle ms at 2^20 rows".

M1 (the line's `value`): complete proofs per second through p2b_prove at the shape of City Rollup's worker jobs —
2^12 rows x 135 wires (the op circuits are padded to exactly 2^12 rows), the gate set of the op
circuits (city_common_circuit/src/builder/pad_circuit.rs:31-55 + the in-tree u32 gates), rate 8, cap 4, 16-bit
proof of work, 28 queries, arities [4,4] (zk_signature2/mod.rs:33-57), synthetic witness.  The job loop mirrors the
reference's worker (city_rollup_core_worker/src/actors/simple.rs:32-113: pop a job, prove, store the proof) with
CONTEXTS worker threads per GPU, one p2b context (= one CUDA stream) each — independent jobs, no collective on the
proof path.  A *step* is PROOFS_PER_STEP proofs per GPU (every worker proves
PROOFS_PER_STEP / CONTEXTS jobs).
  value  witness already resident in HBM (p2b_prove_dev), proof read back; device time = CUDA events on the
         contexts' streams, earliest start to latest stop, max over ranks
  e2e    the same jobs through p2b_prove with HOST witness buffers: pinned host memory (`value`) and pageable
         separately allocated columns, plonky2's Vec<PolynomialValues> (`pageable_value`); H2D of the witness and D2H of
         the proof inside the timed region
M2: one PolynomialBatch::from_values commit of 2^20 rows x 135 columns (ms), N=1 only, with its stage split.

`--impl reference` times the CPU prover of the same jobs (oracle/prove.c, C + OpenMP on all host threads — the
reference's own Rust prover lives in an un-vendored dependency and cannot be built here) on the same
config / metric / unit, each step a bounded sample (one proof), plus the M2 commit on a scaled sample.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)

# more hardware work queues than the default 8: 16 worker streams (+ their copy streams) otherwise share queues and
# serialise behind each other (measured: 16 workers 665 -> 724 proofs/s).  Must be set before CUDA initialises.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

import numpy as np  # noqa: E402

DEGREE_BITS, N_WIRES, RATE_BITS, CAP_HEIGHT = 12, 135, 3, 4
FP = dict(rate_bits=3, cap_height=4, proof_of_work_bits=16, num_query_rounds=28, reduction_arity_bits=[4, 4])
# worker threads (p2b contexts / CUDA streams) per GPU; profiles/r02_bench_v12_c*.json: 12 -> 845, 16 -> 870, 20 -> 879,
# 24 -> 890, 32 -> 895 proofs/s (workers sleep in a blocking sync while their proof runs: ~0.3 ms of host CPU per proof)
CONTEXTS = 24
PROOFS_PER_STEP = 96  # 4 jobs per worker and step
METRIC = "proofs/sec on qbench-shaped worker jobs (2^12 rows x 135 wires, City op-circuit gate set, 28 queries); LDE+Merkle ms at 2^20 rows x 135 cols beside it"
UNIT = "proofs/s"
WORKLOAD = ("City Rollup worker proof jobs: CircuitData::prove at 2^12 rows x 135 wires, 21 gate kinds (add_city_common_gates + "
            "u32 gates), rate_bits=3, cap_height=4, pow 16 bits, 28 queries, arities [4,4], synthetic witness (BASELINE.json "
            "metric M1 / configs[3] job shape); M2 = standalone commit 2^20 rows x 135 cols")
M2_LOG_N = 20
FULL_LOG_N = 16  # roofline launch of the leaf-hash kernel with every SM occupied: 2^16 rows x 135 columns

# Algorithmic int32-op model of one Poseidon permutation: a field multiplication = 4 32x32 multiplies
# + 14 32-bit add/carry ops (18), a modular add = 5, the MDS layer on 32-bit halves = 2*144 multiply-adds + 12*6 folds:
#   full round 12 * 4 * 18 + 360 = 1224; partial round 4 * 18 + 360 = 432; first constant layer 60
OPS_PER_PERM = 8 * 1224 + 22 * 432 + 60  # = 19356


def config_dict():
    """the same keys on both arms (the driver compares them)"""
    return {"workload": WORKLOAD, "rows": 1 << DEGREE_BITS, "wires": N_WIRES, "rate_bits": RATE_BITS, "cap_height": CAP_HEIGHT,
            "pow_bits": 16, "queries": 28, "arity_bits": [4, 4], "gate_set": "city (21 kinds, 6 selector groups)",
            "proofs_per_step": PROOFS_PER_STEP, "contexts_per_gpu": CONTEXTS,
            "sharding": "independent proof jobs per GPU, no data-path collective",
            "l2": "24 proofs in flight x ~60 MB of LDE / coefficient / digest working set each > 126 MB L2; nothing is reused across proofs"}


def perms_per_proof(circ_desc):
    """Poseidon permutations of one proof: leaf sponges and tree nodes of the three commits, FRI layers, PoW mean, transcript"""
    n, N = 1 << DEGREE_BITS, (1 << DEGREE_BITS) << RATE_BITS
    nch = circ_desc["num_challenges"]
    widths = [N_WIRES, nch * (1 + circ_desc["num_partial_products"]), nch * circ_desc["quotient_degree_factor"]]
    leaf = sum(-(-w // 8) * N for w in widths)
    nodes = 3 * (N - (1 << CAP_HEIGHT))
    fri, ln = 0, N
    for a in FP["reduction_arity_bits"]:
        ln >>= a
        fri += ln * ((2 << a) // 8) + ln - (1 << CAP_HEIGHT)
    return {"leaf": leaf, "nodes": nodes, "fri": fri, "pow_mean": 1 << FP["proof_of_work_bits"], "transcript": 110,
            "total": leaf + nodes + fri + (1 << FP["proof_of_work_bits"]) + 110}


def algorithmic_bytes(n_cols, log_n, rate_bits=RATE_BITS, cap_height=CAP_HEIGHT):
    """values in + coeffs out + leaf-ordered LDE out + digests out."""
    n = 1 << log_n
    N = n << rate_bits
    return 8 * n_cols * n * 2 + 8 * n_cols * N + 2 * (N - (1 << cap_height)) * 32


def load_json(path, default=None):
    try:
        return json.load(open(path))
    except Exception:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={gpu_index}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            pass

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [r.strip().split(", ") for r in open(self.f.name) if r.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) < 9:
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                power.append(float(r[3]))
            except ValueError:
                continue
            for nme, v in zip(names, r[5:9]):
                if v.strip().lower() == "active":
                    reasons.add(nme)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm), "power_w_max": max(power)}


def build_job(pi_hash_fn):
    """the synthetic City-shaped circuit + witness (tests/plonk_ref.py is the witness generator)"""
    import plonk_ref as R
    from plonk_ref import CITY_GATES, CITY_GROUPS

    pis = [7, 2, 3, 4]
    circ = R.SyntheticCircuit(DEGREE_BITS, CITY_GATES, CITY_GROUPS, 7, pi_hash=[int(x) for x in pi_hash_fn(pis)])
    return circ, [1, 2, 3, 4], pis


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_commit_seconds(O, cols):
    t0 = time.perf_counter()
    O.batch_from_values(cols, RATE_BITS, CAP_HEIGHT, want_leaves=True, want_digests=True)
    return time.perf_counter() - t0


def cpu_m2_sample(O, budget_s=8.0):
    """the M2 commit on the CPU, on a row sample scaled linearly to 2^20 rows"""
    from util import rand_felts

    probe = [rand_felts(0x5EED0001 + c, 1 << 10) for c in range(N_WIRES)]
    t_probe = cpu_commit_seconds(O, probe)
    log_rows = 10
    while log_rows < M2_LOG_N and t_probe * (1 << (log_rows + 1 - 10)) * 1.15 < budget_s:
        log_rows += 1
    cols = [rand_felts(0x5EED0001 + c, 1 << log_rows) for c in range(N_WIRES)] if log_rows > 10 else probe
    dt = cpu_commit_seconds(O, cols)
    return {"lde_merkle_ms": dt * 1e3 * (1 << (M2_LOG_N - log_rows)), "rows": 1 << M2_LOG_N, "cols": N_WIRES,
            "sample": f"2^{log_rows} of 2^{M2_LOG_N} rows x {N_WIRES} cols, scaled linearly in rows", "sample_seconds": dt}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is a CPU run on ALL host cores
    # (libgomp reads the variable when the oracle library is loaded)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import p2oracle as O

    cores = O.num_threads()
    circ, digest, pis = build_job(O.hash_no_pad)
    pd = O.ProverData(circ.desc(), circ.constants_sigmas_values(), FP)
    wv = circ.wire_values()
    for _ in range(max(1, min(args.warmup, 2))):
        pd.prove(digest, wv, pis)
    steps = args.steps
    t0 = time.perf_counter()
    for _ in range(steps):
        pd.prove(digest, wv, pis)
    dt = (time.perf_counter() - t0) / steps  # seconds per proof
    value = 1.0 / dt
    sample = f"1 proof per step (the CUDA arm's step is {PROOFS_PER_STEP} proofs per GPU); proofs/s = 1 / seconds per proof"
    m2 = cpu_m2_sample(O)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 * PROOFS_PER_STEP, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": config_dict(),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "ms_per_proof": dt * 1e3,
                         "note": "CPU restatement of the prover (oracle/prove.c, C + OpenMP); the Rust reference cannot be built here"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "m2_lde_merkle_2p20x135": m2,
        "gpu_launches": 0,
    }
    emit(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ CUDA arm
class ProofFarm:
    """CONTEXTS worker threads of one GPU, each with its own p2b context / stream, proving the same job shape"""

    def __init__(self, m, device, n_ctx, circ, digest, pis, to_device=None):
        """to_device(np.int64 matrix) -> (object kept alive, device pointer); default = a torch CUDA tensor"""
        if to_device is None:
            import torch

            def to_device(a):
                t = torch.from_numpy(a).cuda()
                torch.cuda.synchronize()
                return t, t.data_ptr()

        self.m, self.n_ctx, self.digest, self.pis = m, n_ctx, digest, pis
        self.params = m.FriParams(FP["rate_bits"], FP["cap_height"], FP["proof_of_work_bits"], FP["num_query_rounds"],
                                  FP["reduction_arity_bits"])
        self.ctxs = [m.Context(device) for _ in range(n_ctx)]
        for c in self.ctxs:
            c.set_blocking_sync(n_ctx > 1)  # sleep while waiting: 8 workers per GPU x 8 GPUs share the host cores
        self.state = []
        wv = circ.wire_values()
        stacked = np.stack(wv)
        self.dev, self.pinned, self.pageable = [], [], []
        for c in self.ctxs:
            cd = m.CircuitData(c, circ.desc())
            cs = m.PolynomialBatch.from_values(c, circ.constants_sigmas_values(), RATE_BITS, False, CAP_HEIGHT, keep_values=True)
            self.state.append((cd, cs))
            self.dev.append(to_device(stacked.view(np.int64)))
            w = c.pinned_empty(stacked.shape)
            w[:] = stacked
            self.pinned.append(w)
            self.pageable.append([col.copy() for col in wv])  # one allocation per column, as plonky2's witness
        self.n_words = None

    def prove_one(self, i, mode):
        c = self.ctxs[i]
        cd, cs = self.state[i]
        if mode == "dev":
            return self.m.prove_native_device(c, cd, cs, self.digest, self.dev[i][1], self.pis, self.params)
        src = self.pinned[i] if mode == "pinned" else self.pageable[i]
        return self.m.prove_native(c, cd, cs, self.digest, src, self.pis, self.params, raw=True)

    def run(self, mode, per_ctx, warm):
        """every worker: `warm` untimed proofs, a thread barrier, then `per_ctx` timed proofs between its context's
        CUDA timer events.  -> (device ms from the earliest start to the latest stop, kernels launched, wall s, cpu s)"""
        bar = threading.Barrier(self.n_ctx + 1)
        launches = [0] * self.n_ctx
        err = []

        def worker(i):
            try:
                c = self.ctxs[i]
                for _ in range(warm):
                    self.prove_one(i, mode)
                bar.wait()
                bar.wait()
                l0 = c.launch_count()
                c.timer_start()
                for _ in range(per_ctx):
                    w = self.prove_one(i, mode)
                c.timer_stop_ms()
                launches[i] = c.launch_count() - l0
                self.n_words = w.size
            except Exception as e:  # noqa: BLE001
                err.append(e)
                try:
                    bar.abort()
                except Exception:
                    pass

        th = [threading.Thread(target=worker, args=(i,)) for i in range(self.n_ctx)]
        for t in th:
            t.start()
        bar.wait()  # all warm
        self.before_timed()
        t0, c0 = time.perf_counter(), time.process_time()
        bar.wait()
        for t in th:
            t.join()
        wall, cpu = time.perf_counter() - t0, time.process_time() - c0
        if err:
            raise err[0]
        ms = max(a.timer_span_ms(b) for a in self.ctxs for b in self.ctxs)
        return ms, sum(launches), wall, cpu

    def before_timed(self):
        pass

    def close(self):
        for (cd, cs), c in zip(self.state, self.ctxs):
            cs.free()
            cd.free()
            c.close()


class RankGroup:
    """the only collectives of the benchmark: the barrier before a timed region and the max / sum over ranks of its
    result (one process per GPU; the proof path itself has no collective).  device = "cuda" under NCCL, "cpu" under gloo."""

    def __init__(self, world, device, sync=None):
        self.world, self.device, self.sync = world, device, sync

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        if self.sync:
            self.sync()

    def _reduce(self, x, op):
        import torch
        t = torch.tensor([x], dtype=torch.float64, device=self.device)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=getattr(dist.ReduceOp, op))
        return float(t.item())

    def max(self, x):
        return self._reduce(x, "MAX")

    def sum(self, x):
        return self._reduce(x, "SUM")


def timed_job_run(farm, group, mode, per_ctx, warm):
    """one timed region on every rank -> (whole-job proofs/s over all ranks, max-over-ranks device ms, launches of all
    ranks, this rank's wall s, this rank's host cpu s)"""
    farm.before_timed = group.barrier  # every rank's workers are warm before any rank starts its timed region
    ms, launches, wall, cpu = farm.run(mode, per_ctx, warm)
    group.barrier()
    ms = group.max(ms)
    total = group.world * farm.n_ctx * per_ctx
    return total / (ms * 1e-3), ms, group.sum(launches), wall, cpu


def run_cuda(args):
    import torch
    import torch.distributed as dist

    import city_rollup_b200 as m

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback "
                         "(use --impl reference for the CPU restatement)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=600))

    group = RankGroup(world, "cuda", torch.cuda.synchronize)

    c0 = m.Context(local)
    circ, digest, pis = build_job(c0.hash_no_pad)
    c0.close()
    n_ctx = args.contexts
    per_ctx_step = max(1, PROOFS_PER_STEP // n_ctx)
    proofs_per_step = per_ctx_step * n_ctx
    farm = ProofFarm(m, local, n_ctx, circ, digest, pis)
    warm = args.warmup * per_ctx_step  # W whole warm-up steps
    total_proofs = world * proofs_per_step * args.steps

    # ---- value: witness resident in HBM
    sampler = ClockSampler(local) if rank == 0 else None
    value, ms_dev, launches_all, wall_dev, cpu_dev = timed_job_run(farm, group, "dev", per_ctx_step * args.steps, warm)
    clocks = sampler.stop() if sampler else None

    # ---- e2e: host witness buffers through p2b_prove (pinned, then pageable per-column allocations)
    _, ms_pin, _, wall_pin, cpu_pin = timed_job_run(farm, group, "pinned", per_ctx_step * args.steps, 2)
    _, ms_pag, _, wall_pag, cpu_pag = timed_job_run(farm, group, "pageable", per_ctx_step * args.steps, 2)
    n_words = farm.n_words
    h2d = proofs_per_step * N_WIRES * (1 << DEGREE_BITS) * 8
    d2h = proofs_per_step * n_words * 8

    if rank != 0:
        farm.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- rank 0: one worker alone (the reference's one-job-at-a-time worker) + stage profile for the roofline
    peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), {})
    hbm_peak = peaks.get("hbm_gbs")
    hbm_src = "measured (MEASURED_PEAKS.json)" if hbm_peak else "fallback (B200_PROFILING.md)"
    hbm_peak = hbm_peak or 6650.0
    ipk = load_json(os.path.join(ROOT, "profiles", "int32_peak.json"), {})
    int_peak = ipk.get("int32_peak_gops")  # dual-pipe IMAD+IADD3 thread-instr/s measured on this pool's B200
    int_src = "measured (profiles/int32_peak.json: IMAD+IADD3 dual-issue microbenchmark)" if int_peak else \
        "nominal 148 SMs x 128 lanes x 1.965 GHz"
    int_peak = int_peak or 148 * 128 * 1.965
    traffic = load_json(os.path.join(ROOT, "profiles", "ncu_traffic.json"), {})

    single = None
    roofline = None
    try:
        c = farm.ctxs[0]
        c.set_blocking_sync(False)
        n1 = 24
        for _ in range(3):
            farm.prove_one(0, "dev")
        best = 1e30
        for _ in range(3):
            c.timer_start()
            for _ in range(n1):
                farm.prove_one(0, "dev")
            best = min(best, c.timer_stop_ms() / n1)
        c.profile_enable(True)
        c.profile_read()
        for _ in range(n1):
            farm.prove_one(0, "dev")
        st_ms, st_cnt = c.profile_read()
        c.profile_enable(False)
        # the same worker in latency mode (p2b_set_latency_mode: it has the GPU to itself)
        best_lat = None
        try:
            c.set_latency_mode(True)
            for _ in range(3):
                farm.prove_one(0, "dev")
            best_lat = 1e30
            for _ in range(3):
                c.timer_start()
                for _ in range(n1):
                    farm.prove_one(0, "dev")
                best_lat = min(best_lat, c.timer_stop_ms() / n1)
        finally:
            c.set_latency_mode(False)
        c.set_blocking_sync(n_ctx > 1)
        single = {"proofs_per_s": 1e3 / best, "ms_per_proof": best, "note": "one context, one proof at a time (spin wait), best of 3 x 24",
                  "stage_ms_per_proof": {k: v / n1 for k, v in st_ms.items() if v},
                  "latency_mode": {"proofs_per_s": 1e3 / best_lat if best_lat else None, "ms_per_proof": best_lat,
                                   "note": "p2b_set_latency_mode(ctx, 1): Merkle trees fused from 2^15 digests, proof-of-work "
                                           "search on every SM — for a worker that has the GPU to itself"}}
        pp = perms_per_proof(circ.desc())
        leaf_ms = st_ms["leaf_hash"] / n1
        leaf_launches = st_cnt["leaf_hash"] / n1
        leaf_gops = pp["leaf"] * OPS_PER_PERM / (leaf_ms * 1e-3) / 1e9
        lde_bytes = 8 * (1 << DEGREE_BITS) * (1 << RATE_BITS) * sum(
            [N_WIRES, 20, 16]) + 32 * 3 * ((1 << DEGREE_BITS) << RATE_BITS)
        # the same kernel with every SM occupied, as in the timed region (24 proofs in flight): one launch over the
        # 2^19 leaves x 135 columns of a 2^16-row commit (BASELINE configs[1]; 2048 CTAs; 17 permutations per leaf,
        # exactly the per-leaf work of a proof's wires commit), CUDA events around the launch on the context's stream
        full = torch.empty((N_WIRES, 1 << FULL_LOG_N), dtype=torch.int64, device="cuda")
        gfull = torch.Generator(device="cuda")
        gfull.manual_seed(11)
        full.random_(0, 2**62, generator=gfull)
        torch.cuda.synchronize()
        c.profile_enable(True)
        for i in range(4):
            b = m.PolynomialBatch.from_values_device(c, full.data_ptr(), N_WIRES, FULL_LOG_N, RATE_BITS, CAP_HEIGHT)
            b.free()
            if i == 0:
                c.profile_read()  # drop the warm-up launch
        st_full, cnt_full = c.profile_read()
        c.profile_enable(False)
        del full
        full_ms = st_full["leaf_hash"] / max(1, cnt_full["leaf_hash"])
        full_perms = -(-N_WIRES // 8) * ((1 << FULL_LOG_N) << RATE_BITS)
        full_gops = full_perms * OPS_PER_PERM / (full_ms * 1e-3) / 1e9
        full_bytes = 8 * N_WIRES * ((1 << FULL_LOG_N) << RATE_BITS) + 32 * ((1 << FULL_LOG_N) << RATE_BITS)
        exec_instr = traffic.get("k_leaf_hash_colmajor_executed_instr_per_perm")
        roofline = {
            "kernel": "k_leaf_hash_colmajor (Poseidon sponge over the LDE rows of a commit: ceil(C / 8) permutations per leaf), "
                      "the largest kernel of the step (55 % of a proof's executed instructions)",
            "bound": "int32", "achieved": full_gops, "peak": int_peak, "unit": "Gop/s (int32)", "frac": full_gops / int_peak,
            "traffic": traffic.get("k_leaf_hash_colmajor_2p16x135"),
            "peak_source": int_src, "ops_per_permutation": OPS_PER_PERM,
            "permutations_per_launch": full_perms, "ms_per_launch": full_ms, "gperm_s": full_perms / (full_ms * 1e-3) / 1e9,
            "algorithmic_bytes_per_launch": full_bytes,
            "timed": "CUDA events around the launch on the context's stream; one launch over 2^19 leaves x 135 columns "
                     "(2048 CTAs: every SM occupied, as in the timed region where the launches of 24 proofs overlap)",
            "frac_executed_instructions": (full_perms * exec_instr / (full_ms * 1e-3) / 1e9 / int_peak) if exec_instr else None,
            "executed_instructions_per_permutation": exec_instr,
            "single_proof_launches": {
                "achieved": leaf_gops, "frac": leaf_gops / int_peak, "permutations_per_proof_in_kernel": pp["leaf"],
                "ms_per_proof_in_kernel": leaf_ms, "launches_per_proof": leaf_launches,
                "share_of_single_worker_proof": leaf_ms / best,
                "traffic": traffic.get("k_leaf_hash_colmajor_m1"),
                "note": "the three launches of ONE proof running alone (17 + 3 + 2 permutations per leaf over 2^15 leaves = "
                        "128 CTAs on 148 SMs, two warps per scheduler): latency, not throughput"},
            "whole_proof": {"permutations_per_proof": pp, "achieved_gperm_s": pp["total"] * value / world / 1e9,
                            "int32_frac_all_permutations": pp["total"] * value / world * OPS_PER_PERM / 1e9 / int_peak,
                            "note": "all Poseidon work of a proof at the measured proofs/s of one GPU against the INT32 issue peak"},
            "hbm_view": {"bound": "hbm", "algorithmic_bytes_per_launch": full_bytes,
                         "achieved": full_bytes / (full_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                         "frac": full_bytes / (full_ms * 1e-3) / 1e9 / hbm_peak, "peak_source": hbm_src},
            "note": "integer-issue bound (SURVEY.md §0.7); `frac` uses the fixed 19356-op scalar model of a permutation, "
                    "`frac_executed_instructions` the instructions the kernel executes per permutation (ncu, profiles/)",
        }
    except Exception as e:  # noqa: BLE001
        single = {"error": str(e)[:200]}

    # ---- CPU baseline beside it (bounded sample, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import p2oracle as O
        cores = O.num_threads()
        pd = O.ProverData(circ.desc(), circ.constants_sigmas_values(), FP)
        wv = circ.wire_values()
        ref_words = pd.prove(digest, wv, pis)  # warm-up + the checker: the GPU proof of the same job must equal it
        same = bool((farm.prove_one(0, "pinned") == ref_words).all())
        t0, k = time.perf_counter(), 0
        while k < 3 or (time.perf_counter() - t0 < 12.0 and k < 64):
            pd.prove(digest, wv, pis)
            k += 1
        dt = (time.perf_counter() - t0) / k
        cpu = {"value": 1.0 / dt, "unit": UNIT, "cores": cores, "kind": "port", "ms_per_proof": dt * 1e3,
               "sample": f"{k} proofs of the same job, one at a time on all host threads",
               "gpu_proof_equals_cpu_proof": same,
               "note": "CPU restatement of the prover (oracle/prove.c, C + OpenMP) — the Rust reference cannot be built here"}
        if not args.no_m2:
            cpu["m2_lde_merkle_2p20x135"] = cpu_m2_sample(O, 6.0)
        pd.free()
    farm.close()

    # ---- M2: LDE+Merkle ms at 2^20 rows x 135 cols, N=1 only
    m2 = None
    if world == 1 and not args.no_m2:
        try:
            ctx = m.Context(local)
            big = torch.empty((N_WIRES, 1 << M2_LOG_N), dtype=torch.int64, device="cuda")
            g = torch.Generator(device="cuda")
            g.manual_seed(7)
            big.random_(0, 2**62, generator=g)
            torch.cuda.synchronize()
            times = []
            ctx.profile_enable(True)
            ctx.profile_read()
            for i in range(4):
                ctx.timer_start()
                b = m.PolynomialBatch.from_values_device(ctx, big.data_ptr(), N_WIRES, M2_LOG_N, RATE_BITS, CAP_HEIGHT)
                times.append(ctx.timer_stop_ms())
                b.free()
                if i == 0:
                    ctx.profile_read()  # drop the warm-up run
            st, _ = ctx.profile_read()
            ctx.profile_enable(False)
            ab = algorithmic_bytes(N_WIRES, M2_LOG_N)
            lde_b = 8 * N_WIRES * (1 << M2_LOG_N) * 10
            ntt2 = (st["intt"] + st["lde"]) / 3
            leaf2 = st["leaf_hash"] / 3
            leaf_perms = -(-N_WIRES // 8) * ((1 << M2_LOG_N) << RATE_BITS)
            m2 = {"rows": 1 << M2_LOG_N, "cols": N_WIRES, "lde_merkle_ms": min(times[1:]),
                  "ntt_lde_ms": ntt2, "leaf_hash_ms": leaf2, "tree_levels_ms": st["tree_levels"] / 3,
                  "ntt_lde_hbm_frac": lde_b / (ntt2 * 1e-3) / 1e9 / hbm_peak,
                  "ntt_lde_dram_traffic": traffic.get("ntt_2p20"),
                  # the transforms are bound by integer issue, not by HBM: executed instructions (ncu) / measured time / peak
                  "ntt_lde_int32_frac_executed": ((traffic.get("ntt_2p20") or {}).get("executed_thread_instructions", 0.0)
                                                  / (ntt2 * 1e-3) / 1e9 / int_peak) or None,
                  "whole_commit_hbm_frac": ab / (min(times[1:]) * 1e-3) / 1e9 / hbm_peak,
                  "poseidon_int32_frac": leaf_perms * OPS_PER_PERM / (leaf2 * 1e-3) / 1e9 / int_peak,
                  "leaf_hash_gperm_s": leaf_perms / (leaf2 * 1e-3) / 1e9,
                  "working_set": "11.9 GB per commit >> L2"}
            del big
            ctx.close()
        except Exception as e:  # noqa: BLE001
            m2 = {"error": str(e)[:200]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": config_dict() if n_ctx == CONTEXTS else dict(config_dict(), contexts_per_gpu=n_ctx, proofs_per_step=proofs_per_step),
        "roofline": roofline,
        "cpu_baseline": cpu,
        "e2e": {"value": total_proofs / (ms_pin * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": ms_pin / args.steps, "host_buffers": "pinned (one matrix per worker)",
                "pageable_value": total_proofs / (ms_pag * 1e-3), "pageable_ms_per_step": ms_pag / args.steps,
                "pageable_buffers": "135 separately allocated pageable columns per job (plonky2's Vec<PolynomialValues>)",
                "host_cpu_ms_per_proof": {"dev": cpu_dev / (proofs_per_step * args.steps) * 1e3,
                                          "pinned": cpu_pin / (proofs_per_step * args.steps) * 1e3,
                                          "pageable": cpu_pag / (proofs_per_step * args.steps) * 1e3}},
        "gpu_launches": int(launches_all), "launches_per_proof": launches_all / total_proofs,
        "clocks": clocks, "single_worker": single, "m2_lde_merkle_2p20x135": m2,
        "proof_words": n_words,
    }
    emit(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


class StdoutGuard:
    """Everything a library prints to fd 1 while the benchmark runs (NCCL's version banner, ...) goes to stderr;
    the one JSON line is written to the real stdout at the end."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text):
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


GUARD = None


def emit(line):
    if GUARD is not None:
        GUARD.emit(line)
    else:
        print(line)


def main():
    global GUARD
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--contexts", type=int, default=CONTEXTS, help="worker threads (p2b contexts) per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-m2", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    with StdoutGuard() as g:
        GUARD = g
        if args.impl == "reference":
            return run_reference(args)
        return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
