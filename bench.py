#!/usr/bin/env python3
"""bench.py — headline benchmark of the B200-native Plonky2 hot path (BASELINE.json).

A *step* is one PolynomialBatch::from_values commit (iNTT -> rate-8 coset LDE -> Poseidon Merkle tree,
cap_height 4) of BASELINE.json configs[1]: 2^16 rows x 135 wire columns — one GPU, synthetic witnesses
(SplitMix64).  `value` is whole-job commits/s with the inputs already resident in
HBM; `e2e` is the same commit through the C-ABI host entry point (p2b_batch_from_values with pinned host
columns: H2D inside the timed region, cap read back).  With --gpus N (torchrun, one rank per GPU) every
rank commits its own independent batch — the path shards by independent proof jobs, no collective on the
data path — and value = N * steps / max-over-ranks time.

`--impl reference` times the CPU restatement of the same commit (oracle/, OpenMP, all host threads): the
reference's own prover is Rust in an un-vendored dependency and cannot be built here.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

LOG_N, N_COLS, RATE_BITS, CAP_HEIGHT = 16, 135, 3, 4
METRIC = "PolynomialBatch commits/sec (LDE+Merkle): 2^16 rows x 135 cols, rate_bits=3, cap_height=4"
UNIT = "commits/s"
WORKLOAD = "standalone PolynomialBatch::from_values commit: 2^16 rows x 135 wire columns, rate_bits=3, Poseidon Merkle cap_height=4 (BASELINE.json configs[1])"

# Algorithmic int32-op model of one Poseidon permutation (DESIGN.md "Rooflines"): the oracle's scalar
# schedule with a field multiplication = 4 32x32 multiplies + 14 32-bit add/carry ops (18), a modular
# add = 5, and the MDS layer on 32-bit halves = 2*144 multiply-adds + 12 * 6 fold ops:
#   full round  : 12 lanes * 4 mul * 18 + (288 + 72)           = 1224
#   partial     : 1 lane  * 4 mul * 18 + (288 + 72)            =  432
#   first layer : 12 adds * 5                                   =   60
OPS_PER_PERM = 8 * 1224 + 22 * 432 + 60  # = 19356


def algorithmic_bytes(n_cols, log_n, rate_bits=RATE_BITS, cap_height=CAP_HEIGHT):
    """values in + coeffs out + leaf-ordered LDE out + digests out."""
    n = 1 << log_n
    N = n << rate_bits
    return 8 * n_cols * n * 2 + 8 * n_cols * N + 2 * (N - (1 << cap_height)) * 32


def n_perms(n_cols, log_n, rate_bits=RATE_BITS, cap_height=CAP_HEIGHT):
    N = (1 << log_n) << rate_bits
    return -(-n_cols // 8) * N, N - (1 << cap_height)


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


def synth_columns(seed_base, n_cols, log_n, out):
    from util import rand_felts
    for c in range(n_cols):
        out[c] = rand_felts(seed_base + c, 1 << log_n)


# ------------------------------------------------------------------------------------------------ reference arm
def cpu_commit_seconds(O, cols, reps=1):
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        O.batch_from_values(cols, RATE_BITS, CAP_HEIGHT, want_leaves=True, want_digests=True)
        best = min(best, time.perf_counter() - t0)
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is a CPU run on ALL host cores
    # (libgomp reads the variable when the oracle library is loaded)
    os.environ["OMP_NUM_THREADS"] = str(os.cpu_count() or 1)
    import p2oracle as O
    from util import rand_felts

    cores = O.num_threads()
    # bounded sample: probe a 2^12-row commit, then pick the largest row count whose K+W steps fit ~150 s
    probe = [rand_felts(0x5EED0001 + c, 1 << 12) for c in range(N_COLS)]
    t_probe = cpu_commit_seconds(O, probe)
    budget = 150.0 / max(1, args.steps + args.warmup)
    log_rows = LOG_N
    while log_rows > 12 and t_probe * (1 << (log_rows - 12)) * 1.15 > budget:
        log_rows -= 1
    cols = [rand_felts(0x5EED0001 + c, 1 << log_rows) for c in range(N_COLS)]
    for _ in range(args.warmup):
        cpu_commit_seconds(O, cols)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_commit_seconds(O, cols)
    dt = (time.perf_counter() - t0) / args.steps
    frac = (1 << log_rows) / float(1 << LOG_N)
    value = frac / dt  # commits of the full workload per second (work is linear in rows up to log factors)
    sample = (f"one full commit per step" if log_rows == LOG_N else
              f"2^{log_rows} of 2^{LOG_N} rows x {N_COLS} cols per step, scaled linearly in rows")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3 / frac, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "note": "CPU restatement (oracle/, C + OpenMP); the Rust reference cannot be built here"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))
    return 0


# ------------------------------------------------------------------------------------------------ CUDA arm
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
        dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=240))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    ctx = m.Context(local)
    n = 1 << LOG_N
    # two independent pinned input batches, alternated between steps (each step's working set, 0.74 GB of
    # LDE + digests, is itself ~6x the 126 MB L2, so nothing survives in L2 from one step to the next)
    host = [ctx.pinned_empty((N_COLS, n)) for _ in range(2)]
    for b, h in enumerate(host):
        synth_columns(0x5EED0001 + 1000 * b + 100000 * rank, N_COLS, LOG_N, h)
    dev = [torch.from_numpy(h.view(np.int64).copy()).cuda() for h in host]  # resident copies for the device-timed arm
    torch.cuda.synchronize()

    def step_dev(i):
        b = m.PolynomialBatch.from_values_device(ctx, dev[i & 1].data_ptr(), N_COLS, LOG_N, RATE_BITS, CAP_HEIGHT)
        b.free()

    def step_e2e(i):
        h = host[i & 1]
        b = m.PolynomialBatch.from_values(ctx, h, RATE_BITS, False, CAP_HEIGHT)  # rows of the pinned matrix = columns
        cap = b.cap  # D2H of the result (16 digests), synchronises
        b.free()
        return cap

    # ---- device-resident throughput (value) + per-stage roofline timing
    for i in range(args.warmup):
        step_dev(i)
    ctx.synchronize()
    ctx.profile_enable(True)
    ctx.profile_read()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ctx.launch_count()
    ctx.timer_start()
    for i in range(args.steps):
        step_dev(i)
    ms = ctx.timer_stop_ms()
    launches = ctx.launch_count() - l0
    barrier()
    clocks = sampler.stop() if sampler else None
    stage_ms, stage_launches = ctx.profile_read()
    ctx.profile_enable(False)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * args.steps / (ms_max * 1e-3)

    # ---- end-to-end through the host C ABI (pinned host columns -> cap on the host)
    for i in range(max(1, args.warmup // 2)):
        step_e2e(i)
    barrier()
    ctx.timer_start()
    for i in range(args.steps):
        step_e2e(i)
    ms_e2e = ctx.timer_stop_ms()
    barrier()
    t = torch.tensor([ms_e2e], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * args.steps / (float(t.item()) * 1e-3)

    # ---- M1: complete synthetic proofs per second at the City Rollup shape (BASELINE.json metric, first half).
    # Independent proof jobs per GPU: every rank proves its own jobs, no collective on the proof
    # path; the multi-GPU aggregate = all proofs / the slowest rank's wall time.
    def measure_m1():
        if args.no_m1:
            return None
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import prove_bench as PB
            circ, digest, pis = PB.build_case(device=local)
            out = {"shape": "2^12 rows x 135 wires, the 13 gate types of the recursion circuits (123 gate constraints), rate 8, cap 4, 16-bit PoW, 28 queries, arities [4,4] "
                            "(city_common_circuit/src/circuits/zk_signature2/mod.rs:33-57); synthetic witness",
                   "call": "p2b_prove (witness columns in pinned host memory -> proof words on the host)"}
            for n_ctx in ((1, 8) if world == 1 else (8,)):
                n_proofs = 40 * n_ctx
                barrier()
                # one context = the reference's one-job-at-a-time worker: spin-wait (lowest latency); several contexts
                # per GPU: sleep on a blocking-sync event (about 1.2 ms of host CPU per proof instead of a busy core
                # per waiting thread, which is what lets 8 GPUs x 8 contexts share the box's host cores)
                # a 40-proof sample lasts 0.17 s and a single host hiccup (the nvidia-smi clock sampler, a page-in)
                # can add half of that: the one-context figure is the best of three samples, and says so
                reps = 3 if n_ctx == 1 else 1
                st, pps, ms_pp = {}, 0.0, 0.0
                for _ in range(reps):
                    st_i = {}
                    pps_i, ms_i = PB.run(n_ctx, n_proofs, circ, digest, pis, device=local, blocking=n_ctx > 1, stats=st_i)
                    if pps_i > pps:
                        st, pps, ms_pp = st_i, pps_i, ms_i
                if world > 1:
                    tt = torch.tensor([n_proofs / pps], dtype=torch.float64, device="cuda")
                    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                    pps = world * n_proofs / float(tt.item())
                out[f"contexts_{n_ctx}"] = {"proofs_per_s": pps, "ms_per_proof_per_context": ms_pp,
                                            "host_cpu_ms_per_proof": st["cpu_s_per_proof"] * 1e3,
                                            "host_wait": "block" if n_ctx > 1 else "spin",
                                            "proofs_per_sample": n_proofs, "samples": reps}
            if world > 1:
                out["aggregate"] = "sum over %d GPUs, 8 contexts each; wall time = slowest rank" % world
            return out
        except Exception as e:  # noqa: BLE001
            if world > 1:
                raise  # a rank that dropped out would leave the others waiting in the barrier
            return {"error": str(e)[:200]}

    m1 = measure_m1() if world > 1 else None

    if rank != 0:
        ctx.close()
        if world > 1:
            dist.destroy_process_group()
        return 0

    # ---- rooflines (rank 0)
    peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), {})
    hbm_peak = peaks.get("hbm_gbs")
    hbm_src = "measured (MEASURED_PEAKS.json)" if hbm_peak else "fallback (B200_PROFILING.md)"
    hbm_peak = hbm_peak or 6650.0
    ipk = load_json(os.path.join(ROOT, "profiles", "int32_peak.json"), {})
    int_peak = ipk.get("int32_peak_gops")  # dual-pipe IMAD+IADD3 thread-instr/s measured on this pool's B200
    int_src = "measured (profiles/int32_peak.json: IMAD+IADD3 dual-issue microbenchmark)" if int_peak else \
        "nominal 148 SMs x 128 lanes x 1.965 GHz"
    int_peak = int_peak or 148 * 128 * 1.965
    leaf_perms, node_perms = n_perms(N_COLS, LOG_N)
    leaf_ms = stage_ms["leaf_hash"] / args.steps
    tree_ms = stage_ms["tree_levels"] / args.steps
    ntt_ms = (stage_ms["intt"] + stage_ms["lde"]) / args.steps
    traffic = load_json(os.path.join(ROOT, "profiles", "ncu_traffic.json"), {})
    leaf_gops = leaf_perms * OPS_PER_PERM / (leaf_ms * 1e-3) / 1e9
    roofline = {
        "kernel": "k_leaf_hash_colmajor (Poseidon sponge over the 135-wide LDE rows; %.0f%% of the step)"
                  % (100.0 * leaf_ms / (ms / args.steps)),
        "bound": "int32", "achieved": leaf_gops, "peak": int_peak, "unit": "Gop/s (int32)",
        "frac": leaf_gops / int_peak, "traffic": traffic.get("k_leaf_hash_colmajor"),
        "peak_source": int_src, "ops_per_permutation": OPS_PER_PERM, "permutations_per_launch": leaf_perms,
        "ms_per_launch": leaf_ms,
        "hbm_view": {"bound": "hbm", "algorithmic_bytes_per_launch": 8 * N_COLS * (n << RATE_BITS) + 32 * (n << RATE_BITS),
                     "achieved": (8 * N_COLS * (n << RATE_BITS) + 32 * (n << RATE_BITS)) / (leaf_ms * 1e-3) / 1e9,
                     "peak": hbm_peak, "unit": "GB/s",
                     "frac": (8 * N_COLS * (n << RATE_BITS) + 32 * (n << RATE_BITS)) / (leaf_ms * 1e-3) / 1e9 / hbm_peak},
        "note": "the dominant kernel is integer-pipe bound (SURVEY.md §0.7: ~380 int32 ops per byte against a machine "
                "balance of ~5), so its roofline is the measured INT32 issue rate; hbm_view gives the same launch against "
                "the HBM peak (its DRAM traffic equals the algorithmic bytes); the HBM-side kernels are in roofline_hbm",
    }
    lde_bytes = 8 * N_COLS * n * 2 + 8 * N_COLS * (n << RATE_BITS)
    roofline_hbm = {
        "kernels": "ntt2::k_strided + ntt2::k_row4096 (iNTT + 8 coset NTTs writing the leaf-ordered LDE)",
        "bound": "hbm", "achieved": lde_bytes / (ntt_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
        "frac": lde_bytes / (ntt_ms * 1e-3) / 1e9 / hbm_peak, "traffic": traffic.get("ntt"),
        "peak_source": hbm_src, "algorithmic_bytes_per_step": lde_bytes, "ms_per_step": ntt_ms,
        "whole_commit": {"algorithmic_bytes": algorithmic_bytes(N_COLS, LOG_N),
                         "achieved": algorithmic_bytes(N_COLS, LOG_N) / (ms / args.steps * 1e-3) / 1e9,
                         "frac": algorithmic_bytes(N_COLS, LOG_N) / (ms / args.steps * 1e-3) / 1e9 / hbm_peak},
    }

    # ---- CPU baseline beside it (bounded sample, rank 0 only, N=1 only)
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        import p2oracle as O
        from util import rand_felts
        cores = O.num_threads()
        probe = [rand_felts(0x5EED0001 + c, 1 << 12) for c in range(N_COLS)]
        t_probe = cpu_commit_seconds(O, probe)
        log_rows = LOG_N
        while log_rows > 12 and t_probe * (1 << (log_rows - 12)) * 1.15 > 25.0:
            log_rows -= 1
        cols = [host[0][c][: 1 << log_rows].copy() for c in range(N_COLS)]
        dt = cpu_commit_seconds(O, cols)
        frac = (1 << log_rows) / float(n)
        cpu = {"value": frac / dt, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": ("one full commit" if log_rows == LOG_N else
                          f"2^{log_rows} of 2^{LOG_N} rows x {N_COLS} cols, scaled linearly in rows"),
               "seconds": dt,
               "note": "CPU restatement (oracle/, C + OpenMP) — the Rust reference cannot be built here"}

    # ---- M2: LDE+Merkle ms at 2^20 rows x 135 cols (BASELINE.json metric, second half), N=1 only
    m2 = None
    if world == 1 and not args.no_m2:
        try:
            big = torch.empty((N_COLS, 1 << 20), dtype=torch.int64, device="cuda")
            g = torch.Generator(device="cuda")
            g.manual_seed(7)
            big.random_(0, 2**62, generator=g)
            times = []
            ctx.profile_enable(True)
            ctx.profile_read()
            for i in range(3):
                ctx.timer_start()
                b = m.PolynomialBatch.from_values_device(ctx, big.data_ptr(), N_COLS, 20, RATE_BITS, CAP_HEIGHT)
                times.append(ctx.timer_stop_ms())
                b.free()
                if i == 0:
                    ctx.profile_read()  # drop the warm-up run
            st, _ = ctx.profile_read()
            ctx.profile_enable(False)
            ab = algorithmic_bytes(N_COLS, 20)
            lde_b = 8 * N_COLS * (1 << 20) * 10
            ntt2 = (st["intt"] + st["lde"]) / 2
            m2 = {"rows": 1 << 20, "cols": N_COLS, "lde_merkle_ms": min(times[1:]),
                  "ntt_lde_ms": ntt2, "leaf_hash_ms": st["leaf_hash"] / 2, "tree_levels_ms": st["tree_levels"] / 2,
                  "ntt_lde_hbm_frac": lde_b / (ntt2 * 1e-3) / 1e9 / hbm_peak,
                  "whole_commit_hbm_frac": ab / (min(times[1:]) * 1e-3) / 1e9 / hbm_peak,
                  "poseidon_int32_frac": n_perms(N_COLS, 20)[0] * OPS_PER_PERM / (st["leaf_hash"] / 2 * 1e-3) / 1e9 / int_peak}
            del big
        except Exception as e:  # noqa: BLE001
            m2 = {"error": str(e)[:200]}

    # ---- M1 (single GPU; the multi-GPU measurement ran above, before the other ranks left)
    if world == 1:
        m1 = measure_m1()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "rows": n, "cols": N_COLS, "rate_bits": RATE_BITS, "cap_height": CAP_HEIGHT,
                   "sharding": "independent commits per GPU, no data-path collective",
                   "l2": "per-step working set 0.74 GB > 126 MB L2; two input batches alternated"},
        "roofline": roofline, "roofline_hbm": roofline_hbm,
        "stage_ms_per_step": {k: v / args.steps for k, v in stage_ms.items() if v},
        "cpu_baseline": cpu,
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": 8 * N_COLS * n,
                "d2h_bytes_per_step": 32 << CAP_HEIGHT, "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": int(launches), "clocks": clocks, "m2_lde_merkle_2p20x135": m2,
        "m1_synthetic_proofs": m1,
    }
    emit(json.dumps(line))
    ctx.close()
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
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-m2", action="store_true")
    ap.add_argument("--no-m1", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "cuda" else args.warmup
    with StdoutGuard() as g:
        GUARD = g
        if args.impl == "reference":
            return run_reference(args)
        return run_cuda(args)


if __name__ == "__main__":
    sys.exit(main())
