"""N > 1 host logic on CPU (gloo, world size 2).  The path shards by independent proof jobs with no data-path
collective; what is specific to N > 1 is bench.py's own code: ProofFarm (worker threads, thread
barrier, per-context timers and launch counters) and timed_job_run / RankGroup (the cross-rank barrier, the
max-over-ranks time, the whole-job proofs/s).  This test runs exactly that code on two gloo ranks against a FAKE
device module (contexts that sleep instead of proving, a different speed per rank) and checks the arithmetic of the
reported number: every worker of every rank proves its share, launches are summed over ranks, and the slowest rank's
time is the denominator.  No oracle, no GPU."""
import os
import sys
import time

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

LAUNCHES_PER_PROOF = 45


class FakeBatch:
    def free(self):
        pass


class FakeContext:
    """timer / launch-count surface of city_rollup_b200.Context, wall-clock based"""

    def __init__(self, device=0):
        self.t0 = self.t1 = 0.0
        self.launches = 0
        self.proofs = 0

    def set_blocking_sync(self, on=True):
        pass

    def pinned_empty(self, shape):
        return np.empty(shape, np.uint64)

    def launch_count(self):
        return self.launches

    def timer_start(self):
        self.t0 = time.perf_counter()

    def timer_stop_ms(self):
        self.t1 = time.perf_counter()
        return (self.t1 - self.t0) * 1e3

    def timer_span_ms(self, last):
        return (last.t1 - self.t0) * 1e3

    def close(self):
        pass


class FakeModule:
    """what ProofFarm uses of the package; a proof = sleep(ms_per_proof)"""

    def __init__(self, ms_per_proof):
        self.ms = ms_per_proof
        self.contexts = []

    def Context(self, device=0):
        c = FakeContext(device)
        self.contexts.append(c)
        return c

    def FriParams(self, *a):
        return a

    def CircuitData(self, ctx, desc):
        return FakeBatch()

    class PolynomialBatch:
        @staticmethod
        def from_values(ctx, values, rate_bits, blinding, cap_height, keep_values=False):
            return FakeBatch()

    def _prove(self, ctx):
        time.sleep(self.ms * 1e-3)
        ctx.launches += LAUNCHES_PER_PROOF
        ctx.proofs += 1
        return np.zeros(7, np.uint64)

    def prove_native_device(self, ctx, cd, cs, digest, ptr, pis, params):
        return self._prove(ctx)

    def prove_native(self, ctx, cd, cs, digest, wires, pis, params, raw=True):
        assert raw
        return self._prove(ctx)


class FakeCircuit:
    def desc(self):
        return {}

    def constants_sigmas_values(self):
        return []

    def wire_values(self):
        return [np.arange(8, dtype=np.uint64) for _ in range(3)]


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    import bench

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ms_per_proof = 4.0 if rank == 0 else 9.0  # rank 1 is the slow GPU
    m = FakeModule(ms_per_proof)
    n_ctx, per_ctx, warm = 3, 5, 2
    farm = bench.ProofFarm(m, rank, n_ctx, FakeCircuit(), [1, 2, 3, 4], [5], to_device=lambda a: (a, 0))
    group = bench.RankGroup(world, "cpu")
    results = {}
    for mode in ("dev", "pinned", "pageable"):
        value, ms, launches, wall, cpu = bench.timed_job_run(farm, group, mode, per_ctx, warm)
        results[mode] = (value, ms, launches)
    proofs = [c.proofs for c in m.contexts]
    farm.close()
    out.put((rank, results, proofs))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_report_whole_job_throughput_over_the_slowest_rank():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(world):
        rank, results, proofs = q.get(timeout=180)
        got[rank] = (results, proofs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_ctx, per_ctx, warm = 3, 5, 2
    for rank in range(world):
        results, proofs = got[rank]
        # every worker proved warm + timed proofs in each of the three modes
        assert proofs == [3 * (warm + per_ctx)] * n_ctx
        for mode, (value, ms, launches) in results.items():
            # both ranks agree on the reduced numbers
            assert (value, ms, launches) == got[0][0][mode]
            # launches: summed over ranks, timed region only
            assert launches == world * n_ctx * per_ctx * LAUNCHES_PER_PROOF
            # the denominator is the SLOW rank's time: >= 5 proofs x 9 ms, and nowhere near rank 0's 20 ms
            assert ms >= per_ctx * 9.0 * 0.98
            assert ms < per_ctx * 9.0 * 2.5
            # whole-job throughput = all proofs of all ranks / that time
            assert abs(value - world * n_ctx * per_ctx / (ms * 1e-3)) < 1e-6 * value
