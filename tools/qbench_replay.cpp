// qbench_replay.cpp — native multi-GPU job loop in the shape of the reference's qbench harness.
//
// The reference measures its workers by replaying a dumped block (`city-rollup-cli qbench -i qbench_data/example.bin
// -o out.json -n 1`, city_common/src/cli/args.rs:105-117): the job planner writes the block's proving jobs into the
// proof store as LEVELS — every level has a counter, a goal (= its size) and the list of next jobs
// (city_rollup_common/src/qworker/proof_store.rs:60-89) — a worker pops a job, proves, stores the proof with
// `bincode::serialize`, bumps the level counter and, when the counter reaches the goal, enqueues the next level
// (city_rollup_core_worker/src/actors/simple.rs:57-113); per job it records `QWorkerJobBenchmark { job_id, duration }`
// with the 24-byte id as hex and the duration in milliseconds (city_rollup_common/src/qworker/job_id.rs:194-202).
//
// This tool keeps that protocol (city_rollup_b200/cpp/qworker.hpp: job ids, level counters, dump reader, ready queue, proof
// store, benchmark records) and replaces the worker: G GPUs x W contexts per GPU, one OS thread and one p2b context each —
// or ONE thread per GPU driving its contexts through p2b_prove_submit_nowait / poll / collect (--async) — all consuming
// ONE ready queue (the reference's many l2-worker processes against one Redis queue,
// city_rollup_worker_dispatch/src/implementations/redis/mod.rs:109).  The block has the job structure decoded from
// qbench_data/example.bin (CityOpJobConfig {register 4, claim 2, transfer 4, add_withdrawal 4,
// process_withdrawal 4, add_deposit 2} -> 20 op leaves, 14 per-op aggregates, 2 block aggregators, the state
// transition, 3 sighash introspections, 3 final-GL jobs: 43 plonky2 jobs, 67 `prove` calls; the 3 Groth16 wrappers are
// out of scope).  The circuits themselves cannot be built here (no Rust CircuitBuilder), so every `prove` call of a
// job proves the synthetic City-shaped circuit of the case file (2^12 rows x 135 wires, the City op-circuit gate set, FRI
// parameters of the stored proofs) and every proof is compared word for word with the expected one; the proof bytes
// that go into the store are produced by p2b_proof_to_bincode.  Several blocks can be in flight at once (-n), which
// is how the orchestrator keeps eight GPUs busy.
//
// Output: -o FILE gets the reference's benchmark JSON ([{"job_id": "<48 hex>", "duration": <ms>}, ...]); stdout gets
// one JSON summary line (proofs/s, jobs/s, wall time from the first enqueue to the last processed job, mismatches).
//
// Build: g++ -O2 -std=c++17 -I. tools/qbench_replay.cpp -Lcity_rollup_b200 -lp2b -lpthread -o tools/qbench_replay
#include <atomic>
#include <cstdlib>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <iterator>
#include <map>
#include <memory>
#include <mutex>
#include <thread>

#include "city_rollup_b200/cpp/qworker.hpp"
#include "tools/prove_case.hpp"

using namespace plonky2_b200::qworker;

// --plan-only: no GPU.  Replays the DAG with zero-cost jobs on one thread and prints what a worker would see.
int plan_only(Block& blk, const char* source) {
  std::deque<int> q(blk.entry_jobs.begin(), blk.entry_jobs.end());
  size_t processed = 0, proving = 0, proofs = 0, groth16 = 0, aggregate = 0, notify = 0, max_ready = q.size();
  std::map<int, int> per_circuit;
  while (!q.empty()) {
    const int j = q.front();
    q.pop_front();
    const Job& job = blk.jobs[j];
    processed++;
    if (job.id[0] == TOPIC_NOTIFY) {
      notify++;
      continue;
    }
    if (job.id[0] == TOPIC_AGGREGATE) aggregate++;
    else if (job.id[9] == CIRCUIT_GROTH16) groth16++;
    else proving++, proofs += (size_t)job.n_proofs, per_circuit[job.id[9]]++;
    after_job(blk, job, [&](int nx) { q.push_back(nx); });
    if (q.size() > max_ready) max_ready = q.size();
  }
  printf("{\"source\": \"%s\", \"checkpoint_id\": %llu, \"jobs_in_store\": %zu, \"counters\": %zu, \"entry_jobs\": %zu, "
         "\"processed\": %zu, \"plonky2_jobs\": %zu, \"plonky2_proofs\": %zu, \"groth16_jobs\": %zu, \"aggregate_jobs\": %zu, "
         "\"notify_orchestrator_complete\": %zu, \"max_ready\": %zu, \"jobs_per_circuit\": {",
         source, (unsigned long long)blk.checkpoint_id, blk.jobs.size(), blk.levels.size(), blk.entry_jobs.size(), processed, proving,
         proofs, groth16, aggregate, notify, max_ready);
  bool first = true;
  for (auto& kv : per_circuit) printf("%s\"%d\": %d", first ? "" : ", ", kv.first, kv.second), first = false;
  printf("}}\n");
  return notify == 1 && processed == blk.jobs.size() ? 0 : 1;
}

int main(int argc, char** argv) {
  // more hardware work queues than the default 8: with more worker streams than queues, streams that share a queue
  // serialise behind each other (bench.py does the same; must be set before CUDA initialises)
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  const char* case_path = nullptr;
  const char* out_path = nullptr;
  const char* dump_path = nullptr;
  int n_gpus = 1, ctx_per_gpu = 8, n_blocks = 4, agg_tree = -1, async_depth = 0;
  double fake_ms = 0.0;
  bool only_plan = false;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto val = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "-i") case_path = val();
    else if (a == "-o") out_path = val();
    else if (a == "-d" || a == "--dump") dump_path = val();
    else if (a == "-n") n_blocks = atoi(val());
    else if (a == "--gpus") n_gpus = atoi(val());
    else if (a == "--contexts") ctx_per_gpu = atoi(val());
    else if (a == "--agg-tree") agg_tree = atoi(val());
    else if (a == "--async") async_depth = atoi(val());
    else if (a == "--plan-only") only_plan = true;
    else if (a == "--fake-ms") fake_ms = atof(val());
  }
  if ((!case_path && !only_plan && fake_ms <= 0.0) || n_gpus < 1 || ctx_per_gpu < 1 || n_blocks < 1 || agg_tree == 0 || agg_tree > 20 || async_depth < 0) {
    fprintf(stderr,
            "usage: %s -i case.bin [-d dump.bin] [-o bench.json] [-n blocks=4] [--gpus G=1] [--contexts W=8] [--agg-tree log2_leaves]\n"
            "          [--async K]   one host thread per GPU drives K contexts through p2b_prove_submit / collect\n"
            "          [--fake-ms X] no GPU: every proof is a sleep of X ms (the store protocol, queue and worker pool alone)\n"
            "          [--plan-only] parse the dump (or the built-in plan), replay the DAG without a GPU, print its census\n"
            "  -d: a bincode BlockProofStoreDump (qbench_data/example.bin, or tests/golden/example_dag.bin = the same with\n"
            "      witness / proof payloads stripped): the job ids, counters, goals and next-job lists come from the file\n",
            argv[0]);
    return 2;
  }
  try {
    std::deque<Block> blocks(n_blocks);
    for (int b = 0; b < n_blocks; b++) {
      if (agg_tree > 0) plan_agg_tree(blocks[b], 4 + (uint64_t)b, agg_tree);
      else if (dump_path) load_dump(blocks[b], dump_path, b == 0 ? 0 : 1000 + (uint64_t)b);
      else plan_block(blocks[b], 4 + (uint64_t)b);  // example.bin is checkpoint 4
    }
    if (only_plan) return plan_only(blocks[0], dump_path ? dump_path : (agg_tree > 0 ? "built-in aggregation tree" : "built-in block plan"));

    const Case cs = fake_ms > 0.0 ? Case{} : load_case(case_path);
    p2b_proof_shape shape{};
    shape.degree_bits = cs.desc.degree_bits;
    shape.num_constants = cs.desc.num_constants;
    shape.num_routed_wires = cs.desc.num_routed_wires;
    shape.num_wires = cs.desc.num_wires;
    shape.num_challenges = cs.desc.num_challenges;
    shape.num_partial_products = cs.desc.num_partial_products;
    shape.quotient_degree_factor = cs.desc.quotient_degree_factor;
    shape.constants_sigmas_cap_height = cs.params.cap_height;
    shape.n_public_inputs = (uint32_t)cs.public_inputs.size();

    // the ready queue, the in-memory proof store and the benchmark records (city_rollup_b200/cpp/qworker.hpp)
    std::atomic<int> mismatches{0}, warm{0};
    std::atomic<bool> go{false};
    const int n_workers = async_depth ? n_gpus : n_gpus * ctx_per_gpu;
    WorkQueue wq(blocks, n_workers);
    const size_t total_jobs = wq.total_jobs(), total_proofs = wq.total_proofs(), proving_jobs = wq.proving_jobs();
    auto finish_job = [&](int w, int b, int ji, std::vector<uint8_t>&& bytes, double sec) { wq.finish(w, {b, ji}, std::move(bytes), sec); };
    auto pop_job = [&](std::pair<int, int>& item, bool block) { return wq.pop(item, block); };
    auto fail_all = [&](const std::exception& e) { wq.fail(e); };

    std::vector<std::thread> pool;
    if (fake_ms > 0.0) {
      // no GPU: a job costs fake_ms per proof of sleep.  What is left is the store protocol, the ready queue and the
      // worker pool themselves — their overhead and the parallelism the DAG offers (CPU test of the scheduler)
      for (int w = 0; w < n_workers; w++) {
        pool.emplace_back([&, w] {
          warm++;
          while (!go.load()) std::this_thread::yield();
          std::pair<int, int> item;
          while (pop_job(item, true)) {
            Job& job = blocks[item.first].jobs[item.second];
            const auto t0 = std::chrono::steady_clock::now();
            if (job.n_proofs) std::this_thread::sleep_for(std::chrono::duration<double, std::milli>(fake_ms * job.n_proofs));
            const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            finish_job(w, item.first, item.second, std::vector<uint8_t>(job.n_proofs ? 8 : 0), sec);
          }
        });
      }
    } else if (!async_depth) {
      // one OS thread and one context per worker (the reference's model: many l2-worker processes, one queue)
      for (int w = 0; w < n_workers; w++) {
        pool.emplace_back([&, w] {
          try {
            Context ctx(w % n_gpus);
            if (ctx_per_gpu > 1) ctx.set_blocking_sync(true);
            CircuitData circuit(ctx, cs.desc, cs.gates, cs.k_is);
            PolynomialBatch constants_sigmas =
                PolynomialBatch::from_values(ctx, cs.cs_values, cs.params.rate_bits, false, cs.params.cap_height, true);
            // the worker's witness buffer: pinned, so that a proof's upload is one DMA
            PinnedColumns witness(ctx, cs.wire_values.size(), cs.wire_values[0].size());
            witness.fill(cs.wire_values);
            const std::vector<const F*>& wire_cols = witness.pointers();
            for (int k = 0; k < 2; k++)  // the second proof of a shape builds the context's prove plan
              if (prove(ctx, circuit, constants_sigmas, cs.digest, wire_cols, cs.public_inputs, cs.params) != cs.expected) mismatches++;
            warm++;
            while (!go.load()) std::this_thread::yield();
            std::pair<int, int> item;
            while (pop_job(item, true)) {
              Job& job = blocks[item.first].jobs[item.second];
              const auto t0 = std::chrono::steady_clock::now();
              std::vector<uint8_t> bytes;
              for (int p = 0; p < job.n_proofs; p++) {  // prover.worker_prove_mut(store, job_id)
                auto words = prove(ctx, circuit, constants_sigmas, cs.digest, wire_cols, cs.public_inputs, cs.params);
                if (words != cs.expected) mismatches++;
                if (p + 1 == job.n_proofs) bytes = proof_to_bincode(shape, cs.params, words);
              }
              const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
              finish_job(w, item.first, item.second, std::move(bytes), sec);
            }
          } catch (const std::exception& e) {
            fail_all(e);
          }
        });
      }
    } else {
      // ONE host thread per GPU keeps `async_depth` contexts busy through p2b_prove_submit / p2b_prove_collect: the
      // thread is free while the GPU proves (in the reference it would generate the next job's witness there), and the
      // contexts share one device copy of the constants|sigmas batch (p2b_batch_attach)
      for (int w = 0; w < n_workers; w++) {
        pool.emplace_back([&, w] {
          try {
            const int K = async_depth;
            std::vector<std::unique_ptr<Context>> ctxs;
            std::vector<std::unique_ptr<CircuitData>> circuits;
            std::vector<std::unique_ptr<PinnedColumns>> witnesses;
            for (int k = 0; k < K; k++) {
              ctxs.emplace_back(new Context(w));
              ctxs.back()->set_blocking_sync(false);
              circuits.emplace_back(new CircuitData(*ctxs[k], cs.desc, cs.gates, cs.k_is));
              witnesses.emplace_back(new PinnedColumns(*ctxs[k], cs.wire_values.size(), cs.wire_values[0].size()));
              witnesses.back()->fill(cs.wire_values);
            }
            PolynomialBatch owner = PolynomialBatch::from_values(*ctxs[0], cs.cs_values, cs.params.rate_bits, false, cs.params.cap_height, true);
            p2b_synchronize(ctxs[0]->get());
            std::vector<PolynomialBatch> views;
            for (int k = 1; k < K; k++) views.push_back(owner.attach(*ctxs[k]));
            auto cs_of = [&](int k) -> const PolynomialBatch& { return k == 0 ? owner : views[k - 1]; };
            for (int k = 0; k < K; k++)
              for (int r = 0; r < 2; r++)
                if (prove(*ctxs[k], *circuits[k], cs_of(k), cs.digest, witnesses[k]->pointers(), cs.public_inputs, cs.params) != cs.expected) mismatches++;
            warm++;
            while (!go.load()) std::this_thread::yield();
            struct Slot { bool busy = false; int b = 0, j = 0, left = 0; size_t len = 0; std::chrono::steady_clock::time_point t0; };
            double t_submit = 0, t_collect = 0, t_finish = 0;  // where the driver thread's own time goes (stderr at the end)
            auto now = [] { return std::chrono::steady_clock::now(); };
            auto since = [&](std::chrono::steady_clock::time_point a) { return std::chrono::duration<double>(now() - a).count(); };
            std::vector<Slot> slots(K);
            int in_flight = 0;
            for (;;) {
              bool progressed = false;
              for (int k = 0; k < K; k++) {
                Slot& sl = slots[k];
                if (sl.busy) {
                  if (!prove_poll(*ctxs[k])) continue;
                  auto tc = now();
                  auto words = prove_collect(*ctxs[k], sl.len);
                  t_collect += since(tc);
                  if (words != cs.expected) mismatches++;
                  progressed = true;
                  if (--sl.left > 0) {  // the job's next proof (minifier chain): dependent on this one in the reference
                    auto ts = now();
                    sl.len = prove_submit_nowait(*ctxs[k], *circuits[k], cs_of(k), cs.digest, witnesses[k]->pointers(), cs.public_inputs, cs.params);
                    t_submit += since(ts);
                    continue;
                  }
                  const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - sl.t0).count();
                  sl.busy = false;
                  in_flight--;
                  auto tf = now();
                  finish_job(w, sl.b, sl.j, proof_to_bincode(shape, cs.params, words), sec);
                  t_finish += since(tf);
                }
                if (!sl.busy) {
                  std::pair<int, int> item;
                  if (!pop_job(item, false)) continue;
                  Job& job = blocks[item.first].jobs[item.second];
                  progressed = true;
                  if (job.n_proofs == 0) {
                    finish_job(w, item.first, item.second, {}, 0.0);
                    continue;
                  }
                  sl = Slot{true, item.first, item.second, job.n_proofs, 0, std::chrono::steady_clock::now()};
                  auto ts = now();
                  sl.len = prove_submit_nowait(*ctxs[k], *circuits[k], cs_of(k), cs.digest, witnesses[k]->pointers(), cs.public_inputs, cs.params);
                  t_submit += since(ts);
                  in_flight++;
                }
              }
              if (in_flight == 0 && wq.all_done()) {
                fprintf(stderr, "driver %d: submit %.3f s, collect %.3f s, store / queue %.3f s\n", w, t_submit, t_collect, t_finish);
                break;
              }
              if (!progressed) {
                if (in_flight == 0) {
                  std::pair<int, int> item;
                  if (!pop_job(item, true)) break;  // sleep until a job arrives or everything is done
                  wq.push_front(item);
                } else {
                  std::this_thread::yield();
                }
              }
            }
          } catch (const std::exception& e) {
            fail_all(e);
          }
        });
      }
    }
    while (warm.load() < n_workers && !wq.failed()) std::this_thread::yield();
    wq.start();  // the entry jobs of every block; the clock runs from here to the last processed job
    go = true;
    for (auto& th : pool) th.join();
    const double wall_join = wq.seconds_since_start();
    if (wq.failed()) throw std::runtime_error(wq.first_error());
    const double wall = wq.wall_seconds() > 0.0 ? wq.wall_seconds() : wall_join;
    const auto& bench = wq.bench();
    const auto& busy = wq.busy();
    const auto& store = wq.store();

    size_t recorded = 0, stored_bytes = 0;
    double sum_ms = 0, busy_sum = 0;
    for (auto& v : bench) recorded += v.size();
    for (auto& v : bench)
      for (auto& e : v) sum_ms += (double)e.ms;
    for (double s : busy) busy_sum += s;
    for (auto& kv : store) stored_bytes += kv.second.size();
    if (out_path) {
      FILE* f = fopen(out_path, "w");
      if (!f) throw std::runtime_error(std::string("cannot write ") + out_path);
      fprintf(f, "[");
      bool first = true;
      for (auto& v : bench)
        for (auto& e : v) {
          fprintf(f, "%s\n  {\"job_id\": \"%s\", \"duration\": %llu}", first ? "" : ",", hex(e.id).c_str(), (unsigned long long)e.ms);
          first = false;
        }
      fprintf(f, "\n]\n");
      fclose(f);
    }
    const int slots_total = async_depth ? n_gpus * async_depth : n_workers;
    printf("{\"harness\": \"%s\", \"dag\": \"%s\", \"rows_log2\": %u, \"gpus\": %d, \"host_threads\": %d, "
           "\"contexts_per_gpu\": %d, \"mode\": \"%s\", \"blocks\": %d, \"jobs\": %zu, \"proving_jobs\": %zu, \"jobs_recorded\": %zu, "
           "\"proofs\": %zu, \"wall_s\": %.4f, \"wall_incl_teardown_s\": %.4f, "
           "\"proofs_per_s\": %.2f, \"jobs_per_s\": %.2f, \"sum_job_duration_ms\": %.0f, \"worker_busy_fraction\": %.3f, "
           "\"stored_proofs\": %zu, \"stored_bytes\": %zu, \"mismatching_proofs\": %d}\n",
           agg_tree > 0 ? "binary aggregation tree, level-synchronous (synthetic City-shaped circuit)"
                        : "qbench replay (synthetic City-shaped circuit)",
           dump_path ? dump_path : "built-in plan", cs.desc.degree_bits, n_gpus, n_workers, async_depth ? async_depth : ctx_per_gpu,
           fake_ms > 0.0 ? "no GPU: every proof is a sleep (--fake-ms), scheduler only"
                       : async_depth ? "one host thread per GPU, p2b_prove_submit / collect" : "one host thread per context, blocking p2b_prove",
           n_blocks, total_jobs, proving_jobs, recorded, total_proofs, wall, wall_join, total_proofs / wall, proving_jobs / wall, sum_ms,
           busy_sum / (wall * slots_total), store.size(), stored_bytes, mismatches.load());
    return (mismatches.load() || recorded != proving_jobs || store.size() != proving_jobs) ? 1 : 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
