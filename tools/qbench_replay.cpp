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
// This tool keeps that protocol and replaces the worker: G GPUs x W contexts per GPU, one OS thread and one p2b
// context each, all consuming ONE ready queue (the reference's many l2-worker processes against one Redis queue,
// city_rollup_worker_dispatch/src/implementations/redis/mod.rs:109).  The block has the job structure decoded from
// qbench_data/example.bin (CityOpJobConfig {register 4, claim 2, transfer 4, add_withdrawal 4,
// process_withdrawal 4, add_deposit 2} -> 20 op leaves, 14 per-op aggregates, 2 block aggregators, the state
// transition, 3 sighash introspections, 3 final-GL jobs: 43 plonky2 jobs, 67 `prove` calls; the 3 Groth16 wrappers are
// out of scope).  The circuits themselves cannot be built here (no Rust CircuitBuilder), so every `prove` call of a
// job proves the synthetic City-shaped circuit of the case file (2^12 rows x 135 wires, the recursion gate set, FRI
// parameters of the stored proofs) and every proof is compared word for word with the expected one; the proof bytes
// that go into the store are produced by p2b_proof_to_bincode.  Several blocks can be in flight at once (-n), which
// is how the orchestrator keeps eight GPUs busy.
//
// Output: -o FILE gets the reference's benchmark JSON ([{"job_id": "<48 hex>", "duration": <ms>}, ...]); stdout gets
// one JSON summary line (proofs/s, jobs/s, wall time, mismatches).
//
// Build: g++ -O2 -std=c++17 -I. tools/qbench_replay.cpp -Lcity_rollup_b200 -lp2b -lpthread -o tools/qbench_replay
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <map>
#include <mutex>
#include <thread>

#include "tools/prove_case.hpp"

namespace {

// ProvingJobCircuitType values (city_rollup_common/src/qworker/job_id.rs:100-165)
enum Circuit : uint8_t {
  REGISTER_USER = 0, REGISTER_USER_AGG = 1, ADD_L1_DEPOSIT = 2, ADD_L1_DEPOSIT_AGG = 3, CLAIM_L1_DEPOSIT = 4,
  CLAIM_L1_DEPOSIT_AGG = 5, TRANSFER_L2 = 6, TRANSFER_L2_AGG = 7, ADD_L1_WITHDRAWAL = 8, ADD_L1_WITHDRAWAL_AGG = 9,
  PROCESS_L1_WITHDRAWAL = 10, PROCESS_L1_WITHDRAWAL_AGG = 11, ROLLUP_STATE_TRANSITION = 32, SIGHASH_INTROSPECTION = 33,
  FINAL_SIGHASH_GL = 34, AGG_PART_1 = 40, AGG_PART_2 = 41,
};

struct Job {
  std::array<uint8_t, 24> id{};  // [topic u8][goal u64][circuit u8][group u32][sub_group u32][task u32][data_type u8][data_index u8]
  int n_proofs = 1;              // plonky2 `prove` calls inside the job
  int level = -1;                // index into Block::levels
};

struct Level {
  std::vector<int> jobs;        // job indices
  std::atomic<uint32_t> counter{0};
  std::vector<int> next_levels;  // enqueued when counter == goal; a level with several parents waits for all of them
  std::atomic<uint32_t> parents_done{0};
  uint32_t n_parents = 0;
};

struct Block {
  uint64_t checkpoint_id = 0;
  std::vector<Job> jobs;
  std::deque<Level> levels;  // deque: Level holds atomics (not movable)
  std::vector<int> entry_levels;
};

std::array<uint8_t, 24> job_id(uint64_t goal, uint8_t circuit, uint32_t group, uint32_t sub_group, uint32_t task) {
  std::array<uint8_t, 24> b{};
  b[0] = 0;  // QJobTopic::GenerateStandardProof
  memcpy(&b[1], &goal, 8);
  b[9] = circuit;
  memcpy(&b[10], &group, 4);
  memcpy(&b[14], &sub_group, 4);
  memcpy(&b[18], &task, 4);
  b[22] = 0;  // ProvingJobDataType::InputWitness (get_output_id: OutputProof = 8)
  b[23] = 0;
  return b;
}

int add_level(Block& blk, uint8_t circuit, uint32_t group, uint32_t sub_group, int n_jobs, int proofs_per_job) {
  blk.levels.emplace_back();
  const int li = (int)blk.levels.size() - 1;
  for (int t = 0; t < n_jobs; t++) {
    Job j;
    j.id = job_id(blk.checkpoint_id, circuit, group, sub_group, (uint32_t)t);
    j.n_proofs = proofs_per_job;
    j.level = li;
    blk.jobs.push_back(j);
    blk.levels[li].jobs.push_back((int)blk.jobs.size() - 1);
  }
  return li;
}

void link(Block& blk, int from, int to) {
  blk.levels[from].next_levels.push_back(to);
  blk.levels[to].n_parents++;
}

// one block in the shape of qbench_data/example.bin
void plan_block(Block& blk, uint64_t checkpoint_id) {
  blk.checkpoint_id = checkpoint_id;
  struct Op { uint8_t leaf, agg; int n; };
  const Op ops[6] = {{REGISTER_USER, REGISTER_USER_AGG, 4}, {CLAIM_L1_DEPOSIT, CLAIM_L1_DEPOSIT_AGG, 2},
                     {TRANSFER_L2, TRANSFER_L2_AGG, 4}, {ADD_L1_WITHDRAWAL, ADD_L1_WITHDRAWAL_AGG, 4},
                     {PROCESS_L1_WITHDRAWAL, PROCESS_L1_WITHDRAWAL_AGG, 4}, {ADD_L1_DEPOSIT, ADD_L1_DEPOSIT_AGG, 2}};
  const int agg = add_level(blk, AGG_PART_1, 100, 0, 1, 3);  // block aggregators: prove + two minifier wrappers
  const int agg2 = add_level(blk, AGG_PART_2, 101, 0, 1, 3);
  for (int o = 0; o < 6; o++) {
    int prev = add_level(blk, ops[o].leaf, (uint32_t)o, 0, ops[o].n, 1);
    blk.entry_levels.push_back(prev);
    uint32_t sub = 1;
    for (int n = ops[o].n / 2; n >= 1; n /= 2, sub++) {  // binary aggregation tree over the op's leaves
      const int lv = add_level(blk, ops[o].agg, (uint32_t)o, sub, n, 1);
      link(blk, prev, lv);
      prev = lv;
    }
    link(blk, prev, o < 3 ? agg : agg2);  // part 1: register / claim / transfer, part 2: withdrawals / deposits
  }
  const int st = add_level(blk, ROLLUP_STATE_TRANSITION, 102, 0, 1, 3);
  link(blk, agg, st);
  link(blk, agg2, st);
  const int sh = add_level(blk, SIGHASH_INTROSPECTION, 103, 0, 3, 5);
  link(blk, st, sh);
  const int gl = add_level(blk, FINAL_SIGHASH_GL, 104, 0, 3, 3);
  link(blk, sh, gl);
}

// BASELINE.json configs[4]: a binary aggregation tree over 2^k leaf proofs — 2^k leaf jobs (circuit 6, an L2 transfer)
// and 2^k - 1 two-verifier aggregation jobs (circuit 7), level-synchronous exactly like the reference's tree prover
// (city_common_circuit/src/treeprover/: every level waits for the one below); one `prove` per job
void plan_agg_tree(Block& blk, uint64_t checkpoint_id, int log_leaves) {
  blk.checkpoint_id = checkpoint_id;
  int prev = add_level(blk, TRANSFER_L2, 2, 0, 1 << log_leaves, 1);
  blk.entry_levels.push_back(prev);
  uint32_t sub = 1;
  for (int n = 1 << (log_leaves - 1); n >= 1; n /= 2, sub++) {
    const int lv = add_level(blk, TRANSFER_L2_AGG, 2, sub, n, 1);
    link(blk, prev, lv);
    prev = lv;
  }
}

std::string hex(const std::array<uint8_t, 24>& b) {
  static const char* d = "0123456789abcdef";
  std::string s;
  for (uint8_t c : b) s += d[c >> 4], s += d[c & 15];
  return s;
}

}  // namespace

int main(int argc, char** argv) {
  const char* case_path = nullptr;
  const char* out_path = nullptr;
  int n_gpus = 1, ctx_per_gpu = 8, n_blocks = 4, agg_tree = -1;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto val = [&]() -> const char* { return i + 1 < argc ? argv[++i] : ""; };
    if (a == "-i") case_path = val();
    else if (a == "-o") out_path = val();
    else if (a == "-n") n_blocks = atoi(val());
    else if (a == "--gpus") n_gpus = atoi(val());
    else if (a == "--contexts") ctx_per_gpu = atoi(val());
    else if (a == "--agg-tree") agg_tree = atoi(val());
  }
  if (!case_path || n_gpus < 1 || ctx_per_gpu < 1 || n_blocks < 1 || agg_tree == 0 || agg_tree > 20) {
    fprintf(stderr, "usage: %s -i case.bin [-o bench.json] [-n blocks=4] [--gpus G=1] [--contexts W=8] [--agg-tree log2_leaves]\n", argv[0]);
    return 2;
  }
  try {
    const Case cs = load_case(case_path);
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

    std::deque<Block> blocks(n_blocks);
    for (int b = 0; b < n_blocks; b++) {
      if (agg_tree > 0) plan_agg_tree(blocks[b], 4 + (uint64_t)b, agg_tree);
      else plan_block(blocks[b], 4 + (uint64_t)b);  // example.bin is checkpoint 4
    }
    size_t total_jobs = 0, total_proofs = 0;
    for (auto& b : blocks)
      for (auto& j : b.jobs) total_jobs++, total_proofs += (size_t)j.n_proofs;

    // the ready queue (WorkerEventReceiverSync::wait_for_next_job / enqueue_jobs) and the in-memory proof store
    std::mutex qm, sm;
    std::condition_variable qcv;
    std::deque<std::pair<int, int>> ready;  // (block, job)
    std::map<std::string, std::vector<uint8_t>> store;
    std::atomic<size_t> jobs_done{0};
    std::atomic<int> mismatches{0}, warm{0};
    std::atomic<bool> go{false};
    struct Bench { std::array<uint8_t, 24> id; uint64_t ms; };
    std::vector<std::vector<Bench>> bench(n_gpus * ctx_per_gpu);
    std::vector<double> busy(n_gpus * ctx_per_gpu, 0.0);

    auto enqueue_level = [&](int b, int lv) {
      std::lock_guard<std::mutex> g(qm);
      for (int j : blocks[b].levels[lv].jobs) ready.emplace_back(b, j);
      qcv.notify_all();
    };

    std::vector<std::thread> pool;
    std::string first_error;
    for (int w = 0; w < n_gpus * ctx_per_gpu; w++) {
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
          if (prove(ctx, circuit, constants_sigmas, cs.digest, wire_cols, cs.public_inputs, cs.params) != cs.expected) mismatches++;
          warm++;
          while (!go.load()) std::this_thread::yield();
          for (;;) {
            std::pair<int, int> item;
            {
              std::unique_lock<std::mutex> g(qm);
              qcv.wait(g, [&] { return !ready.empty() || jobs_done.load() == total_jobs; });
              if (ready.empty()) return;
              item = ready.front();
              ready.pop_front();
            }
            Block& blk = blocks[item.first];
            Job& job = blk.jobs[item.second];
            const auto t0 = std::chrono::steady_clock::now();
            std::vector<uint8_t> bytes;
            for (int p = 0; p < job.n_proofs; p++) {  // prover.worker_prove_mut(store, job_id)
              auto words = prove(ctx, circuit, constants_sigmas, cs.digest, wire_cols, cs.public_inputs, cs.params);
              if (words != cs.expected) mismatches++;
              if (p + 1 == job.n_proofs) bytes = proof_to_bincode(shape, cs.params, words);
            }
            {  // store.set_proof_by_id(job_id.get_output_id(), &proof)
              auto out_id = job.id;
              out_id[22] = 8;
              std::lock_guard<std::mutex> g(sm);
              store[hex(out_id)] = std::move(bytes);
            }
            const auto t1 = std::chrono::steady_clock::now();
            const double sec = std::chrono::duration<double>(t1 - t0).count();
            bench[w].push_back({job.id, (uint64_t)(sec * 1e3)});  // start_time.elapsed().as_millis()
            busy[w] += sec;
            // store.inc_counter_by_id(...) == goal  =>  enqueue_jobs(get_next_jobs_by_job_id(...))
            Level& lv = blk.levels[job.level];
            if (lv.counter.fetch_add(1) + 1 == lv.jobs.size()) {
              for (int nx : lv.next_levels)
                if (blk.levels[nx].parents_done.fetch_add(1) + 1 == blk.levels[nx].n_parents) enqueue_level(item.first, nx);
            }
            if (jobs_done.fetch_add(1) + 1 == total_jobs) {
              std::lock_guard<std::mutex> g(qm);
              qcv.notify_all();  // NotifyOrchestratorComplete of the last block
            }
          }
        } catch (const std::exception& e) {
          std::lock_guard<std::mutex> g(qm);
          if (first_error.empty()) first_error = e.what();
          jobs_done = total_jobs;
          qcv.notify_all();
        }
      });
    }
    while (warm.load() < n_gpus * ctx_per_gpu && first_error.empty()) std::this_thread::yield();
    const auto t0 = std::chrono::steady_clock::now();
    for (int b = 0; b < n_blocks; b++)
      for (int lv : blocks[b].entry_levels) enqueue_level(b, lv);
    go = true;
    for (auto& th : pool) th.join();
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!first_error.empty()) throw std::runtime_error(first_error);

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
    printf("{\"harness\": \"%s\", \"rows_log2\": %u, \"gpus\": %d, "
           "\"contexts_per_gpu\": %d, \"blocks\": %d, \"jobs\": %zu, \"jobs_recorded\": %zu, \"proofs\": %zu, \"wall_s\": %.4f, "
           "\"proofs_per_s\": %.2f, \"jobs_per_s\": %.2f, \"sum_job_duration_ms\": %.0f, \"worker_busy_fraction\": %.3f, "
           "\"stored_proofs\": %zu, \"stored_bytes\": %zu, \"mismatching_proofs\": %d}\n",
           agg_tree > 0 ? "binary aggregation tree, level-synchronous (synthetic City-shaped circuit)"
                        : "qbench replay (job DAG of qbench_data/example.bin, synthetic City-shaped circuit)",
           cs.desc.degree_bits, n_gpus, ctx_per_gpu, n_blocks, total_jobs, recorded, total_proofs, wall, total_proofs / wall, total_jobs / wall, sum_ms,
           busy_sum / (wall * n_gpus * ctx_per_gpu), store.size(), stored_bytes, mismatches.load());
    return (mismatches.load() || recorded != total_jobs || store.size() != total_jobs) ? 1 : 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
