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
  int n_proofs = 1;              // plonky2 `prove` calls inside the job; 0 for AggregateJobs / Groth16
  int level = -1;                // index into Block::levels (the job's sub-group counter), -1 = no counter
};

// One sub-group counter of the proof store (city_rollup_common/src/qworker/proof_store.rs:60-89): value, goal and the
// jobs to enqueue when the value reaches the goal.
struct Level {
  std::atomic<uint32_t> counter{0};
  uint32_t goal = 0;
  std::vector<int> next_jobs;  // job indices
};

struct Block {
  uint64_t checkpoint_id = 0;
  std::vector<Job> jobs;
  std::deque<Level> levels;  // deque: Level holds an atomic (not movable)
  std::vector<int> entry_jobs;
  std::map<std::string, int> level_of_key;  // counter key (id with task = 0, data type Counter, index 0) -> level
  std::map<std::string, int> job_of_key;
};

constexpr uint8_t TOPIC_PROOF = 0, TOPIC_NOTIFY = 3, TOPIC_AGGREGATE = 4, DT_WITNESS = 0, DT_COUNTER = 16;
constexpr uint8_t CIRCUIT_GROTH16 = 36, CIRCUIT_NONE = 255;

std::array<uint8_t, 24> job_id(uint8_t topic, uint64_t goal, uint8_t circuit, uint32_t group, uint32_t sub_group, uint32_t task) {
  std::array<uint8_t, 24> b{};
  b[0] = topic;
  memcpy(&b[1], &goal, 8);
  b[9] = circuit;
  memcpy(&b[10], &group, 4);
  memcpy(&b[14], &sub_group, 4);
  memcpy(&b[18], &task, 4);
  b[22] = DT_WITNESS;  // ProvingJobDataType::InputWitness (get_output_id: OutputProof = 8)
  b[23] = 0;
  return b;
}
std::string key_of(const std::array<uint8_t, 24>& id) { return std::string((const char*)id.data(), 24); }
// QProvingJobDataID::get_sub_group_counter_id (job_id.rs:542-549): task 0, data type Counter, index 0
std::string counter_key(std::array<uint8_t, 24> id) {
  memset(&id[18], 0, 4);
  id[22] = DT_COUNTER;
  id[23] = 0;
  return key_of(id);
}

// plonky2 `prove` calls inside a job of this circuit type (from the circuit code): op leaves and
// per-op aggregates 1; block aggregators and the state transition 1 + a two-step minifier chain; sighash introspection
// ~5; final GL 1 + chain; the Groth16 wrapper (36) is out of scope (north_star) and AggregateJobs prove nothing.
int proofs_of(uint8_t topic, uint8_t circuit) {
  if (topic != TOPIC_PROOF || circuit == CIRCUIT_GROTH16) return 0;
  if (circuit <= PROCESS_L1_WITHDRAWAL_AGG) return 1;
  if (circuit == SIGHASH_INTROSPECTION) return 5;
  return 3;
}

int add_job(Block& blk, const std::array<uint8_t, 24>& id_in) {
  std::array<uint8_t, 24> id = id_in;
  id[22] = DT_WITNESS;
  id[23] = 0;
  auto it = blk.job_of_key.find(key_of(id));
  if (it != blk.job_of_key.end()) return it->second;
  Job j;
  j.id = id;
  j.n_proofs = proofs_of(id[0], id[9]);
  auto lv = blk.level_of_key.find(counter_key(id));
  j.level = lv == blk.level_of_key.end() ? -1 : lv->second;
  blk.jobs.push_back(j);
  blk.job_of_key[key_of(id)] = (int)blk.jobs.size() - 1;
  return (int)blk.jobs.size() - 1;
}
int add_level_key(Block& blk, const std::array<uint8_t, 24>& any_id_of_the_sub_group) {
  const std::string k = counter_key(any_id_of_the_sub_group);
  auto it = blk.level_of_key.find(k);
  if (it != blk.level_of_key.end()) return it->second;
  blk.levels.emplace_back();
  blk.level_of_key[k] = (int)blk.levels.size() - 1;
  return (int)blk.levels.size() - 1;
}

// ---- the job DAG of a dumped block: bincode BlockProofStoreDump (city_rollup_core_worker_qbench/src/dump.rs:16-27) =
// DumpProofStoreConfig {checkpoint_id u64, rpc_node_id u32, CityOpJobConfig 6 x u64} then SimpleProofStoreMemory
// {proofs: map<[u8; 24], Vec<u8>>, counters: map}.  The store's Counter entries carry, per sub-group, the goal (index 1,
// u32 LE) and the next-job list (index 2, bincode Vec<[u8; 24]>); every key with data type InputWitness is a job.  Works
// on the full qbench_data/example.bin and on tests/golden/example_dag.bin (the same file with the witness / proof
// payloads stripped).  Entry jobs = the proving jobs no next-job list mentions (what plan_jobs returns as leaves,
// qbench.rs:44-52).
void load_dump(Block& blk, const char* path, uint64_t checkpoint_override) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error(std::string("cannot open ") + path);
  std::vector<uint8_t> d((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  size_t off = 0;
  auto need = [&](size_t n) {
    if (off + n > d.size()) throw std::runtime_error("dump truncated");
  };
  auto u64_at = [&]() {
    need(8);
    uint64_t v;
    memcpy(&v, &d[off], 8);
    off += 8;
    return v;
  };
  blk.checkpoint_id = u64_at();
  need(4 + 48);
  off += 4 + 48;  // rpc_node_id, CityOpJobConfig
  const uint64_t n_entries = u64_at();
  struct Entry { std::array<uint8_t, 24> key; size_t off, len; };
  std::vector<Entry> entries;
  for (uint64_t i = 0; i < n_entries; i++) {
    Entry e;
    need(24);
    memcpy(e.key.data(), &d[off], 24);
    off += 24;
    e.len = (size_t)u64_at();
    need(e.len);
    e.off = off;
    off += e.len;
    entries.push_back(e);
  }
  // levels first (goals), then jobs, then the next lists (they may name jobs without a witness: AggregateJobs, notify)
  for (auto& e : entries)
    if (e.key[22] == DT_COUNTER && e.key[23] == 1) {
      if (e.len != 4) throw std::runtime_error("dump: a counter goal is not a u32");
      uint32_t g;
      memcpy(&g, &d[e.off], 4);
      blk.levels[add_level_key(blk, e.key)].goal = g;
    }
  for (auto& e : entries)
    if (e.key[22] == DT_WITNESS && e.key[0] == TOPIC_PROOF) add_job(blk, e.key);
  std::vector<char> mentioned;
  for (auto& e : entries)
    if (e.key[22] == DT_COUNTER && e.key[23] == 2) {
      if (e.len < 8) throw std::runtime_error("dump: truncated next-job list");
      uint64_t m;
      memcpy(&m, &d[e.off], 8);
      if (e.len != 8 + 24 * m) throw std::runtime_error("dump: next-job list length mismatch");
      Level& lv = blk.levels[add_level_key(blk, e.key)];
      for (uint64_t k = 0; k < m; k++) {
        std::array<uint8_t, 24> id;
        memcpy(id.data(), &d[e.off + 8 + 24 * k], 24);
        const int j = add_job(blk, id);
        lv.next_jobs.push_back(j);
        if ((size_t)j >= mentioned.size()) mentioned.resize(j + 1, 0);
        mentioned[j] = 1;
      }
    }
  mentioned.resize(blk.jobs.size(), 0);
  for (size_t j = 0; j < blk.jobs.size(); j++)
    if (!mentioned[j] && blk.jobs[j].id[0] == TOPIC_PROOF) blk.entry_jobs.push_back((int)j);
  if (checkpoint_override) {
    // several copies of the block in flight: give each its own checkpoint id (bytes 1..9 of every id)
    blk.checkpoint_id = checkpoint_override;
    for (auto& j : blk.jobs) memcpy(&j.id[1], &checkpoint_override, 8);
  }
}

// ---- built-in plans, expressed in the store's own terms (counters with goals and next-job lists; a job with several
// parent levels waits behind an AggregateJobs counter, exactly how the reference's planner joins sub-trees)
int add_level(Block& blk, uint8_t circuit, uint32_t group, uint32_t sub_group, int n_jobs, int proofs_per_job, std::vector<int>* jobs_out) {
  const int li = add_level_key(blk, job_id(TOPIC_PROOF, blk.checkpoint_id, circuit, group, sub_group, 0));
  blk.levels[li].goal = (uint32_t)n_jobs;
  for (int t = 0; t < n_jobs; t++) {
    const int j = add_job(blk, job_id(TOPIC_PROOF, blk.checkpoint_id, circuit, group, sub_group, (uint32_t)t));
    blk.jobs[j].n_proofs = proofs_per_job;
    if (jobs_out) jobs_out->push_back(j);
  }
  return li;
}

// one block in the shape of qbench_data/example.bin (used when no dump is given)
void plan_block(Block& blk, uint64_t checkpoint_id) {
  blk.checkpoint_id = checkpoint_id;
  struct Op { uint8_t leaf, agg; int n; };
  const Op ops[6] = {{REGISTER_USER, REGISTER_USER_AGG, 4}, {CLAIM_L1_DEPOSIT, CLAIM_L1_DEPOSIT_AGG, 2},
                     {TRANSFER_L2, TRANSFER_L2_AGG, 4}, {ADD_L1_WITHDRAWAL, ADD_L1_WITHDRAWAL_AGG, 4},
                     {PROCESS_L1_WITHDRAWAL, PROCESS_L1_WITHDRAWAL_AGG, 4}, {ADD_L1_DEPOSIT, ADD_L1_DEPOSIT_AGG, 2}};
  // joins: AggregateJobs counters (topic 4, circuit 255) with goal = number of parents
  auto join = [&](uint32_t group, uint32_t n_parents, const std::vector<int>& next) {
    const int li = add_level_key(blk, job_id(TOPIC_AGGREGATE, checkpoint_id, CIRCUIT_NONE, group, 0, 0));
    blk.levels[li].goal = n_parents;
    blk.levels[li].next_jobs = next;
    std::vector<int> tokens;
    for (uint32_t t = 0; t < n_parents; t++) tokens.push_back(add_job(blk, job_id(TOPIC_AGGREGATE, checkpoint_id, CIRCUIT_NONE, group, 0, t)));
    return tokens;
  };
  std::vector<int> agg1_jobs, agg2_jobs, st_jobs, sh_jobs, gl_jobs;
  const int agg1 = add_level(blk, AGG_PART_1, 100, 0, 1, 3, &agg1_jobs);  // block aggregators: prove + two minifier wrappers
  const int agg2 = add_level(blk, AGG_PART_2, 101, 0, 1, 3, &agg2_jobs);
  const int st = add_level(blk, ROLLUP_STATE_TRANSITION, 102, 0, 1, 3, &st_jobs);
  const int sh = add_level(blk, SIGHASH_INTROSPECTION, 103, 0, 3, 5, &sh_jobs);
  const int gl = add_level(blk, FINAL_SIGHASH_GL, 104, 0, 3, 3, &gl_jobs);
  const std::vector<int> tok1 = join(11, 3, agg1_jobs), tok2 = join(12, 3, agg2_jobs), tok_st = join(6, 2, st_jobs);
  for (int o = 0; o < 6; o++) {
    std::vector<int> cur;
    int prev = add_level(blk, ops[o].leaf, (uint32_t)o, 0, ops[o].n, 1, &cur);
    for (int j : cur) blk.entry_jobs.push_back(j);
    uint32_t sub = 1;
    for (int n = ops[o].n / 2; n >= 1; n /= 2, sub++) {  // binary aggregation tree over the op's leaves
      std::vector<int> nxt;
      const int lv = add_level(blk, ops[o].agg, (uint32_t)o, sub, n, 1, &nxt);
      blk.levels[prev].next_jobs = nxt;
      prev = lv;
    }
    blk.levels[prev].next_jobs = {o < 3 ? tok1[o] : tok2[o - 3]};  // part 1: register / claim / transfer, part 2: the rest
  }
  blk.levels[agg1].next_jobs = {tok_st[0]};
  blk.levels[agg2].next_jobs = {tok_st[1]};
  blk.levels[st].next_jobs = sh_jobs;
  blk.levels[sh].next_jobs = gl_jobs;
  blk.levels[gl].next_jobs = {add_job(blk, job_id(TOPIC_NOTIFY, checkpoint_id, CIRCUIT_NONE, 0, 0, 0))};
}

// BASELINE.json configs[4]: a binary aggregation tree over 2^k leaf proofs — 2^k leaf jobs (circuit 6, an L2 transfer)
// and 2^k - 1 two-verifier aggregation jobs (circuit 7), level-synchronous exactly like the reference's tree prover
// (city_common_circuit/src/treeprover/: every level waits for the one below); one `prove` per job
void plan_agg_tree(Block& blk, uint64_t checkpoint_id, int log_leaves) {
  blk.checkpoint_id = checkpoint_id;
  std::vector<int> cur;
  int prev = add_level(blk, TRANSFER_L2, 2, 0, 1 << log_leaves, 1, &cur);
  blk.entry_jobs = cur;
  uint32_t sub = 1;
  for (int n = 1 << (log_leaves - 1); n >= 1; n /= 2, sub++) {
    std::vector<int> nxt;
    const int lv = add_level(blk, TRANSFER_L2_AGG, 2, sub, n, 1, &nxt);
    blk.levels[prev].next_jobs = nxt;
    prev = lv;
  }
  blk.levels[prev].next_jobs = {add_job(blk, job_id(TOPIC_NOTIFY, checkpoint_id, CIRCUIT_NONE, 0, 0, 0))};
}

std::string hex(const std::array<uint8_t, 24>& b) {
  static const char* d = "0123456789abcdef";
  std::string s;
  for (uint8_t c : b) s += d[c >> 4], s += d[c & 15];
  return s;
}

}  // namespace

// the worker's decision after a job, in the store's terms (city_rollup_core_worker/src/actors/simple.rs:97-105):
// goal = store.get_goal_by_job_id(job); if goal != 0 and inc_counter == goal, enqueue store.get_next_jobs_by_job_id(job)
template <class Enqueue>
void after_job(Block& blk, const Job& job, Enqueue&& enqueue) {
  if (job.level < 0) return;
  Level& lv = blk.levels[job.level];
  if (lv.goal != 0 && lv.counter.fetch_add(1) + 1 == lv.goal)
    for (int nx : lv.next_jobs) enqueue(nx);
}

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

    size_t total_jobs = 0, total_proofs = 0, proving_jobs = 0;
    for (auto& b : blocks)
      for (auto& j : b.jobs) total_jobs++, total_proofs += (size_t)j.n_proofs, proving_jobs += j.n_proofs ? 1 : 0;

    // the ready queue (WorkerEventReceiverSync::wait_for_next_job / enqueue_jobs) and the in-memory proof store
    std::mutex qm, sm;
    std::condition_variable qcv;
    std::deque<std::pair<int, int>> ready;  // (block, job)
    std::map<std::string, std::vector<uint8_t>> store;
    std::atomic<size_t> jobs_done{0};
    std::atomic<long long> t_done_ns{0};
    std::atomic<int> mismatches{0}, warm{0};
    std::atomic<bool> go{false};
    struct Bench { std::array<uint8_t, 24> id; uint64_t ms; };
    const int n_workers = async_depth ? n_gpus : n_gpus * ctx_per_gpu;
    std::vector<std::vector<Bench>> bench(n_workers);
    std::vector<double> busy(n_workers, 0.0);

    auto enqueue_job = [&](int b, int j) {
      std::lock_guard<std::mutex> g(qm);
      ready.emplace_back(b, j);
      qcv.notify_one();
    };
    // everything a worker does after the proof(s) of a job exist
    auto finish_job = [&](int w, int b, int ji, std::vector<uint8_t>&& bytes, double sec) {
      Block& blk = blocks[b];
      Job& job = blk.jobs[ji];
      if (job.n_proofs) {
        {  // store.set_proof_by_id(job_id.get_output_id(), &proof)
          auto out_id = job.id;
          out_id[22] = 8;
          std::lock_guard<std::mutex> g(sm);
          store[hex(out_id)] = std::move(bytes);
        }
        bench[w].push_back({job.id, (uint64_t)(sec * 1e3)});  // start_time.elapsed().as_millis()
        busy[w] += sec;
      }
      if (job.id[0] != TOPIC_NOTIFY) after_job(blk, job, [&](int nx) { enqueue_job(b, nx); });
      if (jobs_done.fetch_add(1) + 1 == total_jobs) {
        // NotifyOrchestratorComplete of the last block: the replay's clock stops HERE (qbench.rs:44-60 stops its timer when
        // the job loop returns), not when the worker threads have torn their contexts down — freeing 24 contexts' pinned
        // buffers, graphs and device memory takes seconds and was counted as proving time before
        t_done_ns.store(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count());
        std::lock_guard<std::mutex> g(qm);
        qcv.notify_all();
      }
    };
    auto pop_job = [&](std::pair<int, int>& item, bool block) {
      std::unique_lock<std::mutex> g(qm);
      if (block) qcv.wait(g, [&] { return !ready.empty() || jobs_done.load() == total_jobs; });
      if (ready.empty()) return false;
      item = ready.front();
      ready.pop_front();
      return true;
    };

    std::vector<std::thread> pool;
    std::string first_error;
    auto fail_all = [&](const std::exception& e) {
      std::lock_guard<std::mutex> g(qm);
      if (first_error.empty()) first_error = e.what();
      jobs_done = total_jobs;
      qcv.notify_all();
    };
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
              if (in_flight == 0 && jobs_done.load() == total_jobs) {
                fprintf(stderr, "driver %d: submit %.3f s, collect %.3f s, store / queue %.3f s\n", w, t_submit, t_collect, t_finish);
                break;
              }
              if (!progressed) {
                if (in_flight == 0) {
                  std::pair<int, int> item;
                  if (!pop_job(item, true)) break;  // sleep until a job arrives or everything is done
                  std::lock_guard<std::mutex> g(qm);
                  ready.push_front(item);
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
    while (warm.load() < n_workers && first_error.empty()) std::this_thread::yield();
    const auto t0 = std::chrono::steady_clock::now();
    for (int b = 0; b < n_blocks; b++)
      for (int j : blocks[b].entry_jobs) enqueue_job(b, j);
    go = true;
    for (auto& th : pool) th.join();
    const double wall_join = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    const double wall = t_done_ns.load()
        ? (double)(t_done_ns.load() - std::chrono::duration_cast<std::chrono::nanoseconds>(t0.time_since_epoch()).count()) * 1e-9
        : wall_join;
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
