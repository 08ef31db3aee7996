// qworker.hpp — the worker side of City Rollup's proving-job protocol, in process: the job DAG of a block as the proof store
// holds it, the ready queue, the in-memory proof store and the benchmark records.  Header-only C++17, no CUDA in here: the
// prover is whatever the caller runs between WorkQueue::pop and WorkQueue::finish (tools/qbench_replay.cpp: G GPUs x W
// p2b contexts, blocking or driven asynchronously; a sleeping stand-in for the CPU test of the scheduler).
//
// What it mirrors in the reference:
//   * job ids (24 bytes) and their counter / output forms — city_rollup_common/src/qworker/job_id.rs:100-165, 542-549;
//   * the store's level counters: value, goal and next-job list per sub-group, "increment, and when the counter reaches the
//     goal enqueue the next jobs" — city_rollup_common/src/qworker/proof_store.rs:60-89 and the worker's use of it,
//     city_rollup_core_worker/src/actors/simple.rs:57-113;
//   * the dumped block (bincode BlockProofStoreDump) — city_rollup_core_worker_qbench/src/dump.rs:16-27;
//   * the proof store itself (SimpleProofStoreMemory: proofs as bincode bytes under the job's output id) —
//     city_store/.../memory_proof_store/mod.rs:11-46;
//   * the queue every worker pops from (WorkerEventReceiverSync::wait_for_next_job / enqueue_jobs; many l2-worker processes
//     against one Redis queue in production, city_rollup_worker_dispatch/src/implementations/redis/mod.rs:109);
//   * QWorkerJobBenchmark { job_id, duration } — job_id.rs:194-202.
#pragma once
#include <array>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <iterator>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace plonky2_b200 {
namespace qworker {


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

inline std::array<uint8_t, 24> job_id(uint8_t topic, uint64_t goal, uint8_t circuit, uint32_t group, uint32_t sub_group, uint32_t task) {
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
inline std::string key_of(const std::array<uint8_t, 24>& id) { return std::string((const char*)id.data(), 24); }
// QProvingJobDataID::get_sub_group_counter_id (job_id.rs:542-549): task 0, data type Counter, index 0
inline std::string counter_key(std::array<uint8_t, 24> id) {
  memset(&id[18], 0, 4);
  id[22] = DT_COUNTER;
  id[23] = 0;
  return key_of(id);
}

// plonky2 `prove` calls inside a job of this circuit type (from the circuit code): op leaves and
// per-op aggregates 1; block aggregators and the state transition 1 + a two-step minifier chain; sighash introspection
// ~5; final GL 1 + chain; the Groth16 wrapper (36) is out of scope (north_star) and AggregateJobs prove nothing.
inline int proofs_of(uint8_t topic, uint8_t circuit) {
  if (topic != TOPIC_PROOF || circuit == CIRCUIT_GROTH16) return 0;
  if (circuit <= PROCESS_L1_WITHDRAWAL_AGG) return 1;
  if (circuit == SIGHASH_INTROSPECTION) return 5;
  return 3;
}

inline int add_job(Block& blk, const std::array<uint8_t, 24>& id_in) {
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
inline int add_level_key(Block& blk, const std::array<uint8_t, 24>& any_id_of_the_sub_group) {
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
inline void load_dump(Block& blk, const char* path, uint64_t checkpoint_override) {
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
inline int add_level(Block& blk, uint8_t circuit, uint32_t group, uint32_t sub_group, int n_jobs, int proofs_per_job, std::vector<int>* jobs_out) {
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
inline void plan_block(Block& blk, uint64_t checkpoint_id) {
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
inline void plan_agg_tree(Block& blk, uint64_t checkpoint_id, int log_leaves) {
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

inline std::string hex(const std::array<uint8_t, 24>& b) {
  static const char* d = "0123456789abcdef";
  std::string s;
  for (uint8_t c : b) s += d[c >> 4], s += d[c & 15];
  return s;
}



// the worker's decision after a job, in the store's terms (city_rollup_core_worker/src/actors/simple.rs:97-105):
// goal = store.get_goal_by_job_id(job); if goal != 0 and inc_counter == goal, enqueue store.get_next_jobs_by_job_id(job)
template <class Enqueue>
void after_job(Block& blk, const Job& job, Enqueue&& enqueue) {
  if (job.level < 0) return;
  Level& lv = blk.levels[job.level];
  if (lv.goal != 0 && lv.counter.fetch_add(1) + 1 == lv.goal)
    for (int nx : lv.next_jobs) enqueue(nx);
}

// The ready queue, the proof store and the benchmark records of one replay: what the reference spreads over Redis, the proof
// store and the qbench harness.  Workers (any number of threads) call pop / finish; the clock of a run stops when the last
// job has been processed (NotifyOrchestratorComplete of the last block), not when the workers have torn their contexts down.
class WorkQueue {
 public:
  struct Bench { std::array<uint8_t, 24> id; uint64_t ms; };

  WorkQueue(std::deque<Block>& blocks, int n_workers) : blocks_(blocks), bench_(n_workers), busy_(n_workers, 0.0) {
    for (auto& b : blocks_)
      for (auto& j : b.jobs) total_jobs_++, total_proofs_ += (size_t)j.n_proofs, proving_jobs_ += j.n_proofs ? 1 : 0;
  }
  size_t total_jobs() const { return total_jobs_; }
  size_t total_proofs() const { return total_proofs_; }
  size_t proving_jobs() const { return proving_jobs_; }
  bool all_done() const { return jobs_done_.load() == total_jobs_; }
  Job& job(const std::pair<int, int>& item) { return blocks_[item.first].jobs[item.second]; }

  // enqueue the entry jobs of every block and start the clock
  void start() {
    t0_ = std::chrono::steady_clock::now();
    for (size_t b = 0; b < blocks_.size(); b++)
      for (int j : blocks_[b].entry_jobs) enqueue((int)b, j);
  }
  void enqueue(int b, int j) {
    std::lock_guard<std::mutex> g(qm_);
    ready_.emplace_back(b, j);
    qcv_.notify_one();
  }
  void push_front(const std::pair<int, int>& item) {
    std::lock_guard<std::mutex> g(qm_);
    ready_.push_front(item);
  }
  // the next ready job; block = sleep until one arrives or everything is done.  false = nothing (left) to do
  bool pop(std::pair<int, int>& item, bool block) {
    std::unique_lock<std::mutex> g(qm_);
    if (block) qcv_.wait(g, [&] { return !ready_.empty() || all_done(); });
    if (ready_.empty()) return false;
    item = ready_.front();
    ready_.pop_front();
    return true;
  }
  // everything a worker does after the proof(s) of a job exist: store.set_proof_by_id(job_id.get_output_id(), &proof), the
  // benchmark record (start_time.elapsed().as_millis()), the level counter and the next jobs (simple.rs:83-105)
  void finish(int worker, const std::pair<int, int>& item, std::vector<uint8_t>&& bytes, double sec) {
    Block& blk = blocks_[item.first];
    Job& jb = blk.jobs[item.second];
    if (jb.n_proofs) {
      {
        auto out_id = jb.id;
        out_id[22] = 8;  // ProvingJobDataType::OutputProof
        std::lock_guard<std::mutex> g(sm_);
        store_[hex(out_id)] = std::move(bytes);
      }
      bench_[worker].push_back({jb.id, (uint64_t)(sec * 1e3)});
      busy_[worker] += sec;
    }
    if (jb.id[0] != TOPIC_NOTIFY) after_job(blk, jb, [&](int nx) { enqueue(item.first, nx); });
    if (jobs_done_.fetch_add(1) + 1 == total_jobs_) {
      t_done_ns_.store(std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0_).count());
      std::lock_guard<std::mutex> g(qm_);
      qcv_.notify_all();
    }
  }
  // a worker died: release everybody, keep the first message
  void fail(const std::exception& e) {
    std::lock_guard<std::mutex> g(qm_);
    if (first_error_.empty()) first_error_ = e.what();
    failed_ = true;
    jobs_done_ = total_jobs_;
    qcv_.notify_all();
  }
  bool failed() const { return failed_.load(); }
  const std::string& first_error() const { return first_error_; }  // after the workers have been joined
  // seconds from start() to the last processed job (0 while running)
  double wall_seconds() const { return (double)t_done_ns_.load() * 1e-9; }
  double seconds_since_start() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count(); }
  const std::vector<std::vector<Bench>>& bench() const { return bench_; }
  const std::vector<double>& busy() const { return busy_; }
  const std::map<std::string, std::vector<uint8_t>>& store() const { return store_; }

 private:
  std::deque<Block>& blocks_;
  std::mutex qm_, sm_;
  std::condition_variable qcv_;
  std::deque<std::pair<int, int>> ready_;  // (block, job)
  std::map<std::string, std::vector<uint8_t>> store_;
  std::atomic<size_t> jobs_done_{0};
  std::atomic<long long> t_done_ns_{0};
  std::atomic<bool> failed_{false};
  std::chrono::steady_clock::time_point t0_{};
  size_t total_jobs_ = 0, total_proofs_ = 0, proving_jobs_ = 0;
  std::vector<std::vector<Bench>> bench_;
  std::vector<double> busy_;
  std::string first_error_;
};

}  // namespace qworker
}  // namespace plonky2_b200
