// prove_bench.cpp — native (C++17) job loop over the C ABI through the header-only mirror
// city_rollup_b200/cpp/plonky2_b200.hpp: T worker threads, one p2b context each (the reference's one-worker-
// per-process model, city_rollup_core_worker/src/actors/simple.rs:32-56, several workers per GPU), every worker
// proving the same synthetic City-shaped circuit in a loop.  Input: a case file written by
// tools/dump_prove_case.py (circuit description, constants|sigmas values, witness columns, expected proof
// words).  Checks every proof word for word against the expected one and prints one JSON line.
//
// Build: g++ -O2 -std=c++17 -I. tools/prove_bench.cpp -Lcity_rollup_b200 -lp2b -Wl,-rpath,'$ORIGIN/../city_rollup_b200' -lpthread -o tools/prove_bench_cpp
#include <atomic>
#include <cstdlib>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <thread>

#include "tools/prove_case.hpp"

int main(int argc, char** argv) {
  // more hardware work queues than the default 8: with more worker streams than queues, streams that share a queue
  // serialise behind each other (bench.py does the same; must be set before CUDA initialises)
  setenv("CUDA_DEVICE_MAX_CONNECTIONS", "32", 0);
  if (argc < 2) {
    fprintf(stderr, "usage: %s case.bin [threads=8] [proofs_per_thread=40] [device=0]\n", argv[0]);
    return 2;
  }
  try {
    Case cs = load_case(argv[1]);
    const int threads = argc > 2 ? atoi(argv[2]) : 8, per = argc > 3 ? atoi(argv[3]) : 40, device = argc > 4 ? atoi(argv[4]) : 0;
    std::atomic<int> mismatches{0}, ready{0};
    std::atomic<bool> go{false};
    std::vector<std::thread> pool;
    std::vector<double> secs(threads, 0.0);
    for (int t = 0; t < threads; t++) {
      pool.emplace_back([&, t] {
        Context ctx(device);
        if (threads > 1) ctx.set_blocking_sync(true);  // several workers per GPU: sleep while waiting, do not spin
        CircuitData circuit(ctx, cs.desc, cs.gates, cs.k_is);
        PolynomialBatch constants_sigmas =
            PolynomialBatch::from_values(ctx, cs.cs_values, cs.params.rate_bits, false, cs.params.cap_height, true);
        auto warm = prove(ctx, circuit, constants_sigmas, cs.digest, cs.wire_values, cs.public_inputs, cs.params);
        if (warm != cs.expected) mismatches++;
        ready++;
        while (!go.load()) std::this_thread::yield();
        auto t0 = std::chrono::steady_clock::now();
        for (int i = 0; i < per; i++) {
          auto proof = prove(ctx, circuit, constants_sigmas, cs.digest, cs.wire_values, cs.public_inputs, cs.params);
          if (proof != cs.expected) mismatches++;
        }
        secs[t] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      });
    }
    while (ready.load() < threads) std::this_thread::yield();
    auto t0 = std::chrono::steady_clock::now();
    go = true;
    for (auto& th : pool) th.join();
    const double wall = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    printf("{\"host\": \"c++ (plonky2_b200.hpp)\", \"threads\": %d, \"proofs\": %d, \"proofs_per_s\": %.2f, "
           "\"ms_per_proof_per_thread\": %.3f, \"mismatching_proofs\": %d, \"proof_words\": %zu}\n",
           threads, threads * per, threads * per / wall, wall / per * 1e3, mismatches.load(), cs.expected.size());
    return mismatches.load() ? 1 : 0;
  } catch (const std::exception& e) {
    fprintf(stderr, "error: %s\n", e.what());
    return 1;
  }
}
