// prove_case.hpp — the case file written by tools/dump_prove_case.py (a synthetic City-shaped circuit: description,
// constants|sigmas values, witness columns, FRI parameters, and the proof words p2b_prove must return for it),
// shared by the native job loops tools/prove_bench.cpp and tools/qbench_replay.cpp.
#pragma once
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "city_rollup_b200/cpp/plonky2_b200.hpp"

using namespace plonky2_b200;

struct Case {
  p2b_circuit_desc desc{};
  std::vector<p2b_gate> gates;
  std::vector<F> k_is;
  std::vector<std::vector<F>> cs_values, wire_values;
  HashOut digest{};
  std::vector<F> public_inputs, expected;
  p2b_fri_params params{};
};

inline std::vector<uint64_t> read_words(std::ifstream& f, size_t n) {
  std::vector<uint64_t> v(n);
  f.read(reinterpret_cast<char*>(v.data()), n * 8);
  if (!f) throw std::runtime_error("case file truncated");
  return v;
}

inline Case load_case(const char* path) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error(std::string("cannot open ") + path);
  Case c;
  auto hdr = read_words(f, 16);
  if (hdr[0] != 0x70326263617365ull) throw std::runtime_error("bad magic");
  uint32_t* d = &c.desc.degree_bits;  // the nine leading u32 fields + n_gates, in declaration order
  for (int i = 0; i < 10; i++) d[i] = (uint32_t)hdr[1 + i];
  const size_t n = size_t(1) << c.desc.degree_bits, n_pis = hdr[11], n_expected = hdr[12];
  auto g = read_words(f, 7 * c.desc.n_gates);
  for (uint32_t i = 0; i < c.desc.n_gates; i++) {
    p2b_gate gt{};
    uint32_t* q = &gt.kind;
    for (int k = 0; k < 7; k++) q[k] = (uint32_t)g[7 * i + k];
    c.gates.push_back(gt);
  }
  c.k_is = read_words(f, c.desc.num_routed_wires);
  for (uint32_t i = 0; i < c.desc.num_constants + c.desc.num_routed_wires; i++) c.cs_values.push_back(read_words(f, n));
  for (uint32_t i = 0; i < c.desc.num_wires; i++) c.wire_values.push_back(read_words(f, n));
  auto dg = read_words(f, 4);
  std::copy(dg.begin(), dg.end(), c.digest.begin());
  c.public_inputs = read_words(f, n_pis);
  auto fp = read_words(f, 5 + 16);
  uint32_t* pp = &c.params.rate_bits;
  for (int i = 0; i < 5 + 16; i++) pp[i] = (uint32_t)fp[i];
  c.expected = read_words(f, n_expected);
  return c;
}

