// pow_debug.cu — times frik::k_pow_search alone for several `count` values (development tool)
#include <cstdio>
#include <cuda_runtime.h>
#include "poseidon.cuh"
#include "fri_kernels.cuh"
int main() {
  uint64_t h[30] = {0};
  for (int i = 0; i < 12; i++) h[i] = 1000 + i;
  h[frik::CH_NIN] = 2; h[frik::CH_IN] = 5; h[frik::CH_IN + 1] = 6;
  uint64_t *d_st; unsigned long long* d_best;
  cudaMalloc(&d_st, sizeof(h)); cudaMalloc(&d_best, 8);
  cudaMemcpy(d_st, h, sizeof(h), cudaMemcpyHostToDevice);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int bits : {0, 8, 16}) for (int lc : {16, 18, 20, 22}) for (int blocks : {148, 444}) {
    float best_ms = 1e9; unsigned long long w = 0;
    for (int rep = 0; rep < 3; rep++) {
      cudaMemset(d_best, 0xFF, 8);
      cudaEventRecord(e0);
      frik::k_pow_search<<<blocks, 256>>>(d_st, 0, 1ull << lc, bits, d_best);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best_ms) best_ms = ms;
      cudaMemcpy(&w, d_best, 8, cudaMemcpyDeviceToHost);
    }
    printf("bits %2d count 2^%d blocks %d: %.3f ms, witness %llu (%s)\n", bits, lc, blocks, best_ms, w, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
