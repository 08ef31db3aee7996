// int32_peak.cu — integer-pipe microbenchmark for the B200 roofline denominators.
// Measures sustained warp-instruction throughput of the instructions the Goldilocks/Poseidon kernels
// are made of: IMAD.WIDE.U32 (accumulate form), IMAD (32-bit), IMAD.HI, IADD3 (+carry chains), and
// 1:1 mixes of FMA-pipe and ALU-pipe instructions.  Output: one JSON object on stdout.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int32_peak int32_peak.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048
#define CHAINS 8

template <int KIND>
__global__ void __launch_bounds__(256) k(uint64_t* out, uint32_t a, uint32_t b) {
  uint64_t acc[CHAINS];
  uint32_t x[CHAINS], y[CHAINS];
  double dd[CHAINS];
  double dm = 1.0 + 1e-9 * a, da = 1e-7 * b;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) {
    acc[i] = threadIdx.x + i;
    x[i] = a + threadIdx.x * 3 + i;
    y[i] = b + i + threadIdx.x * 5;
    dd[i] = (double)(threadIdx.x + i) * 1e-3;
  }
#pragma unroll 1
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++) {
#pragma unroll
      for (int i = 0; i < CHAINS; i++) {
        if (KIND == 0) {  // IMAD.WIDE.U32 accumulate
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((uint32_t)acc[i]), "r"(y[i]));
        } else if (KIND == 1) {  // IMAD 32-bit
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(a));
        } else if (KIND == 2) {  // IMAD.HI
          asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(a));
        } else if (KIND == 3) {  // IADD3 (3-input add)
          asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(x[(i + 1) % CHAINS]));
        } else if (KIND == 4) {  // 64-bit add: IADD3 + IADD3.X
          asm volatile("add.cc.u32 %0, %0, %2; addc.u32 %1, %1, %3;" : "+r"(x[i]), "+r"(y[i]) : "r"(y[(i + 1) % CHAINS]), "r"(x[(i + 3) % CHAINS]));
        } else if (KIND == 5) {  // mix: IMAD.WIDE.U32 accumulate + IADD3
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((uint32_t)acc[i]), "r"(b));
          asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(y[i]) : "r"(x[i]), "r"(b));
        } else if (KIND == 6) {  // mix: IMAD 32 + IADD3
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a));
          asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(y[i]) : "r"(a), "r"(b));
        } else if (KIND == 7) {  // IMAD.WIDE.U32 without addend (RZ) + separate 64-bit add on ALU (ptxas' MDS form)
          uint64_t t;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"((uint32_t)acc[i]), "r"(y[i]));
          asm volatile("add.u64 %0, %0, %1;" : "+l"(acc[i]) : "l"(t));
        } else if (KIND == 8) {  // LOP3
          asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(y[i]), "r"(a));
        } else if (KIND == 9) {  // SHF (funnel shift)
          asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(y[i]), "r"(a));
        } else if (KIND == 10) {  // DFMA (FP64 pipe)
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(dm), "d"(da));
        } else if (KIND == 11) {  // DFMA + IMAD 1:1
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(dm), "d"(da));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a));
        } else if (KIND == 12) {  // DFMA + IADD3 1:1
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(dm), "d"(da));
          asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(y[i]) : "r"(a), "r"(b));
        } else if (KIND == 13) {  // DFMA + IMAD + IADD3 1:1:1
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(dm), "d"(da));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a));
          asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(y[i]) : "r"(a), "r"(b));
        } else if (KIND == 14) {  // DADD
          asm volatile("add.rn.f64 %0, %0, %1;" : "+d"(dd[i]) : "d"(da));
        } else if (KIND == 15) {  // DFMA + IMAD.WIDE 1:1
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(dm), "d"(da));
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((uint32_t)acc[i]), "r"(b));
        } else if (KIND == 16) {  // 2 DFMA + IMAD + IADD3
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(dm), "d"(da));
          asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a));
          asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(dd[i]) : "d"(dm), "d"(da));
          asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(y[i]) : "r"(a), "r"(b));
        } else if (KIND == 18) {  // IMAD.HI + IADD3 1:1
          asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(b), "r"(a));
          asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(y[i]) : "r"(a), "r"(b));
        } else if (KIND == 19) {  // IMAD.WIDE accumulate + 2 IADD3
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((uint32_t)acc[i]), "r"(b));
          asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(y[i]) : "r"(a), "r"(b));
          asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
        } else if (KIND == 20) {  // IMAD.WIDE accumulate + 4 IADD3
          asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[i]) : "r"((uint32_t)acc[i]), "r"(b));
          asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(y[i]) : "r"(a), "r"(b));
          asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
          asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(y[i]) : "r"(b), "r"(a));
          asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x[i]) : "r"(b), "r"(a));
        } else if (KIND == 21) {  // IMAD.WIDE without addend (low word fed back) + 2 IADD3
          uint64_t t;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"((uint32_t)acc[i]), "r"((uint32_t)(acc[i] >> 32)));
          acc[i] = t;
          asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(y[i]) : "r"(a), "r"(b));
          asm volatile("add.u32 %0, %0, %1; add.u32 %0, %0, %2;" : "+r"(x[i]) : "r"(a), "r"(b));
        } else if (KIND == 22) {  // IMAD.WIDE without addend alone
          uint64_t t;
          asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(t) : "r"((uint32_t)acc[i]), "r"((uint32_t)(acc[i] >> 32)));
          acc[i] = t;
        } else if (KIND == 17) {  // I2F.F64.U32 conversion
          asm volatile("cvt.rn.f64.u32 %0, %1;" : "=d"(dd[i]) : "r"(x[i]));
          asm volatile("mov.b64 {%0, %1}, %2;" : "=r"(x[i]), "=r"(y[i]) : "d"(dd[i]));
        }
      }
    }
  }
  uint64_t r = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; i++) r += acc[i] + x[i] + y[i] + (uint64_t)__double_as_longlong(dd[i]);
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = r;  // keep results live
}

template <int KIND>
double run(const char* name, int instr_per_slot, uint64_t* d, int sms) {
  int blocks = sms * 8;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  k<KIND><<<blocks, 256>>>(d, 12345, 678);
  cudaDeviceSynchronize();
  float best = 1e30f;
  for (int rep = 0; rep < 5; rep++) {
    cudaEventRecord(e0);
    k<KIND><<<blocks, 256>>>(d, 12345, 678);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  double thread_instr = (double)blocks * 256 * ITERS * 4 * CHAINS * instr_per_slot;
  double gops = thread_instr / (best * 1e-3) / 1e9;
  printf("  \"%s\": {\"ms\": %.4f, \"thread_instr_per_s_G\": %.1f, \"per_sm_per_clk_at_1965MHz\": %.2f},\n", name, best,
         gops, gops * 1e9 / sms / 1.965e9);
  return gops;
}

int main() {
  cudaDeviceProp p;
  cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount;
  uint64_t* d;
  cudaMalloc(&d, (size_t)sms * 8 * 256 * 8);
  printf("{\n  \"device\": \"%s\", \"sms\": %d,\n", p.name, sms);
  run<0>("imad_wide_u32_acc", 1, d, sms);
  run<1>("imad_lo_u32", 1, d, sms);
  run<2>("imad_hi_u32", 1, d, sms);
  run<3>("iadd3", 1, d, sms);
  run<4>("iadd3_plus_iadd3x", 2, d, sms);
  run<5>("mix_imadwide_iadd3", 2, d, sms);
  run<6>("mix_imad_iadd3", 2, d, sms);
  run<7>("imadwide_rz_plus_add64", 3, d, sms);
  run<8>("lop3", 1, d, sms);
  run<9>("shf", 1, d, sms);
  run<10>("dfma", 1, d, sms);
  run<11>("mix_dfma_imad", 2, d, sms);
  run<12>("mix_dfma_iadd3", 2, d, sms);
  run<13>("mix_dfma_imad_iadd3", 3, d, sms);
  run<14>("dadd", 1, d, sms);
  run<15>("mix_dfma_imadwide", 2, d, sms);
  run<16>("mix_2dfma_imad_iadd3", 4, d, sms);
  run<17>("cvt_f64_u32_plus_mov", 1, d, sms);
  // do the wide multiplies run beside ALU work? (profiles/r01_poseidon_v6_experiments.md)
  run<18>("mix_imadhi_iadd3", 2, d, sms);
  run<19>("mix_imadwide_2iadd3", 3, d, sms);
  run<20>("mix_imadwide_4iadd3", 5, d, sms);
  run<21>("mix_imadwide_rz_2iadd3", 3, d, sms);
  run<22>("imadwide_rz", 1, d, sms);
  printf("  \"note\": \"thread-level instructions per second; per_sm_per_clk normalised to 1965 MHz max clock\"\n}\n");
  return 0;
}
