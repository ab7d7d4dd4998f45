// int_peaks.cu -- integer-pipe throughput micro-benchmark for the Hamming roofline ("microbenchmark it").
// Measures warp-wide results / clk / SM of POPC, LOP3, IADD3 and VIMNMX3 on the device it runs on (the pipes the
// Hamming inner loop uses: 8 XOR + 8 POPC + adds per pair, or carry-save: 8 XOR + 14 LOP3 + 4 POPC).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_peaks tools/int_peaks.cu && ./int_peaks > profiles/int_peaks.json
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>

constexpr int THREADS = 512;
constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ unsigned popc_asm(unsigned a) {
  unsigned r;
  asm volatile("popc.b32 %0, %1;" : "=r"(r) : "r"(a));
  return r;
}
__device__ __forceinline__ unsigned xor3_asm(unsigned a, unsigned b, unsigned c) {
  unsigned r;
  asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}
__device__ __forceinline__ unsigned maj_asm(unsigned a, unsigned b, unsigned c) {
  unsigned r;
  asm volatile("lop3.b32 %0, %1, %2, %3, 0xe8;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
  return r;
}

// KIND 0: POPC   1: LOP3   2: IADD3 (a = a + b + c)   3: VIMNMX3
template <int KIND>
__global__ void __launch_bounds__(THREADS) peak_kernel(unsigned* out, long long* cycles, unsigned seed) {
  unsigned a[CHAINS];
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) a[i] = seed * (threadIdx.x + 1) + i * 0x9e3779b9u;
  unsigned b = seed ^ 0x55aa55aau, c = seed + threadIdx.x;
  __syncthreads();
  const long long t0 = clock64();
  {
#pragma unroll 1
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
          if (KIND == 0) a[i] = popc_asm(a[i]);
          if (KIND == 1) a[i] = xor3_asm(a[i], b, c);
          if (KIND == 2) asm volatile("{ .reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2; }" : "+r"(a[i]) : "r"(b), "r"(c));
          if (KIND == 3) asm volatile("min.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b + i));
        }
      }
    }
  }
  const long long t1 = clock64();
  unsigned r = 0;
#pragma unroll
  for (int i = 0; i < CHAINS; ++i) r ^= a[i];
  out[blockIdx.x * THREADS + threadIdx.x] = r;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int KIND>
static void run(const char* name, double ops_per_iter, int sms, int blocks_per_sm, double sm_mhz, bool last) {
  const int grid = sms * blocks_per_sm;
  unsigned* out;
  long long* cyc;
  CK(cudaMalloc(&out, sizeof(unsigned) * grid * THREADS));
  CK(cudaMalloc(&cyc, sizeof(long long) * grid));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int w = 0; w < 3; ++w) peak_kernel<KIND><<<grid, THREADS>>>(out, cyc, 12345u + w);
  CK(cudaDeviceSynchronize());
  float best_ms = 1e30f;
  std::vector<long long> h(grid);
  double med_cycles = 0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0));
    peak_kernel<KIND><<<grid, THREADS>>>(out, cyc, 777u + rep);
    CK(cudaEventRecord(e1));
    CK(cudaEventSynchronize(e1));
    float ms;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (ms < best_ms) {
      best_ms = ms;
      CK(cudaMemcpy(h.data(), cyc, sizeof(long long) * grid, cudaMemcpyDeviceToHost));
      std::sort(h.begin(), h.end());
      med_cycles = (double) h[grid / 2];
    }
  }
  const double ops_per_sm = ops_per_iter * ITERS * 4.0 * THREADS * blocks_per_sm;
  const double per_clk_clock64 = ops_per_sm / med_cycles;
  const double per_clk_event = ops_per_sm / (best_ms * 1e-3 * sm_mhz * 1e6);
  printf("  \"%s\": {\"results_per_clk_per_sm\": %.2f, \"by_event_time_at_%.0f_mhz\": %.2f, \"ms\": %.4f, \"blocks_per_sm\": %d}%s\n",
         name, per_clk_clock64, sm_mhz, per_clk_event, best_ms, blocks_per_sm, last ? "" : ",");
  CK(cudaFree(out));
  CK(cudaFree(cyc));
}

int main() {
  cudaDeviceProp p;
  CK(cudaGetDeviceProperties(&p, 0));
  int khz = 0;
  CK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0));
  const double mhz = khz / 1e3;
  const int sms = p.multiProcessorCount;
  printf("{\n  \"device\": \"%s\", \"sms\": %d, \"sm_max_mhz\": %.0f,\n", p.name, sms, mhz);
  printf("  \"how\": \"tools/int_peaks.cu: %d threads x 4 CTAs per SM, %d independent dependent-chains per thread, clock64 around the loop (median CTA)\",\n", THREADS, CHAINS);
  run<0>("popc", CHAINS, sms, 4, mhz, false);
  run<1>("lop3", CHAINS, sms, 4, mhz, false);
  run<2>("iadd3", CHAINS, sms, 4, mhz, false);          // ptxas fuses the two PTX adds into one IADD3 (checked in SASS)
  run<3>("vimnmx3", CHAINS / 2.0, sms, 4, mhz, true);    // two chained PTX min -> one VIMNMX3 (checked in SASS)
  printf("}\n");
  return 0;
}
