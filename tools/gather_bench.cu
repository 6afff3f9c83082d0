// Random-sector gather micro-benchmark: the secondary roofline denominator for the hash-probe kernel
// (SURVEY.md section 8d).  Every thread reads one random, 32-byte-aligned 32 B sector per iteration
// (two ld.global.cg.v2.u64, exactly the access k_insert makes) from a table of the given size.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/gather_bench.cu -o build/gather_bench
//   build/gather_bench > gpurun_out/gather.json
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
__device__ __forceinline__ uint64_t fmix64(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33; return x;
}
__global__ void k_gather(const uint64_t* table, uint64_t bucket_mask, uint64_t n, unsigned long long* sink) {
  uint64_t acc = 0;
  uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint64_t b = fmix64(i + 1) & bucket_mask;
    const uint64_t* p = table + (b << 2);
    uint64_t a0, a1, a2, a3;
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(a0), "=l"(a1) : "l"(p));
    asm volatile("ld.global.cg.v2.u64 {%0, %1}, [%2];" : "=l"(a2), "=l"(a3) : "l"(p + 2));
    acc += a0 ^ a1 ^ a2 ^ a3;
  }
  if (acc == 0x1234567ull) atomicAdd(sink, 1ull);
}
int main() {
  int sms = 148;
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  sms = prop.multiProcessorCount;
  unsigned long long* sink;
  cudaMalloc(&sink, 8);
  printf("{\"device\": \"%s\", \"results\": [", prop.name);
  bool first = true;
  for (int lg = 25; lg <= 32; ++lg) {           // 2^lg slots of 8 B: 256 MB .. 32 GB
    uint64_t slots = 1ull << lg;
    uint64_t* table;
    if (cudaMalloc(&table, slots * 8) != cudaSuccess) break;
    cudaMemset(table, 0, slots * 8);
    uint64_t n = 1ull << 28;                    // 268 M probes
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
      cudaEventRecord(a);
      k_gather<<<sms * 8, 256>>>(table, (slots >> 2) - 1, n, sink);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms; cudaEventElapsedTime(&ms, a, b);
      if (rep > 0 && ms < best) best = ms;
    }
    printf("%s{\"table_bytes\": %llu, \"probes\": %llu, \"ms\": %.3f, \"probes_per_s\": %.4g, \"gbs_32B\": %.1f}",
           first ? "" : ", ", (unsigned long long)(slots * 8), (unsigned long long)n, best, n / (best * 1e-3), n * 32.0 / (best * 1e-3) / 1e9);
    first = false;
    cudaFree(table);
  }
  printf("]}\n");
  return 0;
}
