// Probe: what one random 8-byte gather from a 16 GB table costs in DRAM bytes and in time on a B200, by load flavour
// (plain ld.global, ld.global.nc, .cg, .cs, L1::no_allocate, with an L2::64B / L2::128B hint, and under the three
// settings of cudaLimitMaxL2FetchGranularity).  Run under ncu to read dram__bytes_read.sum per launch:
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_modes gather_modes.cu
//   ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,lts__t_sectors_srcunit_tex_op_read.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum ./gather_modes
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

template <int MODE>
__device__ __forceinline__ uint64_t ld(const uint64_t* p) {
  uint64_t v;
  if (MODE == 0) { asm volatile("ld.global.u64 %0, [%1];" : "=l"(v) : "l"(p)); }
  else if (MODE == 1) { asm volatile("ld.global.nc.u64 %0, [%1];" : "=l"(v) : "l"(p)); }
  else if (MODE == 2) { asm volatile("ld.global.cg.u64 %0, [%1];" : "=l"(v) : "l"(p)); }
  else if (MODE == 3) { asm volatile("ld.global.cs.u64 %0, [%1];" : "=l"(v) : "l"(p)); }
  else if (MODE == 4) { asm volatile("ld.global.nc.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p)); }
  else if (MODE == 5) { asm volatile("ld.global.nc.L2::64B.u64 %0, [%1];" : "=l"(v) : "l"(p)); }
  else if (MODE == 6) { asm volatile("ld.global.nc.L2::128B.u64 %0, [%1];" : "=l"(v) : "l"(p)); }
  else if (MODE == 7) { asm volatile("ld.global.cv.u64 %0, [%1];" : "=l"(v) : "l"(p)); }
  else { asm volatile("ld.global.nc.L1::evict_first.u64 %0, [%1];" : "=l"(v) : "l"(p)); }
  return v;
}

template <int MODE>
__global__ void gather(const uint64_t* __restrict__ tab, uint64_t mask, int per_thread, uint64_t* __restrict__ out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t acc = 0;
#pragma unroll 4
  for (int i = 0; i < per_thread; ++i) acc += ld<MODE>(tab + (mix(t * 1315423911ull + i) & mask));
  if (acc == 0x1234567) out[0] = acc;
}

// dependent chains: one outstanding load per thread (latency-bound form)
template <int MODE>
__global__ void chase(const uint64_t* __restrict__ tab, uint64_t mask, int per_thread, uint64_t* __restrict__ out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  uint64_t x = t;
  for (int i = 0; i < per_thread; ++i) x = mix(x + ld<MODE>(tab + (mix(x + i) & mask)));
  if (x == 0x1234567) out[0] = x;
}

template <int MODE>
static void run(const char* name, const uint64_t* tab, uint64_t mask, uint64_t* out) {
  const int per = 64, threads = 256, blocks = 148 * 64;
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  gather<MODE><<<blocks, threads>>>(tab, mask, per, out);
  cudaEventRecord(a);
  gather<MODE><<<blocks, threads>>>(tab, mask, per, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms = 0;
  cudaEventElapsedTime(&ms, a, b);
  const double n = (double)blocks * threads * per;
  cudaEventRecord(a);
  chase<MODE><<<148 * 4, 256>>>(tab, mask, 256, out);
  cudaEventRecord(b);
  cudaEventSynchronize(b);
  float ms2 = 0;
  cudaEventElapsedTime(&ms2, a, b);
  printf("%-28s independent: %.2f G gathers/s (%.2f ms)   chains(151552 lanes): %.2f G/s, %.0f ns per hop  %s\n", name, n / ms / 1e6, ms,
         148.0 * 4 * 256 * 256 / ms2 / 1e6, ms2 * 1e6 / 256, cudaGetErrorString(cudaGetLastError()));
}

int main() {
  const uint64_t n = 1ull << 31;  // 16 GB of u64
  uint64_t *tab, *out;
  cudaMalloc(&tab, n * 8);
  cudaMalloc(&out, 8);
  cudaMemset(tab, 1, n * 8);
  for (int gran : {0, 32, 64, 128}) {
    if (gran) {
      cudaError_t e = cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, gran);
      size_t got = 0;
      cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
      printf("--- cudaLimitMaxL2FetchGranularity = %d (%s), now %zu\n", gran, cudaGetErrorString(e), got);
    } else {
      size_t got = 0;
      cudaDeviceGetLimit(&got, cudaLimitMaxL2FetchGranularity);
      printf("--- default cudaLimitMaxL2FetchGranularity = %zu\n", got);
    }
    run<0>("ld.global", tab, n - 1, out);
    run<1>("ld.global.nc", tab, n - 1, out);
    run<2>("ld.global.cg", tab, n - 1, out);
    run<3>("ld.global.cs", tab, n - 1, out);
    run<4>("ld.global.nc.L1::no_allocate", tab, n - 1, out);
    run<5>("ld.global.nc.L2::64B", tab, n - 1, out);
    run<6>("ld.global.nc.L2::128B", tab, n - 1, out);
    run<7>("ld.global.cv", tab, n - 1, out);
    run<8>("ld.global.nc.L1::evict_first", tab, n - 1, out);
  }
  return 0;
}
