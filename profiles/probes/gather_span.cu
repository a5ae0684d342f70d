// Probe: (1) random 8-byte gathers as a function of the SPAN they are spread over (1 GB .. 96 GB): does address
// translation, not DRAM, bound the prefix-table lookups of a 68.7 GB table?  (2) the same gathers when the stream of
// addresses is ordered into windows of 256 MB (what sorting the lookups by key would give).  (3) same-address atomicAdd
// rate with one atomic per warp (the list appends of the wave path).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o gather_span gather_span.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}

// span_words: gathers fall uniformly into [0, span_words); window_words != 0: gather i of the whole launch falls into
// window (global index * n_windows / total), i.e. the launch sweeps the span once, window by window
__global__ void gather(const uint64_t* __restrict__ tab, uint64_t span_words, uint64_t window_words, int per_thread,
                       uint64_t* __restrict__ out) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t total = (uint64_t)gridDim.x * blockDim.x;
  uint64_t acc = 0;
#pragma unroll 4
  for (int i = 0; i < per_thread; ++i) {
    const uint64_t h = mix(t * 1315423911ull + i);
    uint64_t a;
    if (window_words) {
      const uint64_t n_win = span_words / window_words;
      // blocks run roughly in launch order: block b works in window b * n_win / gridDim.x
      const uint64_t win = (uint64_t)blockIdx.x * n_win / gridDim.x;
      a = win * window_words + h % window_words;
    } else {
      a = h % span_words;
    }
    acc += __ldg(tab + a);
  }
  if (acc == 0x1234567) out[0] = acc;
  (void)total;
}

__global__ void atomics_one_per_warp(unsigned int* counter, int per_thread) {
  for (int i = 0; i < per_thread; ++i)
    if ((threadIdx.x & 31) == 0) atomicAdd(counter, 32u);
}
__global__ void atomics_one_per_block(unsigned int* counter, int per_thread) {
  for (int i = 0; i < per_thread; ++i) {
    if (threadIdx.x == 0) atomicAdd(counter, 256u);
    __syncthreads();
  }
}

int main() {
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  const uint64_t bytes = 96ull << 30;
  uint64_t *tab, *out;
  if (cudaMalloc(&tab, bytes) != cudaSuccess) { printf("no 96 GB\n"); return 1; }
  cudaMalloc(&out, 8);
  cudaMemset(tab, 1, bytes);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  const int per = 32, threads = 256, blocks = 148 * 64;
  const double n = (double)blocks * threads * per;
  for (uint64_t gb : {1ull, 4ull, 16ull, 32ull, 64ull, 96ull}) {
    const uint64_t span = (gb << 30) / 8;
    for (uint64_t win_mb : {0ull, 256ull, 1024ull}) {
      const uint64_t win = (win_mb << 20) / 8;
      if (win >= span) continue;
      gather<<<blocks, threads>>>(tab, span, win, per, out);
      cudaEventRecord(a);
      gather<<<blocks, threads>>>(tab, span, win, per, out);
      cudaEventRecord(b);
      cudaEventSynchronize(b);
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      printf("span %3llu GB, window %4llu MB: %.2f G gathers/s (%.2f ms) %s\n", (unsigned long long)gb, (unsigned long long)win_mb,
             n / ms / 1e6, ms, cudaGetErrorString(cudaGetLastError()));
    }
  }
  unsigned int* ctr;
  cudaMalloc(&ctr, 4);
  cudaMemset(ctr, 0, 4);
  for (int which = 0; which < 2; ++which) {
    cudaEventRecord(a);
    if (which == 0) atomics_one_per_warp<<<148 * 4, 256>>>(ctr, 256);
    else atomics_one_per_block<<<148 * 4, 256>>>(ctr, 256);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    const double cnt = which == 0 ? 148.0 * 4 * 8 * 256 : 148.0 * 4 * 256;
    printf("same-address atomicAdd, one per %s: %.1f M atomics/s (%.3f ms for %.0f)\n", which == 0 ? "warp" : "block", cnt / ms / 1e3, ms, cnt);
  }
  return 0;
}
