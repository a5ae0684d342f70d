// Ordering of the per-read records for the writers: CAligner::SortReadHits(eRSMHitMatch) with the comparator
// SortHitMatch (biokanga/Aligner.cpp:9917-9991, 10067-10114), which the reference runs as a multi-threaded
// quicksort over 20-100 M record pointers.  Here: two stable LSD radix passes on the device over (key, index)
// pairs -- the low-order fields first, then the high-order ones.
//   order:  NAR class, uniquely-hit records first, then (for those) chromosome id, locus, match length, strand,
//           mismatches;  records without a unique hit order by NumHits.
// The reference's quicksort is unstable and leaves equal keys in unspecified order; ties here are by record index
// (= read load order), so the output is deterministic.
#include <cstdint>
#include <cub/device/device_radix_sort.cuh>

#include "../../include/bkx.h"

int bkx_fail(int code, const char* fmt, ...);

namespace {

__global__ void sort_low_keys(const bkx_read_result* __restrict__ r, uint32_t n, uint64_t* __restrict__ keys,
                              uint32_t* __restrict__ vals) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const bkx_read_result a = r[i];
    uint64_t k = 0;
    if (a.num_hits == 1)
      k = ((uint64_t)a.match_loci << 32) | ((uint64_t)a.match_len << 16) | ((uint64_t)a.strand << 8) | (uint8_t)a.low_mm;
    keys[i] = k;
    vals[i] = i;
  }
}

__global__ void sort_high_keys(const bkx_read_result* __restrict__ r, const uint32_t* __restrict__ vals, uint32_t n,
                               uint64_t* __restrict__ keys) {
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    const bkx_read_result a = r[vals[j]];
    const bool uniq = a.num_hits == 1;
    keys[j] = ((uint64_t)a.nar << 41) | ((uint64_t)(uniq ? 0 : 1) << 40) | ((uint64_t)(uniq ? 0 : a.num_hits) << 32) |
              (uint64_t)(uniq ? a.chrom_id : 0u);
  }
}

}  // namespace

extern "C" int bkx_sort_hits(const bkx_read_result* results, uint32_t n, uint32_t* order_out, int device) {
  if (!results || !order_out) return bkx_fail(BKX_ERR_PARAM, "null argument");
  if (n == 0) return BKX_OK;
  int rc = BKX_OK;
  cudaStream_t st = nullptr;
  bkx_read_result* d_res = nullptr;
  uint64_t *k0 = nullptr, *k1 = nullptr;
  uint32_t *v0 = nullptr, *v1 = nullptr;
  void* tmp = nullptr;
  size_t need = 0, need2 = 0;
#define SO_CU(call)                                                                                              \
  do {                                                                                                           \
    cudaError_t e__ = (call);                                                                                    \
    if (e__ != cudaSuccess) { rc = bkx_fail(BKX_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); goto done; } \
  } while (0)
  SO_CU(cudaSetDevice(device));
  SO_CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  SO_CU(cudaMalloc((void**)&d_res, (size_t)n * sizeof(bkx_read_result)));
  SO_CU(cudaMalloc((void**)&k0, (size_t)n * 8));
  SO_CU(cudaMalloc((void**)&k1, (size_t)n * 8));
  SO_CU(cudaMalloc((void**)&v0, (size_t)n * 4));
  SO_CU(cudaMalloc((void**)&v1, (size_t)n * 4));
  SO_CU(cudaMemcpyAsync(d_res, results, (size_t)n * sizeof(bkx_read_result), cudaMemcpyHostToDevice, st));
  {
    const int grid = (int)std::min<uint64_t>(((uint64_t)n + 255) / 256, 148 * 16);
    cub::DoubleBuffer<uint64_t> dk(k0, k1);
    cub::DoubleBuffer<uint32_t> dv(v0, v1);
    SO_CU(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (int)n, 0, 64, st));
    SO_CU(cub::DeviceRadixSort::SortPairs(nullptr, need2, dk, dv, (int)n, 0, 49, st));
    need = std::max(need, need2);
    SO_CU(cudaMalloc(&tmp, need));
    sort_low_keys<<<grid, 256, 0, st>>>(d_res, n, dk.Current(), dv.Current());
    SO_CU(cudaGetLastError());
    SO_CU(cub::DeviceRadixSort::SortPairs(tmp, need, dk, dv, (int)n, 0, 64, st));
    sort_high_keys<<<grid, 256, 0, st>>>(d_res, dv.Current(), n, dk.Current());
    SO_CU(cudaGetLastError());
    SO_CU(cub::DeviceRadixSort::SortPairs(tmp, need, dk, dv, (int)n, 0, 49, st));
    SO_CU(cudaMemcpyAsync(order_out, dv.Current(), (size_t)n * 4, cudaMemcpyDeviceToHost, st));
    SO_CU(cudaStreamSynchronize(st));
  }
done:
  if (st) cudaStreamDestroy(st);
  cudaFree(d_res); cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(tmp);
  return rc;
#undef SO_CU
}
