// GPU suffix-array construction for genomes of any size the index format allows (up to 2^40 symbols),
// in bounded memory: the builder behind the 14 Gbp / 5-byte-element configuration.
//
// Same order as bkx_sabuild.cu and as `biokanga index` (4-bit symbols A<C<G<T<N<EOS compared straight
// through the terminators, CSfxArrayV3::QSortSeq / SfxOfsCompare, libbiokanga/SfxArrayV2.cpp:9451-9542);
// different method, because prefix doubling needs a rank per suffix (5 more bytes x n) that does not fit
// beside a 70 GB suffix array:
//   1. histogram of the leading 7 symbols (2M bins); contiguous bin ranges of at most `cap` suffixes are
//      the BATCHES -- each is a contiguous range of the final array;
//   2. per batch: gather the member positions with their first 21-symbol window (63-bit key), radix sort;
//   3. suffixes still tied are re-sorted inside their group on the NEXT 21-symbol window, read straight
//      from the sequence (two stable radix passes: window, then group id), until every group is a
//      singleton.  Unlike doubling this is linear in the longest repeat (a 5 kb exact repeat takes 240
//      rounds), but the tied set shrinks geometrically for diverged copies (homeologs at 2-5 % keep a
//      factor ~0.5 per round) and the exact repeats left are few.
// Output: u32 low plane + u8 high plane, the layout DevIndex reads (bkx_index.cuh).
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "../../include/bkx.h"

int bkx_fail(int code, const char* fmt, ...);

namespace {

constexpr int kWin = 21;        // symbols per 63-bit window
constexpr int kBinSyms = 7;     // symbols per histogram bin
constexpr uint32_t kBins = 1u << (3 * kBinSyms);
constexpr int kTile = 2048;     // positions per block iteration of the scan kernels
constexpr int kThreads = 256;

#define SL_CU(call)                                                                                   \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess) {                                                                         \
      rc = bkx_fail(BKX_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      goto done;                                                                                      \
    }                                                                                                 \
  } while (0)

__device__ __forceinline__ unsigned sym3(const uint8_t* __restrict__ seq, uint64_t n, uint64_t p) {
  if (p >= n) return 0;  // past the end sorts lowest
  unsigned b = __ldg(seq + p) & 0x0f;
  return b < 5 ? b + 1 : 7;
}

__device__ __forceinline__ uint64_t window_key(const uint8_t* __restrict__ seq, uint64_t n, uint64_t p) {
  uint64_t k = 0;
#pragma unroll
  for (int j = 0; j < kWin; ++j) k = (k << 3) | sym3(seq, n, p + j);
  return k;
}

// stage the 3-bit symbols of [t0, t0 + kTile + kWin) in shared memory
__device__ __forceinline__ void load_tile(const uint8_t* __restrict__ seq, uint64_t n, uint64_t t0, uint8_t* sm) {
  for (int i = threadIdx.x; i < kTile + kWin; i += kThreads) sm[i] = (uint8_t)sym3(seq, n, t0 + i);
  __syncthreads();
}

__global__ void __launch_bounds__(kThreads) sl_histogram(const uint8_t* __restrict__ seq, uint64_t n,
                                                         unsigned long long* __restrict__ hist) {
  __shared__ uint8_t sm[kTile + kWin];
  for (uint64_t t0 = (uint64_t)blockIdx.x * kTile; t0 < n; t0 += (uint64_t)gridDim.x * kTile) {
    load_tile(seq, n, t0, sm);
    for (int i = threadIdx.x; i < kTile && t0 + i < n; i += kThreads) {
      uint32_t b = 0;
#pragma unroll
      for (int j = 0; j < kBinSyms; ++j) b = (b << 3) | sm[i + j];
      atomicAdd(hist + b, 1ull);
    }
    __syncthreads();
  }
}

// append (first window, position) of every suffix whose bin lies in [blo, bhi)
__global__ void __launch_bounds__(kThreads) sl_gather(const uint8_t* __restrict__ seq, uint64_t n, uint32_t blo,
                                                      uint32_t bhi, uint64_t* __restrict__ keys,
                                                      uint64_t* __restrict__ vals, unsigned long long* __restrict__ count) {
  __shared__ uint8_t sm[kTile + kWin];
  const int lane = threadIdx.x & 31;
  for (uint64_t t0 = (uint64_t)blockIdx.x * kTile; t0 < n; t0 += (uint64_t)gridDim.x * kTile) {
    load_tile(seq, n, t0, sm);
    for (int i0 = 0; i0 < kTile; i0 += kThreads) {
      const int i = i0 + threadIdx.x;
      bool take = false;
      if (t0 + i < n) {
        uint32_t b = 0;
#pragma unroll
        for (int j = 0; j < kBinSyms; ++j) b = (b << 3) | sm[i + j];
        take = b >= blo && b < bhi;
      }
      const unsigned m = __ballot_sync(0xffffffffu, take);
      if (m) {
        unsigned long long base = 0;
        if (lane == __ffs(m) - 1) base = atomicAdd(count, (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, __ffs(m) - 1);
        if (take) {
          uint64_t k = 0;
#pragma unroll
          for (int j = 0; j < kWin; ++j) k = (k << 3) | sm[i + j];
          const unsigned long long at = base + __popc(m & ((1u << lane) - 1));
          keys[at] = k;
          vals[at] = t0 + i;
        }
      }
    }
    __syncthreads();
  }
}

// lo == nullptr: `hi` is an array of 5-byte elements back to back (the layout of the .sfx file and of DevIndex::sa5)
__device__ __forceinline__ void put_sa(uint32_t* __restrict__ lo, uint8_t* __restrict__ hi, uint64_t at, uint64_t pos) {
  if (!lo) {
    uint8_t* p = hi + at * 5;
    p[0] = (uint8_t)pos; p[1] = (uint8_t)(pos >> 8); p[2] = (uint8_t)(pos >> 16); p[3] = (uint8_t)(pos >> 24);
    p[4] = (uint8_t)(pos >> 32);
    return;
  }
  lo[at] = (uint32_t)pos;
  if (hi) hi[at] = (uint8_t)(pos >> 32);
}

// after the first sort of a batch: group heads, "still tied" flags, and the provisional array slice
__global__ void sl_first_pass(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ vals, uint64_t c,
                              uint64_t base, uint32_t* __restrict__ lo, uint8_t* __restrict__ hi,
                              uint8_t* __restrict__ head, uint8_t* __restrict__ tied) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < c; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    const bool h = (i == 0) || keys[i - 1] != k;
    const bool nh = (i + 1 == c) || keys[i + 1] != k;
    head[i] = h;
    tied[i] = !(h && nh);
    put_sa(lo, hi, base + i, vals[i]);
  }
}

struct U8ToU32 {
  __host__ __device__ __forceinline__ uint32_t operator()(uint8_t v) const { return v; }
};
struct MaxU32 {
  __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

// keep the tied elements: (position, slot in the batch, head flag), order preserved
__global__ void sl_compact(const uint64_t* __restrict__ pos, const uint32_t* __restrict__ slot,
                           const uint8_t* __restrict__ head, const uint8_t* __restrict__ tied,
                           const uint32_t* __restrict__ idx, uint64_t m, uint64_t* __restrict__ pos_o,
                           uint32_t* __restrict__ slot_o, uint8_t* __restrict__ head_o) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
    if (!tied[i]) continue;
    const uint32_t o = idx[i];
    pos_o[o] = pos[i];
    slot_o[o] = slot ? slot[i] : (uint32_t)i;
    head_o[o] = head[i];
  }
}

// round r: window r of every tied suffix, its group id seed (own index at heads, 0 elsewhere), identity permutation
__global__ void sl_round_keys(const uint8_t* __restrict__ seq, uint64_t n, const uint64_t* __restrict__ pos,
                              const uint8_t* __restrict__ head, uint64_t m, uint64_t shift,
                              uint64_t* __restrict__ keys, uint32_t* __restrict__ perm, uint32_t* __restrict__ grp) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
    keys[i] = window_key(seq, n, pos[i] + shift);
    perm[i] = (uint32_t)i;
    grp[i] = head[i] ? (uint32_t)i : 0u;
  }
}

__global__ void sl_gather_grp(const uint32_t* __restrict__ grp, const uint32_t* __restrict__ perm, uint64_t m,
                              uint32_t* __restrict__ out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = grp[perm[i]];
}

// apply the (group, window) order: new positions and their windows
__global__ void sl_apply(const uint8_t* __restrict__ seq, uint64_t n, const uint64_t* __restrict__ pos,
                         const uint32_t* __restrict__ perm, uint64_t m, uint64_t shift, uint64_t* __restrict__ pos_o,
                         uint64_t* __restrict__ keys_o) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t p = pos[perm[i]];
    pos_o[i] = p;
    keys_o[i] = window_key(seq, n, p + shift);
  }
}

// split groups where the new window differs; write the slice; flag what is still tied
__global__ void sl_regroup(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ pos,
                           const uint32_t* __restrict__ slot, const uint8_t* __restrict__ head_in, uint64_t m,
                           uint64_t base, uint32_t* __restrict__ lo, uint8_t* __restrict__ hi,
                           uint8_t* __restrict__ head_out, uint8_t* __restrict__ tied) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t k = keys[i];
    const bool h = head_in[i] || keys[i - (i ? 1 : 0)] != k;
    const bool nh = (i + 1 == m) || head_in[i + 1] || keys[i + 1] != k;
    head_out[i] = h;
    tied[i] = !(h && nh);
    put_sa(lo, hi, base + slot[i], pos[i]);
  }
}

int grid_for(uint64_t n) {
  uint64_t g = (n + 255) / 256;
  return (int)std::min<uint64_t>(std::max<uint64_t>(g, 1), 148 * 32);
}

size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }

struct Carver {  // bump allocation inside the one arena
  uint8_t* base;
  size_t off, cap;
  template <typename T>
  T* take(size_t count) {
    size_t bytes = align256(count * sizeof(T));
    if (off + bytes > cap) return nullptr;
    T* p = (T*)(base + off);
    off += bytes;
    return p;
  }
};

}  // namespace

static int build_large(const uint8_t* d_seq, uint64_t n, uint32_t* d_sa_lo, uint8_t* d_sa_hi, int device, uint64_t max_batch);

extern "C" int bkx_build_suffix_array_planes(const uint8_t* d_seq, uint64_t n, uint32_t* d_sa_lo, uint8_t* d_sa_hi,
                                             int device, uint64_t max_batch) {
  if (!d_seq || !d_sa_lo) return bkx_fail(BKX_ERR_PARAM, "null argument");
  if (n > 0xffffffffull && !d_sa_hi) return bkx_fail(BKX_ERR_PARAM, "more than 2^32 symbols need the high plane");
  return build_large(d_seq, n, d_sa_lo, d_sa_hi, device, max_batch);
}

extern "C" int bkx_build_suffix_array_packed5(const uint8_t* d_seq, uint64_t n, uint8_t* d_sa5, int device, uint64_t max_batch) {
  if (!d_seq || !d_sa5) return bkx_fail(BKX_ERR_PARAM, "null argument");
  return build_large(d_seq, n, nullptr, d_sa5, device, max_batch);
}

static int build_large(const uint8_t* d_seq, uint64_t n, uint32_t* d_sa_lo, uint8_t* d_sa_hi, int device, uint64_t max_batch) {
  if (n < 2) return bkx_fail(BKX_ERR_PARAM, "sequence too short");
  if (n >= (1ull << 40)) return bkx_fail(BKX_ERR_PARAM, "sequence longer than 2^40 symbols");
  int rc = BKX_OK;
  cudaStream_t st = nullptr;
  unsigned long long* d_hist = nullptr;
  unsigned long long* d_count = nullptr;
  uint8_t* arena = nullptr;
  void* cub_tmp = nullptr;
  size_t cub_bytes = 0;
  std::vector<unsigned long long> hist(kBins);
  uint64_t cap = 0, max_bin = 0, done_sfx = 0;
  size_t free_b = 0, total_b = 0, arena_bytes = 0;
  const size_t kPerElem = 60;  // arena bytes per batch element (see the layout below: 13c + max(38c, 46m), m <= c)
  int n_batches = 0;
  uint64_t rounds_total = 0;

  SL_CU(cudaSetDevice(device));
  SL_CU(cudaDeviceSynchronize());  // whatever stream produced d_seq: this call works on its own non-blocking stream
  SL_CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  SL_CU(cudaMalloc((void**)&d_hist, (size_t)kBins * 8));
  SL_CU(cudaMalloc((void**)&d_count, 8));
  SL_CU(cudaMemsetAsync(d_hist, 0, (size_t)kBins * 8, st));
  sl_histogram<<<148 * 8, kThreads, 0, st>>>(d_seq, n, d_hist);
  SL_CU(cudaGetLastError());
  SL_CU(cudaMemcpyAsync(hist.data(), d_hist, (size_t)kBins * 8, cudaMemcpyDeviceToHost, st));
  SL_CU(cudaStreamSynchronize(st));
  for (uint32_t b = 0; b < kBins; ++b) max_bin = std::max<uint64_t>(max_bin, hist[b]);

  // batch capacity from the memory that is free right now
  SL_CU(cudaMemGetInfo(&free_b, &total_b));
  cap = (uint64_t)((double)free_b * 0.90 / (double)kPerElem);
  cap = std::min<uint64_t>(cap, 0x7fffff00ull);  // slots are u32, CUB item counts stay below 2^31
  cap = std::min<uint64_t>(cap, n);
  if (max_batch) cap = std::min<uint64_t>(cap, std::max<uint64_t>(max_batch, max_bin));
  if (cap < max_bin) {
    rc = bkx_fail(BKX_ERR_UNSUPPORTED, "suffix array: %llu suffixes share their first %d symbols, more than one batch of %llu holds",
                  (unsigned long long)max_bin, kBinSyms, (unsigned long long)cap);
    goto done;
  }
  arena_bytes = (size_t)cap * kPerElem + (64u << 10);
  SL_CU(cudaMalloc((void**)&arena, arena_bytes));
  {  // CUB scratch sized once for the largest calls
    size_t need = 0, a = 0;
    cub::DoubleBuffer<uint64_t> dk((uint64_t*)nullptr, (uint64_t*)nullptr);
    cub::DoubleBuffer<uint64_t> dv((uint64_t*)nullptr, (uint64_t*)nullptr);
    cub::DoubleBuffer<uint32_t> dp((uint32_t*)nullptr, (uint32_t*)nullptr);
    cub::DoubleBuffer<uint32_t> dg((uint32_t*)nullptr, (uint32_t*)nullptr);
    SL_CU(cub::DeviceRadixSort::SortPairs(nullptr, a, dk, dv, (long long)cap, 0, 63, st));
    need = std::max(need, a);
    SL_CU(cub::DeviceRadixSort::SortPairs(nullptr, a, dk, dp, (long long)cap, 0, 63, st));
    need = std::max(need, a);
    SL_CU(cub::DeviceRadixSort::SortPairs(nullptr, a, dg, dp, (long long)cap, 0, 32, st));
    need = std::max(need, a);
    cub::TransformInputIterator<uint32_t, U8ToU32, const uint8_t*> it((const uint8_t*)nullptr, U8ToU32());
    SL_CU(cub::DeviceScan::ExclusiveSum(nullptr, a, it, (uint32_t*)nullptr, (long long)cap, st));
    need = std::max(need, a);
    SL_CU(cub::DeviceScan::InclusiveScan(nullptr, a, (uint32_t*)nullptr, (uint32_t*)nullptr, MaxU32(), (long long)cap, st));
    need = std::max(need, a);
    cub_bytes = need + 256;
    SL_CU(cudaMalloc(&cub_tmp, cub_bytes));
  }

  for (uint32_t b0 = 0; b0 < kBins;) {
    // next batch: bins [b0, b1) holding c suffixes
    uint64_t c = 0;
    uint32_t b1 = b0;
    while (b1 < kBins && c + hist[b1] <= cap) c += hist[b1++];
    if (b1 == b0) { rc = bkx_fail(BKX_ERR_CUDA, "suffix array: batch planning failed"); goto done; }
    if (c == 0) { b0 = b1; continue; }
    ++n_batches;
    const uint64_t base = done_sfx;
    // arena: [tied set A, 13c][round 0, 38c  |  re-used by the tie rounds, 46m]
    Carver cv{arena, 0, arena_bytes};
    uint64_t* pos_a = cv.take<uint64_t>(c);
    uint32_t* slot_a = cv.take<uint32_t>(c);
    uint8_t* head_a = cv.take<uint8_t>(c);
    const size_t rounds_off = cv.off;
    uint64_t* k0 = cv.take<uint64_t>(c);
    uint64_t* k1 = cv.take<uint64_t>(c);
    uint64_t* v0 = cv.take<uint64_t>(c);
    uint64_t* v1 = cv.take<uint64_t>(c);
    uint8_t* head = cv.take<uint8_t>(c);
    uint8_t* tied = cv.take<uint8_t>(c);
    uint32_t* idx = cv.take<uint32_t>(c + 1);
    if (!idx) { rc = bkx_fail(BKX_ERR_CUDA, "suffix array: arena too small"); goto done; }
    SL_CU(cudaMemsetAsync(d_count, 0, 8, st));
    sl_gather<<<148 * 8, kThreads, 0, st>>>(d_seq, n, b0, b1, k0, v0, d_count);
    SL_CU(cudaGetLastError());
    {
      unsigned long long got = 0;
      SL_CU(cudaMemcpyAsync(&got, d_count, 8, cudaMemcpyDeviceToHost, st));
      SL_CU(cudaStreamSynchronize(st));
      if (got != c) { rc = bkx_fail(BKX_ERR_CUDA, "suffix array: batch gather count mismatch"); goto done; }
    }
    {
      cub::DoubleBuffer<uint64_t> dk(k0, k1);
      cub::DoubleBuffer<uint64_t> dv(v0, v1);
      size_t need = cub_bytes;
      SL_CU(cub::DeviceRadixSort::SortPairs(cub_tmp, need, dk, dv, (long long)c, 0, 63, st));
      k0 = dk.Current();
      v0 = dv.Current();
    }
    sl_first_pass<<<grid_for(c), 256, 0, st>>>(k0, v0, c, base, d_sa_lo, d_sa_hi, head, tied);
    SL_CU(cudaGetLastError());
    uint64_t m = 0;
    {
      cub::TransformInputIterator<uint32_t, U8ToU32, const uint8_t*> it(tied, U8ToU32());
      size_t need = cub_bytes;
      SL_CU(cub::DeviceScan::ExclusiveSum(cub_tmp, need, it, idx, (long long)c, st));
      uint32_t last_idx = 0;
      uint8_t last_flag = 0;
      SL_CU(cudaMemcpyAsync(&last_idx, idx + (c - 1), 4, cudaMemcpyDeviceToHost, st));
      SL_CU(cudaMemcpyAsync(&last_flag, tied + (c - 1), 1, cudaMemcpyDeviceToHost, st));
      SL_CU(cudaStreamSynchronize(st));
      m = (uint64_t)last_idx + last_flag;
    }
    if (m > 0) {
      sl_compact<<<grid_for(c), 256, 0, st>>>(v0, nullptr, head, tied, idx, c, pos_a, slot_a, head_a);
      SL_CU(cudaGetLastError());
      // the round-0 buffers are dead from here on
      Carver cr{arena, rounds_off, arena_bytes};
      uint64_t* pos_b = cr.take<uint64_t>(m);
      uint32_t* slot_b = cr.take<uint32_t>(m);
      uint8_t* head_b = cr.take<uint8_t>(m);
      uint32_t* grp = cr.take<uint32_t>(m);
      uint64_t* ka = cr.take<uint64_t>(m);
      uint64_t* kb = cr.take<uint64_t>(m);
      uint32_t* pa = cr.take<uint32_t>(m);
      uint32_t* pb = cr.take<uint32_t>(m);
      uint8_t* tied2 = cr.take<uint8_t>(m);
      uint32_t* idx2 = cr.take<uint32_t>(m + 1);
      if (!idx2) { rc = bkx_fail(BKX_ERR_CUDA, "suffix array: arena too small for the tied set"); goto done; }
      for (uint64_t r = 1; m > 0; ++r) {
        if (r * kWin > n + kWin) { rc = bkx_fail(BKX_ERR_CUDA, "suffix array: did not converge"); goto done; }
        ++rounds_total;
        const uint64_t shift = r * kWin;
        sl_round_keys<<<grid_for(m), 256, 0, st>>>(d_seq, n, pos_a, head_a, m, shift, ka, pa, grp);
        SL_CU(cudaGetLastError());
        size_t need = cub_bytes;
        SL_CU(cub::DeviceScan::InclusiveScan(cub_tmp, need, grp, grp, MaxU32(), (long long)m, st));
        uint32_t* perm;
        {
          cub::DoubleBuffer<uint64_t> dk(ka, kb);
          cub::DoubleBuffer<uint32_t> dp(pa, pb);
          need = cub_bytes;
          SL_CU(cub::DeviceRadixSort::SortPairs(cub_tmp, need, dk, dp, (long long)m, 0, 63, st));
          perm = dp.Current();
        }
        {  // stable pass on the group id; the window buffers are free and hold the two id arrays
          uint32_t* ga = (uint32_t*)ka;
          uint32_t* gb = (uint32_t*)kb;
          uint32_t* other = perm == pa ? pb : pa;
          sl_gather_grp<<<grid_for(m), 256, 0, st>>>(grp, perm, m, ga);
          SL_CU(cudaGetLastError());
          int bits = 1;
          while (bits < 32 && (1ull << bits) < m) ++bits;
          cub::DoubleBuffer<uint32_t> dg(ga, gb);
          cub::DoubleBuffer<uint32_t> dp(perm, other);
          need = cub_bytes;
          SL_CU(cub::DeviceRadixSort::SortPairs(cub_tmp, need, dg, dp, (long long)m, 0, bits, st));
          perm = dp.Current();
        }
        sl_apply<<<grid_for(m), 256, 0, st>>>(d_seq, n, pos_a, perm, m, shift, pos_b, ka);
        sl_regroup<<<grid_for(m), 256, 0, st>>>(ka, pos_b, slot_a, head_a, m, base, d_sa_lo, d_sa_hi, head_b, tied2);
        SL_CU(cudaGetLastError());
        cub::TransformInputIterator<uint32_t, U8ToU32, const uint8_t*> it(tied2, U8ToU32());
        need = cub_bytes;
        SL_CU(cub::DeviceScan::ExclusiveSum(cub_tmp, need, it, idx2, (long long)m, st));
        uint32_t last_idx = 0;
        uint8_t last_flag = 0;
        SL_CU(cudaMemcpyAsync(&last_idx, idx2 + (m - 1), 4, cudaMemcpyDeviceToHost, st));
        SL_CU(cudaMemcpyAsync(&last_flag, tied2 + (m - 1), 1, cudaMemcpyDeviceToHost, st));
        SL_CU(cudaStreamSynchronize(st));
        const uint64_t m2 = (uint64_t)last_idx + last_flag;
        if (m2 > 0) {
          // survivors back into the A set (slot/head of the survivors come from slot_a / head_b)
          sl_compact<<<grid_for(m), 256, 0, st>>>(pos_b, slot_a, head_b, tied2, idx2, m, (uint64_t*)kb, slot_b, (uint8_t*)pb);
          SL_CU(cudaGetLastError());
          SL_CU(cudaMemcpyAsync(pos_a, kb, m2 * 8, cudaMemcpyDeviceToDevice, st));
          SL_CU(cudaMemcpyAsync(slot_a, slot_b, m2 * 4, cudaMemcpyDeviceToDevice, st));
          SL_CU(cudaMemcpyAsync(head_a, pb, m2, cudaMemcpyDeviceToDevice, st));
        }
        m = m2;
      }
      (void)head_b;
    }
    SL_CU(cudaStreamSynchronize(st));
    done_sfx += c;
    b0 = b1;
  }
  if (done_sfx != n) { rc = bkx_fail(BKX_ERR_CUDA, "suffix array: %llu of %llu suffixes placed", (unsigned long long)done_sfx, (unsigned long long)n); goto done; }
  if (getenv("BKX_TRACE"))
    fprintf(stderr, "[bkx trace] suffix array: %llu symbols, %d batch(es) of <= %llu, %llu tie rounds\n", (unsigned long long)n,
            n_batches, (unsigned long long)cap, (unsigned long long)rounds_total);
done:
  if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  cudaFree(d_hist); cudaFree(d_count); cudaFree(arena); cudaFree(cub_tmp);
  return rc;
}
