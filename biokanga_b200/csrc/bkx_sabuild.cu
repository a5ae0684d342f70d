// GPU suffix-array construction for synthetic / bench genomes, and a writer for the reference's .sfx
// container -- the two halves of what `biokanga index` does (kangax.cpp:774-926 ->
// CSfxArrayV3::AddEntry / Finalise / QSortSeq, SfxArrayV2.cpp:1439, 814-870, 9451-9487).
//
// The reference sorts suffix offsets with a comparison quicksort over 4-bit symbols
// (A<C<G<T<N<EOS), comparing straight through the EOS terminators (SfxArrayV2.cpp:9491-9510).
// Here the same order is produced by prefix doubling on the device:
//   round 0  radix-sort all suffixes by their first 21 symbols (3 bits each in a 63-bit key);
//   round r  only suffixes still tied: radix-sort (group rank, rank of suffix + h) pairs, h = 21*2^(r-1).
// Suffixes that tie all the way to the end of the concatenation order shorter-first; the reference
// reads past its buffer there, so that corner has no defined order to reproduce.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>

#include "../../include/bkx.h"

int bkx_fail(int code, const char* fmt, ...);

namespace {

constexpr int kSymPerKey = 21;

#define SA_CU(call)                                                                                   \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess) {                                                                         \
      rc = bkx_fail(BKX_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
      goto done;                                                                                      \
    }                                                                                                 \
  } while (0)

__global__ void sa_init_keys(const uint8_t* __restrict__ seq, uint64_t n, uint64_t* __restrict__ keys,
                             uint32_t* __restrict__ vals) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint64_t k = 0;
#pragma unroll
    for (int j = 0; j < kSymPerKey; ++j) {
      uint64_t p = i + j;
      unsigned s = 0;  // past the end sorts lowest
      if (p < n) {
        unsigned b = __ldg(seq + p) & 0x0f;
        s = b < 5 ? b + 1 : 7;
      }
      k = (k << 3) | s;
    }
    keys[i] = k;
    vals[i] = (uint32_t)i;
  }
}

// head[i] = 1 if sorted key i starts a new group
__global__ void sa_heads(const uint64_t* __restrict__ keys, uint64_t m, uint8_t* __restrict__ head) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x)
    head[i] = (i == 0 || keys[i] != keys[i - 1]) ? 1 : 0;
}

// gstart[i] = head[i] ? i : 0  (an inclusive max-scan turns this into "index of my group's head")
__global__ void sa_head_index(const uint8_t* __restrict__ head, uint64_t m, uint32_t* __restrict__ gstart) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x)
    gstart[i] = head[i] ? (uint32_t)i : 0u;
}

__global__ void sa_unresolved(const uint8_t* __restrict__ head, uint64_t m, uint8_t* __restrict__ unres) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x)
    unres[i] = (head[i] && (i + 1 == m || head[i + 1])) ? 0 : 1;
}

// round 0: rank[sa[i]] = group head index
__global__ void sa_scatter_rank0(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ gstart, uint64_t n,
                                 uint32_t* __restrict__ rank) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
    rank[sa[i]] = gstart[i];
}

// later rounds: key = (rank of own group, 1 + rank of the suffix h further on), value = suffix position
__global__ void sa_round_keys(const uint32_t* __restrict__ sa, const uint32_t* __restrict__ slots, uint64_t m,
                              const uint32_t* __restrict__ rank, uint64_t n, uint64_t h, uint64_t* __restrict__ keys,
                              uint32_t* __restrict__ vals) {
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t p = sa[slots[k]];
    uint64_t q = (uint64_t)p + h;
    uint64_t r2 = q < n ? (uint64_t)rank[q] + 1 : 0;
    keys[k] = ((uint64_t)rank[p] << 32) | r2;
    vals[k] = p;
  }
}

// write the re-ordered suffixes back and give every suffix the slot index of its (sub)group head
__global__ void sa_round_apply(const uint32_t* __restrict__ slots, const uint32_t* __restrict__ vals,
                               const uint32_t* __restrict__ ghead, uint64_t m, uint32_t* __restrict__ sa,
                               uint32_t* __restrict__ rank) {
  for (uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; k < m; k += (uint64_t)gridDim.x * blockDim.x) {
    uint32_t p = vals[k];
    sa[slots[k]] = p;
    rank[p] = slots[ghead[k]];
  }
}

struct MaxOp {
  __device__ __forceinline__ uint32_t operator()(uint32_t a, uint32_t b) const { return a > b ? a : b; }
};

int grid_for(uint64_t n) {
  uint64_t g = (n + 255) / 256;
  return (int)std::min<uint64_t>(std::max<uint64_t>(g, 1), 148 * 32);
}

// compact `slots_in[i]` (or i itself when slots_in == nullptr) where flag[i] != 0; chunked so that
// every CUB call sees fewer than 2^30 items.
int compact(const uint32_t* slots_in, const uint8_t* flag, uint64_t m, uint32_t* out, uint64_t* out_count,
            cudaStream_t st) {
  int rc = BKX_OK;
  const uint64_t chunk = 1ull << 30;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  int* d_num = nullptr;
  uint64_t total = 0;
  SA_CU(cudaMalloc((void**)&d_num, sizeof(int)));
  for (uint64_t s = 0; s < m; s += chunk) {
    int cnt = (int)std::min<uint64_t>(chunk, m - s);
    size_t need = 0;
    if (slots_in) {
      SA_CU(cub::DeviceSelect::Flagged(nullptr, need, slots_in + s, flag + s, out + total, d_num, cnt, st));
    } else {
      cub::CountingInputIterator<uint32_t> it((uint32_t)s);
      SA_CU(cub::DeviceSelect::Flagged(nullptr, need, it, flag + s, out + total, d_num, cnt, st));
    }
    if (need > tmp_bytes) {
      if (tmp) cudaFree(tmp);
      tmp = nullptr;
      SA_CU(cudaMalloc(&tmp, need));
      tmp_bytes = need;
    }
    if (slots_in) {
      SA_CU(cub::DeviceSelect::Flagged(tmp, need, slots_in + s, flag + s, out + total, d_num, cnt, st));
    } else {
      cub::CountingInputIterator<uint32_t> it((uint32_t)s);
      SA_CU(cub::DeviceSelect::Flagged(tmp, need, it, flag + s, out + total, d_num, cnt, st));
    }
    int h_num = 0;
    SA_CU(cudaMemcpyAsync(&h_num, d_num, sizeof(int), cudaMemcpyDeviceToHost, st));
    SA_CU(cudaStreamSynchronize(st));
    total += (uint64_t)h_num;
  }
  *out_count = total;
done:
  if (tmp) cudaFree(tmp);
  if (d_num) cudaFree(d_num);
  return rc;
}

int count_flags(const uint8_t* flag, uint64_t m, uint64_t* out, cudaStream_t st);

__global__ void sa_count_flags(const uint8_t* __restrict__ flag, uint64_t m, unsigned long long* __restrict__ out) {
  unsigned long long c = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < m; i += (uint64_t)gridDim.x * blockDim.x)
    c += flag[i];
  for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if ((threadIdx.x & 31) == 0 && c) atomicAdd(out, c);
}

int count_flags(const uint8_t* flag, uint64_t m, uint64_t* out, cudaStream_t st) {
  int rc = BKX_OK;
  unsigned long long* d = nullptr;
  unsigned long long h = 0;
  SA_CU(cudaMalloc((void**)&d, 8));
  SA_CU(cudaMemsetAsync(d, 0, 8, st));
  sa_count_flags<<<grid_for(m), 256, 0, st>>>(flag, m, d);
  SA_CU(cudaGetLastError());
  SA_CU(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, st));
  SA_CU(cudaStreamSynchronize(st));
  *out = h;
done:
  if (d) cudaFree(d);
  return rc;
}

int max_scan_inplace(uint32_t* data, uint64_t m, cudaStream_t st) {
  int rc = BKX_OK;
  void* tmp = nullptr;
  size_t need = 0;
  SA_CU(cub::DeviceScan::InclusiveScan(nullptr, need, data, data, MaxOp(), (long long)m, st));
  SA_CU(cudaMalloc(&tmp, need));
  SA_CU(cub::DeviceScan::InclusiveScan(tmp, need, data, data, MaxOp(), (long long)m, st));
  SA_CU(cudaStreamSynchronize(st));
done:
  if (tmp) cudaFree(tmp);
  return rc;
}

int sort_pairs(uint64_t*& k_in, uint64_t*& k_alt, uint32_t*& v_in, uint32_t*& v_alt, uint64_t m, int end_bit,
               cudaStream_t st) {
  int rc = BKX_OK;
  void* tmp = nullptr;
  size_t need = 0;
  cub::DoubleBuffer<uint64_t> dk(k_in, k_alt);
  cub::DoubleBuffer<uint32_t> dv(v_in, v_alt);
  SA_CU(cub::DeviceRadixSort::SortPairs(nullptr, need, dk, dv, (long long)m, 0, end_bit, st));
  SA_CU(cudaMalloc(&tmp, need));
  SA_CU(cub::DeviceRadixSort::SortPairs(tmp, need, dk, dv, (long long)m, 0, end_bit, st));
  SA_CU(cudaStreamSynchronize(st));
  k_in = dk.Current(); k_alt = dk.Alternate();
  v_in = dv.Current(); v_alt = dv.Alternate();
done:
  if (tmp) cudaFree(tmp);
  return rc;
}

}  // namespace

extern "C" int bkx_build_suffix_array_device(const uint8_t* d_seq, uint64_t n, uint32_t* d_sa, int device) {
  if (!d_seq || !d_sa) return bkx_fail(BKX_ERR_PARAM, "null argument");
  if (n < 2) return bkx_fail(BKX_ERR_PARAM, "sequence too short");
  if (n >= 4000000000ull)
    return bkx_fail(BKX_ERR_UNSUPPORTED, "suffix arrays with 5-byte elements (>= 4e9 symbols) are not built on the GPU yet");
  int rc = BKX_OK;
  cudaStream_t st = nullptr;
  uint64_t *k0 = nullptr, *k1 = nullptr;
  uint32_t *v0 = nullptr, *v1 = nullptr, *rank = nullptr, *gstart = nullptr, *slots = nullptr, *slots2 = nullptr;
  uint8_t *head = nullptr, *unres = nullptr;
  uint64_t m = 0;
  SA_CU(cudaSetDevice(device));
  SA_CU(cudaDeviceSynchronize());  // whatever stream produced d_seq: this call works on its own non-blocking stream
  SA_CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  SA_CU(cudaMalloc((void**)&k0, n * 8));
  SA_CU(cudaMalloc((void**)&k1, n * 8));
  SA_CU(cudaMalloc((void**)&v0, n * 4));
  SA_CU(cudaMalloc((void**)&v1, n * 4));
  sa_init_keys<<<grid_for(n), 256, 0, st>>>(d_seq, n, k0, v0);
  SA_CU(cudaGetLastError());
  if ((rc = sort_pairs(k0, k1, v0, v1, n, 63, st)) < 0) goto done;
  SA_CU(cudaMalloc((void**)&head, n));
  SA_CU(cudaMalloc((void**)&unres, n));
  sa_heads<<<grid_for(n), 256, 0, st>>>(k0, n, head);
  SA_CU(cudaGetLastError());
  SA_CU(cudaStreamSynchronize(st));
  cudaFree(k0); cudaFree(k1); k0 = k1 = nullptr;
  SA_CU(cudaMemcpyAsync(d_sa, v0, n * 4, cudaMemcpyDeviceToDevice, st));
  SA_CU(cudaStreamSynchronize(st));
  cudaFree(v1); v1 = nullptr;
  SA_CU(cudaMalloc((void**)&gstart, n * 4));
  SA_CU(cudaMalloc((void**)&rank, n * 4));
  sa_head_index<<<grid_for(n), 256, 0, st>>>(head, n, gstart);
  SA_CU(cudaGetLastError());
  if ((rc = max_scan_inplace(gstart, n, st)) < 0) goto done;
  sa_scatter_rank0<<<grid_for(n), 256, 0, st>>>(d_sa, gstart, n, rank);
  sa_unresolved<<<grid_for(n), 256, 0, st>>>(head, n, unres);
  SA_CU(cudaGetLastError());
  SA_CU(cudaStreamSynchronize(st));
  cudaFree(gstart); gstart = nullptr;
  cudaFree(v0); v0 = nullptr;
  if ((rc = count_flags(unres, n, &m, st)) < 0) goto done;
  if (m > 0) {
    SA_CU(cudaMalloc((void**)&slots, m * 4));
    uint64_t got = 0;
    if ((rc = compact(nullptr, unres, n, slots, &got, st)) < 0) goto done;
    if (got != m) { rc = bkx_fail(BKX_ERR_CUDA, "suffix array: compaction count mismatch"); goto done; }
  }
  cudaFree(head); head = nullptr;
  cudaFree(unres); unres = nullptr;
  for (uint64_t h = kSymPerKey; m > 0; h *= 2) {
    if (h > (1ull << 40)) { rc = bkx_fail(BKX_ERR_CUDA, "suffix array: did not converge"); goto done; }
    SA_CU(cudaMalloc((void**)&k0, m * 8));
    SA_CU(cudaMalloc((void**)&k1, m * 8));
    SA_CU(cudaMalloc((void**)&v0, m * 4));
    SA_CU(cudaMalloc((void**)&v1, m * 4));
    SA_CU(cudaMalloc((void**)&head, m));
    SA_CU(cudaMalloc((void**)&unres, m));
    SA_CU(cudaMalloc((void**)&gstart, m * 4));
    sa_round_keys<<<grid_for(m), 256, 0, st>>>(d_sa, slots, m, rank, n, h, k0, v0);
    SA_CU(cudaGetLastError());
    if ((rc = sort_pairs(k0, k1, v0, v1, m, 64, st)) < 0) goto done;
    sa_heads<<<grid_for(m), 256, 0, st>>>(k0, m, head);
    sa_head_index<<<grid_for(m), 256, 0, st>>>(head, m, gstart);
    SA_CU(cudaGetLastError());
    if ((rc = max_scan_inplace(gstart, m, st)) < 0) goto done;
    sa_round_apply<<<grid_for(m), 256, 0, st>>>(slots, v0, gstart, m, d_sa, rank);
    sa_unresolved<<<grid_for(m), 256, 0, st>>>(head, m, unres);
    SA_CU(cudaGetLastError());
    uint64_t m2 = 0;
    if ((rc = count_flags(unres, m, &m2, st)) < 0) goto done;
    if (m2 > 0) {
      SA_CU(cudaMalloc((void**)&slots2, m2 * 4));
      uint64_t got = 0;
      if ((rc = compact(slots, unres, m, slots2, &got, st)) < 0) goto done;
      if (got != m2) { rc = bkx_fail(BKX_ERR_CUDA, "suffix array: compaction count mismatch"); goto done; }
    }
    SA_CU(cudaStreamSynchronize(st));
    cudaFree(slots); slots = slots2; slots2 = nullptr;
    cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(head); cudaFree(unres); cudaFree(gstart);
    k0 = k1 = nullptr; v0 = v1 = nullptr; head = unres = nullptr; gstart = nullptr;
    m = m2;
  }
done:
  if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  cudaFree(k0); cudaFree(k1); cudaFree(v0); cudaFree(v1); cudaFree(rank); cudaFree(gstart); cudaFree(slots);
  cudaFree(slots2); cudaFree(head); cudaFree(unres);
  return rc;
}

// ---- .sfx writer: tsSfxHeaderV3 (pack 4, 1224 bytes), tsSfxBlock (pack 1, 20-byte header), then
//      tsSfxEntriesBlock with 111-byte tsSfxEntry records (SfxArrayV2.h:79-104,174-187) ---------------
static uint16_t gen_hash16(const char* s) {  // CUtility::GenHash16, libbiokanga/Utility.cpp:17-35
  int h = 19937;
  if (!s || !*s) return 0;
  for (; *s; ++s) {
    int c = *s;
    if (c >= 'A' && c <= 'Z') c += 32;
    h = (h ^ c) * 3119;
    h ^= (h >> 13);
    h &= 0xffff;
  }
  if (h == 0) h = 19937;
  return (uint16_t)h;
}

extern "C" int bkx_write_sfx(const char* path, const uint8_t* seq, uint64_t n, const void* sa, uint32_t el,
                             const bkx_entry* entries, uint32_t n_ent, const char* name) {
  if (!path || !seq || !sa || !entries) return bkx_fail(BKX_ERR_PARAM, "null argument");
  if (el != 4 && el != 5) return bkx_fail(BKX_ERR_PARAM, "bad suffix element size %u", el);
  FILE* f = fopen(path, "wb");
  if (!f) return bkx_fail(BKX_ERR_FILE, "unable to create '%s'", path);
  const uint64_t blk_ofs = 1224, blk_size = 20 + n + n * el, ent_ofs = blk_ofs + blk_size;
  const uint32_t ent_size = 8 + 111 * n_ent;
  std::vector<uint8_t> hdr(1224, 0);
  memcpy(hdr.data(), "sfx5", 4);
  uint32_t u32; uint64_t u64;
  auto p32 = [&](size_t o, uint32_t v) { u32 = v; memcpy(hdr.data() + o, &u32, 4); };
  auto p64 = [&](size_t o, uint64_t v) { u64 = v; memcpy(hdr.data() + o, &u64, 8); };
  p32(4, 5); p32(8, 0); p64(12, ent_ofs + ent_size); p64(20, ent_ofs); p32(28, ent_size); p32(32, 1);
  p64(36, blk_size); p64(44, blk_ofs);
  const char* nm = name ? name : "bkx";
  snprintf((char*)hdr.data() + 52, 81, "%s", nm);
  snprintf((char*)hdr.data() + 133, 1024, "%s", nm);
  snprintf((char*)hdr.data() + 1157, 64, "%s", nm);
  bool ok = fwrite(hdr.data(), 1, hdr.size(), f) == hdr.size();
  uint8_t bh[20];
  u32 = 1; memcpy(bh, &u32, 4);
  u32 = n_ent; memcpy(bh + 4, &u32, 4);
  u64 = n; memcpy(bh + 8, &u64, 8);
  u32 = el; memcpy(bh + 16, &u32, 4);
  ok = ok && fwrite(bh, 1, 20, f) == 20;
  auto wr_big = [&](const uint8_t* p, uint64_t len) {
    while (len && ok) {
      size_t c = (size_t)std::min<uint64_t>(len, 1ull << 30);
      ok = fwrite(p, 1, c, f) == c;
      p += c; len -= c;
    }
  };
  wr_big(seq, n);
  wr_big((const uint8_t*)sa, n * el);
  std::vector<uint8_t> eb(ent_size, 0);
  u32 = n_ent; memcpy(eb.data(), &u32, 4); memcpy(eb.data() + 4, &u32, 4);
  for (uint32_t i = 0; i < n_ent; ++i) {
    uint8_t* e = eb.data() + 8 + 111 * (size_t)i;
    u32 = entries[i].entry_id; memcpy(e, &u32, 4);
    u32 = 1; memcpy(e + 4, &u32, 4);
    snprintf((char*)e + 8, 81, "%s", entries[i].name);
    uint16_t h16 = gen_hash16(entries[i].name); memcpy(e + 89, &h16, 2);
    u32 = entries[i].seq_len; memcpy(e + 91, &u32, 4);
    u64 = entries[i].start_ofs; memcpy(e + 95, &u64, 8);
    u64 = entries[i].end_ofs; memcpy(e + 103, &u64, 8);
  }
  ok = ok && fwrite(eb.data(), 1, eb.size(), f) == eb.size();
  ok = (fclose(f) == 0) && ok;
  if (!ok) return bkx_fail(BKX_ERR_FILE, "write to '%s' failed", path);
  return BKX_OK;
}
