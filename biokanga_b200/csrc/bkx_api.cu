// Host side of libbkx.so: the extern "C" entry points declared in include/bkx.h.
// Owns device memory, streams and staging; launches the kernels of bkx_kernels.cu.
// There is deliberately no CPU fallback: every compute entry point fails with BKX_ERR_CUDA when
// no CUDA device is usable.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>

#include "bkx_align.cuh"
#include "bkx_fast.cuh"
#include "bkx_wave.cuh"
#include "bkx_rescue.cuh"
#include "bkx_kernels.h"

using namespace bkx;

static thread_local std::string g_err;

int bkx_fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define fail bkx_fail

// BKX_TRACE=1: stage timings of the index load on stderr
#include <chrono>
struct StageTimer {
  bool on = getenv("BKX_TRACE") != nullptr;
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  void mark(const char* what) {
    if (!on) return;
    auto t1 = std::chrono::steady_clock::now();
    fprintf(stderr, "[bkx trace] %-34s %.3f s\n", what, std::chrono::duration<double>(t1 - t0).count());
    t0 = t1;
  }
};

#define CU(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess) return fail(BKX_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(e__), \
                                        __FILE__, __LINE__);                                          \
  } while (0)

// Host-buffer calls pipeline slices of reads through this many slots (own stream, staging, cursors, lane hash sets).
// Three, not two: a slice's general kernel only gets SMs after the NEXT slice's fast kernel (whose blocks move in as
// this slice's fast blocks retire), so a slice completes one kernel late and two slots would run in lock step.
constexpr int kSlots = 3;
// All kernels of the host-buffer calls go down ONE compute stream (bkx_index::cst) in slice order; the slots' own
// streams only carry their H2D / D2H copies, tied in with events.  (With kernels on the slots' streams the persistent
// fast kernel of the next slice moves onto the SMs as this slice's fast blocks retire, and this slice's general kernel --
// stream-ordered behind its fast kernel -- waits a whole kernel for room: measured 72 ms instead of 40 ms per 20 M reads.)

struct Slot {
  cudaStream_t st = nullptr;
  uint8_t* d_bases = nullptr;
  size_t bases_cap = 0;
  uint8_t* d_packed = nullptr;   // staging for 4-bit packed input (bkx_align_reads_packed4)
  size_t packed_cap = 0;
  bkx_multi_hit* d_multi = nullptr;   // -r5 loci of the slice's reads
  size_t multi_cap = 0;
  uint32_t* d_orphans = nullptr;      // fused PE call: pairs of the slice whose mate needs recovery
  size_t orphans_cap = 0;
  uint64_t* d_offs = nullptr;
  bkx_read_result* d_out = nullptr;
  // compact host interface (bkx_align_reads_packed2): 16-byte records out, u16 lengths / exception list in
  bkx_read_result16* d_out16 = nullptr;
  uint16_t* d_lens = nullptr;
  uint64_t* d_exc_pos = nullptr;
  uint8_t* d_exc_code = nullptr;
  size_t exc_cap = 0;
  void* d_scan_tmp = nullptr;
  size_t scan_tmp_bytes = 0;
  uint8_t* d_rflags = nullptr;   // per read of the slice: holds a non-ACGT base
  uint32_t* d_hard = nullptr;   // reads the fast kernel deferred to the general kernel
  size_t hard_cap = 0;
  WaveBuf wave;                 // scratch of the wave path (bkx_wave.cuh), for wave_cap reads
  size_t wave_cap = 0;
  size_t reads_cap = 0;
  cudaEvent_t k0 = nullptr, k1 = nullptr;
  cudaEvent_t in_ready = nullptr;   // this slot's H2D copies are done (the compute stream waits on it)
  bool timed = false;
};

struct bkx_index {
  int device = 0;
  DevIndex d{};
  std::vector<void*> owned;
  struct Arr { size_t field_ofs; size_t bytes; };
  std::vector<Arr> arrs;   // every device array the DevIndex points at (for peer replication)
  std::vector<bkx_entry> entries;
  bkx_index_info info{};
  // runtime workspace
  std::mutex mtx;
  Slot slot[kSlots];
  cudaStream_t cst = nullptr;
  unsigned int* d_cursor[kSlots] = {};   // per slot: [0] fast cursor, [1] general cursor, [2] deferred count
  int fast_grid = 0;
  int fast_W = 0;
  uint64_t max_len_prepared = 0;
  uint32_t* d_pe_list = nullptr;
  size_t pe_list_cap = 0;
  uint8_t* d_chrom_keep = nullptr;   // bkx_set_chrom_filter: AcceptThisChromID per chromosome id, NULL = no -Z / -z
  uint64_t* fast_hash[kSlots] = {};   // per slot (launches of different slots overlap): lane-private overflow sets
                                      // of the fast kernel, fast_grid x 256 lanes x 1024 slots, allocated on first use
  size_t fast_hash_lanes = 0;
  uint32_t fast_epoch = 1;
  HashPool hp{};
  int grid = 0;
  int grid_W = 0;
  bkx_align_stats* d_stats = nullptr;
  bkx_pe_stats* d_pe_stats = nullptr;
  uint32_t* d_len_dist = nullptr;
  float last_ms = -1.f;
  uint64_t launches = 0;
};

extern "C" int bkx_abi_version(void) { return BKX_ABI_VERSION; }
extern "C" const char* bkx_last_error(void) { return g_err.c_str(); }
extern "C" int bkx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

template <typename T>
static int dev_alloc(bkx_index* x, T** p, size_t count, bool zero) {
  void* q = nullptr;
  size_t bytes = count * sizeof(T);
  if (bytes == 0) bytes = sizeof(T);
  CU(cudaMalloc(&q, bytes));
  // NOT cudaMemset: that is asynchronous on the legacy default stream, which the library's non-blocking streams do
  // not wait for -- the kernels that fill these arrays run on slot[0].st
  if (zero) CU(cudaMemsetAsync(q, 0, bytes, x->slot[0].st));
  x->owned.push_back(q);
  x->info.device_bytes += bytes;
  *p = (T*)q;
  return BKX_OK;
}

#define REG_ARR(x, field, nbytes) (x)->arrs.push_back({offsetof(DevIndex, field), (size_t)(nbytes)})

static int choose_k(uint64_t n, int requested, size_t free_bytes, bool wide) {
  if (requested > 0) return std::min(std::max(requested, 4), 17);
  // ceil(log4 n) + 1: on average 1/4 suffix per bucket, so most absent cores die at the table lookup
  // without touching the suffix array (measured: k=17 beats 16 by 8 % at 3.1 G symbols).  HBM is there
  // to be used (180 GB): the table may take up to 55 % of what is free -- for indexes of >= 2^32 symbols (wheat scale:
  // 70 GB of suffix array, where a table one size down leaves 3.3 suffixes per bucket and nearly every absent core then
  // costs a suffix-array and a genome fetch) everything but 14 GB for reads, records and the kernels' hash sets.
  int k = 8;
  while (k < 17 && (1ull << (2 * (k - 1))) < n) ++k;
  const double budget = wide ? std::max(0.55 * (double)free_bytes, (double)free_bytes - 14.0 * (double)(1ull << 30))
                             : 0.55 * (double)free_bytes;
  while (k > 8 && (double)((1ull << (2 * k)) + 1) * 4.0 > budget) --k;
  return k;
}

// Where the suffix array comes from: raw elements as the .sfx stores them (copied / split into planes here), or
// planes that are already in place on the device (used as they are; the caller settles who frees them).
static thread_local bool g_selfcheck_failed = false;   // set by finish_index, read by bkx_open_index's retry loop

struct SaSrc {
  const void* raw = nullptr;      // elements as the .sfx stores them (4 or 5 bytes each): copied
  const uint32_t* lo = nullptr;   // planes already on the device: used as they are (4-byte elements) or merged into 5-byte
  const uint8_t* hi = nullptr;    //   elements (copied)
  const uint8_t* packed5 = nullptr;   // 5-byte elements already on the device, 8-byte aligned with 16 bytes of slack: used as they are
};

// Build every derived structure from a device-resident 1-byte/base sequence and the suffix array.
static int finish_index(bkx_index* x, const uint8_t* d_seq, uint64_t n, const SaSrc& sa, uint32_t el,
                        const bkx_entry* entries, uint32_t n_ent, const char* name, int prefix_k) {
  if (el != 4 && el != 5) return fail(BKX_ERR_FORMAT, "unsupported suffix element size %u", el);
  if (n < 2 || n_ent == 0) return fail(BKX_ERR_FORMAT, "empty index");
  x->info.concat_len = n;
  x->info.sfx_el_size = el;
  x->info.num_entries = n_ent;
  x->info.device = (uint32_t)x->device;
  if (name) snprintf(x->info.dataset_name, sizeof(x->info.dataset_name), "%s", name);
  x->entries.assign(entries, entries + n_ent);
  std::sort(x->entries.begin(), x->entries.end(),
            [](const bkx_entry& a, const bkx_entry& b) { return a.start_ofs < b.start_ofs; });
  uint64_t tot = 0;
  for (auto& e : x->entries) tot += e.seq_len;
  x->info.tot_seq_len = tot;

  StageTimer tm;
  cudaStream_t st = x->slot[0].st;
  uint64_t* g2; uint64_t* gx; uint32_t* gxc;
  size_t g2w = ((n + 63) >> 6) * 2 + 4, gxw = ((n + 63) >> 6) + 2, gcw = (((n + 63) >> 6) + 31) / 32 + 2;
  int rc;
  if ((rc = dev_alloc(x, &g2, g2w, true)) < 0) return rc;
  if ((rc = dev_alloc(x, &gx, gxw, true)) < 0) return rc;
  if ((rc = dev_alloc(x, &gxc, gcw, true)) < 0) return rc;
  unsigned long long* d_bad;
  CU(cudaMalloc((void**)&d_bad, 8));
  CU(cudaMemsetAsync(d_bad, 0, 8, st));
  CU(launch_pack_genome(d_seq, n, g2, gx, gxc, d_bad, st));
  unsigned long long bad = 0;
  CU(cudaMemcpyAsync(&bad, d_bad, 8, cudaMemcpyDeviceToHost, st));
  CU(cudaStreamSynchronize(st));
  cudaFree(d_bad);
  x->launches += 1;
  tm.mark("finish_index: pack genome");
  if (bad) return fail(BKX_ERR_UNSUPPORTED, "%llu symbols other than A,C,G,T,N,EOS in the index sequence", bad);
  x->d.g2 = g2; x->d.gx = gx; x->d.gxc = gxc; x->d.n = n;
  REG_ARR(x, g2, g2w * 8); REG_ARR(x, gx, gxw * 8); REG_ARR(x, gxc, gcw * 4);

  if (sa.packed5) {
    if ((uintptr_t)sa.packed5 & 7) return fail(BKX_ERR_PARAM, "5-byte suffix elements must start on an 8-byte boundary");
    x->d.sa5 = sa.packed5;
    REG_ARR(x, sa5, n * 5 + 16);
  } else if (sa.lo && el == 4) {
    x->d.sa_lo = sa.lo;
    REG_ARR(x, sa_lo, n * 4);
  } else if (sa.lo) {   // borrowed planes, 5-byte elements: merged into the one-fetch layout (the planes stay the caller's)
    if (!sa.hi) return fail(BKX_ERR_PARAM, "5-byte suffix elements need the high plane");
    uint8_t* s5;
    if ((rc = dev_alloc(x, &s5, n * 5 + 16, false)) < 0) return rc;
    CU(launch_merge_sa5(sa.lo, sa.hi, n, s5, st));
    x->launches += 1;
    x->d.sa5 = s5;
    REG_ARR(x, sa5, n * 5 + 16);
  } else if (el == 4) {
    uint32_t* lo;
    if ((rc = dev_alloc(x, &lo, n, false)) < 0) return rc;
    CU(cudaMemcpyAsync(lo, sa.raw, n * 4, cudaMemcpyDeviceToDevice, st));
    x->d.sa_lo = lo;
    REG_ARR(x, sa_lo, n * 4);
  } else {
    uint8_t* s5;
    if ((rc = dev_alloc(x, &s5, n * 5 + 16, false)) < 0) return rc;
    CU(cudaMemcpyAsync(s5, sa.raw, n * 5, cudaMemcpyDeviceToDevice, st));
    x->d.sa5 = s5;
    REG_ARR(x, sa5, n * 5 + 16);
  }
  // chromosome table
  std::vector<uint64_t> es(n_ent), ee(n_ent);
  std::vector<uint32_t> ei(n_ent);
  for (uint32_t i = 0; i < n_ent; ++i) { es[i] = x->entries[i].start_ofs; ee[i] = x->entries[i].end_ofs; ei[i] = x->entries[i].entry_id; }
  uint64_t *d_es, *d_ee; uint32_t* d_ei;
  if ((rc = dev_alloc(x, &d_es, n_ent, false)) < 0) return rc;
  if ((rc = dev_alloc(x, &d_ee, n_ent, false)) < 0) return rc;
  if ((rc = dev_alloc(x, &d_ei, n_ent, false)) < 0) return rc;
  CU(cudaMemcpy(d_es, es.data(), n_ent * 8, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_ee, ee.data(), n_ent * 8, cudaMemcpyHostToDevice));
  CU(cudaMemcpy(d_ei, ei.data(), n_ent * 4, cudaMemcpyHostToDevice));
  x->d.ent_start = d_es; x->d.ent_end = d_ee; x->d.ent_id = d_ei; x->d.n_ent = n_ent;
  REG_ARR(x, ent_start, (size_t)n_ent * 8); REG_ARR(x, ent_end, (size_t)n_ent * 8); REG_ARR(x, ent_id, (size_t)n_ent * 4);
  {  // entry id -> sorted index
    uint32_t max_id = 0;
    for (uint32_t i = 0; i < n_ent; ++i) max_id = std::max(max_id, ei[i]);
    std::vector<uint32_t> inv((size_t)max_id + 1, 0xffffffffu);
    for (uint32_t i = 0; i < n_ent; ++i) inv[ei[i]] = i;
    uint32_t* d_inv;
    if ((rc = dev_alloc(x, &d_inv, inv.size(), false)) < 0) return rc;
    CU(cudaMemcpy(d_inv, inv.data(), inv.size() * 4, cudaMemcpyHostToDevice));
    x->d.ent_of_id = d_inv;
    x->d.max_ent_id = max_id;
    REG_ARR(x, ent_of_id, inv.size() * 4);
  }
  {  // coarse offset -> entry table (<= 1M blocks): first entry whose end is at or after the block start
    uint32_t shift = 8;
    while (((n >> shift) + 1) > (1u << 20)) ++shift;
    size_t nb = (size_t)(n >> shift) + 1;
    std::vector<uint32_t> lut(nb);
    uint32_t e = 0;
    for (size_t b = 0; b < nb; ++b) {
      uint64_t start = (uint64_t)b << shift;
      while (e < n_ent && ee[e] < start) ++e;
      lut[b] = e;
    }
    uint32_t* d_lut;
    if ((rc = dev_alloc(x, &d_lut, nb, false)) < 0) return rc;
    CU(cudaMemcpy(d_lut, lut.data(), nb * 4, cudaMemcpyHostToDevice));
    x->d.ent_lut = d_lut;
    x->d.lut_shift = shift;
    REG_ARR(x, ent_lut, nb * 4);
  }
  // prefix table
  size_t free_b = 0, total_b = 0;
  CU(cudaMemGetInfo(&free_b, &total_b));
  bool wide = n >= (1ull << 32) || getenv("BKX_FORCE_WIDE_PT") != nullptr;  // (env: test hook for the two-level table)
  int k = choose_k(n, prefix_k, free_b, wide);
  uint64_t pt_entries = (1ull << (2 * k)) + 1;
  uint32_t* pt = nullptr;
  uint64_t* pt_hi = nullptr;
  if ((rc = dev_alloc(x, &pt, pt_entries, false)) < 0) return rc;
  x->d.pt32 = pt;
  x->d.pt_hi = nullptr;   // set once the table is built: the histogram kernel reads the genome through x->d only
  x->d.k = k;
  x->info.prefix_k = (uint32_t)k;
  REG_ARR(x, pt32, pt_entries * 4);
  const uint64_t pt_blocks = (pt_entries + (1u << kPtBlockShift) - 1) >> kPtBlockShift;
  if (wide) {
    if ((rc = dev_alloc(x, &pt_hi, pt_blocks, false)) < 0) return rc;
  }
  tm.mark("finish_index: small tables, allocs");
  CU(build_prefix_table(x->d, k, pt, pt_hi, st));
  if (wide) { x->d.pt_hi = pt_hi; REG_ARR(x, pt_hi, pt_blocks * 8); }
  x->launches += 2;
  CU(cudaStreamSynchronize(st));
  tm.mark("finish_index: prefix table");
  // self-check: every suffix-array element inside the table bucket of its suffix (0.3 s at 3.1 G symbols); a failure
  // means the suffix array does not belong to this sequence or an array was damaged on its way to the device
  if (!getenv("BKX_NO_VERIFY")) {
    unsigned long long* d_bad = nullptr;
    unsigned long long n_bad = 0;
    CU(cudaMalloc((void**)&d_bad, 8));
    CU(cudaMemsetAsync(d_bad, 0, 8, st));
    CU(launch_verify_index(x->d, k, d_bad, st));
    CU(cudaMemcpyAsync(&n_bad, d_bad, 8, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    cudaFree(d_bad);
    x->launches += 1;
    g_selfcheck_failed = n_bad != 0;
    if (n_bad) {
      cudaDeviceSynchronize();
      return fail(BKX_ERR_FORMAT, "index self-check failed: %llu inconsistencies (suffix-array elements outside the bucket of their suffix, or "
                  "exception-map blocks that disagree) in an index of %llu symbols", n_bad, (unsigned long long)n);
    }
  }
  // the small tables above went up with cudaMemcpy on the legacy stream, which the work streams do not wait for
  CU(cudaDeviceSynchronize());
  tm.mark("finish_index: self-check");
  return BKX_OK;
}

static int new_index(int device, bkx_index** out) {
  int nd = 0;
  if (cudaGetDeviceCount(&nd) != cudaSuccess || nd <= 0)
    return fail(BKX_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
  if (device < 0 || device >= nd) return fail(BKX_ERR_PARAM, "device %d out of range (%d devices)", device, nd);
  CU(cudaSetDevice(device));
  bkx_index* x = new bkx_index();
  x->device = device;
  for (int s = 0; s < kSlots; ++s) {
    CU(cudaStreamCreateWithFlags(&x->slot[s].st, cudaStreamNonBlocking));
    CU(cudaEventCreate(&x->slot[s].k0));
    CU(cudaEventCreate(&x->slot[s].k1));
    CU(cudaEventCreateWithFlags(&x->slot[s].in_ready, cudaEventDisableTiming));
    CU(cudaMalloc((void**)&x->d_cursor[s], 4 * sizeof(unsigned int)));
  }
  CU(cudaStreamCreateWithFlags(&x->cst, cudaStreamNonBlocking));
  CU(cudaMalloc((void**)&x->d_stats, sizeof(bkx_align_stats)));
  CU(cudaMalloc((void**)&x->d_pe_stats, sizeof(bkx_pe_stats)));
  *out = x;
  return BKX_OK;
}

extern "C" void bkx_close_index(bkx_index* x) {
  if (!x) return;
  cudaSetDevice(x->device);
  cudaDeviceSynchronize();
  for (void* p : x->owned) cudaFree(p);
  for (int s = 0; s < kSlots; ++s) {
    if (x->fast_hash[s]) cudaFree(x->fast_hash[s]);
    if (x->slot[s].d_bases) cudaFree(x->slot[s].d_bases);
    if (x->slot[s].d_packed) cudaFree(x->slot[s].d_packed);
    if (x->slot[s].d_multi) cudaFree(x->slot[s].d_multi);
    if (x->slot[s].d_orphans) cudaFree(x->slot[s].d_orphans);
    if (x->slot[s].d_offs) cudaFree(x->slot[s].d_offs);
    if (x->slot[s].d_out) cudaFree(x->slot[s].d_out);
    if (x->slot[s].d_out16) cudaFree(x->slot[s].d_out16);
    if (x->slot[s].d_lens) cudaFree(x->slot[s].d_lens);
    if (x->slot[s].d_exc_pos) cudaFree(x->slot[s].d_exc_pos);
    if (x->slot[s].d_exc_code) cudaFree(x->slot[s].d_exc_code);
    if (x->slot[s].d_scan_tmp) cudaFree(x->slot[s].d_scan_tmp);
    if (x->slot[s].d_rflags) cudaFree(x->slot[s].d_rflags);
    if (x->slot[s].d_hard) cudaFree(x->slot[s].d_hard);
    {
      WaveBuf& B = x->slot[s].wave;
      void* wp[] = {B.cnt, B.ph, B.fb, B.acc, B.ncand, B.cand, B.items, B.fb_ids};
      for (void* q : wp) if (q) cudaFree(q);
    }
    if (x->slot[s].k0) cudaEventDestroy(x->slot[s].k0);
    if (x->slot[s].k1) cudaEventDestroy(x->slot[s].k1);
    if (x->slot[s].in_ready) cudaEventDestroy(x->slot[s].in_ready);
    if (x->slot[s].st) cudaStreamDestroy(x->slot[s].st);
    if (x->d_cursor[s]) cudaFree(x->d_cursor[s]);
  }
  if (x->cst) cudaStreamDestroy(x->cst);
  if (x->hp.tables) cudaFree(x->hp.tables);
  if (x->hp.locks) cudaFree(x->hp.locks);
  if (x->hp.epochs) cudaFree(x->hp.epochs);
  if (x->d_stats) cudaFree(x->d_stats);
  if (x->d_pe_stats) cudaFree(x->d_pe_stats);
  if (x->d_len_dist) cudaFree(x->d_len_dist);
  if (x->d_pe_list) cudaFree(x->d_pe_list);
  if (x->d_chrom_keep) cudaFree(x->d_chrom_keep);
  delete x;
}

extern "C" int bkx_open_index_dev(const uint8_t* d_seq, uint64_t concat_len, const void* d_sa, uint32_t el,
                                  const bkx_entry* entries, uint32_t n_ent, const char* name, int device, int prefix_k,
                                  bkx_index** out) {
  if (!d_seq || !d_sa || !entries || !out) return fail(BKX_ERR_PARAM, "null argument");
  bkx_index* x = nullptr;
  int rc = new_index(device, &x);
  if (rc < 0) return rc;
  x->info.version = 5;
  CU(cudaDeviceSynchronize());  // the caller's buffers may come from any stream; the library's streams are non-blocking
  SaSrc src;
  src.raw = d_sa;
  rc = finish_index(x, d_seq, concat_len, src, el, entries, n_ent, name, prefix_k);
  if (rc < 0) { bkx_close_index(x); return rc; }
  *out = x;
  return BKX_OK;
}

extern "C" int bkx_open_index_planes(const uint8_t* d_seq, uint64_t concat_len, const uint32_t* d_sa_lo,
                                     const uint8_t* d_sa_hi, const bkx_entry* entries, uint32_t n_ent, const char* name,
                                     int device, int prefix_k, bkx_index** out) {
  if (!d_seq || !d_sa_lo || !entries || !out) return fail(BKX_ERR_PARAM, "null argument");
  if (!d_sa_hi && concat_len >= 4000000000ull)
    return fail(BKX_ERR_PARAM, "%llu symbols use 5-byte suffix elements: the high plane is required", (unsigned long long)concat_len);
  bkx_index* x = nullptr;
  int rc = new_index(device, &x);
  if (rc < 0) return rc;
  x->info.version = 5;
  CU(cudaDeviceSynchronize());  // the caller's buffers may come from any stream; the library's streams are non-blocking
  SaSrc src;  // borrowed: not entered in x->owned, so bkx_close_index leaves the planes alone
  src.lo = d_sa_lo;
  src.hi = d_sa_hi;
  rc = finish_index(x, d_seq, concat_len, src, d_sa_hi ? 5 : 4, entries, n_ent, name, prefix_k);
  if (rc < 0) { bkx_close_index(x); return rc; }
  *out = x;
  return BKX_OK;
}

extern "C" int bkx_open_index_packed5(const uint8_t* d_seq, uint64_t concat_len, const uint8_t* d_sa5, const bkx_entry* entries,
                                      uint32_t n_ent, const char* name, int device, int prefix_k, bkx_index** out) {
  if (!d_seq || !d_sa5 || !entries || !out) return fail(BKX_ERR_PARAM, "null argument");
  bkx_index* x = nullptr;
  int rc = new_index(device, &x);
  if (rc < 0) return rc;
  x->info.version = 5;
  CU(cudaDeviceSynchronize());  // the caller's buffers may come from any stream; the library's streams are non-blocking
  SaSrc src;  // borrowed: not entered in x->owned, so bkx_close_index leaves the array alone
  src.packed5 = d_sa5;
  rc = finish_index(x, d_seq, concat_len, src, 5, entries, n_ent, name, prefix_k);
  if (rc < 0) { bkx_close_index(x); return rc; }
  *out = x;
  return BKX_OK;
}

extern "C" int bkx_open_index_mem(const uint8_t* seq, uint64_t concat_len, const void* sa, uint32_t el,
                                  const bkx_entry* entries, uint32_t n_ent, const char* name, int device, int prefix_k,
                                  bkx_index** out) {
  if (!seq || !sa || !entries || !out) return fail(BKX_ERR_PARAM, "null argument");
  bkx_index* x = nullptr;
  int rc = new_index(device, &x);
  if (rc < 0) return rc;
  x->info.version = 5;
  if (el != 4 && el != 5) { bkx_close_index(x); return fail(BKX_ERR_FORMAT, "unsupported suffix element size %u", el); }
  uint8_t* d_seq = nullptr;
  void* d_sa = nullptr;
  cudaError_t e = cudaMalloc((void**)&d_seq, concat_len);
  if (e == cudaSuccess) e = cudaMemcpy(d_seq, seq, concat_len, cudaMemcpyHostToDevice);
  bool keep_sa = (el == 4);
  if (e == cudaSuccess) e = cudaMalloc(&d_sa, concat_len * el);
  if (e == cudaSuccess) e = cudaMemcpy(d_sa, sa, concat_len * el, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    cudaFree(d_seq); cudaFree(d_sa); bkx_close_index(x);
    return fail(BKX_ERR_CUDA, "index upload: %s", cudaGetErrorString(e));
  }
  if (cudaDeviceSynchronize() != cudaSuccess) { /* the uploads above ran on the legacy stream */ }
  SaSrc src;
  if (keep_sa) { x->owned.push_back(d_sa); x->info.device_bytes += concat_len * 4; src.lo = (const uint32_t*)d_sa; }
  else src.raw = d_sa;
  rc = finish_index(x, d_seq, concat_len, src, el, entries, n_ent, name, prefix_k);
  cudaFree(d_seq);
  if (!keep_sa) cudaFree(d_sa);
  if (rc < 0) { bkx_close_index(x); return rc; }
  *out = x;
  return BKX_OK;
}

// ---- .sfx file: header tsSfxHeaderV3/Vv (SfxArrayV2.h:174-203), entries tsSfxEntry (:79-88),
//      block tsSfxBlock (:98-104); see Disk2Hdr / Disk2Entries (SfxArrayV2.cpp:551-747) ----------------
static bool pread_all(int fd, void* buf, size_t len, off_t ofs) {
  uint8_t* p = (uint8_t*)buf;
  while (len) {
    ssize_t n = pread(fd, p, std::min(len, (size_t)1 << 30), ofs);
    if (n <= 0) return false;
    p += n; len -= (size_t)n; ofs += n;
  }
  return true;
}
static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

static int open_index_once(const char* path, int device, int prefix_k, bkx_index** out) {
  if (!path || !out) return fail(BKX_ERR_PARAM, "null argument");
  int fd = open(path, O_RDONLY);
  if (fd < 0) return fail(BKX_ERR_FILE, "unable to open '%s'", path);
  struct stat sb;
  fstat(fd, &sb);
  uint8_t hdr[1224];
  memset(hdr, 0, sizeof(hdr));
  size_t hlen = std::min((size_t)sb.st_size, sizeof(hdr));
  if (hlen < 52 || !pread_all(fd, hdr, hlen, 0) || memcmp(hdr, "sfx", 3) != 0) {
    close(fd);
    return fail(BKX_ERR_FORMAT, "'%s' is not a biokanga suffix array file", path);
  }
  uint32_t version = rd32(hdr + 4), attributes = rd32(hdr + 8);
  if (version < 3 || version > 5 || hdr[3] != (uint8_t)('0' + version)) {
    close(fd);
    return fail(BKX_ERR_FORMAT, "'%s': unsupported sfx version %u", path, version);
  }
  if (attributes & 0x03) {
    close(fd);
    return fail(BKX_ERR_UNSUPPORTED, "'%s': bisulfite / colorspace indexes are not supported", path);
  }
  uint64_t entries_ofs = rd64(hdr + 20), blk_ofs = rd64(hdr + 44);
  uint32_t n_blocks = rd32(hdr + 32);
  int name_len = (version <= 3) ? 36 : 81;
  char dataset[88] = {0};
  memcpy(dataset, hdr + 52, (size_t)std::min(name_len, 83));
  if (n_blocks != 1 || entries_ofs == 0 || blk_ofs == 0 || entries_ofs + 8 > (uint64_t)sb.st_size) {
    close(fd);
    return fail(BKX_ERR_FORMAT, "'%s': no suffix block / entries", path);
  }
  uint8_t eh[8];
  if (!pread_all(fd, eh, 8, (off_t)entries_ofs)) { close(fd); return fail(BKX_ERR_FILE, "'%s': short read", path); }
  uint32_t n_ent = rd32(eh);
  size_t esz = (size_t)(8 + name_len + 2 + 4 + 8 + 8);
  std::vector<uint8_t> eb((size_t)n_ent * esz);
  if (n_ent == 0 || !pread_all(fd, eb.data(), eb.size(), (off_t)entries_ofs + 8)) {
    close(fd);
    return fail(BKX_ERR_FORMAT, "'%s': bad entries block", path);
  }
  std::vector<bkx_entry> ents(n_ent);
  for (uint32_t i = 0; i < n_ent; ++i) {
    const uint8_t* e = eb.data() + esz * i;
    bkx_entry& o = ents[i];
    memset(&o, 0, sizeof(o));
    o.entry_id = rd32(e);
    memcpy(o.name, e + 8, (size_t)name_len);
    o.seq_len = rd32(e + 8 + name_len + 2);
    o.start_ofs = rd64(e + 8 + name_len + 6);
    o.end_ofs = rd64(e + 8 + name_len + 14);
  }
  uint8_t bh[20];
  if (!pread_all(fd, bh, 20, (off_t)blk_ofs)) { close(fd); return fail(BKX_ERR_FILE, "'%s': short read", path); }
  uint64_t n = rd64(bh + 8);
  uint32_t el = rd32(bh + 16);
  if ((el != 4 && el != 5) || blk_ofs + 20 + n + n * el > (uint64_t)sb.st_size) {
    close(fd);
    return fail(BKX_ERR_FORMAT, "'%s': bad suffix block", path);
  }
  StageTimer tm;
  bkx_index* x = nullptr;
  int rc = new_index(device, &x);
  if (rc < 0) { close(fd); return rc; }
  tm.mark("open: CUDA context, streams");
  x->info.version = version;
  x->info.attributes = attributes;
  // Stream the block to the GPU: a few reader threads, each with its own pinned staging buffer and CUDA stream, take
  // the 60 MB chunks round-robin (the page-cache copy, not PCIe, is the slow half).  4-byte elements land in their final
  // array; 5-byte elements pass through a device staging buffer and are split into the two planes chunk by chunk, so
  // an 84 GB index never needs more than its final footprint.
  uint8_t* d_seq = nullptr;
  uint8_t* d_sa = nullptr;   // the elements as the file holds them: u32, or 5 bytes each (+ 16 bytes of slack)
  const size_t chunk = (size_t)60 << 20;  // a multiple of 5 and of 4
  cudaError_t e = cudaMalloc((void**)&d_seq, n);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_sa, n * el + 16);
  bool io_ok = true;
  tm.mark("open: device allocations");
  if (e == cudaSuccess) {
    const uint64_t seq_chunks = (n + chunk - 1) / chunk, sa_bytes = n * el, sa_chunks = (sa_bytes + chunk - 1) / chunk;
    const uint64_t n_chunks = seq_chunks + sa_chunks;
    int n_workers = (int)std::min<uint64_t>(std::max(4u, std::min(12u, std::thread::hardware_concurrency() / 2)), n_chunks);
    if (const char* ev = getenv("BKX_LOAD_THREADS")) n_workers = (int)std::min<uint64_t>((uint64_t)std::max(1, atoi(ev)), n_chunks);
    std::vector<cudaError_t> werr((size_t)n_workers, cudaSuccess);
    std::vector<char> wio((size_t)n_workers, 1);
    std::vector<uint64_t> wlaunch((size_t)n_workers, 0);
    std::vector<std::thread> workers;
    for (int w = 0; w < n_workers; ++w)
      workers.emplace_back([&, w]() {
        cudaError_t ce = cudaSetDevice(device);
        uint8_t* pin = nullptr;
        cudaStream_t st = nullptr;
        if (ce == cudaSuccess) ce = cudaMallocHost((void**)&pin, chunk);
        if (ce == cudaSuccess) ce = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
        for (uint64_t c = (uint64_t)w; c < n_chunks && ce == cudaSuccess && wio[(size_t)w]; c += (uint64_t)n_workers) {
          const bool is_seq = c < seq_chunks;
          const uint64_t ofs = is_seq ? c * chunk : (c - seq_chunks) * chunk;
          const size_t len = (size_t)std::min<uint64_t>(chunk, (is_seq ? n : sa_bytes) - ofs);
          if (!pread_all(fd, pin, len, (off_t)(blk_ofs + 20 + (is_seq ? ofs : n + ofs)))) { wio[(size_t)w] = 0; break; }
          ce = cudaMemcpyAsync((is_seq ? d_seq : d_sa) + ofs, pin, len, cudaMemcpyHostToDevice, st);
          if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);  // the pinned buffer is about to be refilled
        }
        if (pin) cudaFreeHost(pin);
        if (st) cudaStreamDestroy(st);
        werr[(size_t)w] = ce;
      });
    for (auto& t : workers) t.join();
    for (int w = 0; w < n_workers; ++w) {
      if (werr[(size_t)w] != cudaSuccess && e == cudaSuccess) e = werr[(size_t)w];
      if (!wio[(size_t)w]) io_ok = false;
      x->launches += wlaunch[(size_t)w];
    }
  }
  close(fd);
  tm.mark("open: file -> device");
  if (e != cudaSuccess || !io_ok) {
    cudaFree(d_seq); cudaFree(d_sa); bkx_close_index(x);
    return e != cudaSuccess ? fail(BKX_ERR_CUDA, "index upload: %s", cudaGetErrorString(e))
                            : fail(BKX_ERR_FILE, "'%s': short read", path);
  }
  x->owned.push_back(d_sa);
  x->info.device_bytes += n * el + 16;
  SaSrc src;
  if (el == 4) src.lo = (const uint32_t*)d_sa; else src.packed5 = d_sa;
  rc = finish_index(x, d_seq, n, src, el, ents.data(), n_ent, dataset, prefix_k);
  cudaFree(d_seq);
  if (rc < 0) { bkx_close_index(x); return rc; }
  *out = x;
  return BKX_OK;
}

// Re-run the self-check of finish_index on a live index: number of suffix-array elements outside their bucket.
extern "C" int64_t bkx_self_check(bkx_index* x) {
  if (!x) return fail(BKX_ERR_PARAM, "null argument");
  std::lock_guard<std::mutex> lk(x->mtx);
  CU(cudaSetDevice(x->device));
  CU(cudaDeviceSynchronize());
  unsigned long long* d_bad = nullptr;
  unsigned long long n_bad = 0;
  CU(cudaMalloc((void**)&d_bad, 8));
  CU(cudaMemsetAsync(d_bad, 0, 8, x->slot[0].st));
  CU(launch_verify_index(x->d, x->d.k, d_bad, x->slot[0].st));
  CU(cudaMemcpyAsync(&n_bad, d_bad, 8, cudaMemcpyDeviceToHost, x->slot[0].st));
  CU(cudaStreamSynchronize(x->slot[0].st));
  cudaFree(d_bad);
  return (int64_t)n_bad;
}

// Diagnostic hook for the open issue in DESIGN.md section 4: bring parts of an index's run-time state back to what a
// freshly opened index has.  what: 1 = overflow pool (tables, locks, epochs), 2 = lane hash sets of the fast kernel,
// 4 = launch geometry (grids, word counts; re-derived at the next call).
extern "C" int bkx_debug_reset(bkx_index* x, int what) {
  if (!x) return fail(BKX_ERR_PARAM, "null argument");
  std::lock_guard<std::mutex> lk(x->mtx);
  CU(cudaSetDevice(x->device));
  CU(cudaDeviceSynchronize());
  if ((what & 1) && x->hp.tables) {
    CU(cudaMemset(x->hp.tables, 0, (size_t)x->hp.n_tables * x->hp.slots * 8));
    CU(cudaMemset(x->hp.locks, 0, (size_t)x->hp.n_tables * 4));
    CU(cudaMemset(x->hp.epochs, 0, (size_t)x->hp.n_tables * 4));
  }
  if (what & 2) {
    for (int s = 0; s < kSlots; ++s)
      if (x->fast_hash[s]) CU(cudaMemset(x->fast_hash[s], 0, x->fast_hash_lanes * kFastHashSlots * 8));
    x->fast_epoch = 1;
  }
  if (what & 4) {
    for (int s = 0; s < kSlots; ++s)
      if (x->fast_hash[s]) { cudaFree(x->fast_hash[s]); x->fast_hash[s] = nullptr; }
    x->fast_hash_lanes = 0;
    x->fast_epoch = 1;
    x->grid = 0; x->grid_W = 0; x->fast_grid = 0; x->fast_W = 0; x->max_len_prepared = 0;
  }
  CU(cudaDeviceSynchronize());  // cudaMemset runs on the legacy stream; the work streams do not wait for it
  return BKX_OK;
}

// A failed self-check of an index that came from a file is retried from the file: if the file is sound the damage
// happened on the way to (or on) the device, and a second upload is the remedy; each retry is reported.
extern "C" int bkx_open_index(const char* path, int device, int prefix_k, bkx_index** out) {
  int rc = BKX_OK;
  for (int attempt = 0; attempt < 3; ++attempt) {
    g_selfcheck_failed = false;
    rc = open_index_once(path, device, prefix_k, out);
    if (rc >= 0 || !g_selfcheck_failed) return rc;
    fprintf(stderr, "[bkx] warning: %s -- uploading '%s' again (attempt %d of 3)\n", bkx_last_error(), path, attempt + 2);
  }
  return rc;
}

// Replicate an open index onto another GPU with peer copies (NVLink when the GPUs are peers).
extern "C" int bkx_clone_index(const bkx_index* src, int device, bkx_index** out) {
  if (!src || !out) return fail(BKX_ERR_PARAM, "null argument");
  bkx_index* x = nullptr;
  int rc = new_index(device, &x);
  if (rc < 0) return rc;
  x->info = src->info;
  x->info.device = (uint32_t)device;
  x->info.device_bytes = 0;
  x->entries = src->entries;
  x->d = src->d;
  x->arrs = src->arrs;
  int can = 0;
  cudaDeviceCanAccessPeer(&can, device, src->device);
  if (can) {
    cudaError_t pe = cudaDeviceEnablePeerAccess(src->device, 0);
    if (pe != cudaSuccess && pe != cudaErrorPeerAccessAlreadyEnabled) can = 0;
    cudaGetLastError();
  }
  for (const auto& a : src->arrs) {
    const void* sp = *(const void* const*)((const char*)&src->d + a.field_ofs);
    void* dp = nullptr;
    cudaError_t e = cudaMalloc(&dp, a.bytes ? a.bytes : 8);
    if (e == cudaSuccess) e = cudaMemcpyPeer(dp, device, sp, src->device, a.bytes);
    if (e != cudaSuccess) {
      if (dp) cudaFree(dp);
      bkx_close_index(x);
      return fail(BKX_ERR_CUDA, "index replication to device %d: %s", device, cudaGetErrorString(e));
    }
    x->owned.push_back(dp);
    x->info.device_bytes += a.bytes;
    *(const void**)((char*)&x->d + a.field_ofs) = dp;
  }
  CU(cudaDeviceSynchronize());
  CU(cudaSetDevice(src->device));   // the peer copies involve both devices' legacy streams
  CU(cudaDeviceSynchronize());
  CU(cudaSetDevice(device));
  *out = x;
  return BKX_OK;
}

extern "C" int bkx_index_info_get(const bkx_index* x, bkx_index_info* out) {
  if (!x || !out) return fail(BKX_ERR_PARAM, "null argument");
  *out = x->info;
  return BKX_OK;
}

extern "C" int bkx_get_entry(const bkx_index* x, uint32_t entry_id, bkx_entry* out) {
  if (!x || !out) return fail(BKX_ERR_PARAM, "null argument");
  for (const auto& e : x->entries)
    if (e.entry_id == entry_id) { *out = e; return BKX_OK; }
  return fail(BKX_ERR_ENTRY, "no entry %u", entry_id);
}

extern "C" int bkx_get_ident(const bkx_index* x, const char* name) {
  if (!x || !name) return fail(BKX_ERR_PARAM, "null argument");
  for (const auto& e : x->entries)
    if (strcasecmp(e.name, name) == 0) return (int)e.entry_id;
  return fail(BKX_ERR_ENTRY, "no entry named '%s'", name);
}

__global__ void unpack_seq_kernel(DevIndex I, uint64_t start, uint64_t len, uint8_t* out) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < len; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = (uint8_t)gsym(I, start + i);
}

extern "C" int64_t bkx_get_seq(const bkx_index* x, uint32_t entry_id, uint64_t loci, uint64_t len, uint8_t* buf) {
  if (!x || !buf) return fail(BKX_ERR_PARAM, "null argument");
  const bkx_entry* e = nullptr;
  for (const auto& t : x->entries) if (t.entry_id == entry_id) e = &t;
  if (!e) return fail(BKX_ERR_ENTRY, "no entry %u", entry_id);
  if (loci >= e->seq_len) return 0;
  len = std::min<uint64_t>(len, e->seq_len - loci);
  if (len == 0) return 0;
  CU(cudaSetDevice(x->device));
  uint8_t* d = nullptr;
  CU(cudaMalloc((void**)&d, len));
  unpack_seq_kernel<<<(unsigned)std::min<uint64_t>((len + 255) / 256, 148 * 8), 256>>>(x->d, e->start_ofs + loci, len, d);
  cudaError_t err = cudaMemcpy(buf, d, len, cudaMemcpyDeviceToHost);
  cudaFree(d);
  if (err != cudaSuccess) return fail(BKX_ERR_CUDA, "bkx_get_seq: %s", cudaGetErrorString(err));
  return (int64_t)len;
}

extern "C" int bkx_default_params(const bkx_index* x, int pmode, bkx_align_params* p) {
  if (!x || !p) return fail(BKX_ERR_PARAM, "null argument");
  memset(p, 0, sizeof(*p));
  uint64_t t = x->info.tot_seq_len;
  int mcl;  // Aligner.cpp:8727-8739
  if (t <= 500000) mcl = 4;
  else if (t <= 20000000) mcl = 7;
  else if (t <= 250000000) mcl = 11;
  else if (t <= 3500000000ull) mcl = 12;
  else mcl = 15;
  int slides, iters;  // Aligner.cpp:8744-8760 and :341-356
  switch (pmode) {
    case BKX_PMODE_ULTRASENS: slides = 9; iters = 20000; break;
    case BKX_PMODE_MORESENS: mcl += 1; slides = 8; iters = 10000; break;
    case BKX_PMODE_DEFAULT: mcl += 2; slides = 8; iters = 5000; break;
    case BKX_PMODE_LESSSENS: mcl += 4; slides = 6; iters = 2500; break;
    default: return fail(BKX_ERR_PARAM, "bad processing mode %d", pmode);
  }
  p->pmode = pmode;
  p->max_subs = 10;
  p->min_edit_dist = 1;
  p->max_ns = 1;
  p->align_strand = BKX_STRAND_BOTH;
  p->max_ml_matches = 1;
  p->min_core_len = mcl;
  p->max_num_slides = slides;
  p->max_iter = iters;
  p->max_ident_nodes = 1024000;
  return BKX_OK;
}

static int check_params(const bkx_align_params* p, KParams* k) {
  if (!p) return fail(BKX_ERR_PARAM, "null parameters");
  if (p->max_subs < 0 || p->max_subs > 15) return fail(BKX_ERR_PARAM, "max_subs %d out of range 0..15", p->max_subs);
  if (p->min_edit_dist < 1 || p->min_edit_dist > 2) return fail(BKX_ERR_PARAM, "min_edit_dist %d out of range 1..2", p->min_edit_dist);
  if (p->max_ns < 0 || p->max_ns > 5) return fail(BKX_ERR_PARAM, "max_ns %d out of range 0..5", p->max_ns);
  if (p->align_strand < 0 || p->align_strand > 2) return fail(BKX_ERR_PARAM, "bad align_strand %d", p->align_strand);
  if (p->ml_mode == BKX_ML_DEFAULT) {
    if (p->max_ml_matches != 1) return fail(BKX_ERR_PARAM, "max_ml_matches %d needs a multi-loci mode (ml_mode)", p->max_ml_matches);
  } else if (p->ml_mode == BKX_ML_DIST) {
    if (p->max_ml_matches < 2 || p->max_ml_matches > 500)  // cMaxMultiHits, Aligner.h:62
      return fail(BKX_ERR_PARAM, "max_ml_matches %d out of range 2..500", p->max_ml_matches);
  } else if (p->ml_mode == BKX_ML_ALL || p->ml_mode == BKX_ML_UNIQ || p->ml_mode == BKX_ML_MULTI) {
    if (p->max_ml_matches < 2 || p->max_ml_matches > 500)  // cMaxMultiHits, Aligner.h:62 (-r5 alone goes on to cMaxAllHits = 100000
      // in the reference, which writes loci out as it finds them; here every read owns max_ml_matches slots of 12 bytes)
      return fail(BKX_ERR_UNSUPPORTED, "max_ml_matches %d with -r3..5: 2..500 loci per read are supported", p->max_ml_matches);
  } else {
    return fail(BKX_ERR_UNSUPPORTED, "ml_mode %d: -r2 picks a locus with libc rand() in the reference and is not reproducible", p->ml_mode);
  }
  if (p->min_core_len < 4 || p->min_core_len > 100) return fail(BKX_ERR_PARAM, "bad min_core_len %d", p->min_core_len);
  if (p->max_num_slides < 1 || p->max_num_slides > 16) return fail(BKX_ERR_PARAM, "bad max_num_slides %d", p->max_num_slides);
  if (p->max_iter <= 100) return fail(BKX_ERR_PARAM, "max_iter %d must exceed 100", p->max_iter);
  if (p->max_ident_nodes < 1) return fail(BKX_ERR_PARAM, "bad max_ident_nodes %d", p->max_ident_nodes);
  k->max_subs = p->max_subs; k->mmd = p->min_edit_dist; k->max_ns = p->max_ns; k->strand_mode = p->align_strand;
  k->max_hits = p->max_ml_matches; k->min_core_len = p->min_core_len; k->slides_per100 = p->max_num_slides;
  k->max_iter = p->max_iter; k->max_nodes = p->max_ident_nodes;
  k->ml_mode = p->ml_mode; k->clamp_ml = p->clamp_max_ml ? 1 : 0;
  k->best = (p->best_matches && p->ml_mode != BKX_ML_DEFAULT) ? 1 : 0;   // kanga.cpp:666, 686: only with a multi-loci mode
  if (k->best) k->clamp_ml = 1;                                           // kanga.cpp:695-696
  k->multi = nullptr;
  k->xdedup = 0;   // measured: faster in some sweep cells (L=100 -s6: +11 %), slower in others (L=150 -s8: -13 %) and at configs[1] (-7 %)
  if (const char* ev = getenv("BKX_XDEDUP")) k->xdedup = atoi(ev) != 0;   // tuning hook
  k->prefetch = 0;   // measured: 31.5 ms without, 33.9 ms with (configs[1])
  if (const char* ev = getenv("BKX_PREFETCH")) k->prefetch = atoi(ev) != 0;   // tuning hook
  k->scan_iters = 0;
  return BKX_OK;
}

// Slots an overflow table needs so that one strand of one phase fits at a load factor <= 1/2: at most
// min(cMaxNumIdentNodes, slides x MaxIter) new candidates (both depend on the run's parameters, not only on the reads).
static uint32_t pool_slots_needed(const KParams& k, uint64_t max_len) {
  int slides = std::max(1, (int)((k.slides_per100 * std::max<uint64_t>(max_len, 100) + 99) / 100));
  uint64_t cap = std::min<uint64_t>((uint64_t)k.max_nodes, (uint64_t)slides * (uint64_t)k.max_iter);
  uint32_t slots = 1024;
  while (slots < 2 * cap) slots <<= 1;
  return slots;
}

// Size the persistent grid and the per-warp overflow hash sets for reads up to max_len bases.
static int prepare_launch(bkx_index* x, const KParams& k, uint32_t max_len, int* W_out) {
  StageTimer tm;
  if (max_len > 2000) return fail(BKX_ERR_PARAM, "read length %u exceeds cMaxSeqLen 2000", max_len);
  int W = (int)((max_len + 31) / 32) + 1;
  if (W < 3) W = 3;
  if (W > x->grid_W || x->grid == 0) {
    int nb = align_blocks_per_sm(W);
    if (nb < 1) return fail(BKX_ERR_CUDA, "align kernel does not fit on an SM (W=%d)", W);
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, x->device));
    x->grid = nb * sms;
    x->grid_W = W;
  }
  const uint32_t slots = pool_slots_needed(k, max_len);
  if (slots > x->hp.slots || !x->hp.tables) {
    // a pool of overflow tables shared by all groups (only reads in high-copy repeats borrow one)
    if (x->hp.tables) { cudaFree(x->hp.tables); x->hp.tables = nullptr; }
    if (x->hp.locks) { cudaFree(x->hp.locks); x->hp.locks = nullptr; }
    if (x->hp.epochs) { cudaFree(x->hp.epochs); x->hp.epochs = nullptr; }
    uint32_t n_tables = 2048;
    while (n_tables > 64 && (size_t)n_tables * slots * 8 > ((size_t)4 << 30)) n_tables >>= 1;
    CU(cudaMalloc((void**)&x->hp.tables, (size_t)n_tables * slots * 8));
    CU(cudaMemset(x->hp.tables, 0, (size_t)n_tables * slots * 8));
    CU(cudaMalloc((void**)&x->hp.locks, (size_t)n_tables * 4));
    CU(cudaMemset(x->hp.locks, 0, (size_t)n_tables * 4));
    CU(cudaMalloc((void**)&x->hp.epochs, (size_t)n_tables * 4));
    CU(cudaMemset(x->hp.epochs, 0, (size_t)n_tables * 4));
    // cudaMemset on device memory is ASYNCHRONOUS (legacy default stream) and the library's streams are non-blocking,
    // i.e. they do not wait for that stream: without this barrier the first kernels could acquire tables while the
    // 2 GiB clear -- and the clear of the epoch counters behind it -- was still on its way, after which epochs repeated
    // and candidates looked "already seen" (the intermittent lost candidates of round 1)
    CU(cudaDeviceSynchronize());
    x->hp.n_tables = n_tables;
    x->hp.slots = slots;
  }
  {
    int Wf = std::min(W, kFastMaxLen / 32 + 1);
    if (Wf > x->fast_W || x->fast_grid == 0) {
      int nb = fast_blocks_per_sm(Wf);
      if (nb < 1) return fail(BKX_ERR_CUDA, "fast align kernel does not fit on an SM (W=%d)", Wf);
      int sms = 0;
      CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, x->device));
      x->fast_grid = nb * sms;
      x->fast_W = Wf;
    }
    size_t lanes = (size_t)x->fast_grid * kFastThreads;
    if (lanes > x->fast_hash_lanes) {  // grid grew: drop the tables, launch_both re-creates them per slot
      CU(cudaDeviceSynchronize());
      for (int s = 0; s < kSlots; ++s)
        if (x->fast_hash[s]) { cudaFree(x->fast_hash[s]); x->fast_hash[s] = nullptr; }
      x->fast_hash_lanes = lanes;
      x->fast_epoch = 1;
    }
  }
  *W_out = x->grid_W;
  tm.mark("prepare_launch");
  return BKX_OK;
}

// fast kernel over all reads, then the general kernel over the reads it deferred (same stream)
// Scratch of the wave path for n reads in slot si; false (and no error) when the device has no room for it
static bool ensure_wave(bkx_index* x, int si, uint32_t n, cudaStream_t st) {
  Slot& s = x->slot[si];
  static const int row = [] { const char* e = getenv("BKX_WAVE_ROW"); int r = e ? atoi(e) : 8; return std::max(2, std::min(r, 32)); }();
  if (n <= s.wave_cap && s.wave.cnt) return true;
  cudaStreamSynchronize(st);
  cudaStreamSynchronize(x->cst);
  WaveBuf& B = s.wave;
  void* ptrs[] = {B.cnt, B.ph, B.fb, B.acc, B.ncand, B.cand, B.items, B.fb_ids};
  for (void* q : ptrs) if (q) cudaFree(q);
  B = WaveBuf();
  s.wave_cap = 0;
  const size_t cap = (size_t)n * 5 / 4 + 1024;
  const size_t item_cap = ((cap * 3 + (1u << 20)) / kWaveChunk) * kWaveChunk;
  const size_t need = cap * (1 + 1 + 8 + 4 + (size_t)row * 8 + 4) + item_cap * 32 + kWaveCounters * 4;
  size_t free_b = 0, total_b = 0;
  if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess || free_b < need + ((size_t)2 << 30)) return false;
  bool ok = cudaMalloc((void**)&B.cnt, kWaveCounters * 4) == cudaSuccess && cudaMalloc((void**)&B.ph, cap) == cudaSuccess &&
            cudaMalloc((void**)&B.fb, cap) == cudaSuccess && cudaMalloc((void**)&B.acc, cap * 8) == cudaSuccess &&
            cudaMalloc((void**)&B.ncand, cap * 4) == cudaSuccess && cudaMalloc((void**)&B.cand, cap * (size_t)row * 8) == cudaSuccess &&
            cudaMalloc((void**)&B.items, item_cap * 32) == cudaSuccess && cudaMalloc((void**)&B.fb_ids, cap * 4) == cudaSuccess;
  if (!ok) {
    cudaGetLastError();
    void* got[] = {B.cnt, B.ph, B.fb, B.acc, B.ncand, B.cand, B.items, B.fb_ids};
    for (void* q : got) if (q) cudaFree(q);
    B = WaveBuf();
    return false;
  }
  B.item_cap = item_cap;
  B.row = row;
  B.sa_split = getenv("BKX_WAVE_SA_SPLIT") ? 1 : 0;   // tuning hook
  B.half_loads = 1;
  s.wave_cap = cap;
  return true;
}

static int launch_both(bkx_index* x, const KParams& k, const uint8_t* d_bases, const uint64_t* d_offs, uint32_t n,
                       int W, bkx_read_result* d_out, bkx_align_stats* d_stats, int si, uint32_t* d_hard,
                       cudaStream_t st, const Packed2Src& p2 = Packed2Src(), uint32_t max_len = 0) {
  unsigned int* cur = x->d_cursor[si];
  if (getenv("BKX_NO_FAST") || k.best) {   // -N: a different search (LocateBestMatches), built in the general kernel only
    CU(launch_align(x->d, k, d_bases, d_offs, n, W, d_out, d_stats, cur + 1, x->hp, nullptr, nullptr, x->grid, st));
    x->launches += 1;
    return BKX_OK;
  }
  // every launch gets its own 2^20-wide epoch range for the lane hash sets; wipe them when the 32-bit tag wraps
  if (!x->fast_hash[si]) {
    CU(cudaMalloc((void**)&x->fast_hash[si], x->fast_hash_lanes * kFastHashSlots * 8));
    CU(cudaMemsetAsync(x->fast_hash[si], 0, x->fast_hash_lanes * kFastHashSlots * 8, st));
  }
  if (x->fast_epoch > 0xfff00000u) {
    CU(cudaDeviceSynchronize());
    for (int s = 0; s < kSlots; ++s)
      if (x->fast_hash[s]) CU(cudaMemset(x->fast_hash[s], 0, x->fast_hash_lanes * kFastHashSlots * 8));
    CU(cudaDeviceSynchronize());
    x->fast_epoch = 1;
  }
  uint32_t epoch_base = x->fast_epoch;
  x->fast_epoch += 1u << 20;
  static const bool trace = getenv("BKX_TRACE") != nullptr;  // diagnostic: split of the two kernels, serialising
  cudaEvent_t t0 = nullptr, t1 = nullptr, t2 = nullptr;
  if (trace) {
    cudaEventCreate(&t0); cudaEventCreate(&t1); cudaEventCreate(&t2);
    cudaEventRecord(t0, st);
  }
  static const bool no_direct = getenv("BKX_NO_DIRECT2") != nullptr;   // tuning hook: ignore the 2-bit copy of the reads
  KParams kf = k;
  if (x->d.n > (1ull << 32)) kf.xdedup = 0;   // 32-bit dedup keys collide beyond 2^32 symbols: keep the reference's key set there
  // The default search goes down the wave path first (bkx_wave.cuh) when the reads came 2-bit packed and the launch is large:
  // its ~20 kernels per launch take 27.5 ms per 20 M reads against 29.5-30.4 ms for the lane-per-read kernel (configs[1]),
  // but 3.7 ms against 2.8 ms per 2 M reads.  The lane-per-read kernel then only redoes what that path hands on.
  // BKX_WAVE=0 / 1: never / always (read per launch: profiles/ab_kernel.py switches it inside one process).
  const char* wave_env = getenv("BKX_WAVE");
  uint32_t wave_min = kWaveAutoReads;
  if (const char* ev = getenv("BKX_WAVE_MIN_READS")) wave_min = (uint32_t)std::max(1, atoi(ev));   // tuning hook
  const int wave_mode = wave_env ? atoi(wave_env) : (n >= wave_min ? 1 : 0);
  const uint32_t* fast_ids = nullptr;
  const unsigned int* fast_n = nullptr;
  cudaEvent_t tw = nullptr;
  if (wave_mode && p2.words && !no_direct && k.ml_mode == BKX_ML_DEFAULT && !k.clamp_ml && k.max_hits == 1 && max_len > 0 &&
      x->d.n < (1ull << 40) && ensure_wave(x, si, n, st)) {
    int sms = 0;
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, x->device));
    if (const char* hv = getenv("BKX_WAVE_HALF")) x->slot[si].wave.half_loads = atoi(hv) != 0;   // tuning hook, read per launch
    CU(launch_wave(x->d, k, d_offs, n, max_len, p2, x->slot[si].wave, d_out, d_stats, sms, st));
    fast_ids = x->slot[si].wave.fb_ids;
    fast_n = x->slot[si].wave.cnt + kWaveCntFallback;
    x->launches += (uint64_t)wave_launches(k, max_len, x->slot[si].wave.sa_split != 0);
    if (trace) { cudaEventCreate(&tw); cudaEventRecord(tw, st); }
  }
  CU(launch_align_fast(x->d, kf, d_bases, d_offs, n, x->fast_W, d_out, d_stats, cur, d_hard, cur + 2, x->fast_hash[si], epoch_base,
                       x->fast_grid, st, no_direct ? Packed2Src() : p2, fast_ids, fast_n));
  if (trace) cudaEventRecord(t1, st);
  CU(launch_align(x->d, k, d_bases, d_offs, n, W, d_out, d_stats, cur + 1, x->hp, d_hard, cur + 2, x->grid, st));
  if (trace) {
    cudaEventRecord(t2, st);
    cudaStreamSynchronize(st);
    unsigned int nh = 0;
    cudaMemcpy(&nh, cur + 2, 4, cudaMemcpyDeviceToHost);
    float a = 0.f, b = 0.f;
    cudaEventElapsedTime(&a, t0, t1);
    cudaEventElapsedTime(&b, t1, t2);
    if (tw) {
      unsigned int nfb = 0;
      float wv = 0.f;
      cudaMemcpy(&nfb, fast_n, 4, cudaMemcpyDeviceToHost);
      cudaEventElapsedTime(&wv, t0, tw);
      cudaEventElapsedTime(&a, tw, t1);
      fprintf(stderr, "[bkx trace] reads %u: wave %.3f ms, handed on %u (%.2f%%)\n", n, wv, nfb, 100.0 * nfb / n);
      cudaEventDestroy(tw);
    }
    fprintf(stderr, "[bkx trace] reads %u: fast %.3f ms, deferred %u (%.2f%%), general %.3f ms\n", n, a, nh, 100.0 * nh / n, b);
    cudaEventDestroy(t0); cudaEventDestroy(t1); cudaEventDestroy(t2);
  }
  x->launches += 2;
  return BKX_OK;
}

extern "C" int bkx_align_reads_device(bkx_index* x, const bkx_align_params* p, const uint8_t* d_bases,
                                      const uint64_t* d_offsets, uint32_t n_reads, uint32_t max_read_len,
                                      bkx_read_result* d_out, bkx_align_stats* d_stats, void* cuda_stream) {
  if (!x || !d_bases || !d_offsets || !d_out) return fail(BKX_ERR_PARAM, "null argument");
  KParams k;
  int rc = check_params(p, &k);
  if (rc < 0) return rc;
  if (p->ml_mode >= BKX_ML_UNIQ) return fail(BKX_ERR_UNSUPPORTED, "-r3..5 need the host call bkx_align_reads_multi");
  if (n_reads == 0) return BKX_OK;
  std::lock_guard<std::mutex> lk(x->mtx);
  CU(cudaSetDevice(x->device));
  int W = 0;
  if ((rc = prepare_launch(x, k, max_read_len, &W)) < 0) return rc;
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : x->slot[0].st;
  Slot& s = x->slot[0];
  if (n_reads > s.hard_cap) {
    CU(cudaStreamSynchronize(st));
    if (s.d_hard) cudaFree(s.d_hard);
    s.hard_cap = (size_t)n_reads * 5 / 4;
    CU(cudaMalloc((void**)&s.d_hard, s.hard_cap * 4));
  }
  CU(cudaEventRecord(s.k0, st));
  if ((rc = launch_both(x, k, d_bases, d_offsets, n_reads, W, d_out, d_stats, 0, s.d_hard, st)) < 0) return rc;
  CU(cudaEventRecord(s.k1, st));
  s.timed = true;
  for (int si = 1; si < kSlots; ++si) x->slot[si].timed = false;
  x->last_ms = -2.f;  // resolved lazily by bkx_last_kernel_ms
  return BKX_OK;
}

// Device-resident reads with a 2-bit copy beside the one-byte-per-base layout (see include/bkx.h)
extern "C" int bkx_align_reads_device_packed2(bkx_index* x, const bkx_align_params* p, const uint8_t* d_bases,
                                              const uint64_t* d_packed2, const uint8_t* d_read_flags, const uint64_t* d_offsets,
                                              uint32_t n_reads, uint32_t max_read_len, bkx_read_result* d_out,
                                              bkx_align_stats* d_stats, void* cuda_stream) {
  if (!x || !d_bases || !d_packed2 || !d_offsets || !d_out) return fail(BKX_ERR_PARAM, "null argument");
  if ((uintptr_t)d_packed2 & 7) return fail(BKX_ERR_PARAM, "the 2-bit stream must start on an 8-byte boundary");
  KParams k;
  int rc = check_params(p, &k);
  if (rc < 0) return rc;
  if (p->ml_mode >= BKX_ML_UNIQ) return fail(BKX_ERR_UNSUPPORTED, "-r3..5 need the host call bkx_align_reads_multi");
  if (n_reads == 0) return BKX_OK;
  std::lock_guard<std::mutex> lk(x->mtx);
  CU(cudaSetDevice(x->device));
  int W = 0;
  if ((rc = prepare_launch(x, k, max_read_len, &W)) < 0) return rc;
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : x->slot[0].st;
  Slot& s = x->slot[0];
  if (n_reads > s.hard_cap) {
    CU(cudaStreamSynchronize(st));
    if (s.d_hard) cudaFree(s.d_hard);
    s.hard_cap = (size_t)n_reads * 5 / 4;
    CU(cudaMalloc((void**)&s.d_hard, s.hard_cap * 4));
  }
  Packed2Src psrc;
  psrc.words = d_packed2; psrc.flags = d_read_flags; psrc.phase = 0;
  CU(cudaEventRecord(s.k0, st));
  if ((rc = launch_both(x, k, d_bases, d_offsets, n_reads, W, d_out, d_stats, 0, s.d_hard, st, psrc, max_read_len)) < 0) return rc;
  CU(cudaEventRecord(s.k1, st));
  s.timed = true;
  for (int si = 1; si < kSlots; ++si) x->slot[si].timed = false;
  x->last_ms = -2.f;
  return BKX_OK;
}

struct PeCall {   // the pairing half of a fused align + pair call
  const bkx_pe_params* pe = nullptr;
  bkx_pe_stats* stats = nullptr;
  uint32_t* len_dist = nullptr;
};

// The reads of a host-buffer call, in one of the three layouts the ABI takes
struct HostReads {
  enum { BYTES = 0, PACKED4 = 1, PACKED2 = 2 };
  int fmt = BYTES;
  const uint8_t* data = nullptr;      // one byte per base / two bases per byte / four bases per byte
  const uint64_t* offsets = nullptr;  // BYTES, PACKED4: start of every read, n_reads + 1 entries
  const uint16_t* lens = nullptr;     // PACKED2: read lengths, or NULL with
  uint32_t fixed_len = 0;             //          every read this long
  const uint64_t* exc_pos = nullptr;  // PACKED2: bases that are not A C G T
  const uint8_t* exc_code = nullptr;
  uint64_t n_exc = 0;
  uint64_t first_base = 0;            // PACKED2: stream position of read 0
};

static int align_host(bkx_index* x, const bkx_align_params* p, const HostReads& hr, uint32_t n_reads, void* out_any,
                      bool compact_out, bkx_align_stats* stats, bkx_multi_hit* multi = nullptr, const PeCall* pec = nullptr) {
  if (!x || !hr.data || !out_any) return fail(BKX_ERR_PARAM, "null argument");
  const bool p2 = hr.fmt == HostReads::PACKED2, packed4 = hr.fmt == HostReads::PACKED4;
  if (!p2 && !hr.offsets) return fail(BKX_ERR_PARAM, "null argument");
  if (p2 && !hr.lens && (hr.fixed_len < 1 || hr.fixed_len > 2000)) return fail(BKX_ERR_PARAM, "fixed read length %u out of range 1..2000", hr.fixed_len);
  if (p2 && hr.n_exc && (!hr.exc_pos || !hr.exc_code)) return fail(BKX_ERR_PARAM, "null exception list");
  const uint64_t* offsets = hr.offsets;
  bkx_read_result* out = compact_out ? nullptr : (bkx_read_result*)out_any;
  bkx_read_result16* out16 = compact_out ? (bkx_read_result16*)out_any : nullptr;
  KParams k;
  int rc = check_params(p, &k);
  if (rc < 0) return rc;
  bool rescue = false;
  if (pec) {
    const bkx_pe_params* pe = pec->pe;
    if (!pe) return fail(BKX_ERR_PARAM, "null argument");
    if (pe->pe_proc < BKX_PE_ORPHAN || pe->pe_proc > BKX_PE_UNIQUE_SE) return fail(BKX_ERR_PARAM, "bad pe_proc %d", pe->pe_proc);
    if (pe->pair_min_len < 25 || pe->pair_max_len > 100000 || pe->pair_min_len > pe->pair_max_len)
      return fail(BKX_ERR_PARAM, "bad insert size range %d..%d", pe->pair_min_len, pe->pair_max_len);
    if (n_reads & 1) return fail(BKX_ERR_PARAM, "paired reads come in twos");
    if (p->ml_mode != BKX_ML_DEFAULT) return fail(BKX_ERR_PARAM, "multi-loci modes do not combine with paired ends (kanga.cpp:535)");
    rescue = pe->pe_proc == BKX_PE_ORPHAN || pe->pe_proc == BKX_PE_ORPHAN_SE;
  }
  if ((p->ml_mode >= BKX_ML_UNIQ) != (multi != nullptr))
    return fail(BKX_ERR_PARAM, "ml_mode -r3..5 and bkx_align_reads_multi go together");
  if (n_reads == 0) return BKX_OK;
  std::lock_guard<std::mutex> lk(x->mtx);
  CU(cudaSetDevice(x->device));
  int W = 0;
  CU(cudaMemsetAsync(x->d_stats, 0, sizeof(bkx_align_stats), x->slot[0].st));
  if (pec) {
    if (!x->d_len_dist) CU(cudaMalloc((void**)&x->d_len_dist, 100001 * 4));
    CU(cudaMemsetAsync(x->d_len_dist, 0, 100001 * 4, x->slot[0].st));
    CU(cudaMemsetAsync(x->d_pe_stats, 0, sizeof(bkx_pe_stats), x->slot[0].st));
  }
  CU(cudaStreamSynchronize(x->slot[0].st));
  // Slice sizes: start small (the first H2D is exposed), grow to kMaxSlice (every slice pays ~0.3 ms of persistent-
  // kernel ramp-up and tail; much larger slices stall on their own H2D), shrink towards the end (the last D2H is
  // exposed).  Measured on configs[1], 20 M x 150 bp packed: 0.5 M..1.5 M -> 42 ms, fixed 1 M -> 43 ms, 0.125 M..4 M -> 46.5 ms.
  uint32_t kMinSlice = 1u << 19, kMaxSlice = 3u << 19;
  if (const char* ev = getenv("BKX_SLICE_READS")) kMinSlice = kMaxSlice = (uint32_t)std::max(1024, atoi(ev));  // tuning hooks
  if (const char* ev = getenv("BKX_SLICE_MIN")) kMinSlice = (uint32_t)std::max(1024, atoi(ev));
  if (const char* ev = getenv("BKX_SLICE_MAX")) kMaxSlice = (uint32_t)std::max((int)kMinSlice, atoi(ev));
  const uint64_t kBatchBases = 1024ull << 20;
  uint32_t ramp = kMinSlice;
  float ms_total = 0.f;
  uint32_t start = 0;
  uint64_t base_pos = hr.first_base;   // PACKED2: position of read `start` in the stream
  uint64_t exc_next = 0;   // PACKED2: first exception at or after base_pos
  int b = 0;
  bool inflight[kSlots] = {};
  // diagnostic (BKX_TIMELINE=1): device timestamps of every pipeline stage of the first slices
  static const bool timeline = getenv("BKX_TIMELINE") != nullptr;
  std::vector<cudaEvent_t> tl;
  cudaEvent_t tl0 = nullptr;
  auto mark = [&](cudaStream_t st) {
    if (!timeline || tl.size() >= 60) return;
    cudaEvent_t e; cudaEventCreate(&e); cudaEventRecord(e, st); tl.push_back(e);
  };
  if (timeline) { cudaEventCreate(&tl0); cudaEventRecord(tl0, x->slot[0].st); }
  auto drain = [&](int si) -> int {
    Slot& s = x->slot[si];
    CU(cudaStreamSynchronize(s.st));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, s.k0, s.k1));
    ms_total += ms;
    inflight[si] = false;
    return BKX_OK;
  };
  // bases, longest read and soundness of reads [from, from + cnt)
  auto measure = [&](uint32_t from, uint32_t cnt, uint64_t& nb, uint64_t& max_len) -> bool {
    if (p2 && !hr.lens) { nb = (uint64_t)cnt * hr.fixed_len; max_len = hr.fixed_len; return true; }
    nb = 0; max_len = 0;
    bool mono = true;
    if (p2) {
      for (uint32_t i = from; i < from + cnt; ++i) { const uint64_t l = hr.lens[i]; nb += l; max_len = l > max_len ? l : max_len; }
    } else {
      for (uint32_t i = from; i < from + cnt; ++i) {
        const uint64_t a = offsets[i], z = offsets[i + 1];
        mono &= (z >= a);
        const uint64_t l = z - a;
        max_len = l > max_len ? l : max_len;
      }
      nb = offsets[from + cnt] - offsets[from];
    }
    return mono;
  };
  while (start < n_reads) {
    const uint32_t left = n_reads - start;
    uint32_t cnt = std::min(ramp, std::max(kMinSlice, left / 2));
    if (cnt > left || left - cnt < kMinSlice / 2) cnt = left;
    ramp = std::min<uint64_t>(kMaxSlice, (uint64_t)ramp * 2);
    if (multi) cnt = std::min<uint32_t>(cnt, std::max<uint32_t>(1024u, (1u << 27) / (uint32_t)k.max_hits));  // <= 1.6 GB of loci slots per slice
    uint64_t nb = 0, max_len = 0;
    // longest read of this slice (overlaps with the GPU work of the previous slices)
    bool mono = measure(start, cnt, nb, max_len);
    while (cnt > 1 && mono && nb > kBatchBases) { cnt = (cnt + 1) / 2; mono = measure(start, cnt, nb, max_len); }
    if (pec && (cnt & 1)) { cnt += (cnt < left) ? 1 : 0; mono = measure(start, cnt, nb, max_len); }  // PE1 / PE2 of a pair stay in one slice
    if (!mono) return fail(BKX_ERR_PARAM, "offsets not monotonic in reads %u..%u", start, start + cnt);
    const uint64_t o0 = p2 ? base_pos : offsets[start];
    if ((int)((max_len + 31) / 32) + 1 > x->grid_W || x->grid == 0 || x->hp.tables == nullptr ||
        max_len > x->max_len_prepared || pool_slots_needed(k, max_len) > x->hp.slots) {
      // (re)sizing the grid / overflow pool: let the slices in flight finish first
      for (int si = 0; si < kSlots; ++si)
        if (inflight[si] && (rc = drain(si)) < 0) return rc;
      if ((rc = prepare_launch(x, k, (uint32_t)std::min<uint64_t>(max_len, 0xffffffffu), &W)) < 0) return rc;
      x->max_len_prepared = std::max<uint64_t>(x->max_len_prepared, max_len);
    }
    W = x->grid_W;
    Slot& s = x->slot[b];
    if (inflight[b] && (rc = drain(b)) < 0) return rc;
    if (nb + 128 > s.bases_cap) {
      if (s.d_bases) cudaFree(s.d_bases);
      s.bases_cap = (size_t)(nb + 128) * 5 / 4;
      CU(cudaMalloc((void**)&s.d_bases, s.bases_cap));
    }
    if (cnt > s.reads_cap) {
      if (s.d_offs) cudaFree(s.d_offs);
      if (s.d_out) cudaFree(s.d_out);
      if (s.d_out16) cudaFree(s.d_out16);
      if (s.d_lens) cudaFree(s.d_lens);
      if (s.d_rflags) cudaFree(s.d_rflags);
      s.reads_cap = (size_t)cnt * 5 / 4;
      CU(cudaMalloc((void**)&s.d_rflags, s.reads_cap + 8));
      CU(cudaMalloc((void**)&s.d_offs, (s.reads_cap + 1) * 8));
      CU(cudaMalloc((void**)&s.d_out, s.reads_cap * sizeof(bkx_read_result)));
      CU(cudaMalloc((void**)&s.d_out16, s.reads_cap * sizeof(bkx_read_result16)));
      CU(cudaMalloc((void**)&s.d_lens, (s.reads_cap + 8) * 2));
      if (s.d_scan_tmp) { cudaFree(s.d_scan_tmp); s.d_scan_tmp = nullptr; }
      size_t tb = 0;
      CU(launch_len_offsets(s.d_lens, (uint32_t)s.reads_cap, s.d_offs, nullptr, &tb, s.st));
      s.scan_tmp_bytes = tb + 256;
      CU(cudaMalloc(&s.d_scan_tmp, s.scan_tmp_bytes));
    }
    if (cnt > s.hard_cap) {
      if (s.d_hard) cudaFree(s.d_hard);
      s.hard_cap = (size_t)cnt * 5 / 4;
      CU(cudaMalloc((void**)&s.d_hard, s.hard_cap * 4));
    }
    uint32_t n_exc_slice = 0;
    if (hr.fmt == HostReads::BYTES) {
      CU(cudaMemcpyAsync(s.d_bases, hr.data + o0, nb, cudaMemcpyHostToDevice, s.st));
    } else {
      // a half / a quarter of the PCIe bytes: ship the packed codes, expand to one byte per base on the device (an
      // HBM-speed pass)
      const int per = packed4 ? 2 : 4, sh = packed4 ? 1 : 2;
      const uint64_t o1 = o0 + nb;
      const uint64_t byte0 = o0 >> sh, nbytes = ((o1 + per - 1) >> sh) - byte0;
      if (nbytes + 128 > s.packed_cap) {
        if (s.d_packed) cudaFree(s.d_packed);
        s.packed_cap = (size_t)(nbytes + 128) * 5 / 4;
        CU(cudaMalloc((void**)&s.d_packed, s.packed_cap));
      }
      mark(s.st);
      if (nbytes) CU(cudaMemcpyAsync(s.d_packed, hr.data + byte0, nbytes, cudaMemcpyHostToDevice, s.st));
      mark(s.st);
      if (p2) {  // the slice's share of the exception list (positions ascend)
        uint64_t e1 = exc_next;
        while (e1 < hr.n_exc && hr.exc_pos[e1] < o1) ++e1;
        if (e1 - exc_next > 0x7fffffffu) return fail(BKX_ERR_PARAM, "too many non-ACGT bases in one slice");
        n_exc_slice = (uint32_t)(e1 - exc_next);
        if (n_exc_slice > s.exc_cap) {
          if (s.d_exc_pos) cudaFree(s.d_exc_pos);
          if (s.d_exc_code) cudaFree(s.d_exc_code);
          s.exc_cap = (size_t)n_exc_slice * 5 / 4 + 1024;
          CU(cudaMalloc((void**)&s.d_exc_pos, s.exc_cap * 8));
          CU(cudaMalloc((void**)&s.d_exc_code, s.exc_cap));
        }
        if (n_exc_slice) {
          CU(cudaMemcpyAsync(s.d_exc_pos, hr.exc_pos + exc_next, (size_t)n_exc_slice * 8, cudaMemcpyHostToDevice, s.st));
          CU(cudaMemcpyAsync(s.d_exc_code, hr.exc_code + exc_next, (size_t)n_exc_slice, cudaMemcpyHostToDevice, s.st));
        }
        exc_next = e1;
        if (hr.lens) CU(cudaMemcpyAsync(s.d_lens, hr.lens + start, (size_t)cnt * 2, cudaMemcpyHostToDevice, s.st));
      }
    }
    if (!p2) CU(cudaMemcpyAsync(s.d_offs, offsets + start, ((size_t)cnt + 1) * 8, cudaMemcpyHostToDevice, s.st));
    CU(cudaEventRecord(s.in_ready, s.st));
    CU(cudaStreamWaitEvent(x->cst, s.in_ready, 0));
    if (packed4 && nb) {
      CU(launch_unpack4(s.d_packed, (unsigned)(o0 & 1), nb, s.d_bases, x->cst));
      x->launches += 1;
    }
    if (p2) {
      if (nb) CU(launch_unpack2(s.d_packed, (unsigned)(o0 & 3), nb, s.d_bases, x->cst));
      if (n_exc_slice) CU(launch_scatter_exceptions(s.d_exc_pos, s.d_exc_code, n_exc_slice, o0, nb, s.d_bases, x->cst));
      if (hr.lens) CU(launch_len_offsets(s.d_lens, cnt, s.d_offs, s.d_scan_tmp, &s.scan_tmp_bytes, x->cst));
      else CU(launch_fixed_offsets(s.d_offs, cnt, hr.fixed_len, x->cst));
      CU(launch_flag_exception_reads(s.d_exc_pos, n_exc_slice, o0, s.d_offs, cnt, s.d_rflags, x->cst));
      x->launches += 2 + (n_exc_slice ? 2 : 0);
    }
    if (multi) {
      const size_t need = (size_t)cnt * (size_t)k.max_hits;
      if (need > s.multi_cap) {
        if (s.d_multi) cudaFree(s.d_multi);
        s.multi_cap = need * 5 / 4;
        CU(cudaMalloc((void**)&s.d_multi, s.multi_cap * sizeof(bkx_multi_hit)));
      }
      CU(cudaMemsetAsync(s.d_multi, 0, need * sizeof(bkx_multi_hit), x->cst));
      k.multi = s.d_multi;
    }
    CU(cudaEventRecord(s.k0, x->cst));
    // BYTES / PACKED4: the offsets stay absolute, so the kernels get a base pointer shifted by the slice start;
    // PACKED2: the device made offsets relative to the slice
    const uint8_t* kbases = p2 ? s.d_bases : s.d_bases - o0;
    Packed2Src psrc;   // PACKED2: the fast kernel takes its reads straight from the 2-bit stream that came over PCIe
    if (p2) { psrc.words = (const uint64_t*)s.d_packed; psrc.flags = s.d_rflags; psrc.phase = (uint32_t)(o0 & 3); }
    if ((rc = launch_both(x, k, kbases, s.d_offs, cnt, W, s.d_out, x->d_stats, b, s.d_hard, x->cst, psrc, (uint32_t)max_len)) < 0) return rc;
    if (pec) {  // pair the slice's reads while they are still on the device (ProcessPairedEnds, Aligner.cpp:2876-3049)
      unsigned int* cur = x->d_cursor[b];
      if ((size_t)cnt / 2 > s.orphans_cap) {
        if (s.d_orphans) cudaFree(s.d_orphans);
        s.orphans_cap = (size_t)cnt / 2 * 5 / 4 + 16;
        CU(cudaMalloc((void**)&s.d_orphans, s.orphans_cap * 4));
      }
      CU(cudaMemsetAsync(cur + 3, 0, sizeof(unsigned int), x->cst));
      CU(launch_pair(*pec->pe, s.d_out, cnt / 2, x->d_pe_stats, x->d_len_dist, s.d_orphans, cur + 3, x->d_chrom_keep, x->cst));
      x->launches += 1;
      if (rescue) {
        CU(launch_rescue(x->d, k, *pec->pe, s.d_out, s.d_orphans, cur + 3, kbases, s.d_offs,
                         std::max<int>((int)max_len, 32), x->d_pe_stats, x->d_len_dist, cur, x->d_chrom_keep, x->cst));
        x->launches += 1;
      }
    }
    if (compact_out) {
      CU(launch_compact_results(s.d_out, cnt, s.d_out16, x->cst));
      x->launches += 1;
    }
    CU(cudaEventRecord(s.k1, x->cst));
    mark(x->cst);
    CU(cudaStreamWaitEvent(s.st, s.k1, 0));
    if (compact_out)
      CU(cudaMemcpyAsync(out16 + start, s.d_out16, (size_t)cnt * sizeof(bkx_read_result16), cudaMemcpyDeviceToHost, s.st));
    else
      CU(cudaMemcpyAsync(out + start, s.d_out, (size_t)cnt * sizeof(bkx_read_result), cudaMemcpyDeviceToHost, s.st));
    if (multi)
      CU(cudaMemcpyAsync(multi + (size_t)start * (size_t)k.max_hits, s.d_multi,
                         (size_t)cnt * (size_t)k.max_hits * sizeof(bkx_multi_hit), cudaMemcpyDeviceToHost, s.st));
    mark(s.st);
    inflight[b] = true;
    start += cnt;
    base_pos += nb;
    b = (b + 1) % kSlots;
  }
  // drain in submission order (the oldest slice sits in slot b)
  for (int q = 0; q < kSlots; ++q) {
    int si = (b + q) % kSlots;
    if (inflight[si] && (rc = drain(si)) < 0) return rc;
  }
  x->last_ms = ms_total;
  for (int si = 0; si < kSlots; ++si) x->slot[si].timed = false;
  if (timeline) {
    cudaDeviceSynchronize();
    fprintf(stderr, "[bkx timeline] ms since call start, per slice: stage marks\n");
    for (size_t i = 0; i < tl.size(); ++i) {
      float ms = 0.f;
      cudaEventElapsedTime(&ms, tl0, tl[i]);
      fprintf(stderr, "%s%.2f", (i % 5) ? " " : "\n  ", ms);
      cudaEventDestroy(tl[i]);
    }
    fprintf(stderr, "\n");
    cudaEventDestroy(tl0);
  }
  if (stats) {
    bkx_align_stats h;
    CU(cudaMemcpy(&h, x->d_stats, sizeof(h), cudaMemcpyDeviceToHost));
    uint64_t* d = (uint64_t*)stats;
    const uint64_t* s = (const uint64_t*)&h;
    for (size_t i = 0; i < sizeof(h) / 8; ++i) d[i] += s[i];
  }
  if (pec && pec->stats) {
    bkx_pe_stats h;
    CU(cudaMemcpy(&h, x->d_pe_stats, sizeof(h), cudaMemcpyDeviceToHost));
    uint64_t* d = (uint64_t*)pec->stats;
    const uint64_t* s = (const uint64_t*)&h;
    for (size_t i = 0; i < sizeof(h) / 8; ++i) d[i] += s[i];
  }
  if (pec && pec->len_dist) {
    std::vector<uint32_t> ld(100001);
    CU(cudaMemcpy(ld.data(), x->d_len_dist, 100001 * 4, cudaMemcpyDeviceToHost));
    for (size_t i = 0; i < ld.size(); ++i) pec->len_dist[i] += ld[i];
  }
  return BKX_OK;
}

static HostReads reads_bytes(const uint8_t* bases, const uint64_t* offsets, bool packed4) {
  HostReads h;
  h.fmt = packed4 ? HostReads::PACKED4 : HostReads::BYTES;
  h.data = bases;
  h.offsets = offsets;
  return h;
}
static HostReads reads_packed2(const uint8_t* packed2, uint64_t first_base, const uint16_t* lens, uint32_t fixed_len,
                               const uint64_t* exc_pos, const uint8_t* exc_code, uint64_t n_exc) {
  HostReads h;
  h.fmt = HostReads::PACKED2;
  h.data = packed2; h.lens = lens; h.fixed_len = fixed_len; h.exc_pos = exc_pos; h.exc_code = exc_code; h.n_exc = n_exc;
  h.first_base = first_base;
  return h;
}

extern "C" int bkx_align_pairs(bkx_index* x, const bkx_align_params* p, const bkx_pe_params* pe, const uint8_t* bases,
                               const uint64_t* offsets, uint32_t n_pairs, bkx_read_result* out, bkx_align_stats* stats,
                               bkx_pe_stats* pe_stats, uint32_t* len_dist) {
  if (n_pairs > 0x7fffffffu) return fail(BKX_ERR_PARAM, "too many pairs");
  PeCall pc;
  pc.pe = pe; pc.stats = pe_stats; pc.len_dist = len_dist;
  return align_host(x, p, reads_bytes(bases, offsets, false), 2 * n_pairs, out, false, stats, nullptr, &pc);
}

extern "C" int bkx_align_pairs_packed4(bkx_index* x, const bkx_align_params* p, const bkx_pe_params* pe, const uint8_t* packed,
                                       const uint64_t* offsets, uint32_t n_pairs, bkx_read_result* out, bkx_align_stats* stats,
                                       bkx_pe_stats* pe_stats, uint32_t* len_dist) {
  if (n_pairs > 0x7fffffffu) return fail(BKX_ERR_PARAM, "too many pairs");
  PeCall pc;
  pc.pe = pe; pc.stats = pe_stats; pc.len_dist = len_dist;
  return align_host(x, p, reads_bytes(packed, offsets, true), 2 * n_pairs, out, false, stats, nullptr, &pc);
}

extern "C" int bkx_align_reads(bkx_index* x, const bkx_align_params* p, const uint8_t* bases, const uint64_t* offsets,
                               uint32_t n_reads, bkx_read_result* out, bkx_align_stats* stats) {
  return align_host(x, p, reads_bytes(bases, offsets, false), n_reads, out, false, stats);
}

extern "C" int bkx_align_reads_multi(bkx_index* x, const bkx_align_params* p, const uint8_t* bases, const uint64_t* offsets,
                                     uint32_t n_reads, bkx_read_result* out, bkx_multi_hit* multi, bkx_align_stats* stats) {
  if (!multi) return fail(BKX_ERR_PARAM, "null argument");
  return align_host(x, p, reads_bytes(bases, offsets, false), n_reads, out, false, stats, multi);
}

extern "C" int bkx_align_reads_packed4(bkx_index* x, const bkx_align_params* p, const uint8_t* packed, const uint64_t* offsets,
                                       uint32_t n_reads, bkx_read_result* out, bkx_align_stats* stats) {
  return align_host(x, p, reads_bytes(packed, offsets, true), n_reads, out, false, stats);
}

extern "C" int bkx_align_reads_packed2(bkx_index* x, const bkx_align_params* p, const uint8_t* packed2, uint64_t first_base,
                                       const uint16_t* lens, uint32_t fixed_len, const uint64_t* exc_pos, const uint8_t* exc_code,
                                       uint64_t n_exc, uint32_t n_reads, bkx_read_result16* out, bkx_align_stats* stats) {
  if (p && p->ml_mode >= BKX_ML_UNIQ) return fail(BKX_ERR_UNSUPPORTED, "-r3..5 need bkx_align_reads_multi");
  return align_host(x, p, reads_packed2(packed2, first_base, lens, fixed_len, exc_pos, exc_code, n_exc), n_reads, out, true, stats);
}

extern "C" int bkx_align_pairs_packed2(bkx_index* x, const bkx_align_params* p, const bkx_pe_params* pe, const uint8_t* packed2,
                                       uint64_t first_base, const uint16_t* lens, uint32_t fixed_len, const uint64_t* exc_pos,
                                       const uint8_t* exc_code, uint64_t n_exc, uint32_t n_pairs, bkx_read_result16* out,
                                       bkx_align_stats* stats, bkx_pe_stats* pe_stats, uint32_t* len_dist) {
  if (n_pairs > 0x7fffffffu) return fail(BKX_ERR_PARAM, "too many pairs");
  PeCall pc;
  pc.pe = pe; pc.stats = pe_stats; pc.len_dist = len_dist;
  return align_host(x, p, reads_packed2(packed2, first_base, lens, fixed_len, exc_pos, exc_code, n_exc), 2 * n_pairs, out, true, stats,
                    nullptr, &pc);
}

// One-byte codes -> 2-bit stream + exception list (the loop a loader fuses into its parser).  A code above T goes into the
// list with its low 3 bits (N = 4 and the InDel / undefined codes the search rejects as eNARNs) and leaves 0 in the stream.
extern "C" int64_t bkx_pack_bases2(const uint8_t* bases, uint64_t n_bases, uint8_t* packed2, uint64_t* exc_pos, uint8_t* exc_code,
                                   uint64_t exc_cap) {
  if ((!bases || !packed2) && n_bases) return fail(BKX_ERR_PARAM, "null argument");
  uint64_t ne = 0;
  const uint64_t nbytes = (n_bases + 3) / 4;
  for (uint64_t b = 0; b < nbytes; ++b) {
    unsigned v = 0;
    for (unsigned j = 0; j < 4; ++j) {
      const uint64_t i = 4 * b + j;
      if (i >= n_bases) break;
      const unsigned c = bases[i] & 7u;
      if (c > 3) {
        if (ne >= exc_cap || !exc_pos || !exc_code) return fail(BKX_ERR_MEM, "exception list too small (%llu entries)", (unsigned long long)exc_cap);
        exc_pos[ne] = i; exc_code[ne] = (uint8_t)c; ++ne;
      } else v |= c << (2 * j);
    }
    packed2[b] = (uint8_t)v;
  }
  return (int64_t)ne;
}

extern "C" int bkx_expand_results16(const bkx_read_result16* in, uint32_t n_reads, const uint16_t* lens, uint32_t fixed_len,
                                    bkx_read_result* out) {
  if ((!in || !out) && n_reads) return fail(BKX_ERR_PARAM, "null argument");
  static const uint8_t kStrand[4] = {0, '+', '-', '?'};
  for (uint32_t i = 0; i < n_reads; ++i) {
    const bkx_read_result16 c = in[i];
    bkx_read_result r;
    memset(&r, 0, sizeof(r));
    r.nar = c.nar_hr & 0x1f; r.hit_rslt = c.nar_hr >> 5;
    r.strand = kStrand[c.strand_flags & 3]; r.flags = (c.strand_flags >> 2) & 3;
    r.num_hits = c.num_hits; r.mismatches = c.mismatches; r.low_mm = c.low_mm; r.nxt_low_mm = c.nxt_low_mm;
    r.low_hit_instances = c.low_hit_instances; r.chrom_id = c.chrom_id; r.match_loci = c.match_loci;
    r.match_len = r.strand ? (uint16_t)(lens ? lens[i] : fixed_len) : 0;
    out[i] = r;
  }
  return BKX_OK;
}

extern "C" int bkx_pack_bases4(const uint8_t* bases, uint64_t n_bases, uint8_t* packed) {
  if ((!bases || !packed) && n_bases) return fail(BKX_ERR_PARAM, "null argument");
  for (uint64_t i = 0; i + 1 < n_bases; i += 2) packed[i >> 1] = (uint8_t)((bases[i] & 0x0f) | (bases[i + 1] << 4));
  if (n_bases & 1) packed[n_bases >> 1] = (uint8_t)(bases[n_bases - 1] & 0x0f);
  return BKX_OK;
}

extern "C" int bkx_align_one(bkx_index* x, const bkx_align_params* p, const uint8_t* probe, int probe_len,
                             int* low_hit_instances, int* low_mm, int* nxt_low_mm, bkx_read_result* hit) {
  if (!probe || probe_len < 1 || !hit) return fail(BKX_ERR_PARAM, "null argument");
  uint64_t offs[2] = {0, (uint64_t)probe_len};
  bkx_read_result r;
  int rc = bkx_align_reads(x, p, probe, offs, 1, &r, nullptr);
  if (rc < 0) return rc;
  *hit = r;
  if (low_hit_instances) *low_hit_instances = r.low_hit_instances;
  if (low_mm) *low_mm = r.low_mm;
  if (nxt_low_mm) *nxt_low_mm = r.nxt_low_mm;
  return r.hit_rslt;
}

extern "C" int bkx_pair_reads(bkx_index* x, const bkx_align_params* p, const bkx_pe_params* pe, bkx_read_result* results,
                              uint32_t n_pairs, const uint8_t* bases, const uint64_t* offsets, bkx_pe_stats* stats,
                              uint32_t* len_dist) {
  if (!x || !pe || !results) return fail(BKX_ERR_PARAM, "null argument");
  if (pe->pe_proc < BKX_PE_ORPHAN || pe->pe_proc > BKX_PE_UNIQUE_SE) return fail(BKX_ERR_PARAM, "bad pe_proc %d", pe->pe_proc);
  const bool rescue = pe->pe_proc == BKX_PE_ORPHAN || pe->pe_proc == BKX_PE_ORPHAN_SE;
  if (pe->pair_min_len < 25 || pe->pair_max_len > 100000 || pe->pair_min_len > pe->pair_max_len)
    return fail(BKX_ERR_PARAM, "bad insert size range %d..%d", pe->pair_min_len, pe->pair_max_len);
  KParams k{};
  if (rescue) {
    int rc = check_params(p, &k);
    if (rc < 0) return rc;
    if (!bases || !offsets) return fail(BKX_ERR_PARAM, "orphan recovery needs the read sequences");
  }
  if (n_pairs == 0) return BKX_OK;
  int Lmax = 0;
  if (rescue) {  // checked before anything is copied or launched: the caller's records stay untouched on error
    for (uint64_t i = 0; i < 2 * (uint64_t)n_pairs; ++i) {
      if (offsets[i + 1] < offsets[i]) return fail(BKX_ERR_PARAM, "offsets not monotonic at read %llu", (unsigned long long)i);
      Lmax = std::max<int>(Lmax, (int)std::min<uint64_t>(offsets[i + 1] - offsets[i], 1u << 30));
    }
    if (Lmax > kRescueMaxLen) return fail(BKX_ERR_PARAM, "read length %d exceeds cMaxSeqLen", Lmax);
  }
  std::lock_guard<std::mutex> lk(x->mtx);
  CU(cudaSetDevice(x->device));
  cudaStream_t st = x->slot[0].st;
  if (!x->d_len_dist) CU(cudaMalloc((void**)&x->d_len_dist, 100001 * 4));
  CU(cudaMemsetAsync(x->d_len_dist, 0, 100001 * 4, st));
  CU(cudaMemsetAsync(x->d_pe_stats, 0, sizeof(bkx_pe_stats), st));
  bkx_read_result* d_res = nullptr;
  uint32_t* d_list = nullptr;
  uint8_t* d_bases = nullptr;
  uint64_t* d_offs = nullptr;
  unsigned int* cnt = x->d_cursor[0];  // [0] rescue cursor, [3] orphan count
  size_t bytes = (size_t)n_pairs * 2 * sizeof(bkx_read_result);
  cudaError_t e = cudaMalloc((void**)&d_res, bytes);
  if (e == cudaSuccess && rescue) {
    uint64_t nb = offsets[2 * (uint64_t)n_pairs] - offsets[0];
    e = cudaMalloc((void**)&d_list, (size_t)n_pairs * 4);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_bases, nb + 64);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_offs, (2 * (size_t)n_pairs + 1) * 8);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_bases, bases + offsets[0], nb, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_offs, offsets, (2 * (size_t)n_pairs + 1) * 8, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaMemsetAsync(cnt + 3, 0, sizeof(unsigned int), st);
  }
  if (e == cudaSuccess) e = cudaMemcpyAsync(d_res, results, bytes, cudaMemcpyHostToDevice, st);
  if (e == cudaSuccess) e = launch_pair(*pe, d_res, n_pairs, x->d_pe_stats, x->d_len_dist, d_list, cnt + 3, x->d_chrom_keep, st);
  if (e == cudaSuccess && rescue)
    e = launch_rescue(x->d, k, *pe, d_res, d_list, cnt + 3, d_bases - offsets[0], d_offs, std::max(Lmax, 32), x->d_pe_stats,
                      x->d_len_dist, cnt, x->d_chrom_keep, st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(results, d_res, bytes, cudaMemcpyDeviceToHost, st);
  bkx_pe_stats hs;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&hs, x->d_pe_stats, sizeof(hs), cudaMemcpyDeviceToHost, st);
  std::vector<uint32_t> ld;
  if (len_dist && e == cudaSuccess) {
    ld.resize(100001);
    e = cudaMemcpyAsync(ld.data(), x->d_len_dist, 100001 * 4, cudaMemcpyDeviceToHost, st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  cudaFree(d_res); cudaFree(d_list); cudaFree(d_bases); cudaFree(d_offs);
  if (e != cudaSuccess) return fail(BKX_ERR_CUDA, "bkx_pair_reads: %s", cudaGetErrorString(e));
  x->launches += rescue ? 2 : 1;
  if (stats) {
    uint64_t* d = (uint64_t*)stats;
    const uint64_t* s = (const uint64_t*)&hs;
    for (size_t i = 0; i < sizeof(hs) / 8; ++i) d[i] += s[i];
  }
  if (len_dist) for (size_t i = 0; i < ld.size(); ++i) len_dist[i] += ld[i];
  return BKX_OK;
}

// -Z / -z in paired-end runs: the keep map the pairing kernels consult (AcceptThisChromID, Aligner.cpp:2651-2710).
// keep[id] != 0: alignments to chromosome id (1..num_entries) stay; keep[0] is not read.  NULL / 0 clears the filter.
extern "C" int bkx_set_chrom_filter(bkx_index* x, const uint8_t* keep, uint32_t n_keep) {
  if (!x) return fail(BKX_ERR_PARAM, "null argument");
  std::lock_guard<std::mutex> lk(x->mtx);
  CU(cudaSetDevice(x->device));
  CU(cudaDeviceSynchronize());   // no pairing launch of this index may still be reading the old map
  if (x->d_chrom_keep) { cudaFree(x->d_chrom_keep); x->d_chrom_keep = nullptr; }
  if (!keep || n_keep == 0) return BKX_OK;
  if (n_keep != x->info.num_entries + 1)
    return fail(BKX_ERR_PARAM, "chromosome filter has %u entries, the index needs %u (ids 0..%u)", n_keep,
                x->info.num_entries + 1, x->info.num_entries);
  CU(cudaMalloc((void**)&x->d_chrom_keep, n_keep));
  CU(cudaMemcpy(x->d_chrom_keep, keep, n_keep, cudaMemcpyHostToDevice));
  CU(cudaDeviceSynchronize());   // the library's streams are non-blocking: the map is in place before any of them runs
  return BKX_OK;
}

// Device-resident variant of bkx_pair_reads: results / reads / stats / histogram already on idx's GPU;
// asynchronous on `cuda_stream` (NULL = the index's own stream).  d_len_dist: 100001 u32 or NULL.
extern "C" int bkx_pair_reads_device(bkx_index* x, const bkx_align_params* p, const bkx_pe_params* pe,
                                     bkx_read_result* d_results, uint32_t n_pairs, const uint8_t* d_bases,
                                     const uint64_t* d_offsets, uint32_t max_read_len, bkx_pe_stats* d_stats,
                                     uint32_t* d_len_dist, void* cuda_stream) {
  if (!x || !pe || !d_results) return fail(BKX_ERR_PARAM, "null argument");
  if (pe->pe_proc < BKX_PE_ORPHAN || pe->pe_proc > BKX_PE_UNIQUE_SE) return fail(BKX_ERR_PARAM, "bad pe_proc %d", pe->pe_proc);
  const bool rescue = pe->pe_proc == BKX_PE_ORPHAN || pe->pe_proc == BKX_PE_ORPHAN_SE;
  if (pe->pair_min_len < 25 || pe->pair_max_len > 100000 || pe->pair_min_len > pe->pair_max_len)
    return fail(BKX_ERR_PARAM, "bad insert size range %d..%d", pe->pair_min_len, pe->pair_max_len);
  KParams k{};
  if (rescue) {
    int rc = check_params(p, &k);
    if (rc < 0) return rc;
    if (!d_bases || !d_offsets) return fail(BKX_ERR_PARAM, "orphan recovery needs the read sequences");
    if (max_read_len > (uint32_t)kRescueMaxLen) return fail(BKX_ERR_PARAM, "read length %u exceeds cMaxSeqLen", max_read_len);
  }
  if (n_pairs == 0) return BKX_OK;
  std::lock_guard<std::mutex> lk(x->mtx);
  CU(cudaSetDevice(x->device));
  cudaStream_t st = cuda_stream ? (cudaStream_t)cuda_stream : x->slot[0].st;
  if (rescue && n_pairs > x->pe_list_cap) {
    CU(cudaStreamSynchronize(st));
    if (x->d_pe_list) cudaFree(x->d_pe_list);
    x->pe_list_cap = (size_t)n_pairs * 5 / 4;
    CU(cudaMalloc((void**)&x->d_pe_list, x->pe_list_cap * 4));
  }
  unsigned int* cnt = x->d_cursor[1];  // [0] rescue cursor, [3] orphan count (slot 1's scalars are free here)
  CU(cudaMemsetAsync(cnt + 3, 0, sizeof(unsigned int), st));
  CU(launch_pair(*pe, d_results, n_pairs, d_stats, d_len_dist, rescue ? x->d_pe_list : nullptr, cnt + 3, x->d_chrom_keep, st));
  if (rescue)
    CU(launch_rescue(x->d, k, *pe, d_results, x->d_pe_list, cnt + 3, d_bases, d_offsets, std::max<int>((int)max_read_len, 32),
                     d_stats, d_len_dist, cnt, x->d_chrom_keep, st));
  x->launches += rescue ? 2 : 1;
  return BKX_OK;
}

// Page-lock / unlock caller memory so the library's H2D / D2H copies run asynchronously at full PCIe speed.
// Page-locked host memory from the start (no registration pass later): for buffers the caller fills itself and then hands
// to the host-buffer calls -- bkx-align's read stream and record array.
extern "C" void* bkx_alloc_host(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); fail(BKX_ERR_MEM, "cudaHostAlloc(%zu bytes) failed", bytes); return nullptr; }
  return p;
}
extern "C" void bkx_free_host(void* ptr) { if (ptr) cudaFreeHost(ptr); }

extern "C" int bkx_pin_host(void* ptr, size_t bytes) {
  if (!ptr || !bytes) return BKX_OK;
  cudaError_t e = cudaHostRegister(ptr, bytes, cudaHostRegisterPortable);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(BKX_ERR_CUDA, "cudaHostRegister(%zu bytes): %s", bytes, cudaGetErrorString(e)); }
  return BKX_OK;
}
extern "C" int bkx_unpin_host(void* ptr) {
  if (!ptr) return BKX_OK;
  cudaError_t e = cudaHostUnregister(ptr);
  if (e != cudaSuccess) { cudaGetLastError(); return fail(BKX_ERR_CUDA, "cudaHostUnregister: %s", cudaGetErrorString(e)); }
  return BKX_OK;
}

extern "C" float bkx_last_kernel_ms(const bkx_index* cx) {
  bkx_index* x = const_cast<bkx_index*>(cx);
  if (!x) return -1.f;
  if (x->last_ms == -2.f) {
    Slot& s = x->slot[0];
    if (cudaEventSynchronize(s.k1) != cudaSuccess) return -1.f;
    float ms = -1.f;
    if (cudaEventElapsedTime(&ms, s.k0, s.k1) != cudaSuccess) return -1.f;
    x->last_ms = ms;
  }
  return x->last_ms;
}

extern "C" uint64_t bkx_kernel_launches(const bkx_index* x) { return x ? x->launches : 0; }
