// TEST INFRASTRUCTURE ONLY -- a test double of the C ABI entry points the front end `bkx-align` calls
// (include/bkx.h), answered by the CPU oracle (oracle/bk_oracle.c) instead of the GPU library.
//
// tests/test_host_cli_cpu.py links biokanga_b200/csrc/host/bkx_align_main.cpp against THIS file instead of
// libbkx.so, so that everything on the host side of the boundary -- option parsing, read parsing and packing, the
// post-alignment passes, the summary block and every writer -- is compared with the reference's own output files in
// the CPU-only test run.  Nothing here is shipped, built by the package Makefile, or reachable from the product:
// the real bkx-align links libbkx.so, whose compute entry points fail without a CUDA device.
#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../oracle/bk_oracle.h"

struct bkx_index { bko_index* o; std::vector<uint8_t> keep; };

static thread_local std::string g_err;

int bkx_fail(int code, const char* fmt, ...) {  // same helper the library's host code (bkx_cluster.cu) expects
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

static std::vector<uint8_t> unpack4(const uint8_t* packed, uint64_t n_bases) {
  std::vector<uint8_t> b(n_bases);
  for (uint64_t i = 0; i < n_bases; ++i) b[i] = (packed[i >> 1] >> ((i & 1) * 4)) & 15;
  return b;
}

extern "C" {

const char* bkx_last_error(void) { return g_err.c_str(); }

int bkx_open_index(const char* sfx_path, int, int, bkx_index** out) {
  bko_index* o = nullptr;
  int rc = bko_open(sfx_path, &o);
  if (rc < 0) return bkx_fail(rc, "unable to open '%s' (oracle code %d)", sfx_path, rc);
  *out = new bkx_index{o, {}};
  return BKX_OK;
}
int bkx_clone_index(const bkx_index*, int, bkx_index**) { return bkx_fail(BKX_ERR_UNSUPPORTED, "test double: one device only"); }
void bkx_close_index(bkx_index* x) { if (x) { bko_close(x->o); delete x; } }
int bkx_index_info_get(const bkx_index* x, bkx_index_info* out) { return bko_info(x->o, out); }
int bkx_get_entry(const bkx_index* x, uint32_t id, bkx_entry* out) { return bko_get_entry(x->o, id, out); }
int64_t bkx_get_seq(const bkx_index* x, uint32_t id, uint64_t loci, uint64_t len, uint8_t* buf) {
  bkx_entry e;
  if (bko_get_entry(x->o, id, &e) < 0) return bkx_fail(BKX_ERR_ENTRY, "no entry %u", id);
  if (loci >= e.seq_len) return 0;
  len = std::min<uint64_t>(len, e.seq_len - loci);
  memcpy(buf, bko_seq(x->o) + e.start_ofs + loci, len);
  return (int64_t)len;
}
int bkx_default_params(const bkx_index* x, int pmode, bkx_align_params* out) { return bko_default_params(x->o, pmode, out); }
int bkx_pin_host(void*, size_t) { return BKX_OK; }
int bkx_unpin_host(void*) { return BKX_OK; }

int bkx_align_reads_packed4(bkx_index* x, const bkx_align_params* p, const uint8_t* packed, const uint64_t* offs, uint32_t n,
                            bkx_read_result* out, bkx_align_stats* st) {
  std::vector<uint8_t> b = unpack4(packed, offs[n]);
  return bko_align_batch(x->o, p, b.data(), offs, n, out, st, 4);
}
int bkx_align_reads_multi(bkx_index* x, const bkx_align_params* p, const uint8_t* bases, const uint64_t* offs, uint32_t n,
                          bkx_read_result* out, bkx_multi_hit* multi, bkx_align_stats* st) {
  return bko_align_batch_multi(x->o, p, bases, offs, n, out, multi, st, 4);
}
int bkx_align_pairs_packed4(bkx_index* x, const bkx_align_params* p, const bkx_pe_params* pe, const uint8_t* packed,
                            const uint64_t* offs, uint32_t n_pairs, bkx_read_result* out, bkx_align_stats* st,
                            bkx_pe_stats* pst, uint32_t* len_dist) {
  std::vector<uint8_t> b = unpack4(packed, offs[2 * (uint64_t)n_pairs]);
  int rc = bko_align_batch(x->o, p, b.data(), offs, 2 * n_pairs, out, st, 4);
  if (rc < 0) return rc;
  return bko_pair_reads_filtered(x->o, p, pe, out, n_pairs, b.data(), offs, pst, len_dist,
                                 x->keep.empty() ? nullptr : x->keep.data());
}
int bkx_set_chrom_filter(bkx_index* x, const uint8_t* keep, uint32_t n_keep) {
  bkx_index_info info;
  bko_info(x->o, &info);
  x->keep.clear();
  if (!keep || n_keep == 0) return BKX_OK;
  if (n_keep != info.num_entries + 1) return bkx_fail(BKX_ERR_PARAM, "chromosome filter has %u entries", n_keep);
  x->keep.assign(keep, keep + n_keep);
  return BKX_OK;
}

// SortHitMatch order (Aligner.cpp:10067-10114) with ties by record index -- the contract of bkx_sort_hits.
int bkx_sort_hits(const bkx_read_result* r, uint32_t n, uint32_t* order, int) {
  std::iota(order, order + n, 0u);
  auto key = [&](uint32_t i, uint64_t& hi, uint64_t& lo) {
    const bkx_read_result& a = r[i];
    const bool u = a.num_hits == 1;
    hi = ((uint64_t)a.nar << 41) | ((uint64_t)(u ? 0 : 1) << 40) | ((uint64_t)(u ? 0 : a.num_hits) << 32) | (u ? a.chrom_id : 0u);
    lo = u ? ((uint64_t)a.match_loci << 32) | ((uint64_t)a.match_len << 16) | ((uint64_t)a.strand << 8) | (uint8_t)a.low_mm : 0;
  };
  std::stable_sort(order, order + n, [&](uint32_t x, uint32_t y) {
    uint64_t xh, xl, yh, yl;
    key(x, xh, xl);
    key(y, yh, yl);
    return xh != yh ? xh < yh : xl < yl;
  });
  return BKX_OK;
}

}  // extern "C"
