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
#include <cstdlib>
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
void* bkx_alloc_host(size_t bytes) { return malloc(bytes ? bytes : 1); }
void bkx_free_host(void* p) { free(p); }
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
int bkx_pair_reads(bkx_index* x, const bkx_align_params* p, const bkx_pe_params* pe, bkx_read_result* results, uint32_t n_pairs,
                   const uint8_t* bases, const uint64_t* offs, bkx_pe_stats* pst, uint32_t* len_dist) {
  return bko_pair_reads_filtered(x->o, p, pe, results, n_pairs, bases, offs, pst, len_dist,
                                 x->keep.empty() ? nullptr : x->keep.data());
}
// the compact host interface: 2 bits per base + exception list in, 16-byte records out
static void unpack2(const uint8_t* packed2, uint64_t first_base, const uint16_t* lens, uint32_t fixed_len, const uint64_t* exc_pos,
                    const uint8_t* exc_code, uint64_t n_exc, uint32_t n, std::vector<uint8_t>& b, std::vector<uint64_t>& offs) {
  offs.assign((size_t)n + 1, 0);
  for (uint32_t i = 0; i < n; ++i) offs[i + 1] = offs[i] + (lens ? lens[i] : fixed_len);
  b.resize(offs[n]);
  for (uint64_t i = 0; i < offs[n]; ++i) { const uint64_t q = first_base + i; b[i] = (packed2[q >> 2] >> ((q & 3) * 2)) & 3; }
  for (uint64_t k = 0; k < n_exc; ++k)
    if (exc_pos[k] >= first_base && exc_pos[k] - first_base < offs[n]) b[exc_pos[k] - first_base] = exc_code[k];
}
static void compact16(const std::vector<bkx_read_result>& r, bkx_read_result16* out) {
  for (size_t i = 0; i < r.size(); ++i) {
    bkx_read_result16 c;
    c.nar_hr = (uint8_t)((r[i].nar & 0x1f) | (r[i].hit_rslt << 5));
    const unsigned sc = r[i].strand == '+' ? 1u : r[i].strand == '-' ? 2u : r[i].strand == '?' ? 3u : 0u;
    c.strand_flags = (uint8_t)(sc | ((r[i].flags & 3u) << 2));
    c.num_hits = r[i].num_hits; c.mismatches = r[i].mismatches; c.low_mm = r[i].low_mm; c.nxt_low_mm = r[i].nxt_low_mm;
    c.low_hit_instances = r[i].low_hit_instances; c.chrom_id = r[i].chrom_id; c.match_loci = r[i].match_loci;
    out[i] = c;
  }
}
int bkx_align_reads_packed2(bkx_index* x, const bkx_align_params* p, const uint8_t* packed2, uint64_t first_base, const uint16_t* lens,
                            uint32_t fixed_len, const uint64_t* exc_pos, const uint8_t* exc_code, uint64_t n_exc, uint32_t n,
                            bkx_read_result16* out, bkx_align_stats* st) {
  std::vector<uint8_t> b;
  std::vector<uint64_t> offs;
  unpack2(packed2, first_base, lens, fixed_len, exc_pos, exc_code, n_exc, n, b, offs);
  std::vector<bkx_read_result> r(n);
  int rc = bko_align_batch(x->o, p, b.data(), offs.data(), n, r.data(), st, 4);
  if (rc < 0) return rc;
  compact16(r, out);
  return BKX_OK;
}
int bkx_align_pairs_packed2(bkx_index* x, const bkx_align_params* p, const bkx_pe_params* pe, const uint8_t* packed2,
                            uint64_t first_base, const uint16_t* lens, uint32_t fixed_len, const uint64_t* exc_pos,
                            const uint8_t* exc_code, uint64_t n_exc, uint32_t n_pairs, bkx_read_result16* out, bkx_align_stats* st,
                            bkx_pe_stats* pst, uint32_t* len_dist) {
  std::vector<uint8_t> b;
  std::vector<uint64_t> offs;
  unpack2(packed2, first_base, lens, fixed_len, exc_pos, exc_code, n_exc, 2 * n_pairs, b, offs);
  std::vector<bkx_read_result> r((size_t)2 * n_pairs);
  int rc = bko_align_batch(x->o, p, b.data(), offs.data(), 2 * n_pairs, r.data(), st, 4);
  if (rc < 0) return rc;
  rc = bko_pair_reads_filtered(x->o, p, pe, r.data(), n_pairs, b.data(), offs.data(), pst, len_dist,
                               x->keep.empty() ? nullptr : x->keep.data());
  if (rc < 0) return rc;
  compact16(r, out);
  return BKX_OK;
}
int bkx_expand_results16(const bkx_read_result16* in, uint32_t n, const uint16_t* lens, uint32_t fixed_len, bkx_read_result* out) {
  static const uint8_t kStrand[4] = {0, '+', '-', '?'};
  for (uint32_t i = 0; i < n; ++i) {
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
