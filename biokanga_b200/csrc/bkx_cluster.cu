// Assignment of multi-loci reads to one locus by clustering (-r3 / -r4): host code, no device work.
//
// Replaces CAligner::AssignMultiMatches with its helper ProcAssignMultiMatches (biokanga/Aligner.cpp:5108-5270,
// 4961-5105) and the orderings SortMultiHits / SortMultiHitReadIDs (:10119-10199).  Behaviour, restated:
//   * every locus of every read whose search ended eHRhits takes part; loci of reads with several loci are "multi",
//     the single locus of a uniquely placed read is "unique" (AddMHitReads, :9560-9610);
//   * loci are ordered by chromosome, start, length, mismatches, strand, read id; each multi locus collects a score
//     from the loci of OTHER reads on the same strand that overlap it by at least 10 bp: first walking up the order
//     (stop at another chromosome or at a start >= longest-read bp away; partial sums clamped to 0x1fff), then down
//     (stop at the first start beyond end - 10; clamp 0x3fff).  An overlapping unique locus adds 1 + overlap*5/10 and
//     switches the score to "near unique" (bit 15), discarding whatever multi loci had contributed; before that an
//     overlapping multi locus adds 1 + overlap/10 (never in -r3).  A locus identical to the one scored just before it
//     copies that score;
//   * per read the best-scoring locus is taken if its score is >= 50 and, when the runner-up is of the same kind
//     (near unique or not), at least twice the runner-up's;
//   * a locus taken on the strength of other multi loci only survives if a unique or a taken locus starts within
//     10 + length of it; survivors become the read's alignment (NAR accepted, one hit, LowHitInstances 1).
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/bkx.h"

int bkx_fail(int code, const char* fmt, ...);

namespace {

constexpr uint16_t kUniqueFlag = 0x8000;   // cUniqueClustFlg
constexpr int kOverlap = 10;               // cClustMultiOverLap
constexpr int kUniqueScore = 5, kMultiScore = 1, kScale = 10;   // cClustUniqueScore / cClustMultiScore / cClustScaleFact
constexpr uint32_t kMinScore = 50;         // cMHminScore

struct Locus {
  uint32_t read;        // 0-based record index (the reference's ReadID - 1)
  uint32_t chrom, start;
  uint16_t len;
  uint8_t mm, strand;
  uint8_t multi;        // FlagMH
  uint8_t taken;        // FlagMHA
  uint8_t near_unique;  // eHLclustunique (1) / eHLclustany (0) once taken
  uint16_t score;
};

bool by_position(const Locus& a, const Locus& b) {  // SortMultiHits
  if (a.chrom != b.chrom) return a.chrom < b.chrom;
  if (a.start != b.start) return a.start < b.start;
  if (a.len != b.len) return a.len < b.len;
  if (a.mm != b.mm) return a.mm < b.mm;
  if (a.strand != b.strand) return a.strand < b.strand;
  return a.read < b.read;
}

bool by_read_then_score(const Locus& a, const Locus& b) {  // SortMultiHitReadIDs
  if (a.read != b.read) return a.read < b.read;
  if (a.score != b.score) return a.score > b.score;
  if (a.chrom != b.chrom) return a.chrom < b.chrom;
  if (a.len != b.len) return a.len < b.len;
  if (a.mm != b.mm) return a.mm < b.mm;
  if (a.start != b.start) return a.start < b.start;
  return a.strand < b.strand;
}

}  // namespace

extern "C" int bkx_assign_multi_matches(bkx_read_result* results, uint32_t n_reads, const bkx_multi_hit* multi,
                                        int max_ml_matches, int ml_mode, uint32_t max_read_len, bkx_cluster_stats* out) {
  if (!results || !multi || !out) return bkx_fail(BKX_ERR_PARAM, "null argument");
  if (ml_mode != BKX_ML_UNIQ && ml_mode != BKX_ML_MULTI) return bkx_fail(BKX_ERR_PARAM, "clustering belongs to -r3 / -r4");
  if (max_ml_matches < 2) return bkx_fail(BKX_ERR_PARAM, "bad max_ml_matches %d", max_ml_matches);
  memset(out, 0, sizeof(*out));
  std::vector<Locus> h;
  for (uint32_t i = 0; i < n_reads; ++i) {
    const bkx_read_result& r = results[i];
    if (r.hit_rslt != BKX_HR_HITS) continue;
    if (r.nar == BKX_NAR_ACCEPTED) {
      h.push_back({i, r.chrom_id, r.match_loci, r.match_len, r.mismatches, r.strand, 0, 0, 0, 0});
    } else if (r.nar == BKX_NAR_MULTIALIGN) {
      int cnt = std::min<int>(r.low_hit_instances, max_ml_matches);
      for (int k = 0; k < cnt; ++k) {
        const bkx_multi_hit& m = multi[(size_t)i * (size_t)max_ml_matches + (size_t)k];
        h.push_back({i, m.chrom_id, m.match_loci, m.match_len, m.mismatches, m.strand, 1, 0, 0, 0});
      }
      out->multi_reads += 1;
    }
  }
  std::sort(h.begin(), h.end(), by_position);
  const size_t n = h.size();
  // ---- scores
  long prev = -1;  // the multi locus scored just before (pPrevProcCurHit)
  for (size_t i = 0; i < n; ++i) {
    Locus& c = h[i];
    if (!c.multi) continue;
    if (prev >= 0) {
      const Locus& q = h[(size_t)prev];
      if (q.start == c.start && q.len == c.len && q.strand == c.strand && q.chrom == c.chrom) { c.score = q.score; continue; }
    }
    c.score = 0;
    const int cs = (int)c.start, clen = (int)c.len, cend = cs + clen - 1;
    for (size_t j = i; j-- > 0;) {  // up the order
      const Locus& u = h[j];
      if (u.chrom != c.chrom) break;
      if (cs - (int)u.start >= (int)max_read_len) break;
      const int uend = (int)u.start + (int)u.len - 1;
      if (uend < cs + kOverlap) continue;
      const int overlap = std::min(clen, uend - cs);
      if ((ml_mode == BKX_ML_UNIQ && u.multi) || ((c.score & kUniqueFlag) && (uint32_t)(c.score & ~kUniqueFlag) >= 0x1fffu)) continue;
      if (u.strand != c.strand || u.read == c.read) continue;
      if (!u.multi) {
        uint32_t s = 1 + (uint32_t)(overlap * kUniqueScore) / kScale;
        if (c.score & kUniqueFlag) s += c.score & ~kUniqueFlag;
        if (s > 0x1fffu) s = 0x1fffu;
        c.score = (uint16_t)(s | kUniqueFlag);
        if (s == 0x1fffu) break;
      } else if (!(c.score & kUniqueFlag)) {
        uint32_t s = 1 + (uint32_t)(overlap * kMultiScore) / kScale;
        s += c.score & ~kUniqueFlag;
        if (s > 0x1fffu) s = 0x1fffu;
        c.score = (uint16_t)s;
      }
    }
    for (size_t j = i + 1; j < n; ++j) {  // down the order
      const Locus& d = h[j];
      if (d.chrom != c.chrom) break;
      if ((int)d.start > cend - kOverlap) break;
      const int overlap = std::min((int)d.len, cend - (int)d.start);
      if ((ml_mode == BKX_ML_UNIQ && d.multi) || ((c.score & kUniqueFlag) && (uint32_t)(c.score & ~kUniqueFlag) >= 0x3fffu)) continue;
      if (d.strand != c.strand || d.read == c.read) continue;
      if (!d.multi) {
        uint32_t s = 1 + (uint32_t)(overlap * kUniqueScore) / kScale;
        if (c.score & kUniqueFlag) s += c.score & ~kUniqueFlag;
        if (s > 0x3fffu) s = 0x3fffu;
        c.score = (uint16_t)(s | kUniqueFlag);
        if (s == 0x3fffu) break;
      } else if (!(c.score & kUniqueFlag)) {
        uint32_t s = 1 + (uint32_t)(overlap * kMultiScore) / kScale;
        s += c.score & ~kUniqueFlag;
        if (s > 0x3fffu) s = 0x3fffu;
        c.score = (uint16_t)s;
      }
    }
    prev = (long)i;
  }
  // ---- best locus per read
  std::sort(h.begin(), h.end(), by_read_then_score);
  {
    long cur = -1;
    for (size_t i = 0; i < n; ++i) {
      Locus& c = h[i];
      if (!c.multi) continue;
      if (cur == (long)c.read) continue;
      cur = (long)c.read;
      out->putative += 1;
      const uint32_t best = c.score & ~kUniqueFlag;
      if (best < kMinScore) continue;
      const uint16_t next_score = i + 1 < n ? h[i + 1].score : 0;
      if ((c.score & kUniqueFlag) == (next_score & kUniqueFlag)) {
        const uint32_t nxt = next_score & ~kUniqueFlag;
        if (best < nxt * 2) continue;
      }
      c.taken = 1;
      c.near_unique = (c.score & kUniqueFlag) ? 1 : 0;
    }
  }
  // ---- loci taken on multi evidence only need a unique or taken neighbour
  std::sort(h.begin(), h.end(), by_position);
  for (size_t i = 0; i < n; ++i) {
    Locus& c = h[i];
    if (!c.taken) continue;
    bool keep = c.near_unique != 0;
    if (!keep) {
      for (size_t j = i; j-- > 0 && !keep;) {
        const Locus& u = h[j];
        const uint32_t dist = c.start - u.start;   // unsigned, as in the reference
        if (dist > (uint32_t)(kOverlap + (int)u.len)) break;
        if (u.chrom != c.chrom) break;
        if (!u.multi || u.taken) keep = true;
      }
      for (size_t j = i + 1; j < n && !keep; ++j) {
        const Locus& d = h[j];
        const uint32_t dist = d.start - c.start;
        if (dist > (uint32_t)(kOverlap + (int)c.len)) break;
        if (d.chrom != c.chrom) break;
        if (!d.multi || d.taken) keep = true;
      }
      if (!keep) c.taken = 0;
    }
    if (keep) {
      bkx_read_result& r = results[c.read];
      r.nar = BKX_NAR_ACCEPTED;
      r.num_hits = 1;
      r.low_hit_instances = 1;
      r.chrom_id = c.chrom; r.match_loci = c.start; r.match_len = c.len; r.strand = c.strand; r.mismatches = c.mm;
      out->assigned += 1;
      if (c.near_unique) out->near_unique += 1; else out->near_multi += 1;
    }
  }
  return BKX_OK;
}
