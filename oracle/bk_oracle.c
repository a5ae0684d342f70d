/* TEST INFRASTRUCTURE ONLY -- see bk_oracle.h.
 *
 * Plain-C, single-read-at-a-time restatement of the reference's `biokanga align` hot path, written
 * from the behaviour of the reference (file:line citations below are relative to /root/reference):
 *   .sfx reading          libbiokanga/SfxArrayV2.cpp:551-747 (Disk2Hdr, Disk2Entries), SfxArrayV2.h:79-104,174-203
 *   SfxOfsToLoci          libbiokanga/SfxArrayV2.cpp:33-44
 *   MapChunkHit2Entry     libbiokanga/SfxArrayV2.cpp:2530-2575
 *   LocateFirst/LastExact libbiokanga/SfxArrayV2.cpp:7765-7876, 7914-8027
 *   LocateCoreMultiples   libbiokanga/SfxArrayV2.cpp:5693-6262   (non-chimeric, basespace branch)
 *   AlignReads            libbiokanga/SfxArrayV2.cpp:7666-7760   (indel/splice/chimeric passes off)
 *   ProcCoredApprox       biokanga/Aligner.cpp:9024-9505
 *   LocateCoredApprox     biokanga/Aligner.cpp:8727-8761         (MinCoreLen / MaxNumSlides rule)
 *   AcceptProvPE etc.     biokanga/Aligner.cpp:2726-2850, 3055-3489
 *   AlignPairedRead       libbiokanga/SfxArrayV2.cpp:8247-8433,  AdaptiveTrim :5482-5682
 *
 * PARITY PIN: the reference ships no tests or golden vectors; this file is
 * pinned against outputs of the reference binary itself (oracle/_ref/biokanga, built by
 * oracle/build_ref.sh) on the fixtures under tests/golden/ (made by oracle/make_fixtures.py; checked by
 * tests/test_oracle_golden.py) and, when oracle/_ref/biokanga is present, on freshly drawn option sets and
 * inputs (tests/test_host_fuzz_cpu.py, tests/fuzz_host_cli.py) and on full-size files (bench.py --dropin).
 */
#define _GNU_SOURCE
#include "bk_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct bko_index {
  uint8_t* file;          /* whole .sfx file when opened from disk, else NULL */
  const uint8_t* seq;     /* ConcatSeqLen symbols, 1 byte each */
  const uint8_t* sa;      /* ConcatSeqLen elements of sfx_el_size bytes */
  uint64_t concat_len;
  uint32_t el_size;
  uint32_t n_entries;
  bkx_entry* entries;     /* sorted by start_ofs */
  uint32_t version, attributes;
  char dataset[84];
};

/* ------------------------------------------------------------------------------------------------ */
static inline uint64_t sa_at(const bko_index* x, int64_t i) {
  const uint8_t* p = x->sa + (uint64_t)i * x->el_size;
  uint32_t lo;
  memcpy(&lo, p, 4);
  uint64_t v = lo;
  if (x->el_size == 5) v |= (uint64_t)p[4] << 32;
  return v;
}

static uint32_t rd32(const uint8_t* p) { uint32_t v; memcpy(&v, p, 4); return v; }
static uint64_t rd64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }

int bko_open(const char* path, bko_index** out) {
  FILE* f = fopen(path, "rb");
  if (!f) return BKX_ERR_FILE;
  fseeko(f, 0, SEEK_END);
  off_t sz = ftello(f);
  fseeko(f, 0, SEEK_SET);
  uint8_t* b = (uint8_t*)malloc((size_t)sz);
  if (!b) { fclose(f); return BKX_ERR_MEM; }
  size_t got = 0;
  while (got < (size_t)sz) {
    size_t n = fread(b + got, 1, (size_t)sz - got > (1u << 30) ? (1u << 30) : (size_t)sz - got, f);
    if (n == 0) break;
    got += n;
  }
  fclose(f);
  if (got != (size_t)sz || sz < 1224 - 45 * 2) { free(b); return BKX_ERR_FORMAT; }
  /* header: magic "sfx" + '3'..'5'; pack(4) fields; v3 used 36-byte names, v4+ 81-byte names */
  if (memcmp(b, "sfx", 3) != 0) { free(b); return BKX_ERR_FORMAT; }
  bko_index* x = (bko_index*)calloc(1, sizeof(*x));
  x->file = b;
  x->version = rd32(b + 4);
  x->attributes = rd32(b + 8);
  uint64_t entries_ofs = rd64(b + 20);
  uint64_t blk_ofs = rd64(b + 44);
  int name_len = (x->version <= 3) ? 36 : 81;
  memcpy(x->dataset, b + 52, (size_t)(name_len < 83 ? name_len : 83));
  const uint8_t* blk = b + blk_ofs;
  x->concat_len = rd64(blk + 8);
  x->el_size = rd32(blk + 16);
  x->seq = blk + 20;
  x->sa = x->seq + x->concat_len;
  const uint8_t* eb = b + entries_ofs;
  x->n_entries = rd32(eb);
  x->entries = (bkx_entry*)calloc(x->n_entries ? x->n_entries : 1, sizeof(bkx_entry));
  size_t esz = (size_t)(8 + name_len + 2 + 4 + 8 + 8);
  for (uint32_t i = 0; i < x->n_entries; i++) {
    const uint8_t* e = eb + 8 + esz * i;
    bkx_entry* o = &x->entries[i];
    o->entry_id = rd32(e);
    memcpy(o->name, e + 8, (size_t)name_len);
    o->name[name_len < 83 ? name_len : 83] = 0;
    o->seq_len = rd32(e + 8 + name_len + 2);
    o->start_ofs = rd64(e + 8 + name_len + 6);
    o->end_ofs = rd64(e + 8 + name_len + 14);
  }
  *out = x;
  return BKX_OK;
}

int bko_open_mem(const uint8_t* seq, uint64_t concat_len, const void* sa, uint32_t el, const bkx_entry* entries,
                 uint32_t n, bko_index** out) {
  bko_index* x = (bko_index*)calloc(1, sizeof(*x));
  x->seq = seq;
  x->sa = (const uint8_t*)sa;
  x->concat_len = concat_len;
  x->el_size = el;
  x->n_entries = n;
  x->entries = (bkx_entry*)malloc(sizeof(bkx_entry) * (n ? n : 1));
  memcpy(x->entries, entries, sizeof(bkx_entry) * n);
  x->version = 5;
  *out = x;
  return BKX_OK;
}

void bko_close(bko_index* x) {
  if (!x) return;
  free(x->file);
  free(x->entries);
  free(x);
}

static uint64_t tot_seq_len(const bko_index* x) { /* SfxArrayV2.cpp:2070-2082 */
  uint64_t t = 0;
  for (uint32_t i = 0; i < x->n_entries; i++) t += x->entries[i].seq_len;
  return t;
}

int bko_info(const bko_index* x, bkx_index_info* o) {
  memset(o, 0, sizeof(*o));
  o->concat_len = x->concat_len;
  o->tot_seq_len = tot_seq_len(x);
  o->num_entries = x->n_entries;
  o->sfx_el_size = x->el_size;
  o->version = x->version;
  o->attributes = x->attributes;
  memcpy(o->dataset_name, x->dataset, sizeof(o->dataset_name));
  return BKX_OK;
}

int bko_get_entry(const bko_index* x, uint32_t id, bkx_entry* o) {
  if (id < 1 || id > x->n_entries) return BKX_ERR_ENTRY;
  *o = x->entries[id - 1];
  return BKX_OK;
}
const uint8_t* bko_seq(const bko_index* x) { return x->seq; }
const void* bko_sa(const bko_index* x) { return x->sa; }

int bko_default_params(const bko_index* x, int pmode, bkx_align_params* p) {
  memset(p, 0, sizeof(*p));
  uint64_t t = tot_seq_len(x);
  int mcl; /* Aligner.cpp:8727-8739, cMinCoreLen = 4 */
  if (t <= 500000) mcl = 4;
  else if (t <= 20000000) mcl = 7;
  else if (t <= 250000000) mcl = 11;
  else if (t <= 3500000000ull) mcl = 12;
  else mcl = 15;
  int slides, iters; /* Aligner.cpp:8744-8760, :341-356 */
  switch (pmode) {
    case BKX_PMODE_ULTRASENS: slides = 9; iters = 20000; break;
    case BKX_PMODE_MORESENS: mcl += 1; slides = 8; iters = 10000; break;
    case BKX_PMODE_DEFAULT: mcl += 2; slides = 8; iters = 5000; break;
    default: mcl += 4; slides = 6; iters = 2500; break;
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

/* ------------------------------------------------------------------------------------------------ */
/* probe-vs-suffix comparison shared by the searches: symbols are compared on their low 4 bits, a
 * target EOS means "probe sorts before target" (SfxArrayV2.cpp:7792-7811). */
static inline int cmp_probe(const bko_index* x, const uint8_t* probe, int len, uint64_t loci) {
  const uint8_t* t = x->seq + loci;
  for (int i = 0; i < len; i++) {
    uint8_t e2 = t[i] & 0x0f;
    if (e2 == BKX_BASE_EOS) return -1;
    uint8_t e1 = probe[i] & 0x0f;
    if (e1 > e2) return 1;
    if (e1 < e2) return -1;
  }
  return 0;
}

/* The reference's two searches return the lowest / highest matching SA index inside [lo,hi]; on a
 * sorted array that is the classic lower / upper bound, which is how they are restated here. */
int64_t bko_locate_first_exact(const bko_index* x, const uint8_t* probe, int len, int64_t lo, int64_t hi) {
  int64_t l = lo, h = hi + 1;
  while (l < h) {
    int64_t m = l + (h - l) / 2;
    if (cmp_probe(x, probe, len, sa_at(x, m)) > 0) l = m + 1; else h = m;
  }
  if (l <= hi && cmp_probe(x, probe, len, sa_at(x, l)) == 0) return l + 1;
  return 0;
}

int64_t bko_locate_last_exact(const bko_index* x, const uint8_t* probe, int len, int64_t lo, int64_t hi) {
  int64_t l = lo, h = hi + 1;
  while (l < h) {
    int64_t m = l + (h - l) / 2;
    if (cmp_probe(x, probe, len, sa_at(x, m)) >= 0) l = m + 1; else h = m;
  }
  if (l - 1 >= lo && cmp_probe(x, probe, len, sa_at(x, l - 1)) == 0) return l;
  return 0;
}

static const bkx_entry* map_entry(const bko_index* x, uint64_t ofs) { /* SfxArrayV2.cpp:2530-2575 */
  int64_t lo = 0, hi = (int64_t)x->n_entries - 1;
  while (hi >= lo) {
    int64_t mid = (hi + lo) / 2;
    const bkx_entry* e = &x->entries[mid];
    if (e->start_ofs <= ofs && e->end_ofs >= ofs) return e;
    if (e->start_ofs > ofs) hi = mid - 1; else lo = mid + 1;
  }
  return NULL;
}

static void revcpl(uint8_t* s, int len) { /* CSeqTrans::ReverseComplement: A<->T, C<->G, others kept */
  for (int i = 0, j = len - 1; i <= j; i++, j--) {
    uint8_t a = s[i], b = s[j];
    a = (a & 0x07) < 4 ? (uint8_t)(3 - (a & 0x07)) : (uint8_t)(a & 0x07);
    b = (b & 0x07) < 4 ? (uint8_t)(3 - (b & 0x07)) : (uint8_t)(b & 0x07);
    s[i] = b;
    s[j] = a;
  }
}

/* "already processed" set of LocateCoreMultiples (SfxArrayV2.cpp:5931-5950): keys are 32-bit
 * truncations of (1 + locus - coreOfs); reset per strand; at most max_ident_nodes keys. */
typedef struct {
  uint32_t* slot;
  uint32_t* gen;
  uint32_t mask, cur_gen;
} seen_set;

static void seen_init(seen_set* s, uint32_t cap_pow2) {
  s->slot = (uint32_t*)calloc(cap_pow2, 4);
  s->gen = (uint32_t*)calloc(cap_pow2, 4);
  s->mask = cap_pow2 - 1;
  s->cur_gen = 0;
}
static void seen_free(seen_set* s) { free(s->slot); free(s->gen); }
static void seen_reset(seen_set* s) {
  if (++s->cur_gen == 0) { memset(s->gen, 0, 4ull * (s->mask + 1)); s->cur_gen = 1; }
}
static int seen_test_and_set(seen_set* s, uint32_t key) { /* 1 if already present */
  uint32_t h = (key * 2654435761u) & s->mask;
  while (s->gen[h] == s->cur_gen) {
    if (s->slot[h] == key) return 1;
    h = (h + 1) & s->mask;
  }
  s->gen[h] = s->cur_gen;
  s->slot[h] = key;
  return 0;
}

typedef struct {
  seen_set seen;
  uint32_t seeds, cands;
} work_ctx;

/* One phase: LocateCoreMultiples, SfxArrayV2.cpp:5693-6262, MinChimericLen = 0, basespace. */
static int locate_core_multiples(const bko_index* x, work_ctx* w, int max_tot_mm, int core_len, int core_delta,
                                 int max_slides, int mm_delta, int align_strand, int* p_inst, int* p_low,
                                 int* p_nxt, uint8_t* probe, int probe_len, int max_hits, bkx_read_result* hits,
                                 int cur_max_iter, int max_nodes) {
  if (*p_inst > max_hits && *p_low == 0) return BKX_HR_HITINSTS;                       /* :5778 */
  if (*p_inst >= 1 && *p_low == 0 && (*p_nxt - *p_low) < mm_delta) return BKX_HR_MMDELTA; /* :5782 */
  int64_t sfx_len = (int64_t)x->concat_len;
  int inst, low, nxt;
  if (*p_inst <= 0 || *p_low < 0 || *p_nxt < 0) {                                      /* :5790 */
    inst = *p_inst = 0;
    low = *p_low = max_tot_mm + mm_delta + 1;
    nxt = *p_nxt = low;
  } else {
    inst = *p_inst; low = *p_low; nxt = *p_nxt;
  }
  int cur_hit = inst < max_hits ? inst : -1; /* index of "pCurHit" (-1 = NULL) */
  char strand;
  if (align_strand == BKX_STRAND_CRICK) { revcpl(probe, probe_len); strand = '-'; } else strand = '+';
  int strands_left = align_strand; /* BOTH -> after '+' do '-' */
  int stop_all = 0;
  for (;;) {
    int cur_delta = core_delta, slides = 0, nodes = 0;
    seen_reset(&w->seen);
    for (int ofs = 0; slides < max_slides && ofs <= probe_len - core_len && cur_delta > core_len / 3 &&
                      nodes < max_nodes; slides++, ofs += cur_delta) {
      if (ofs + core_len + cur_delta > probe_len) cur_delta = probe_len - (ofs + core_len); /* :5846 */
      w->seeds++;
      int64_t ti = bko_locate_first_exact(x, probe + ofs, core_len, 0, sfx_len - 1);
      if (ti == 0) continue;
      ti -= 1;
      int iter = 0;
      uint32_t num_copies = 0;
      int first = 1;
      while (!cur_max_iter || iter < cur_max_iter) {
        if (nodes >= max_nodes) break;
        if (!first) {
          if (ti + 1 >= sfx_len || (int64_t)sa_at(x, ti + 1) + core_len > sfx_len) break;
          if (iter == 100 && !num_copies) {                                            /* :5868-5875 */
            int64_t last = bko_locate_last_exact(x, probe + ofs, core_len, ti - 1, sfx_len - 1);
            num_copies = last > 0 ? (uint32_t)(1 + last - ti) : 0;
            if (cur_max_iter && num_copies > (uint32_t)cur_max_iter) break;
          }
          if (cmp_probe(x, probe + ofs, core_len, sa_at(x, ti + 1)) != 0) break;
          ti += 1;
        }
        first = 0;
        uint64_t loci = sa_at(x, ti);
        if (loci < (uint32_t)ofs) continue;                                           /* :5918 */
        uint64_t left = loci - (uint64_t)ofs;
        const bkx_entry* e = map_entry(x, left);
        if (e == NULL || !probe_len || left + (uint32_t)probe_len - 1 > e->end_ofs) continue; /* :5928 */
        uint32_t key = (uint32_t)(1 + loci - (uint32_t)ofs);                          /* :5932 */
        if (seen_test_and_set(&w->seen, key)) continue;
        nodes++;
        iter++;
        w->cands++;
        /* Hamming, :6093-6154 */
        const uint8_t* t = x->seq + left;
        int mm = 0, i;
        for (i = 0; i < probe_len; i++) {
          uint8_t tb = t[i] & 0x0f, pb = probe[i] & 0x0f;
          if (tb == BKX_BASE_EOS) break;
          if (pb == tb) continue;
          if (++mm > max_tot_mm) break;
          if (mm >= nxt) break;
        }
        if (i != probe_len) continue;
        if (mm < low) {                                                               /* :6157 */
          cur_hit = 0;
          inst = 1;
          nxt = low;
          low = mm;
        } else if (mm == low) {
          inst += 1;
          if (cur_hit >= 0 && inst <= max_hits) cur_hit += 1; else goto no_store;
        } else {
          if (mm < nxt) nxt = mm;
          goto no_store;
        }
        {
          const bkx_entry* he = map_entry(x, left);
          bkx_read_result* h = &hits[cur_hit];
          memset(h, 0, sizeof(*h));
          h->strand = (uint8_t)strand;
          h->chrom_id = he->entry_id;
          h->match_loci = (uint32_t)(left - he->start_ofs);
          h->match_len = (uint16_t)probe_len;
          h->mismatches = (uint8_t)mm;
        }
      no_store:
        if (inst > max_hits && low == 0) break;                                        /* :6206 */
      }
      if (inst > max_hits && low == 0) { stop_all = 1; break; }                        /* :6210 */
    }
    if (stop_all) break;
    if (strand == '+' && strands_left == BKX_STRAND_BOTH) {                            /* :6216 */
      revcpl(probe, probe_len);
      strand = '-';
      strands_left = BKX_STRAND_CRICK;
    } else break;
    if (inst > max_hits && low == 0) break;
  }
  if (strand == '-') revcpl(probe, probe_len);                                        /* :6231 */

  if (*p_low == low && *p_inst == inst) {                                             /* :6238 */
    if (*p_nxt > nxt) {
      *p_nxt = nxt;
      if ((nxt - *p_low) < mm_delta) return BKX_HR_MMDELTA;
      return BKX_HR_RMMDELTA;
    }
    return BKX_HR_NONE;
  }
  *p_low = low; *p_inst = inst; *p_nxt = nxt;
  if (inst >= 1 && (nxt - low) < mm_delta) return BKX_HR_MMDELTA;
  if (inst > max_hits) return BKX_HR_HITINSTS;
  return BKX_HR_HITS;
}

/* -N: LocateBestMatches, SfxArrayV2.cpp:6654-7019 -- ONE pass (no staged phases) over the cores of both strands that keeps
 * the max_hits loci with the fewest mismatches (<= max_tot_mm), ordered by mismatches, equal ones in discovery order.
 * Returns 0 (no locus), else the number kept, plus one if a locus was turned away from a full list; *p_inst = number kept. */
static int locate_best_matches(const bko_index* x, work_ctx* w, int max_tot_mm, int core_len, int core_delta, int max_slides,
                               int align_strand, uint8_t* probe, int probe_len, int max_hits, int* p_inst,
                               bkx_read_result* hits, int cur_max_iter, int max_nodes) {
  int64_t sfx_len = (int64_t)x->concat_len;
  int inst = 0, sloughed = 0;
  char strand;
  *p_inst = 0;
  if (align_strand == BKX_STRAND_CRICK) { revcpl(probe, probe_len); strand = '-'; } else strand = '+';
  int strands_left = align_strand;
  do {
    int cur_delta = core_delta, slides = 0, nodes = 0;
    seen_reset(&w->seen);
    for (int ofs = 0; slides < max_slides && ofs <= probe_len - core_len && cur_delta > core_len / 3 && nodes < max_nodes;
         slides++, ofs += cur_delta) {                                                 /* :6748-6756 */
      if (ofs + core_len + cur_delta > probe_len) cur_delta = probe_len - (ofs + core_len);
      w->seeds++;
      int64_t ti = bko_locate_first_exact(x, probe + ofs, core_len, 0, sfx_len - 1);
      if (ti != 0) {
        ti -= 1;
        int iter = 0;
        uint32_t num_copies = 0;
        int first = 1;
        while (!cur_max_iter || iter < cur_max_iter) {
          if (nodes >= max_nodes) break;
          if (!first) {
            if (ti + 1 >= sfx_len || (int64_t)sa_at(x, ti + 1) + core_len > sfx_len) break;
            if (iter == 100 && !num_copies) {                                          /* :6781-6788 */
              int64_t last = bko_locate_last_exact(x, probe + ofs, core_len, ti - 1, sfx_len - 1);
              num_copies = last > 0 ? (uint32_t)(1 + last - ti) : 0;
              if (cur_max_iter && num_copies > (uint32_t)cur_max_iter) break;
            }
            if (cmp_probe(x, probe + ofs, core_len, sa_at(x, ti + 1)) != 0) break;
            ti += 1;
          }
          first = 0;
          uint64_t loci = sa_at(x, ti);
          if (loci < (uint32_t)ofs) continue;                                         /* :6832 */
          uint64_t left = loci - (uint64_t)ofs;
          if (!probe_len || left + (uint32_t)probe_len > x->concat_len) continue;     /* :6840 (no entry test here) */
          uint32_t key = (uint32_t)(1 + loci - (uint32_t)ofs);
          if (seen_test_and_set(&w->seen, key)) continue;
          nodes++;
          iter++;
          w->cands++;
          const uint8_t* t = x->seq + left;
          int mm = 0, i;
          for (i = 0; i < probe_len; i++) {                                           /* :6878-6934: EOS ends the match */
            uint8_t tb = t[i] & 0x0f, pb = probe[i] & 0x0f;
            if (tb == BKX_BASE_EOS) break;
            if (pb == tb) continue;
            if (++mm > max_tot_mm) break;
          }
          if (i != probe_len) continue;
          int at = -1;                                                                /* :6938-6959 */
          if (inst) {
            if (inst == max_hits) sloughed = 1;
            int b;
            for (b = 0; b < inst; b++)
              if (hits[b].mismatches > mm) {
                at = b;
                int move = inst - b;
                if (b + move >= max_hits) move = max_hits - 1 - b;   /* the last one falls off a full list */
                if (move > 0) memmove(&hits[b + 1], &hits[b], sizeof(*hits) * (size_t)move);
                break;
              }
            if (b == inst && inst < max_hits) at = inst;
          } else at = 0;
          if (at >= 0) {
            const bkx_entry* he = map_entry(x, left);
            bkx_read_result* h = &hits[at];
            memset(h, 0, sizeof(*h));
            h->strand = (uint8_t)strand;
            h->chrom_id = he ? he->entry_id : 0;
            h->match_loci = he ? (uint32_t)(left - he->start_ofs) : 0;
            h->match_len = (uint16_t)probe_len;
            h->mismatches = (uint8_t)mm;
            if (inst < max_hits) inst += 1;
            else max_tot_mm = hits[inst - 1].mismatches;                              /* :6980-6984 */
          }
        }
      }
      if (inst == max_hits && max_tot_mm == 0 && !sloughed) { strands_left = -1; break; }   /* :6988-6992 */
    }
    if (strands_left == -1) break;
    if (strand == '+' && strands_left == BKX_STRAND_BOTH) {
      revcpl(probe, probe_len);
      strand = '-';
      strands_left = BKX_STRAND_CRICK;
    } else strands_left = -1;
  } while (!(inst == max_hits && max_tot_mm == 0 && !sloughed) && strands_left != -1);
  if (strand == '-') revcpl(probe, probe_len);
  *p_inst = inst;
  if (inst == 0) return 0;
  return sloughed ? inst + 1 : inst;
}

static int align_reads(const bko_index* x, work_ctx* w, const bkx_align_params* p, int max_tot_mm, int core_len,
                       int core_delta, int max_slides, uint8_t* probe, int probe_len, int* inst, int* low,
                       int* nxt, bkx_read_result* hits) { /* SfxArrayV2.cpp:7666-7760 */
  int r = 0, allow = 0;
  int mmd = p->min_edit_dist;
  if (max_tot_mm > 0) {
    for (allow = 0; allow <= max_tot_mm; allow++) {
      int cl = probe_len / (allow + mmd);
      if (cl <= core_len) break;
      r = locate_core_multiples(x, w, allow, cl, cl, max_slides, mmd, p->align_strand, inst, low, nxt, probe,
                                probe_len, p->max_ml_matches, hits, p->max_iter, p->max_ident_nodes);
      if (r != 0) return r;
    }
  }
  if (allow <= max_tot_mm) {
    r = locate_core_multiples(x, w, max_tot_mm, core_len, core_delta, max_slides, mmd, p->align_strand, inst, low,
                              nxt, probe, probe_len, p->max_ml_matches, hits, p->max_iter, p->max_ident_nodes);
    if (r != 0) return r;
  }
  return 0;
}

int bko_align_reads_one(const bko_index* x, const bkx_align_params* p, int max_tot_mm, int core_len, int core_delta,
                        int max_slides, uint8_t* probe, int probe_len, int* inst, int* low, int* nxt,
                        bkx_read_result* hit, uint32_t* seeds, uint32_t* cands) {
  work_ctx w;
  memset(&w, 0, sizeof(w));
  seen_init(&w.seen, 1u << 12);
  /* grow on demand is not needed for the unit shim: cap the table at 2x max nodes */
  seen_free(&w.seen);
  uint32_t cap = 1u << 12;
  while (cap < 2u * (uint32_t)p->max_ident_nodes && cap < (1u << 22)) cap <<= 1;
  seen_init(&w.seen, cap);
  bkx_read_result* hits = (bkx_read_result*)calloc((size_t)p->max_ml_matches + 1, sizeof(*hits));
  int r = align_reads(x, &w, p, max_tot_mm, core_len, core_delta, max_slides, probe, probe_len, inst, low, nxt, hits);
  if (hit) *hit = hits[0];
  if (seeds) *seeds = w.seeds;
  if (cands) *cands = w.cands;
  free(hits);
  seen_free(&w.seen);
  return r;
}

/* Per-read driver: ProcCoredApprox body, Aligner.cpp:9027-9504 (default -r0 multi-loci mode). */
static void proc_read(const bko_index* x, work_ctx* w, const bkx_align_params* p, const uint8_t* rd, int len,
                      bkx_read_result* out, bkx_read_result* hits, uint8_t* seqbuf, bkx_align_stats* st) {
  memset(hits, 0, ((size_t)p->max_ml_matches + 1) * sizeof(*hits));
  memset(out, 0, sizeof(*out));
  out->nar = BKX_NAR_NOHIT;
  w->seeds = w->cands = 0;
  int max_ns_seq = 0, num_ns = 0, i;
  if (p->max_ns) {
    max_ns_seq = (len * p->max_ns) / 100;
    if (max_ns_seq < p->max_ns) max_ns_seq = p->max_ns;
  }
  for (i = 0; i < len; i++) {                                                          /* :9045-9055 */
    uint8_t b = rd[i] & 0x07;
    if (b > BKX_BASE_N) break;
    seqbuf[i] = b;
    if (b == BKX_BASE_N && ++num_ns > max_ns_seq) break;
  }
  if (i != len) {
    out->nar = BKX_NAR_NS;
    st->num_sloughed_ns++;
    st->nar[BKX_NAR_NS]++;
    return;
  }
  int max_tot_mm = p->max_subs == 0 ? 0 : (len * p->max_subs + 50) / 100;               /* :9085 */
  if (p->max_subs != 0 && max_tot_mm < 1) max_tot_mm = 1;
  if (max_tot_mm > 63) max_tot_mm = 63;
  int core_len = len / (p->min_edit_dist == 1 ? max_tot_mm + 1 : max_tot_mm + 2);       /* :9093 */
  if (core_len < p->min_core_len) core_len = p->min_core_len;
  int slides = (p->max_num_slides * len + 99) / 100;
  if (slides < 1) slides = 1;
  int core_delta = len / slides - 1;
  if (core_delta < core_len) core_delta = core_len;
  int inst = 0, low = 0, nxt = 0;
  memset(hits, 0, sizeof(*hits));
  int hr;
  if (p->best_matches && p->ml_mode != BKX_ML_DEFAULT) {                                 /* -N, Aligner.cpp:9197-9218 */
    hr = locate_best_matches(x, w, max_tot_mm, core_len, core_delta, slides, p->align_strand, seqbuf, len, p->max_ml_matches,
                             &inst, hits, p->max_iter, p->max_ident_nodes);
    hr = hr >= 1 ? BKX_HR_HITS : BKX_HR_NONE;
  } else
    hr = align_reads(x, w, p, max_tot_mm, core_len, core_delta, slides, seqbuf, len, &inst, &low, &nxt, hits);
  if (inst > p->max_ml_matches) inst = p->max_ml_matches + 1;                           /* :9241 */
  if ((p->clamp_max_ml || p->best_matches) && hr == BKX_HR_HITINSTS) { inst = p->max_ml_matches; hr = BKX_HR_HITS; } /* :9243 */
  out->hit_rslt = (uint8_t)hr;
  out->seeds = w->seeds;
  out->cands = w->cands;
  switch (hr) {
    case BKX_HR_NONE:
      out->nar = BKX_NAR_NOHIT;
      st->tot_non_aligned++;
      break;
    case BKX_HR_HITS:                                                                   /* :9328-9400 */
      st->tot_accepted_aligned++;
      st->tot_loci_aligned += (uint64_t)inst;
      if (inst == 1) st->tot_accepted_unique++; else st->tot_accepted_multi++;
      if (inst == 1 || p->ml_mode == BKX_ML_DEFAULT || p->ml_mode == BKX_ML_ALL) {  /* unique, or every locus is wanted */
        out->nar = BKX_NAR_ACCEPTED;
        out->num_hits = (p->ml_mode == BKX_ML_ALL) ? (uint8_t)(inst > 255 ? 255 : inst) : 1;  /* -r5: every hit is reported (:9336-9352) */
        out->strand = hits[0].strand;
        out->chrom_id = hits[0].chrom_id;
        out->match_loci = hits[0].match_loci;
        out->match_len = hits[0].match_len;
        out->mismatches = hits[0].mismatches;
        if (p->ml_mode != BKX_ML_ALL) inst = 1;
        if (out->strand == '+') st->plus_hits++; else st->minus_hits++;
      } else {                                       /* -r1 / -r3 / -r4: counted, not placed (:9383-9397) */
        out->nar = BKX_NAR_MULTIALIGN;
        out->num_hits = 0;
      }
      out->low_hit_instances = (int16_t)inst;
      out->low_mm = (int8_t)low;
      out->nxt_low_mm = (int8_t)nxt;
      break;
    case BKX_HR_MMDELTA:                                                                /* :9426-9438 */
      st->tot_not_accepted_delta++;
      out->nar = BKX_NAR_MMDELTA;
      out->strand = '?';
      out->match_len = (uint16_t)len;
      out->low_hit_instances = (int16_t)inst;
      out->low_mm = (int8_t)low;
      out->nxt_low_mm = (int8_t)nxt;
      break;
    case BKX_HR_HITINSTS:                                                               /* :9440-9479 */
      st->tot_non_aligned++;
      out->nar = BKX_NAR_MULTIALIGN;
      out->strand = '?';
      out->match_len = (uint16_t)len;
      out->low_hit_instances = (int16_t)inst;
      out->low_mm = (int8_t)low;
      out->nxt_low_mm = (int8_t)nxt;
      break;
    case BKX_HR_RMMDELTA:
      out->nxt_low_mm = (int8_t)nxt;
      break;
  }
  st->nar[out->nar]++;
}

typedef struct {
  const bko_index* x;
  const bkx_align_params* p;
  const uint8_t* bases;
  const uint64_t* offs;
  uint32_t n;
  bkx_read_result* out;
  bkx_multi_hit* multi;
  bkx_align_stats st;
  uint32_t* cursor;
  pthread_mutex_t* mtx;
} thr_arg;

static void* thr_main(void* a_) {
  thr_arg* a = (thr_arg*)a_;
  work_ctx w;
  memset(&w, 0, sizeof(w));
  uint32_t cap = 1u << 12;
  while (cap < 2u * (uint32_t)a->p->max_ident_nodes && cap < (1u << 22)) cap <<= 1;
  seen_init(&w.seen, cap);
  bkx_read_result* hits = (bkx_read_result*)calloc((size_t)a->p->max_ml_matches + 1, sizeof(*hits));
  uint8_t* seqbuf = (uint8_t*)malloc(1 << 16);
  for (;;) {
    pthread_mutex_lock(a->mtx);
    uint32_t s = *a->cursor;
    uint32_t e = s + 1024 > a->n ? a->n : s + 1024;
    *a->cursor = e;
    pthread_mutex_unlock(a->mtx);
    if (s >= e) break;
    for (uint32_t i = s; i < e; i++) {
      int len = (int)(a->offs[i + 1] - a->offs[i]);
      proc_read(a->x, &w, a->p, a->bases + a->offs[i], len, &a->out[i], hits, seqbuf, &a->st);
      if (a->multi) {  /* -r5: the pMultiHits list WriteHitLoci walks */
        bkx_multi_hit* m = a->multi + (size_t)i * (size_t)a->p->max_ml_matches;
        memset(m, 0, (size_t)a->p->max_ml_matches * sizeof(*m));
        int nh = a->out[i].nar == BKX_NAR_ACCEPTED ? (a->p->ml_mode == BKX_ML_ALL ? a->out[i].low_hit_instances : a->out[i].num_hits)
                 : (a->out[i].nar == BKX_NAR_MULTIALIGN && a->out[i].hit_rslt == BKX_HR_HITS) ? a->out[i].low_hit_instances : 0;
        if (nh > 0)
          for (int h = 0; h < nh && h < a->p->max_ml_matches; h++) {
            m[h].chrom_id = hits[h].chrom_id; m[h].match_loci = hits[h].match_loci; m[h].match_len = hits[h].match_len;
            m[h].strand = hits[h].strand; m[h].mismatches = hits[h].mismatches;
          }
      }
      a->st.seeds += a->out[i].seeds;
      a->st.cands += a->out[i].cands;
      a->st.reads++;
    }
  }
  free(seqbuf);
  free(hits);
  seen_free(&w.seen);
  return NULL;
}

int bko_align_batch_multi(const bko_index* x, const bkx_align_params* p, const uint8_t* bases, const uint64_t* offs,
                          uint32_t n, bkx_read_result* out, bkx_multi_hit* multi, bkx_align_stats* stats, int nthreads);

int bko_align_batch(const bko_index* x, const bkx_align_params* p, const uint8_t* bases, const uint64_t* offs,
                    uint32_t n, bkx_read_result* out, bkx_align_stats* stats, int nthreads) {
  return bko_align_batch_multi(x, p, bases, offs, n, out, NULL, stats, nthreads);
}

int bko_align_batch_multi(const bko_index* x, const bkx_align_params* p, const uint8_t* bases, const uint64_t* offs,
                          uint32_t n, bkx_read_result* out, bkx_multi_hit* multi, bkx_align_stats* stats, int nthreads) {
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  pthread_t tid[256];
  thr_arg* args = (thr_arg*)calloc((size_t)nthreads, sizeof(thr_arg));
  pthread_mutex_t mtx = PTHREAD_MUTEX_INITIALIZER;
  uint32_t cursor = 0;
  for (int t = 0; t < nthreads; t++) {
    args[t].x = x; args[t].p = p; args[t].bases = bases; args[t].offs = offs; args[t].n = n; args[t].out = out;
    args[t].multi = multi;
    args[t].cursor = &cursor; args[t].mtx = &mtx;
    if (nthreads == 1) thr_main(&args[t]); else pthread_create(&tid[t], NULL, thr_main, &args[t]);
  }
  for (int t = 0; t < nthreads; t++) {
    if (nthreads > 1) pthread_join(tid[t], NULL);
    if (stats) {
      uint64_t* d = (uint64_t*)stats;
      const uint64_t* s = (const uint64_t*)&args[t].st;
      for (size_t k = 0; k < sizeof(bkx_align_stats) / 8; k++) d[k] += s[k];
    }
  }
  free(args);
  return BKX_OK;
}

/* ------------------------------------------------------------------------------------------------ */
/* AdaptiveTrim restricted to what AlignPairedRead needs is added with orphan recovery (see
 * bko_pair_reads); PEInsertSize / AcceptProvPE: Aligner.cpp:2726-2850 with zero trims, no chrom filter. */
static int pe_insert_size(const bkx_pe_params* pe, uint8_t s1, uint32_t st1, uint32_t en1, uint8_t s2, uint32_t st2,
                          uint32_t en2) {
  if ((pe->pair_strand && s1 != s2) || (!pe->pair_strand && s1 == s2)) return -1;
  int frag;
  if (pe->circularised) {
    if (s1 == '+') frag = 1 + (int)en1 - (int)st2; else frag = 1 + (int)st2 - (int)en1;
  } else {
    if (s1 == '+') frag = 1 + (int)en2 - (int)st1; else frag = 1 + (int)en1 - (int)st2;
  }
  if (frag < 0) return -1;
  if (frag < pe->pair_min_len) return -6;
  if (frag > pe->pair_max_len) return -7;
  return frag;
}

/* keep: AcceptThisChromID per entry id (Aligner.cpp:2651-2710: exclude expressions first, then the include ones), NULL = no
 * -Z / -z filters; -3 both ends on a filtered chromosome, -4 the 5' end, -5 the 3' end (:2771-2786). */
static int accept_prov_pe(const bkx_pe_params* pe, const bkx_read_result* f, const bkx_read_result* r, const uint8_t* keep) {
  if (!(f->num_hits == 1 && r->num_hits == 1)) return 0;
  int bf = keep ? keep[f->chrom_id] : 1;
  if (f->chrom_id != r->chrom_id) {
    int br = keep ? keep[r->chrom_id] : 1;
    if (br && bf) return -2;
    if (!br && !bf) return -3;
    return !bf ? -4 : -5;
  }
  if (!bf) return -3;
  return pe_insert_size(pe, f->strand, f->match_loci, f->match_loci + f->match_len - 1, r->strand, r->match_loci,
                        r->match_loci + r->match_len - 1);
}

/* SfxArrayV2.cpp:5482-5682 restated for the only way AlignPairedRead calls it on this path:
 * MinTrimLen == SeqLen (no chimeric trimming) -- see the orphan-recovery notes in DESIGN.md. */
static int adaptive_trim_full(int len, const uint8_t* probe, const uint8_t* targ, int max_mm_rate, int min_flank,
                              uint32_t* mm_out);

static int align_paired_read(const bko_index* x, const bkx_pe_params* pe, const bkx_align_params* p,
                             int b3prime, int antisense, uint32_t chrom_id, uint32_t start_loci, uint32_t end_loci,
                             const uint8_t* read, int read_len, bkx_read_result* hit);

int bko_pair_reads(const bko_index* x, const bkx_align_params* p, const bkx_pe_params* pe, bkx_read_result* res,
                   uint32_t n_pairs, const uint8_t* bases, const uint64_t* offs, bkx_pe_stats* st, uint32_t* len_dist) {
  return bko_pair_reads_filtered(x, p, pe, res, n_pairs, bases, offs, st, len_dist, NULL);
}

int bko_pair_reads_filtered(const bko_index* x, const bkx_align_params* p, const bkx_pe_params* pe, bkx_read_result* res,
                            uint32_t n_pairs, const uint8_t* bases, const uint64_t* offs, bkx_pe_stats* st, uint32_t* len_dist,
                            const uint8_t* keep) {
  bkx_pe_stats z;
  memset(&z, 0, sizeof(z));
  int mode = pe->pe_proc;
  for (uint32_t i = 0; i < n_pairs; i++) {                                             /* :3107-3478 */
    bkx_read_result* f = &res[2 * i];
    bkx_read_result* r = &res[2 * i + 1];
    f->flags &= (uint8_t)~(BKX_FLG_PE_ALIGNED | BKX_FLG_PE_RECOVERED);
    r->flags &= (uint8_t)~(BKX_FLG_PE_ALIGNED | BKX_FLG_PE_RECOVERED);
    int f_un = (f->nar == BKX_NAR_NS || f->nar == BKX_NAR_NOHIT || f->nar == BKX_NAR_UNALIGNED);
    int r_un = (r->nar == BKX_NAR_NS || r->nar == BKX_NAR_NOHIT || r->nar == BKX_NAR_UNALIGNED);
    if (!(f->nar == BKX_NAR_ACCEPTED || r->nar == BKX_NAR_ACCEPTED)) { z.unaligned_pairs++; continue; }
    if (mode == BKX_PE_UNIQUE && (f_un || r_un)) {                                      /* :3125 */
      f->num_hits = r->num_hits = 0;
      f->low_hit_instances = r->low_hit_instances = 0;
      if (f->nar == BKX_NAR_ACCEPTED) f->nar = BKX_NAR_PENOHIT;
      if (r->nar == BKX_NAR_ACCEPTED) r->nar = BKX_NAR_PENOHIT;
      z.partner_unpaired++;
      continue;
    }
    if (f->nar == BKX_NAR_ACCEPTED && r->nar == BKX_NAR_ACCEPTED) {
      int frag = accept_prov_pe(pe, f, r, keep);
      if (frag > 0) {
        f->flags |= BKX_FLG_PE_ALIGNED;
        r->flags |= BKX_FLG_PE_ALIGNED;
        if (len_dist) len_dist[frag]++;
        z.accepted_num_paired++;
        continue;
      }
      switch (frag) {
        case -1: f->nar = r->nar = BKX_NAR_PESTRAND; break;
        case -2: f->nar = r->nar = BKX_NAR_PECHROM; break;
        case -6: f->nar = r->nar = BKX_NAR_PEINSERTMIN; break;
        case -7: f->nar = r->nar = BKX_NAR_PEINSERTMAX; break;
        case -3:                                                                        /* :3170 */
          z.num_filtered_by_chrom++;
          f->num_hits = r->num_hits = 0;
          f->low_hit_instances = r->low_hit_instances = 0;
          f->nar = r->nar = BKX_NAR_CHROMFILT;
          break;
        case -4: f->nar = BKX_NAR_CHROMFILT; f->low_hit_instances = 0; f->num_hits = 0; break;
        case -5: r->nar = BKX_NAR_CHROMFILT; r->low_hit_instances = 0; r->num_hits = 0; break;
        default: break; /* 0 cannot happen: both accepted => NumHits == 1 */
      }
      if (frag == -3) continue;
      if (mode == BKX_PE_UNIQUE) {                                                      /* :3203 */
        f->num_hits = r->num_hits = 0;
        f->low_hit_instances = r->low_hit_instances = 0;
        if (f->nar == BKX_NAR_ACCEPTED) f->nar = BKX_NAR_PENOHIT;
        if (r->nar == BKX_NAR_ACCEPTED) r->nar = BKX_NAR_PENOHIT;
        z.partner_unpaired++;
        continue;
      }
    }
    z.partner_unpaired++;                                                               /* :3219 */
    if (mode == BKX_PE_ORPHAN || mode == BKX_PE_ORPHAN_SE) {
      int done = 0;
      if (f->num_hits == 1 && !r_un && keep && !keep[f->chrom_id]) {                    /* :3224, :3296-3302 */
        if (f->nar == BKX_NAR_ACCEPTED) { f->num_hits = 0; f->low_hit_instances = 0; f->nar = BKX_NAR_CHROMFILT; }
      } else if (f->num_hits == 1 && !r_un) {                                           /* 5' anchor :3222 */
        int b3 = f->strand == '+';
        int anti = pe->pair_strand ? (f->strand == '+' ? 0 : 1) : (f->strand == '+' ? 1 : 0);
        if (pe->circularised) b3 = !b3;
        uint32_t os = f->match_loci, oe = f->match_loci + f->match_len - 1;
        bkx_read_result h;
        int rl = (int)(offs[2 * i + 2] - offs[2 * i + 1]);
        int rs = align_paired_read(x, pe, p, b3, anti, f->chrom_id, os, oe, bases + offs[2 * i + 1], rl, &h);
        int frag = 0;
        if (rs == 1) {
          frag = pe_insert_size(pe, f->strand, os, oe, h.strand, h.match_loci, h.match_loci + h.match_len - 1);
          if (frag <= 0) rs = 0;
        }
        if (rs == 1) {
          r->strand = h.strand; r->chrom_id = h.chrom_id; r->match_loci = h.match_loci; r->match_len = h.match_len;
          r->mismatches = h.mismatches;
          r->num_hits = 1;
          r->low_mm = (int8_t)h.mismatches;
          r->low_hit_instances = 1;
          f->flags |= BKX_FLG_PE_ALIGNED;
          r->flags |= BKX_FLG_PE_ALIGNED | BKX_FLG_PE_RECOVERED;
          f->nar = r->nar = BKX_NAR_ACCEPTED;
          if (len_dist) len_dist[frag]++;
          z.accepted_num_paired++;
          z.partner_paired++;
          done = 1;
        }
      }
      if (!done && r->num_hits == 1 && !f_un && keep && !keep[r->chrom_id]) {           /* :3323, :3411-3417: the 5' end is marked */
        if (r->nar == BKX_NAR_ACCEPTED) { f->num_hits = 0; f->low_hit_instances = 0; f->nar = BKX_NAR_CHROMFILT; }
      } else if (!done && r->num_hits == 1 && !f_un) {                                  /* 3' anchor :3321 */
        int b3 = r->strand == '+';
        int anti = r->strand == '+';
        if (pe->pair_strand) { b3 = !b3; anti = !anti; }
        if (pe->circularised) b3 = !b3;
        uint32_t os = r->match_loci, oe = r->match_loci + r->match_len - 1;
        bkx_read_result h;
        int rl = (int)(offs[2 * i + 1] - offs[2 * i]);
        int rs = align_paired_read(x, pe, p, b3, anti, r->chrom_id, os, oe, bases + offs[2 * i], rl, &h);
        int frag = 0;
        if (rs == 1) {
          frag = pe_insert_size(pe, h.strand, h.match_loci, h.match_loci + h.match_len - 1, r->strand, os, oe);
          if (frag <= 0) rs = 0;
        }
        if (rs == 1) {
          f->strand = h.strand; f->chrom_id = h.chrom_id; f->match_loci = h.match_loci; f->match_len = h.match_len;
          f->mismatches = h.mismatches;
          f->low_mm = (int8_t)h.mismatches;
          f->num_hits = 1;
          f->low_hit_instances = 1;
          f->flags |= BKX_FLG_PE_ALIGNED | BKX_FLG_PE_RECOVERED;
          r->flags |= BKX_FLG_PE_ALIGNED;
          f->nar = r->nar = BKX_NAR_ACCEPTED;
          if (len_dist) len_dist[frag]++;
          z.accepted_num_paired++;
          z.partner_paired++;
          done = 1;
        }
      }
      if (done) continue;
    }
    if (f->nar == BKX_NAR_CHROMFILT || r->nar == BKX_NAR_CHROMFILT) z.num_filtered_by_chrom++;  /* :3422 */
    if (f->nar == BKX_NAR_PEINSERTMIN || r->nar == BKX_NAR_PEINSERTMIN) z.under_len_pairs++;
    if (f->nar == BKX_NAR_PEINSERTMAX || r->nar == BKX_NAR_PEINSERTMAX) z.over_len_pairs++;
    if (!(mode == BKX_PE_ORPHAN_SE || mode == BKX_PE_UNIQUE_SE)) {
      f->num_hits = r->num_hits = 0;
      f->low_hit_instances = r->low_hit_instances = 0;
      if (f->nar == BKX_NAR_ACCEPTED) f->nar = BKX_NAR_PENOHIT;
      if (r->nar == BKX_NAR_ACCEPTED) r->nar = BKX_NAR_PENOHIT;
      continue;
    }
    /* SE fallback :3442-3477: an end counts when it has its one locus and that chromosome passes AcceptThisChromID (the
     * eNARChromFilt arm of the reference's ternary is unreachable) */
    bkx_read_result* ends[2] = {f, r};
    for (int k = 0; k < 2; k++) {
      bkx_read_result* e = ends[k];
      int ok = e->num_hits == 1 && (!keep || keep[e->chrom_id]);
      if (!ok) {
        e->num_hits = 0;
        e->low_hit_instances = 0;
        if (e->nar == BKX_NAR_ACCEPTED) e->nar = BKX_NAR_PEUNALIGN;
      } else {
        e->nar = BKX_NAR_ACCEPTED;
        z.accepted_num_se++;
      }
    }
  }
  if (st) {
    uint64_t* d = (uint64_t*)st;
    const uint64_t* s = (const uint64_t*)&z;
    for (size_t k = 0; k < sizeof(z) / 8; k++) d[k] += s[k];
  }
  return BKX_OK;
}

/* ---- orphan recovery --------------------------------------------------------------------------- */
#include "bk_oracle_pe.inc"
