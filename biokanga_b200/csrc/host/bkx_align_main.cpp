// bkx-align -- host side of the drop-in for `biokanga align`.
//
// Same option letters as the reference front-end (biokanga/kanga.cpp:194-294), same on-disk .sfx index,
// same CSV / BED / SAM / BAM records and the same alignment-summary log block
// (biokanga/Aligner.cpp:486-535, 3000-3008, 3726-3769).  The search itself goes through the C ABI of
// libbkx.so (include/bkx.h); this file holds what the reference keeps in CAligner around that call:
//   read ingest       CAligner::LoadRawReads / AddEntry   Aligner.cpp:10724-11427, 10572-10677
//                     CFasta::ReadSequence / Ascii2Sense  libbiokanga/Fasta.cpp:907-1129, 1518-1570
//   hit ordering      CAligner::SortHitMatch              Aligner.cpp:10067-10114
//   CSV rows          CAligner::WriteReadHits             Aligner.cpp:6336-6664
//   SAM records       WriteBAMReadHits / ReportBAMread    Aligner.cpp:5543-5725, 5768-6126
//                     CSAMfile::AddAlignment (text form)  libbiokanga/SAMfile.cpp:2100-2262
//   summary           CAligner::ReportAlignStats          Aligner.cpp:3493-3822
//   post-alignment    IdentifyConstraintViolations (-5)   Aligner.cpp:1245-1441, 2480-2647
//                     ReducePCRduplicates (-k)            Aligner.cpp:2184-2282
//                     PCR5PrimerCorrect (-6)              Aligner.cpp:1996-2107
//                     AutoTrimFlanks (-x)                 Aligner.cpp:1608-1812
//                     FiltByChroms (-Z / -z)              Aligner.cpp:4019-4124, 4736-4798
//                     priority regions (-B / -V)          Aligner.cpp:280-299, 9102-9186, 4126-4186
//                     adaptor trimming at load (-H)       Aligner.cpp:250-268, 11036-11084; Contaminants.cpp:204-431, 1227-1310
//                     ReportNoneAligned / ReportMultiAlign (-j / -J)   Aligner.cpp:3826-4016
//                     WriteSubDist / WriteBasicCountStats / ReportTargHitCnts (-O)   Aligner.cpp:6275-6331, 4191-4332, 5475-5537
// Written from the behaviour of those functions; no reference code is reused.  Options of the
// reference that select paths outside SURVEY section 8 (-r2 random locus, -c chimeric, -a/-A indel
// and splice, -p SNP calling, -b/-C bisulfite/SOLiD) are recognised and
// rejected with a clear message.  Output formats: CSV -M0..3, BED -M4, SAM -M5/-M6 (gzip when the name ends in
// .gz), BAM + BAI when the name ends in .bam (kanga.cpp:849-857).
#include <algorithm>
#include <unordered_map>
#include <chrono>
#include <cmath>
#include <cerrno>
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <string>
#include <memory>
#include <thread>
#include <utility>
#include <vector>

#include <glob.h>
#include <regex.h>
#include <sys/mman.h>
#include <unistd.h>
#include <zlib.h>

#include "../../../include/bkx.h"

static FILE* g_log = nullptr;
static int g_loglevel = 2;

static void diag(const char* fmt, ...) {  // CDiagnostics::DiagOut format: "[Mon DD HH:MM:SS.mmm YYYY](biokanga) text"
  char msg[4096];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(msg, sizeof(msg), fmt, ap);
  va_end(ap);
  auto now = std::chrono::system_clock::now();
  time_t t = std::chrono::system_clock::to_time_t(now);
  int ms = (int)(std::chrono::duration_cast<std::chrono::milliseconds>(now.time_since_epoch()).count() % 1000);
  struct tm tmv;
  localtime_r(&t, &tmv);
  char ts[64], yr[8];
  strftime(ts, sizeof(ts), "%b %d %H:%M:%S", &tmv);
  strftime(yr, sizeof(yr), "%Y", &tmv);
  char line[4400];
  snprintf(line, sizeof(line), "[%s.%03d %s](biokanga) %s\n", ts, ms, yr, msg);
  fputs(line, stdout);
  if (g_log) { fputs(line, g_log); fflush(g_log); }
}

// Buffers of several GB (file text, read arena) must not be value-initialised by one thread: resize() leaves new
// elements untouched, the threads that fill them take the page faults.
template <class T>
struct NoInit : std::allocator<T> {
  template <class U> struct rebind { using other = NoInit<U>; };
  template <class U, class... A>
  void construct(U* p, A&&... a) {
    if constexpr (sizeof...(A) == 0) ::new ((void*)p) U; else ::new ((void*)p) U(std::forward<A>(a)...);
  }
};
template <class T> using Buf = std::vector<T, NoInit<T>>;

// A buffer that crosses PCIe: page-locked by the library from the start (bkx_alloc_host), plain memory if that fails.
template <class T>
struct PinnedBuf {
  T* p = nullptr;
  size_t n = 0;
  bool pinned = false;
  PinnedBuf() = default;
  PinnedBuf(const PinnedBuf&) = delete;
  PinnedBuf& operator=(const PinnedBuf&) = delete;
  ~PinnedBuf() { release(); }
  void release() {
    if (p) { if (pinned) bkx_free_host(p); else free(p); }
    p = nullptr; n = 0;
  }
  bool resize(size_t count) {   // contents are not kept
    release();
    p = (T*)bkx_alloc_host(count * sizeof(T));
    pinned = p != nullptr;
    if (!p) p = (T*)malloc(std::max<size_t>(count * sizeof(T), 1));
    n = p ? count : 0;
    return p != nullptr;
  }
  T* data() { return p; }
  const T* data() const { return p; }
  size_t size() const { return n; }
  T& operator[](size_t i) { return p[i]; }
  const T& operator[](size_t i) const { return p[i]; }
};
using ResVec = Buf<bkx_read_result>;

// ---- read ingest ---------------------------------------------------------------------------------
struct Reads {
  // what crosses PCIe (the library's compact host interface): the bases 2 bits each (base i at bits [2(i%4), +2) of byte i/4,
  // non-ACGT bases as 0 and listed in exc_*), read lengths (empty when all reads are equally long)
  PinnedBuf<uint8_t> packed2;
  std::vector<uint64_t> exc_pos;
  std::vector<uint8_t> exc_code;
  Buf<uint16_t> lens;
  uint32_t fixed_len = 0;
  PinnedBuf<bkx_read_result16> res16;   // the records of the compact interface, allocated along with the stream
  Buf<uint8_t> bases;             // 1 byte/base, etSeqBase code in the low 3 bits
  Buf<uint64_t> offs = Buf<uint64_t>(1, 0);
  Buf<char> names;                // NUL-terminated descriptors, back to back
  Buf<uint64_t> name_ofs;
  uint32_t n() const { return (uint32_t)(offs.size() - 1); }
  const char* name(uint32_t i) const { return names.data() + name_ofs[i]; }
  int len(uint32_t i) const { return (int)(offs[i + 1] - offs[i]); }
};

static inline uint8_t base_code(char c) {  // CFasta::Ascii2Sense, Fasta.cpp:1518-1570 (soft-mask bit dropped)
  switch (c) {
    case 'a': case 'A': return 0;
    case 'c': case 'C': return 1;
    case 'g': case 'G': return 2;
    case 't': case 'T': case 'u': case 'U': return 3;
    case '-': return 6;
    default: return 4;
  }
}

// 4-bit quality packed next to the base (bits 4..7), CAligner::LoadRawReads, Aligner.cpp:11130-11196
static inline unsigned qual4(int qmode, unsigned c) {
  unsigned q;
  switch (qmode) {
    case 0:  // Sanger / Illumina 1.8+
      if (c < 33) c = 33; else if (c >= 126) c = 125;
      q = c - 33;
      break;
    case 1:  // Illumina 1.3+
      if (c < 64) c = 64; else if (c >= 126) c = 125;
      q = c - 64;
      break;
    default:  // Solexa < 1.3
      if (c < 59 || c >= 126) c = c < 64 ? 64 : 125;
      q = c - 59;
      q = (uint8_t)(10 * log(1 + pow(10.0, ((double)q / 10.0) / log(10.0))));
      break;
  }
  if (q > 40) q = 40;
  return ((q + 2) * 15) / 40;
}

struct Opts {
  int pmode = 0, strand = 0, max_subs = 10, edit_delta = 1, max_ns = 1, fmt = 5, pe_mode = 0, pair_min = 100,
      pair_max = 1000, trim5 = 0, trim3 = 0, min_len = 50, max_len = 500, threads = 0, gpus = 1, sam_seq_thres = 10000,
      qmode = 3;
  int ml_mode = 0, max_ml = 0;   // -r / -R (kanga.cpp:482-486, 667-696); max_ml 0 = not given
  bool clamp_ml = false;         // -X
  bool best_matches = false;     // -N: the -R loci with the fewest mismatches from one un-staged pass (only with -r; implies -X)
  bool pair_strand = false, pe_circ = false;
  int sample_nth = 1;            // -#: sample every Nth raw read or read pair
  int pcr_primer = 0;            // -6: align with -s + this many substitutions, then correct 5' primer artefacts back to -s
  int pcr_win = -1;              // -k: PCR artefact reduction window, -1 = off (kanga.cpp:719-724)
  int min_flank = 0;             // -x: auto-trim flanks back to this many exactly matching bases (kanga.cpp:497, 804)
  std::vector<std::string> excl, incl;   // -Z / -z chromosome filters (POSIX extended, case insensitive)
  std::string constraints_file;          // -5: loci base constraints CSV (chrom, start, end, bases)
  std::string contam_file;               // -H: adaptor (contaminant) sequences trimmed off the read ends at load
  std::string priority_file;             // -B: BED file of priority regions (reads with one locus in them are taken as unique)
  bool priority_nofilt = false;          // -V: keep the accepted alignments outside the priority regions
  std::string stats_file;                // -O: substitution / quality / multi-hit / insert-length distributions (CSV)
  std::string none_file, multi_file;     // -j / -J: FASTA of the reads without a locus / with too many loci
  std::vector<std::string> in, pair;
  std::string sfx, out, logfile, title;
  std::string sqlite_file, exp_name, exp_descr;   // -q / -w / -W: SQLite results-summary database (validated, not written)
};

// ---- read files: slurped whole (gzip inflated on the way), split at record boundaries, parsed by all host threads.
struct RawRec {        // one FASTA / FASTQ record as pointers into the file text
  const char* name;    // descriptor: text after '>' / '@' up to the first white space, <= 127 chars
  const char* sb;      // sequence region [sb, se); may hold line breaks (multi-line FASTA)
  const char* se;
  const char* q;       // FASTQ quality line or nullptr
  uint32_t name_n, len, qlen;
};

template <class Text>
static bool slurp(const std::string& path, Text& text, unsigned threads) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  unsigned char magic[2] = {0, 0};
  size_t got = fread(magic, 1, 2, f);
  fseeko(f, 0, SEEK_END);
  size_t size = (size_t)ftello(f);
  fclose(f);
  if (got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) {
    gzFile g = gzopen(path.c_str(), "rb");
    if (!g) return false;
    gzbuffer(g, 1 << 20);
    text.resize(std::max<size_t>(size * 4, (size_t)1 << 20));
    size_t n = 0;
    for (;;) {
      if (n == text.size()) text.resize(text.size() * 2);
      int r = gzread(g, text.data() + n, (unsigned)std::min<size_t>(text.size() - n, (size_t)1 << 30));
      if (r < 0) { gzclose(g); return false; }
      if (r == 0) break;
      n += (size_t)r;
    }
    gzclose(g);
    text.resize(n);
    return true;
  }
  text.resize(size);
  if (size == 0) return true;
  threads = std::max(1u, std::min(threads, (unsigned)((size >> 24) + 1)));
  std::vector<std::thread> th;
  std::vector<char> okv(threads, 1);
  for (unsigned t = 0; t < threads; ++t)
    th.emplace_back([&, t]() {
      size_t b = size * t / threads, e = size * (t + 1) / threads;
      FILE* h = fopen(path.c_str(), "rb");
      if (!h) { okv[t] = 0; return; }
      fseeko(h, (off_t)b, SEEK_SET);
      while (b < e) {
        size_t r = fread(text.data() + b, 1, std::min<size_t>(e - b, (size_t)64 << 20), h);
        if (r == 0) { okv[t] = 0; break; }
        b += r;
      }
      fclose(h);
    });
  for (auto& x : th) x.join();
  for (char k : okv) if (!k) return false;
  return true;
}

static inline const char* line_end(const char* p, const char* end) {
  const char* nl = (const char*)memchr(p, '\n', (size_t)(end - p));
  return nl ? nl : end;
}
static inline const char* next_line(const char* eol, const char* end) { return eol < end ? eol + 1 : end; }

// first record start at or after `from` (a line start): FASTA '>' lines; FASTQ '@' lines whose line + 2 starts with '+'
static const char* record_start(const char* text, const char* from, const char* end, bool fastq) {
  const char* p = from;
  if (p > text && p[-1] != '\n') p = next_line(line_end(p, end), end);
  while (p < end) {
    if (!fastq) {
      if (*p == '>') return p;
    } else if (*p == '@') {
      const char* l1 = next_line(line_end(p, end), end);
      const char* l2 = next_line(line_end(l1, end), end);
      if (l2 < end && *l2 == '+' && (l1 >= end || *l1 != '@')) return p;
    }
    p = next_line(line_end(p, end), end);
  }
  return end;
}

// parse the records that START inside [b, e); returns false at a malformed record (everything before it is kept)
static bool parse_records(const char* b, const char* e, const char* end, bool fastq, std::vector<RawRec>& out) {
  const char* p = b;
  auto trim_eol = [](const char* s, const char* t) { while (t > s && (t[-1] == '\r' || t[-1] == '\n')) --t; return t; };
  while (p < e) {
    const char* le = line_end(p, end);
    const char* lt = trim_eol(p, le);
    if (lt == p) { p = next_line(le, end); continue; }  // blank line
    if (*p != (fastq ? '@' : '>')) return false;
    RawRec r;
    const char* d = p + 1;
    const char* de = d;
    while (de < lt && !isspace((unsigned char)*de) && de - d < 127) ++de;
    r.name = d;
    r.name_n = (uint32_t)(de - d);
    r.q = nullptr;
    r.qlen = 0;
    p = next_line(le, end);
    if (fastq) {
      if (p >= end) return true;  // header without a sequence line: end of data
      const char* se = line_end(p, end);
      r.sb = p;
      r.se = trim_eol(p, se);
      r.len = (uint32_t)(r.se - r.sb);
      p = next_line(se, end);
      p = next_line(line_end(p, end), end);  // '+' line
      if (p < end) {
        const char* qe = line_end(p, end);
        r.q = p;
        r.qlen = (uint32_t)(trim_eol(p, qe) - p);
        p = next_line(qe, end);
      }
    } else {
      r.sb = p;
      uint32_t len = 0;
      while (p < end && *p != '>') {
        const char* se = line_end(p, end);
        len += (uint32_t)(trim_eol(p, se) - p);
        p = next_line(se, end);
      }
      r.se = p;
      r.len = len;
    }
    out.push_back(r);
  }
  return true;
}

static bool parse_file(const std::string& path, unsigned threads, Buf<char>& text, std::vector<RawRec>& recs) {
  if (!slurp(path, text, threads)) return false;
  const char* b = text.data();
  const char* end = b + text.size();
  const char* first = b;
  while (first < end && (*first == '\n' || *first == '\r')) ++first;
  if (first == end) return true;
  if (*first != '>' && *first != '@') return true;  // not a sequence file: no records (as the line reader behaved)
  const bool fastq = *first == '@';
  size_t min_chunk = (size_t)4 << 20;
  if (const char* ev = getenv("BKX_PARSE_MIN_CHUNK")) min_chunk = std::max<size_t>(64, (size_t)atoll(ev));  // test hook
  unsigned T = std::max(1u, std::min(threads, (unsigned)(text.size() / min_chunk + 1)));
  std::vector<const char*> cut(T + 1, end);
  cut[0] = first;
  for (unsigned t = 1; t < T; ++t) cut[t] = record_start(b, b + text.size() * t / T, end, fastq);
  for (unsigned t = 1; t <= T; ++t) if (cut[t] < cut[t - 1]) cut[t] = cut[t - 1];
  std::vector<std::vector<RawRec>> part(T);
  std::vector<char> good(T, 1);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < T; ++t)
    th.emplace_back([&, t]() { good[t] = parse_records(cut[t], cut[t + 1], end, fastq, part[t]) ? 1 : 0; });
  for (auto& x : th) x.join();
  size_t total = 0;
  for (unsigned t = 0; t < T; ++t) { total += part[t].size(); if (!good[t]) break; }
  recs.reserve(total);
  for (unsigned t = 0; t < T; ++t) {
    recs.insert(recs.end(), part[t].begin(), part[t].end());
    if (!good[t]) break;  // a malformed record ends the file there
  }
  return true;
}

// Single-end runs: every -i is a file specification that may hold wildcards (CSimpleGlob with SG_GLOB_FULLSORT,
// ProcLoadReadFiles, Aligner.cpp:10474-10520): the matches of one specification are loaded in case-insensitive name
// order; a specification without a match ends the run.  Paired-end file names are taken as they are.
static int expand_input_specs(Opts& o) {
  if (o.pe_mode) return 0;
  std::vector<std::string> files;
  for (const std::string& spec : o.in) {
    std::vector<std::string> found;
    if (spec.find('*') == std::string::npos && spec.find('?') == std::string::npos) {
      if (access(spec.c_str(), F_OK) == 0) found.push_back(spec);
    } else {
      glob_t g;
      memset(&g, 0, sizeof(g));
      if (glob(spec.c_str(), GLOB_MARK | GLOB_NOSORT, nullptr, &g) == 0)
        for (size_t k = 0; k < g.gl_pathc; ++k) found.push_back(g.gl_pathv[k]);
      globfree(&g);
      std::sort(found.begin(), found.end(), [](const std::string& a, const std::string& b) { return strcasecmp(a.c_str(), b.c_str()) < 0; });
    }
    if (found.empty()) { diag("Unable to glob '%s", spec.c_str()); return -1; }
    files.insert(files.end(), found.begin(), found.end());
  }
  o.in.swap(files);
  return 0;
}

// ---- -H: adaptor sequences ("contaminants") overlapping the read ends are trimmed off at load (Aligner.cpp:250-268,
//      11036-11084, 11283-11293; CContaminants::LoadContaminantsFile / MatchContaminants, Contaminants.cpp:204-431, 1227-1310).
//      A multi-FASTA file; the name of a sequence may end in '@' + digits saying which read ends it is tried on: 1 / 2 the 5'
//      end of SE-PE1 / PE2 reads, 3 / 4 their 3' end, 5..8 the same four with the sequence reverse complemented; no suffix
//      means 1, 2, 5 and 6.  On a 5' end the read's first L bases are compared with the LAST L bases of a sequence, on a 3'
//      end the read's last L bases with its FIRST L; the longest L (down from min(read, longest sequence) to the fixed
//      trim + 1) at which some sequence fits with at most one substitution is trimmed -- one substitution is granted at
//      every length, so a read loses at least one base at every end some sequence is tried on.  An N in the read never
//      matches, an N in the sequence always does.  The reference walks a trie over the sequences (RecursiveMatch); the
//      sequences are few and short, so here they are simply tried one by one.
struct Contaminants {
  std::vector<std::vector<uint8_t>> seqs[4];   // [0] 5' SE/PE1, [1] 5' PE2, [2] 3' SE/PE1, [3] 3' PE2
  int max_len[4] = {0, 0, 0, 0};
  // '&' entries: vector sequences a read may lie INSIDE, as it is (sense) or reverse complemented (antisense); a read that
  // does is reported as overlapped over its whole length and so falls to the length filter (MatchVectContams /
  // MatchVectContam, Contaminants.cpp:1038-1225).  Found through the 12-mers of the vector (the reference seeds with
  // windows of read length / (allowed mismatches + 1) >= 13 bases through a suffix array of the vector).
  struct Vector {
    std::vector<uint8_t> seq;
    bool pe1_sense = false, pe1_anti = false, pe2_sense = false, pe2_anti = false;
    std::unordered_map<uint32_t, std::vector<uint32_t>> kmers;   // 12-mer without N -> positions
  };
  std::vector<Vector> vectors;
  bool loaded = false;
  static constexpr int kSeed = 12;
  static bool seed_key(const uint8_t* p, uint32_t& key) {
    key = 0;
    for (int i = 0; i < kSeed; ++i) {
      const uint8_t b = p[i] & 7;
      if (b > 3) return false;
      key = (key << 2) | b;
    }
    return true;
  }
  // does the vector hold the query with at most `allowed` substitutions (an N only equals an N)
  static bool inside(const Vector& v, const uint8_t* q, int qlen, int allowed) {
    const int n = (int)v.seq.size();
    auto fits = [&](long p) {
      if (p < 0 || p + qlen > n) return false;
      int mm = 0;
      for (int i = 0; i < qlen; ++i)
        if ((q[i] & 7) != v.seq[(size_t)p + (size_t)i] && ++mm > allowed) return false;
      return true;
    };
    const int win = qlen > 25 ? qlen / (allowed + 1) : qlen;
    for (int ofs = 0; ofs < qlen; ofs += win) {
      if (ofs + win > qlen) ofs = qlen - win;
      // any 12-mer of an exactly matching window matches exactly as well: take the first one without an N
      bool seeded = false;
      for (int j = 0; j + kSeed <= win && !seeded; ++j) {
        uint32_t key;
        if (!seed_key(q + ofs + j, key)) continue;
        seeded = true;
        auto it = v.kmers.find(key);
        if (it == v.kmers.end()) break;
        for (uint32_t hit : it->second)
          if (fits((long)hit - (long)(ofs + j))) return true;
      }
      if (!seeded)   // a window of Ns: every placement
        for (long p = 0; p + qlen <= n; ++p) if (fits(p)) return true;
      if (ofs + win >= qlen) break;
    }
    return false;
  }
  int match_vectors(bool pe2_flags, const uint8_t* q, int qlen) const {
    if (vectors.empty() || qlen < 20 || qlen > 2000) return 0;
    const int allowed = qlen > 25 ? qlen / 25 : 0;
    std::vector<uint8_t> rc;
    for (const auto& v : vectors) {
      const bool sense = pe2_flags ? v.pe2_sense : v.pe1_sense, anti = pe2_flags ? v.pe2_anti : v.pe1_anti;
      if (!(sense || anti) || (int)v.seq.size() < qlen) continue;
      if (sense && inside(v, q, qlen, allowed)) return qlen;
      if (anti) {
        if (rc.empty()) {
          rc.resize((size_t)qlen);
          for (int i = 0; i < qlen; ++i) { const uint8_t b = q[qlen - 1 - i] & 7; rc[(size_t)i] = b < 4 ? (uint8_t)(3 - b) : b; }
        }
        if (inside(v, rc.data(), qlen, allowed)) return qlen;
      }
    }
    return 0;
  }
  int match(int type, const uint8_t* q, int qlen, int min_overlap) const {
    if (qlen < 20 || qlen > 2000) return 0;   // cMinContamQuerySeqLen .. cMaxContamQuerySeqLen
    // vectors first; the reference picks the PE1 flags for the 5' end of SE / PE1 reads only -- its test for the 3' end
    // names the 5' type twice (Contaminants.cpp:1257) -- and the PE2 flags for the other three
    if (const int whole = match_vectors(type != 0, q, qlen)) return whole;
    if (seqs[type].empty()) return 0;
    if (min_overlap < 1) min_overlap = 1;
    const int cur = std::min(qlen, max_len[type]);
    const bool five = type < 2;
    for (int L = cur; L >= min_overlap; --L)
      for (const auto& c : seqs[type]) {
        if ((int)c.size() < L) continue;
        const uint8_t* qs = five ? q : q + qlen - L;
        const uint8_t* cs = five ? c.data() + c.size() - (size_t)L : c.data();
        int mm = 0;
        for (int i = 0; i < L && mm <= 1; ++i) {
          const uint8_t tb = qs[i] & 7, cb = cs[i];
          if (tb == 4 || (cb != 4 && tb != cb)) ++mm;
        }
        if (mm <= 1) return L;
      }
    return 0;
  }
};
static Contaminants g_contam;

static int load_contaminants(const std::string& path, Contaminants& C) {
  FILE* fp = fopen(path.c_str(), "r");
  if (!fp) { diag("Unable to load contaminate sequences file '%s'", path.c_str()); return -1; }
  diag("LoadContaminantsFile:- Processing %s..", path.c_str());
  std::string text;
  char buf[65536];
  size_t got;
  while ((got = fread(buf, 1, sizeof(buf), fp)) > 0) text.append(buf, got);
  fclose(fp);
  bool flags[8] = {true, true, false, false, true, true, false, false};   // a file without any descriptor: 1, 2, 5, 6
  bool vector_type = false;
  std::string name = "ContamSeq.1";
  std::vector<uint8_t> seq;
  int n_seqs = 0;
  bool ok = true;
  auto flush = [&]() {
    if (seq.empty()) return;
    ++n_seqs;
    if (vector_type) {
      if (seq.size() < 100 || seq.size() > 0x0ffffff) {   // cMinVectorSeqLen .. cMaxVectorSeqLen
        diag("LoadContamiantsFile: Vector sequence for '%s' outside of accepted length range %d..%d", name.c_str(), 100, 0x0ffffff);
        ok = false;
        return;
      }
      if (C.vectors.size() >= 10) {   // cMaxNumVectors
        diag("AddVectContam: Too many vector contaminants (max allowed %d) current contaminant is '%s'", 10, name.c_str());
        ok = false;
        return;
      }
      Contaminants::Vector v;
      v.seq = seq;
      v.pe1_sense = flags[0] || flags[2]; v.pe1_anti = flags[4] || flags[6];
      v.pe2_sense = flags[1] || flags[3]; v.pe2_anti = flags[5] || flags[7];
      for (size_t i = 0; i + Contaminants::kSeed <= v.seq.size(); ++i) {
        uint32_t key;
        if (Contaminants::seed_key(v.seq.data() + i, key)) v.kmers[key].push_back((uint32_t)i);
      }
      C.vectors.push_back(std::move(v));
      seq.clear();
      return;
    }
    if (seq.size() < 4 || seq.size() > 200) {   // cMinContaminantLen .. cMaxContaminantLen
      diag("LoadContamiantsFile: Sequence for '%s' outside of accepted length range %d..%d", name.c_str(), 4, 200);
      ok = false;
      return;
    }
    for (int t = 0; t < 4; ++t) if (flags[t]) { C.seqs[t].push_back(seq); C.max_len[t] = std::max(C.max_len[t], (int)seq.size()); }
    if (flags[4] || flags[5] || flags[6] || flags[7]) {
      std::vector<uint8_t> rc(seq.rbegin(), seq.rend());
      for (auto& b : rc) if (b < 4) b = (uint8_t)(3 - b);
      for (int t = 0; t < 4; ++t) if (flags[4 + t]) { C.seqs[t].push_back(rc); C.max_len[t] = std::max(C.max_len[t], (int)rc.size()); }
    }
    seq.clear();
  };
  size_t pos = 0;
  while (pos < text.size() && ok) {
    size_t eol = text.find('\n', pos);
    if (eol == std::string::npos) eol = text.size();
    std::string line = text.substr(pos, eol - pos);
    pos = eol + 1;
    while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
    if (line.empty()) continue;
    if (line[0] == '>') {
      flush();
      if (!ok) break;
      for (bool& f : flags) f = false;
      vector_type = false;
      size_t a = 1;
      while (a < line.size() && isspace((unsigned char)line[a])) ++a;
      size_t b = a;
      while (b < line.size() && !isspace((unsigned char)line[b])) ++b;
      name = line.substr(a, b - a);
      bool coded = false;
      if (name.empty()) {
        name = "ContamSeq." + std::to_string(n_seqs + 1);
      } else {
        // overlay codes: the digits 1..8 at the very end of the name, behind an '@' (adaptor) or '&' (vector)
        size_t k = name.size() - 1;
        for (size_t left = name.size(); left > 1; --left, --k)
          if (name[k] == '@' || name[k] == '&' || !(name[k] >= '1' && name[k] <= '8')) break;
        vector_type = name[k] == '&';
        if ((name[k] == '@' || name[k] == '&') && k + 1 < name.size()) {
          for (size_t i = k + 1; i < name.size(); ++i) flags[name[i] - '1'] = true;
          name.resize(k);
          coded = true;
        }
      }
      if (!coded) flags[0] = flags[1] = flags[4] = flags[5] = true;
      continue;
    }
    for (char ch : line) {
      if (isspace((unsigned char)ch)) continue;
      const uint8_t c = base_code(ch);
      if (c > 4) { diag("LoadContaminantsFile: Illegal base in %s sequence, only bases A,C,G,T,N accepted", name.c_str()); ok = false; break; }
      seq.push_back(c);
    }
  }
  if (ok) flush();
  if (!ok) { diag("Unable to load contaminate sequences file '%s'", path.c_str()); return -1; }
  if (n_seqs == 0) diag("No contaminant sequences loaded from '%s'", path.c_str());
  C.loaded = true;
  return 0;
}

static int load_reads(const Opts& o, Reads& R) {  // Aligner.cpp:10724-11427 (default -g3: qualities ignored)
  bool pe = o.pe_mode != 0;
  uint8_t code_tab[256];
  for (int c = 0; c < 256; ++c) code_tab[c] = base_code((char)c);
  const unsigned T = o.threads > 0 ? (unsigned)o.threads : std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  diag("Loading reads from file...");
  for (size_t fi = 0; fi < o.in.size(); ++fi) {
    Buf<char> text[2];
    std::vector<RawRec> recs[2];
    bool opened[2] = {true, true};
    {
      std::thread t2;
      if (pe) t2 = std::thread([&]() { opened[1] = parse_file(o.pair[fi], std::max(1u, T / 2), text[1], recs[1]); });
      opened[0] = parse_file(o.in[fi], pe ? std::max(1u, T / 2) : T, text[0], recs[0]);
      if (pe) t2.join();
    }
    if (!opened[0]) { diag("Unable to open '%s'", o.in[fi].c_str()); return -1; }
    if (pe && !opened[1]) { diag("Unable to open '%s'", o.pair[fi].c_str()); return -1; }
    const size_t nrec = recs[0].size();
    // -H: bases an adaptor overlaps at the 5' / 3' end of every read that is going to be looked at, beyond the fixed trims
    // (Aligner.cpp:11036-11084).  Computed on the raw read, by all threads, before the length filter that counts them in.
    std::vector<uint16_t> c5[2], c3[2];
    if (g_contam.loaded) {
      for (int f = 0; f < (pe ? 2 : 1); ++f) { c5[f].assign(recs[f].size(), 0); c3[f].assign(recs[f].size(), 0); }
      std::vector<std::thread> th;
      for (unsigned t = 0; t < T; ++t)
        th.emplace_back([&, t]() {
          std::vector<uint8_t> q;
          for (int f = 0; f < (pe ? 2 : 1); ++f) {
            const size_t nr = recs[f].size(), b = nr * t / T, e = nr * (t + 1) / T;
            for (size_t i = b; i < e; ++i) {
              if (o.sample_nth > 1 && i % (size_t)o.sample_nth) continue;
              const RawRec& r = recs[f][i];
              q.clear();
              for (const char* c = r.sb; c < r.se; ++c) if (*c != '\n' && *c != '\r') q.push_back(code_tab[(unsigned char)*c]);
              int a5 = g_contam.match(f, q.data(), (int)q.size(), o.trim5 + 1);
              int a3 = g_contam.match(2 + f, q.data(), (int)q.size(), o.trim3 + 1);
              a5 = a5 <= o.trim5 ? 0 : a5 - o.trim5;
              a3 = a3 <= o.trim3 ? 0 : a3 - o.trim3;
              c5[f][i] = (uint16_t)a5;
              c3[f][i] = (uint16_t)a3;
            }
          }
        });
      for (auto& x : th) x.join();
    }
    auto contam5 = [&](int f, size_t i) -> int { return c5[f].empty() ? 0 : c5[f][i]; };
    auto contam3 = [&](int f, size_t i) -> int { return c3[f].empty() ? 0 : c3[f][i]; };
    // length filter, pairwise for PE (Aligner.cpp:10895-10935, 11086-11121)
    uint32_t accepted = 0, under = 0, over = 0;
    uint32_t n_c5[2] = {0, 0}, n_c3[2] = {0, 0};
    std::vector<uint8_t> keep(nrec, 0);
    auto bad_len = [&](uint32_t len, int extra, uint32_t& u, uint32_t& ov) {
      if (o.trim5 + o.trim3 + extra + o.min_len > (int)len) { ++u; return true; }
      if (o.trim5 + o.trim3 + extra + o.max_len < (int)len) { ++ov; return true; }
      return false;
    };
    for (size_t i = 0; i < nrec; ++i) {
      if (pe && i >= recs[1].size()) { diag("Problem parsing sequence after %u reads parsed", accepted); return -1; }
      if (o.sample_nth > 1 && i % (size_t)o.sample_nth) continue;   // -#: every Nth raw read / pair of a file, from its first (Aligner.cpp:10943, 11027-11033)
      if (bad_len(recs[0][i].len, contam5(0, i) + contam3(0, i), under, over)) continue;
      if (pe && bad_len(recs[1][i].len, contam5(1, i) + contam3(1, i), under, over)) continue;
      keep[i] = 1;
      ++accepted;
      for (int f = 0; f < (pe ? 2 : 1); ++f) { n_c5[f] += contam5(f, i) > 0; n_c3[f] += contam3(f, i) > 0; }
    }
    // offsets of every accepted read in the arena, then a parallel fill
    const int per = pe ? 2 : 1;
    const size_t first_read = R.offs.size() - 1, add_reads = (size_t)accepted * per;
    R.offs.resize(first_read + add_reads + 1);
    R.name_ofs.resize(first_read + add_reads);
    std::vector<size_t> src(add_reads);  // accepted read -> record index (file = read index & 1 for PE)
    uint64_t bo = R.bases.size(), no = R.names.size();
    size_t w = first_read;
    const size_t cut = (size_t)(o.trim5 + o.trim3);
    for (size_t i = 0; i < nrec; ++i) {
      if (!keep[i]) continue;
      for (int f = 0; f < per; ++f) {
        const RawRec& r = recs[f][i];
        R.offs[w] = bo;
        R.name_ofs[w] = no;
        src[w - first_read] = i;
        bo += r.len - cut - (size_t)contam5(f, i) - (size_t)contam3(f, i);
        no += r.name_n + 1;
        ++w;
      }
    }
    R.offs[w] = bo;
    R.bases.resize(bo);
    R.names.resize(no);
    {
      std::vector<std::thread> th;
      for (unsigned t = 0; t < T; ++t)
        th.emplace_back([&, t]() {
          size_t b = add_reads * t / T, e = add_reads * (t + 1) / T;
          std::vector<uint8_t> tmp;
          for (size_t k = b; k < e; ++k) {
            const int fk = pe ? (int)(k & 1) : 0;
            const RawRec& r = recs[fk][src[k]];
            uint8_t* dst = R.bases.data() + R.offs[first_read + k];
            const size_t lead = (size_t)o.trim5 + (size_t)contam5(fk, src[k]);   // fixed 5' trim + what an adaptor overlaps (-H)
            const size_t L = r.len - cut - (size_t)contam5(fk, src[k]) - (size_t)contam3(fk, src[k]);
            const bool useq = o.qmode != 3 && r.q && r.qlen == r.len;
            if ((size_t)(r.se - r.sb) == r.len) {  // one line
              const char* sp = r.sb + lead;
              if (useq) {
                const char* qp = r.q + lead;
                for (size_t i = 0; i < L; ++i)
                  dst[i] = (uint8_t)(code_tab[(unsigned char)sp[i]] | (qual4(o.qmode, (unsigned char)qp[i]) << 4));
              } else {
                for (size_t i = 0; i < L; ++i) dst[i] = code_tab[(unsigned char)sp[i]];
              }
            } else {  // multi-line FASTA: drop the line breaks first
              tmp.clear();
              for (const char* c = r.sb; c < r.se; ++c) if (*c != '\n' && *c != '\r') tmp.push_back(code_tab[(unsigned char)*c]);
              memcpy(dst, tmp.data() + lead, L);
            }
            char* nm = R.names.data() + R.name_ofs[first_read + k];
            memcpy(nm, r.name, r.name_n);
            nm[r.name_n] = '\0';
          }
        });
      for (auto& x : th) x.join();
    }
    diag("LoadReads: Total of %1.9d reads parsed and loaded from %s", accepted, o.in[fi].c_str());
    if (under) diag("Load: total of %d under length sequences sloughed from file '%s'", under, o.in[fi].c_str());
    if (over) diag("Load: total of %d over length sequences sloughed from file '%s'", over, o.in[fi].c_str());
    if (g_contam.loaded) {   // Aligner.cpp:11403-11425 (the PE2 lines say "PE1" there too)
      diag("Load: total of %d sequences PE1 sequences were 5' contaminate trimmed", (int)n_c5[0]);
      diag("Load: total of %d sequences PE1 sequences were 3' contaminate trimmed", (int)n_c3[0]);
      if (pe) {
        diag("Load: total of %d sequences PE1 sequences were 5' contaminant trimmed", (int)n_c5[1]);
        diag("Load: total of %d sequences PE1 sequences were 3' contaminant trimmed", (int)n_c3[1]);
      }
    }
  }
  if (o.ml_mode < 3) {  // the 2-bit stream + exception list + lengths for the H2D copies (the multi-loci calls -r3..5 take the
                        // one-byte-per-base arena); threads own disjoint byte ranges of the stream
    const size_t nb = R.bases.size(), np = (nb + 3) / 4;
    if (!R.packed2.resize(np + 16) || !R.res16.resize(R.n())) { diag("Fatal: out of memory"); return -1; }
    std::vector<std::vector<uint64_t>> epos(T);
    std::vector<std::vector<uint8_t>> ecode(T);
    const uint32_t nreads = R.n();
    R.lens.resize(nreads);
    std::vector<uint8_t> uniform(T, 1);
    const uint32_t len0 = nreads ? (uint32_t)R.len(0) : 0;
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t)
      th.emplace_back([&, t]() {
        size_t b = np * t / T, e = np * (t + 1) / T;
        if (t + 1 == T) e = np + 16;   // the slack the device copies may read
        const uint8_t* src = R.bases.data();
        for (size_t i = b; i < e; ++i) {
          unsigned v = 0;
          for (unsigned j = 0; j < 4; ++j) {
            const size_t q = 4 * i + j;
            if (q >= nb) break;
            const unsigned c = src[q] & 7u;
            if (c > 3) { epos[t].push_back(q); ecode[t].push_back((uint8_t)c); } else v |= c << (2 * j);
          }
          R.packed2[i] = (uint8_t)v;
        }
        const uint32_t rb = (uint32_t)((uint64_t)nreads * t / T), re = (uint32_t)((uint64_t)nreads * (t + 1) / T);
        for (uint32_t r = rb; r < re; ++r) {
          const uint32_t l = (uint32_t)R.len(r);
          R.lens[r] = (uint16_t)l;
          if (l != len0) uniform[t] = 0;
        }
      });
    for (auto& x : th) x.join();
    for (unsigned t = 0; t < T; ++t) {
      R.exc_pos.insert(R.exc_pos.end(), epos[t].begin(), epos[t].end());
      R.exc_code.insert(R.exc_code.end(), ecode[t].begin(), ecode[t].end());
    }
    bool all_same = nreads > 0;
    for (unsigned t = 0; t < T; ++t) all_same = all_same && uniform[t];
    R.fixed_len = all_same ? len0 : 0;
  }
  return 0;
}

// ---- option parsing: the option table of kanga.cpp:194-294 -- short letters (values attached or separate, literal flags
//      may be bundled), the long names (--name=value or --name value), and parameter files: an argument "@file" is replaced
//      by the options read from that file, one or more per line, white space separated unless quoted, lines starting with
//      '#', ';' or "//" skipped (CUtility::arg_parsefromfile, libbiokanga/Utility.cpp:793-912; called at kanga.cpp:298).
static bool takes_value(char c) { return strchr("fFqwWm#QcaAkgryYlLR4esnx6pKGP1MtBiUdDuISo7jJO89H5ZzT", c) != nullptr; }

static const struct { const char* name; char letter; } kLongOpts[] = {
    {"help", 'h'}, {"version", 'v'}, {"ver", 'v'}, {"FileLogLevel", 'f'}, {"log", 'F'}, {"mode", 'm'}, {"format", 'M'}, {"pemode", 'U'},
    {"bisulfite", 'b'}, {"colorspace", 'C'}, {"pcrwin", 'k'}, {"in", 'i'}, {"pair", 'u'}, {"priorityregionfile", 'B'},
    {"nofiltpriority", 'V'}, {"pairminlen", 'd'}, {"pairmaxlen", 'D'}, {"pairstrand", 'E'}, {"alignstrand", 'Q'},
    {"samplenthrawread", '#'}, {"minchimeric", 'c'}, {"chimericrpt", '0'}, {"quality", 'g'}, {"sfx", 'I'}, {"out", 'o'},
    {"microindellen", 'a'}, {"splicejunctlen", 'A'}, {"stats", 'O'}, {"nonealign", 'j'}, {"multialign", 'J'}, {"siteprefs", '8'},
    {"siteprefsofs", '9'}, {"lociconstraints", '5'}, {"contaminants", 'H'}, {"snpfile", 'S'}, {"snpcentroid", '7'},
    {"markerlen", 'K'}, {"markerpolythres", 'G'}, {"editdelta", 'e'}, {"substitutions", 's'}, {"minflankexacts", 'x'},
    {"pcrprimersubs", '6'}, {"maxns", 'n'}, {"title", 't'}, {"chromexclude", 'Z'}, {"chromeinclude", 'z'}, {"threads", 'T'},
    {"maxmulti", 'R'}, {"clampmaxmulti", 'X'}, {"bestmatches", 'N'}, {"mlmode", 'r'}, {"snpreadsmin", 'p'}, {"snpnonrefpcnt", '1'},
    {"pecircularised", '2'}, {"petranslendist", '3'}, {"rptsamseqsthres", '4'}, {"qvalue", 'P'}, {"trim5", 'y'}, {"trim3", 'Y'},
    {"minacceptreadlen", 'l'}, {"maxacceptreadlen", 'L'}, {"sumrslts", 'q'}, {"experimentname", 'w'}, {"experimentdescr", 'W'}};

// "@file" arguments replaced by the options the file holds
static bool expand_param_files(int argc, char** argv, std::vector<std::string>& out) {
  for (int i = 0; i < argc; ++i) {
    if (argv[i][0] != '@') { out.push_back(argv[i]); continue; }
    std::string path = argv[i] + 1;
    while (!path.empty() && isspace((unsigned char)path.front())) path.erase(path.begin());
    while (!path.empty() && isspace((unsigned char)path.back())) path.pop_back();
    if (path.empty()) { printf("No options file specified following '@' switch"); return false; }
    FILE* f = fopen(path.c_str(), "r");
    if (!f) { printf("Unable to open options file '%s'\nError: %s", path.c_str(), strerror(errno)); return false; }
    char line[8196];
    while (fgets(line, sizeof(line), f)) {
      char* b = line;
      while (*b && isspace((unsigned char)*b)) ++b;
      char* e = b + strlen(b);
      while (e > b && isspace((unsigned char)e[-1])) --e;
      *e = '\0';
      if (!*b || *b == '#' || *b == ';' || (b[0] == '/' && b[1] == '/')) continue;
      std::string opt;
      bool in_quotes = false, in_param = false;
      for (const char* c = b;; ++c) {
        char ch = *c == 0x16 ? '-' : *c;   // "on some systems '-' is represented as 0x16"
        if (ch == '"' || ch == '\'') { in_quotes = !in_quotes; in_param = true; opt += ch; continue; }
        if (ch && ((ch != ' ' && ch != '\t') || in_quotes)) { in_param = true; opt += ch; continue; }
        if (in_param) { out.push_back(opt); opt.clear(); in_param = false; in_quotes = false; }
        if (!ch) break;
      }
    }
    fclose(f);
  }
  return true;
}

static int parse(int argc0, char** argv0, Opts& o) {
  std::vector<std::string> args;
  if (!expand_param_files(argc0, argv0, args)) return -1;
  const int argc = (int)args.size();
  int i = 1;
  if (i < argc && (args[i] == "align" || args[i] == "kanga")) ++i;
  std::vector<std::string> unsupported;
  bool pe_insert_dist = false;   // -3 (experimental in the reference): per-transcript insert lengths beside the -O file of a paired-end run
  for (; i < argc; ++i) {
    std::string a = args[i];
    if (a == "--gpus" && i + 1 < argc) { o.gpus = atoi(args[++i].c_str()); continue; }
    if (a.rfind("--gpus=", 0) == 0) { o.gpus = atoi(a.c_str() + 7); continue; }
    // one argument may hold several short options (-EX, -Xs3); a long option is rewritten into its short form first
    std::string shorts;
    std::string attached;
    bool have_attached = false;
    if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
      const size_t eq = a.find('=');
      const std::string name = a.substr(2, eq == std::string::npos ? std::string::npos : eq - 2);
      char letter = 0;
      for (const auto& lo : kLongOpts) if (name == lo.name) letter = lo.letter;
      if (!letter) { fprintf(stderr, "unrecognised argument '%s'\n", a.c_str()); return -1; }
      shorts = std::string(1, letter);
      if (eq != std::string::npos) { attached = a.substr(eq + 1); have_attached = true; }
    } else if (a.size() >= 2 && a[0] == '-' && a[1] != '-') {
      shorts = a.substr(1);
    } else {
      fprintf(stderr, "unrecognised argument '%s'\n", a.c_str());
      return -1;
    }
   for (size_t ci = 0; ci < shorts.size(); ++ci) {
    char c = shorts[ci];
    std::string v;
    if (takes_value(c)) {
      if (have_attached) v = attached;
      else if (ci + 1 < shorts.size()) v = shorts.substr(ci + 1);
      else if (i + 1 < argc) v = args[++i];
      else { fprintf(stderr, "option -%c needs a value\n", c); return -1; }
      ci = shorts.size();
      if (v.size() >= 2 && (v.front() == '"' || v.front() == '\'') && v.back() == v.front()) v = v.substr(1, v.size() - 2);
    }
    int iv = atoi(v.c_str());
    switch (c) {
      case 'f': g_loglevel = iv; break;
      case 'F': o.logfile = v; break;
      case 'm': o.pmode = iv; break;
      case 'Q': o.strand = iv; break;
      case 's': o.max_subs = iv; break;
      case 'e': o.edit_delta = iv; break;
      case 'n': o.max_ns = iv; break;
      case 'M': o.fmt = iv; break;
      case 'U': o.pe_mode = iv; break;
      case 'd': o.pair_min = iv; break;
      case 'D': o.pair_max = iv; break;
      case 'E': o.pair_strand = true; break;
      case '2': o.pe_circ = true; break;
      case 'y': o.trim5 = iv; break;
      case 'Y': o.trim3 = iv; break;
      case 'l': o.min_len = iv; break;
      case 'L': o.max_len = iv; break;
      case 'T': o.threads = iv; break;
      case '4': o.sam_seq_thres = iv; break;
      case 'i': o.in.push_back(v); break;
      case 'u': o.pair.push_back(v); break;
      case 'I': o.sfx = v; break;
      case 'o': o.out = v; break;
      case 't': o.title = v; break;
      case 'g': o.qmode = iv; break;
      case '#': o.sample_nth = std::min(10000, std::max(1, iv)); break;   // clamped like kanga.cpp:459-464
      case 'r': o.ml_mode = iv; if (iv == 2) unsupported.push_back("-r2 (random locus: libc rand() in the reference, not reproducible)"); break;
      case 'R': o.max_ml = iv; break;  // only meaningful with -r
      case 'c': if (iv) unsupported.push_back("-c chimeric trimming"); break;
      case 'a': if (iv) unsupported.push_back("-a microInDels"); break;
      case 'A': if (iv) unsupported.push_back("-A splice junctions"); break;
      case 'k': o.pcr_win = iv; if (iv < 0 || iv > 250) { fprintf(stderr, "Error: PCR differential amplification artefacts window length '-k%d' specified outside of range 0..250\n", iv); return -1; } break;
      case 'x': o.min_flank = iv; break;
      case 'Z': o.excl.push_back(v); break;
      case 'z': o.incl.push_back(v); break;
      case 'j': o.none_file = v; break;
      case 'J': o.multi_file = v; break;
      case 'p': if (iv) unsupported.push_back("-p SNP calling"); break;
      case '6': o.pcr_primer = iv; break;
      case 'b': unsupported.push_back("-b bisulfite"); break;
      case 'C': unsupported.push_back("-C colorspace"); break;
      case 'N': o.best_matches = true; break;
      case 'X': o.clamp_ml = true; break;
      case '5': o.constraints_file = v; break;
      case 'O': o.stats_file = v; break;
      case 'q': o.sqlite_file = v; break;   // results-summary database: accepted, see below
      case 'w': o.exp_name = v; break;
      case 'W': o.exp_descr = v; break;
      case 'B': o.priority_file = v; break;
      case 'V': o.priority_nofilt = true; break;
      case 'H': o.contam_file = v; break;
      case '3': pe_insert_dist = true; break;
      case 'S': case '7': case '8':
        unsupported.push_back(std::string("-") + c + " (outside the accelerated path)"); break;
      case 'v':
        printf("\nbiokanga align Version 4.4.2 (bkx B200 path)\n");
        return 1;
      case 'h':
        printf("bkx-align: B200 drop-in for `biokanga align` -- options -m -Q -s -e -n -r{1,3,4,5} -R -X -N -M0..6 -g -t -U -d -D -E "
               "-y -Y -l -L -# -H -5 -k -6 -x -Z -z -B -V -j -J -O -4 -T -i -u -I -o -F -q -w -W [--gpus N], the reference's long option names, "
               "and @file parameter files\n");
        return 1;
      default: break;  // remaining reference options have no effect on this path (-K -G -P -1 -9 -0: parameters of -p, -8 and -c, which are refused)
    }
   }
  }
  if (!unsupported.empty()) {
    for (auto& u : unsupported) fprintf(stderr, "bkx-align: option %s is not supported by the accelerated path\n", u.c_str());
    return -1;
  }
  // validation mirrors kanga.cpp:452-862
  if (o.sfx.empty() || o.in.empty() || o.out.empty()) { fprintf(stderr, "bkx-align: -I, -i and -o are required\n"); return -1; }
  if (!o.sqlite_file.empty()) {   // kanga.cpp:381-404: the summary database needs an experiment name and description
    if (o.exp_name.empty()) { fprintf(stderr, "Error: After removal of whitespace, no SQLite experiment name specified with '-w<str>' option\n"); return -1; }
    if (o.exp_descr.empty()) { fprintf(stderr, "Error: After removal of whitespace, no SQLite experiment description specified with '-W<str>' option\n"); return -1; }
  }
  if (o.pmode < 0 || o.pmode > 3) { fprintf(stderr, "Error: Processing mode '-m%d' must be in range 0..3\n", o.pmode); return -1; }
  if (o.max_subs < 0 || o.max_subs > 15) { fprintf(stderr, "Error: max substitutions '-s%d' must be in range 0..15\n", o.max_subs); return -1; }
  if (o.edit_delta < 1 || o.edit_delta > 2) { fprintf(stderr, "Error: Min Hamming edit distance '-e%d' must be in range 1..2\n", o.edit_delta); return -1; }
  if (o.max_ns < 0 || o.max_ns > 5) { fprintf(stderr, "Error: Allowed number of indeterminate 'N's '-n%d' must be in range 0..5\n", o.max_ns); return -1; }
  if (o.fmt < 0 || o.fmt > 6) { fprintf(stderr, "Error: Output format mode '-M%d' specified outside of range 0..6\n", o.fmt); return -1; }
  if (o.qmode < 0 || o.qmode > 3) { fprintf(stderr, "Error: fastq quality scoring method '-g%d' specified outside of range 0..3\n", o.qmode); return -1; }
  if (o.pe_mode && !(o.fmt == 0 || o.fmt >= 4)) { fprintf(stderr, "Error: paired end processing supports output formats -M0, -M4, -M5 and -M6 only\n"); return -1; }
  if (o.pe_mode < 0 || o.pe_mode > 4) { fprintf(stderr, "Error: paired end mode '-U%d' must be in range 0..4\n", o.pe_mode); return -1; }
  if (o.pe_mode && o.pair.size() != o.in.size()) { fprintf(stderr, "Error: Paired end processing '-U%d' requested but number of PE1 files not same as PE2 files\n", o.pe_mode); return -1; }
  if (o.min_len < 15 || o.min_len > 2000 || o.max_len < o.min_len || o.max_len > 2000) { fprintf(stderr, "Error: read length limits out of range\n"); return -1; }
  if (o.pcr_primer < 0 || o.pcr_primer > 5) { fprintf(stderr, "Error: PCR primer correction subs '-6%d' specified outside of range 0..5\n", o.pcr_primer); return -1; }  // kanga.cpp:784-789
  if (o.min_flank < 0 || o.min_flank > 7) { fprintf(stderr, "Error: Max flank trimming '-x%d' specified outside of range 0..7\n", o.min_flank); return -1; }  // kanga.cpp:804-808
  if (!o.stats_file.empty() && o.fmt == 6) { fprintf(stderr, "Error: Output induced substitution mode '-O<file>' not available in '-M6' output mode\n"); return -1; }  // kanga.cpp:1015-1021
  if (o.excl.size() > 20 || o.incl.size() > 20) { fprintf(stderr, "Error: at most 20 '-Z' and 20 '-z' chromosome expressions\n"); return -1; }
  if (o.gpus < 1) o.gpus = 1;
  if (pe_insert_dist && o.pe_mode && !o.stats_file.empty()) {   // Aligner.cpp:175-185, 5340-5470: <stats file>.peinserts.csv
    fprintf(stderr, "bkx-align: option -3 (paired-end insert lengths per target sequence, experimental in the reference) is not supported by the accelerated path\n");
    return -1;
  }
  if (o.priority_file.empty()) o.priority_nofilt = false;   // kanga.cpp:1071-1082: -V is only read together with -B
  if (o.ml_mode != 0) {  // kanga.cpp:535-537, 667-696
    if (o.pe_mode) { fprintf(stderr, "Error: Sorry, currently multiloci processing '-r%d' not supported in paired end '-U%d' processing\n", o.ml_mode, o.pe_mode); return -1; }
    if (o.max_ml == 0) o.max_ml = 5;  // cDfltMaxMultiHits
    const int lim = 500;  // cMaxMultiHits; -r5 alone goes on to 100000 in the reference (it writes loci out as it finds them),
                          // here every read owns -R slots
    if (o.max_ml < 2 || o.max_ml > lim) { fprintf(stderr, "Error: multiple aligned reads '-R%d' specified outside of range 2..%d\n", o.max_ml, lim); return -1; }
    if (o.ml_mode == 5) { o.none_file.clear(); o.multi_file.clear(); }   // kanga.cpp:1045-1069: -j / -J are dropped under -r5
    if (o.ml_mode == 5 && o.fmt == 6 && (!o.excl.empty() || !o.incl.empty())) {
      // the reference then turns reads whose loci were all filtered into accepted records without a locus (WriteHitLoci, Aligner.cpp:6755-6771)
      fprintf(stderr, "bkx-align: options -Z / -z are not supported together with '-r5' in output format '-M6'\n");
      return -1;
    }
    if (o.ml_mode == 5 && o.pcr_primer) {   // there every record owns a copy of its read; here the loci of a read share one
      fprintf(stderr, "bkx-align: option -6 is not supported together with '-r5'\n");
      return -1;
    }
    if (o.ml_mode == 5 && !(o.fmt == 0 || o.fmt == 4 || o.fmt == 5 || o.fmt == 6)) {  // kanga.cpp:830
      fprintf(stderr, "Error: '-r5' is only reported as -M0, -M4, -M5 or -M6\n");
      return -1;
    }
    if (o.best_matches) o.clamp_ml = true;   // kanga.cpp:695-696
  } else {
    o.max_ml = 1;
    o.clamp_ml = false;  // kanga.cpp:691-694: -X only counts together with -R
    o.best_matches = false;   // kanga.cpp:666, 686: -N is only read together with a multi-loci mode
  }
  return 0;
}

// ---- ordering: bkx_sort_hits (SortHitMatch, Aligner.cpp:10067-10114; ties broken by read id for determinism)

static const char* kNarCode[] = {"NA", "AA", "EN", "NL", "MH", "ML", "ET", "OJ", "OM", "DP", "DS", "FC", "PR", "UI", "OI", "UP", "IS", "IT", "NP", "LC"};
static const char* kNarText[] = {
    "Not processed for alignment", "Alignment accepted", "Excessive indeterminate (Ns) bases", "No potential alignment loci",
    "Mismatch delta (minimum Hamming) criteria not met", "Aligned to multiloci", "Excessively end trimmed",
    "Aligned as orphaned splice junction", "Aligned as orphaned microInDel", "Duplicate PCR", "Duplicate read sequence",
    "Aligned to filtered target sequence", "Aligned to a priority region", "PE under minimum insert size",
    "PE over maximum insert size", "PE partner not aligned", "PE partner aligned to inconsistent strand",
    "PE partner aligned to different target sequence", "PE alignment not accepted", "Alignment violated loci base constraints"};

struct OutBuf {  // plain or gzip (when the output name ends in .gz, as the reference does)
  FILE* f = nullptr;
  gzFile gz = nullptr;
  std::string s;
  bool open(const std::string& path) {
    s.reserve(1 << 22);
    if (path.size() > 3 && path.compare(path.size() - 3, 3, ".gz") == 0) { gz = gzopen(path.c_str(), "wb"); return gz != nullptr; }
    f = fopen(path.c_str(), "w+b");   // readable too: emit_rows maps the file to let all threads fill it
    if (!f) f = fopen(path.c_str(), "wb");
    return f != nullptr;
  }
  void flush() {
    if (s.empty()) return;
    if (gz) gzwrite(gz, s.data(), (unsigned)s.size()); else fwrite(s.data(), 1, s.size(), f);
    s.clear();
  }
  void maybe() { if (s.size() > (1 << 22) - 8192) flush(); }
  void write(const std::string& t) {
    flush();
    if (t.empty()) return;
    if (gz) gzwrite(gz, t.data(), (unsigned)t.size()); else fwrite(t.data(), 1, t.size(), f);
  }
  void close() { flush(); if (gz) gzclose(gz); if (f) fclose(f); gz = nullptr; f = nullptr; }
};

// Format rows [0, n) with `threads` workers, each filling its own buffer for a contiguous run of rows; the buffers
// are written in row order.  row(k, out) appends the text of row k (possibly nothing).
template <class F>
static void emit_rows(OutBuf& ob, uint32_t n, unsigned threads, F&& row) {
  const uint32_t chunk = 1u << 16;
  if (threads < 1) threads = 1;
  std::vector<std::string> bufs(threads);
  // plain files: every worker writes its own rows at their final offset (pwrite), so that formatting AND writing run on
  // all threads; gzip streams are sequential by nature
  int fd = -1;
  off_t file_ofs = 0;
  if (ob.f && !ob.gz) {
    ob.flush();
    fflush(ob.f);
    fd = fileno(ob.f);
    file_ofs = ftello(ob.f);
    if (file_ofs < 0) fd = -1;
  }
  for (uint64_t base = 0; base < n; base += (uint64_t)chunk * threads) {
    std::vector<std::thread> th;
    unsigned used = 0;
    for (unsigned t = 0; t < threads; ++t) {
      uint64_t b = base + (uint64_t)t * chunk;
      if (b >= n) break;
      uint64_t e = std::min<uint64_t>(n, b + chunk);
      ++used;
      th.emplace_back([&row, &bufs, t, b, e]() {
        std::string& s = bufs[t];
        s.clear();
        for (uint64_t k = b; k < e; ++k) row((uint32_t)k, s);
      });
    }
    for (auto& x : th) x.join();
    if (fd < 0) {
      for (unsigned t = 0; t < used; ++t) ob.write(bufs[t]);
      continue;
    }
    std::vector<off_t> at(used);
    const off_t batch_ofs = file_ofs;
    for (unsigned t = 0; t < used; ++t) { at[t] = file_ofs; file_ofs += (off_t)bufs[t].size(); }
    if (file_ofs == batch_ofs) continue;
    // the file grows by the batch and the workers copy their rows into a shared mapping of that stretch: page faults and
    // copies run on all threads (concurrent write() calls on one file serialise on its inode lock); pwrite if mmap fails
    const off_t page = (off_t)sysconf(_SC_PAGESIZE), map_ofs = batch_ofs & ~(page - 1);
    char* map = nullptr;
    if (ftruncate(fd, file_ofs) == 0) {
      void* m = mmap(nullptr, (size_t)(file_ofs - map_ofs), PROT_READ | PROT_WRITE, MAP_SHARED, fd, map_ofs);
      if (m != MAP_FAILED) map = (char*)m;
    }
    std::vector<char> ok(used, 1);
    th.clear();
    for (unsigned t = 0; t < used; ++t)
      th.emplace_back([&bufs, &at, &ok, fd, t, map, map_ofs]() {
        const char* p = bufs[t].data();
        size_t left = bufs[t].size();
        off_t o = at[t];
        if (map) { memcpy(map + (o - map_ofs), p, left); return; }
        while (left) {
          ssize_t w = pwrite(fd, p, left, o);
          if (w <= 0) { ok[t] = 0; return; }
          p += w; left -= (size_t)w; o += w;
        }
      });
    for (auto& x : th) x.join();
    if (map) munmap(map, (size_t)(file_ofs - map_ofs));
    for (unsigned t = 0; t < used; ++t) if (!ok[t]) { fd = -2; break; }
    if (fd == -2) { diag("Fatal: write to the result file failed"); exit(1); }
  }
  if (fd >= 0) fseeko(ob.f, file_ofs, SEEK_SET);
}

// ---- BAM (BGZF) + BAI, as CSAMfile writes them (libbiokanga/SAMfile.cpp:1383-1660, 1839-2030, 2286-2556; bgzf.cpp) ----
// BGZF: 0xff00-byte blocks, raw deflate level 6 (Aligner.cpp:722), standard 18-byte header / 8-byte footer, an empty
// block at close.  Virtual address = (file offset of the block << 16) | offset inside the block.
struct Bgzf {
  FILE* f = nullptr;
  std::vector<uint8_t> ubuf, cbuf;
  size_t uofs = 0;
  uint64_t block_addr = 0;
  int level = 6;
  bool open(const std::string& path) {
    f = fopen(path.c_str(), "wb");
    ubuf.resize(0x10000);
    cbuf.resize(0x10000 + 1024);
    return f != nullptr;
  }
  bool deflate_block(size_t len) {
    z_stream zs;
    memset(&zs, 0, sizeof(zs));
    zs.next_in = ubuf.data();
    zs.avail_in = (uInt)len;
    zs.next_out = cbuf.data() + 18;
    zs.avail_out = (uInt)(cbuf.size() - 18 - 8);
    if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) return false;
    if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { deflateEnd(&zs); return false; }
    deflateEnd(&zs);
    size_t dlen = zs.total_out + 18 + 8;
    static const uint8_t magic[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0, 0};
    memcpy(cbuf.data(), magic, 18);
    uint16_t bs = (uint16_t)(dlen - 1);
    memcpy(cbuf.data() + 16, &bs, 2);
    uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), ubuf.data(), (uInt)len), isz = (uint32_t)len;
    memcpy(cbuf.data() + dlen - 8, &crc, 4);
    memcpy(cbuf.data() + dlen - 4, &isz, 4);
    if (fwrite(cbuf.data(), 1, dlen, f) != dlen) return false;
    block_addr += dlen;
    uofs = 0;
    return true;
  }
  bool flush() { return uofs == 0 || deflate_block(uofs); }
  bool write(const void* p, size_t len) {
    const uint8_t* b = (const uint8_t*)p;
    while (len) {
      size_t c = std::min(len, (size_t)0xff00 - uofs);
      memcpy(ubuf.data() + uofs, b, c);
      uofs += c; b += c; len -= c;
      if (uofs == 0xff00 && !deflate_block(uofs)) return false;
    }
    return true;
  }
  uint64_t tell() const { return (block_addr << 16) | (uint64_t)(uofs & 0xffff); }
  // Whole-stream form of write(): the blocks a sequence of write() calls would have produced -- a cut every 0xff00
  // bytes and one more at *flush_at (an explicit flush()) -- deflated by `threads` workers and written in order.  A
  // trailing partial block stays pending for close(), as it would after write().  vaddr() then maps stream offsets to
  // the virtual addresses tell() would have returned at those points.
  std::vector<uint64_t> cut_u, cut_c;   // block starts in the stream / in the file, plus the end sentinel
  bool write_stream(const uint8_t* data, uint64_t len, const uint64_t* flush_at, unsigned threads) {
    if (uofs) return false;  // only from a block boundary
    cut_u.clear(); cut_c.clear();
    uint64_t pos = 0;
    while (pos < len) {
      cut_u.push_back(pos);
      uint64_t nxt = std::min<uint64_t>(pos + 0xff00, len);
      if (flush_at && *flush_at > pos && *flush_at < nxt) nxt = *flush_at;
      pos = nxt;
    }
    const size_t nblk = cut_u.size();
    cut_u.push_back(len);
    // the last block stays in ubuf unless it is full or ends at the explicit flush
    size_t nfull = nblk;
    if (nblk) {
      uint64_t lb = cut_u[nblk - 1], le = len;
      bool closed = (le - lb == 0xff00) || (flush_at && *flush_at == le);
      if (!closed) nfull = nblk - 1;
    }
    std::vector<std::vector<uint8_t>> comp(nfull);
    std::vector<char> good(std::max<unsigned>(threads, 1), 1);
    {
      std::vector<std::thread> th;
      unsigned T = std::max<unsigned>(threads, 1);
      for (unsigned t = 0; t < T; ++t)
        th.emplace_back([&, t]() {
          for (size_t j = t; j < nfull; j += T) {
            size_t ulen = (size_t)(cut_u[j + 1] - cut_u[j]);
            std::vector<uint8_t>& out = comp[j];
            out.resize(ulen + 1024 + 26);
            z_stream zs;
            memset(&zs, 0, sizeof(zs));
            zs.next_in = const_cast<uint8_t*>(data + cut_u[j]);
            zs.avail_in = (uInt)ulen;
            zs.next_out = out.data() + 18;
            zs.avail_out = (uInt)(out.size() - 18 - 8);
            if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { good[t] = 0; return; }
            if (deflate(&zs, Z_FINISH) != Z_STREAM_END) { deflateEnd(&zs); good[t] = 0; return; }
            deflateEnd(&zs);
            size_t dlen = zs.total_out + 18 + 8;
            static const uint8_t magic[18] = {0x1f, 0x8b, 8, 4, 0, 0, 0, 0, 0, 0xff, 6, 0, 'B', 'C', 2, 0, 0, 0};
            memcpy(out.data(), magic, 18);
            uint16_t bs = (uint16_t)(dlen - 1);
            memcpy(out.data() + 16, &bs, 2);
            uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), data + cut_u[j], (uInt)ulen), isz = (uint32_t)ulen;
            memcpy(out.data() + dlen - 8, &crc, 4);
            memcpy(out.data() + dlen - 4, &isz, 4);
            out.resize(dlen);
          }
        });
      for (auto& x : th) x.join();
    }
    for (char g : good) if (!g) return false;
    for (size_t j = 0; j < nfull; ++j) {
      cut_c.push_back(block_addr);
      if (fwrite(comp[j].data(), 1, comp[j].size(), f) != comp[j].size()) return false;
      block_addr += comp[j].size();
    }
    cut_c.push_back(block_addr);  // address of the block that follows (pending or next)
    if (nfull < nblk) {           // pending partial block
      uofs = (size_t)(len - cut_u[nblk - 1]);
      memcpy(ubuf.data(), data + cut_u[nblk - 1], uofs);
    }
    return true;
  }
  uint64_t vaddr(uint64_t u) const {  // what tell() returned when the stream stood at offset u (after any flush due there)
    if (u >= cut_u.back()) return (cut_c.back() << 16) | (uint64_t)(uofs & 0xffff);  // end of stream: pending bytes, if any
    size_t j = (size_t)(std::upper_bound(cut_u.begin(), cut_u.end(), u) - cut_u.begin()) - 1;  // block with start <= u
    return (cut_c[j] << 16) | (uint64_t)((u - cut_u[j]) & 0xffff);
  }
  bool close() {
    bool ok = flush() && deflate_block(0);  // trailing empty block = BGZF EOF marker
    if (f) fclose(f);
    f = nullptr;
    return ok;
  }
};

static int bai_reg2bin(int beg, int end) {  // SAM spec section 5.3 (half-open interval)
  --end;
  if (beg >> 14 == end >> 14) return ((1 << 15) - 1) / 7 + (beg >> 14);
  if (beg >> 17 == end >> 17) return ((1 << 12) - 1) / 7 + (beg >> 17);
  if (beg >> 20 == end >> 20) return ((1 << 9) - 1) / 7 + (beg >> 20);
  if (beg >> 23 == end >> 23) return ((1 << 6) - 1) / 7 + (beg >> 23);
  if (beg >> 26 == end >> 26) return ((1 << 3) - 1) / 7 + (beg >> 26);
  return 0;
}

// BAI built the way CSAMfile::AddChunk / UpdateSAIIndex do: per reference, bins in ascending number, chunks of a bin
// merged while the next alignment starts no later than the chunk's last end + 1, a sparse 16 kb linear index
// (windows in which no alignment starts stay 0), references after the last aligned one get no block.
struct BaiBuilder {
  struct Chunk { uint64_t sva, eva; uint32_t start, end; };
  std::vector<std::pair<int, std::vector<Chunk>>> bins;  // filled through `slot`
  std::vector<int> slot;                                   // bin number -> index into bins, -1 if empty
  std::vector<uint64_t> lin;
  uint32_t n_lin = 0;
  std::string out;
  BaiBuilder() : slot(37450, -1) {}
  void add(uint64_t sva, uint32_t start, uint64_t eva, uint32_t end) {
    uint32_t k = start / 0x4000;
    if (lin.size() <= k) lin.resize(k + 1, 0);
    if (lin[k] == 0) { n_lin = k + 1; lin[k] = sva; }
    int bin = bai_reg2bin((int)start, (int)end);
    if (slot[bin] < 0) { slot[bin] = (int)bins.size(); bins.push_back({bin, {}}); bins.back().second.push_back({sva, eva, start, end}); return; }
    std::vector<Chunk>& cs = bins[slot[bin]].second;
    Chunk& c = cs.back();
    if (start > c.end + 1) cs.push_back({sva, eva, start, end});
    else {
      if (c.start > start) { c.start = start; c.sva = sva; }
      if (c.end < end) c.end = end;
      c.eva = eva;
    }
  }
  void put32(uint32_t v) { out.append((const char*)&v, 4); }
  void put64(uint64_t v) { out.append((const char*)&v, 8); }
  void end_ref() {  // UpdateSAIIndex: emit this reference's block and reset
    put32((uint32_t)bins.size());
    if (!bins.empty()) {
      std::sort(bins.begin(), bins.end(), [](const auto& a, const auto& b) { return a.first < b.first; });
      for (auto& b : bins) {
        put32((uint32_t)b.first);
        put32((uint32_t)b.second.size());
        for (auto& c : b.second) { put64(c.sva); put64(c.eva); }
      }
      put32(n_lin);
      for (uint32_t i = 0; i < n_lin; ++i) put64(lin[i]);
    } else {
      put32(0);
    }
    bins.clear();
    std::fill(slot.begin(), slot.end(), -1);
    lin.clear();
    n_lin = 0;
  }
};

static bool is_bam_name(const std::string& p) {  // kanga.cpp:849-857: longer than 5 chars and ending in ".bam"
  if (p.size() <= 5) return false;
  std::string e = p.substr(p.size() - 4);
  for (auto& c : e) c = (char)tolower((unsigned char)c);
  return e == ".bam";
}

static void append_uint(std::string& s, uint64_t v) {   // decimal digits without printf (hundreds of millions of calls per run)
  char b[24];
  int n = 24;
  do { b[--n] = (char)('0' + v % 10); v /= 10; } while (v);
  s.append(b + n, (size_t)(24 - n));
}

// ---- -5: loci base constraints, CAligner::LoadLociConstraints, Aligner.cpp:1245-1441.  CSV rows chrom,start,end,bases:
//      reads aligned over [start, end] of chrom are only kept if their base there is one of `bases` (A C G T; R = the
//      target's own base).  An all-text first row is taken for a title row (CCSVFile::IsLikelyHeaderLine).
struct LociConstraint { uint32_t chrom, start, end; uint8_t mask; };

static int load_constraints(const std::string& path, const std::vector<bkx_entry>& ents, std::vector<LociConstraint>& out) {
  diag("Loading loci base constraints from CSV file '%s' ...", path.c_str());
  std::vector<char> text;
  if (!slurp(path, text, 1)) { diag("Unable to open '%s' for processing", path.c_str()); return -1; }
  const char *p = text.data(), *end = p + text.size();
  int line_no = 0, rows = 0;
  std::vector<uint32_t> chroms;
  while (p < end) {
    const char* eol = line_end(p, end);
    std::string line(p, eol);
    p = next_line(eol, end);
    ++line_no;
    while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
    if (line.empty()) continue;
    std::vector<std::string> f;
    std::vector<char> quoted;
    for (size_t b = 0; b <= line.size();) {
      size_t c = line.find(',', b);
      if (c == std::string::npos) c = line.size();
      std::string v = line.substr(b, c - b);
      size_t l = v.find_first_not_of(" \t"), r = v.find_last_not_of(" \t");
      v = l == std::string::npos ? std::string() : v.substr(l, r - l + 1);
      bool q = v.size() >= 2 && (v.front() == '"' || v.front() == '\'') && v.back() == v.front();
      if (q) v = v.substr(1, v.size() - 2);
      f.push_back(v);
      quoted.push_back(q);
      b = c + 1;
    }
    if (f.size() < 4) { diag("Expected at least 4 fields at line %d in '%s', GetCurFields() returned '%d'", line_no, path.c_str(), (int)f.size()); return -1; }
    if (++rows == 1) {  // title row: no unquoted field that parses as a number, at most two empty ones
      bool header = true;
      int empties = 0;
      for (size_t k = 0; k < f.size() && header; ++k) {
        if (quoted[k]) continue;
        if (f[k].empty()) { if (++empties > 2) header = false; continue; }
        char* term = nullptr;
        strtod(f[k].c_str(), &term);
        if (term && *term == 0) header = false;
      }
      if (header) continue;
    }
    uint32_t chrom = 0;
    for (uint32_t e = 1; e < ents.size(); ++e) if (!strcasecmp(f[0].c_str(), ents[e].name)) { chrom = e; break; }   // GetIdent: case insensitive
    if (!chrom) { diag("Unable to find matching indexed identifier for '%s' at line %d in '%s'", f[0].c_str(), line_no, path.c_str()); return -1; }
    const int start = atoi(f[1].c_str()), stop = atoi(f[2].c_str());
    if (start < 0 || start > stop) { diag("Start loci must be >= 0 and <= end loci for '%s' at line %d in '%s'", f[0].c_str(), line_no, path.c_str()); return -1; }
    if ((uint32_t)stop >= ents[chrom].seq_len) { diag("End loci must be > targeted sequence length for '%s' at line %d in '%s'", f[0].c_str(), line_no, path.c_str()); return -1; }
    uint8_t mask = 0;
    bool bad = false;
    for (char c : f[3]) switch (c) {
      case 'a': case 'A': mask |= 0x01; break;
      case 'c': case 'C': mask |= 0x02; break;
      case 'g': case 'G': mask |= 0x04; break;
      case 't': case 'T': mask |= 0x08; break;
      case 'r': case 'R': mask |= 0x10; break;
      case ' ': case '\t': break;
      default: bad = true;
    }
    if (bad || !mask) { diag("Illegal base specifiers for '%s' at line %d in '%s'", f[0].c_str(), line_no, path.c_str()); return -1; }
    if (std::find(chroms.begin(), chroms.end(), chrom) == chroms.end()) {
      if (chroms.size() == 64) { diag("Number of constrained chroms would be more than max (64) allowed for '%s' at line %d in '%s'", f[0].c_str(), line_no, path.c_str()); return -1; }
      chroms.push_back(chrom);
    }
    out.push_back({chrom, (uint32_t)start, (uint32_t)stop, mask});
  }
  std::sort(out.begin(), out.end(), [](const LociConstraint& a, const LociConstraint& b) {
    return a.chrom != b.chrom ? a.chrom < b.chrom : a.start != b.start ? a.start < b.start : a.end < b.end;
  });
  diag("Completed loading %d loci base constraints for %d target sequences from CSV file '%s'", (int)out.size(), (int)chroms.size(), path.c_str());
  return (int)out.size();
}

// ---- what the post-alignment passes and the writers share about a run ----------------------------------
struct Records {
  const Opts& o;
  Reads& R;                                          // -6 rewrites read bases
  ResVec& res;                 // one per record: per read, or per reported locus under -r5
  const std::vector<uint32_t>& src;                  // record -> read (empty: identity)
  const std::vector<bkx_entry>& ents;                // [1 .. num_entries]
  const std::vector<std::vector<uint8_t>>& genome;   // host copy of the chromosomes (1 byte/base), filled when a pass or writer needs it
  unsigned threads;
  // -x: Seg[0].TrimLeft / TrimRight (in READ orientation) and TrimMismatches per record; empty unless the trimming ran
  std::vector<uint16_t> trim_l, trim_r;
  std::vector<uint8_t> trim_mm;
  uint32_t elim_plus = 0, elim_minus = 0;            // alignments sloughed by the trimming, per strand
  // HitLoci.FlagIA: accepted alignment of a `simreads` read that is not where its descriptor says (sim_truth_check); empty
  // unless the read set carries such descriptors
  std::vector<uint8_t> flag_ia;
  const char* align_type(uint32_t i) const { return !flag_ia.empty() && flag_ia[i] ? "iar" : "ar"; }   // Aligner.cpp:6391-6430

  uint32_t n() const { return (uint32_t)res.size(); }
  uint32_t rix(uint32_t i) const { return src.empty() ? i : src[i]; }
  // the alignment as reported: AdjStartLoci / AdjHitLen / TrimMismatches (Aligner.cpp:1528-1552)
  uint32_t tleft(uint32_t i) const { return trim_l.empty() ? 0u : trim_l[i]; }
  uint32_t tright(uint32_t i) const { return trim_r.empty() ? 0u : trim_r[i]; }
  uint32_t adj_start(uint32_t i) const { return res[i].match_loci + (res[i].strand == '+' ? tleft(i) : tright(i)); }
  uint32_t adj_len(uint32_t i) const { return (uint32_t)res[i].match_len - tleft(i) - tright(i); }
  uint32_t adj_mm(uint32_t i) const { return trim_mm.empty() ? res[i].mismatches : trim_mm[i]; }
};

// ---- -5: IdentifyConstraintViolations / AcceptLociConstraints / AcceptBaseConstraint, Aligner.cpp:2480-2647.  An accepted
//      alignment becomes eNARLociConstrained when, at some constrained locus it covers, the read's base (as aligned:
//      complemented on '-') is none of the bases that constraint allows; in paired-end runs the mate goes with it.
static void identify_constraint_violations(Records& rc, const std::vector<LociConstraint>& constraints) {
  const Opts& o = rc.o;
  const Reads& R = rc.R;
  auto& res = rc.res;
  const uint32_t nrec = rc.n();
  const auto& genome = rc.genome;
  diag("Identifying %s loci base constraint violations ...", o.pe_mode ? "PE" : "SE");
  std::vector<std::pair<size_t, size_t>> span(rc.ents.size() + 1, {0, 0});   // constraints of a chromosome: [first, last)
  for (size_t k = 0; k < constraints.size(); ++k) {
    auto& sp = span[constraints[k].chrom];
    if (sp.second == 0) sp.first = k;
    sp.second = k + 1;
  }
  auto violates = [&](const bkx_read_result& r, uint32_t ri) -> bool {
    const auto sp = span[r.chrom_id];
    if (sp.second == 0) return false;
    const uint8_t* b = R.bases.data() + R.offs[ri];
    const uint32_t L = r.match_len, lo = r.match_loci, hi = r.match_loci + L - 1;
    for (size_t k = sp.first; k < sp.second; ++k) {
      const LociConstraint& c = constraints[k];
      if (c.end < lo || c.start > hi) continue;
      for (uint32_t l = std::max(lo, c.start); l <= std::min(hi, c.end); ++l) {
        uint8_t base = r.strand == '-' ? b[L - 1 - (l - lo)] & 7 : b[l - lo] & 7;
        if (r.strand == '-' && base < 4) base = 3 - base;
        if ((c.mask & 0x10) && (genome[r.chrom_id][l] & 0x0f) == base) continue;
        if (base < 4 && (c.mask & (1u << base))) continue;
        return true;
      }
    }
    return false;
  };
  std::vector<uint8_t> bad(nrec, 0);
  {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < rc.threads; ++t)
      th.emplace_back([&, t]() {
        const uint32_t b0 = (uint32_t)((uint64_t)nrec * t / rc.threads), e0 = (uint32_t)((uint64_t)nrec * (t + 1) / rc.threads);
        for (uint32_t i = b0; i < e0; ++i) bad[i] = res[i].nar == BKX_NAR_ACCEPTED && violates(res[i], rc.rix(i));
      });
    for (auto& x : th) x.join();
  }
  int identified = 0;
  auto mark = [&](uint32_t i) {
    if (res[i].nar == BKX_NAR_LOCICONSTRAINED) return;
    res[i].nar = BKX_NAR_LOCICONSTRAINED; res[i].num_hits = 0; res[i].low_hit_instances = 0;
    ++identified;
  };
  for (uint32_t i = 0; i < nrec; ++i)
    if (bad[i]) { mark(i); if (o.pe_mode) mark(i ^ 1u); }
  diag("Identified %d %s loci base constraint violations", identified, o.pe_mode ? "PE" : "SE");
}

// ---- -k: ReducePCRduplicates, Aligner.cpp:2184-2282 (single-end runs only, :599).  In hit order, alignments that
//      share chromosome, start, strand and length with an earlier one are dropped as eNARPCRdup once the allowance
//      is used up; with a window the allowance grows with the number of distinct start loci within WinLen either
//      side (NumUpUniques / NumDnUniques, :9817-9914).  Which of several reads with identical sort keys survives is
//      left open by the reference's unstable sort; here it is the one loaded first.
static int reduce_pcr_duplicates(Records& rc) {
  const Opts& o = rc.o;
  auto& res = rc.res;
  const uint32_t nrec = rc.n();
  diag("Processing to reduce PCR differential amplification artefacts processing started..");
  std::vector<uint32_t> ord(nrec);
  if (nrec && bkx_sort_hits(res.data(), nrec, ord.data(), 0) < 0) { diag("Fatal: %s", bkx_last_error()); return -1; }
  const int W = o.pcr_win;
  auto acc = [&](uint32_t k) { return res[ord[k]].nar == BKX_NAR_ACCEPTED; };
  auto dn_uniques = [&](uint32_t k) -> int {
    const bkx_read_result& c = res[ord[k]];
    int n_u = 0;
    uint32_t prv = c.match_loci;
    for (uint32_t j = k + 1; j < nrec; ++j) {
      const bkx_read_result& x = res[ord[j]];
      if (x.chrom_id != c.chrom_id) break;
      if (!acc(j)) continue;
      if ((int64_t)c.match_loci + W < (int64_t)x.match_loci) break;
      if (x.strand != c.strand) continue;
      if (x.match_loci != prv) { ++n_u; prv = x.match_loci; }
    }
    return n_u;
  };
  auto up_uniques = [&](uint32_t k) -> int {
    if (k == 0) return 0;
    const bkx_read_result& c = res[ord[k]];
    int n_u = 0;
    uint32_t prv = c.match_loci;
    for (uint32_t j = k + 1; j-- > 0;) {   // starts at the current record itself, as the reference does
      const bkx_read_result& x = res[ord[j]];
      if (x.chrom_id != c.chrom_id) break;
      if (!acc(j)) continue;
      if ((int64_t)c.match_loci > W && (int64_t)c.match_loci - W > (int64_t)x.match_loci) break;
      if (x.strand != c.strand) continue;
      if (x.match_loci != prv) { ++n_u; prv = x.match_loci; }
    }
    return n_u;
  };
  int removed = 0;
  for (uint32_t k = 0; k < nrec; ++k) {
    if (!acc(k)) continue;
    int limit = 0;
    if (W > 0) {
      limit = std::max(up_uniques(k), dn_uniques(k));
      const int prop = (int)(((double)limit / W) * 100.0);
      limit = prop < 5 ? 1 : prop <= 10 ? 2 : prop <= 20 ? 3 : prop <= 40 ? 4 : prop <= 60 ? 5 : prop <= 80 ? 10 : 50;
    }
    const bkx_read_result c = res[ord[k]];
    uint32_t mark = k;
    for (uint32_t j = k + 1; j < nrec; ++j) {
      bkx_read_result& x = res[ord[j]];
      if (x.nar != BKX_NAR_ACCEPTED) continue;
      if (x.chrom_id != c.chrom_id || x.match_loci != c.match_loci || x.strand != c.strand) break;
      if (x.match_len != c.match_len) continue;
      if (limit > 0) { --limit; continue; }
      x.num_hits = 0; x.low_hit_instances = 0; x.nar = BKX_NAR_PCRDUP;
      mark = j;
      ++removed;
    }
    k = mark;
  }
  diag("Removed %d potential PCR artefact reads", removed);
  diag("PCR differential amplification artefacts processing completed");
  return 0;
}

// ---- -6: PCR5PrimerCorrect, Aligner.cpp:1996-2107.  The search ran with -s raised by -6; an accepted alignment with more
//      mismatches than -s allows is kept only if replacing mismatching bases among the first 12 of the read by the
//      target's (5' random-primer artefacts) brings it within -s: the read is rewritten and LowMMCnt / Mismatches
//      lowered; otherwise it becomes eNARNoHit.  Seg[0].TrimMismatches -- what the CSV reports -- keeps its value.
static void pcr_5prime_correct(Records& rc) {
  const Opts& o = rc.o;
  auto& res = rc.res;
  const uint32_t nrec = rc.n();
  const int KLen = 12;
  diag("Starting PCR 5' %dbp primer correction processing targeting substitution rate of %d ...", KLen, o.max_subs);
  rc.trim_mm.resize(nrec);
  for (uint32_t i = 0; i < nrec; ++i) rc.trim_mm[i] = res[i].mismatches;
  int reads_fixed = 0, bases_fixed = 0, rejected = 0;
  for (uint32_t i = 0; i < nrec; ++i) {
    bkx_read_result& r = res[i];
    if (r.nar != BKX_NAR_ACCEPTED) continue;
    const uint32_t ri = rc.rix(i);
    const int L = rc.R.len(ri), max_mm = (o.max_subs * L + 50) / 100;
    if (r.low_mm <= max_mm || r.match_len != L) continue;
    uint8_t* b = rc.R.bases.data() + rc.R.offs[ri];
    const uint8_t* g = rc.genome[r.chrom_id].data() + r.match_loci;
    auto targ = [&](int q) -> uint8_t {
      if (r.strand != '-') return g[q];
      const uint8_t c = g[L - 1 - q];
      return c < 4 ? (uint8_t)(3 - c) : c;
    };
    int cur = r.low_mm;
    for (int q = 0; q < KLen && q < L; ++q)
      if ((b[q] & 7) != targ(q) && --cur <= max_mm) break;
    if (cur > max_mm) { r.num_hits = 0; r.nar = BKX_NAR_NOHIT; ++rejected; continue; }
    cur = r.low_mm;
    for (int q = 0; q < KLen && q < L; ++q)
      if ((b[q] & 7) != targ(q)) {
        b[q] = (uint8_t)((b[q] & 0xf8) | targ(q));
        ++bases_fixed;
        if (--cur <= max_mm) break;
      }
    r.low_mm = (int8_t)cur;
    r.mismatches = (uint8_t)cur;
    ++reads_fixed;
  }
  diag("Completed PCR 5' primer correction, %d reads with %d bases corrected, %d reads with excessive substitutions rejected", reads_fixed, bases_fixed, rejected);
}

// ---- -x: AutoTrimFlanks, Aligner.cpp:1608-1812.  Each accepted alignment is cut back from both ends to the first run
//      of MinFlankExacts matching bases; the rest must keep at least half the read (>= 15 bp) or the read is sloughed
//      as eNARTrim.  In paired-end runs the 5' scan stays inside the first third of the read and the 3' scan inside
//      the last third, and nothing is sloughed (the pairing stands).  TrimLeft / TrimRight are in READ orientation
//      (Aligner.cpp:1528-1549).
static void auto_trim_flanks(Records& rc) {
  const Opts& o = rc.o;
  const Reads& R = rc.R;
  auto& res = rc.res;
  const uint32_t nrec = rc.n();
  const auto& genome = rc.genome;
  auto &trim_l = rc.trim_l, &trim_r = rc.trim_r;
  auto& trim_mm = rc.trim_mm;
  const unsigned fmt_threads = rc.threads;
  diag("Autotrim aligned read flank processing started..");
  diag("Starting 5' and 3' flank sequence autotrim processing...");
  trim_l.assign(nrec, 0); trim_r.assign(nrec, 0);
  const bool fresh_mm = trim_mm.empty();   // after -6 TrimMismatches already holds the count from before the correction
  if (fresh_mm) trim_mm.assign(nrec, 0);
  std::vector<uint32_t> ep(fmt_threads, 0), em(fmt_threads, 0);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < fmt_threads; ++t)
    th.emplace_back([&, t]() {
      std::vector<uint8_t> rd, tg;
      const uint32_t b0 = (uint32_t)((uint64_t)nrec * t / fmt_threads), e0 = (uint32_t)((uint64_t)nrec * (t + 1) / fmt_threads);
      for (uint32_t i = b0; i < e0; ++i) {
        bkx_read_result& r = res[i];
        if (fresh_mm) trim_mm[i] = r.mismatches;
        if (r.nar != BKX_NAR_ACCEPTED) continue;
        const int L = r.match_len, minlen = std::max(15, (L + 1) / 2), X = o.min_flank;
        const uint8_t* b = R.bases.data() + R.offs[rc.rix(i)];
        const uint8_t* g = genome[r.chrom_id].data() + r.match_loci;
        rd.resize(L); tg.resize(L);
        for (int q = 0; q < L; ++q) rd[q] = b[q] & 7;
        if (r.strand == '-') for (int q = 0; q < L; ++q) { uint8_t c = g[L - 1 - q] & 7; tg[q] = c < 4 ? 3 - c : c; }
        else for (int q = 0; q < L; ++q) tg[q] = g[q] & 7;
        auto slough = [&]() { r.num_hits = 0; r.nar = BKX_NAR_TRIM; ++(r.strand == '+' ? ep[t] : em[t]); };
        const bool pe = o.pe_mode != 0;
        int exact = 0, tmm = 0, k;
        for (k = 0; k <= L - minlen && k < (pe ? L / 3 : L); ++k) {       // 5' -> 3'
          if (rd[k] != tg[k]) { exact = 0; ++tmm; continue; }
          if (++exact == X) break;
        }
        if (!pe && (k + minlen > L || exact < X)) { slough(); continue; }
        const int left = k - (X - 1);
        exact = 0;
        for (k = L - 1; k >= left + minlen && k > (pe ? (L * 2) / 3 : 0); --k) {  // 3' -> 5'
          if (rd[k] != tg[k]) { exact = 0; ++tmm; continue; }
          if (++exact == X) break;
        }
        if (!pe && (exact != X || k < left + minlen)) { slough(); continue; }
        const int right = k + X;
        trim_l[i] = (uint16_t)left; trim_r[i] = (uint16_t)(L - right);
        if (left || L - right) trim_mm[i] = (uint8_t)(r.mismatches - tmm);
      }
    });
  for (auto& x : th) x.join();
  for (unsigned t = 0; t < fmt_threads; ++t) { rc.elim_plus += ep[t]; rc.elim_minus += em[t]; }
  diag("Finished 5' and 3' flank sequence autotriming, %d plus strand and %d minus strand aligned reads removed", (int)rc.elim_plus, (int)rc.elim_minus);
  diag("Autotrim aligned read flank processing completed");
}

// ---- -Z / -z: FiltByChroms, Aligner.cpp:4019-4124.  A chromosome matching an include expression stays; with no
//      include expressions given it stays unless an exclude expression matches; otherwise its alignments become
//      eNARChromFilt.  (Exclude expressions are not consulted once include expressions exist -- as in the reference.)
static void filter_by_chroms(Records& rc, std::vector<regex_t>& rin, std::vector<regex_t>& rex) {
  auto& res = rc.res;
  const auto& ents = rc.ents;
  const uint32_t nrec = rc.n(), num_entries = (uint32_t)ents.size() - 1;
  diag("Filtering aligned reads by chromosome started..");
  diag("Now filtering matches by chromosome");
  std::vector<char> keep(num_entries + 1, 1);
  for (uint32_t e = 1; e <= num_entries; ++e) {
    regmatch_t mc;
    bool ok = false;
    for (auto& re : rin) if (!regexec(&re, ents[e].name, 1, &mc, 0)) { ok = true; break; }
    if (!ok && rin.empty()) {
      ok = true;
      for (auto& re : rex) if (!regexec(&re, ents[e].name, 1, &mc, 0)) { ok = false; break; }
    }
    keep[e] = ok;
  }
  for (auto& re : rin) regfree(&re);
  for (auto& re : rex) regfree(&re);
  int removed = 0;
  for (uint32_t i = 0; i < nrec; ++i) {
    bkx_read_result& r = res[i];
    if (r.nar == BKX_NAR_ACCEPTED && !keep[r.chrom_id]) { r.nar = BKX_NAR_CHROMFILT; r.num_hits = 0; r.low_hit_instances = 0; ++removed; }
  }
  diag("Filtering by chromosome completed - removed %d  matches", removed);
  diag("Filtering aligned reads by chromosome completed");
}

// ---- -B: priority regions (CAligner::Align loads them as a CBEDfile, Aligner.cpp:280-299; BED lines "chrom start end
//      [name [score [strand ...]]]" separated by white space -- or by commas when the first feature line does not parse
//      otherwise --, '#' comments, header lines in front of the first feature; BEDfile.cpp:760-905).  A feature covers
//      [start, end); chromosome names are compared without case, "chloroplast" / "ChrC" and "mitochondria" / "ChrM" are
//      synonyms (LocateChromIDbyName, BEDfile.cpp:2588-2613).  Kept per chromosome of the index as sorted, merged intervals.
struct PriorityRegions {
  std::vector<std::vector<std::pair<uint32_t, uint32_t>>> by_chrom;   // [chrom id]: inclusive [first, last], disjoint, ascending
  // InAnyFeature: does [first, last] overlap a feature of this chromosome (BEDfile.cpp:3509-3526)
  bool in_any(uint32_t chrom, uint32_t first, uint32_t last) const {
    if (chrom >= by_chrom.size()) return false;
    const auto& v = by_chrom[chrom];
    auto it = std::lower_bound(v.begin(), v.end(), first, [](const std::pair<uint32_t, uint32_t>& a, uint32_t x) { return a.second < x; });
    return it != v.end() && it->first <= last;
  }
};

static int load_priority_regions(const std::string& path, const std::vector<bkx_entry>& ents, PriorityRegions& out) {
  diag("Loading high priority regions BED file '%s'", path.c_str());
  FILE* fp = fopen(path.c_str(), "r");
  if (!fp) {
    diag("Unable to open high priority regions BED file '%s'", path.c_str());
    return -1;
  }
  struct Feat { std::string chrom; long start, end; };
  std::vector<Feat> feats;
  char line[16384];
  int line_num = 0;
  bool csv = false, bad = false;
  while (fgets(line, sizeof(line) - 1, fp)) {
    ++line_num;
    if (feats.empty() && line_num >= 20) { bad = true; break; }   // no feature in the first 20 lines: not a BED file
    char* t = line;
    while (*t && isspace((unsigned char)*t)) ++t;
    size_t len = strlen(t);
    while (len && isspace((unsigned char)t[len - 1])) t[--len] = '\0';
    if (*t == '\0' || *t == '#') continue;
    char chrom[300];
    int a = 0, b = 0, cnt = 0;
    if (!csv) {
      cnt = sscanf(t, " %140s %d %d", chrom, &a, &b);
      if (feats.empty() && cnt < 3) csv = true;
    }
    if (csv) cnt = sscanf(t, " %140s , %d , %d", chrom, &a, &b);
    if (cnt < 3) {
      if (feats.empty()) { csv = false; continue; }   // could be a header line
      bad = true;
      break;
    }
    chrom[35] = '\0';   // cMaxDatasetSpeciesChrom - 1
    feats.push_back({chrom, a, b});
  }
  fclose(fp);
  if (bad || feats.empty()) {
    diag("Unable to open high priority regions BED file '%s'", path.c_str());
    return -1;
  }
  auto canon = [](const char* name) {   // one spelling per synonym class, lower case
    std::string s(name);
    for (auto& ch : s) ch = (char)tolower((unsigned char)ch);
    if (s == "chloroplast") s = "chrc";
    if (s == "mitochondria") s = "chrm";
    return s;
  };
  out.by_chrom.assign(ents.size(), {});
  for (size_t e = 1; e < ents.size(); ++e) {
    // LocateChromIDbyName resolves the name of the INDEX chromosome against the BED names: the name itself first
    const std::string want = canon(ents[e].name);
    std::string exact(ents[e].name);
    for (auto& ch : exact) ch = (char)tolower((unsigned char)ch);
    bool have_exact = false;
    for (const auto& f : feats) {
      std::string fl(f.chrom);
      for (auto& ch : fl) ch = (char)tolower((unsigned char)ch);
      if (fl == exact) { have_exact = true; break; }
    }
    for (const auto& f : feats) {
      std::string fl(f.chrom);
      for (auto& ch : fl) ch = (char)tolower((unsigned char)ch);
      const bool match = have_exact ? fl == exact : canon(f.chrom.c_str()) == want && fl != exact;
      if (!match || f.end <= f.start || f.end <= 0) continue;
      out.by_chrom[e].push_back({(uint32_t)std::max<long>(0, f.start), (uint32_t)(f.end - 1)});
    }
    auto& v = out.by_chrom[e];
    std::sort(v.begin(), v.end());
    size_t w = 0;
    for (size_t i = 0; i < v.size(); ++i) {
      if (w && v[i].first <= v[w - 1].second + 1) v[w - 1].second = std::max(v[w - 1].second, v[i].second);
      else v[w++] = v[i];
    }
    v.resize(w);
  }
  diag("High priority regions BED file '%s' loaded", path.c_str());
  return 0;
}

// FiltByPriorityRegions, Aligner.cpp:4126-4186: accepted alignments outside every priority region become eNARRegionFilt
static void filter_by_priority_regions(Records& rc, const PriorityRegions& pr) {
  diag("Now filtering matches by prioritorised regions");
  uint32_t kept = 0, removed = 0;
  for (uint32_t i = 0; i < rc.n(); ++i) {
    bkx_read_result& r = rc.res[i];
    if (r.nar != BKX_NAR_ACCEPTED) continue;
    if (pr.in_any(r.chrom_id, r.match_loci, r.match_loci + r.match_len - 1)) { ++kept; continue; }
    ++removed;
    r.nar = BKX_NAR_REGIONFILT;
    r.num_hits = 0;
    r.chrom_id = 0;
  }
  diag("Filtering by prioritorised regions completed - retained %u, removed %u matches", kept, removed);
}

// ---- simulated-read truth check of ReportAlignStats, Aligner.cpp:3556-3657.  Reads written by `biokanga simreads` name
//      their origin in the descriptor: <type>|usimreads|<n>|<chrom>|<start>|<end>|<len>|<strand>|<errs>..., where <chrom> may
//      itself be three '|' separated parts (gnl|UG|Ta#S58887126).  Records are walked in load order; the first accepted one
//      decides whether the set is simulated, and from then on every accepted record is parsed -- until one does not parse,
//      which ends the checking for good (later records are skipped because they are not the first accepted one any more).
//      An accepted alignment on another chromosome, or with neither edge on the source's edges, is flagged (HitLoci.FlagIA)
//      and reported as "iar" by the CSV / BED writers; the edges compared are the untrimmed ones.
struct SimTruth { bool sim = false; uint32_t edge2 = 0, edge1 = 0, misaligned = 0; };

static SimTruth sim_truth_check(Records& rc) {
  SimTruth t;
  const uint32_t nrec = rc.n();
  uint32_t accepted = 0;
  for (uint32_t i = 0; i < nrec; ++i) {
    const bkx_read_result& r = rc.res[i];
    if (r.nar != BKX_NAR_ACCEPTED) continue;
    ++accepted;
    if (!t.sim && accepted != 1) continue;
    const char* d = rc.R.name(rc.rix(i));
    char type[100], c0[100], c1[100], c2[100], strand;
    int seq, start = 0, end = 0, len, errs;
    std::string chrom;
    int its = sscanf(d, "%99[^|]|usimreads|%d|%99[^|]|%d|%d|%d|%c|%d", type, &seq, c0, &start, &end, &len, &strand, &errs);
    if (its >= 6) { t.sim = true; chrom = c0; }
    else {
      its = sscanf(d, "%99[^|]|usimreads|%d|%99[^|]|%99[^|]|%99[^|]|%d|%d|%d|%c|%d", type, &seq, c0, c1, c2, &start, &end, &len, &strand, &errs);
      t.sim = its >= 8;
      if (t.sim) { chrom = c0; chrom += '|'; chrom += c1; chrom += '|'; chrom += c2; }
    }
    if (!t.sim) continue;
    if (rc.flag_ia.empty()) rc.flag_ia.assign(nrec, 0);
    const int left = (int32_t)r.match_loci, right = left + (int)r.match_len - 1;
    if (strcasecmp(rc.ents[r.chrom_id].name, chrom.c_str())) { ++t.misaligned; rc.flag_ia[i] = 1; }
    else if (left == start || right == end) ++(left == start && right == end ? t.edge2 : t.edge1);
    else { ++t.misaligned; rc.flag_ia[i] = 1; }
  }
  return t;
}

// ---- writers.  Rows go out in hit order (`order`); every writer reports the alignment as trimmed by -x, if at all.
static const char kAsc[] = "ACGTN";

static void write_bed(const Records& rc, const std::vector<uint32_t>& order, const bkx_index_info& info, OutBuf& ob) {
  const Opts& o = rc.o;
  const auto& res = rc.res;
  const auto& ents = rc.ents;
  const uint32_t nrec = rc.n();
  const unsigned fmt_threads = rc.threads;
  auto adj_start = [&](uint32_t i) { return rc.adj_start(i); };
  auto adj_len = [&](uint32_t i) { return rc.adj_len(i); };
  // UCSC BED: track line then chrom, start, end+1, "ar", score, strand (Aligner.cpp:6355-6362, 6463-6466)
  const char* title = o.title.empty() ? "kanga" : o.title.c_str();
  // under -r5 the reference writes the track line twice: when it creates the file and when it reports (Aligner.cpp:4405, 6358)
  for (int rep = 0; rep < ((o.ml_mode == BKX_ML_ALL) ? 2 : 1); ++rep) { ob.s += "track type=bed name=\""; ob.s += title; ob.s += "\" description=\""; ob.s += title; ob.s += "\"\n"; }
  emit_rows(ob, nrec, fmt_threads, [&](uint32_t k, std::string& s) {
    uint32_t i = order[k];
    const bkx_read_result& r = res[i];
    if (r.nar != BKX_NAR_ACCEPTED) return;
    s += ents[r.chrom_id].name; s += '\t';
    append_uint(s, adj_start(i)); s += '\t';
    append_uint(s, (uint64_t)adj_start(i) + adj_len(i)); s += '\t'; s += rc.align_type(i); s += "\t0\t"; s += (char)r.strand; s += '\n';
  });
}


static void write_csv(const Records& rc, const std::vector<uint32_t>& order, const bkx_index_info& info, OutBuf& ob) {
  const Opts& o = rc.o;
  const Reads& R = rc.R;
  const auto& res = rc.res;
  const auto& ents = rc.ents;
  const auto& genome = rc.genome;
  const uint32_t nrec = rc.n();
  const unsigned fmt_threads = rc.threads;
  auto rix = [&](uint32_t i) { return rc.rix(i); };
  auto tleft = [&](uint32_t i) { return rc.tleft(i); };
  auto adj_start = [&](uint32_t i) { return rc.adj_start(i); };
  auto adj_len = [&](uint32_t i) { return rc.adj_len(i); };
  auto adj_mm = [&](uint32_t i) { return rc.adj_mm(i); };
  // ReadID,"ar","species","chrom",start,end,len,"strand",score,0,NumReads,TrimMismatches,"N/A","descriptor"
  // -M2/-M3 append the read sequence, -M1/-M3 the matched genome sequence in read orientation (Aligner.cpp:6612-6621)
  emit_rows(ob, nrec, fmt_threads, [&](uint32_t k, std::string& s) {
    uint32_t i = order[k];
    const bkx_read_result& r = res[i];
    if (r.nar != BKX_NAR_ACCEPTED) return;
    append_uint(s, (uint64_t)i + 1);
    s += ",\""; s += rc.align_type(i); s += "\",\""; s += info.dataset_name; s += "\",\""; s += ents[r.chrom_id].name; s += "\",";
    const uint32_t start = adj_start(i), alen = adj_len(i);   // the alignment as trimmed by -x, if at all
    append_uint(s, start); s += ',';
    append_uint(s, (uint64_t)start + alen - 1); s += ',';
    append_uint(s, alen); s += ",\""; s += (char)r.strand; s += "\",0,0,1,";
    const uint32_t ri = rix(i);
    append_uint(s, adj_mm(i)); s += ",\"N/A\",\""; s += R.name(ri); s += '"';
    if (o.fmt >= 2) {
      s += ",\"";
      const uint8_t* b = R.bases.data() + R.offs[ri] + tleft(i);
      for (uint32_t q = 0; q < alen; ++q) { uint8_t c = b[q] & 7; s += c < 4 ? kAsc[c] : 'N'; }
      s += '"';
    }
    if (o.fmt == 1 || o.fmt == 3) {
      s += ",\"";
      const uint8_t* g = genome[r.chrom_id].data() + start;
      if (r.strand == '-') for (int q = (int)alen - 1; q >= 0; --q) { uint8_t c = g[q] & 7; s += c < 4 ? kAsc[3 - c] : 'N'; }
      else for (uint32_t q = 0; q < alen; ++q) { uint8_t c = g[q] & 7; s += c < 4 ? kAsc[c] : 'N'; }
      s += '"';
    }
    s += '\n';
  });
}


static bool write_bam(const Records& rc, const std::vector<uint32_t>& order, const bkx_index_info& info, OutBuf& ob) {
  const Opts& o = rc.o;
  const Reads& R = rc.R;
  const auto& res = rc.res;
  const auto& ents = rc.ents;
  const uint32_t nrec = rc.n();
  const unsigned fmt_threads = rc.threads;
  auto rix = [&](uint32_t i) { return rc.rix(i); };
  auto tleft = [&](uint32_t i) { return rc.tleft(i); };
  auto tright = [&](uint32_t i) { return rc.tright(i); };
  auto adj_start = [&](uint32_t i) { return rc.adj_start(i); };
  auto adj_len = [&](uint32_t i) { return rc.adj_len(i); };
  // ---- BAM + BAI (same record content as the SAM branch below, binary form)
  ob.close();
  Bgzf bz;
  if (!bz.open(o.out)) { diag("Fatal: unable to create '%s'", o.out.c_str()); return false; }
  std::vector<char> hit(info.num_entries + 1, 0);
  for (uint32_t i = 0; i < nrec; ++i) if (res[i].nar == BKX_NAR_ACCEPTED) hit[res[i].chrom_id] = 1;
  bool all = (uint32_t)o.sam_seq_thres >= info.num_entries;
  std::vector<int> refid(info.num_entries + 1, -1);
  std::string text = "@HD\tVN:1.4\tSO:coordinate", refs;
  uint32_t nref = 0;
  for (uint32_t e = 1; e <= info.num_entries; ++e)
    if (all || hit[e]) {
      text += "\n@SQ\tAS:"; text += info.dataset_name; text += "\tSN:"; text += ents[e].name; text += "\tLN:";
      append_uint(text, ents[e].seq_len);
      refid[e] = (int)nref++;
      uint32_t ln = (uint32_t)strlen(ents[e].name) + 1, sl = ents[e].seq_len;
      refs.append((const char*)&ln, 4); refs.append(ents[e].name, ln); refs.append((const char*)&sl, 4);
    }
  text += "\n@PG\tID:biokanga\tVN:4.4.2\n";
  std::string hdr = "BAM\1";
  uint32_t lt = (uint32_t)text.size();
  hdr.append((const char*)&lt, 4); hdr += text; hdr.append((const char*)&nref, 4); hdr += refs;
  // The whole uncompressed BAM stream is laid out first (record sizes -> offsets -> records written in place by all
  // threads), then cut into BGZF blocks exactly where the reference's sequential writer cuts them (every 0xff00
  // bytes, plus once behind the last aligned record), the blocks are deflated in parallel and written in order; the
  // BAI follows from the record offsets and the block addresses.
  BaiBuilder bai;
  bai.out = "BAI\1";
  bai.put32(nref);
  struct RecMeta { uint64_t uofs; uint32_t len; int32_t rid, pos; int32_t L, alen; uint16_t lead, trail; uint8_t acc; };
  std::vector<RecMeta> meta;
  meta.reserve(nrec);
  uint64_t utotal = hdr.size();
  for (uint32_t k = 0; k < nrec; ++k) {
    uint32_t i = order[k];
    const bkx_read_result& r = res[i];
    bool acc = r.nar == BKX_NAR_ACCEPTED;
    if (!acc && o.fmt != 6) continue;
    const uint32_t ri = rix(i);
    const int L = R.len(ri);
    uint32_t lname = (uint32_t)strlen(R.name(ri)) + 1;
    // -x trims become soft clips either side of the M operation (Aligner.cpp:5960-5985)
    const uint32_t lead = !acc ? 0u : r.strand == '+' ? tleft(i) : tright(i), trail = !acc ? 0u : r.strand == '+' ? tright(i) : tleft(i);
    const uint32_t ncig = 1 + (lead ? 1 : 0) + (trail ? 1 : 0);
    uint32_t len = 4 + 32 + lname + 4 * ncig + (uint32_t)((L + 1) / 2) + (uint32_t)L + (acc ? 0u : 6u);
    meta.push_back({utotal, len, acc ? refid[r.chrom_id] : -1, acc ? (int32_t)adj_start(i) : -1, L, acc ? (int32_t)adj_len(i) : L,
                    (uint16_t)lead, (uint16_t)trail, (uint8_t)acc});
    utotal += len;
  }
  std::vector<uint8_t> U(utotal);
  memcpy(U.data(), hdr.data(), hdr.size());
  // record k' (index into meta) <- sorted record; meta and the sorted walk advance together
  std::vector<uint32_t> kept;
  kept.reserve(meta.size());
  for (uint32_t k = 0; k < nrec; ++k) {
    const bkx_read_result& r = res[order[k]];
    if (r.nar == BKX_NAR_ACCEPTED || o.fmt == 6) kept.push_back(order[k]);
  }
  auto write_record = [&](size_t mi) {
    const uint32_t i = kept[mi];
    const RecMeta& M = meta[mi];
    const bkx_read_result& r = res[i];
    const bool acc = M.acc != 0;
    int flags = 0, tlen = 0;
    long pnext = -1;
    if (!o.pe_mode) flags = acc ? (r.strand == '+' ? 0 : 0x10) : 0x04;
    else {
      bool pe2 = i & 1;
      const bkx_read_result& m = res[pe2 ? i - 1 : i + 1];
      flags = 0x01 | 0x02 | (pe2 ? 0x80 : 0x40);
      if (acc) flags |= r.strand == '+' ? 0 : 0x10; else flags |= 0x04;
      bool both = (r.flags & BKX_FLG_PE_ALIGNED) && (m.flags & BKX_FLG_PE_ALIGNED) && m.nar == BKX_NAR_ACCEPTED;
      if (both) {
        flags |= m.strand == '+' ? 0 : 0x20;
        if (acc) {
          const uint32_t mi2 = pe2 ? i - 1 : i + 1;   // both alignments as trimmed by -x, if at all
          long se = adj_start(i), pes = adj_start(mi2);
          tlen = se <= pes ? (int)(pes - se) + (int)adj_len(mi2) : (int)(se - pes) + (int)adj_len(i);
          pnext = pes;
        }
      } else flags |= 0x08;
    }
    const uint32_t ri = rix(i);
    const int L = M.L;
    const uint8_t* b = R.bases.data() + R.offs[ri];
    const char* qn = R.name(ri);
    uint32_t lname = (uint32_t)strlen(qn) + 1;
    int32_t rid = M.rid, pos = M.pos;
    const uint32_t ncig = 1 + (M.lead ? 1 : 0) + (M.trail ? 1 : 0);
    uint32_t bin = acc ? (uint32_t)bai_reg2bin(pos, pos + M.alen) : 0;
    uint32_t bmn = bin << 16 | 255u << 8 | lname, fnc = (uint32_t)flags << 16 | ncig;
    int32_t nrid = (acc && pnext >= 0) ? rid : -1, npos = acc ? (int32_t)pnext : -1, tl = acc ? tlen : 0, lseq = L;
    uint32_t cigar = (uint32_t)M.alen << 4, clip_lead = (uint32_t)M.lead << 4 | 4u, clip_trail = (uint32_t)M.trail << 4 | 4u;
    uint8_t* w = U.data() + M.uofs;
    uint32_t bsz = M.len - 4;
    auto p32 = [&](const void* v) { memcpy(w, v, 4); w += 4; };
    p32(&bsz); p32(&rid); p32(&pos); p32(&bmn); p32(&fnc); p32(&lseq); p32(&nrid); p32(&npos); p32(&tl);
    memcpy(w, qn, lname); w += lname;
    if (M.lead) p32(&clip_lead);
    p32(&cigar);
    if (M.trail) p32(&clip_trail);
    bool rc = acc && r.strand != '+';
    for (int q = 0; q < L; q += 2) {
      auto nib = [&](int idx) -> unsigned {
        if (idx >= L) return 0u;
        uint8_t c = b[rc ? L - 1 - idx : idx] & 7;
        if (c < 4) { if (rc) c = 3 - c; return 1u << c; }
        return 15u;
      };
      *w++ = (uint8_t)(nib(q) << 4 | nib(q + 1));
    }
    int sumq = 0;
    for (int q = 0; q < L; ++q) sumq += (b[q] >> 4) & 0x0f;
    if (sumq == 0) { memset(w, 0xff, (size_t)L); w += L; }
    else for (int q = 0; q < L; ++q) *w++ = (uint8_t)(33 + (((b[rc ? L - 1 - q : q] >> 4) & 0x0f) * 40) / 15);
    if (!acc) { memcpy(w, "YUZ", 3); w += 3; memcpy(w, kNarCode[r.nar], 2); w += 2; *w++ = 0; }
  };
  {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < fmt_threads; ++t)
      th.emplace_back([&, t]() {
        size_t b0 = meta.size() * t / fmt_threads, e0 = meta.size() * (t + 1) / fmt_threads;
        for (size_t mi = b0; mi < e0; ++mi) write_record(mi);
      });
    for (auto& x : th) x.join();
  }
  // explicit flush behind the last aligned record (bLastAligned, SAMfile.cpp), if any
  uint64_t flush_at = 0;
  bool have_flush = false;
  for (size_t mi = meta.size(); mi-- > 0;)
    if (meta[mi].acc) { flush_at = meta[mi].uofs + meta[mi].len; have_flush = true; break; }
  bool ok = bz.write_stream(U.data(), U.size(), have_flush ? &flush_at : nullptr, fmt_threads);
  // BAI bins end at 512 Mbp.  For longer reference sequences the reference switches to a CSI index (SAMfile.cpp:1602-1622)
  // but never opens that file -- the branch that would (:1664) hangs off `if (type >= BAI)` and is unreachable -- so its
  // first index flush fails (WriteIdxToDisk, :1812).  Here the BAM is written complete and the index is left to
  // `samtools index -c`, with a note in the log, rather than an index with overflowing bins.
  uint32_t longest_ref = 0;
  for (uint32_t e = 1; e <= info.num_entries; ++e) if (refid[e] >= 0) longest_ref = std::max(longest_ref, ents[e].seq_len);
  const bool write_bai = longest_ref < 0x20000000u;
  if (!write_bai) diag("Note: reference sequences of 512Mbp or more (max %u): no BAI index written for '%s', index it with a CSI indexer", longest_ref, o.out.c_str());
  int cur_ref = -1;  // reference whose index block is being accumulated
  for (size_t mi = 0; mi < meta.size() && ok && write_bai; ++mi) {
    const RecMeta& M = meta[mi];
    if (!M.acc) continue;
    while (cur_ref < M.rid) { if (cur_ref >= 0) bai.end_ref(); ++cur_ref; }
    bai.add(bz.vaddr(M.uofs), (uint32_t)M.pos, bz.vaddr(M.uofs + M.len), (uint32_t)(M.pos + M.alen - 1));
  }
  bai.end_ref();  // Close(): the reference in progress (an empty block when nothing aligned)
  ok = ok && bz.close();
  if (write_bai) {
    FILE* fb = fopen((o.out + ".bai").c_str(), "wb");
    if (fb) { fwrite(bai.out.data(), 1, bai.out.size(), fb); fclose(fb); } else ok = false;
  }
  if (!ok) { diag("Fatal: write to '%s' failed", o.out.c_str()); return false; }
  return true;
}


static void write_sam(const Records& rc, const std::vector<uint32_t>& order, const bkx_index_info& info, OutBuf& ob) {
  const Opts& o = rc.o;
  const Reads& R = rc.R;
  const auto& res = rc.res;
  const auto& ents = rc.ents;
  const uint32_t nrec = rc.n();
  const unsigned fmt_threads = rc.threads;
  auto rix = [&](uint32_t i) { return rc.rix(i); };
  auto tleft = [&](uint32_t i) { return rc.tleft(i); };
  auto tright = [&](uint32_t i) { return rc.tright(i); };
  auto adj_start = [&](uint32_t i) { return rc.adj_start(i); };
  auto adj_len = [&](uint32_t i) { return rc.adj_len(i); };
  // SAM header: @HD, @SQ for every chromosome (or only those hit if more than -4 threshold), @PG
  std::vector<char> hit(info.num_entries + 1, 0);
  for (uint32_t i = 0; i < nrec; ++i) if (res[i].nar == BKX_NAR_ACCEPTED) hit[res[i].chrom_id] = 1;
  bool all = (uint32_t)o.sam_seq_thres >= info.num_entries;
  ob.s += "@HD\tVN:1.4\tSO:coordinate\n";
  for (uint32_t e = 1; e <= info.num_entries; ++e)
    if (all || hit[e]) {
      ob.s += "@SQ\tAS:"; ob.s += info.dataset_name; ob.s += "\tSN:"; ob.s += ents[e].name; ob.s += "\tLN:";
      append_uint(ob.s, ents[e].seq_len); ob.s += '\n';
    }
  ob.s += "@PG\tID:biokanga\tVN:4.4.2\n";
  // a run without a single read: the reference's -M6 writer still walks one (empty, unprocessed) record
  if (nrec == 0 && o.fmt == 6 && R.n() == 0) ob.s += "\t4\t*\t0\t255\t0M\t*\t0\t0\t\t\t\tYU:Z:NA\n";
  emit_rows(ob, nrec, fmt_threads, [&](uint32_t k, std::string& s) {
    uint32_t i = order[k];
    const bkx_read_result& r = res[i];
    bool acc = r.nar == BKX_NAR_ACCEPTED;
    if (!acc && o.fmt != 6) return;
    int flags = 0, tlen = 0;
    long pnext = -1;
    if (!o.pe_mode) {
      flags = acc ? (r.strand == '+' ? 0 : 0x10) : 0x04;
    } else {  // ReportBAMread, Aligner.cpp:5864-5925
      bool pe2 = i & 1;
      const bkx_read_result& m = res[pe2 ? i - 1 : i + 1];
      flags = 0x01 | 0x02 | (pe2 ? 0x80 : 0x40);
      if (acc) flags |= r.strand == '+' ? 0 : 0x10; else flags |= 0x04;
      bool both = (r.flags & BKX_FLG_PE_ALIGNED) && (m.flags & BKX_FLG_PE_ALIGNED) && m.nar == BKX_NAR_ACCEPTED;
      if (both) {
        flags |= m.strand == '+' ? 0 : 0x20;
        if (acc) {
          const uint32_t mi2 = pe2 ? i - 1 : i + 1;   // both alignments as trimmed by -x, if at all
          long se = adj_start(i), pes = adj_start(mi2);
          tlen = se <= pes ? (int)(pes - se) + (int)adj_len(mi2) : (int)(se - pes) + (int)adj_len(i);
          pnext = pes;
        }
      } else flags |= 0x08;
    }
    const uint32_t ri = rix(i);
    s += R.name(ri); s += '\t';
    append_uint(s, (uint64_t)flags); s += '\t';
    if (acc) { s += ents[r.chrom_id].name; s += '\t'; append_uint(s, (uint64_t)adj_start(i) + 1); }
    else s += "*\t0";
    s += "\t255\t";
    if (acc) {  // flanks trimmed by -x are soft clipped, in reference orientation (Aligner.cpp:5960-5985)
      const uint32_t lead = r.strand == '+' ? tleft(i) : tright(i), trail = r.strand == '+' ? tright(i) : tleft(i);
      if (lead) { append_uint(s, lead); s += 'S'; }
      append_uint(s, adj_len(i)); s += 'M';
      if (trail) { append_uint(s, trail); s += 'S'; }
      s += '\t';
    } else { append_uint(s, (uint64_t)R.len(ri)); s += "M\t"; }
    if (acc && pnext >= 0) { s += "=\t"; append_uint(s, (uint64_t)pnext + 1); s += '\t'; append_uint(s, (uint64_t)tlen); s += '\t'; }
    else s += "*\t0\t0\t";
    const uint8_t* b = R.bases.data() + R.offs[ri];
    int L = R.len(ri);
    if (acc && r.strand != '+') {
      for (int q = L - 1; q >= 0; --q) { uint8_t c = b[q] & 7; s += c < 4 ? kAsc[3 - c] : 'N'; }
    } else {
      for (int q = 0; q < L; ++q) { uint8_t c = b[q] & 7; s += c < 4 ? kAsc[c] : 'N'; }
    }
    // QUAL: '*' unless qualities were kept (-g0..2); 33 + q4*40/15, reversed with the sequence (Aligner.cpp:5929-5955)
    int sumq = 0;
    for (int q = 0; q < L; ++q) sumq += (b[q] >> 4) & 0x0f;
    s += '\t';
    if (sumq == 0) s += '*';
    else if (acc && r.strand != '+') for (int q = L - 1; q >= 0; --q) s += (char)(33 + (((b[q] >> 4) & 0x0f) * 40) / 15);
    else for (int q = 0; q < L; ++q) s += (char)(33 + (((b[q] >> 4) & 0x0f) * 40) / 15);
    if (!acc) { s += "\t\tYU:Z:"; s += kNarCode[r.nar]; }
    s += '\n';
  });
}

// ---- -j / -J: ReportNoneAligned / ReportMultiAlign, Aligner.cpp:3826-4016 -- FASTA, 70 columns, in hit order
template <class Want>
static bool report_reads(const Records& rc, const std::vector<uint32_t>& order, const std::string& path, const char* cls, Want&& want) {
  const Reads& R = rc.R;
  const auto& res = rc.res;
  OutBuf fb;
  if (!fb.open(path)) { diag("Fatal: unable to create '%s'", path.c_str()); return false; }
  emit_rows(fb, rc.n(), rc.threads, [&](uint32_t k, std::string& s) {
    const uint32_t i = order[k];
    if (!want(res[i])) return;
    const uint32_t ri = rc.rix(i);
    const int L = R.len(ri);
    s += ">lcl|"; s += cls; s += '|'; append_uint(s, (uint64_t)i + 1); s += ' '; s += R.name(ri); s += ' ';
    append_uint(s, (uint64_t)i + 1); s += "|1|"; append_uint(s, (uint64_t)L); s += '\n';
    const uint8_t* b = R.bases.data() + R.offs[ri];
    for (int q = 0; q < L; ++q) {
      uint8_t c = b[q] & 7;
      s += c < 4 ? "ACGT"[c] : 'N';
      if (q + 1 == L || (q + 1) % 70 == 0) s += '\n';
    }
  });
  fb.close();
  return true;
}

// ---- -O: WriteSubDist per reported alignment (Aligner.cpp:6275-6331), then WriteBasicCountStats (:4191-4332) and
//      ReportTargHitCnts (:5475-5537): per read offset the bases seen and the aligner induced substitutions by Phred
//      band of the 4-bit quality, the number of substitutions per alignment, alignments per target sequence.
static void write_statistics(const Records& rc, const std::vector<uint32_t>& multi_dist, FILE* stats_fp) {
  const Opts& o = rc.o;
  const Reads& R = rc.R;
  const auto& res = rc.res;
  const auto& ents = rc.ents;
  const auto& genome = rc.genome;
  const uint32_t nrec = rc.n(), num_entries = (uint32_t)ents.size() - 1;
  uint32_t max_len = 0;
  uint64_t n_acc = 0;
  for (uint32_t i = 0; i < nrec; ++i) if (res[i].nar == BKX_NAR_ACCEPTED) { ++n_acc; max_len = std::max<uint32_t>(max_len, (uint32_t)R.len(rc.rix(i))); }
  if (o.fmt == 4) max_len = 0;   // the BED branch of WriteReadHits never reaches WriteSubDist (Aligner.cpp:6448-6556): nothing beyond the histogram
  if (n_acc && max_len) {
    diag("Writing out basic count stats to file");
    std::vector<uint32_t> qinst(4 * (size_t)max_len, 0), qsubs(4 * (size_t)max_len, 0), msub((size_t)max_len + 1, 0), per_chrom(num_entries + 1, 0);
    for (uint32_t i = 0; i < nrec; ++i) {
      const bkx_read_result& r = res[i];
      if (r.nar != BKX_NAR_ACCEPTED || r.chrom_id == 0) continue;
      ++per_chrom[r.chrom_id];
      const uint32_t ri = rc.rix(i), L = (uint32_t)R.len(ri), start = rc.adj_start(i), alen = rc.adj_len(i);
      const uint8_t* b = R.bases.data() + R.offs[ri];
      const uint8_t* g = genome[r.chrom_id].data() + start;
      uint32_t subs = 0;
      for (uint32_t q = rc.tleft(i), k = 0; q < L - rc.tright(i); ++q, ++k) {
        const unsigned q4 = (b[q] >> 4) & 0x0f, band = q4 <= 3 ? 0 : q4 <= 7 ? 1 : q4 <= 11 ? 2 : 3;
        uint8_t t = r.strand == '-' ? g[alen - 1 - k] & 7 : g[k] & 7;
        if (r.strand == '-' && t < 4) t = 3 - t;
        ++qinst[band * (size_t)max_len + q];
        if ((b[q] & 7) != t) { ++qsubs[band * (size_t)max_len + q]; ++subs; }
      }
      ++msub[std::min<uint32_t>(subs, max_len)];
    }
    std::string s;
    auto row = [&](const char* label, auto&& value, uint32_t cnt) {
      s += label;
      for (uint32_t k = 0; k < cnt; ++k) { s += ','; append_uint(s, value(k)); }
    };
    if (o.ml_mode != BKX_ML_DEFAULT) {
      row("\"Multihit distribution\",", [](uint32_t k) { return (uint64_t)k + 1; }, (uint32_t)o.max_ml);
      row("\n,\"Instances\"", [&](uint32_t k) { return (uint64_t)multi_dist[k]; }, (uint32_t)o.max_ml);
      s += '\n';
    }
    static const char* kInstBand[4] = {"\n,\"Phred 0..9\"", "\n,\"Phred 10..19\"", "\n,\"Phred 20..29\"", "\n,\"Phred 30+\""};
    static const char* kSubsBand[4] = {"\n,\"Phred 0..8\"", "\n,\"Phred 9..19\"", "\n,\"Phred 20..29\"", "\n,\"Phred 30+\""};
    row("\"Phred Score Instances\",", [](uint32_t k) { return (uint64_t)k + 1; }, max_len);
    for (int band = 0; band < 4; ++band) row(kInstBand[band], [&](uint32_t k) { return (uint64_t)qinst[band * (size_t)max_len + k]; }, max_len);
    row("\n\n\"Aligner Induced Subs\",", [](uint32_t k) { return (uint64_t)k + 1; }, max_len);
    for (int band = 0; band < 4; ++band) row(kSubsBand[band], [&](uint32_t k) { return (uint64_t)qsubs[band * (size_t)max_len + k]; }, max_len);
    row("\n\n\"Multiple substitutions\",", [](uint32_t k) { return (uint64_t)k; }, max_len);
    row("\n,\"Instances\"", [&](uint32_t k) { return (uint64_t)msub[k]; }, max_len);
    s += '\n';
    diag("Reporting accepted read alignment counts on to targeted transcripts or sequences, sorting reads");
    diag("Completed sort");
    s += "\"TargSeq\",\"TargLen\",\"NumHits\"\n";
    int n_targ = 0;
    for (uint32_t e = 1; e <= num_entries; ++e)
      if (per_chrom[e]) {
        s += '"'; s += ents[e].name; s += "\","; append_uint(s, ents[e].seq_len); s += ','; append_uint(s, per_chrom[e]); s += '\n';
        ++n_targ;
      }
    fwrite(s.data(), 1, s.size(), stats_fp);
    diag("Completed reporting read alignment counts on to %d targeted transcripts or sequences", n_targ);
  }
}

int main(int argc, char** argv) {
  Opts o;
  int pr = parse(argc, argv, o);
  if (pr != 0) return pr < 0 ? 1 : 0;
  if (!o.logfile.empty()) g_log = fopen(o.logfile.c_str(), "w");
  auto t_start = std::chrono::steady_clock::now();
  diag("Subprocess align Version 4.4.2 (bkx B200 path) starting");
  if (!o.sqlite_file.empty())
    diag("Note: results summary database '%s' (-q, experiment '%s') is outside the accelerated path and is not written", o.sqlite_file.c_str(), o.exp_name.c_str());
  if (expand_input_specs(o) < 0) return 1;

  // ---- chromosome filter expressions are compiled up front, so that a malformed one stops the run before any work
  //      (CompileChromRegExprs, Aligner.cpp:4736-4798: POSIX extended, case insensitive)
  std::vector<regex_t> rin(o.incl.size()), rex(o.excl.size());
  {
    auto compile = [&](const std::vector<std::string>& src, std::vector<regex_t>& dst, const char* what) -> bool {
      for (size_t k = 0; k < src.size(); ++k) {
        int e = regcomp(&dst[k], src[k].c_str(), REG_EXTENDED | REG_ICASE);
        if (e) {
          char msg[128];
          regerror(e, &dst[k], msg, sizeof(msg));
          diag("Unable to process %s chrom '%s' error: %s", what, src[k].c_str(), msg);
          return false;
        }
      }
      return true;
    };
    if (!compile(o.incl, rin, "include") || !compile(o.excl, rex, "exclude")) return 1;
  }

  FILE* stats_fp = nullptr;   // created up front like every result file (CreateOrTruncResultFiles, Aligner.cpp:4351)
  if (!o.stats_file.empty() && !(stats_fp = fopen(o.stats_file.c_str(), "wb"))) { diag("Fatal: unable to create '%s'", o.stats_file.c_str()); return 1; }

  if (!o.contam_file.empty() && load_contaminants(o.contam_file, g_contam) < 0) return 1;

  // ---- reads load on their own thread while the index streams to the GPU (the reference also loads in the background)
  Reads R;
  int reads_rc = 0;
  std::thread reads_thread([&]() { reads_rc = load_reads(o, R); });
  struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } reads_joiner{reads_thread};

  // ---- index: one handle per GPU
  diag("Loading suffix array file '%s'", o.sfx.c_str());
  std::vector<bkx_index*> idx((size_t)o.gpus, nullptr);
  if (bkx_open_index(o.sfx.c_str(), 0, 0, &idx[0]) < 0) { diag("Fatal: %s", bkx_last_error()); return 1; }
  for (int g = 1; g < o.gpus; ++g)
    if (bkx_clone_index(idx[0], g, &idx[(size_t)g]) < 0) { diag("Fatal: %s", bkx_last_error()); return 1; }
  bkx_index_info info;
  bkx_index_info_get(idx[0], &info);
  diag("Genome Assembly Name: '%s' Descr: '%s' Title: '%s' Version: %d", info.dataset_name, info.dataset_name, info.dataset_name, info.version);
  bkx_align_params P;
  if (bkx_default_params(idx[0], o.pmode, &P) < 0) { diag("Fatal: %s", bkx_last_error()); return 1; }
  const int align_subs = o.pcr_primer > 0 ? std::min(o.max_subs + o.pcr_primer, 15) : o.max_subs;   // m_InitalAlignSubs, Aligner.cpp:208-213
  P.max_subs = align_subs; P.min_edit_dist = o.edit_delta; P.max_ns = o.max_ns; P.align_strand = o.strand;
  P.ml_mode = o.ml_mode; P.max_ml_matches = o.max_ml; P.clamp_max_ml = o.clamp_ml ? 1 : 0;
  P.best_matches = o.best_matches ? 1 : 0;

  std::vector<bkx_entry> ents(info.num_entries + 1);
  for (uint32_t e = 1; e <= info.num_entries; ++e) bkx_get_entry(idx[0], e, &ents[e]);
  std::vector<LociConstraint> constraints;
  if (!o.constraints_file.empty() && load_constraints(o.constraints_file, ents, constraints) < 0) return 1;
  PriorityRegions priority;
  if (!o.priority_file.empty() && load_priority_regions(o.priority_file, ents, priority) < 0) return 1;

  // -Z / -z as AcceptThisChromID evaluates them (Aligner.cpp:2651-2710: exclude expressions first, then -- if any -- the
  // include expressions).  Two callers in the reference: the pairing of paired-end runs (AcceptProvPE, the orphan-recovery
  // arms and the SE fallback), which the library's pairing kernels reproduce from this map, and WriteHitLoci under -r5.
  // FiltByChroms (include expressions first) still runs over the records afterwards, in every mode.
  std::vector<uint8_t> accept_keep;
  if (!rin.empty() || !rex.empty()) {
    accept_keep.assign(info.num_entries + 1, 1);
    for (uint32_t e = 1; e <= info.num_entries; ++e) {
      regmatch_t mc;
      bool ok = true;
      for (auto& re : rex) if (!regexec(&re, ents[e].name, 1, &mc, 0)) { ok = false; break; }
      if (ok && !rin.empty()) {
        ok = false;
        for (auto& re : rin) if (!regexec(&re, ents[e].name, 1, &mc, 0)) { ok = true; break; }
      }
      accept_keep[e] = ok;
    }
    if (o.pe_mode)
      for (int g = 0; g < o.gpus; ++g)
        if (bkx_set_chrom_filter(idx[(size_t)g], accept_keep.data(), (uint32_t)accept_keep.size()) < 0) { diag("Fatal: %s", bkx_last_error()); return 1; }
  }

  reads_thread.join();
  if (reads_rc < 0) return 1;
  const uint32_t n = R.n();
  // no read survived the load filters: the reference goes on and reports an empty run (Aligner.cpp:486-535 then print an
  // average of 0 with a minimum of -1, and its class summary counts one unprocessed record)
  diag("Genome assembly suffix array loaded");
  diag("Now aligning with minimum core size of %dbp...\n", P.min_core_len);

  // ---- align: contiguous, even-sized read ranges, one host thread per GPU (reads shard with no exchange)
  const bool all_loci = o.ml_mode == BKX_ML_ALL;   // -r5: every locus of a read becomes a record of its own
  const bool clustered = o.ml_mode == BKX_ML_UNIQ || o.ml_mode == BKX_ML_MULTI;   // -r3 / -r4: one locus by clustering
  const bool compact = !(all_loci || clustered);   // the compact host interface: 2 bits per base in, 16-byte records out
  ResVec res(n);            // filled below (expanded from the 16-byte records, or written by the multi-loci call)
  // what crosses PCIe is page-locked: the 2-bit read stream and the 16-byte records from the moment they are allocated (by the
  // reads thread, while the index streams in); the multi-loci calls register their one-byte-per-base arena here
  PinnedBuf<bkx_read_result16>& res16 = R.res16;
  bool pinned = true;
  if (compact) pinned = R.packed2.pinned && res16.pinned;
  else pinned = bkx_pin_host(R.bases.data(), R.bases.size()) >= 0 && bkx_pin_host(res.data(), res.size() * sizeof(bkx_read_result)) >= 0;
  if (!pinned) diag("Note: unable to page-lock host buffers (%s); continuing with pageable copies", bkx_last_error());
  std::vector<bkx_multi_hit> multi;
  if (all_loci || clustered) multi.resize((size_t)n * (size_t)o.max_ml);
  // paired ends: aligned and paired in one pass over the data (each GPU pairs its own contiguous, even-sized range)
  bkx_pe_params PE;
  memset(&PE, 0, sizeof(PE));
  PE.pe_proc = o.pe_mode; PE.pair_min_len = o.pair_min; PE.pair_max_len = o.pair_max; PE.pair_strand = o.pair_strand;
  PE.circularised = o.pe_circ;
  // -6: search, pairing and the recovery's acceptance run at -s plus -6, the recovery's core lengths still come from -s
  // (m_MaxSubs at Aligner.cpp:3256 against pPars->MaxSubs at :3275)
  PE.rescue_core_subs_p1 = o.pcr_primer > 0 ? o.max_subs + 1 : 0;
  // -O with paired ends: the insert-size histogram m_pLenDist[0..100000] (Aligner.cpp:2908-2915), one per GPU, summed below
  const size_t kLenDist = 100001;
  std::vector<std::vector<uint32_t>> len_dist((size_t)o.gpus);
  if (o.pe_mode && !o.stats_file.empty()) for (auto& v : len_dist) v.assign(kLenDist, 0);
  std::vector<bkx_pe_stats> pst((size_t)o.gpus);
  for (auto& q : pst) memset(&q, 0, sizeof(q));
  std::vector<bkx_align_stats> st((size_t)o.gpus);
  std::vector<int> rcs((size_t)o.gpus, 0);
  std::vector<std::string> errs((size_t)o.gpus);
  {
    std::vector<std::thread> th;
    uint32_t per = ((n + o.gpus - 1) / o.gpus + 1) & ~1u;
    for (int g = 0; g < o.gpus; ++g) {
      uint32_t b = std::min<uint64_t>(n, (uint64_t)per * g), e = std::min<uint64_t>(n, (uint64_t)per * (g + 1));
      memset(&st[(size_t)g], 0, sizeof(bkx_align_stats));
      th.emplace_back([&, g, b, e]() {
        if (e > b) {
          // this GPU's share of the stream: its first base and its slice of the (ascending) exception list
          const uint64_t o_b = R.offs[b], o_e = R.offs[e];
          const size_t x0 = (size_t)(std::lower_bound(R.exc_pos.begin(), R.exc_pos.end(), o_b) - R.exc_pos.begin());
          const size_t x1 = (size_t)(std::lower_bound(R.exc_pos.begin(), R.exc_pos.end(), o_e) - R.exc_pos.begin());
          const uint16_t* lens = R.fixed_len ? nullptr : R.lens.data() + b;
          if (o.pe_mode && o.priority_file.empty())
            rcs[(size_t)g] = bkx_align_pairs_packed2(idx[(size_t)g], &P, &PE, R.packed2.data(), o_b, lens, R.fixed_len,
                                                     R.exc_pos.data() + x0, R.exc_code.data() + x0, x1 - x0, (e - b) / 2,
                                                     res16.data() + b, &st[(size_t)g], &pst[(size_t)g],
                                                     len_dist[(size_t)g].empty() ? nullptr : len_dist[(size_t)g].data());
          else if (!compact)
            rcs[(size_t)g] = bkx_align_reads_multi(idx[(size_t)g], &P, R.bases.data(), R.offs.data() + b, e - b, res.data() + b,
                                                   multi.data() + (size_t)b * (size_t)o.max_ml, &st[(size_t)g]);
          else
            rcs[(size_t)g] = bkx_align_reads_packed2(idx[(size_t)g], &P, R.packed2.data(), o_b, lens, R.fixed_len,
                                                     R.exc_pos.data() + x0, R.exc_code.data() + x0, x1 - x0, e - b,
                                                     res16.data() + b, &st[(size_t)g]);
          if (rcs[(size_t)g] < 0) errs[(size_t)g] = bkx_last_error();
        }
      });
    }
    for (auto& t : th) t.join();
  }
  for (int g = 0; g < o.gpus; ++g)
    if (rcs[(size_t)g] < 0) { diag("Fatal: %s", errs[(size_t)g].c_str()); return 1; }
  bkx_align_stats S = st[0];  // the only reduction of the path: element-wise sum of the counters
  for (int g = 1; g < o.gpus; ++g) {
    uint64_t* d = (uint64_t*)&S;
    const uint64_t* s = (const uint64_t*)&st[(size_t)g];
    for (size_t k = 0; k < sizeof(S) / 8; ++k) d[k] += s[k];
  }
  if (!compact) { bkx_unpin_host(R.bases.data()); bkx_unpin_host(res.data()); }
  if (compact) {   // 16-byte records -> the 32-byte form the passes and writers work on, by all host threads
    const unsigned T = o.threads > 0 ? (unsigned)o.threads : std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    std::vector<std::thread> th;
    for (unsigned t = 0; t < T; ++t)
      th.emplace_back([&, t]() {
        const uint32_t b = (uint32_t)((uint64_t)n * t / T), e = (uint32_t)((uint64_t)n * (t + 1) / T);
        if (e > b) bkx_expand_results16(res16.data() + b, e - b, R.fixed_len ? nullptr : R.lens.data() + b, R.fixed_len, res.data() + b);
      });
    for (auto& x : th) x.join();
    res16.release();
    R.packed2.release();
  }
  // ---- -B: ProcCoredApprox first searches every read with room for cPriorityExacts = 10 more loci (Aligner.cpp:9102-9145); when
  //      that search ends eHRhits and, of its loci, no more than -R (here: one) lie inside a priority region, the read is
  //      accepted at that locus as if it were unique (:9146-9186) -- otherwise the ordinary search decides.  The first
  //      search is the library's all-loci call with 1 + 10 slots per read; the ordinary search above already ran for
  //      every read, so its record is simply replaced, and the provisional totals move with it.
  if (!o.priority_file.empty()) {
    const int slots = o.max_ml + 10;   // m_MaxMLmatches + cPriorityExacts
    bkx_align_params PP = P;
    PP.ml_mode = BKX_ML_ALL; PP.max_ml_matches = slots; PP.clamp_max_ml = 0; PP.best_matches = 0;
    ResVec pre(n);
    std::vector<bkx_multi_hit> phits((size_t)n * (size_t)slots);
    std::vector<bkx_align_stats> pst2((size_t)o.gpus);
    std::vector<int> prc((size_t)o.gpus, 0);
    std::vector<std::string> perr((size_t)o.gpus);
    std::vector<std::thread> th;
    const uint32_t per = ((n + o.gpus - 1) / o.gpus + 1) & ~1u;
    for (int g = 0; g < o.gpus; ++g) {
      const uint32_t b = std::min<uint64_t>(n, (uint64_t)per * g), e = std::min<uint64_t>(n, (uint64_t)per * (g + 1));
      th.emplace_back([&, g, b, e]() {
        if (e <= b) return;
        prc[(size_t)g] = bkx_align_reads_multi(idx[(size_t)g], &PP, R.bases.data(), R.offs.data() + b, e - b, pre.data() + b,
                                               phits.data() + (size_t)b * (size_t)slots, &pst2[(size_t)g]);
        if (prc[(size_t)g] < 0) perr[(size_t)g] = bkx_last_error();
      });
    }
    for (auto& t : th) t.join();
    for (int g = 0; g < o.gpus; ++g)
      if (prc[(size_t)g] < 0) { diag("Fatal: %s", perr[(size_t)g].c_str()); return 1; }
    std::vector<bkx_multi_hit> h((size_t)slots);
    for (uint32_t i = 0; i < n; ++i) {
      const bkx_read_result& q = pre[i];
      if (q.hit_rslt != BKX_HR_HITS) continue;
      const int nh = std::min<int>(q.low_hit_instances, slots);
      for (int k = 0; k < nh; ++k) h[(size_t)k] = phits[(size_t)i * (size_t)slots + (size_t)k];
      // the reference's in-place compaction: its cursor only advances when it copies, and it copies only once a locus
      // outside the regions has been seen (Aligner.cpp:9152-9176)
      int in_pri = 0, not_pri = 0, cursor = 0;
      for (int k = 0; k < nh; ++k) {
        const bkx_multi_hit m = h[(size_t)k];
        if (!priority.in_any(m.chrom_id, m.match_loci, m.match_loci + m.match_len - 1)) { ++not_pri; continue; }
        if (not_pri > 0) h[(size_t)cursor++] = m;
        ++in_pri;
      }
      if (in_pri == 0 || !(o.clamp_ml || in_pri <= o.max_ml)) continue;
      // eHRhits with LowHitInstances = the loci inside the regions (at most -R of them), the compacted list as the hit list:
      // from here on the record is what ProcCoredApprox makes of any such search (Aligner.cpp:9328-9420)
      const int k = std::min(in_pri, o.max_ml);
      const bkx_multi_hit& m = h[0];
      bkx_read_result r = q;
      r.hit_rslt = BKX_HR_HITS; r.flags = 0;
      if (k == 1 || o.ml_mode == BKX_ML_DEFAULT || o.ml_mode == BKX_ML_ALL) {
        r.nar = BKX_NAR_ACCEPTED;
        r.num_hits = o.ml_mode == BKX_ML_ALL ? (uint8_t)std::min(k, 255) : 1;
        r.low_hit_instances = o.ml_mode == BKX_ML_ALL ? (int16_t)k : 1;
        r.chrom_id = m.chrom_id; r.match_loci = m.match_loci; r.match_len = m.match_len; r.strand = m.strand;
        r.mismatches = m.mismatches;
      } else {   // -r1 / -r3 / -r4 with several loci: counted, not placed
        r.nar = BKX_NAR_MULTIALIGN;
        r.num_hits = 0; r.low_hit_instances = (int16_t)k;
        r.chrom_id = 0; r.match_loci = 0; r.match_len = 0; r.strand = 0; r.mismatches = 0;
      }
      if (!multi.empty())   // -r3 / -r4 / -r5: the loci the clustering or the writers go on with
        for (int j = 0; j < o.max_ml; ++j) {
          bkx_multi_hit z;
          memset(&z, 0, sizeof(z));
          multi[(size_t)i * (size_t)o.max_ml + (size_t)j] = j < k ? h[(size_t)j] : z;
        }
      // the provisional totals move with the record (NumAcceptedAsAligned, NumLociAligned, unique / multi, and what counts
      // as not aligned: no hit, or too many)
      auto count = [&](const bkx_read_result& x, int sign) {
        const bool placed = x.nar == BKX_NAR_ACCEPTED, counted = x.nar == BKX_NAR_MULTIALIGN && x.hit_rslt == BKX_HR_HITS;
        if (placed || counted) {
          const int loci = (placed && !(o.ml_mode == BKX_ML_ALL && x.low_hit_instances > 1)) ? 1 : x.low_hit_instances;
          S.tot_accepted_aligned += (uint64_t)sign;
          S.tot_loci_aligned += (uint64_t)(sign * loci);
          if (loci == 1 && placed) S.tot_accepted_unique += (uint64_t)sign; else S.tot_accepted_multi += (uint64_t)sign;
        }
        if (x.nar == BKX_NAR_NOHIT || (x.nar == BKX_NAR_MULTIALIGN && x.hit_rslt != BKX_HR_HITS)) S.tot_non_aligned += (uint64_t)sign;
      };
      count(res[i], -1);
      count(r, +1);
      res[i] = r;
    }
  }
  // paired ends behind -B: the reads were aligned one by one above (the priority regions decide between the two searches
  // before ProcessPairedEnds sees a record), now they are paired -- the same kernels as the fused call, as a call of their own
  if (o.pe_mode && !o.priority_file.empty()) {
    std::vector<std::thread> th;
    const uint32_t per = ((n + o.gpus - 1) / o.gpus + 1) & ~1u;
    for (int g = 0; g < o.gpus; ++g) {
      const uint32_t b = std::min<uint64_t>(n, (uint64_t)per * g), e = std::min<uint64_t>(n, (uint64_t)per * (g + 1));
      th.emplace_back([&, g, b, e]() {
        if (e <= b) return;
        rcs[(size_t)g] = bkx_pair_reads(idx[(size_t)g], &P, &PE, res.data() + b, (e - b) / 2, R.bases.data(), R.offs.data() + b,
                                        &pst[(size_t)g], len_dist[(size_t)g].empty() ? nullptr : len_dist[(size_t)g].data());
        if (rcs[(size_t)g] < 0) errs[(size_t)g] = bkx_last_error();
      });
    }
    for (auto& t : th) t.join();
    for (int g = 0; g < o.gpus; ++g)
      if (rcs[(size_t)g] < 0) { diag("Fatal: %s", errs[(size_t)g].c_str()); return 1; }
  }
  diag("Alignment of %u from %u loaded completed", n, n);

  // -O: m_MultiHitDist (Aligner.cpp:9364, 9521) -- reads whose search ended eHRhits, by their number of equally good loci;
  //     taken now, before clustering / filters rewrite the records (not counted under -r5, :9336-9352)
  std::vector<uint32_t> multi_dist((size_t)std::max(1, o.max_ml), 0);
  if (!o.stats_file.empty() && o.ml_mode != BKX_ML_DEFAULT && o.ml_mode != BKX_ML_ALL)
    for (uint32_t i = 0; i < n; ++i)
      if (res[i].hit_rslt == BKX_HR_HITS && res[i].low_hit_instances >= 1 && res[i].low_hit_instances <= o.max_ml)
        ++multi_dist[(size_t)res[i].low_hit_instances - 1];

  // -r5 with -Z / -z: the loci of a read are filtered as they are recorded (WriteHitLoci -> AcceptThisChromID, Aligner.cpp:6737-6757,
  // 2651-2710: exclude expressions first, then -- if any -- the include expressions), so the provisional totals count the
  // loci that stay; reads left without a locus leave no record.
  std::vector<uint8_t> r5_keep;
  if (all_loci && !accept_keep.empty()) {
    r5_keep = accept_keep;
    S.tot_accepted_aligned = S.tot_accepted_unique = S.tot_accepted_multi = S.tot_loci_aligned = 0;
    for (uint32_t i = 0; i < n; ++i) {
      if (res[i].nar != BKX_NAR_ACCEPTED) continue;
      int kept = 0;
      for (int h = 0; h < res[i].low_hit_instances; ++h) kept += r5_keep[multi[(size_t)i * (size_t)o.max_ml + (size_t)h].chrom_id];
      if (!kept) continue;
      ++S.tot_accepted_aligned;
      S.tot_loci_aligned += (uint64_t)kept;
      ++(kept == 1 ? S.tot_accepted_unique : S.tot_accepted_multi);
    }
  }

  // ---- read-length summary, Aligner.cpp:486-535
  uint64_t tot_len = R.bases.size();
  int minl = n ? R.len(0) : -1, maxl = n ? R.len(0) : 0;
  for (uint32_t i = 1; i < n; ++i) { minl = std::min(minl, R.len(i)); maxl = std::max(maxl, R.len(i)); }
  int avl = n ? (int)(tot_len / n) : 0;
  diag("Average length of all reads was: %d (min: %d, max: %d)", avl, minl, maxl);
  if (align_subs != 0)
    diag("Typical allowed aligner induced substitutions was: %d (min: %d, max: %d)", std::max(1, avl * align_subs / 100),
         std::max(1, minl * align_subs / 100), std::max(1, maxl * align_subs / 100));
  diag("Provisionally accepted %d aligned reads (%d uniquely, %d aligning to multiloci) aligning to a total of %d loci",
       (int)S.tot_accepted_aligned, (int)S.tot_accepted_unique, (int)S.tot_accepted_multi, (int)S.tot_loci_aligned);

  // ---- -r5: the record array is replaced by one record per reported locus, numbered in read order (the reference
  //      numbers them in arrival order, which is the read order with one thread; WriteHitLoci / AddMultiHit,
  //      Aligner.cpp:6668-6790, and the swap at :538-560).  Reads without a hit only stay for -M6; reads rejected for
  //      their Hamming margin or for Ns leave no record at all.
  std::vector<uint32_t> src;   // record -> read (empty: identity)
  if (all_loci) {
    diag("Treating accepted %d multialigned reads as uniquely aligned %d source reads in subsequent processing",
         (int)S.tot_accepted_multi, (int)(S.tot_loci_aligned - S.tot_accepted_unique));
    ResVec rec;
    rec.reserve((size_t)S.tot_loci_aligned + 16);
    std::vector<int> kept((size_t)std::max(1, o.max_ml));
    for (uint32_t i = 0; i < n; ++i) {
      const bkx_read_result& r = res[i];
      if (r.nar == BKX_NAR_ACCEPTED) {
        // loci that pass the chromosome filter.  The reference compacts the hit list in place with a cursor that only
        // advances when it copies (WriteHitLoci, Aligner.cpp:6744-6754: `if (pAcceptHit != pHit) *pAcceptHit++ = *pHit`):
        // a kept FIRST locus leaves the cursor behind, the following kept loci overwrite it, and the tail of the reported
        // list (its length is the number of kept loci) is whatever the slots held before -- a repeated or even a filtered
        // locus, which FiltByChroms then removes as eNARChromFilt.  Reproduced by running the same compaction.
        int nk = 0;
        const int nh = std::min<int>(r.low_hit_instances, o.max_ml);
        for (int h = 0; h < nh; ++h) kept[h] = h;
        if (r5_keep.empty()) nk = nh;
        else
          for (int h = 0, cursor = 0; h < nh; ++h)
            if (r5_keep[multi[(size_t)i * (size_t)o.max_ml + (size_t)h].chrom_id]) {
              ++nk;
              if (cursor != h) kept[cursor++] = h;
            }
        for (int k = 0; k < nk; ++k) {
          const bkx_multi_hit& m = multi[(size_t)i * (size_t)o.max_ml + (size_t)kept[k]];
          bkx_read_result q = r;
          q.num_hits = 1;
          q.chrom_id = m.chrom_id; q.match_loci = m.match_loci; q.match_len = m.match_len; q.strand = m.strand;
          q.mismatches = m.mismatches;
          rec.push_back(q);
          src.push_back(i);
        }
      } else if (o.fmt == 6 && (r.hit_rslt == BKX_HR_NONE || r.hit_rslt == BKX_HR_HITINSTS) && r.nar != BKX_NAR_NS) {
        bkx_read_result q = r;
        q.num_hits = 0;
        rec.push_back(q);
        src.push_back(i);
      }
    }
    res.swap(rec);
    std::vector<bkx_multi_hit>().swap(multi);
  }
  const uint32_t nrec = (uint32_t)res.size();

  // ---- -r3 / -r4: AssignMultiMatches, Aligner.cpp:583-592, 5108-5270
  if (clustered) {
    diag("Multialignment processing started..");
    uint32_t longest = 0;
    for (uint32_t i = 0; i < n; ++i) longest = std::max<uint32_t>(longest, (uint32_t)R.len(i));
    bkx_cluster_stats cs;
    diag("Assigning %d reads which aligned to multiple loci to a single loci", (int)S.tot_accepted_multi);
    diag("Sorting...");
    diag("Sorting completed, now clustering...");
    diag("Assigning...");
    memset(&cs, 0, sizeof(cs));
    if (n && bkx_assign_multi_matches(res.data(), n, multi.data(), o.max_ml, o.ml_mode, longest, &cs) < 0) {
      diag("Fatal: %s", bkx_last_error());
      return 1;
    }
    diag("Checking for orphans (unclustered) from %d putative assignments..", (int)cs.putative);
    diag("Clustering completed, removed %d unclustered orphans from %d putative resulting in %d (%d clustered near unique, %d clustered near other multiloci reads) multihit reads accepted as assigned",
         (int)(cs.putative - cs.assigned), (int)cs.putative, (int)cs.assigned, (int)cs.near_unique, (int)cs.near_multi);
    diag("Multialignment processing completed");
    std::vector<bkx_multi_hit>().swap(multi);
  }

  // ---- paired ends, Aligner.cpp:2876-3049
  if (o.pe_mode) {
    diag("Paired end association and partner alignment processing started..");
    diag("Generating paired reads index over %d paired reads", n / 2);
    diag("Starting to associate Paired End reads to be within insert size range ...");
    diag("Processed putative 0 pairs, accepted 0");
    bkx_pe_stats ps = pst[0];
    for (int g = 1; g < o.gpus; ++g) {
      uint64_t* d = (uint64_t*)&ps;
      const uint64_t* q = (const uint64_t*)&pst[(size_t)g];
      for (size_t k = 0; k < sizeof(ps) / 8; ++k) d[k] += q[k];
    }
    diag("Completed association of Paired End reads from %u pairs, accepted %u pairs", n / 2, (unsigned)ps.accepted_num_paired);
    diag("From %d Paired End pairs there were %d accepted (of which %d pairs were from recovered orphans)", (int)(n / 2),
         (int)ps.accepted_num_paired, (int)ps.partner_paired);
    // the reference adds PartnerUnpaired twice when it joins its threads (Aligner.cpp:2990,2997)
    diag("%d Paired End pairs unrecoverable as still orphan partnered", (int)(2 * ps.partner_unpaired - ps.partner_paired));
    diag("%d Paired End pairs were under length, %d over length, not accepted as being paired", (int)ps.under_len_pairs, (int)ps.over_len_pairs);
    diag("%d Paired End aligned pairs were filtered out by chromosome", (int)ps.num_filtered_by_chrom);
    diag("%d Paired End pairs have neither end uniquely aligned", (int)ps.unaligned_pairs);
    if (o.pe_mode == BKX_PE_UNIQUE_SE || o.pe_mode == BKX_PE_ORPHAN_SE)
      diag("%d Paired End reads were unable to be associated with partner read and accepted as if SE aligned", (int)ps.accepted_num_se);
    if (stats_fp) {   // Aligner.cpp:3024-3041: "<insert length>,<pairs>" for 0..cPairMaxLen
      std::vector<uint32_t> tot(kLenDist, 0);
      for (auto& v : len_dist) for (size_t k = 0; k < v.size(); ++k) tot[k] += v[k];
      std::string s;
      for (size_t k = 0; k < kLenDist; ++k) { append_uint(s, k); s += ','; append_uint(s, tot[k]); s += '\n'; }
      fwrite(s.data(), 1, s.size(), stats_fp);
    }
    diag("Paired end association and partner alignment processing completed..");
  }

  unsigned fmt_threads = o.threads > 0 ? (unsigned)o.threads : std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  // host copy of the chromosomes (1 byte/base) for the passes and writers that compare with / print the target sequence
  std::vector<std::vector<uint8_t>> genome(info.num_entries + 1);
  bool need_genome = o.fmt == 1 || o.fmt == 3 || o.min_flank > 0 || o.pcr_primer > 0 || stats_fp;
  for (auto& c : constraints) need_genome = need_genome || (c.mask & 0x10);
  if (need_genome)
    for (uint32_t e = 1; e <= info.num_entries; ++e) {
      genome[e].resize(ents[e].seq_len);
      if (bkx_get_seq(idx[0], e, 0, ents[e].seq_len, genome[e].data()) < 0) { diag("Fatal: %s", bkx_last_error()); return 1; }
    }

  Records rc{o, R, res, src, ents, genome, fmt_threads, {}, {}, {}};
  // the passes, in the order of CAligner::Align (Aligner.cpp:596-655)
  if (!constraints.empty()) identify_constraint_violations(rc, constraints);
  if (o.pcr_win >= 0 && !o.pe_mode && reduce_pcr_duplicates(rc) < 0) return 1;
  if (o.pcr_primer > 0) {
    diag("PCR 5' Primer correction processing started..");
    pcr_5prime_correct(rc);
    diag("PCR 5' Primer correction processing completed");
  }
  if (o.min_flank > 0) auto_trim_flanks(rc);
  if (!o.excl.empty() || !o.incl.empty()) filter_by_chroms(rc, rin, rex);
  if (!o.priority_file.empty() && !o.priority_nofilt) filter_by_priority_regions(rc, priority);
  const auto& trim_l = rc.trim_l;
  const auto& trim_r = rc.trim_r;
  const uint32_t elim_plus = rc.elim_plus, elim_minus = rc.elim_minus;
  uint32_t num_trimmed = 0;

  // "were trimmed" of the summary: FlagTR of the alignments still accepted at this point (Aligner.cpp:3558)
  if (!trim_l.empty())
    for (uint32_t i = 0; i < nrec; ++i) if (res[i].nar == BKX_NAR_ACCEPTED && (trim_l[i] || trim_r[i])) ++num_trimmed;

  // ---- order (SortReadHits(eRSMHitMatch): on the alignment as reported, i.e. after trimming)
  std::vector<uint32_t> order(nrec);
  {
    ResVec keyed;
    if (!trim_l.empty()) {
      keyed = res;
      for (uint32_t i = 0; i < nrec; ++i) { keyed[i].match_loci = rc.adj_start(i); keyed[i].match_len = (uint16_t)rc.adj_len(i); }
    }
    if (nrec && bkx_sort_hits(keyed.empty() ? res.data() : keyed.data(), nrec, order.data(), 0) < 0) { diag("Fatal: %s", bkx_last_error()); return 1; }
  }

  if (!o.none_file.empty()) {
    diag("Reporting of non-aligned reads started..");
    if (!report_reads(rc, order, o.none_file, "na", [](const bkx_read_result& r) { return r.nar == BKX_NAR_NS || r.nar == BKX_NAR_NOHIT; })) return 1;
    diag("Reporting of non-aligned reads completed");
  }
  if (!o.multi_file.empty()) {
    diag("Reporting of multialigned reads started..");
    if (!report_reads(rc, order, o.multi_file, "ml", [](const bkx_read_result& r) { return r.nar == BKX_NAR_MULTIALIGN; })) return 1;
    diag("Reporting of multialigned reads completed");
  }

  // ---- summary, Aligner.cpp:3535-3769
  uint64_t nar[BKX_NAR_COUNT] = {0};
  uint64_t plus = 0;
  for (uint32_t i = 0; i < nrec; ++i) {
    nar[res[i].nar]++;
    if (res[i].nar == BKX_NAR_ACCEPTED && res[i].strand == '+') ++plus;
  }
  if (nrec == 0) nar[BKX_NAR_UNALIGNED] = 1;   // with no record left (-r5, every locus filtered) the reference's walk still classes one empty slot
  uint64_t no_match = nar[BKX_NAR_NOHIT] + S.num_sloughed_ns;
  if (all_loci) {  // Aligner.cpp:3706-3708, 3727-3730: the non-aligned total of the search stands in for the record count
    no_match = S.tot_non_aligned + S.num_sloughed_ns;
    nar[BKX_NAR_NOHIT] = no_match;
  }
  diag("From %u source reads there are %u accepted alignments, %u on '+' strand, %u on '-' strand", n, (unsigned)nar[BKX_NAR_ACCEPTED],
       (unsigned)plus, (unsigned)(nar[BKX_NAR_ACCEPTED] - plus));
  const SimTruth truth = sim_truth_check(rc);
  if (truth.sim)
    diag("There are %u (%u 2 edge, %u 1 edge) high confidence aligned simulated reads with %u misaligned", truth.edge2 + truth.edge1,
         truth.edge2, truth.edge1, truth.misaligned);
  diag("A further %u multiloci aligned reads could not accepted as hits because they were unresolvable", (unsigned)nar[BKX_NAR_MULTIALIGN]);
  diag("A further %u aligned reads were not accepted as hits because of insufficient Hamming edit distance", (unsigned)nar[BKX_NAR_MMDELTA]);
  diag("A further %u '+' and %u '-' strand aligned reads not accepted because of flank trimming (%d were trimmed) requirements", (unsigned)elim_plus, (unsigned)elim_minus, (int)num_trimmed);
  diag("Unable to align %u source reads of which %d were not aligned as they contained excessive number of indeterminate 'N' bases",
       (unsigned)no_match, (int)S.num_sloughed_ns);
  diag("Read nonalignment reason summary:");
  for (int k = 0; k < BKX_NAR_COUNT; ++k) {
    uint64_t v = nar[k];
    if (k == BKX_NAR_NS) v = S.num_sloughed_ns;  // the reference substitutes m_NumSloughedNs (Aligner.cpp:3724)
    diag("   %u (%s) %s", (unsigned)v, kNarCode[k], kNarText[k]);
  }

  // ---- write
  OutBuf ob;
  if (!ob.open(o.out)) { diag("Fatal: unable to create '%s'", o.out.c_str()); return 1; }
  diag("Reporting of aligned result set started...");
  if (o.fmt == 4) write_bed(rc, order, info, ob);
  else if (o.fmt <= 3) write_csv(rc, order, info, ob);
  else if (is_bam_name(o.out)) { if (!write_bam(rc, order, info, ob)) return 1; }
  else write_sam(rc, order, info, ob);
  ob.close();
  diag("Reporting of aligned result set completed");

  if (stats_fp) {
    write_statistics(rc, multi_dist, stats_fp);
    fclose(stats_fp);
  }
  for (auto* x : idx) bkx_close_index(x);
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
  int hh = (int)(secs / 3600), mm = (int)(secs / 60) % 60;
  diag("Exit code: 0 Total processing time: %3.2d:%2.2d:%06.3f seconds", hh, mm, secs - 3600.0 * hh - 60.0 * mm);
  if (g_log) fclose(g_log);
  return 0;
}
