/*
 * bkx.h -- C ABI of the B200-native `biokanga align` hot path.
 *
 * The reference has no plugin/FFI layer: CAligner (biokanga/Aligner.cpp) calls the C++ class
 * CSfxArrayV3 (libbiokanga/SfxArrayV2.h:298-990) one read at a time from <=128 pthreads.  A per-read
 * synchronous call is unusable for a GPU, so the boundary sits one level up: the host aligner hands
 * BATCHES of reads to this library and gets back one fixed 32-byte record per read that carries
 * exactly the fields CAligner::ProcCoredApprox stores into tsReadHit (Aligner.cpp:9311-9479).
 * Everything here is plain C: pointers, sizes, PODs; no C++/torch types.  All functions return
 * >= 0 on success and a negative teBSFrsltCodes-style code on failure (libbiokanga/ErrorCodes.h:15-96);
 * bkx_last_error() returns the text of the last failure on the calling thread.
 *
 * Each entry point names the reference interface it replaces (file:line under /root/reference).
 */
#ifndef BKX_H
#define BKX_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BKX_ABI_VERSION 1

/* ---- result codes (sign convention of libbiokanga/ErrorCodes.h) ------------------------------ */
enum {
  BKX_OK = 0,
  BKX_ERR_PARAM = -1,     /* eBSFerrParams   */
  BKX_ERR_FILE = -2,      /* eBSFerrOpnFile / eBSFerrFileAccess */
  BKX_ERR_FORMAT = -3,    /* eBSFerrNotBioseq / eBSFerrFileVer  */
  BKX_ERR_MEM = -4,       /* eBSFerrMem      */
  BKX_ERR_CUDA = -5,      /* any CUDA runtime failure (no CPU fallback exists) */
  BKX_ERR_UNSUPPORTED = -6,
  BKX_ERR_ENTRY = -7      /* eBSFerrEntry    */
};

/* ---- base codes: libbiokanga/commdefs.h:108-123 (etSeqBase) ---------------------------------- */
enum { BKX_BASE_A = 0, BKX_BASE_C = 1, BKX_BASE_G = 2, BKX_BASE_T = 3, BKX_BASE_N = 4,
       BKX_BASE_UNDEF = 5, BKX_BASE_INDEL = 6, BKX_BASE_EOS = 7 };

/* ---- tHRslt: libbiokanga/SfxArrayV2.h:68-74 -------------------------------------------------- */
enum { BKX_HR_NONE = 0, BKX_HR_HITS = 1, BKX_HR_MMDELTA = 2, BKX_HR_HITINSTS = 3, BKX_HR_RMMDELTA = 4 };

/* ---- teNAR: biokanga/Aligner.h:106-128 -------------------------------------------------------- */
enum {
  BKX_NAR_UNALIGNED = 0, BKX_NAR_ACCEPTED, BKX_NAR_NS, BKX_NAR_NOHIT, BKX_NAR_MMDELTA, BKX_NAR_MULTIALIGN,
  BKX_NAR_TRIM, BKX_NAR_SPLICEJCTN, BKX_NAR_MICROINDEL, BKX_NAR_PCRDUP, BKX_NAR_NONUNIQUE, BKX_NAR_CHROMFILT,
  BKX_NAR_REGIONFILT, BKX_NAR_PEINSERTMIN, BKX_NAR_PEINSERTMAX, BKX_NAR_PENOHIT, BKX_NAR_PESTRAND,
  BKX_NAR_PECHROM, BKX_NAR_PEUNALIGN, BKX_NAR_LOCICONSTRAINED, BKX_NAR_COUNT
};

/* ---- eALStrand: libbiokanga/SfxArrayV2.h:61-66 ------------------------------------------------ */
enum { BKX_STRAND_BOTH = 0, BKX_STRAND_WATSON = 1, BKX_STRAND_CRICK = 2 };

/* ---- etPMode: biokanga/Aligner.h:214-220 ------------------------------------------------------ */
enum { BKX_PMODE_DEFAULT = 0, BKX_PMODE_MORESENS = 1, BKX_PMODE_ULTRASENS = 2, BKX_PMODE_LESSSENS = 3 };

/* ---- etMLMode: biokanga/Aligner.h:224-231.  Built: DEFAULT (max_ml_matches must be 1), DIST ("-r1", stats only:
 * reads with 2..max_ml_matches equally good loci are reported eNARMultiAlign with their exact LowHitInstances,
 * Aligner.cpp:9328-9400), ALL ("-r5": every one of up to max_ml_matches <= 500 equally good loci is returned, through
 * bkx_align_reads_multi) and UNIQ / MULTI ("-r3" / "-r4": the same call returns the loci, records of multi-loci reads
 * stay eNARMultiAlign until bkx_assign_multi_matches picks a locus by clustering).  RAND uses libc rand() in the
 * reference and cannot be reproduced -- rejected with BKX_ERR_UNSUPPORTED. */
enum { BKX_ML_DEFAULT = 0, BKX_ML_DIST = 1, BKX_ML_RAND = 2, BKX_ML_UNIQ = 3, BKX_ML_MULTI = 4, BKX_ML_ALL = 5 };

/* ---- etPEproc: biokanga/Aligner.h:252-259 ----------------------------------------------------- */
enum { BKX_PE_NONE = 0, BKX_PE_ORPHAN = 1, BKX_PE_UNIQUE = 2, BKX_PE_ORPHAN_SE = 3, BKX_PE_UNIQUE_SE = 4 };

/* One chromosome / contig of the index: tsSfxEntry, libbiokanga/SfxArrayV2.h:79-88. */
typedef struct bkx_entry {
  uint32_t entry_id;     /* 1..n */
  uint32_t seq_len;      /* excludes the EOS terminator */
  uint64_t start_ofs;    /* offset of first base in the concatenated sequence */
  uint64_t end_ofs;      /* offset of last base (inclusive) */
  char name[88];         /* szSeqName[81], NUL terminated; sizeof(bkx_entry) == 112 */
} bkx_entry;

/* tsSfxHeaderV3 + tsSfxBlock summary: libbiokanga/SfxArrayV2.h:98-104,174-187. */
typedef struct bkx_index_info {
  uint64_t concat_len;     /* tsSfxBlock::ConcatSeqLen (bases + one EOS per entry) */
  uint64_t tot_seq_len;    /* CSfxArrayV3::GetTotSeqsLen(), SfxArrayV2.cpp:2070 */
  uint32_t num_entries;
  uint32_t sfx_el_size;    /* 4 or 5 */
  uint32_t version;        /* 3..5 */
  uint32_t attributes;     /* bit0 bisulfite, bit1 colorspace (both rejected) */
  uint32_t prefix_k;       /* k of the device k-mer prefix table */
  uint32_t device;         /* CUDA ordinal the index lives on */
  uint64_t device_bytes;   /* HBM bytes held by this index */
  char dataset_name[84];
} bkx_index_info;

/* Everything CAligner feeds into CSfxArrayV3::AlignReads per run (tsThreadMatchPars, Aligner.h, and
 * the per-read derivations at Aligner.cpp:9041-9095).  bkx_default_params() fills it the way
 * kanga.cpp:322-1082 + CAligner::LocateCoredApprox (Aligner.cpp:8727-8761) do. */
typedef struct bkx_align_params {
  int32_t pmode;            /* -m  BKX_PMODE_*                                   */
  int32_t max_subs;         /* -s  substitutions per 100 bp (default 10, max 15) */
  int32_t min_edit_dist;    /* -e  MMDelta 1..2                                  */
  int32_t max_ns;           /* -n  max N per read / per 100 bp (default 1)       */
  int32_t align_strand;     /* -Q  BKX_STRAND_*                                  */
  int32_t max_ml_matches;   /* MaxHits handed to AlignReads (1 unless -r/-R)     */
  int32_t min_core_len;     /* m_MinCoreLen after the genome-size + mode rule    */
  int32_t max_num_slides;   /* per 100 bp: 8 default, 9 ultra, 6 less sensitive  */
  int32_t max_iter;         /* CSfxArrayV3::m_MaxIter: 5000/10000/20000/2500     */
  int32_t max_ident_nodes;  /* cMaxNumIdentNodes = 1 024 000 (SfxArrayV2.h:15)   */
  int32_t ml_mode;          /* -r  BKX_ML_*: what to do with reads hitting several loci (default slough)  */
  int32_t clamp_max_ml;     /* -X  treat reads with more than max_ml_matches loci as if exactly that many */
  int32_t best_matches;     /* -N  with a multi-loci mode: the max_ml_matches loci with the fewest mismatches (<= the -s limit),
                             *     found in ONE un-staged pass -- CSfxArrayV3::LocateBestMatches (SfxArrayV2.cpp:6654-7019)
                             *     instead of AlignReads; implies clamp_max_ml (kanga.cpp:695-696) */
  int32_t reserved[3];
} bkx_align_params;

/* Fixed 32-byte per-read record: the tsReadHit fields written by ProcCoredApprox
 * (Aligner.cpp:9311-9479) plus two work counters that define the algorithmic-bytes numerator
 * (SURVEY.md section 8(d)): seeds = LocateFirstExact calls the reference issues for this read,
 * cands = candidate loci that reach its Hamming loop. */
typedef struct bkx_read_result {
  uint8_t nar;                /* BKX_NAR_*                                   */
  uint8_t hit_rslt;           /* BKX_HR_* returned by AlignReads              */
  uint8_t strand;             /* '+', '-', '?' or 0                           */
  uint8_t num_hits;           /* tsReadHit::NumHits (-r5: loci returned, saturating at 255 -- low_hit_instances has the count) */
  int8_t low_mm;              /* tsReadHit::LowMMCnt                          */
  int8_t nxt_low_mm;          /* tsReadHit::NxtLowMMCnt                       */
  int16_t low_hit_instances;  /* tsReadHit::LowHitInstances (clamped MaxML+1) */
  uint32_t chrom_id;          /* Seg[0].ChromID = EntryID, 0 if no hit        */
  uint32_t match_loci;        /* Seg[0].MatchLoci (0-based in chromosome)     */
  uint16_t match_len;         /* Seg[0].MatchLen                              */
  uint8_t mismatches;         /* Seg[0].Mismatches                            */
  uint8_t flags;              /* BKX_FLG_*                                    */
  uint32_t seeds;
  uint32_t cands;
  uint32_t reserved;
} bkx_read_result;

enum { BKX_FLG_PE_ALIGNED = 1, BKX_FLG_PE_RECOVERED = 2 };

/* The same record as it crosses PCIe in the compact host interface (bkx_align_reads_packed2 / bkx_align_pairs_packed2):
 * 16 bytes.  Left out: seeds / cands (their sums are in bkx_align_stats) and match_len, which is the read's length
 * whenever the record names a strand and 0 otherwise.  bkx_expand_results16 rebuilds the 32-byte form. */
typedef struct bkx_read_result16 {
  uint8_t nar_hr;             /* BKX_NAR_* in bits 0..4, BKX_HR_* in bits 5..7            */
  uint8_t strand_flags;       /* bits 0..1: 0 none, 1 '+', 2 '-', 3 '?'; bits 2..3: BKX_FLG_* */
  uint8_t num_hits;
  uint8_t mismatches;
  int8_t low_mm;
  int8_t nxt_low_mm;
  int16_t low_hit_instances;
  uint32_t chrom_id;
  uint32_t match_loci;
} bkx_read_result16;

/* One locus of a read under -r5 (BKX_ML_ALL): the tsHitLoci fields CAligner::WriteHitLoci copies into the record it
 * appends per hit (Aligner.cpp:6723-6790).  bkx_align_reads_multi returns max_ml_matches of these per read; the first
 * bkx_read_result::num_hits of a read's slots are valid, in the order LocateCoreMultiples found them. */
typedef struct bkx_multi_hit {
  uint32_t chrom_id;
  uint32_t match_loci;
  uint16_t match_len;
  uint8_t strand;       /* '+' / '-' */
  uint8_t mismatches;
} bkx_multi_hit;

/* Paired-end parameters: CAligner::ProcessPairedEnds arguments, Aligner.cpp:2876-2881. */
typedef struct bkx_pe_params {
  int32_t pe_proc;          /* -U BKX_PE_*                         */
  int32_t pair_min_len;     /* -d (default 100)                    */
  int32_t pair_max_len;     /* -D (default 1000)                   */
  int32_t pair_strand;      /* -E: 1 = both ends on same strand    */
  int32_t circularised;     /* -c PE circularised fragments        */
  int32_t rescue_core_subs_p1; /* 0: orphan recovery derives its core length from the run's max_subs.  s + 1: from the rate s
                                * instead -- a -6 run searches and accepts at -s plus -6 (m_InitalAlignSubs) while the recovery's
                                * cores are still sized from -s (m_MaxSubs), Aligner.cpp:3256 vs :3275 */
  int32_t reserved[2];
} bkx_pe_params;

/* The eight PE counters of tsPEThreadPars (Aligner.cpp:3479-3486).  partner_unpaired is the plain
 * per-pair count; the reference's log line adds it twice (Aligner.cpp:2990,2997) -- the host
 * reporter reproduces that, the library does not. */
typedef struct bkx_pe_stats {
  uint64_t unaligned_pairs, accepted_num_paired, accepted_num_se, partner_paired, partner_unpaired,
      num_filtered_by_chrom, under_len_pairs, over_len_pairs;
} bkx_pe_stats;

/* Per-batch counters merged by the reference under m_hMtxIterReads (Aligner.cpp:9507-9525). */
typedef struct bkx_align_stats {
  uint64_t nar[BKX_NAR_COUNT];
  uint64_t plus_hits, minus_hits;
  uint64_t num_sloughed_ns, tot_non_aligned, tot_accepted_unique, tot_accepted_multi, tot_accepted_aligned,
      tot_loci_aligned, tot_not_accepted_delta;
  uint64_t seeds, cands;        /* sum of the per-read work counters */
  uint64_t reads;
} bkx_align_stats;

typedef struct bkx_index bkx_index; /* opaque: device-resident index + streams + staging */

/* ---- library ---------------------------------------------------------------------------------- */
int bkx_abi_version(void);
const char* bkx_last_error(void);
int bkx_device_count(void);

/* ---- index: replaces CSfxArrayV3::Open + SetTargBlock (SfxArrayV2.cpp:891-1100, 1836-1925) ----
 * Reads a .sfx written by the reference's `biokanga index` (header tsSfxHeaderV3 / tsSfxHeaderVv,
 * entries tsSfxEntry, one tsSfxBlock), uploads it to `device`, 2-bit packs the genome, builds the
 * exception masks and the k-mer prefix table there.  prefix_k = 0 chooses k from the genome size. */
int bkx_open_index(const char* sfx_path, int device, int prefix_k, bkx_index** out);
/* Same, from host memory laid out like tsSfxBlock::SeqSuffix (1 byte/base then SA elements). */
int bkx_open_index_mem(const uint8_t* seq, uint64_t concat_len, const void* sa, uint32_t sfx_el_size,
                       const bkx_entry* entries, uint32_t num_entries, const char* dataset_name, int device,
                       int prefix_k, bkx_index** out);
/* Same, from buffers already resident on `device` (bench / GPU-built suffix arrays).  The buffers must be complete when
 * the call is made (synchronise the stream that produced them): the library works on its own non-blocking streams. */
int bkx_open_index_dev(const uint8_t* d_seq, uint64_t concat_len, const void* d_sa, uint32_t sfx_el_size,
                       const bkx_entry* entries, uint32_t num_entries, const char* dataset_name, int device,
                       int prefix_k, bkx_index** out);
/* Same, from a suffix array split into two device planes (u32 low words and, for 5-byte elements, u8 high bytes; d_sa_hi
 * may be NULL below 4e9 symbols).  Without a high plane the low plane is BORROWED, not copied, and must outlive the index
 * (bkx_close_index leaves it alone); with one, the planes are merged into an array of 5-byte elements owned by the index
 * -- at wheat scale use bkx_build_suffix_array_packed5 + bkx_open_index_packed5, which need no second copy. */
int bkx_open_index_planes(const uint8_t* d_seq, uint64_t concat_len, const uint32_t* d_sa_lo, const uint8_t* d_sa_hi,
                          const bkx_entry* entries, uint32_t num_entries, const char* dataset_name, int device,
                          int prefix_k, bkx_index** out);
/* Same, from 5-byte suffix elements back to back on the device (the element layout of the .sfx file; what
 * bkx_build_suffix_array_packed5 writes): the layout the search reads for indexes of >= 4e9 symbols -- one memory fetch per
 * element.  d_sa5 must start on an 8-byte boundary and be followed by 16 readable bytes; it is BORROWED like the planes. */
int bkx_open_index_packed5(const uint8_t* d_seq, uint64_t concat_len, const uint8_t* d_sa5, const bkx_entry* entries,
                           uint32_t num_entries, const char* dataset_name, int device, int prefix_k, bkx_index** out);
/* Every bkx_open_index* ends with a device self-check: each suffix-array element must lie inside the prefix-table bucket
 * of the suffix it names (ties genome words, table and suffix array together; BKX_NO_VERIFY=1 skips it).  This re-runs
 * it on a live index and returns the number of elements that fail (0 = sound), < 0 on error. */
int64_t bkx_self_check(bkx_index* idx);
/* Diagnostic: reset run-time state of an index to that of a freshly opened one (1 overflow pool, 2 lane hash sets,
 * 4 launch geometry; OR-able).  Results never depend on that state by design; tests/test_gpu_fuzz.py uses this to
 * narrow down the open issue described in DESIGN.md section 4. */
int bkx_debug_reset(bkx_index* idx, int what);
/* Replicate an open index onto another GPU by peer copies (multi-GPU read sharding). */
int bkx_clone_index(const bkx_index* src, int device, bkx_index** out);
void bkx_close_index(bkx_index* idx); /* CSfxArrayV3::Reset / Close */

int bkx_index_info_get(const bkx_index* idx, bkx_index_info* out); /* GetSfxHeader/GetTotSeqsLen/GetNumEntries */
int bkx_get_entry(const bkx_index* idx, uint32_t entry_id, bkx_entry* out); /* GetIdentName/GetSeqLen (SfxArrayV2.h) */
int bkx_get_ident(const bkx_index* idx, const char* name);                  /* GetIdent(name) -> entry id or <0 */
/* GetSeq(EntryID, Loci, buf, Len): 1 byte/base codes, returns number of bases copied. */
int64_t bkx_get_seq(const bkx_index* idx, uint32_t entry_id, uint64_t loci, uint64_t len, uint8_t* buf);

/* ---- parameters ------------------------------------------------------------------------------- */
/* Defaults of kanga.cpp:322-1082 and the genome-size rule of Aligner.cpp:8727-8761 for this index. */
int bkx_default_params(const bkx_index* idx, int pmode, bkx_align_params* out);

/* ---- the hot path: replaces the ProcCoredApprox -> CSfxArrayV3::AlignReads loop ---------------
 * (Aligner.cpp:9024-9505 -> SfxArrayV2.cpp:7666-7760, 5693-6262, 7765-8027)
 * bases: concatenated reads, 1 byte/base, low 3 bits = etSeqBase code (as CAligner::AddEntry packs
 * them, Aligner.cpp:10572-10677; bits 3..7 are ignored).  offsets[n_reads+1]: start of each read.
 * Host variant: plain (pageable or pinned) host pointers; H2D, kernels and D2H run on the index's
 * own streams, double buffered; returns when `out[0..n_reads)` is complete. */
int bkx_align_reads(bkx_index* idx, const bkx_align_params* p, const uint8_t* bases, const uint64_t* offsets,
                    uint32_t n_reads, bkx_read_result* out, bkx_align_stats* stats /* may be NULL; accumulated */);
/* -r5 (params->ml_mode == BKX_ML_ALL): as bkx_align_reads, plus the loci of every read.  multi holds
 * n_reads * params->max_ml_matches entries; read i owns multi[i * max_ml_matches ..], of which out[i].num_hits are
 * valid when out[i].nar == BKX_NAR_ACCEPTED (out[i] itself carries the first one, and low_hit_instances the count).
 * Replaces the MaxHits > 1 use of CSfxArrayV3::AlignReads (pMultiHits, Aligner.cpp:9220-9238) + WriteHitLoci. */
int bkx_align_reads_multi(bkx_index* idx, const bkx_align_params* params, const uint8_t* bases, const uint64_t* offsets,
                          uint32_t n_reads, bkx_read_result* out, bkx_multi_hit* multi, bkx_align_stats* stats);
/* Same call with the reads 4-bit packed on the host: base i of the concatenation in the low (i even) or high (i odd)
 * nibble of byte i/2, reads back to back at nibble granularity, offsets still counted in BASES.  Halves the bytes
 * that cross PCIe (and the host-memory traffic of an 8-GPU node); the nibbles are expanded on the device.  Nibble
 * values are etSeqBase codes 0..4; anything above counts as an N-class symbol.  bkx_pack_bases4 packs n_bases
 * one-byte codes (low 4 bits kept) into (n_bases + 1) / 2 bytes -- the loop a loader would fuse into its parser. */
int bkx_align_reads_packed4(bkx_index* idx, const bkx_align_params* params, const uint8_t* packed, const uint64_t* offsets,
                            uint32_t n_reads, bkx_read_result* out, bkx_align_stats* stats);
int bkx_pack_bases4(const uint8_t* bases, uint64_t n_bases, uint8_t* packed);
/* The compact host interface: what crosses PCIe per read is 2 bits per base in and 16 bytes out (the 4-bit call moves 4 bits
 * + 8 bytes of offset in and 32 bytes out) -- on a multi-GPU node the host memory system, not the GPUs, bounds the
 * host-buffer calls.
 *   packed2   base i of the concatenated reads at bits [2(i%4), +2) of byte i/4, codes A0 C1 G2 T3; reads back to back at
 *             base granularity.  Bases that are not A C G T are stored as 0 and listed in
 *   exc_pos / exc_code   (n_exc entries, exc_pos ascending): position in the concatenation and etSeqBase code (4 = N, ...).
 *   first_base  position in the stream of the first base of read 0 (a caller that shards one stream over several GPUs
 *             hands each its reads' lengths and the position they start at; exc_pos are stream positions too, the first
 *             entry at or after first_base).
 *   lens      read lengths (n_reads u16; cMaxSeqLen is 2000), or NULL when every read has fixed_len bases.
 *   out       n_reads 16-byte records.
 * Same search, same results as bkx_align_reads (the device expands the reads to the layout the kernels take).
 * bkx_pack_bases2 packs n_bases one-byte codes (low 3 bits) that way: packed2 needs (n_bases + 3) / 4 bytes (+ 8 bytes of
 * slack so that device copies may run in words); returns the number of exceptions, or BKX_ERR_MEM if exc_cap is too small. */
int bkx_align_reads_packed2(bkx_index* idx, const bkx_align_params* params, const uint8_t* packed2, uint64_t first_base,
                            const uint16_t* lens, uint32_t fixed_len, const uint64_t* exc_pos, const uint8_t* exc_code,
                            uint64_t n_exc, uint32_t n_reads, bkx_read_result16* out, bkx_align_stats* stats);
int64_t bkx_pack_bases2(const uint8_t* bases, uint64_t n_bases, uint8_t* packed2, uint64_t* exc_pos, uint8_t* exc_code,
                        uint64_t exc_cap);
/* 16-byte records -> 32-byte records (seeds / cands 0); lens / fixed_len as above. */
int bkx_expand_results16(const bkx_read_result16* in, uint32_t n_reads, const uint16_t* lens, uint32_t fixed_len,
                         bkx_read_result* out);
/* Device variant: all pointers are device pointers on idx's GPU; asynchronous on `cuda_stream`
 * (a cudaStream_t, NULL = the index's compute stream).  d_stats (device, may be NULL) is accumulated. */
int bkx_align_reads_device(bkx_index* idx, const bkx_align_params* p, const uint8_t* d_bases,
                           const uint64_t* d_offsets, uint32_t n_reads, uint32_t max_read_len,
                           bkx_read_result* d_out, bkx_align_stats* d_stats, void* cuda_stream);
/* Same, for reads that are resident in BOTH layouts: d_bases / d_offsets as above, and d_packed2 -- the same concatenation 2
 * bits per base (base i at bits [2(i%32), +2) of 64-bit word i/32; 8-byte aligned, 16 readable bytes beyond the last base;
 * non-ACGT bases stored as 0).  d_read_flags (one byte per read, or NULL when no read holds a non-ACGT base) marks the reads
 * that must be taken from d_bases.  The search takes every other read straight from the 2-bit words (no packing pass in the
 * kernel); reads it hands to the general kernel, and paired-end recovery, still use d_bases.  Same records as
 * bkx_align_reads_device. */
int bkx_align_reads_device_packed2(bkx_index* idx, const bkx_align_params* p, const uint8_t* d_bases, const uint64_t* d_packed2,
                                   const uint8_t* d_read_flags, const uint64_t* d_offsets, uint32_t n_reads,
                                   uint32_t max_read_len, bkx_read_result* d_out, bkx_align_stats* d_stats, void* cuda_stream);
/* Per-read shim with the exact CSfxArrayV3::AlignReads contract (SfxArrayV2.h:585-606) for unit
 * parity tests: one read in, tHRslt out, In/Out (LowHitInstances, LowMMCnt, NxtLowMMCnt). */
int bkx_align_one(bkx_index* idx, const bkx_align_params* p, const uint8_t* probe, int probe_len,
                  int* low_hit_instances, int* low_mm, int* nxt_low_mm, bkx_read_result* hit);

/* ---- paired ends: replaces CAligner::ProcessPairedEnds (Aligner.cpp:2726-2850, 3055-3489) -------
 * results[2*i], results[2*i+1] are PE1 / PE2 of pair i (as after SortReadHits(eRSMPairReadID),
 * Aligner.cpp:2899).  Updated in place; len_dist (may be NULL) is the insert-size histogram
 * m_pLenDist[0..100000] (Aligner.cpp:2908-2915), accumulated.  bases/offsets are needed only for
 * orphan recovery (pe_proc ORPHAN / ORPHAN_SE) and may be NULL otherwise. */
int bkx_pair_reads(bkx_index* idx, const bkx_align_params* p, const bkx_pe_params* pe, bkx_read_result* results,
                   uint32_t n_pairs, const uint8_t* bases, const uint64_t* offsets, bkx_pe_stats* stats,
                   uint32_t* len_dist);

/* Device variant: every pointer is a device pointer on idx's GPU; asynchronous on `cuda_stream`. */
int bkx_pair_reads_device(bkx_index* idx, const bkx_align_params* p, const bkx_pe_params* pe, bkx_read_result* d_results,
                          uint32_t n_pairs, const uint8_t* d_bases, const uint64_t* d_offsets, uint32_t max_read_len,
                          bkx_pe_stats* d_stats, uint32_t* d_len_dist, void* cuda_stream);

/* -Z / -z chromosome filters of a paired-end run: in the reference they act INSIDE the pairing (AcceptThisChromID called
 * from AcceptProvPE, the orphan-recovery arms and the SE fallback of ProcessPairedEnds: Aligner.cpp:2651-2710, 2771-2786,
 * 3296-3302, 3411-3417, 3442-3477), so the pairing kernels take them as a per-chromosome keep map held by the index:
 * keep[id] != 0 -- alignments to chromosome id stay (ids 1..num_entries; keep[0] is not read); n_keep must be
 * num_entries + 1.  The caller evaluates the expressions (exclude first, then -- if any -- the include ones).
 * keep = NULL clears the filter.  Consulted by bkx_pair_reads, bkx_pair_reads_device and bkx_align_pairs[_packed4] on
 * this index (not copied by bkx_clone_index); single-end filtering (FiltByChroms, Aligner.cpp:4019-4124) stays host work. */
int bkx_set_chrom_filter(bkx_index* idx, const uint8_t* keep, uint32_t n_keep);

/* Alignment and pairing of paired-end reads in ONE pass over the data: reads 2i / 2i+1 are PE1 / PE2 of pair i; each
 * slice of the internal pipeline is aligned and then paired (and its orphans recovered) while it is still on the GPU, so
 * the reads cross PCIe once.  Same records, counters and histogram as bkx_align_reads followed by bkx_pair_reads
 * (CAligner::ProcCoredApprox then ProcessPairedEnds, Aligner.cpp:8943-9527, 2876-3049).  The _packed4 form takes the
 * reads 4-bit packed (see bkx_align_reads_packed4).  stats / pe_stats / len_dist (100001 u32) are accumulated into. */
int bkx_align_pairs(bkx_index* idx, const bkx_align_params* params, const bkx_pe_params* pe, const uint8_t* bases,
                    const uint64_t* offsets, uint32_t n_pairs, bkx_read_result* out, bkx_align_stats* stats,
                    bkx_pe_stats* pe_stats, uint32_t* len_dist);
int bkx_align_pairs_packed4(bkx_index* idx, const bkx_align_params* params, const bkx_pe_params* pe, const uint8_t* packed,
                            const uint64_t* offsets, uint32_t n_pairs, bkx_read_result* out, bkx_align_stats* stats,
                            bkx_pe_stats* pe_stats, uint32_t* len_dist);

/* Paired ends through the compact host interface (see bkx_align_reads_packed2): reads 2i / 2i+1 are PE1 / PE2 of pair i. */
int bkx_align_pairs_packed2(bkx_index* idx, const bkx_align_params* params, const bkx_pe_params* pe, const uint8_t* packed2,
                            uint64_t first_base, const uint16_t* lens, uint32_t fixed_len, const uint64_t* exc_pos,
                            const uint8_t* exc_code, uint64_t n_exc, uint32_t n_pairs, bkx_read_result16* out,
                            bkx_align_stats* stats, bkx_pe_stats* pe_stats, uint32_t* len_dist);

/* ---- -r3 / -r4: one locus for reads that hit several, by clustering with the loci of other reads.  Replaces
 * CAligner::AssignMultiMatches + ProcAssignMultiMatches (Aligner.cpp:5108-5270, 4961-5105).  Host code (no device
 * work).  results / multi are what bkx_align_reads_multi returned under ml_mode BKX_ML_UNIQ or BKX_ML_MULTI: records of
 * reads with several loci are eNARMultiAlign with their count in low_hit_instances and their loci in multi; records of
 * reads that get a locus assigned are rewritten in place (accepted, one hit).  max_read_len: longest read of the run.
 * The counters are the ones of the reference's log lines (:5125, :5195, :5264). */
typedef struct bkx_cluster_stats {
  uint32_t multi_reads;   /* reads that aligned to several loci (m_NumProvMultiAligned)            */
  uint32_t putative;      /* reads considered for assignment                                        */
  uint32_t assigned;      /* reads given a locus                                                    */
  uint32_t near_unique;   /* ... because uniquely aligned reads overlap it                          */
  uint32_t near_multi;    /* ... because loci of other multi-loci reads overlap it (-r4 only)       */
  uint32_t reserved[3];
} bkx_cluster_stats;
int bkx_assign_multi_matches(bkx_read_result* results, uint32_t n_reads, const bkx_multi_hit* multi, int max_ml_matches,
                             int ml_mode, uint32_t max_read_len, bkx_cluster_stats* out);

/* ---- output order: replaces CAligner::SortReadHits(eRSMHitMatch) + SortHitMatch (Aligner.cpp:9917-9991,
 * 10067-10114).  order_out[k] = index of the k-th record under the reference's hit ordering (NAR class; uniquely
 * hit records by chromosome id, locus, match length, strand, mismatches; the others by NumHits); ties, which the
 * reference's unstable quicksort leaves unspecified, go by ascending index.  Host arrays in and out; the sort itself
 * is two radix passes on `device`. */
int bkx_sort_hits(const bkx_read_result* results, uint32_t n_reads, uint32_t* order_out, int device);

/* ---- index construction for synthetic / bench genomes: the two halves of `biokanga index` ---------
 * (kangax.cpp:774-926 -> CSfxArrayV3::QSortSeq, SfxArrayV2.cpp:9451-9542; file layout SfxArrayV2.h:79-104,
 * 174-187).  d_seq: device, 1 byte/base incl. one EOS(7) after every entry; d_sa: device, concat_len
 * u32 elements out, sorted like the reference sorts (4-bit symbol order, through the terminators). */
int bkx_build_suffix_array_device(const uint8_t* d_seq, uint64_t concat_len, uint32_t* d_sa, int device);
/* Same order, any size up to 2^40 symbols, in bounded device memory (batches of suffixes sharing leading symbols,
 * ties broken 21 symbols at a time straight from the sequence): the builder for genomes of >= 4e9 symbols, whose
 * elements take 5 bytes (SfxArrayV2.cpp:33-44).  Output as planes: d_sa_lo[i] = low 32 bits, d_sa_hi[i] = bits 32-39
 * (d_sa_hi may be NULL when concat_len <= 2^32).  max_batch = 0 sizes the batches from free device memory. */
int bkx_build_suffix_array_planes(const uint8_t* d_seq, uint64_t concat_len, uint32_t* d_sa_lo, uint8_t* d_sa_hi,
                                  int device, uint64_t max_batch);
/* Same builder, output as 5-byte elements back to back (concat_len * 5 bytes; allocate 16 more for bkx_open_index_packed5). */
int bkx_build_suffix_array_packed5(const uint8_t* d_seq, uint64_t concat_len, uint8_t* d_sa5, int device, uint64_t max_batch);
/* Write host-resident sequence + suffix array as a version-5 .sfx the reference's `biokanga align` loads. */
int bkx_write_sfx(const char* path, const uint8_t* seq, uint64_t concat_len, const void* sa, uint32_t sfx_el_size,
                  const bkx_entry* entries, uint32_t num_entries, const char* dataset_name);

/* ---- host memory ------------------------------------------------------------------------------
 * Page-lock caller buffers (read arena, result array) so bkx_align_reads() copies asynchronously at PCIe speed;
 * the reference keeps its reads in one mmap'd arena (Aligner.cpp:10572-10677) -- pin that arena once. */
int bkx_pin_host(void* ptr, size_t bytes);
/* Page-locked host memory from the start (NULL on failure; bkx_last_error() says why): for buffers the caller fills itself
 * and then hands to the host-buffer calls, instead of a registration pass over them afterwards. */
void* bkx_alloc_host(size_t bytes);
void bkx_free_host(void* ptr);
int bkx_unpin_host(void* ptr);

/* ---- instrumentation -------------------------------------------------------------------------- */
/* Device time (ms) of the kernels of the last bkx_align_reads* call on this index, measured with
 * CUDA events on the launching stream; <0 if none. */
float bkx_last_kernel_ms(const bkx_index* idx);
/* Number of kernel launches issued by this library on this index since it was opened. */
uint64_t bkx_kernel_launches(const bkx_index* idx);

#ifdef __cplusplus
}
#endif
#endif /* BKX_H */
