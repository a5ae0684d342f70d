/* TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's `biokanga align` hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this.  The product (biokanga_b200/) never links or calls it.
 * Shares only POD layouts with include/bkx.h so results can be compared field by field. */
#ifndef BK_ORACLE_H
#define BK_ORACLE_H
#include "../include/bkx.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct bko_index bko_index;

int bko_open(const char* sfx_path, bko_index** out);
int bko_open_mem(const uint8_t* seq, uint64_t concat_len, const void* sa, uint32_t sfx_el_size,
                 const bkx_entry* entries, uint32_t num_entries, bko_index** out); /* borrows the buffers */
void bko_close(bko_index* idx);
int bko_info(const bko_index* idx, bkx_index_info* out);
int bko_get_entry(const bko_index* idx, uint32_t entry_id, bkx_entry* out);
const uint8_t* bko_seq(const bko_index* idx);
const void* bko_sa(const bko_index* idx);

int bko_default_params(const bko_index* idx, int pmode, bkx_align_params* out);

/* SfxArrayV2.cpp:7765-7876 / 7914-8027: index+1 of first / last exact match, 0 if none. */
int64_t bko_locate_first_exact(const bko_index* idx, const uint8_t* probe, int probe_len, int64_t lo, int64_t hi);
int64_t bko_locate_last_exact(const bko_index* idx, const uint8_t* probe, int probe_len, int64_t lo, int64_t hi);

/* SfxArrayV2.cpp:7666-7760 (AlignReads) with indel/splice/chimeric passes disabled. */
int bko_align_reads_one(const bko_index* idx, const bkx_align_params* p, int max_tot_mm, int core_len,
                        int core_delta, int max_num_core_slides, uint8_t* probe, int probe_len,
                        int* low_hit_instances, int* low_mm, int* nxt_low_mm, bkx_read_result* hit,
                        uint32_t* seeds, uint32_t* cands);

/* Aligner.cpp:9024-9505 (ProcCoredApprox body) for a batch; nthreads pthreads. */
int bko_align_batch(const bko_index* idx, const bkx_align_params* p, const uint8_t* bases,
                    const uint64_t* offsets, uint32_t n_reads, bkx_read_result* out, bkx_align_stats* stats,
                    int nthreads);
/* same with the -r5 loci of every read: max_ml_matches slots per read (see bkx_align_reads_multi) */
int bko_align_batch_multi(const bko_index* idx, const bkx_align_params* p, const uint8_t* bases, const uint64_t* offsets,
                          uint32_t n_reads, bkx_read_result* out, bkx_multi_hit* multi, bkx_align_stats* stats, int nthreads);

/* Aligner.cpp:2726-2850, 3055-3489 (ProcessPairedEnds). */
int bko_pair_reads(const bko_index* idx, const bkx_align_params* p, const bkx_pe_params* pe,
                   bkx_read_result* results, uint32_t n_pairs, const uint8_t* bases, const uint64_t* offsets,
                   bkx_pe_stats* stats, uint32_t* len_dist);

/* Same with the -Z / -z chromosome filters acting inside the pairing (AcceptThisChromID in AcceptProvPE and around the
 * orphan recovery, Aligner.cpp:2771-2786, 3170-3190, 3224, 3296-3302, 3323, 3411-3417, 3442-3477): keep[entry id] != 0
 * for chromosomes that pass, NULL for none.  Ahead of the CUDA path, which does not take the filter yet. */
int bko_pair_reads_filtered(const bko_index* idx, const bkx_align_params* p, const bkx_pe_params* pe,
                            bkx_read_result* results, uint32_t n_pairs, const uint8_t* bases, const uint64_t* offsets,
                            bkx_pe_stats* stats, uint32_t* len_dist, const uint8_t* keep);

#ifdef __cplusplus
}
#endif
#endif
