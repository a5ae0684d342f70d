// Host-callable launchers of the kernels in bkx_kernels.cu (internal to libbkx.so).
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/bkx.h"

namespace bkx {
struct DevIndex;
struct KParams;
struct HashPool;
struct Packed2Src;

cudaError_t launch_pack_genome(const uint8_t* seq, uint64_t n, uint64_t* g2, uint64_t* gx, uint32_t* gxc,
                               unsigned long long* bad, cudaStream_t st);
cudaError_t launch_verify_index(const DevIndex& I, int k, unsigned long long* n_bad, cudaStream_t st);
cudaError_t launch_unpack4(const uint8_t* packed, unsigned phase, uint64_t n_bases, uint8_t* out, cudaStream_t st);
cudaError_t launch_unpack2(const uint8_t* packed, unsigned phase, uint64_t n_bases, uint8_t* out, cudaStream_t st);
cudaError_t launch_scatter_exceptions(const uint64_t* pos, const uint8_t* code, uint32_t n_exc, uint64_t first_base,
                                      uint64_t n_bases, uint8_t* out, cudaStream_t st);
cudaError_t launch_fixed_offsets(uint64_t* offs, uint32_t n_reads, uint32_t len, cudaStream_t st);
// tmp == nullptr: size query into *tmp_bytes
cudaError_t launch_len_offsets(const uint16_t* lens, uint32_t n_reads, uint64_t* offs, void* tmp, size_t* tmp_bytes, cudaStream_t st);
cudaError_t launch_compact_results(const bkx_read_result* in, uint32_t n, bkx_read_result16* out, cudaStream_t st);
cudaError_t launch_split_sa5(const uint8_t* sa5, uint64_t n, uint32_t* lo, uint8_t* hi, cudaStream_t st);
cudaError_t build_prefix_table(const DevIndex& I, int k, uint32_t* table, uint64_t* block_starts, cudaStream_t st);
cudaError_t launch_merge_sa5(const uint32_t* lo, const uint8_t* hi, uint64_t n, uint8_t* sa5, cudaStream_t st);

size_t align_smem_bytes(int W);
int align_blocks_per_sm(int W);
cudaError_t launch_align(const DevIndex& I, const KParams& P, const uint8_t* bases, const uint64_t* offs,
                         uint32_t n_reads, int W, bkx_read_result* out, bkx_align_stats* stats, unsigned int* cursor,
                         const HashPool& hp, const uint32_t* ids, const unsigned int* n_ids, int grid, cudaStream_t st);
int fast_blocks_per_sm(int W);
cudaError_t launch_align_fast(const DevIndex& I, const KParams& P, const uint8_t* bases, const uint64_t* offs,
                              uint32_t n_reads, int W, bkx_read_result* out, bkx_align_stats* stats,
                              unsigned int* cursor, uint32_t* hard_ids, unsigned int* n_hard, uint64_t* lane_hash,
                              uint32_t epoch_base, int grid, cudaStream_t st, const Packed2Src& p2,
                              const uint32_t* ids = nullptr, const unsigned int* n_ids = nullptr);
struct WaveBuf;
int wave_launches(const KParams& P, uint32_t max_len, bool sa_split);   // kernels one launch_wave starts
// the wave path (bkx_wave.cuh): results for the reads it finishes, B.fb_ids / B.cnt[kWaveCntFallback] for the others
cudaError_t launch_wave(const DevIndex& I, const KParams& P, const uint64_t* offs, uint32_t n_reads, uint32_t max_len,
                        const Packed2Src& p2, const WaveBuf& B, bkx_read_result* out, bkx_align_stats* stats, int sms,
                        cudaStream_t st);
cudaError_t launch_flag_exception_reads(const uint64_t* pos, uint32_t n_exc, uint64_t first_base, const uint64_t* offs,
                                        uint32_t n_reads, uint8_t* flags, cudaStream_t st);
cudaError_t launch_pair(const bkx_pe_params& pe, bkx_read_result* res, uint32_t n_pairs, bkx_pe_stats* stats,
                        uint32_t* len_dist, uint32_t* orphan_list, unsigned int* n_orphans, const uint8_t* keep,
                        cudaStream_t st);
cudaError_t launch_rescue(const DevIndex& I, const KParams& P, const bkx_pe_params& pe, bkx_read_result* res,
                          const uint32_t* orphan_list, const unsigned int* n_orphans, const uint8_t* bases,
                          const uint64_t* offs, int Lmax, bkx_pe_stats* stats, uint32_t* len_dist, unsigned int* cursor,
                          const uint8_t* keep, cudaStream_t st);
}  // namespace bkx
