// fasta.cuh -- launchers of the GPU FASTA front end (fasta.cu).
#pragma once
#include "common.cuh"

namespace dsmfm {

// tiles of the per-byte passes over m input bytes / of the per-record passes over nrec records
uint64_t fasta_tiles(uint64_t m);
uint64_t fasta_rec_tiles(uint64_t nrec);

// Pass 1: line structure.  Scratch arrays hold fasta_tiles(m) entries each.  totals[0] = sequence bytes,
// totals[1] = header lines (device memory, two u64).
void launch_fasta_scan_lines(cudaStream_t st, const uint8_t *text, uint64_t m, long long *last_nl, long long *entry,
                             uint32_t *cnt_seq, uint32_t *cnt_hdr, uint64_t *off_seq, uint64_t *off_hdr, uint64_t *totals,
                             uint32_t *launches);

// Pass 2: records.  nrec = header lines + 1 (record 0 = rows in front of the first header).  B has nrec + 1
// entries; the caller sets B[0] = 0 and B[nrec] = sequence bytes before the launch.  O[nrec] receives the
// document offsets, rec_total[0] the number of non-empty records; counters[0] counts blank headers.
void launch_fasta_records(cudaStream_t st, const uint8_t *text, uint64_t m, const long long *entry, const uint64_t *off_seq,
                          const uint64_t *off_hdr, uint64_t nrec, uint64_t *B, uint64_t *O, uint32_t *rec_cnt,
                          uint64_t *rec_off, uint64_t *rec_total, unsigned long long *counters, uint32_t *launches);

// Pass 3: documents into `out` (2 * sequence bytes + 2 * non-empty records bytes).  bad_bitmap: one zeroed bit
// per record; counters[1] counts records with symbols normalize() turns into N, counters[2] (preset to ~0)
// receives the smallest input offset of such a symbol.
void launch_fasta_emit(cudaStream_t st, const uint8_t *text, uint64_t m, const long long *entry, const uint64_t *off_seq,
                       const uint64_t *off_hdr, const uint64_t *B, const uint64_t *O, uint8_t *out, uint32_t *bad_bitmap,
                       unsigned long long *counters, uint32_t *launches);

} // namespace dsmfm
