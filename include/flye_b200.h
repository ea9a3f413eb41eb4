/* flye_b200 — C ABI of the B200-native read-overlap hot path.
 *
 * The reference (Flye 2.8.1, /root/reference) has no FFI layer: the boundary of this path is the C++
 * class surface of src/sequence/ (SURVEY.md §8b).  The C++ mirror of those classes lives in
 * flye_b200/host/ and is a thin shim over the entry points below; each entry point names the
 * reference member function(s) whose work it replaces.
 *
 * Conventions: every function returns FG_OK (0) or a negative status; no exception crosses the ABI;
 * fg_last_error(ctx) returns the message the C++ shim rethrows as std::runtime_error (the reference's
 * own error texts are reproduced where it has them).  Host buffers are caller-owned unless stated.
 * A context is bound to one CUDA device and serialises its calls internally (one mutex), so the
 * shim's thread-safe members (lazySeqOverlaps/quickSeqOverlaps, overlap.h:397-411) may call in
 * concurrently.  There is no CPU fallback: if no CUDA device is usable fg_ctx_create fails.
 */
#ifndef FLYE_B200_H
#define FLYE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FG_OK              0
#define FG_ERR_CUDA       -1   /* CUDA runtime error (message has the call site)                          */
#define FG_ERR_ARG        -2   /* bad argument / call order                                               */
#define FG_ERR_KMER_SIZE  -3   /* "Can't use flat counter for k-mer size > 17"  (vertex_index.cpp:504-507) */
#define FG_ERR_TOO_FREQ   -4   /* "k-mer is too frequent"                       (vertex_index.cpp:372-375) */
#define FG_ERR_OVERFLOW   -5   /* "Input overflow" (>= 2^40 bases over both strands, sequence_container.cpp:386-391) */
#define FG_ERR_NCCL       -6
#define FG_ERR_INTERNAL   -7

typedef struct fg_ctx fg_ctx;

/* ---- context ------------------------------------------------------------------------------------ */
int         fg_ctx_create(int cuda_device, fg_ctx** out);
void        fg_ctx_destroy(fg_ctx* ctx);
const char* fg_last_error(const fg_ctx* ctx);
/* the CUDA stream (cudaStream_t) every kernel of this context is launched on — for external event timing */
void*       fg_stream(const fg_ctx* ctx);
/* number of kernels this library has launched on the context since creation (bench.py "gpu_launches") */
uint64_t    fg_kernel_launches(const fg_ctx* ctx);
/* device time (ms, CUDA events on the context's stream) of the phases of the most recent call;
 * names e.g. "count_clear","count","count_hist","count_table","select","emit","index_sort","index_table","lookup","expand",
 *        "hit_sort_radix","group","chain_dp","chain_fill","chain_walk","edit", host wall clocks as "host_*";
 *        calls[i] = how many times the phase ran (sub-batches);
 *        returns number written */
int         fg_last_timings(const fg_ctx* ctx, const char** names, float* ms, int* calls, int cap);

/* ---- reads: replaces SequenceContainer's storage for the device side ----------------------------------
 * (sequence.h:54-69 packing; sequence_container.cpp:48-79 ids; :359-392 global offsets)
 * Forward strands only; read i becomes sequence ids 2i (forward) and 2i+1 (reverse complement).
 * `packed`: 2-bit codes A0 C1 G2 T3, 32 per uint64, base j of a read at bits 2*(j%32) of word j/32
 * (exactly DnaSequence's chunk vector), read i starting at word word_offsets[i]; word_offsets has n+1 entries. */
int fg_reads_upload(fg_ctx* ctx, const uint64_t* packed, const uint64_t* word_offsets,
                    const uint32_t* lengths, uint32_t n_reads);
/* Same, from ASCII letters (ACGT/acgt only) — the 2-bit packing runs on the device (K1). */
int fg_reads_upload_ascii(fg_ctx* ctx, const char* bases, const uint64_t* base_offsets, uint32_t n_reads);

/* Optional second sequence set for queries that are not among the indexed reads — OverlapContainer(detector,
 * queryContainer) with a different container, as ReadAligner::alignReads does (read_aligner.cpp:178-217).  Same layout as
 * fg_reads_upload; query ids 2i / 2i+1 then name sequence i of THIS set and its reverse complement. */
int fg_queries_upload(fg_ctx* ctx, const uint64_t* packed, const uint64_t* word_offsets,
                      const uint32_t* lengths, uint32_t n_seqs);

/* ---- KmerCounter::count / getKmerHist (vertex_index.cpp:499-590) ----------------------------------- */
int fg_count_kmers(fg_ctx* ctx, int k, uint64_t* n_distinct);
/* histogram freq -> number of distinct canonical k-mers; call with NULLs to get the bin count */
int fg_kmer_hist(fg_ctx* ctx, uint64_t* freqs, uint64_t* counts, uint64_t* n_bins);
/* exact frequency of canonical k-mers (KmerCounter::getFreq, :593-616); parity tests only */
int fg_kmer_freq(fg_ctx* ctx, const uint64_t* canonical_kmers, uint32_t n, uint32_t* freq_out);

typedef struct {
    uint64_t n_keys;              /* k-mers present in the index (_kmerIndex.size())                      */
    uint64_t n_entries;           /* total positions stored                                               */
    uint64_t n_repetitive;        /* |_repetitiveKmers|                                                   */
    uint64_t repetitive_frequency;/* _repetitiveFrequency (vertex_index.cpp:186)                          */
    float    mean_frequency;      /* meanFrequency (:185)                                                 */
    float    sample_rate;         /* getSampleRate(): ctor value, or totalLen/totalEntries for minimizers */
} fg_index_stats;

/* VertexIndex::buildIndexUnevenCoverage (vertex_index.cpp:25-125) incl. yieldFrequentKmers (:316-358),
 * filterFrequentKmers (:173-212); needs fg_count_kmers first.  sample_rate = VertexIndex ctor argument. */
int fg_build_index_solid(fg_ctx* ctx, int min_freq, float select_rate, int tandem_freq,
                         float repeat_kmer_rate, float sample_rate, fg_index_stats* stats);
/* VertexIndex::buildIndexMinimizers (vertex_index.cpp:389-483) incl. yieldMinimizers (kmer.h:206-262). */
int fg_build_index_minimizers(fg_ctx* ctx, int k, int min_coverage, int window,
                              float repeat_kmer_rate, fg_index_stats* stats);
/* VertexIndex::clear() */
int fg_index_clear(fg_ctx* ctx);

/* iterKmerPos / isRepetitive / kmerFreq (vertex_index.h:220-246) for a batch of k-mers given in any
 * orientation.  For k-mer i: is_repetitive[i], size[i] (= kmerFreq) and first[i] = index of its first
 * position in the list store; fetch positions with fg_index_positions.  Positions are reported the way
 * KmerPosIterator::operator* reports them (already flipped to the query's orientation, :158-174). */
int fg_index_lookup(fg_ctx* ctx, const uint64_t* kmers, uint32_t n, uint8_t* is_repetitive,
                    uint32_t* size, uint64_t* first, uint8_t* rev_comp);
int fg_index_positions(fg_ctx* ctx, uint64_t first, uint32_t n, int rev_comp,
                       uint32_t* seq_ids, int32_t* positions);
/* whole index to the host (tests): keys ascending; call with NULL arrays to get the sizes */
int fg_index_export(fg_ctx* ctx, uint64_t* keys, uint8_t* is_repetitive, uint64_t* first, uint32_t* size,
                    uint64_t* n_keys, uint32_t* entry_seq_ids, int32_t* entry_pos, uint64_t* n_entries);

/* ---- OverlapDetector::getSeqOverlaps (overlap.cpp:99-508) for a batch of query sequences ------------- */
typedef struct {
    int32_t max_jump;             /* _maxJump                                                            */
    int32_t min_overlap;          /* _minOverlap                                                         */
    int32_t max_overhang;         /* _maxOverhang (checkOverhang = max_overhang > 0, overlap.h:324)      */
    int32_t max_overlaps;         /* maxOverlaps argument, 0 = unlimited (:218-219)                      */
    int32_t force_local;          /* forceLocal argument                                                 */
    int32_t keep_alignment;       /* _keepAlignment: return kmerMatches                                  */
    int32_t only_max_ext;         /* _onlyMaxExt                                                         */
    int32_t nucl_alignment;       /* _nuclAlignment: divergence = edit distance of (HPC) substrings      */
    int32_t use_hpc;              /* _useHpc                                                             */
    float   max_divergence;       /* _maxDivergence                                                      */
    int32_t query_set;            /* 0: query ids name the indexed reads; 1: the set of fg_queries_upload  */
    const float* query_max_divergence; /* optional (NULL = max_divergence for every query): n_queries thresholds, one per
                                          query.  estimateOverlaperParameters' unfiltered sample (threshold 1.0,
                                          overlap.cpp:744-790) can then share a batch with the thresholded main pass.  With
                                          nucl_alignment the threshold also bounds the edit-distance work: an overlap that
                                          cannot pass it (overlap.cpp:470-473) is dropped without finishing its alignment */
    int32_t keep_rejected;        /* _partitionBadMappings support: primary overlaps that FAIL the divergence test are returned too,
                                     in the reference's order, with fg_overlap.reserved = 1 — the caller hands them to
                                     checkIdyAndTrim (overlap.cpp:475-485, alignment.cpp:306-495).  Needs max_overlaps = 0 */
    int32_t pad_;
} fg_overlap_params;

typedef struct {                  /* OverlapRange (overlap.h:20-279) plus what seqDivergence is made of  */
    uint32_t cur_id;  int32_t cur_begin, cur_end, cur_len;
    uint32_t ext_id;  int32_t ext_begin, ext_end, ext_len;
    int32_t  score;
    float    seq_divergence;
    int32_t  chain_length;        /* k-mer divergence inputs (overlap.cpp:409-423)                        */
    int32_t  filtered_positions;
    int32_t  edit_distance;       /* -1 unless nucl_alignment                                            */
    int32_t  aln_len;             /* max(len(HPC cur), len(HPC ext)), 0 unless nucl_alignment            */
    uint64_t aln_first;           /* keep_alignment: index of first kmerMatches pair, count in aln_count  */
    uint32_t aln_count;
    uint32_t reserved;            /* 0; 1 = failed the divergence test, returned because keep_rejected was set */
} fg_overlap;

typedef struct {
    uint32_t    n_queries;
    const uint64_t*   offsets;    /* n_queries+1: overlaps of query i are [offsets[i], offsets[i+1])      */
    const fg_overlap* overlaps;   /* in the reference's order (ascending extId group order)              */
    const int32_t*    aln_pairs;  /* keep_alignment: (cur,ext) pairs, 2 ints each                        */
    uint64_t    n_aln_pairs;
    /* workload counters of this call */
    uint64_t    n_hits, n_pairs, n_dp_pairs, n_dp_cells;
} fg_overlap_result;

/* Results live in library-owned host memory, valid until the next fg_overlaps_batch on the same ctx
 * or fg_ctx_destroy.  query_ids may name either strand. */
int fg_overlaps_batch(fg_ctx* ctx, const uint32_t* query_ids, uint32_t n_queries,
                      const fg_overlap_params* params, fg_overlap_result* result);

/* OverlapContainer::setDivergenceThreshold after estimateOverlaperParameters (overlap.cpp:744-827): re-applies a (tighter)
 * threshold to the overlaps of the LAST fg_overlaps_batch on this ctx — records of the queries [first_query, n_queries) with
 * seq_divergence >= max_divergence are removed (queries before first_query, the estimate sample, stay as they are).
 * `result` is refreshed (same lifetime rules).  Presets with a RELATIVE threshold (asm_raw_reads) only know it after the
 * sample; this keeps the sample and the main pass in one device batch.
 * Only a threshold at most as large as the one the batch was computed with can be applied, and only to a batch run with
 * max_overlaps = 0 (FG_ERR_ARG otherwise): records the batch dropped are gone. */
int fg_overlaps_refilter(fg_ctx* ctx, uint32_t first_query, float max_divergence, fg_overlap_result* result);

/* The rest of OverlapContainer::findAllOverlaps (overlap.cpp:630-668) on the device: ensureTransitivity(false) (:576-627) and
 * filterOverlaps() (:681-741).  `records`: the getSeqOverlaps vectors of the FORWARD sequences (any order; host memory).  Result:
 * for every sequence id 0 .. n_seqs-1 (both strands) its overlaps after the symmetric closure, the clustering of near-identical
 * overlaps (ends within max_ends_diff = Parameters::kmerSize on both sequences, best score stays) and the sort by cur_begin.
 * Every result record is a variant of an input record: reserved = 4 * input index + variant, variant bit 0 = complement(),
 * bit 1 = reverse() applied after it (overlap.h:95-147) — enough to rebuild kmerMatches on the host.  The order among equal
 * cur_begin is unspecified, as in the reference.  Same result lifetime rules as fg_overlaps_batch. */
int fg_overlaps_closure(fg_ctx* ctx, const fg_overlap* records, uint64_t n_records, uint32_t n_seqs, int32_t max_ends_diff,
                        fg_overlap_result* result);

/* ---- multi-GPU (one context per rank; reads are partitioned, the index replicated) ------- */
#define FG_NCCL_ID_BYTES 128
int fg_comm_unique_id(uint8_t id[FG_NCCL_ID_BYTES]);
int fg_comm_init(fg_ctx* ctx, int n_ranks, int rank, const uint8_t id[FG_NCCL_ID_BYTES]);
/* this rank's share of the forward reads for counting / index emission: [first_read, first_read+n) */
int fg_comm_set_shard(fg_ctx* ctx, uint32_t first_read, uint32_t n_reads);

/* ---- measurement hook: the chip's integer-issue ceiling, in Gop/s (one op = one IMAD / LOP3 / SHF thread
 * instruction; all SMs, full occupancy, no memory traffic).  bench.py reports the integer-bound kernels against it. ---- */
int fg_debug_int_peak(fg_ctx* ctx, double* gops);

/* ---- test hook: the std::sort-exact warp introsort on caller data (segments sorted independently) ---- */
int fg_debug_edit_distance(fg_ctx* ctx, const uint8_t* a, int n, const uint8_t* b, int m, int* distance);
/* the same with either string read as its reverse complement (base c -> 3 - c), as the kernel reads a reverse strand */
int fg_debug_edit_distance_rc(fg_ctx* ctx, const uint8_t* a, int n, int rc_a, const uint8_t* b, int m, int rc_b, int* distance);
int fg_debug_warp_sort(fg_ctx* ctx, uint64_t* keys, uint32_t* vals, const uint64_t* seg_offsets, uint32_t n_segments);
/* ---- getAlignmentCigarKsw (alignment.cpp:102-216) for a batch of sequence pairs: the banded affine-gap global alignment behind
 * checkIdyAndTrim (overlap.cpp:475-485, _partitionBadMappings) — minimap2's ksw_extz2_sse restated byte for byte (csrc/ksw_core.cuh; match 2,
 * mismatch -4, gap open 4, gap extend 2).  Base codes 0..3; pair i aligns queries[query_offsets[i] .. query_offsets[i+1]) against
 * targets[target_offsets[i] .. target_offsets[i+1]); band 64, doubled while it cannot connect the corners (alignment.cpp:150-165).
 * cigars[i * cigar_cap ..] receives n_cigar[i] entries len << 4 | op (0 = M, 1 = I: query only, 2 = D: target only); status[i]: 0 ok, 1 no band up
 * to the longer sequence connects the corners (empty CIGAR, as the reference), 2 cigar_cap too small (tlen + qlen entries always suffice).
 * The host mirror calls it for the rejected primaries of a batch when FLYE_B200_DEVICE_KSW=1; by default it still hands them to the
 * reference's own checkIdyAndTrim on host threads. ---- */
int fg_align_cigar_batch(fg_ctx* ctx, const uint8_t* targets, const uint64_t* target_offsets, const uint8_t* queries,
                       const uint64_t* query_offsets, uint32_t n_pairs, uint32_t cigar_cap, uint32_t* cigars, uint32_t* n_cigar,
                       int32_t* status);

#ifdef __cplusplus
}
#endif
#endif /* FLYE_B200_H */
