/* ma_b200 — C ABI of the B200-native (sm_100a) implementation of MA's read-alignment hot path.
 *
 * This is the drop-in boundary: plain C, pointers and sizes only, no torch / STL types.  The reference
 * (ITBE-Lab/ma) has no FFI for this path; each entry point below names the reference interface it replaces.
 * The reference-side binding a maintainer would add is shown in INTEGRATION.md.
 *
 * Conventions: every function returns 0 on success and a negative MA_B200_E* code on failure (never throws across
 * the ABI); ma_b200_last_error() returns a message for the calling context.  A context owns one CUDA device, one
 * stream and its device buffers; a context is not re-entrant (use one per host thread / per GPU).  All `*_ms`
 * out-parameters are CUDA-event times measured on the context's stream.  There is NO CPU fallback: without a
 * CUDA device ma_b200_create fails.
 *
 * Nucleotide code: A=0 C=1 G=2 T=3 N=4 (reference: libs/ma/src/container/nucSeq.cpp:17-28).
 */
#ifndef MA_B200_H
#define MA_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MA_B200_OK 0
#define MA_B200_ECUDA -1 /* CUDA runtime error */
#define MA_B200_EINVAL -2 /* invalid argument */
#define MA_B200_ENOMEM -3 /* an output slab supplied by the caller is too small */
#define MA_B200_ESTATE -4 /* call order violated (e.g. no index uploaded) */

typedef struct ma_b200_ctx ma_b200_ctx;

/* Flattened copy of the values the hot path reads from the reference's ParameterSetManager presets and from
 * pGlobalParams (libs/ms/inc/ms/util/parameter.h:521-917, 1004-1064, 1079-1132).  ma_b200_params_preset() fills
 * it for "default" | "illumina" | "illumina_paired" | "pacbio" | "nanopore". */
typedef struct
{
    /* pGlobalParams: scoring (penalties are positive) — parameter.h:1032-1046 */
    int32_t match, mismatch, gap, extend, gap2, extend2, sv_penalty;
    /* seeding — parameter.h:671-701 */
    int32_t seeding_technique; /* 0 = maxSpan, 1 = SMEMs */
    int32_t min_seed_length; /* "Minimal Seed Length", filter is on segment size => length >= this + 1 */
    int32_t min_ambiguity, max_ambiguity;
    int32_t seed_drop_min_size; /* "Seeding Drop-off A" */
    double seed_drop_factor; /* "Seeding Drop-off B" */
    /* strip of consideration — parameter.h:703-718 */
    int32_t max_num_soc, min_num_soc, soc_width, rectangular_soc;
    /* harmonization heuristics — parameter.h:822-880 */
    double soc_score_drop; /* "SoC Score Drop-off" */
    int32_t harm_score_min;
    double harm_score_min_rel;
    double score_diff_tolerance;
    int32_t max_score_lookahead, switch_qlen;
    double max_delta_dist;
    int32_t min_delta_dist;
    int32_t optimistic_gap_estimation, gap_cost_cutting;
    int32_t max_gap_area; /* "Maximal Gap Size": larger gaps use the dual extension */
    int64_t genome_size_disable; /* "Minimum Genome Size for Heuristics" */
    int32_t disable_heuristics;
    /* dynamic programming — parameter.h:621-639 */
    int32_t padding, bandwidth_ext, min_bandwidth_gap, zdrop;
    /* RANSAC draws come from glibc's TYPE_3 rand(); the reference never seeds it itself.  Parity contract
     * (SURVEY.md A-5): the stream for read i is seeded with srand(srand_base + i). */
    uint32_t srand_base;
    /* MappingQuality (mappingQuality.h:26-36, parameter.h:721-737) and PairedReads (pairedReads.h:43-55,
     * parameter.h:649-668); use_paired_reads: reads 2k and 2k+1 of a batch are mates ("Use Paired Reads") */
    int32_t report_n, min_alignment_score, max_supplementary_per_prim, use_paired_reads;
    double max_overlap_supplementary, paired_mean, paired_std, paired_bonus;
} ma_b200_params;

int ma_b200_params_preset( const char* name, ma_b200_params* out );

/* ---- context ---------------------------------------------------------------------------------------------- */
int ma_b200_create( int device, ma_b200_ctx** out );
void ma_b200_destroy( ma_b200_ctx* ctx );
/* A second context on the same device that SHARES the index of ctx (a view: ctx keeps owning it and must outlive the
 * sibling; re-uploading / rebuilding the index of ctx invalidates the view until ma_b200_create_sibling is called again)
 * and starts with its parameters. For a stream of batches: two host threads, each calling ma_b200_align_batch on its own
 * context, keep two batches in flight — the host<->device copies of one run under the kernels of the other (what
 * maCMD_b200 does per GPU, and bench.py's e2e figure). A context itself stays single-threaded. */
int ma_b200_create_sibling( ma_b200_ctx* ctx, ma_b200_ctx** out );
const char* ma_b200_last_error( const ma_b200_ctx* ctx );
int ma_b200_set_params( ma_b200_ctx* ctx, const ma_b200_params* params );
/* number of kernels this context has launched so far (for bench.py's gpu_launches) */
int64_t ma_b200_launch_count( const ma_b200_ctx* ctx );
/* Page-locked host memory for the caller's batch buffers (reads slab, record arrays): copies from pageable memory are
 * staged by the driver at a fraction of the link rate.  The role AlignedMemoryManager (kswcpp_mem.h:319-334) plays for
 * the reference's DP scratch: caller-owned, reused across calls.  NULL if the allocation fails. */
void* ma_b200_host_alloc( int64_t bytes );
void ma_b200_host_free( void* p );

/* ---- banded DP: replaces kswcpp_dispatch (libs/kswcpp/inc/kswcpp.h:165-190) -------------------------------- */
#define MA_B200_KSW_RIGHT 0x02 /* KSW_EZ_RIGHT */
#define MA_B200_KSW_EXTZ_ONLY 0x40 /* KSW_EZ_EXTZ_ONLY */
#define MA_B200_KSW_REV_CIGAR 0x80 /* KSW_EZ_REV_CIGAR */

typedef struct
{
    int64_t qoff, toff; /* byte offsets of query / target in the sequence slab */
    int32_t qlen, tlen, w, zdrop, flag, tag; /* tag: caller's cookie, ignored */
} ma_b200_ksw_task;

/* kswcpp_extz_t (kswcpp.h:31-41); the cigar lives at cigar[cigar_off .. cigar_off + n_cigar), word = len<<4 | op,
 * op 0 = M, 1 = I, 2 = D.  cells = band cells processed (the GCUPS unit). */
typedef struct
{
    int32_t max, zdropped, max_q, max_t, mqe, mqe_t, mte, mte_q, score, n_cigar, reach_end, status;
    int64_t cigar_off;
    int64_t cells;
} ma_b200_ksw_result;

/* Extension tasks (flag & 0x40, KSW_EZ_EXTZ_ONLY) of the batches uploaded afterwards run in the early-termination
 * mode the alignment path uses for NeedlemanWunsch's end extensions (needlemanWunsch.cpp:486-541 reads only the
 * maximum, its position and the CIGAR): max, max_q, max_t, n_cigar and the CIGAR are identical to the reference's,
 * the remaining result fields (zdropped, mqe, mte, score, reach_end) are undefined. Off by default. */
int ma_b200_ksw_set_extension_only( ma_b200_ctx* ctx, int32_t on );

/* Three-step form: inputs stay resident in HBM between upload and run (what bench.py times as `value`). */
int ma_b200_ksw_upload( ma_b200_ctx* ctx, int64_t n, const ma_b200_ksw_task* tasks, const uint8_t* seq,
                        int64_t seq_bytes );
int ma_b200_ksw_run( ma_b200_ctx* ctx, float* kernel_ms );
int ma_b200_ksw_download( ma_b200_ctx* ctx, ma_b200_ksw_result* results, uint32_t* cigar, int64_t cigar_cap_words,
                          int64_t* cigar_words );
/* One-call form with host buffers (host<->device copies inside; what bench.py times as `e2e`). */
int ma_b200_ksw_batch( ma_b200_ctx* ctx, int64_t n, const ma_b200_ksw_task* tasks, const uint8_t* seq,
                       int64_t seq_bytes, ma_b200_ksw_result* results, uint32_t* cigar, int64_t cigar_cap_words,
                       int64_t* cigar_words );

/* ---- index: replaces FMIndex / Pack loading (fMIndex.h:854-884, pack.h:513-525, 799-812) --------------------- */
/* Uploads (replicates) the reference's own index arrays into HBM:
 *   bwt_words : FMIndex::bwt, 16 x u32 per 128 symbols = 4 x u64 counts + 8 x u32 symbols (fMIndex.h:434-437)
 *   L2[5]     : cumulative counts, L2[0] = 0 (fMIndex.h:195);  primary (fMIndex.h:200);  ref_len = forward + reverse
 *   sa        : one sample per sa_intv rows, sa[0] = -1 (fMIndex.h:602-663)
 *   pac       : Pack::xPackedNucSeqs, 2 bit per forward-strand base (pack.h:172-176);  fwd_len = forward length
 *   contig_start / contig_len : Pack::xVectorOfSequenceDescriptors (pack.h:823-860) */
int ma_b200_index_upload( ma_b200_ctx* ctx, const uint32_t* bwt_words, int64_t n_words, const int64_t* L2,
                          int64_t primary, int64_t ref_len, const int64_t* sa, int64_t n_sa, int32_t sa_intv,
                          const uint8_t* pac, int64_t n_pac_bytes, int64_t fwd_len, const int64_t* contig_start,
                          const int64_t* contig_len, int32_t n_contigs );
/* Builds the same index on the GPU from the forward strand (1 byte per base, codes 0..3, contigs concatenated):
 * suffix array of forward ++ reverse-complement by prefix doubling, BWT, occurrence blocks, SA samples.  The result
 * is bit-identical to FMIndex(pPack) (fMIndex.cpp:316-391) and stays resident; ma_b200_index_download copies it
 * out in the layout of ma_b200_index_upload (sizes via ma_b200_index_sizes). */
int ma_b200_index_build( ma_b200_ctx* ctx, const uint8_t* fwd, int64_t fwd_len, const int64_t* contig_start,
                         const int64_t* contig_len, int32_t n_contigs );
int ma_b200_index_sizes( ma_b200_ctx* ctx, int64_t* n_words, int64_t* n_sa, int64_t* n_pac_bytes, int64_t* primary,
                         int64_t* L2 /* 5 */ );
int ma_b200_index_download( ma_b200_ctx* ctx, uint32_t* bwt_words, int64_t* sa, uint8_t* pac );

/* ---- the alignment path ------------------------------------------------------------------------------------ */
/* replaces, per read, the chain wired by setUpCompGraph (libs/ma/src/util/export.cpp:99-126):
 *   BinarySeeding::execute -> StripOfConsideration::execute (ExtractSeeds + StripOfConsiderationSeeds)
 *   -> Harmonization::execute -> NeedlemanWunsch::execute */
#define MA_B200_STAGE_SEEDS 1 /* BinarySeeding + ExtractSeeds: located seeds in emission order */
#define MA_B200_STAGE_SETS 2 /* + StripOfConsiderationSeeds + Harmonization: harmonized seed sets */
#define MA_B200_STAGE_ALIGN 3 /* + NeedlemanWunsch: alignments */
#define MA_B200_STAGE_MAPQ 4 /* + MappingQuality, and PairedReads over reads (2k, 2k+1) if use_paired_reads */

typedef struct /* Seed (seed.h:34-43) */
{
    int32_t q, len;
    int64_t r; /* start on the (folded) forward strand, see segment.h:99-105 */
    uint32_t ambiguity;
    int32_t on_forward;
    int64_t delta;
} ma_b200_seed;

typedef struct /* Segment (segment.h:31-115): query interval (size = length - 1) + SAInterval */
{
    int32_t start, size;
    int64_t sa_start, sa_rev_start, sa_size;
} ma_b200_segment;

typedef struct /* one harmonized seed set (Seeds + xStats.index_of_strip) */
{
    int32_t read, ordinal;
    uint32_t soc_index;
    int32_t n;
    int64_t seed_off;
    int32_t task_off, n_tasks;
    uint64_t win_begin, win_end;
    int32_t valid, pad;
} ma_b200_seed_set;

typedef struct /* Alignment (alignment.h:55-95); runs: word = len << 3 | MatchType (seed 0, match 1, missmatch 2, */
{ /*              insertion 3, deletion 4), alignment.h:40-47 */
    int64_t begin_ref, end_ref, score;
    int32_t begin_q, end_q;
    int32_t length, n_runs;
    uint32_t soc_index;
    int32_t read;
    int64_t run_off;
    int32_t rank; /* position in the read's result vector after the reference's final sort */
    int32_t flags; /* MA_B200_ALN_*; valid after MA_B200_STAGE_MAPQ */
    double mapq; /* Alignment::fMappingQuality (SAM MAPQ = ceil(mapq * 254)); NaN where the reference leaves it unset */
    int32_t rank_mq; /* position in MappingQuality's result vector, -1: not reported (n-best / minimal score) */
    int32_t pair_rank; /* paired mode: position in PairedReads' result vector, -1: not in it */
} ma_b200_alignment;
#define MA_B200_ALN_SECONDARY 1
#define MA_B200_ALN_SUPPLEMENTARY 2
#define MA_B200_ALN_FIRST_MATE 4

typedef struct /* per read: where its seeds / sets / alignments are */
{
    int64_t seed_off;
    int32_t n_seeds;
    int32_t set_off; /* also the offset of the read's alignments (one per set) */
    int32_t n_sets;
    int32_t status; /* 0, or MA_B200_READ_* bits: the read exceeded a capacity of this implementation; it has no (ELISTS,
                     * ESEGMENTS, ESETS) or partial (EBAND: the seed sets concerned give empty records) results. The
                     * reference has no such capacities; the other reads of the batch are not affected. */
} ma_b200_read_info;
#define MA_B200_READ_ELISTS 1 /* more than 1 024 SMEM interval-list entries for one seed centre */
#define MA_B200_READ_ESEGMENTS 2 /* more than min(2 L + 8, 65 536) filtered segments */
#define MA_B200_READ_ESETS 4 /* more than 128 harmonized seed sets (ma_b200_set_params rejects max_num_soc > 128) */
#define MA_B200_READ_EBAND 8 /* a DP problem wider than the largest band window (about 2 000 columns) */

typedef struct
{
    int64_t n_reads, n_seeds, n_sets, n_set_seeds, n_tasks, n_runs, n_cigar_words;
    int64_t n_ext; /* FMIndex::extend_backward calls: 128 algorithmic bytes each */
    int64_t n_invpsi; /* bwt_invPsi steps: 64 algorithmic bytes each */
    int64_t n_dropped; /* reads cleared by the seeding drop-off */
    int64_t dp_cells; /* band cells */
    int64_t n_lookup; /* extend_backward calls that really read the occurrence table (the rest reuse the previous
                         result of the same SA interval, see fmindex.cuh SeederSM) */
    float ms_seed, ms_locate, ms_socharm, ms_plan, ms_dp, ms_assemble, ms_total;
    int32_t launches;
    int32_t n_failed; /* reads with a non-zero ma_b200_read_info::status */
    int64_t n_reported; /* alignment records with rank_mq >= 0 (what ma_b200_set_reported_only downloads) */
} ma_b200_align_stats;

/* reads: concatenated, 1 byte per base; offsets[n_reads + 1].  Stays resident until the next upload. */
int ma_b200_align_upload( ma_b200_ctx* ctx, int64_t n_reads, const uint8_t* reads, const int64_t* offsets );
/* keep_segments > 0: additionally record up to keep_segments segments per read (parity tests of BinarySeeding). */
int ma_b200_align_run( ma_b200_ctx* ctx, int32_t upto_stage, int32_t keep_segments, ma_b200_align_stats* stats );
int ma_b200_align_download_info( ma_b200_ctx* ctx, ma_b200_read_info* info );
int ma_b200_align_download_segments( ma_b200_ctx* ctx, ma_b200_segment* segs, int32_t* n_segs );
int ma_b200_align_download_seeds( ma_b200_ctx* ctx, ma_b200_seed* seeds, int64_t cap );
int ma_b200_align_download_sets( ma_b200_ctx* ctx, ma_b200_seed_set* sets, int64_t cap_sets, ma_b200_seed* seeds,
                                 int64_t cap_seeds );
int ma_b200_align_download( ma_b200_ctx* ctx, ma_b200_read_info* info, ma_b200_alignment* alns, int64_t cap_alns,
                            uint32_t* runs, int64_t cap_runs );
/* Optional pipelined form of ma_b200_align_batch: with reads_per_subbatch > 0, batches of at least 2 x that many reads
 * are cut into sub-batches that alternate between two sets of device slabs (two host threads, two streams), so that
 * the host<->device copies of one sub-batch run under the kernels of the other; results do not depend on the split.
 * Off (0) by default: on a PCIe Gen5 B200 box the copies of a 2 M read batch are 7 % of the step and the smaller
 * launches cost more than the overlap returns (bench.py --split); it pays when the link is slower or shared. */
int ma_b200_set_batch_split( ma_b200_ctx* ctx, int64_t reads_per_subbatch );

/* Reported-only output: with on != 0, ma_b200_align_download / ma_b200_align_batch after MA_B200_STAGE_MAPQ deliver only
 * the records MappingQuality (and PairedReads) return — rank_mq >= 0, i.e. what the reference's writers consume — packed
 * per read (info[i].set_off / n_sets index them; stats.n_reported of them in all; the run words stay complete). A
 * human-sized genome gives ~3 seed sets but ~1.1 reported alignments per Illumina read: the record download shrinks 3x.
 * Off by default: the staged downloads and the parity tests see every alignment NeedlemanWunsch computed. */
int ma_b200_set_reported_only( ma_b200_ctx* ctx, int32_t on );

/* One call, host buffers in and out (upload + all stages + download): the drop-in for a batch of reads. */
int ma_b200_align_batch( ma_b200_ctx* ctx, int64_t n_reads, const uint8_t* reads, const int64_t* offsets,
                         ma_b200_read_info* info, ma_b200_alignment* alns, int64_t cap_alns, uint32_t* runs,
                         int64_t cap_runs, ma_b200_align_stats* stats );

/* PairedReads::execute (pairedReads.cpp:15-121) for ONE pair whose records are on the host: for graphs that run a host
 * module between MappingQuality and PairedReads (SmallInversions adds records, export.cpp:176-184). Runs the same
 * routine as the device stage (ma_b200/csrc/mapq.cuh paired_reads_pair, compiled for the host) on the records of the two
 * mates — entries with rank_mq >= 0 take part, in that order — and sets pair_rank / flags / mapq on them in place.
 * runs: the run words the records' run_off / n_runs point into. Returns the size of the module's result vector, or a
 * negative MA_B200_E* code (MA_B200_EINVAL also where the reference itself would index an empty vector). */
int ma_b200_paired_reads_host( const ma_b200_params* params, int64_t ref_len, ma_b200_alignment* mate1, int32_t n1,
                               int64_t qlen1, ma_b200_alignment* mate2, int32_t n2, int64_t qlen2, const uint32_t* runs );

/* ---- measurement helper ----------------------------------------------------------------------------------- */
/* Measured bandwidth (GB/s) of independent random 64-byte block reads over a buffer of buffer_bytes: the roofline
 * of the seeding kernels (two such reads per extend_backward), SURVEY.md §8(d). */
int ma_b200_gather_probe( ma_b200_ctx* ctx, int64_t buffer_bytes, double* gbs );

#ifdef __cplusplus
}
#endif
#endif
