/*
 * swb.h -- C ABI of the B200-native Smith-Waterman database-scan engine (libswb.so).
 *
 * Drop-in boundary for the hot path of MattAgostini/ECE1782-Smith-Waterman-CUDA:
 *
 *     void smith_waterman_cuda(FASTAQuery&, FASTADatabase&, std::vector<seqid_score>&)
 *         reference: src/SWSolver.h:9, implemented in src/SWSolver.cu:266-404
 *
 * The reference does everything inside that one call (encode, pack, upload, launch, gather). Behind
 * this ABI the same work is split into the pieces a binding needs:
 *
 *   reference step (file:line)                               entry point here
 *   -------------------------------------------------------  ---------------------------------
 *   residue encoding  convertStringToFloat  SWSolver.cu:91-120   swb_encode
 *   blosum50[25][25] + GAP_PENALTY          SWSolver.cu:7,54-81  swb_scoring_matrix / swb_set_scoring*
 *   (+3/-3 scheme of the CPU solver         cpu.cpp:6-8,57-59)   SWB_SCORING_IDENT3
 *   host packing loop + managed upload      SWSolver.cu:301-359  swb_db_load   (once per database)
 *   query upload to constQuery              SWSolver.cu:291-298  swb_search / swb_search_batch
 *   f_scoreSequenceTiledCoalesced launches  SWSolver.cu:201-264, 346, 379      "
 *   result gather                           SWSolver.cu:383-390  scores[] in database order; swb_search_batch_topk
 *   the whole call on every GPU of the box  main.cpp:52-56       swb_group_* (one process, one engine per GPU)
 *   traceback of a pair (CPU solver)        cpu.cpp:39-108       swb_align (one hit of a scan, on the GPU)
 *   text parsing of the database            FASTAParsers.h:73-136 swb_read_fasta / swb_dbfile_* (optional fast path)
 *
 * Plain pointers and sizes only; all buffers are caller-owned host memory unless stated otherwise.
 * Every function returns SWB_OK (0) or a negative error code; swb_last_error() gives the text.
 * Scores are the exact int32 values of the recurrence (the reference stores `short`, SWSolver.cu:263).
 * There is no CPU fallback: without a CUDA device swb_create fails.
 *
 * Residue codes: one byte per residue, 0..31. 0..23 = ARNDCQEGHILKMFPSTWYVBJZX, 24 = '*'/unknown,
 * 25..30 free for other alphabets, 31 = padding (its matrix row and column are forced to zero).
 */
#ifndef SWB_H
#define SWB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWB_OK 0
#define SWB_ERR_ARG (-1)    /* bad argument */
#define SWB_ERR_CUDA (-2)   /* a CUDA call failed (text in swb_last_error) */
#define SWB_ERR_STATE (-3)  /* call out of order (e.g. search before db_load) */
#define SWB_ERR_NOMEM (-4)

#define SWB_SCORING_BLOSUM50_REF 0 /* SWSolver.cu:54-81, gap 2 */
#define SWB_SCORING_IDENT3 1       /* cpu.cpp:6-8, gap 2 */

typedef struct swb_engine swb_engine;

typedef struct swb_stats_t {
    double device_ms;        /* CUDA-event time of the last search call: first kernel -> last result */
    double load_ms;          /* wall time of the last swb_db_load */
    uint64_t cells;          /* true cells (query length x shard residues) of the last search call */
    uint64_t padded_cells;   /* cells actually computed, including row/column padding */
    uint64_t db_residues;    /* residues of this shard */
    uint64_t db_residues_total;
    uint32_t db_sequences;   /* sequences of this shard */
    uint32_t tiles;          /* warp tiles of this shard */
    uint32_t tiles_by_group[6]; /* tiles with 1,2,4,8,16,32 lanes per sequence pair */
    uint32_t recomputed_tiles;  /* tiles re-scored in int32 during the last search call */
    uint32_t kernel_launches;   /* kernels launched by the last search call */
    uint32_t last_k;            /* query rows per lane used by the last score kernel */
    uint32_t sm_count;
    uint32_t pack_us;           /* device time of the pack kernel of the last swb_db_load, microseconds (CUDA events) */
} swb_stats_t;

/* ---- engine ------------------------------------------------------------------------------ */
int swb_create(swb_engine **out, int device);
void swb_destroy(swb_engine *e);
/* e == NULL: error text of the last failed swb_create of this thread */
const char *swb_last_error(const swb_engine *e);
/* options: "group_len" (longest sequence handled by one lane per pair; set before db_load. Default 0 = chosen per
 *          load from the shard size: 1536 for ~450 k sequences and more, 768 down to ~110 k, 384 below),
 *          "k" (query rows per lane: 0 = chosen per lane-group size and query, else 8, 16, 32),
 *          "streams" (queries of a batch in flight at once, 1..24, default 16; their scratch is allocated on first use),
 *          "group_order" (0 = auto, 1 = launch the long-sequence tiles first, 2 = launch the bulk first),
 *          "batch_order" (0 = a batch runs its longest query first (default), 1 = in the caller's order; the
 *          results are always in the caller's order),
 *          "split" (1 = the passes of sequences longer than "xl_len" run as pipelined work items on
 *          different warps, 0 = never, -1 (default) = only on small shards, where those few tiles are the critical
 *          path of a query: +10 % at 1/8 of Swiss-Prot per GPU; on a large shard it costs ~0.5 %),
 *          "xl_len" (lane-group tiles wider than this are the ones "split" applies to, default 3072; before db_load),
 *          "split_k" (query rows per lane of the pipelined groups: 0 = chosen per query, 8, 16, 32),
 *          "exact" (scores beyond the s16 range: 0 (default) = the rebased s16 policy where the scoring scheme allows it
 *          -- steps between neighbouring cells small enough for a 16-bit window -- else int32; 1 = always int32),
 *          "direct_len" (pipelined tiles at least this wide, against a query at least this long, skip the plain s16
 *          pass and are scored by the rebased policy at once; default 16000, 0 = never),
 *          "load_threads" (host threads that gather the residues of a sharded load, default 4),
 *          "static_wave" (1 = the first work item of every warp of a bulk launch is taken by position, so the warps of a
 *          block start on tiles of similar length and the block leaves the SM together; 0 = all items from the shared
 *          counter in arrival order; -1 (default) = by position when the launch has fewer than two tiles per warp),
 *          "chunk_rows" (query rows per launch for queries beyond shared memory; multiple of 1024, <= 7168) */
int swb_set_option(swb_engine *e, const char *key, int64_t value);
/* run on the caller's CUDA stream (cudaStream_t as void*); NULL = the engine's own stream */
int swb_set_stream(swb_engine *e, void *cuda_stream);

/* ---- scoring (replaces blosum50 / GAP_PENALTY / convertStringToFloat) ---------------------- */
/* matrix: alpha x alpha int8, row-major, alpha <= 32; S + gap must fit int8; gap >= 0 (linear gap) */
int swb_set_scoring(swb_engine *e, const int8_t *matrix, int alpha, int gap);
int swb_set_scoring_preset(swb_engine *e, int preset);
/* Affine gaps (Gotoh): a gap of length L costs gap_open + (L-1) * gap_extend, 0 <= gap_extend <= gap_open <= 64.
 * The reference only has the linear model ("define affine penalty ?", SWSolver.cu:8); gap_open == gap_extend is
 * that model and runs the same kernels as swb_set_scoring. swb_align / swb_align_batch follow the model that is set:
 * under affine gaps they walk Gotoh's three states (H sources in cpu.cpp's order LEFT, TOP, DIAG with strict '>', a gap
 * prefers to open over to extend on a tie); with gap_open == gap_extend that is exactly the linear walk. */
int swb_set_scoring_affine(swb_engine *e, const int8_t *matrix, int alpha, int gap_open, int gap_extend);
/* host helpers, usable without a GPU: the 32 x 32 preset matrix and the preset's char -> code map */
int swb_scoring_matrix(int preset, int8_t *matrix32x32, int *gap);
int swb_encode(int preset, const char *text, size_t n, uint8_t *codes);

/* ---- database (replaces the packing loop, SWSolver.cu:301-359) ----------------------------- */
/* codes: concatenated residue codes; offsets: n+1 entries (64-bit, cf. FASTAParsers.h:69-71 overflow).
 * shard / nshards: this engine keeps the shard-th of nshards residue-balanced parts (0,1 = everything).
 * The packed database stays resident on the GPU until the next swb_db_load or swb_destroy. */
int swb_db_load(swb_engine *e, const uint8_t *codes, const uint64_t *offsets, uint32_t n, uint32_t shard,
                uint32_t nshards);
uint32_t swb_db_count(const swb_engine *e);      /* sequences of this shard */
int swb_db_ids(const swb_engine *e, uint32_t *ids); /* their database ids, ascending = order of scores[] */

/* ---- search (replaces the kernel launches + gather, SWSolver.cu:346-390) ------------------- */
/* scores: swb_db_count() entries, scores[k] belongs to database id ids[k] (id k when nshards == 1) */
int swb_search(swb_engine *e, const uint8_t *query, uint32_t qlen, int32_t *scores);
/* nq queries, concatenated codes + nq+1 offsets; scores: nq x swb_db_count(), or NULL to leave the
 * results on the device (read them later with swb_fetch_scores) */
int swb_search_batch(swb_engine *e, const uint8_t *qcodes, const uint64_t *qoffsets, uint32_t nq, int32_t *scores);
int swb_fetch_scores(swb_engine *e, uint32_t query_index, int32_t *scores);
/* The same scan, results written by DATABASE ID into vectors of the whole database: scores_full is nq x n_total
 * (n_total = the n of swb_db_load), query q, database id i -> scores_full[q * n_total + i]; this engine fills the ids
 * of its own shard and leaves the others alone, so the engines of all shards (one per GPU) can share one matrix. */
int swb_search_batch_scatter(swb_engine *e, const uint8_t *qcodes, const uint64_t *qoffsets, uint32_t nq,
                             int32_t *scores_full, uint64_t n_total);
/* The same scan with the hit list selected ON THE DEVICE: per query the k best sequences of this shard, score
 * descending, database id ascending on equal scores (k <= 1024): ids[q * k + j], top[q * k + j]; when the shard has
 * fewer than k sequences the rest is id 0xffffffff, score -1. Copies 8 k bytes per query back instead of 4 n (the
 * reference returns every score, SWSolver.cu:383-390); the full vectors stay on the device for swb_fetch_scores. */
int swb_search_batch_topk(swb_engine *e, const uint8_t *qcodes, const uint64_t *qoffsets, uint32_t nq, uint32_t k,
                          uint32_t *ids, int32_t *top);
/* k best (score desc, id asc) of one score vector of this shard that is already in HOST memory; ids are database ids */
int swb_topk(const swb_engine *e, const int32_t *scores, uint32_t k, uint32_t *ids, int32_t *top);
int swb_stats(const swb_engine *e, swb_stats_t *out);
/* Alignment with traceback of the query against ONE database sequence of this shard (the top hits of a scan) -- what
 * the reference's cpu.cpp prints for two strings (cpu.cpp:39-108): same update order LEFT, TOP, DIAG with strict '>',
 * first row-major maximum, walk back until H == 0. end_i / end_j: 1-based cell of the maximum. ops: the alignment from
 * its start to its end, one byte per column: 1 = gap in the query (consumes a subject residue, cpu.cpp FROM_LEFT),
 * 2 = gap in the subject (FROM_TOP), 3 = aligned pair (FROM_TOP_LEFT). cap >= qlen + subject length always suffices. */
int swb_align(swb_engine *e, const uint8_t *query, uint32_t qlen, uint32_t db_id, int32_t *score, uint32_t *end_i,
              uint32_t *end_j, uint8_t *ops, uint32_t cap, uint32_t *nops);
/* The same for a list of hits in ONE launch (one thread block per hit; e.g. the top-k lists of a batch): hit h aligns query
 * hit_query[h] of the batch (qcodes / qoffsets as in swb_search_batch) with database sequence hit_db_id[h] of this shard.
 * ops of hit h go to ops[ops_offsets[h] .. ops_offsets[h+1]) (nhits + 1 offsets; room of qlen + subject length always
 * suffices); ops may be NULL (scores and end cells only). end_i, end_j, nops may be NULL. */
int swb_align_batch(swb_engine *e, const uint8_t *qcodes, const uint64_t *qoffsets, uint32_t nq, const uint32_t *hit_query,
                    const uint32_t *hit_db_id, uint32_t nhits, int32_t *scores, uint32_t *end_i, uint32_t *end_j,
                    uint8_t *ops, const uint64_t *ops_offsets, uint32_t *nops);

/* ---- engine group: every GPU of the box in one process (the reference is one process, main.cpp:52-56) --- */
/* One engine, one host worker thread and one stream set per device. The devices form P database parts x R query groups
 * (P * R = devices): device (p, r) keeps part p of the residue-balanced sharding resident and scores the queries of
 * group r of a batch (groups of equal total length, longest-processing-time first). P = the largest divisor of the
 * device count whose parts keep at least "min_part_sequences" sequences (default SWB_MIN_PART_SEQUENCES: measured on
 * fractions of the benchmark database a B200 keeps its full rate down to about that size and loses 3 / 5 / 9 % at
 * 1/2, 1/4, 1/8 of it, while a batch of a quarter of the 20 reference queries against the whole database loses 4 %);
 * option "db_parts" forces it (= devices: pure database sharding). */
#define SWB_MIN_PART_SEQUENCES 450000u
typedef struct swb_group swb_group;
/* devices == NULL: devices 0 .. ndev-1; ndev == 0: every visible device */
int swb_group_create(swb_group **out, const int *devices, int ndev);
/* the same from the environment: SWB_DEVICES=<i,j,...> (indices may repeat: several engines on one device), else
 * SWB_GPUS=<n> (the first n devices), else every visible device */
int swb_group_create_env(swb_group **out);
void swb_group_destroy(swb_group *g);
const char *swb_group_last_error(const swb_group *g); /* g == NULL: last failed swb_group_create of this thread */
int swb_group_size(const swb_group *g);
swb_engine *swb_group_engine(swb_group *g, int i); /* the engine of device i (stats, options) */
int swb_group_db_parts(const swb_group *g);        /* P of the loaded layout, 0 before swb_group_db_load */
/* "db_parts", "min_part_sequences", or any swb_set_option key (applied to every engine) */
int swb_group_set_option(swb_group *g, const char *key, int64_t value);
int swb_group_set_scoring(swb_group *g, const int8_t *matrix, int alpha, int gap);
int swb_group_set_scoring_preset(swb_group *g, int preset);
int swb_group_set_scoring_affine(swb_group *g, const int8_t *matrix, int alpha, int gap_open, int gap_extend);
/* one length sort for all parts; the devices load their parts concurrently */
int swb_group_db_load(swb_group *g, const uint8_t *codes, const uint64_t *offsets, uint32_t n);
/* scores: nq x n in database order, as swb_search_batch returns them on one GPU */
int swb_group_search_batch(swb_group *g, const uint8_t *qcodes, const uint64_t *qoffsets, uint32_t nq, int32_t *scores);
/* per-GPU device-side hit lists merged on the host: ids / top are nq x k, score descending, id ascending */
int swb_group_search_batch_topk(swb_group *g, const uint8_t *qcodes, const uint64_t *qoffsets, uint32_t nq, uint32_t k,
                                uint32_t *ids, int32_t *top);
/* totals over the devices of the last call (device_ms = the slowest device) */
int swb_group_stats(const swb_group *g, swb_stats_t *out);
/* the layout rules, CPU only: P for n sequences on ndev devices; group_of[q] for R groups of a batch (LPT) */
int swb_layout_parts(uint32_t n, int ndev, uint32_t min_part_sequences);
/* the same with the batch in view: P = ndev (one query group) when R = ndev / P query groups would hold fewer than 8
 * queries each or cannot be formed with (nearly) equal total length from this batch (heaviest group more than 2 % above
 * the mean); bench.py lays out its ranks with it */
int swb_layout_parts_batch(uint32_t n, int ndev, uint32_t min_part_sequences, const uint64_t *qoffsets, uint32_t nq);
int swb_layout_query_groups(const uint64_t *qoffsets, uint32_t nq, int groups, uint32_t *group_of);

/* ---- plan introspection, CPU only (host logic of swb_db_load) ------------------------------ */
typedef struct swb_plan_info_t {
    uint32_t n_total, n_local, tiles, max_len;
    uint64_t residues_local, residues_total, res_bytes, bnd_elems, padded_cols;
    uint32_t tiles_by_group[6];
} swb_plan_info_t;
/* sorted_ids / shard_ids may be NULL; otherwise n_local entries each (query n_local with NULLs first) */
int swb_plan_describe(const uint64_t *offsets, uint32_t n, uint32_t shard, uint32_t nshards, uint32_t group_len,
                      swb_plan_info_t *info, uint32_t *sorted_ids, uint32_t *shard_ids);

/* ---- database ingest and the encoded on-disk database, CPU only (the reference re-parses and ---- */
/* ---- re-packs text on every run, FASTAParsers.h:73-136, SWSolver.cu:301-359) ------------------------------- */
typedef struct swb_dbfile swb_dbfile;
/* Text -> encoded arrays (malloc'ed, release with swb_free). swb_read_fasta follows the record rules of the reference
 * parser (records at '>' lines; a file without '>' is one record whose reference id is -1, reported in *first_id)
 * without its '/' padding; swb_read_uniprot_dat takes the SQ blocks of a UniProt flat file (parse.py:24-35). */
int swb_read_fasta(const char *path, int preset, uint8_t **codes, uint64_t **offsets, uint32_t *n, int32_t *first_id);
int swb_read_uniprot_dat(const char *path, int preset, uint8_t **codes, uint64_t **offsets, uint32_t *n);
void swb_free(void *p);
/* One memory-mappable file: 32-byte header, n+1 64-bit offsets, the codes. The pointers of an open file stay valid
 * until swb_dbfile_close and can be passed straight to swb_db_load. */
int swb_dbfile_write(const char *path, const uint8_t *codes, const uint64_t *offsets, uint32_t n, int32_t first_id);
int swb_dbfile_open(const char *path, swb_dbfile **out);
uint32_t swb_dbfile_count(const swb_dbfile *d);
int32_t swb_dbfile_first_id(const swb_dbfile *d);
const uint64_t *swb_dbfile_offsets(const swb_dbfile *d);
const uint8_t *swb_dbfile_codes(const swb_dbfile *d);
void swb_dbfile_close(swb_dbfile *d);

/* ---- measurement support (not on the scoring path) ------------------------------------------- */
/* Issue rate of the integer SIMD instructions the score kernel is built from, whole GPU, in giga
 * lane-instructions/s. kind: 0 viaddmax.s16x2.relu, 1 vimax3.s16x2, 2 vadd2, 3 prmt, 4 a dependent-chain loop of the
 * round-1 per-cell mix (two viaddmax), 5 viaddmax+imad, 6 imad, 7 scalar add+max, 8 the mix of the (rejected) biased
 * FMA-pipe variant, 9 hmnmx2 (+ a mask), 10 viaddmax+hmnmx2 (18.3 T: the fp16 comparator shares the ALU pipe, so it
 * cannot take over the max), 11 viaddmax+vadd2, 12 viaddmax+prmt, 13 viaddmax+vimax3 in one loop, 14 a dependent-chain
 * loop of the present per-cell mix.
 * Roofline accounting (bench.py): the peak of the score kernel is the ALU-pipe issue rate (kind 0: 64 lanes/clk/SM)
 * divided by the ALU-pipe instructions per cell. Per cell PAIR the V16 kernel issues prmt + vimax3.relu + 1/2 vimax3
 * on that pipe (2.5) plus two vadd2; kinds 11..13 decide where the vadd2 goes: viaddmax+vadd2 runs at twice the single
 * rate (vadd2 issues on another pipe -> 2.5 per pair), the other pairs at the single rate (same pipe). If kind 11 did
 * not double, the count would be 4.5. (Until r2x the cell was prmt + viaddmax.relu + viaddmax + 1/2 vimax3 + one vadd2:
 * 3.5 on the ALU pipe; bench.py prints the fraction against that ceiling too.) */
int swb_microbench(int device, int kind, double *glane_instr_per_s, double *ms);
/* Re-runs the pack kernel of the loaded database `reps` times (same inputs, same output) and reports the mean device time
 * per launch in microseconds and the bytes one launch reads + writes: the HBM figure of the one bandwidth-bound kernel of
 * the path at steady clocks. (swb_stats_t::pack_us is the same kernel inside swb_db_load, right after an upload during
 * which the SMs were idle.) */
int swb_pack_time(swb_engine *e, int reps, double *us_per_launch, uint64_t *bytes_per_launch);

#ifdef __cplusplus
}
#endif
#endif /* SWB_H */
