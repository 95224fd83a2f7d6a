/* mbfbam.h -- C ABI of the B200 read-counting library (libmbfbam_cuda.so).
 *
 * This is the boundary a host extension binds (the reference's host is a
 * Rust/PyO3 cdylib; here the host is mbf_bam_b200/csrc/pymod.cpp because no Rust
 * toolchain exists in this image -- INTEGRATION.md shows the Rust `extern "C"`
 * block that binds the same symbols).  Plain C: pointers, sizes, PODs; no C++
 * or torch types; bindgen/cbindgen friendly.
 *
 * Every entry point names the reference interface it replaces
 * (paths relative to the reference tree, MarcoMernberger/mbf_bam).
 *
 * Error convention: functions returning int give 0 on success or a negative
 * code; a NUL-terminated message is written to `err` (may be NULL).  The host
 * turns it into ValueError exactly like BamError -> ValueError in
 * src/lib.rs:26-32.
 */
#ifndef MBFBAM_H
#define MBFBAM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBFBAM_OK 0
#define MBFBAM_ERR_IO (-1)        /* "Could not read bam: ..." (src/bam_ext.rs:16-18) */
#define MBFBAM_ERR_FORMAT (-2)    /* corrupt BGZF / BAM / BAI */
#define MBFBAM_ERR_ARG (-3)       /* bad arguments (e.g. contig missing from `intervals`) */
#define MBFBAM_ERR_CUDA (-4)      /* CUDA runtime failure / no device */
#define MBFBAM_ERR_UNSUPPORTED (-5)

/* Counting modes. */
#define MBFBAM_MODE_UNSTRANDED 0  /* src/count_reads/counters.rs:26-106  */
#define MBFBAM_MODE_STRANDED 1    /* src/count_reads/counters.rs:114-202 */
#define MBFBAM_MODE_WEIGHTED 2    /* not in the reference; 1/NH weights */

/* Flattened form of the Python argument `{contig: [(gene_id, strand, [starts], [stops])]}`
 * that src/lib.rs:118-134 (py_intervals_to_trees) + src/count_reads/mod.rs:27-45
 * (build_tree) turn into one bio IntervalTree<u32,(u32,i8)> per contig.
 * Intervals of contig c are entries iv_offsets[c] .. iv_offsets[c+1]-1, in list
 * order (any order is accepted; the library sorts its own copy).  iv_gene is the
 * gene_no LOCAL to the contig (index in that contig's list, mod.rs:31);
 * n_genes[c] is the length of that list.  Gene-id strings never cross the ABI. */
typedef struct mbfbam_intervals {
    int32_t n_contigs;
    const char* const* contig_names; /* n_contigs NUL-terminated names            */
    const uint32_t* n_genes;         /* n_contigs                                  */
    const uint64_t* iv_offsets;      /* n_contigs + 1                              */
    const uint32_t* iv_start;        /* half-open [start, end)                     */
    const uint32_t* iv_end;
    const uint32_t* iv_gene;
    const int8_t* iv_strand;         /* +1 / -1 / 0 (0 behaves like -1, counters.rs:151) */
} mbfbam_intervals;

/* Opaque reader: replaces rust-htslib's bam::IndexedReader as opened by
 * open_bam() in src/bam_ext.rs:6-21 (header + BAI loaded once, not per chunk). */
typedef struct mbfbam_reader mbfbam_reader;

/* src/bam_ext.rs:6-21 open_bam(bam_filename, bai_filename).  bai_path == NULL
 * tries "<bam>.bai" then "<bam with .bam replaced by .bai>". */
int mbfbam_open(const char* bam_path, const char* bai_path, mbfbam_reader** out,
                char* err, size_t errlen);
void mbfbam_close(mbfbam_reader* r);

/* bam.header().target_names()/target_len()/tid() (src/count_reads/chunked_genome.rs:24,36-40,81-82) */
int32_t mbfbam_n_targets(const mbfbam_reader* r);
const char* mbfbam_target_name(const mbfbam_reader* r, int32_t tid);
uint32_t mbfbam_target_len(const mbfbam_reader* r, int32_t tid);
int32_t mbfbam_tid(const mbfbam_reader* r, const char* name); /* -1 if absent */
uint64_t mbfbam_file_size(const mbfbam_reader* r);

/* Optional: copy the compressed BGZF byte ranges of the whole file into HBM of
 * `device` once, so later count calls on that device start from device-resident
 * input (bench.py's `value`; e2e calls do not use it). */
int mbfbam_preload(mbfbam_reader* r, int32_t device, char* err, size_t errlen);
void mbfbam_drop_preload(mbfbam_reader* r);

/* Host side of the ingest (replaces the hFILE reads under htslib's bgzf_read_block).  The library keeps a
 * process-wide cache of private file mappings keyed by (device, inode, size, mtime); jobs copy out of the
 * mapping into pinned staging buffers with a few threads.  Once MBF_BAM_PIN_AFTER (default 3; "never")
 * jobs have streamed a file, the byte ranges a job touches are page-locked (cudaHostRegister, 64 MiB
 * segments, at most MBF_BAM_PIN_MAX_MB, default half of RAM) and later jobs DMA straight out of the page
 * cache.  mbfbam_pin_file page-locks the whole file now (MBFBAM_ERR_UNSUPPORTED when the platform refuses;
 * the reader keeps working through the staging path); mbfbam_unpin_file releases the page locks of this
 * file (no job may be running on it).  Two more process-wide caches keep repeated calls cheap: parsed indices, keyed
 * by the index file's identity (name, inode, size, mtime) and the header's targets, and count tables (chunk table,
 * sorted interval arrays), keyed by the content of the two interval arguments.  mbfbam_host_cache_clear drops all
 * three. */
int mbfbam_pin_file(mbfbam_reader* r, char* err, size_t errlen);
void mbfbam_unpin_file(mbfbam_reader* r);
void mbfbam_host_cache_clear(void);
/* The library keeps one pipeline (streams, device and pinned buffers, sized by the largest job so far) per
 * device between calls; this frees the one of `device` (it is rebuilt by the next job). */
void mbfbam_release_device(int32_t device);

/* Per-call measurements (all optional output; zeroed by the library first). */
typedef struct mbfbam_stats {
    uint64_t n_records;          /* BAM records framed (all refs in the scanned ranges) */
    uint64_t n_records_owned;    /* records that passed the chunk-ownership filter      */
    uint64_t n_bgzf_blocks;
    uint64_t compressed_bytes;   /* BGZF bytes fed to the GPU                           */
    uint64_t uncompressed_bytes;
    uint64_t h2d_bytes;
    uint64_t d2h_bytes;
    uint64_t n_chunks;           /* ChunkedGenome chunks in this shard                  */
    uint64_t n_kernel_launches;
    uint64_t n_windows;
    uint64_t n_mm_keys;          /* NH != 1 (chunk,gene,bucket,qname) keys emitted      */
    uint64_t n_fixups;           /* BGZF blocks whose first record was not block-aligned */
    double ms_total;             /* wall clock of the call (host)                       */
    double ms_gpu_scan;          /* sum over windows, CUDA events on the launch stream  */
    double ms_gpu_inflate;
    double ms_gpu_frame;
    double ms_gpu_count;
    double ms_gpu_other;
    double ms_gpu_all;           /* first copy/kernel start -> last kernel end (CUDA events) */
    double ms_gpu_tokens;        /* pass 1 of the two-pass inflate (part of ms_gpu_inflate) */
    uint64_t n_tokens;           /* DEFLATE symbols (literals + matches) decoded by pass 1  */
    uint64_t h2d_pinned_bytes;   /* part of h2d_bytes that went by DMA from the page-locked file mapping */
    uint64_t n_devices;          /* devices that worked on the call                     */
    uint64_t n_retries;          /* job restarts with smaller windows (inflate ratio jump) */
    double ms_pin;               /* time spent page-locking file segments in this call  */
} mbfbam_stats;
/* With several devices the additive fields are summed and the times are the maximum over devices. */

typedef struct mbfbam_count_job {
    const mbfbam_intervals* intervals;      /* overlap targets (`intervals` argument)        */
    const mbfbam_intervals* gene_intervals; /* chunk edges only (`gene_intervals` argument)  */
    int32_t mode;                           /* MBFBAM_MODE_*                                  */
    int32_t each_read_counts_once;          /* lib.rs:143,152 (default false)                 */
    int32_t device;                         /* CUDA ordinal (used when n_devices == 0)        */
    int32_t shard_index;                    /* this call owns chunks of shard i of n ...      */
    int32_t shard_count;                    /* ... (1 => everything); see DESIGN.md (e)       */
    int32_t n_devices;                      /* > 0: fan out over devices[0..n_devices) from this one call:    */
    const int32_t* devices;                 /* the shard is cut into n_devices chunk ranges, one pipeline per  */
                                            /* device on its own thread, dense results summed on the host --   */
                                            /* the rayon pool of counters.rs:217-259 at GPU granularity        */
} mbfbam_count_job;

/* Results are dense and caller-allocated; gene slot = gene_base[c] + gene_no where
 * gene_base is the exclusive prefix sum of intervals->n_genes.  n_total = sum n_genes.
 *   fwd/rev : u64[n_total]   (rev may be NULL for unstranded; weighted leaves them 0)
 *   wfwd/wrev: f64[n_total]  (weighted mode only, may be NULL otherwise)
 *   scanned : u8[intervals->n_contigs] 1 if the contig was scanned (in gene_intervals ∩ header)
 *   outside : reads in an owned chunk that hit no interval (counters.rs:95-97)
 * each_read_counts_once != 0 (counters.rs:83-92,173-182): a read counts only for the first hit of its first
 * block that has one, "first" in the iteration order of bio's IntervalTree (restated in reader.hpp). */
typedef struct mbfbam_counts {
    uint64_t* fwd;
    uint64_t* rev;
    double* wfwd;
    double* wrev;
    uint8_t* scanned;
    uint64_t outside;
} mbfbam_counts;

/* Replaces py_count_reads_unstranded / py_count_reads_stranded
 * (src/count_reads/counters.rs:204-260, :276-323) minus the String-keyed map
 * assembly, which stays in the host extension. */
int mbfbam_count_reads(mbfbam_reader* r, const mbfbam_count_job* job, mbfbam_counts* out,
                       mbfbam_stats* stats, char* err, size_t errlen);

/* Replaces py_count_introns (src/count_reads/introns.rs:56-88).  Library-allocated
 * result: n entries (tid, start, stop, count), including the sentinel keys
 * (u32::MAX, u32::MAX - n_introns) of introns.rs:46-48; `contig_seen[tid]` = 1 for
 * every header contig that gets a (possibly empty) inner dict. */
typedef struct mbfbam_introns {
    uint64_t n;
    int32_t* tid;
    uint32_t* start;
    uint32_t* stop;
    uint64_t* count;
    int32_t n_targets;
    uint8_t* contig_seen;
} mbfbam_introns;

int mbfbam_count_introns(mbfbam_reader* r, int32_t device, int32_t shard_index,
                         int32_t shard_count, mbfbam_introns* out, mbfbam_stats* stats,
                         char* err, size_t errlen);
/* Same over several devices from one call (introns.rs:66-87 is a rayon map/reduce over chunks);
 * every (tid, start, stop) appears once in the result. */
int mbfbam_count_introns_multi(mbfbam_reader* r, const int32_t* devices, int32_t n_devices,
                               int32_t shard_index, int32_t shard_count, mbfbam_introns* out,
                               mbfbam_stats* stats, char* err, size_t errlen);
void mbfbam_free_introns(mbfbam_introns* x);

/* Replaces py_quantify_gene_reads (src/count_reads/quantify.rs:20-75, :98-146; wrapper src/lib.rs:247-264): per gene the
 * number of reads starting at each position, split into reads congruent (bucket 0) and incongruent (bucket 1)
 * with the gene's strand; a read counts once per gene, no NH logic.  job->mode and each_read_counts_once are
 * ignored.  Library-allocated result: n entries (global gene slot as in mbfbam_counts, bucket, pos, count) in no
 * particular order; scanned[c] has the meaning it has in the mbfbam_counts struct: every gene of a scanned contig gets
 * a -- possibly empty -- list in the Python result, quantify.rs:148-153. */
typedef struct mbfbam_gene_pos_counts {
    uint64_t n;
    uint32_t* gene;
    uint32_t* pos;
    uint32_t* count;
    uint8_t* bucket;
    uint8_t* scanned;
} mbfbam_gene_pos_counts;
int mbfbam_quantify_gene_reads(mbfbam_reader* r, const mbfbam_count_job* job, mbfbam_gene_pos_counts* out,
                               mbfbam_stats* stats, char* err, size_t errlen);
void mbfbam_free_gene_pos_counts(mbfbam_gene_pos_counts* x);

/* Replaces py_calculate_duplicate_distribution (src/duplicate_distribution.rs:44-86; wrapper src/lib.rs:108-116):
 * {k: how many distinct (reference, position, strand, sequence) occur exactly k times}.  Reads are keyed by a
 * 96-bit fingerprint of (tid, pos, is_reverse, SEQ) instead of the byte string (deviations). */
typedef struct mbfbam_dup_hist {
    uint64_t n;
    uint32_t* k;
    uint64_t* count;
} mbfbam_dup_hist;
int mbfbam_duplicate_distribution(mbfbam_reader* r, const int32_t* devices, int32_t n_devices, int32_t shard_index,
                                  int32_t shard_count, mbfbam_dup_hist* out, mbfbam_stats* stats, char* err,
                                  size_t errlen);
void mbfbam_free_dup_hist(mbfbam_dup_hist* x);

/* Replaces py_count_reads_primary_only_right_strand_only_by_barcode (src/count_reads/by_barcode.rs:18-92, :101-182;
 * wrapper src/lib.rs:184-211), umi_strategy "straight": per chunk and (gene, XC barcode) the number of distinct XM UMIs,
 * summed over (pos, is_reverse), of the reads that are not secondary and lie on the gene's strand.  One entry per
 * (chunk, gene, barcode), sorted by chunk, gene slot, barcode bytes; a chunk in which a right-strand hit lacks a string
 * XC or XM tag contributes nothing, as in the reference.  Keys are 128-bit fingerprints on the device (DESIGN.md,
 * deviations); barcodes longer than 44 bytes are an error.  job->mode and each_read_counts_once are ignored. */
typedef struct mbfbam_barcode_counts {
    uint64_t n;
    uint32_t* gene;      /* global gene slot as in the mbfbam_counts struct */
    uint32_t* chunk;     /* index into the chunk table of job->gene_intervals */
    uint32_t* count;
    uint64_t* bc_off;    /* n + 1 offsets into bc_bytes */
    char* bc_bytes;
} mbfbam_barcode_counts;
int mbfbam_count_by_barcode(mbfbam_reader* r, const mbfbam_count_job* job, mbfbam_barcode_counts* out, mbfbam_stats* stats,
                            char* err, size_t errlen);
void mbfbam_free_barcode_counts(mbfbam_barcode_counts* x);

/* ChunkedGenome (src/count_reads/chunked_genome.rs:17-43,72-119): writes up to `cap`
 * chunks (tid,start,stop) and returns the total count.  gene_intervals == NULL gives
 * new_without_tree().  Exposed for tests and for sharding decisions made by the host. */
int64_t mbfbam_chunks(const mbfbam_reader* r, const mbfbam_intervals* gene_intervals,
                      int32_t* tid, uint32_t* start, uint32_t* stop, int64_t cap);

/* Shard planning (host only, no GPU needed): the owned chunk range [*chunk_lo, *chunk_hi) of shard
 * `shard_index` of `shard_count` over the ChunkedGenome table of `gene_intervals` (NULL = all header
 * contigs), and the BGZF byte ranges [cbeg[i], cend[i]) that can hold its records.  Returns the number
 * of byte ranges (up to `cap` are written) or -1.  See DESIGN.md (e). */
int64_t mbfbam_plan_shard(const mbfbam_reader* r, const mbfbam_intervals* gene_intervals, int32_t shard_index,
                          int32_t shard_count, int64_t* chunk_lo, int64_t* chunk_hi, uint64_t* cbeg,
                          uint64_t* cend, int64_t cap);

/* Stage-level entry points, used by the parity tests (and usable on their own). */
/* Inflate every BGZF member of `bgzf` (n bytes, host memory) on `device`; `out` must hold
 * the sum of ISIZEs (returned through *out_len; query with out == NULL). */
int mbfbam_inflate_buffer(const uint8_t* bgzf, uint64_t n, uint8_t* out, uint64_t out_cap,
                          uint64_t* out_len, int32_t device, mbfbam_stats* stats, char* err,
                          size_t errlen);
/* CIGAR block expander on the device (replaces BamRecordExtensions::{blocks,introns},
 * src/bam_ext.rs:22-42): n_ops packed BAM cigar words at `pos`; writes (start,stop) pairs. */
int mbfbam_cigar_expand(const uint32_t* cigar, int32_t n_ops, int32_t pos, int32_t want_introns,
                        uint32_t* out_pairs, int32_t cap_pairs, int32_t* n_pairs, int32_t device,
                        char* err, size_t errlen);

const char* mbfbam_version(void);
int32_t mbfbam_device_count(void);

#ifdef __cplusplus
}
#endif
#endif /* MBFBAM_H */
