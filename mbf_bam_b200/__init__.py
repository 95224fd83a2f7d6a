"""mbf_bam_b200 -- B200-native (sm_100a) read counters with mbf_bam's Python API.

Drop-in for the counting functions of `mbf_bam` (reference src/mbf_bam/__init__.py:11
star-imports the native module; reference src/lib.rs:286-299 lists its functions):

    count_reads_unstranded(filename, index_filename, intervals, gene_intervals, each_read_counts_once=None)
    count_reads_stranded(...)      -> (forward_dict, reverse_dict)
    count_reads_weighted(...)      -> (forward_dict, reverse_dict) of 1/NH weighted floats (new)
    count_introns(filename, index_filename=None)
    quantify_gene_reads(filename, index_filename, intervals, gene_intervals)
                                   -> (congruent, incongruent) {gene_id: [(pos, n_reads)]}
    calculate_duplicate_distribution(filename, index_filename=None) -> {k: occurrences}
    count_reads_primary_only_right_strand_only_by_barcode(filename, index_filename, intervals, gene_intervals,
                                   umi_strategy) -> (genes, barcodes, [(gene_idx, barcode_idx, n_umis)])

The native module is built in-tree by `__graft_entry__.build()` (or `make -C mbf_bam_b200/csrc`).
There is no CPU implementation behind these names: importing works anywhere, calling needs a GPU.  One call uses
every visible GPU (`MBF_BAM_DEVICES`, `Reader.set_devices`); `mbf_bam_b200.dist` is the one-process-per-GPU form.

Where results can differ from the reference (DESIGN.md section 4 has the full list):
  * counts are u64, the reference's are wrapping u32;
  * keys the reference compares as byte strings (multi-mapper read names, duplicate-distribution sequences, cell
    barcodes and UMIs) are compared as 96/128-bit fingerprints: a false merge needs a hash collision (< 1e-13 for 1e8
    keys);
  * errors the reference swallows or panics on (corrupt BGZF/DEFLATE data, malformed records, an NH tag that is not an
    integer, a contig of `gene_intervals` missing from `intervals`) raise ValueError; CRC32 of BGZF members is not
    checked;
  * lists and index assignments the reference returns in HashMap order (quantify_gene_reads,
    count_reads_primary_only_right_strand_only_by_barcode) come back sorted;
  * count_reads_weighted does not exist in the reference: every record adds 1/NH once per distinct (gene, strand
    bucket) it hits.
"""
try:
    from .mbf_bam import (  # noqa: F401
        Reader,
        __version__,
        calculate_duplicate_distribution,
        cigar_expand,
        count_introns,
        count_reads_primary_only_right_strand_only_by_barcode,
        count_reads_stranded,
        count_reads_unstranded,
        count_reads_weighted,
        device_count,
        host_cache_clear,
        release_device,
        inflate_buffer,
        last_stats,
        quantify_gene_reads,
    )
except ImportError as e:  # pragma: no cover
    raise ImportError(
        "mbf_bam_b200: the native extension is not built (run `python -c 'import __graft_entry__ as g; "
        "g.build()'` or `make -C mbf_bam_b200/csrc`): %s" % (e,)
    ) from e

MODE_UNSTRANDED, MODE_STRANDED, MODE_WEIGHTED = 0, 1, 2
