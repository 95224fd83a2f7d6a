"""Compatibility shim: `import mbf_bam` resolves to the B200 implementation, mirroring
reference src/mbf_bam/__init__.py:11 (`from .mbf_bam import *`).

The reference's native module exports eight functions (reference src/lib.rs:288-295).  The six read counters are here
(plus count_reads_weighted); the two BAM *writers* are not part of the read-counting path this package replaces and say
so when called."""
from mbf_bam_b200 import (  # noqa: F401
    __version__,
    calculate_duplicate_distribution,
    count_introns,
    count_reads_primary_only_right_strand_only_by_barcode,
    count_reads_stranded,
    count_reads_unstranded,
    count_reads_weighted,
    quantify_gene_reads,
)


def subtract_bam(output_filename, minuend_filename, subtrahend_filename):
    """reference src/lib.rs:229-243 -> src/bam_manipulation.rs: writes a BAM; not implemented here."""
    raise NotImplementedError("subtract_bam writes BAM files (needs a deflate path); mbf_bam_b200 covers the read counters only")


def annotate_barcodes_from_fastq(*args, **kwargs):
    """reference src/lib.rs:266-283 -> src/bam_manipulation.rs: writes a BAM; not implemented here."""
    raise NotImplementedError("annotate_barcodes_from_fastq writes BAM files; mbf_bam_b200 covers the read counters only")
