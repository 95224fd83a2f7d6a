"""Compatibility shim: `import mbf_bam` resolves to the B200 implementation, mirroring
reference src/mbf_bam/__init__.py:11 (`from .mbf_bam import *`)."""
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
