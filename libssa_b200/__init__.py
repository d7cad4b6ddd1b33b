"""libssa_b200 - B200-native (sm_100a) implementation of libssa's database-search hot path.

The product is the C-ABI shared library ``libssa_b200/libssa_b200.so`` (sources in ``csrc/``), which
exports the unchanged ``libssa.h`` API (include/libssa.h) plus a few additive ``ssa_b200_*`` calls
(include/libssa_b200.h).  This package is the thin Python host-side mirror of that interface, used by
the tests and by bench.py.  There is no CPU fallback: importing works anywhere, but every search
needs the CUDA library and a GPU and fails loudly otherwise.
"""
from .api import (  # noqa: F401
    LibSSA, Hit, Stats, load, library_path, build_library, LibraryMissing,
    NUCLEOTIDE, AMINOACID, TRANS_QUERY, TRANS_DB, TRANS_BOTH,
    FORWARD_STRAND, COMPLEMENTARY_STRAND, BOTH_STRANDS,
    OUTPUT_SILENT, OUTPUT_ERROR, OUTPUT_WARNING, OUTPUT_INFO,
    COMPUTE_SCORE, COMPUTE_ALIGNMENT, READ_FROM_FILE, READ_FROM_STRING, MATRIX_BUILDIN,
    BIT_WIDTH_8, BIT_WIDTH_16, BIT_WIDTH_64,
)

__version__ = "0.1.0"
