"""GPU parity of the searches that multiply the query and/or the database into strands and reading frames
(NUCLEOTIDE on the complementary / both strands, TRANS_QUERY, TRANS_DB, TRANS_BOTH) against the committed
fixture tests/golden/frames.json, which tests/golden/make_frames_golden.py generated from the compiled reference
(and tests/test_host.py re-checks against a live reference run where that is available).

For TRANS_DB / TRANS_BOTH the reference crashes while it builds the result records (us_prepare_sequence reallocs
an uninitialised pointer, src/util/util_sequence.c:385,408), so for those two the fixture holds the hit list as it
leaves the reference's search phase and the alignment is checked against the oracle instead."""
import pytest

import frames as F


@pytest.mark.gpu
@pytest.mark.parametrize("mode", F.MODES, ids=F.MODE_IDS)
def test_strand_and_frame_searches_match_reference(prod, port, tmp_path, mode):
    F.check_product_mode(prod, port, mode, tmp_path)


TRANSLATE_QUERY = "ATGCCCAAAATTGACTTGAATAGGGGCGTGAGGAGGGGTTTTCAATACTATTTTTTG"


@pytest.mark.gpu
@pytest.mark.parametrize("bits", [64, 16, 8])
def test_reference_translated_query_vectors(prod, bits):
    """Known answers of the reference's own searcher tests (tests/algo/test_searcher.c:247-325): a 57-nt query, all
    six frames (genetic code 3) against short_AA.fas, match/mismatch 1/-1, gaps -1/-1, six hits."""
    import os
    import cases
    prod.set_output_mode(0)
    prod.init_symbol_translation(F.TRANS_QUERY, F.BOTH, 3, 3)
    prod.init_constant_scores(1, -1)
    prod.init_gap_penalties(-1, -1)
    prod.set_chunk_size(6)
    prod.init_db(os.path.join(cases.DATA, "short_AA.fas"))
    q = prod.init_sequence_fasta(cases.READ_FROM_STRING, TRANSLATE_QUERY)
    try:
        nw = prod.nw_align(q, 6, bits, cases.COMPUTE_SCORE)
        sw = prod.sw_align(q, 6, bits, cases.COMPUTE_SCORE)
        assert [(h.score, h.db_id) for h in nw] == [(-109, 0), (-109, 0), (-110, 0), (-110, 0), (-110, 0), (-113, 0)]
        assert [(h.score, h.db_id) for h in sw] == [(2, 0), (2, 0), (2, 0), (2, 0), (2, 0), (1, 0)]
        assert sorted((h.q_strand, h.q_frame) for h in sw) == [(s, f) for s in range(2) for f in range(3)]
    finally:
        prod.free_sequence(q)
        prod.set_chunk_size(1000)
        prod.init_symbol_translation(cases.AMINOACID, 1, 1, 1)
