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
