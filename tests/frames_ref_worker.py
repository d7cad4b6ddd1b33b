"""Runs the compiled reference on one strand / reading-frame workload (tests/frames.py) and dumps what it
produced as JSON.  One process per mode: the reference's init_symbol_translation refuses every later call once
the symbol type is TRANS_DB or TRANS_BOTH (it validates the previous global, src/libssa.c:146-151), so those
modes cannot share a process with anything else.

usage: frames_ref_worker.py SYMTYPE STRANDS OUT.json [--entries]
"""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
import frames as F  # noqa: E402

CHUNK, K = 7, 25


def main():
    symtype, strands, out = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3]
    want_entries = "--entries" in sys.argv
    from oracle import refapi
    ref = refapi.reference()
    assert ref is not None, "oracle/_ref/libssa_ref.so is not built"
    ref.set_output_mode(0)
    recs, query = F.workload(symtype)
    res = {"symtype": symtype, "strands": strands, "chunk": CHUNK, "k": K}
    if want_entries:
        # everything that needs other symbol types first (see the module docstring)
        res["entries"] = [e.tolist() for e in F.entries_via_reference(ref, symtype, strands, recs)]
        if symtype == F.TRANS_DB:
            qf = F.query_frames(ref, F.AMINOACID, F.FORWARD, query)
        elif symtype == F.TRANS_BOTH:
            qf = F.query_frames(ref, F.TRANS_QUERY, strands, query)
        else:
            qf = F.query_frames(ref, symtype, strands, query)
        res["query_frames"] = [(c.tolist(), s, f) for c, s, f in qf]
    with tempfile.TemporaryDirectory() as tmp:
        db = os.path.join(tmp, "db.fas")
        with open(db, "w") as f:
            for i, r in enumerate(recs):
                f.write(">r%d\n%s\n" % (i, r))
        F.configure(ref, symtype, strands)
        ref.set_chunk_size(CHUNK)
        ref.init_db(db)
        q = ref.init_sequence_fasta(cases.READ_FROM_STRING, query)
        # "phase": the hit list as it leaves the search phase [(score, db_id, db_frame, db_strand, query_id)] plus the
        # overflow counters; "hits": the finished alignment records, for the modes whose result creation does not crash
        res["searches"] = {}
        for nw in (False, True):
            for bits in (64, 16, 8):
                phase, o8, o16 = refapi.search_phase(ref, q, K, bits, nw)
                rec = {"phase": phase, "overflow_8": o8, "overflow_16": o16}
                if symtype not in (F.TRANS_DB, F.TRANS_BOTH):
                    hits = (ref.nw_align if nw else ref.sw_align)(q, K, bits, cases.COMPUTE_ALIGNMENT)
                    rec["hits"] = [F.hit_record(h) for h in hits]
                    assert [(h.score, h.db_id, h.db_strand, h.db_frame) for h in hits] == [tuple(e[:4]) for e in phase]
                res["searches"]["%s%d" % ("nw" if nw else "sw", bits)] = rec
    with open(out, "w") as f:
        json.dump(res, f)


if __name__ == "__main__":
    main()
