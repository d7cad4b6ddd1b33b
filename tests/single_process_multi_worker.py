"""Worker of tests/test_gpu_multi.py: ONE process driving several GPUs (LIBSSA_B200_GPUS / ssa_b200_set_device_count) - the
library shards the database over its devices, one host thread per device, ncclAllGather inside the library."""
import json
import os
import sys

import cases
import libssa_b200 as L


def main():
    expected = json.load(open(sys.argv[1]))
    ngpus = int(sys.argv[2])
    lib = L.load()
    if os.environ.get("LIBSSA_B200_GPUS") is None:
        lib.set_device_count(ngpus)          # the API route; the environment route is exercised by the C host test
    for entry in expected:
        got = cases.run_case(lib, entry["case"])
        for run, g, e in zip(entry["case"]["runs"], got, entry["result"]):
            assert g == e, (entry["case"]["name"], run, g[:3], e[:3])
        assert lib.device_count() == ngpus
        st = lib.stats()
        assert st.local_seqs == st.n_seqs          # the process holds every record, spread over its devices
    # the sparse replay, batched queries and the multi-frame searches on the multi-device path
    big_entry = [e for e in expected if e["case"]["name"] == "synth_protein_30000"][0]
    os.environ["LIBSSA_B200_REPLAY_PREFIX"] = "500"
    try:
        got = cases.run_case(lib, big_entry["case"])
    finally:
        del os.environ["LIBSSA_B200_REPLAY_PREFIX"]
    for run, g, e in zip(big_entry["case"]["runs"], got, big_entry["result"]):
        assert g == e, ("sparse replay", run, g[:3], e[:3])
    q0 = cases.configure(lib, big_entry["case"])
    ids, sc = lib.score_all(q0, False)
    assert len(ids) == 30000 and ids.tolist() == list(range(30000))
    lib.free_sequence(q0)
    from libssa_b200 import synth
    import numpy as np
    rng = np.random.default_rng(9)
    qs = [lib.init_sequence_fasta(cases.READ_FROM_STRING, synth.random_protein(rng, n)) for n in (120, 380, 45, 700, 1500)]
    for nw, bits, k, align in ((False, 8, 50, 0), (True, 16, 10, 1)):
        single = [[(h.score, h.db_id, h.cigar) for h in (lib.nw_align if nw else lib.sw_align)(q, k, bits, align)] for q in qs]
        batch = [[(h.score, h.db_id, h.cigar) for h in hits] for hits in lib.align_batch(qs, k, bits, align, nw=nw)]
        assert batch == single, (nw, bits)
    for q in qs:
        lib.free_sequence(q)
    import tempfile
    import frames as F
    from oracle import port
    with tempfile.TemporaryDirectory() as tmp:
        for mode in ((F.NUCLEOTIDE, F.BOTH), (F.TRANS_QUERY, F.BOTH), (F.TRANS_DB, F.FORWARD)):
            F.check_product_mode(lib, port, mode, tmp)
    lib.ssa_exit()                                   # joins the device threads, destroys the communicators
    print("SINGLE_PROCESS_MULTI_OK %d GPUs" % ngpus)


if __name__ == "__main__":
    main()
