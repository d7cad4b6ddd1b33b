"""Worker of tests/test_gpu_multi.py: one rank per GPU, sharded searches, every rank checks the merged result."""
import json
import os
import sys

import cases
import libssa_b200 as L
from libssa_b200 import dist as ldist


def main():
    expected = json.load(open(sys.argv[1]))
    lib = L.load()
    rank, world, local = ldist.init_from_env(lib)
    for entry in expected:
        got = cases.run_case(lib, entry["case"])
        for run, g, e in zip(entry["case"]["runs"], got, entry["result"]):
            assert g == e, (rank, entry["case"]["name"], run, g[:3], e[:3])
        st = lib.stats()
        assert st.local_seqs < st.n_seqs or st.n_seqs < world
    # the sparse form of the heap replay (what shards of more than 262 144 records in total use), forced on the 30 000-record case
    big_entry = [e for e in expected if e["case"]["name"] == "synth_protein_30000"][0]
    os.environ["LIBSSA_B200_REPLAY_PREFIX"] = "500"
    try:
        got = cases.run_case(lib, big_entry["case"])
    finally:
        del os.environ["LIBSSA_B200_REPLAY_PREFIX"]
    for run, g, e in zip(big_entry["case"]["runs"], got, big_entry["result"]):
        assert g == e, (rank, "sparse replay", run, g[:3], e[:3])
    # batched queries on the sharded database: the pipelined batch must equal single calls (which equal the reference)
    big = big_entry["case"]
    q0 = cases.configure(lib, big)
    lib.free_sequence(q0)
    from libssa_b200 import synth
    import numpy as np
    rng = np.random.default_rng(9)
    qs = [lib.init_sequence_fasta(cases.READ_FROM_STRING, synth.random_protein(rng, n)) for n in (120, 380, 45, 700)]
    for nw, bits, k, align in ((False, 8, 50, 0), (True, 16, 10, 1)):
        single = [[(h.score, h.db_id, h.cigar) for h in (lib.nw_align if nw else lib.sw_align)(q, k, bits, align)] for q in qs]
        batch = [[(h.score, h.db_id, h.cigar) for h in hits] for hits in lib.align_batch(qs, k, bits, align, nw=nw)]
        assert batch == single, (rank, nw, bits)
    for q in qs:
        lib.free_sequence(q)
    # strand / reading-frame searches, sharded: every frame's shard scores are gathered, every rank replays
    import tempfile
    import frames as F
    from oracle import port
    with tempfile.TemporaryDirectory() as tmp:
        for mode in ((F.NUCLEOTIDE, F.BOTH), (F.TRANS_QUERY, F.BOTH), (F.TRANS_DB, F.FORWARD), (F.TRANS_BOTH, F.BOTH)):
            F.check_product_mode(lib, port, mode, tmp)
    import torch.distributed as dist
    dist.barrier()
    print("MULTI_OK rank %d of %d" % (rank, world))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
