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
