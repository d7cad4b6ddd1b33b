#!/usr/bin/env python3
"""BASELINE config 4: Smith-Waterman, entry widths 8 and 16, P07327 (375 aa) and P33450 (5147 aa) against ONE synthetic
Swiss-Prot-sized database (570 000 records, ~187 M residues) sharded over the N GPUs of a box - strong scaling.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/config4_scale.py

Every rank loads the same FASTA file and owns the records at length-sorted positions rank, rank + N, ...; a search is the
shard's scoring, one NCCL all-gather of the shard scores on the library's stream, hit selection on every rank.  Timed:
wall clock around the libssa.h call between two barriers, median of 10, max over ranks.  After the timed runs rank 0
repeats every search unsharded (world size 1) and compares the hit lists.  Prints one JSON line on rank 0."""
import json
import os
import statistics
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
D = os.path.join(ROOT, "tests", "golden", "data")
N_SEQS, SEED = 570000, 20150406


def main():
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    from libssa_b200 import synth
    q0 = open(os.path.join(D, "P07327")).read().split("\n", 1)[1].replace("\n", "")
    path = os.path.join(tempfile.gettempdir(), "libssa_b200_c4_S%d_%d.fas" % (N_SEQS, SEED))
    if rank == 0 and not os.path.exists(path + ".ok"):
        synth.write_protein_db(path, N_SEQS, SEED, query=q0)      # before CUDA / NCCL are up (fork pool)
        open(path + ".ok", "w").write("ok")
    import torch
    import torch.distributed as dist
    import libssa_b200 as L
    from libssa_b200 import dist as ldist
    lib = L.load()
    lib.set_output_mode(L.OUTPUT_SILENT)
    ldist.init_from_env(lib)
    barrier = dist.barrier if world > 1 else (lambda: None)
    barrier()
    lib.init_symbol_translation(L.AMINOACID)
    lib.init_score_matrix(L.MATRIX_BUILDIN, "blosum62")
    lib.init_gap_penalties(-11, -1)
    lib.init_db(path)
    lib.prepare_db()
    st = lib.stats()
    rows, lists = [], {}
    for qn in ("P07327", "P33450"):
        q = lib.init_sequence_fasta(L.READ_FROM_FILE, os.path.join(D, qn))
        for bw in (8, 16):
            times = []
            for it in range(13):
                barrier()
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                hits = lib.sw_align(q, 10, bw, L.COMPUTE_SCORE)
                dt = time.perf_counter() - t0
                if world > 1:
                    t = torch.tensor([dt], device="cuda")
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    dt = float(t.item())
                if it >= 3:
                    times.append(dt)
            s = lib.stats()
            lists[(qn, bw)] = [(h.score, h.db_id) for h in hits]
            rows.append({"query": qn, "bit_width": bw, "ms": round(1e3 * statistics.median(times), 3), "kernel_ms_rank0": round(s.score_kernel_ms, 3),
                         "qlen": {"P07327": 375, "P33450": 5147}[qn]})
        lib.free_sequence(q)
    for r in rows:
        r["gcups"] = round(r["qlen"] * st.n_residues / (r["ms"] * 1e-3) / 1e9, 1)
    same = None
    if world > 1:
        dist.barrier()
    if rank == 0 and world > 1:   # the unsharded search on one GPU must return the same lists
        lib.set_shard(0, 1)
        lib.set_allgather(None)
        lib.init_db(path)
        same = True
        for qn in ("P07327", "P33450"):
            q = lib.init_sequence_fasta(L.READ_FROM_FILE, os.path.join(D, qn))
            for bw in (8, 16):
                same = same and [(h.score, h.db_id) for h in lib.sw_align(q, 10, bw, L.COMPUTE_SCORE)] == lists[(qn, bw)]
            lib.free_sequence(q)
    if rank == 0:
        print(json.dumps({"config": "BASELINE configs[3]: SW vs S-570k sharded by length-sorted round robin, strong scaling", "n_gpus": world,
                          "db_records": int(st.n_seqs), "db_residues": int(st.n_residues), "records_per_gpu": int(st.local_seqs),
                          "runs": rows, "hit_lists_equal_unsharded": same}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
