"""Worker of tests/test_host.py::test_sharded_merge_two_ranks_gloo (launched by torch.distributed.run)."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

import cases
import libssa_b200
from libssa_b200 import synth
from oracle import port


def main():
    dist.init_process_group(backend="gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    lib = libssa_b200.load()
    q = synth.random_protein(np.random.default_rng(1), 60)
    seqs = synth.protein_db(400, 77, query=q, homolog_frac=0.05, n_dup=40, mu=4.0, sigma=0.5, min_len=0)
    seqs[5] = ""
    codes = [port.map_db(s, False) for s in seqs]
    qc = port.map_query(q, False)
    M = port.constant_matrix(3, -2)
    lens = np.array([len(s) for s in seqs], dtype=np.uint64)

    # shard exactly like ssa_b200_set_shard: global length-sorted order dealt round-robin
    sorted_all = lib.host_sorted_order(lens)
    nglob = len(sorted_all)
    stride = (nglob + world - 1) // world
    mine = sorted_all[rank::world]
    local = np.full(stride, np.iinfo(np.int32).min, dtype=np.int32)
    for mode_nw in (False, True):
        sc = port.score_db([codes[i] for i in mine], qc, M, -4, -1, mode_nw)
        local[:len(mine)] = sc
        gathered = [torch.zeros(stride, dtype=torch.int32) for _ in range(world)]
        dist.all_gather(gathered, torch.from_numpy(local.copy()))
        g = torch.stack(gathered).numpy()                    # [rank][local position]
        for bw in (64, 16):
            order = lib.host_arrival_order(lens, 100, bw)
            pos_of_id = np.full(len(lens), -1, dtype=np.int64)
            pos_of_id[order] = np.arange(len(order))
            arrival = np.zeros(len(order), dtype=np.int32)
            for r in range(world):
                for s in range(stride):
                    gi = s * world + r
                    if gi < nglob and pos_of_id[sorted_all[gi]] >= 0:
                        arrival[pos_of_id[sorted_all[gi]]] = g[r, s]
            hits = lib.host_replay(arrival, order, 25)
            expect = port.search(codes, qc, M, -4, -1, mode_nw, 25, bw, 100)
            assert hits == expect, (rank, mode_nw, bw, hits[:5], expect[:5])
    print("MERGE_OK rank %d" % rank)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
