"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): the database sharded over one process per GPU,
shard scores all-gathered over NCCL, must give exactly the hit lists and alignments of the compiled reference."""
import json
import os
import subprocess
import sys

import pytest

import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sharded_search_matches_reference(ref, tmp_path):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    selected = [c for c in cases.CASES if c["name"] in ("synth_protein_2000_P07327", "synth_protein_edge_O74807", "bigger_db_AF091148",
                                                         "synth_nucleotide_800", "overflow_8_16_64_cascade")]
    big = {"name": "synth_protein_30000", "symtype": cases.AMINOACID, "scoring": "blosum62", "gaps": [-11, -1],
           "db": ["protein", 30000, 123, "P07327"], "query": "P07327",
           "runs": [["sw", 8, 50, 0], ["sw", 16, 50, 0], ["nw", 16, 50, 0], ["sw", 64, 20, 1]]}
    expected = []
    for case in selected + [big]:
        case = dict(case, runs=[r for r in case["runs"] if not cases.TIE_AWARE(r[0], r[1])])
        expected.append({"case": case, "result": cases.run_case(ref, case)})
    path = tmp_path / "expected.json"
    json.dump(expected, open(path, "w"))
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
                          "--master-addr", "127.0.0.1", "--master-port", "29621",
                          os.path.join(ROOT, "tests", "dist_gpu_worker.py"), str(path)],
                         capture_output=True, text=True, env=env, timeout=1200)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-5000:]
    assert out.stdout.count("MULTI_OK") == world
