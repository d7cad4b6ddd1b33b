"""Multi-GPU parity (needs >= 2 GPUs; skipped otherwise): the database sharded over the GPUs - one process per GPU
(torchrun) or ONE process driving all of them - shard scores all-gathered with NCCL inside the library, must give exactly
the hit lists and alignments of the compiled reference."""
import json
import os
import subprocess
import sys

import pytest

import cases

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _expected(ref, tmp_path):
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
    return path


def _gpus():
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    return 2 if n < 4 else 4


@pytest.mark.parametrize("collective", ["nccl", "torch"])
def test_sharded_search_matches_reference(ref, tmp_path, collective):
    """one process per GPU; collective = the library's own NCCL communicator, or the host-supplied callback"""
    world = _gpus()
    path = _expected(ref, tmp_path)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"), LIBSSA_B200_COLLECTIVE=collective)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
                          "--master-addr", "127.0.0.1", "--master-port", "29621" if collective == "nccl" else "29622",
                          os.path.join(ROOT, "tests", "dist_gpu_worker.py"), str(path)],
                         capture_output=True, text=True, env=env, timeout=1200)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-5000:]
    assert out.stdout.count("MULTI_OK") == world


def test_single_process_multi_gpu_matches_reference(ref, tmp_path):
    """ONE process, several GPUs (ssa_b200_set_device_count): host thread per device, ncclCommInitAll inside the library"""
    n = _gpus()
    path = _expected(ref, tmp_path)
    env = dict(os.environ, PYTHONPATH=ROOT + os.pathsep + os.path.join(ROOT, "tests"))
    env.pop("LIBSSA_B200_GPUS", None)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "single_process_multi_worker.py"), str(path), str(n)],
                         capture_output=True, text=True, env=env, timeout=1200)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-5000:]
    assert "SINGLE_PROCESS_MULTI_OK %d GPUs" % n in out.stdout


def test_unchanged_c_host_uses_all_gpus_through_the_environment(prod, tmp_path):
    """a libssa.h-only C program gets several GPUs with zero source change: LIBSSA_B200_GPUS=n"""
    n = _gpus()
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "examples")])
    db = cases.db_path(["protein", 2000, 11, "P07327"])
    qf = os.path.join(cases.DATA, "P07327")
    outs = []
    for gpus in (1, n):
        env = dict(os.environ, LIBSSA_B200_GPUS=str(gpus))
        r = subprocess.run([os.path.join(ROOT, "examples", "search_example"), db, qf, "sw", "8", "25", "align"],
                           capture_output=True, text=True, timeout=600, env=env)
        assert r.returncode == 0, r.stderr[-3000:]
        # (with NCCL_DEBUG=VERSION in the environment NCCL prints its banner on stdout)
        outs.append([l for l in r.stdout.split("\n") if l and not l.startswith("libssa") and not l.startswith("NCCL version")])
    assert outs[0] == outs[1] and len(outs[0]) == 25
