#!/usr/bin/env python3
"""bench.py - GCUPS of the libssa database-search hot path on B200 (BASELINE.json metric).

Workload (BASELINE.json configs[1], SURVEY.md 8d "Config 2"): Smith-Waterman, entry bit width 8 (the
reference's 8 -> 16 -> 64 bit overflow cascade is reproduced), COMPUTE_SCORE, hitcount 10, BLOSUM62,
gap open 11 / extend 1, the five UniProt queries O74807 (110 aa), P07327 (375), P27895 (1000),
P04775 (2005) and P33450 (5147) against a seeded synthetic Swiss-Prot-like protein database of
500 000 records per GPU (~165 M residues; packed size > L2, so consecutive searches cannot live in L2).
One STEP = the five searches.  GCUPS = sum(query length) x database residues / seconds / 1e9, real cells
only (reference benchmark/scripts/evaluate_threads.r:111-127).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU under torchrun)
    python bench.py --impl reference [--steps K] [--warmup W]      the reference's own AVX2 path on the host cores

value : whole-job GCUPS from the device time (CUDA events on the library's stream) of everything the K timed
        steps enqueue - database resident in HBM, max over ranks.
e2e   : the same metric from the host wall clock around the K steps of libssa.h calls (sw_align through the
        C-ABI: host query -> profile upload, kernels, score download, host hit selection), barrier +
        synchronize on both sides, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
DATA = os.path.join(ROOT, "tests", "golden", "data")

QUERIES = ["O74807", "P07327", "P27895", "P04775", "P33450"]
DB_SEQS_PER_GPU = 500000
DB_SEED = 20150404                 # 20150402 + config number
REF_SAMPLE_SEQS = 20000            # bounded sample of the same database for the CPU arm
HITCOUNT = 10
BIT_WIDTH = 8
SM_COUNT = 148
LANE_OPS_PER_CLK_PER_SM = 64       # measured: tools/microbench, profiles/pipes_r01.md (18.55 T lane-ops/s @ 1.958 GHz)
CELLS_PER_WORD = 2                 # int16 x 2 per 32-bit register
ALU_OPS_PER_WORD = 4.5             # binding (ALU/DPX) pipe instructions per packed word of the SW kernel


def query_text(name):
    txt = open(os.path.join(DATA, name)).read()
    return "".join(l.strip() for l in txt.split("\n") if l and not l.startswith(">"))


def query_lengths():
    return [len(query_text(q)) for q in QUERIES]


def db_file(n_seqs, rank0_builds=True, barrier=None):
    """The synthetic database as a FASTA file in the temp dir (built once per size, by rank 0)."""
    from libssa_b200 import synth
    path = os.path.join(tempfile.gettempdir(), "libssa_b200_bench_S%d_%d.fas" % (n_seqs, DB_SEED))
    meta = path + ".json"
    if rank0_builds and not os.path.exists(meta):
        residues = synth.write_protein_db(path + ".tmp", n_seqs, DB_SEED, query=query_text("P07327"))
        os.replace(path + ".tmp", path)
        json.dump({"n_seqs": n_seqs, "residues": residues}, open(meta, "w"))
    if barrier:
        barrier()
    return path, json.load(open(meta))


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line)

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, power, reasons = [], [], [], set()
        for line in self.lines:
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                mx.append(float(p[2]))
                power.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def run_reference(steps, warmup, n_gpus):
    """The reference's own CPU implementation (compiled unmodified: oracle/_ref) on a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    from oracle import refapi
    ref = refapi.reference()
    if ref is None:
        return {"impl": "reference", "unavailable": "oracle/_ref/libssa_ref.so was not built (needs /root/reference at build time)"}
    path, meta = db_file(REF_SAMPLE_SEQS)
    cores = os.cpu_count() or 1
    ref.set_output_mode(refapi.OUTPUT_SILENT)
    ref.set_simd_compute_mode(refapi.COMPUTE_ON_AVX2)
    ref.set_thread_count(cores)
    ref.set_chunk_size(1000)
    ref.init_symbol_translation(refapi.AMINOACID)
    ref.init_score_matrix(refapi.MATRIX_BUILDIN, "blosum62")
    ref.init_gap_penalties(-11, -1)
    ref.init_db(path)
    queries = [ref.init_sequence_fasta(refapi.READ_FROM_FILE, os.path.join(DATA, q)) for q in QUERIES]
    cells_per_step = sum(query_lengths()) * meta["residues"]

    def step():
        for q in queries:
            ref.sw_align(q, HITCOUNT, BIT_WIDTH, refapi.COMPUTE_SCORE)

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    gcups = cells_per_step * steps / dt / 1e9
    sample = "%d-record / %d-residue prefix sample of the same seeded database, all 5 queries, bit width 8" % (
        meta["n_seqs"], meta["residues"])
    return {
        "impl": "reference", "metric": "GCUPS", "value": round(gcups, 3), "unit": "GCUPS", "n_gpus": n_gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": round(dt / steps * 1e3, 3), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "int8/int16 saturating (AVX2)", "data": "synthetic",
        "config": workload_config(1, meta, sampled=True),
        "cpu_baseline": {"value": round(gcups, 3), "unit": "GCUPS", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": round(gcups, 3), "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def workload_config(n_gpus, meta, sampled=False):
    return {
        "workload": "SW bit_width=8 (8->16->64 overflow cascade), COMPUTE_SCORE, hitcount 10, BLOSUM62, gaps -11/-1, "
                    "queries O74807/P07327/P27895/P04775/P33450 (110/375/1000/2005/5147 aa) vs synthetic protein DB",
        "baseline_config": "configs[1]",
        "db_records": meta["n_seqs"], "db_residues": meta["residues"],
        "db_records_per_gpu": meta["n_seqs"] // n_gpus, "query_lengths": query_lengths(),
        "db_sampled_for_cpu": sampled,
        "l2_policy": "inputs larger than L2: the packed database of one GPU is %.0f MB and is streamed once per query"
                     % (meta["residues"] / max(n_gpus, 1) * 1.0 / 1e6),
        "parallelism": "db-shard x%d (one process per GPU, NCCL all-gather of shard scores)" % n_gpus,
        "arithmetic": "int16 x 2 SIMD-in-word (DPX), exact int32 tier for records whose scores may leave the int16 range",
    }


def run_ours(steps, warmup, n_gpus):
    import torch
    import libssa_b200 as L
    from libssa_b200 import dist as ldist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - libssa_b200 has no CPU fallback")
    lib = L.load()
    # rank 0 generates the database BEFORE any CUDA/NCCL initialisation (the generator forks worker processes)
    if int(os.environ.get("RANK", "0")) == 0:
        db_file(DB_SEQS_PER_GPU * n_gpus, rank0_builds=True)
    rank, world, local = ldist.init_from_env(lib)
    if world != n_gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch with torch.distributed.run for N > 1)" % (n_gpus, world))
    import torch.distributed as tdist
    torch.cuda.set_device(local)

    def barrier():
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        tdist.all_reduce(t, op=tdist.ReduceOp.MAX)
        return float(t.item())

    barrier()
    path, meta = db_file(DB_SEQS_PER_GPU * n_gpus, rank0_builds=False)
    lib.set_output_mode(L.OUTPUT_SILENT)
    lib.set_thread_count(1)
    lib.set_chunk_size(1000)
    lib.init_symbol_translation(L.AMINOACID)
    lib.init_score_matrix(L.MATRIX_BUILDIN, "blosum62")
    lib.init_gap_penalties(-11, -1)
    t0 = time.perf_counter()
    lib.init_db(path)
    t_load = time.perf_counter() - t0
    t0 = time.perf_counter()
    lib.prepare_db()
    t_prepare = time.perf_counter() - t0
    st = lib.stats()
    queries = [lib.init_sequence_fasta(L.READ_FROM_FILE, os.path.join(DATA, q)) for q in QUERIES]
    qlens = query_lengths()
    cells_per_step = sum(qlens) * meta["residues"]          # whole job, all ranks

    acc = {"device_ms": 0.0, "kernel_ms": 0.0, "launches": 0, "h2d": 0, "d2h": 0, "ovf8": 0, "ovf16": 0, "wide": 0,
           "kernel_ms_by_query": [0.0] * len(QUERIES)}
    top = []

    def step(record):
        for i, q in enumerate(queries):
            hits = lib.sw_align(q, HITCOUNT, BIT_WIDTH, L.COMPUTE_SCORE)
            if record:
                s = lib.stats()
                acc["device_ms"] += s.device_ms
                acc["kernel_ms"] += s.score_kernel_ms
                acc["kernel_ms_by_query"][i] += s.score_kernel_ms
                acc["launches"] += s.launches
                acc["h2d"] += s.h2d_bytes
                acc["d2h"] += s.d2h_bytes
                acc["ovf8"] += s.overflow_8
                acc["ovf16"] += s.overflow_16
                acc["wide"] += s.wide_tier_seqs
                if len(top) < len(QUERIES):
                    top.append([hits[0].score, hits[0].db_id] if hits else None)

    for _ in range(warmup):
        step(False)
    sampler = ClockSampler(local)
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    for _ in range(steps):
        step(True)
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None

    wall = max_over_ranks(wall)
    device_s = max_over_ranks(acc["device_ms"] / 1e3)
    kernel_s = max_over_ranks(acc["kernel_ms"] / 1e3)
    value = cells_per_step * steps / device_s / 1e9
    e2e = cells_per_step * steps / wall / 1e9

    line = None
    if rank == 0:
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        peak = SM_COUNT * sm_mhz * 1e6 * LANE_OPS_PER_CLK_PER_SM * CELLS_PER_WORD / ALU_OPS_PER_WORD / 1e9   # GCUPS per GPU
        local_cells = cells_per_step / n_gpus * steps
        achieved = local_cells / kernel_s / 1e9
        traffic = None
        prof = os.path.join(ROOT, "profiles", "roofline_r01.json")
        if os.path.exists(prof):
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        # the stream is read once per pass: 2 bytes per pair column ~ 1 byte per residue; 4 bytes per score written
        stream_bytes = sum(((q + 1023) // 1024) for q in qlens) * (meta["residues"] / n_gpus) * 1.0 + 4.0 * len(qlens) * meta["n_seqs"] / n_gpus
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        line = {
            "metric": "GCUPS", "value": round(value, 2), "unit": "GCUPS", "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": round(device_s / steps * 1e3, 3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "int16", "data": "synthetic",
            "config": workload_config(n_gpus, meta),
            "e2e": {"value": round(e2e, 2), "unit": "GCUPS", "ms_per_step": round(wall / steps * 1e3, 3),
                    "h2d_bytes_per_step": acc["h2d"] // steps, "d2h_bytes_per_step": acc["d2h"] // steps,
                    "note": "wall clock around the libssa.h sw_align calls; database resident (as in the reference's own timing, "
                            "benchmark/src/benchmark_util.c:29-53); cold start reported in `cold_start`"},
            "gpu_launches": acc["launches"],
            "clocks": clocks,
            "roofline": {
                "bound": "int_alu", "kernel": "ssa::score_kernel<PACKED,...> (SW, s16x2)", "achieved": round(achieved, 1),
                "peak": round(peak, 1), "unit": "GCUPS", "frac": round(achieved / peak, 4), "traffic": traffic,
                "traffic_note": "ncu dram bytes per launch of the dominant kernel = one 32 x 32-row pass of the 5147-aa query over the whole "
                                "database (profiles/prof_r01_top.md): 0.16 GB of column stream, the rest are the pass-boundary rows it reads "
                                "from the previous pass and writes for the next one (8 B per column each way; a single-pass launch moves "
                                "0.17 GB) - about 48 GB/s, HBM is not the bound",
                "peak_formula": "SMs x f_clk x L x c / I = %d x %.0f MHz x %d lane-ops/clk/SM x %d cells/word / %.1f ALU-pipe ops/word"
                                % (SM_COUNT, sm_mhz, LANE_OPS_PER_CLK_PER_SM, CELLS_PER_WORD, ALU_OPS_PER_WORD),
                "peak_source": "L measured on this pool (profiles/pipes_r01.md); f_clk = median SM clock sampled during the timed region",
                "per_query_kernel_gcups": [round(q * meta["residues"] / n_gpus * steps / (ms / 1e3) / 1e9, 1) if ms > 0 else None
                                           for q, ms in zip(qlens, acc["kernel_ms_by_query"])],
                "hbm": {"algorithmic_bytes_per_step": int(stream_bytes), "achieved_gbs": round(stream_bytes * steps / kernel_s / 1e9, 2),
                        "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                        "note": "not the bound: %.0f cell updates per database byte" % (sum(qlens) / len(qlens))},
            },
            "overflow": {"reference_8bit_overflows_per_step": acc["ovf8"] // steps, "reference_16bit_overflows_per_step": acc["ovf16"] // steps,
                         "wide_tier_records_per_step": acc["wide"] // steps},
            "cold_start": {"fasta_load_s": round(t_load, 3), "pack_and_upload_s": round(t_prepare, 3), "pack_ms": round(st.pack_ms, 1),
                           "upload_ms": round(st.upload_ms, 1), "db_device_bytes": st.db_device_bytes},
            "top_hits": top,
        }
        if n_gpus == 1:
            try:
                ref_line = run_reference(3, 1, 1)
                line["cpu_baseline"] = ref_line.get("cpu_baseline") or {"unavailable": ref_line.get("unavailable")}
            except Exception as e:  # the CPU leg must never take the GPU number down with it
                line["cpu_baseline"] = {"unavailable": repr(e)}
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()
    return line


def main():
    # the contract is ONE JSON line on stdout: park the real stdout and send any library chatter
    # (e.g. NCCL's version banner) to stderr
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    if args.impl == "reference":
        line = run_reference(args.steps, args.warmup, args.gpus)
    else:
        line = run_ours(args.steps, max(args.warmup, 0), args.gpus)
    if line is not None:
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
