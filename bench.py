#!/usr/bin/env python3
"""bench.py - GCUPS of the libssa database-search hot path on B200 (BASELINE.json metric).

Headline workload, every N (BASELINE.json configs[3], SURVEY.md 8d "Config 4"): Smith-Waterman, entry bit widths 8 AND 16
(the reference's 8 -> 16 -> 64 bit overflow cascade is reproduced), COMPUTE_SCORE, hitcount 10, BLOSUM62, gap open 11 /
extend 1, the UniProt queries P07327 (375 aa) and P33450 (5147 aa) against ONE seeded synthetic Swiss-Prot-sized protein
database (570 000 records, ~187 M residues; packed size > L2, so consecutive searches cannot live in L2), sharded by
residue count over the N GPUs - STRONG scaling.  One STEP = the four searches (2 queries x 2 bit widths).
GCUPS = sum(query length) x database residues / seconds / 1e9, real cells only (reference
benchmark/scripts/evaluate_threads.r:111-127).

    python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (one process per GPU under torchrun)
    python bench.py --impl reference [--steps K] [--warmup W]      the reference's own AVX2 path on the host cores,
                                                                   SAME database, queries, bit widths and step

value : whole-job GCUPS from the device time (CUDA events on the library's stream) of everything the K timed steps
        enqueue - database resident in HBM, max over ranks.
e2e   : the same metric from the host wall clock around the K steps of libssa.h calls (sw_align through the C-ABI: host
        query -> profile upload, kernels, NCCL all-gather of the shard scores, hit download), barrier + synchronize on both
        sides, max over ranks.
per_config (N = 1 only, --no-per-config skips them): BASELINE configs[1] (SW-8, five queries, S-500k), configs[2] (NW-16,
        1 000 queries vs a length-matched 100 k database, score pass + alignments) and configs[4] (nucleotide, 10 kbp
        query, SW + NW with alignments of the top 10), each with its own roofline and cpu_baseline.
parity_check : untimed, after the timed region - hit lists against the compiled reference (one thread) on the full
        database and, at N > 1, sharded == unsharded.
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

HITCOUNT = 10
SM_COUNT = 148
LANE_OPS_PER_CLK_PER_SM = 64       # measured: tools/microbench, profiles/pipes_r01.md (18.55 T lane-ops/s @ 1.958 GHz)
CELLS_PER_WORD = 2                 # int16 x 2 per 32-bit register
ALU_OPS_PER_WORD = 4.5             # binding (ALU/DPX) pipe instructions per packed word of the SW kernel

# seeds: 20150402 + config number
HEAD = {"name": "configs[3]", "db_seqs": 570000, "seed": 20150406, "queries": ["P07327", "P33450"], "bit_widths": [8, 16]}
CFG1 = {"name": "configs[1]", "db_seqs": 500000, "seed": 20150404, "queries": ["O74807", "P07327", "P27895", "P04775", "P33450"],
        "bit_widths": [8]}


def query_text(name):
    txt = open(os.path.join(DATA, name)).read()
    return "".join(l.strip() for l in txt.split("\n") if l and not l.startswith(">"))


def protein_db_file(n_seqs, seed, build=True):
    """The synthetic protein database as a FASTA file in the temp dir (built once per (size, seed))."""
    from libssa_b200 import synth
    path = os.path.join(tempfile.gettempdir(), "libssa_b200_bench_S%d_%d.fas" % (n_seqs, seed))
    meta = path + ".json"
    if build and not os.path.exists(meta):
        residues = synth.write_protein_db(path + ".tmp", n_seqs, seed, query=query_text("P07327"))
        os.replace(path + ".tmp", path)
        json.dump({"n_seqs": n_seqs, "residues": residues}, open(meta + ".tmp", "w"))
        os.replace(meta + ".tmp", meta)
    return path, json.load(open(meta))


def length_matched_db_file():
    """configs[2]: 100 000 proteins within +-10 % of 350 residues (SURVEY.md 8d "Config 3"); returns path, sequences."""
    import numpy as np
    from libssa_b200 import synth
    seqs = synth.protein_db(100000, 20150405, mu=float(np.log(350)), sigma=0.03, min_len=315, max_len=385, n_dup=50)
    path = os.path.join(tempfile.gettempdir(), "libssa_b200_bench_LM100k_20150405.fas")
    if not os.path.exists(path):
        synth.write_fasta(path + ".tmp", seqs)
        os.replace(path + ".tmp", path)
    return path, seqs


def nucleotide_db_file():
    """configs[4]: 10 kbp query with 2 % IUPAC codes, 50 000 records around 1.5 kbp plus 200 planted homologs, 100 of them
    8-12 kbp (SURVEY.md 8d "Config 5"); returns path, sequences, query, indices of the long homologs."""
    import numpy as np
    from libssa_b200 import synth
    q = synth.random_nucleotide(20150407, 10000, iupac_rate=0.02)
    seqs = synth.nucleotide_db(50000, 20150407, query=q, n_homologs=200, mu=7.25, sigma=0.35, min_len=100, max_len=12000)
    hom = np.random.default_rng(7).choice(50000, 200, replace=False)
    for h in hom[:100]:
        seqs[h] = synth.mutate(np.random.default_rng(int(h)), q[: 8000 + int(h) % 4000], 0.15, "ACGT")
    path = os.path.join(tempfile.gettempdir(), "libssa_b200_bench_NT50k_20150407.fas")
    if not os.path.exists(path):
        synth.write_fasta(path + ".tmp", seqs)
        os.replace(path + ".tmp", path)
    return path, seqs, q, [int(h) for h in hom[:100]]


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


def qlens_of(cfg):
    return [len(query_text(q)) for q in cfg["queries"]]


def workload_config(cfg, n_gpus, meta):
    ql = qlens_of(cfg)
    return {
        "workload": "SW entry bit widths %s (8->16->64 overflow cascade reproduced), COMPUTE_SCORE, hitcount %d, BLOSUM62, gaps -11/-1, "
                    "queries %s (%s aa) vs ONE synthetic Swiss-Prot-like protein DB of %d records sharded over the GPUs"
                    % ("/".join(str(b) for b in cfg["bit_widths"]), HITCOUNT, "/".join(cfg["queries"]), "/".join(str(x) for x in ql),
                       meta["n_seqs"]),
        "baseline_config": cfg["name"],
        "db_records": meta["n_seqs"], "db_residues": meta["residues"], "db_seed": cfg["seed"],
        "query_lengths": ql, "bit_widths": cfg["bit_widths"], "searches_per_step": len(ql) * len(cfg["bit_widths"]),
        "l2_policy": "inputs larger than L2: the packed database (%.0f MB in total) is streamed once per query pass"
                     % (meta["residues"] / 1e6),
    }


def setup_protein(api, path):
    api.set_output_mode(0)
    api.set_chunk_size(1000)
    api.init_symbol_translation(1, 1, 1, 1)          # AMINOACID, FORWARD_STRAND
    api.init_score_matrix(2, "blosum62")             # MATRIX_BUILDIN
    api.init_gap_penalties(-11, -1)
    api.init_db(path)


def keyed(hits, align=False):
    return [(h.score, h.db_id) + ((h.cigar, h.q_start, h.q_end, h.d_start, h.d_end) if align else ()) for h in hits]


# ---------------------------------------------------------------------------------------------------------
# the reference's own CPU implementation (compiled unmodified: oracle/_ref), all host threads
# ---------------------------------------------------------------------------------------------------------

def reference_api():
    from oracle import refapi
    ref = refapi.reference()
    if ref is not None:
        ref.set_output_mode(refapi.OUTPUT_SILENT)
        ref.set_simd_compute_mode(refapi.COMPUTE_ON_AVX2)
    return ref


def reference_protein_steps(ref, cfg, path, meta, steps, warmup, cores):
    """K steps of cfg's searches on the full database; returns (GCUPS, seconds per step)."""
    ref.set_thread_count(cores)
    setup_protein(ref, path)
    queries = [ref.init_sequence_fasta(0, os.path.join(DATA, q)) for q in cfg["queries"]]
    cells_per_step = sum(qlens_of(cfg)) * len(cfg["bit_widths"]) * meta["residues"]

    def step():
        for bw in cfg["bit_widths"]:
            for q in queries:
                ref.sw_align(q, HITCOUNT, bw, 0)

    # warm-up of a CPU arm only has to page the database in: one search of the first query per warm-up step (at least one)
    for _ in range(max(1, warmup)):
        ref.sw_align(queries[0], HITCOUNT, 16, 0)
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    for q in queries:
        ref.free_sequence(q)
    return cells_per_step * steps / dt / 1e9, dt / steps


def run_reference_port(path, meta, steps, warmup, n_gpus):
    """oracle/_ref is not there (the reference was not compiled on this machine): the oracle's C restatement of the same algorithm,
    one scalar thread, on a bounded sample of the same database - the other place bench.py may execute oracle/."""
    from oracle import port
    n_sample = 600
    import numpy as np
    codes = [port.map_db(s, False) for _, s in port.read_fasta(path)[:n_sample]]
    residues = sum(len(c) for c in codes)
    M = np.array(json.load(open(os.path.join(ROOT, "tests", "golden", "matrices.json")))["blosum62"], dtype=np.int64).reshape(32, 32)
    qcodes = [port.map_query(query_text(q), False) for q in HEAD["queries"]]

    def step():
        for bw in HEAD["bit_widths"]:
            for qc in qcodes:
                port.search(codes, qc, M, -11, -1, False, HITCOUNT, bw)

    step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    gcups = sum(qlens_of(HEAD)) * len(HEAD["bit_widths"]) * residues * steps / dt / 1e9
    sample = "oracle port (oracle/restate.c, scalar, 1 thread) on the first %d records / %d residues of the %d-record database, same %d searches per step" % (
        len(codes), residues, meta["n_seqs"], len(HEAD["queries"]) * len(HEAD["bit_widths"]))
    return {
        "impl": "reference", "metric": "GCUPS", "value": round(gcups, 3), "unit": "GCUPS", "n_gpus": n_gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": round(dt / steps * 1e3, 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int64", "data": "synthetic", "config": workload_config(HEAD, n_gpus, meta),
        "cpu_baseline": {"value": round(gcups, 3), "unit": "GCUPS", "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": round(gcups, 3), "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def run_reference(steps, warmup, n_gpus):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return None
    ref = reference_api()
    path, meta = protein_db_file(HEAD["db_seqs"], HEAD["seed"])
    if ref is None:
        return run_reference_port(path, meta, steps, warmup, n_gpus)
    
    cores = os.cpu_count() or 1
    gcups, s_per_step = reference_protein_steps(ref, HEAD, path, meta, steps, warmup, cores)
    sample = "the full workload: all %d records / %d residues, %d searches per timed step, %d threads, chunk size 1000 (warm-up steps: one search each)" % (
        meta["n_seqs"], meta["residues"], len(HEAD["queries"]) * len(HEAD["bit_widths"]), cores)
    return {
        "impl": "reference", "metric": "GCUPS", "value": round(gcups, 3), "unit": "GCUPS", "n_gpus": n_gpus, "steps": steps,
        "warmup": warmup, "ms_per_step": round(s_per_step * 1e3, 3), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int16", "data": "synthetic",
        "config": workload_config(HEAD, n_gpus, meta),
        "arithmetic": "int8/int16 saturating AVX2 (the reference's own kernels), 64-bit scalar fallback",
        "cpu_baseline": {"value": round(gcups, 3), "unit": "GCUPS", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": round(gcups, 3), "unit": "GCUPS", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


# ---------------------------------------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------------------------------------

class Acc:
    def __init__(self, nq):
        self.device_ms = 0.0
        self.kernel_ms = 0.0
        self.launches = self.h2d = self.d2h = self.ovf8 = self.ovf16 = self.wide = 0
        self.kernel_ms_by_search = [0.0] * nq

    def add(self, s, i):
        self.device_ms += s.device_ms
        self.kernel_ms += s.score_kernel_ms
        self.kernel_ms_by_search[i] += s.score_kernel_ms
        self.launches += s.launches
        self.h2d += s.h2d_bytes
        self.d2h += s.d2h_bytes
        self.ovf8 += s.overflow_8
        self.ovf16 += s.overflow_16
        self.wide += s.wide_tier_seqs


def peak_gcups(sm_mhz):
    return SM_COUNT * sm_mhz * 1e6 * LANE_OPS_PER_CLK_PER_SM * CELLS_PER_WORD / ALU_OPS_PER_WORD / 1e9


def roofline_block(cells_local, kernel_s, sm_mhz, extra=None):
    peak = peak_gcups(sm_mhz)
    achieved = cells_local / kernel_s / 1e9 if kernel_s > 0 else 0.0
    r = {"bound": "int_alu", "kernel": "ssa::sw32_kernel<R, MULTIPASS, NW, G, DUAL> / ssa::score_kernel<PACKED, ...> (s16x2, DPX)", "achieved": round(achieved, 1),
         "peak": round(peak, 1), "unit": "GCUPS", "frac": round(achieved / peak, 4),
         "peak_formula": "SMs x f_clk x L x c / I = %d x %.0f MHz x %d lane-ops/clk/SM x %d cells/word / %.1f ALU-pipe ops/word"
                         % (SM_COUNT, sm_mhz, LANE_OPS_PER_CLK_PER_SM, CELLS_PER_WORD, ALU_OPS_PER_WORD),
         "peak_source": "L measured on this pool (profiles/pipes_r01.md); f_clk = median SM clock sampled during the timed region"}
    if extra:
        r.update(extra)
    return r


def run_protein_workload(lib, L, cfg, path, meta, steps, warmup, barrier, max_over_ranks, rank, local, n_gpus):
    """warm-up + K timed steps of cfg on the resident database; returns a dict of raw measurements"""
    queries = [lib.init_sequence_fasta(L.READ_FROM_FILE, os.path.join(DATA, q)) for q in cfg["queries"]]
    searches = [(bw, i) for bw in cfg["bit_widths"] for i in range(len(queries))]
    acc = Acc(len(searches))
    top = []

    def step(record):
        for si, (bw, qi) in enumerate(searches):
            hits = lib.sw_align(queries[qi], HITCOUNT, bw, L.COMPUTE_SCORE)
            if record:
                acc.add(lib.stats(), si)
                if len(top) < len(searches):
                    top.append(keyed(hits))

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
    for q in queries:
        lib.free_sequence(q)
    return {"wall": max_over_ranks(wall), "device_s": max_over_ranks(acc.device_ms / 1e3), "kernel_s": max_over_ranks(acc.kernel_ms / 1e3),
            "acc": acc, "clocks": clocks, "top": top, "searches": searches}


def config2_record(lib, L, ref, sm_mhz, cores):
    """BASELINE configs[2]: NW-16, 1 000 queries vs a length-matched 100 k database; score pass, then alignments of the top 10."""
    path, seqs = length_matched_db_file()
    residues = sum(len(s) for s in seqs)
    setup_protein(lib, path)
    lib.prepare_db()
    NQ = 1000
    queries = [lib.init_sequence_fasta(L.READ_FROM_STRING, seqs[i]) for i in range(NQ)]
    cells = sum(len(seqs[i]) for i in range(NQ)) * residues
    lib.align_batch(queries[:8], HITCOUNT, 16, L.COMPUTE_ALIGNMENT, nw=True)         # warm-up (shapes, buffers)
    t0 = time.perf_counter()
    lists = lib.align_batch(queries, HITCOUNT, 16, L.COMPUTE_SCORE, nw=True)
    t_score = time.perf_counter() - t0
    st = lib.stats()
    dev_score_s, launches, h2d, d2h = st.device_ms / 1e3, st.launches, st.h2d_bytes, st.d2h_bytes
    t0 = time.perf_counter()
    alists = lib.align_batch(queries, HITCOUNT, 16, L.COMPUTE_ALIGNMENT, nw=True)
    t_align = time.perf_counter() - t0
    # the reference API proper: one nw_align call per query (src/libssa.h:250) - a 100-query subset, also the kernel time
    NS = 100
    kernel_ms = 0.0
    t0 = time.perf_counter()
    for q in queries[:NS]:
        lib.nw_align(q, HITCOUNT, 16, L.COMPUTE_SCORE)
        kernel_ms += lib.stats().score_kernel_ms
    t_single = time.perf_counter() - t0
    cells_ns = sum(len(seqs[i]) for i in range(NS)) * residues
    rec = {
        "config": "configs[2]",
        "workload": "NW entry bit width 16, %d queries (the first records of the database, 315-385 aa) vs length-matched protein DB of "
                    "100 000 records / %d residues, BLOSUM62 -11/-1, hitcount %d: score pass, then COMPUTE_ALIGNMENT of the top hits"
                    % (NQ, residues, HITCOUNT),
        "value": round(cells / dev_score_s / 1e9, 1), "unit": "GCUPS",
        "value_note": "device time of the score pass of all %d queries through ssa_b200_align_batch" % NQ,
        "e2e": {"value": round(cells / t_score / 1e9, 1), "unit": "GCUPS", "ms_per_query": round(t_score / NQ * 1e3, 3),
                "h2d_bytes": h2d, "d2h_bytes": d2h, "note": "wall clock around ONE ssa_b200_align_batch call, host queries in, hit lists out"},
        "e2e_with_alignments": {"value": round(cells / t_align / 1e9, 1), "unit": "GCUPS", "ms_per_query": round(t_align / NQ * 1e3, 3)},
        "e2e_single_calls": {"value": round(cells_ns / t_single / 1e9, 1), "unit": "GCUPS", "ms_per_query": round(t_single / NS * 1e3, 3),
                             "note": "one nw_align call per query (the reference API), %d-query subset" % NS},
        "gpu_launches": launches,
        "roofline": roofline_block(cells, dev_score_s, sm_mhz, {
            "note": "device time of the batch call (scoring two queries per word, selection, copies) against the Smith-Waterman "
                    "roofline of one query per word (4.5 ALU-pipe instructions per word; the two-query NW recurrence needs 3)",
            "single_call_kernel_gcups": round(cells_ns / (kernel_ms / 1e3) / 1e9, 1) if kernel_ms > 0 else None}),
    }
    parity = {}
    if ref is not None:
        # parity (one thread, deterministic) on 5 queries with CIGARs; throughput on a 50-query subset with all cores
        ref.set_thread_count(1)
        setup_protein(ref, path)
        ok = True
        for i in range(5):
            rq = ref.init_sequence_fasta(1, seqs[i])
            ok = ok and keyed(ref.nw_align(rq, HITCOUNT, 16, 1), True) == keyed(alists[i], True)
            ok = ok and keyed(ref.nw_align(rq, HITCOUNT, 16, 0)) == keyed(lists[i])
            ref.free_sequence(rq)
        parity = {"hit_lists_and_cigars_equal_reference": bool(ok), "queries_checked": 5}
        ref.set_thread_count(cores)
        NR = 50
        rqs = [ref.init_sequence_fasta(1, seqs[i]) for i in range(NR)]
        ref.nw_align(rqs[0], HITCOUNT, 16, 0)
        t0 = time.perf_counter()
        for rq in rqs:
            ref.nw_align(rq, HITCOUNT, 16, 0)
        t_ref = time.perf_counter() - t0
        for rq in rqs:
            ref.free_sequence(rq)
        rec["cpu_baseline"] = {"value": round(sum(len(seqs[i]) for i in range(NR)) * residues / t_ref / 1e9, 2), "unit": "GCUPS",
                               "cores": cores, "kind": "reference",
                               "sample": "the first %d of the %d queries, score pass, full 100 000-record database, %.1f s" % (NR, NQ, t_ref)}
    rec["parity_check"] = parity
    for q in queries:
        lib.free_sequence(q)
    return rec


def config4_record(lib, L, ref, sm_mhz, cores):
    """BASELINE configs[4]: nucleotide SW / NW, 10 kbp query, COMPUTE_ALIGNMENT of the top 10."""
    path, seqs, qtext, long_hom = nucleotide_db_file()
    residues = sum(len(s) for s in seqs)

    def setup(api, db):
        api.set_output_mode(0)
        api.set_chunk_size(1000)
        api.init_symbol_translation(0, 1, 1, 1)      # NUCLEOTIDE, FORWARD_STRAND
        api.init_constant_scores(5, -4)
        api.init_gap_penalties(-4, -2)
        api.init_db(db)

    setup(lib, path)
    lib.prepare_db()
    q = lib.init_sequence_fasta(L.READ_FROM_STRING, qtext)
    qlen = len(qtext)
    rec = {"config": "configs[4]",
           "workload": "NUCLEOTIDE, match 5 / mismatch -4, gaps -4/-2, %d bp query (2 %% IUPAC codes) vs %d records / %d nt incl. 100 planted "
                       "8-12 kbp homologs, entry bit width 16, hitcount %d, COMPUTE_ALIGNMENT" % (qlen, len(seqs), residues, HITCOUNT),
           "unit": "GCUPS", "modes": {}}
    reps = 3
    kernel_s_total, cells_total = 0.0, 0
    for mode in ("sw", "nw"):
        f = lib.sw_align if mode == "sw" else lib.nw_align
        f(q, HITCOUNT, 16, L.COMPUTE_ALIGNMENT)
        dev = ker = 0.0
        launches = wide = 0
        t0 = time.perf_counter()
        for _ in range(reps):
            hits = f(q, HITCOUNT, 16, L.COMPUTE_ALIGNMENT)
            s = lib.stats()
            dev += s.device_ms
            ker += s.score_kernel_ms
            launches += s.launches
            wide = s.wide_tier_seqs
        wall = time.perf_counter() - t0
        cells = qlen * residues * reps
        kernel_s_total += ker / 1e3
        cells_total += cells
        rec["modes"][mode] = {"e2e": round(cells / wall / 1e9, 1), "ms_per_search": round(wall / reps * 1e3, 2),
                              "score_kernel_gcups": round(cells / (ker / 1e3) / 1e9, 1), "gpu_launches": launches // reps,
                              "wide_tier_records": wide, "longest_cigar": max(len(h.cigar) for h in hits), "top_score": hits[0].score}
    rec["value"] = round(cells_total / kernel_s_total / 1e9, 1)
    rec["value_note"] = "score kernels of both modes (device time); e2e per mode includes the alignments of the top %d" % HITCOUNT
    rec["e2e"] = {"value": round(2 * qlen * residues / ((rec["modes"]["sw"]["ms_per_search"] + rec["modes"]["nw"]["ms_per_search"]) / 1e3) / 1e9, 1),
                  "unit": "GCUPS", "note": "wall clock around sw_align + nw_align with COMPUTE_ALIGNMENT, host query in, CIGARs out"}
    rec["roofline"] = roofline_block(cells_total, kernel_s_total, sm_mhz)
    lib.free_sequence(q)
    if ref is not None:
        # bounded sample for the CPU: the reference falls to its scalar 64-bit tier on the long homologs (NW leaves int16)
        idx = long_hom[:6] + list(range(1000))
        from libssa_b200 import synth
        small = os.path.join(tempfile.gettempdir(), "libssa_b200_bench_NT_sample.fas")
        synth.write_fasta(small, [seqs[i] for i in idx])
        res_small = sum(len(seqs[i]) for i in idx)
        setup(lib, small)
        q = lib.init_sequence_fasta(L.READ_FROM_STRING, qtext)
        ours = {m: keyed(getattr(lib, m + "_align")(q, HITCOUNT, 16, L.COMPUTE_ALIGNMENT), True) for m in ("sw", "nw")}
        lib.free_sequence(q)
        ref.set_thread_count(1)
        setup(ref, small)
        rq = ref.init_sequence_fasta(1, qtext)
        ok = all(keyed(getattr(ref, m + "_align")(rq, HITCOUNT, 16, 1), True) == ours[m] for m in ("sw", "nw"))
        ref.set_thread_count(cores)
        t0 = time.perf_counter()
        n = 0
        while time.perf_counter() - t0 < 6.0:
            ref.sw_align(rq, HITCOUNT, 16, 1)
            ref.nw_align(rq, HITCOUNT, 16, 1)
            n += 1
        t_ref = time.perf_counter() - t0
        ref.free_sequence(rq)
        rec["cpu_baseline"] = {"value": round(2 * n * qlen * res_small / t_ref / 1e9, 2), "unit": "GCUPS", "cores": cores, "kind": "reference",
                               "sample": "%d-record sample (six 8-12 kbp homologs + the first 1000 records, %d nt), SW + NW with alignments, "
                                         "%d repetitions in %.1f s" % (len(idx), res_small, n, t_ref)}
        rec["parity_check"] = {"hit_lists_and_cigars_equal_reference": bool(ok), "sample_records": len(idx)}
    return rec


def run_ours(steps, warmup, n_gpus, per_config, parity_checks=True):
    import torch
    import libssa_b200 as L
    from libssa_b200 import dist as ldist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - libssa_b200 has no CPU fallback")
    lib = L.load()
    # rank 0 generates the databases BEFORE any CUDA/NCCL initialisation (the generator forks worker processes)
    if int(os.environ.get("RANK", "0")) == 0:
        protein_db_file(HEAD["db_seqs"], HEAD["seed"])
        if n_gpus == 1 and per_config:
            protein_db_file(CFG1["db_seqs"], CFG1["seed"])
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
    path, meta = protein_db_file(HEAD["db_seqs"], HEAD["seed"], build=False)
    lib.set_thread_count(1)
    t0 = time.perf_counter()
    setup_protein(lib, path)
    t_load = time.perf_counter() - t0
    t0 = time.perf_counter()
    lib.prepare_db()
    t_prepare = time.perf_counter() - t0
    st = lib.stats()
    qlens = qlens_of(HEAD)
    cells_per_step = sum(qlens) * len(HEAD["bit_widths"]) * meta["residues"]          # whole job, all ranks

    m = run_protein_workload(lib, L, HEAD, path, meta, steps, warmup, barrier, max_over_ranks, rank, local, n_gpus)
    acc = m["acc"]
    value = cells_per_step * steps / m["device_s"] / 1e9
    e2e = cells_per_step * steps / m["wall"] / 1e9

    # ---- untimed parity checks: sharded == unsharded, and the compiled reference (one thread) on the full database ----
    parity = {}
    if not parity_checks:
        parity["skipped"] = "--no-parity"
    q0 = lib.init_sequence_fasta(L.READ_FROM_FILE, os.path.join(DATA, HEAD["queries"][0]))
    sharded = {(bw, k): keyed(lib.sw_align(q0, k, bw, L.COMPUTE_SCORE)) for bw in (8, 16) for k in (10, 100)}
    barrier()
    if rank == 0 and parity_checks:
        if world > 1:
            lib.set_shard(0, 1)          # rank 0 alone re-searches the whole database on its GPU
            unsharded = {key: keyed(lib.sw_align(q0, key[1], key[0], L.COMPUTE_SCORE)) for key in sharded}
            parity["sharded_equals_unsharded"] = bool(unsharded == sharded)
            lib.set_shard(rank, world)
        else:
            parity["sharded_equals_unsharded"] = None
    lib.free_sequence(q0)

    line = None
    if rank == 0:
        clocks = m["clocks"]
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        cores = os.cpu_count() or 1
        ref = None
        try:
            ref = reference_api()
        except Exception as e:
            parity["reference_error"] = repr(e)
        if ref is not None and parity_checks:
            ref.set_thread_count(1)
            setup_protein(ref, path)
            rq = ref.init_sequence_fasta(0, os.path.join(DATA, HEAD["queries"][0]))
            t0 = time.perf_counter()
            want = {(8, 10): keyed(ref.sw_align(rq, 10, 8, 0)), (16, 100): keyed(ref.sw_align(rq, 100, 16, 0))}
            parity["hit_lists_equal_reference"] = bool(all(sharded[k] == v for k, v in want.items()))
            parity["reference_run"] = "%s, bit width 8 (top 10) and 16 (top 100), one thread, full %d-record database, %.1f s" % (
                HEAD["queries"][0], meta["n_seqs"], time.perf_counter() - t0)
            ref.free_sequence(rq)
        local_cells = cells_per_step / n_gpus * steps
        traffic = None
        prof = os.path.join(ROOT, "profiles", "roofline_r02.json")
        if not os.path.exists(prof):
            prof = os.path.join(ROOT, "profiles", "roofline_r01.json")
        if os.path.exists(prof):
            traffic = json.load(open(prof)).get("dram_bytes_per_launch")
        # the stream is read once per pass: 2 bytes per pair column ~ 1 byte per residue; 4 bytes per score written
        stream_bytes = (sum(((q + 1023) // 1024) for q in qlens) * (meta["residues"] / n_gpus) * 1.0
                        + 4.0 * len(qlens) * meta["n_seqs"] / n_gpus) * len(HEAD["bit_widths"])
        peaks = {}
        pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(pk):
            peaks = json.load(open(pk))
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        searches = m["searches"]
        line = {
            "metric": "GCUPS", "value": round(value, 2), "unit": "GCUPS", "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
            "ms_per_step": round(m["device_s"] / steps * 1e3, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "int16", "data": "synthetic",
            "config": workload_config(HEAD, n_gpus, meta),
            "arithmetic": "int16 x 2 SIMD-in-word (DPX), exact int32 tier for records whose scores leave the int16 range",
            "parallelism": "db-shard x%d (one process per GPU; ncclAllGather of the shard scores inside libssa_b200.so)" % n_gpus,
            "db_records_per_gpu": meta["n_seqs"] // n_gpus,
            "e2e": {"value": round(e2e, 2), "unit": "GCUPS", "ms_per_step": round(m["wall"] / steps * 1e3, 3),
                    "h2d_bytes_per_step": acc.h2d // steps, "d2h_bytes_per_step": acc.d2h // steps,
                    "note": "wall clock around the libssa.h sw_align calls (host query in, host hit list out); database resident (as in the "
                            "reference's own timing, benchmark/src/benchmark_util.c:29-53); cold start reported in `cold_start`"},
            "gpu_launches": acc.launches,
            "clocks": clocks,
            "roofline": roofline_block(local_cells, m["kernel_s"], sm_mhz, {
                "traffic": traffic,
                "traffic_note": "ncu dram bytes per launch of the dominant kernel, sw32_kernel<32, true, false>: ALL five full passes "
                                "(5 x 1024 rows) of the 5147-aa query over the whole database in one launch = the column stream five times "
                                "(0.94 GB) plus the pass-boundary rows every pass writes for the next one and reads from the previous one "
                                "(16 bytes per column pair and pass); 52 GB/s - HBM is not the bound (profiles/prof_r02_top.md)",
                "per_search_kernel_gcups": {"%s_bw%d" % (HEAD["queries"][qi], bw):
                                            (round(qlens[qi] * meta["residues"] / n_gpus * steps / (ms / 1e3) / 1e9, 1) if ms > 0 else None)
                                            for (bw, qi), ms in zip(searches, acc.kernel_ms_by_search)},
                "hbm": {"algorithmic_bytes_per_step": int(stream_bytes), "achieved_gbs": round(stream_bytes * steps / m["kernel_s"] / 1e9, 2),
                        "peak_gbs": hbm_peak, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback",
                        "note": "not the bound: %.0f cell updates per database byte" % (sum(qlens) / len(qlens))},
            }),
            "overflow": {"reference_8bit_overflows_per_step": acc.ovf8 // steps, "reference_16bit_overflows_per_step": acc.ovf16 // steps,
                         "wide_tier_records_per_step": acc.wide // steps},
            "cold_start": {"fasta_load_s": round(t_load, 3), "pack_and_upload_s": round(t_prepare, 3), "pack_ms": round(st.pack_ms, 1),
                           "upload_ms": round(st.upload_ms, 1), "db_device_bytes": st.db_device_bytes},
            "top_hits": [t[0] if t else None for t in m["top"]],
            "parity_check": parity,
        }
        if n_gpus == 1 and ref is not None:
            try:   # bounded CPU leg: the same database and step, all host threads, 2 timed steps
                g, s_step = reference_protein_steps(ref, HEAD, path, meta, 2, 1, cores)
                line["cpu_baseline"] = {"value": round(g, 3), "unit": "GCUPS", "cores": cores, "kind": "reference",
                                        "sample": "the full workload (all %d records, same 4 searches per step), 2 timed steps of %.1f s after "
                                                  "paging the database in with one search" % (meta["n_seqs"], s_step)}
            except Exception as e:  # the CPU leg must never take the GPU number down with it
                line["cpu_baseline"] = {"unavailable": repr(e)}
        elif n_gpus == 1:
            line["cpu_baseline"] = {"unavailable": "oracle/_ref/libssa_ref.so was not built"}
        if n_gpus == 1 and per_config:
            line["per_config"] = []
            try:
                p1, meta1 = protein_db_file(CFG1["db_seqs"], CFG1["seed"], build=False)
                setup_protein(lib, p1)
                lib.prepare_db()
                m1 = run_protein_workload(lib, L, CFG1, p1, meta1, 3, 2, barrier, max_over_ranks, rank, local, 1)
                cells1 = sum(qlens_of(CFG1)) * meta1["residues"]
                a1 = m1["acc"]
                rec1 = {"config": "configs[1]", "workload": workload_config(CFG1, 1, meta1)["workload"], "steps": 3, "warmup": 2,
                        "value": round(cells1 * 3 / m1["device_s"] / 1e9, 1), "unit": "GCUPS",
                        "e2e": {"value": round(cells1 * 3 / m1["wall"] / 1e9, 1), "unit": "GCUPS", "h2d_bytes_per_step": a1.h2d // 3,
                                "d2h_bytes_per_step": a1.d2h // 3},
                        "gpu_launches": a1.launches,
                        "roofline": roofline_block(cells1 * 3, m1["kernel_s"], sm_mhz, {
                            "per_query_kernel_gcups": [round(q * meta1["residues"] * 3 / (ms / 1e3) / 1e9, 1) if ms > 0 else None
                                                       for q, ms in zip(qlens_of(CFG1), a1.kernel_ms_by_search)]}),
                        "overflow": {"reference_8bit_overflows_per_step": a1.ovf8 // 3, "wide_tier_records_per_step": a1.wide // 3},
                        "top_hits": [t[0] if t else None for t in m1["top"]]}
                if ref is not None:
                    g, s_step = reference_protein_steps(ref, CFG1, p1, meta1, 2, 1, cores)
                    rec1["cpu_baseline"] = {"value": round(g, 3), "unit": "GCUPS", "cores": cores, "kind": "reference",
                                            "sample": "the full workload (all %d records, 5 queries), 2 timed steps of %.1f s" % (meta1["n_seqs"], s_step)}
                line["per_config"].append(rec1)
            except Exception as e:
                line["per_config"].append({"config": "configs[1]", "error": repr(e)})
            for fn in (config2_record, config4_record):
                try:
                    line["per_config"].append(fn(lib, L, ref, sm_mhz, cores))
                except Exception as e:
                    line["per_config"].append({"config": fn.__name__, "error": repr(e)})
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
    ap.add_argument("--no-per-config", action="store_true", help="N = 1: skip the configs[1]/[2]/[4] sub-records")
    ap.add_argument("--no-parity", action="store_true", help="development runs: skip the untimed parity checks")
    args = ap.parse_args()
    if args.impl == "reference":
        line = run_reference(args.steps, args.warmup, args.gpus)
    else:
        line = run_ours(args.steps, max(args.warmup, 0), args.gpus, not args.no_per_config, not args.no_parity)
    if line is not None:
        sys.stdout.flush()
        os.write(real_stdout, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
