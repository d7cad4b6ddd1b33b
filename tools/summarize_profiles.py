#!/usr/bin/env python3
"""Turns the raw ncu outputs in gpurun_out/ into the committed summaries under profiles/."""
import collections, csv, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
G = os.path.join(ROOT, sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"); P = os.path.join(ROOT, "profiles")

def launches():
    rows = [r for r in csv.reader(open(os.path.join(G, "launches_%s.csv" % tag))) if len(r) > 10]
    hdr, data = None, []
    for r in rows:
        if r[0] == "ID": hdr = r; continue
        if hdr: data.append(dict(zip(hdr, r)))
    agg = collections.OrderedDict()
    for d in data:
        v = float(d["Metric Value"].replace(",", "")); u = d["Metric Unit"]
        us = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
        k = d["Kernel Name"]
        a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += us
    tot = sum(a[1] for a in agg.values())
    out = ["# ncu launch list of `python bench.py --steps 2 --warmup 1` (%s)" % tag, "",
           "`ncu --metrics gpu__time_duration.sum --clock-control none -c 400` on one B200; raw CSV: `profiles/launches_%s.csv`." % tag,
           "Times under ncu are cold-cache and serialised: compare SHARES.  %d launches, %.1f ms in total." % (len(data), tot / 1e3), "",
           "| kernel | launches | total ms | share |", "|---|---|---|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append("| `%s` | %d | %.2f | %.2f %% |" % (k.replace("|", "/"), a[0], a[1] / 1e3, 100 * a[1] / tot))
    share = sum(a[1] for k, a in agg.items() if "score_kernel<1" in k or "score_kernel<(bool)1" in k or "sw32_kernel" in k) / tot
    out += ["", "Packed scoring kernels (`sw32_kernel<R, MULTIPASS>`, `score_kernel<1, R, ...>`): **%.1f %%** of the GPU time of a step; the wide tier (`score_kernel<0, ...>`), "
            "arrival-order, unpack, flag, scatter and fill kernels share the rest." % (100 * share)]
    open(os.path.join(P, "launches_%s.md" % tag), "w").write("\n".join(out) + "\n")
    import shutil; shutil.copy(os.path.join(G, "launches_%s.csv" % tag), os.path.join(P, "launches_%s.csv" % tag))
    return share

WANT = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max", "smsp__warps_eligible.avg.per_cycle_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio"]

def kernel_report(rep, title, cells_per_launch, note, outname=None):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    out = ["# %s" % title, "", note, ""]
    res = []
    for vals in rows[2:]:
        d = dict(zip(hdr, vals)); u = dict(zip(hdr, units))
        out += ["## launch %s: `%s`" % (d.get("ID"), d.get("Kernel Name")), "", "| metric | value | unit |", "|---|---|---|"]
        for w in WANT:
            if w in d: out.append("| %s | %s | %s |" % (w, d[w], u[w]))
        def num(k):
            v = float(d[k].replace(",", "")); un = u[k]
            return v * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1}.get(un, 1)
        t = num("gpu__time_duration.sum"); dram = num("dram__bytes_read.sum") + num("dram__bytes_write.sum")
        inst = float(d["smsp__inst_executed.sum"].replace(",", ""))
        out += ["", "derived: %.1f GCUPS under ncu (padded rows not counted), DRAM traffic %.1f MB per launch, %.2f warp instructions per packed word-row"
                % (cells_per_launch / t / 1e9, dram / 1e6, inst / (cells_per_launch / 64.0)), ""]
        res.append({"kernel": d.get("Kernel Name"), "seconds": t, "dram_bytes": dram, "instructions": inst})
    open(os.path.join(P, outname or rep.replace(".ncu-rep", ".md")), "w").write("\n".join(out) + "\n")
    return res

def stall_breakdown(rep, outname=None):
    """appends the warp-state samples by reason and by opcode (ncu --page source) of the first captured launch to the summary"""
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep), "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, data = rows[1], []
    for r in rows[2:]:
        if r and r[0] == "Kernel Name": break
        if len(r) < 10 or r[0] == "Address": continue
        data.append(r)
    ia, ie = hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    sc = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot, exe = sum(int(r[ia]) for r in data), sum(int(r[ie]) for r in data)
    reasons = sorted(((sum(int(r[i] or 0) for r in data), hdr[i]) for i in sc), reverse=True)
    agg = collections.defaultdict(lambda: [0, 0])
    for r in data:
        t = r[1].split(); op = t[1] if t[0].startswith("@") else t[0]
        agg[op][0] += int(r[ia]); agg[op][1] += int(r[ie])
    out = ["", "## warp-state samples by reason and by opcode (first launch above; `ncu --page source`, %d samples, %d SASS instructions, "
           "%.2f G warp instructions executed)" % (tot, len(data), exe / 1e9), "", "| reason | share |", "|---|---|"]
    out += ["| %s | %.1f %% |" % (n, 100.0 * v / tot) for v, n in reasons if v * 1000 >= tot]
    out += ["", "| opcode | share of samples | share of executed instructions |", "|---|---|---|"]
    out += ["| `%s` | %.1f %% | %.1f %% |" % (op, 100.0 * sv / tot, 100.0 * ev / exe) for op, (sv, ev) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:14]]
    open(os.path.join(P, outname or rep.replace(".ncu-rep", ".md")), "a").write("\n".join(out) + "\n")


def round1():
        share = launches()
        meta = json.load(open(os.path.join(G, "bench_%s.json" % tag)))
        res_per_gpu = meta["config"]["db_residues"] / meta["n_gpus"]
        # prof_<tag>_top: P33450 (5147 aa) vs the 500k database: one full pass = 32 lanes x 32 rows = 1024 query rows
        top = kernel_report("prof_%s_top.ncu-rep" % tag, "ncu --set full: dominant kernel of the benchmark step (P33450, 5147 aa, one of the 5 full passes of 1024 rows)",
                            res_per_gpu * 1024.0,
                            "`ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k 'regex:score_kernel<.bool.1, .int.32, .bool.0, .int.1' -s 5 -c 2 "
                            "python tools/profile_target.py P33450 sw 500000` (same 500 000-record database shape as bench.py; the capture follows a plain run of the same command that exited 0).")
        stall_breakdown("prof_%s_top.ncu-rep" % tag)
        json.dump({"kernel": top[-1]["kernel"], "dram_bytes_per_launch": top[-1]["dram_bytes"], "seconds_under_ncu": top[-1]["seconds"],
                   "algorithmic_cells_per_launch": res_per_gpu * 1024.0, "algorithmic_bytes_per_launch": res_per_gpu * 1.0 + 4.0 * meta["config"]["db_records"] / meta["n_gpus"],
                   "score_kernel_share_of_step": share, "source": "profiles/prof_%s_top.md" % tag}, open(os.path.join(P, "roofline_%s.json" % tag), "w"), indent=1)
        q375 = kernel_report("prof_%s_q375.ncu-rep" % tag, "ncu --set full: bulk kernel of a 375-aa query (P07327, groups of 16 lanes x 24 rows, two columns per step)",
                             res_per_gpu * 375, "`ncu --set full ... -k 'regex:score_kernel<.bool.1, .int.24,' ... python tools/profile_target.py P07327 sw 500000`; pairs longer than the "
                             "long-class threshold run in a separate (32 x 12) launch and are not part of this launch, so the derived GCUPS slightly overstates.")
        print("share", share, top[-1], q375[-1])


def round2():
    """r02: the dominant kernel is sw32_kernel<32, true> - ALL five full passes (5 x 1024 rows) of P33450 in one launch."""
    share = launches()
    residues = 187296691    # tools/profile_target.py database: synth.write_protein_db(570000, seed 20150402)
    rows = 5 * 1024
    top = kernel_report("prof_%s_P33450.ncu-rep" % tag, "ncu --set full: dominant kernel of the benchmark step (P33450, 5147 aa: the five full passes of 32 lanes x 32 rows in ONE launch)",
                        residues * float(rows),
                        "`ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:sw32_kernel -s 2 -c 1 python tools/profile_target.py P33450 sw 570000` "
                        "(570 000-record database of bench.py's shape, %d residues; the capture follows a plain run of the same command that exited 0: `kernel_ms 155.6 gcups 6195`).  "
                        "The remaining 27 query rows run in a `score_kernel<1, 8, 0, 2, 0>` launch that is not part of this capture." % residues, "prof_%s_top.md" % tag)
    stall_breakdown("prof_%s_P33450.ncu-rep" % tag, "prof_%s_top.md" % tag)
    json.dump({"kernel": top[-1]["kernel"], "dram_bytes_per_launch": top[-1]["dram_bytes"], "seconds_under_ncu": top[-1]["seconds"],
               "algorithmic_cells_per_launch": residues * float(rows), "algorithmic_bytes_per_launch": residues * 1.0 * 5 + 4.0 * 570000,
               "passes_per_launch": 5, "instructions_per_word_row": top[-1]["instructions"] / (residues * float(rows) / 64.0),
               "score_kernel_share_of_step": share, "source": "profiles/prof_%s_top.md" % tag}, open(os.path.join(P, "roofline_%s.json" % tag), "w"), indent=1)
    if os.path.exists(os.path.join(G, "prof_%s_P07327.ncu-rep" % tag)):
        q375 = kernel_report("prof_%s_P07327.ncu-rep" % tag, "ncu --set full: chained kernel on half-warp arrays, 375-aa query (P07327: two arrays of 16 lanes x 24 rows per warp)",
                             residues * 375.0, "`ncu --set full ... -k regex:sw32_kernel -s 2 -c 1 python tools/profile_target.py P07327 sw 570000` "
                             "(plain run first: `kernel_ms 11.83 gcups 5937`).", "prof_%s_q375.md" % tag)
        stall_breakdown("prof_%s_P07327.ncu-rep" % tag, "prof_%s_q375.md" % tag)
        print(q375[-1])
    print("share", share, top[-1])


if __name__ == "__main__":
    (round1 if tag == "r01" else round2)()
