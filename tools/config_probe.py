#!/usr/bin/env python3
"""Development aid: one of bench.py's per-config records on its own (no reference arm, no parity leg).
   python tools/config_probe.py 2|4"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import libssa_b200 as L
lib = L.load()
lib.set_output_mode(L.OUTPUT_SILENT)
which = sys.argv[1] if len(sys.argv) > 1 else "4"
rec = (bench.config2_record if which == "2" else bench.config4_record)(lib, L, None, 1965.0, 1)
rec.pop("workload", None)
rec.get("roofline", {}).pop("peak_formula", None)
rec.get("roofline", {}).pop("peak_source", None)
print(json.dumps(rec))
