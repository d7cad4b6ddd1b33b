"""Generates tests/golden/frames.json: what the compiled reference produces for the strand / reading-frame
workloads of tests/frames.py (one reference process per mode, see tests/frames_ref_worker.py).
Run in the build container (needs oracle/_ref/libssa_ref.so):  python tests/golden/make_frames_golden.py"""
import json
import os
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import frames as F  # noqa: E402

out = {}
with tempfile.TemporaryDirectory() as tmp:
    for (symtype, strands), name in zip(F.MODES, F.MODE_IDS):
        path = os.path.join(tmp, name + ".json")
        F.run_reference_worker(symtype, strands, path)
        out[name] = json.load(open(path))
with open(os.path.join(HERE, "frames.json"), "w") as f:
    json.dump(out, f, indent=0, separators=(",", ":"))
print("wrote", len(out), "modes")
