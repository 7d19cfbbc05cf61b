"""Measures the chip's FP64 peak (DMMA / DFMA) on the GPU box and writes gpurun_out/fp64_peak.json; the committed copy
is profiles/fp64_peak.json, which bench.py reads for the roofline of the reduced-system factorisation."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = os.path.join(ROOT, "tools", "ubench", "fp64_peak.cu")
exe = os.path.join(ROOT, "tools", "ubench", "fp64_peak")
subprocess.check_call(["nvcc", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-o", exe, src])
out = subprocess.check_output([exe], text=True).strip().splitlines()[-1]
d = json.loads(out)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(d, open(os.path.join(ROOT, "gpurun_out", "fp64_peak.json"), "w"), indent=1)
print(json.dumps(d))
