"""Print the per-kernel breakdown bench.py leaves in gpurun_out/bench_detail.json (development helper)."""
import json, os, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
d = json.load(open(os.path.join(root, "gpurun_out", "bench_detail.json")))
tot = sum(v["ms_per_step"] for v in d.values())
for k, v in list(d.items())[:n]:
    tf = f"{v['tflops']:.0f}" if v["tflops"] else "-"
    print(f"{k:34s} n={v['launches_per_step']:5.1f} ms={v['ms_per_step']:8.3f} TF={tf}")
print("sum of timed kernels per step: %.1f ms" % tot)
try:
    b = json.load(open(os.path.join(root, "gpurun_out", "bench_line.json")))
    print("frames/s", round(b["value"], 2), "ms/step", round(b["ms_per_step"], 1), "e2e", round(b["e2e"]["value"], 2))
    print({k: (round(v["ms"] / b["steps"], 1), round(v["tflops"])) for k, v in b["roofline"]["by_kind"].items()})
except Exception as e:
    print("no bench line:", e)
