import sys, json
d = json.loads([l for l in sys.stdin.read().strip().splitlines() if l.startswith("{")][-1])
print("value %.1f bags/s  ms/step %.2f | streaming %.1f bags/s %.2f ms | e2e %s | launches %s" % (d["value"], d["ms_per_step"], d["streaming_only"]["value"], d["streaming_only"]["ms_per_step"], (d.get("e2e") or {}).get("value"), d.get("gpu_launches")))
print("roofline_streaming", d["roofline_streaming"].get("achieved"), d["roofline_streaming"].get("frac"), d["roofline_streaming"].get("kernel_ms_per_step"))
for k in d["kernels"][:10]: print("  %-26s %8.4f ms  share %.3f  %s %s frac %s" % (k["kernel"], k["ms_per_step"], k["share_of_kernel_time"], k.get("achieved"), k.get("unit"), k.get("frac")))
