"""Pretty-print bench.py JSON lines (helper for reading gpurun_out logs)."""
import json
import sys

for path in sys.argv[1:]:
    for line in open(path):
        line = line.strip()
        if not line.startswith("{"):
            continue
        d = json.loads(line)
        if d.get("impl") == "reference":
            print(path, "REFERENCE", f"{d['value']:.3e}", d["cpu_baseline"]["sample"])
            continue
        print(path, d["config"]["name"], "gpus", d["n_gpus"], "ms/step", round(d["ms_per_step"], 2), "value", f"{d['value']:.3e}",
              "e2e", f"{d['e2e']['value']:.3e}", "tc", d.get("tensor_cores"), "TF useful", round(d.get("step_tflops_useful", 0), 1),
              "launches", d.get("gpu_launches"))
        print("   kernels ms/step:", d.get("kernel_ms_per_step"))
        print("   roofline:", d.get("roofline"))
        print("   clocks:", d.get("clocks"), " cpu:", d.get("cpu_baseline", {}).get("value"))
