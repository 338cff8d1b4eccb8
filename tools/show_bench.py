"""Print the headline figures of a bench.py JSON line: python tools/show_bench.py file.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "frac", round(d["roofline"]["frac"], 4), d["roofline"].get("frac_by_n"))
for k, v in d.get("other_configs", {}).items():
    if isinstance(v, dict):
        print(" ", k, v.get("bo_iterations_per_s") or v.get("seconds_per_run") or v.get("eager") or v.get("kernels_ms"))
