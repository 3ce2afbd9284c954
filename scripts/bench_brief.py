"""Pretty-print the last JSON line of a bench.py run read from stdin."""
import json
import sys
tag = sys.argv[1] if len(sys.argv) > 1 else ""
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])
ks = {k: round(v["ms_per_launch"] * v["launches"] / d["steps"], 4) for k, v in d.get("kernels", {}).items()}
print(tag, "pairs/s", round(d["value"]), "ms/step", round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), ks)
