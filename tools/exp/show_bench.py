"""Print the few numbers of a bench.py JSON line that an A/B run compares."""
import json
import sys

for f in sys.argv[1:]:
    d = json.loads(open(f).read().strip().splitlines()[-1])
    t = d["config"]["tuning"]
    print(f, "value", d["value"], "ms/step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"],
          "achieved", d["roofline"]["achieved"], d["clocks"])
