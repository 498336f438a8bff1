#!/bin/bash
# ON THE GPU BOX: bench.py over trunk micro-batch sizes; prints one summary line per chunk.
for c in "$@"; do
  python bench.py --steps 10 --no-cpu --chunk $c > /tmp/b_$c.log 2>&1
  python - "$c" /tmp/b_$c.log <<'PY'
import sys, json
c, path = sys.argv[1], sys.argv[2]
try:
    d = json.loads(open(path).read().strip().splitlines()[-1])
    print("chunk", c, "ms", round(d["ms_per_step"], 3), "samples/s", round(d["value"]), "e2e ms",
          round(d["e2e"]["ms_per_step"], 3), "igemm frac", round(d["roofline"]["frac"], 3), "whole-step frac",
          round(d["roofline"]["whole_step_frac"], 3))
except Exception as e:
    print("chunk", c, "failed", e, open(path).read()[-400:])
PY
done
