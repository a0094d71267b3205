#!/bin/bash
# prints value / e2e / clocks / per-layer ms per step of a bench run: tools/bench_layers.sh <out.json> [bench args]
out=$1; shift
python bench.py --steps 10 --warmup 3 --no-cpu-baseline "$@" > "$out" 2> "${out%.json}.err" || tail -5 "${out%.json}.err"
python - "$out" <<'PY'
import json, sys
d = json.load(open(sys.argv[1]))
L = d["layers"]
print(round(d["value"]), round(d["e2e"]["value"]), d["clocks"].get("sm_mhz"),
      {k: round(v["ms_per_launch"] * v["launches"] / d["steps"], 3) for k, v in L.items() if v["launches"]},
      d["parity_spot_check"]["ok"])
PY
