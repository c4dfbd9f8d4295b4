#!/bin/bash
# e2e of the drop-in call on one GPU for several pack fractions (run under gpurun)
mkdir -p gpurun_out
for f in "$@"; do
  SUCHTREE_B200_PACK_FRACTION=$f timeout 300 python bench.py --steps 20 --warmup 3 --no-other-workloads --no-cpu-baseline > gpurun_out/packfrac_n1_$f.json 2> gpurun_out/packfrac_n1_$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/packfrac_n1_$f.json"))
e=d["e2e"]; print("N=1 frac=$f e2e %.3e roofline %.3e variants %s" % (e["value"], e["roofline"]["peak_pairs_per_s"], {k: "%.3e" % v for k, v in e["variants_pairs_per_s"].items()}))
PY
done
