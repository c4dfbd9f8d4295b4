#!/bin/bash
# e2e of the drop-in call at N ranks for several pack fractions (run under gpurun --gpus N)
N=$1; shift
mkdir -p gpurun_out
for f in "$@"; do
  SUCHTREE_B200_PACK_FRACTION=$f timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 10 --warmup 3 --no-other-workloads --no-cpu-baseline > gpurun_out/packfrac_n${N}_$f.json 2> gpurun_out/packfrac_n${N}_$f.err
  python - <<PY
import json
d=json.load(open("gpurun_out/packfrac_n${N}_$f.json"))
e=d["e2e"]; print("N=$N frac=$f e2e %.3e roofline %.3e variants %s" % (e["value"], e["roofline"]["peak_pairs_per_s"], {k: "%.3e" % v for k, v in e["variants_pairs_per_s"].items()}))
PY
done
