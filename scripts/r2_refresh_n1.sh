#!/bin/bash
# ncu capture for roofline.traffic + the N=1 bench line (run under gpurun)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pairs -s 3 -c 1 -f \
    -o gpurun_out/r02_k_pairs_yule_final2 python scripts/ncu_target.py pairs yule 100000000 > gpurun_out/r2z_ncu1.log 2>&1
python scripts/update_traffic.py gpurun_out/r02_k_pairs_yule_final2.ncu-rep 100000000 gpurun_out/traffic.json | cut -c1-160
timeout 600 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
tail -c 300 gpurun_out/r2z_bench.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2z_bench.json"))
e = d["e2e"]
print("value %.4e  e2e %.4e  first %.3f  frac %.3f  traffic %s" % (d["value"], e["value"], e["first_call_s"], e["roofline"]["frac"], d["roofline"].get("traffic")))
print({k: e[k] for k in ("h2d_bytes_per_step", "h2d_bytes_over_pcie_per_step", "d2h_bytes_per_step", "pack_fraction", "id_bits")}, e["roofline"]["packed_bound"]["frac"])
PY
