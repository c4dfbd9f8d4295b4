#!/bin/bash
# Final one-GPU evidence pass of the round (run under gpurun): GPU tests, the ncu capture the
# bench's roofline.traffic is stamped from, the bench line, the reference arm, the launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/r2z_pytest.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pairs -s 3 -c 1 -f \
    -o gpurun_out/r02_k_pairs_yule_final2 python scripts/ncu_target.py pairs yule 100000000 > gpurun_out/r2z_ncu1.log 2>&1
python scripts/update_traffic.py gpurun_out/r02_k_pairs_yule_final2.ncu-rep 100000000 gpurun_out/traffic.json | cut -c1-200
timeout 600 python bench.py > gpurun_out/r2z_bench.json 2> gpurun_out/r2z_bench.err
tail -c 300 gpurun_out/r2z_bench.err
if [ -z "$SKIP_REF" ]; then
  timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2z_reference.json 2> gpurun_out/r2z_reference.err
  tail -c 300 gpurun_out/r2z_reference.err
fi
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r02_launches_bench.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --e2e-pairs 8000000 --cfg3-pairs 100000000 --cfg4-samples 10000000 \
    --quartets 10000000 > gpurun_out/r2z_launch.log 2>&1
wc -l gpurun_out/r02_launches_bench.csv
python - <<'PY'
import json
d = json.load(open("gpurun_out/r2z_bench.json"))
e = d["e2e"]
print("value %.4e  e2e %.4e  first %.3f  frac %.3f  traffic %s" % (d["value"], e["value"], e["first_call_s"], e["roofline"]["frac"], d["roofline"].get("traffic")))
o = d["other_workloads"]
print({k: (v.get("samples_per_s") or v.get("link_pairs_per_s") or v.get("pairs_per_s") or v.get("quartets_per_s") or v.get("pairs_per_s_host_call")) for k, v in o.items()})
print(e["roofline"].get("packed_bound"))
PY
