"""profiles/traffic.json from an `ncu --set full` capture of the pair kernel:

    python scripts/update_traffic.py gpurun_out/<capture>.ncu-rep <pairs per launch> [copy-to]

Stamps the figure with st_pairs_kernel_id() of the library in the tree (hash of the pair
kernel's sources and compiler flags): bench.py only reports `roofline.traffic` when the stamp
matches the library it is running."""
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, REPO)
import ncu_summarize  # noqa: E402
from suchtree_b200 import build as B  # noqa: E402


def main():
    rep, pairs = sys.argv[1], int(sys.argv[2])
    recs = [r for r in ncu_summarize.summarize(rep) if "k_pairs" in r["kernel"]]
    r = recs[0]
    name = os.path.basename(rep)[:-len(".ncu-rep")]
    out = {
        "k_pairs_dram_bytes_per_launch": int(r["dram__bytes_read.sum"] + r["dram__bytes_write.sum"]),
        "pairs_per_launch": pairs, "kernel": r["kernel"], "pairs_kernel_id": B.pairs_kernel_id(),
        "capture": "profiles/%s_ncu_details.txt" % name,
        "dram_bytes_read": r["dram__bytes_read.sum"], "dram_bytes_write": r["dram__bytes_write.sum"],
        "ncu_duration_ms": 1e3 * r["gpu__time_duration.sum"],
        "how": "ncu --set full --clock-control none -k regex:k_pairs -s 3 -c 1 python scripts/ncu_target.py pairs yule %d" % pairs,
    }
    path = os.path.join(REPO, "profiles", "traffic.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    if len(sys.argv) > 3:
        shutil.copyfile(path, sys.argv[3])
    print(json.dumps(out))


if __name__ == "__main__":
    main()
