"""Summary of an ncu launch list (long CSV of `--metrics ... --csv --log-file`): per kernel the launches, time share,
time-weighted tensor-pipe activity, SM-active share of the elapsed cycles, tensor-core shared-memory wavefront share, DRAM
and L2 bytes.      python tools/ncu_launch_summary.py gpurun_out/r2_ncu_launches.csv > profiles/r2_ncu_launch_summary.txt"""
import csv
import re
import sys
from collections import defaultdict

rows = csv.DictReader(l for l in open(sys.argv[1]) if l.startswith('"'))
launch = defaultdict(dict)
for r in rows:
    v = float(r["Metric Value"].replace(",", "") or 0)
    u = r["Metric Unit"].lower()
    v *= {"kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "usecond": 1e3, "msecond": 1e6, "us": 1e3, "ms": 1e6}.get(u, 1)
    launch[int(r["ID"])]["name"] = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "")
    launch[int(r["ID"])][r["Metric Name"]] = v
T = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
TC = "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"
agg = defaultdict(lambda: defaultdict(float))
for L in launch.values():
    a = agg[L["name"]]
    t = L.get("gpu__time_duration.sum", 0.0)
    a["n"] += 1
    a["ns"] += t
    a["tensor_w"] += t * L.get(T, 0.0)
    a["tc_w"] += t * L.get(TC, 0.0)
    a["active"] += L.get("sm__cycles_active.avg", 0.0)
    a["elapsed"] += L.get("sm__cycles_elapsed.avg", 0.0)
    a["dram"] += L.get("dram__bytes_read.sum", 0.0) + L.get("dram__bytes_write.sum", 0.0)
    a["l2"] += L.get("lts__t_bytes.sum", 0.0)
tot = sum(a["ns"] for a in agg.values())
print(f"# {len(launch)} launches of one step, {tot / 1e6:.3f} ms summed (ncu: cold caches, serialised - compare SHARES)")
print(f"{'kernel':42s} {'n':>4s} {'ms':>8s} {'share':>6s} {'tensor%':>8s} {'active/elapsed':>14s} {'tc smem%':>9s} {'DRAM GB':>8s} {'L2 GB':>7s}")
conv = defaultdict(float)
for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["ns"]):
    print(f"{name[:42]:42s} {int(a['n']):4d} {a['ns'] / 1e6:8.3f} {100 * a['ns'] / tot:5.1f}% {a['tensor_w'] / max(a['ns'], 1):8.1f} "
          f"{a['active'] / max(a['elapsed'], 1):14.2f} {a['tc_w'] / max(a['ns'], 1):9.1f} {a['dram'] / 1e9:8.2f} {a['l2'] / 1e9:7.2f}")
    if "conv_igemm" in name:
        for k in ("ns", "tensor_w", "dram", "n", "active", "elapsed"):
            conv[k] += a[k]
print(f"# conv_igemm* launches: {int(conv['n'])}, {conv['ns'] / 1e6:.3f} ms = {100 * conv['ns'] / tot:.1f} % of the step, time-weighted tensor-pipe active "
      f"{conv['tensor_w'] / conv['ns']:.1f} % (of the SM-active cycles; SMs active {100 * conv['active'] / conv['elapsed']:.0f} % of the launches' elapsed cycles), "
      f"DRAM traffic {conv['dram'] / 1e9:.2f} GB")
