#!/bin/bash
# round 2, second session: fp16 activation storage (default) vs bf16, deep TMA rings A/B, full gpu tests with the printed precision numbers
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -s -m gpu > gpurun_out/r2b_gputests.log 2>&1; echo "pytest rc=$?"
grep -E "^E  |passed|failed|FAILED|heads vs reference|decoded|per-stage|mode=|vs .* oracle" gpurun_out/r2b_gputests.log | cut -c1-900 | tail -40
for cfg in "bf16 1" "fp16 0" "fp16 1"; do
  set -- $cfg
  VGGHEADS_B200_DEEP_RINGS=$2 timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline --act-dtype $1 > gpurun_out/r2b_bench_$1_rings$2.json 2> gpurun_out/r2b_bench_$1_rings$2.err; echo "bench $cfg rc=$?"
done
SECONDS=0; timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2b_bench_1gpu_full.json 2> gpurun_out/r2b_bench_1gpu_full.err; echo "full bench rc=$? wall ${SECONDS}s"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2b_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], d["dtype"], round(d["value"]), "img/s e2e", round(d["e2e"]["value"]), "frac_serial", round(d["roofline"]["frac_serial"], 3),
              "frac_step", round(d["roofline"]["frac_step"], 3), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
        if "end_to_end" in d.get("parity", {}):
            e = d["parity"]["end_to_end"]
            for k in e:
                if isinstance(e[k], dict) and "boxes_max_abs_err_px" in e[k]:
                    print("  ", k, {kk: (round(vv, 6) if isinstance(vv, float) else vv) for kk, vv in e[k].items() if kk != "stage_rel_err"}, "stages", [round(v, 5) for v in e[k]["stage_rel_err"].values()])
            print("  dense", d.get("dense_heads"), "cfg4", d.get("config4_shard"))
    except Exception as ex:
        print(f, "unreadable", ex)
PY
