"""Print ms/step, e2e ms/step, roofline.frac and stage times of bench logs: python tools/bench_summary.py gpurun_out/x/*.log"""
import json
import sys

for f in sys.argv[1:]:
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            if "roofline" not in d:
                print(f, d.get("value"), d.get("unit"))
                continue
            print(f, "ms %.2f" % d["ms_per_step"], "e2e %.2f" % d["e2e"]["ms_per_step"], "frac %.3f" % d["roofline"]["frac"],
                  "fit %.2f" % d["stages_ms"]["fit"], "pf %.3f" % d["stages_ms"]["prefilter"], "warps", d["launch"]["warps_per_cta"],
                  "regs", d["launch"]["regs_per_thread"])
