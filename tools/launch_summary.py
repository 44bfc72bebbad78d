"""Summarise an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`): this repository's
kernels with their durations, and the shares of the last complete step (prefilter .. stats).
usage: python tools/launch_summary.py launches.csv [bench log of the same build] > profiles/x.summary.txt"""
import csv
import json
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "")
    ms = v / 1e6 if unit.startswith("ns") else v / 1e3 if unit.startswith("us") else v if unit.startswith("ms") else v * 1e3
    rows.append((int(r["ID"]), r["Kernel Name"], r["Block Size"], r["Grid Size"], ms))
ours = [r for r in rows if "xp_" in r[1]]
print("# launches of this repository's kernels in `ncu --metrics gpu__time_duration.sum --clock-control none -c 400`")
print("# over `python bench.py --steps 2 --warmup 1 --no-cpu-baseline` (c2, 1M positions); the other launches of the list")
print("# are the torch kernels of the synthetic generator.  Per-launch times under ncu are serialised and cold-cache:")
print("# compare SHARES with bench.py's stages_ms, not absolutes.")
for i, k, b, g, ms in ours:
    print("%4d %-44s block %-14s grid %-16s %10.4f ms" % (i, k.split("(")[0][-44:], b, g, ms))
# last complete step: from the last prefilter that is followed by a stats kernel
idx = [j for j, r in enumerate(ours) if "xp_stats_kernel" in r[1]]
if idx:
    end = idx[-1]
    start = max(j for j in range(end) if "xp_prefilter_kernel" in ours[j][1])
    if start > 0 and "xp_prefilter_u16_kernel" in ours[start - 1][1]:  # u16 batches: the pair of prefilter launches
        start -= 1
    step = ours[start:end + 1]
    tot = sum(r[4] for r in step)
    print("# the last complete step in this list is launches %d-%d:" % (step[0][0], step[-1][0]))
    for i, k, b, g, ms in step:
        print("#   %-36s %9.4f ms  %5.1f%%" % (k.split("(")[0].replace("void ", "").replace("xpk::", ""), ms, 100 * ms / tot))
    line = "#   total %.4f ms" % tot
    if len(sys.argv) > 2:
        for l in open(sys.argv[2]):
            if l.startswith("{"):
                st = json.loads(l)["stages_ms"]
                line += "; bench.py stages_ms of the same build: " + ", ".join("%s %.3f" % kv for kv in st.items()) + \
                        " ms (fit share %.1f%%)" % (100 * st["fit"] / sum(st.values()))
    print(line)
