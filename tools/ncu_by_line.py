"""Join an ncu report's per-SASS-instruction counts with source lines (nvdisasm -g on the in-tree .so).

usage: python tools/ncu_by_line.py report.ncu-rep <mangled kernel substring> [lib.so]
Prints executed warp-instructions, FP64 share and stall samples per source line, hottest first.
Only valid when the .so is the build that was profiled (addresses are matched by offset).
"""
import csv
import io
import re
import subprocess
import sys
import tempfile
import collections
import os


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    lib = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(__file__), "..", "xpore_b200", "libxpore_b200.so")
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr = rows[1]
    iA, iS, iN, iM = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples")
    data = rows[2:]
    base = int(data[0][iA], 16)
    prof = {int(r[iA], 16) - base: (r[iS].strip(), int(r[iN]), int(r[iM])) for r in data}
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
        cub = [f for f in os.listdir(td) if f.endswith(".cubin")]
        dis = ""
        for c in cub:
            dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, c)], capture_output=True, text=True).stdout
    line_of = {}
    cur = None
    infn = False
    for l in dis.split("\n"):
        if l.startswith(".text."):
            infn = kern in l
            continue
        if not infn:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
        if m:
            line_of[int(m.group(1), 16)] = cur
    tot = sum(v[1] for v in prof.values())
    tots = sum(v[2] for v in prof.values())
    agg = collections.defaultdict(lambda: [0, 0, 0])
    for a, (src, n, s) in prof.items():
        k = line_of.get(a, ("?", 0))
        agg[k][0] += n
        agg[k][2] += s
        if re.match(r"(@!?U?P\d+\s+)?D(FMA|ADD|MUL|SETP)", src):
            agg[k][1] += n
    print(f"total warp-instructions {tot}, samples {tots}")
    top = int(os.environ.get("TOP", "60"))
    for k, (n, f, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{k[0]:>14s}:{k[1]:<5d} inst {n / tot * 100:6.2f}%  fp64 {f / tot * 100:6.2f}%  samples {s / tots * 100:6.2f}%")


if __name__ == "__main__":
    main()
