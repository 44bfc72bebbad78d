"""Executed warp-instructions of every loop (backward branch) of the profiled kernel, from an ncu report's SASS page.

usage: python tools/ncu_loops.py report.ncu-rep [min share %]
Nested loops are listed innermost first; a loop's share includes the loops inside it.
"""
import csv
import io
import re
import subprocess
import sys


def main():
    rep = sys.argv[1]
    thr = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    h0 = next(i for i, r in enumerate(rows) if "Address" in r)
    hdr = rows[h0]
    rows = rows[h0 - 1:]
    iA, iS, iN = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
    data = [(int(r[iA], 16), r[iS].strip(), int(r[iN])) for r in rows[2:]]
    base = data[0][0]
    data = [(a - base, s, n) for a, s, n in data]
    total = sum(n for _, _, n in data)
    print("total warp-instructions %d" % total)
    loops = []
    for a, s, n in data:
        if "BRA" in s and "BRA.DIV" not in s:
            m = re.search(r"0x([0-9a-f]+)", s)
            if m:
                tgt = int(m.group(1), 16) - base
                if 0 <= tgt < a:
                    loops.append((tgt, a, n))
    loops.sort(key=lambda x: x[1] - x[0])
    for tgt, a, nb in loops:
        body = [(x, s, n) for x, s, n in data if tgt <= x <= a]
        ex = sum(n for _, _, n in body)
        if 100.0 * ex / total < thr:
            continue
        fp = sum(n for _, s, n in body if re.match(r"(@!?U?P\d+\s+)?D(FMA|ADD|MUL|SETP)", s))
        trips = max(n for _, _, n in body)
        print("loop 0x%05x-0x%05x  static %4d  executed %6.2f%%  fp64 %5.2f%% of total  (hottest instruction ran %d times)"
              % (tgt, a, len(body), 100.0 * ex / total, 100.0 * fp / total, trips))


main()
