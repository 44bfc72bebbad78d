"""Annotated SASS of one kernel from an ncu report: executed warp-instructions per address divided by a unit
count (e.g. position-sweeps), with the source line of the in-tree library (must be the profiled build).
usage: python tools/ncu_annot.py report.ncu-rep <mangled kernel substring> <units> > annot.txt
columns: address, executions per unit, stall samples, file:line, SASS"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, kern, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
lib = os.path.join(os.path.dirname(__file__), "..", "xpore_b200", "libxpore_b200.so")
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = rows[1]
iA, iN, iM = hdr.index("Address"), hdr.index("Instructions Executed"), hdr.index("# Samples")
data = rows[2:]
base = int(data[0][iA], 16)
prof = {int(r[iA], 16) - base: (int(r[iN]), int(r[iM])) for r in data}
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
    dis = ""
    for c in os.listdir(td):
        if c.endswith(".cubin"):
            dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, c)], capture_output=True, text=True).stdout
infn, cur = False, ("?", 0)
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
        a = int(m.group(1), 16)
        p = prof.get(a, (-1, -1))
        print("%06x %8.2f %6d %s:%-4d %s" % (a, p[0] / units, p[1], cur[0], cur[1], m.group(2)))
    elif l.strip().startswith(".L_"):
        print(l.strip())
