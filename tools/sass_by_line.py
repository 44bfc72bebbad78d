"""Static SASS instruction count per source line of one kernel of the in-tree library.
usage: python tools/sass_by_line.py <mangled kernel substring> [first_line last_line [file]]   (prints the SASS of that range too)"""
import collections
import os
import re
import subprocess
import sys
import tempfile

lib = os.environ.get("XPLIB") or os.path.join(os.path.dirname(__file__), "..", "xpore_b200", "libxpore_b200.so")
kern = sys.argv[1]
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (0, 0)
fname = sys.argv[4] if len(sys.argv) > 4 else "xp_fit.cuh"
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
    dis = ""
    for c in os.listdir(td):
        if c.endswith(".cubin"):
            dis += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, c)], capture_output=True, text=True).stdout
infn = False
cur = None
cnt = collections.Counter()
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
    if m and cur:
        cnt[cur] += 1
        if cur[0] == fname and lo <= cur[1] <= hi:
            print("%s:%d  %s" % (cur[0], cur[1], m.group(2)))
    elif l.strip().startswith(".L_") and lo:
        print(l.strip())
if not lo:
    for k, v in sorted(cnt.items()):
        print("%s:%d %d" % (k[0], k[1], v))
    print("total", sum(cnt.values()))
