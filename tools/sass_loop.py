"""Static dump of the E-step trip loop (the loop holding the first 'magic' DFMA) of one fit-kernel instance.
usage: python tools/sass_loop.py [mangled kernel substring] [-v]"""
import collections
import os
import re
import subprocess
import sys
import tempfile

kern = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("-") else "xp_fit_kernelILi1ELi896"
lib = os.environ.get("XPLIB") or os.path.join(os.path.dirname(__file__), "..", "xpore_b200", "libxpore_b200.so")
with tempfile.TemporaryDirectory() as td:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=td, capture_output=True)
    txt = []
    for c in os.listdir(td):
        if c.endswith(".cubin"):
            txt += subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, c)], capture_output=True, text=True).stdout.split("\n")
s = [i for i, l in enumerate(txt) if l.startswith(".text.") and kern in l][0]
e = s + 1
while e < len(txt) and not txt[e].startswith(".text."):
    e += 1
ins, cur = [], None
for l in txt[s:e]:
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = m.group(2)
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), cur, m.group(2)))
    elif l.strip().startswith(".L_"):
        ins.append((None, None, l.strip()))
i0 = [i for i, x in enumerate(ins) if x[0] is not None and "6.7553994410557440" in x[2] and "DFMA" in x[2]][0]
j = i0
while ins[j][0] is not None:
    j -= 1
k = i0
while not (ins[k][0] is not None and "BRA" in ins[k][2] and ins[j][2].rstrip(":") in ins[k][2]):
    k += 1
body = ins[j:k + 1]
ops = collections.Counter((x[2].split()[1] if x[2].startswith("@") else x[2].split()[0]) for x in body[1:])
nf = sum(v for o, v in ops.items() if re.match(r"D(FMA|ADD|MUL|SETP)", o))
print(len(body) - 1, "instructions in the loop,", nf, "FP64 -> dispatch cycles", len(body) - 1 + nf)
print(ops.most_common())
if "-v" in sys.argv:
    for x in body:
        print(x[1], x[2])
