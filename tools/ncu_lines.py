"""Attribute ncu per-SASS-instruction counters to CUDA source lines (no GPU needed).

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-substring> [launch-index] [lib.so]
Needs the library built with -lineinfo; matches the i-th SASS instruction of the kernel in the
report with the i-th instruction nvdisasm prints for the same function.
"""
import csv, io, os, re, subprocess, sys, tempfile
from collections import Counter

rep, kname = sys.argv[1], sys.argv[2]
which = int(sys.argv[3]) if len(sys.argv) > 3 else 0
so = sys.argv[4] if len(sys.argv) > 4 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "squlearn_b200", "libb200sq.so")
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
lines = None
for f in sorted(os.listdir(tmp)):
    if not f.endswith(".cubin"):
        continue
    txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
    # split into functions
    parts = re.split(r"\n\s*\.text\.", txt)
    for p in parts:
        head = p.split("\n", 1)[0]
        if kname in head:
            cur = None
            lines = []
            for ln in p.splitlines():
                m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
                if m:
                    cur = (os.path.basename(m.group(1)), int(m.group(2)))
                    continue
                if re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
                    lines.append(cur)
            break
    if lines:
        break
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name" and kname in r[1]]
s = starts[which]
hdr = rows[s + 1]
end = next((i for i in range(s + 2, len(rows)) if rows[i] and rows[i][0] == "Kernel Name"), len(rows))
data = rows[s + 2:end]
ci, wi = hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
print("sass rows", len(data), "nvdisasm instrs", len(lines))
inst, stall = Counter(), Counter()
for r, l in zip(data, lines):
    inst[l] += int(r[ci]); stall[l] += int(r[wi])
ti, ts = sum(inst.values()), sum(stall.values())
print(f"total warp-instructions {ti}, stall samples {ts}")
print("--- by instructions executed")
for l, n in inst.most_common(28):
    print(f"{str(l):36s} inst {100*n/ti:5.1f}%  stall {100*stall[l]/ts:5.1f}%")
print("--- by stall samples")
for l, n in stall.most_common(12):
    print(f"{str(l):36s} stall {100*n/ts:5.1f}%  inst {100*inst[l]/ti:5.1f}%")
