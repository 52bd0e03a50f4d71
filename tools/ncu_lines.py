"""Summarise an .ncu-rep per CUDA source line: executed warp-instructions and stall samples.
usage: python tools/ncu_lines.py report.ncu-rep [kernel-index]"""
import csv, subprocess, sys, collections, io

rep = sys.argv[1]
which = int(sys.argv[2]) if len(sys.argv) > 2 else 0
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
# sections start with "File Path"; a kernel may span several files
kernels = collections.OrderedDict()
cur_file = cur_fn = None
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Function Name": cur_fn = r[1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None: continue
    kernels.setdefault(cur_fn, []).append((cur_file, hdr, r))
names = list(kernels)
fn = names[min(which, len(names) - 1)]
per_line = collections.Counter(); samples = collections.Counter(); src = {}
tot = 0
line = None
for f, h, r in kernels[fn]:
    ie = h.index("Instructions Executed"); sm = h.index("# Samples")
    if r[0].strip():
        line = (f.split("/")[-1], int(r[0])); src[line] = r[1]
    if len(r) > ie and r[ie].isdigit():
        per_line[line] += int(r[ie]); tot += int(r[ie])
        if r[sm].isdigit(): samples[line] += int(r[sm])
print(fn[:100]); print("total warp-instructions", tot)
ts = sum(samples.values()) or 1
for line, v in per_line.most_common(45):
    print(f"{line[0]:>14s}:{line[1]:<4d} {100*v/tot:5.1f}% inst {100*samples[line]/ts:5.1f}% samples | {src.get(line,'')[:110]}")
