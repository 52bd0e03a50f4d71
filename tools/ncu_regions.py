"""Instruction and stall-sample share per code REGION of bc_device.cuh (phase A helpers, candidate arithmetic, ordering,
bound table, bound test, batch evaluation, ...) from the source page of an .ncu-rep captured with --import-source on.
The regions are found by marker comments in the current source, so run it on a capture of the current build.
usage: python tools/ncu_regions.py report.ncu-rep"""
import csv, subprocess, sys, collections, io, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rep=sys.argv[1]
txt=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(txt)))
cur_file=None;hdr=None
per=collections.Counter(); samp=collections.Counter()
line=None
for r in rows:
    if not r: continue
    if r[0]=="File Path": cur_file=r[1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; continue
    if hdr is None: continue
    ie=hdr.index("Instructions Executed"); sm=hdr.index("# Samples")
    if r[0].strip(): line=(cur_file.split("/")[-1], int(r[0]))
    if len(r)>ie and r[ie].isdigit():
        per[line]+=int(r[ie])
        if r[sm].isdigit(): samp[line]+=int(r[sm])
# regions from source file line ranges
src=open(os.path.join(ROOT, 'tileconv_b200', 'csrc', 'bc_device.cuh')).read().splitlines()
def find(pat,start=0):
    for i in range(start,len(src)):
        if pat in src[i]: return i+1
marks=[('helpers/phaseA', 1), ('cluster arithmetic (solve_tail/eval)', find('sm_100 packed binary32 arithmetic')), ('construct_ordering', find('// ConstructOrdering: stable ascending')), ('Q build', find('bound terms of the pruned sweep')), ('run_sweep: head', find('// WALK: step a pointer')), ('run_sweep: bound test', find('// bound test, 64 candidates per step')), ('run_sweep: batch eval', find('      if (waiting == 0) break;')), ('run_sweep: argmin', find('// arg-min over the warp with the lowest candidate')), ('sweep_endpoints+cluster_fit', find('// Endpoints of candidate c of the current ordering'))]
tot=sum(per.values()); ts=sum(samp.values())
reg=collections.Counter(); regs=collections.Counter()
for (f,l),v in per.items():
    if f!='bc_device.cuh': name=f
    else:
        name=None
        for nm,st in marks:
            if st and l>=st: name=nm
    reg[name]+=v; regs[name]+=samp[(f,l)]
for k,v in reg.most_common(): print(f"{k:45s} {100*v/tot:5.1f}% inst {100*regs[k]/ts:5.1f}% samples")
