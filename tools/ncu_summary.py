"""Key metrics of an .ncu-rep (first matching kernel): usage: python tools/ncu_summary.py report.ncu-rep"""
import csv, io, re, subprocess, sys
txt = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt))); hdr, units = rows[0], rows[1]
want = re.compile(r"gpu__time_duration.sum|dram__bytes_(read|write).sum$|sm__throughput.avg.pct|sm__inst_executed_pipe_(fma|fmaheavy|fmalite|alu|xu|lsu)\.avg.pct_of_peak_sustained_active|"
                  r"smsp__inst_executed.sum$|sm__warps_active.avg.pct_of_peak|launch__registers_per_thread$|launch__occupancy_limit|smsp__issue_active.avg.pct|"
                  r"smsp__thread_inst_executed_per_inst_executed.ratio|smsp__average_warps_issue_stalled.*ratio|launch__grid_size|launch__block_size|"
                  r"launch__shared_mem_per_block_dynamic|smsp__cycles_active.avg$|sm__cycles_elapsed.avg$|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|smsp__inst_executed_pipe_fma")
for r in rows[2:3]:
    print("kernel:", r[hdr.index("Kernel Name")])
    for h, u, v in zip(hdr, units, r):
        if want.search(h):
            print(f"{h:88s} {u:16s} {v}")
