"""Summarise an ncu report of the ensemble kernel: headline metrics + SASS regions grouped by
lane utilisation (RK / jump / refill sections).  Usage: python scripts/ncu_regions.py rep.ncu-rep"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
det = subprocess.run(["ncu", "-i", rep, "--page", "details"], capture_output=True, text=True).stdout
keys = ("Duration", "Registers Per", "Theoretical Occ", "Achieved Occ", "Executed Ipc Active", "Issue Slots Busy",
        "Warp Cycles Per Issued", "Avg. Active Threads", "Avg. Not Predicated", "No Eligible", "Eligible Warps",
        "DRAM Throughput", "Block Limit Registers", "fused and", "L2 Hit")
for ln in det.splitlines():
    if any(k in ln for k in keys):
        print(ln.strip()[:150])
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
if len(rr) >= 3:
    h = rr[0]
    for name in ("sm__inst_executed_pipe_fp64.sum", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
                 "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
                 "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
                 "sm__cycles_elapsed.max", "sm__inst_executed_pipe_fma.sum", "sm__inst_executed_pipe_alu.sum",
                 "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_uniform.sum"):
        if name in h:
            print(name, rr[2][h.index(name)], rr[1][h.index(name)])
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ci = {n: i for i, n in enumerate(hdr)}
tot = sum(int(r[ci["Instructions Executed"]]) for r in data)
tthr = sum(int(r[ci["Thread Instructions Executed"]]) for r in data)
smp = sum(int(r[ci["# Samples"]]) for r in data)
print(f"SASS instrs {len(data)}  warp-instr {tot:.4g}  avg threads {tthr / tot:.2f}")
regions, cur = [], None
for k, r in enumerate(data):
    ie, te = int(r[ci["Instructions Executed"]]), int(r[ci["Thread Instructions Executed"]])
    if ie == 0:
        continue
    avg = te / ie
    tok = r[ci["Source"]].split()
    op = tok[1] if tok[0].startswith("@") else tok[0]
    if cur and abs(cur["last"] - ie) <= 0.15 * max(ie, cur["last"]) and abs(cur["avg"] - avg) < 3:
        cur["n"] += 1; cur["ie"] += ie; cur["te"] += te; cur["last"] = ie; cur["ops"].append(op); cur["s"] += int(r[ci["# Samples"]])
    else:
        cur = dict(start=k, n=1, ie=ie, te=te, last=ie, avg=avg, ops=[op], s=int(r[ci["# Samples"]]))
        regions.append(cur)
for rg in regions:
    if rg["ie"] / tot < 0.005:
        continue
    ops = collections.Counter(o.split(".")[0] for o in rg["ops"]).most_common(7)
    print(f"@{rg['start']:5d} n={rg['n']:4d} issue={100 * rg['ie'] / tot:5.1f}% stall-samples={100 * rg['s'] / smp:5.1f}% "
          f"exec/instr={rg['ie'] / rg['n']:.3g} avg-thr={rg['te'] / rg['ie']:.1f} {ops}")
