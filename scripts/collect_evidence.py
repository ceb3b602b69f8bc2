"""Copies the evidence of the last scripts/gpu_round4.sh + gpu_configs.sh pass from gpurun_out/ into profiles/
under a version tag and prints the headline numbers.  Usage: python scripts/collect_evidence.py v9"""
import collections, csv, glob, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
for old in glob.glob(os.path.join(P, "r1_tcp_kernel_v[0-9]*_full.txt")) + glob.glob(os.path.join(P, "r1_compact_kernel_v*_full.txt")) + glob.glob(os.path.join(P, "r1_bench_1e7_v*.json")):
    if "_v1_" in old or "_v2_" in old:
        continue
    os.remove(old)
for rep, out in (("prof_tcp_r3", f"r1_tcp_kernel_{tag}_full.txt"), ("prof_compact_r3", f"r1_compact_kernel_{tag}_full.txt")):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_regions.py"), os.path.join(G, rep + ".ncu-rep")], capture_output=True, text=True).stdout
    open(os.path.join(P, out), "w").write(txt)
shutil.copy(os.path.join(G, "launches_r1.csv"), os.path.join(P, "r1_launches_bench.csv"))
shutil.copy(os.path.join(G, "bench_1e7.log"), os.path.join(P, f"r1_bench_1e7_{tag}.json"))
shutil.copy(os.path.join(G, "bench_ref.log"), os.path.join(P, "r1_bench_reference_arm.json"))
if os.path.exists(os.path.join(G, "configs_n1.jsonl")):
    shutil.copy(os.path.join(G, "configs_n1.jsonl"), os.path.join(P, "r1_configs_n1.jsonl"))
for ln in open(os.path.join(G, "bench_1e7.log")):
    if ln.startswith("{"):
        d = json.loads(ln)
        print({k: d[k] for k in ("value", "ms_per_step", "trajectories_per_sec", "rk_attempts_per_jump")})
        print(d["e2e"]); print({k: d["roofline"][k] for k in ("achieved", "peak", "frac", "kernel_ms_per_step")})
        print({k: d["roofline_output"][k] for k in ("achieved", "frac", "compact_ms_per_step")}); print(d["cpu_baseline"]["value"], d["clocks"])
        km, cm, ms = d["roofline"]["kernel_ms_per_step"], d["roofline_output"]["compact_ms_per_step"], d["ms_per_step"]
for ln in open(os.path.join(G, "bench_ref.log")):
    if ln.startswith("{"):
        print("REF", json.loads(ln)["value"])
if os.path.exists(os.path.join(G, "configs_n1.jsonl")):
    for ln in open(os.path.join(G, "configs_n1.jsonl")):
        if ln.startswith("{"):
            d = json.loads(ln)
            if d.get("config") in ("C3", "C4", "C5"):
                print(d["config"], "jumps/s %.4g traj/s %.4g cand/s %.4g step %.2f ms kernel %.2f frac %.3f" % (
                    d["jumps_per_sec"], d["trajectories_per_sec"], d["candidates_per_sec"], d["ms_step_max_over_ranks"], d["kernel_ms"], d["fp64_roofline"]["frac"]))
rows = [r for r in csv.reader(open(os.path.join(G, "launches_r1.csv"))) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    agg[r[4]][0] += 1; agg[r[4]][1] += float(r[-1])
tot = sum(v[1] for v in agg.values())
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1",
       "# (per-launch times are serialised/cold-cache; the kernel's SHARE is what must agree with the live CUDA-event split:",
       "#  live bench (profiles/r1_bench_1e7_%s.json): ensemble kernel %.2f ms of %.2f ms/step = %.1f %%, scan+compaction %.2f ms = %.1f %%)" % (tag, km, ms, 100 * km / ms, cm, 100 * cm / ms)]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k[:100]:100s} n={v[0]:4d} total_ms={v[1] / 1e6:9.2f} share={100 * v[1] / tot:5.1f}%")
open(os.path.join(P, "r1_launches_bench_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[2:5]))
print(open(os.path.join(P, f"r1_tcp_kernel_{tag}_full.txt")).read().split("SASS instrs")[0])
