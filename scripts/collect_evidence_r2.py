"""Copies the evidence of the last scripts/gpu_r2_evidence.sh + gpu_configs.sh 1 pass from gpurun_out/ (scratch) into
profiles/ (tracked) and prints the headline numbers.  Usage: python scripts/collect_evidence_r2.py"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
def regions(rep, out):
    txt = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_regions.py"), os.path.join(G, rep + ".ncu-rep")], capture_output=True, text=True).stdout
    open(os.path.join(P, out), "w").write(txt)
regions("r2_prof_ensemble_kernel", "r2_tcp_kernel_final_4M_full.txt")
regions("r2_prof_compact_kernel", "r2_compact_kernel_final_full.txt")
regions("r2_prof_stats_reduce_kernel", "r2_stats_reduce_kernel_full.txt")
for m in ("morris_lecar", "tcp2d", "sir_rejection"):
    regions("r2_prof_" + m, f"r2_{m}_kernel_final_full.txt")
def json_line(src, dst):
    for ln in open(os.path.join(G, src)):
        if ln.startswith("{"):
            open(os.path.join(P, dst), "w").write(ln)
            return json.loads(ln)
d = json_line("r2_bench_1e7.log", "r2_bench_1e7.json")
ref = json_line("r2_bench_ref.log", "r2_bench_reference_arm.json")
shutil.copy(os.path.join(G, "r2_launches.csv"), os.path.join(P, "r2_launches_bench.csv"))
shutil.copy(os.path.join(G, "r2_pytest_gpu.log"), os.path.join(P, "r2_pytest_gpu.log"))
if os.path.exists(os.path.join(G, "configs_n1.jsonl")):
    shutil.copy(os.path.join(G, "configs_n1.jsonl"), os.path.join(P, "r2_configs_n1.jsonl"))
print({k: d[k] for k in ("value", "ms_per_step", "trajectories_per_sec", "rk_attempts_per_jump", "value_tstop_policy0")})
print(d["e2e"]); print(d["roofline_e2e"]["frac"], d["roofline_e2e"]["peak"])
print({k: d["roofline"][k] for k in ("achieved", "peak", "frac", "kernel_ms_per_step")})
print({k: d["roofline_output"][k] for k in ("achieved", "frac", "compact_ms_per_step")}); print(d["cpu_baseline"]["value"], d["clocks"])
print("REF", ref["value"])
km, cm, ms = d["roofline"]["kernel_ms_per_step"], d["roofline_output"]["compact_ms_per_step"], d["ms_per_step"]
rows = [r for r in csv.reader(open(os.path.join(G, "r2_launches.csv"))) if len(r) > 10 and r[0].isdigit()]
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    agg[r[4]][0] += 1; agg[r[4]][1] += float(r[-1])
tot = sum(v[1] for v in agg.values())
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 --no-policy0",
       "# (per-launch times are serialised/cold-cache; the kernel's SHARE is what must agree with the live CUDA-event split:",
       "#  live bench (profiles/r2_bench_1e7.json): ensemble kernel %.2f ms of %.2f ms/step = %.1f %%, scan+compaction %.2f ms = %.1f %%)" % (km, ms, 100 * km / ms, cm, 100 * cm / ms)]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k[:100]:100s} n={v[0]:4d} total_ms={v[1] / 1e6:9.2f} share={100 * v[1] / tot:5.1f}%")
open(os.path.join(P, "r2_launches_bench_summary.txt"), "w").write("\n".join(out) + "\n")
print("\n".join(out[2:6]))
