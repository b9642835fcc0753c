"""Development: turn one tools/profile_round.sh pass (gpurun_out/<tag>_*) into the tracked files under profiles/."""
import collections, csv, json, os, shutil, subprocess, sys
tag = sys.argv[1]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
R = tag.split("_")[0]            # round prefix of the tracked files: r1, r2, ...
shutil.copy(f"{G}/{tag}_bench_n1.json", f"{P}/{R}_bench_n1.json")
shutil.copy(f"{G}/{tag}_kbench.txt", f"{P}/{R}_kbench.txt")
shutil.copy(f"{G}/{tag}_launches_bench.csv", f"{P}/{R}_launches_bench.csv")
for rep, out in ((f"{tag}_tds_offsets_fullsize", f"{R}_tds_offsets_fullsize_ncu_raw.csv"), (f"{tag}_st_post", f"{R}_st_post_ncu_raw.csv"),
                 (f"{tag}_resample", f"{R}_resample_ncu_raw.csv"), (f"{tag}_stft", f"{R}_stft_ncu_raw.csv")):
    with open(f"{P}/{out}", "w") as f:
        subprocess.run(["ncu", "-i", f"{G}/{rep}.ncu-rep", "--page", "raw", "--csv"], stdout=f, stderr=subprocess.DEVNULL, check=True)
rows = list(csv.reader(open(f"{P}/{R}_launches_bench.csv")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    try: v = float(r[mv].replace(",", ""))
    except ValueError: continue
    ms = v / 1e6 if r[mu] in ("ns", "nsecond") else (v / 1e3 if r[mu] in ("us", "usecond") else v)
    name = r[kn].split("(")[0].replace("void ", "").replace("nodey::", "")
    tot[name] += ms; cnt[name] += 1
T = sum(tot.values())
out = [f"# ncu launch list summary (profiles/{R}_launches_bench.csv)", "",
       "command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv python bench.py --tracks 64 --seconds 20 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e`",
       "(per-launch times are cold-cache and serialised: compare SHARES; all launches of the run incl. source synthesis, warm-up steps, timed step and the profiled steps)", "",
       "| kernel | launches | total ms | share |", "|---|---:|---:|---:|"]
for n, ms in tot.most_common():
    out.append(f"| `{n}` | {cnt[n]} | {ms:.3f} | {100 * ms / T:.1f}% |")
open(f"{P}/{R}_launches_bench_summary.md", "w").write("\n".join(out) + "\n")
print("\n".join(out[5:]))
rows = list(csv.reader(open(f"{P}/{R}_tds_offsets_fullsize_ncu_raw.csv")))
d, u = dict(zip(rows[0], rows[2])), dict(zip(rows[0], rows[1]))
b = lambda k: float(d[k]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u[k]]
tr = b("dram__bytes_read.sum") + b("dram__bytes_write.sum")
json.dump({"tds_offsets_kernel": tr, "_note": "dram__bytes_read.sum + dram__bytes_write.sum of the pitch-node launch of one full-size step (256 tracks x 180 s: %.2f GB read + %.1f MB written); ncu --set full --clock-control none, profiles/%s_tds_offsets_fullsize_ncu_raw.csv (commit %s). The search windows AND the regions the next mid buffer can come from are staged, hence more than the 10.4 GB of windows + mid buffers a perfect implementation reads" % (b("dram__bytes_read.sum") / 1e9, b("dram__bytes_write.sum") / 1e6, R, subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()),
           "_commit": subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()},
          open(f"{P}/roofline_traffic.json", "w"), indent=1)
j = json.load(open(f"{P}/{R}_bench_n1.json"))
print("value", j["value"], "ms", j["ms_per_step"], "e2e", j["e2e"]["value"], j["e2e"]["ms_per_step"], "launches", j["gpu_launches"], "cpu", j["cpu_baseline"]["value"])
print(json.dumps(j["roofline"], indent=0)[:1500])
print(open(f"{P}/{R}_kbench.txt").read())
shutil.copy(f"{G}/{tag}_stft_timing.txt", f"{P}/{R}_stft_timing.txt")
for f in (f"{R}_tds_offsets_fullsize_ncu_raw.csv", f"{R}_st_post_ncu_raw.csv", f"{R}_resample_ncu_raw.csv", f"{R}_stft_ncu_raw.csv"):
    rows = list(csv.reader(open(f"{P}/{f}")))
    for r in rows[2:]:
        d, u = dict(zip(rows[0], r)), dict(zip(rows[0], rows[1]))
        print(f, d["Kernel Name"][:44], "ms", d["gpu__time_duration.sum"], "dram", d["dram__bytes_read.sum"], u["dram__bytes_read.sum"], d["dram__bytes_write.sum"], u["dram__bytes_write.sum"],
              "issue", d["smsp__issue_active.avg.pct_of_peak_sustained_active"][:5], "fma", d["sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"][:5])
