"""usage: python scratch/mk_profile.py TAG  — copies the launch list and writes the full-set summary
and the per-entry-point DRAM traffic of gpurun_out/prof_TAG.ncu-rep into profiles/."""
import csv, json, re, shutil, subprocess, sys
from collections import defaultdict
tag = sys.argv[1]
out = "r01" + tag[-1] if tag.startswith("r1") else tag
shutil.copy(f"gpurun_out/launches_{tag}.csv", f"profiles/{out}_launches.csv")
summ = subprocess.run([sys.executable, "scratch/ncu_summary.py", f"gpurun_out/prof_{tag}.ncu-rep"], capture_output=True, text=True).stdout
hdr = f"# {out}: ncu --set full --clock-control none --import-source on, python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e (B=8 volumes of 256^3 per launch)\n"
open(f"profiles/{out}_ncu_full_summary.txt", "w").write(hdr + summ)
# launch shares
rows = [r for r in csv.DictReader(l for l in open(f"gpurun_out/launches_{tag}.csv") if l.startswith('"'))]
t = defaultdict(lambda: [0, 0.0])
for r in rows:
    if r["Metric Name"] == "gpu__time_duration.sum":
        k = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("fsg::", "")
        t[k][0] += 1; t[k][1] += float(r["Metric Value"]) / 1e6
tot = sum(v[1] for v in t.values())
with open(f"profiles/{out}_launch_shares.txt", "w") as f:
    f.write(f"# {out}: kernel, launches, total ms, share of all captured launches (cold-cache, serialised)\n")
    for k, v in sorted(t.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k:60s} {v[0]:4d} {v[1]:9.3f} ms {100*v[1]/tot:5.1f} %\n")
# traffic per entry point (dominant device kernel of each)
ENTRY = {"warp_fast_kernel": "fsg_warp", "warp_kernel<1": "fsg_warp", "gmm_kernel": "fsg_gmm", "zoom_rows_kernel<0": "fsg_zoom", "zoom_rows_kernel<1": "fsg_zoom_minmax", "sep_": "fsg_sepconv"}
per_kernel = defaultdict(lambda: [0.0, 0]); cur = None
for line in summ.splitlines():
    if line.startswith("---"):
        cur = re.sub(r"\(.*", "", line[4:]).replace("void ", "").strip()
        per_kernel[cur][1] += 1
    m = re.match(r"\s+dram__bytes_(read|write)\.sum: ([0-9.]+) (\w+)", line)
    if m and cur:
        mul = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}[m.group(3)]
        per_kernel[cur][0] += float(m.group(2)) * mul
nsteps = min(v[0] for v in t.values())  # every kernel runs >= once per step
traffic = defaultdict(float)
for k, (byt, cnt) in per_kernel.items():
    e = next((e for p, e in ENTRY.items() if p in k), None)
    full = next((kk for kk in t if kk.startswith(k[:40])), None)
    if e and cnt and full:
        traffic[e] += byt / cnt * (t[full][0] / nsteps)
json.dump({k: v for k, v in traffic.items()}, open("profiles/dominant_kernel_traffic.json", "w"), indent=1)
print(open(f"profiles/{out}_launch_shares.txt").read()); print(dict(traffic))
