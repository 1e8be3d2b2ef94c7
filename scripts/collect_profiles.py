#!/usr/bin/env python
"""Copy the evidence of one gpu_round.sh pass from gpurun_out/ (scratch) into profiles/<round>/ (tracked):
bench lines, launch-list shares, ncu --set full summary, phase shares, and the per-launch DRAM traffic bench.py reports.
usage: python scripts/collect_profiles.py <tag> [round_dir]"""
import collections, csv, json, os, shutil, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]; rnd = sys.argv[2] if len(sys.argv) > 2 else "r01"
src = os.path.join(ROOT, "gpurun_out"); dst = os.path.join(ROOT, "profiles", rnd)
os.makedirs(dst, exist_ok=True)
for a, b in ((f"bench_{tag}.json", f"bench_{tag}_cfg2.json"), (f"bench_ref_{tag}.json", f"bench_ref_{tag}_cfg2.json"), (f"phase_{tag}.txt", f"phase_{tag}.txt")):
    if os.path.exists(os.path.join(src, a)):
        shutil.copy(os.path.join(src, a), os.path.join(dst, b))
# launch list
rows = list(csv.reader(open(os.path.join(src, f"launches_{tag}.csv"))))
for i, r in enumerate(rows):
    if r and r[0] == "ID":
        hdr, data = r, rows[i + 1:]
        break
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in data:
    if len(r) <= iv:
        continue
    a = agg.setdefault(r[ik].split("(")[0], [0, 0.0]); a[0] += 1; a[1] += float(r[iv].replace(",", ""))
tot = sum(v[1] for v in agg.values()); unit = data[0][hdr.index("Metric Unit")]
out = [f"ncu --metrics gpu__time_duration.sum --clock-control none -c 400: python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu (kernel {tag}; cold-cache, serialised: compare SHARES)",
       f"{'kernel':70s} {'launches':>8s} {'total ' + unit:>16s} {'share':>7s}"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k[:70]:70s} {v[0]:8d} {v[1]:16.0f} {100 * v[1] / tot:6.2f}%")
# share of the hot kernel among the kernels of the timed step only (the command also generates the data and runs bm25)
setup = ("sequential_mean", "bm25_apply", "doclen_df", "idf_kernel", "at_cuda", "native::", "cuda::", "at::", "cub::", "sort_rows")
step = {k: v for k, v in agg.items() if "spy::" in k and not any(x in k for x in setup)}
step_tot = sum(v[1] for v in step.values())
hot = sum(v[1] for k, v in step.items() if "knn_flat_kernel" in k or "knn_stream_kernel" in k)
out.append("")
out.append(f"kernels of the timed step only (without data generation and bm25): hot kernel {100 * hot / step_tot:.1f} % of {step_tot / 1e6:.1f} ms "
           f"-- bench.py reports kernel_share_of_step = {json.load(open(os.path.join(src, f'bench_{tag}.json')))['roofline']['kernel_share_of_step']}")
open(os.path.join(dst, f"launches_{tag}_cfg2.txt"), "w").write("\n".join(out) + "\n")
# ncu full summary
rep = os.path.join(src, f"prof_knn_{tag}.ncu-rep")
summ = subprocess.run([sys.executable, os.path.join(ROOT, "scripts", "ncu_summary.py"), rep], capture_output=True, text=True).stdout
bench = json.load(open(os.path.join(src, f"bench_{tag}.json")))
r = bench["roofline"]
head = [f"ncu --set full --clock-control none --import-source on -k regex:knn_ -c 1   (bench.py cfg2 full size, kernel {tag})",
        f"plan {r['plan']}; algorithmic bytes {r['algorithmic_bytes'] / 1e9:.1f} GB; CUDA-event kernel time in bench.py {r['kernel_ms']} ms -> {r['achieved']} GB/s = {r['frac']} of the measured {r['peak']} GB/s"]
open(os.path.join(dst, f"knn_{tag}_ncu_summary.txt"), "w").write("\n".join(head) + "\n" + summ)
rd = wr = None
for line in summ.splitlines():
    if line.startswith("dram__bytes_read.sum"): rd = float(line.split()[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Tbyte": 1e12}[line.split()[2]]
    if line.startswith("dram__bytes_write.sum"): wr = float(line.split()[1]) * {"Gbyte": 1e9, "Mbyte": 1e6, "Tbyte": 1e12}[line.split()[2]]
if rd is not None and wr is not None:
    json.dump({"kernel": f"{r.get('kernel', 'knn_stream_kernel')} {tag}", "workload_nnz": bench["config"]["nnz"], "dram_bytes_per_launch": int(rd + wr),
               "source": f"profiles/{rnd}/knn_{tag}_ncu_summary.txt (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum)"},
              open(os.path.join(ROOT, "profiles", "knn_traffic.json"), "w"), indent=1)
print("\n".join(out[:8])); print(summ[:1500])
