"""Turn gpurun_out/<round>_launches.csv and <round>_prof.ncu-rep into the small text summaries kept under profiles/."""
import csv, subprocess, sys, collections, io
R = sys.argv[1] if len(sys.argv) > 1 else "r1"
src = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out"
out = open(f"profiles/{R}_launch_list_summary.txt", "w")
rows = list(csv.reader(open(f"{src}/{R}_launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]; ki, vi = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi: continue
    n = r[ki].split("(")[0][:70]; v = float(r[vi].replace(",", ""))
    a = agg.setdefault(n, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(v[1] for v in agg.values())
out.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised: compare SHARES)\n")
out.write(f"# command: python bench.py --steps 2 --warmup 1 --no-cpu --scale-leg off --min-seconds 0 --no-check   (3 warm-up + 2 timed + 3+2 e2e steps)\n")
out.write(f"{'kernel':72s} {'launches':>8s} {'total_ms':>10s} {'avg_us':>10s} {'share':>7s}\n")
for n, (c, v) in sorted(agg.items(), key=lambda x: -x[1][1]):
    out.write(f"{n:72s} {c:8d} {v / 1e6:10.3f} {v / c / 1e3:10.1f} {v / tot:7.3f}\n")
out.close()
raw = subprocess.run(["ncu", "-i", f"{src}/{R}_prof.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
H = rows[0]
keys = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.sum", "smsp__inst_executed.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio"]
with open(f"profiles/{R}_ncu_full_summary.txt", "w") as f:
    f.write("# ncu --set full --clock-control none --import-source on; units as printed by ncu (dram bytes in MB, time in us or ms)\n")
    units = rows[1]
    for r in rows[2:]:
        f.write("\n")
        for k in keys:
            if k in H:
                i = H.index(k)
                f.write(f"{k:75s} {r[i][:110]:>20s} {units[i]}\n")
# measured DRAM traffic per launch for bench.py's roofline.traffic
import json, os
tr = {}
ir, iw, ik = H.index("dram__bytes_read.sum"), H.index("dram__bytes_write.sum"), H.index("Kernel Name")
def to_bytes(v, unit):
    return float(v) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[unit]
sp = [to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in rows[2:] if r[ik].startswith("void k_spmm") or r[ik].startswith("k_spmm_fixed")]
sc = [to_bytes(r[ir], units[ir]) + to_bytes(r[iw], units[iw]) for r in rows[2:] if "k_score_topk_tc" in r[ik] or "k_score_topk_gq" in r[ik]]
key = os.environ.get("LGX_TRAFFIC_KEY", "amazon-book:bf16")
path = "profiles/traffic.json"
alltr = json.load(open(path)) if os.path.exists(path) else {}
alltr[key] = {"spmm_layer_bytes": sum(sp) / len(sp) if sp else None, "score_bytes": sum(sc) / len(sc) if sc else None,
              "source": f"profiles/{R}_ncu_full_summary.txt"}
json.dump(alltr, open(path, "w"), indent=1)
print(open(f"profiles/{R}_launch_list_summary.txt").read())
print(open(f"profiles/{R}_ncu_full_summary.txt").read())
