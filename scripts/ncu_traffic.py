"""Merges one workload's entry into profiles/<round>/traffic.json from an `ncu --set full --page raw --csv` dump: DRAM and L2
bytes per launch of one kernel, stamped with the hash of the CUDA sources the capture ran (written on the GPU box next to the
dump by scripts/gpu_bench_profile.sh) and the commit. bench.py reports the figures only while that hash equals the hash of the
sources it is running (bench.kernel_source_hash), so a capture can not silently go stale.
usage: ncu_traffic.py raw.csv kernel-substring out.json workload-key kernel-hash "source note" """
import csv, json, os, subprocess, sys
raw, kernel, out_path, key, khash, source = sys.argv[1:7]
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
def to_bytes(r, name):
    return float(r[idx[name]]) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[units[idx[name]]]
def num(r, name):
    return float(r[idx[name]]) if name in idx and r[idx[name]] not in ("", "n/a") else None
sel = [r for r in rows[2:] if kernel in r[idx["Kernel Name"]]]
dram = [to_bytes(r, "dram__bytes_read.sum") + to_bytes(r, "dram__bytes_write.sum") for r in sel]
l2 = [float(r[idx["lts__t_sectors.sum"]]) * 32.0 for r in sel]            # 32-byte L2 sectors (all sources)
l1 = [float(r[idx["l1tex__t_sectors.sum"]]) * 32.0 for r in sel] if "l1tex__t_sectors.sum" in idx else []
dur = [float(r[idx["gpu__time_duration.sum"]]) for r in sel]
def mean(name):
    v = [num(r, name) for r in sel]
    v = [x for x in v if x is not None]
    return sum(v) / len(v) if v else None
try:
    commit = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=os.path.dirname(os.path.abspath(__file__))).stdout.strip()
except Exception:
    commit = None
entry = {"kernel": kernel, "kernel_hash": khash, "commit": commit, "launches_profiled": len(sel),
         "dram_bytes_per_launch": sum(dram) / max(len(dram), 1), "l2_bytes_per_launch": sum(l2) / max(len(l2), 1),
         "l1_bytes_per_launch": (sum(l1) / len(l1)) if l1 else None, "dram_bytes": dram, "l2_bytes": l2, "duration_" + units[idx["gpu__time_duration.sum"]]: dur,
         "ncu": {"dram_throughput_pct": mean("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                 "l1_hit_pct": mean("l1tex__t_sector_hit_rate.pct"), "l2_hit_pct": mean("lts__t_sector_hit_rate.pct"),
                 "lanes_per_instruction": mean("smsp__thread_inst_executed_per_inst_executed.ratio"),
                 "issue_active_pct": mean("smsp__issue_active.avg.pct_of_peak_sustained_active"),
                 "warps_active_pct": mean("sm__warps_active.avg.pct_of_peak_sustained_active")},
         "source": source}
data = {}
if os.path.exists(out_path):
    with open(out_path) as f:
        data = json.load(f)
data[key] = entry
with open(out_path, "w") as f:
    json.dump(data, f, indent=1)
print(key, {k: v for k, v in entry.items() if k not in ("dram_bytes", "l2_bytes")})
