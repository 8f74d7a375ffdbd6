"""Writes profiles/<round>/traffic.json from an `ncu --set full --page raw --csv` dump: DRAM bytes per launch of one kernel.
usage: ncu_traffic.py raw.csv kernel-substring out.json "source note" """
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
def to_bytes(v, u):
    v = float(v)
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
sel = [r for r in rows[2:] if sys.argv[2] in r[idx["Kernel Name"]]]
tot = [to_bytes(r[idx["dram__bytes_read.sum"]], units[idx["dram__bytes_read.sum"]]) + to_bytes(r[idx["dram__bytes_write.sum"]], units[idx["dram__bytes_write.sum"]]) for r in sel]
dur = [float(r[idx["gpu__time_duration.sum"]]) for r in sel]
out = {"kernel": sys.argv[2], "dram_bytes_per_launch": sum(tot) / max(len(tot), 1), "launches_profiled": len(tot),
       "per_launch": tot, "duration_" + units[idx["gpu__time_duration.sum"]]: dur, "source": sys.argv[4]}
json.dump(out, open(sys.argv[3], "w"), indent=1)
print(out)
