"""Per-source-line instruction counts from `ncu --page source --csv --print-source cuda,sass` (k-th kernel in the file).
usage: ncu_lines.py file.csv [min_share_pct] [kernel_index=0]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
want = int(sys.argv[3]) if len(sys.argv) > 3 else 0
seen = -1
hdr = None
lines = []
n_fn = 0
cur_file = ""
for r in rows:
    if r and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    if r and r[0] == "Function Name":
        if seen < 0 or not r[1].startswith(fn):
            seen += 1
            if seen > want:
                break
            hdr, lines = None, []
        fn = r[1]
    if r and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) < 10:
        continue
    if r[0].strip().isdigit() and len(r) == len(hdr):  # a source-line summary row
        r[1] = cur_file[:12] + ": " + r[1].strip()
        lines.append(r)
i_inst = hdr.index("Instructions Executed"); i_thr = hdr.index("Thread Instructions Executed"); i_smp = hdr.index("# Samples")
tot_i = sum(float(r[i_inst]) for r in lines if r[i_inst] not in ("-", ""))
tot_t = sum(float(r[i_thr]) for r in lines if r[i_thr] not in ("-", ""))
tot_s = sum(float(r[i_smp]) for r in lines if r[i_smp] not in ("-", ""))
print(f"total warp-inst {tot_i:.3e} thread-inst {tot_t:.3e} avg lanes {tot_t/tot_i:.2f} samples {tot_s:.0f}")
for r in lines:
    if r[i_inst] in ("-", ""): continue
    wi, ti, sm = float(r[i_inst]), float(r[i_thr]), float(r[i_smp])
    if wi / tot_i * 100 >= thr or sm / max(tot_s, 1) * 100 >= thr:
        print(f"{r[0]:>5s} inst {100*wi/tot_i:5.1f}% lanes {ti/max(wi,1):5.1f} samples {100*sm/max(tot_s,1):5.1f}% | {r[1][:110]}")
