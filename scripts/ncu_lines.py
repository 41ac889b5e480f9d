"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line:
    ncu -i rep.ncu-rep --page source --csv --print-source cuda,sass > x.csv ; python scripts/ncu_lines.py x.csv [top]
Prints per (file, line): warp-stall samples, instructions executed, top stall reasons, shared-memory wavefronts."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file, hdr = None, None
agg = collections.OrderedDict()
tot = collections.Counter()
cur = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; continue
    if hdr is None or len(r) < len(hdr):
        continue
    d = dict(zip(range(len(hdr)), r))
    if r[0] != "":
        cur = (cur_file, int(r[0]), r[1].strip()[:110])
        agg.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    a = agg[cur]
    for i, h in enumerate(hdr):
        if i < 4:
            continue
        if h in ("# Samples", "Instructions Executed", "L1 Wavefronts Shared", "L1 Wavefronts Shared Excessive") or (h.startswith("stall_") and "Not Issued" not in h):
            try:
                v = float(r[i])
            except ValueError:
                continue
            a[h] += v
            tot[h] += v
print("totals: samples %d  inst %.3e  smem wavefronts %.3e (excess %.3e)" % (tot["# Samples"], tot["Instructions Executed"], tot["L1 Wavefronts Shared"], tot["L1 Wavefronts Shared Excessive"]))
print("stalls:", ", ".join("%s %.1f%%" % (k[6:], 100 * v / max(tot["# Samples"], 1)) for k, v in sorted(((k, v) for k, v in tot.items() if k.startswith("stall_")), key=lambda kv: -kv[1])[:10]))
for key, a in sorted(agg.items(), key=lambda kv: -kv[1]["# Samples"])[:top]:
    st = sorted(((k[6:], v) for k, v in a.items() if k.startswith("stall_") and v > 0), key=lambda kv: -kv[1])[:3]
    print("%5.1f%% inst %5.1f%% wf %5.1f%%  %s:%d  %s   [%s]" % (100 * a["# Samples"] / max(tot["# Samples"], 1), 100 * a["Instructions Executed"] / max(tot["Instructions Executed"], 1),
          100 * a["L1 Wavefronts Shared"] / max(tot["L1 Wavefronts Shared"], 1), key[0], key[1], key[2], ", ".join("%s %d" % (k, v) for k, v in st)))
