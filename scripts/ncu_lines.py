"""Summarise an ncu report's source page: per source line, instructions executed and stall
samples (top N).  usage: ncu_lines.py report.ncu-rep [kernel-substring] [topN]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
filt = sys.argv[2] if len(sys.argv) > 2 else ""
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
cur_file, cur_fn, hdr = None, None, None
done_fns = set()
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
    elif r[0] == "Function Name":
        cur_fn = r[1]
    elif r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
    elif hdr and cur_fn and filt in cur_fn and r[0].isdigit():
        key = (cur_fn, cur_file, int(r[0]), r[1][:90])
        try:
            ins = int(r[hdr["Instructions Executed"]])
            smp = int(r[hdr["# Samples"]])
        except Exception:
            continue
        a = agg.setdefault(key, [0, 0, {}])
        a[0] += ins
        a[1] += smp
        for st in ("stall_long_sb", "stall_barrier", "stall_short_sb", "stall_wait", "stall_math", "stall_mio", "stall_lg", "stall_not_selected"):
            try:
                a[2][st] = a[2].get(st, 0) + int(r[hdr[st]])
            except Exception:
                pass
fns = sorted(set(k[0] for k in agg))
for fn in fns[:1] if filt else fns:
    items = [(k, v) for k, v in agg.items() if k[0] == fn]
    tot_i = sum(v[0] for _, v in items) or 1
    tot_s = sum(v[1] for _, v in items) or 1
    print("==", fn[:120], "total inst", tot_i, "samples", tot_s)
    for k, v in sorted(items, key=lambda kv: -kv[1][1])[:top]:
        stalls = ",".join("%s=%d" % (s.replace("stall_", ""), n) for s, n in sorted(v[2].items(), key=lambda x: -x[1])[:3] if n)
        print("%5.1f%%smp %5.1f%%ins  %s:%d  %s   [%s]" % (100.0 * v[1] / tot_s, 100.0 * v[0] / tot_i, k[1], k[2], k[3].strip(), stalls))
