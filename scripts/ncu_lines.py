"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump per source line.
usage: ncu_lines.py file.csv [top]"""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, hdr, out = None, None, []
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) >= 12 and r[0] not in ("", "Function Name"):
        try:
            out.append(dict(file=cur, line=int(r[0]), src=r[1], samples=int(r[6]), inst=int(r[7]), thr=int(r[8]),
                            avg_thr=r[10]))
        except ValueError:
            pass
ti = sum(d["inst"] for d in out)
ts = sum(d["samples"] for d in out)
print(f"total warp-inst {ti}  samples {ts}")
out.sort(key=lambda d: -d["inst"])
for d in out[:top]:
    print(f"{d['inst'] / ti * 100:5.1f}% inst {d['samples'] / max(ts, 1) * 100:5.1f}% smp  thr/inst {d['thr'] / max(d['inst'], 1):5.1f}  "
          f"{d['file']}:{d['line']:<4} {d['src'].strip()[:100]}")
