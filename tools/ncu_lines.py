"""Per-source-line stall samples from `ncu --page source --csv --print-source cuda,sass` output.
usage: ncu -i rep --page source --csv --print-source cuda,sass > src.csv; python tools/ncu_lines.py src.csv [top]"""
import csv
import sys

top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
cur_file = cur_fn = None
acc = {}
for r in csv.reader(open(sys.argv[1])):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1].split("::")[-1][:40]
        continue
    if r[0] in ("Line No", "") or not r[0].isdigit():
        continue
    try:
        samples, execd = int(r[4]), int(r[7])
    except ValueError:
        continue
    acc.setdefault(cur_fn, []).append((samples, execd, cur_file, int(r[0]), r[1].strip()[:110]))
for fn, rows in acc.items():
    tot = sum(x[0] for x in rows) or 1
    print(f"==== {fn}: {tot} samples, {sum(x[1] for x in rows)} warp-instructions")
    for s, e, f, ln, src in sorted(rows, reverse=True)[:top]:
        print(f"{s:6d} {100 * s / tot:5.1f}%  exec={e:9d}  {f}:{ln}  {src}")
