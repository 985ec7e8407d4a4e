#!/usr/bin/env python
"""Per-SASS-instruction view of an .ncu-rep (source page): address, samples, executed count, text.
   python tools/ncu_sass_hot.py file.ncu-rep [min_samples]   -- prints the whole listing with counts,
   lines below min_samples are elided in runs."""
import csv, subprocess, sys
path = sys.argv[1]
min_s = int(sys.argv[2]) if len(sys.argv) > 2 else 0
out = subprocess.run(['ncu', '-i', path, '--page', 'source', '--csv', '--print-source', 'sass'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
iS, iE, iSrc = hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Source')
tot_s = sum(int(r[iS]) for r in rows[2:] if len(r) > iE)
tot_e = sum(int(r[iE]) for r in rows[2:] if len(r) > iE)
print(f"total samples {tot_s}, total warp-instructions {tot_e}")
skipped = 0
for n, r in enumerate(rows[2:]):
    if len(r) <= iE:
        continue
    s, e = int(r[iS]), int(r[iE])
    if s < min_s:
        skipped += 1
        continue
    if skipped:
        print(f"      ... {skipped} instructions below {min_s} samples")
        skipped = 0
    print(f"{n:5d} {s:6d} {100*s/max(tot_s,1):5.1f}% {e:9d}  {r[iSrc].strip()[:110]}")
