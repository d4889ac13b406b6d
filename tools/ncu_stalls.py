#!/usr/bin/env python3
"""Summarise `ncu --page source --csv` output: per kernel, stall-reason totals and the top stalled instructions.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv ; tools/ncu_stalls.py src.csv [top_n]"""
import csv, sys, collections

def main():
    path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    kernels = []; cur = None; hdr = None
    for row in csv.reader(open(path)):
        if not row: continue
        if row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": []}; kernels.append(cur); hdr = None; continue
        if row[0] == "Address":
            hdr = row; cur["hdr"] = hdr; continue
        if cur is not None and hdr is not None:
            cur["rows"].append(row)
    for k in kernels:
        hdr = k["hdr"]; rows = k["rows"]
        si = hdr.index("# Samples"); src = hdr.index("Source")
        stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        tot = collections.Counter(); total = 0
        for r in rows:
            total += int(r[si] or 0)
            for i in stall_cols:
                tot[hdr[i]] += int(r[i] or 0)
        print(f"== {k['name']}: {len(rows)} instr, {total} samples")
        print("   " + ", ".join(f"{n}:{100*v/max(total,1):.1f}%" for n, v in tot.most_common(8)))
        top = sorted(rows, key=lambda r: -int(r[si] or 0))[:topn]
        for r in top:
            why = max(stall_cols, key=lambda i: int(r[i] or 0))
            print(f"   {int(r[si]):7d} {100*int(r[si])/max(total,1):5.1f}%  {hdr[why]:18s} {r[src].strip()[:90]}")

if __name__ == "__main__":
    main()
