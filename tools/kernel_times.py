"""ncu --csv launch list -> per-kernel total ms (short names).  python tools/kernel_times.py file.csv"""
import csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]; ki = hdr.index("Kernel Name"); vi = hdr.index("Metric Value"); gi = hdr.index("Grid Size")
tot = {}
for r in rows[1:]:
    nm = r[ki].split("(")[0]
    if "cutlass" in nm:
        nm = "cutlass i8gemm " + ("batched" if not r[gi].endswith(" 1)") else "single")
    nm = nm.replace("gpz::", "").replace("void ", "")
    try: t = float(r[vi].replace(",", "")) / 1e6
    except ValueError: continue
    tot.setdefault(nm, [0, 0.0]); tot[nm][0] += 1; tot[nm][1] += t
for k, (c, t) in sorted(tot.items(), key=lambda x: -x[1][1])[:int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print(f"{t:9.3f} ms x{c:4d}  {k}")
