"""Summarise an `ncu --page source --print-source cuda,sass --csv` dump per CUDA source line (dev tool)."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Line No")
hdr = rows[hi]
idx = {}
for i, h in enumerate(hdr):
    idx.setdefault(h, i)
data = [r for r in rows[hi + 1:] if r and r[0] not in ("", "Line No", "File Path", "Function Name") and len(r) > 10]
def f(r, c):
    try: return float(r[idx[c]] or 0)
    except Exception: return 0.0
ti = sum(f(r, "Instructions Executed") for r in data); ts = sum(f(r, "# Samples") for r in data)
print("total warp-inst %.4g  samples %d" % (ti, ts))
stall = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
data.sort(key=lambda r: -f(r, "# Samples"))
for r in data[:top]:
    st = sorted(((f(r, c), c) for c in stall), reverse=True)[:3]
    print("%5s %5.1f%%smp %5.1f%%inst wf_sh=%-9d ideal=%-9d | %-90s | %s" % (r[0], 100 * f(r, "# Samples") / ts, 100 * f(r, "Instructions Executed") / ti,
          f(r, "L1 Wavefronts Shared"), f(r, "L1 Wavefronts Shared Ideal"), r[1].strip()[:90], " ".join("%s=%d" % (c[6:], v) for v, c in st if v > 0)))
