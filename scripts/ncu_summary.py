"""Transpose `ncu -i REPORT --page raw --csv` into metric rows x kernel columns (the summaries kept under profiles/)."""
import csv, subprocess, sys

rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
rows = [r for r in csv.reader(raw.splitlines()) if len(r) > 10]
hdr, units, kern = rows[0], rows[1], rows[2:]
ki = hdr.index("Kernel Name")
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["metric", "unit"] + [k[ki].split("(")[0] for k in kern])
    for i, h in enumerate(hdr):
        if h in ("ID", "Process ID", "Process Name", "Host Name", "Context", "Stream", "Device", "CC"):
            continue
        w.writerow([h, units[i]] + [k[i] for k in kern])
print("wrote", out, len(hdr), "metrics x", len(kern), "kernels")
