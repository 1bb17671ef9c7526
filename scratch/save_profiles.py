#!/usr/bin/env python
"""Copy the ncu evidence of one gpurun call from gpurun_out/ (scratch) into profiles/ (tracked).
usage: save_profiles.py TAG [launches.csv] [name=file.ncu-rep ...]; writes profiles/TAG_launches.csv, profiles/TAG_<name>_raw.csv and
refreshes profiles/traffic.json from the k_sweep capture (dram bytes read + written per launch)."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1]
for a in sys.argv[2:]:
    if "=" not in a:
        rows = [l for l in open(a) if l.startswith('"')]
        open(os.path.join(ROOT, "profiles", tag + "_launches.csv"), "w").writelines(rows)
        continue
    name, rep = a.split("=")
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    open(os.path.join(ROOT, "profiles", "%s_%s_raw.csv" % (tag, name)), "w").write(out)
    rows = list(csv.reader(io.StringIO(out)))
    h, u, r = rows[0], rows[1], rows[2]
    def val(m):
        i = h.index(m); x = float(r[i].replace(",", "")); un = u[i]
        return x * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}.get(un, 1.0)
    if "sweep" in name:
        t = {"k_residual_dram_bytes_per_launch": val("dram__bytes_read.sum") + val("dram__bytes_write.sum"),
             "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
             "kernel": r[h.index("Kernel Name")], "grid": r[h.index("launch__grid_size")], "gpu_time_ms_under_ncu": val("gpu__time_duration.sum"),
             "source": "profiles/%s_%s_raw.csv (ncu --set full --clock-control none, one launch at the bench.py default 256^3 workload)" % (tag, name)}
        json.dump(t, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
        print(t)
