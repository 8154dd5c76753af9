"""Extract dram__bytes_read.sum + dram__bytes_write.sum (per launch) of the solver kernel from `ncu --set full`
reports and write profiles/ncu_traffic.json, which bench.py quotes as roofline.traffic.
usage: python scripts/ncu_traffic.py c3=gpurun_out/prof_c3.ncu-rep c4=... c5=... [--scale c3=1.0]
A capture taken on a reduced batch (fewer columns) is scaled to the full workload with name=factor:scale,
stated in the `source` string."""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def read(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        rd = float(d["dram__bytes_read.sum"].replace(",", "")) * UNIT[u["dram__bytes_read.sum"]]
        wr = float(d["dram__bytes_write.sum"].replace(",", "")) * UNIT[u["dram__bytes_write.sum"]]
        out.append((d["Kernel Name"], rd, wr, float(d["gpu__time_duration.sum"].replace(",", "")), u["gpu__time_duration.sum"]))
    return out


def main():
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        table = json.load(open(path))
    except (OSError, ValueError):
        table = {}
    for arg in sys.argv[1:]:
        name, rest = arg.split("=", 1)
        rep, _, scale = rest.partition(":")
        scale = float(scale) if scale else 1.0
        ks = read(rep)
        kname, rd, wr, dur, dur_u = ks[-1]
        fam = ("single_stream" if "single_stream" in kname else "multi_reg3" if "multi_reg3" in kname else "multi_reg2" if "multi_reg2" in kname else
               "multi_reg" if "multi_reg" in kname else
               "multi_tiled" if "multi_tiled" in kname else "multi_resident" if "multi" in kname else
               "single_resident_mw" if "mw" in kname else "single_resident")
        table[name] = {"kernel_family": fam, "kernel": kname, "dram_bytes_read": rd * scale, "dram_bytes_write": wr * scale,
                       "dram_bytes_per_launch": (rd + wr) * scale,
                       "source": f"ncu --set full --clock-control none, {os.path.basename(rep)}"
                                 + (f", capture on 1/{scale:g} of the columns scaled x{scale:g}" if scale != 1.0 else "")
                                 + f", launch {dur:g} {dur_u} under the profiler"}
    json.dump(table, open(path, "w"), indent=1)
    print(json.dumps(table, indent=1))


if __name__ == "__main__":
    main()
