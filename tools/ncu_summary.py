"""Summaries of ncu output for profiles/ (run in the build container; ncu reads reports without a GPU).

    python tools/ncu_summary.py launches gpurun_out/x/ncu_launches.csv        # per-kernel time shares
    python tools/ncu_summary.py full [--sites=N] gpurun_out/x/full_conv.ncu-rep [...]   # one line per profiled launch

`full` reads `ncu -i REPORT --page raw --csv` and prints duration, tensor-pipe activity, SM / L2 / L1 / DRAM
throughput (% of peak), DRAM bytes read / written per launch (the roofline "traffic"), registers, SM clock."""
import csv
import re
import io
import subprocess
import sys
from collections import OrderedDict

METRICS = OrderedDict([
    ("dur[us]", ("gpu__time_duration.sum", 1e-3)),                     # ns -> us
    ("tensor%act", ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", 1)),
    ("sm%", ("sm__throughput.avg.pct_of_peak_sustained_elapsed", 1)),
    ("lts%", ("lts__throughput.avg.pct_of_peak_sustained_elapsed", 1)),
    ("l1tex%", ("l1tex__throughput.avg.pct_of_peak_sustained_active", 1)),
    ("dram%", ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", 1)),
    ("dramR[MB]", ("dram__bytes_read.sum", None)),
    ("dramW[MB]", ("dram__bytes_write.sum", None)),
    ("regs", ("launch__registers_per_thread", 1)),
    ("sm_GHz", ("smsp__cycles_elapsed.avg.per_second", None)),
])
UNIT_TO_MB = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "Tbyte": 1e6}
UNIT_TO_GHZ = {"hz": 1e-9, "Khz": 1e-6, "Mhz": 1e-3, "Ghz": 1.0, "cycle/second": 1e-9, "cycle/usecond": 1e-3,
               "cycle/nsecond": 1.0}
UNIT_TO_NS = {"nsecond": 1.0, "usecond": 1e3, "msecond": 1e6, "second": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}


def short(name: str) -> str:
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"svx::(<unnamed>|\(anonymous namespace\))::", "", name)
    name = re.sub(r"^(<unnamed>|unnamed>)::", "", name)
    return name.split("(")[0][:44]


def full(paths):
    sites = [p.split("=", 1)[1] for p in paths if p.startswith("--sites=")]
    paths = [p for p in paths if not p.startswith("--sites=")]
    if sites:
        print(f"# sites per launch: {int(sites[0])}   (bench.py reads this line for roofline.traffic)")
    print("# ncu --set full --clock-control none --import-source on (per profiled launch); dramR/dramW = "
          "dram__bytes_read/write.sum per launch (the roofline \"traffic\")")
    print(f"{'kernel':46s}{'grid':>8s}" + "".join(f"{k:>12s}" for k in METRICS))
    for path in paths:
        raw = subprocess.check_output(["ncu", "-i", path, "--page", "raw", "--csv"]).decode()
        rows = list(csv.reader(io.StringIO(raw)))
        header, units, data = rows[0], rows[1], rows[2:]
        col = {h: i for i, h in enumerate(header)}
        for r in data:
            out = []
            for label, (metric, scale) in METRICS.items():
                if metric not in col:
                    out.append(f"{'n/a':>12s}")
                    continue
                txt = r[col[metric]].replace(",", "")
                try:
                    v = float(txt)
                except ValueError:
                    out.append(f"{'n/a':>12s}")
                    continue
                u = units[col[metric]]
                if label.startswith("dram") and label.endswith("[MB]"):
                    v *= UNIT_TO_MB.get(u, 1.0)
                elif label == "sm_GHz":
                    v *= UNIT_TO_GHZ.get(u, 1.0)
                elif label == "dur[us]":
                    v *= UNIT_TO_NS.get(u, 1.0) * 1e-3
                out.append(f"{v:12.3f}")
            grid = r[col["Grid Size"]] if "Grid Size" in col else ""
            print(f"{short(r[col['Kernel Name']]):46s}{grid.replace(' ', ''):>8s}" + "".join(out))


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(io.StringIO("".join(lines))))
    agg = OrderedDict()
    total = 0.0
    for r in rows:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", "")) * UNIT_TO_NS.get(r.get("Metric Unit", "ns"), 1.0) * 1e-3
        key = (short(r["Kernel Name"]), r.get("Grid Size", ""), r.get("Block Size", ""))
        a = agg.setdefault(key, [0, 0.0])
        a[0] += 1
        a[1] += v
        total += v
    print("# ncu --metrics gpu__time_duration.sum --clock-control none; per-launch times are cold-cache and "
          "serialised: compare SHARES")
    print(f"{'kernel':46s}{'grid':>16s}{'block':>16s}{'n':>6s}{'mean_us':>12s}{'share':>8s}")
    for (k, g, b), (n, us) in agg.items():
        print(f"{k:46s}{g:>16s}{b:>16s}{n:6d}{us / n:12.1f}{100 * us / total:7.1f}%")


if __name__ == "__main__":
    if len(sys.argv) < 3 or sys.argv[1] not in ("full", "launches"):
        sys.exit(__doc__)
    if sys.argv[1] == "full":
        full(sys.argv[2:])
    else:
        launches(sys.argv[2])
