"""One line per captured launch of an ncu report (--set full): duration, DRAM traffic, DRAM / tensor-pipe / issue-slot
utilisation. usage: python tools/ncu_summary.py report.ncu-rep [> profiles/xxx_summary.txt]"""
import csv
import io
import re
import subprocess
import sys


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {k: i for i, k in enumerate(hdr)}

    def get(r, k, scale=1.0):
        try:
            return float(r[col[k]].replace(",", "")) * scale
        except (KeyError, ValueError):
            return float("nan")

    def to_unit(k, want):
        u = units[col[k]] if k in col else ""
        table = {("ns", "us"): 1e-3, ("us", "us"): 1.0, ("ms", "us"): 1e3, ("byte", "MB"): 1e-6, ("Kbyte", "MB"): 1e-3,
                 ("Mbyte", "MB"): 1.0, ("Gbyte", "MB"): 1e3}
        return table.get((u, want), 1.0)

    print("columns: duration us | dram read MB | dram write MB | DRAM throughput % of peak | tensor pipe active % | issue slots active % | kernel")
    for r in rows[2:]:
        name = re.sub(r"\(.*", "", r[col["Kernel Name"]]).replace("<unnamed>::", "").replace("void ", "")
        d = get(r, "gpu__time_duration.sum", to_unit("gpu__time_duration.sum", "us"))
        rd = get(r, "dram__bytes_read.sum", to_unit("dram__bytes_read.sum", "MB"))
        wr = get(r, "dram__bytes_write.sum", to_unit("dram__bytes_write.sum", "MB"))
        print(f"{d:9.1f} | {rd:7.1f} | {wr:7.1f} | {get(r, 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'):5.1f} | "
              f"{get(r, 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed'):5.1f} | "
              f"{get(r, 'sm__issue_active.avg.pct_of_peak_sustained_elapsed'):5.1f} | {name}")


if __name__ == "__main__":
    main(sys.argv[1])
