"""Source-page digest of an `ncu --set full --import-source on` report of ONE kernel: warp-stall samples and executed
instructions per SASS opcode, the stall-reason mix, and the hottest lines.
usage: python tools/ncu_stalls.py report.ncu-rep [top_lines] [> profiles/xxx.txt]"""
import collections
import csv
import io
import re
import subprocess
import sys


def num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return 0.0


def main(path, top=24):
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    print(rows[0][1][:150] if rows and len(rows[0]) > 1 else "")
    hdr = rows[1]
    col = {k: i for i, k in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    stall = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    ops_i, ops_s, st = collections.Counter(), collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[col["Source"]])
        op = ".".join(m.group(2).split(".")[:2]) if m else r[col["Source"]][:16]
        ops_i[op] += num(r[col["Instructions Executed"]])
        ops_s[op] += num(r[col["# Samples"]])
        for k in stall:
            st[k] += num(r[col[k]])
    T, I = sum(ops_s.values()), sum(ops_i.values())
    print(f"warp-stall samples {T:.0f}, warp instructions executed {I:.0f}")
    print("stall reasons (% of samples): " + ", ".join(f"{k[6:]} {100 * v / max(1, sum(st.values())):.1f}" for k, v in st.most_common(8)))
    print("opcode                    samples     %   instructions     %")
    for op, s in ops_s.most_common(18):
        print(f"{op:24s} {s:8.0f} {100 * s / T:5.1f} {ops_i[op]:14.0f} {100 * ops_i[op] / I:5.1f}")
    print(f"hottest {top} lines (index, executed, samples, SASS, top stall reasons)")
    hot = sorted(range(len(data)), key=lambda i: -num(data[i][col["# Samples"]]))[:top]
    for i in sorted(hot):
        r = data[i]
        ss = sorted(((num(r[col[k]]), k[6:]) for k in stall), reverse=True)[:2]
        print(f"{i:5d} {num(r[col['Instructions Executed']]):10.0f} {num(r[col['# Samples']]):6.0f}  {r[col['Source']].strip()[:62]:62s} "
              + " ".join(f"{k}:{int(v)}" for v, k in ss if v > 0))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 24)
