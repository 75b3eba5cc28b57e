"""Per-launch duration and DRAM bytes out of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
--csv` log (tools/gpu_evidence.sh writes them).  Usage: python tools/ncu_dram_table.py <csv> [min_us]"""
import collections
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    min_us = float(sys.argv[2]) if len(sys.argv) > 2 else 100.0
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    d = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(hdr):
            continue
        rec = dict(zip(hdr, r))
        e = d.setdefault(rec["ID"], {"k": rec["Kernel Name"][:64], "grid": rec["Grid Size"]})
        scale = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}[rec["Metric Unit"]]
        e[rec["Metric Name"]] = float(rec["Metric Value"].replace(",", "")) * scale
    for k, v in d.items():
        t = v.get("gpu__time_duration.sum", 0.0)
        if t >= min_us:
            print(f"{k:>4} {v['k']:64s} {v['grid']:14s} {t:8.1f} us  rd {v.get('dram__bytes_read.sum', 0):7.0f} MB  wr {v.get('dram__bytes_write.sum', 0):7.0f} MB")


if __name__ == "__main__":
    main()
