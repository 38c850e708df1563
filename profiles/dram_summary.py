"""ncu `--metrics dram__bytes_read.sum,dram__bytes_write.sum,... --csv --log-file X` -> the two-column summary that
bench.py reads for `roofline.traffic` (profiles/r<round>_dram_<config>_<size>_<pre|post>.csv): per-launch averages
over the captured launches of the step kernel.

    python profiles/dram_summary.py gpurun_out/r2l_dram_c2_512_pre.csv > profiles/r2_dram_c2_512_pre.csv
"""
import csv
import sys

SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
         "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "%": 1.0}


def main(path):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    hdr = rows[0]
    name, unit, value, kern, ident = (hdr.index(k) for k in ("Metric Name", "Metric Unit", "Metric Value",
                                                             "Kernel Name", "ID"))
    acc, kernels, launches = {}, set(), set()
    for r in rows[1:]:
        v = float(r[value].replace(",", "")) * SCALE.get(r[unit], 1.0)
        acc.setdefault(r[name], []).append(v)
        kernels.add(r[kern])
        launches.add(r[ident])
    w = csv.writer(sys.stdout)
    w.writerow(["metric", "value", "unit"])
    for k, vals in acc.items():
        u = "byte" if "bytes" in k else ("us" if "time" in k else "%")
        w.writerow([k, sum(vals) / len(vals), u])
    w.writerow(["launches_averaged", len(launches), ""])
    for k in sorted(kernels):
        w.writerow(["kernel", k, ""])


if __name__ == "__main__":
    main(sys.argv[1])
