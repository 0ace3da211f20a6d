"""ncu launch-list CSV (gpu__time_duration.sum [+ dram bytes]) -> per-kernel shares CSV under profiles/."""
import collections
import csv
import sys


def summarize(path, out, title):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
    hdr = rows[hi]
    ki, vi, ui, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Metric Name")
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])
    for r in rows[hi + 1:]:
        if len(r) <= vi:
            continue
        try:
            v = float(r[vi].replace(",", ""))
        except ValueError:
            continue
        k = r[ki].split("(")[0]
        if r[mi] == "gpu__time_duration.sum":
            v = v / 1e3 if r[ui] == "ns" else (v * 1e3 if r[ui] == "ms" else v)
            agg[k][0] += 1
            agg[k][1] += v
        else:
            agg[k][2] += v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r[ui], 1)
    tot = sum(v[1] for v in agg.values())
    with open(out, "w", newline="") as f:
        f.write(f"# {title}; {sum(v[0] for v in agg.values())} launches, {tot / 1e3:.2f} ms serialised cold-cache total\n")
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_us", "share_pct", "dram_MB_per_launch"])
        for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
            w.writerow([k, v[0], f"{v[1]:.1f}", f"{v[1] / tot * 100:.1f}", f"{v[2] / max(v[0], 1) / 1e6:.2f}"])
    g = [(v[0], v[2]) for k, v in agg.items() if "gemm_tc_kernel" in k]
    n = sum(a for a, _ in g)
    return n, (sum(b for _, b in g) / n if n else 0.0)


if __name__ == "__main__":
    print(summarize(sys.argv[1], sys.argv[2], sys.argv[3]))
