"""DRAM bytes of ONE training step, per kernel, from
  ncu --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
      --clock-control none --csv --log-file step_dram.csv python bench.py --ncu-step --no-graph --no-cpu-baseline
usage: python tools/step_dram.py step_dram.csv [title] > profiles/rN_step_dram.txt"""
import csv
import sys


def main():
    rows = list(csv.reader(open(sys.argv[1])))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    h = rows[hi]
    ik, im, iv, iu = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("Metric Unit")
    agg, tot = {}, {"r": 0.0, "w": 0.0, "t": 0.0}
    for r in rows[hi + 1:]:
        if len(r) <= iv:
            continue
        v, u, m = float(r[iv].replace(",", "")), r[iu], r[im]
        a = agg.setdefault(r[ik][:100], {"n": 0, "r": 0.0, "w": 0.0, "t": 0.0})
        if m.startswith("dram"):
            v *= {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
            key = "r" if "read" in m else "w"
            if key == "r":
                a["n"] += 1
        else:
            v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
            key = "t"
        a[key] += v
        tot[key] += v
    print(f"# {sys.argv[2] if len(sys.argv) > 2 else sys.argv[1]}")
    print("# one eager training step between cudaProfilerStart/Stop (bench.py --ncu-step); ncu serialises and replays kernels, so the")
    print("# per-kernel times are cold-cache (compare shares); the byte counts are what the step moves through DRAM")
    print(f"# TOTAL: {(tot['r'] + tot['w']) / 1e9:.2f} GB per step (read {tot['r'] / 1e9:.2f} GB, write {tot['w'] / 1e9:.2f} GB), "
          f"{sum(a['n'] for a in agg.values())} kernels, {tot['t'] / 1e3:.2f} ms summed kernel time under ncu")
    print("#   GB     read    write   ms(ncu)  launches  kernel")
    for k, a in sorted(agg.items(), key=lambda kv: -(kv[1]["r"] + kv[1]["w"])):
        if a["r"] + a["w"] < 1e6:
            continue
        print(f"{(a['r'] + a['w']) / 1e9:7.3f} {a['r'] / 1e9:7.3f} {a['w'] / 1e9:7.3f} {a['t'] / 1e3:8.3f} {a['n']:6d}    {k}")


if __name__ == "__main__":
    main()
