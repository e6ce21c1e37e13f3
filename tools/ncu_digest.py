"""Digest of an ncu report: key raw metrics of every captured kernel and its hottest stall sites (SASS level).
usage: python tools/ncu_digest.py report.ncu-rep [title] > profiles/xxx.txt   (needs the `ncu` CLI, no GPU)"""
import csv
import io
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
        "launch__registers_per_thread", "launch__block_size", "launch__grid_size"]


def page(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", name, "--csv"], capture_output=True, text=True).stdout
    return list(csv.reader(io.StringIO(out)))


def main():
    rep = sys.argv[1]
    print(f"# {sys.argv[2] if len(sys.argv) > 2 else rep}")
    print("# ncu --set full --clock-control none --import-source on (cold-cache, serialised replay: read shares and ratios)")
    raw = page(rep, "raw")
    hdr, units = raw[0], raw[1]
    for r in raw[2:]:
        d = dict(zip(hdr, r))
        print(f"\nkernel: {d['Kernel Name'][:110]}")
        for w in WANT:
            if w in d:
                print(f"  {w:72s} {d[w]:>16s} {units[hdr.index(w)]}")
        for k in hdr:
            if "pcsamp_warps_issue_stalled" in k and "not_issued" not in k:
                try:
                    if float(d[k]) > 200:
                        print(f"  {k:72s} {d[k]:>16s} warp samples")
                except ValueError:
                    pass
    src = page(rep, "source")
    if len(src) > 2:
        h = src[1]
        try:
            i_s, i_n, i_e = h.index("Source"), h.index("# Samples"), h.index("Instructions Executed")
        except ValueError:
            return
        rows = []
        for r in src[2:]:
            try:
                rows.append((r[i_s].strip(), int(r[i_n]), int(r[i_e])))
            except (ValueError, IndexError):
                pass
        n = len(rows)
        if n > 20 and rows[:10] == rows[n // 2:n // 2 + 10]:
            rows = rows[:n // 2]
        tot = sum(x[1] for x in rows) or 1
        print(f"\nhottest SASS sites of the last kernel ({tot} warp samples, {len(rows)} instructions):")
        print("  index  samples  share  executed  instruction")
        for k in sorted(sorted(range(len(rows)), key=lambda k: -rows[k][1])[:14]):
            print(f"  {k:5d} {rows[k][1]:8d} {100 * rows[k][1] / tot:5.1f}% {rows[k][2]:9d}  {rows[k][0][:80]}")


if __name__ == "__main__":
    main()
