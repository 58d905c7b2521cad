"""Aggregate an ncu report's per-SASS executed-instruction and sample counts by CUDA source line.

usage: python tools/ncu_by_line.py report.ncu-rep kernel_regex [top_n]
(needs a capture made with --import-source on and a -lineinfo build)"""
import collections
import csv
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", f"regex:{kern}"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    agg = collections.OrderedDict()
    fname, hdr = None, None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            iex, ismp = hdr.index("Instructions Executed"), hdr.index("# Samples")
            continue
        if hdr is None or len(r) <= iex:
            continue
        try:
            ex, smp = int(r[iex] or 0), int(r[ismp] or 0)
        except ValueError:
            continue
        key = (fname, r[0], r[1].strip()[:110])
        a = agg.setdefault(key, [0, 0, 0])
        a[0] += ex; a[1] += smp; a[2] += 1
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[1] for a in agg.values()) or 1
    print(f"total warp-instructions {tot}, samples {tots}")
    for (f, ln, src), (ex, smp, n) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * ex / tot:5.1f}% inst {100 * smp / tots:5.1f}% smp {n:4d} sass  {f}:{ln}  {src}")


if __name__ == "__main__":
    main()
