"""Aggregate an `ncu --page source --print-source cuda,sass --csv` dump by CUDA source line: share of executed warp
instructions, share of stall samples, average active threads per instruction.  Usage: ncu_source_lines.py rep [top]"""
import csv
import subprocess
import sys


def main(rep, top=40):
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = next(r for r in rows if r and r[0] == "Line No")
    iI, iT, iS = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"), hdr.index("# Samples")
    agg, cur = {}, None
    for r in rows:
        if len(r) <= iI or r[0] == "Line No":
            continue
        if r[0] != "":
            cur = (r[0], r[1].strip()[:120])
            continue
        try:
            n, t, s = int(r[iI]), int(r[iT]), int(r[iS])
        except ValueError:
            continue
        a = agg.setdefault(cur, [0, 0, 0])
        a[0] += n; a[1] += t; a[2] += s
    tot = sum(a[0] for a in agg.values()) or 1
    tots = sum(a[2] for a in agg.values()) or 1
    print("total warp instructions %d, samples %d" % (tot, tots))
    for (line, src), (n, t, s) in sorted(agg.items(), key=lambda kv: -kv[1][2])[:top]:
        print("%5.1f%% inst %5.1f%% samples  %4.1f thr/inst  L%-5s %s" % (100.0 * n / tot, 100.0 * s / tots, t / max(n, 1), line, src))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
