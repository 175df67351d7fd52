"""Attribute the executed warp instructions of an ncu capture to CUDA source lines.

    python tools/ncu_lines.py REP CUBIN KERNEL_SUBSTRING [top]

REP: .ncu-rep of one kernel launch; CUBIN: the cubin of the SAME build (cuobjdump -xelf all
libturtle_b200.so), disassembled here with `nvdisasm -g -c` for its line table. Prints, per
source line (innermost frame) and per kernel-level line (outermost frame of the inlining
chain): warp instructions executed, share, average active threads, stall samples.
Development tool (reads profiles, no GPU)."""
import collections
import csv
import re
import subprocess
import sys


def main():
    rep, cubin, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], stdout=subprocess.PIPE, text=True).stdout
    lines = dis.splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and kernel in l)
    locs, cur_in, cur_out = [], ("?", 0), ("?", 0)
    for l in lines[start + 1:]:
        if l.startswith("\t.section") or l.startswith(".text."):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
        if m:
            cur_in = (m.group(1).split("/")[-1], int(m.group(2)))
            chain = re.findall(r'inlined at "([^"]+)", line (\d+)', m.group(3))
            cur_out = (chain[-1][0].split("/")[-1], int(chain[-1][1])) if chain else cur_in
            continue
        if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
            locs.append((cur_in, cur_out))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], stdout=subprocess.PIPE,
                         text=True).stdout
    rows = list(csv.reader([l for l in raw.splitlines() if not l.startswith("==")]))
    hdr, data = rows[1], rows[2:]
    ie, it, isamp = (hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed"),
                     hdr.index("# Samples"))
    if len(data) != len(locs):
        sys.stderr.write("instruction counts differ: capture %d, cubin %d -- not the same build\n"
                         % (len(data), len(locs)))
        sys.exit(1)
    total = sum(int(r[ie]) for r in data)
    for name, pick in (("innermost line", 0), ("kernel-level line", 1)):
        agg = collections.defaultdict(lambda: [0, 0, 0])
        for r, loc in zip(data, locs):
            a = agg[loc[pick]]
            a[0] += int(r[ie])
            a[1] += int(r[it])
            a[2] += int(r[isamp])
        print("== by %s (total %.2f G warp instructions)" % (name, total / 1e9))
        for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            print("%-28s %8.1f M  %5.1f %%  threads %5.1f  samples %d" % (
                "%s:%d" % k, a[0] / 1e6, 100. * a[0] / total, a[1] / max(a[0], 1), a[2]))


if __name__ == "__main__":
    main()
