#!/usr/bin/env python
"""Per-source-line summary of an `ncu --set full --import-source on` capture: joins the SASS page of the report with
`nvdisasm --print-line-info` of the library that was profiled (same build!) and prints instruction and warp-sample shares
per source line and per named line range.
Usage: python tools/ncu_source_summary.py REPORT.ncu-rep KERNEL_SUBSTRING [SOURCE.cu] [--top N]"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    src_path = sys.argv[3] if len(sys.argv) > 3 and not sys.argv[3].startswith("--") else None
    top = int(sys.argv[sys.argv.index("--top") + 1]) if "--top" in sys.argv else 25
    lib = os.path.join(ROOT, "visual_odometry_ros_b200", "libvo_b200.so")
    page = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(page)))
    print("kernel:", rows[0][1])
    hdr, data = rows[1], rows[2:]
    iS, iI, iSrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", lib], cwd=td, capture_output=True)
        sass = ""
        for f in sorted(os.listdir(td)):
            out = subprocess.run(["nvdisasm", "--print-line-info", os.path.join(td, f)], capture_output=True, text=True).stdout
            if kern in out:
                sass = out
                break
    lines, cur, inside = [], None, False
    want = rows[0][1].split("(")[0].replace("void ", "").strip()
    for l in sass.split("\n"):
        if l.startswith(".text."):
            inside = kern in l
            if inside and lines:
                break          # first matching function only unless the caller's substring is specific
            continue
        if not inside:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            lines.append((m.group(2), cur))
    if len(lines) != len(data):
        print(f"SASS of the current build has {len(lines)} instructions, the report {len(data)}: not the same build", file=sys.stderr)
        sys.exit(1)
    inst, samp = collections.Counter(), collections.Counter()
    for r, (_, c) in zip(data, lines):
        inst[c] += int(r[iI])
        samp[c] += int(r[iS])
    ti, ts = sum(inst.values()), max(1, sum(samp.values()))
    print(f"warp-instructions executed: {ti}, warp samples: {ts}")
    src = open(src_path).read().split("\n") if src_path else None
    print(f"{'file:line':24s} {'inst %':>7s} {'samples %':>9s}  source")
    for c, v in sorted(samp.items(), key=lambda x: -x[1])[:top]:
        text = ""
        if src and c and c[0] == os.path.basename(src_path) and c[1] <= len(src):
            text = src[c[1] - 1].strip()[:100]
        print(f"{(c[0] + ':' + str(c[1])) if c else '-':24s} {100 * inst[c] / ti:7.1f} {100 * v / ts:9.1f}  {text}")


if __name__ == "__main__":
    main()
