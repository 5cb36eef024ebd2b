#!/usr/bin/env python
"""Attribute ncu warp-stall samples to CUDA source lines.

    python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> [launch-skip] [top]

ncu's CSV source page lists SASS instructions with their sample counts but no line numbers;
`nvdisasm -g` of the cubin embedded in the library gives the line of every instruction.  The two
listings are joined by instruction order within the function.
"""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "real-time-multi-object-detection---tracking-system_b200", "librtmodt_b200.so")


def sass_lines(kernel_regex):
    """[(line, sass text)] of the first function whose mangled name matches, via nvdisasm -g."""
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", LIB], cwd=tmp, check=True, capture_output=True)
    out = []
    for f in sorted(os.listdir(tmp)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        cur_fn, line, take = None, None, False
        for l in txt.splitlines():
            m = re.match(r"\s*\.text\.(\S+):", l)
            if m:
                cur_fn = m.group(1)
                take = re.search(kernel_regex, cur_fn) is not None and not out
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                line = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if take and re.match(r"\s*/\*[0-9a-f]{4,}\*/", l):
                out.append((line, l.split("*/", 1)[1].strip().rstrip(";").strip()))
        if out:
            break
    return out


def main():
    rep, rx = sys.argv[1], sys.argv[2]
    skip = sys.argv[3] if len(sys.argv) > 3 else "0"
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 25
    csv_txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}",
                              "--launch-skip", skip, "--launch-count", "1"], capture_output=True, text=True).stdout
    rows = list(csv.reader(csv_txt.splitlines()))
    hdr = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
    si, ci = rows[hdr].index("Source"), rows[hdr].index("# Samples")
    ei = rows[hdr].index("Instructions Executed")
    inst = [(r[si].strip(), int(r[ci] or 0), int(r[ei] or 0)) for r in rows[hdr + 1:] if len(r) > ci and r[ci].isdigit()]
    lines = sass_lines(rx)
    if len(inst) > len(lines) and len(inst) % len(lines) == 0:      # ncu lists the function once per launch in the filter
        inst = inst[:len(lines)]
    if len(inst) != len(lines):
        print(f"warning: {len(inst)} profiled instructions vs {len(lines)} disassembled", file=sys.stderr)
    agg = collections.defaultdict(lambda: [0, 0])
    for (src, n, ex), (line, _) in zip(inst, lines):
        agg[line][0] += n
        agg[line][1] += ex
    total = sum(v[0] for v in agg.values()) or 1
    cache = {}
    print(f"{total} samples, {sum(v[1] for v in agg.values())} warp-instructions")
    for line, (n, ex) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        text = ""
        if line:
            path = next((os.path.join(dp, line[0]) for dp, _, fs in os.walk(ROOT) if line[0] in fs), None)
            if path:
                cache.setdefault(path, open(path).read().splitlines())
                text = cache[path][line[1] - 1].strip()[:100] if line[1] <= len(cache[path]) else ""
        print(f"{100 * n / total:5.1f}%  exec={ex:8d}  {line[0] if line else '?'}:{line[1] if line else 0:<5d} {text}")


if __name__ == "__main__":
    main()
