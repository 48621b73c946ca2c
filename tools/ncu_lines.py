"""Per-source-line stall samples of one kernel from an ncu report: joins `ncu --page source --csv` (SASS rows with samples)
with `nvdisasm -g` line markers of the same cubin by instruction order.

    python tools/ncu_lines.py <report.ncu-rep> <kernel regex> <cubin> [launch-skip] [top-n] [section regex]

``section regex`` selects the function inside the cubin by its MANGLED name when the kernel regex matches several
instantiations (e.g. "match_tc_kernelILb0" for match_tc_kernel<false>).
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, kern, cubin = sys.argv[1:4]
skip = sys.argv[4] if len(sys.argv) > 4 else "0"
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 40
sect = sys.argv[6] if len(sys.argv) > 6 else kern
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kern}", "--launch-skip", skip,
                      "--launch-count", "1"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ci, cs = hdr.index("Source"), hdr.index("# Samples")
ce = hdr.index("Instructions Executed")
sass = []
for r in rows[hi + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):          # a second kernel's section follows: stop
        break
    if len(r) > cs:
        sass.append((r[ci].strip(), int(r[cs] or 0), int(r[ce] or 0)))
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout
# walk the listing of the matching function: line markers then instructions
lines, cur, infn = [], ("?", 0), False
for l in dis.splitlines():
    if l.startswith(".text.") or l.lstrip().startswith(".section"):
        infn = bool(re.search(sect, l))
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
if len(lines) != len(sass):
    print(f"warning: {len(lines)} disassembled instructions vs {len(sass)} profiled rows", file=sys.stderr)
agg, agi = collections.Counter(), collections.Counter()
for (f, ln), (_, smp, ex) in zip(lines, sass):
    agg[(f, ln)] += smp
    agi[(f, ln)] += ex
tot = sum(agg.values())
print(f"total samples {tot}, warp instructions {sum(agi.values())}")
for (f, ln), s in agg.most_common(topn):
    print(f"{100.0 * s / tot:6.2f}%  {s:8d} smp {agi[(f, ln)]:12d} inst  {f}:{ln}")
