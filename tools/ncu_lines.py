"""Development aid: attribute an ncu SASS-level source page to CUDA source lines, with inline call chains.

usage: ncu_lines.py <source-page.csv> <nvdisasm -gi output> <mangled kernel name> [top]
  source-page.csv  = `ncu -i rep --page source --csv`
  nvdisasm output  = `nvdisasm -gi -c <cubin>` of the SAME build
Prints the hottest (stall samples, executed warp instructions, L2 sectors) innermost lines and, per line of the
outermost step file (f2d_step.h), the inclusive totals: where a phase spends its issue slots and memory traffic.
"""
import csv
import re
import sys
from collections import defaultdict

page, sass, kernel = sys.argv[1:4]
BUSY = "--busy" in sys.argv  # count only the samples of warps that are not waiting at a barrier
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40

# ---- address offset -> inline chain [(file, line) innermost first]
chains = {}
cur = []
pending = []
infunc = False
line_re = re.compile(r'//## File "([^"]+)", line (\d+)')
ins_re = re.compile(r"/\*([0-9a-f]{4,})\*/")
for ln in open(sass):
    if ln.startswith(".text."):
        infunc = ln.strip() == ".text.%s:" % kernel
        continue
    if not infunc:
        continue
    m = line_re.search(ln)
    if m:
        pending.append((m.group(1).split("/")[-1], int(m.group(2))))
        continue
    m = ins_re.search(ln)
    if m:
        if pending:
            cur = pending
            pending = []
        chains[int(m.group(1), 16)] = cur

rows = list(csv.reader(open(page)))
hdr = None
for i, r in enumerate(rows):
    if r and r[0] == "Address":
        hdr = r
        body = rows[i + 1:]
        break
ci = {n: k for k, n in enumerate(hdr)}
base = int(body[0][0], 16)

inner = defaultdict(lambda: [0, 0, 0, 0])
outer = defaultdict(lambda: [0, 0, 0, 0])
phase = defaultdict(lambda: [0, 0, 0, 0])
tot = [0, 0, 0, 0]


def num(r, name):
    v = r[ci[name]] if name in ci and ci[name] < len(r) else "0"
    try:
        return int(float(v))
    except ValueError:
        return 0


for r in body:
    if len(r) < 5:
        continue
    off = int(r[0], 16) - base
    chain = chains.get(off, [("?", 0)])
    vals = [num(r, "Warp Stall Sampling (All Samples)") - (num(r, "stall_barrier") if BUSY else 0), num(r, "Instructions Executed"),
            num(r, "L2 Theoretical Sectors Global"), num(r, "stall_long_sb")]
    for k in range(4):
        tot[k] += vals[k]
    key_in = chain[0]
    step = [c for c in chain if c[0] == "f2d_step.h"]
    # the outermost frame inside a step phase function (skip the stepWorld dispatcher lines)
    key_out = step[-2] if len(step) >= 2 else (step[-1] if step else chain[-1])
    # top-level phase: the call site inside stepWorld (or the outermost frame of a non-inlined function)
    key_top = step[-1] if step else ("noinline:" + chain[-1][0], chain[-1][1])
    for k in range(4):
        phase[key_top][k] += vals[k]
    for k in range(4):
        inner[key_in][k] += vals[k]
        outer[key_out][k] += vals[k]

print("totals: samples %d, warp-instructions %d, L2 sectors %d, long-scoreboard samples %d" % tuple(tot))
for title, table in (("top-level call sites (inclusive)", phase), ("phase-level lines (inclusive)", outer), ("innermost lines", inner)):
    print("\n== %s, by stall samples" % title)
    print("%-28s %9s %6s %12s %6s %12s %6s" % ("file:line", "samples", "%", "warp-inst", "%", "L2 sectors", "%"))
    for key, v in sorted(table.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-28s %9d %6.2f %12d %6.2f %12d %6.2f" % ("%s:%d" % key, v[0], 100.0 * v[0] / max(1, tot[0]), v[1], 100.0 * v[1] / max(1, tot[1]),
                                                      v[2], 100.0 * v[2] / max(1, tot[2])))

# optional: --under file:line  -> innermost-line table restricted to instructions whose chain contains that frame
if "--under" in sys.argv:
    f, l = sys.argv[sys.argv.index("--under") + 1].split(":")
    want = (f, int(l))
    sub = defaultdict(lambda: [0, 0, 0, 0])
    stot = [0, 0, 0, 0]
    for r in body:
        if len(r) < 5:
            continue
        off = int(r[0], 16) - base
        chain = chains.get(off, [("?", 0)])
        if want not in chain:
            continue
        vals = [num(r, "Warp Stall Sampling (All Samples)"), num(r, "Instructions Executed"), num(r, "L2 Theoretical Sectors Global"),
                num(r, "L2 Theoretical Sectors Local")]
        k = chain.index(want)
        key = chain[k - 1] if k > 0 else chain[0]
        for q in range(4):
            sub[key][q] += vals[q]
            stot[q] += vals[q]
    print("\n== callees directly under %s:%d: samples %d (%.2f%% of all), warp-inst %d, L2 global sectors %d, local %d" % (
        want[0], want[1], stot[0], 100.0 * stot[0] / tot[0], stot[1], stot[2], stot[3]))
    for key, v in sorted(sub.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-28s %9d %6.2f %12d %12d %12d" % ("%s:%d" % key, v[0], 100.0 * v[0] / max(1, stot[0]), v[1], v[2], v[3]))

# optional: --functions  -> inclusive totals per function of f2d_step.h (outermost frame below stepWorld)
if "--functions" in sys.argv:
    import os
    src = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "forge2d_b200", "csrc", "f2d_step.h")
    starts = []
    fre = re.compile(r"inline\s+\w+\s+(\w+)\s*\(")
    for k, ln in enumerate(open(src), 1):
        m = fre.search(ln)
        if m and not ln.startswith("\t\t"):
            starts.append((k, m.group(1)))

    def fn_of(line):
        name = "?"
        for k, n in starts:
            if k <= line:
                name = n
        return name

    ft = defaultdict(lambda: [0, 0, 0, 0])
    for r in body:
        if len(r) < 5:
            continue
        off = int(r[0], 16) - base
        chain = chains.get(off, [("?", 0)])
        vals = [num(r, "Warp Stall Sampling (All Samples)"), num(r, "Instructions Executed"), num(r, "L2 Theoretical Sectors Global"),
                num(r, "L2 Theoretical Sectors Local")]
        step = [c for c in chain if c[0] == "f2d_step.h" and fn_of(c[1]) != "stepWorld"]
        if step:
            key = fn_of(step[-1][1]) + " > " + (fn_of(step[-2][1]) if len(step) > 1 else "-")
        else:
            key = "noinline " + chain[-1][0]
        for q in range(4):
            ft[key][q] += vals[q]
    print("\n== per function (outermost frame below stepWorld > next frame)")
    print("%-52s %9s %6s %12s %6s %12s %12s" % ("function", "samples", "%", "warp-inst", "%", "L2 global", "L2 local"))
    for key, v in sorted(ft.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-52s %9d %6.2f %12d %6.2f %12d %12d" % (key, v[0], 100.0 * v[0] / max(1, tot[0]), v[1], 100.0 * v[1] / max(1, tot[1]), v[2], v[3]))

# optional: --callers file:line -> who calls into that frame (next outer frame), by samples
if "--callers" in sys.argv:
    f, l = sys.argv[sys.argv.index("--callers") + 1].split(":")
    want = (f, int(l))
    sub = defaultdict(lambda: [0, 0, 0, 0])
    for r in body:
        if len(r) < 5:
            continue
        off = int(r[0], 16) - base
        chain = chains.get(off, [("?", 0)])
        if want not in chain:
            continue
        k = chain.index(want)
        key = chain[k + 1] if k + 1 < len(chain) else ("<top>", 0)
        vals = [num(r, "Warp Stall Sampling (All Samples)"), num(r, "Instructions Executed"), num(r, "L2 Theoretical Sectors Global"), 0]
        for q in range(4):
            sub[key][q] += vals[q]
    print("\n== callers of %s:%d" % want)
    for key, v in sorted(sub.items(), key=lambda kv: -kv[1][0])[:top]:
        print("%-28s %9d %12d %12d" % ("%s:%d" % key, v[0], v[1], v[2]))
