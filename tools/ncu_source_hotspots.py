"""Warp-stall samples of one kernel by CUDA source line, from an `ncu --set full --import-source on` report.

  python tools/ncu_source_hotspots.py <report.ncu-rep> <object.o> <kernel substring> [top_n]

The report's SASS page (`ncu -i ... --page source --csv`) carries the samples per instruction; the line each
instruction belongs to comes from `nvdisasm -g` of the same object (compiled with -lineinfo), matched by the
instruction's offset inside the kernel.  Build the object from the commit the report was captured on.
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, obj, pattern = sys.argv[1:4]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 30

# -- samples per SASS offset --------------------------------------------------------------------------------------
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True, check=True).stdout
blocks, cur = [], None
for line in raw.splitlines():
    if line.startswith('"Kernel Name"'):
        cur = [line]
        blocks.append(cur)
    elif cur is not None:
        cur.append(line)
samples = None
for b in blocks:
    name = next(csv.reader([b[0]]))[1]
    if pattern not in name:
        continue
    rows = list(csv.DictReader(io.StringIO("\n".join(b[1:]))))
    base = int(rows[0]["Address"], 16)
    samples = [(int(r["Address"], 16) - base, r) for r in rows]
    kernel_name = name
    break
if samples is None:
    sys.exit(f"no kernel matching {pattern!r} in {rep}")

# -- offset -> (file, line) from nvdisasm -g -----------------------------------------------------------------------
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True, check=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
line_of, where, in_kernel = {}, None, False
mangled = None
for line in dis.splitlines():
    m = re.match(r"\s*\.text\.(\S+):", line)
    if m:
        demangled = subprocess.run(["cu++filt", m.group(1)], capture_output=True, text=True).stdout.strip() or m.group(1)
        in_kernel = demangled.replace("(int)", "").replace(" ", "") == kernel_name.replace("(int)", "").replace(" ", "")
        continue
    if not in_kernel:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', line)
    if m:
        where = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", line)
    if m and where:
        line_of[int(m.group(1), 16)] = where
if not line_of:
    sys.exit("the object holds no line information for this kernel (compile with -lineinfo)")

# -- aggregate ---------------------------------------------------------------------------------------------------
stall_cols = [c for c in samples[0][1] if c.startswith("stall_") and "Not Issued" not in c]
by_line = collections.defaultdict(lambda: collections.Counter())
total = 0
for off, r in samples:
    n = int(r["# Samples"] or 0)
    if not n:
        continue
    w = line_of.get(off, ("?", 0))
    by_line[w]["n"] += n
    total += n
    for c in stall_cols:
        by_line[w][c[6:]] += int(r[c] or 0)
src_cache = {}


def text(w):
    f, ln = w
    for root in ("mobrob_b200/csrc",):
        p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), root, f)
        if os.path.exists(p):
            if p not in src_cache:
                src_cache[p] = open(p).read().splitlines()
            return src_cache[p][ln - 1].strip()[:110] if 0 < ln <= len(src_cache[p]) else ""
    return ""


print(f"{kernel_name}: warp-stall samples by CUDA source line ({total} samples; ncu --set full, SASS page mapped to lines "
      f"through nvdisasm -g of {os.path.basename(obj)})")
for w, c in sorted(by_line.items(), key=lambda kv: -kv[1]["n"])[:top_n]:
    top = ", ".join(f"{k}={v}" for k, v in c.most_common(4) if k != "n")
    print(f"  {100.0 * c['n'] / total:4.1f}%  {w[0]}:{w[1]}  {text(w)}   [{top}]")
