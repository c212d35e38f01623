#!/usr/bin/env python3
"""Correlate an ncu source-page CSV (SASS view) with CUDA source lines via nvdisasm line info.

usage: ncu_by_line.py <report.ncu-rep> <librsb.so> [kernel-substring] [top-N]
Prints executed warp-instructions and stall samples aggregated per source line (outermost,
i.e. the line in the kernel body that an inlined callee was expanded at) and per stage.
"""
import csv, os, re, subprocess, sys, tempfile, collections

rep, so = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "rsb_step_kernelILi28"
topn = int(sys.argv[4]) if len(sys.argv) > 4 else 40
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith("batch") and "model" not in f][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
addr2line = {}
in_k, cur = False, None
for ln in sass.splitlines():
    if ln.startswith("\t.section") or ".text." in ln and ln.strip().startswith(".section"):
        in_k = kern in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        inner = int(m.group(2))
        outer = int(m.group(4)) if m.group(4) else inner
        ofile = os.path.basename(m.group(3) if m.group(3) else m.group(1))
        cur = (inner, (ofile, outer))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/", ln)
    if m and in_k and cur:
        addr2line[int(m.group(1), 16)] = cur
csvtxt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(csvtxt.splitlines()))
# the report may hold several launches; take the first kernel section
hdr_i = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hdr_i]
ci = {n: hdr.index(n) for n in ("Address", "Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
by_outer, by_inner = collections.Counter(), collections.Counter()
samp_outer = collections.Counter()
total = tot_s = tot_thr = 0
base = None
for r in rows[hdr_i + 1:]:
    if not r or r[0] in ("Kernel Name", "Address"):
        break
    a = int(r[ci["Address"]], 16) if r[ci["Address"]].startswith("0x") else int(r[ci["Address"]])
    if base is None:
        base = a
    ex = int(float(r[ci["Instructions Executed"]] or 0)); sm = int(float(r[ci["# Samples"]] or 0))
    thr = int(float(r[ci["Thread Instructions Executed"]] or 0))
    inner, outer = addr2line.get(a - base, (0, ("?", 0)))
    by_outer[outer] += ex; by_inner[inner] += ex; samp_outer[outer] += sm
    total += ex; tot_s += sm; tot_thr += thr
csrc = os.path.join(os.path.dirname(os.path.abspath(so)), "csrc")
files = {f: open(os.path.join(csrc, f)).read().splitlines() for f in os.listdir(csrc) if f.endswith((".cuh", ".cu", ".hpp"))}
src = files["step_kernel.cuh"]
print(f"total warp-instructions {total}  samples {tot_s}  avg active threads {tot_thr / max(total, 1):.1f}")
# stage boundaries from the marker comments in the kernel
marks = [(i + 1, l.strip()) for i, l in enumerate(src) if "=====" in l and "stage" in l]
marks = [(1, "prologue / loads")] + marks + [(len(src) + 1, "end")]
print("\n-- by stage (outer line) --")
for (a, name), (b, _) in zip(marks, marks[1:]):
    ex = sum(v for k, v in by_outer.items() if k[0] == "step_kernel.cuh" and a <= k[1] < b); sm = sum(v for k, v in samp_outer.items() if k[0] == "step_kernel.cuh" and a <= k[1] < b)
    print(f"  lines {a:4d}-{b - 1:4d}  instr {ex:11d} ({100.0 * ex / total:5.1f}%)  samples {100.0 * sm / max(tot_s, 1):5.1f}%  {name[:70]}")
print(f"\n-- top {topn} kernel-body lines by executed warp-instructions --")
for f in sorted(set(k[0] for k in by_outer) - {"step_kernel.cuh"}):
    ex = sum(v for k, v in by_outer.items() if k[0] == f); sm = sum(v for k, v in samp_outer.items() if k[0] == f)
    print(f"  {f:28s} instr {ex:11d} ({100.0 * ex / total:5.1f}%)  samples {100.0 * sm / max(tot_s, 1):5.1f}%  (out-of-line functions of this file)")
for (f, line), ex in by_outer.most_common(topn):
    text = files[f][line - 1].strip() if f in files and 0 < line <= len(files[f]) else "?"
    print(f"  {f[:14]:14s}{line:5d} {ex:11d} {100.0 * ex / total:5.1f}%  samp {100.0 * samp_outer[(f, line)] / max(tot_s, 1):5.1f}%  {text[:100]}")
