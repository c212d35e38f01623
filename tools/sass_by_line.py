#!/usr/bin/env python3
"""Static SASS size per source line of the step kernel (code-size / I-cache budget)."""
import os, re, subprocess, sys, tempfile, collections
so = sys.argv[1]; kern = sys.argv[2] if len(sys.argv) > 2 else "rsb_step_kernelILi28"
tmp = tempfile.mkdtemp()
subprocess.check_call(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, stdout=subprocess.DEVNULL)
cubin = [f for f in os.listdir(tmp) if f.startswith("batch") and "model" not in f][0]
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
cnt_outer, cnt_inner = collections.Counter(), collections.Counter()
in_k, cur, total = False, (0, 0), 0
for ln in sass.splitlines():
    if ln.strip().startswith(".section") and ".text." in ln:
        in_k = kern in ln
    m = re.search(r'//## File "([^"]+)", line (\d+)(?: inlined at "([^"]+)", line (\d+))?', ln)
    if m:
        inner = int(m.group(2)); outer = int(m.group(4)) if m.group(4) else inner
        cur = (inner, outer); continue
    if in_k and re.match(r"\s+/\*[0-9a-f]{4,}\*/", ln):
        cnt_outer[cur[1]] += 1; cnt_inner[cur[0]] += 1; total += 1
src = open(os.path.join(os.path.dirname(os.path.abspath(so)), "csrc", "step_kernel.cuh")).read().splitlines()
print("total SASS instructions", total, "=", total * 16 // 1024, "KB")
marks = [(i + 1, l.strip()) for i, l in enumerate(src) if "=====" in l and "stage" in l]
marks = [(1, "prologue / helpers")] + marks + [(len(src) + 1, "end")]
for (a, name), (b, _) in zip(marks, marks[1:]):
    print(f"  lines {a:4d}-{b-1:4d}: {sum(v for k, v in cnt_outer.items() if a <= k < b):6d}  {name[:60]}")
print("top lines (outer):")
for line, c in cnt_outer.most_common(int(sys.argv[3]) if len(sys.argv) > 3 else 30):
    print(f"  {line:4d} {c:5d}  {src[line-1].strip()[:120] if 0 < line <= len(src) else '?'}")
