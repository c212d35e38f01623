#!/usr/bin/env python3
"""Per-stage summary of an ncu source-page capture of the step kernel (SASS view).

usage: ncu_by_stage.py <report.ncu-rep> [n_env_substeps] [top-N]

The kernel body is split at the SM-clock stamps of the stage timers (CS2R ... SR_CLOCKLO, the same boundaries tools/stage_probe.py
reports) and the out-of-line functions at their RET instructions; the functions are named by what the kernel body calls between which
stamps.  Prints executed warp-instructions (per environment sub-step), stall samples and the stall reasons of every part, the
local / global memory instructions executed, and the most sampled instructions.
"""
import collections, csv, subprocess, sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 4096 * 4
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 25
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
print(rows[hi - 1][1] if hi else "")
hdr = rows[hi]
ci = {n: i for i, n in enumerate(hdr)}
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
recs, base = [], None
for r in rows[hi + 1:]:
    if len(r) < 10 or r[0] in ("Kernel Name", "Address"):
        break
    a = int(r[0], 16) if r[0].startswith("0x") else int(r[0])
    base = a if base is None else base
    f = lambda k: float(r[ci[k]] or 0)
    recs.append(dict(off=a - base, src=r[ci["Source"]].strip(), samp=int(f("# Samples")), ex=f("Instructions Executed"), thr=f("Thread Instructions Executed"),
                     st={k: int(f(k)) for k in stalls}))
tot_s = sum(x["samp"] for x in recs); tot_i = sum(x["ex"] for x in recs); tot_t = sum(x["thr"] for x in recs)
print(f"executed warp-instructions {tot_i:.0f} = {tot_i / units:.0f} per env-sub-step; stall samples {tot_s}; average active lanes {tot_t / tot_i:.1f}")


def opcode(src):
    t = src.split()
    return (t[1] if t[0].startswith("@") else t[0]).rstrip(";") if t else ""


clocks = [x["off"] for x in recs if "SR_CLOCKLO" in x["src"]]
rets = [x["off"] for x in recs if opcode(x["src"]).startswith("RET")]
main_end = rets[0] + 16
in_main = [c for c in clocks if c < main_end]
if len(in_main) >= 8:      # the in-kernel arrival wait of the fused gather reads the clock twice (bounded spin) at the very end: not stage stamps
    in_main = in_main[:-2]
names = ["prologue (TMA of the model block, state rows in)", "A  FK + RNEA + CRBA (incl. the sub-step barrier)", "B  narrow phase", "C  factorisation, z, Y, G", "D  call site",
         "E  v+, integration, records out", "epilogue / cold paths of the kernel body"]
if len(in_main) == 7:      # stage B compiled in line: its inner stamp splits it
    names[2:3] = ["B  candidates, sphere groups", "B  selection, contact records"]
parts = []
edges = [0] + in_main + [main_end]
for i, (a, b) in enumerate(zip(edges, edges[1:])):
    parts.append((a, b, names[i] if i < len(names) else f"part {i}"))
# out-of-line functions: [previous RET + 16, RET + 16); named by their callers
calls = collections.defaultdict(set)
for x in recs:
    if opcode(x["src"]).startswith("CALL") and x["off"] < main_end:
        tgt = int([t for t in x["src"].replace(";", " ").split() if t.startswith("0x")][-1], 16)
        tgt = tgt - base if tgt >= base else tgt
        seg = max(i for i, (a, b, n) in enumerate(parts) if a <= x["off"])
        calls[tgt].add(seg)
fstart = main_end
for rr in rets[1:]:
    callers = set()
    for tgt, segs in calls.items():
        if fstart <= tgt < rr + 16:
            callers |= segs
    nm = lambda c: parts[c][2][:1]
    first = nm(min(callers)) if callers else ""
    label = "helper"
    if callers:
        label = "D  gs_solve()" if first == "D" else ("B  out-of-line shape routine" if callers == {min(callers)} and first == "B" else "helper (called from %s)" % ",".join(sorted({nm(c) for c in callers})))
    parts.append((fstart, rr + 16, label))
    fstart = rr + 16
print(f"\n{'part':52s} {'instr/env-sub-step':>18s} {'instr %':>8s} {'samples %':>9s}   top stall reasons")
for a, b, n in parts:
    sel = [x for x in recs if a <= x["off"] < b]
    ii = sum(x["ex"] for x in sel); ss = sum(x["samp"] for x in sel)
    if ii == 0 and ss == 0:
        continue
    c = collections.Counter()
    for x in sel:
        c.update(x["st"])
    top = ", ".join(f"{k[6:]} {100 * v / max(ss, 1):.0f}%" for k, v in c.most_common(4))
    lm = sum(x["ex"] for x in sel if opcode(x["src"])[:3] in ("LDL", "STL")) / units
    gm = sum(x["ex"] for x in sel if opcode(x["src"]).split(".")[0] in ("LD", "LDG", "ST", "STG")) / units
    print(f"{n:52s} {ii / units:18.1f} {100 * ii / tot_i:7.1f}% {100 * ss / max(tot_s, 1):8.1f}%   {top}   [local {lm:.1f}, global {gm:.1f} per env-sub-step]")
print(f"\n-- {topn} most sampled instructions --")
for x in sorted(recs, key=lambda x: -x["samp"])[:topn]:
    part = [n for a, b, n in parts if a <= x["off"] < b]
    top = max(x["st"].items(), key=lambda kv: kv[1])
    print(f"  {x['off']:#8x} {100 * x['samp'] / tot_s:5.2f}%  {top[0][6:]:>14s}  {(part[0] if part else '?')[:28]:28s} {x['src'][:70]}")
