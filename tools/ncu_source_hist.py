#!/usr/bin/env python
"""Developer script: aggregate the ncu source page (per-SASS-instruction stall samples) by
opcode and list the hottest instructions.  usage: ncu_source_hist.py <src.csv>"""
import csv, sys, re, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
S, E = ix["# Samples"], ix["Instructions Executed"]
byop = collections.Counter(); byop_exec = collections.Counter(); tot = 0; totexec = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
st_tot = collections.Counter()
inst = []
for r in rows[2:]:
    if len(r) < len(hdr) or r[0] == "Address" or not r[S].isdigit(): continue
    src = r[ix["Source"]].strip()
    m = re.match(r"(@!?U?P\w+\s+)?([A-Z0-9_]+)", src)
    op = m.group(2) if m else "?"
    s = int(r[S] or 0); e = int(r[E] or 0)
    byop[op] += s; byop_exec[op] += e; tot += s; totexec += e
    for h in stall_cols: st_tot[h] += int(r[ix[h]] or 0)
    inst.append((s, e, src, {h: int(r[ix[h]] or 0) for h in stall_cols}))
print("static instructions:", len(inst), " samples:", tot, " warp-inst executed:", totexec)
print("\nby opcode: samples%  exec%  samples/exec(k)")
for op, s in byop.most_common(22):
    print(f"  {op:10s} {100*s/tot:6.2f}%  {100*byop_exec[op]/totexec:6.2f}%")
print("\nstall totals:")
for h, v in st_tot.most_common(10): print(f"  {h:28s} {100*v/tot:6.2f}%")
print("\nhottest instructions:")
for s, e, src, st in sorted(inst, key=lambda x: -x[0])[:25]:
    top = sorted(st.items(), key=lambda kv: -kv[1])[:2]
    print(f"  {s:6d} {e:9d}  {src[:60]:60s} {top}")
