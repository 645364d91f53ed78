#!/usr/bin/env python
"""Join the SASS page of an ncu report with nvdisasm line info: instructions executed and stall
samples per CUDA source line of one kernel.

    cuobjdump -xelf all cns_b200/libcns_sm100.so; nvdisasm -g -c encoder_tc.sm_100a.cubin > enc.sass
    ncu -i rep.ncu-rep --page source --csv > src.csv
    python scripts/ncu_by_line.py src.csv enc.sass _ZN3cns18enc_edge_tc_kernelILi32ELb0E cns_b200/csrc/encoder_tc.cu
"""
import csv, re, sys, collections

src_csv, sass, mangled, cu = sys.argv[1:5]
rows = list(csv.reader(open(src_csv)))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
inst = rows[2:]
# nvdisasm: walk the function, remember current line for each instruction
lines, cur, on = [], None, False
for l in open(sass):
    if l.startswith(".text."):
        on = mangled in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append((cur, l.split("*/", 1)[1].strip()))
print("sass instructions: ncu %d, nvdisasm %d" % (len(inst), len(lines)))
n = min(len(inst), len(lines))
agg = collections.defaultdict(lambda: [0, 0, 0, collections.Counter()])
tot_i = tot_s = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for k in range(n):
    r = inst[k]
    ie, sm = int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]])
    a = agg[lines[k][0]]
    a[0] += ie; a[1] += sm; a[2] += 1
    for h in stall_cols:
        v = int(r[ix[h]] or 0)
        if v:
            a[3][h] += v
    tot_i += ie; tot_s += sm
text = {}
try:
    for i, l in enumerate(open(cu), 1):
        text[i] = l.rstrip()
except Exception:
    pass
print("total inst %d samples %d" % (tot_i, tot_s))
top = sorted(agg.items(), key=lambda kv: -kv[1][1])[: int(sys.argv[5]) if len(sys.argv) > 5 else 45]
for key, (ie, sm, cnt, st) in top:
    f, ln = key if key else ("?", 0)
    print("%5.1f%% smp %5.1f%% inst  %-16s:%-4d  %s  | %s" % (100 * sm / tot_s, 100 * ie / tot_i, f, ln,
          ", ".join("%s %d" % (h[6:], v) for h, v in st.most_common(3)), text.get(ln, "")[:90] if f.endswith(cu.split("/")[-1]) else ""))
