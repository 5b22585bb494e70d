"""Per-source-line instruction / stall-sample shares of one kernel from an `ncu --set full --import-source on` capture.

    ncu -i cap.ncu-rep --page source --csv > cap_src.csv          # SASS-level rows of the captured launch
    nvcc <the library's flags> -lineinfo -cubin -o k.cubin multi_agent_aac_b200/csrc/aac_kernels.cu
    nvdisasm -g -c k.cubin > k.sass                                # the same SASS with `//## File ..., line N` marks
    python tests/tools/hotspots.py cap_src.csv k.sass <mangled kernel name> multi_agent_aac_b200/csrc/aac_kernels.cu [units]

Rows of the two listings are matched one to one (same build => same instruction sequence; the opcode of every row is
checked).  `units` (e.g. 655360 agent-steps per launch) adds a per-unit column."""
import collections
import csv
import re
import sys

src_csv, sass, kernel, cu = sys.argv[1:5]
units = float(sys.argv[5]) if len(sys.argv) > 5 else None

seq, cur, inside = [], None, False
for l in open(sass):
    if l.startswith(".text."):
        inside = l.strip() == ".text.%s:" % kernel
        continue
    if not inside:
        continue
    m = re.search(r'//## File "(.*?)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    m = re.match(r"\s+/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
    if m:
        seq.append((cur, m.group(2)))
rows = list(csv.reader(open(src_csv)))
hdr, data = rows[1], rows[2:]
col = {n: hdr.index(n) for n in ("Source", "# Samples", "Instructions Executed", "Thread Instructions Executed")}
stalls = [(i, n[6:]) for i, n in enumerate(hdr) if n.startswith("stall_") and "Not Issued" not in n]
assert len(seq) == len(data), (len(seq), len(data))
for (_, txt), r in zip(seq, data):
    op = (txt.split()[1] if txt.startswith("@") else txt.split()[0]).split(".")[0]
    assert op in r[col["Source"]], (txt, r[col["Source"]])

inst, smp, thr, why = collections.Counter(), collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
for (key, _), r in zip(seq, data):
    n = int(r[col["Instructions Executed"]])
    inst[key] += n
    smp[key] += int(r[col["# Samples"]])
    thr[key] += int(r[col["Thread Instructions Executed"]])
    for i, name in stalls:
        if r[i] and int(r[i]):
            why[key][name] += int(r[i])
tot_i, tot_s = sum(inst.values()), sum(smp.values())
text = open(cu).read().split("\n")
base = cu.split("/")[-1]
print("%d SASS instructions, %d warp instructions executed%s, %d stall samples" %
      (len(seq), tot_i, (" = %.1f per unit" % (tot_i / units)) if units else "", tot_s))
print("top source lines by stall samples (inst share | sample share | threads per instruction%s | line | source | top stall reasons)" %
      (" | warp-inst per unit" if units else ""))
for key, s in smp.most_common(60):
    line = text[key[1] - 1].strip()[:90] if key and key[0] == base else (key[0] if key else "?")
    per = (" %5.1f |" % (inst[key] / units)) if units else ""
    print("%5.2f%% inst %5.2f%% smp thr %4.1f |%s %4s | %-90s | %s" %
          (100.0 * inst[key] / tot_i, 100.0 * s / tot_s, thr[key] / max(inst[key], 1), per, key[1] if key else "", line,
           ",".join("%s:%d" % kv for kv in why[key].most_common(3))))
allw = collections.Counter()
for c in why.values():
    allw.update(c)
print("ALL stalls: " + ", ".join("%s:%.1f%%" % (k, 100.0 * v / max(sum(allw.values()), 1)) for k, v in allw.most_common(10)))
