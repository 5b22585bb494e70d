"""Per-source-line instruction / stall-sample shares of one kernel from an `ncu --set full --import-source on` capture.

    ncu -i cap.ncu-rep --page source --csv > cap_src.csv          # SASS-level rows of the captured launch
    nvcc <the library's flags> -lineinfo -cubin -o k.cubin multi_agent_aac_b200/csrc/aac_kernels.cu
    nvdisasm -g -c k.cubin > k.sass                                # the same SASS with `//## File ..., line N` marks
    python tests/tools/hotspots.py cap_src.csv k.sass <mangled kernel name> multi_agent_aac_b200/csrc/aac_kernels.cu [units]

(`--page source --csv` prints every captured launch twice, SASS view then source view: pass the first block of the launch you
want.)  Lines of the other files next to the .cu (aac_radar.cuh) are resolved too; a per-function summary follows the lines.

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
import os
srcdir = os.path.dirname(cu)
texts = {}


def src_line(key):
    if not key:
        return "?"
    if key[0] not in texts:
        try:
            texts[key[0]] = open(os.path.join(srcdir, key[0])).read().split("\n")
        except OSError:
            texts[key[0]] = None
    t = texts[key[0]]
    return t[key[1] - 1].strip()[:90] if t and key[1] - 1 < len(t) else key[0]


def func_of(key):
    """Enclosing top-level function of a source line (last non-indented definition above it)."""
    if not key or key[0] not in texts or not texts[key[0]]:
        return "other:" + (key[0] if key else "?")
    t = texts[key[0]]
    for i in range(min(key[1], len(t)) - 1, -1, -1):
        m = re.match(r"^(?:__device__|__global__|static|AAC_HD|AAC_HD_NOINLINE|inline|template).*?\b([A-Za-z_0-9]+)\s*\(", t[i])
        if m and not t[i].startswith(" "):
            return key[0][:10] + ":" + ("env_kernel (body)" if m.group(1) == "__launch_bounds__" else m.group(1))
    return key[0]
print("%d SASS instructions, %d warp instructions executed%s, %d stall samples" %
      (len(seq), tot_i, (" = %.1f per unit" % (tot_i / units)) if units else "", tot_s))
print("top source lines by stall samples (inst share | sample share | threads per instruction%s | line | source | top stall reasons)" %
      (" | warp-inst per unit" if units else ""))
for key, s in smp.most_common(60):
    line = src_line(key)
    per = (" %5.1f |" % (inst[key] / units)) if units else ""
    print("%5.2f%% inst %5.2f%% smp thr %4.1f |%s %4s | %-90s | %s" %
          (100.0 * inst[key] / tot_i, 100.0 * s / tot_s, thr[key] / max(inst[key], 1), per, key[1] if key else "", line,
           ",".join("%s:%d" % kv for kv in why[key].most_common(3))))
allw = collections.Counter()
for c in why.values():
    allw.update(c)
print("ALL stalls: " + ", ".join("%s:%.1f%%" % (k, 100.0 * v / max(sum(allw.values()), 1)) for k, v in allw.most_common(10)))

byf_i, byf_s = collections.Counter(), collections.Counter()
for key in inst:
    src_line(key)
    f = func_of(key)
    byf_i[f] += inst[key]
    byf_s[f] += smp[key]
print("by function (warp-inst%s | stall-sample share):" % (" per unit" if units else ""))
for f, v in byf_i.most_common(22):
    print("  %-40s %8.2f  %5.1f%%" % (f, v / units if units else v, 100.0 * byf_s[f] / max(tot_s, 1)))
