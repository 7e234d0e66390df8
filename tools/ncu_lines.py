"""Per-source-line instruction counts of one kernel: joins `ncu --page source --csv` (SASS view: executed instructions per instruction)
with `nvdisasm --print-line-info` of the same cubin (the n-th SASS instruction of the function is the n-th row of the ncu table).
usage: ncu_lines.py report.ncu-rep kernel_regex cubin mangled_function_substring [top]"""
import csv, re, subprocess, sys, collections
rep, kre, cubin, fn = sys.argv[1:5]
top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
tables, cur, hdr = [], None, None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = []; tables.append((r[1], cur)); hdr = None; continue
    if r and r[0] == "Address":
        hdr = r; continue
    if hdr and cur is not None and len(r) == len(hdr):
        cur.append(dict(zip(hdr, r)))
name, table = tables[0]
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
lines, infn, curline = [], False, "?"
for l in dis:
    if l.startswith(".text."):
        infn = fn in l
        continue
    if not infn:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        curline = "%s:%s" % (m.group(1).split("/")[-1], m.group(2)); continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+\S", l):
        lines.append(curline)
print(name, "ncu instructions", len(table), "nvdisasm instructions", len(lines))
n = min(len(table), len(lines))
acc, stall = collections.Counter(), collections.Counter()
for i in range(n):
    acc[lines[i]] += float(table[i]["Instructions Executed"] or 0)
    stall[lines[i]] += float(table[i]["# Samples"] or 0)
tot, stot = sum(acc.values()), sum(stall.values())
print("total warp instructions %.0f, samples %.0f" % (tot, stot))
for k, v in acc.most_common(top):
    print("%6.2f%% instr %6.2f%% samples  %s" % (100 * v / tot, 100 * stall[k] / max(stot, 1), k))
