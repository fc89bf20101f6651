"""Executed warp instructions and stall samples of one kernel aggregated by CUDA source line:
the ncu SASS page (per-instruction counts, by address) joined with nvdisasm --print-line-info
of the object the kernel was built from.
usage: ncu_lines.py report.ncu-rep kernel_regex build/file.o [top]"""
import csv, subprocess, sys, io, collections, re, tempfile, os, glob
rep, rx, obj = sys.argv[1], sys.argv[2], sys.argv[3]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{rx}"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
blocks, cur = [], None
for r in rows:
    if r and r[0] == "Kernel Name":
        cur = {"name": r[1], "rows": []}; blocks.append(cur)
    elif cur is not None:
        cur["rows"].append(r)
b = blocks[0]
hdr = b["rows"][0]; data = b["rows"][1:]
ix = {h: i for i, h in enumerate(hdr)}
base = int(data[0][ix["Address"]], 16)
per_off = {}
for r in data:
    try: per_off[int(r[ix["Address"]], 16) - base] = (int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]] or 0))
    except Exception: pass
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=tmp, capture_output=True)
cubin = glob.glob(tmp + "/*.cubin")[0]
dis = subprocess.run(["nvdisasm", "--print-line-info", "-c", cubin], capture_output=True, text=True).stdout
short = re.sub(r"[^A-Za-z0-9_]", "", b["name"].split("(")[0].split("::")[-1].split("<")[0])
# several template instances may share the name: take the section whose size matches the report
secs, cur = {}, None
for l in dis.splitlines():
    if l.startswith("\t.section\t.text."):
        cur = l if short in l else None
        if cur: secs[cur] = []
        continue
    if cur is not None: secs[cur].append(l)
def ninstr(ls): return sum(1 for l in ls if re.match(r"\s*/\*([0-9a-f]{4,})\*/", l))
best = min(secs, key=lambda k: abs(ninstr(secs[k]) - len(per_off)))
line, agg, ins = None, collections.Counter(), collections.Counter()
for l in secs[best]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: line = (os.path.basename(m.group(1)), int(m.group(2))); continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
    if m and line:
        off = int(m.group(1), 16)
        if off in per_off:
            agg[line] += per_off[off][0]; ins[line] += per_off[off][1]
tot = sum(agg.values()); stot = sum(ins.values()) or 1
print(b["name"][:70], "| warp instr", tot, "| samples", stot)
srcs = {}
for (f, ln), n in agg.most_common(top):
    if f not in srcs:
        p = [q for q in glob.glob(os.path.dirname(os.path.abspath(obj)) + "/../" + f)]
        srcs[f] = open(p[0]).read().splitlines() if p else []
    text = srcs[f][ln - 1].strip()[:80] if srcs[f] and ln <= len(srcs[f]) else ""
    print(f"  {100*n/tot:5.1f}% instr {100*ins[(f,ln)]/stot:5.1f}% smp  {f}:{ln:<4d} {text}")
