"""Instruction mix of one kernel from an ncu report's source page: executed warp instructions
per opcode.  usage: ncu_opmix.py report.ncu-rep kernel_regex [top]"""
import csv, subprocess, sys, io, collections
rep, rx = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
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
mix = collections.Counter(); tot = 0
for r in data:
    try: n = int(r[ix["Instructions Executed"]])
    except: continue
    s = r[ix["Source"]].strip().split()
    op = s[1] if s and s[0].startswith("@") and len(s) > 1 else (s[0] if s else "?")
    op = op.split(".")[0]
    mix[op] += n; tot += n
print(b["name"][:80], "total warp instr", tot, "static", len(data))
for op, n in mix.most_common(top): print(f"  {op:12s} {n:12d} {100*n/tot:5.1f}%")
