"""Summarise the ncu source page of one kernel: total samples per stall reason and the top
instructions by samples.  usage: ncu_stalls.py report.ncu-rep kernel_regex [top]"""
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
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = collections.Counter()
for r in data:
    for h in stall_cols:
        try: tot[h] += int(r[ix[h]])
        except: pass
alls = sum(tot.values())
print(b["name"][:90]); print("total samples", alls)
for h, v in tot.most_common(8): print(f"  {h:28s} {v:8d} {100*v/alls:5.1f}%")
data.sort(key=lambda r: -int(r[ix["# Samples"]] or 0))
print("top instructions:")
for r in data[:top]:
    st = sorted(((int(r[ix[h]] or 0), h) for h in stall_cols), reverse=True)[:2]
    print(f"  {int(r[ix['# Samples']]):7d} exec={r[ix['Instructions Executed']]:>9s} {r[ix['Source']].strip()[:70]:70s} {st}")
