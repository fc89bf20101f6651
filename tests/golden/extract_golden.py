#!/usr/bin/env python3
"""Extracts the reference's own known-answer vectors into reference_vectors.json.

Run in the build container (the reference tree does not exist on the GPU box):
    python tests/golden/extract_golden.py [/root/reference]

Sources (paths relative to the reference root):
  * src/Data/RLE.hs:279-311   rle1, s1, rle2, s2 and the four tests at :313-320
  * src/Data/MTF.hs:287-299   the two MTF tests
  * src/Data/FMIndex/Internal.hs:49-113  the worked "abracadabra" example
    (BWT string, C[c] table, Occ(c,k) table) from the module documentation.
The Haskell literals are parsed textually; nothing is computed here.
"""
import json
import os
import re
import sys

ref = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"
here = os.path.dirname(os.path.abspath(__file__))


def read(p):
    with open(os.path.join(ref, p), encoding="utf-8") as f:
        return f.read()


def maybe_list(src):
    """Parse `Just "x"` / `Nothing` items in order."""
    out = []
    for m in re.finditer(r'Just\s+"((?:[^"\\]|\\.)*)"|Nothing', src):
        out.append(None if m.group(0) == "Nothing" else m.group(1))
    return out


rle_hs = read("src/Data/RLE.hs")
mtf_hs = read("src/Data/MTF.hs")
fmi_hs = read("src/Data/FMIndex/Internal.hs")

rle1 = maybe_list(re.search(r"rle1 = RLE \(fromList \[(.*?)\]\)", rle_hs, re.S).group(1))
rle2 = maybe_list(re.search(r"rle2 = RLE \(fromList \[(.*?)\]\)", rle_hs, re.S).group(1))
s1 = re.search(r'^s1 = "(.*)"$', rle_hs, re.M).group(1)
s2 = re.search(r'^s2 = "(.*)"$', rle_hs, re.M).group(1)
assert "textToBWTToRLET s1" in rle_hs and "textToBWTToRLEB s2" in rle_hs
assert "textFromBWTFromRLET rle1" in rle_hs and "textFromBWTFromRLET rle2" in rle_hs

m = re.search(r"\(MTF \(\[([0-9,]+)\],\s*\[(.*?)\]\)\)\s*\(textToBWTToMTFB \"(\w+)\"\)", mtf_hs, re.S)
mtf_idx = [int(x) for x in m.group(1).split(",")]
mtf_list = maybe_list(m.group(2))
mtf_text = m.group(3)
assert re.search(r'"%s"\s*\(textFromBWTFromMTFB' % mtf_text, mtf_hs, re.S)

# abracadabra documentation tables
fm_text = re.search(r'Given the following input, "(\w+)"', fmi_hs).group(1)
fm_bwt = re.search(r'C\[c\] of "([^"]+)"', fmi_hs).group(1)


def table_rows(block):
    rows = []
    for line in block.splitlines():
        if "|" in line:
            rows.append([c.strip() for c in line.strip().lstrip("-").strip().strip("|").split("|")])
    return rows


cblock = fmi_hs.split('C[c] of "%s"' % fm_bwt)[1].split("-- and")[0]
crow = table_rows(cblock)
c_syms, c_vals = crow[0][1:], [int(x) for x in crow[1][1:]]
oblock = fmi_hs.split('Occ(c,k) of "%s"' % fm_bwt)[1].split("Keep in mind")[0]
orow = table_rows(oblock)
occ_cols = orow[0][1:]
occ = {r[0]: [int(x) for x in r[1:]] for r in orow[2:] if r and r[0]}

out = {
    "_source": "Matthew-Mosior/text-compression v0.1.0.25; see extract_golden.py for file:line",
    "rle": [
        {"fn": "textToBWTToRLET", "text": s1, "rle": rle1, "src": "src/Data/RLE.hs:279-288,316,318"},
        {"fn": "textToBWTToRLEB", "text": s2, "rle": rle2, "src": "src/Data/RLE.hs:290-311,317,319"},
    ],
    "mtf": [
        {"fn": "textToBWTToMTFB", "text": mtf_text, "indices": mtf_idx, "final_list": mtf_list,
         "src": "src/Data/MTF.hs:290-298"},
    ],
    "fmindex_doc": {
        "text": fm_text, "bwt": fm_bwt, "C_syms": c_syms, "C_vals": c_vals,
        "occ_cols": occ_cols, "occ": occ, "src": "src/Data/FMIndex/Internal.hs:49-113",
    },
}
with open(os.path.join(here, "reference_vectors.json"), "w") as f:
    json.dump(out, f, indent=1)
print("rle1", len(rle1), "rle2", len(rle2), "s2", len(s2), "mtf", mtf_idx, mtf_list, "fm", fm_text, fm_bwt, c_syms, c_vals,
      {k: len(v) for k, v in occ.items()})
