"""Extracts the export lists of the reference's eight modules (names only) into reference_exports.json:
the contract the Haskell drop-in modules under haskell/src must match name for name.
Run in the build container (reads /root/reference); tests read only the committed JSON."""
import json
import os
import re
import sys

REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference/src"
MODS = ["Data/BWT.hs", "Data/BWT/Internal.hs", "Data/MTF.hs", "Data/MTF/Internal.hs", "Data/RLE.hs",
        "Data/RLE/Internal.hs", "Data/FMIndex.hs", "Data/FMIndex/Internal.hs"]


def exports(path):
    src = open(path).read()
    src = re.sub(r"\{-.*?-\}", "", src, flags=re.S)
    m = re.search(r"^module\s+([\w.]+)\s*\((.*?)\)\s*where", src, flags=re.S | re.M)
    body = re.sub(r"--[^\n]*", "", m.group(2))
    names, depth, cur = [], 0, ""
    for ch in body:            # split on top-level commas: Pack(pck, unpck, Itm, one) is one export
        if ch == "(":
            depth += 1
        if ch == ")":
            depth -= 1
        if ch == "," and depth == 0:
            names.append(cur)
            cur = ""
        else:
            cur += ch
    names.append(cur)
    return m.group(1), sorted(re.sub(r"\s+", "", n) for n in names if n.strip())


if __name__ == "__main__":
    out = dict(exports(os.path.join(REF, m)) for m in MODS)
    json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_exports.json"), "w"),
              indent=1, sort_keys=True)
    print({k: len(v) for k, v in out.items()})
