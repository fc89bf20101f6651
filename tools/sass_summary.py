"""SASS evidence for profiles/: per hot kernel of libtc_b200.so the static instruction count and the mnemonic
histogram (cuobjdump -sass), plus registers / shared memory / spills from cuobjdump -res-usage.  Runs without a GPU.

    python tools/sass_summary.py > profiles/r2_sass_kernels.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "text_compression_b200", "libtc_b200.so")
HOT = ["uk_keys_kernel<0>", "uk_keys_raw_kernel", "seg_hist_kernel", "part_kernel<true>", "part_kernel<false>",
       "final_sort_kernel", "bwt_emit_kernel", "mtf3_tile_last_kernel<(anonymous namespace)::SrcU8>", "mtf3_tile_scan_kernel",
       "mtf3_starts_kernel", "mtf3_replay_kernel<(anonymous namespace)::SrcU8, 160, true>",
       "mtfa_summary_kernel<(anonymous namespace)::SrcU8>", "mtfa_replay_kernel<(anonymous namespace)::SrcU8, true>",
       "rle_emit_tiled_kernel<(anonymous namespace)::In16<true>, true>", "rs_scatter_kernel", "rs_hist_kernel",
       "fm_count_kernel", "fm_locate_kernel", "mtfd_perm_kernel", "mtfd_replay_kernel", "inv_walk1_kernel", "inv_walk2_kernel"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    usage = {}
    cur = None
    for line in res.split("\n"):
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        if cur and "REG:" in line:
            usage[cur] = line.strip()
            cur = None
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.split("\n"):
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append(m.group(1))
    dm = demangle(list(funcs))
    tensor = re.compile(r"\b(UTCMMA|UTCHMMA|UTCQMMA|HMMA|IMMA|DMMA|UTMALDG|UTMASTG|UBLKCP|LDGSTS)\b")
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(funcs)} kernels; static SASS of the hot ones (sm_100a)")
    for want in HOT:
        hit = [f for f in funcs if want in dm[f]]
        if not hit:
            print(f"\n## {want}: not in the library", file=sys.stderr)
            continue
        f = hit[0]
        ins = funcs[f]
        hist = collections.Counter()
        for i in ins:
            t = i.split()
            op = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
            hist[op.split(".")[0]] += 1
        print(f"\n## {dm[f][:150]}")
        print(f"   {len(ins)} instructions; {usage.get(f, '')}")
        print("   " + "  ".join(f"{k}:{v}" for k, v in hist.most_common(24)))
        t = sorted({m.group(1) for i in ins for m in [tensor.search(i)] if m})
        print(f"   tensor-core / TMA / async-copy mnemonics: {t if t else 'none (integer, logic, shuffle, vote, LDG/STG/LDS/STS)'}")


if __name__ == "__main__":
    main()
