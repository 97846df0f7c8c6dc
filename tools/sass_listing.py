#!/usr/bin/env python
"""profiles/<tag>_sass_{fast,contact}.txt: SASS of the two step kernels of marl_soccer_b200/libmsoc.so
(cuobjdump -sass), each preceded by its mnemonic histogram and the lines that prove the Blackwell-specific paths
(UBLKCP = cp.async.bulk, FENCE.VIEW.ASYNC, UTMACMDFLUSH).  Encodings are stripped to keep the files readable."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
lib = os.path.join(ROOT, "marl_soccer_b200", "libmsoc.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = re.split(r"(?=\t*Function : )", sass)
for name in ("fast", "contact"):
    blk = next(b for b in blocks if f"msoc_step_{name}_kernel" in b.split("\n", 1)[0])
    lines = []
    for ln in blk.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);\s*/\*", ln)
        if m:
            lines.append((m.group(1), m.group(2).strip()))
    hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for _, t in lines)
    special = [f"  {a}: {t}" for a, t in lines if re.search(r"UBLKCP|FENCE\.VIEW\.ASYNC|UTMACMDFLUSH|MUFU\.RSQ|LDG\.E\.128|STG\.E\.128|REDUX|CCTL", t)]
    out = os.path.join(ROOT, "profiles", f"{tag}_sass_{name}.txt")
    with open(out, "w") as f:
        f.write(f"msoc_step_{name}_kernel, sm_100a, {len(lines)} SASS instructions (cuobjdump -sass marl_soccer_b200/libmsoc.so; tools/sass_listing.py)\n\n")
        f.write("mnemonic histogram:\n")
        for k, v in hist.most_common():
            f.write(f"  {k:12s} {v}\n")
        f.write("\nBlackwell-specific / notable instructions (address: instruction):\n")
        seen = collections.Counter()
        for s in special:
            key = s.split(":", 1)[1].split()[0]
            seen[key] += 1
            if seen[key] <= 12:
                f.write(s + "\n")
        f.write("  totals: " + ", ".join(f"{k} x{v}" for k, v in seen.items()) + "\n")
        f.write("\nfull listing:\n")
        for a, t in lines:
            f.write(f"/*{a}*/ {t}\n")
    print(out, len(lines), "instructions", os.path.getsize(out) // 1024, "KiB")
