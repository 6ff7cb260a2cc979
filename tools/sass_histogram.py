"""SASS opcode histogram per CUDA source (cuobjdump -sass on the in-tree objects) -> profiles/rNN_sass_opcodes.md.
Shows which kernels really use the Blackwell paths: UTCHMMA (tcgen05.mma), LDTM / STTM (tcgen05.ld / st), UTMALDG /
UTMASTG (TMA load / store), UTCBAR, versus legacy HMMA (mma.sync) and LDGSTS (cp.async).

    python tools/sass_histogram.py r02
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "marie-icr_b200", "build")
KEY = ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTCBAR", "SYNCS", "HMMA", "IMMA", "IDP", "LDGSTS",
       "MUFU", "REDUX", "ATOMG", "ATOMS", "RED", "BAR")


def histogram(obj):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    per_fn, fn = collections.OrderedDict(), None
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            per_fn[fn] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_]+)*)", line)
        if m and fn:
            per_fn[fn][m.group(1)] += 1
            if m.group(1) in ("UTCHMMA", "UTMALDG", "UTMASTG", "HMMA") and m.group(2):
                per_fn[fn][m.group(1) + m.group(2)] += 1
    return per_fn


def demangle(names):
    out = subprocess.run(["cu++filt"] + list(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out)) if len(out) == len(names) else {n: n for n in names}


def main(tag):
    lines = [f"# SASS opcode histogram per source file ({tag}; `cuobjdump -sass marie-icr_b200/build/*.o`, sm_100a)", ""]
    for obj in sorted(os.listdir(BUILD)):
        if not obj.endswith(".o"):
            continue
        per_fn = histogram(os.path.join(BUILD, obj))
        if not per_fn:
            continue
        total = collections.Counter()
        for c in per_fn.values():
            total.update({k: v for k, v in c.items() if "." not in k})
        n_inst = sum(total.values())
        lines.append(f"## {obj.replace('.o', '.cu')} — {len(per_fn)} kernels, {n_inst} instructions")
        lines.append("")
        lines.append("key opcodes: " + ", ".join(f"{k} {total[k]}" for k in KEY if total.get(k)))
        lines.append("")
        lines.append("top 12: " + ", ".join(f"{k} {v}" for k, v in total.most_common(12)))
        lines.append("")
        names = demangle(list(per_fn))
        lines.append("| kernel | instr | " + " | ".join(KEY[:12]) + " |")
        lines.append("|---|---|" + "---|" * 12)
        for fn, c in per_fn.items():
            n = sum(v for k, v in c.items() if "." not in k)
            short = re.sub(r"\(.*", "", re.sub(r"\((?:bool|int|unsigned int|unsigned)\)", "", names[fn])).replace("(anonymous namespace)::", "").replace("void ", "")
            if len(short) > 70:
                short = short[:67] + "..."
            lines.append(f"| `{short}` | {n} | " + " | ".join(str(c.get(k, 0) or "") for k in KEY[:12]) + " |")
        variants = collections.Counter()
        for c in per_fn.values():
            variants.update({k: v for k, v in c.items() if "." in k})
        if variants:
            lines.append("")
            lines.append("variants: " + ", ".join(f"{k} {v}" for k, v in sorted(variants.items())))
        lines.append("")
    path = os.path.join(ROOT, "profiles", f"{tag}_sass_opcodes.md")
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    print("wrote", path)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
