"""Developer tool (CPU only): per-kernel SASS evidence for the shipped library.
   python tools/sass_evidence.py [profiles/sass_r2.md]
Runs `cuobjdump -sass` on cruse_b200/libcruse_sm100.so, splits the listing per kernel and counts the mnemonics that prove which
hardware paths a kernel uses (B200_PROFILING.md): UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (tensor memory), UTMALDG =
TMA tensor loads, UBLKCP = non-tensor bulk copies, STAS = st.async to a peer CTA's shared memory, SYNCS = mbarrier, UTCBAR =
tcgen05.commit, MUFU.* = special-function unit; plus one excerpt (the instructions around the first MMA) per tensor-core kernel."""
import collections
import os
import re
import subprocess
import sys

root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(root, "cruse_b200", "libcruse_sm100.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(root, "profiles", "sass_r2.md")
txt = subprocess.run(["cuobjdump", "-sass", lib], check=True, capture_output=True, text=True).stdout
names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", txt)), capture_output=True, text=True).stdout.split("\n")
parts = re.split(r"\n\s*Function : \S+\n", txt)[1:]
KEYS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UBLKCP", "STAS", "SYNCS", "MUFU", "LDG.E.128", "LDG.E.64", "STG.E.128", "STG.E.64", "HMMA", "FFMA"]


def short(n):
    n = re.sub(r"\(anonymous namespace\)::", "", n)
    n = re.sub(r"^void ", "", n)
    n = re.sub(r"\((int|bool|unsigned int)\)", "", n).replace("cruse::<unnamed>::", "").replace("<unnamed>::", "")
    return n.split("(")[0][:110]


rows, excerpts = [], []
for name, body in zip(names, parts):
    ins = re.findall(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", body)
    c = collections.Counter()
    for m in ins:
        if m.startswith(("LDG", "STG")):
            w = "128" if ".128" in m else "64" if ".64" in m else None
            if w:
                c[f"{m[:3]}.E.{w}"] += 1
            continue
        for k in KEYS:
            if m == k or m.startswith(k + "."):
                c[k] += 1
                break
    kinds = collections.Counter(re.findall(r"MUFU\.([A-Z0-9]+)", body))
    rows.append((short(name), len(ins), c, kinds))
    if c["UTCHMMA"]:
        lines = body.split("\n")
        i = next(k for k, l in enumerate(lines) if "UTCHMMA" in l)
        excerpts.append((short(name), [re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip() for l in lines[max(0, i - 3):i + 3] if "/*" in l]))

rows.sort(key=lambda r: (-r[2]["UTCHMMA"], -r[2]["UTMALDG"], r[0]))
with open(out, "w") as f:
    f.write(f"SASS evidence for `cruse_b200/libcruse_sm100.so` (sm_100a; `cuobjdump -sass`, {len(rows)} kernels, made by `tools/sass_evidence.py`)\n\n")
    f.write("UTCHMMA = tcgen05.mma (kind::tf32 and kind::f16 share the mnemonic; the operand kind is in the instruction descriptor), UTCBAR = tcgen05.commit, "
            "LDTM / STTM = tcgen05.ld / tcgen05.st, UTMALDG = TMA tensor load, UBLKCP = bulk copy, STAS = st.async into a peer CTA's shared memory, "
            "SYNCS = mbarrier operations, HMMA = warp-level mma.sync (the one-warp-per-frame fused decoder).  Kernels without any of these are the streaming (HBM-bound) kernels.\n\n")
    f.write("| kernel | instr | UTCHMMA | UTCBAR | LDTM | STTM | UTMALDG | UBLKCP | STAS | SYNCS | HMMA (mma.sync) | MUFU (kinds) | LDG.128 | STG.128 | FFMA |\n|---|---:|---:|---:|---:|---:|---:|---:|---:|---:|---:|---|---:|---:|---:|\n")
    for n, ni, c, kinds in rows:
        mu = " ".join(f"{k}:{v}" for k, v in sorted(kinds.items())) or "-"
        f.write(f"| `{n}` | {ni} | {c['UTCHMMA'] or ''} | {c['UTCBAR'] or ''} | {c['LDTM'] or ''} | {c['STTM'] or ''} | {c['UTMALDG'] or ''} | {c['UBLKCP'] or ''} | {c['STAS'] or ''} | "
                f"{c['SYNCS'] or ''} | {c['HMMA'] or ''} | {mu} | {c['LDG.E.128'] or ''} | {c['STG.E.128'] or ''} | {c['FFMA'] or ''} |\n")
    tot = collections.Counter()
    for _, _, c, _ in rows:
        tot.update(c)
    f.write("\nTotals: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]) + "\n\n## Excerpts: the instructions around the first MMA of each tensor-core kernel\n\n")
    for n, ex in excerpts:
        f.write(f"`{n}`\n```\n" + "\n".join(ex) + "\n```\n")
print(out, len(rows), "kernels")
