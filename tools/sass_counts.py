"""Per-kernel SASS opcode counts of the shipped library (cuobjdump -sass): the mnemonics that prove which hardware paths the
kernels use (B200_PROFILING.md): UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR = tcgen05.commit, LDGSTS = cp.async,
UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier ops, REDG = red.global, LDGMC = multimem.ld_reduce, USETMAXREG = setmaxnreg.

    python tools/sass_counts.py > profiles/r2_sass_counts.md
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "fruitnerf_b200" / "csrc" / "libfruitnerf_b200.so"
KEYS = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "LDGSTS", "UBLKCP", "UTMALDG", "SYNCS", "REDG", "ATOMG", "LDGMC", "USETMAXREG", "NANOSLEEP", "ELECT",
        "SHFL", "MUFU", "FFMA"]


def demangle(name: str) -> str:
    out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    out = out.replace("(anonymous namespace)::", "")
    depth, cut = 0, len(out)
    for i, ch in enumerate(out):  # cut the parameter list, keep template arguments
        if ch == "<":
            depth += 1
        elif ch == ">":
            depth -= 1
        elif ch == "(" and depth == 0:
            cut = i
            break
    return out[:cut].replace("void ", "").replace("fnr::", "")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    pat = re.compile(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Za-z0-9_.]+)")
    counts, fn = collections.defaultdict(collections.Counter), None
    variants = collections.Counter()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = pat.match(line)
        if m and fn:
            op = m.group(1)
            counts[fn][op.split(".")[0]] += 1
            if op.startswith(("LDGMC", "REDG", "UBLKCP", "LDGSTS", "UTCHMMA")):
                variants[op] += 1
    res, cur = {}, None
    for line in subprocess.run(["cuobjdump", "-res-usage", str(LIB)], capture_output=True, text=True, check=True).stdout.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            cur = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+)", line)
        if m and cur:
            res[cur] = tuple(int(v) for v in m.groups())
    rows = sorted(((demangle(f), sum(c.values()), c, res.get(f, (0, 0, 0))) for f, c in counts.items()), key=lambda r: -r[2]["UTCHMMA"] * 10**6 - r[1])
    print(f"# SASS opcode counts of `{LIB.relative_to(ROOT)}` (`python tools/sass_counts.py`; sm_100a, static instruction counts)\n")
    print("| kernel | regs | stack B | static smem B | instr | " + " | ".join(KEYS) + " |")
    print("|---|---|---|---|---|" + "---|" * len(KEYS))
    for name, tot, c, (reg, stack, shared) in rows:
        print(f"| `{name}` | {reg} | {stack} | {shared} | {tot} | " + " | ".join(str(c.get(k, 0)) for k in KEYS) + " |")
    print("\n(regs / stack / static shared memory: `cuobjdump -res-usage`; the tensor-core kernels take their dynamic shared memory at launch and "
          "re-balance registers per role with `setmaxnreg`; stack bytes on the simt kernels are per-thread activation arrays, on the tensor kernels at most 128 B (a dynamically indexed register array, not spill traffic in the loops).)")
    total = collections.Counter()
    for c in counts.values():
        total.update(c)
    print("\nLibrary totals: " + ", ".join(f"{k} {total[k]}" for k in KEYS if total[k]))
    print("\nVariants: " + ", ".join(f"`{k}` {v}" for k, v in sorted(variants.items())))
    absent = [k for k in ("HMMA", "IMMA", "UTMALDG", "WGMMA") if not total[k]]
    print("\nAbsent: " + ", ".join(absent) + " (no mma.sync / wgmma tensor-core fallback; bulk copies are the non-tensor-map `cp.async.bulk` form = `UBLKCP`).")


if __name__ == "__main__":
    sys.exit(main())
