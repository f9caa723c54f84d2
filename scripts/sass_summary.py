"""Per-kernel SASS opcode summary of the shipped library (cuobjdump -sass splatter_a_video_b200/libspv_b200.so):
instruction count, the mnemonics that prove the Blackwell / Hopper+ data paths (UBLKCP = cp.async.bulk, SYNCS = mbarrier,
multimem.* = LDGMC / STGMC ..., REDG / RED = global reductions, SHFL, MUFU) and the ten most frequent opcodes.
    python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "splatter_a_video_b200", "libspv_b200.so")
MARK = ["UBLKCP", "UTMALDG", "SYNCS", "LDGMC", "STGMC", "REDG", "RED", "ATOMG", "ATOMS", "SHFL", "MUFU", "HMMA", "UTCHMMA", "LDGSTS", "BAR", "LDS", "STS", "FFMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip() or n
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*(?:\.[A-Z0-9_]+)*)", line)
        if m and cur is not None:
            cur[m.group(1)] += 1
    print(f"# {os.path.relpath(LIB, ROOT)}: {len(kernels)} kernels, sm_100a SASS (cuobjdump -sass), opcode = mnemonic with its modifiers")
    arch = subprocess.run(["cuobjdump", "-lelf", LIB], capture_output=True, text=True).stdout
    print("# ELF images:", ", ".join(sorted(set(re.findall(r"sm_\d+a?", arch)))))
    total = collections.Counter()
    for name, ops in kernels.items():
        base = collections.Counter()
        for op, n in ops.items():
            base[op.split(".")[0]] += n
            total[op.split(".")[0]] += n
        short = re.sub(r"\(anonymous namespace\)::", "", demangle(name))
        short = re.sub(r"\(.*", "", short)
        marks = "  ".join(f"{k}={base[k]}" for k in MARK if base.get(k))
        special = sorted({op for op in ops if any(op.startswith(p) for p in ("UBLKCP", "SYNCS", "LDGMC", "STGMC", "UTMA", "REDG", "RED."))})
        print(f"\n{short}\n  instructions {sum(ops.values())}   {marks}")
        if special:
            print("  data-path opcodes: " + ", ".join(f"{op} x{ops[op]}" for op in special))
        print("  top: " + ", ".join(f"{op} {n}" for op, n in base.most_common(10)))
    print("\n# library totals: " + ", ".join(f"{k}={total[k]}" for k in MARK if total.get(k)))


if __name__ == "__main__":
    sys.exit(main())
