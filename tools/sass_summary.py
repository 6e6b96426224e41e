"""Per-kernel counts of the Blackwell-specific SASS mnemonics and the resource usage of every kernel of the built library
(no GPU needed):
    python tools/sass_summary.py sass > profiles/r2_sass_mnemonics.txt
    python tools/sass_summary.py res  > profiles/r2_ptxas_resources.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "robust-audio-deepfake-evolution_b200", "csrc")
LIB = os.path.join(ROOT, "robust-audio-deepfake-evolution_b200", "libbimamba_sm100.so")
MN = ("UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "LDGSTS", "FFMA2", "FMUL2", "FADD2", "MUFU.EX2", "ACQBULK")


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), text=True, capture_output=True).stdout.splitlines()
    return dict(zip(names, out))


def sass():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], text=True, capture_output=True).stdout
    counts, order, cur = {}, [], None
    for ln in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur and "/*" in ln:
            for k in MN:
                if re.search(r"\b" + re.escape(k) + r"\b", ln):
                    counts[cur][k] += 1
    dm = demangle(order)
    print("# cuobjdump -sass libbimamba_sm100.so : occurrences of the Blackwell-specific mnemonics per kernel (last build of round 2)")
    print("# UTCHMMA = tcgen05.mma, UTMALDG = TMA tensor load (cp.async.bulk.tensor), LDTM = tcgen05.ld (TMEM -> registers), "
          "UTCBAR = tcgen05.commit, LDGSTS = cp.async, FFMA2/FMUL2/FADD2 = packed fp32, ACQBULK = griddepcontrol.wait (PDL)")
    for f in order:
        c = counts[f]
        if not c:
            continue
        print(f"{dm[f][:110]:112s}" + " ".join(f"{k}={c[k]}" for k in MN if c[k]))


def res():
    print("# cuobjdump -res-usage of every object of the library: registers / stack (spill) bytes / shared memory per kernel (last build of round 2)")
    print("# source | kernel | usage")
    for o in sorted(os.listdir(CSRC)):
        if not o.endswith(".o"):
            continue
        txt = subprocess.run(["cuobjdump", "-res-usage", os.path.join(CSRC, o)], text=True, capture_output=True).stdout
        fn = None
        rows = []
        for ln in txt.splitlines():
            m = re.match(r"\s*Function (\S+):", ln)
            if m:
                fn = m.group(1)
                continue
            if fn and "REG:" in ln:
                rows.append((fn, " ".join(ln.split())))
                fn = None
        dm = demangle([r[0] for r in rows]) if rows else {}
        for f, u in rows:
            print(f"{o[:-2]}.cu | {dm[f]} | {u}")


if __name__ == "__main__":
    (sass if sys.argv[1:] == ["sass"] else res)()
