"""SASS evidence of the Blackwell-native instructions per kernel of libgeossl_b200.so.

    python profiles/sass_summary.py > profiles/rNN_sass_summary.txt

Counts, per kernel function, the mnemonics that prove tcgen05 / TMEM / TMA / async-copy use (B200_PROFILING.md):
UTCHMMA (tcgen05.mma kind::f16), LDTM / STTM (tcgen05.ld / .st), UTCBAR (tcgen05.commit), UBLKCP (cp.async.bulk), UBLKPF (cp.async.bulk.prefetch.L2),
LDGSTS (cp.async), SYNCS (mbarrier), MUFU, and prints the first UTCHMMA / LDTM / UBLKCP line of each kernel as an excerpt."""
import collections
import os
import re
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "geossl_b200", "libgeossl_b200.so")
MNEMONICS = ("UTCHMMA", "LDTM", "STTM", "UTCBAR", "UBLKCP", "UBLKPF", "LDGSTS", "SYNCS", "MUFU", "ATOMG", "REDG", "RED.")

out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kernels, cur = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(re.sub(r"\(.*", "", demangle(m.group(1))), {"n": 0, "c": collections.Counter(), "ex": {}})
        continue
    if cur is None or "/*" not in line:
        continue
    m = re.search(r"/\*[0-9a-f]+\*/\s+(.*?);", line)
    if not m:
        continue
    ins = m.group(1)
    cur["n"] += 1
    for mn in MNEMONICS:
        if re.search(r"(^|\s)" + re.escape(mn), ins):
            cur["c"][mn.rstrip(".")] += 1
            cur["ex"].setdefault(mn.rstrip("."), ins.strip())
print(f"# cuobjdump -sass {os.path.relpath(LIB, REPO)} (sm_100a): instruction counts per kernel")
tot = collections.Counter()
for name, k in kernels.items():
    tot.update(k["c"])
    if not k["c"]:
        continue
    print(f"{name}\n    {k['n']} instructions; " + ", ".join(f"{m} {c}" for m, c in k["c"].items()))
    for m in ("UTCHMMA", "LDTM", "STTM", "UBLKCP", "UBLKPF", "LDGSTS"):
        if m in k["ex"]:
            print(f"        {k['ex'][m]}")
print("# library totals: " + ", ".join(f"{m} {c}" for m, c in tot.items()))
