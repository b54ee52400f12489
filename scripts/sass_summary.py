"""SASS evidence for profiles/: per hot kernel of libf2d.so, the opcode histogram of its
sm_100a code (cuobjdump -sass), with the Blackwell-specific mnemonics called out:
UTMALDG (cp.async.bulk.tensor = TMA loads), SYNCS (mbarrier), LDGSTS (cp.async),
SHFL, DFMA / DADD / DMUL (fp64 pipe), F2F (fp32 <-> fp64), BAR, LDS / STS.

    python scripts/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fluids2d_b200", "libf2d.so")
# the instantiations the 4096^2 Euler benchmark launches (+ the round-1 per-point stage kernel for comparison,
# the rsw diagnostic kernel and the scalar transport kernel of the rsw / boussinesq configurations)
HOT = ["k_stage_tma<0, 0, 1>", "k_stage_tma<0, 0, 2>", "k_stage_tma<0, 0, 3>", "k_diag_tma<0, true>", "k_diag_tma<0, false>", "k_transport_tma<0, 2>", "k_rhs_mom<0, 0, 2>",
       "k_mg_down<float, float, double, float, true, true, 2, 64", "k_mg_up<float, float, double, float, true, true, 2, 64",
       "k_mg_down<float, float, float, float, false, true, 2, 64", "k_mg_up<float, float, float, float, false, false, 2, 64",
       "k_cg_dir_apply<float>", "k_cg_update_p", "k_cg_resid_guess", "k_mg_tail", "k_div_u", "k_p2p_exchange"]
KEYS = ["UTMALDG", "SYNCS", "LDGSTS", "LDG", "STG", "LDS", "STS", "SHFL", "BAR", "DFMA", "DADD", "DMUL", "FFMA", "FADD",
        "FMUL", "F2F", "MUFU", "IMAD", "LDL", "STL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            funcs[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            funcs[cur][m.group(1)] += 1
    dm = demangle(list(funcs))
    print(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}  (arch sm_100a), opcode counts per kernel instantiation")
    print(f"# {'kernel':100s} {'instr':>6s} " + " ".join(f"{k:>7s}" for k in KEYS))
    tot = collections.Counter()
    for f, c in funcs.items():
        name = dm.get(f, f)
        if not any(h in name for h in HOT):
            continue
        short = re.sub(r"f2d::", "", name)
        short = re.sub(r"^void ", "", re.sub(r"\(.*", "", short))[:100]
        n = sum(c.values())
        row = [sum(v for k, v in c.items() if k.startswith(key)) for key in KEYS]
        for key, v in zip(KEYS, row):
            tot[key] += v
        print(f"{short:102s} {n:6d} " + " ".join(f"{v:7d}" for v in row))
    print("# library totals over the kernels above: " + ", ".join(f"{k} {tot[k]}" for k in KEYS if tot[k]))


if __name__ == "__main__":
    main()
