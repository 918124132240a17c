#!/usr/bin/env python
"""profiles/r2_sass_tma_excerpts.txt: static SASS evidence (UBLKCP = TMA bulk copies, SYNCS = mbarrier operations) for the persistent
kernels, from `cuobjdump -sass` of the built library (no GPU needed).  Usage: python tools/sass_excerpts.py"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["k_div3qILi8", "div3q_elementILi8", "k_div3pILi8", "k_axhelm3pILi8", "ax3p_elementILi8", "k_pcg_fused_pILi6", "k_gradt3ILi8ELi2",
        "k_advab2ILi8ELi0"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "nekstab_b200", "libnekstab_b200.so")], capture_output=True, text=True).stdout
    out = ["SASS evidence (cuobjdump -sass nekstab_b200/libnekstab_b200.so, sm_100a) for the TMA / mbarrier kernels of round 2.",
           "Mnemonics: UBLKCP = cp.async.bulk (TMA bulk copy global->shared); SYNCS.* = mbarrier operations (EXCH = init, ARRIVE.TRANS64 =",
           "arrive.expect_tx, PHASECHK.TRANS64.TRYWAIT = try_wait.parity); DFMA with a UR operand fed by LDCU.128 c[0x3][..] = operator matrices",
           "from the constant bank; LDGSTS = cp.async.  Static instruction counts per function; the element bodies of the persistent kernels are",
           "separate (__noinline__) functions.  Regenerate: python tools/sass_excerpts.py", ""]
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        if not any(w in name for w in WANT):
            continue
        ins = [l for l in f.split("\n") if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        cnt = lambda p: sum(1 for l in ins if re.search(p, l))
        out.append(f"== {name}")
        out.append(f"   instructions {len(ins)}; UBLKCP {cnt('UBLKCP')}; SYNCS {cnt('SYNCS')}; DFMA {cnt('DFMA')} (with uniform-register operand: "
                   f"{cnt(r'DFMA[^;]*UR')}); LDCU {cnt('LDCU')}; LDS {cnt(r'LDS')}; STS {cnt(r'STS')}; LDG {cnt(r'LDG')}; STG {cnt(r'STG')}; "
                   f"BAR.SYNC {cnt('BAR.SYNC')}; LDGSTS {cnt('LDGSTS')}")
        shown = 0
        for l in ins:
            if re.search("UBLKCP|SYNCS", l) and shown < 8:
                out.append("   " + re.sub(r"\s+", " ", l.strip())[:150])
                shown += 1
        out.append("")
    with open(os.path.join(ROOT, "profiles", "r2_sass_tma_excerpts.txt"), "w") as fh:
        fh.write("\n".join(out))


if __name__ == "__main__":
    main()
