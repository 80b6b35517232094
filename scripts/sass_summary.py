#!/usr/bin/env python
"""Opcode histogram, resources and the first global-memory instructions of the step kernels, from the built library
(cuobjdump -sass / -res-usage; runs without a GPU).  Writes profiles/r02_sass_step_kernels.txt."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "badchimp-cpp_b200", "libchimp_b200.so")
KERNELS = {
    "default single-GPU step kernel  collideStreamKernel<D3Q19, BGK, ONEPHASE=0, MOM=0, IDX_COMPACT, PEER=0>": "_ZN5chimp19collideStreamKernelINS_5D3Q19ELi0ELi0ELb0ELi1ELb0EEEvNS_8StepArgsE",
    "step kernel, skip-mask form of the compact index (untimed variant; bench.py tries it against the default before timing)  collideStreamKernel<D3Q19, BGK, 0, 0, IDX_COMPACT_MASK, 0>": "_ZN5chimp19collideStreamKernelINS_5D3Q19ELi0ELi0ELb0ELi2ELb0EEEvNS_8StepArgsE",
    "N-GPU step kernel (peer exchange fused)  collideStreamKernel<D3Q19, BGK, 0, 0, IDX_COMPACT, PEER=1>": "_ZN5chimp19collideStreamKernelINS_5D3Q19ELi0ELi0ELb0ELi1ELb1EEEvNS_8StepArgsE",
    "one_phase step kernel, packed attribute word (untimed variant)  collideStreamKernel<D3Q19, TRT, OP_PACKED, 0, IDX_COMPACT, 0>": "_ZN5chimp19collideStreamKernelINS_5D3Q19ELi1ELi2ELb0ELi1ELb0EEEvNS_8StepArgsE",
    "two-phase collide pass, derived phi index (untimed variant)  twoPhaseCollideKernel<D3Q19, MOM=0, IDX_COMPACT, DERIVED=1>": "_ZN5chimp21twoPhaseCollideKernelINS_5D3Q19ELb0ELi1ELb1EEEvNS_12TwoPhaseArgsE",
    "two-phase collide pass, full phi table  twoPhaseCollideKernel<D3Q19, MOM=0, IDX_COMPACT, DERIVED=0>": "_ZN5chimp21twoPhaseCollideKernelINS_5D3Q19ELb0ELi1ELb0EEEvNS_12TwoPhaseArgsE",
}
HEADER = """# r02: SASS of the step kernels (cuobjdump -sass / -res-usage of badchimp-cpp_b200/libchimp_b200.so, sm_100a, nvcc 12.9,
# -O3 -fmad=false -lineinfo; scripts/sass_summary.py).  Opcode histogram, resources, and the instructions that touch global memory.
# What it shows: no spills (no STL/LDL, STACK:0); no contracted multiply-add -- the 31 DFMA are the Newton steps of the
# three IEEE double divisions (MUFU.RCP64H + DFMA sequences, correctly rounded like the reference's divsd), every
# a*b+c of the collision is a separate DMUL and DADD (the reference binary has no FMA: bit parity);
# population loads are 64-bit LDG.E.64.CONSTANT (read-only path; consecutive lanes read consecutive doubles, so a
# warp load is one 256-byte request), index words are 32-bit coalesced LDG and the per-tile bases 128-bit broadcast
# loads (LDG.E.128.CONSTANT), stores are 64-bit STG.E.64 to consecutive slots; zero tensor-pipe instructions.
# The PEER instantiation adds, for the halo-coupled blocks only: counter polls (LDG.E.64.STRONG.SYS), system-scope fences
# (MEMBAR.SC.SYS + CCTL.IVALL), the remote copies of outgoing populations (STG.E.64.STRONG.SYS) and the block counter."""


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout.splitlines()
    out = [HEADER]
    for title, fun in KERNELS.items():
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", fun, LIB], capture_output=True, text=True).stdout
        ins = [l for l in sass.splitlines() if re.search(r"/\*[0-9a-f]{4}\*/", l)]
        ops = collections.Counter()
        for l in ins:
            m = re.search(r"/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", l)
            if m:
                ops[m.group(1)] += 1
        usage = next((res[i + 1].strip() for i, l in enumerate(res) if fun in l), "?")
        ordered = sorted(ops.items(), key=lambda kv: -kv[1])
        pick = lambda pre: ", ".join("%s x%d" % kv for kv in ordered if kv[0].startswith(pre))
        out += ["", "## " + title, "resources: " + usage, "instructions: %d" % sum(ops.values()),
                "memory opcodes: " + pick(("LDG", "STG", "LDL", "STL", "LDS", "STS", "LDC", "ATOM", "RED", "MEMBAR", "CCTL")),
                "fp64 opcodes:   " + pick(("DADD", "DMUL", "DFMA", "DSETP", "MUFU", "DMNMX")),
                "tensor / TMA opcodes: " + (pick(("HMMA", "IMMA", "DMMA", "UTC", "TCGEN", "UTMA")) or "none (memory-bound path, no contraction)"),
                "local-memory (spill) opcodes: " + (pick(("LDL", "STL")) or "none"),
                "top opcodes: " + ", ".join("%s x%d" % kv for kv in ordered[:14]), "first global loads / stores as emitted:"]
        shown = 0
        for l in ins:
            if ("LDG" in l or "STG" in l) and shown < 12:
                out.append("   " + re.sub(r"\s+", " ", l.strip())[:150])
                shown += 1
    with open(os.path.join(ROOT, "profiles", "r02_sass_step_kernels.txt"), "w") as fh:
        fh.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
