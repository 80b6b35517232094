"""DEVELOPMENT TOOL (see include/cuda_runtime.h).  Rewrites the CUDA sources of the engine so that g++ accepts them:
kernel<<<grid, block, smem, stream>>>(args)  ->  emu::bind(kernel, emu::Cfg(grid, block, smem, stream))(args)
and the one PTX statement (%globaltimer) becomes a host clock.  usage: prep.py SRC_DIR DST_DIR"""
import os
import re
import sys

LAUNCH = re.compile(r"([A-Za-z_]\w*(?:<[^<>;(){}]*>)?)\s*<<<(.*?)>>>\s*\(", re.S)
TIMER = re.compile(r'asm volatile\("mov\.u64 %0, %%globaltimer;"\s*:\s*"=l"\((\w+)\)\);')


def convert(text):
    text, n = LAUNCH.subn(lambda m: "emu::bind(%s, emu::Cfg(%s))(" % (m.group(1), m.group(2)), text)
    text, t = TIMER.subn(lambda m: "%s = emu::nowNs();" % m.group(1), text)
    if "<<<" in text or "asm volatile" in text:
        raise SystemExit("prep.py: a launch or asm statement was not recognised")
    return text, n


def main(src, dst):
    os.makedirs(dst, exist_ok=True)
    total = 0
    for name in sorted(os.listdir(src)):
        if not name.endswith((".cu", ".cuh")):
            continue
        text, n = convert(open(os.path.join(src, name)).read())
        total += n
        out = name[:-3] + ".cpp" if name.endswith(".cu") else name
        open(os.path.join(dst, out), "w").write(text)
    print("prep.py: %d launches rewritten" % total)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
