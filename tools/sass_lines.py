#!/usr/bin/env python
"""Static view of one kernel of csrc/librthx.so, no GPU needed: SASS instructions and local-memory (spill) instructions per
source line, from `nvdisasm -g` on the embedded sm_100a cubin (the library is built with -lineinfo).

    python tools/sass_lines.py 'trace_exchange_queue_kernel<3, 2, false, true, false>'          # spills per line + hottest lines
    python tools/sass_lines.py 'trace_exchange_sq_kernel<4, false>' --dump sq.sass               # also write the kernel's SASS

This is how the parameter-bank alignment regression of round 2 was found ("A code-generation trap"): the spill
instructions of the MULTI_BOUNCE queue kernel had gone from 127 to 377 without a change to its source.
"""
import argparse
import collections
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def kernel_sections(so, workdir):
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=workdir, check=True, capture_output=True)
    cubin = max((os.path.join(workdir, f) for f in os.listdir(workdir) if f.endswith(".cubin")), key=os.path.getsize)
    text = subprocess.run(["nvdisasm", "-g", cubin], check=True, capture_output=True, text=True, errors="replace").stdout
    sections, name, cur = {}, None, []
    for line in text.splitlines():
        if line.startswith(".text."):
            name, cur = line.strip().rstrip(":")[len(".text."):], []
            sections[name] = cur
        elif line.startswith("//---------------------"):
            name = None
        elif name is not None:
            cur.append(line)
    return sections


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("kernel", help="demangled name (prefix) without the rthx:: namespace, e.g. 'trace_exchange_sq_kernel<4, false>'")
    ap.add_argument("--so", default=os.path.join(ROOT, "raytraceheattransfer.jl_b200", "csrc", "librthx.so"))
    ap.add_argument("--dump", help="write the kernel's annotated SASS to this file")
    ap.add_argument("--top", type=int, default=15, help="source lines with the most instructions to list")
    args = ap.parse_args()
    with tempfile.TemporaryDirectory() as d:
        sections = kernel_sections(args.so, d)
    mangled = list(sections)
    dem = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True, check=True).stdout.splitlines()
    hits = [m for m, n in zip(mangled, dem) if n.replace("void ", "").replace("rthx::", "").startswith(args.kernel)]
    if len(hits) != 1:
        sys.exit(f"{len(hits)} kernels match {args.kernel!r}: " + ", ".join(n for n in dem if "trace_exchange" in n))
    body = sections[hits[0]]
    if args.dump:
        open(args.dump, "w").write("\n".join(body) + "\n")
    line_no, total, spills = None, collections.Counter(), collections.Counter()
    for l in body:
        m = re.search(r'//## File "[^"]+", line (\d+)', l)
        if m:
            line_no = int(m.group(1))
        elif re.search(r"/\*[0-9a-f]{4,}\*/", l):
            total[line_no] += 1
            if re.search(r"\b(STL|LDL)\b", l):
                spills[line_no] += 1
    print(f"{args.kernel}: {sum(total.values())} SASS instructions, {sum(spills.values())} local-memory (spill) instructions")
    print("spill instructions by source line of rthx_kernels.cu:", dict(sorted(spills.items())))
    print("source lines with the most instructions:", total.most_common(args.top))


if __name__ == "__main__":
    main()
