#!/usr/bin/env python
"""Offline look at the run-time compiled kernels of a model (no GPU needed).

Dumps the translation unit of one kernel group as the run-time compiler sees it, compiles it with
nvcc for sm_100a (-lineinfo, --fmad=true: the NVRTC options of csrc/jit.cpp) and prints, per kernel,
registers / spills / stack (ptxas -v) and the SASS opcode histogram with the fp64 instruction count
(DFMA + DADD + DMUL + DSETP ...).  The numbers behind profiles/ncu_traffic.json's
`fp64_instr_per_cell`: a straight-line kernel executes every instruction once per thread, i.e. once
per lattice cell for the one-thread-per-cell kernels; for the looping (marching / tile) kernels the
histogram is printed for the whole kernel and, with --loop, for the hottest loop body (the longest
backward-branch span).

  python tools/sass_count.py                       # bench model, tile group
  python tools/sass_count.py --group 5 --kernel dc_k_struct_apply_0
  python tools/sass_count.py --case cell3d --group 2
"""
from __future__ import annotations

import argparse
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def source_for(args):
    import dune_copasi_b200 as D
    if args.case:
        import cases as K
        cfg, model, _ = K.product_objects(K.ALL_CASES[args.case] if hasattr(K, "ALL_CASES") else K.CASES[args.case])
    else:
        import bench
        ns = argparse.Namespace(rk="Alexander2", prec="Jacobi", matrix_free=True, scheme="auto", b200=args.b200,
                                element=args.element)
        model = D.Model(D.Config(bench.ini_for(ns)), args.dim)
    return model.cuda_source_group(args.group)


def histogram(lines):
    ops = collections.Counter()
    for ln in lines:
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", ln)
        if m:
            ops[m.group(1)] += 1
    return ops


def hottest_loop(lines):
    """instructions between the target of the longest backward branch and the branch"""
    addr = []
    for i, ln in enumerate(lines):
        m = re.match(r"\s+/\*([0-9a-f]+)\*/", ln)
        if m:
            addr.append((int(m.group(1), 16), i))
    best = None
    for a, i in addr:
        m = re.search(r"\bBRA\b.*?0x([0-9a-f]+)", lines[i])
        if m:
            tgt = int(m.group(1), 16)
            if tgt < a and (best is None or a - tgt > best[1] - best[0]):
                best = (tgt, a)
    if not best:
        return lines
    return [lines[i] for a, i in addr if best[0] <= a <= best[1]]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--group", type=int, default=7)
    ap.add_argument("--case", default="")
    ap.add_argument("--dim", type=int, default=3)
    ap.add_argument("--element", default="p1")
    ap.add_argument("--b200", default="")
    ap.add_argument("--kernel", default="")
    ap.add_argument("--loop", action="store_true")
    ap.add_argument("--out", default="/tmp/dcb_sass")
    args = ap.parse_args()
    os.makedirs(args.out, exist_ok=True)
    src = os.path.join(args.out, f"group{args.group}.cu")
    cubin = os.path.join(args.out, f"group{args.group}.cubin")
    open(src, "w").write(source_for(args))
    r = subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-std=c++17", "-lineinfo", "--fmad=true",
                        "-cubin", "-o", cubin, src, "-Xptxas", "-v"], capture_output=True, text=True)
    if r.returncode != 0:
        print(r.stderr)
        raise SystemExit(1)
    info = {}
    cur = None
    for ln in r.stderr.splitlines():
        m = re.search(r"Compiling entry function '(\w+)'", ln)
        if m:
            cur = m.group(1)
        elif cur and ("registers" in ln or "spill" in ln):
            info.setdefault(cur, []).append(ln.replace("ptxas info    :", "").strip())
    names = [args.kernel] if args.kernel else sorted(info)
    for k in names:
        sass = subprocess.run(["cuobjdump", "-sass", "-fun", k, cubin], capture_output=True, text=True).stdout.splitlines()
        body = hottest_loop(sass) if args.loop else sass
        ops = histogram(body)
        f64 = sum(v for o, v in ops.items() if o in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"))
        top = ", ".join(f"{o} {v}" for o, v in ops.most_common(12))
        print(f"{k}: {' | '.join(info.get(k, []))}")
        print(f"   {'loop body' if args.loop else 'kernel'}: {sum(ops.values())} instr, fp64 {f64}  [{top}]")


if __name__ == "__main__":
    main()
