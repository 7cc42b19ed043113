#!/usr/bin/env python
"""Warm the in-tree JIT cache with the tile-kernel variants tests/test_gpu_tile.py launches (small and
odd tile shapes), so that the GPU box does not spend its minutes in NVRTC.  Runs without a GPU."""
import os
import sys
from concurrent.futures import ProcessPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def job(args):
    name, over = args
    import cases as K
    import dune_copasi_b200 as D
    case = K.ALL_CASES[name]
    model = D.Model(D.Config(case.ini_with(**over)), case.dim, [])
    model.precompile_group(8 if case.element == "cube" else 7)
    return name


def main():
    import test_gpu_tile as T
    jobs = [(n, s) for n in T.P1 + T.Q1 for s in T.SHAPES.values()]
    with ProcessPoolExecutor(int(os.environ.get("JOBS", "4"))) as ex:
        for n in ex.map(job, jobs):
            pass
    print(len(jobs), "variants compiled")


if __name__ == "__main__":
    main()
