#!/usr/bin/env python
"""Headline benchmark: implicit time stepping of the 3-D two-species Gray-Scott system, CG-P1 on the
Kuhn (6 tets per cube) split of an n^3 lattice, Newton + BiCGSTAB/Jacobi -- BASELINE.json's metric
"DOF-updates/s & time steps/s, 3D Gray-Scott" on configs[3] (256^3: 16 974 593 vertices,
33 949 186 DOFs, 100 663 296 tets).

One "step" = one accepted time step of the reference's step operator (RungeKutta o Newton o
LinearSolver o instationary operator, dune/copasi/model/make_step_operator.hh:164-444).

  python bench.py --gpus N --steps K --warmup W          (N > 1: launched under torchrun)
  python bench.py --impl reference ...                   CPU arm: the oracle restatement on the
                                                         host cores, bounded sample of the workload

Prints ONE JSON line (rank 0).  value = DOFs * K / device time with the state resident in HBM;
e2e = the same through the host-buffer C-ABI calls (state uploaded before and downloaded after
every step inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import re
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "DOF-updates/s (3D Gray-Scott, CG-P1 Kuhn tets, implicit time stepping)"


def metric_name(args):
    if getattr(args, "element", "p1") == "q1":
        return "DOF-updates/s (3D Gray-Scott, CG-Q1 cubes, implicit time stepping)"
    return METRIC


def ini_for(args):
    over = {
        "model.time_step_operator.type": args.rk,
        "model.time_step_operator.linear_solver.type": "BiCGSTAB",
        "model.time_step_operator.linear_solver.preconditioner.type": args.prec,
        "model.time_step_operator.linear_solver.matrix_free": "true" if args.matrix_free else "false",
        "model.time_step_operator.linear_solver.convergence_condition.relative_tolerance": "1e-8",
        "model.time_step_operator.nonlinear_solver.convergence_condition.relative_tolerance": "1e-8",
        "model.time_step_operator.nonlinear_solver.dx_inverse_fixed_tolerance": "true",
        "model.assembly.b200.scheme": args.scheme,
    }
    for kv in filter(None, getattr(args, "b200", "").split(",")):
        k, v = kv.split("=")
        over["model.assembly.b200." + k] = v
    for kv in filter(None, getattr(args, "set", "").split(",")):
        k, v = kv.split("=")
        over[k] = v
    from dune_copasi_b200 import workloads as W
    name = getattr(args, "workload", "grayscott")
    if getattr(args, "mesh", "lattice") == "nested":
        name += "_nested"
    return W.ini_text(name, **over)


def precompile():
    """Warm the in-tree JIT cache with the bench model (called from __graft_entry__.build())."""
    import dune_copasi_b200 as D
    ns = argparse.Namespace(rk="Alexander2", prec="Jacobi", matrix_free=True, scheme="auto", b200="")
    for mf in (True, False):
        ns.matrix_free = mf
        D.Model(D.Config(ini_for(ns)), 3).precompile()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md).  The
    sampler runs from before the warm-up; begin()/end() bracket the timed region and only rows that
    arrived inside it (or, for very short regions, the nearest ones) are summarised."""

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index
        self.t0 = self.t1 = None

    def start(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [c.strip() for c in line.split(",")]))

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc:
            time.sleep(0.12)
            self.proc.terminate()
        rows = [(t, r) for t, r in self.rows if len(r) >= 8 and r[1].replace(".", "").isdigit()]
        inside = [r for t, r in rows if self.t0 is not None and self.t0 <= t <= self.t1 + 0.06]
        if not inside and rows and self.t0 is not None:
            mid = 0.5 * (self.t0 + self.t1)
            inside = [min(rows, key=lambda tr: abs(tr[0] - mid))[1]]
        sm = [float(r[1]) for r in inside]
        mx = [float(r[2]) for r in inside]
        pw = [float(r[3]) for r in inside if r[3].replace(".", "").isdigit()]
        reasons = set()
        for r in inside:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w": float(np.median(pw)) if pw else None, "reasons": sorted(reasons), "samples": len(sm)}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def measured_traffic(kernel, cells, world, elem=""):
    """DRAM bytes per launch (dram__bytes_read.sum + dram__bytes_write.sum) of the dominant kernel
    from the committed `ncu --set full` capture of this workload (profiles/ncu_traffic.json), or
    None when that configuration has not been captured."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if not os.path.exists(path):
        return None
    return json.load(open(path)).get(f"{kernel}@{cells}^3/n{world}" + elem)


def timeline(torch, run, rank, path):
    """Device timeline of `run()` from CUPTI activity records (torch.profiler): per kernel name the launches, the
    busy time and the idle time on the device right before it; what nsys would show (not installed here)."""
    import tempfile
    from torch.profiler import profile, ProfilerActivity
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as p:
        run()
    if rank != 0:
        return None
    with tempfile.NamedTemporaryFile(suffix=".json") as f:
        p.export_chrome_trace(f.name)
        ev = json.load(open(f.name))["traceEvents"]
    dev = sorted((e for e in ev if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset") and "dur" in e),
                 key=lambda e: e["ts"])
    if not dev:
        return {"error": "no device records"}
    per, end = {}, None
    for e in dev:
        name = e["name"]
        m = re.search(r"\b(k_\w+|dc_k_\w+)", name)       # our kernels: the function name without namespaces / arguments
        name = m.group(1) if m else name.split("(")[0].split("<")[0]
        if e["cat"] != "kernel":
            name = e["cat"] + ":" + name
        r = per.setdefault(name, {"n": 0, "busy_us": 0.0, "gap_before_us": 0.0})
        r["n"] += 1
        r["busy_us"] += e["dur"]
        if end is not None:
            r["gap_before_us"] += max(0.0, e["ts"] - end)
        end = max(end or 0.0, e["ts"] + e["dur"])
    span = dev[-1]["ts"] + dev[-1]["dur"] - dev[0]["ts"]
    busy = sum(r["busy_us"] for r in per.values())
    out = {"span_ms": span / 1e3, "busy_ms": busy / 1e3, "idle_ms": (span - busy) / 1e3, "records": len(dev),
           "kernels": {k: {"n": v["n"], "busy_ms": round(v["busy_us"] / 1e3, 4), "gap_before_ms": round(v["gap_before_us"] / 1e3, 4)}
                       for k, v in sorted(per.items(), key=lambda kv: -kv[1]["busy_us"] - kv[1]["gap_before_us"])}}
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    json.dump(out, open(path, "w"), indent=1)
    return {k: out[k] for k in ("span_ms", "busy_ms", "idle_ms", "records")}


def cpu_baseline(args, steps, warmup, cells):
    """The oracle restatement (kind "port": the reference itself needs a DUNE stack that is not
    available) with OpenMP over all host cores, on a bounded sample of the workload."""
    from oracle import core as ORC, ini as INI, mesh as OMESH
    # all host cores, whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1)
    ncores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    ORC.lib().orc_set_num_threads(int(ncores))
    cfg = INI.parse_ini(ini_for(args))
    dim = getattr(args, "dim", 3)
    if getattr(args, "mesh", "lattice") == "nested":
        from dune_copasi_b200 import meshgen
        coords, elems, keys, data = meshgen.nested_compartments(cells)
        mesh = OMESH.Mesh(dim=3, coords=coords, elems=elems)
        mesh.cell_keys, mesh.cell_data = keys, data
    else:
        mesh = OMESH.structured(dim, [cells] * dim, element="cube" if getattr(args, "element", "p1") == "q1" else "simplex")
    om = ORC.Model(cfg, mesh)
    S = ORC.StepOperator(om, par=1)
    u = om.initial(0.0)
    t = 0.0
    for _ in range(warmup):
        u, ok = S.apply(u, t, args.dt)
        t += args.dt
    t0 = time.perf_counter()
    for _ in range(steps):
        u, ok = S.apply(u, t, args.dt)
        assert ok
        t += args.dt
    wall = time.perf_counter() - t0
    cores = ORC.lib().orc_num_threads()
    return {"value": om.ndofs * steps / wall, "unit": "DOF-updates/s", "cores": int(cores), "kind": "port",
            "sample": f"{steps} steps (after {warmup} warm-up) of the same model on a {cells}^{dim} lattice "
                      f"({om.ndofs} DOFs), matrix based (the reference's Jacobi needs the assembled matrix), "
                      f"OpenMP over {cores} threads (assembly, SpMV, dot products, vector sweeps)",
            "ms_per_step": 1e3 * wall / steps}, om.ndofs, wall


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb, ndofs, wall = cpu_baseline(args, args.steps, args.warmup, args.cpu_cells)
    line = {"impl": "reference", "metric": metric_name(args), "value": cb["value"], "unit": "DOF-updates/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": cb["ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(config_dict(args, args.cpu_cells), matrix_free=False), "dofs": int(ndofs), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "DOF-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def config_dict(args, cells):
    dim = getattr(args, "dim", 3)
    name = {"cell": "cell3d_3comp_6species", "cell10": "cell3d_3comp_10species"}.get(getattr(args, "workload", "grayscott"),
                                                                                       f"grayscott{dim}d")
    elem = "q1_cubes" if getattr(args, "element", "p1") == "q1" else "p1_kuhn"
    if getattr(args, "mesh", "lattice") == "nested":
        elem = "p1_tets_nested_compartments"      # unstructured tetrahedral mesh of dune_copasi_b200/meshgen.py, 6 n^3 tets
    return {"workload": f"{name}_{elem}_{cells}^{dim}", "element": elem, "cells": cells,
            "collectives": getattr(args, "collectives", "none (1 GPU)"), "dt": args.dt, "rk": args.rk,
            "linear_solver": "BiCGSTAB", "preconditioner": args.prec, "matrix_free": bool(args.matrix_free),
            "linear_rel_tol": 1e-8, "newton_rel_tol": 1e-8, "assembly": args.scheme,
            "l2": getattr(args, "l2_note", "inputs larger than L2 (every vector exceeds 126 MB at the default size)"),
            "reference_reachability": ("matrix-free + Jacobi is the same operator the reference assembles: its registry "
                                       "offers Jacobi for the assembled matrix only (solver/istl/factory/preconditioner.hh:"
                                       "98-104); the matrix-based run of the same model is `assembled_variant`")
            if args.matrix_free and args.prec == "Jacobi" else "as the reference's registry offers it"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--cells", type=int, default=256)
    ap.add_argument("--dim", type=int, default=3, choices=[2, 3], help="3: the headline lattice; 2: SURVEY 8d S-2D squares")
    ap.add_argument("--element", default="p1", choices=["p1", "q1"],
                    help="p1: Kuhn simplices, the reference's element (headline); q1: the lattice cells as Q1 "
                         "elements, BASELINE configs[3]'s wording (not a reference capability, own oracle)")
    ap.add_argument("--cpu-cells", type=int, default=72, help="lattice of the bounded CPU sample (10-30 s of CPU work)")
    ap.add_argument("--no-assembled", action="store_true", help="skip the matrix-based variant that rides along")
    ap.add_argument("--assembled-cells", type=int, default=160, help="lattice of the matrix-based variant")
    ap.add_argument("--dt", type=float, default=1.0)
    ap.add_argument("--rk", default="Alexander2")
    ap.add_argument("--prec", default="Jacobi")
    ap.add_argument("--scheme", default="auto")
    ap.add_argument("--matrix-free", type=int, default=1)
    ap.add_argument("--workload", default="grayscott", choices=["grayscott", "cell", "cell10"],
                    help="grayscott: BASELINE configs[3] (headline); cell / cell10: 3-compartment cell model with "
                         "6 / 10 species (configs[4] in miniature, general unstructured kernels)")
    ap.add_argument("--mesh", default="lattice", choices=["lattice", "nested"],
                    help="nested: the cell workloads on the unstructured tetrahedral mesh of three nested "
                         "compartments (BASELINE configs[4]; dune_copasi_b200/meshgen.py, 6 cells^3 tets, RCB partition)")
    ap.add_argument("--b200", default="", help="model.assembly.b200.* overrides, e.g. patch_elements=768,patch_min_blocks=4")
    ap.add_argument("--set", default="", help="any ini key overrides, e.g. model.time_step_operator.linear_solver.b200.speculation=false")
    ap.add_argument("--timeline", default="", help="also run 2 steps under torch.profiler (CUPTI kernel records) and "
                    "write the per-kernel busy time and the idle gaps on the stream to this JSON file (rank 0)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-q1", action="store_true", help="skip the Q1 variant that rides along with the P1 headline")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import dune_copasi_b200 as D
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if D.lib().dcb_device_count() < 1:
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def measure(args):
        # ---- problem
        cfg = D.Config(ini_for(args))
        t_setup = time.perf_counter()
        if args.mesh == "nested":
            from dune_copasi_b200 import meshgen
            coords, elems, keys, data = meshgen.nested_compartments(args.cells)
            model = D.Model(cfg, 3, keys)
            gglobal = D.Grid.from_arrays(3, coords, elems, keys, data)
            del coords, elems, data
        else:
            model = D.Model(cfg, args.dim)
            gglobal = D.Grid.structured(args.dim, [args.cells] * args.dim, element="cube" if args.element == "q1" else "simplex")
        ne_global = gglobal.ne
        nv_global = gglobal.nv
        grid = gglobal.partition(rank, world) if world > 1 else gglobal
        grid.bind(model)
        op = D.Operator(model, grid)
        comm = None
        if world > 1:
            uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
            if rank == 0:
                uid = torch.tensor(list(D.Comm.unique_id()), dtype=torch.uint8, device="cuda")
            dist.broadcast(uid, 0)
            comm = D.Comm(bytes(uid.cpu().tolist()), rank, world, op)
            args.collectives = ("NVLink peer memory (own kernels: all-reduce, slab halo) + NCCL for setup"
                                if comm.uses_peer_memory else "NCCL")
            del gglobal
        owned = torch.tensor([sum(e - b for b, e in op.owned_ranges())], dtype=torch.int64, device="cuda")
        if dist is not None:
            dist.all_reduce(owned)
        ndofs_global = int(owned.item())
        st = D.Stepper(op, cfg, comm)
        # per-rank working set of a Krylov iteration: ~10 solver vectors + the state, plus the mesh arrays of unstructured grids
        ws_mb = (op.ndofs * 8 * 11 + (0 if args.mesh == "lattice" else grid.ne * (16 + 32) + grid.nv * 24)) / 1e6
        args.l2_note = (f"per-rank working set of a Krylov iteration ~{ws_mb:.0f} MB "
                        + ("exceeds the 126 MB L2: inputs larger than L2, no flush" if ws_mb > 126.0 else
                           "does NOT exceed the 126 MB L2 (small per-rank share): the sweeps run out of L2, stated here instead of flushing"))
        u0 = grid.interpolate(model, 0.0)
        st.set_state(u0, 0.0)
        t_setup = time.perf_counter() - t_setup
        stream = torch.cuda.ExternalStream(op.stream)

        def barrier():
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()

        def timed(nsteps, e2e):
            """-> device ms for nsteps (max over ranks)"""
            u_host = None
            if e2e:
                # host buffers of the step's input/result live in pinned memory
                u_host = torch.empty(op.ndofs, dtype=torch.float64).pin_memory().numpy()
                st.get_state(u_host)
            barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            for _ in range(nsteps):
                if e2e:
                    st.set_state(u_host, st.time)        # H2D of the step's input through the C ABI
                ok = st.step(args.dt)
                if not ok:
                    raise SystemExit("time step failed")
                if e2e:
                    st.get_state(u_host)                 # D2H of the step's result
            e1.record(stream)
            barrier()
            ms = torch.tensor([e0.elapsed_time(e1)], device="cuda", dtype=torch.float64)
            if dist is not None:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            return float(ms.item())

        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
        for _ in range(args.warmup):
            assert st.step(args.dt)
        # the state the timed steps start from: the end-to-end pass below is rewound to it, so that both
        # passes integrate the same K steps (same Newton / Krylov iteration counts)
        u_start = torch.empty(op.ndofs, dtype=torch.float64).pin_memory().numpy()
        st.get_state(u_start)
        t_start = st.time
        s0 = st.stats()
        sampler.begin()
        ms = timed(args.steps, False)
        sampler.end()
        s_timed = st.stats()
        clocks = sampler.stop() if rank == 0 else None
        # per-kernel durations: the same K steps replayed (state rewound) with a CUDA event pair around every
        # launch on the operator's stream -- kept out of the timed pass, where the event records would sit
        # between the kernels (measured: 2-7 % of the step at N = 8)
        st.set_state(u_start, t_start)
        D.lib().dcb_operator_profile(op.h, 1)
        ms_prof = timed(args.steps, False)
        prof = op.profile()
        host = {k: v for k, v in prof.items() if k.startswith("host_")}     # host-side timers (ms)
        prof = {k: v for k, v in prof.items() if not k.startswith("host_")}
        D.lib().dcb_operator_profile(op.h, 0)
        s1 = st.stats()
        tl = None
        if args.timeline:
            st.set_state(u_start, t_start)
            tl = timeline(torch, lambda: timed(min(args.steps, 2), False), rank, args.timeline)
        e2e = None
        if not args.no_e2e:
            st.set_state(u_start, t_start)
            s1e = st.stats()
            ms_e2e = timed(args.steps, True)
            s2 = st.stats()
            nbytes = op.ndofs * 8
            e2e = {"value": ndofs_global * args.steps / (ms_e2e * 1e-3), "unit": "DOF-updates/s",
                   "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes, "ms_per_step": ms_e2e / args.steps,
                   "solver_stats": {k: s2[k] - s1e[k] for k in s2},
                   "note": "same K steps as `value` (state rewound), host buffers in pinned memory, "
                           "H2D of the state before and D2H after every step inside the timed region"}

        def shutdown():
            # every rank tears down in the same order: library objects (their NCCL communicator)
            # first; torch's process group goes last, in main()
            nonlocal st, comm, op
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            st = None
            comm = None
            op = None

        if rank != 0:
            shutdown()
            return None
        d = {k: s_timed[k] - s0[k] for k in s_timed}
        value = ndofs_global * args.steps / (ms * 1e-3)
        # ---- roofline of the dominant kernel (by accumulated device time in the timed region)
        peak, peak_src = peaks()
        dim = args.dim
        nodes, tets = nv_global / world, (6 if dim == 3 else 2) * args.cells ** dim / world
        yfree_apply = (args.element == "p1" and args.workload == "grayscott" and args.matrix_free and args.prec == "Jacobi"
                       and "yfree=false" not in args.set and "tile=true" not in args.b200)
        alg = {  # algorithmic bytes per launch, SURVEY.md 8(d) (per rank)
            "patch_apply": nodes * (16 * 2 + 8 * dim + 8 * 2) + tets * 4 * (dim + 1),
            "patch_residual": nodes * (16 * 2 + 8 * dim) + tets * 4 * (dim + 1),
            "patch_bdiag": nodes * (8 * 2 + 8 * dim + 8 * 4) + tets * 4 * (dim + 1),
            "elem_apply": nodes * (16 * 2 + 8 * dim + 8 * 2) + tets * 4 * (dim + 1),
            "elem_residual": nodes * (16 * 2 + 8 * dim) + tets * 4 * (dim + 1),
            "spmv": op.ndofs * (54 if args.element == "q1" else 30) * 12 + op.ndofs * 20,
            # structured-implicit variant: no connectivity, no coordinates (32 B/vertex)
            "struct_residual": nodes * (16 * 2),
            # P1 default (linear_solver.b200.yfree): the apply also reads D^-1 and forms the Jacobi application itself:
            # read u, z, dinv / write y = 64 B per vertex at two species; otherwise read u, z / write y = 48
            "struct_apply": nodes * (16 * 2 + 8 * 2 + (8 * 2 if yfree_apply else 0)),
            "struct_bdiag": nodes * (8 * 2 + 8 * 4),
            # tile-marching sweeps with the BiCGSTAB updates fused in (mean of the two sweeps of an iteration:
            # read u, r, p, v, dinv, rt / write p, v = 128 B per vertex; read u, r, v, dinv / write r, t = 96)
            "tile_apply": nodes * 112,
            "tile_residual": nodes * (16 * 2 + 8 * 2),
        }
        if args.workload != "grayscott":
            # general meshes, several compartments (one launch per compartment and application): per
            # application read x, z and write y (24 B per dof), read the coordinates once per vertex and
            # compartment (8 d) and the connectivity (4 (d + 1) per element) -- SURVEY.md 8(d)
            ranges = op.owned_ranges()
            nsp = {}
            for _, c in model.species():
                nsp[c] = nsp.get(c, 0) + 1
            comps = [c for c in sorted(nsp)]
            verts = sum((e - b) / nsp[c] for (b, e), c in zip(ranges, comps))
            per_apply = op.ndofs * 24 + verts * 8 * dim + grid.ne * 4 * (dim + 1)
            alg["elem_apply"] = per_apply / max(1, len(comps))
            alg["elem_residual"] = (op.ndofs * 16 + verts * 8 * dim + grid.ne * 4 * (dim + 1)) / max(1, len(comps))
        top = max(prof, key=lambda k: prof[k]["ms"]) if prof else None
        roof = None
        if top:
            avg_ms = prof[top]["ms"] / max(1, prof[top]["launches"])
            ach = alg.get(top, 0.0) / (avg_ms * 1e-3) / 1e9
            roof = {"bound": "hbm", "kernel": top, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": measured_traffic(top, args.cells, world, "/q1" if args.element == "q1" else ""), "algorithmic_bytes": alg.get(top, 0.0),
                    "peak_source": peak_src, "avg_launch_ms": avg_ms, "launches": prof[top]["launches"],
                    "share_of_step": prof[top]["ms"] / ms_prof, "profiled_ms_per_step": ms_prof / args.steps,
                    "breakdown_ms_per_step": {k: v["ms"] / args.steps for k, v in prof.items()},
                    "host_ms_per_step": {k: v["ms"] / args.steps for k, v in host.items()}}
            # the assembly kernels are fp64-pipe bound on B200 (64 fp64 lanes/SM/clk), not HBM bound:
            # report the live fp64 instruction rate against that peak next to the HBM fraction
            path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
            key = f"{top}.fp64_instr_per_cell" + (".q1" if args.element == "q1" else "")
            per_cell = json.load(open(path)).get(key) if os.path.exists(path) else None
            if per_cell and dim == 3 and clocks and clocks.get("sm_mhz"):
                cells_rank = args.cells ** dim / world
                rate = cells_rank * per_cell / (avg_ms * 1e-3)
                peak64 = 148 * 64 * clocks["sm_mhz"] * 1e6
                roof["fp64_pipe"] = {"instr_per_cell": per_cell, "achieved_ginstr_s": rate / 1e9,
                                     "peak_ginstr_s": peak64 / 1e9, "frac": rate / peak64,
                                     "note": "dominant kernel is bound by the fp64 pipe; HBM traffic is ~1.3x algorithmic"}
        cb = None
        line = {"metric": metric_name(args), "value": value, "unit": "DOF-updates/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config_dict(args, args.cells),
                "time_steps_per_s": args.steps / (ms * 1e-3), "dofs": int(ndofs_global), "elements": int(ne_global), "e2e": e2e,
                "gpu_launches": int(d["kernel_launches"]), "clocks": clocks, "roofline": roof, "cpu_baseline": cb,
                "solver_stats": d, "setup_s": t_setup}
        if tl is not None:
            line["timeline"] = tl
        shutdown()
        return line

    line = measure(args)
    # BASELINE configs[3] words the lattice as "Q1": the same run on the cells as Q1 elements rides along
    # (the headline stays on the reference's own element, P1 on the Kuhn split -- SURVEY.md F3); single-GPU
    # runs only: under torchrun use --element q1 for the Q1 numbers
    q1 = None
    if args.element == "p1" and args.dim == 3 and args.workload == "grayscott" and not args.no_q1 and world == 1:
        qargs = argparse.Namespace(**vars(args))
        qargs.element = "q1"
        q1 = measure(qargs)
    # the reference-reachable Jacobi configuration: the assembled matrix (CSR fill + SpMV), on the largest
    # lattice whose matrix set-up fits the bench's time budget; single-GPU runs only
    asm = None
    if (args.matrix_free and args.element == "p1" and args.dim == 3 and args.workload == "grayscott" and world == 1
            and not args.no_assembled):
        aargs = argparse.Namespace(**vars(args))
        aargs.matrix_free, aargs.cells = 0, min(args.cells, args.assembled_cells)
        aargs.steps, aargs.warmup, aargs.no_e2e = min(args.steps, 3), min(args.warmup, 2), True
        asm = measure(aargs)
    if dist is not None:
        dist.destroy_process_group()
    if rank != 0:
        return
    if asm is not None:
        line["assembled_variant"] = {k: asm[k] for k in ("value", "unit", "ms_per_step", "dofs", "steps", "warmup", "gpu_launches",
                                                         "solver_stats", "setup_s")}
        line["assembled_variant"]["config"] = asm["config"]
        if asm["roofline"]:
            line["assembled_variant"]["roofline"] = {k: asm["roofline"][k] for k in
                                                     ("kernel", "achieved", "peak", "frac", "avg_launch_ms", "share_of_step",
                                                      "breakdown_ms_per_step") if k in asm["roofline"]}
    if q1 is not None:
        keep = ("metric", "value", "unit", "ms_per_step", "time_steps_per_s", "dofs", "e2e", "gpu_launches", "clocks",
                "solver_stats")
        line["q1_variant"] = {k: q1[k] for k in keep}
        line["q1_variant"]["config"] = q1["config"]
        line["q1_variant"]["roofline"] = {k: q1["roofline"][k] for k in
                                          ("kernel", "achieved", "peak", "frac", "avg_launch_ms", "share_of_step",
                                           "breakdown_ms_per_step") if q1["roofline"] and k in q1["roofline"]}
        if q1["roofline"] and "fp64_pipe" in q1["roofline"]:
            line["q1_variant"]["roofline"]["fp64_pipe"] = q1["roofline"]["fp64_pipe"]
    if not args.no_cpu_baseline and world == 1:
        # the CPU arm runs after the GPU is released (single-GPU runs only: the scaling runs carry no CPU leg)
        line["cpu_baseline"], _, _ = cpu_baseline(args, 2, 1, args.cpu_cells)
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
