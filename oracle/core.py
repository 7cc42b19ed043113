"""ORACLE (test infrastructure, not product code) -- model tables, assembly drivers, Newton,
Runge-Kutta and the adaptive stepper of the reference, restated on top of oracle.c.

Follows (paths relative to /root/reference):
  Model.from_config      dune/copasi/model/diffusion_reaction/local_equations.hh:617-700 (which
                         terms exist), functor_factory_parser.impl.hh:116-182 (symbols, absent terms)
  pattern                local_operator.hh:276-399 + make_step_operator.hh:378-384 (sorted CSR)
  constraints            dune/copasi/model/constraints.hh:93-195 (translation constraints, no_value)
  Instationary / RK      make_step_operator.hh:408-443 + PDELab RK tables (third party, App. C.2)
  newton / linear        make_step_operator.hh:210-288 + PDELab NewtonOperator (third party, App. C.3)
  LinearSolver           make_step_operator.hh:102-146 (solver rebuilt on every call)
  stepper                dune/copasi/common/stepper.hh:95-125, 337-368
  reduce_l2              dune/copasi/model/diffusion_reaction/reduce.hh:157-203 (functional only)
PARITY STATUS: pinned by the reference's coarse KATs only; see oracle.c header.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess
from dataclasses import dataclass

import numpy as np

from . import expr as E
from . import ini as INI
from . import mesh as MESH

HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

(K_REACTION, K_REACTION_JAC, K_STORAGE, K_STORAGE_JAC, K_DIFF, K_DIFF_JAC, K_OUTFLOW, K_OUTFLOW_JAC,
 K_VEL, K_VEL_JAC, K_DIFF_T, K_DIFF_T_JAC, K_NKIND) = range(13)
AXES = "xyz"


def build_lib(force=False):
    out = os.path.join(HERE, "_build", "liboracle.so")
    src = os.path.join(HERE, "oracle.c")
    if force or not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["gcc", "-O3", "-march=native", "-fopenmp", "-shared", "-fPIC",
                               "-o", out, src, "-lm"])
    return out


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build_lib())
        _LIB.orc_eval.restype = C.c_double
    return _LIB


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


class CMesh(C.Structure):
    _fields_ = [("dim", C.c_int32), ("nv", C.c_int64), ("coords", C.POINTER(C.c_double)),
                ("ne", C.c_int64), ("elems", C.POINTER(C.c_int32)), ("elem_comp", C.POINTER(C.c_int32)),
                ("nkeys", C.c_int32), ("cell_data", C.POINTER(C.c_double)),
                ("elem_dof", C.POINTER(C.c_int64)), ("nf", C.c_int64),
                ("f_in", C.POINTER(C.c_int64)), ("f_out", C.POINTER(C.c_int64)),
                ("f_lin", C.POINTER(C.c_int32)), ("f_lout", C.POINTER(C.c_int32)), ("etype", C.c_int32)]


class CModel(C.Structure):
    _fields_ = [("ncomp", C.c_int32), ("nspec", C.c_int32), ("spec_comp", C.POINTER(C.c_int32)),
                ("spec_local", C.POINTER(C.c_int32)), ("comp_ptr", C.POINTER(C.c_int32)),
                ("comp_spec", C.POINTER(C.c_int32)), ("nterms", C.c_int32),
                ("terms", C.POINTER(C.c_int32)), ("tptr", C.POINTER(C.c_int32)),
                ("prog_ptr", C.POINTER(C.c_int32)), ("code", C.POINTER(C.c_int32)),
                ("const_ptr", C.POINTER(C.c_int32)), ("consts", C.POINTER(C.c_double)),
                ("nslots", C.c_int32), ("spec_base", C.c_int32)]


class CResult(C.Structure):
    _fields_ = [("iterations_x2", C.c_int32), ("converged", C.c_int32), ("reduction", C.c_double),
                ("norm0", C.c_double)]


@dataclass
class Species:
    name: str
    comp: int
    local: int
    cfg: dict


class Model:
    """Species, compartments and term tables of one ini file, bound to a mesh."""

    def __init__(self, cfg: dict, mesh: MESH.Mesh):
        self.cfg, self.mesh = cfg, mesh
        mcfg = INI.sub(cfg, "model")
        self.ctx = E.Context.from_config(INI.sub(cfg, "parser_context"))
        self.comp_names = [k for k, v in INI.sub(cfg, "compartments").items() if isinstance(v, dict)]
        fields = [(k, v) for k, v in INI.sub(mcfg, "scalar_field").items() if isinstance(v, dict)]
        self.species: list[Species] = []
        for c, cname in enumerate(self.comp_names):
            loc = 0
            for name, sc in fields:
                if sc.get("compartment") == cname:
                    self.species.append(Species(name, c, loc, sc))
                    loc += 1
        self.names = [s.name for s in self.species]
        self.ncomp, self.nspec = len(self.comp_names), len(self.species)
        self.comp_nspec = [sum(1 for s in self.species if s.comp == c) for c in range(self.ncomp)]
        self.is_linear = str(mcfg.get("is_linear", "false")).lower() in ("true", "1", "yes")
        # local_operator.hh:193-202, 234-235: numerical (FD) Jacobian unless the operator is linear
        self.numerical = INI.get(mcfg, "jacobian.type", "analytical") == "numerical" and not self.is_linear
        self.fd_eps = float(INI.get(mcfg, "jacobian.epsilon", 1e-7))
        self.sym = E.Symbols(mesh.dim, self.names, mesh.cell_keys)
        # --- compartments on the mesh (make_multi_domain_grid.hh:118-155)
        self._mark_compartments()
        MESH.build_facets(mesh)
        MESH.build_dofmap(mesh, self.comp_nspec)
        # --- terms
        self.progs: list = []
        terms = []

        def add(kind, i, j, k, text):
            if E.is_absent(text):
                return False
            self.progs.append(E.compile_expr(text, self.sym, self.ctx))
            terms.append((kind, i, j, k, len(self.progs) - 1))
            return True

        idx = {n: g for g, n in enumerate(self.names)}
        for g, sp in enumerate(self.species):
            sc = sp.cfg
            for kind, jkind, key in ((K_REACTION, K_REACTION_JAC, "reaction"), (K_STORAGE, K_STORAGE_JAC, "storage")):
                t = INI.sub(sc, key)
                if add(kind, g, -1, -1, t.get("expression")):
                    for wrt, jc in INI.sub(t, "jacobian").items():
                        if wrt in idx and isinstance(jc, dict):
                            add(jkind, g, idx[wrt], -1, jc.get("expression"))
            for wrt, dc in INI.sub(sc, "cross_diffusion").items():
                if wrt in idx and isinstance(dc, dict):
                    # make_tensor_apply (functor_factory_parser.impl.hh:65-112): `type` = scalar | tensor,
                    # read separately for the function and for each of its jacobian entries
                    def diffusion(cfgd, kind_s, kind_t, kk):
                        typ = cfgd.get("type", "scalar")
                        if typ == "scalar":
                            return add(kind_s, g, idx[wrt], kk if kk >= 0 else -1, cfgd.get("expression"))
                        if typ != "tensor":
                            raise ValueError(f"not known type 'scalar_value.{sp.name}.cross_diffusion.type = {typ}'")
                        active = False
                        for r in range(mesh.dim):
                            for c in range(mesh.dim):
                                ent = cfgd.get(AXES[r] + AXES[c])
                                if isinstance(ent, dict):
                                    code3 = 3 * r + c + (9 * kk if kk >= 0 else 0)
                                    active |= add(kind_t, g, idx[wrt], code3, ent.get("expression"))
                        return active
                    if diffusion(dc, K_DIFF, K_DIFF_T, -1):
                        for k, jc in INI.sub(dc, "jacobian").items():
                            if k in idx and isinstance(jc, dict):
                                diffusion(jc, K_DIFF_JAC, K_DIFF_T_JAC, idx[k])
            for cname, oc in INI.sub(sc, "outflow").items():
                if cname in self.comp_names and isinstance(oc, dict):
                    l = self.comp_names.index(cname)
                    if add(K_OUTFLOW, g, l, -1, oc.get("expression")):
                        for k, jc in INI.sub(oc, "jacobian").items():
                            if k in idx and isinstance(jc, dict):
                                add(K_OUTFLOW_JAC, g, l, idx[k], jc.get("expression"))
            # velocity.<axis>.expression (make_vector, functor_factory_parser.impl.hh:36-61) with
            # velocity.jacobian.<wrt>.<axis>.expression (local_equations.hh:663-669)
            vc = INI.sub(sc, "velocity")
            active = False
            for ax in range(mesh.dim):
                ent = vc.get(AXES[ax])
                if isinstance(ent, dict):
                    active |= add(K_VEL, g, ax, -1, ent.get("expression"))
            if active:
                for k, jc in INI.sub(vc, "jacobian").items():
                    if k in idx and isinstance(jc, dict):
                        for ax in range(mesh.dim):
                            ent = jc.get(AXES[ax])
                            if isinstance(ent, dict):
                                add(K_VEL_JAC, g, idx[k], ax, ent.get("expression"))
        terms.sort(key=lambda t: (t[1], t[0]))
        self.terms = np.asarray(terms, dtype=np.int32).reshape(-1, 5)
        tptr = np.zeros(self.nspec * K_NKIND + 1, dtype=np.int32)
        for t in terms:
            tptr[t[1] * K_NKIND + t[0] + 1] += 1
        self.tptr = np.cumsum(tptr).astype(np.int32)
        self._pack()
        self.has_outflow = bool((self.terms[:, 0] == K_OUTFLOW).any()) if len(terms) else False
        if mesh.etype == 1 and self.has_outflow:
            raise NotImplementedError("outflow terms on Q1 cube grids are out of scope")

    # ------------------------------------------------------------------ compartments
    def _mark_compartments(self):
        m = self.mesh
        comp = -np.ones(m.ne, dtype=np.int32)
        sym = E.Symbols(m.dim, [], m.cell_keys)
        cen = m.centers()
        for c, cname in enumerate(self.comp_names):
            text = INI.sub(self.cfg, "compartments")[cname].get("expression", "0")
            ast = E.resolve(E.Parser(text).parse(), self.ctx)
            code, consts = [], []
            E.emit(ast, sym, code, consts)
            ctx = np.zeros((m.ne, sym.nslots))
            ctx[:, E.SLOT_POS:E.SLOT_POS + m.dim] = cen
            for k in range(len(m.cell_keys)):
                ctx[:, E.SLOT_CELL + k] = m.cell_data[k]
            val = eval_program(code, consts, ctx)
            hit = np.abs(val) > 1e-8 * np.maximum(1.0, np.abs(val))   # FloatCmp::ne(val, 0)
            if np.any(hit & (comp >= 0)):
                raise NotImplementedError("overlapping compartments are out of scope")
            comp[hit] = c
        m.elem_comp = comp

    def _pack(self):
        code, consts, pptr, cptr = [], [], [0], []
        for c, k in self.progs:
            cptr.append(len(consts))
            code += c
            consts += k
            pptr.append(len(code))
        self._code = np.asarray(code or [0], dtype=np.int32)
        self._consts = np.asarray(consts or [0.0], dtype=np.float64)
        self._pptr = np.asarray(pptr, dtype=np.int32)
        self._cptr = np.asarray(cptr or [0], dtype=np.int32)
        self._spec_comp = np.asarray([s.comp for s in self.species], dtype=np.int32)
        self._spec_local = np.asarray([s.local for s in self.species], dtype=np.int32)
        self._comp_ptr = np.concatenate([[0], np.cumsum(self.comp_nspec)]).astype(np.int32)
        self._comp_spec = np.arange(self.nspec, dtype=np.int32)   # species are compartment-major
        m = self.mesh
        self._cell = np.ascontiguousarray(m.cell_data if m.cell_data is not None else np.zeros((0, m.ne)))
        self._ed = np.ascontiguousarray(m.elem_dof)
        self.cmodel = CModel(self.ncomp, self.nspec, _p(self._spec_comp, C.c_int32),
                             _p(self._spec_local, C.c_int32), _p(self._comp_ptr, C.c_int32),
                             _p(self._comp_spec, C.c_int32), len(self.terms),
                             _p(np.ascontiguousarray(self.terms), C.c_int32), _p(self.tptr, C.c_int32),
                             _p(self._pptr, C.c_int32), _p(self._code, C.c_int32),
                             _p(self._cptr, C.c_int32), _p(self._consts, C.c_double),
                             self.sym.nslots, self.sym.spec_base)
        self._terms_c = np.ascontiguousarray(self.terms)
        self.cmodel.terms = _p(self._terms_c, C.c_int32)
        self.cmesh = CMesh(m.dim, m.nv, _p(m.coords, C.c_double), m.ne, _p(m.elems, C.c_int32),
                           _p(m.elem_comp, C.c_int32), len(m.cell_keys), _p(self._cell, C.c_double),
                           _p(self._ed, C.c_int64), len(m.f_in), _p(m.f_in, C.c_int64),
                           _p(m.f_out, C.c_int64), _p(m.f_lin, C.c_int32), _p(m.f_lout, C.c_int32), m.etype)

    @property
    def ndofs(self):
        return self.mesh.ndofs

    # ------------------------------------------------------------------ DOF helpers
    def dof_positions(self):
        """-> (coords [ndofs, dim], species id [ndofs])"""
        m = self.mesh
        pos = np.zeros((m.ndofs, m.dim))
        spec = np.zeros(m.ndofs, dtype=np.int32)
        for g, sp in enumerate(self.species):
            v = m.comp_vertices[sp.comp]
            d = m.comp_offset[sp.comp] + np.arange(v.size) * self.comp_nspec[sp.comp] + sp.local
            pos[d] = m.coords[v]
            spec[d] = g
        return pos, spec

    def eval_at_dofs(self, key: str, time: float, default=None, boundary=False):
        """Evaluate <species>.<key>.expression at every DOF's vertex (P1 interpolation of initial
        conditions: make_initial.hh:26-90, interpolate.hh:22-78)."""
        pos, spec = self.dof_positions()
        out = np.zeros(self.ndofs)
        for g, sp in enumerate(self.species):
            text = INI.get(sp.cfg, key + ".expression", default)
            sel = spec == g
            if text is None or E.is_absent(text):
                out[sel] = 0.0
                continue
            code, consts = E.compile_expr(text, self.sym, self.ctx)
            ctx = np.zeros((int(sel.sum()), self.sym.nslots))
            ctx[:, E.SLOT_TIME] = time
            ctx[:, E.SLOT_POS:E.SLOT_POS + self.mesh.dim] = pos[sel]
            if boundary:
                ctx[:, E.SLOT_INBND] = 1.0
            else:
                ctx[:, E.SLOT_INVOL] = 1.0
            out[sel] = eval_program(code, consts, ctx)
        return out

    def initial(self, time: float):
        return self.eval_at_dofs("initial", time)

    def constraints(self):
        """-> (dof indices, values) of Dirichlet translation constraints (constraints.hh:114-195).
        `time` is NaN there; a value of no_value (DBL_MAX) means unconstrained.  A vertex of a face keeps its
        value when it is a boundary vertex exactly if the face is a boundary face (:182): constrain.boundary
        binds the boundary vertices, constrain.skeleton the others (every one of them lies on a face with a
        neighbour); constrain.volume binds codim-0 dofs, which P1 / Q1 do not have (:93-112).  The
        face-dependent symbols (normal_*, entity_volume) read 0."""
        m = self.mesh
        isb = np.zeros(m.nv, dtype=bool)
        isb[m.boundary_vertices] = True
        dofs, vals = [], []
        for g, sp in enumerate(self.species):
            v = m.comp_vertices[sp.comp]
            d = m.comp_offset[sp.comp] + np.arange(v.size) * self.comp_nspec[sp.comp] + sp.local
            dd, vv = [], []
            for key, slot, sel in (("constrain.boundary.expression", E.SLOT_INBND, isb[v]),
                                   ("constrain.skeleton.expression", E.SLOT_INSKEL, ~isb[v])):
                text = INI.get(sp.cfg, key)
                if text is None or E.is_absent(text):
                    continue
                code, consts = E.compile_expr(text, self.sym, self.ctx)
                ctx = np.zeros((int(sel.sum()), self.sym.nslots))
                ctx[:, E.SLOT_TIME] = np.nan
                ctx[:, slot] = 1.0
                ctx[:, E.SLOT_POS:E.SLOT_POS + m.dim] = m.coords[v[sel]]
                val = eval_program(code, consts, ctx)
                ok = val != E.DBL_MAX
                dd.append(d[sel][ok])
                vv.append(val[ok])
            if dd:
                dd, vv = np.concatenate(dd), np.concatenate(vv)
                order = np.argsort(dd, kind="stable")       # ascending dofs within a species, as the product walks them
                dofs.append(dd[order])
                vals.append(vv[order])
        if not dofs:
            return np.zeros(0, dtype=np.int64), np.zeros(0)
        return np.concatenate(dofs), np.concatenate(vals)

    # ------------------------------------------------------------------ sparsity pattern
    def species_pairs(self):
        """(i, j) species couplings of the volume pattern, local_operator.hh:276-338."""
        pairs = set()
        for kind, i, j, k, _ in self.terms:
            if kind in (K_REACTION_JAC, K_STORAGE_JAC):
                pairs.add((i, j))
            elif kind == K_STORAGE:
                pairs.add((i, i))
            elif kind in (K_DIFF, K_DIFF_T):
                pairs.add((i, j))
            elif kind == K_DIFF_JAC:
                pairs.add((i, k))
            elif kind == K_DIFF_T_JAC:
                pairs.add((i, k // 9))
            elif kind == K_VEL:                 # local_operator.hh:304-307
                pairs.add((i, i))
            elif kind == K_VEL_JAC:             # :308-313
                pairs.add((i, j))
        return sorted(p for p in pairs if self.species[p[0]].comp == self.species[p[1]].comp)

    def pattern(self):
        """Sorted CSR pattern: volume links + skeleton/boundary links (local_operator.hh:340-399)."""
        import scipy.sparse as sp
        m = self.mesh
        nd = m.elems.shape[1]
        rows, cols = [], []
        for (i, j) in self.species_pairs():
            c = self.species[i].comp
            sel = m.elem_comp == c
            ed = m.elem_dof[sel]
            r = ed[:, :, None] + self.species[i].local
            cc = ed[:, None, :] + self.species[j].local
            rows.append(np.broadcast_to(r, (ed.shape[0], nd, nd)).ravel())
            cols.append(np.broadcast_to(cc, (ed.shape[0], nd, nd)).ravel())
        # skeleton links
        for kind, i, l, k, _ in self.terms:
            if kind != K_OUTFLOW_JAC:
                continue
            ci, ck = self.species[i].comp, self.species[k].comp
            for f in range(len(m.f_in)):
                ei, eo = m.f_in[f], m.f_out[f]
                for (e, eother) in ((ei, eo), (eo, ei)):
                    if e < 0 or m.elem_comp[e] != ci:
                        continue
                    target = m.elem_comp[eother] if eo >= 0 else ci
                    if target != l:
                        continue
                    ew = e if ck == ci else (eother if (eo >= 0 and m.elem_comp[eother] == ck) else -1)
                    if ew < 0:
                        continue
                    fv = [a for a in range(nd) if a != (m.f_lin[f] if e == ei else m.f_lout[f])]
                    gv = m.elems[e, fv]
                    if ew != e and self.reference_compat:
                        fw = fv      # column dof_j of the other element carries phi_j of this one (:1133-1143)
                    else:
                        fw = [int(np.nonzero(m.elems[ew] == g)[0][0]) for g in gv]
                    r = m.elem_dof[e, fv] + self.species[i].local
                    cc = m.elem_dof[ew, fw] + self.species[k].local
                    rows.append(np.repeat(r, len(fv)))
                    cols.append(np.tile(cc, len(fv)))
        n = m.ndofs
        if rows:
            rows, cols = np.concatenate(rows), np.concatenate(cols)
        else:
            rows, cols = np.zeros(0, np.int64), np.zeros(0, np.int64)
        A = sp.coo_matrix((np.ones(rows.size, dtype=np.int8), (rows, cols)), shape=(n, n)).tocsr()
        A.sum_duplicates()
        A.sort_indices()
        return A.indptr.astype(np.int64), A.indices.astype(np.int32)

    # ------------------------------------------------------------------ operators
    @property
    def reference_compat(self):
        """model.b200.reference_compat (default true): facet terms exactly as local_operator.hh:903-916, :1298"""
        v = str(INI.get(self.cfg, "model.b200.reference_compat", "true")).strip().lower()
        return v not in ("false", "0", "no", "off")

    def residual(self, form, time, w, x, r, par=0):
        L = lib()
        L.orc_set_reference_compat(C.c_int(int(self.reference_compat)))
        L.orc_residual_volume(C.byref(self.cmesh), C.byref(self.cmodel), C.c_int(form), C.c_double(time),
                              C.c_double(w), _p(x, C.c_double), _p(r, C.c_double), C.c_int(par))
        if form == 0 and self.has_outflow:
            L.orc_residual_skeleton(C.byref(self.cmesh), C.byref(self.cmodel), C.c_double(time),
                                    C.c_double(w), _p(x, C.c_double), _p(r, C.c_double))

    def jacobian(self, form, time, w, x, rowptr, colidx, vals, par=0, numerical=False, eps=1e-7):
        L = lib()
        L.orc_set_reference_compat(C.c_int(int(self.reference_compat)))
        if numerical:
            L.orc_jacobian_volume_numerical(C.byref(self.cmesh), C.byref(self.cmodel), C.c_int(form),
                                            C.c_double(time), C.c_double(w), C.c_double(eps),
                                            _p(x, C.c_double), _p(rowptr, C.c_int64),
                                            _p(colidx, C.c_int32), _p(vals, C.c_double))
        else:
            L.orc_jacobian_volume(C.byref(self.cmesh), C.byref(self.cmodel), C.c_int(form), C.c_double(time),
                                  C.c_double(w), _p(x, C.c_double), _p(rowptr, C.c_int64),
                                  _p(colidx, C.c_int32), _p(vals, C.c_double), C.c_int(par))
        if form == 0 and self.has_outflow and numerical:
            L.orc_jacobian_skeleton_numerical(C.byref(self.cmesh), C.byref(self.cmodel), C.c_double(time),
                                              C.c_double(w), C.c_double(eps), C.c_int64(self.ndofs),
                                              _p(x, C.c_double), _p(rowptr, C.c_int64),
                                              _p(colidx, C.c_int32), _p(vals, C.c_double))
        elif form == 0 and self.has_outflow:
            L.orc_jacobian_skeleton(C.byref(self.cmesh), C.byref(self.cmodel), C.c_double(time),
                                    C.c_double(w), _p(x, C.c_double), _p(rowptr, C.c_int64),
                                    _p(colidx, C.c_int32), _p(vals, C.c_double))

    def jacobian_apply(self, form, time, w, x, z, y, par=0):
        L = lib()
        L.orc_set_reference_compat(C.c_int(int(self.reference_compat)))
        L.orc_jacobian_apply_volume(C.byref(self.cmesh), C.byref(self.cmodel), C.c_int(form),
                                    C.c_double(time), C.c_double(w), _p(x, C.c_double),
                                    _p(z, C.c_double), _p(y, C.c_double), C.c_int(par))
        if form == 0 and self.has_outflow:
            L.orc_jacobian_apply_skeleton(C.byref(self.cmesh), C.byref(self.cmodel), C.c_double(time),
                                          C.c_double(w), _p(x, C.c_double), _p(z, C.c_double),
                                          _p(y, C.c_double))


def eval_program(code, consts, ctx: np.ndarray) -> np.ndarray:
    """Run one RPN program on many contexts through the C VM."""
    L = lib()
    code = np.asarray(code, dtype=np.int32)
    consts = np.asarray(consts if len(consts) else [0.0], dtype=np.float64)
    ctx = np.ascontiguousarray(ctx, dtype=np.float64)
    out = np.empty(ctx.shape[0])
    n, ns = ctx.shape
    L.orc_eval_rows(_p(code, C.c_int32), C.c_int(code.size), _p(consts, C.c_double), _p(ctx, C.c_double),
                    C.c_int64(n), C.c_int64(ns), _p(out, C.c_double))
    return out


# ---------------------------------------------------------------------- linear solver
def linear_solve(rowptr, colidx, vals, b, cfg: dict, rel_tol: float, par=0):
    """LinearSolver::apply (make_step_operator.hh:102-146): a new solver + preconditioner per call.
    Returns (z, result).  Registry names: factory/iterative.hh:87-106, preconditioner.hh:96-113."""
    typ = cfg.get("type", "BiCGSTAB")
    n = b.size
    z = np.zeros(n)
    rhs = b.copy()
    pc = INI.sub(cfg, "preconditioner")
    ptype = pc.get("type", "SSOR")     # DUNE_COPASI_DEFAULT_PRECONDITIONER (factory/preconditioner.hh:17)
    kind = {"Richardson": 0, "Jacobi": 1, "BlockJacobi": 2, "SSOR": 3, "SOR": 4, "GaussSeidel": 5}[ptype]
    bs = int(pc.get("block_size", 1))
    relax = float(pc.get("relaxation", 1.0))
    maxit = int(str(INI.get(cfg, "convergence_condition.iteration_range", "1 500")).split()[-1])
    res = CResult()
    L = lib()
    L.orc_set_preconditioner_iterations(C.c_int(int(pc.get("iterations", 1))))   # preconditioner.hh:96-113
    if typ == "RestartedGMRes":
        restart = int(cfg.get("restart", 40))          # factory/iterative.hh:64
        L.orc_gmres(C.c_int64(n), _p(rowptr, C.c_int64), _p(colidx, C.c_int32), _p(vals, C.c_double),
                    _p(z, C.c_double), _p(rhs, C.c_double), C.c_double(rel_tol), C.c_int(maxit), C.c_int(restart),
                    C.c_int(kind), C.c_int(bs), C.c_double(relax), C.c_int(par), C.byref(res))
        return z, res
    fn = {"BiCGSTAB": L.orc_bicgstab, "CG": L.orc_cg}[typ]
    fn(C.c_int64(n), _p(rowptr, C.c_int64), _p(colidx, C.c_int32), _p(vals, C.c_double),
       _p(z, C.c_double), _p(rhs, C.c_double), C.c_double(rel_tol), C.c_int(maxit), C.c_int(kind),
       C.c_int(bs), C.c_double(relax), C.c_int(par), C.byref(res))
    return z, res


# ---------------------------------------------------------------------- RK tables (App. C.2)
def rk_table(name: str):
    if name == "ImplicitEuler":
        return np.array([[-1.0, 1.0]]), np.array([[0.0, 1.0]]), np.array([0.0, 1.0])
    if name == "Alexander2":
        al = 1.0 - math.sqrt(2.0) / 2.0
        return (np.array([[-1.0, 1.0, 0.0], [-1.0, 0.0, 1.0]]),
                np.array([[0.0, al, 0.0], [0.0, 1.0 - al, al]]), np.array([0.0, al, 1.0]))
    # explicit schemes: the stage system only carries the mass form (b_ii = 0)
    if name == "ExplicitEuler":
        return np.array([[-1.0, 1.0]]), np.array([[1.0, 0.0]]), np.array([0.0, 1.0])
    if name == "Heun":
        return (np.array([[-1.0, 1.0, 0.0], [-0.5, -0.5, 1.0]]),
                np.array([[1.0, 0.0, 0.0], [0.0, 0.5, 0.0]]), np.array([0.0, 1.0, 1.0]))
    if name == "Shu3":
        return (np.array([[-1.0, 1.0, 0.0, 0.0], [-0.75, -0.25, 1.0, 0.0], [-1.0 / 3.0, 0.0, -2.0 / 3.0, 1.0]]),
                np.array([[1.0, 0.0, 0.0, 0.0], [0.0, 0.25, 0.0, 0.0], [0.0, 0.0, 2.0 / 3.0, 0.0]]),
                np.array([0.0, 1.0, 0.5, 1.0]))
    if name == "RungeKutta4":
        return (np.array([[-1.0, 1.0, 0.0, 0.0, 0.0], [-1.0, 0.0, 1.0, 0.0, 0.0], [-1.0, 0.0, 0.0, 1.0, 0.0],
                          [-1.0, 0.0, 0.0, 0.0, 1.0]]),
                np.array([[0.5, 0.0, 0.0, 0.0, 0.0], [0.0, 0.5, 0.0, 0.0, 0.0], [0.0, 0.0, 1.0, 0.0, 0.0],
                          [1.0 / 6.0, 1.0 / 3.0, 1.0 / 3.0, 1.0 / 6.0, 0.0]]),
                np.array([0.0, 0.5, 0.5, 1.0, 1.0]))
    if name == "Alexander3":
        al = 0.4358665215
        b1 = -(6.0 * al * al - 16.0 * al + 1.0) / 4.0
        b2 = (6.0 * al * al - 20.0 * al + 5.0) / 4.0
        return (np.array([[-1.0, 1.0, 0.0, 0.0], [-1.0, 0.0, 1.0, 0.0], [-1.0, 0.0, 0.0, 1.0]]),
                np.array([[0.0, al, 0.0, 0.0], [0.0, (1.0 - al) / 2.0, al, 0.0], [0.0, b1, b2, al]]),
                np.array([0.0, al, (1.0 + al) / 2.0, 1.0]))
    if name == "FractionalStepTheta":
        # PDELab::FractionalStepParameter (third party; make_step_operator.hh:434-435): three
        # sub-steps theta, 1-2theta, theta with the implicit weight alpha*theta in each
        th = 1.0 - 0.5 * math.sqrt(2.0)
        alpha = 2.0 - math.sqrt(2.0)
        beta = 1.0 - alpha
        return (np.array([[-1.0, 1.0, 0.0, 0.0], [0.0, -1.0, 1.0, 0.0], [0.0, 0.0, -1.0, 1.0]]),
                np.array([[beta * th, alpha * th, 0.0, 0.0], [0.0, alpha * (1.0 - 2.0 * th), alpha * th, 0.0],
                          [0.0, 0.0, beta * th, alpha * th]]),
                np.array([0.0, th, 1.0 - th, 1.0]))
    raise NotImplementedError(name)


class StepOperator:
    """RungeKutta o (Newton | linear defect correction) o LinearSolver o instationary assembler
    (make_step_operator.hh:164-444), matrix based."""

    def __init__(self, model: Model, cfg: dict | None = None, par=0):
        self.model, self.par = model, par
        cfg = cfg if cfg is not None else INI.sub(model.cfg, "model.time_step_operator")
        self.cfg = cfg
        self.rk = cfg.get("type", "Alexander2")
        self.lin_cfg = INI.sub(cfg, "linear_solver")
        self.lin_rel = float(INI.get(self.lin_cfg, "convergence_condition.relative_tolerance", 1e-4))
        nl = INI.sub(cfg, "nonlinear_solver")
        self.n_rel = float(INI.get(nl, "convergence_condition.relative_tolerance", 1e-4))
        self.n_abs = float(INI.get(nl, "convergence_condition.absolute_tolerance", 0.0))
        rng = str(INI.get(nl, "convergence_condition.iteration_range", "0 40")).split()
        self.n_maxit = int(rng[-1])
        self.fixed_tol = str(nl.get("dx_inverse_fixed_tolerance", "false")).lower() in ("true", "1")
        self.min_rel_tol = float(nl.get("dx_inverse_min_relative_tolerance", 0.1))
        self.rowptr, self.colidx = model.pattern()
        self.cdofs, self.cvals = model.constraints()
        self.stats = {"newton_its": 0, "linear_its_x2": 0, "linear_solves": 0, "residuals": 0}

    # r += sum of mass/stiffness residual contributions of unknown stage u
    def _stage_residual(self, u, t, wM, wA, r):
        if wM != 0.0:
            self.model.residual(1, t, wM, u, r, self.par)
        if wA != 0.0:
            self.model.residual(0, t, wA, u, r, self.par)
        self.stats["residuals"] += 1

    def _stage_jacobian(self, u, t, wM, wA):
        vals = np.zeros(self.colidx.size)
        num, eps = self.model.numerical, self.model.fd_eps
        if wM != 0.0:
            self.model.jacobian(1, t, wM, u, self.rowptr, self.colidx, vals, self.par, numerical=num, eps=eps)
        if wA != 0.0:
            self.model.jacobian(0, t, wA, u, self.rowptr, self.colidx, vals, self.par, numerical=num, eps=eps)
        self._constrain_matrix(vals)
        return vals

    def _constrain_matrix(self, vals):
        if self.cdofs.size == 0:
            return
        isc = np.zeros(self.model.ndofs, dtype=bool)
        isc[self.cdofs] = True
        rows = np.repeat(np.arange(self.model.ndofs), np.diff(self.rowptr))
        kill = isc[rows] | isc[self.colidx]
        vals[kill] = 0.0
        vals[(rows == self.colidx) & isc[rows]] = 1.0

    def apply(self, u0, t0, dt):
        """One time step. Returns (u_new, ok)."""
        a, b, d = rk_table(self.rk)
        nst = a.shape[0]
        us = [u0.copy()]
        for s in range(nst):
            const = np.zeros_like(u0)
            for j in range(s + 1):
                tj = t0 + d[j] * dt
                self._stage_residual(us[j], tj, a[s, j], dt * b[s, j], const)
            wM, wA, ts = a[s, s + 1], dt * b[s, s + 1], t0 + d[s + 1] * dt
            x = us[-1].copy()
            x[self.cdofs] = self.cvals
            ok = self._solve_stage(x, ts, wM, wA, const)
            if not ok:
                return u0, False
            us.append(x)
        return us[-1], True

    def _residual(self, x, ts, wM, wA, const):
        r = const.copy()
        self._stage_residual(x, ts, wM, wA, r)
        r[self.cdofs] = 0.0
        return r

    def _solve_stage(self, x, ts, wM, wA, const):
        if self.model.is_linear:
            # defect correction, make_step_operator.hh:215-243
            r = self._residual(x, ts, wM, wA, const)
            vals = self._stage_jacobian(x, ts, wM, wA)
            z, res = linear_solve(self.rowptr, self.colidx, vals, r, self.lin_cfg, self.lin_rel, self.par)
            self.stats["linear_its_x2"] += res.iterations_x2
            self.stats["linear_solves"] += 1
            if not res.converged:
                return False
            x -= z
            return True
        # Newton (PDELab NewtonOperator, third party; classic PDELab rule restated, App. C.3):
        # "defect" is whatever the configured norm returns -- ||r||_2^2 by default
        # (make_step_operator.hh:274-276).
        r = self._residual(x, ts, wM, wA, const)
        cur = float(r @ r)
        first, prev = cur, cur
        stop = max(first * self.n_rel, self.n_abs)
        it = 0
        while cur > stop:
            if it >= self.n_maxit or not math.isfinite(cur):
                return False
            if self.fixed_tol:
                lin_tol = self.lin_rel
            elif stop / (10 * cur) > cur * cur / (prev * prev):
                lin_tol = stop / (10 * cur)
            else:
                lin_tol = min(self.min_rel_tol, cur * cur / (prev * prev))
            vals = self._stage_jacobian(x, ts, wM, wA)
            z, res = linear_solve(self.rowptr, self.colidx, vals, r, self.lin_cfg, lin_tol, self.par)
            self.stats["linear_its_x2"] += res.iterations_x2
            self.stats["linear_solves"] += 1
            if not res.converged:
                return False
            x -= z
            r = self._residual(x, ts, wM, wA, const)
            prev, cur = cur, float(r @ r)
            it += 1
            self.stats["newton_its"] += 1
        return True


def _fc_eq(a, b):
    """dune-common FloatCmp::eq with its defaults (relativeWeak, 8 ulp)"""
    return abs(a - b) <= 8.0 * 2.220446049250313e-16 * max(abs(a), abs(b))


def _fc_le(a, b):
    return a < b or _fc_eq(a, b)


def _fc_lt(a, b):
    return a < b and not _fc_eq(a, b)


def evolve(step: StepOperator, u, t0, t_end, dt0, dt_max=None, dt_min=None, inc=1.1, dec=0.5,
           on_step=None, steps_taken=None):
    """TimeStepper::evolve with snap_to_end_time (stepper.hh:145-176), snap_to_time (:192-239) and
    SimpleAdaptiveStepper::do_step / check_dt (:337-386), literally: full steps while two of them still
    fit before t_end, then the remainder in ceil(remainder / dt) equal steps, recomputed after every step
    (dt keeps growing by `inc`), halved after a failure (at most 100 times)."""
    state = {"u": u, "t": t0, "dt": dt0, "n": 0}

    def check_dt(dt):
        if dt_min is not None and _fc_lt(abs(dt), abs(dt_min)):
            return False
        if dt_max is not None and not _fc_le(abs(dt), abs(dt_max)):
            return False
        return True

    def do_step():
        if not check_dt(state["dt"]):
            return False
        un, ok = step.apply(state["u"], state["t"], state["dt"])
        while not ok:
            state["dt"] *= dec
            if not check_dt(state["dt"]):
                return False
            un, ok = step.apply(state["u"], state["t"], state["dt"])
        state["u"], state["t"] = un, state["t"] + state["dt"]
        if steps_taken is not None:
            steps_taken.append(state["dt"])
        nxt = state["dt"] * inc
        if dt_max is not None:
            nxt = min(max(nxt, -abs(dt_max)), abs(dt_max))
        state["dt"] = nxt
        state["n"] += 1
        if on_step:
            on_step(state["t"], state["u"])
        return True

    while _fc_le(state["t"] + 2.0 * state["dt"], t_end):
        if not do_step():
            raise RuntimeError("Evolving system could not approach final time")
    snap_count = 0
    while _fc_lt(state["t"], t_end):
        n = int(np.ceil((t_end - state["t"]) / state["dt"]))
        if n <= 0:
            raise ArithmeticError("Timestep doesn't make advances towards snap step")
        state["dt"] = (t_end - state["t"]) / n
        t_before, dt_try = state["t"], state["dt"]
        if not do_step():
            state["t"] = t_before
            if snap_count == 100:
                raise RuntimeError("Snapping time exceeded maximum iteration count")
            snap_count += 1
            state["dt"] = dt_try * 0.5
    return state["u"], state["t"], state["n"]


# ---------------------------------------------------------------------- reduce (L2 functional)
def _simplex_rule(dim, n=3):
    """Collapsed Gauss-Jacobi (Stroud conical product) rule, exact to degree 2n-1 on the unit simplex."""
    from scipy.special import roots_jacobi
    pts, wts = [np.zeros(0)], np.ones(1)
    out_p, out_w = np.zeros((1, 0)), np.ones(1)
    for k in range(dim):
        x, w = roots_jacobi(n, dim - 1 - k, 0)
        x, w = (x + 1) / 2, w / 2 ** (dim - k)
        P, W = [], []
        for p0, w0 in zip(out_p, out_w):
            for xi, wi in zip(x, w):
                P.append(np.concatenate([p0, [xi]]))
                W.append(w0 * wi)
        out_p, out_w = np.asarray(P), np.asarray(W)
    # map collapsed coordinates t -> simplex coordinates
    xi = np.zeros_like(out_p)
    rem = np.ones(out_p.shape[0])
    for k in range(dim):
        xi[:, k] = out_p[:, k] * rem
        rem = rem * (1 - out_p[:, k])
    return xi, out_w


def reduce_l2(model: Model, u, species: str, exact, time: float):
    """sum_T sum_q (u_h - exact(x,t))^2 * integration_factor, then sqrt (reduce.hh:157-203 with the
    u_error functional of test/gauss.ini:53-55)."""
    m = model.mesh
    g = model.names.index(species)
    sp = model.species[g]
    xi, w = _simplex_rule(m.dim, 3)
    phi = np.concatenate([1 - xi.sum(axis=1, keepdims=True), xi], axis=1)   # [nq, nd]
    sel = m.elem_comp == sp.comp
    X = m.coords[m.elems[sel]]                                   # [ne, nd, dim]
    B = X[:, 1:, :] - X[:, :1, :]
    det = np.abs(np.linalg.det(B))
    uh = u[m.elem_dof[sel] + sp.local] @ phi.T                   # [ne, nq]
    pos = np.einsum("qa,ead->eqd", phi, X)
    ex = exact(pos, time)
    return math.sqrt(float((((uh - ex) ** 2) * w[None, :] * det[:, None]).sum()))


# ---------------------------------------------------------------------- [model.reduce] (general)
def _reduce_rule(dim):
    """Quadrature of the reduce functionals in barycentric coordinates (weights sum to the reference
    volume).  dune-geometry's order-4 tables are third party and absent from the reference tree
    (parity unpinned): triangle = 6-point rule of degree 4 (Dunavant), tetrahedron = 15-point rule
    of degree 5 (Stroud T3:5-1) -- the same tables as kernels/reduce.cuh."""
    if dim == 2:
        lam, w = [], []
        for a, wt in ((0.445948490915965, 0.223381589678011), (0.091576213509771, 0.109951743655322)):
            for odd in range(3):
                lam.append([1 - 2 * a if k == odd else a for k in range(3)])
                w.append(0.5 * wt)
        return np.asarray(lam), np.asarray(w)
    s15 = math.sqrt(15.0)
    lam, w = [[0.25] * 4], [(16.0 / 135.0) / 6.0]
    for a, wt in (((7 - s15) / 34, (2665 + 14 * s15) / 37800), ((7 + s15) / 34, (2665 - 14 * s15) / 37800)):
        for odd in range(4):
            lam.append([1 - 3 * a if k == odd else a for k in range(4)])
            w.append(wt / 6.0)
    b = (10 - 2 * s15) / 40
    for i0 in range(4):
        for i1 in range(i0 + 1, 4):
            lam.append([b if k in (i0, i1) else 0.5 - b for k in range(4)])
            w.append((10.0 / 189.0) / 6.0)
    return np.asarray(lam), np.asarray(w)


def _function(text, ctx, nargs, what):
    head, body = text.split(":", 1)
    args = [a.strip() for a in head.split(",") if a.strip()]
    if len(args) != nargs:
        raise ValueError(f"{what} must have exactly {nargs} argument(s)")
    ast = E.resolve(E.Parser(body.strip()).parse(), ctx)
    return lambda *v: E.py_eval(ast, dict(zip(args, v)))


def reduce(model: Model, u, time: float, cfg: dict | None = None):
    """reduce.hh:38-285, sequential path: for every cell and quadrature point
    value = reduction(evaluation(), value) from `initial.value`; the gather step applies the
    reduction once more against `initial.value` (:205-210); then transformation / error / warn.
    Returns (values, status) with status 0 fine / 1 warn / 2 error."""
    rc = INI.sub(INI.sub(cfg if cfg is not None else model.cfg, "model"), "reduce")
    keys = [(k, v) for k, v in rc.items() if isinstance(v, dict)]
    m = model.mesh
    dim = m.dim
    X = m.coords[m.elems]                                        # [ne, nd, dim]
    sym = model.sym
    if m.etype == 1:
        # Q1 cells (not a reference element): the order-4 rule of a cube is the 3-point Gauss rule per
        # axis; lam[q, a] = multilinear shape function a at point q, dlam its reference gradient
        g1 = np.array([0.5 - math.sqrt(0.15), 0.5, 0.5 + math.sqrt(0.15)])
        w1 = np.array([5.0, 8.0, 5.0]) / 18.0
        pts = np.stack([a.ravel() for a in np.meshgrid(*[g1] * dim, indexing="ij")], axis=1)[:, ::-1]   # x fastest
        w = np.prod(np.stack([a.ravel() for a in np.meshgrid(*[w1] * dim, indexing="ij")], axis=1), axis=1)
        nd, nq = 1 << dim, w.size
        lam = np.ones((nq, nd))
        dlam = np.ones((nq, nd, dim))
        for a in range(nd):
            for k in range(dim):
                f = pts[:, k] if (a >> k) & 1 else 1.0 - pts[:, k]
                lam[:, a] *= f
                for r in range(dim):
                    dlam[:, a, r] *= (1.0 if (a >> k) & 1 else -1.0) if r == k else f
        h = np.stack([X[:, 1 << k, k] - X[:, 0, k] for k in range(dim)], axis=1)       # [ne, dim]
        det = np.prod(h, axis=1)
        entvol = det
    else:
        nd = dim + 1
        lam, w = _reduce_rule(dim)
        nq = w.size
        Bm = np.transpose(X[:, 1:, :] - X[:, :1, :], (0, 2, 1))      # columns = edges
        det = np.abs(np.linalg.det(Bm))
        Binv = np.linalg.inv(Bm)                                     # rows = grad of xi_k
        G = np.concatenate([-Binv.sum(axis=1, keepdims=True), Binv], axis=1)   # [ne, nd, dim]
        entvol = det / (2.0 if dim == 2 else 6.0)
    ctx = np.zeros((m.ne, nq, sym.nslots))
    ctx[..., E.SLOT_TIME] = time
    ctx[..., E.SLOT_INTFAC] = w[None, :] * det[:, None]
    ctx[..., E.SLOT_ENTVOL] = entvol[:, None]
    ctx[..., E.SLOT_INVOL] = 1.0
    ctx[..., E.SLOT_POS:E.SLOT_POS + dim] = np.einsum("qa,ead->eqd", lam, X)
    for k in range(len(m.cell_keys)):
        ctx[..., E.SLOT_CELL + k] = m.cell_data[k][:, None]
    for g, sp in enumerate(model.species):
        sel = m.elem_comp == sp.comp
        xl = u[m.elem_dof[sel] + sp.local]                       # [nsel, nd]
        base = sym.spec_base + 4 * g
        ctx[sel, :, base] = xl @ lam.T
        if m.etype == 1:
            ctx[sel, :, base + 1:base + 1 + dim] = np.einsum("ea,qad->eqd", xl, dlam) / h[sel][:, None, :]
        else:
            ctx[sel, :, base + 1:base + 1 + dim] = np.einsum("ea,ead->ed", xl, G[sel])[:, None, :]
    rows = ctx.reshape(-1, sym.nslots)
    values, status = {}, {}
    for key, sub in keys:
        init = float(INI.get(sub, "initial.value", 0.0))
        ev = INI.get(sub, "evaluation.expression", None)
        if ev is None or E.is_absent(ev):
            vals = np.zeros(rows.shape[0])
        else:
            code, consts = E.compile_expr(ev, sym, model.ctx)
            vals = eval_program(code, consts, rows)
        red = INI.get(sub, "reduction.expression", None)
        if red is None:
            total = init + float(vals.sum())
            total = total + init
        else:
            op = _function(red, model.ctx, 2, "Reduction arguments")
            total = init
            for v in vals:
                total = op(float(v), total)
            total = op(total, init)
        tr = INI.get(sub, "transformation.expression", None)
        if tr is not None:
            total = _function(tr, model.ctx, 1, "Warning function")(total)
        st = 0
        nz = lambda v: abs(v) > 1e-8 * max(1.0, abs(v))   # noqa: E731  FloatCmp::ne(v, 0)
        er = INI.get(sub, "error.expression", None)
        wa = INI.get(sub, "warn.expression", None)
        if er is not None and nz(_function(er, model.ctx, 1, "Error function")(total)):
            st = 2
        elif wa is not None and nz(_function(wa, model.ctx, 1, "Warning function")(total)):
            st = 1
        values[key], status[key] = total, st
    return values, status
