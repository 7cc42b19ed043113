"""ORACLE (test infrastructure, not product code) -- expression compiler.

Restates the *semantics* of the reference's run-time math expressions so the CPU
oracle can evaluate them per quadrature point the way the reference does
(interpreted, one call per functor per point):

* reference evaluation chain: dune/copasi/model/local_equations.hh:101-146 ->
  dune/copasi/model/functor_factory_parser.impl.hh:116-182 -> parser back-ends
  src/dune/copasi/parser/{exprtk,mu,symengine}.cc (third party: ExprTk 0.0.3,
  muParser, SymEngine -- none vendored in /root/reference; their published
  grammars are restated here for the subset the reference's inis use,
  SURVEY.md App. D).
* symbol table: functor_factory_parser.impl.hh:133-169 (time, integration_factor,
  entity_volume, in_volume, in_boundary, in_skeleton, no_value, position_{x,y,z},
  normal_{x,y,z}, cell-data keys, species names, grad_<species>_{x,y,z}).
* literal fast path / "zero removes the term": impl.hh:122-128.
* parser context (constants, inline functions): src/dune/copasi/parser/context.cc:56-97.

The expression is parsed to an AST (precedence climbing), user functions and
constants are inlined, and the AST is emitted as RPN byte-code that
``oracle.c:orc_eval`` runs on a small stack (muParser itself is a byte-code
interpreter: src/dune/copasi/parser/mu.cc:214-219).

Deliberately written independently from the product's C++ front-end
(dune_copasi_b200/csrc/expr.cpp): different language, different algorithm
(shunting by precedence-climbing to RPN vs recursive descent to CUDA text).
"""
from __future__ import annotations

import math
import re
from dataclasses import dataclass, field

DBL_MAX = 1.7976931348623157e308

# ---- VM op-codes (must match oracle.c) ---------------------------------------------------------
OP_CONST, OP_VAR = 0, 1
OP_ADD, OP_SUB, OP_MUL, OP_DIV, OP_POW, OP_NEG = 2, 3, 4, 5, 6, 7
OP_LT, OP_GT, OP_LE, OP_GE, OP_EQ, OP_NE = 8, 9, 10, 11, 12, 13
OP_AND, OP_OR, OP_NOT, OP_SEL = 14, 15, 16, 17
OP_F1, OP_F2, OP_MOD = 18, 19, 20
OP_TAB = 21          # tabulated context function: arg = offset of its record in the constants
OP_TAB2 = 22         # image function of two arguments (tiff): arg = offset of its record

F1 = {"sqrt": 0, "exp": 1, "log": 2, "sin": 3, "cos": 4, "tan": 5, "abs": 6, "floor": 7,
      "ceil": 8, "tanh": 9, "sinh": 10, "cosh": 11, "asin": 12, "acos": 13, "atan": 14,
      "log10": 15, "log2": 16, "sgn": 17, "ln": 2, "sign": 17, "exp2": 18, "round": 19}
F2 = {"min": 0, "max": 1, "atan2": 2, "pow": 3}

# ---- evaluation-context slots (must match oracle.c) -------------------------------------------
SLOT_TIME, SLOT_INTFAC, SLOT_ENTVOL, SLOT_INVOL, SLOT_INBND, SLOT_INSKEL = 0, 1, 2, 3, 4, 5
SLOT_POS, SLOT_NORMAL, SLOT_CELL = 6, 9, 12
AXES = "xyz"

_zero_re = re.compile(r"^\s*[+-]?\s*(0+\.?0*|\.0+)([eE][+-]?\d+)?\s*$")
_num_re = re.compile(r"^\s*[+-]?(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?\s*$")
_tok_re = re.compile(r"""
    (?P<num>(\d+\.?\d*|\.\d+)([eE][+-]?\d+)?)
  | (?P<id>[A-Za-z_][A-Za-z_0-9]*)
  | (?P<op>\*\*|<=|>=|==|!=|&&|\|\||[-+*/^%<>()?:,&|!])
  | (?P<ws>\s+)
""", re.X)


class ExprError(ValueError):
    pass


def is_absent(expr: str | None) -> bool:
    """Empty or literally-zero expression => the term does not exist (impl.hh:122-124)."""
    return expr is None or expr.strip() == "" or bool(_zero_re.match(expr))


def tokenize(s: str):
    pos, out = 0, []
    while pos < len(s):
        m = _tok_re.match(s, pos)
        if not m:
            raise ExprError(f"bad character {s[pos]!r} at {pos} in {s!r}")
        pos = m.end()
        if m.lastgroup == "ws":
            continue
        out.append((m.lastgroup, m.group(m.lastgroup)))
    out.append(("end", ""))
    return out


def lerp_std(a: float, b: float, t: float) -> float:
    """std::lerp as libstdc++ implements it (the reference calls std::lerp, context.cc:93, :262)."""
    if (a <= 0 and b >= 0) or (a >= 0 and b <= 0):
        return t * b + (1 - t) * a
    if t == 1:
        return b
    x = a + t * (b - a)
    return max(b, x) if (t > 1) == (b > a) else min(b, x)


@dataclass
class Table:
    """Tabulated one-argument context function.
    kind 0: `type = interpolation` (context.cc:72-97): lower_bound on the sorted domain, the end
            values outside, std::lerp inside.
    kind 1: `type = function` with `interpolate = true` (context.cc:237-283), restated literally:
            the sample index i = max(0, k) and j = min(k, intervals + 1) come from the same interval
            number k, so lerp(g[i], g[j], t) = g[k]; `clamp` is std::clamp(domain[0], pos, domain[1])
            with the arguments in that order, i.e. max(domain[0], pos); `error` raises (NaN in the C VM)."""
    kind: int
    domain: list
    range: list
    clamp: bool = False
    name: str = ""

    def __call__(self, x: float) -> float:
        import bisect
        if self.kind == 0:
            d = bisect.bisect_left(self.domain, x)
            if d == 0:
                return self.range[0]
            if d == len(self.domain):
                return self.range[-1]
            return lerp_std(self.range[d - 1], self.range[d],
                            (x - self.domain[d - 1]) / (self.domain[d] - self.domain[d - 1]))
        d0, d1 = self.domain
        n = len(self.range) - 1
        if self.clamp:
            x = x if d0 < x else d0
        elif x < d0 or x > d1 or x != x:
            raise ExprError(f"interpolation of function {self.name} is out of bounds: {x} not in [{d0}, {d1}]")
        whole = math.modf((x - d0) * (n / (d1 - d0)))[1]
        k = 0 if whole <= 0 else (n if whole >= n else int(whole))
        return self.range[k]

    def record(self) -> list:
        """what the C VM reads at the OP_TAB offset: kind, samples, clamp, domain, range"""
        return [float(self.kind), float(len(self.range)), float(self.clamp)] + list(self.domain) + list(self.range)


@dataclass
class Context:
    """parser_context: constants, inline functions and tabulated functions (context.cc:56-97, 237-283)."""
    constants: dict = field(default_factory=dict)          # name -> float
    functions: dict = field(default_factory=dict)          # name -> (argnames, body string)
    tables: dict = field(default_factory=dict)             # name -> Table

    @staticmethod
    def from_config(cfg: dict) -> "Context":
        """cfg: nested dict of the [parser_context] section."""
        ctx = Context()
        for name, sub in cfg.items():
            if not isinstance(sub, dict):
                continue
            typ = sub.get("type")
            if typ == "constant":
                ctx.constants[name] = float(sub["value"])
            elif typ == "function":
                head, body = sub["expression"].split(":", 1)
                args = [a.strip() for a in head.split(",") if a.strip()]
                ctx.functions[name] = (args, body.strip())
            elif typ == "tiff":
                from . import tiff as TIFF
                ctx.tables[name] = TIFF.read(str(sub["path"]))
            elif typ == "interpolation":
                dom = [float(v) for v in str(sub["domain"]).split()]
                rng = [float(v) for v in str(sub["range"]).split()]
                if dom != sorted(dom):
                    raise ExprError("The interpolation domain must be sorted")
                if len(dom) < 2 or len(dom) != len(rng):
                    raise ExprError("Interpolation range and domain must have at least two points and be the same size")
                ctx.tables[name] = Table(0, dom, rng, name=name)
        truthy = ("1", "true", "yes", "on")
        for name, sub in cfg.items():
            if not isinstance(sub, dict) or sub.get("type") != "function":
                continue
            if str(sub.get("interpolate", "false")).strip().lower() not in truthy:
                continue
            args, body = ctx.functions[name]
            if len(args) != 1:
                raise ExprError(f"Cannot interpolate '{name}' function with {len(args)} arguments")
            isub = sub.get("interpolation", {}) if isinstance(sub.get("interpolation", {}), dict) else {}
            n = int(isub.get("intervals", 1000))
            if n > 1e5:
                raise ExprError("Number of interpolation intervals is too big!")
            if n < 1:
                raise ExprError(f"At least one interval is required in function {name}")
            dsub = isub.get("domain", {})
            dom = [float(v) for v in str(dsub.get(args[0], "0 1")).split()] if isinstance(dsub, dict) else [0.0, 1.0]
            if len(dom) != 2 or dom[0] >= dom[1]:
                raise ExprError(f"Domain arguments of function {name} are not ordered")
            ooo = str(isub.get("out_of_bounds", "error"))
            if ooo not in ("clamp", "error"):
                raise ExprError(f"Not known '{name}.interpolation.out_of_bounds = {ooo}'")
            plain = Context(ctx.constants, {k: v for k, v in ctx.functions.items()}, {})
            ast = resolve(Parser(body).parse(), plain)
            width = (dom[1] - dom[0]) / float(n)
            rng = [py_eval(ast, {args[0]: dom[0] + float(i) * width}) for i in range(n + 1)]
            ctx.tables[name] = Table(1, dom, rng, clamp=(ooo == "clamp"), name=name)
        return ctx


class Parser:
    """Precedence climbing. Result AST: ('num',v) ('var',name) ('un',op,a) ('bin',op,a,b)
    ('sel',c,a,b) ('call',name,[args])."""

    BIN = [  # (level, ops) low -> high
        ({"or", "||", "|"}),
        ({"and", "&&", "&"}),
        ({"==", "!="}),
        ({"<", ">", "<=", ">="}),
        ({"+", "-"}),
        ({"*", "/", "%"}),
    ]

    def __init__(self, text: str):
        self.toks = tokenize(text)
        self.i = 0
        self.text = text

    def peek(self):
        return self.toks[self.i]

    def take(self):
        t = self.toks[self.i]
        self.i += 1
        return t

    def accept(self, val):
        k, v = self.peek()
        if v == val and k in ("op", "id"):
            self.i += 1
            return True
        return False

    def expect(self, val):
        if not self.accept(val):
            raise ExprError(f"expected {val!r} near token {self.i} in {self.text!r}")

    def parse(self):
        e = self.ternary()
        if self.peek()[0] != "end":
            raise ExprError(f"trailing input {self.peek()[1]!r} in {self.text!r}")
        return e

    def ternary(self):
        c = self.binary(0)
        if self.accept("?"):
            a = self.ternary()
            self.expect(":")
            b = self.ternary()
            return ("sel", c, a, b)
        return c

    def binary(self, lvl):
        if lvl == len(self.BIN):
            return self.unary()
        a = self.binary(lvl + 1)
        while True:
            k, v = self.peek()
            if k in ("op", "id") and v in self.BIN[lvl]:
                self.i += 1
                b = self.binary(lvl + 1)
                a = ("bin", _canon(v), a, b)
            else:
                return a

    def unary(self):
        if self.accept("-"):
            return ("un", "neg", self.unary())
        if self.accept("+"):
            return self.unary()
        if self.accept("!") or self.accept("not"):
            return ("un", "not", self.unary())
        return self.power()

    def power(self):
        base = self.atom()
        if self.accept("^") or self.accept("**"):
            # right associative; exponent may carry its own sign:  2^-x^2 == 2^(-(x^2))
            expo = self.unary()
            return ("bin", "^", base, expo)
        return base

    def atom(self):
        k, v = self.take()
        if k == "num":
            return ("num", float(v))
        if k == "id":
            if self.accept("("):
                args = []
                if not self.accept(")"):
                    while True:
                        args.append(self.ternary())
                        if self.accept(")"):
                            break
                        self.expect(",")
                return ("call", v, args)
            return ("var", v)
        if k == "op" and v == "(":
            e = self.ternary()
            self.expect(")")
            return e
        raise ExprError(f"unexpected token {v!r} in {self.text!r}")


def _canon(op):
    return {"||": "or", "|": "or", "&&": "and", "&": "and"}.get(op, op)


def _subst(ast, env):
    k = ast[0]
    if k == "num":
        return ast
    if k == "var":
        return env.get(ast[1], ast)
    if k == "un":
        return ("un", ast[1], _subst(ast[2], env))
    if k == "bin":
        return ("bin", ast[1], _subst(ast[2], env), _subst(ast[3], env))
    if k == "sel":
        return ("sel", _subst(ast[1], env), _subst(ast[2], env), _subst(ast[3], env))
    if k == "call":
        return ("call", ast[1], [_subst(a, env) for a in ast[2]])
    if k == "tab":
        return ("tab", ast[1], _subst(ast[2], env))
    if k == "tab2":
        return ("tab2", ast[1], _subst(ast[2], env), _subst(ast[3], env))
    raise AssertionError(k)


def resolve(ast, ctx: Context, depth=0):
    """Inline context constants and functions (function bodies may use constants)."""
    if depth > 16:
        raise ExprError("context functions nested too deep (recursion?)")
    k = ast[0]
    if k == "num":
        return ast
    if k == "var":
        name = ast[1]
        if name in ctx.constants:
            return ("num", float(ctx.constants[name]))
        if name == "no_value":
            return ("num", DBL_MAX)
        if name == "pi":
            return ("num", math.pi)
        return ast
    if k == "un":
        return ("un", ast[1], resolve(ast[2], ctx, depth))
    if k == "bin":
        return ("bin", ast[1], resolve(ast[2], ctx, depth), resolve(ast[3], ctx, depth))
    if k == "sel":
        return ("sel",) + tuple(resolve(a, ctx, depth) for a in ast[1:])
    if k == "call":
        name, args = ast[1], [resolve(a, ctx, depth) for a in ast[2]]
        if name in ctx.tables and not isinstance(ctx.tables[name], Table):      # image: two arguments
            if len(args) != 2:
                raise ExprError(f"function {name} expects 2 arguments, got {len(args)}")
            return ("tab2", ctx.tables[name], args[0], args[1])
        if name in ctx.tables:
            if len(args) != 1:
                raise ExprError(f"function {name} expects 1 argument, got {len(args)}")
            return ("tab", ctx.tables[name], args[0])
        if name in ctx.functions:
            argn, body = ctx.functions[name]
            if len(argn) != len(args):
                raise ExprError(f"function {name} expects {len(argn)} args, got {len(args)}")
            b = Parser(body).parse()
            b = _subst(b, dict(zip(argn, args)))
            return resolve(b, ctx, depth + 1)
        if name == "if" and len(args) == 3:
            return ("sel", args[0], args[1], args[2])
        return ("call", name, args)
    if k == "tab":
        return ("tab", ast[1], resolve(ast[2], ctx, depth))
    if k == "tab2":
        return ("tab2", ast[1], resolve(ast[2], ctx, depth), resolve(ast[3], ctx, depth))
    raise AssertionError(k)


class Symbols:
    """Maps names to context slots for one model (species across all compartments)."""

    def __init__(self, dim: int, species: list[str], cell_keys: list[str] = ()):  # noqa
        self.dim = dim
        self.species = list(species)
        self.cell_keys = list(cell_keys)
        self.spec_base = SLOT_CELL + len(self.cell_keys)
        self.nslots = self.spec_base + 4 * len(self.species)
        self.table = {
            "time": SLOT_TIME, "integration_factor": SLOT_INTFAC, "entity_volume": SLOT_ENTVOL,
            "in_volume": SLOT_INVOL, "in_boundary": SLOT_INBND, "in_skeleton": SLOT_INSKEL,
        }
        for a, ax in enumerate(AXES):
            self.table[f"position_{ax}"] = SLOT_POS + a      # axes >= dim hold 0 (impl.hh:148-152)
            self.table[f"normal_{ax}"] = SLOT_NORMAL + a
        for j, key in enumerate(self.cell_keys):
            self.table[key] = SLOT_CELL + j
        for g, name in enumerate(self.species):
            self.table[name] = self.spec_base + 4 * g
            for a, ax in enumerate(AXES):
                self.table[f"grad_{name}_{ax}"] = self.spec_base + 4 * g + 1 + a

    def value_slot(self, g):
        return self.spec_base + 4 * g


def emit(ast, sym: Symbols, code: list, consts: list):
    k = ast[0]
    if k == "num":
        code += [OP_CONST, len(consts)]
        consts.append(ast[1])
    elif k == "var":
        if ast[1] not in sym.table:
            raise ExprError(f"unknown symbol {ast[1]!r}")
        code += [OP_VAR, sym.table[ast[1]]]
    elif k == "un":
        emit(ast[2], sym, code, consts)
        code += [OP_NEG if ast[1] == "neg" else OP_NOT, 0]
    elif k == "bin":
        emit(ast[2], sym, code, consts)
        emit(ast[3], sym, code, consts)
        op = {"+": OP_ADD, "-": OP_SUB, "*": OP_MUL, "/": OP_DIV, "^": OP_POW, "%": OP_MOD,
              "<": OP_LT, ">": OP_GT, "<=": OP_LE, ">=": OP_GE, "==": OP_EQ, "!=": OP_NE,
              "and": OP_AND, "or": OP_OR}[ast[1]]
        code += [op, 0]
    elif k == "sel":
        for a in ast[1:]:
            emit(a, sym, code, consts)
        code += [OP_SEL, 0]
    elif k == "call":
        name, args = ast[1], ast[2]
        if name in F1 and len(args) == 1:
            emit(args[0], sym, code, consts)
            code += [OP_F1, F1[name]]
        elif name in F2 and len(args) >= 2:
            emit(args[0], sym, code, consts)
            for a in args[1:]:          # min/max are variadic in ExprTk
                emit(a, sym, code, consts)
                code += [OP_F2, F2[name]]
        else:
            raise ExprError(f"unknown function {name!r}/{len(args)}")
    elif k == "tab":
        emit(ast[2], sym, code, consts)
        code += [OP_TAB, len(consts)]
        consts.extend(ast[1].record())
    elif k == "tab2":
        emit(ast[2], sym, code, consts)
        emit(ast[3], sym, code, consts)
        code += [OP_TAB2, len(consts)]
        consts.extend(ast[1].record())
    else:
        raise AssertionError(k)


def compile_expr(text: str, sym: Symbols, ctx: Context | None = None):
    """-> (code int list, consts float list).  A plain number never reaches a parser in the
    reference (impl.hh:125-128); here it simply compiles to one CONST op."""
    ctx = ctx or Context()
    ast = resolve(Parser(text).parse(), ctx)
    code, consts = [], []
    emit(ast, sym, code, consts)
    return code, consts


# ---- pure-python evaluator (small cases, cross-checks the C VM) --------------------------------
def py_eval(ast, env: dict):
    k = ast[0]
    if k == "num":
        return ast[1]
    if k == "var":
        return float(env[ast[1]])
    if k == "un":
        a = py_eval(ast[2], env)
        return -a if ast[1] == "neg" else float(a == 0.0)
    if k == "bin":
        a, b = py_eval(ast[2], env), py_eval(ast[3], env)
        op = ast[1]
        if op == "+": return a + b
        if op == "-": return a - b
        if op == "*": return a * b
        if op == "/": return a / b if b != 0 else math.copysign(math.inf, a) if a != 0 else math.nan
        if op == "^": return math.pow(a, b)
        if op == "%": return math.fmod(a, b)
        if op == "<": return float(a < b)
        if op == ">": return float(a > b)
        if op == "<=": return float(a <= b)
        if op == ">=": return float(a >= b)
        if op == "==": return float(a == b)
        if op == "!=": return float(a != b)
        if op == "and": return float(a != 0 and b != 0)
        if op == "or": return float(a != 0 or b != 0)
    if k == "sel":
        return py_eval(ast[2], env) if py_eval(ast[1], env) != 0 else py_eval(ast[3], env)
    if k == "call":
        args = [py_eval(a, env) for a in ast[2]]
        n = ast[1]
        if n in ("min", "max"):
            return min(args) if n == "min" else max(args)
        if n == "atan2": return math.atan2(*args)
        if n == "pow": return math.pow(*args)
        if n in ("ln", "log"): return math.log(args[0])
        if n == "abs": return abs(args[0])
        if n in ("sgn", "sign"): return float((args[0] > 0) - (args[0] < 0))
        return getattr(math, n)(args[0])
    if k == "tab":
        return ast[1](py_eval(ast[2], env))
    if k == "tab2":
        return ast[1](py_eval(ast[2], env), py_eval(ast[3], env))
    raise AssertionError(k)
