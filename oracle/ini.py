"""ORACLE (test infrastructure) -- Dune::ParameterTree-style INI reader.

Format as consumed by the reference CLI (src/dune_copasi.cc:270-282, dune-common
ParameterTreeParser): ``[a.b]`` section headers prefix the keys that follow, keys may contain dots
to address sub-trees, ``#`` starts a comment, values are strings.  Sub-key order is insertion order
(compartment ids and species order depend on it: grid/make_multi_domain_grid.hh:118-124).
"""
from __future__ import annotations


def parse_ini(text: str) -> dict:
    root: dict = {}
    prefix: list[str] = []
    for raw in text.splitlines():
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        if line.startswith("[") and line.endswith("]"):
            name = line[1:-1].strip()
            prefix = [p for p in name.split(".") if p] if name else []
            continue
        if "=" not in line:
            continue
        key, val = line.split("=", 1)
        path = prefix + [p for p in key.strip().split(".") if p]
        set_key(root, path, val.strip())
    return root


def set_key(root: dict, path, val: str):
    if isinstance(path, str):
        path = path.split(".")
    node = root
    for p in path[:-1]:
        nxt = node.get(p)
        if not isinstance(nxt, dict):
            nxt = {} if nxt is None else {"": nxt}
            node[p] = nxt
        node = nxt
    node[path[-1]] = val


def get(root: dict, key: str, default=None):
    node = root
    for p in key.split("."):
        if not isinstance(node, dict) or p not in node:
            return default
        node = node[p]
    return node


def sub(root: dict, key: str) -> dict:
    v = get(root, key, {})
    return v if isinstance(v, dict) else {}


def to_text(root: dict) -> str:
    """Flatten back to ``a.b.c = v`` lines (what the product's C-ABI config loader accepts too)."""
    out = []

    def rec(node, pre):
        for k, v in node.items():
            if isinstance(v, dict):
                rec(v, pre + [k])
            else:
                out.append(".".join(pre + [k]) + " = " + str(v))
    rec(root, [])
    return "\n".join(out) + "\n"
