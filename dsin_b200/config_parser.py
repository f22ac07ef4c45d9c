"""Mini parser for the reference's run-config files.

The reference reads ``run_configs/ae_run_configs`` and ``run_configs/pc_run_configs``
through ``fjcommon.config_parser.parse`` (/root/reference/src/main.py:184-185), which is
not installed here (fjcommon 0.1.69, pinned in requirements.txt:11).  The grammar that
those two files actually use is small and is restated here:

  * ``# ...``                      comment (also trailing)
  * ``key = <python expression>``  e.g. ``H_target = 2*0.02``, ``crop_size = (320,1224)``
  * ``constrain key :: A, B, C``   declares bare words usable unquoted as values of
                                   ``key`` (``normalization = FIXED``); they evaluate to
                                   their own name as a string, and assigning a value
                                   outside the set raises ``ValueError``.

``parse(path)`` returns ``(config, rel_path)`` like the fjcommon call it replaces.
"""
from __future__ import annotations

import ast
import operator
import os

# The value grammar the two run-config files use: literals, tuples / lists, earlier keys and constraint words as names,
# and + - * / ** with unary minus.  Evaluated over the AST (never eval()): a config file cannot reach attributes, calls,
# subscripts or comprehensions.
_BINOPS = {ast.Add: operator.add, ast.Sub: operator.sub, ast.Mult: operator.mul, ast.Div: operator.truediv,
           ast.FloorDiv: operator.floordiv, ast.Mod: operator.mod, ast.Pow: operator.pow}


def _evaluate(node, env):
    if isinstance(node, ast.Expression):
        return _evaluate(node.body, env)
    if isinstance(node, ast.Constant):
        return node.value
    if isinstance(node, ast.Name):
        if node.id in env:
            return env[node.id]
        raise NameError("unknown name {!r}".format(node.id))
    if isinstance(node, ast.Tuple):
        return tuple(_evaluate(e, env) for e in node.elts)
    if isinstance(node, ast.List):
        return [_evaluate(e, env) for e in node.elts]
    if isinstance(node, ast.UnaryOp) and isinstance(node.op, (ast.USub, ast.UAdd)):
        v = _evaluate(node.operand, env)
        return -v if isinstance(node.op, ast.USub) else +v
    if isinstance(node, ast.BinOp) and type(node.op) in _BINOPS:
        a, b = _evaluate(node.left, env), _evaluate(node.right, env)
        if isinstance(node.op, ast.Pow) and isinstance(b, (int, float)) and abs(b) > 64:
            raise ValueError("exponent too large")
        return _BINOPS[type(node.op)](a, b)
    raise ValueError("unsupported expression element {}".format(type(node).__name__))


class Config(object):
    """Attribute bag (what AE.__init__ reads: /root/reference/src/AE.py:13-29)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)

    def all_params_and_values(self):
        return sorted((k, v) for k, v in self.__dict__.items() if not k.startswith("_"))

    def __str__(self):
        return "\n".join("{} = {!r}".format(k, v) for k, v in self.all_params_and_values())

    __repr__ = __str__


def _strip_comment(line):
    out, quote = [], None
    for ch in line:
        if quote:
            out.append(ch)
            if ch == quote:
                quote = None
        elif ch in "'\"":
            quote = ch
            out.append(ch)
        elif ch == "#":
            break
        else:
            out.append(ch)
    return "".join(out).strip()


def parse_string(text, base=None):
    values = dict(base.__dict__) if base is not None else {}
    constraints = {}
    words = {}
    for lineno, raw in enumerate(text.splitlines(), 1):
        line = _strip_comment(raw)
        if not line:
            continue
        if line.startswith("constrain "):
            body = line[len("constrain "):]
            if "::" not in body:
                raise ValueError("line {}: expected 'constrain key :: A, B'".format(lineno))
            key, opts = body.split("::", 1)
            opts = [o.strip() for o in opts.split(",") if o.strip()]
            constraints[key.strip()] = opts
            for o in opts:
                words[o] = o
            continue
        if "=" not in line:
            raise ValueError("line {}: expected 'key = value', got {!r}".format(lineno, raw))
        key, expr = line.split("=", 1)
        key, expr = key.strip(), expr.strip()
        if not key.isidentifier():
            raise ValueError("line {}: bad key {!r}".format(lineno, key))
        env = dict(words)
        env.update(values)
        try:
            val = _evaluate(ast.parse(expr, mode="eval"), env)
        except Exception as e:  # noqa: BLE001
            raise ValueError("line {}: cannot evaluate {!r}: {}".format(lineno, expr, e))
        if key in constraints and val not in constraints[key]:
            raise ValueError("line {}: {} = {!r} violates constraint {}".format(
                lineno, key, val, constraints[key]))
        values[key] = val
    return Config(**values)


def parse(config_path):
    with open(config_path, "r") as f:
        cfg = parse_string(f.read())
    return cfg, os.path.basename(config_path)
