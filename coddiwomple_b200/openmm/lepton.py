"""A small evaluator for the Lepton expressions the reference writes its alchemical protocols in
(`openmm_pdf_update_functions` of OMMLIAIS, coddiwomple/openmm/integrators.py:483-500; examples in
coddiwomple/tests/utils.py:63-67,100-110: "select(step(fractional_iteration - 0.5), 1.0, 2.0 * fractional_iteration)").

Host-side only: the protocol is lowered once to a dense [n_steps + 1][n_lambda] table that the kernels index.
Grammar: numbers, identifiers, + - * / ^, unary minus, parentheses, and the Lepton functions
select, step, delta, min, max, abs, sqrt, exp, log, sin, cos.
"""
import math
import re

_TOKEN = re.compile(r"\s*(?:(\d+\.?\d*(?:[eE][-+]?\d+)?|\.\d+(?:[eE][-+]?\d+)?)|([A-Za-z_]\w*)|(.))")

_FUNCTIONS = {
    "select": lambda a, b, c: b if a != 0 else c,
    "step": lambda x: 1.0 if x >= 0 else 0.0,
    "delta": lambda x: 1.0 if x == 0 else 0.0,
    "min": min, "max": max, "abs": abs, "sqrt": math.sqrt, "exp": math.exp, "log": math.log, "sin": math.sin, "cos": math.cos,
}


class LeptonError(ValueError):
    pass


def _tokens(text):
    pos, out = 0, []
    text = text.strip().rstrip(";")
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m or m.end() == pos:
            raise LeptonError(f"cannot tokenise {text[pos:]!r}")
        pos = m.end()
        num, name, op = m.groups()
        out.append(("num", float(num)) if num is not None else (("name", name) if name is not None else ("op", op)))
    out.append(("end", None))
    return out


class _Parser:
    def __init__(self, text):
        self.toks, self.i = _tokens(text), 0

    def peek(self):
        return self.toks[self.i]

    def take(self, kind=None, value=None):
        t = self.toks[self.i]
        if (kind and t[0] != kind) or (value is not None and t[1] != value):
            raise LeptonError(f"expected {value or kind}, found {t[1]!r}")
        self.i += 1
        return t

    def expr(self):  # + -
        node = self.term()
        while self.peek() in (("op", "+"), ("op", "-")):
            op = self.take()[1]
            rhs = self.term()
            node = ("add" if op == "+" else "sub", node, rhs)
        return node

    def term(self):  # * /
        node = self.unary()
        while self.peek() in (("op", "*"), ("op", "/")):
            op = self.take()[1]
            rhs = self.unary()
            node = ("mul" if op == "*" else "div", node, rhs)
        return node

    def unary(self):
        if self.peek() == ("op", "-"):
            self.take()
            return ("neg", self.unary())
        if self.peek() == ("op", "+"):
            self.take()
            return self.unary()
        return self.power()

    def power(self):  # right associative, binds tighter than unary minus on its left operand
        base = self.atom()
        if self.peek() == ("op", "^"):
            self.take()
            return ("pow", base, self.unary())
        return base

    def atom(self):
        kind, val = self.peek()
        if kind == "num":
            self.take()
            return ("num", val)
        if kind == "name":
            self.take()
            if self.peek() == ("op", "("):
                self.take()
                args = [self.expr()]
                while self.peek() == ("op", ","):
                    self.take()
                    args.append(self.expr())
                self.take("op", ")")
                if val not in _FUNCTIONS:
                    raise LeptonError(f"unknown function {val!r}")
                return ("call", val, args)
            return ("var", val)
        if (kind, val) == ("op", "("):
            self.take()
            node = self.expr()
            self.take("op", ")")
            return node
        raise LeptonError(f"unexpected {val!r}")


def _eval(node, env):
    k = node[0]
    if k == "num":
        return node[1]
    if k == "var":
        if node[1] not in env:
            raise LeptonError(f"unknown variable {node[1]!r}")
        return float(env[node[1]])
    if k == "neg":
        return -_eval(node[1], env)
    if k == "call":
        return float(_FUNCTIONS[node[1]](*[_eval(a, env) for a in node[2]]))
    a, b = _eval(node[1], env), _eval(node[2], env)
    return {"add": a + b, "sub": a - b, "mul": a * b, "div": a / b if k == "div" else 0.0, "pow": a ** b if k == "pow" else 0.0}[k]


def compile_expression(text):
    """Returns f(**variables) -> float for a Lepton expression string."""
    p = _Parser(text)
    tree = p.expr()
    p.take("end")
    return lambda **env: _eval(tree, env)


def protocol_table(functions, n_steps, variable="fractional_iteration"):
    """[{name: value}] for fractional_iteration = k / n_steps, k = 1 .. n_steps (the values the 'H' step of OMMLIAIS
    assigns after the 'I' step has advanced the iteration, integrators.py:548-590)."""
    compiled = {name: compile_expression(expr) for name, expr in functions.items()}
    return [{name: f(**{variable: k / float(n_steps)}) for name, f in compiled.items()} for k in range(1, n_steps + 1)]
