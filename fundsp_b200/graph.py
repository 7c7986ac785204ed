"""Graph expressions: the host-side mirror of the reference's `An<X>` wrapper and its operators.

Mirrors `src/combinator.rs:178-488` of SamiPerttu/fundsp v0.23.0: `>>` (Pipe), `|` (Stack), `&` (Bus),
`^` (Branch), `+ - *` (Binop, or Unop with an f32 scalar), unary `-` (Unop neg) and `!` (Thru; spelled `~`
in Python), plus the builder methods `.phase()`, `.seed()` (`combinator.rs:263-276`).

An `An` is an immutable *description* (op name, arguments, children).  It owns no DSP state: lowering
(`An.lower(backend)`) walks the tree once and emits the backend's builder calls in depth-first,
left-to-right order, which is exactly the order the reference constructs (and pings) its nodes in.
The GPU backend is `fundsp_b200.capi` (the C-ABI of libfundsp_b200.so).
"""
from __future__ import annotations

import numpy as np

# Binop / Unop kinds (src/audionode.rs:724-1240)
OP_ADD, OP_SUB, OP_MUL = 0, 1, 2
U_NEG, U_ADD, U_NEGADD, U_MUL = 0, 1, 2, 3
# Multi-node kinds = reference node IDs (src/audionode.rs:2103,2265,2414,2576,2714)
M_BUS, M_STACK, M_REDUCE, M_BRANCH, M_CHAIN = 28, 30, 31, 33, 32


def f32(x) -> float:
    """Round a Python float to f32 the way a Rust `as f32` / f32 literal does."""
    return float(np.float32(x))


class An:
    __slots__ = ("op", "args", "kids", "nin", "nout")

    def __init__(self, op, args=(), kids=(), nin=0, nout=0):
        self.op, self.args, self.kids, self.nin, self.nout = op, tuple(args), tuple(kids), int(nin), int(nout)

    # -- arity (AudioNode::inputs / outputs)
    def inputs(self):
        return self.nin

    def outputs(self):
        return self.nout

    # -- operators (src/combinator.rs:289-488)
    def __rshift__(self, y):
        _arity(self.nout == y.nin, f"pipe: {self.nout} outputs >> {y.nin} inputs")
        return An("pipe", (), (self, y), self.nin, y.nout)

    def __or__(self, y):
        return An("stack", (), (self, y), self.nin + y.nin, self.nout + y.nout)

    def __and__(self, y):
        _arity(self.nin == y.nin and self.nout == y.nout, "bus: mismatched arity")
        return An("bus", (), (self, y), self.nin, self.nout)

    def __xor__(self, y):
        _arity(self.nin == y.nin, "branch: mismatched inputs")
        return An("branch", (), (self, y), self.nin, self.nout + y.nout)

    def __invert__(self):  # Rust `!x`
        return An("thru", (), (self,), self.nin, self.nin)

    def __neg__(self):
        return An("unop", (U_NEG, 0.0), (self,), self.nin, self.nout)

    def _bin(self, op, y):
        _arity(self.nout == y.nout, "binop: mismatched outputs")
        return An("binop", (op,), (self, y), self.nin + y.nin, self.nout)

    def __add__(self, y):
        return self._bin(OP_ADD, y) if isinstance(y, An) else An("unop", (U_ADD, f32(y)), (self,), self.nin, self.nout)

    def __radd__(self, y):
        return An("unop", (U_ADD, f32(y)), (self,), self.nin, self.nout)

    def __sub__(self, y):
        return self._bin(OP_SUB, y) if isinstance(y, An) else An("unop", (U_ADD, f32(-f32(y))), (self,), self.nin, self.nout)

    def __rsub__(self, y):
        return An("unop", (U_NEGADD, f32(y)), (self,), self.nin, self.nout)

    def __mul__(self, y):
        return self._bin(OP_MUL, y) if isinstance(y, An) else An("unop", (U_MUL, f32(y)), (self,), self.nin, self.nout)

    def __rmul__(self, y):
        return An("unop", (U_MUL, f32(y)), (self,), self.nin, self.nout)

    # -- builder methods (src/combinator.rs:263-286)
    def phase(self, p):
        return An("phase", (f32(p),), (self,), self.nin, self.nout)

    def seed(self, s):
        return An("seed", (int(s) & 0xFFFFFFFFFFFFFFFF,), (self,), self.nin, self.nout)

    def set(self, kind, values=(), seed=0, address=()):
        """Apply `Setting{parameter(kind, values), address}` (src/setting.rs:52-211) after construction.
        `address` is a sequence of (type, value) with type 1 = Index, 2 = Node."""
        return An("set", (int(kind), tuple(f32(v) for v in values), int(seed), tuple(address)), (self,), self.nin, self.nout)

    # -- lowering
    def lower(self, backend):
        """Build this expression on `backend` (an object with one method per primitive op)."""
        kids = [k.lower(backend) for k in self.kids]
        return getattr(backend, "b_" + self.op)(*self.args, *kids)

    def __repr__(self):
        a = ",".join(repr(x) for x in self.args)
        k = ",".join(repr(x) for x in self.kids)
        return f"{self.op}({a}{';' if a and k else ''}{k})"


class ArityError(ValueError):
    """The reference rejects these at compile time (typenum arity mismatch)."""


def _arity(ok, msg):
    if not ok:
        raise ArityError(msg)


def multi(kind, op, nodes):
    nodes = list(nodes)
    _arity(len(nodes) > 0, "multi: no nodes")
    x = nodes[0]
    nin = x.nin * len(nodes) if kind in (M_STACK, M_REDUCE) else x.nin
    nout = x.nout * len(nodes) if kind in (M_STACK, M_BRANCH) else x.nout
    return An("multi", (kind, op, len(nodes)), nodes, nin, nout)
