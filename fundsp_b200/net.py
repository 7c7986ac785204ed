"""`Net`: host-side mirror of the reference's dynamic graph (src/net.rs) — vertices of units, edges, global I/O —
and of its graph algebra (`&`, `>>`, `|`, `^`, `+`, `-`, `*` on nets, src/net.rs:1447-1832).

Like `An`, a `Net` here is a pure description; `lower(backend)` replays it onto a backend (oracle or the C ABI) as
`net_new / net_push / net_connect*` calls in vertex order, which is the order `Net::ping` threads the phase hash
through (src/net.rs:1383-1389). Vertex ids are indices (nothing is ever removed here).
"""
from __future__ import annotations

from .graph import An, OP_ADD, OP_MUL, OP_SUB
from .prelude import pass_

ZERO = ("zero",)


def G(i):
    return ("global", i)


def L(node, port):
    return ("local", node, port)


class Net:
    def __init__(self, inputs, outputs):  # Net::new (src/net.rs:175-200)
        self.nin, self.nout = inputs, outputs
        self.vertex = []  # [unit(An), [source ports]]
        self.out = [ZERO] * outputs

    def inputs(self):
        return self.nin

    def outputs(self):
        return self.nout

    def size(self):
        return len(self.vertex)

    # ---- construction (src/net.rs:204-213, 472-787)
    def push(self, unit: An):
        self.vertex.append([unit, [ZERO] * unit.inputs()])
        return len(self.vertex) - 1

    def connect(self, source, source_port, target, target_port):
        assert source != target
        self.vertex[target][1][target_port] = L(source, source_port)

    def connect_input(self, global_input, target, target_port):
        self.vertex[target][1][target_port] = G(global_input)

    def connect_output(self, source, source_port, global_output):
        self.out[global_output] = L(source, source_port)

    def pass_through(self, global_input, global_output):
        self.out[global_output] = G(global_input)

    def pipe_input(self, target):
        u = self.vertex[target][0]
        self.vertex[target][1] = [G(c % self.nin) if self.nin > 0 else ZERO for c in range(u.inputs())]

    def pipe_output(self, source):
        no = self.vertex[source][0].outputs()
        self.out = [L(source, c % no) if no > 0 else ZERO for c in range(self.nout)]

    def pipe_all(self, source, target):
        if source == target:
            return
        no = self.vertex[source][0].outputs()
        self.vertex[target][1] = [L(source, c % no) if no > 0 else ZERO for c in range(self.vertex[target][0].inputs())]

    def chain(self, unit: An):  # src/net.rs:764-787
        idx = self.push(unit)
        if self.size() == 1:
            if self.nin > 0:
                self.pipe_input(idx)
        else:
            self.vertex[idx][1] = [self.out[i % self.nout] if self.nout > 0 else ZERO for i in range(unit.inputs())]
        self.pipe_output(idx)
        return idx

    @staticmethod
    def wrap(unit: An):  # src/net.rs:925-935
        n = Net(unit.inputs(), unit.outputs())
        i = n.push(unit)
        if n.nin > 0:
            n.pipe_input(i)
        if n.nout > 0:
            n.pipe_output(i)
        return n

    # ---- algebra (src/net.rs:1447-1832); operands are consumed conceptually (we copy)
    def _append(self, other, input_offset=0, global_to=None):
        off = len(self.vertex)
        for unit, src in other.vertex:
            ns = []
            for s in src:
                if s[0] == "local":
                    ns.append(L(s[1] + off, s[2]))
                elif s[0] == "global":
                    ns.append(global_to[s[1]] if global_to is not None else G(s[1] + input_offset))
                else:
                    ns.append(ZERO)
            self.vertex.append([unit, ns])
        return off

    def _copy(self):
        n = Net(self.nin, self.nout)
        n.vertex = [[u, list(s)] for u, s in self.vertex]
        n.out = list(self.out)
        return n

    def _binary(self, other, op, share_inputs):
        assert self.nout == other.nout and (not share_inputs or self.nin == other.nin)
        n = self._copy()
        o1, o2 = list(n.out), list(other.out)
        input_offset = 0 if share_inputs else n.nin
        off = n._append(other, input_offset)
        if not share_inputs:
            n.nin += other.nin
        add = len(n.vertex)
        for i in range(n.nout):
            n.push(pass_() + pass_() if op == OP_ADD else (pass_() - pass_() if op == OP_SUB else pass_() * pass_()))
            n.out[i] = L(add + i, 0)
        for i, s in enumerate(o1):
            if s[0] == "local":
                n.vertex[add + i][1][0] = L(s[1], s[2])
            elif s[0] == "global":
                n.vertex[add + i][1][0] = G(s[1])
        for i, s in enumerate(o2):
            if s[0] == "local":
                n.vertex[add + i][1][1] = L(s[1] + off, s[2])
            elif s[0] == "global":
                n.vertex[add + i][1][1] = G(s[1] + input_offset)
        return n

    def __and__(self, other):  # Net::bus
        return self._binary(other, OP_ADD, True)

    def __add__(self, other):  # Net::sum
        return self._binary(other, OP_ADD, False)

    def __sub__(self, other):
        return self._binary(other, OP_SUB, False)

    def __mul__(self, other):  # Net::product
        return self._binary(other, OP_MUL, False)

    def __or__(self, other):  # Net::stack
        n = self._copy()
        off = n._append(other, n.nin)
        for s in other.out:
            n.out.append(L(s[1] + off, s[2]) if s[0] == "local" else (G(s[1] + n.nin) if s[0] == "global" else ZERO))
        n.nin += other.nin
        n.nout += other.nout
        return n

    def __xor__(self, other):  # Net::branch
        assert self.nin == other.nin
        n = self._copy()
        off = n._append(other, 0)
        for s in other.out:
            n.out.append(L(s[1] + off, s[2]) if s[0] == "local" else s)
        n.nout += other.nout
        return n

    def __rshift__(self, other):  # Net::pipe
        assert self.nout == other.nin
        n = self._copy()
        oe1 = list(n.out)
        off = n._append(other, 0, global_to=oe1)
        n.out = [L(s[1] + off, s[2]) if s[0] == "local" else (oe1[s[1]] if s[0] == "global" else ZERO) for s in other.out]
        n.nout = other.nout
        return n

    # ---- lowering
    def node(self):
        """This Net as a graph node (an `An`), e.g. `noise() >> net.node()`: the reference's `An<Net>` / `Net` inside expressions."""
        return An("netnode", (self,), (), self.inputs(), self.outputs())

    def lower(self, backend):
        h = backend.net_new(self.nin, self.nout)
        for unit, _ in self.vertex:
            backend.net_push(h, unit.lower(backend))
        for t, (_, src) in enumerate(self.vertex):
            for p, s in enumerate(src):
                if s[0] == "local":
                    backend.net_connect(h, s[1], s[2], t, p)
                elif s[0] == "global":
                    backend.net_connect_input(h, s[1], t, p)
        for o, s in enumerate(self.out):
            if s[0] == "local":
                backend.net_connect_output(h, s[1], s[2], o)
            elif s[0] == "global":
                backend.net_pass_through(h, s[1], o)
        return h


def balanced_bus(nets):
    """Bus many nets as a balanced tree: level-wise adjacent pairing, the odd one carried up (SURVEY.md §8d item 5:
    a 65536-long left-leaning `&` chain would recurse 65536 frames deep in the reference's order computation)."""
    cur = list(nets)
    while len(cur) > 1:
        nxt = [cur[i] & cur[i + 1] for i in range(0, len(cur) - 1, 2)]
        if len(cur) & 1:
            nxt.append(cur[-1])
        cur = nxt
    return cur[0]


def voice_net(voices):
    """The dynamic-Net form of a voice bank (config 5): each voice wrapped as a Net and bussed to the outputs."""
    return balanced_bus([Net.wrap(v) for v in voices])
