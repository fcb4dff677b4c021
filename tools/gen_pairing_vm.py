#!/usr/bin/env python
"""Generator of the pairing VM program (dvt_circuits_b200/csrc/pairing_prog.inc).

The batched BLS check e(pk, H(m)) == e(G1, sig) (crates/dkg/src/crypto/bls_common.rs:26-35) runs on the GPU as a small
virtual machine (csrc/pairing_vm.cuh): a block is R warps ("roles") that work TOGETHER on 32 checks (lane = check).  Every
Fp2 value of a check lives in a shared-memory slot (96 B, lane-interleaved: conflict-free), each role executes its own stream
of register-machine instructions (two Fp2 registers X, Y per thread; loads / signed sums from slots, product, square, store)
and the roles meet at a block barrier after every dependency level.  Nothing of a check's working set (an Fp12 is 576 B) is
ever in local memory - round 1's one-thread-per-check kernel moved ~1.1 MB of spill traffic per check.

This script writes the formulas ONCE, at Fp2 level, in a tiny DSL (`Prog`):
    Miller loop     per step: the point step produces the merged line L_j = (prepared line of H(m), scaled by pk) * (line through
                    the running multiple of sig, evaluated at -G1) - a 014 x 014 sparse product (6 Fp2 products) - then
                    f <- f^2 * L_j (or f * L_j) with ONE 17-product multiplication instead of two 13-product sparse ones
    final exponentiation   f^(3 (p^12 - 1) / r) exactly as csrc/tower.cuh (easy part, (x-1)^2 (x+p) (x^2+p^2-1) + 3, Granger-Scott
                    cyclotomic squarings)
schedules every segment onto R roles by dependency level (longest-processing-time first within a level), allocates the
temporaries to slots by liveness, emits the instruction streams, and SIMULATES them with Python integers against the Python
restatement of the reference (oracle/pyref) - `python tools/gen_pairing_vm.py --check` - so the program is proven before any
CUDA runs.  tests/test_pairing_vm.py runs the same check in the CPU suite and the C++ interpreter on the host.
"""
import argparse
import os
import sys
from collections import defaultdict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
X_ABS = 0xD201000000010000
G1X = 0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB
G1Y = 0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1

# ---------------------------------------------------------------------------------------------- instruction set
OPS = ["END", "BAR",
       "LDX", "ADDX", "SUBX", "LDY", "ADDY", "SUBY", "STX",          # arg = slot
       "LDXK", "LDYK",                                                  # arg = constant index (Fp2 constants, __constant__)
       "LDXL", "ADDXL",                                                 # arg = coefficient 0..2 of the current prepared line of H(m)
       "LDXIN", "ADDXIN", "SUBXIN", "LDYIN", "ADDYIN",                  # arg = input 0: (xp, yp) of pk, 1: sig.x, 2: sig.y
       "LDYS",                                                          # arg = component 0 / 1 of input 0: Y = (xp or yp, 0)
       "STXL",                                                          # arg = coefficient 0..2: X -> the prepared-line buffer (line preparation)
       "MUL", "SQR", "XI", "NEGX", "DBLX", "TPLX", "CONJX", "INVX"]   # X <- X*Y, X^2, X*(1+u), -X, 2X, 3X, conj X, 1/X
OP = {n: i for i, n in enumerate(OPS)}
COST = {"mul": 888, "sqr": 600, "lin": 0, "inv": 21 * 888}  # scheduling weights in wide-MAC issue slots (inversion: binary extended Euclid, ~38 k ALU instructions)
ADD_COST = 45


# ---------------------------------------------------------------------------------------------- Fp2 on Python integers
def f2add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2neg(a): return ((-a[0]) % P, (-a[1]) % P)
def f2mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2xi(a): return ((a[0] - a[1]) % P, (a[0] + a[1]) % P)
def f2conj(a): return (a[0], (-a[1]) % P)


def f2inv(a):
    n = pow((a[0] * a[0] + a[1] * a[1]) % P, P - 2, P)
    return (a[0] * n % P, (-a[1]) * n % P)


def f2pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = f2mul(r, a)
        a = f2mul(a, a)
        e >>= 1
    return r


# ---------------------------------------------------------------------------------------------- DSL
class Var:
    """one Fp2 value; loc: None (temporary, allocated later) | ('S', fixed slot name) | ('K', const) | ('L', k) | ('IN', k)"""
    __slots__ = ("name", "loc", "node", "slot")

    def __init__(self, name, loc=None):
        self.name, self.loc, self.node, self.slot = name, loc, None, None

    def __repr__(self):
        return self.name


class Lin:
    """(sum_k xi^k * sum_v c_{v,k} v) * xi^pk * m  with small integer c, m = 2^a 3^b; `conj` only for single-term expressions"""

    def __init__(self, terms=None, mult=1, pk=0, conj=False):
        self.t = dict(terms or {})  # (var, k) -> c
        self.mult, self.pk, self.conj = mult, pk, conj

    @staticmethod
    def of(x):
        if isinstance(x, Lin):
            return x
        return Lin({(x, 0): 1})

    def _flat(self):
        """terms with the outer xi-power folded in (outer multiplier must be 1)"""
        assert self.mult == 1 and not self.conj
        return {(v, k + self.pk): c for (v, k), c in self.t.items()}

    def __add__(self, o):
        o = Lin.of(o)
        d = defaultdict(int, self._flat())
        for key, c in o._flat().items():
            d[key] += c
        return Lin({k: c for k, c in d.items() if c})

    def __sub__(self, o):
        return self + (-Lin.of(o))

    def __neg__(self):
        if self.mult != 1:
            return Lin(self.t, -self.mult, self.pk, self.conj)
        return Lin({k: -c for k, c in self.t.items()}, 1, self.pk, self.conj)

    def __rmul__(self, m):  # small integer * expression
        assert isinstance(m, int) and m != 0
        if abs(m) <= 3 and self.mult == 1 and not self.conj:
            return Lin({k: c * m for k, c in self.t.items()}, 1, self.pk)
        return Lin(self.t, self.mult * m, self.pk, self.conj)

    def xi(self):
        return Lin(self.t, self.mult, self.pk + 1, self.conj)

    def cj(self):
        assert len(self.t) == 1 and self.mult == 1 and self.pk == 0
        return Lin(self.t, 1, 0, True)

    def vars(self):
        return {v for (v, _k) in self.t}


def V(x):
    return Lin.of(x)


class Node:
    def __init__(self, kind, x, y, out, post=(0, 1)):
        self.kind, self.x, self.y, self.out, self.post = kind, x, y, out, post
        self.level = self.role = None
        self.min_level = 1

    def inputs(self):
        s = set(self.x.vars())
        if self.y is not None:
            s |= self.y.vars()
        return s

    def cost(self):
        def adds(l):
            g = max(abs(c) for c in l.t.values())
            if g in (2, 3) and len(l.t) > 1 and all(k == 0 for (_v, k), c in l.t.items() if abs(c) != g):
                return sum(1 if abs(c) == g else abs(c) for c in l.t.values()) + 1
            return sum(abs(c) for c in l.t.values())
        n_add = adds(self.x) + (adds(self.y) if self.y is not None else 0)
        return COST[self.kind] + ADD_COST * (n_add + 2) + (COST["mul"] if getattr(self, "then_const", None) is not None else 0)


def factor_chain(m):
    """ops multiplying X by |m| = 2^a 3^b (then NEGX when m < 0)"""
    ops, n = [], abs(m)
    while n % 3 == 0:
        ops.append("TPLX")
        n //= 3
    while n % 2 == 0:
        ops.append("DBLX")
        n //= 2
    assert n == 1, m
    if m < 0:
        ops.append("NEGX")
    return ops


class Prog:
    """one segment: a straight-line Fp2 program over fixed state slots, constants, the prepared line and the inputs"""

    def __init__(self, name, state_slots):
        self.name, self.state_slots = name, state_slots
        self.nodes = []
        self.cur = {}      # fixed slot name -> Var currently holding it
        self.readers = defaultdict(list)  # Var -> nodes reading it
        self.n_tmp = 0

    # ---- operands
    def state(self, slot):
        if slot not in self.cur:
            assert slot in self.state_slots, slot
            self.cur[slot] = Var(slot + "@in", ("S", slot))
        return Lin.of(self.cur[slot])

    @staticmethod
    def const(k): return Lin.of(Var(f"K{k}", ("K", k)))
    @staticmethod
    def line(k): return Lin.of(Var(f"LINE{k}", ("L", k)))
    @staticmethod
    def inp(k): return Lin.of(Var(f"IN{k}", ("IN", k)))

    def at_least(self, level):
        """nodes created from now on sit at dependency level >= `level` (hand-placed pipelining inside a segment)"""
        self.floor = level

    def _node(self, kind, x, y, dst, post):
        if dst is None:
            self.n_tmp += 1
            out = Var(f"{self.name}.t{self.n_tmp}")
        else:
            assert dst in self.state_slots, dst
            out = Var(f"{dst}@{len(self.nodes)}", ("S", dst))
        n = Node(kind, x, y, out, post)
        n.min_level = getattr(self, "floor", 1)
        out.node = n
        for v in n.inputs():
            self.readers[v].append(n)
        self.nodes.append(n)
        if dst is not None:
            n.overwrites = self.cur.get(dst) or self.state(dst).vars().pop()
            self.cur[dst] = out
        else:
            n.overwrites = None
        return Lin.of(out)

    def mul(self, a, b, dst=None, post=(0, 1)):
        """xi^post[0] * post[1] * a * b"""
        a, b = Lin.of(a), Lin.of(b)
        return self._node("mul", a, b, dst, post)

    def sqr(self, a, dst=None, post=(0, 1), then_const=None):
        """xi^post[0] * post[1] * a^2 [* constant then_const]"""
        r = self._node("sqr", Lin.of(a), None, dst, post)
        self.nodes[-1].then_const = then_const
        return r

    def lin(self, a, dst=None):
        return self._node("lin", Lin.of(a), None, dst, (0, 1))

    def inv(self, a, dst=None):
        return self._node("inv", Lin.of(a), None, dst, (0, 1))

    def line_out(self, k, a):
        """coefficient k of the line being prepared <- a (global memory, written once, never read back by the segment)"""
        r = self._node("lin", Lin.of(a), None, None, (0, 1))
        self.nodes[-1].line_out = k
        return r

    # ---- scheduling
    def schedule(self, R, balance=True):
        nodes = self.nodes
        changed = True
        while changed:
            changed = False
            for n in nodes:
                lv = n.min_level
                for v in n.inputs():
                    if v.node is not None:
                        lv = max(lv, v.node.level + 1 if v.node.level else 1)
                if n.overwrites is not None:  # write-after-read on a fixed slot: strictly after every OTHER reader of the old value
                    for r in self.readers.get(n.overwrites, []):
                        if r is not n and r.level:
                            lv = max(lv, r.level + 1)
                    if n.overwrites.node is not None and n.overwrites.node.level:
                        lv = max(lv, n.overwrites.node.level + 1)
                if n.level != lv:
                    n.level = lv
                    changed = True
        self.n_levels = max(n.level for n in nodes)
        self.R = R
        if balance:
            self._balance(R)
        for lv in range(1, self.n_levels + 1):
            load = [0] * R
            for n in sorted([n for n in nodes if n.level == lv], key=lambda n: -n.cost()):
                r = min(range(R), key=lambda i: load[i])
                n.role = r
                load[r] += n.cost()
        # a role executes its nodes of a level in program order; a node writing a fixed slot must not run before a node of the SAME
        # role and level that still reads the old value - program order guarantees that (the reader was created first)

    def _makespan(self, R, lv):
        load = [0] * R
        for n in sorted([n for n in self.nodes if n.level == lv], key=lambda n: -n.cost()):
            r = min(range(R), key=lambda i: load[i])
            load[r] += n.cost()
        return max(load)

    def _latest(self, n):
        """latest level node n may move to with every other node where it is"""
        hi = self.n_levels
        for c in self.readers.get(n.out, []):
            hi = min(hi, c.level - 1)
        for m in self.nodes:
            if m is n or m.overwrites is None:
                continue
            if m.overwrites is n.out:                      # a later write to the same fixed slot
                hi = min(hi, m.level - 1)
            if m.overwrites in n.inputs():                 # n still reads the value m overwrites
                hi = min(hi, m.level - 1)
        return hi

    def _balance(self, R):
        """nodes with slack move to a later level when that shortens the sum over levels of the busiest role's load"""
        improved = True
        while improved:
            improved = False
            for n in sorted(self.nodes, key=lambda n: -n.cost()):
                lo, hi = n.level, self._latest(n)
                if hi <= lo:
                    continue
                base = {lv: self._makespan(R, lv) for lv in range(lo, hi + 1)}
                best, best_gain = None, 0
                for lv in range(lo + 1, hi + 1):
                    n.level = lv
                    gain = (base[lo] + base[lv]) - (self._makespan(R, lo) + self._makespan(R, lv))
                    if gain > best_gain:
                        best, best_gain = lv, gain
                n.level = best if best is not None else lo
                improved |= best is not None

    def allocate(self, fixed_index, n_fixed):
        """temporaries -> slots by liveness over levels: first the dead windows of the fixed state slots (between the last read of a
        value and the level that writes the next one - e.g. the merged line's slots while the step is still computing it), then
        slots >= n_fixed; returns number of slots used"""
        last = {}
        for n in self.nodes:
            for v in n.inputs():
                last[v] = max(last.get(v, 0), n.level)
        # dead windows [lo, hi] of fixed slots
        versions = defaultdict(list)  # slot name -> [(def level, var)] in program order
        for n in self.nodes:
            for v in n.inputs():
                if v.loc is not None and v.loc[0] == "S" and v.node is None and not versions[v.loc[1]]:
                    versions[v.loc[1]].append((0, v))
        for n in self.nodes:
            if n.out.loc is not None:
                name = n.out.loc[1]
                if not versions[name]:
                    versions[name].append((0, None))  # the incoming value is never read in this segment: dead on entry
                versions[name].append((n.level, n.out))
        windows = []  # (slot index, lo, hi)
        for name, vs in versions.items():
            for (d0, v0), (d1, _v1) in zip(vs, vs[1:]):
                lo = (max(last.get(v0, d0), d0) if v0 is not None else 0) + 1
                if lo <= d1 - 1:
                    windows.append([fixed_index[name], lo, d1 - 1])
        free_at = []  # (slot, free from level)
        used = n_fixed
        for n in sorted(self.nodes, key=lambda n: n.level):
            v = n.out
            if v.loc is not None:
                v.slot = fixed_index[v.loc[1]]
                continue
            if getattr(n, "line_out", None) is not None:
                v.slot = 0  # never stored to a slot
                continue
            end = last.get(v, n.level)
            pick = None
            for w in windows:  # a fixed slot that is dead for the whole life of this temporary
                if w[1] <= n.level and end <= w[2]:
                    v.slot = w[0]
                    w[1] = end + 1
                    pick = -1
                    break
            if pick == -1:
                continue
            for i, (s, fr) in enumerate(free_at):
                if fr <= n.level:
                    pick = i
                    break
            if pick is None:
                v.slot = used
                used += 1
                free_at.append((v.slot, end + 1))
            else:
                v.slot = free_at[pick][0]
                free_at[pick] = (v.slot, end + 1)
        for n in self.nodes:
            for v in n.inputs():
                if v.loc is not None and v.loc[0] == "S":
                    v.slot = fixed_index[v.loc[1]]
        return used

    # ---- emission
    def _emit_operand(self, lin, reg):
        """instructions computing `lin` into register X or Y"""
        ins = []
        if lin.conj:
            assert reg == "X"
        # 3 (a + xi b) - 2 z  /  2 (a + b - c): sum the terms that share the largest coefficient once, scale, then add the rest
        if reg == "X" and not lin.conj and len(lin.t) > 1:
            g = max(abs(c) for c in lin.t.values())
            rest = {k: c for k, c in lin.t.items() if abs(c) != g}
            if g in (2, 3) and all(k == 0 for (_v, k) in rest) and \
                    any(c > 0 and (v.loc is None or v.loc[0] == "S") for (v, _k), c in lin.t.items() if abs(c) == g):
                head = Lin({k: (1 if c > 0 else -1) for k, c in lin.t.items() if abs(c) == g})
                ins = self._emit_operand(head, "X") + [("TPLX" if g == 3 else "DBLX", 0)]
                for (v, _k), c in sorted(rest.items(), key=lambda kv: kv[1] < 0):
                    assert v.loc is None or v.loc[0] == "S"
                    ins += [(("ADDX" if c > 0 else "SUBX"), v.slot)] * abs(c)
                for _ in range(lin.pk):
                    ins.append(("XI", 0))
                if lin.mult != 1:
                    ins += [(o, 0) for o in factor_chain(lin.mult)]
                return ins
        by_k = defaultdict(list)
        for (v, k), c in lin.t.items():
            by_k[k].append((v, c))
        kmax = max(by_k)
        if reg == "Y":
            assert kmax == 0 and lin.pk == 0 and lin.mult == 1 and not lin.conj, "Y operand must be a plain signed sum"
        empty = True
        for k in range(kmax, -1, -1):
            if not empty:
                ins.append(("XI", 0))
            terms = sorted(by_k.get(k, []), key=lambda vc: (vc[1] < 0, vc[0].loc is None or vc[0].loc[0] == "S"))
            # positive terms first; among them pseudo-variables (constants / line / inputs) first: they may only START a sum
            for v, c in terms:
                for _ in range(abs(c)):
                    kind = v.loc[0] if v.loc is not None else "S"
                    if empty:
                        ld = {"S": "LD" + reg, "K": "LD" + reg + "K", "L": "LDXL", "IN": "LD" + reg + "IN"}[kind]
                        assert not (kind == "L" and reg == "Y")
                        ins.append((ld, v.slot if kind == "S" else v.loc[1]))
                        if c < 0:
                            assert reg == "X", "negative leading term in Y"
                            ins.append(("NEGX", 0))
                        empty = False
                    else:
                        assert kind in ("S", "IN", "L"), "constants may only start a sum"
                        if kind == "S":
                            ins.append((("ADD" if c > 0 else "SUB") + reg, v.slot))
                        elif kind == "IN":
                            assert c > 0 or reg == "X"
                            ins.append((("ADD" if c > 0 else "SUB") + reg + "IN", v.loc[1]))
                        else:
                            assert c > 0 and reg == "X"
                            ins.append(("ADDXL", v.loc[1]))
        if lin.conj:
            ins.append(("CONJX", 0))
        for _ in range(lin.pk):
            ins.append(("XI", 0))
        if lin.mult != 1:
            ins += [(o, 0) for o in factor_chain(lin.mult)]
        return ins

    def emit(self):
        """-> streams[role] = list of (op, arg), one BAR per level, END at the end"""
        streams = [[] for _ in range(self.R)]
        for lv in range(1, self.n_levels + 1):
            for n in self.nodes:
                if n.level != lv:
                    continue
                s = streams[n.role]
                x, y = n.x, n.y
                if n.kind == "mul":
                    # Y takes plain sums only: move an outer xi-power / multiplier / conjugation-free factor of y over to x, or swap
                    def plain(l):
                        return l.pk == 0 and l.mult == 1 and not l.conj and all(k == 0 for (_v, k) in l.t) and \
                            any(c > 0 for c in l.t.values()) and \
                            all(v.loc is None or v.loc[0] in ("S", "IN") or len(l.t) == 1 for (v, _k) in l.t) and \
                            not any(v.loc is not None and v.loc[0] == "L" for (v, _k) in l.t)
                    if not plain(y) and plain(x):
                        x, y = y, x
                    if not plain(y):  # factor xi^pk * mult out of y into the post-processing of the product
                        assert all(k == 0 for (_v, k) in y.t) and not y.conj, (self.name, "unsupported Y operand")
                        n.post = (n.post[0] + y.pk, n.post[1] * y.mult)
                        y = Lin(y.t)
                        assert plain(y), (self.name, "unsupported Y operand")
                    # scale by a component of pk: Y = (xp, 0) or (yp, 0) is requested with the pseudo-input 3 / 4
                    if len(y.t) == 1 and list(y.t)[0][0].loc == ("IN", 3):
                        yi = [("LDYS", 0)]
                    elif len(y.t) == 1 and list(y.t)[0][0].loc == ("IN", 4):
                        yi = [("LDYS", 1)]
                    else:
                        yi = self._emit_operand(y, "Y")
                    s += self._emit_operand(x, "X") + yi + [("MUL", 0)]
                elif n.kind == "sqr":
                    s += self._emit_operand(x, "X") + [("SQR", 0)]
                    if getattr(n, "then_const", None) is not None:
                        s += [("LDYK", n.then_const), ("MUL", 0)]
                elif n.kind == "inv":
                    s += self._emit_operand(x, "X") + [("INVX", 0)]
                else:
                    s += self._emit_operand(x, "X")
                for _ in range(n.post[0]):
                    s.append(("XI", 0))
                if n.post[1] != 1:
                    s += [(o, 0) for o in factor_chain(n.post[1])]
                if getattr(n, "line_out", None) is not None:
                    s.append(("STXL", n.line_out))
                else:
                    s.append(("STX", n.out.slot))
            for s in streams:
                s.append(("BAR", 0))
        for s in streams:
            s.append(("END", 0))
        return streams


# ---------------------------------------------------------------------------------------------- the formulas
# fixed state (Fp2 slots).  RA = f / the accumulator of the final exponentiation, coefficient order of tower.cuh's Fp12:
# c0.c0 c0.c1 c0.c2 c1.c0 c1.c1 c1.c2.  The Miller loop's T / L and the final exponentiation's RB share slots 6..13.
STATE_MILLER = ["A0", "A1", "A2", "A3", "A4", "A5", "TX", "TY", "TZ", "L00", "L01", "L02", "L11", "L12"]
STATE_FE = ["A0", "A1", "A2", "A3", "A4", "A5", "B0", "B1", "B2", "B3", "B4", "B5"]
FIXED_INDEX = {n: i for i, n in enumerate(STATE_MILLER)}
FIXED_INDEX.update({n: i for i, n in enumerate(STATE_FE)})
N_FIXED = len(STATE_MILLER)
IN_PK, IN_QX, IN_QY, IN_XP, IN_YP = 0, 1, 2, 3, 4
K_ONE, K_M3XG, K_M2YG, K_MXG, K_MYG, K_F1, K_F2, K_F12, K_F12F1, K_F12F2, K_ZERO = range(11)


def constants():
    """Fp2 constants of the program (canonical integers)"""
    from oracle.pyref import bls12_381 as B
    xi = (1, 1)
    f1 = B.f2_pow(xi, (P - 1) // 3)
    f2 = B.f2_pow(xi, 2 * (P - 1) // 3)
    f12 = B.f2_pow(xi, (P - 1) // 6)
    return [(1, 0), ((-3 * G1X) % P, 0), ((-2 * G1Y) % P, 0), ((-G1X) % P, 0), ((-G1Y) % P, 0),
            f1, f2, f12, f2mul(f12, f1), f2mul(f12, f2), (0, 0)]


def fp6_mul(p, x, y, ytag=None):
    """Karatsuba (6 products); x, y = 3 Lin each.  Returns 3 Lin over the product variables.  The y side may carry xi."""
    v0, v1, v2 = p.mul(y[0], x[0]), p.mul(y[1], x[1]), p.mul(y[2], x[2])
    t3 = p.mul(y[1] + y[2], x[1] + x[2])
    t4 = p.mul(y[0] + y[1], x[0] + x[1])
    t5 = p.mul(y[0] + y[2], x[0] + x[2])
    return [v0 + (t3 - v1 - v2).xi(), t4 - v0 - v1 + v2.xi(), t5 - v0 - v2 + v1]


def mul_v(a):
    return [a[2].xi(), a[0], a[1]]


def f_times_line(p, f0, f1, L):
    """(f0 + f1 w) * ((L00, L01, L02) + (0, L11, L12) w): 6 + 5 + 6 products -> 6 Lin"""
    L00, L01, L02, L11, L12 = L
    A = fp6_mul(p, [L00, L01, L02], f0)                       # f0 * L0  (operands swapped: sums with xi go to X)
    u1, u2 = p.mul(f1[1], L11), p.mul(f1[2], L12)
    u3 = p.mul(f1[1] + f1[2], L11 + L12)
    x0y, x0z = p.mul(f1[0], L11), p.mul(f1[0], L12)
    Bv = [(u3 - u1 - u2).xi(), x0y + u2.xi(), x0z + u1]      # f1 * (0, L11, L12)
    C = fp6_mul(p, [L00, L01 + L11, L02 + L12], [f0[i] + f1[i] for i in range(3)])
    vB = mul_v(Bv)
    return [A[0] + vB[0], A[1] + vB[1], A[2] + vB[2], C[0] - A[0] - Bv[0], C[1] - A[1] - Bv[1], C[2] - A[2] - Bv[2]]


def merged_line(p, h, l):
    """(h00, h01, h11) x (l00, l01, l11), both of the sparse 014 shape -> the five coefficients L00 L01 L02 L11 L12"""
    a1, b1, c1 = h
    a2, b2, c2 = l
    M1, M2, M3 = p.mul(a1, a2), p.mul(b1, b2), p.mul(c1, c2)
    M4, M5, M6 = p.mul(a1 + b1, a2 + b2), p.mul(a1 + c1, a2 + c2), p.mul(b1 + c1, b2 + c2)
    return [M1 + M3.xi(), M4 - M1 - M2, M2, M5 - M1 - M3, M6 - M2 - M3]


def h_line(p):
    """prepared line of H(m) for this step, scaled by the key: (c00, c01 * xp, c11 * yp)  (tower.cuh fp12_mul_line)"""
    return [p.line(0), p.mul(p.line(1), p.inp(IN_XP)), p.mul(p.line(2), p.inp(IN_YP))]


def t_part_dbl(p):
    """doubling step of the running point T (RCB Alg. 9 over Fp2, b3 = 12 xi) + the tangent line at -G1; writes T and L"""
    X, Y, Z = p.state("TX"), p.state("TY"), p.state("TZ")
    h = h_line(p)
    A, C = p.sqr(Y), p.sqr(X)
    t2 = p.sqr(Z, post=(1, 12))            # b3 Z^2
    D, E = p.mul(Y, Z), p.mul(X, Y)
    P1 = p.mul(t2, A, post=(0, 8))         # x3 = t2 * 8 Y^2
    P2 = p.mul(A - 3 * t2, A + t2)
    p.mul(D, A, dst="TZ", post=(0, 8))     # Z3 = Y Z * 8 Y^2
    p.mul(A - 3 * t2, E, dst="TX", post=(0, 2))
    l01 = p.mul(C, p.const(K_M3XG))        # -3 X^2 * xG
    l11 = p.mul(D, p.const(K_M2YG))        # 2 Y Z * (-yG)
    p.lin(P1 + P2, dst="TY")
    Lc = merged_line(p, h, [A - t2, l01, l11])
    for nm, e in zip(["L00", "L01", "L02", "L11", "L12"], Lc):
        p.lin(e, dst=nm)


def t_part_add(p):
    """addition step T <- T + sig (RCB Alg. 8, mixed) + the chord through T and sig at -G1; writes T and L"""
    X, Y, Z = p.state("TX"), p.state("TY"), p.state("TZ")
    xq, yq = p.inp(IN_QX), p.inp(IN_QY)
    h = h_line(p)
    t0, t1 = p.mul(X, xq), p.mul(Y, yq)
    t3p = p.mul(X + Y, xq + yq)
    n_, d_ = p.mul(Z, yq), p.mul(Z, xq)
    t2 = p.lin(12 * Z.xi())
    t3 = t3p - t0 - t1
    t4, y3 = n_ + Y, d_ + X
    x3 = 3 * t0
    z3, t1m = t1 + t2, t1 - t2
    y3b = 12 * y3.xi()
    Pa, Pb = p.mul(t3, t1m), p.mul(y3b, t4)
    Pc, Pd = p.mul(t1m, z3), p.mul(y3b, x3)
    Pe, Pf = p.mul(z3, t4), p.mul(x3, t3)
    n, d = n_ - Y, d_ - X
    Pg, Ph = p.mul(n, xq), p.mul(d, yq)
    l01 = p.mul(n, p.const(K_MXG))         # -n * xG
    l11 = p.mul(d, p.const(K_MYG))         # d * (-yG)
    p.lin(Pa - Pb, dst="TX")
    p.lin(Pc + Pd, dst="TY")
    p.lin(Pe + Pf, dst="TZ")
    Lc = merged_line(p, h, [Pg - Ph, l01, l11])
    for nm, e in zip(["L00", "L01", "L02", "L11", "L12"], Lc):
        p.lin(e, dst=nm)


def prep_dbl(p):
    """line preparation, doubling step of the running multiple of H(m) (tower.cuh g2_line_dbl): the P-independent coefficients
    c00 = Y^2 - b3 Z^2, c01 = -3 X^2, c11 = 2 Y Z go to the line buffer, T <- 2T"""
    X, Y, Z = p.state("TX"), p.state("TY"), p.state("TZ")
    A, C = p.sqr(Y), p.sqr(X)
    t2 = p.sqr(Z, post=(1, 12))
    D, E = p.mul(Y, Z), p.mul(X, Y)
    p.line_out(0, A - t2)
    p.line_out(1, -(3 * C))
    p.line_out(2, 2 * D)
    P1 = p.mul(t2, A, post=(0, 8))
    P2 = p.mul(A - 3 * t2, A + t2)
    p.mul(D, A, dst="TZ", post=(0, 8))
    p.mul(A - 3 * t2, E, dst="TX", post=(0, 2))
    p.lin(P1 + P2, dst="TY")


def prep_add(p):
    """line preparation, addition step (tower.cuh g2_line_add): c00 = n xQ - d yQ, c01 = -n, c11 = d with n = yQ Z - Y, d = xQ Z - X; T <- T + Q"""
    X, Y, Z = p.state("TX"), p.state("TY"), p.state("TZ")
    xq, yq = p.inp(IN_QX), p.inp(IN_QY)
    t0, t1 = p.mul(X, xq), p.mul(Y, yq)
    t3p = p.mul(X + Y, xq + yq)
    n_, d_ = p.mul(Z, yq), p.mul(Z, xq)
    t2 = p.lin(12 * Z.xi())
    t3 = t3p - t0 - t1
    t4, y3 = n_ + Y, d_ + X
    x3 = 3 * t0
    z3, t1m = t1 + t2, t1 - t2
    y3b = 12 * y3.xi()
    Pa, Pb = p.mul(t3, t1m), p.mul(y3b, t4)
    Pc, Pd = p.mul(t1m, z3), p.mul(y3b, x3)
    Pe, Pf = p.mul(z3, t4), p.mul(x3, t3)
    n, d = n_ - Y, d_ - X
    Pg, Ph = p.mul(n, xq), p.mul(d, yq)
    p.line_out(0, Pg - Ph)
    p.line_out(1, Y - n_)
    p.line_out(2, d)
    p.lin(Pa - Pb, dst="TX")
    p.lin(Pc + Pd, dst="TY")
    p.lin(Pe + Pf, dst="TZ")


def f_state(p, pre="A"):
    return [p.state(pre + str(i)) for i in range(6)]


def f_part_sqr(p):
    """f <- f^2: complex squaring over Fp6 (2 x 6 products), back into f's own slots"""
    a = f_state(p)
    f0, f1 = a[:3], a[3:]
    ab = fp6_mul(p, f1, f0)
    vb = mul_v(f1)
    t = fp6_mul(p, [f0[i] + vb[i] for i in range(3)], [f0[i] + f1[i] for i in range(3)])
    vab = mul_v(ab)
    for i in range(3):
        p.lin(t[i] - ab[i] - vab[i], dst=f"A{i}")
        p.lin(2 * ab[i], dst=f"A{3 + i}")


def f_part_mul(p):
    a = f_state(p)
    L = [p.state(n) for n in ["L00", "L01", "L02", "L11", "L12"]]
    out = f_times_line(p, a[:3], a[3:], L)
    for i, e in enumerate(out):
        p.lin(e, dst=f"A{i}")


def f_part_copy(p):
    """f <- L (the first step: f was 1)"""
    L = [p.state(n) for n in ["L00", "L01", "L02", "L11", "L12"]]
    for i, e in zip([0, 1, 2, 4, 5], L):
        p.lin(e, dst=f"A{i}")
    p.lin(p.const(K_ZERO), dst="A3")


def seg_fe_mul(p):
    """RA <- RA * RB (Karatsuba over Fp6: 18 products)"""
    a, b = f_state(p, "A"), f_state(p, "B")
    A = fp6_mul(p, a[:3], b[:3])
    Bm = fp6_mul(p, a[3:], b[3:])
    C = fp6_mul(p, [a[i] + a[3 + i] for i in range(3)], [b[i] + b[3 + i] for i in range(3)])
    vB = mul_v(Bm)
    for i in range(3):
        p.lin(A[i] + vB[i], dst=f"A{i}")
        p.lin(C[i] - A[i] - Bm[i], dst=f"A{3 + i}")


def seg_fe_cyc(p):
    """RA <- RA^2 in the cyclotomic subgroup (Granger-Scott, tower.cuh fp12_cyc_sqr): 9 squarings"""
    a = f_state(p)
    z0, z4, z3, z2, z1, z5 = a

    def fp4(x, y):
        s0, s1, s2 = p.sqr(x), p.sqr(y), p.sqr(x + y)
        return s0 + s1.xi(), s2 - s0 - s1
    # (placing the third Fp4 squaring one level down, next to the sums of the first, balances the roles better on paper - 9
    # squarings on 6 roles - but costs a third barrier: measured 246.8 ms against 239.1 ms for 262 144 checks on B200)
    t0, t1 = fp4(z0, z1)
    u0, u1 = fp4(z2, z3)
    w0, w1 = fp4(z4, z5)
    p.lin(3 * t0 - 2 * z0, dst="A0")
    p.lin(3 * t1 + 2 * z1, dst="A4")
    p.lin(3 * u0 - 2 * z4, dst="A1")
    p.lin(3 * u1 + 2 * z5, dst="A5")
    p.lin(3 * w1.xi() + 2 * z2, dst="A3")
    p.lin(3 * w0 - 2 * z3, dst="A2")


def seg_fe_inv(p):
    """RB <- RA^-1 (tower.cuh fp12_inv / fp6_inv / fp2_inv)"""
    a = f_state(p)
    f0, f1 = a[:3], a[3:]
    s0 = fp6_mul(p, f0, f0)
    s1 = fp6_mul(p, f1, f1)
    vs1 = mul_v(s1)
    d = [p.lin(s0[i] - vs1[i]) for i in range(3)]
    q0, q1, q2 = p.sqr(d[0]), p.sqr(d[1]), p.sqr(d[2])
    m12, m01, m02 = p.mul(d[1], d[2]), p.mul(d[0], d[1]), p.mul(d[0], d[2])
    t0 = p.lin(q0 - m12.xi())
    t1 = p.lin(q2.xi() - m01)
    t2 = p.lin(q1 - m02)
    e0, e1, e2 = p.mul(d[0], t0), p.mul(d[2], t1), p.mul(d[1], t2)
    ninv = p.inv(e0 + (e1 + e2).xi())
    di = [p.mul(t0, ninv), p.mul(t1, ninv), p.mul(t2, ninv)]
    r0 = fp6_mul(p, f0, di)
    r1 = fp6_mul(p, f1, di)
    for i in range(3):
        p.lin(r0[i], dst=f"B{i}")
        p.lin(-r1[i], dst=f"B{3 + i}")


def seg_conj(p, pre):
    for i in range(3, 6):
        p.lin(-p.state(f"{pre}{i}"), dst=f"{pre}{i}")


def seg_frob_b(p):
    """RB <- RB^p (tower.cuh fp12_frob with the constants combined)"""
    b = f_state(p, "B")
    p.lin(b[0].cj(), dst="B0")
    for i, k in ((1, K_F1), (2, K_F2), (3, K_F12), (4, K_F12F1), (5, K_F12F2)):
        p.mul(b[i].cj(), p.const(k), dst=f"B{i}")


def build_segments(R):
    """name -> scheduled Prog.  (Running the f part of step j next to the point part of step j + 1 in one segment was tried: the
    critical path per step drops only from 13.6 to 12.7 product times - the roles are busy either way - while the live
    temporaries grow from 32 to 49 slots, i.e. from two resident blocks per SM to one.)"""
    segs = {}

    def seg(name, state, *parts):
        p = Prog(name, state)
        for part in parts:
            part(p)
        p.schedule(R)
        segs[name] = p
    # one segment per Miller step.  Doubling step: f^2 (which does not need the line) runs next to the point step that produces the
    # merged line, then f takes the line in: 6 dependency levels instead of 4 + 4 for the two halves run one after the other.
    seg("S_0", STATE_MILLER, t_part_dbl, f_part_copy)           # first step: f <- L_0
    seg("S_D", STATE_MILLER, f_part_sqr, t_part_dbl, f_part_mul)
    seg("S_A", STATE_MILLER, t_part_add, f_part_mul)
    seg("P_D", STATE_MILLER, prep_dbl)   # preparation of the lines of a hashed message (its own small kernel)
    seg("P_A", STATE_MILLER, prep_add)
    seg("FE_MUL", STATE_FE, seg_fe_mul)
    seg("FE_CYC", STATE_FE, seg_fe_cyc)
    seg("FE_INV", STATE_FE, seg_fe_inv)
    seg("FE_CONJA", STATE_FE, lambda p: seg_conj(p, "A"))
    seg("FE_CONJB", STATE_FE, lambda p: seg_conj(p, "B"))
    seg("FE_FROBB", STATE_FE, seg_frob_b)
    n_slots = 0
    for p in segs.values():
        n_slots = max(n_slots, p.allocate(FIXED_INDEX, N_FIXED))
    return segs, n_slots


SEG_ORDER = ["S_0", "S_D", "S_A", "P_D", "P_A", "FE_MUL", "FE_CYC", "FE_INV", "FE_CONJA", "FE_CONJB", "FE_FROBB"]


def miller_steps():
    """kinds of the Miller-loop steps ('D' doubling, 'A' addition) over the bits of |x| below the top one"""
    steps = []
    for b in range(62, -1, -1):
        steps.append("D")
        if (X_ABS >> b) & 1:
            steps.append("A")
    return steps


def driver_sequence():
    """the segment calls of one whole check: list of (segment name, line index or None) + the copies / spills the driver does
    itself, as tuples ('COPY', dst, src) / ('SPILL', k, reg) / ('FILL', reg, k) with reg in 'A', 'B'"""
    steps = miller_steps()
    seq = []
    for j, k in enumerate(steps):
        seq.append((("S_0" if j == 0 else "S_" + k), j))
    # f = conj(f) (x < 0), then the final exponentiation of tower.cuh final_exponentiation()
    seq.append(("FE_CONJA", None))
    seq += [("FE_INV", None), ("FE_CONJA", None), ("FE_MUL", None)]                 # f^(p^6 - 1)
    seq += [("COPY", "B", "A"), ("FE_FROBB", None), ("FE_FROBB", None), ("FE_MUL", None)]  # ^(p^2 + 1)
    seq.append(("SPILL", 0, "A"))                                                          # G0 = f

    def pow_x():  # RA <- RA^x (x < 0), clobbers RB
        out = [("COPY", "B", "A")]
        for b in range(62, -1, -1):
            out.append(("FE_CYC", None))
            if (X_ABS >> b) & 1:
                out.append(("FE_MUL", None))
        out.append(("FE_CONJA", None))
        return out
    seq += pow_x()[:1]
    seq += pow_x()[1:] + [("FE_CONJB", None), ("FE_MUL", None)]      # t0 = f^x * conj(f)   (RB still holds f)
    seq += pow_x() + [("FE_CONJB", None), ("FE_MUL", None)]          # t1 = t0^x * conj(t0)
    seq += pow_x() + [("FE_FROBB", None), ("FE_MUL", None)]          # t2 = t1^x * t1^p
    seq.append(("SPILL", 1, "A"))                                    # G1 = t2
    seq += pow_x() + pow_x()                                         # t2^(x^2)
    seq += [("FILL", "B", 1), ("FE_FROBB", None), ("FE_FROBB", None), ("FE_MUL", None)]
    seq += [("FILL", "B", 1), ("FE_CONJB", None), ("FE_MUL", None)]  # t3
    seq.append(("SPILL", 1, "A"))
    seq += [("FILL", "A", 0), ("COPY", "B", "A"), ("FE_MUL", None), ("FE_MUL", None)]  # f^3
    seq += [("FILL", "B", 1), ("FE_MUL", None)]
    return seq


# ---------------------------------------------------------------------------------------------- simulation (validation)
def simulate_segment(p, streams, slots, line, inputs, consts):
    """execute the per-role streams level by level on Python integers; slots: dict slot -> Fp2"""
    pcs = [0] * len(streams)
    done = False
    while not done:
        writes = {}
        for r, s in enumerate(streams):
            X = Y = None
            while True:
                op, arg = s[pcs[r]]
                pcs[r] += 1
                if op == "BAR":
                    break
                if op == "END":
                    done = True
                    break
                if op == "LDX": X = slots[arg]
                elif op == "ADDX": X = f2add(X, slots[arg])
                elif op == "SUBX": X = f2sub(X, slots[arg])
                elif op == "LDY": Y = slots[arg]
                elif op == "ADDY": Y = f2add(Y, slots[arg])
                elif op == "SUBY": Y = f2sub(Y, slots[arg])
                elif op == "STX":
                    assert arg not in writes or writes[arg][0] == r, "two roles write one slot in one level"
                    writes[arg] = (r, X)
                    slots[arg] = X  # same-role later reads see it; cross-role same-level reads are excluded by the schedule (checked below)
                elif op == "LDXK": X = consts[arg]
                elif op == "LDYK": Y = consts[arg]
                elif op == "LDXL": X = line[arg]
                elif op == "ADDXL": X = f2add(X, line[arg])
                elif op == "LDXIN": X = inputs[arg]
                elif op == "ADDXIN": X = f2add(X, inputs[arg])
                elif op == "SUBXIN": X = f2sub(X, inputs[arg])
                elif op == "LDYIN": Y = inputs[arg]
                elif op == "ADDYIN": Y = f2add(Y, inputs[arg])
                elif op == "LDYS": Y = (inputs[0][arg], 0)
                elif op == "STXL": line[arg] = X
                elif op == "MUL": X = f2mul(X, Y)
                elif op == "SQR": X = f2mul(X, X)
                elif op == "XI": X = f2xi(X)
                elif op == "NEGX": X = f2neg(X)
                elif op == "DBLX": X = f2add(X, X)
                elif op == "TPLX": X = f2add(f2add(X, X), X)
                elif op == "CONJX": X = f2conj(X)
                elif op == "INVX": X = f2inv(X)
                else:
                    raise AssertionError(op)


def check_hazards(p):
    """no slot is written by one role and read or written by another role within the same level"""
    for lv in range(1, p.n_levels + 1):
        w, r = {}, defaultdict(set)
        for n in p.nodes:
            if n.level != lv:
                continue
            if getattr(n, "line_out", None) is None:
                assert n.out.slot not in w or w[n.out.slot] == n.role, (p.name, lv, "write/write")
                w[n.out.slot] = n.role
            for v in n.inputs():
                if v.slot is not None and (v.loc is None or v.loc[0] == "S"):
                    r[v.slot].add(n.role)
        for s, role in w.items():
            assert r[s] <= {role}, (p.name, lv, s, "read/write across roles")
    # a temporary's slot must not be re-used while it is live: allocate() guarantees it; a fixed slot is written only after its
    # last reader (schedule()).  Same-role read-after-write inside a level follows program order.


def simulate_check(segs, streams, pk, sig, hm, n_slots):
    """whole check on integers -> RA (six Fp2).  pk: G1 affine ints, sig / hm: G2 affine ((x0,x1),(y0,y1))."""
    from oracle.pyref import bls12_381 as B
    consts = constants()
    # the prepared lines of hm: by the VM's own preparation segments (P_D / P_A, what k_g2_prepare_vm runs)
    fi = FIXED_INDEX
    lines = []
    pslots = {i: (0, 0) for i in range(n_slots)}
    pslots[fi["TX"]], pslots[fi["TY"]], pslots[fi["TZ"]] = hm[0], hm[1], (1, 0)
    for k in miller_steps():
        line = [None, None, None]
        simulate_segment(segs["P_" + k], streams["P_" + k], pslots, line, [None, hm[0], hm[1]], consts)
        assert None not in line
        lines.append(tuple(line))
    slots = {i: (0, 0) for i in range(n_slots)}
    G = {0: None, 1: None}
    slots[fi["TX"]], slots[fi["TY"]], slots[fi["TZ"]] = sig[0], sig[1], (1, 0)
    inputs = [(pk[0], pk[1]), sig[0], sig[1]]
    for item in driver_sequence():
        if item[0] == "COPY":
            for i in range(6):
                slots[fi[item[1] + str(i)]] = slots[fi[item[2] + str(i)]]
        elif item[0] == "SPILL":
            G[item[1]] = [slots[fi[item[2] + str(i)]] for i in range(6)]
        elif item[0] == "FILL":
            for i in range(6):
                slots[fi[item[1] + str(i)]] = G[item[2]][i]
        else:
            name, li = item
            simulate_segment(segs[name], streams[name], slots, lines[li] if li is not None else None, inputs, consts)
    return [slots[fi[f"A{i}"]] for i in range(6)]


def self_check(R, verbose=True):
    from oracle.pyref import bls12_381 as B
    segs, n_slots = build_segments(R)
    streams = {}
    for name, p in segs.items():
        check_hazards(p)
        streams[name] = p.emit()
    if verbose:
        for name in SEG_ORDER:
            p = segs[name]
            loads = [sum(n.cost() for n in p.nodes if n.role == r) for r in range(R)]
            crit = sum(max((sum(n.cost() for n in p.nodes if n.role == r and n.level == lv) for r in range(R)), default=0)
                       for lv in range(1, p.n_levels + 1))
            print(f"{name:9s} levels {p.n_levels:2d} nodes {len(p.nodes):3d} words {sum(len(s) for s in streams[name]):5d} "
                  f"work {sum(loads) / 888:6.1f} mul-eq, critical path {crit / 888:5.1f}, balance {sum(loads) / (R * crit):.2f}")
        for name in SEG_ORDER:
            p = segs[name]
            print(f"{name:9s} slots {max(n.out.slot for n in p.nodes) + 1}")
        print("slots per check:", n_slots, "=", n_slots * 96, "B; per block of 32 checks:", n_slots * 96 * 32 / 1024, "KB")
    # a valid signature, a wrong one, random points
    sk = 0x1234567890ABCDEF1234567890ABCDEF % B.R
    msg = b"Sign with new partial key"
    hm = B.hash_to_g2(msg)
    pk = B.g1_mul(B.G1, sk)
    sig = B.g2_mul(hm, sk)
    one = [(1, 0)] + [(0, 0)] * 5
    got = simulate_check(segs, streams, pk, sig, hm, n_slots)
    assert got == one, "valid signature must give 1"
    sig2 = B.g2_mul(hm, sk + 1)
    got = simulate_check(segs, streams, pk, sig2, hm, n_slots)
    assert got != one
    # the value itself: e(pk, hm)^3 * e(-G, sig2)^3 in the convention of tower.cuh
    e1 = B.pairing(pk, hm)
    e2 = B.pairing((B.G1[0], (-B.G1[1]) % P), sig2)
    want = B.f12_pow(B.f12_mul(e1, e2), 3)
    flat = [want[0][0], want[0][1], want[0][2], want[1][0], want[1][1], want[1][2]]
    assert got == flat, "pairing product differs from the Python restatement"
    if verbose:
        print("simulation against oracle/pyref: OK (valid -> 1, invalid -> e(pk,H)^3 e(-G,sig)^3)")
    return segs, streams, n_slots


# ---------------------------------------------------------------------------------------------- output
def mont_words(v):
    v = v * (1 << 384) % P
    return [(v >> (32 * i)) & 0xFFFFFFFF for i in range(12)]


def write_inc(path, R):
    segs, streams, n_slots = self_check(R, verbose=False)
    words, table = [], []
    for name in SEG_ORDER:
        offs = []
        for r in range(R):
            offs.append(len(words))
            for op, arg in streams[name][r]:
                assert 0 <= arg < (1 << 16)
                words.append(OP[op] | (arg << 8))
        table.append(offs)
    out = ["// GENERATED by tools/gen_pairing_vm.py - do not edit (python tools/gen_pairing_vm.py --write)",
           f"// pairing VM program for R = {R} roles; {len(words)} instruction words, {n_slots} Fp2 slots per check",
           "#pragma once",
           f"#define PVM_R {R}",
           f"#define PVM_SLOTS {n_slots}",
           f"#define PVM_N_FIXED {N_FIXED}",
           f"#define PVM_N_WORDS {len(words)}",
           f"#define PVM_N_CONSTS {len(constants())}",
           "enum PvmOp : uint32_t { " + ", ".join(f"PVM_{n} = {i}" for i, n in enumerate(OPS)) + " };",
           "enum PvmSeg : uint32_t { " + ", ".join(f"SEG_{n} = {i}" for i, n in enumerate(SEG_ORDER)) + ", SEG_COUNT };",
           "enum PvmSlot : uint32_t { " + ", ".join(f"SLOT_{n} = {i}" for n, i in sorted(FIXED_INDEX.items(), key=lambda kv: (kv[1], kv[0]))) + " };"]
    out.append("PVM_CONST uint32_t pvm_seg_start[SEG_COUNT][PVM_R] = {" + ", ".join("{" + ", ".join(str(o) for o in offs) + "}" for offs in table) + "};")
    out.append("PVM_CONST uint32_t pvm_prog[PVM_N_WORDS] = {")
    for i in range(0, len(words), 16):
        out.append("  " + ", ".join("0x%06xu" % w for w in words[i:i + 16]) + ",")
    out.append("};")
    out.append("// Fp2 constants in Montgomery form (c0 then c1, 12 limbs each): 1, -3 xG, -2 yG, -xG, -yG, Frobenius coefficients")
    out.append("PVM_CONST uint32_t pvm_consts[PVM_N_CONSTS][24] = {")
    for c in constants():
        out.append("  {" + ", ".join("0x%08xu" % w for w in mont_words(c[0]) + mont_words(c[1])) + "},")
    out.append("};")
    # the driver's call sequence, run-length friendly: (kind, a, b): kind 0 = segment a with line index b (0xffff: none),
    # 1 = COPY reg a <- reg b, 2 = SPILL global a <- reg b, 3 = FILL reg a <- global b   (reg: 0 = RA, 1 = RB)
    seq = []
    reg = {"A": 0, "B": 1}
    for item in driver_sequence():
        if item[0] == "COPY":
            seq.append((1, reg[item[1]], reg[item[2]]))
        elif item[0] == "SPILL":
            seq.append((2, item[1], reg[item[2]]))
        elif item[0] == "FILL":
            seq.append((3, reg[item[1]], item[2]))
        else:
            seq.append((0, SEG_ORDER.index(item[0]), 0xFFFF if item[1] is None else item[1]))
    prep = [(0, SEG_ORDER.index("P_" + k), j) for j, k in enumerate(miller_steps())]
    out.append(f"#define PVM_N_PREP_CALLS {len(prep)}")
    out.append("PVM_CONST uint32_t pvm_prep_calls[PVM_N_PREP_CALLS] = {" + ", ".join("0x%07xu" % (k | (a << 4) | (b << 12)) for k, a, b in prep) + "};")
    out.append(f"#define PVM_N_CALLS {len(seq)}")
    out.append("PVM_CONST uint32_t pvm_calls[PVM_N_CALLS] = {")
    enc = [k | (a << 4) | (b << 12) for k, a, b in seq]
    for i in range(0, len(enc), 16):
        out.append("  " + ", ".join("0x%07xu" % w for w in enc[i:i + 16]) + ",")
    out.append("};")
    with open(path, "w") as f:
        f.write("\n".join(out) + "\n")
    # executed work of one whole check, for the roofline of bench.py: Fp2 products / squarings / barriers over the driver's call list
    import json
    cnt = defaultdict(int)
    for item in driver_sequence():
        if item[0] in ("COPY", "SPILL", "FILL"):
            cnt["copies"] += 1
            continue
        st = streams[item[0]]
        for r in range(R):
            for op, _arg in st[r]:
                if op in ("MUL", "SQR", "INVX"):
                    cnt[op] += 1
        cnt["BAR"] += sum(1 for op, _ in st[0] if op == "BAR")
        cnt["segments"] += 1
    # wide multiply-accumulates (32x32->64): an Fp2 product = two fused sums of two products (2 x 444), an Fp2 squaring = two
    # products (2 x 300); the inversion (binary extended Euclid) runs on the ALU pipe + 6 products
    macs = cnt["MUL"] * 888 + cnt["SQR"] * 600 + cnt["INVX"] * 6 * 300
    meta = {"roles": R, "slots": n_slots, "words": len(words), "fp2_mul_per_check": cnt["MUL"], "fp2_sqr_per_check": cnt["SQR"],
            "fp2_inv_per_check": cnt["INVX"], "barriers_per_check": cnt["BAR"], "segments_per_check": cnt["segments"],
            "wide_macs_per_check": macs, "fp_mul_equivalents_per_check": macs / 300}
    with open(os.path.splitext(path)[0] + ".json", "w") as f:
        json.dump(meta, f, indent=1)
        f.write("\n")
    return len(words), n_slots


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--roles", type=int, default=6)
    ap.add_argument("--check", action="store_true", help="simulate the program against oracle/pyref and print the schedule")
    ap.add_argument("--write", action="store_true", help="write dvt_circuits_b200/csrc/pairing_prog.inc")
    ap.add_argument("--out", default=None, help="other output path (experiments with another number of roles)")
    a = ap.parse_args()
    if a.check or not a.write:
        self_check(a.roles)
    if a.write:
        nw, ns = write_inc(a.out or os.path.join(ROOT, "dvt_circuits_b200", "csrc", "pairing_prog.inc"), a.roles)
        print(f"wrote pairing_prog.inc: {nw} words, {ns} slots")
