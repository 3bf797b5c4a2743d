"""Restatement of src/curves/multimult.ts (Relation, MultiMult, Bos-Coster heap).

TEST INFRASTRUCTURE (oracle) — see oracle/__init__.py.
"""
from __future__ import annotations


class Pair:
    # multimult.ts:19-29
    __slots__ = ('pt', 'scalar')

    def __init__(self, pt, scalar):
        self.pt, self.scalar = pt, scalar

    def cmp(self, b) -> int:
        return self.scalar.cmp(b.scalar)


class MultiMult:
    # multimult.ts:31-90
    def __init__(self, g):
        self.group = g
        self.pairs = []
        self.known = []

    def add_known(self, pt):
        self.group.is_compat_point(pt)
        if not any(pt.eq(x[0]) for x in self.known):
            self.pairs.append(Pair(pt, self.group.new_scalar(0)))
            self.known.append((pt, len(self.pairs) - 1))

    def insert(self, pt, s):
        self.group.is_compat_point(pt)
        self.group.is_compat_scalar(s)
        for kpt, idx in self.known:
            if pt.eq(kpt):
                self.pairs[idx].scalar = self.pairs[idx].scalar.add(s)
                return
        self.pairs.append(Pair(pt, s))

    def evaluate(self):
        # multimult.ts:61-89 (Bos-Coster)
        pairs = self.pairs
        if len(pairs) == 0:
            return self.group.identity()
        if len(pairs) == 1:
            return pairs[0].pt.mul(pairs[0].scalar)
        _heapify(pairs)
        while True:
            if len(pairs) == 1:
                return pairs[0].pt.mul(pairs[0].scalar)
            a = _extract_max(pairs)
            b = pairs[0]
            if b.scalar.is_zero():
                return a.pt.mul(a.scalar)
            c = Pair(a.pt, a.scalar.sub(b.scalar))
            d = Pair(b.pt.add(a.pt), b.scalar)
            pairs[0] = d
            if not c.scalar.is_zero():
                pairs.append(c)
                _bubbleup(pairs, len(pairs))

    def evaluate_naive(self):
        """Not in the reference: plain sum of pt*scalar (test cross-check)."""
        acc = self.group.identity()
        for pr in self.pairs:
            acc = acc.add(pr.pt.mul(pr.scalar))
        return acc


def _extract_max(arr):
    # multimult.ts:92-103
    arr[0], arr[-1] = arr[-1], arr[0]
    mx = arr.pop()
    _pushdown(arr, 1)
    return mx


def _heapify(arr):
    for i in range(len(arr)):
        _bubbleup(arr, i + 1)


def _bubbleup(arr, index):
    # 1-based, iterative form of multimult.ts:111-123
    while index > 1:
        parent = index // 2
        if arr[parent - 1].cmp(arr[index - 1]) < 0:
            arr[parent - 1], arr[index - 1] = arr[index - 1], arr[parent - 1]
            index = parent
        else:
            return


def _pushdown(arr, parent):
    # multimult.ts:125-145
    while True:
        son, daughter = 2 * parent, 2 * parent + 1
        if son > len(arr):
            return
        child = son
        if daughter <= len(arr) and arr[daughter - 1].cmp(arr[son - 1]) > 0:
            child = daughter
        if arr[parent - 1].cmp(arr[child - 1]) < 0:
            arr[parent - 1], arr[child - 1] = arr[child - 1], arr[parent - 1]
            parent = child
        else:
            return


class Relation:
    # multimult.ts:147-174
    def __init__(self, g, tape):
        self.group = g
        self.pairs = []
        self.tape = tape

    def insert_m(self, pts, scalars):
        if len(pts) != len(scalars):
            raise ValueError('arrays are not the same length')
        for pt, s in zip(pts, scalars):
            self.insert(pt, s)

    def insert(self, pt, s):
        self.group.is_compat_point(pt)
        self.group.is_compat_scalar(s)
        self.pairs.append(Pair(pt, s))

    def drain(self, m: MultiMult):
        # ONE fresh random scalar per relation (multimult.ts:168-173)
        randomizer = self.group.random_scalar(self.tape)
        for pr in self.pairs:
            m.insert(pr.pt, pr.scalar.mul(randomizer))
