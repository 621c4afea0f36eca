#!/usr/bin/env python
"""Extracts the Butcher tableaus of the reference's explicit Runge-Kutta methods
(/root/reference/src/time_steppers/RKMethods.jl) into tests/golden/rk_tableaus.json, as exact rationals ("p/q") or --
for HEM5, whose entries contain sqrt(6) -- as floats.  The methods built by linear solves (SSP104, rSSPs2, rSSPs3)
are not literal tableaus and are skipped.  Run in the build container (the GPU boxes have no /root/reference):

    python tests/golden/make_rk_tableaus.py
"""
import json
import math
import os
import re
from fractions import Fraction

SRC = "/root/reference/src/time_steppers/RKMethods.jl"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "rk_tableaus.json")


def ev(expr, A=None):
    """Evaluate a Julia scalar expression made of integers, `//`, + - * /, sqrt() and A[i, j] references."""
    e = expr.strip()
    ns = {"Fraction": Fraction, "sqrt": math.sqrt}

    def ref(m):
        key = f"A_{m.group(1)}_{m.group(2)}"
        ns[key] = A[int(m.group(1)) - 1][int(m.group(2)) - 1]
        return key

    e = re.sub(r"A\[(\d+),\s*(\d+)\]", ref, e)
    e = re.sub(r"(\d+)\s*//\s*(\d+)", r"Fraction(\1,\2)", e)
    return eval(e, ns)


def split_commas(txt):
    out, depth, cur = [], 0, ""
    for ch in txt:
        depth += ch in "([" 
        depth -= ch in ")]"
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    out.append(cur)
    return [x for x in out if x.strip()]


def split_row(row):
    """Entries of a matrix row: whitespace separated, parenthesised expressions kept together."""
    out, depth, cur = [], 0, ""
    for ch in row.strip():
        if ch == "(":
            depth += 1
        if ch == ")":
            depth -= 1
        if ch.isspace() and depth == 0:
            if cur:
                out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur:
        out.append(cur)
    # a leading sign belongs to the entry ("-3//8"); binary operators never stand alone inside these literals
    return out


def enc(v):
    if isinstance(v, float):
        return repr(v)
    v = Fraction(v)
    return f"{v.numerator}/{v.denominator}"


def main():
    src = open(SRC).read()
    out = {}
    for m in re.finditer(r"^function (\w+)\((?:s = \d+)?; kwargs\.\.\.\)\n(.*?)^end", src, flags=re.S | re.M):
        name, body = m.group(1), m.group(2)
        if "\\" in body or "I(" in body:
            continue  # built by a linear solve
        s = None
        ms = re.search(r"^\s*s = (\d+)", body, flags=re.M)
        if ms:
            s = int(ms.group(1))
        # A
        mA = re.search(r"^\s*A = \[(.*?)\]\n", body, flags=re.S | re.M)
        if mA:
            txt = mA.group(1)
            rows = [r for r in re.split(r";|\n", txt) if r.strip()]
            A = [[ev(x) for x in split_row(r)] for r in rows]
        elif re.search(r"A = fill\(0, 1, 1\)", body):
            A = [[Fraction(0)]]
        else:
            mz = re.search(r"A = zeros\(Rational, (\d+), (\d+)\)", body)
            assert mz, name
            n = int(mz.group(1))
            A = [[Fraction(0)] * n for _ in range(n)]
            for ma in re.finditer(r"^\s*A\[(\d+), (\d+)\] = (.*)$", body, flags=re.M):
                A[int(ma.group(1)) - 1][int(ma.group(2)) - 1] = ev(ma.group(3), A)
        n = len(A)
        assert all(len(r) == n for r in A), (name, A)
        # b
        mb = re.search(r"^\s*b = \[(.*?)\]$", body, flags=re.S | re.M)
        if mb:
            b = [ev(x, A) for x in split_commas(mb.group(1).replace("\n", " "))]
        else:
            mf = re.search(r"b = fill\((.*?), s\)", body)
            assert mf and s, name
            b = [ev(mf.group(1))] * s
        # c
        if re.search(r"c = sum\(eachcol\(A\)\)", body):
            c = [sum(r) for r in A]
        else:
            mc = re.search(r"^\s*c = \[(.*?)\]$", body, flags=re.M)
            assert mc, name
            c = [ev(x, A) for x in split_commas(mc.group(1))]
        mr = re.search(r"^\s*r = (.*)$", body, flags=re.M)
        r = ev(mr.group(1)) if mr else Fraction(0)
        out[name] = dict(A=[[enc(v) for v in row] for row in A], b=[enc(v) for v in b], c=[enc(v) for v in c], r=enc(r))
    json.dump(out, open(OUT, "w"), indent=1, sort_keys=True)
    print(len(out), "tableaus:", " ".join(sorted(out)))


if __name__ == "__main__":
    main()
