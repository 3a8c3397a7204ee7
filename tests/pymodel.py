"""A second, independent restatement of the reference loop in plain Python (small cases only).

Purpose: pin the C++ oracle against transcription errors.  This file was written from the
reference source (src/J_Space.jl:320-416, 556-613, 638-842) and DESIGN.md §3 (the draw contract),
not from oracle/jspace_oracle.cpp; tests/test_oracle_model.py requires both to produce the same
event log, bit for bit, on small lattices.  Method 1 (random drivers) with all three contact rules
and single-entry excision; Julia's data structures are mirrored literally (lists with `remove`,
linear `in` tests).  Python floats are IEEE doubles and CPython never fuses multiply-add, so the
arithmetic is reproducible.
"""
from __future__ import annotations

import math
import struct

M32 = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for rnd in range(10):
        if rnd:
            k0 = (k0 + 0x9E3779B9) & M32
            k1 = (k1 + 0xBB67AE85) & M32
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & M32, p1 & M32, ((p0 >> 32) ^ c3 ^ k1) & M32, p0 & M32
    return c0, c1, c2, c3


class Draws:
    def __init__(self, seed, stream):
        self.key = (seed & M32, (seed >> 32) & M32)
        self.stream = stream

    def block(self, it, attempt, site, seq):
        o = philox4x32_10((it & M32, ((it >> 32) & 0xFFFF) | (attempt << 16), self.stream, (site << 28) | (seq & 0x0FFFFFFF)),
                          self.key)
        return o[0] | (o[1] << 32), o[2] | (o[3] << 32)

    def index(self, it, site, seq, half, s):
        attempt = 0
        w = self.block(it, attempt, site, seq)[half]
        m = w * s
        lo = m & 0xFFFFFFFFFFFFFFFF
        if lo < s:
            t = ((1 << 64) - s) % s
            while lo < t:
                attempt += 1
                w = self.block(it, attempt, site, seq)[half]
                m = w * s
                lo = m & 0xFFFFFFFFFFFFFFFF
        return m >> 64


def u53(w):
    return float(w >> 11) * 2.0 ** -53


def contract_log(x):
    ln2_hi, ln2_lo = float.fromhex("0x1.62e42fee00000p-1"), float.fromhex("0x1.a39ef35793c76p-33")
    Lg = [float.fromhex(h) for h in ("0x1.5555555555593p-1", "0x1.999999997fa04p-2", "0x1.2492494229359p-2",
                                     "0x1.c71c51d8e78afp-3", "0x1.7466496cb03dep-3", "0x1.39a09d078c69fp-3",
                                     "0x1.2f112df3e5244p-3")]
    bits = struct.unpack("<Q", struct.pack("<d", x))[0]
    hx, lx = bits >> 32, bits & M32
    hx = (hx + 0x3ff00000 - 0x3fe6a09e) & M32
    k = (hx >> 20) - 0x3ff
    hx = (hx & 0x000fffff) + 0x3fe6a09e
    x = struct.unpack("<d", struct.pack("<Q", (hx << 32) | lx))[0]
    f = x - 1.0
    hfsq = (0.5 * f) * f
    s = f / (2.0 + f)
    z = s * s
    w = z * z
    t1 = w * (Lg[1] + w * (Lg[3] + w * Lg[5]))
    t2 = z * (Lg[0] + w * (Lg[2] + w * (Lg[4] + w * Lg[6])))
    R = t2 + t1
    dk = float(k)
    return (((s * (hfsq + R) + dk * ln2_lo) - hfsq) + f) + dk * ln2_hi


def grid_neighbors(row, col):
    """Graphs.grid((row, col, 1)): vertex = x + (y-1)*row, neighbours ascending."""
    nb = {}
    for y in range(1, col + 1):
        for x in range(1, row + 1):
            v = x + (y - 1) * row
            lst = []
            if y > 1:
                lst.append(v - row)
            if x > 1:
                lst.append(v - 1)
            if x < row:
                lst.append(v + 1)
            if y < col:
                lst.append(v + row)
            nb[v] = lst
    return nb


def cumsum_pairwise(v):
    """Base.accumulate_pairwise!(add_sum, ...) of Julia 1.7."""
    n = len(v)
    c = [0.0] * n
    c[0] = v[0]

    def rec(s, i1, m):
        if m < 128:
            s_ = v[i1]
            c[i1] = s + s_
            for i in range(i1 + 1, i1 + m):
                s_ = s_ + v[i]
                c[i] = s + s_
            return s_
        n2 = m >> 1
        sl = rec(s, i1, n2)
        return sl + rec(s + sl, i1 + n2, m - n2)

    if n > 1:
        rec(v[0], 1, n - 1)
    return c


def simulate(row, col, Tf, rate_birth, rate_death, rate_migration, mu, adv_mean, adv_std, model, seed, stream,
             t_bottleneck=(), ratio_bottleneck=()):
    """Returns the event log as tuples (type, time, cell, parent, site, site2, clone, label, flags)."""
    D = Draws(seed, stream)
    nbrs = grid_neighbors(row, col)
    nv = row * col
    v0 = col * (row // 2) + (col // 2 + 1) if nv % 2 == 0 else nv // 2 + 1     # :131-133
    ident = {v0: 1}            # :id
    subpop = {v0: 1}           # :Subpop
    next_id = 2
    cs_alive = [v0]
    n = 1
    set_mut = [(0, 1)]         # (parent clone, label)
    alpha = [rate_birth]
    ca_subpop = [[v0]]
    used = {1}
    df = []
    t = 0.0
    it = 0
    id_b = 1 if t_bottleneck else 0

    def death(cell, time, flags):
        df.append((2, time, ident[cell], 0, cell, 0, subpop[cell], 0, flags))
        del ident[cell], subpop[cell]

    while (t < Tf and n > 0) or (rate_death == 0 and n == nv):
        if not (t < Tf and n > 0):
            break
        it += 1
        a_sub = [alpha[i] * float(len(ca_subpop[i])) for i in range(len(alpha))]
        birth = a_sub[0]
        for x in a_sub[1:]:
            birth = birth + x
        dth = rate_death * float(n)
        mig = rate_migration * float(n)
        lam = birth + dth + mig
        wa, wb = D.block(it, 0, 0, 0)
        t_old = t
        t = t + (-contract_log((float(wa >> 12) + 0.5) * 2.0 ** -52)) * (1.0 / lam)
        prob = cumsum_pairwise([x / lam for x in a_sub] + [dth / lam, mig / lam])
        k = u53(wb)
        cls = next((i + 1 for i, p in enumerate(prob) if k <= p), 0)
        S = len(alpha)
        if cls == 0:
            pass
        elif cls == S + 2:                                             # migration :725-742
            cell = cs_alive[D.index(it, 1, 0, 0, n)]
            nb = nbrs[cell]
            pos = nb[D.index(it, 1, 0, 1, len(nb))]
            if pos not in cs_alive:
                sp = subpop[cell]
                ident[pos], subpop[pos] = ident.pop(cell), subpop.pop(cell)
                df.append((3, t, ident[pos], 0, pos, cell, sp, 0, 1))
                cs_alive.append(pos)
                ca_subpop[sp - 1].append(pos)
                cs_alive.remove(cell)
                ca_subpop[sp - 1].remove(cell)
        elif cls == S + 1:                                             # death :744-759
            cell = cs_alive[D.index(it, 1, 0, 0, n)]
            sp = subpop[cell]
            death(cell, t, 1)
            cs_alive.remove(cell)
            ca_subpop[sp - 1].remove(cell)
            n -= 1
        elif ca_subpop[cls - 1]:                                       # birth :761-810
            members = ca_subpop[cls - 1]
            cell = members[D.index(it, 1, 0, 0, len(members))]
            nb = nbrs[cell]
            pos = nb[D.index(it, 1, 0, 1, len(nb))]
            occupied = pos in cs_alive
            rule = True
            if model == "contact":
                rule = not occupied
            elif model == "voter":
                rule = not (occupied and subpop[cell] == subpop[pos])
            elif model in ("hvoter", "h_voter"):
                rule = not (occupied and alpha[subpop[cell] - 1] <= alpha[subpop[pos] - 1])
            if rule:
                wd, wcoin = D.block(it, 0, 4, 0)
                r = u53(wd)
                driver = not (r > mu)
                sp = subpop[cell]
                if driver and len(set_mut) >= 1000:
                    break
                flag = 1
                if pos in ident:                                        # :335-338
                    ca_subpop[subpop[pos] - 1].remove(pos)
                    death(pos, t, 2 | flag)
                    flag = 0
                nid, nid2 = next_id, next_id + 1
                next_id += 2
                parent = ident[cell]
                if not driver:                                          # :351-363
                    df.append((0, t, nid, parent, cell, 0, sp, 0, flag))
                    df.append((0, t, nid2, parent, pos, 0, sp, 0, 0))
                    ident[cell] = nid
                    ident[pos], subpop[pos] = nid2, sp
                    ca_subpop[cls - 1].append(pos)
                else:                                                   # :365-414
                    free = [m for m in range(1, 1001) if m not in used]
                    label = free[D.index(it, 5, 0, 0, len(free))]
                    used.add(label)
                    attempt = 0
                    while True:
                        a, b = D.block(it, 0, 6, attempt)
                        v1, v2 = 2.0 * u53(a) - 1.0, 2.0 * u53(b) - 1.0
                        s = v1 * v1 + v2 * v2
                        if s >= 1.0 or s == 0.0:
                            attempt += 1
                            continue
                        z = v1 * math.sqrt((-2.0 * contract_log(s)) / s)
                        break
                    alpha.append(alpha[sp - 1] + abs(adv_mean + adv_std * z))
                    set_mut.append((sp, label))
                    nsp = len(set_mut)
                    coin = u53(wcoin) < 0.5
                    df.append((1, t, nid2, parent, pos if coin else cell, 0, nsp, label, flag))
                    df.append((0, t, nid, parent, cell if coin else pos, 0, sp, 0, 0))
                    if coin:
                        ident[cell] = nid
                        ident[pos], subpop[pos] = nid2, nsp
                        ca_subpop.append([pos])
                    else:
                        ident[pos], subpop[pos] = nid, sp
                        ident[cell], subpop[cell] = nid2, nsp
                        ca_subpop[cls - 1].append(pos)
                        ca_subpop[cls - 1].remove(cell)
                        ca_subpop.append([cell])
                if not occupied:
                    cs_alive.append(pos)
                    n += 1
        if id_b:                                                        # excision :813-838 (index never advances)
            tb = t_bottleneck[id_b - 1]
            if t_old < tb <= t:
                kk = int(math.trunc(float(n) * ratio_bottleneck[id_b - 1]))
                q = 0
                taken = []
                if kk == 1:
                    taken = [cs_alive[D.index(it, 7, 0, 0, n)]]
                elif kk == 2:
                    i1 = D.index(it, 7, 0, 0, n) + 1
                    i2 = D.index(it, 7, 1, 0, n - 1) + 1
                    taken = [cs_alive[i1 - 1], cs_alive[(n if i2 == i1 else i2) - 1]]
                elif kk > 2 and n < kk * 24:
                    inds = list(range(n))
                    for i in range(kk):
                        j = i + D.index(it, 7, i, 0, n - i)
                        inds[i], inds[j] = inds[j], inds[i]
                        taken.append(cs_alive[inds[i]])
                elif kk > 2:
                    seen = set()
                    for i in range(kk):
                        while True:
                            idx = D.index(it, 7, q, 0, n)
                            q += 1
                            if idx not in seen:
                                break
                        seen.add(idx)
                        taken.append(cs_alive[idx])
                for i, ct in enumerate(taken):
                    sp = subpop[ct]
                    death(ct, tb, 4 | (1 if i == 0 else 0))
                    cs_alive.remove(ct)
                    ca_subpop[sp - 1].remove(ct)
                    n -= 1
    return df, alpha, set_mut, it, t
