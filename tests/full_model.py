"""numpy model of the FULL (un-reduced, symmetrised) Newton system in the FullLayout order of
ipm-zoo_b200/csrc/ipmz_device.cuh (TEST INFRASTRUCTURE).

Reference: the un-reduced system of SymbolicOptimization.cpp:417-433 with the residual
definitions of :480-492.  tests/test_full_model.py checks on the CPU that solving this system
gives the oracle's augmented Newton step (so the layout, signs and symmetrisation are right before
any GPU time is spent); tests/test_gpu_parity.py compares the CUDA assembly entry by entry to it.
"""
import numpy as np

import oracle_lib as ol


def _sides(p):
    ylo = p.var_bounds in (ol.LOWER, ol.BOTH)
    zup = p.var_bounds in (ol.UPPER, ol.BOTH)
    mi = p.m_ineq if p.ineq_bounds != ol.NONE else 0
    me = p.m_eq if p.equalities else 0
    ilo = mi > 0 and p.ineq_bounds in (ol.LOWER, ol.BOTH)
    iup = mi > 0 and p.ineq_bounds in (ol.UPPER, ol.BOTH)
    return ylo, zup, mi, me, ilo, iup


def layout(p):
    ylo, zup, mi, me, ilo, iup = _sides(p)
    n, m = p.n, mi + me
    hasl = m > 0 and (ilo or me > 0)
    hasu = m > 0 and (iup or me > 0)
    o, L = 0, {}
    for name, size in (("y", n if ylo else 0), ("z", n if zup else 0), ("sl", m if hasl else 0),
                       ("su", m if hasu else 0), ("ly", n if ylo else 0), ("lz", n if zup else 0),
                       ("ll", m if hasl else 0), ("lu", m if hasu else 0), ("s", m), ("x", n), ("lam", m)):
        L[name] = o
        o += size
    L["N"] = o
    return L


def stacked(p, it):
    """Row-stacked (inequalities then equalities) views of the packed iterate."""
    ylo, zup, mi, me, ilo, iup = _sides(p)
    off = p.offsets()
    g = lambda k: it[off[k][0]:off[k][0] + off[k][1]]
    cat = lambda a, b: np.concatenate([g(a)[:mi] if mi else np.zeros(0), g(b)[:me] if me else np.zeros(0)])
    M = np.vstack([p.A if mi else np.zeros((0, p.n)), p.C if me else np.zeros((0, p.n))])
    lo = np.concatenate([p.l_A if mi else np.zeros(0), p.d if me else np.zeros(0)])
    up = np.concatenate([p.u_A if mi else np.zeros(0), p.d if me else np.zeros(0)])
    has_lo = np.concatenate([np.full(mi, bool(ilo)), np.full(me, True)])
    has_up = np.concatenate([np.full(mi, bool(iup)), np.full(me, True)])
    return dict(x=g("x"), lam=cat("lamA", "lamC"), s=cat("s", "t"), ll=cat("lamg", "lamv"), lu=cat("lamh", "lamw"),
                sl=cat("g", "v"), su=cat("h", "w"), ly=g("lamy"), lz=g("lamz"), y=g("y"), z=g("z"),
                M=M, lo=lo, up=up, has_lo=has_lo, has_up=has_up)


def full_system(p, it, mu=0.0):
    """K (N x N symmetric) and right-hand side of the symmetrised full Newton system at iterate `it`."""
    ylo, zup, mi, me, ilo, iup = _sides(p)
    n, m = p.n, mi + me
    v = stacked(p, it)
    L = layout(p)
    K = np.zeros((L["N"], L["N"]))
    b = np.zeros(L["N"])
    I = np.arange(n)
    x, lam, s = v["x"], v["lam"], v["s"]
    rx = p.c + p.Q @ x
    if zup:
        rx = rx + v["lz"]
    if m:
        rx = rx + v["M"].T @ lam
    if ylo:
        rx = rx - v["ly"]
    ox, olam, os_ = L["x"], L["lam"], L["s"]
    K[ox:ox + n, ox:ox + n] = p.Q
    b[ox:ox + n] = -rx
    if ylo:
        y, ly = v["y"], v["ly"]
        K[L["y"] + I, L["y"] + I] = ly / y
        K[L["y"] + I, L["ly"] + I] = 1.0
        K[L["ly"] + I, L["y"] + I] = 1.0
        K[L["ly"] + I, ox + I] = -1.0
        K[ox + I, L["ly"] + I] = -1.0
        b[L["y"]:L["y"] + n] = -(y * ly - mu) / y
        b[L["ly"]:L["ly"] + n] = -((p.l_x + y) - x)
    if zup:
        z, lz = v["z"], v["lz"]
        K[L["z"] + I, L["z"] + I] = lz / z
        K[L["z"] + I, L["lz"] + I] = 1.0
        K[L["lz"] + I, L["z"] + I] = 1.0
        K[L["lz"] + I, ox + I] = 1.0
        K[ox + I, L["lz"] + I] = 1.0
        b[L["z"]:L["z"] + n] = -(z * lz - mu) / z
        b[L["lz"]:L["lz"] + n] = -((x + z) - p.u_x)
    if m:
        J = np.arange(m)
        K[olam:olam + m, ox:ox + n] = v["M"]
        K[ox:ox + n, olam:olam + m] = v["M"].T
        K[olam + J, os_ + J] = -1.0
        K[os_ + J, olam + J] = -1.0
        b[olam:olam + m] = -(v["M"] @ x - s)
        rs = -lam.copy()
        for side, sgn, has in (("l", -1.0, v["has_lo"]), ("u", 1.0, v["has_up"])):
            if not has.any():
                continue
            sk, lk = v["s" + side], v["l" + side]
            osk, olk = L["s" + side], L["l" + side]
            for j in J:
                if has[j]:
                    K[osk + j, osk + j] = lk[j] / sk[j]
                    K[osk + j, olk + j] = K[olk + j, osk + j] = 1.0
                    K[olk + j, os_ + j] = K[os_ + j, olk + j] = sgn
                    b[osk + j] = -(sk[j] * lk[j] - mu) / sk[j]
                    b[olk + j] = -((v["lo"][j] + sk[j]) - s[j]) if side == "l" else -((sk[j] + s[j]) - v["up"][j])
                    rs[j] += sgn * lk[j]
                else:  # decoupled identity rows
                    K[osk + j, osk + j] = 1.0
                    K[olk + j, olk + j] = 1.0
        b[os_:os_ + m] = -rs
    return K, b, L


def augmented_part(p, u, L):
    ylo, zup, mi, me, ilo, iup = _sides(p)
    return np.concatenate([u[L["x"]:L["x"] + p.n], u[L["lam"]:L["lam"] + mi + me]])
