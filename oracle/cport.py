"""ORACLE -- test / measurement infrastructure, NOT product code (see oracle/ops.py header for who may import this package).

ctypes front end of oracle/cport.c: the 3-D linearised step of oracle/stepper.py (PCG mode, direct problem) with every hot
loop in C / OpenMP, so that the CPU baseline beside the GPU numbers uses all host cores at compiled-code speed.  The shared
object is built on first use with the host's own gcc (`-O3 -march=native -fopenmp`) into oracle/_build/, keyed by the
CPU's feature flags (the build container and the GPU box may differ).
"""
from __future__ import annotations

import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

from .ops import SEM
from .stepper import AB, BD

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _p(a):
    return None if a is None else a.ctypes.data_as(_dp)


def _pi(a):
    return a.ctypes.data_as(_ip)


def _cpu_key() -> str:
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                return hashlib.sha1(line.encode()).hexdigest()[:10]
    except OSError:
        pass
    return "generic"


def build(force: bool = False) -> str:
    out_dir = os.path.join(HERE, "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, f"libcport_{_cpu_key()}.so")
    src = os.path.join(HERE, "cport.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        cmd = ["gcc", "-O3", "-march=native", "-fopenmp", "-fPIC", "-shared", "-std=c99", "-o", so, src, "-lm"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("gcc failed for oracle/cport.c:\n" + r.stderr)
    return so


def set_threads(n: int) -> int:
    """Use n OpenMP threads (n <= 0: every core the process may run on); returns the count in effect."""
    lib = C.CDLL(build())
    if n <= 0:
        try:
            n = len(os.sched_getaffinity(0))
        except AttributeError:
            n = os.cpu_count() or 1
    lib.c_set_threads(int(n))
    return int(lib.c_max_threads())


class _PMG(C.Structure):
    _fields_ = [("nv", C.c_int), ("nagg", C.c_int), ("S", _dp), ("deninv", _dp), ("phi", _dp), ("vid", _ip), ("voff", _ip),
                ("vent", _ip), ("d1inv", _dp), ("agg", _ip), ("A2inv", _dp), ("rc", _dp), ("xv", _dp), ("ra", _dp), ("x2", _dp)]


class _PRES(C.Structure):
    _fields_ = [("nel", C.c_int), ("N", C.c_int), ("L", C.c_int), ("nseg", C.c_int), ("J12", _dp), ("D12", _dp), ("RW2", _dp),
                ("mbinv", _dp), ("bm2inv", _dp), ("dinvE", _dp), ("seg_off", _ip), ("seg_idx", _ip), ("vol2", C.c_double)]


class _HELM(C.Structure):
    _fields_ = [("nel", C.c_int), ("N", C.c_int), ("nseg", C.c_int), ("D", _dp), ("G6", _dp), ("bm1", _dp), ("dinv", _dp),
                ("mult", _dp), ("binv", _dp), ("seg_off", _ip), ("seg_idx", _ip), ("vol", C.c_double)]


def _c(a, dtype=np.float64):
    return np.ascontiguousarray(a, dtype=dtype)


class CPort:
    """Operators and solvers of one 3-D SEM object in C.  `pmg`: an oracle.pmg.PMG instance or None (Jacobi)."""

    def __init__(self, s: SEM, pmg=None):
        assert s.ldim == 3, "oracle/cport.c covers the 3-D path"
        self.s, self.lib = s, C.CDLL(build())
        self.lib.c_pressure_pcg.restype = C.c_int
        self.lib.c_helmholtz_pcg.restype = C.c_int
        self.nel, self.N, self.L, self.M = s.nel, s.lx1, s.lx2, s.lxd
        self.n, self.n2, self.nd = s.nel * s.lx1 ** 3, s.nel * s.lx2 ** 3, s.nel * s.lxd ** 3
        self.D, self.J12, self.D12, self.Jd, self.Dd = _c(s.D), _c(s.J12), _c(s.D12), _c(s.Jd), _c(s.Dd)
        order = [(0, 0), (1, 1), (2, 2), (0, 1), (0, 2), (1, 2)]
        self.G6 = _c(np.stack([s.G[i, j].ravel() for i, j in order]))
        self.bm1, self.binv, self.mult = _c(s.bm1.ravel()), _c(s.binv.ravel()), _c(s.mult.ravel())
        self.mask = _c(s.mask.reshape(3, -1))
        self.mbinv = _c(self.mask * self.binv[None, :])
        self.RW2 = _c(np.stack([(s.R2[i, c] * s.w32).ravel() for i in range(3) for c in range(3)]))
        self.Rd = _c(np.stack([s.Rd[i, c].ravel() for i in range(3) for c in range(3)]))
        self.bm2inv = _c(1.0 / s.bm2.ravel())
        self.dinvE = _c(1.0 / s.e_diag().ravel())
        # dssum segments: copies of every global node with more than one local copy
        g = s.glo.ravel()
        o = np.argsort(g, kind="stable")
        gs = g[o]
        starts = np.flatnonzero(np.r_[True, gs[1:] != gs[:-1]])
        lens = np.diff(np.r_[starts, gs.size])
        keep = lens > 1
        self.seg_off = _c(np.r_[0, np.cumsum(lens[keep])], np.int32)
        idx = [o[a:a + l] for a, l in zip(starts[keep], lens[keep])]
        self.seg_idx = _c(np.concatenate(idx) if idx else np.zeros(0), np.int32)
        self.nseg = int(keep.sum())
        self.pres = _PRES(self.nel, self.N, self.L, self.nseg, _p(self.J12), _p(self.D12), _p(self.RW2), _p(self.mbinv),
                          _p(self.bm2inv), _p(self.dinvE), _pi(self.seg_off), _pi(self.seg_idx), float(s.vol2))
        self._helm_cache = {}
        self.pm = None
        if pmg is not None:
            m = pmg
            nv = m.nv
            vid = _c(m.vid, np.int32)
            cnt = np.bincount(vid.ravel(), minlength=nv)
            voff = _c(np.r_[0, np.cumsum(cnt)], np.int32)
            vent = _c(np.argsort(vid.ravel(), kind="stable"), np.int32)
            self._pm_keep = dict(S=_c(m.S), deninv=_c(m.deninv), phi=_c(m.phi.reshape(8, -1)), vid=vid, voff=voff, vent=vent,
                                 d1inv=_c(1.0 / m.d1), agg=_c(m.agg, np.int32), A2inv=_c(m.A2inv), rc=np.zeros(self.nel * 8),
                                 xv=np.zeros(nv), ra=np.zeros(m.nagg), x2=np.zeros(m.nagg))
            k = self._pm_keep
            self.pm = _PMG(nv, m.nagg, _p(k["S"]), _p(k["deninv"]), _p(k["phi"]), _pi(k["vid"]), _pi(k["voff"]), _pi(k["vent"]),
                           _p(k["d1inv"]), _pi(k["agg"]), _p(k["A2inv"]), _p(k["rc"]), _p(k["xv"]), _p(k["ra"]), _p(k["x2"]))
        # work arrays
        self.w3 = np.zeros(3 * self.n)
        self.wp = [np.zeros(self.n2) for _ in range(4)]
        self.wv = [np.zeros(self.n) for _ in range(3)]

    # ------------------------------------------------------------------ operators (for the parity tests)
    def axhelm(self, u, h1, h2):
        u = _c(u).ravel()
        w = np.zeros(self.n)
        self.lib.c_axhelm(self.nel, self.N, _p(self.D), _p(self.G6), _p(self.bm1), C.c_double(h1), C.c_double(h2), 1, _p(u), _p(w),
                          None, 0, C.c_longlong(self.n))
        return w

    def dssum(self, u, nf=1):
        u = _c(u).ravel().copy()
        self.lib.c_dssum(self.nseg, _pi(self.seg_off), _pi(self.seg_idx), _p(u), nf, C.c_longlong(self.n))
        return u

    def opgradt(self, p):
        p = _c(p).ravel()
        w = np.zeros(3 * self.n)
        self.lib.c_gradt(self.nel, self.N, self.L, _p(self.J12), _p(self.D12), _p(self.RW2), _p(p), _p(w), C.c_longlong(self.n),
                         C.c_longlong(self.n2))
        return w.reshape(3, -1)

    def opdiv(self, u):
        u = _c(u).ravel()
        q = np.zeros(self.n2)
        self.lib.c_div(self.nel, self.N, self.L, _p(self.J12), _p(self.D12), _p(self.RW2), _p(u), None, _p(q), C.c_double(1.0),
                       C.c_longlong(self.n), C.c_longlong(self.n2))
        return q

    def cdabdtp(self, p):
        p = _c(p).ravel()
        ep = np.zeros(self.n2)
        self.lib.c_apply_E(C.byref(self.pres), _p(p), _p(self.w3), _p(ep))
        return ep

    def advab_direct(self, up, ub, spng=None):
        up, ub = _c(up).ravel(), _c(ub).ravel()
        f = np.zeros(3 * self.n)
        sp = None if spng is None else _c(spng).ravel()
        self.lib.c_advab_direct(self.nel, self.N, self.M, _p(self.Jd), _p(self.Dd), _p(self.Rd), _p(self.bm1), _p(sp), _p(up), _p(ub),
                                _p(f), C.c_longlong(self.n), C.c_longlong(self.nd))
        return f.reshape(3, -1)

    def pmg_apply(self, r):
        r = _c(r).ravel()
        z = np.zeros(self.n2)
        self.lib.c_pmg_apply(self.nel, self.L, C.byref(self.pm), _p(r), _p(z))
        return z

    # ------------------------------------------------------------------ solvers
    def pressure_pcg(self, g, tol, maxit):
        g = _c(g).ravel().copy()
        x = np.zeros(self.n2)
        it = self.lib.c_pressure_pcg(C.byref(self.pres), C.byref(self.pm) if self.pm is not None else None, _p(g), _p(x),
                                     _p(self.wp[0]), _p(self.wp[1]), _p(self.wp[2]), _p(self.w3), C.c_double(tol), int(maxit))
        return x, it

    def _helm(self, h1, h2):
        key = (h1, h2)
        if key not in self._helm_cache:
            dinv = _c(1.0 / self.s.helm_diag(h1, h2).ravel())
            self._helm_cache[key] = (dinv, _HELM(self.nel, self.N, self.nseg, _p(self.D), _p(self.G6), _p(self.bm1), _p(dinv),
                                                 _p(self.mult), _p(self.binv), _pi(self.seg_off), _pi(self.seg_idx), float(self.s.vol)))
        return self._helm_cache[key][1]

    def helmholtz_pcg(self, rhs, h1, h2, tol, maxit):
        hs = self._helm(h1, h2)
        rhs = _c(rhs).reshape(3, -1).copy()
        out, its = np.zeros((3, self.n)), []
        for c in range(3):
            its.append(self.lib.c_helmholtz_pcg(C.byref(hs), _p(self.mask[c]), C.c_double(h1), C.c_double(h2), _p(rhs[c]), _p(out[c]),
                                                _p(self.wv[0]), _p(self.wv[1]), C.c_double(tol), int(maxit)))
        return out, its


class CStepper:
    """oracle/stepper.py LinearizedStepper.linearized_map (direct problem, PCG mode) on the C port."""

    def __init__(self, s: SEM, ubase, re, spng_fun=None, tol_v=1e-9, tol_p=1e-7, max_iter_v=1000, max_iter_p=20000, ifvcor=False,
                 pmg=None):
        self.s, self.cp = s, CPort(s, pmg)
        self.ub = _c(ubase).reshape(3, -1)
        self.h1 = 1.0 / re
        self.spng = None if spng_fun is None else _c(spng_fun).ravel()
        self.tol_v, self.tol_p, self.max_iter_v, self.max_iter_p = tol_v, tol_p, max_iter_v, max_iter_p
        self.ifvcor = bool(ifvcor)
        self.iters_v, self.iters_p = [], []

    def linearized_map(self, v, p, nsteps, dt, budget_s=None):
        """nsteps of the direct perturbation stepper from (v, p), order ramp restarted (core/matvec.f:216).  budget_s: stop
        after that many seconds of wall time (at a step boundary, at least one step); self.steps_done / self.step_seconds
        report what ran -- used by bench.py to bound the CPU baseline."""
        import time as _time
        t_start = _time.perf_counter()
        self.steps_done, self.step_seconds = 0, []
        cp, lib = self.cp, self.cp.lib
        n, n2 = cp.n, cp.n2
        u = _c(v).reshape(3, n).copy()
        pr = _c(p).ravel().copy()
        ulag = [np.zeros((3, n)), np.zeros((3, n))]
        flag = [np.zeros((3, n)), np.zeros((3, n))]
        plag = np.zeros(n2)
        b, r, w3 = np.zeros((3, n)), np.zeros((3, n)), np.zeros((3, n))
        for istep in range(1, nsteps + 1):
            t_step = _time.perf_counter()
            if budget_s is not None and istep > 1 and t_step - t_start > budget_s:
                break
            k = min(istep, 3)
            bd, ab = np.array(BD[k] + [0.0] * 3), np.array(AB[k] + [0.0] * 3)
            h2 = BD[k][0] / dt
            f = cp.advab_direct(u, self.ub, self.spng)
            lib.c_make_rhs(C.c_longlong(n), k, _p(ab), _p(bd), C.c_double(dt), _p(cp.bm1), _p(f), _p(flag[0]), _p(flag[1]), _p(u),
                           _p(ulag[0]), _p(ulag[1]), _p(b))
            pt = 2.0 * pr - plag if k == 3 else pr.copy()
            lib.c_gradt(cp.nel, cp.N, cp.L, _p(cp.J12), _p(cp.D12), _p(cp.RW2), _p(pt), _p(r), C.c_longlong(n), C.c_longlong(n2))
            lib.c_axhelm(cp.nel, cp.N, _p(cp.D), _p(cp.G6), _p(cp.bm1), C.c_double(self.h1), C.c_double(h2), 3, _p(u), _p(r), _p(b), 1,
                         C.c_longlong(n))                                    # r = b + D^T pt - H u
            lib.c_dssum(cp.nseg, _pi(cp.seg_off), _pi(cp.seg_idx), _p(r), 3, C.c_longlong(n))
            lib.c_mask3(C.c_longlong(n), _p(cp.mask), _p(r))
            du, its = cp.helmholtz_pcg(r, self.h1, h2, self.tol_v, self.max_iter_v)
            self.iters_v.append(its)
            uh = u + du
            g = np.zeros(n2)
            lib.c_div(cp.nel, cp.N, cp.L, _p(cp.J12), _p(cp.D12), _p(cp.RW2), _p(uh), None, _p(g), C.c_double(-1.0), C.c_longlong(n),
                      C.c_longlong(n2))
            if self.ifvcor:
                g -= g.mean()
            phi, itp = cp.pressure_pcg(g, self.tol_p, self.max_iter_p)
            if self.ifvcor:
                phi -= phi.mean()
            self.iters_p.append(itp)
            lib.c_gradt(cp.nel, cp.N, cp.L, _p(cp.J12), _p(cp.D12), _p(cp.RW2), _p(phi), _p(w3), C.c_longlong(n), C.c_longlong(n2))
            lib.c_dssum(cp.nseg, _pi(cp.seg_off), _pi(cp.seg_idx), _p(w3), 3, C.c_longlong(n))
            unew, pnew = np.zeros((3, n)), np.zeros(n2)
            lib.c_final_update(C.c_longlong(n), C.c_longlong(n2), _p(u), _p(du), _p(cp.mbinv), _p(w3), _p(unew), _p(pt), _p(phi),
                               C.c_double(h2), _p(pnew))
            ulag = [u, ulag[0]]
            flag = [f, flag[0]]
            plag = pr
            u, pr = unew, pnew
            self.steps_done += 1
            self.step_seconds.append(_time.perf_counter() - t_step)
        return u, pr
