"""ctypes binding of libnekstab_b200.so (include/nekstab_b200.h) -- the stand-in for the Fortran ISO_C_BINDING
shim (nekstab_b200/fortran/nekstab_b200_c.f90) in this Fortran-less image.  Every method maps 1:1 to a C-ABI
entry point; arrays are host numpy float64 buffers exactly as a Nek5000 executable would pass its COMMON arrays.

There is deliberately NO fallback: a missing library or a missing GPU raises.
"""
from __future__ import annotations

import ctypes as C
import glob
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libnekstab_b200.so")

DIRECT, ADJOINT, DIRECT_ADJOINT, NEWTON, FORCE_SENS = 1, 2, 3, 4, 5

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)
_lp = C.POINTER(C.c_longlong)


class NsbError(RuntimeError):
    pass


class Stats(C.Structure):
    _fields_ = [("steps", C.c_longlong), ("helm_iters", C.c_longlong), ("pres_iters", C.c_longlong),
                ("kernel_launches", C.c_longlong), ("step_ms", C.c_double)]


def _arr(a, dtype=np.float64):
    if a is None:
        return None
    b = np.ascontiguousarray(a, dtype=dtype)
    return b


def _p(a):
    if a is None:
        return None
    if a.dtype == np.float64:
        return a.ctypes.data_as(_dp)
    if a.dtype == np.int64:
        return a.ctypes.data_as(_lp)
    if a.dtype == np.int32:
        return a.ctypes.data_as(_ip)
    raise TypeError(a.dtype)


_SIGS = {
    "nsb_comm_unique_id": [C.c_char_p],
    "nsb_comm_init": [C.c_int, C.c_int, C.c_char_p, C.c_int],
    "nsb_init": [C.c_int] * 5 + [C.c_longlong] + [_dp] * 6 + [_lp, C.c_int],
    "nsb_finalize": [],
    "nsb_set_params": [C.c_double] * 4 + [C.c_int] * 2,
    "nsb_set_weights": [_dp],
    "nsb_set_baseflow": [_dp] * 3,
    "nsb_set_sponge": [_dp],
    "nsb_set_floquet": [C.c_int, _dp],
    "nsb_set_upo": [C.c_int],
    "nsb_set_scalar": [C.c_int, C.c_double, C.c_double, _dp, C.c_double, C.c_int],
    "nsb_set_scalar_base": [_dp],
    "nsb_vec_upload_scalar": [C.c_int, _dp],
    "nsb_vec_download_scalar": [C.c_int, _dp],
    "nsb_op_conv_scalar": [_dp] * 5,
    "nsb_vec_set_time": [C.c_int, C.c_double],
    "nsb_vec_get_time": [C.c_int, _dp],
    "nsb_set_dns_sponge": [C.c_double, _dp, _dp, _dp],
    "nsb_get_orbit": [C.c_int, _dp, _dp, _dp],
    "nsb_prepare_linearized_solver": [C.c_double, C.c_double, _dp, _ip, _dp],
    "nsb_set_timestep": [C.c_double, C.c_int],
    "nsb_set_ifvcor": [C.c_int, C.c_int],
    "nsb_set_projection": [C.c_int],
    "nsb_set_step_callback": [C.c_void_p, C.c_void_p],
    "nsb_set_pressure_preconditioner": [C.c_int, C.c_int],
    "nsb_op_pc_apply": [C.c_int, _dp, _dp],
    "nsb_pc_get": [C.c_int, _dp, _lp],
    "nsb_set_adjoint_masks": [_dp] * 3,
    "nsb_vec_alloc": [C.c_int],
    "nsb_vec_upload": [C.c_int] + [_dp] * 4,
    "nsb_vec_download": [C.c_int] + [_dp] * 4,
    "nsb_vec_copy": [C.c_int, C.c_int],
    "nsb_vec_zero": [C.c_int],
    "nsb_vec_cmult": [C.c_int, C.c_double],
    "nsb_vec_add2": [C.c_int, C.c_int],
    "nsb_vec_sub2": [C.c_int, C.c_int],
    "nsb_vec_inner_product": [C.c_int, C.c_int, _dp],
    "nsb_vec_norm": [C.c_int, _dp],
    "nsb_vec_normalize": [C.c_int, _dp],
    "nsb_basis_gemv": [C.c_int, C.c_int, _dp, C.c_int],
    "nsb_basis_gemv_complex": [C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int],
    "nsb_basis_rotate": [C.c_int, C.c_int, _dp, C.c_int],
    "nsb_biorthogonalize": [C.c_int] * 4,
    "nsb_wave_maker": [C.c_int] * 4 + [_dp],
    "nsb_orthonormalize": [C.c_int, C.c_int, C.c_int, _dp],
    "nsb_matvec": [C.c_int] * 3,
    "nsb_nonlinear_forward_map": [C.c_int, C.c_int],
    "nsb_prepare_solver_from_slot": [C.c_int, C.c_double, C.c_double, _dp, _ip, _dp],
    "nsb_newton_krylov": [C.c_int] * 6 + [C.c_double] * 3 + [C.c_int, C.c_int, _ip, _dp, _dp, _lp],
    "nsb_get_stats": [C.POINTER(Stats), C.c_int],
    "nsb_profile": [C.c_int, _dp, _lp],
    "nsb_fp64_peak": [_dp],
    "nsb_op_axhelm": [_dp, C.c_double, C.c_double, _dp],
    "nsb_op_dssum": [_dp],
    "nsb_op_glsc3": [_dp] * 4,
    "nsb_op_opgradt": [_dp] * 4,
    "nsb_op_opdiv": [_dp] * 4,
    "nsb_op_cdabdtp": [_dp] * 2,
    "nsb_op_advab": [C.c_int] + [_dp] * 6,
    "nsb_op_hmholtz": [_dp] * 6 + [C.c_double, C.c_double, _ip],
    "nsb_op_esolver": [_dp, _dp, _ip],
    "nsb_op_cfl": [_dp] * 3 + [C.c_double, _dp],
    "nsb_get_field": [C.c_char_p, _dp, _lp],
    "nsb_gs_host_candidates": [C.c_int, C.c_int, C.c_int, _lp, _lp, _lp],
    "nsb_gs_host_plan": [C.c_int, C.c_int, _lp, _lp, _ip],
    "nsb_gs_host_get": [C.c_int, _ip],
    "nsb_pm_host_aggregates": [C.c_int, C.c_int, _dp, C.c_int, _ip],
    "nsb_pm_host_colouring": [C.c_int, C.c_int, _lp, _ip, _ip],
    "nsb_pm_host_fdm_1d": [C.c_int, C.c_double, C.c_double, _dp, _dp],
    "nsb_pm_host_spd_inverse": [C.c_int, _dp],
    "nsb_arnoldi_factorization": [C.c_int, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int],
    "nsb_krylov_schur": [C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_int, _dp, _dp, _dp, _dp, _ip, _ip, C.c_int],
    "nsb_schur_condensation": [_ip, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_double],
    "nsb_select_eigenvalues": [_ip, _ip, _dp, _dp, C.c_double, C.c_int, C.c_int],
    "nsb_ts_gmres": [C.c_int] * 7 + [C.c_double, _ip, _dp],
    "nsb_lapack_eig": [_dp, C.c_int, _dp, _dp, _dp],
    "nsb_lapack_schur": [_dp, C.c_int, _dp, _dp, _dp],
    "nsb_lapack_ordschur": [_dp, _dp, _ip, C.c_int],
    "nsb_lapack_lstsq": [_dp, _dp, _dp, C.c_int, C.c_int],
    "nsb_lapack_load": [C.c_char_p],
}
EXPORTED = sorted(list(_SIGS) + ["nsb_last_error", "nsb_n", "nsb_n2"])

_lib = None
_comm = None          # (rank, nranks) once the process-wide NCCL communicator exists (a unique id is single-use)


def find_lapack() -> Optional[str]:
    """A shared library exporting LP64 dgeev_/dgees_/dtrsen_/dgels_ (scipy's bundled OpenBLAS in this image)."""
    env = os.environ.get("NSB_LAPACK_LIB")
    if env:
        return env
    try:
        import scipy
        cands = glob.glob(os.path.join(os.path.dirname(scipy.__file__), "..", "scipy.libs", "libscipy_openblas-*.so"))
        cands = [c for c in cands if "64_" not in os.path.basename(c)]
        if cands:
            return os.path.abspath(cands[0])
    except Exception:
        pass
    return None


def load_library(path: str = LIB_PATH):
    """dlopen the C-ABI library (no CUDA call is made here, so this works on a GPU-less build box)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise NsbError(f"{path} not found: run `python -m nekstab_b200.build` (no CPU fallback exists)")
    # Python hosts only: PyTorch bundles its own, newer libnccl.so.2.  The dynamic loader keeps ONE library per SONAME, so if this
    # library pulled in the system NCCL first, a later `import torch` in the same process fails (undefined symbol ncclDevCommCreate).
    # Importing torch first makes both use torch's NCCL (a superset of the API used here).  A Fortran host never loads torch.
    if os.environ.get("NSB_NO_TORCH_PRELOAD") != "1":
        try:
            import torch  # noqa: F401
        except Exception:
            pass
    lib = C.CDLL(path, mode=C.RTLD_GLOBAL)
    for name, args in _SIGS.items():
        fn = getattr(lib, name)
        fn.argtypes = args
        fn.restype = C.c_int
    lib.nsb_last_error.restype = C.c_char_p
    lib.nsb_last_error.argtypes = []
    lib.nsb_n.restype = C.c_longlong
    lib.nsb_n2.restype = C.c_longlong
    _lib = lib
    return lib


def _ck(rc):
    if rc != 0:
        raise NsbError(_lib.nsb_last_error().decode())


class NekStabB200:
    """One process-wide solver context on one GPU (mirrors the COMMON-block state of a Nek5000+nekStab rank)."""

    def __init__(self, case, device: int = 0, rank: int = 0, nranks: int = 1, nccl_id: Optional[bytes] = None):
        self.lib = load_library()
        lp = find_lapack()
        if lp:
            _ck(self.lib.nsb_lapack_load(lp.encode()))
        global _comm
        if _comm != (rank, nranks):
            if nranks > 1:
                assert nccl_id is not None and len(nccl_id) == 128
                _ck(self.lib.nsb_comm_init(rank, nranks, nccl_id, device))
            else:
                _ck(self.lib.nsb_comm_init(0, 1, b"\0" * 128, device))
            _comm = (rank, nranks)
        self.case = case
        self.ldim, self.lx1, self.lx2, self.lxd = case.ldim, case.lx1, case.lx1 - 2, 3 * case.lx1 // 2
        self.nel = case.nel
        self.np1, self.np2 = self.lx1 ** self.ldim, self.lx2 ** self.ldim
        self.n, self.n2 = self.nel * self.np1, self.nel * self.np2
        xyz = [_arr(case.xyz[d]) for d in range(self.ldim)] + [None] * (3 - self.ldim)
        msk = [_arr(case.mask[d]) for d in range(self.ldim)] + [None] * (3 - self.ldim)
        glo = _arr(case.glo, np.int64)
        nelg = int(case.nelg or case.nel)
        self.nelg = nelg
        _ck(self.lib.nsb_init(self.ldim, self.lx1, self.lxd, self.lx2, self.nel, nelg, _p(xyz[0]), _p(xyz[1]), _p(xyz[2]),
                              _p(msk[0]), _p(msk[1]), _p(msk[2]), _p(glo), device))
        _ck(self.lib.nsb_set_params(1.0 / case.re, 1.0, case.tol_v, case.tol_p, 0, 0))
        ub = [_arr(case.ubase[d]) for d in range(self.ldim)] + [None] * (3 - self.ldim)
        _ck(self.lib.nsb_set_baseflow(_p(ub[0]), _p(ub[1]), _p(ub[2])))
        if case.spng_fun is not None:
            sp = _arr(case.spng_fun)
            _ck(self.lib.nsb_set_sponge(_p(sp)))
            bm1 = self.get_field("bm1")
            bm1s = np.where(sp.ravel() != 0.0, 0.0, bm1)          # core/usr_extra.f:116-118
            _ck(self.lib.nsb_set_weights(_p(_arr(bm1s))))
        if isinstance(case.extra, dict) and "mask_adjoint" in case.extra:
            ma = [_arr(case.extra["mask_adjoint"][d]) for d in range(self.ldim)] + [None] * (3 - self.ldim)
            _ck(self.lib.nsb_set_adjoint_masks(_p(ma[0]), _p(ma[1]), _p(ma[2])))

        iv = getattr(case, "ifvcor", None)
        iva = getattr(case, "ifvcor_adjoint", None)
        _ck(self.lib.nsb_set_ifvcor(-1 if iv is None else int(iv), -1 if iva is None else int(iva)))

    # ---- setup
    def set_params(self, viscosity, density=1.0, tol_v=1e-9, tol_p=1e-7, maxit_v=0, maxit_p=0):
        _ck(self.lib.nsb_set_params(viscosity, density, tol_v, tol_p, maxit_v, maxit_p))

    def prepare_linearized_solver(self, end_time, cfl_target=0.5):
        dt, ns, ct = C.c_double(), C.c_int(), C.c_double()
        _ck(self.lib.nsb_prepare_linearized_solver(end_time, cfl_target, C.byref(dt), C.byref(ns), C.byref(ct)))
        return dt.value, ns.value, ct.value

    def set_pressure_preconditioner(self, kind, nagg=0):
        """0: Jacobi (north-star); 1: FDM element blocks + Q1 vertex-mesh Jacobi + aggregate coarse solve (csrc/pmg.cu)."""
        _ck(self.lib.nsb_set_pressure_preconditioner(int(kind), int(nagg)))

    def op_pc_apply(self, r, adjoint=False):
        r = _arr(r).ravel(); z = np.empty(self.n2)
        _ck(self.lib.nsb_op_pc_apply(int(adjoint), _p(r), _p(z)))
        return z

    def pc_get(self, which):
        cnt = C.c_longlong(0)
        _ck(self.lib.nsb_pc_get(which, None, C.byref(cnt)))
        out = np.zeros(cnt.value)
        _ck(self.lib.nsb_pc_get(which, _p(out), C.byref(cnt)))
        return out

    STEP_CB = C.CFUNCTYPE(None, C.c_int, C.c_double, C.c_void_p)

    def set_step_callback(self, fn):
        """fn(istep, time) is called on the host before every step of a matvec (nekstab_usrchk, core/matvec.f:221); None removes it."""
        if fn is None:
            self._step_cb = None
            _ck(self.lib.nsb_set_step_callback(None, None))
            return
        self._step_cb = self.STEP_CB(lambda istep, t, _u: fn(istep, t))        # keep a reference: ctypes callbacks must outlive the call
        _ck(self.lib.nsb_set_step_callback(C.cast(self._step_cb, C.c_void_p), None))

    def set_floquet(self, enable, pbase=None):
        """Floquet / UPO mode: base-flow co-evolution + orbit storage (core/matvec.f:187-236)."""
        pb = None if pbase is None else _arr(pbase).reshape(self.n2)
        _ck(self.lib.nsb_set_floquet(int(enable), _p(pb)))

    def set_upo(self, enable):
        """Newton-GMRES for UPOs (uparam(1) = 2.1): time component of the Krylov vectors, orbit storage in the nonlinear map, border
        terms of newton_linearized_map (core/matvec.f:407-419)."""
        _ck(self.lib.nsb_set_upo(int(enable)))

    def vec_set_time(self, slot, t):
        _ck(self.lib.nsb_vec_set_time(int(slot), float(t)))

    def vec_get_time(self, slot):
        t = C.c_double()
        _ck(self.lib.nsb_vec_get_time(int(slot), C.byref(t)))
        return t.value

    def set_scalar(self, enable, conductivity=0.0, rhocp=1.0, tmask=None, ri=0.0, gdir=1):
        """Scalar transport (`ifheat`): theta travels in the Krylov vectors ([v|theta|pr]); discards the slots (vec_alloc again)."""
        tm = None if tmask is None else _arr(tmask).reshape(self.n)
        _ck(self.lib.nsb_set_scalar(int(enable), float(conductivity), float(rhocp), _p(tm), float(ri), int(gdir)))
        self.scalar_on = bool(enable)

    def set_scalar_base(self, tbase):
        _ck(self.lib.nsb_set_scalar_base(_p(_arr(tbase).reshape(self.n))))

    def vec_upload_scalar(self, slot, theta):
        _ck(self.lib.nsb_vec_upload_scalar(int(slot), _p(_arr(theta).reshape(self.n))))

    def vec_download_scalar(self, slot):
        t = np.empty(self.n)
        _ck(self.lib.nsb_vec_download_scalar(int(slot), _p(t)))
        return t

    def op_conv_scalar(self, a, phi):
        a3 = self._vec3(a)
        out = np.empty(self.n)
        _ck(self.lib.nsb_op_conv_scalar(_p(a3[0]), _p(a3[1]), _p(a3[2]), _p(_arr(phi).reshape(self.n)), _p(out)))
        return out

    def set_dns_sponge(self, spng_str, ref=None):
        r = [None] * 3 if ref is None else self._vec3(ref)
        _ck(self.lib.nsb_set_dns_sponge(float(spng_str), _p(r[0]), _p(r[1]), _p(r[2])))

    def get_orbit(self, istep):
        u, us = self._out3()
        _ck(self.lib.nsb_get_orbit(int(istep), _p(us[0]), _p(us[1]), _p(us[2])))
        return u

    def set_projection(self, mxprev):
        _ck(self.lib.nsb_set_projection(int(mxprev)))

    def set_timestep(self, dt, nsteps):
        _ck(self.lib.nsb_set_timestep(dt, nsteps))

    def get_field(self, name):
        cnt = C.c_longlong()
        _ck(self.lib.nsb_get_field(name.encode(), None, C.byref(cnt)))
        out = np.empty(cnt.value)
        _ck(self.lib.nsb_get_field(name.encode(), _p(out), C.byref(cnt)))
        return out

    def stats(self, reset=False):
        s = Stats()
        _ck(self.lib.nsb_get_stats(C.byref(s), int(reset)))
        return {k: getattr(s, k) for k, _ in Stats._fields_}

    PROFILE_KINDS = ["pcg_gradt", "dssum", "pcg_div", "pcg_update", "hcg_axhelm", "hcg_update", "advab", "hcg_dssum",
                     "pcg_pc_restrict", "pcg_pc_coarse", "pcg_pc_apply", "orth_multidot", "orth_multiaxpy", "pcg_fused"]

    def profile(self, enable=-1):
        ms = np.zeros(16); cnt = np.zeros(16, dtype=np.int64)
        _ck(self.lib.nsb_profile(enable, _p(ms), _p(cnt)))
        return {k: (float(ms[i]), int(cnt[i])) for i, k in enumerate(self.PROFILE_KINDS)}

    def fp64_peak(self):
        t = C.c_double()
        _ck(self.lib.nsb_fp64_peak(C.byref(t)))
        return t.value

    # ---- krylov vectors
    def vec_alloc(self, nslots):
        _ck(self.lib.nsb_vec_alloc(nslots))

    def vec_upload(self, slot, v, p=None):
        v = _arr(v).reshape(self.ldim, self.n)
        comps = [np.ascontiguousarray(v[d]) for d in range(self.ldim)] + [None] * (3 - self.ldim)
        pp = None if p is None else _arr(p).reshape(self.n2)
        _ck(self.lib.nsb_vec_upload(slot, _p(comps[0]), _p(comps[1]), _p(comps[2]), _p(pp)))

    def vec_download(self, slot, out=None):
        """Device slot -> host arrays; `out` = (v, p) reuses caller-owned (e.g. pinned) buffers of shapes (ldim, n) and (n2,)."""
        if out is None:
            v = np.empty((self.ldim, self.n))
            p = np.empty(self.n2)
        else:
            v, p = out
            assert v.shape == (self.ldim, self.n) and p.shape == (self.n2,) and v.flags.c_contiguous and v.dtype == np.float64
        comps = [v[d] for d in range(self.ldim)] + [None] * (3 - self.ldim)
        _ck(self.lib.nsb_vec_download(slot, _p(comps[0]), _p(comps[1]), _p(comps[2]), _p(p)))
        return v, p

    def vec_copy(self, dst, src): _ck(self.lib.nsb_vec_copy(dst, src))
    def vec_zero(self, s): _ck(self.lib.nsb_vec_zero(s))
    def vec_cmult(self, s, a): _ck(self.lib.nsb_vec_cmult(s, float(a)))
    def vec_add2(self, p, q): _ck(self.lib.nsb_vec_add2(p, q))
    def vec_sub2(self, p, q): _ck(self.lib.nsb_vec_sub2(p, q))

    def inner_product(self, p, q):
        a = C.c_double()
        _ck(self.lib.nsb_vec_inner_product(p, q, C.byref(a)))
        return a.value

    def norm(self, p):
        a = C.c_double()
        _ck(self.lib.nsb_vec_norm(p, C.byref(a)))
        return a.value

    def normalize(self, p):
        a = C.c_double()
        _ck(self.lib.nsb_vec_normalize(p, C.byref(a)))
        return a.value

    def basis_gemv(self, k, first, y, out):
        y = _arr(y)
        _ck(self.lib.nsb_basis_gemv(k, first, _p(y), out))

    def basis_gemv_complex(self, k, first, y, slot_re, slot_im):
        """mode = sum_i y_i Q_i for a complex coefficient vector (outpost_ks, core/eigensolvers.f:554-564)."""
        y = np.asarray(y, dtype=complex)
        yr, yi = np.ascontiguousarray(y.real), np.ascontiguousarray(y.imag)
        _ck(self.lib.nsb_basis_gemv_complex(k, first, _p(yr), _p(yi), slot_re, slot_im))

    def biorthogonalize(self, dre, dim, are, aim):
        _ck(self.lib.nsb_biorthogonalize(dre, dim, are, aim))

    def wave_maker(self, dre, dim, are, aim):
        wm = np.empty(self.n)
        _ck(self.lib.nsb_wave_maker(dre, dim, are, aim, _p(wm)))
        return wm.reshape(self.nel, -1)

    def basis_rotate(self, k, first, S):
        S = np.asfortranarray(S, dtype=np.float64)
        _ck(self.lib.nsb_basis_rotate(k, first, S.ctypes.data_as(_dp), S.shape[0]))

    def orthonormalize(self, k, first, slot_f):
        h = np.zeros(k + 1)
        _ck(self.lib.nsb_orthonormalize(k, first, slot_f, _p(h)))
        return h

    def matvec(self, mode, slot_in, slot_out):
        _ck(self.lib.nsb_matvec(mode, slot_in, slot_out))

    def nonlinear_forward_map(self, slot_q, slot_f):
        _ck(self.lib.nsb_nonlinear_forward_map(slot_q, slot_f))

    def newton_krylov(self, q_slot, f_slot, dq_slot, work_slot, first_slot, k_dim, end_time, tol, cfl_target=0.5,
                      maxiter_newton=100, maxiter_gmres=100):
        it, res, calls = C.c_int(), C.c_double(), C.c_longlong()
        hist = np.zeros(maxiter_newton)
        rc = self.lib.nsb_newton_krylov(q_slot, f_slot, dq_slot, work_slot, first_slot, k_dim, end_time, cfl_target, tol,
                                        maxiter_newton, maxiter_gmres, C.byref(it), C.byref(res), _p(hist), C.byref(calls))
        if rc not in (0, 3):
            _ck(rc)
        return rc == 0, it.value, res.value, hist[:min(it.value, maxiter_newton)], calls.value

    # ---- host drivers
    def arnoldi_factorization(self, mode, first_slot, H, mstart, mend, ksize):
        assert H.flags.f_contiguous and H.shape == (ksize + 1, ksize)
        _ck(self.lib.nsb_arnoldi_factorization(mode, first_slot, H.ctypes.data_as(_dp), ksize + 1, mstart, mend, ksize))

    def krylov_schur(self, mode, k_dim, schur_tgt, eigen_tol=1e-6, schur_del=0.1, seed_slot=0, max_restarts=-1):
        vr, vi, res = np.zeros(k_dim), np.zeros(k_dim), np.zeros(k_dim)
        vecs = np.zeros(2 * k_dim * k_dim)
        ncv, scnt = C.c_int(), C.c_int()
        rc = self.lib.nsb_krylov_schur(mode, k_dim, schur_tgt, eigen_tol, schur_del, seed_slot, _p(vr), _p(vi), _p(res),
                                       _p(vecs), C.byref(ncv), C.byref(scnt), max_restarts)
        if rc not in (0, 3):
            _ck(rc)
        V = (vecs[0::2] + 1j * vecs[1::2]).reshape(k_dim, k_dim).T      # column i = eigenvector i
        return vr + 1j * vi, res, V, ncv.value, scnt.value

    def ts_gmres(self, mode, rhs_slot, sol_slot, first_slot, work_slot, maxiter, ksize, tol):
        calls, res = C.c_int(), C.c_double()
        _ck(self.lib.nsb_ts_gmres(mode, rhs_slot, sol_slot, first_slot, work_slot, maxiter, ksize, tol, C.byref(calls), C.byref(res)))
        return calls.value, res.value

    # ---- operator-level entry points (host arrays in/out)
    def op_axhelm(self, u, h1, h2):
        u = _arr(u).ravel(); w = np.empty_like(u)
        _ck(self.lib.nsb_op_axhelm(_p(u), h1, h2, _p(w)))
        return w

    def op_dssum(self, u):
        u = _arr(u).ravel().copy()
        _ck(self.lib.nsb_op_dssum(_p(u)))
        return u

    def op_glsc3(self, a, b, c):
        a, b, c = (_arr(x).ravel() for x in (a, b, c))
        out = C.c_double()
        _ck(self.lib.nsb_op_glsc3(_p(a), _p(b), _p(c), C.byref(out)))
        return out.value

    def _vec3(self, v):
        v = _arr(v).reshape(self.ldim, self.n)
        return [np.ascontiguousarray(v[d]) for d in range(self.ldim)] + [None] * (3 - self.ldim)

    def _out3(self):
        v = np.empty((self.ldim, self.n))
        return v, [v[d] for d in range(self.ldim)] + [None] * (3 - self.ldim)

    def op_opgradt(self, p):
        p = _arr(p).ravel(); w, ws = self._out3()
        _ck(self.lib.nsb_op_opgradt(_p(p), _p(ws[0]), _p(ws[1]), _p(ws[2])))
        return w

    def op_opdiv(self, u):
        us = self._vec3(u); q = np.empty(self.n2)
        _ck(self.lib.nsb_op_opdiv(_p(us[0]), _p(us[1]), _p(us[2]), _p(q)))
        return q

    def op_cdabdtp(self, p):
        p = _arr(p).ravel(); ep = np.empty(self.n2)
        _ck(self.lib.nsb_op_cdabdtp(_p(p), _p(ep)))
        return ep

    def op_advab(self, adjoint, up):
        us = self._vec3(up); f, fs = self._out3()
        _ck(self.lib.nsb_op_advab(int(adjoint), _p(us[0]), _p(us[1]), _p(us[2]), _p(fs[0]), _p(fs[1]), _p(fs[2])))
        return f

    def op_hmholtz(self, rhs, h1, h2):
        rs = self._vec3(rhs); x, xs = self._out3(); it = C.c_int()
        _ck(self.lib.nsb_op_hmholtz(_p(xs[0]), _p(xs[1]), _p(xs[2]), _p(rs[0]), _p(rs[1]), _p(rs[2]), h1, h2, C.byref(it)))
        return x, it.value

    def op_esolver(self, g):
        g = _arr(g).ravel(); phi = np.empty(self.n2); it = C.c_int()
        _ck(self.lib.nsb_op_esolver(_p(g), _p(phi), C.byref(it)))
        return phi, it.value

    def op_cfl(self, u, dt):
        us = self._vec3(u); out = C.c_double()
        _ck(self.lib.nsb_op_cfl(_p(us[0]), _p(us[1]), _p(us[2]), dt, C.byref(out)))
        return out.value

    def close(self):
        self.lib.nsb_finalize()


def nccl_unique_id() -> bytes:
    lib = load_library()
    buf = C.create_string_buffer(128)
    _ck(lib.nsb_comm_unique_id(buf))
    return buf.raw
