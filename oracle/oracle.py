"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes view of oracle/_build/liboracle.so, the CPU restatement of the reference hot path
(src/equations/mod_uvwp.f90, src/modules/mod_solver.f90, src/modules/mod_subdomains.f90) and
of the mesh set-up that feeds it.  Only tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs may import this module.  Parity is pinned to the reference's SOURCE TEXT
executed by oracle/f90run/f90py.py (tests/test_oracle_vs_reference_source.py, bit for bit) — the reference has no golden
vectors for this path and cannot be compiled here (no Fortran compiler in the image or on the GPU box).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(force=False):
    if force or not os.path.exists(_LIB):
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.orc_last_error.restype = C.c_char_p
        _lib.orc_create.restype = C.c_void_p
        _lib.orc_real.restype = _dp
        _lib.orc_int.restype = _ip
        _lib.orc_real.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_long)]
        _lib.orc_int.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_long)]
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def _i(a):
    return a.ctypes.data_as(_ip)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class OracleCase:
    """One reference run: mesh set-up (cell_input) + construct_physics, then the hot path."""

    REAL = ("xc yc zc aip rip vol rho mu ap anb b phic u v w p gu gv gw gp gpc mip mip0 "
            "u0 v0 w0 bu bv bw d dc tc cp t gt h h0 gh s s0 gs").split()
    INT = "ef2nb_idx ef2nb_nb ef2nb_fg s2g bs gf2g g2gf_p g2gf_idx".split()

    def __init__(self, raw, n_subdomains=1, geom=None):
        """raw: CGNS-like mesh content.  geom (optional): prebuilt geometry_t arrays — skips the
        oracle's own (slow, faithful) face matching; used for timing runs on large meshes."""
        L = lib()
        self.raw = raw
        et, es = _i32(raw["etype"]), _i32(raw["esec"])
        names = raw["names"]
        if geom is not None:
            L.orc_create_from_geom.restype = C.c_void_p
            a = {k: _i32(geom[k]) for k in ("ef2nb_idx", "ef2nb_nb", "ef2nb_fg", "s2g", "bs")}
            r = {k: _f64(geom[k]) for k in ("xc", "yc", "zc", "aip", "rip", "vol")}
            self.h = L.orc_create_from_geom(C.c_int(int(geom["ne"])), C.c_int(int(geom["nf"])), C.c_int(int(geom["nbf"])),
                                            _i(a["ef2nb_idx"]), _i(a["ef2nb_nb"]), _i(a["ef2nb_fg"]), _i(a["s2g"]), _i(a["bs"]),
                                            _d(r["xc"]), _d(r["yc"]), _d(r["zc"]), _d(r["aip"]), _d(r["rip"]), _d(r["vol"]),
                                            C.c_int(len(et)), _i(et), _i(es), C.c_char_p(names), C.c_int(n_subdomains))
        else:
            x, y, z = _f64(raw["x"]), _f64(raw["y"]), _f64(raw["z"])
            e2vx = _i32(raw["e2vx"])
            nelem = int(raw["ne"] + raw["nbf"])
            self.h = L.orc_create(C.c_int(len(x)), _d(x), _d(y), _d(z), C.c_int(len(et)), _i(et), _i(es),
                                  C.c_char_p(names), C.c_int(int(raw["ne2vx_max"])), C.c_int(nelem),
                                  _i(e2vx), C.c_int(n_subdomains))
        if not self.h:
            raise RuntimeError("oracle: " + L.orc_last_error().decode())
        self.h = C.c_void_p(self.h)
        ne, nf, nbf, nbc = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        L.orc_dims(self.h, C.byref(ne), C.byref(nf), C.byref(nbf), C.byref(nbc))
        self.ne, self.nf, self.nbf, self.nbc = ne.value, nf.value, nbf.value, nbc.value
        self.n_subdomains = n_subdomains

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_destroy(self.h)
            self.h = None

    def view(self, name):
        """numpy view (no copy) of an oracle array; writing to it sets the oracle's state."""
        n = C.c_long()
        if name in self.REAL:
            p = lib().orc_real(self.h, name.encode(), C.byref(n))
            dt = np.float64
        elif name in self.INT:
            p = lib().orc_int(self.h, name.encode(), C.byref(n))
            dt = np.int32
        else:
            raise KeyError(name)
        if n.value < 0:
            raise KeyError(name)
        if n.value == 0:
            return np.zeros(0, dtype=dt)
        return np.ctypeslib.as_array(p, shape=(n.value,))

    def __getitem__(self, name):
        return self.view(name)

    def bc_table(self):
        esec = np.zeros(2 * self.nbc, np.int32)
        kind = np.zeros(self.nbc, np.int32)
        uvw = np.zeros(3 * self.nbc, np.float64)
        lib().orc_bc_table(self.h, _i(esec), _i(kind), _d(uvw))
        return esec, kind, uvw

    def set_bc(self, i, kind, uvw=(0.0, 0.0, 0.0)):
        rc = lib().orc_set_bc(self.h, C.c_int(i), C.c_int(kind), C.c_double(uvw[0]), C.c_double(uvw[1]), C.c_double(uvw[2]))
        assert rc == 0

    def set_param(self, key, v):
        assert lib().orc_set_param(self.h, key.encode(), C.c_double(v)) == 0

    def _chk(self, rc):
        if rc != 0:
            raise RuntimeError("oracle: " + lib().orc_last_error().decode())

    def update_boundaries(self):
        self._chk(lib().orc_update_boundaries(self.h))

    def update_time(self):
        self._chk(lib().orc_update_time(self.h))

    def solve_uvwp(self):
        hist = np.zeros(16)
        self._chk(lib().orc_solve_uvwp(self.h, _d(hist)))
        return hist.reshape(4, 4)

    def run(self, ntstep=10, ncoef=3):
        hist = np.zeros(ntstep * ncoef * 16)
        sec = C.c_double()
        self._chk(lib().orc_run(self.h, C.c_int(ntstep), C.c_int(ncoef), _d(hist), C.byref(sec)))
        return hist.reshape(ntstep * ncoef, 4, 4), sec.value

    def construct_energy(self):
        """construct_energy (mod_energy.f90:14-48): t = 273, phi = cp t, tc = 5, cp = 1000"""
        self._chk(lib().orc_construct_energy(self.h))

    def solve_energy(self):
        out = np.zeros(4)
        self._chk(lib().orc_solve_energy(self.h, _d(out)))
        return out

    def construct_scalar(self, dcoef=1.0, vel=(0.0, 0.0, -100.0), bc_value=None):
        """construct_scalar (mod_scalar.f90:16-46) with a Dirichlet value per boundary section"""
        vel = _f64(vel)
        bcv = _f64(bc_value) if bc_value is not None else None
        self._chk(lib().orc_construct_scalar(self.h, C.c_double(dcoef), _d(vel), _d(bcv) if bcv is not None else None))

    def solve_scalar(self):
        out = np.zeros(4)
        self._chk(lib().orc_solve_scalar(self.h, _d(out)))
        return out

    def calc_coef_uvw(self):
        self._chk(lib().orc_calc_coef_uvw(self.h))

    def calc_coef_p(self):
        self._chk(lib().orc_calc_coef_p(self.h))

    def calc_mip(self, rhie_chow=True):
        self._chk(lib().orc_calc_mip(self.h, C.c_int(1 if rhie_chow else 0)))

    def adjust_pc(self):
        self._chk(lib().orc_adjust_pc(self.h))

    def update_uvwp(self):
        self._chk(lib().orc_update_uvwp(self.h))

    def calc_grad(self, phi):
        phi = _f64(phi)
        grad = np.zeros(3 * (self.ne + self.nbf))
        self._chk(lib().orc_calc_grad(self.h, _d(phi), _d(grad)))
        return grad

    def solve(self, is_pc, ap, anb, b, phi, nit=100):
        """solve(cname, subdomain, intf, ...) — multi_subdomain_solver when n_subdomains>1."""
        ap, anb, b = _f64(ap), _f64(anb), _f64(b)
        phi = _f64(phi).copy()
        out = np.zeros(4)
        self._chk(lib().orc_solve(self.h, C.c_int(int(is_pc)), _d(ap), _d(anb), _d(b), _d(phi), C.c_int(nit), _d(out)))
        return phi, out


def flat_solve_gs(is_pc, phi, ap, anb, b, ef2nb_idx, ef2nb_nb, nit=100):
    phi = _f64(phi).copy()
    ap, anb, b = _f64(ap), _f64(anb), _f64(b)
    idx, nb = _i32(ef2nb_idx), _i32(ef2nb_nb)
    out = np.zeros(4)
    rc = lib().orc_flat_solve_gs(C.c_int(int(is_pc)), _d(phi), _d(ap), _d(anb), _d(b), _i(idx), _i(nb),
                                 C.c_int(len(ap)), C.c_int(nit), _d(out))
    assert rc == 0
    return phi, out


def flat_smoother_gs(is_pc, phi, ap, anb, b, ef2nb_idx, ef2nb_nb, nit=1):
    phi = _f64(phi).copy()
    ap, anb, b = _f64(ap), _f64(anb), _f64(b)
    idx, nb = _i32(ef2nb_idx), _i32(ef2nb_nb)
    rc = lib().orc_flat_smoother_gs(C.c_int(int(is_pc)), _d(phi), _d(ap), _d(anb), _d(b), _i(idx), _i(nb),
                                    C.c_int(len(ap)), C.c_int(nit))
    assert rc == 0
    return phi


def flat_calc_residual(phi, ap, anb, b, ef2nb_idx, ef2nb_nb):
    phi, ap, anb, b = _f64(phi), _f64(ap), _f64(anb), _f64(b)
    idx, nb = _i32(ef2nb_idx), _i32(ef2nb_nb)
    res, res_max = C.c_double(), C.c_double()
    rc = lib().orc_flat_calc_residual(_d(phi), _d(ap), _d(anb), _d(b), _i(idx), _i(nb), C.c_int(len(ap)),
                                      C.byref(res), C.byref(res_max))
    assert rc == 0
    return res.value, res_max.value


def flat_calc_grad(phi, xc, yc, zc, ef2nb_idx, ef2nb_nb):
    phi, xc, yc, zc = _f64(phi), _f64(xc), _f64(yc), _f64(zc)
    idx, nb = _i32(ef2nb_idx), _i32(ef2nb_nb)
    grad = np.zeros(3 * len(phi))
    rc = lib().orc_flat_calc_grad(_d(phi), _d(grad), _d(xc), _d(yc), _d(zc), _i(idx), _i(nb), C.c_int(len(idx) - 1))
    assert rc == 0
    return grad
