"""ctypes binding of libcfdl.so (include/cfdl.h) — the host-side mirror of the reference's
Fortran call sites for the SIMPLE hot path (src/equations/mod_uvwp.f90:95-134,
src/main.f90:50-63).  Thin by design: every method is one C-ABI call with host (numpy)
buffers.  There is no CPU fallback: compute entry points raise CfdlError when the CUDA
library or a GPU is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libcfdl.so")
_lib = None

_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)

MESH_HEX, MESH_TET = 0, 1
BC_WALL, BC_LID, BC_SYMMETRY = 0, 1, 2
SOLVER_PARITY, SOLVER_MCSGS, SOLVER_PCG = 0, 1, 2
EQ_U, EQ_V, EQ_W, EQ_PC = 0, 1, 2, 3
FIELDS = ("u v w p u0 v0 w0 pc gu gv gw gp gpc mip mip0 bu bv bw d dc ap b anb t h h0 s s0 gt gh gs").split()
FIELD_ID = {n: i for i, n in enumerate(FIELDS)}


class CfdlError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise CfdlError("libcfdl.so is not built (run `python -c 'import __graft_entry__ as g; g.build()'`): " + LIB_PATH)
        _lib = C.CDLL(LIB_PATH)
        _lib.cfdl_last_error.restype = C.c_char_p
    return _lib


def _chk(rc):
    if rc != 0:
        raise CfdlError("cfdl error %d: %s" % (rc, lib().cfdl_last_error().decode()))


def _d(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _i(a):
    return a.ctypes.data_as(_ip) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def device_count():
    return int(lib().cfdl_device_count())


# ---- host mesh tooling -----------------------------------------------------------------------
def meshgen(kind, n, jitter=0.0, shuffle=False, seed=12345):
    """Synthetic unit-cube mesh in CGNS conventions (what cell_input.f90:36-99 reads)."""
    L = lib()
    nvx, ne, nbf = C.c_int64(), C.c_int64(), C.c_int64()
    nsec, w = C.c_int(), C.c_int()
    _chk(L.cfdl_meshgen_sizes(C.c_int(kind), C.c_int(n), C.byref(nvx), C.byref(ne), C.byref(nbf), C.byref(nsec), C.byref(w)))
    x, y, z = (np.zeros(nvx.value) for _ in range(3))
    e2vx = np.zeros(w.value * (ne.value + nbf.value), np.int32)
    etype = np.zeros(nsec.value, np.int32)
    esec = np.zeros(2 * nsec.value, np.int32)
    names = C.create_string_buffer(32 * nsec.value)
    _chk(L.cfdl_meshgen_fill(C.c_int(kind), C.c_int(n), C.c_double(jitter), C.c_int(int(shuffle)), C.c_uint64(seed),
                             _d(x), _d(y), _d(z), _i(e2vx), _i(etype), _i(esec), names))
    return dict(kind=kind, n=n, nvx=nvx.value, ne=ne.value, nbf=nbf.value, nsec=nsec.value, ne2vx_max=w.value,
                x=x, y=y, z=z, e2vx=e2vx, etype=etype, esec=esec, names=names.raw)


def rawmesh_write(raw, path):
    """Write a mesh (dict as returned by meshgen) to the raw mesh file format (cfdl_rawmesh_write)."""
    x, y, z = _f64(raw["x"]), _f64(raw["y"]), _f64(raw["z"])
    et, es, e2vx = _i32(raw["etype"]), _i32(raw["esec"]), _i32(raw["e2vx"])
    nelem = int(raw["ne"]) + int(raw["nbf"])
    _chk(lib().cfdl_rawmesh_write(path.encode(), C.c_int64(len(x)), _d(x), _d(y), _d(z), C.c_int(len(et)), _i(et), _i(es),
                                  C.c_char_p(bytes(raw["names"])), C.c_int(int(raw["ne2vx_max"])), _i(e2vx), C.c_int64(nelem)))


def rawmesh_read(path):
    """Read a raw mesh file into the dict format of meshgen (kind / n are not stored: None)."""
    L = lib()
    nvx, nelem = C.c_int64(), C.c_int64()
    nsec, w = C.c_int(), C.c_int()
    _chk(L.cfdl_rawmesh_sizes(path.encode(), C.byref(nvx), C.byref(nelem), C.byref(nsec), C.byref(w)))
    x, y, z = (np.zeros(nvx.value) for _ in range(3))
    e2vx = np.zeros(w.value * nelem.value, np.int32)
    etype = np.zeros(nsec.value, np.int32)
    esec = np.zeros(2 * nsec.value, np.int32)
    names = C.create_string_buffer(32 * nsec.value)
    _chk(L.cfdl_rawmesh_read(path.encode(), _d(x), _d(y), _d(z), _i(etype), _i(esec), names, _i(e2vx)))
    ne = sum(int(esec[2 * s + 1] - esec[2 * s] + 1) for s in range(nsec.value) if etype[s] >= 10)
    return dict(kind=None, n=None, nvx=nvx.value, ne=ne, nbf=nelem.value - ne, nsec=nsec.value, ne2vx_max=w.value,
                x=x, y=y, z=z, e2vx=e2vx, etype=etype, esec=esec, names=names.raw)


def mesh_build(raw, gpu=False, device=0):
    """Connectivity + geometry (find_element_nb, calc_aip_xyzip_uns, calc_vol_cv_centers_uns); gpu=True: cfdl_mesh_build_gpu."""
    L = lib()
    ne, nbf = int(raw["ne"]), int(raw["nbf"])
    nface_per = {17: 6, 10: 4, 12: 5, 14: 5}
    tot = 0
    for s in range(raw["nsec"]):
        t = int(raw["etype"][s])
        cnt = int(raw["esec"][2 * s + 1] - raw["esec"][2 * s] + 1)
        if t >= 10:
            tot += nface_per[t] * cnt
    nf = (tot + nbf) // 2
    Z = 2 * nf - nbf
    H = ne + nbf
    g = dict(ne=ne, nf=nf, nbf=nbf,
             ef2nb_idx=np.zeros(ne + 1, np.int32), ef2nb_nb=np.zeros(Z, np.int32), ef2nb_fg=np.zeros(Z, np.int32),
             s2g=np.zeros(nf, np.int32), bs=np.zeros(nbf, np.int32),
             xc=np.zeros(H), yc=np.zeros(H), zc=np.zeros(H), aip=np.zeros(3 * nf), rip=np.zeros(3 * nf), vol=np.zeros(ne))
    x, y, z = _f64(raw["x"]), _f64(raw["y"]), _f64(raw["z"])
    et, es, e2vx = _i32(raw["etype"]), _i32(raw["esec"]), _i32(raw["e2vx"])
    args = (C.c_int64(len(x)), _d(x), _d(y), _d(z), C.c_int(len(et)), _i(et), _i(es),
            C.c_int(int(raw["ne2vx_max"])), _i(e2vx), C.c_int32(ne), C.c_int32(nf), C.c_int32(nbf),
            _i(g["ef2nb_idx"]), _i(g["ef2nb_nb"]), _i(g["ef2nb_fg"]), _i(g["s2g"]), _i(g["bs"]),
            _d(g["xc"]), _d(g["yc"]), _d(g["zc"]), _d(g["aip"]), _d(g["rip"]), _d(g["vol"]))
    if gpu:
        _chk(L.cfdl_mesh_build_gpu(C.c_int32(device), *args))
    else:
        _chk(L.cfdl_mesh_build(*args))
    return g


def partition_rcb(geom, P, want_order=True):
    """The reference's RCB blocks: (cell2sub, g2gf_p, g2gf_idx), all 1-based.  want_order=False
    skips the block-local order (the reference's unstable quicksort, quadratic in cells per
    block), which only the exact pc block solver needs."""
    ne = int(geom["ne"])
    c2s = np.zeros(ne, np.int32)
    p, idx = (np.zeros(ne, np.int32), np.zeros(P + 1, np.int32)) if want_order else (None, None)
    xc, yc, zc, vol = (_f64(geom[k]) for k in ("xc", "yc", "zc", "vol"))
    _chk(lib().cfdl_partition_rcb(C.c_int32(ne), _d(xc), _d(yc), _d(zc), _d(vol), C.c_int32(P), _i(c2s), _i(p), _i(idx)))
    return c2s, p, idx


def structured_hex_arrays(n, nranks=1):
    """The connectivity / geometry Solver.structured_hex(n) is built from (CPU only)."""
    ne, nbf, nf = n ** 3, 6 * n * n, 3 * n * n * (n + 1)
    o = dict(nb=np.empty(6 * ne, np.int32), fg=np.empty(6 * ne, np.int32), xc=np.empty(ne + nbf), yc=np.empty(ne + nbf),
             zc=np.empty(ne + nbf), vol=np.empty(ne), aip=np.empty(3 * nf), rip=np.empty(3 * nf), cell2rank=np.empty(ne, np.int32))
    _chk(lib().cfdl_structured_hex_arrays(C.c_int32(n), C.c_int32(nranks), _i(o["nb"]), _i(o["fg"]), _d(o["xc"]), _d(o["yc"]),
                                          _d(o["zc"]), _d(o["vol"]), _d(o["aip"]), _d(o["rip"]), _i(o["cell2rank"])))
    return o


def comm_unique_id():
    """128-byte NCCL unique id (create on rank 0, broadcast to the other ranks)."""
    buf = (C.c_uint8 * 128)()
    _chk(lib().cfdl_comm_unique_id(buf))
    return bytes(buf)


def partition_plan(geom, cell2rank, nranks, rank):
    """Host-only view of rank's partition: owned/ghost cells (1-based global ids, device order),
    neighbour ranks and the send/receive lists."""
    ne = int(geom["ne"])
    a = {k: _i32(geom[k]) for k in ("ef2nb_idx", "ef2nb_nb", "ef2nb_fg", "s2g", "bs")}
    r = {k: _f64(geom[k]) for k in ("xc", "yc", "zc")}
    c2r = _i32(cell2rank)
    owned, ghost, send = (np.zeros(ne, np.int32) for _ in range(3))
    nbr = np.zeros(nranks, np.int32)
    sp, rp = np.zeros(nranks + 1, np.int32), np.zeros(nranks + 1, np.int32)
    cp = np.zeros(33, np.int32)
    no, ng, nn, nc = C.c_int32(), C.c_int32(), C.c_int32(), C.c_int32()
    _chk(lib().cfdl_partition_plan(C.c_int32(ne), C.c_int32(int(geom["nf"])), C.c_int32(int(geom["nbf"])),
                                   _i(a["ef2nb_idx"]), _i(a["ef2nb_nb"]), _i(a["ef2nb_fg"]), _i(a["s2g"]), _i(a["bs"]),
                                   _d(r["xc"]), _d(r["yc"]), _d(r["zc"]), _i(c2r), C.c_int32(nranks), C.c_int32(rank),
                                   C.byref(no), _i(owned), C.byref(ng), _i(ghost), C.byref(nn), _i(nbr), _i(sp), _i(send), _i(rp),
                                   C.byref(nc), _i(cp)))
    k = nn.value
    return dict(owned=owned[:no.value].copy(), ghost=ghost[:ng.value].copy(), nbr_rank=nbr[:k].copy(), send_ptr=sp[:k + 1].copy(),
                send_cells=send[:sp[k]].copy(), recv_ptr=rp[:k + 1].copy(), ncolors=nc.value, color_ptr=cp[:nc.value + 1].copy())


def default_bcs(raw):
    """The reference's hard-wired BCs (mod_uvwp.f90:73-78): 'top' = lid u=1, others no-slip;
    one BC per 2-D section in section order (mod_eqn_setup.f90:46-68)."""
    esec, kind, uvw = [], [], []
    for s in range(raw["nsec"]):
        if int(raw["etype"][s]) >= 10:
            continue
        name = raw["names"][32 * s:32 * s + 32].decode().strip()
        esec += [int(raw["esec"][2 * s]), int(raw["esec"][2 * s + 1])]
        if name and name in "top":
            kind.append(BC_LID)
            uvw += [1.0, 0.0, 0.0]
        else:
            kind.append(BC_WALL)
            uvw += [0.0, 0.0, 0.0]
    return np.array(esec, np.int32), np.array(kind, np.int32), np.array(uvw, np.float64)


def host_register(arr):
    """Page-lock a numpy array the caller owns (cfdl_host_register); call host_unregister before it is freed."""
    _chk(lib().cfdl_host_register(C.c_void_p(arr.ctypes.data), C.c_uint64(arr.nbytes)))


def host_unregister(arr):
    _chk(lib().cfdl_host_unregister(C.c_void_p(arr.ctypes.data)))


class PinnedBuffer:
    """Page-locked host array (cudaMallocHost) for the end-to-end path."""

    def __init__(self, n):
        self.ptr = C.c_void_p()
        _chk(lib().cfdl_host_alloc(C.byref(self.ptr), C.c_uint64(8 * max(int(n), 1))))
        self.array = np.ctypeslib.as_array(C.cast(self.ptr, _dp), shape=(max(int(n), 1),))[:int(n)]

    def free(self):
        if self.ptr:
            self.array = None
            lib().cfdl_host_free(self.ptr)
            self.ptr = None


class Solver:
    """Device-resident SIMPLE hot path for one mesh (one GPU).  Mirrors phys_t + uvwp_t."""

    def __init__(self, geom, bcs, rho=5.0, mu=0.01, n_subdomains=1, g2gf_p=None, g2gf_idx=None, device=0,
                 cell2rank=None, rank=0, nranks=1):
        """geom: the GLOBAL mesh in the reference's array format.  With nranks > 1 this handle is
        rank `rank`'s partition (cell2rank in 1..nranks); call comm_init() before computing."""
        L = lib()
        self.rank, self.nranks = rank, nranks
        self.ne, self.nf, self.nbf = int(geom["ne"]), int(geom["nf"]), int(geom["nbf"])
        self.H = self.ne + self.nbf
        self.Z = 2 * self.nf - self.nbf
        a = {k: _i32(geom[k]) for k in ("ef2nb_idx", "ef2nb_nb", "ef2nb_fg", "s2g", "bs")}
        r = {k: _f64(geom[k]) for k in ("xc", "yc", "zc", "aip", "rip", "vol")}
        rho = _f64(np.broadcast_to(rho, (self.ne,)))
        mu = _f64(np.broadcast_to(mu, (self.ne,)))
        esec, kind, uvw = (_i32(bcs[0]), _i32(bcs[1]), _f64(bcs[2]))
        p = _i32(g2gf_p) if n_subdomains > 1 else None
        pi = _i32(g2gf_idx) if n_subdomains > 1 else None
        h = C.c_void_p()
        if nranks > 1:
            c2r = _i32(cell2rank)
            _chk(L.cfdl_create_distributed(C.byref(h), C.c_int32(self.ne), C.c_int32(self.nf), C.c_int32(self.nbf),
                                           _i(a["ef2nb_idx"]), _i(a["ef2nb_nb"]), _i(a["ef2nb_fg"]), _i(a["s2g"]), _i(a["bs"]),
                                           _d(r["xc"]), _d(r["yc"]), _d(r["zc"]), _d(r["aip"]), _d(r["rip"]), _d(r["vol"]),
                                           _d(rho), _d(mu), C.c_int32(len(kind)), _i(esec), _i(kind), _d(uvw),
                                           _i(c2r), C.c_int32(rank), C.c_int32(nranks), C.c_int32(device)))
        else:
            _chk(L.cfdl_create(C.byref(h), C.c_int32(self.ne), C.c_int32(self.nf), C.c_int32(self.nbf),
                               _i(a["ef2nb_idx"]), _i(a["ef2nb_nb"]), _i(a["ef2nb_fg"]), _i(a["s2g"]), _i(a["bs"]),
                               _d(r["xc"]), _d(r["yc"]), _d(r["zc"]), _d(r["aip"]), _d(r["rip"]), _d(r["vol"]),
                               _d(rho), _d(mu), C.c_int32(len(kind)), _i(esec), _i(kind), _d(uvw),
                               C.c_int32(n_subdomains), _i(p), _i(pi), C.c_int32(device)))
        self.h = h
        self.n_subdomains = n_subdomains

    @classmethod
    def structured_hex(cls, n, rho=5.0, mu=0.01, device=0, rank=0, nranks=1, slabs=False, nz=None):
        """The n^3 lid-driven cavity generated per rank without the packed int32 arrays
        (cfdl_create_structured_hex): the way to the 512^3 target, identical to
        Solver(mesh_build(meshgen("hex", n)), default_bcs(...)) in everything but the host-side
        numbering of face fields."""
        self = cls.__new__(cls)
        self.rank, self.nranks = rank, nranks
        nz = nz or n  # nz != n (slabs only): n x n x nz cells of edge 1/n
        self.ne, self.nbf = n * n * nz, 2 * n * n + 4 * n * nz
        self.nf = 2 * (n + 1) * n * nz + n * n * (nz + 1)
        self.H = self.ne + self.nbf
        self.Z = 6 * self.ne
        self.n_subdomains = 1
        h = C.c_void_p()
        if slabs:  # nranks z-slabs instead of bisection
            _chk(lib().cfdl_create_structured_hex_slabs(C.byref(h), C.c_int32(n), C.c_int32(nz), C.c_double(rho), C.c_double(mu), C.c_int32(rank),
                                                        C.c_int32(nranks), C.c_int32(device)))
        else:
            _chk(lib().cfdl_create_structured_hex(C.byref(h), C.c_int32(n), C.c_double(rho), C.c_double(mu), C.c_int32(rank), C.c_int32(nranks), C.c_int32(device)))
        self.h = h
        return self

    def close(self):
        if getattr(self, "h", None):
            lib().cfdl_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def field_size(self, name):
        if name in ("u", "v", "w", "p", "u0", "v0", "w0", "pc", "t", "h", "h0", "s", "s0"):
            return self.H
        if name in ("gu", "gv", "gw", "gp", "gpc", "gt", "gh", "gs"):
            return 3 * self.H
        if name in ("mip", "mip0"):
            return self.nf
        if name == "anb":
            return self.Z
        return self.ne

    def set_option(self, key, value):
        _chk(lib().cfdl_set_option(self.h, key.encode(), C.c_double(value)))

    def get_info(self, key):
        v = C.c_double()
        _chk(lib().cfdl_get_info(self.h, key.encode(), C.byref(v)))
        return v.value

    def cell_order(self):
        nc = int(self.get_info("ncolors"))
        c2o = np.zeros(self.ne, np.int32)
        cp = np.zeros(nc + 1, np.int32)
        _chk(lib().cfdl_get_cell_order(self.h, _i(c2o), _i(cp)))
        return c2o, cp

    def comm_init(self, unique_id):
        """Collective: join the NCCL communicator identified by the 128-byte id of rank 0."""
        buf = (C.c_uint8 * 128).from_buffer_copy(bytes(unique_id))
        _chk(lib().cfdl_comm_init(self.h, buf, C.c_int32(self.rank), C.c_int32(self.nranks)))

    def ipc_handle(self):
        """64-byte CUDA IPC handle of this rank's exchange slab (gather over ranks, then ipc_connect)."""
        buf = (C.c_uint8 * 64)()
        _chk(lib().cfdl_comm_ipc_handle(self.h, buf))
        return bytes(buf)

    def ipc_connect(self, handles):
        """handles: list of the nranks 64-byte handles in rank order.  Enables peer-to-peer ghost exchange."""
        blob = b"".join(bytes(x) for x in handles)
        assert len(blob) == 64 * self.nranks
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        _chk(lib().cfdl_comm_ipc_connect(self.h, buf))

    def local_size(self, name):
        n = C.c_int64()
        _chk(lib().cfdl_field_local_size(self.h, C.c_int(FIELD_ID[name]), C.byref(n)))
        return n.value

    def upload_local(self, name, arr):
        assert arr.size == self.local_size(name) and arr.dtype == np.float64
        _chk(lib().cfdl_upload_field_local(self.h, C.c_int(FIELD_ID[name]), _d(arr)))

    def download_local(self, name, out):
        assert out.size == self.local_size(name) and out.dtype == np.float64
        _chk(lib().cfdl_download_field_local(self.h, C.c_int(FIELD_ID[name]), _d(out)))
        return out

    def timer_record(self, slot):
        _chk(lib().cfdl_timer_record(self.h, C.c_int32(slot)))

    def timer_elapsed_ms(self, a, b):
        ms = C.c_double()
        _chk(lib().cfdl_timer_elapsed_ms(self.h, C.c_int32(a), C.c_int32(b), C.byref(ms)))
        return ms.value

    def upload(self, name, arr):
        arr = _f64(arr)
        assert arr.size == self.field_size(name), (name, arr.size, self.field_size(name))
        _chk(lib().cfdl_upload_field(self.h, C.c_int(FIELD_ID[name]), _d(arr)))

    def download(self, name):
        out = np.zeros(self.field_size(name))
        _chk(lib().cfdl_download_field(self.h, C.c_int(FIELD_ID[name]), _d(out)))
        return out

    def download_into(self, name, out):
        assert out.size == self.field_size(name) and out.dtype == np.float64 and out.flags.c_contiguous
        _chk(lib().cfdl_download_field(self.h, C.c_int(FIELD_ID[name]), _d(out)))
        return out

    def update_boundaries(self):
        _chk(lib().cfdl_update_boundaries(self.h))

    def update_time(self):
        _chk(lib().cfdl_update_time(self.h))

    def write_vtu(self, path, raw, equation=0):
        """write_vtubin (mod_vtu_output.f90:6-326) from the device fields; raw = the mesh file content (meshgen / rawmesh)"""
        x, y, z = _f64(raw["x"]), _f64(raw["y"]), _f64(raw["z"])
        esec, etype, e2vx = _i32(raw["esec"]), _i32(raw["etype"]), _i32(raw["e2vx"])
        _chk(lib().cfdl_write_vtu(self.h, path.encode(), C.c_int32(equation), C.c_int32(len(x)), _d(x), _d(y), _d(z), C.c_int32(len(etype)),
                                  _i(esec), _i(etype), C.c_int32(int(raw["ne2vx_max"])), _i(e2vx)))

    def energy_init(self, tc=None, cp=None):
        """construct_energy (mod_energy.f90:14-48); tc / cp per cell (reference numbering) or None for 5 / 1000"""
        tc = _f64(tc) if tc is not None else None
        cp = _f64(cp) if cp is not None else None
        _chk(lib().cfdl_energy_init(self.h, _d(tc) if tc is not None else None, _d(cp) if cp is not None else None))

    def solve_energy(self, dt=0.01, nit=100):
        out = np.zeros(4)
        _chk(lib().cfdl_solve_energy(self.h, C.c_double(dt), C.c_int32(nit), _d(out)))
        return out

    def scalar_init(self, dcoef=1.0, vel=(0.0, 0.0, -100.0), bc_value=None):
        """construct_scalar (mod_scalar.f90:16-46) with a Dirichlet value per boundary section of this mesh"""
        vel = _f64(vel)
        bcv = _f64(bc_value) if bc_value is not None else None
        _chk(lib().cfdl_scalar_init(self.h, C.c_double(dcoef), _d(vel), _d(bcv) if bcv is not None else None))

    def solve_scalar(self, dt=0.01, nit=100):
        out = np.zeros(4)
        _chk(lib().cfdl_solve_scalar(self.h, C.c_double(dt), C.c_int32(nit), _d(out)))
        return out

    def solve_uvwp(self, dt=0.01, nit=100):
        hist = np.zeros(16)
        _chk(lib().cfdl_solve_uvwp(self.h, C.c_double(dt), C.c_int32(nit), _d(hist)))
        return hist.reshape(4, 4)

    def step_host(self, ins, outs, dt=0.01, nit=100, apply_bcs=True, local=False):
        """One SIMPLE iteration with host arrays (cfdl_step_host): ins/outs map field names to float64
        arrays (page-locked ones let the transfers overlap the computation)."""
        def pack(d, const):
            ids = (C.c_int32 * len(d))(*[FIELD_ID[k] for k in d])
            ptrs = (_dp * len(d))(*[_d(a) for a in d.values()])
            return ids, ptrs
        size = self.local_size if local else self.field_size
        for k, a in list(ins.items()) + list(outs.items()):
            assert a.dtype == np.float64 and a.flags.c_contiguous and a.size == size(k), k
        iid, iptr = pack(ins, True)
        oid, optr = pack(outs, False)
        hist = np.zeros(16)
        _chk(lib().cfdl_step_host(self.h, C.c_double(dt), C.c_int32(nit), C.c_int32(1 if apply_bcs else 0), C.c_int32(1 if local else 0),
                                  C.c_int32(len(ins)), iid, iptr, C.c_int32(len(outs)), oid, optr, _d(hist)))
        return hist.reshape(4, 4)

    def checkpoint_write(self, path):
        _chk(lib().cfdl_checkpoint_write(self.h, path.encode()))

    def checkpoint_read(self, path):
        _chk(lib().cfdl_checkpoint_read(self.h, path.encode()))

    def run(self, dt=0.01, nit=100, ntstep=10, ncoef=3, want_hist=True):
        hist = np.zeros(ntstep * ncoef * 16) if want_hist else None
        _chk(lib().cfdl_run(self.h, C.c_double(dt), C.c_int32(nit), C.c_int32(ntstep), C.c_int32(ncoef), _d(hist)))
        return hist.reshape(ntstep * ncoef, 4, 4) if want_hist else None

    def calc_coef_uvw(self, dt=0.01):
        _chk(lib().cfdl_calc_coef_uvw(self.h, C.c_double(dt)))

    def calc_mip(self, rhie_chow=True, dt=0.01):
        _chk(lib().cfdl_calc_mip(self.h, C.c_int32(1 if rhie_chow else 0), C.c_double(dt)))

    def calc_coef_p(self):
        _chk(lib().cfdl_calc_coef_p(self.h))

    def adjust_pc(self):
        _chk(lib().cfdl_adjust_pc(self.h))

    def update_uvwp(self):
        _chk(lib().cfdl_update_uvwp(self.h))

    def calc_grad(self, phi_name, grad_name):
        _chk(lib().cfdl_calc_grad(self.h, C.c_int(FIELD_ID[phi_name]), C.c_int(FIELD_ID[grad_name])))

    def solve_eq(self, eq, nit=100):
        out = np.zeros(4)
        _chk(lib().cfdl_solve_eq(self.h, C.c_int(eq), C.c_int32(nit), _d(out)))
        return out

    # stand-alone drop-ins (host arrays in, host arrays out)
    def host_calc_grad(self, phi):
        phi = _f64(phi)
        grad = np.zeros(3 * self.H)
        _chk(lib().cfdl_host_calc_grad(self.h, _d(phi), _d(grad)))
        return grad

    def host_solve_gs(self, eq, phi, ap, anb, b, nit=100):
        phi = _f64(phi).copy()
        ap, anb, b = _f64(ap), _f64(anb), _f64(b)
        out = np.zeros(4)
        _chk(lib().cfdl_host_solve_gs(self.h, C.c_int(eq), _d(phi), _d(ap), _d(anb), _d(b), C.c_int32(nit), _d(out)))
        return phi, out

    def host_solve(self, eq, phi, ap, anb, b, nit=100):
        phi = _f64(phi).copy()
        ap, anb, b = _f64(ap), _f64(anb), _f64(b)
        out = np.zeros(4)
        _chk(lib().cfdl_host_solve(self.h, C.c_int(eq), _d(phi), _d(ap), _d(anb), _d(b), C.c_int32(nit), _d(out)))
        return phi, out

    def host_calc_coef_uvw(self, f, dt=0.01):
        """f: dict of host arrays u,v,w,u0,v0,w0,gu,gv,gw,gp,mip -> dict ap,anb,bu,bv,bw,d,dc."""
        ins = [_f64(f[k]) for k in ("u", "v", "w", "u0", "v0", "w0", "gu", "gv", "gw", "gp", "mip")]
        out = {k: np.empty(self.field_size(k)) for k in ("ap", "anb", "bu", "bv", "bw", "d", "dc")}
        _chk(lib().cfdl_host_calc_coef_uvw(self.h, C.c_double(dt), *[_d(a) for a in ins], *[_d(out[k]) for k in out]))
        return out

    def host_calc_mip(self, f, rhie_chow=True, dt=0.01):
        """f: dict u,v,w,u0,v0,w0,p,gp,d,mip0,mip -> new mip (boundary entries pass through)."""
        ins = [_f64(f[k]) for k in ("u", "v", "w", "u0", "v0", "w0", "p", "gp", "d", "mip0")]
        mip = _f64(f["mip"]).copy()
        _chk(lib().cfdl_host_calc_mip(self.h, C.c_int32(1 if rhie_chow else 0), C.c_double(dt), *[_d(a) for a in ins], _d(mip)))
        return mip

    def host_calc_coef_p(self, dc, mip):
        out = {k: np.empty(self.field_size(k)) for k in ("ap", "anb", "b")}
        _chk(lib().cfdl_host_calc_coef_p(self.h, _d(_f64(dc)), _d(_f64(mip)), *[_d(out[k]) for k in out]))
        return out

    def host_adjust_pc(self, pc):
        pc = _f64(pc).copy()
        _chk(lib().cfdl_host_adjust_pc(self.h, _d(pc)))
        return pc

    def host_update_uvwp(self, pc, gpc, dc, p, gp, mip):
        p, gp, mip = _f64(p).copy(), _f64(gp).copy(), _f64(mip).copy()
        _chk(lib().cfdl_host_update_uvwp(self.h, _d(_f64(pc)), _d(_f64(gpc)), _d(_f64(dc)), _d(p), _d(gp), _d(mip)))
        return p, gp, mip

    def host_calc_residual(self, phi, ap, anb, b):
        phi, ap, anb, b = _f64(phi), _f64(ap), _f64(anb), _f64(b)
        res, res_max = C.c_double(), C.c_double()
        _chk(lib().cfdl_host_calc_residual(self.h, _d(phi), _d(ap), _d(anb), _d(b), C.byref(res), C.byref(res_max)))
        return res.value, res_max.value
