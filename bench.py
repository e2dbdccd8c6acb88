#!/usr/bin/env python
"""bench.py — SIMPLE hot path of CFD-Lite on B200: cell-iterations/s.

A "step" is one SIMPLE iteration of src/main.f90:50-63 on the synthetic lid-driven-cavity
mesh: update_boundaries + solve_uvwp (assembly, three momentum solves, gradients, Rhie-Chow,
pressure-correction assembly + solve, correction), plus update_time after every third step.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--size 128] [--mesh hex|tet] [--solver mcsgs|parity|pcg]

ours       the CUDA library through its C ABI (include/cfdl.h).  `value` is timed with CUDA
           events on the library's stream with all state resident in HBM; `e2e` repeats the
           steps with the step's input fields uploaded from pinned host memory and the results
           downloaded inside the timed region (what a host driver that owns the arrays pays).
reference  the reference's own CPU algorithm (the C++ oracle, one thread: the reference is
           serial Fortran and cannot be compiled here) on the same workload.
One JSON line on stdout (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "cfd-lite_b200", "python"))
# stdout carries exactly one JSON line, but native libraries print there too (NCCL's version banner and, with
# NCCL_DEBUG=INFO, its whole log).  NCCL_DEBUG is left as the caller set it; instead file descriptor 1 is pointed at
# stderr for the life of the process and the JSON line goes to a private duplicate of the original stdout.
sys.stdout.flush()
_JSON_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)

SOLVERS = {"parity": 0, "mcsgs": 1, "pcg": 2}
DT, NIT, NCOEF = 0.01, 100, 3  # reference defaults, src/modules/mod_physics.f90:15-18


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock and throttle reasons of one GPU sampled during the timed region: NVML inside this process (nvidia_ml_py; initialised
    before the warm-up, then one cheap query every 20 ms from a thread; only the samples of the timed region count), or — without NVML bindings — an `nvidia-smi -lms 50` child.
    Measured (8 GPUs, 30 timed steps, round 2, profiles/r02_exp_8gpu_stack_sampler_*.json): 5.90 ms per step without a sampler, 5.85 with
    the in-process queries, 5.88 with the nvidia-smi child — neither disturbs the run; the in-process form needs no child process
    and no start-up time."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device, how="nvml"):
        self.device, self.proc, self.lines, self.how = device, None, [], how
        self.nvml, self.handle, self.t, self.samples, self.stop_flag = None, None, None, [], threading.Event()
        self.samples_before = self.lines_before = 0

    def _nvml_id(self):
        """the GPU this process computes on, as NVML / nvidia-smi name it: CUDA device i is entry i of CUDA_VISIBLE_DEVICES
        (an index or a UUID) when that is set, else index i"""
        vis = [v.strip() for v in os.environ.get("CUDA_VISIBLE_DEVICES", "").split(",") if v.strip()]
        if self.device < len(vis):
            return vis[self.device]
        return str(self.device)

    def start(self):
        if self.how == "off":
            return
        ident = self._nvml_id()
        if self.how == "nvml":
            try:
                import pynvml
                pynvml.nvmlInit()
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(int(ident)) if ident.isdigit() else pynvml.nvmlDeviceGetHandleByUUID(ident.encode())
                self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM))
                pynvml.nvmlDeviceGetClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
                self.nvml = pynvml
                self.t = threading.Thread(target=self._poll, daemon=True)
                self.t.start()
                return
            except Exception as ex:
                dbg("NVML sampler unavailable (%s): nvidia-smi child instead" % ex)
                self.nvml = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", ident, "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def mark(self):
        """the timed region starts here: earlier samples (warm-up, clocks still ramping up from idle) are dropped"""
        self.samples_before = len(self.samples)
        self.lines_before = len(self.lines)

    def _poll(self):
        n = self.nvml
        bits = (("hw_slowdown", n.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", n.nvmlClocksThrottleReasonHwThermalSlowdown),
                ("sw_thermal_slowdown", n.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", n.nvmlClocksThrottleReasonSwPowerCap))
        while not self.stop_flag.is_set():
            try:
                mhz = float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM))
                mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
                self.samples.append((mhz, [name for name, b in bits if mask & b]))
            except Exception:
                pass
            self.stop_flag.wait(0.02)

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if self.how == "off":
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["not sampled (--clock-sampler off)"]}
        if self.nvml is not None:
            self.stop_flag.set()
            self.t.join(timeout=2)
            timed = self.samples[self.samples_before:] or self.samples
            sm = [m for m, _ in timed]
            reasons = sorted({r for _, rs in timed for r in rs})
            try:
                self.nvml.nvmlShutdown()
            except Exception:
                pass
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm),
                    "source": "NVML in-process, every 20 ms, samples of the timed region"}
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()  # exact PID of the sampler we started
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in (self.lines[self.lines_before:] or self.lines):
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": "nvidia-smi -lms 50 (child process)"}


def ncu_traffic(kernel_key, args, world):
    """DRAM bytes per pass of the dominant kernel from the committed ncu capture (profiles/r02_traffic.json:
    dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` launch, divided by the passes the launch
    executes); only quoted for the kernel and the configuration that capture was taken on, else null."""
    p = os.path.join(ROOT, "profiles", "r02_traffic.json")
    if not (world == 1 and args.mesh == "hex" and args.n == 128 and os.path.exists(p)):
        return None
    try:
        return float(json.load(open(p))["dram_bytes_per_pass"][kernel_key])
    except Exception:
        return None


SETUP_S = {}  # seconds per set-up stage of the product arm (reported in config.setup_breakdown_seconds)


def build_mesh(cfdl, kind, n, device=None):
    """device: connectivity + geometry on that GPU (cfdl_mesh_build_gpu); None: the host builder (the CPU reference arm)"""
    t0 = time.perf_counter()
    raw = cfdl.meshgen(cfdl.MESH_HEX if kind == "hex" else cfdl.MESH_TET, n, jitter=0.0 if kind == "hex" else 0.2,
                       shuffle=(kind == "tet"), seed=12345)
    t1 = time.perf_counter()
    geom = cfdl.mesh_build(raw) if device is None else cfdl.mesh_build(raw, gpu=True, device=device)
    SETUP_S["meshgen (vertices + element lists, host)"] = round(t1 - t0, 2)
    SETUP_S["mesh_build (connectivity + geometry, %s; the first CUDA call of the process creates the context)" % ("host" if device is None else "GPU")] = round(time.perf_counter() - t1, 2)
    return raw, geom


def workload_name(kind, n, ne, nz=None):
    if nz and nz != n:
        return ("lid-driven cavity of depth %d, synthetic %dx%dx%d hex mesh = %d stacked %d^3 blocks of cell edge 1/%d (%d cells), lid on the top z face, "
                "rho=5 mu=0.01 dt=0.01 nit=100" % (nz // n, n, n, nz, nz // n, n, n, ne))
    return "lid-driven cavity, synthetic %s mesh n=%d (%d cells), rho=5 mu=0.01 dt=0.01 nit=100" % (
        "%d^3 hex" % n if kind == "hex" else "Kuhn-tet (jitter 0.2h, shuffled ids)", n, ne)


def algorithmic_bytes(ne, Z, H):
    """SURVEY §8(d) / BASELINE.md §3 algorithmic bytes of one full sweep / one residual pass over
    ne rows with Z coefficient slots gathering from H values (cells + ghosts + halos)."""
    return {"sgs_sweep": 28 * ne + 12 * Z + 8 * H, "residual": 20 * ne + 12 * Z + 8 * H}


def run_reference(args, rank):
    """Reference arm: the reference's CPU algorithm (oracle port, 1 thread) on the same workload."""
    if rank != 0:
        return
    import cfdl
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    raw, geom = build_mesh(cfdl, args.mesh, args.n)
    oc = oracle.OracleCase(raw, n_subdomains=4, geom=geom)  # reference default n_subdomains=4
    oc.set_param("dt", DT); oc.set_param("nit", NIT)

    def step(i):
        oc.update_boundaries()
        oc.solve_uvwp()
        if (i + 1) % NCOEF == 0:
            oc.update_time()

    for i in range(args.warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        step(args.warmup + i)
    sec = time.perf_counter() - t0
    value = oc.ne * args.steps / sec
    line = {"impl": "reference", "metric": "cell-iterations/s (SIMPLE)", "value": value, "unit": "cell-iterations/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sec / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.mesh, args.n, oc.ne), "n_subdomains": 4,
                       "solver": "reference block-SGS (multi_subdomain_solver) + natural-order SGS"},
            "cpu_baseline": {"value": value, "unit": "cell-iterations/s", "cores": 1, "kind": "port",
                             "sample": "%d SIMPLE iterations of the full workload after %d warm-up iterations; C++ oracle "
                                       "(g++ -O3 -ffp-contract=off, 1 thread: the reference is serial and has no OpenMP)" % (args.steps, args.warmup)},
            "e2e": {"value": value, "unit": "cell-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), file=_JSON_OUT, flush=True)


def dbg(*a):
    if os.environ.get("CFDL_BENCH_DEBUG"):
        print("[rank %s]" % os.environ.get("RANK", "0"), *a, file=sys.stderr, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=int(os.environ.get("WORLD_SIZE", "1")))
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", dest="n", type=int, default=128, help="cells per edge per GPU (torchrun would swallow --n)")
    ap.add_argument("--mesh", default="hex", choices=["hex", "tet"])
    ap.add_argument("--solver", default="mcsgs", choices=list(SOLVERS))
    ap.add_argument("--unfused", action="store_true", help="one launch per colour + residual pass (no fused two-colour passes)")
    ap.add_argument("--no-p2p", action="store_true", help="multi-GPU: NCCL send/recv instead of peer-to-peer ghost stores")
    ap.add_argument("--global-size", type=int, default=None, help="cells per edge of the WHOLE cube (overrides the weak-scaling rule), e.g. 512 with --gpus 8")
    ap.add_argument("--structured", action="store_true",
                    help="hex cavity generated per rank without the reference's packed int32 arrays (automatic beyond their 2^26 limit)")
    ap.add_argument("--partition", default=None, choices=["slabs", "rcb"],
                    help="several GPUs, hex: slabs along z (default; the persistent pc solve synchronises chunk to chunk across ranks) or the "
                         "reference's recursive bisection (x, y, z in turn; pc solve pass by pass)")
    ap.add_argument("--weak-cubes", action="store_true", help="weak-scaling series of cubes n = size * N^(1/3) (round 1) instead of N stacked size^3 blocks")
    ap.add_argument("--no-parity-check", action="store_true")
    ap.add_argument("--no-pdl", action="store_true", help="fused passes fully serialised (no programmatic dependent launch)")
    ap.add_argument("--opt", action="append", default=[], metavar="KEY=VALUE", help="cfdl_set_option(KEY, VALUE) after creation (tuning experiments)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--clock-sampler", default="nvml", choices=["nvml", "smi", "off"], help="how clocks / throttle reasons are sampled during the timed region")
    ap.add_argument("--bind-cores", action="store_true", help="several ranks: pin each rank's host threads to its own 1/local-world share of the cores this process may use")
    ap.add_argument("--e2e-separate", action="store_true", help="e2e through separate upload/solve/download calls instead of cfdl_step_host")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if args.steps is None:
            args.steps = 2
        if args.warmup is None:
            args.warmup = 1
        run_reference(args, rank)
        return
    if args.steps is None:
        args.steps = 30
    if args.warmup is None:
        args.warmup = 3
    args.warmup = max(args.warmup, 3)

    bound = None
    if args.bind_cores and world > 1:
        try:
            lw = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
            cores = sorted(os.sched_getaffinity(0))
            share = len(cores) // lw
            if share >= 2:
                mine = cores[local_rank * share:(local_rank + 1) * share]
                os.sched_setaffinity(0, mine)
                bound = "%d cores from %d" % (len(mine), mine[0])
        except Exception as ex:
            dbg("core binding failed:", ex)
    import cfdl
    if cfdl.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device; the product path has no CPU fallback")
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # weak scaling: the global cube grows with the GPU count so that every GPU keeps ~n^3 cells;
    # the mesh is split by the reference's own RCB (x-slabs, columns, octants on a cube)
    # Several GPUs, hex (default): N blocks of size^3 cells stacked along z — a cavity of depth N with the lid on top, cut into
    # one slab per GPU: every GPU keeps exactly the N=1 work (same cells, same cell size), so the series is a weak-scaling
    # series by construction.  --global-size G: the G^3 cube on N GPUs (strong series of BASELINE.json configs 3 and 5).
    # --weak-cubes: round 1's cubes of edge size * N^(1/3).
    exchange = "NCCL send/recv after every pass, residual norms by NCCL all-reduce"
    partition = args.partition or ("slabs" if args.mesh == "hex" else "rcb")
    nz_global = None
    if args.global_size:
        n_global = args.global_size
    elif world == 1 or args.mesh != "hex":
        n_global = args.n if world == 1 else int(round(args.n * world ** (1.0 / 3.0)))
    elif args.weak_cubes or partition != "slabs":
        n_global = int(round(args.n * world ** (1.0 / 3.0)))
    else:
        n_global, nz_global = args.n, args.n * world
    cells_global = n_global * n_global * (nz_global or n_global)
    structured = args.mesh == "hex" and (args.structured or nz_global is not None or cells_global + 6 * n_global ** 2 >= 2 ** 26)
    if world > 1 and args.solver == "parity":
        raise SystemExit("bench.py: the exact natural-order solver is single-GPU; use --solver mcsgs with --gpus > 1")
    raw = geom = None
    t_setup = time.perf_counter()

    def make_solver(n, nz, use_structured, nranks, rk):
        """One handle of the mesh n x n x (nz or n) (rank rk of nranks), connected to its peers."""
        r = g = None
        if use_structured:  # same mesh, same numbering, generated analytically on every rank (cfdl_create_structured_hex[_slabs])
            sv = cfdl.Solver.structured_hex(n, device=local_rank, rank=rk, nranks=nranks, slabs=(partition == "slabs"), nz=nz)
        else:
            r, g = build_mesh(cfdl, args.mesh, n, device=local_rank)
            if nranks == 1:
                t0 = time.perf_counter()
                sv = cfdl.Solver(g, cfdl.default_bcs(r), device=local_rank)
                SETUP_S["cfdl_create (colouring, renumbering, ELL + statics on the host, upload)"] = round(time.perf_counter() - t0, 2)
            else:
                if partition == "slabs" and args.mesh == "hex":  # cell id = i + n (j + n k)
                    kk = np.arange(n ** 3) // (n * n)
                    c2r = np.zeros(n ** 3, np.int32)
                    for q in range(nranks):
                        c2r[(kk >= n * q // nranks) & (kk < n * (q + 1) // nranks)] = q + 1
                else:
                    c2r, _, _ = cfdl.partition_rcb(g, nranks, want_order=False)
                sv = cfdl.Solver(g, cfdl.default_bcs(r), device=local_rank, cell2rank=c2r, rank=rk, nranks=nranks)
        how = None
        if nranks > 1:
            ids = [cfdl.comm_unique_id() if rk == 0 else None]
            dist.broadcast_object_list(ids, src=0)
            sv.comm_init(ids[0])
        sv.set_option("solver", SOLVERS[args.solver])
        if nranks > 1 and not args.no_p2p:
            try:  # peer-to-peer ghost exchange over NVLink (CUDA IPC); NCCL send/recv stays the fallback
                handles = [None] * nranks
                dist.all_gather_object(handles, sv.ipc_handle())
                sv.ipc_connect(handles)
                how = "peer-to-peer stores into the neighbours' ghost cells + flag words (CUDA IPC over NVLink), residual norms through peer slots"
            except cfdl.CfdlError as ex:
                dbg("p2p unavailable:", ex)
        return sv, r, g, how

    s, raw, geom, how = make_solver(n_global, nz_global, structured, world, rank)
    if how:
        exchange = how
    if args.unfused:
        s.set_option("fused", 0)
    if args.no_pdl:
        s.set_option("pdl", 0)
    for kv in args.opt:
        k, v = kv.split("=")
        s.set_option(k, float(v))
    ne, nf, nbf, H = s.ne, s.nf, s.nbf, s.H  # global sizes

    def step(i):
        s.update_boundaries()
        s.solve_uvwp(DT, NIT)
        if (i + 1) % NCOEF == 0:
            s.update_time()

    def barrier():
        if dist is not None:
            dist.barrier()

    # ---- value: device-resident steps, CUDA events on the library's stream -------------------
    setup_s = time.perf_counter() - t_setup
    dbg("setup %.1f s;" % setup_s, "solver ready; owned", int(s.get_info("owned_cells")), "ghost", int(s.get_info("ghost_cells")))
    # the clock sampler starts before the warm-up and covers the timed region; it samples the GPU under the same load
    # throughout (rank 0 only: one sample stream describes the box)
    sampler = ClockSampler(local_rank, args.clock_sampler)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(i)
        dbg("warmup step", i)
    hist_last = None
    hist_all = []
    barrier()
    sampler.mark()
    s.set_option("reset_counters", 1)
    s.timer_record(0)
    if args.warmup % NCOEF == 0 and args.steps % NCOEF == 0:
        # the reference's loop (main.f90:50-63: ntstep x (ncoef x (update_boundaries; solve_uvwp); update_time)) as ONE C-ABI
        # call, cfdl_run: the same steps as the per-step calls below without a host round trip between them
        hist_all = s.run(dt=DT, nit=NIT, ntstep=args.steps // NCOEF, ncoef=NCOEF)
        hist_last = hist_all[-1]
        loop = "cfdl_run (one call for all timed steps)"
    else:
        for i in range(args.steps):
            s.update_boundaries()
            hist_last = s.solve_uvwp(DT, NIT)
            hist_all.append(hist_last)
            if (args.warmup + i + 1) % NCOEF == 0:
                s.update_time()
        loop = "cfdl_update_boundaries + cfdl_solve_uvwp (+ cfdl_update_time) per step"
    s.timer_record(1)
    ms = s.timer_elapsed_ms(0, 1)
    dbg("timed steps done", ms)
    barrier()
    clocks = sampler.stop()
    launches = int(s.get_info("launches"))
    ms_by_rank = None
    if dist is not None:
        import torch
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        every = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(every, t)
        ms_by_rank = [round(float(x.item()) / args.steps, 4) for x in every]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = ne * args.steps / (ms * 1e-3)  # ne = cells of the whole (global) mesh over all ranks
    n_owned = int(s.get_info("owned_cells"))
    n_local_faces, n_local_halos = int(s.get_info("local_faces")), int(s.get_info("local_halos"))

    # ---- roofline: per-launch CUDA-event timing of the dominant kernel over the same steps ----
    peak, peak_src = peaks()
    s.set_option("profile", 1)
    s.set_option("reset_counters", 1)
    nprof = min(args.steps, 6)
    for i in range(nprof):
        step(args.warmup + args.steps + i)
    kinds = ("sgs", "residual", "residual3", "levels", "coef_uvw", "coef_p", "mip", "grad", "grad1", "pcg", "sgs3")
    prof = {k: (s.get_info("prof_ms_" + k), int(s.get_info("prof_n_" + k))) for k in kinds}
    s.set_option("profile", 0)
    # fused passes are launched back to back and overlap head/tail (programmatic dependent launch);
    # bracketing every launch with its own event pair serialises them.  Second instrumented pass:
    # one event pair around each batch of consecutive passes, divided by the passes that did work.
    batch = None
    try:
        s.set_option("profile", 2)
        s.set_option("reset_counters", 1)
        for i in range(nprof):
            step(args.warmup + args.steps + nprof + i)
        bms, bn = s.get_info("prof_ms_sgs"), int(s.get_info("prof_n_sgs"))
        if bn > 0 and bms > 0:
            batch = (bms, bn)
    except cfdl.CfdlError as ex:
        dbg("batch profiling unavailable:", ex)
    s.set_option("profile", 0)
    rbq_dist = int(s.get_info("rbq_dist")) if world > 1 else 0
    passes_per_step = (batch[1] / nprof) if batch else None  # two-colour passes that did work, per SIMPLE iteration (pc + momentum one-by-one)
    K = int(s.get_info("ell_width"))  # faces per cell (hex 6, tet 4: no ELL padding)
    ab = algorithmic_bytes(n_owned, K * n_owned, n_owned + int(s.get_info("ghost_cells")) + n_local_halos)  # this rank's partition
    ncol = int(s.get_info("ncolors"))
    roof = None
    extra_roof = {}
    fused = ncol == 2 and args.solver != "parity" and not args.unfused
    if prof["sgs"][1] > 0:
        if fused:
            # fused two-colour passes: a red pass reads ap,b,anb,idx (16+12K B/row), its own value,
            # gathers the black values, writes mid+new; a black pass gathers two red arrays and writes one value.
            n_r = n_owned // 2
            halo_vals = int(s.get_info("ghost_cells")) + n_local_halos
            # pc passes (nearly all of them: the momentum passes have their own entry when they run side by side) rebuild
            # ap from the row's anb instead of reading it (pc_sumap): 8 + 12K bytes of matrix per row instead of 16 + 12K
            # chosen form of the pc passes (candidate c: persistent = c & 1, L2 hint = (c >> 1) % 3 > 0, 16-bit neighbour offsets = c >= 6)
            rbq = int(s.get_info("rbq_active")) == 1 if world == 1 else rbq_dist == 1
            i16 = int(s.get_info("rb_idx16")) == 1 and not (world > 1 and rbq_dist == 1)  # the partitioned persistent form reads 32-bit positions
            row = (8 if int(s.get_info("pc_sumap")) else 16) + (10 if i16 else 12) * K
            red = n_r * (row + 8 + 16) + 8 * (n_owned - n_r + halo_vals)
            black = (n_owned - n_r) * (row + 8 + 8) + 16 * (n_r + halo_vals)
            per_launch = (red + black) / 2.0
            kname = "rb_red_kernel / rb_black_kernel (fused two-colour SGS pass incl. residual; average of the two)"
            kkey = "rb_pass_i16" if i16 else "rb_pass_i32"
            if rbq:
                kkey = ("rbq_i16" if i16 else "rbq_i32") + ("_counter" if int(s.get_info("rbq_chunks")) > int(s.get_info("rbq_grid")) else "")
                kname = ("rbq_kernel (all red/black passes of a pc solve in one persistent launch, neighbour-only synchronisation; "
                         "launch time / passes; the 'isolated' figure is the pass-by-pass kernels rb_red_kernel / rb_black_kernel, one event pair per launch)")
        else:
            # one colour launch updates 1/ncolors of the cells: a full sweep = ncolors launches
            per_launch = ab["sgs_sweep"] / ncol
            kname = "sgs_range_kernel (one colour of a Gauss-Seidel sweep)"
            kkey = "sgs_range"
        iso_ms = prof["sgs"][0] / prof["sgs"][1]
        iso = per_launch / (iso_ms * 1e-3) / 1e9
        if fused and batch:  # average duration of a pass inside its batch, as it runs in the timed steps
            avg_ms, timed, how = batch[0] / batch[1], batch[1], "CUDA events around each batch of consecutive passes / passes that did work"
        else:
            avg_ms, timed, how = iso_ms, prof["sgs"][1], "one CUDA-event pair per launch"
        ach = per_launch / (avg_ms * 1e-3) / 1e9
        roof = {"kernel": kname, "bound": "hbm", "achieved": ach, "peak": peak,
                "unit": "GB/s", "frac": ach / peak, "traffic": ncu_traffic(kkey, args, world), "peak_source": peak_src,
                "bytes_per_launch": per_launch, "avg_launch_ms": avg_ms, "launches_timed": timed, "timing": how,
                "isolated": {"achieved": iso, "frac": iso / peak, "avg_launch_ms": iso_ms, "launches_timed": prof["sgs"][1],
                             "timing": "one CUDA-event pair per launch (launches serialised, no overlap)"}}
        if roof["traffic"]:
            # `achieved` counts ALGORITHMIC bytes; the persistent pc solve keeps its value arrays in the L2, so the DRAM moves fewer
            # (the ncu figure): frac can approach or pass 1 while the DRAM itself is at dram_frac of its measured copy rate
            roof["dram_achieved"] = roof["traffic"] / (avg_ms * 1e-3) / 1e9
            roof["dram_frac"] = roof["dram_achieved"] / peak
            roof["note"] = ("achieved/frac relate the algorithmic bytes of a pass to the measured HBM copy rate; traffic is the DRAM traffic ncu measured "
                            "for the same pass (the L2 serves the rest), dram_achieved/dram_frac relate that to the same peak")
    if prof["residual"][1] > 0:
        # residual_kernel only (one right-hand side: 20N + 12Z + 8H); the three-RHS residual3_kernel of the
        # side-by-side momentum solves and its cross-rank reduction are timed under their own kind (residual3)
        avg_ms = prof["residual"][0] / prof["residual"][1]
        ach = ab["residual"] / (avg_ms * 1e-3) / 1e9
        extra_roof["residual_spmv"] = {"kernel": "residual_kernel (SpMV + norm, one right-hand side)", "achieved": ach, "frac": ach / peak, "unit": "GB/s",
                                       "bytes_per_launch": ab["residual"], "avg_launch_ms": avg_ms, "launches_timed": prof["residual"][1]}
    if prof["residual3"][1] > 0 and world == 1:
        # residual3_kernel: the matrix row once (16 + 12K B), per equation b + own value (16 B) and the gathered values (8 B per cell or halo)
        per3r = n_owned * (16 + 12 * K + 3 * 16) + 3 * 8 * (n_owned + n_local_halos)
        avg_ms = prof["residual3"][0] / prof["residual3"][1]
        ach = per3r / (avg_ms * 1e-3) / 1e9
        extra_roof["residual3_spmv"] = {"kernel": "residual3_kernel (SpMV + norm for u, v, w in one pass over the matrix)", "achieved": ach, "frac": ach / peak,
                                        "unit": "GB/s", "bytes_per_launch": per3r, "avg_launch_ms": avg_ms, "launches_timed": prof["residual3"][1]}
    # assembly kernels against SURVEY 8(d)'s algorithmic bytes (reference layout: every distinct input element once, every output once)
    Fl, Hl, Zl = n_local_faces, n_owned + int(s.get_info("ghost_cells")) + n_local_halos, K * n_owned
    asm_bytes = {"coef_uvw": ("calc_coef_uvw", 196 * n_owned + 16 * Zl + 56 * Fl + 48 * Hl),
                 "coef_p": ("calc_coef_p", 36 * n_owned + 16 * Zl + 56 * Fl + 24 * Hl),
                 "mip": ("calc_mip", 124 * n_owned + 72 * Fl),
                 "grad": ("calc_grad x3 in one pass (geometry and connectivity once: 3(28N+4Z+32H) - 2(4Z+24H))", 3 * (28 * n_owned + 4 * Zl + 32 * Hl) - 2 * (4 * Zl + 24 * Hl)),
                 "grad1": ("calc_grad", 28 * n_owned + 4 * Zl + 32 * Hl)}
    for kk, (nm, nbytes) in asm_bytes.items():
        if prof[kk][1] > 0:
            avg_ms = prof[kk][0] / prof[kk][1]
            ach = nbytes / (avg_ms * 1e-3) / 1e9
            extra_roof[kk] = {"kernel": nm, "achieved": ach, "frac": ach / peak, "unit": "GB/s", "bytes_per_launch": nbytes,
                              "avg_launch_ms": avg_ms, "launches_timed": prof[kk][1]}
    if prof["sgs3"][1] > 0:
        # u, v, w side by side (kernels_rb3.inc): a pass reads a row's ap, anb, ids once (16+12K B) and per equation
        # b + own value (16 B), writes 8 B (red passes 16 B) and gathers the other colour's values (8 B per cell, 16 in black passes)
        n_r = n_owned // 2
        per3 = (n_r * (16 + 12 * K + 3 * (16 + 16 + 8)) + (n_owned - n_r) * (16 + 12 * K + 3 * (16 + 8 + 16))) / 2.0
        avg_ms = prof["sgs3"][0] / prof["sgs3"][1]
        ach = per3 / (avg_ms * 1e-3) / 1e9
        extra_roof["momentum_side_by_side_pass"] = {"achieved": ach, "frac": ach / peak, "unit": "GB/s", "bytes_per_launch": per3,
                                                    "avg_launch_ms": avg_ms, "launches_timed": prof["sgs3"][1]}
    if roof is None and prof["levels"][1] > 0:
        # parity mode: one persistent launch = one (or two) symmetric iterations over all levels
        it_per_launch = 2 if (args.solver == "parity" and False) else 1
        avg_ms = prof["levels"][0] / prof["levels"][1]
        per_launch = 2 * ab["sgs_sweep"] * it_per_launch
        ach = per_launch / (avg_ms * 1e-3) / 1e9
        roof = {"kernel": "level_sgs_kernel (persistent level-scheduled exact SGS, latency-bound by design)", "bound": "hbm",
                "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": None, "peak_source": peak_src,
                "bytes_per_launch": per_launch, "avg_launch_ms": avg_ms, "launches_timed": prof["levels"][1]}
    # per-phase time of a step: kernels bracketed one by one (profile = 1), except the two-colour passes, which are taken from
    # the batch timing (they overlap / run inside one persistent launch in production; one event pair per launch would serialise them)
    phase_ms = {k: prof[k][0] / nprof for k in kinds if prof[k][1] > 0}
    if fused and batch:
        phase_ms["sgs"] = batch[0] / nprof
    phase_ms["sum_of_phases"] = sum(phase_ms.values())

    # ---- e2e: same steps through the C ABI with host buffers inside the timed region ----------
    e2e = None
    if not args.no_e2e:
        ins = "u v w p u0 v0 w0 gu gv gw gp mip mip0".split()
        # everything solve_uvwp writes that the next call reads back (or that the driver's output needs): the host
        # arrays stay authoritative, so the run through host buffers is the same simulation as the resident one
        outs = "u v w p gu gv gw gp gpc mip".split()
        # one GPU: host arrays in the reference's numbering (cfdl_upload_field / cfdl_download_field);
        # several GPUs: every rank moves its own partition (cfdl_*_field_local), bytes summed over ranks
        if world == 1:
            size_of, up, down = s.field_size, s.upload, s.download_into
        else:
            size_of, up, down = s.local_size, s.upload_local, s.download_local
        bufs = {k: cfdl.PinnedBuffer(size_of(k)) for k in set(ins + outs)}
        for k in ins:
            down(k, bufs[k].array)
        h2d = 8 * sum(size_of(k) for k in ins)
        d2h = 8 * sum(size_of(k) for k in outs) + 8 * 16
        if dist is not None:
            import torch
            t = torch.tensor([h2d, d2h], dtype=torch.float64, device="cuda")
            dist.all_reduce(t)
            h2d, d2h = int(t[0].item()), int(t[1].item())
        ne2e = min(args.steps, 10)
        base = args.warmup + args.steps + 2 * nprof

        # one C-ABI call per step (cfdl_step_host: the transfers that the iteration allows run beside the
        # computation); the separate upload/solve/download calls remain as the fallback and are named in the line
        e2e_path = {"how": "cfdl_step_host (mip0 uploaded during the momentum phase; u,v,w,gu,gv,gw downloaded during the pc solve)"}
        in_arrays = {k: bufs[k].array for k in ins}
        out_arrays = {k: bufs[k].array for k in outs}

        def e2e_step(i):
            if "failed" not in e2e_path and not args.e2e_separate:
                try:
                    s.step_host(in_arrays, out_arrays, dt=DT, nit=NIT, apply_bcs=True, local=(world > 1))
                except cfdl.CfdlError as ex:
                    e2e_path["failed"] = str(ex)
                    e2e_path["how"] = "separate cfdl_upload_field / cfdl_solve_uvwp / cfdl_download_field calls"
            if "failed" in e2e_path or args.e2e_separate:
                for k in ins:
                    up(k, bufs[k].array)
                s.update_boundaries()
                s.solve_uvwp(DT, NIT)
                for k in outs:
                    down(k, bufs[k].array)
            if (i + 1) % NCOEF == 0:
                s.update_time()
                for k in ("u0", "v0", "w0", "mip0"):
                    down(k, bufs[k].array)

        e2e_step(base)
        barrier()
        t0 = time.perf_counter()
        s.timer_record(2)
        for i in range(ne2e):
            e2e_step(base + 1 + i)
        s.timer_record(3)
        ms_e = s.timer_elapsed_ms(2, 3)
        wall_e = (time.perf_counter() - t0) * 1e3
        ms_e = max(ms_e, wall_e)  # the host side of the copies counts too
        if dist is not None:
            import torch
            t = torch.tensor([ms_e], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms_e = float(t.item())
        e2e = {"value": ne * ne2e / (ms_e * 1e-3), "unit": "cell-iterations/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": d2h, "steps": ne2e, "ms_per_step": ms_e / ne2e,
               "what": "per step: upload u,v,w,p,u0,v0,w0,gu,gv,gw,gp,mip,mip0 from pinned host memory, update_boundaries + "
                       "solve_uvwp through the C ABI, download u,v,w,p,gu,gv,gw,gp,gpc,mip and the residual history",
               "path": ("separate cfdl_upload_field / cfdl_solve_uvwp / cfdl_download_field calls" if args.e2e_separate else e2e_path["how"]),
               "path_error": e2e_path.get("failed")}
        for b in bufs.values():
            b.free()

    # ---- parity_check (several GPUs): the partitioned run against a single-GPU run of the same mesh ------------------------------
    # fresh handles, ncoef = 3 SIMPLE iterations + update_time from the initial state; every rank reports the cells it owns, rank 0
    # runs the whole mesh on its own GPU and compares u, v, w, p and the solver iteration counts.  The bench mesh itself up to
    # 5 M cells, beyond that the same shape with 64^2 cells per plane (same code paths: partition, exchange mode, persistent solve).
    parity = None
    if world > 1 and not args.no_parity_check:
        import torch
        try:
            pn, pnz = n_global, nz_global
            if cells_global > 5000000:
                pn, pnz = 64, (64 * world if nz_global else None)
            sp, _, _, _ = make_solver(pn, pnz, structured or pnz is not None, world, rank)
            for kv in args.opt:
                k, v = kv.split("=")
                sp.set_option(k, float(v))
            hist_p = sp.run(dt=DT, nit=NIT, ntstep=1, ncoef=NCOEF)
            mine = {}
            for f in ("u", "v", "w", "p"):
                a = np.full(sp.field_size(f), np.nan)
                sp.download_into(f, a)  # writes only the entries this rank owns
                own = ~np.isnan(a)
                t = torch.from_numpy(np.where(own, a, 0.0)).cuda()
                c = torch.from_numpy(own.astype(np.float64)).cuda()
                dist.all_reduce(t); dist.all_reduce(c)  # every entry is owned by exactly one rank: x + 0 + ... + 0 = x
                mine[f] = (t.cpu().numpy(), c.cpu().numpy())
            rbq_dist_par = int(sp.get_info("rbq_dist"))
            sp.close()
            if rank == 0:
                one, _, _, _ = make_solver(pn, pnz, structured or pnz is not None, 1, 0)
                for kv in args.opt:
                    k, v = kv.split("=")
                    one.set_option(k, float(v))
                hist_1 = one.run(dt=DT, nit=NIT, ntstep=1, ncoef=NCOEF)
                worst, owned_once = 0.0, True
                for f in ("u", "v", "w", "p"):
                    w = one.download(f)
                    owned_once = owned_once and bool(np.all(mine[f][1] == 1.0))
                    worst = max(worst, float(np.abs(mine[f][0] - w).max() / max(np.abs(w).max(), 1e-300)))
                one.close()
                same_it = bool(np.array_equal(hist_p[:, :, 0], hist_1[:, :, 0]))
                parity = {"result": "ok" if (worst <= 1e-12 and same_it and owned_once) else "fail", "max_rel_err(u,v,w,p)": worst,
                          "iteration_counts_equal": same_it, "every_cell_reported_once": owned_once, "tolerance": 1e-12,
                          "mesh": "%dx%dx%d" % (pn, pn, pnz or pn), "steps": NCOEF, "persistent_partitioned_pc_solve": rbq_dist_par == 1,
                          "what": "%d-GPU run vs a single-GPU run of the same mesh on rank 0 (fresh handles, %d SIMPLE iterations from rest)" % (world, NCOEF)}
        except Exception as ex:  # the check must not take the measurement down with it
            parity = {"result": "error", "error": repr(ex)[:300]}
        barrier()

    # ---- cpu_baseline: the oracle on a bounded sample of the same workload (rank 0, N=1) -------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline and n_global ** 3 + 6 * n_global ** 2 < 2 ** 26:
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import oracle
        if raw is None:
            raw, geom = build_mesh(cfdl, args.mesh, n_global)
        oc = oracle.OracleCase(raw, n_subdomains=4, geom=geom)
        nsamp = 2 if ne > 500000 else 6
        _, sec = oc.run(1, nsamp)
        cpu = {"value": ne * nsamp / sec, "unit": "cell-iterations/s", "cores": 1, "kind": "port",
               "sample": "%d SIMPLE iterations of the same mesh from the initial state, reference defaults incl. n_subdomains=4; "
                         "C++ oracle, 1 thread (the reference is serial Fortran, not buildable here)" % nsamp,
               "seconds": sec}
        del oc

    if rank == 0:
        line = {"metric": "cell-iterations/s (SIMPLE)", "value": value, "unit": "cell-iterations/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": workload_name(args.mesh, n_global, ne, nz_global), "solver": args.solver, "ncolors": ncol, "options": args.opt, "mesh_source": "structured per-rank generator" if structured else "reference-format arrays", "fused_two_colour_passes": bool(fused), "timed_loop": loop, "programmatic_dependent_launch": bool(fused and not args.no_pdl),
                           "cells_per_gpu": ne // world, "ms_per_step_by_rank": ms_by_rank, "host_cores": bound or "not pinned", "host_cpus": os.cpu_count(), "setup_seconds": round(setup_s, 1), "setup_breakdown_seconds": dict(SETUP_S), "l2": "working set (>1 GB per GPU) exceeds the 126 MB L2; no flush needed" if ne // world > 1000000 else
                           "working set may fit L2",
                           "parallelism": "1 GPU" if world == 1 else
                           "%d GPUs, one %s of the global mesh per GPU; ghost-cell exchange: %s" % (world, "z-slab" if partition == "slabs" else "RCB block", exchange),
                           "pc_solve": ("one persistent launch per rank, chunks synchronise with neighbouring chunks only (across NVLink too), residual history combined once per launch"
                                        if rbq_dist == 1 else ("one persistent launch, neighbour-only synchronisation" if (world == 1 and fused and int(s.get_info("rbq_active")) == 1) else "one launch per pass")),
                           "pc_solve_chunks": ({"chunks_per_colour": int(s.get_info("rbq_chunks")), "ctas": int(s.get_info("rbq_grid")), "rows_per_chunk": int(s.get_info("rbq_chunk_rows")),
                                                "assignment": "handed out in order from a counter" if int(s.get_info("rbq_chunks")) > int(s.get_info("rbq_grid")) else "one chunk per CTA"}
                                               if int(s.get_info("rbq_chunks")) > 0 else None),
                           "passes_per_step": passes_per_step,
                           "solver_iterations_last_step(u,v,w,pc)": [int(x) for x in hist_last[:, 0]] if hist_last is not None else None,
                           "solver_iterations_mean_over_timed_steps(u,v,w,pc)": [round(float(x), 2) for x in np.asarray(hist_all)[:, :, 0].mean(axis=0)] if len(hist_all) else None,
                           "pc_residual_reduction_mean(res_f/res_i)": round(float(np.mean([h[3, 2] / h[3, 1] for h in np.asarray(hist_all) if h[3, 1] > 0])), 4) if len(hist_all) else None,
                           "last_step_history(it,res_i,res_f,res_max)": hist_last.tolist() if hist_last is not None else None},
                "roofline": roof, "roofline_other": extra_roof, "phase_ms_per_step": phase_ms, "cpu_baseline": cpu, "e2e": e2e,
                "parity_check": parity,
                "gpu_launches": launches, "clocks": clocks}
        print(json.dumps(line), file=_JSON_OUT, flush=True)
    s.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
