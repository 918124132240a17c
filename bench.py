#!/usr/bin/env python
"""bench.py -- linearised-NS DOF*timesteps/s of the nekStab matvec hot path on N B200s (one rank per GPU).

Workload (BASELINE.json configs[4]): the shipped 2-D cylinder mesh (1996 elements, Re=50 base flow,
tests/golden/cyl.npz) re-interpolated to lx1=8 and extruded into periodic z-layers, lxd=12, lx2=6, BDF3/EXT3, dt from
CFL 0.5, tolerances 1e-8/1e-8, seed = nekStab's deterministic noise (core/utils.f:344-408) normalised in the energy norm.
A "step" is one linearised time step: dealiased advection + 3 Helmholtz solves + one pressure solve + projection.
The timed region is one forward_linearized_map call (core/matvec.f:163) of K steps after a warm-up call of W steps.

One JSON line (rank 0) carries
  * the headline `value`: WEAK scaling, 19 960 hexahedra = 1.02e7 grid points per GPU (10*N layers over Lz = 2 pi N), timed with
    the sampling profiler OFF (CUDA graphs on); per-kernel averages come from a separate, sampled call (`roofline.kernels`);
  * `strong` (N > 1): the named fixed-size mesh (19 960 hexahedra in total) split over the N GPUs, same steps;
  * `arnoldi`: M Arnoldi iterations of the Krylov loop on the weak workload (one matvec = `nsteps` steps to T = 1, then the
    CGS2/DGKS orthogonalisation), the matvec / orthogonalisation split, the roofline entry of the tall-skinny GEMV pair at
    k = k_dim and the wall time to converged eigenpairs (measured once by tools/run_arnoldi_cfg5.py, profiles/);
  * `parity_n` (N > 1): relative energy-norm difference between a 2-step matvec computed on N GPUs and on rank 0 alone on
    the same small global mesh;
  * `cpu_baseline` (N = 1) / `--impl reference`: the C / OpenMP restatement of the same step (oracle/cport.c) run to the SAME
    tolerances from the SAME seed for the SAME warm-up and timed steps on a bounded sample of the workload (the full 2-D mesh,
    fewer z-layers), with its own measured iteration counts printed next to the GPU's.
"""
from __future__ import annotations

import argparse
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from nekstab_b200 import cases  # noqa: E402

PEAKS_FILE = os.path.join(ROOT, "MEASURED_PEAKS.json")
HBM_FALLBACK_GBS = 6650.0      # B200_PROFILING.md fallback
METRIC, UNIT = "linearized-NS DOF*timesteps/s", "DOF*steps/s"

# Algorithmic traffic model, words (8 B) per velocity grid point, 3-D lx1=8 (DESIGN.md "Kernels"):
R2 = (6.0 / 8.0) ** 3           # n2/n
FS = 1.0 - (6.0 / 8.0) ** 3     # surface-node fraction
WORDS = {
    # pressure-CG iteration pieces
    "pcg_gradt": (3 + 9) * R2 + R2 + 3.0,            # read z,(1),pdir + 9 metrics (mesh 2); write pdir; write 3 fields
    "dssum": 3 * 2 * FS + 0.5 * FS,                  # surface values R+W for 3 fields + int32 index
    "pcg_div": 3.0 + 1.0 + (9 + 1 + 1) * R2,         # read 3 fields + mask*binv; 9 metrics + pdir read, Ep write (mesh 2)
    "pcg_update": 8 * R2,                            # x,p,r,Ep,dinvE,bm2inv read; x,r write
    # three-level preconditioner (csrc/pmg.cu): restriction reads r; the element-block kernel reads r and 126 words of FDM
    # factors per element (216 points) and writes z; the vertex/aggregate levels move O(nel) words (latency-bound launches)
    "pcg_pc_restrict": 1.0 * R2,
    "pcg_pc_coarse": 0.0,
    "pcg_pc_apply": (2 + 126.0 / 216.0) * R2,
    # fused CG tail (k_pcg_fused): r, Ep, p, x, 1/bm2 read, r, x, zloc written + the FDM factors; the direction kernel then reads
    # zloc and pdir (no separate z, no vector of ones)
    "pcg_fused": (5 + 3 + 126.0 / 216.0) * R2,
    "pcg_gradt_fused": (2 + 9) * R2 + R2 + 3.0,
    # Helmholtz-CG iteration pieces (3 components batched)
    "hcg_axhelm": 3 * (1 + 1 + 1 + 1) + 6 + 1 + 1,   # per comp r,p read, p,w write; 6 G + bm1 + dinv once
    "hcg_dssum": 3 * 2 * FS + 0.5 * FS,
    "hcg_update": 3 * (4 + 2) + 3 + 3,               # per comp x,p,r,w read x,r write; mask, dinv, mult, binv
}
# SURVEY 8d contract figures (per component / per call) for the same pieces, reported next to the model above
SURVEY_WORDS = {"helmholtz_cg_iteration_per_component": 19.45, "E_apply": 21.2, "pcg_vectors": 9 * R2, "ADV": 39.4, "RHS": 26.0,
                "RES": 20.6, "PCOR": 15.0}


def cyl2d():
    g = np.load(os.path.join(ROOT, "tests", "golden", "cyl.npz"))
    return cases.cylinder_case(g, lx1=8, sponge=False)            # sponge off for benchmarks


def build_workload(nz: int, world: int = 1, rank: int = 0, name: str = ""):
    """This rank's share of the synthetic 3-D cylinder-wake mesh with `nz` periodic layers over Lz = 2 pi nz/10.  The 2-D mesh
    is partitioned by Nek5000's rule on the shipped RSB keys; every rank extrudes its own 2-D elements (global node ids stay
    consistent: id2d*levels + level).  Returns (case, global grid points)."""
    c2 = cyl2d()
    nel_glob = c2.nel * nz
    if world > 1:
        c2 = c2.local_part(rank, world)
    c3 = cases.extrude(c2, nz, 2 * np.pi * nz / 10.0, name=name or f"cyl3d_1996x{nz}_lx8", compress_ids=False)
    c3.nelg = nel_glob
    return c3, nel_glob * 512


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons DURING the timed region (B200_PROFILING.md).  In-process NVML (the library nvidia-smi itself
    queries) every 50 ms; falls back to spawning nvidia-smi when pynvml is missing.  Spawning a process per sample steals host time
    from the thread that polls the CG loops, NVML does not."""
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, device):
        super().__init__(daemon=True)
        self.device, self.stop_flag, self.rows, self.how = device, False, [], "nvidia-smi"
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[device]) if vis and vis.split(",")[device].strip().isdigit() else device
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.nv, self.how = pynvml, "nvml"
        except Exception:
            self.nv = None

    def _nvml_row(self):
        nv = self.nv
        sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
        mx = nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM)
        r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
        bits = [getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8), getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4)]
        return [str(sm), str(mx)] + ["Active" if (r & b) else "Not Active" for b in bits]

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                if self.nv is not None:
                    self.rows.append(self._nvml_row())
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.device}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        self.rows.append([t.strip() for t in out.split(",")])
            except Exception:
                pass
            time.sleep(0.05)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        reasons = [n for i, n in enumerate(self.NAMES) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows), "how": self.how}


KERNEL_FILES = ("elem_common.cuh", "elem_kernels.cu", "pcg_kernels.cu", "pmg.cu", "vec_kernels.cu", "gs.cu")


def kernel_source_hash():
    """sha1 of the CUDA files that hold the single-GPU kernels of the step (what the ncu captures profile); stamped into
    profiles/ncu_dram_traffic.json by tools/ncu_summarise.py: a mismatch means the committed capture predates the current kernels."""
    h = hashlib.sha1()
    d = os.path.join(ROOT, "nekstab_b200", "csrc")
    for f in KERNEL_FILES:
        with open(os.path.join(d, f), "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()[:16]


NCU_NAMES = {"pcg_div": "k_div3q", "pcg_gradt": "k_gradt3<8, 2>", "dssum": "k_gs_sum", "pcg_update": "k_pcg_update",
             "hcg_axhelm": "k_axhelm3p", "hcg_update": "k_hcg_update", "pcg_pc_apply": "k_pm_apply2", "pcg_fused": "k_pcg_fused_p"}


def ncu_traffic(kernel_kind):
    """(dram__bytes_read+write per launch, provenance) of a kernel from the committed `ncu --set full` capture."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_dram_traffic.json")) as f:
            d = json.load(f)
        meta = d.get("_meta", {})
        key = next(k for k in d if k != "_meta" and k.startswith(NCU_NAMES[kernel_kind]))
        recs = d[key]
        nbytes = float(np.median([r["dram_bytes"] for r in recs]))
        src = {"report": meta.get("report"), "source_hash": meta.get("source_hash"), "kernel": key, "launches_in_capture": len(recs),
               "stale": meta.get("source_hash") not in (None, kernel_source_hash())}
        return nbytes, src
    except Exception:
        return None, None


def hbm_peak():
    try:
        with open(PEAKS_FILE) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md)"


def pressure_solver_name(precond):
    return ("Jacobi-PCG (north-star)" if precond == "jacobi" else
            "PCG, three-level additive preconditioner (FDM element blocks + Q1 vertex-mesh Jacobi + aggregate coarse solve)")


def workload_config(world, precond, tol, small=False):
    """The part of `config` that names the workload: identical in the GPU arm and in the `--impl reference` arm."""
    nz = 3 if small else 10 * world
    return {"workload": "cyl3d_small_1996x3_lx8" if small else f"cyl3d_1996x{nz}_lx8", "elements": 1996 * nz, "dof": 1996 * nz * 512,
            "lx1": 8, "lxd": 12, "lx2": 6, "re": 50.0, "tol_v": tol, "tol_p": tol, "pressure_solver": pressure_solver_name(precond)}


# ------------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(nsteps: int, nwarm: int, nz: int, max_seconds: float, precond: str = "pmg", tol: float = 1e-8):
    """The CPU restatement of the step (oracle/cport.c, C / OpenMP on every host core; same algorithm, preconditioner, stopping
    norms and tolerances as the GPU path) on the full 2-D cylinder mesh extruded to `nz` periodic layers: a warm-up call of
    `nwarm` steps and a timed call of `nsteps` steps from the same normalised noise seed and dt as the GPU arm.  The solvers run
    to tolerance: the iteration counts are MEASURED here and printed next to the GPU's.  The timed call stops at a step
    boundary once `max_seconds` are spent.  Returns a dict (value in DOF*steps/s)."""
    from oracle.ops import SEM
    from oracle import cport
    t_setup = time.perf_counter()
    thr = cport.set_threads(0)
    c3, _ = build_workload(nz)
    s = SEM(3, 8, c3.xyz, c3.glo, c3.mask)
    pc = None
    if precond == "pmg":
        from oracle.pmg import PMG
        cp0 = cport.CPort(s, None)
        pc = PMG(s, nagg=max(1, min(512, c3.nel // 32)), apply_e=lambda p: cp0.cdabdtp(p).reshape(s.eshape2))
    st = cport.CStepper(s, c3.ubase, c3.re, None, tol_v=tol, tol_p=tol, max_iter_v=2000, max_iter_p=100000, ifvcor=False, pmg=pc)
    v = cases.add_noise(c3).reshape((3,) + s.eshape)
    v = v / np.sqrt(sum(float(np.sum(v[k] * s.bm1 * v[k])) for k in range(3)))        # krylov_normalize (bm1s = bm1: no sponge)
    p = np.zeros(s.eshape2)
    ctarg = s.cfl_sum(c3.ubase.reshape((3,) + s.eshape))
    ns = int(np.ceil(1.0 / (0.5 / ctarg)))
    dt = 1.0 / ns                                                                      # core/matvec.f:28-36
    t_setup = time.perf_counter() - t_setup
    warm_done = 0
    if nwarm > 0:
        v, p = st.linearized_map(v, p, nwarm, dt, budget_s=max_seconds * 0.5)
        warm_done = st.steps_done
    st.iters_v, st.iters_p = [], []
    v, p = st.linearized_map(v, p, nsteps, dt, budget_s=max_seconds)
    done, secs = st.steps_done, st.step_seconds
    tstep = float(np.mean(secs))
    ip = float(np.mean(st.iters_p))
    iv = float(np.mean([np.mean(x) for x in st.iters_v]))
    sample = (f"{c3.nel} hexahedra = the full 1996-element 2-D cylinder mesh x {nz} periodic layers (Lz = 2 pi {nz}/10), lx1=8, "
              f"n={c3.n}; {done} timed step(s) after {warm_done} warm-up step(s) from the normalised noise seed, dt={dt:.6g}, solvers run to "
              f"tol {tol:g} (measured: {ip:.1f} pressure-CG and {iv:.1f} Helmholtz-CG iterations/component/step, preconditioner "
              f"{precond}); C/OpenMP port (oracle/cport.c), {thr} threads; setup {t_setup:.0f} s")
    return {"value": c3.n / tstep, "ms_per_step": tstep * 1e3, "cores": thr, "sample": sample, "points": int(c3.n), "nz": nz,
            "steps_done": done, "warmup_done": warm_done, "pres_iters_per_step": ip, "helm_iters_per_comp_per_step": iv}


# ------------------------------------------------------------------------------------------------ GPU arm
class Plumbing:
    """Host-side plumbing shared by the legs: rank info, barrier, max over ranks (torch.distributed / gloo)."""

    def __init__(self):
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        import torch
        self.torch = torch
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            torch.cuda.set_device(self.local_rank)
            # torch.distributed is host plumbing only (id broadcast, barrier, max over ranks): gloo.  The data path (halo
            # exchange, all-reduces) runs on the library's own communicator over NVLink (nsb_comm_init).
            dist.init_process_group("gloo")
            self.dist = dist
        self.nccl_id = None

    def barrier(self):
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def maxr(self, x):
        if self.dist:
            t = self.torch.tensor([x], dtype=self.torch.float64)
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
            return float(t.item())
        return x

    def comm_id(self):
        from nekstab_b200 import lib
        if self.world > 1 and self.nccl_id is None:
            ids = [lib.nccl_unique_id() if self.rank == 0 else None]
            self.dist.broadcast_object_list(ids, src=0)
            self.nccl_id = ids[0]
        return self.nccl_id


def open_context(pl: Plumbing, case, args, single=False):
    """Library context on this rank's GPU with the bench parameters, the noise seed in slot 0 (normalised) and dt from CFL."""
    from nekstab_b200 import lib
    if single:
        ctx = lib.NekStabB200(case, device=pl.local_rank)
    else:
        ctx = lib.NekStabB200(case, device=pl.local_rank, rank=pl.rank, nranks=pl.world, nccl_id=pl.comm_id())
    ctx.set_params(1.0 / case.re, 1.0, args.tol, args.tol, 2000, 100000)
    ctx.set_projection(args.mxprev)
    if args.precond == "pmg":
        ctx.set_pressure_preconditioner(1, args.nagg)
    dt, nsteps, _ = ctx.prepare_linearized_solver(1.0, 0.5)
    return ctx, dt, nsteps


def seed_slot0(ctx, case, nslots):
    # nekStab's noise seed (core/utils.f:344-408): mth_rand, then the direct-stiffness AVERAGE -- done with the library's
    # (multi-rank) dssum: avg = dssum(q)/dssum(1), then the Dirichlet mask; krylov_normalize
    raw = cases.raw_noise(case)
    mult = ctx.op_dssum(np.ones(case.n))
    seed = np.stack([ctx.op_dssum(raw[k].ravel()) / mult for k in range(3)]).reshape(3, case.nel, -1) * case.mask
    ctx.vec_alloc(nslots)
    ctx.vec_upload(0, seed, None)
    ctx.normalize(0)


def timed_matvec(pl: Plumbing, ctx, dt, K, W, sampler=None):
    """Warm-up call (W steps) from slot 0 into slot 1, timed call (K steps) from slot 1 into slot 2: inputs resident in HBM, the
    sampling profiler off (CUDA-graph replay on).  Device time = CUDA events on the library stream, max over ranks."""
    from nekstab_b200 import lib
    if W > 0:
        ctx.set_timestep(dt, W)
        ctx.matvec(lib.DIRECT, 0, 1)
    else:
        ctx.vec_copy(1, 0)
    ctx.set_timestep(dt, K)
    ctx.stats(reset=True)
    ctx.profile(0)
    use_cuda_profiler = os.environ.get("NSB_CUDA_PROFILER") == "1"    # ncu --profile-from-start off: capture the timed call only
    if use_cuda_profiler:
        pl.torch.cuda.profiler.start()
    if sampler:
        sampler.start()
    pl.barrier()
    t0 = time.perf_counter()
    ctx.matvec(lib.DIRECT, 1, 2)
    pl.barrier()
    wall = time.perf_counter() - t0
    if use_cuda_profiler:
        pl.torch.cuda.profiler.stop()
    if sampler:
        sampler.stop_flag = True
    st = ctx.stats()
    return {"dev_ms": pl.maxr(st["step_ms"]), "wall_s": wall, "pres_iters": st["pres_iters"], "helm_iters": st["helm_iters"],
            "launches": int(st["kernel_launches"])}


def sampled_kernels(ctx, dt, nsteps):
    """A separate call with the sampling profiler on (CUDA events around single launches; graphs off): per-kernel averages."""
    from nekstab_b200 import lib
    ctx.set_timestep(dt, nsteps)
    ctx.profile(1)
    ctx.matvec(lib.DIRECT, 2, 3)                  # slot 3 is scratch: slot 1 stays the input of the timed and the end-to-end calls
    return ctx.profile(0)


def roofline_block(prof, st, K, n_loc, precond, world):
    peak, peak_src = hbm_peak()
    kern = {}
    for kname, (ms, cnt) in prof.items():
        if cnt > 0 and WORDS.get(kname, 0.0) > 0:
            t = ms / cnt * 1e-3
            gbs = WORDS[kname] * 8.0 * n_loc / t / 1e9
            kern[kname] = {"avg_ms": ms / cnt, "samples": cnt, "alg_GBs": gbs, "frac": gbs / peak}
        elif cnt > 0:
            kern[kname] = {"avg_ms": ms / cnt, "samples": cnt}
    fused = "pcg_fused" in kern
    if fused and "pcg_gradt" in kern:                      # the direction kernel no longer reads a vector of ones
        t = kern["pcg_gradt"]["avg_ms"] * 1e-3
        gbs = WORDS["pcg_gradt_fused"] * 8.0 * n_loc / t / 1e9
        kern["pcg_gradt"].update({"alg_GBs": gbs, "frac": gbs / peak})
    names = ("pcg_gradt", "dssum", "pcg_div", "pcg_fused", "pcg_pc_coarse") if fused else \
        ("pcg_gradt", "dssum", "pcg_div", "pcg_update", "pcg_pc_restrict", "pcg_pc_coarse", "pcg_pc_apply")
    wname = lambda k: "pcg_gradt_fused" if (fused and k == "pcg_gradt") else k
    pc = [k for k in names if k in kern]
    cands = [k for k in pc if "frac" in kern[k]]
    if not cands:
        return None
    dom = max(cands, key=lambda k: kern[k]["avg_ms"])
    iter_ms = sum(kern[k]["avg_ms"] for k in pc)
    iter_words = sum(WORDS.get(wname(k), 0.0) for k in pc)
    traffic, tsrc = ncu_traffic(dom) if world == 1 else (None, None)
    roof = {"bound": "hbm", "kernel": dom, "achieved": kern[dom]["alg_GBs"], "peak": peak, "unit": "GB/s", "frac": kern[dom]["frac"],
            "traffic": traffic, "traffic_source": tsrc, "peak_source": peak_src, "alg_bytes_per_launch": WORDS[wname(dom)] * 8.0 * n_loc,
            "pressure_iteration": {"alg_GBs": iter_words * 8.0 * n_loc / (iter_ms * 1e-3) / 1e9,
                                   "frac": iter_words * 8.0 * n_loc / (iter_ms * 1e-3) / 1e9 / peak, "ms": iter_ms,
                                   "note": "sampled (profiler on, graphs off)"},
            "kernels": kern}
    # whole-step roofline: algorithmic bytes of everything a step executes (SURVEY 8d W_step with the measured iteration counts;
    # the advection's HBM traffic = the fine-mesh metrics + fields) over the measured device time per step of the TIMED call
    ip, ih = st["pres_iters"] / K, st["helm_iters"] / K / 3
    iter_w = iter_words if fused else sum(WORDS[k] for k in ("pcg_gradt", "dssum", "pcg_div", "pcg_update")) + \
        (sum(WORDS[k] for k in ("pcg_pc_restrict", "pcg_pc_coarse", "pcg_pc_apply")) if precond == "pmg" else 0.0)
    helm_w = WORDS["hcg_axhelm"] + WORDS["hcg_dssum"] + WORDS["hcg_update"]
    other_w = SURVEY_WORDS["ADV"] + SURVEY_WORDS["RHS"] + SURVEY_WORDS["RES"] + SURVEY_WORDS["PCOR"]
    step_words = ip * iter_w + ih * helm_w + other_w
    step_s = st["dev_ms"] / K * 1e-3
    step_gbs = step_words * 8.0 * n_loc / step_s / 1e9
    sw = SURVEY_WORDS
    survey_words = other_w + 3 * ih * sw["helmholtz_cg_iteration_per_component"] + ip * (sw["E_apply"] + sw["pcg_vectors"] +
                                                                                         (WORDS["pcg_pc_restrict"] + WORDS["pcg_pc_apply"] if precond == "pmg" else R2))
    roof["step"] = {"alg_words_per_point": step_words, "alg_GBs": step_gbs, "frac": step_gbs / peak, "frac_of_nominal_8TBs": step_gbs / 8000.0,
                    "survey_contract_words_per_point": survey_words, "survey_contract_frac": survey_words * 8.0 * n_loc / step_s / 1e9 / peak,
                    "note": "algorithmic bytes of the whole time step / device time per step of the timed call (this rank's points); "
                            "alg_words = this design's minimal traffic (3 components batched), survey_contract = SURVEY 8d per-call figures"}
    return roof


def arnoldi_leg(pl: Plumbing, ctx, case, n_glob, nsteps_full, dt, m_iters, k_dim):
    """M iterations of arnoldi_factorization (core/krylov_decomposition.f:73-99) on the resident basis: wall time per iteration
    ("Time per iteration", :92-99), matvec / orthogonalisation split, then the CGS2/DGKS pass against k_dim resident vectors."""
    from nekstab_b200 import lib
    ctx.set_timestep(dt, nsteps_full)
    H = np.zeros((k_dim + 1, k_dim), order="F")
    ctx.stats(reset=True)
    pl.barrier()
    t0 = time.perf_counter()
    ctx.arnoldi_factorization(lib.DIRECT, 0, H, 1, m_iters, k_dim)
    pl.barrier()
    wall = pl.maxr(time.perf_counter() - t0)
    st = ctx.stats()
    matvec_s = pl.maxr(st["step_ms"]) * 1e-3
    # orthogonalisation against a full basis: slots 1..k_dim filled with copies (bandwidth does not depend on the values)
    for j in range(m_iters + 1, k_dim + 1):
        ctx.vec_copy(j, j % (m_iters + 1))
    ctx.vec_copy(k_dim + 1, 0)
    ctx.orthonormalize(k_dim, 0, k_dim + 1)                       # warm-up
    ctx.profile(1)
    reps = 3
    pl.barrier()
    t0 = time.perf_counter()
    for _ in range(reps):
        ctx.orthonormalize(k_dim, 0, k_dim + 1)
    pl.barrier()
    orth_s = pl.maxr(time.perf_counter() - t0) / reps
    prof = ctx.profile(0)
    peak, _ = hbm_peak()
    n, n2 = case.n, case.nel * 216
    vlen = 3 * n + n2
    md_bytes = (k_dim * 3 * n + 3 * n + n) * 8.0               # Q (velocity part), f, bm1s
    ma_bytes = (k_dim * vlen + 2 * vlen) * 8.0                 # Q, f read, f written
    out = {"k_dim": k_dim, "iterations_timed": m_iters, "nsteps_per_matvec": nsteps_full, "tau": dt * nsteps_full,
           "wall_s_per_iteration": wall / m_iters, "matvec_s_per_iteration": matvec_s / m_iters,
           "other_s_per_iteration": (wall - matvec_s) / m_iters,
           "pres_iters_per_step": st["pres_iters"] / max(st["steps"], 1), "helm_iters_per_comp_per_step": st["helm_iters"] / max(st["steps"], 1) / 3,
           "dof_steps_per_s_in_matvec": n_glob * st["steps"] / max(matvec_s, 1e-30),
           "orthogonalisation_at_k_dim": {"wall_ms": orth_s * 1e3, "alg_bytes": 2 * (md_bytes + ma_bytes),
                                          "alg_GBs": 2 * (md_bytes + ma_bytes) / orth_s / 1e9, "frac": 2 * (md_bytes + ma_bytes) / orth_s / 1e9 / peak,
                                          "survey_ORTH_words": "4 k v + 6 v, v = 3 n (SURVEY 8d)"}}
    for kname, nbytes in (("orth_multidot", md_bytes), ("orth_multiaxpy", ma_bytes)):
        ms, cnt = prof.get(kname, (0.0, 0))
        if cnt:
            out["orthogonalisation_at_k_dim"][kname] = {"avg_ms": ms / cnt, "samples": cnt, "alg_GBs": nbytes / (ms / cnt * 1e-3) / 1e9,
                                                         "frac": nbytes / (ms / cnt * 1e-3) / 1e9 / peak}
    try:      # the one-off full run (tools/run_arnoldi_cfg5.py): iterations needed for converged eigenpairs on this workload
        with open(os.path.join(ROOT, "profiles", "arnoldi_cfg5_summary.json")) as f:
            full = json.load(f)
        out["wall_s_to_k_eigenpairs"] = {"measured_full_run": {k: full.get(k) for k in ("wall_s_arnoldi", "k_dim", "converged_ritz_pairs", "iterations_to_converge",
                                                                                        "leading_lambda", "n_gpus", "source")},
                                         "projected_here": {str(k): v * wall / m_iters for k, v in (full.get("iterations_to_converge") or {}).items()},
                                         "note": "projected = iterations needed in the committed full run x this run's wall time per iteration"}
    except Exception:
        out["wall_s_to_k_eigenpairs"] = None
    return out


def parity_n(pl: Plumbing, args):
    """N > 1: a 2-step matvec on the small global mesh (1996 x 2 layers) computed by rank 0 alone and by all N ranks; relative
    difference in the energy norm, gathered on rank 0.  Tolerances 1e-12 so that the solvers do not hide a wrong halo."""
    from nekstab_b200 import lib
    import copy
    a2 = copy.copy(args)
    a2.tol = 1e-12
    nz, K = 2, 2
    ref = None
    if pl.rank == 0:
        cg, _ = build_workload(nz)
        ctx, dt, _ = open_context(pl, cg, a2, single=True)
        seed_slot0(ctx, cg, 3)
        ctx.set_timestep(dt, K)
        ctx.matvec(lib.DIRECT, 0, 1)
        ref = ctx.vec_download(1)[0].reshape(3, nz, 1996, 512)
        bm1 = ctx.get_field("bm1").reshape(nz, 1996, 512)
        ctx.close()
    pl.barrier()
    c, _ = build_workload(nz, pl.world, pl.rank)
    ctx, dt, _ = open_context(pl, c, a2)
    seed_slot0(ctx, c, 3)
    ctx.set_timestep(dt, K)
    ctx.matvec(lib.DIRECT, 0, 1)
    mine = ctx.vec_download(1)[0]
    plane = "p2p" if ctx.get_field("p2p")[0] > 0 else "nccl"
    ctx.close()
    part = cases.partition(cyl2d().key, pl.world, cyl2d().d2)
    sel = np.nonzero(part == pl.rank)[0]
    gathered = [None] * pl.world
    pl.dist.gather_object((sel, mine.reshape(3, nz, sel.size, 512)), gathered if pl.rank == 0 else None, dst=0)
    if pl.rank != 0:
        return None, plane
    full = np.zeros_like(ref)
    for s_, v_ in gathered:
        full[:, :, s_, :] = v_
    num = float(np.sqrt(np.sum((full - ref) ** 2 * bm1[None])))
    den = float(np.sqrt(np.sum(ref ** 2 * bm1[None])))
    return num / den, plane


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--small", action="store_true", help="tiny 3-layer mesh (debugging / contract tests only; not a valid bench line)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--tol", type=float, default=1e-8)
    ap.add_argument("--precond", default=os.environ.get("NSB_BENCH_PRECOND", "pmg"), choices=["jacobi", "pmg"],
                    help="pressure-CG preconditioner: jacobi (north-star) or pmg (FDM element blocks + vertex-mesh Jacobi + "
                         "aggregate coarse solve, csrc/pmg.cu: the reference's class of preconditioner)")
    ap.add_argument("--nagg", type=int, default=0)
    ap.add_argument("--mxprev", type=int, default=0,
                    help="pressure residual projection size (reference: residualProj=yes, mxprev=20). Measured on this workload "
                         "(noise-seeded first steps): 20 -> 4317 its/step vs 2427 without, so the bench default is 0 = off")
    ap.add_argument("--scaling", default="both", choices=["weak", "strong", "both"],
                    help="N > 1: weak = 19 960 hexahedra per GPU (headline), strong = 19 960 in total; both = weak headline + `strong` object")
    ap.add_argument("--arnoldi", type=int, default=int(os.environ.get("NSB_BENCH_ARNOLDI", "1")),
                    help="Arnoldi iterations timed in the `arnoldi` leg (0 = skip)")
    ap.add_argument("--k-dim", type=int, default=100, help="Krylov basis size of the arnoldi leg (nekStab default k_dim = 100)")
    ap.add_argument("--no-parity", action="store_true", help="N > 1: skip the multi-GPU vs single-GPU self-check")
    ap.add_argument("--cpu-nz", type=int, default=0, help="z-layers of the CPU sample (0: 2 for cpu_baseline, 4 for --impl reference)")
    ap.add_argument("--cpu-seconds", type=float, default=0.0, help="time budget of the CPU timed call (0: 45 s / 200 s)")
    args = ap.parse_args()
    K, W = args.steps, max(args.warmup, 0)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        if rank != 0:
            return
        nz = args.cpu_nz or (3 if args.small else 4)
        r = cpu_reference(K, W, nz, args.cpu_seconds or 200.0, precond=args.precond, tol=args.tol)
        cfg = workload_config(args.gpus, args.precond, args.tol, args.small)
        cfg.update({"sample_points": r["points"], "sample_layers": nz, "same_mesh_as_gpu_arm": bool(args.small and nz == 3),
                    "residual_difference": None if args.small else
                    f"CPU arm: the same 2-D mesh, base flow, seed, dt, tolerances and steps with {nz} of the {10 * args.gpus} periodic z-layers",
                    "pres_iters_per_step": r["pres_iters_per_step"], "helm_iters_per_comp_per_step": r["helm_iters_per_comp_per_step"]})
        line = {"impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": K, "warmup": W,
                "steps_done": r["steps_done"], "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # NCCL writes its debug output (the "NCCL version ..." banner at NCCL_DEBUG >= VERSION) to stdout: keep stdout for the one
    # JSON line and send NCCL's output to stderr
    # (NCCL honours NCCL_DEBUG_FILE only above the VERSION level, so a bare VERSION setting is dropped instead)
    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        del os.environ["NCCL_DEBUG"]
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    from nekstab_b200 import lib
    pl = Plumbing()
    torch = pl.torch

    # ---- N > 1: multi-GPU result == single-GPU result on the same small global mesh (before the big contexts exist)
    par_err, plane = None, None
    if world > 1 and not args.no_parity:
        par_err, plane = parity_n(pl, args)

    # ---- strong scaling (N > 1): the named fixed-size mesh split over the ranks
    strong = None
    if world > 1 and args.scaling in ("strong", "both") and not args.small:
        cs, ng = build_workload(10, world, rank)
        ctx, dt, _ = open_context(pl, cs, args)
        seed_slot0(ctx, cs, 4)
        tm = timed_matvec(pl, ctx, dt, K, W)
        prof = sampled_kernels(ctx, dt, min(K, 2))
        slow = sorted(((k, ms / cnt) for k, (ms, cnt) in prof.items() if cnt > 0 and k != "advab"), key=lambda t: -t[1])[:4]
        strong = {"value": ng * K / (tm["dev_ms"] * 1e-3), "unit": UNIT, "ms_per_step": tm["dev_ms"] / K, "elements_total": 19960,
                  "elements_this_rank": int(cs.nel), "pres_iters_per_step": tm["pres_iters"] / K,
                  "helm_iters_per_comp_per_step": tm["helm_iters"] / K / 3, "gpu_launches": tm["launches"],
                  "slowest_kernels_ms": {k: v for k, v in slow}, "workload": "cyl3d_1996x10_lx8 split over the ranks"}
        ctx.close()
        pl.barrier()

    # ---- headline: weak scaling
    t_setup = time.time()
    nz = 3 if args.small else 10 * world
    case, n_glob = build_workload(nz, world, rank)
    ctx, dt, nsteps_full = open_context(pl, case, args)
    k_dim = min(args.k_dim, 20) if args.small else args.k_dim
    do_arn = args.arnoldi > 0
    seed_slot0(ctx, case, (k_dim + 2) if do_arn else 4)
    if plane is None and world > 1:
        plane = "p2p" if ctx.get_field("p2p")[0] > 0 else "nccl"
    t_setup = time.time() - t_setup
    sampler = ClockSampler(pl.local_rank)
    tm = timed_matvec(pl, ctx, dt, K, W, sampler)
    value = n_glob * K / (tm["dev_ms"] * 1e-3)
    prof = sampled_kernels(ctx, dt, min(K, 3))

    # ---- end to end: host buffers in, host buffers out, through the public C-ABI calls (K steps, like the timed call)
    ctx.set_timestep(dt, K)
    v_h, p_h = ctx.vec_download(1)
    vin = torch.from_numpy(v_h).pin_memory().numpy()
    pin = torch.from_numpy(p_h).pin_memory().numpy()
    vout = torch.empty(v_h.shape, dtype=torch.float64).pin_memory().numpy()      # pinned result buffers owned by the caller
    pout = torch.empty(p_h.shape, dtype=torch.float64).pin_memory().numpy()
    pl.barrier()
    t0 = time.perf_counter()
    ctx.vec_upload(1, vin, pin)
    ctx.matvec(lib.DIRECT, 1, 2)
    ctx.vec_download(2, out=(vout, pout))
    pl.barrier()
    e2e_s = pl.maxr(time.perf_counter() - t0)
    e2e_value = n_glob * K / e2e_s
    vec_bytes = (vin.size + pin.size) * 8

    fp64_tf = ctx.fp64_peak()                 # FP64 FMA throughput measured live (register-resident DFMA loop, csrc/vec_kernels.cu)
    arn = None
    if do_arn:
        arn = arnoldi_leg(pl, ctx, case, n_glob, nsteps_full, dt, args.arnoldi, k_dim)
    ctx.close()

    if rank == 0:
        roof = roofline_block(prof, tm, K, case.n, args.precond, world)
        if roof and "advab" in roof["kernels"]:
            adv = roof["kernels"]["advab"]
            adv["flops_per_point"] = 2.7e3
            adv["TFLOPs"] = 2.7e3 * case.n / (adv["avg_ms"] * 1e-3) / 1e12
            adv["fp64_fma_peak_TFLOPs_measured"] = fp64_tf
            adv["frac_of_measured_fp64_peak"] = adv["TFLOPs"] / fp64_tf if fp64_tf else None
        cfg = workload_config(world, args.precond, args.tol, args.small)
        cfg.update({"dt": dt, "nsteps_per_matvec_T1": nsteps_full, "residual_projection_mxprev": args.mxprev,
                    "pres_iters_per_step": tm["pres_iters"] / K, "helm_iters_per_comp_per_step": tm["helm_iters"] / K / 3,
                    "l2": "per-iteration working set (2 GB) >> L2 (126 MB): no flush needed", "parallelism": f"elements/{world}",
                    "timing": "sampling profiler off, CUDA-graph replay " + ("on" if os.environ.get("NSB_GRAPHS", "1") != "0" and (world == 1 or plane == "p2p")
                                                                             else "off" + (" (NCCL data plane)" if world > 1 and plane != "p2p" else "")),
                    "setup_s": t_setup, "wall_s_timed": tm["wall_s"]})
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": tm["dev_ms"] / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": cfg,
                "clocks": sampler.summary(), "gpu_launches": tm["launches"],
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": vec_bytes / K, "d2h_bytes_per_step": vec_bytes / K,
                        "note": "one matvec call = K steps; vector copied in/out once per call"},
                "roofline": roof}
        if world > 1:
            line["data_plane"] = plane
            line["parity_n"] = par_err
            line["strong"] = strong
        if arn:
            line["arnoldi"] = arn
        if world == 1 and not args.no_cpu_baseline:
            try:
                nzc = args.cpu_nz or (3 if args.small else 2)
                r = cpu_reference(K, W, nzc, args.cpu_seconds or 45.0, precond=args.precond, tol=args.tol)
                line["cpu_baseline"] = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": "port", "sample": r["sample"],
                                        "steps_done": r["steps_done"], "pres_iters_per_step": r["pres_iters_per_step"],
                                        "helm_iters_per_comp_per_step": r["helm_iters_per_comp_per_step"],
                                        "gpu_pres_iters_per_step": tm["pres_iters"] / K, "gpu_helm_iters_per_comp_per_step": tm["helm_iters"] / K / 3}
            except Exception as exc:      # the GPU line must be printed whatever happens to the CPU leg
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": f"failed: {exc}"}
        print(json.dumps(line))
    if pl.dist:
        pl.dist.destroy_process_group()


if __name__ == "__main__":
    main()
