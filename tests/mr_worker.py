"""Worker for the multi-GPU parity test: run under torchrun (one rank per GPU).  Every rank builds the same global
case, takes its share under Nek5000's partition rule, runs the CUDA path on its elements and compares with the
oracle evaluated on the GLOBAL mesh (restricted to the rank's elements)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from nekstab_b200 import cases, lib
    from oracle.stepper import LinearizedStepper
    from util import make_oracle, random_nodal, rel, small_cases, smooth_field
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("gloo")   # host plumbing only; the data path uses the library's own NCCL communicator
    ids = [lib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    failures = []
    for name in ("box3d_n8_outflow", "box2d_n6_outflow", "box3d_n6_dirichlet"):
        gc = small_cases()[name]
        part = cases.partition(gc.key, world, gc.d2)
        if np.bincount(part, minlength=world).min() == 0:       # this small mesh leaves a rank without elements at this world size
            continue
        s = make_oracle(gc)
        sel = np.nonzero(part == rank)[0]
        c = gc.local_part(rank, world)
        g = lib.NekStabB200(c, device=lr, rank=rank, nranks=world, nccl_id=ids[0])
        d = gc.ldim

        def loc(a):              # (.., nel, npts)-shaped global oracle array -> this rank's elements
            a = np.asarray(a)
            return a.reshape(a.shape[:-(d + 1)] + (gc.nel, -1))[..., sel, :] if a.ndim > d + 1 else a.reshape(gc.nel, -1)[sel]

        def check(tag, got, ref, tol):
            e = rel(got, ref)
            if not e < tol:
                failures.append((name, tag, e))

        rng = np.random.default_rng(3)
        u = rng.standard_normal(gc.n)
        check("dssum", g.op_dssum(u.reshape(gc.nel, -1)[sel]), loc(s.dssum(u.reshape(s.eshape))), 2e-12)
        check("binvm1", g.get_field("binvm1"), loc(s.binv), 2e-12)
        a, b = rng.standard_normal(gc.n), rng.standard_normal(gc.n)
        ref = s.glsc3(a.reshape(s.eshape), s.bm1, b.reshape(s.eshape))
        got = g.op_glsc3(a.reshape(gc.nel, -1)[sel], s.bm1.reshape(gc.nel, -1)[sel], b.reshape(gc.nel, -1)[sel])
        if abs(got - ref) > 1e-11 * max(abs(ref), np.sqrt(gc.n)):
            failures.append((name, "glsc3", abs(got - ref)))
        p = rng.standard_normal(gc.nel * s.lx2 ** d)
        check("cdabdtp", g.op_cdabdtp(p.reshape(gc.nel, -1)[sel]), s.cdabdtp(p.reshape(s.eshape2)).reshape(gc.nel, -1)[sel], 5e-12)
        # full linearised maps
        nsteps, dt = 4, 2.0e-3
        g.set_params(1.0 / gc.re, 1.0, 1e-13, 1e-13, 3000, 100000)
        g.set_timestep(dt, nsteps)
        g.vec_alloc(3)
        st = LinearizedStepper(s, gc.ubase, gc.re, gc.spng_fun, solver="direct", ifvcor=gc.ifvcor)
        v0 = smooth_field(gc, 21).reshape((d,) + s.eshape)
        p0 = 0.1 * np.random.default_rng(5).standard_normal(s.eshape2)
        g.vec_upload(0, v0.reshape(d, gc.nel, -1)[:, sel], p0.reshape(gc.nel, -1)[sel])
        for mode, adj in ((lib.DIRECT, False), (lib.ADJOINT, True)):
            g.matvec(mode, 0, 1)
            v, pp = g.vec_download(1)
            vo, po = st.linearized_map(v0, p0, nsteps, dt, adjoint=adj)
            check(f"matvec{mode}", v, vo.reshape(d, gc.nel, -1)[:, sel], 1e-10)
        # inner product is global
        g.vec_copy(2, 0)
        ip = g.inner_product(0, 2)
        bm1s = s.bm1 * (gc.spng_fun.reshape(s.eshape) == 0)
        refip = float(sum(np.sum(v0[k] ** 2 * bm1s) for k in range(d)))
        if abs(ip - refip) > 1e-11 * refip:
            failures.append((name, "inner_product", abs(ip - refip)))
        # ---- pressure preconditioner (csrc/pmg.cu) across ranks: vertex sums over the peer-memory halo channel, aggregate
        #      sums over the all-reduce; compared with the oracle on the GLOBAL mesh using the gathered aggregate map
        from oracle import pmg
        g.set_pressure_preconditioner(1, 2 * world)
        info = g.pc_get(3)
        mine = (sel, g.pc_get(0).astype(np.int64) + int(info[4]))
        allagg = [None] * world
        dist.all_gather_object(allagg, mine)
        agg = np.zeros(gc.nel, dtype=np.int64)
        for se, ag in allagg:
            agg[se] = ag
        M = pmg.PMG(s, agg=agg, ifvcor=bool(gc.ifvcor))
        r = np.random.default_rng(8).standard_normal(s.eshape2)
        if gc.ifvcor:
            r -= r.mean()
        check("pmg_apply", g.op_pc_apply(r.reshape(gc.nel, -1)[sel]), M.apply(r).reshape(gc.nel, -1)[sel], 1e-10)
        gg = -s.opdiv(smooth_field(gc, 11).reshape((d,) + s.eshape))
        g.set_params(1.0 / gc.re, 1.0, 1e-13, 1e-12, 2000, 50000)
        phi1, it1 = g.op_esolver(gg.reshape(gc.nel, -1)[sel])
        g.set_pressure_preconditioner(0)
        phi0, it0 = g.op_esolver(gg.reshape(gc.nel, -1)[sel])
        check("pmg_esolver", phi1, phi0, 1e-7)
        if not it1 < it0:
            failures.append((name, "pmg_iterations", (it1, it0)))
        g.set_pressure_preconditioner(1, 2 * world)
        g.set_params(1.0 / gc.re, 1.0, 1e-13, 1e-13, 3000, 100000)
        for mode, adj in ((lib.DIRECT, False), (lib.ADJOINT, True)):
            g.matvec(mode, 0, 1)
            v, pp = g.vec_download(1)
            vo, po = st.linearized_map(v0, p0, nsteps, dt, adjoint=adj)
            check(f"pmg_matvec{mode}", v, vo.reshape(d, gc.nel, -1)[:, sel], 1e-10)
        g.close()
    t = torch.tensor([len(failures)])
    dist.all_reduce(t)
    if failures:
        print(f"[rank {rank}] FAILURES: {failures}", flush=True)
    if rank == 0:
        print("MULTIRANK_OK" if int(t.item()) == 0 else "MULTIRANK_FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()
