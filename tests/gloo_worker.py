"""World-size-2 (or more) CPU worker: exercises the library's multi-rank gather-scatter PLAN (host C++ code of gs.cu)
with gloo as the transport and a numpy emulation of the pack / exchange / segmented-sum kernels, and checks the result
against the single-domain direct-stiffness sum.  Also checks the partition and that both ranks derive identical,
mirror-ordered interface lists."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def get_plan(L, lib, c, rank, world, dist, torch):
    glo = np.ascontiguousarray(c.glo, dtype=np.int64).ravel()
    cnt = C.c_longlong()
    L.nsb_gs_host_candidates(c.ldim, c.lx1, c.nel, lib._p(glo), None, C.byref(cnt))
    ids = np.zeros(max(cnt.value, 1), dtype=np.int64)
    L.nsb_gs_host_candidates(c.ldim, c.lx1, c.nel, lib._p(glo), lib._p(ids), C.byref(cnt))
    ids = ids[:cnt.value]
    gathered = [None] * world
    dist.all_gather_object(gathered, ids)
    counts = np.array([len(g) for g in gathered], dtype=np.int64)
    allids = np.ascontiguousarray(np.concatenate(gathered), dtype=np.int64)
    sizes = np.zeros(8, dtype=np.int32)
    assert L.nsb_gs_host_plan(rank, world, lib._p(counts), lib._p(allids), lib._p(sizes)) == 0
    lens = {0: sizes[0] + 1, 1: sizes[1], 2: sizes[2], 3: sizes[2] + 1, 4: sizes[3], 5: sizes[3], 6: sizes[3],
            7: sizes[0] + 1, 8: sizes[4], 9: sizes[4], 10: sizes[0]}
    plan = {}
    names = ["seg_off", "seg_idx", "nbr_rank", "nbr_off", "send_seg", "send_base", "send_cnt", "rseg_off", "rseg_pos", "rseg_cnt", "nbefore"]
    for w, nm in enumerate(names):
        a = np.zeros(max(int(lens[w]), 1), dtype=np.int32)
        assert L.nsb_gs_host_get(w, lib._p(a)) == 0
        plan[nm] = a[:int(lens[w])]
    plan["nshared"] = int(sizes[3])
    return plan


def emulate_dssum(plan, u, nf, rank, dist, torch):
    """numpy restatement of k_gs_pack -> grouped send/recv -> k_gs_sum<HALO> (csrc/gs.cu), same summation order."""
    n = u.shape[1]
    so, si = plan["seg_off"], plan["seg_idx"]
    nseg = len(so) - 1
    loc = np.zeros((nf, nseg))
    for s in range(nseg):
        for j in range(so[s], so[s + 1]):
            loc[:, s] += u[:, si[j]]
    send = np.zeros(3 * max(plan["nshared"], 1))
    for e in range(plan["nshared"]):
        for f in range(nf):
            send[plan["send_base"][e] + f * plan["send_cnt"][e]] = loc[f, plan["send_seg"][e]]
    recv = np.zeros_like(send)
    reqs = []
    for i, r in enumerate(plan["nbr_rank"]):
        a, b = plan["nbr_off"][i], plan["nbr_off"][i + 1]
        sbuf = torch.from_numpy(send[3 * a:3 * a + nf * (b - a)].copy())
        rbuf = torch.zeros(nf * (b - a), dtype=torch.float64)
        reqs.append((dist.isend(sbuf, int(r)), dist.irecv(rbuf, int(r)), rbuf, a, b))
    for s_req, r_req, rbuf, a, b in reqs:
        s_req.wait(); r_req.wait()
        recv[3 * a:3 * a + nf * (b - a)] = rbuf.numpy()
    out = u.copy()
    ro, rp, rc, nb = plan["rseg_off"], plan["rseg_pos"], plan["rseg_cnt"], plan["nbefore"]
    for s in range(nseg):
        acc = np.zeros(nf)
        for j in range(ro[s], ro[s] + nb[s]):
            acc += [recv[rp[j] + f * rc[j]] for f in range(nf)]
        acc = acc + loc[:, s] if nb[s] > 0 else loc[:, s].copy()
        for j in range(ro[s] + nb[s], ro[s + 1]):
            acc += [recv[rp[j] + f * rc[j]] for f in range(nf)]
        for j in range(so[s], so[s + 1]):
            out[:, si[j]] = acc
    return out


def main():
    import torch
    import torch.distributed as dist
    from nekstab_b200 import cases, lib
    from util import small_cases
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    L = lib.load_library()
    bad = []
    for name in ("box2d_n6_outflow", "box3d_n4_outflow", "box2d_n4_periodic"):
        gc = small_cases()[name]
        part = cases.partition(gc.key, world, gc.d2)
        sel = np.nonzero(part == rank)[0]
        c = gc.local_part(rank, world)
        assert c.nel == sel.size and c.nelg == gc.nel
        plan = get_plan(L, lib, c, rank, world, dist, torch)
        # interface lists mirror each other: what I send to r has the length r sends to me
        mine = {int(r): int(plan["nbr_off"][i + 1] - plan["nbr_off"][i]) for i, r in enumerate(plan["nbr_rank"])}
        allm = [None] * world
        dist.all_gather_object(allm, mine)
        for r, cnt in mine.items():
            if allm[r].get(rank) != cnt:
                bad.append((name, "asymmetric interface", rank, r))
        nf = gc.ldim
        rng = np.random.default_rng(11)
        ug = rng.standard_normal((nf, gc.n))
        g = gc.glo.ravel()
        ref = np.stack([np.bincount(g, weights=ug[f], minlength=g.max() + 1)[g] for f in range(nf)])
        ul = ug.reshape(nf, gc.nel, -1)[:, sel].reshape(nf, -1)
        got = emulate_dssum(plan, ul, nf, rank, dist, torch)
        refl = ref.reshape(nf, gc.nel, -1)[:, sel].reshape(nf, -1)
        err = np.abs(got - refl).max()
        if not err < 1e-12:
            bad.append((name, "dssum", err))
        # every rank holds bit-identical values on shared nodes: gather interface values and compare
        gl = gc.glo[sel].ravel()
        mine_vals = {int(k): got[0, i] for i, k in enumerate(gl)}
        allv = [None] * world
        dist.all_gather_object(allv, mine_vals)
        for r in range(world):
            if r == rank:
                continue
            for k, v in allv[r].items():
                if k in mine_vals and mine_vals[k] != v:
                    bad.append((name, "not bit-identical across ranks", k))
                    break
        # ---- the same plan code on the element-VERTEX mesh (lx1 = 2): the gather-scatter of the pressure preconditioner's Q1
        #      level across GPUs (csrc/pmg.cu builds it with gs_build(.., N1 = 2, ..) from the corner ids)
        from types import SimpleNamespace
        N = gc.lx1 - 1
        G = gc.glo.reshape((gc.nel,) + (gc.lx1,) * gc.ldim)
        nk = 2 ** gc.ldim
        vg = np.stack([G[(slice(None),) + tuple(N * ((k >> (gc.ldim - 1 - a)) & 1) for a in range(gc.ldim))] for k in range(nk)], axis=1)
        vc = SimpleNamespace(ldim=gc.ldim, lx1=2, nel=sel.size, glo=np.ascontiguousarray(vg[sel]))
        vplan = get_plan(L, lib, vc, rank, world, dist, torch)
        rc_g = rng.standard_normal((1, gc.nel * nk))                       # one value per (element, corner) entry
        vref = np.bincount(vg.ravel(), weights=rc_g[0], minlength=vg.max() + 1)[vg.ravel()]
        rc_l = rc_g.reshape(1, gc.nel, nk)[:, sel].reshape(1, -1)
        vgot = emulate_dssum(vplan, rc_l, 1, rank, dist, torch)
        verr = np.abs(vgot[0] - vref.reshape(gc.nel, nk)[sel].ravel()).max()
        if not verr < 1e-12:
            bad.append((name, "vertex dssum", verr))
    t = torch.tensor([len(bad)])
    dist.all_reduce(t)
    if bad:
        print(f"[rank {rank}] {bad}", flush=True)
    if rank == 0:
        print("GLOO_OK" if int(t.item()) == 0 else "GLOO_FAIL", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(t.item()) == 0 else 1)


if __name__ == "__main__":
    main()
