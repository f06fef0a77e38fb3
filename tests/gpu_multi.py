"""Row-sharded run on N GPUs (one process per GPU, NCCL), checked against the single-GPU run.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29512 \
        tests/gpu_multi.py [log2n]

Checks (exit status 0 iff all hold):
  * every rank sees bitwise identical scalars (step, f, phi'(0), trials) at every iteration -- the
    rank-ordered combination makes all ranks take identical branch decisions;
  * the gathered shards of the first directions and of the minimiser EQUAL the 1-GPU run of the same global
    problem BIT FOR BIT, and so do all per-iteration scalars and the iteration count: the reductions are
    partition-independent (include/flgpu_reduce.cuh) and these shards hold the same power-of-two number of chunks;
    (the AugmentedLagrangian case at the end uses 512-row shards -- half a chunk each -- and is compared within
    tolerance: ragged shards give a deterministic, rank-identical sum, only not the single-GPU bits).
Not a pytest file; tests/test_gpu.py::test_row_sharded_nccl launches it when >= 2 GPUs are visible."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import ctypes as C          # noqa: E402

import numpy as np          # noqa: E402
import torch                # noqa: E402
import torch.distributed as dist   # noqa: E402

import fortran_library_b200 as fl  # noqa: E402


def main():
    log2n = int(sys.argv[1]) if len(sys.argv) > 1 else 20
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def bcast(data):
        t = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(data), dtype=torch.uint8))
        dist.broadcast(t, 0)
        return bytes(t.cpu().numpy().tobytes())
    comm = fl.comm_create(rank, world, bcast)
    os.environ["FLGPU_EXCHANGE"] = "nccl"           # second communicator: the ncclAllGather fallback
    comm_nccl = fl.comm_create(rank, world, bcast)
    del os.environ["FLGPU_EXCHANGE"]
    fl.lib().flgpu_comm_uses_peer_memory.argtypes = [C.c_void_p]
    p2p = int(fl.lib().flgpu_comm_uses_peer_memory(comm))
    if rank == 0:
        print(f"[{world} ranks] peer-memory exchange: {'yes' if p2p else 'NO (fallback to ncclAllGather)'}", flush=True)
    n = 1 << log2n
    lo = (n * rank // world) // 2 * 2
    hi = n if rank == world - 1 else (n * (rank + 1) // world) // 2 * 2
    ok = True
    # (algorithm, objective, start, seed, options); runs that stop at MaxIteration mid-descent or crawl to
    # x* = 0 are chaotic in their tails, so the minimiser is compared where the run converges to an isolated
    # minimiser (Rosenbrock, x* = 1) and scaled by |x0| otherwise; exactness is carried by the other checks.
    cases = [("lbfgs", fl.OBJ_ROSENBROCK, fl.START_ROSEN_PERT, 7, dict(Memory=10)),
             ("lbfgs", fl.OBJ_DIAGQUAD, fl.START_ZERO, 0, dict(Memory=30, MaxIteration=40)),
             ("lbfgs", fl.OBJ_ROSENBROCK, fl.START_ROSEN_PERT, 7, dict(Memory=5, MaxIteration=40, fused=False)),
             ("cg", fl.OBJ_QUARTIC, fl.START_QUARTIC_U, 12345, dict(Method="DY")),
             ("cg", fl.OBJ_QUARTIC, fl.START_QUARTIC_U, 12345, dict(Method="PR"))]
    # the optional FLGPU_LS_FAST policy on row shards (not a reference routine; same exchange, same bitwise bar)
    cases += [("lbfgs", fl.OBJ_ROSENBROCK, fl.START_ROSEN_PERT, 7, dict(Memory=10, line_search="fast", MaxIteration=60)),
              ("cg", fl.OBJ_QUARTIC, fl.START_QUARTIC_U, 12345, dict(Method="DY", line_search="fast"))]
    for algo, kind, start, seed, kw in cases:
        run = fl.LBFGS if algo == "lbfgs" else fl.ConjugateGradient
        prob = fl.builtin_problem(kind)
        x = fl.DeviceVector.start(start, hi - lo, seed=seed, offset=lo, n_global=n)
        ob = fl.Observer(keep_vectors=True, max_vec_iters=10)
        # device-resident search forced on where the case allows it (auto mode uses it up to 2^18 rows per GPU only)
        ds_on = p2p and kw.get("fused", True) and kw.get("line_search") != "fast"
        st = run(prob, x, observer=ob, Warning=False, comm=comm, offset=lo, n_global=n,
                 device_search=True if ds_on else None, **kw)
        # identical scalars on every rank
        mine = torch.tensor([v for r in ob.rows for v in (r[1], r[2], r[3], float(r[4]))] + [float(st.iterations)],
                            dtype=torch.float64, device="cuda")
        count = torch.tensor([mine.numel()], device="cuda")
        cmax = count.clone()
        dist.all_reduce(cmax, op=dist.ReduceOp.MAX)
        same = int(cmax.item()) == int(count.item())
        if same:
            ref = mine.clone()
            dist.broadcast(ref, 0)
            same = bool(torch.equal(ref.view(torch.int64), mine.view(torch.int64)))
        flag = torch.tensor([int(same)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        same = bool(flag.item())
        # gather shards on rank 0
        def gather(v):
            t = torch.from_numpy(np.ascontiguousarray(v)).cuda()
            outs = []
            for r in range(world):
                size = (n if r == world - 1 else (n * (r + 1) // world) // 2 * 2) - (n * r // world) // 2 * 2
                buf = t if r == rank else torch.empty(size, dtype=torch.float64, device="cuda")
                dist.broadcast(buf, r)
                outs.append(buf.clone())
            return torch.cat(outs).cpu().numpy()
        # the two exchange implementations sum in the same (rank) order: identical bits
        x2 = fl.DeviceVector.start(start, hi - lo, seed=seed, offset=lo, n_global=n)
        st2 = run(prob, x2, Warning=False, comm=comm_nccl, offset=lo, n_global=n, **kw)
        modes_equal = bool(np.array_equal(x.numpy(), x2.numpy())) and st2.iterations == st.iterations
        x2.free()
        # the first run used the device-resident search (exchanges inside the cooperative kernel); a
        # host-driven search over the same peer-memory exchange must give the same bits again
        x3 = fl.DeviceVector.start(start, hi - lo, seed=seed, offset=lo, n_global=n)
        st3 = run(prob, x3, Warning=False, comm=comm, offset=lo, n_global=n, device_search=False, **kw)
        modes_equal = modes_equal and bool(np.array_equal(x.numpy(), x3.numpy())) and st3.iterations == st.iterations \
            and (not ds_on or st3.host_syncs > st.host_syncs)
        x3.free()
        flag = torch.tensor([int(modes_equal)], device="cuda")
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        modes_equal = bool(flag.item())
        xg = gather(x.numpy())
        pg = [gather(p) for p in ob.p[:6]]
        if rank == 0:
            x1 = fl.DeviceVector.start(start, n, seed=seed)
            ob1 = fl.Observer(keep_vectors=True, max_vec_iters=10)
            st1 = run(prob, x1, observer=ob1, Warning=False, **kw)
            xs = x1.numpy()
            x0n = np.linalg.norm(fl.DeviceVector.start(start, n, seed=seed).numpy())
            dx = np.linalg.norm(xg - xs) / max(x0n, np.linalg.norm(xs))
            dp = max(np.linalg.norm(a - b) / np.linalg.norm(b) for a, b in zip(pg, ob1.p[:6]))
            bitwise = bool(np.array_equal(xg, xs) and all(np.array_equal(a, b) for a, b in zip(pg, ob1.p[:6]))
                           and ob.rows == ob1.rows and st.iterations == st1.iterations and st.status == st1.status)
            good = same and modes_equal and bitwise
            print(f"[{world} ranks] {algo} kind={kind} {kw}: iterations {st.iterations}/{st1.iterations} "
                  f"ranks_identical={same} device-search==host-driven==nccl:{modes_equal} bitwise==1-GPU:{bitwise} "
                  f"|dx|={dx:.2e} max|dp|(first 6)={dp:.2e} -> {'OK' if good else 'FAIL'}", flush=True)
            ok = ok and good
            x1.free()
        x.free()
    # a USER objective written with include/flgpu_objective.cuh (tests/link/user_objective.cu, compiled by the caller):
    # its device-resident search trades the partial sums inside the kernel (flgpu_comm_search_exchange) and must give
    # the bits of the host-driven search on the same shards and of the single-GPU run
    if len(sys.argv) > 2 and p2p:
        ulib = C.CDLL(sys.argv[2])
        nu = 1 << 16
        lo_u, hi_u = nu * rank // world, nu * (rank + 1) // world
        x0u = np.random.default_rng(99).uniform(-0.5, 1.5, nu)
        for which in (0, 1):
            uprob = fl.capi.Problem()
            ulib.user_problem(which, C.byref(uprob))
            res = []
            for ds in (True, False):
                xu = fl.DeviceVector.from_numpy(x0u[lo_u:hi_u])
                obu = fl.Observer()
                stu = fl.LBFGS(uprob, xu, Memory=6, Warning=False, MaxIteration=60, observer=obu, comm=comm, offset=lo_u,
                               n_global=nu, device_search=ds)
                res.append((xu.numpy(), obu.rows, stu.iterations, stu.n_trials, stu.host_syncs))
                xu.free()
            same_modes = bool(np.array_equal(res[0][0], res[1][0])) and res[0][1:4] == res[1][1:4] and res[0][4] < res[1][4]
            t = torch.from_numpy(res[0][0]).cuda()
            parts = []
            for r in range(world):
                buf = t if r == rank else torch.empty(nu * (r + 1) // world - nu * r // world, dtype=torch.float64, device="cuda")
                dist.broadcast(buf, r)
                parts.append(buf.clone())
            xg = torch.cat(parts).cpu().numpy()
            flag = torch.tensor([int(same_modes)], device="cuda")
            dist.all_reduce(flag, op=dist.ReduceOp.MIN)
            same_modes = bool(flag.item())
            if rank == 0:
                x1 = fl.DeviceVector.from_numpy(x0u)
                ob1 = fl.Observer()
                st1 = fl.LBFGS(uprob, x1, Memory=6, Warning=False, MaxIteration=60, observer=ob1, device_search=False)
                bitwise = bool(np.array_equal(xg, x1.numpy())) and ob1.rows == res[0][1] and st1.iterations == res[0][2]
                good = same_modes and bitwise
                print(f"[{world} ranks] user functor {which}: in-kernel exchange == host-driven: {same_modes}, "
                      f"bitwise == 1 GPU: {bitwise}, round trips {res[0][4]} vs {res[1][4]} -> {'OK' if good else 'FAIL'}", flush=True)
                ok = ok and good
                x1.free()
    # the two-loop operator (flgpu_history_*) on row shards: K1's dots are rank-summed before K2
    nh = 1 << 14
    lo_h, hi_h = nh * rank // world, nh * (rank + 1) // world
    rng = np.random.default_rng(5)
    d = np.exp(rng.uniform(0, 1, nh))
    L = fl.lib()
    L.flgpu_history_create.restype = C.c_void_p
    L.flgpu_history_create.argtypes = [C.c_int64, C.c_int, C.c_void_p, C.c_void_p]
    L.flgpu_history_push.argtypes = [C.c_void_p] * 5
    L.flgpu_history_direction.argtypes = [C.c_void_p] * 7
    L.flgpu_history_destroy.argtypes = [C.c_void_p]
    hs = L.flgpu_history_create(hi_h - lo_h, 6, None, comm)
    h1 = L.flgpu_history_create(nh, 6, None, None) if rank == 0 else None
    x0 = rng.standard_normal(nh)
    worst = 0.0
    for _ in range(9):
        x1 = x0 + 0.1 * rng.standard_normal(nh)
        sh = [fl.DeviceVector.from_numpy(v[lo_h:hi_h]) for v in (x1, x0, d * x1, d * x0)]
        L.flgpu_history_push(hs, *[v.ptr for v in sh])
        ps, xs = fl.DeviceVector(hi_h - lo_h), fl.DeviceVector(hi_h - lo_h)
        gp, pp = C.c_double(), C.c_double()
        L.flgpu_history_direction(hs, sh[2].ptr, sh[0].ptr, ps.ptr, xs.ptr, C.addressof(gp), C.addressof(pp))
        t = torch.from_numpy(ps.numpy()).cuda()
        parts = []
        for r in range(world):
            buf = t if r == rank else torch.empty(nh * (r + 1) // world - nh * r // world, dtype=torch.float64, device="cuda")
            dist.broadcast(buf, r)
            parts.append(buf.clone())
        pg = torch.cat(parts).cpu().numpy()
        if rank == 0:
            fu = [fl.DeviceVector.from_numpy(v) for v in (x1, x0, d * x1, d * x0)]
            L.flgpu_history_push(h1, *[v.ptr for v in fu])
            p1, xt1 = fl.DeviceVector(nh), fl.DeviceVector(nh)
            gp1, pp1 = C.c_double(), C.c_double()
            L.flgpu_history_direction(h1, fu[2].ptr, fu[0].ptr, p1.ptr, xt1.ptr, C.addressof(gp1), C.addressof(pp1))
            worst = max(worst, np.linalg.norm(pg - p1.numpy()) / np.linalg.norm(p1.numpy()),
                        abs(gp.value - gp1.value) / abs(gp1.value), abs(pp.value - pp1.value) / abs(pp1.value))
        x0 = x1
    L.flgpu_history_destroy(hs)
    if rank == 0:
        L.flgpu_history_destroy(h1)
        good = worst == 0.0            # 2^14 rows over <= 8 ranks: whole chunks per rank -> the single-GPU bits
        print(f"[{world} ranks] two-loop operator on shards vs 1 GPU: worst relative difference {worst:.2e} -> "
              f"{'OK' if good else 'FAIL'}", flush=True)
        ok = ok and good
    # AugmentedLagrangian: the constraint values are exchanged inside the composed callbacks
    na = 4096
    lo_a, hi_a = na * rank // world, na * (rank + 1) // world
    prob, con = fl.builtin_problem(fl.OBJ_QUARTIC), fl.builtin_constraints()
    for solver in ("LBFGS", "ConjugateGradient"):
        xa = fl.DeviceVector.start(fl.START_QUARTIC_U, hi_a - lo_a, seed=12345, offset=lo_a, n_global=na)
        sa = fl.AugmentedLagrangian(prob, con, xa, UnconstrainedSolver=solver, Warning=False, MaxIteration=60,
                                    Precision=1e-6, comm=comm, offset=lo_a, n_global=na)
        t = torch.from_numpy(xa.numpy()).cuda()
        parts = [torch.empty(na * (r + 1) // world - na * r // world, dtype=torch.float64, device="cuda") for r in range(world)]
        for r in range(world):
            buf = t if r == rank else parts[r]
            dist.broadcast(buf, r)
            parts[r] = buf.clone()
        xg = torch.cat(parts).cpu().numpy()
        if rank == 0:
            x1 = fl.DeviceVector.start(fl.START_QUARTIC_U, na, seed=12345)
            s1 = fl.AugmentedLagrangian(prob, con, x1, UnconstrainedSolver=solver, Warning=False, MaxIteration=60,
                                        Precision=1e-6)
            dx = np.linalg.norm(xg - x1.numpy()) / np.linalg.norm(x1.numpy())
            good = (sa.status == 0 and s1.status == 0 and abs(np.linalg.norm(xg) - 1.0) < 1e-6
                    and abs(sa.outer_iterations - s1.outer_iterations) <= 1 and dx < 1e-6 * na)
            print(f"[{world} ranks] AugmentedLagrangian/{solver}: outer {sa.outer_iterations}/{s1.outer_iterations} "
                  f"| |x|-1 = {abs(np.linalg.norm(xg) - 1.0):.1e} |dx| = {dx:.1e} -> {'OK' if good else 'FAIL'}", flush=True)
            ok = ok and good
            x1.free()
        xa.free()
    fl.lib().flgpu_comm_destroy(comm)
    fl.lib().flgpu_comm_destroy(comm_nccl)
    okt = torch.tensor([int(ok)], device="cuda")
    dist.broadcast(okt, 0)
    dist.destroy_process_group()
    sys.exit(0 if okt.item() else 1)


if __name__ == "__main__":
    main()
