"""Multi-GPU parity worker (run under torchrun, one rank per GPU): device partition maps, halo'd SpMV and
sharded solves are compared bit-for-bit with the oracle run with the same number of shards."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def empty_shard_cases(kb, o, parallel, ctx, rank, world):
    """the chunk partition leaves ranks without rows whenever ceil(n/p)*(p-1) >= n (n = 1, or
    n = p+1 for p >= 3).  An empty rank must still launch the reducing kernels and join every in-kernel all-reduce."""
    for nn in (1, world + 1):
        # 1-D Laplacian tridiag(-1, 2, -1)
        rp, ci, v = [0], [], []
        for i in range(nn):
            for j, a in ((i - 1, -1.0), (i, 2.0), (i + 1, -1.0)):
                if 0 <= j < nn:
                    ci.append(j); v.append(a)
            rp.append(len(ci))
        rp = np.array(rp, dtype=np.uint64); ci = np.array(ci, dtype=np.uint64); v = np.array(v)
        Ao = o.OCsr(nn, nn, rp, ci, v)
        lo, hi = kb.partition_range(nn, world, rank)
        base = int(rp[lo])
        A = kb.DeviceCsr.from_csr_shard(nn, lo, hi, rp[lo:hi + 1] - np.uint64(base), ci[base:int(rp[hi])], v[base:int(rp[hi])], ctx)
        bg = o.spmv(Ao, np.ones(nn))
        b = bg[lo:hi].copy()
        x = np.zeros(hi - lo)
        st = kb.PcgSolver(1e-10, 100).solve(A, kb.Jacobi().setup(A), b, x)
        rc, xo, so, _ = o.pcg(Ao, o.OPc.jacobi(Ao), bg, np.zeros(nn), 1e-10, 100, nshards=world)
        assert rc == 0 and (st.iterations, st.converged) == (so.iterations, bool(so.converged)), ("empty-shard pcg", nn, st.iterations, so.iterations)
        assert st.final_residual == so.final_residual and np.array_equal(x, xo[lo:hi]), "empty-shard pcg"
        x = np.zeros(hi - lo)
        st = kb.GmresSolver(5, 1e-10, 100).solve(A, kb.Ilu0().setup(A), b, x)
        rc, xo, so = o.gmres(Ao, o.OPc.ilu0(Ao, nblocks=world), bg, np.zeros(nn), 5, 1e-10, 100, mode=1, variant=o.GMRES_CGS2, nshards=world)
        assert (st.iterations, st.converged) == (so.iterations, bool(so.converged)), ("empty-shard gmres", nn, st.iterations, so.iterations)
        assert st.final_residual == so.final_residual and np.array_equal(x, xo[lo:hi]), "empty-shard gmres"
        x = np.zeros(hi - lo)
        st = kb.BiCgStabSolver(1e-10, 100, textbook=True).solve(A, kb.Jacobi().setup(A), b, x)
        rc, xo, so = o.bicgstab(Ao, o.OPc.jacobi(Ao), bg, np.zeros(nn), 1e-10, 100, variant=o.BICG_TEXTBOOK, nshards=world)
        assert (st.iterations, st.converged) == (so.iterations, bool(so.converged)), ("empty-shard bicgstab", nn)
        assert np.array_equal(x, xo[lo:hi]), "empty-shard bicgstab"
        assert ctx.dot(b, b) == o.dot(bg, bg, nshards=world)
        A.close()


def main():
    import torch
    import torch.distributed as dist
    import kryst_b200 as kb
    from kryst_b200 import parallel
    import oracle_ffi as o

    rank, world, local = parallel.dist_env()
    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
    ctx = kb.Context(local)
    parallel.init_comm(ctx)
    assert ctx.rank() == rank and ctx.size() == world
    assert ctx.all_reduce(float(rank + 1)) == float(sum(range(1, world + 1)))

    # rest of the Comm surface (parallel/mod.rs:9-22; mpi_comm.rs:74-109; core/wrappers.rs:134-156)
    g = np.arange(3 * world, dtype=np.float64) * 1.5
    out = np.zeros(3)
    ctx.scatter(g if rank == 1 % world else None, out, root=1 % world)
    assert np.array_equal(out, g[3 * rank:3 * rank + 3]), "comm scatter"
    got = ctx.gather(np.array([rank, 10 * rank], dtype=np.int64), root=0)
    if rank == 0:
        assert np.array_equal(got, np.array([v for r in range(world) for v in (r, 10 * r)], dtype=np.int64)), "comm gather"
    else:
        assert got.size == 0
    xs = np.random.default_rng(11).standard_normal(1000 * world)
    ys = np.random.default_rng(12).standard_normal(1000 * world)
    lo_, hi_ = kb.partition_range(xs.size, world, rank)
    assert ctx.comm_dot(xs[lo_:hi_], ys[lo_:hi_]) == o.dot(xs, ys, nshards=world), "comm dot"
    assert ctx.comm_norm(xs[lo_:hi_]) == float(np.sqrt(o.dot(xs, xs, nshards=world))), "distributed norm"
    empty_shard_cases(kb, o, parallel, ctx, rank, world)
    for kind, N in (("poisson3d", 12), ("convdiff3d", 10), ("convdiff2d", 20), ("varcoef27", 7)):
        n, lo, hi, rp, ci, v = parallel.shard_stencil(kind, N, world, rank)
        A = kb.DeviceCsr.from_csr_shard(n, lo, hi, rp, ci, v, ctx)
        Ao = o.stencil(kind, N)
        # partition maps: bit-exact
        assert np.array_equal(A.ghosts(), o.ghost_list(Ao, lo, hi)), "ghost list"
        gh, _, _ = parallel.host_ghost_plan(n, world, rank, rp, ci)
        assert np.array_equal(A.ghosts(), gh)
        # distributed SpMV == rows [lo,hi) of the global product, bit-exact
        xg = np.random.default_rng(5).standard_normal(n)
        y = np.zeros(hi - lo)
        A.matvec(xg[lo:hi].copy(), y)
        assert np.array_equal(y, o.spmv(Ao, xg)[lo:hi]), "dist spmv"
        bg = o.spmv(Ao, np.ones(n))
        b = bg[lo:hi].copy()
        # PCG + Jacobi
        # (all four stencils: on the nonsymmetric ones PCG either converges or fails with the same KError as the oracle)
        x = np.zeros(hi - lo)
        pcg = kb.PcgSolver(1e-8, 3000)
        rc, xo, so, _ = o.pcg(Ao, o.OPc.jacobi(Ao), bg, np.zeros(n), 1e-8, 3000, nshards=world)
        try:
            st = pcg.solve(A, kb.Jacobi().setup(A), b, x)
            dev_rc = 0
        except kb.IndefiniteMatrix:
            dev_rc, st = 3, pcg.last_stats
        except kb.IndefinitePreconditioner:
            dev_rc, st = 4, pcg.last_stats
        assert dev_rc == rc, ("dist pcg status", kind, dev_rc, rc)
        assert (st.iterations, st.converged) == (so.iterations, bool(so.converged)), (kind, st.iterations, so.iterations)
        assert st.final_residual == so.final_residual, "dist pcg residual"
        if rc == 0:
            assert np.array_equal(x, xo[lo:hi]), "dist pcg"
        else:
            assert not x.any(), "x must stay untouched on Err (pcg.rs:171,212)"
        # single-reduction PCG: one all-reduce of three sums per iteration.  On the nonsymmetric
        # operators the recurrence for p.Ap goes non-positive: device and oracle must raise the same IndefiniteMatrix.
        x = np.zeros(hi - lo)
        rc, xo, so, _ = o.pcg_sr(Ao, o.OPc.jacobi(Ao), bg, np.zeros(n), 1e-8, 3000, nshards=world)
        sr = kb.PcgSolver(1e-8, 3000).with_fused_reduction()
        try:
            st = sr.solve(A, kb.Jacobi().setup(A), b, x)
            assert rc == 0, ("sr: oracle failed, device did not", kind, rc)
        except kb.IndefiniteMatrix:
            assert rc == 3 and not x.any(), ("sr: device raised IndefiniteMatrix", kind, rc)
            st = sr.last_stats
        assert (st.iterations, st.converged) == (so.iterations, bool(so.converged)), ("sr", kind, st.iterations, so.iterations)
        assert st.final_residual == so.final_residual, "dist single-reduction pcg residual"
        if rc == 0:
            assert np.array_equal(x, xo[lo:hi]), "dist single-reduction pcg"
        # pipelined PCG: the iteration's one reduction is sent by the update kernel and received at the end of
        # the SpMV; same status / bits as the oracle's restatement
        x = np.zeros(hi - lo)
        rc, xo, so, _ = o.pcg_pipe(Ao, o.OPc.jacobi(Ao), bg, np.zeros(n), 1e-8, 3000, nshards=world)
        pp = kb.PcgSolver(1e-8, 3000).with_pipelined()
        try:
            st = pp.solve(A, kb.Jacobi().setup(A), b, x)
            assert rc == 0, ("pipelined: oracle failed, device did not", kind, rc)
        except kb.IndefiniteMatrix:
            assert rc == 3 and not x.any(), ("pipelined: device raised IndefiniteMatrix", kind, rc)
            st = pp.last_stats
        except kb.IndefinitePreconditioner:
            assert rc == 4 and not x.any(), ("pipelined: device raised IndefinitePreconditioner", kind, rc)
            st = pp.last_stats
        assert (st.iterations, st.converged) == (so.iterations, bool(so.converged)), ("pipelined", kind, st.iterations, so.iterations)
        assert st.final_residual == so.final_residual, "dist pipelined pcg residual"
        if rc == 0:
            assert np.array_equal(x, xo[lo:hi]), "dist pipelined pcg"
        # BiCGStab (textbook) + Jacobi
        x = np.zeros(hi - lo)
        st = kb.BiCgStabSolver(1e-8, 3000, textbook=True).solve(A, kb.Jacobi().setup(A), b, x)
        rc, xo, so = o.bicgstab(Ao, o.OPc.jacobi(Ao), bg, np.zeros(n), 1e-8, 3000, variant=o.BICG_TEXTBOOK, nshards=world)
        assert (st.iterations, st.converged, st.breakdown) == (so.iterations, bool(so.converged), so.breakdown), "dist bicgstab stats"
        assert st.final_residual == so.final_residual and np.array_equal(x, xo[lo:hi]), "dist bicgstab"
        # GMRES(10) + block-Jacobi ILU(0) (one block per GPU), left and none
        for mode, use_pc in ((1, True), (0, False), (2, True)):
            x = np.zeros(hi - lo)
            pc = kb.Ilu0().setup(A) if use_pc else None
            st = kb.GmresSolver(10, 1e-8, 3000).with_preconditioning(mode).solve(A, pc, b, x)
            rc, xo, so = o.gmres(Ao, o.OPc.ilu0(Ao, nblocks=world) if use_pc else None, bg, np.zeros(n), 10, 1e-8, 3000,
                                 mode=mode, variant=o.GMRES_CGS2, nshards=world)
            assert (st.iterations, st.converged) == (so.iterations, bool(so.converged)), ("dist gmres", kind, mode, st.iterations, so.iterations)
            assert st.final_residual == so.final_residual and np.array_equal(x, xo[lo:hi]), "dist gmres"
        # GMRES with block orthogonalisation: ONE all-reduce of j+2 sums per Arnoldi step
        x = np.zeros(hi - lo)
        st = kb.GmresSolver(10, 1e-8, 3000).with_block_orthogonalisation().solve(A, kb.Ilu0().setup(A), b, x)
        rc, xo, so = o.gmres(Ao, o.OPc.ilu0(Ao, nblocks=world), bg, np.zeros(n), 10, 1e-8, 3000, mode=1, variant=o.GMRES_BLOCK, nshards=world)
        assert (st.iterations, st.converged) == (so.iterations, bool(so.converged)), ("dist block-orth gmres", kind, st.iterations, so.iterations)
        assert st.final_residual == so.final_residual and np.array_equal(x, xo[lo:hi]), "dist block-orth gmres"
        # FGMRES (flexible, block-Jacobi ILU(0) as the fixed preconditioner)
        x = np.zeros(hi - lo)
        st = kb.FgmresSolver(1e-8, 3000, 10).solve_flex(A, kb.Ilu0().setup(A), b, x)
        rc, xo, so = o.fgmres(Ao, o.OPc.ilu0(Ao, nblocks=world), bg, np.zeros(n), 10, 1e-8, 3000, nshards=world)
        assert (st.iterations, st.converged) == (so.iterations, bool(so.converged)), ("dist fgmres", kind, st.iterations, so.iterations)
        assert st.final_residual == so.final_residual and np.array_equal(x, xo[lo:hi]), "dist fgmres"
        pc = None
        A.close()
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()
    print("DIST_WORKER_OK rank %d/%d" % (rank, world))


if __name__ == "__main__":
    main()
