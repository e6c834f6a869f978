"""Worker of tests/test_gpu_multi.py: launched by torch.distributed.run with one process per GPU.
Row-sharded pqrfact / idfact / psvdfact on a tall matrix against the single-GPU result of the SAME library with the
same fast-mode seed (Omega is keyed by the global row index, so only the summation order differs)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "lowrankapprox.jl_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import numpy as np
import torch
import torch.distributed as dist

import brapprox
import lra_oracle as o


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    ctx = brapprox.Context(local)
    brapprox.init_comm(ctx, rank, world, device=dev)

    m, n, r = 6000, 512, 90
    A = o.decaying_matrix(m, n, r, 12.0, r, seed=3)          # identical on every rank
    row0, ml = brapprox.row_shard(m, rank, world)
    Aloc = np.asfortranarray(A[row0:row0 + ml])
    ctx.set_row_shard(row0, m)

    F = brapprox.pqrfact(Aloc, rtol=1e-10, seed=7, ctx=ctx)
    assert ctx.collective_count() > 0
    S = brapprox.psvdfact(Aloc, rtol=1e-10, seed=7, ctx=ctx)

    # gather Q and U row blocks on rank 0
    def gather_rows(X):
        k = X.shape[1]
        chunk = brapprox.row_shard(m, 0, world)[1]
        buf = torch.zeros((chunk, k), dtype=torch.float64, device=dev)
        buf[:X.shape[0]] = torch.from_numpy(np.ascontiguousarray(X)).to(dev)
        parts = [torch.zeros_like(buf) for _ in range(world)]
        dist.all_gather(parts, buf)
        return torch.cat([parts[g][:brapprox.row_shard(m, g, world)[1]] for g in range(world)]).cpu().numpy()

    Q = gather_rows(F.Q)
    U = gather_rows(S.U)
    ks = torch.tensor([F.k, S.k_id, len(S.S)], device=dev)
    kall = [torch.zeros_like(ks) for _ in range(world)]
    dist.all_gather(kall, ks)
    assert all(bool((t == ks).all()) for t in kall), "ranks disagree on k"

    if rank == 0:
        ref = brapprox.Context(local)                         # no communicator: plain single-GPU run
        F1 = brapprox.pqrfact(A, rtol=1e-10, seed=7, ctx=ref)
        S1 = brapprox.psvdfact(A, rtol=1e-10, seed=7, ctx=ref)
        assert F.k == F1.k and np.array_equal(F.p, F1.p), (F.k, F1.k)
        nrm = np.linalg.norm(A, 2)
        assert np.max(np.abs(F.R - F1.R)) <= 1e-10 * nrm
        assert np.max(np.abs(Q @ F.R - F1.Q @ F1.R)) <= 1e-10 * nrm
        assert np.linalg.norm(Q.T @ Q - np.eye(F.k)) <= 1e-12
        P = np.zeros((n, n))
        P[F.p - 1, np.arange(n)] = 1.0
        err = np.linalg.norm(A @ P - Q @ F.R, 2) / nrm
        assert err <= 1e-8, err
        assert len(S.S) == len(S1.S)
        assert np.max(np.abs(S.S - S1.S)) <= 1e-10 * S1.S[0]
        errs = np.linalg.norm(A - (U * S.S) @ S.Vt, 2) / nrm
        assert errs <= 1e-8, errs
        print(f"dist ok: world={world} k={F.k} ksvd={len(S.S)} qr_err={err:.2e} svd_err={errs:.2e} "
              f"collectives={ctx.collective_count()}", flush=True)
        ref.close()
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
