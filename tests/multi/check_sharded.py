"""torchrun worker (N GPUs): the landmark-sharded solver against the unsharded oracle.  Run through
`gpurun --gpus 2 -- python -m torch.distributed.run --nproc-per-node 2 ... tests/multi/check_sharded.py`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch, torch.distributed as dist
from dynosam_b200 import synth
from dynosam_b200.binding import Solver
import bench
from oracle import oracle as O

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

def allreduce(dev, n, stream):
    class _A:
        __cuda_array_interface__ = {"shape": (n,), "typestr": "<f8", "data": (dev, False), "version": 3, "strides": None}
    t = torch.as_tensor(_A(), device=f"cuda:{local}")
    with torch.cuda.stream(torch.cuda.ExternalStream(stream)):
        dist.all_reduce(t)

for name, kw in (("C1", {}), ("C3", dict(scale=0.03))):
    p = synth.make_config(name, **kw)
    bw = bench.problem_bandwidth(p)
    sh = bench.shard_problem(p, rank, world)
    s = Solver(sh, device=local); s.set_shard(rank, world, allreduce, bw)
    o = O.OracleProblem(p)
    e = s.error(); eo = o.error()
    assert abs(e - eo) <= 1e-9*eo, (e, eo)
    lam = 1e-4
    d = s.solve(lam)
    rc, do = o.schur_solve(lam)
    # compare the pose part (replicated) and this rank's landmarks
    npose = p.n_pose
    keep = (np.arange(p.n_point) % world) == rank
    dl = do[6*npose:].reshape(-1, 3)[keep].reshape(-1)
    err_p = np.linalg.norm(d[:6*npose] - do[:6*npose])/np.linalg.norm(do[:6*npose])
    err_l = np.linalg.norm(d[6*npose:] - dl)/np.linalg.norm(dl)
    assert err_p < 1e-6 and err_l < 1e-6, (err_p, err_l)
    st = s.optimize(max_iterations=8); so = o.optimize(max_iterations=8)
    assert st["iterations"] == so["iterations"] and st["inner_iterations"] == so["inner_iterations"], (st, so)
    assert abs(st["error_final"] - so["error_final"]) <= 1e-6*so["error_final"], (st["error_final"], so["error_final"])
    if rank == 0:
        print(f"{name}: sharded x{world} ok  chi2 {st['error_final']:.9f} vs oracle {so['error_final']:.9f}  step err {err_p:.2e}/{err_l:.2e}", flush=True)
    s.close()
dist.destroy_process_group()
