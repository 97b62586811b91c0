"""torchrun --nproc-per-node 2: the erase driver with projections sharded over ranks and ONE NCCL all-gather must reproduce the
single-GPU result: every rank ends up with all 32 edited projections, equal to the single-GPU ones to fp32 rounding (a rank's
shard is a different list of projections, so the row-block plan and with it the summation order of the tensor-core kernels differ:
not bit-equal) and equal bit for bit ACROSS ranks (the gathered bytes are the same everywhere)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch, torch.distributed as dist
from oracle.fake_pipe import FakePipe, layer_table
from uce_b200.erase import UCE

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
pipe = FakePipe(layer_table("sd14"), seed=1, correlated=True)
edit = [f"artist {i}" for i in range(20)]; guide = ["art"] * 20; pres = [f"thing {i}" for i in range(30)]
single = UCE(pipe, edit, guide, pres, 1.0, 1.0, 0.5, None, "x", device=f"cuda:{local}", verbose=False)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
sharded = UCE(pipe, edit, guide, pres, 1.0, 1.0, 0.5, None, "x", device=f"cuda:{local}", verbose=False)
os.environ["UCE_SHARD_CHUNKS"] = "3"          # the chunked form large edits take: a chunk's all-gather overlaps the next chunk's kernels
chunked = UCE(pipe, edit, guide, pres, 1.0, 1.0, 0.5, None, "x", device=f"cuda:{local}", verbose=False)
del os.environ["UCE_SHARD_CHUNKS"]
def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())
worst = max(max(rel(sharded[k].cpu(), single[k].cpu()), rel(chunked[k].cpu(), single[k].cpu())) for k in single)
ok = worst <= 3e-6 and len(sharded) == 32 and len(chunked) == 32
chk = torch.stack([sharded[k].double().sum() for k in sorted(sharded)] + [chunked[k].double().sum() for k in sorted(chunked)])   # the same bytes on every rank?
lo, hi = chk.clone(), chk.clone()
dist.all_reduce(lo, op=dist.ReduceOp.MIN); dist.all_reduce(hi, op=dist.ReduceOp.MAX)
ok = ok and bool(torch.equal(lo, hi))
t = torch.tensor([int(ok)], device=f"cuda:{local}")
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("sharded == single on every rank:", bool(t.item()), "keys", len(sharded), "worst rel-Frobenius difference", worst)
dist.destroy_process_group()
sys.exit(0 if t.item() else 1)
