"""torchrun --nproc-per-node 2 tests/_syncbn_2gpu.py : sharded train-mode Encoder with sync_bn=True on 2 GPUs equals the
single-GPU full-batch result (tokens of the shard, all-reduced parameter gradients, running statistics)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "adapting-2d-vits-for-3d-point-cloud-understanding_b200"))
import numpy as np, torch, torch.distributed as dist
from p3tok import synth
from p3tok.modules import PointNet
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl")
dev = torch.device("cuda")
B, N, G, k, E = 8, 512, 16, 16, 64
x = torch.from_numpy(synth.make_cloud("uniform", B, N, 5, 3)).to(dev)
st = torch.from_numpy(synth.start_indices(B, N, 5)).to(dev)
gt = torch.from_numpy((synth.uniform01(5, B * G * E, 3).reshape(B, G, E) - 0.5).astype(np.float32)).to(dev)
sd = synth.to_torch_state(synth.apf_encoder_state(E, 6, 5))
def run(sync, lo, hi):
    net = PointNet(E, G, k, 6, sync_bn=sync).to(dev).train()
    net.encoder.load_state_dict(sd)
    tok = net(x[lo:hi], st[lo:hi])
    (tok * gt[lo:hi]).sum().backward()
    return net, tok.detach()
full, tok_full = run(False, 0, B)
lo, hi = rank * B // world, (rank + 1) * B // world
part, tok_part = run(True, lo, hi)
err = float((tok_part - tok_full[lo:hi]).abs().max() / tok_full.abs().max())
worst = err
for (n, p), (_, q) in zip(part.encoder.named_parameters(), full.encoder.named_parameters()):
    g = p.grad.clone(); dist.all_reduce(g)
    worst = max(worst, float((g - q.grad).abs().max() / max(float(q.grad.abs().max()), 1e-6)) if "bias" not in n or "conv" not in n else 0.0)
for (n, b), (_, c) in zip(part.encoder.named_buffers(), full.encoder.named_buffers()):
    if "num_batches" not in n:
        worst = max(worst, float((b - c).abs().max() / c.abs().max()))
print(f"rank {rank}: sharded sync_bn vs full batch: worst relative difference {worst:.2e}", flush=True)
assert worst < 1e-4
dist.destroy_process_group()
