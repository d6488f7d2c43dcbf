"""Debug helper (not a test): runs one stage-kernel call at the c5 chunk shape; with P3TOK_TC_TRACE=1 the library
prints the leader CTA's per-tile timeline.  python tests/_stage_trace.py [ngroups]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "adapting-2d-vits-for-3d-point-cloud-understanding_b200"))
import torch
from p3tok import synth, fold, ops, _lib

ng = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
W = int(sys.argv[2]) if len(sys.argv) > 2 else 128
mlp = fold.fold_p3embed_stage(synth.to_torch_state(synth.p3embed_state(3, 0.25, 4, 4, W, 77)), 0).to(torch.device("cuda:0"), torch.bfloat16)
rows = torch.randn(ng * 32, 6, device="cuda") * 0.5
for _ in range(2):
    tok = ops.patch_embed(_lib.ROWS_DIRECT, rows, None, None, None, None, ng, 32, mlp.tensors(), mlp.meta(), True)
torch.cuda.synchronize()
if not os.environ.get("P3TOK_TC_TRACE"):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        tok = ops.patch_embed(_lib.ROWS_DIRECT, rows, None, None, None, None, ng, 32, mlp.tensors(), mlp.meta(), True)
    e1.record(); torch.cuda.synchronize()
    print(f"patch_embed ng={ng} W={W}: {e0.elapsed_time(e1) / 20 * 1000:.1f} us per call")
