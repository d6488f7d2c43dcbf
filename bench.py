#!/usr/bin/env python
"""bench.py - tokenized clouds/sec of the point-patch tokenizer (FPS + kNN + gather/normalise + embed).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl p3tok|reference] [--workload c2|c2v|c1|c3|c4|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one pass of the hot path over one batch of synthetic clouds.  Default workload = BASELINE.json
configs[1] ("c2"): APF tokenizer PointNet(E=384, G=128, k=32, in_channel=6), B=128 clouds of N=2048 points
per GPU (batch sharding, weak scaling: every rank tokenizes its own B clouds, no collective on the path).

Printed JSON line (rank 0): `value` = whole-job clouds/s with inputs resident in HBM (device-timed, CUDA
events, max over ranks); `e2e` = the same metric through the public module call with HOST buffers (pinned
H2D copy of the clouds and D2H read of the tokens inside the timed region); `roofline` = the dominant
kernel family (patch embedding) against the measured bf16 peak of MEASURED_PEAKS.json; `cpu_baseline` =
the reference tokenizer's torch-CPU port (oracle/port.py) timed on this box's host cores.
`--impl reference` times that port alone (the reference is pure Python and /root/reference does not travel
to the GPU box) and prints the same line with "impl": "reference".
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "adapting-2d-vits-for-3d-point-cloud-understanding_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "tokenized clouds/sec (FPS+kNN+embed)"
UNIT = "clouds/s"

# name -> (family, B per GPU, N, centres G, k, embed E, description)
WORKLOADS = {
    "c1": dict(family="p4p", B=32, N=1024, k=32, embed_dim=256, sample_ratio=1 / 16,
               desc="Pix4Point P3Embed 2-stage 1024->256->64, k=32 (BASELINE configs[0])"),
    "c2": dict(family="apf", B=128, N=2048, G=128, k=32, E=384, C=3,
               desc="APF PointNet tokenizer E=384 G=128 k=32 (BASELINE configs[1])"),
    "c2v": dict(family="apf", vit=True, B=128, N=2048, G=128, k=32, E=384, C=3,
                desc="APF tokenizer + ViT-S forward: PointNet E=384 G=128 k=32 -> 12 APFViTLayers (12 heads) -> encoder_norm -> "
                     "token max (BASELINE configs[1] with its ViT tail, SURVEY 8f next #3)"),
    "c3": dict(family="p4p", B=256, N=8192, k=32, embed_dim=256, sample_ratio=1 / 16,
               desc="Pix4Point P3Embed 2-stage 8192->2048->512, k=32 (BASELINE configs[2])"),
    "c4": dict(family="apf", B=16, N=65536, G=2048, k=64, E=384, C=3,
               desc="large-scene APF Group(2048,64)+Encoder(384) (BASELINE configs[3])"),
    "c5": dict(family="p4p", B=512, N=1024, k=32, embed_dim=256, sample_ratio=1 / 16,
               desc="batch-sharded sweep, C1 shapes, 512 clouds per GPU (BASELINE configs[4])"),
}


def algorithmic_work(w):
    """Per-cloud algorithmic work: minimal embed FLOPs, FPS/kNN pairs, compulsory bytes."""
    k = w["k"]
    if w["family"] == "apf":
        E, G, N, C = w["E"], w["G"], w["N"], w["C"]
        per_pt = 2 * C * 256 + 256 * 512 + 512 * E + 4 * E * E      # minimal: global half applied once/group
        per_grp = 2 * E * E
        macs = G * k * per_pt + G * per_grp
        as_written = G * k * (2 * C * 256 + 256 * 512 + 512 * E + 6 * E * E)
        pairs = G * N
        byts = 4 * C * N + 4 * G * E
        if w.get("vit"):      # 12 x (qkv + proj + fc1 + fc2 + adapter down/up) per token + the two attention products per head
            vit_macs = 12 * (G * (3 * E * E + E * E + 8 * E * E + 2 * E * 64) + 2 * G * G * E)
            return dict(embed_flops=2 * macs, embed_flops_as_written=2 * as_written, fps_pairs=pairs, knn_pairs=pairs,
                        compulsory_bytes=4 * C * N + 4 * E, vit_flops=2 * vit_macs)
        return dict(embed_flops=2 * macs, embed_flops_as_written=2 * as_written, fps_pairs=pairs, knn_pairs=pairs,
                    compulsory_bytes=byts)
    n, cin, wd = w["N"], 6, int(w["embed_dim"] // 2)
    macs = aw = pairs = 0
    byts = 12 * n
    for _ in range(2):
        g = n // 4
        macs += g * k * (cin * wd + 4 * wd * wd) + g * 2 * wd * wd
        aw += g * k * (cin * wd + 7 * wd * wd)
        pairs += g * n
        n, cin, wd = g, wd + 3, wd * 2
    byts += 4 * n * (wd // 2)
    return dict(embed_flops=2 * macs, embed_flops_as_written=2 * aw, fps_pairs=pairs, knn_pairs=pairs,
                compulsory_bytes=byts)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_tflops=d["bf16_tflops"],
                    bf16_tflops_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


# ------------------------------------------------------------------------------------------ inputs / models
def make_inputs(w, seed):
    from p3tok import synth
    x = synth.make_cloud("clustered" if w.get("clustered") else "uniform", w["B"], w["N"], seed, w.get("C", 3))
    if w["family"] == "apf":
        starts = [synth.start_indices(w["B"], w["N"], seed)]
    else:
        starts = [synth.start_indices(w["B"], w["N"], seed, 0), synth.start_indices(w["B"], w["N"] // 4, seed, 1)]
    return x, starts


def make_vit_state(w):
    from p3tok import synth
    return synth.apf_vit_state(w["E"], 12, 15, 0)


def make_state(w):
    from p3tok import synth
    if w["family"] == "apf":
        return synth.apf_encoder_state(w["E"], 2 * w["C"], 0)
    return synth.p3embed_state(3, w["sample_ratio"], 4, 4, w["embed_dim"], 0)


def build_gpu_model(w, precision, device):
    from p3tok import synth
    from p3tok.modules import P3Embed, PointNet
    sd = synth.to_torch_state(make_state(w))
    if w.get("vit"):
        from p3tok.apf_model import AdaptPointFormer
        net = AdaptPointFormer(num_classes=15, embedding_dim=w["E"], npoint=w["G"], nsample=w["k"], in_channels=w["C"],
                               precision=precision).eval().to(device)
        net.point_encoder.encoder.load_state_dict(sd, strict=True)
        net.load_state_dict(synth.to_torch_state(make_vit_state(w)), strict=False)
        return net, (lambda x, st: net.features(x, st[0]))
    if w["family"] == "apf":
        net = PointNet(w["E"], w["G"], w["k"], 2 * w["C"], precision=precision).eval().to(device)
        net.encoder.load_state_dict(sd, strict=True)
        return net, (lambda x, st: net(x, st[0]))
    net = P3Embed(sample_ratio=w["sample_ratio"], k=w["k"], embed_dim=w["embed_dim"], precision=precision).eval().to(device)
    net.load_state_dict(sd, strict=True)

    def run(x, st):
        ps, fs = net(x, x.transpose(1, 2), st)
        return fs[-1]
    return net, run


def cpu_port_runner(w):
    """The reference's tokenizer on CPU (oracle/port.py), all host threads."""
    from oracle import port
    from p3tok import synth
    sd = synth.to_torch_state(make_state(w))
    if w.get("vit"):
        vsd = synth.to_torch_state(make_vit_state(w))
        return lambda x, st: port.apf_vit_features(vsd, port.apf_pointnet(sd, x, w["G"], w["k"], st[0]), 12, 12)
    if w["family"] == "apf":
        return lambda x, st: port.apf_pointnet(sd, x, w["G"], w["k"], st[0])
    return lambda x, st: port.p3embed(sd, x, x.transpose(1, 2).contiguous(), w["k"], 2, st)[1][-1]


def time_cpu_port(w, clouds, repeats, seed=4321):
    torch.set_num_threads(os.cpu_count() or 1)
    ww = dict(w, B=clouds)
    x, st = make_inputs(ww, seed)
    xt, stt = torch.from_numpy(x), [torch.from_numpy(s) for s in st]
    run = cpu_port_runner(ww)
    with torch.no_grad():
        run(xt, stt)                                   # warm-up
        best = float("inf")
        for _ in range(repeats):
            t0 = time.perf_counter()
            run(xt, stt)
            best = min(best, time.perf_counter() - t0)
    return clouds / best, best


# ------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self.stop_flag, self.th = index, [], threading.Event(), None

    def _loop(self):
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [v.strip() for v in out.strip().split(",")]
                if len(f) >= 6:
                    self.samples.append(f)
            except Exception:
                pass
            self.stop_flag.wait(0.05)

    def start(self):
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag.set()
        if self.th:
            self.th.join(timeout=6)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm = [float(s[0]) for s in self.samples]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": float(self.samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------ arms
def run_reference(args, w, rank, world):
    """--impl reference: the reference's own CPU implementation of the path (torch-CPU port), rank 0 only."""
    if rank != 0:
        return
    sample = min(w["B"], {"c2": 16, "c2v": 16, "c1": 16, "c5": 16, "c3": 2, "c4": 1}[args.workload])
    torch.set_num_threads(os.cpu_count() or 1)
    ww = dict(w, B=sample)
    x, st = make_inputs(ww, 4321)
    xt, stt = torch.from_numpy(x), [torch.from_numpy(s) for s in st]
    run = cpu_port_runner(ww)
    with torch.no_grad():
        for _ in range(max(args.warmup, 1)):
            run(xt, stt)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            run(xt, stt)
        dt = time.perf_counter() - t0
    v = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": max(args.warmup, 1), "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['desc']}", "clouds_per_step": sample,
                   "note": "CPU port of the reference tokenizer (oracle/port.py), same torch calls in the same order"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sample} clouds per step x {args.steps} steps of the {args.workload} workload"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_p3tok(args, w, rank, world, local_rank):
    import torch.distributed as dist
    from p3tok import ops
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    precision = args.precision
    net, run = build_gpu_model(w, precision, device)
    x_np, st_np = make_inputs(w, 1234 + rank)
    x_host = torch.from_numpy(x_np).pin_memory()
    st_host = [torch.from_numpy(s).pin_memory() for s in st_np]

    # rotating pool of distinct input batches, total footprint > 2x L2 (126 MB), so every timed step reads
    # its clouds from HBM ("inputs larger than L2")
    in_bytes = x_host.numel() * 4
    pool_n = max(2, min(256, int(2.2 * 126e6 / in_bytes) + 1))
    base = x_host.to(device)
    pool = [torch.roll(base, shifts=i, dims=1).contiguous() for i in range(pool_n)]
    st_pool = [[(s.to(device) + i) % (w["N"] if j == 0 else w["N"] // 4) for j, s in enumerate(st_host)]
               for i in range(pool_n)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for i in range(max(args.warmup, 3)):
            out = run(pool[i % pool_n], st_pool[i % pool_n])
        barrier()
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        from p3tok.graph import GraphedTokenizer
        launches0 = ops.kernel_launches()
        out = run(pool[0], st_pool[0])
        per_step_launches = ops.kernel_launches() - launches0
        # Device-resident loop: the step is the module call captured once into a CUDA graph (p3tok.graph, the serving
        # wrapper); every step first copies its clouds and start indices from the HBM-resident pool into the graph's
        # input buffers (device-to-device, inside the timed region), then replays.  --eager times the Python-dispatched
        # module call instead (launch-bound for the small workloads).
        gdev = None if args.eager else GraphedTokenizer(lambda x, *st: run(x, list(st)), [base] + st_pool[0])
        if gdev is not None:
            for i in range(3):
                out = gdev(pool[i % pool_n], *st_pool[i % pool_n])
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            if gdev is not None:
                out = gdev(pool[i % pool_n], *st_pool[i % pool_n])
            else:
                out = run(pool[i % pool_n], st_pool[i % pool_n])
        e1.record()
        barrier()
        dev_ms = e0.elapsed_time(e1)
        launches = per_step_launches * args.steps       # kernels of libp3tok.so executed in the timed region

        # ---- e2e: host buffers in, host tokens out, through the public serving API (p3tok.graph.GraphedHostTokenizer):
        # one CUDA graph per step = pinned-host -> device copies of the clouds and start indices, the captured module call,
        # device -> pinned-host copy of the tokens.  Two instances on two streams are replayed alternately, so the copies
        # of step i+1 overlap the kernels of step i; the host pays one graph launch per step.
        from p3tok.graph import GraphedHostTokenizer
        graphs = [GraphedHostTokenizer(lambda x, *st: run(x, list(st)), [x_host] + st_host, device) for _ in range(2)]
        out_host = [g.host_output for g in graphs]
        for i in range(4):
            graphs[i & 1].replay()
        for g in graphs:
            g.synchronize()
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            graphs[i & 1].replay()
        for g in graphs:
            g.synchronize()
        barrier()
        e2e_ms = 1e3 * (time.perf_counter() - t0)
        # the host-to-host graph returns the same tokens as the eager module call on the same clouds
        assert torch.equal(out_host[0].to(device), run(x_host.to(device), [s.to(device) for s in st_host])), "e2e graph != eager"
        # the graph path returns the same tokens as the eager path
        assert torch.equal(gdev(pool[0], *st_pool[0]), run(pool[0], st_pool[0])) if gdev is not None else True, "graph replay != eager"
        clocks = sampler.stop() if rank == 0 else None

        # ---- per-stage device times (separate pass, not part of the timed region)
        sink = []
        ops.set_profile(sink)
        for i in range(4):
            run(pool[i % pool_n], st_pool[i % pool_n])
        torch.cuda.synchronize()
        ops.set_profile(None)
        stage_ms = {}
        for name, a, b in sink:
            stage_ms[name] = stage_ms.get(name, 0.0) + a.elapsed_time(b) / 4

    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = float(t[0]), float(t[1])

    allgather_ms = None
    if world > 1:   # optional epilogue of BASELINE config 5: gather every rank's tokens over NVLink (not in `value`)
        gathered = torch.empty((world,) + tuple(out.shape), dtype=out.dtype, device=device)
        dist.all_gather_into_tensor(gathered, out.contiguous())
        barrier()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        g0.record()
        dist.all_gather_into_tensor(gathered, out.contiguous())
        g1.record()
        barrier()
        allgather_ms = g0.elapsed_time(g1)

    if rank != 0:
        return
    B = w["B"]
    clouds = B * world * args.steps
    value = clouds / (dev_ms / 1e3)
    work = algorithmic_work(w)
    pk = peaks()
    embed_ms = stage_ms.get("embed", 0.0)
    peak_tf = pk["bf16_tflops_sustained"]
    ach_tf = (work["embed_flops"] * B / (embed_ms / 1e3) / 1e12) if embed_ms > 0 else None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16" if precision == "bf16" else "f32", "data": "synthetic",
        "config": {"workload": f"{args.workload}: {w['desc']}", "clouds_per_gpu_per_step": B,
                   "global_clouds_per_step": B * world, "points": w["N"], "k": w["k"],
                   "parallelism": f"batch-shard x{world}, no collective on the path",
                   "l2": f"rotating pool of {pool_n} distinct input batches ({pool_n * in_bytes / 1e6:.0f} MB > L2)",
                   "embed_precision": precision, "step": "eager module call" if args.eager else "CUDA-graph replay of the module call"},
        "e2e": {"value": clouds / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": in_bytes + sum(s.numel() * 8 for s in st_host),
                "d2h_bytes_per_step": out_host[0].numel() * out_host[0].element_size(), "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": int(launches),
        "clocks": clocks,
        "stage_ms_per_step": {k: round(v, 4) for k, v in sorted(stage_ms.items())},
        "roofline": {"bound": "tensor", "kernel": "patch embedding (p3tok_patch_embed)",
                     "achieved": ach_tf, "peak": peak_tf, "unit": "TFLOP/s",
                     "frac": (ach_tf / peak_tf) if ach_tf else None,
                     # dram__bytes_read.sum + dram__bytes_write.sum of the embed launches of one step (first layer 216 MB,
                     # fused 256->512->384 pair 631 MB, group-bias GEMM 14 MB, fused 384->768->384 pair 483 MB), from the
                     # ncu --set full capture summarised in profiles/r01_c2_ncu_full.txt (c2 / bf16 only; the
                     # layer-by-layer path of the first session moved 3.72 GB)
                     "traffic": 1.344e9 if (args.workload == "c2" and precision == "bf16") else None,
                     "peak_source": f"{pk['source']} bf16 sustained (MEASURED_PEAKS.json)",
                     "algorithmic_flops_per_launch_group": work["embed_flops"] * B,
                     "as_written_flops": work["embed_flops_as_written"] * B, "duration_ms": embed_ms},
        "roofline_hbm": {
            "fps_streaming_equiv_GBps": (16 * work["fps_pairs"] * B / (stage_ms["fps"] / 1e3) / 1e9) if stage_ms.get("fps") else None,
            "knn_pairs_per_s": (work["knn_pairs"] * B / (stage_ms["knn"] / 1e3)) if stage_ms.get("knn") else None,
            "compulsory_bytes_per_step": work["compulsory_bytes"] * B, "peak_GBps": pk["hbm_gbs"]},
    }
    if w.get("vit") and stage_ms.get("apf_vit"):
        line["vit"] = {"ms_per_step": stage_ms["apf_vit"], "algorithmic_flops": work["vit_flops"] * B,
                       "achieved_tflops": work["vit_flops"] * B / (stage_ms["apf_vit"] / 1e3) / 1e12,
                       "frac_of_bf16_peak": work["vit_flops"] * B / (stage_ms["apf_vit"] / 1e3) / 1e12 / peak_tf}
    if allgather_ms is not None:
        line["token_allgather_ms"] = allgather_ms
    if world == 1 and not args.no_cpu_baseline:
        sample = {"c2": 32, "c2v": 32, "c1": 32, "c5": 32, "c3": 4, "c4": 1}[args.workload]
        v, secs = time_cpu_port(w, sample, 2)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                                "sample": f"{sample} clouds of the {args.workload} workload, best of 2 after 1 warm-up ({secs:.2f} s per pass)"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="p3tok", choices=["p3tok", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default=os.environ.get("P3TOK_BENCH_PRECISION", "bf16"), choices=["fp32", "bf16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--eager", action="store_true", help="time the Python-dispatched module call instead of the CUDA-graph replay")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, w, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    elif args.gpus > 1:
        print(json.dumps({"error": "launch with torch.distributed.run for --gpus > 1"}))
        sys.exit(2)
    try:
        run_p3tok(args, w, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
