#!/usr/bin/env python
"""Headline benchmark: WSI bags/s, forward+backward of the IMP prototype-fusion hot path.

Workload (BASELINE.json configs[1]): survival training on synthetic TCGA-shaped bags -- 16384 patches
x 512 features per slide, 32 prototypes, 6 genomic pathway groups, bf16 features/MMA operands, fp32
statistics.  A step = one fwd+bwd of the hot path over one batch of bags (path_net, two prototype
blocks, omic encoders, the training-time modularity term, all parameter gradients; for N > 1 GPUs
also the NCCL gradient all-reduce).  Weak scaling: every rank owns `--bags` slides per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Prints ONE JSON line (see the keys below).  `--impl reference` times the CPU restatement of the
reference arithmetic (oracle/imp_oracle.py; the reference itself is Python and cannot travel to the
GPU box) on the host cores for the same config.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_PATCH, D_IN, N_PROTO, N_PATHWAYS = 16384, 512, 32, 6
GROUP_SIZES = [82, 330, 513, 440, 1538, 451]
WORKLOAD = "configs[1]: survival training, synthetic TCGA-shaped bags 16384x512, 32 prototypes, 6 pathways, bf16"


def workload_label(patches, protos):
    """The BASELINE configs[1] label; a run at other sizes (parity / smoke sizes) says so."""
    if patches == N_PATCH and protos == N_PROTO:
        return WORKLOAD
    return "configs[1] shape at a reduced size (%dx512 patches, %d prototypes, 6 pathways, bf16): not the headline configuration" % (patches, protos)


def ncu_traffic(bags):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch, per kernel, from the tracked summary of the
    `ncu --set full` capture (profiles/ncu_traffic.json, written by profiles/summarize.py from the .ncu-rep of the
    same bench command).  The file names the capture it came from; it applies only to the bags-per-step it was
    captured at -- otherwise (or if absent) the traffic fields are null."""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
    except Exception:
        return {}, None
    if int(t.get("bags_per_step", -1)) != int(bags):
        return {}, None
    return t.get("bytes_per_launch", {}), t.get("source")


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--bags", type=int, default=32, help="slides per step per GPU")
    ap.add_argument("--e2e-bags", type=int, default=32, help="slides per step of the host-buffer (e2e) leg")
    ap.add_argument("--patches", type=int, default=N_PATCH)
    ap.add_argument("--protos", type=int, default=N_PROTO)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="report the eager (Python-launched) steps instead of CUDA-graph replays")
    ap.add_argument("--missing-omics", type=float, default=0.0,
                    help="configs[2]: fraction of slides whose genomics are missing (gene values imputed by the cohort means "
                         "in the encoder gather, umeml_gan.py:391-392); 0 = fully paired batches (configs[1])")
    ap.add_argument("--workload", default="fusion", choices=["fusion", "kmeans", "giant"],
                    help="fusion: the headline (configs[1]); kmeans: configs[4], 2^20 x 512 fp32 -> 32 centroids, one assignment pass "
                         "per step; giant: configs[3], ONE 120k-patch bag sharded by rows over the ranks (LSE-merged pooling, "
                         "two-phase modularity), fwd+bwd per step")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference arithmetic
# ------------------------------------------------------------------------------------------------
def cpu_port_step(n_patch, n_proto, seed=0):
    """One slide through the oracle: path_net -> 2 prototype blocks -> modularity (both token groups)
    -> omic encoders, forward + backward.  Returns seconds."""
    import torch
    from oracle import imp_oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util_hotpath import make_params
    g = torch.Generator().manual_seed(seed)
    params = make_params(0)
    x = torch.randn(n_patch, D_IN, generator=g)
    p_proto = (torch.rand(n_proto, 256, generator=g) * 2 - 1) / n_proto
    cot = torch.randn(1, n_proto, 256, generator=g)
    omic = torch.rand(1, sum(GROUP_SIZES), generator=g)
    groups, o = [], 0
    for s in GROUP_SIZES:
        groups.append(list(range(o, o + s))); o += s
    ws = [(torch.randn(256, s, generator=g) / s ** 0.5).requires_grad_(True) for s in GROUP_SIZES]
    bs = [torch.zeros(256, requires_grad=True) for _ in GROUP_SIZES]
    t0 = time.perf_counter()
    res = O.hot_path_step([x], params, p_proto, with_modularity=True, grad_seed=cot)
    h_omic = O.omic_encode(omic, groups, ws, bs)
    tok = torch.cat([torch.rand(1, 1, 256), h_omic], dim=1)[0]
    h = O.path_net(x, params["path_net.0.weight"], params["path_net.0.bias"])
    _, dtok = O.modularity(tok, h)
    (h_omic[0] * dtok[1:]).sum().backward()
    return time.perf_counter() - t0, float(res["modularity"])


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # bounded sample: one 16384-patch slide costs ~20 s of host work, so at most 4 timed steps of 1 slide each
    # (the GPU arm runs `--steps` steps of 32 slides); the metric (bags/s) is per slide and comparable
    k = max(1, min(args.steps, 4))
    w = min(args.warmup, 1)
    for _ in range(w):
        cpu_port_step(args.patches, args.protos)
    t = 0.0
    for i in range(k):
        dt, _ = cpu_port_step(args.patches, args.protos, seed=i)
        t += dt
    val = k / t
    sample = "%d steps of 1 slide %dx%d, P=%d, fwd+bwd incl. modularity (oracle port, fp32)" % (k, args.patches, D_IN, args.protos)
    print(json.dumps({
        "impl": "reference", "metric": "wsi_bags_per_s_fwd_bwd", "value": val, "unit": "bags/s",
        "n_gpus": args.gpus, "steps": k, "warmup": w, "ms_per_step": 1e3 * t / k, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_label(args.patches, args.protos).replace("bf16", "bf16 on the GPU arm; this CPU arm computes in f32"),
                   "bags_per_step": 1, "patches": args.patches, "prototypes": args.protos, "modularity": True,
                   "steps_note": "at most 4 timed steps of 1 slide (~20 s each) so that the arm ends within minutes"},
        "cpu_baseline": {"value": val, "unit": "bags/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "bags/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# clocks sampler
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def count(self, windows):
        return sum(1 for ts, _ in list(self.rows) if any(a <= ts <= b + 0.1 for a, b in windows))

    def stop(self, windows):
        """windows: [(t0, t1), ...] host-clock intervals during which the timed steps were running."""
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ts, line in self.rows:
            if not any(a <= ts <= b + 0.1 for a, b in windows):
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for nm, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}



# ------------------------------------------------------------------------------------------------
# the reference arithmetic, eager PyTorch on the GPU (the honest speed-up denominator)
# ------------------------------------------------------------------------------------------------
def gpu_eager_baseline(dev):
    """path_net -> 2 prototype blocks -> compute_modularity for both token groups, forward + backward, written with
    the ATen ops the reference launches (oracle.prototype_pool / modularity_literal restate umeml_gan.py:410,425-434
    and ops/utils.py:188-228 literally, N x N and P x N x N tensors included).  fp32, TF32 off, 1 slide per step.
    configs[0] (4096, P = 16) always; configs[1] (16384, P = 32 + 7) if it fits in HBM -- the OOM is reported."""
    import torch
    from oracle import imp_oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from util_hotpath import block_tensors, make_params
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    names = ["in_proj_weight", "in_proj_bias", "out_proj.weight", "out_proj.bias", "norm1.weight", "norm1.bias"]
    out = {"device": torch.cuda.get_device_name(dev), "dtype": "f32 (TF32 off)", "kind": "port of the reference op sequence (oracle/imp_oracle.py) in eager PyTorch on the GPU",
           "unit": "bags/s"}
    for label, n, p in (("configs[0] 4096x512 P=16", 4096, 16), ("configs[1] 16384x512 P=32+7", 16384, 32)):
        try:
            params = {k: v.to(dev).requires_grad_(True) for k, v in make_params(0).items()}
            blocks = [dict(zip(names, block_tensors(params, b))) for b in range(2)]
            g = torch.Generator(device=dev).manual_seed(0)
            x = torch.randn(n, D_IN, device=dev, generator=g)
            p_proto = (torch.rand(p, 256, device=dev, generator=g) * 2 - 1) / p
            tok = torch.rand(N_PATHWAYS + 1, 256, device=dev, generator=g).requires_grad_(True)
            cot = torch.randn(p, 256, device=dev, generator=g) * 1e-2

            def step():
                for t in list(params.values()) + [tok]:
                    t.grad = None
                c, h = O.prototype_pool(x, p_proto, params["path_net.0.weight"], params["path_net.0.bias"], blocks)
                loss = (c * cot).sum() + O.modularity_literal(c, h) + O.modularity_literal(tok, h)
                loss.backward()
            step()
            torch.cuda.synchronize(dev)
            k = 3
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(k):
                step()
            e1.record()
            torch.cuda.synchronize(dev)
            ms = e0.elapsed_time(e1) / k
            out[label] = {"value": 1e3 / ms, "ms_per_bag": ms, "peak_mem_gb": round(torch.cuda.max_memory_allocated(dev) / 1e9, 1)}
        except torch.cuda.OutOfMemoryError as exc:
            out[label] = {"value": None, "oom": str(exc)[:160]}
        finally:
            params = blocks = x = tok = None
            torch.cuda.empty_cache()
            torch.cuda.reset_peak_memory_stats(dev)
    return out


def drop_in_model_step(dev, x, cu, omic, B, N, P):
    """Whole-model training step through the reference-facing API (registry name, cfg, batch dict, 7-tuple)."""
    import torch
    from types import SimpleNamespace as NS
    from imp_b200 import survival
    from imp_b200.registry import build_model
    import imp_b200.umeml_gan  # noqa: F401
    cfg = NS(DATASET=NS(ROOT=".", PATH=NS(DIM=D_IN), OMIC=NS(DIM=sum(GROUP_SIZES))),
             MODEL=NS(DROPOUT=0.25, HIDDEN_DIM=256, PROJECT_DIM=256, FUSION="concat", SIZE="small",
                      UMEML=NS(PROTOTYPES=P, REGISTERS=3, GENE_GROUP_INDEXES=None, IMPORTANCE_LOG="defer")),
             TRAINER=NS(PREC="fp32"))
    old = os.getcwd()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    os.chdir(os.path.join(ROOT, "gpurun_out"))                       # the model appends <set>_path.txt / _omic.txt in the CWD
    try:
        model = build_model("umeml_gan", verbose=False, cfg=cfg, num_classes=4, omic_sizes=1000).to(dev).train()
        model.plot_set = "bench"
        batch = {"x_packed": x, "cu_seqlens": cu, "max_len": N, "omic": omic, "patient_id": [str(i) for i in range(B)]}
        y = torch.randint(0, 4, (B,), device=dev)
        c = torch.randint(0, 2, (B,), device=dev)
        params = [p for p in model.parameters()]

        def step():
            for p in params:
                p.grad = None
            out = model(batch)
            loss = survival.nll_loss_new(out, y, c) + out[5] + out[1]
            loss.backward()
            return loss
        def timeit(fn, k=5):
            for _ in range(3):
                fn()
            torch.cuda.synchronize(dev)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(k):
                out = fn()
            e1.record()
            torch.cuda.synchronize(dev)
            # only the VALUE leaves: a retained loss tensor keeps its autograd graph and with it AccumulateGrad nodes bound
            # to the stream of this eager run, which a later capture on another stream is not allowed to synchronise with
            return e0.elapsed_time(e1) / k, float(out)
        ms_eager, loss = timeit(step)
        rec = {"unit": "bags/s", "bags_per_step": B, "loss": loss, "eager": {"value": B * 1e3 / ms_eager, "ms_per_step": ms_eager},
               "what": "build_model('umeml_gan', cfg) -> model(batch) 7-tuple -> NLL + KD + modularity -> backward; token tail in "
                       "batched torch (fp32) with the Nystrom attention core in csrc/nystrom.cu; importance rows kept on the device "
                       "(IMPORTANCE_LOG='defer') and appended afterwards"}
        try:
            from imp_b200 import step as S

            def loss_fn():
                out = model(batch)
                return survival.nll_loss_new(out, y, c) + out[5] + out[1]
            gs = S.GraphedStep(None).capture_fn(loss_fn, params, dev)
            ms_graph, _ = timeit(gs.replay)
            gs.close()
            rec.update({"value": B * 1e3 / ms_graph, "ms_per_step": ms_graph, "launch_mode": "whole step replayed from one CUDA graph"})
        except Exception as exc:
            import traceback
            traceback.print_exc(file=sys.stderr)
            try:
                torch.cuda.synchronize(dev)
            except Exception:
                pass
            rec.update({"value": B * 1e3 / ms_eager, "ms_per_step": ms_eager,
                        "launch_mode": "eager (graph capture failed: %s: %s)" % (type(exc).__name__, str(exc)[:160])})
        model.flush_importance_logs()
        return rec
    finally:
        os.chdir(old)


# ------------------------------------------------------------------------------------------------
# multi-GPU correctness record (rank 0 recomputes everything alone and compares)
# ------------------------------------------------------------------------------------------------
def run_dp_check(dev, rank, world):
    """(1) slide-parallel: every rank runs fwd+bwd on its own two small bags, gradients are all-reduced (mean);
    rank 0 then runs ALL bags in one process and compares every parameter gradient.  (2) giant-bag mode on a
    4096-patch bag: sharded pooling (LSE merge) + sharded modularity against the single-GPU result."""
    import torch
    import torch.distributed as dist
    from imp_b200 import model as M, modularity as MOD, ops, parallel as PAR, step as S

    def rel(a, b):
        return float(((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30)).item())

    P = 16
    torch.manual_seed(4321)                                             # identical parameters on every rank
    net = M.IMPHotPath(n_proto=P, dropout=0.0, seed=0).to(dev)
    runner = S.HotPathStep(net, with_modularity=True).to(dev).train()

    def data(r):
        g = torch.Generator().manual_seed(900 + r)
        lens = [384 + 64 * (r % 3), 200 + 17 * r]
        x = torch.cat([torch.randn(n, D_IN, generator=g) for n in lens]).bfloat16()
        omic = torch.rand(2, sum(GROUP_SIZES), generator=g)
        cp = torch.randn(2, P, 256, generator=g) * 1e-2
        co = torch.randn(2, N_PATHWAYS + 1, 256, generator=g) * 1e-2
        return lens, x, omic, cp, co

    def grads_of(rs, scale):
        lens, xs, om, cps, cos = [], [], [], [], []
        for r in rs:
            l, x, o, cp, co = data(r)
            lens += l; xs.append(x); om.append(o); cps.append(cp); cos.append(co)
        cu = torch.tensor([0] + torch.tensor(lens).cumsum(0).tolist(), dtype=torch.int32, device=dev)
        batch = {"x_packed": torch.cat(xs).to(dev), "cu_seqlens": cu, "max_len": max(lens), "omic": torch.cat(om).to(dev)}
        for p in runner.parameters():
            p.grad = None
        loss = runner(batch, torch.cat(cps).to(dev) * scale, torch.cat(cos).to(dev) * scale)
        loss.backward()
        return [p.grad.clone() if p.grad is not None else torch.zeros_like(p) for p in runner.parameters()]

    grads_of([rank], 1.0)
    S.allreduce_gradients(runner, world)
    reduced = [p.grad.clone() for p in runner.parameters()]
    out = {}
    if rank == 0:
        # mean over ranks of [sum_b c.cot + mean_b modularity] == [sum over all bags of c.cot/W + mean over all bags]
        single = grads_of(list(range(world)), 1.0 / world)
        out["slide_parallel_grad_rel_err_max"] = max(rel(a, b) for a, b in zip(reduced, single))
        out["slide_parallel_what"] = "all-reduced (mean) gradients of %d ranks x 2 bags vs one process running all %d bags" % (world, 2 * world)
    # ---- giant-bag mode ----
    n = 4096
    g = torch.Generator().manual_seed(77)
    xg = torch.randn(n, D_IN, generator=g).bfloat16().to(dev)
    cot = (torch.randn(1, P, 256, generator=g) * 1e-2).to(dev)
    blocks = [ops.block_params(b) for b in net.proto_g_blocks]
    a, b = PAR.shard_bounds(n, world)[rank]

    def giant(x, cu, max_len, group, row0):
        for p in net.parameters():
            p.grad = None
        c, h = ops.proto_fusion(x, cu, max_len, net.p_proto, net.path_net[0].weight, net.path_net[0].bias, blocks,
                                shard_group=group)
        cl = c.detach().clone().requires_grad_(True)
        if group is not None:
            t = MOD.modularity_terms_sharded(h, row0, n, cl, group=group)[0, 0]
        else:
            t = MOD.modularity_terms(h, cu, n, cl)[0, 0]
        t.backward()
        (c * cot).sum().backward()
        return c.detach(), t.detach(), cl.grad, net.path_net[0].weight.grad.clone()

    cs, ts, dcs, dws = giant(xg[a:b].contiguous(), torch.tensor([0, b - a], dtype=torch.int32, device=dev), max(1, b - a),
                             dist.group.WORLD, a)
    if rank == 0:
        c1, t1, dc1, dw1 = giant(xg, torch.tensor([0, n], dtype=torch.int32, device=dev), n, None, 0)
        out["sharded_pool_token_rel_err"] = rel(cs, c1)
        out["sharded_pool_dw1_rel_err"] = rel(dws, dw1)
        out["sharded_modularity_loss_rel_err"] = abs(float(ts) - float(t1)) / abs(float(t1))
        out["sharded_modularity_grad_rel_err"] = rel(dcs, dc1)
        out["sharded_what"] = "one %d-patch bag split by rows over %d ranks (LSE-merged pooling, two-phase modularity) vs one GPU" % (n, world)
    dist.barrier()
    return out


def giant_measure(dev, rank, world, P, steps, warmup, with_single):
    """configs[3]: ONE 120 000-patch bag, fwd + bwd incl. the modularity term, rows sharded over `world` ranks;
    with_single: rank 0 also runs the whole bag alone (the strong-scaling denominator) and the merged tokens are
    compared.  Returns a dict on rank 0 (None elsewhere)."""
    import torch
    import torch.distributed as dist
    from imp_b200 import _lib, model as M, modularity as MOD, ops, parallel as PAR
    n = 120000
    torch.manual_seed(1234)                                   # identical parameters on every rank
    net = M.IMPHotPath(n_proto=P, dropout=0.0 if with_single else 0.25, seed=0).to(dev).train()
    bounds = PAR.shard_bounds(n, world)
    a, b = bounds[rank]

    def shard(r):
        lo, hi = bounds[r]
        return torch.randn(hi - lo, D_IN, device=dev, generator=torch.Generator(device=dev).manual_seed(100 + r)).bfloat16()

    x = shard(rank)
    cot = torch.randn(1, P, 256, device=dev, generator=torch.Generator(device=dev).manual_seed(100)) * 1e-2
    group = dist.group.WORLD if world > 1 else None
    params = [p for p in net.parameters()]
    blocks = [ops.block_params(blk) for blk in net.proto_g_blocks]

    def make_step(xs, rows, grp, row0):
        cu = torch.tensor([0, rows], dtype=torch.int32, device=dev)

        def step():
            for p in params:
                p.grad = None
            c, h = ops.proto_fusion(xs, cu, max(1, rows), net.p_proto, net.path_net[0].weight, net.path_net[0].bias, blocks,
                                    p_drop=net.dropout, seed=net._seed() if net.dropout > 0 else 0, shard_group=grp)
            loss = (c * cot).sum()
            if grp is not None:
                loss = loss + MOD.modularity_terms_sharded(h, row0, n, c, group=grp)[0, 0]
            else:
                loss = loss + MOD.modularity_terms(h, cu, rows, c)[0, 0]
            loss.backward()
            return c.detach()
        return step

    def make_loss(xs, rows, grp, row0):
        cu = torch.tensor([0, rows], dtype=torch.int32, device=dev)
        seed = net._seed() if net.dropout > 0 else 0          # a constant inside a graph; GraphedStep's seed word varies it

        def loss_fn():
            c, h = ops.proto_fusion(xs, cu, max(1, rows), net.p_proto, net.path_net[0].weight, net.path_net[0].bias, blocks,
                                    p_drop=net.dropout, seed=seed, shard_group=grp)
            loss = (c * cot).sum()
            if grp is not None:
                return loss + MOD.modularity_terms_sharded(h, row0, n, c, group=grp)[0, 0]
            return loss + MOD.modularity_terms(h, cu, rows, c)[0, 0]
        return loss_fn

    graph_check = []                                              # |graph loss - eager loss| / |eager loss| per captured step

    def time_graph(loss_fn, sync_group):
        """The same step replayed from ONE CUDA graph (collectives captured with it): no launch gaps between the ~160
        small launches and 12 collectives of a step.  Returns ms per step, or raises."""
        from imp_b200 import step as S
        with torch.no_grad():
            eager_loss = float(loss_fn())                         # dropout is off when this is compared (with_single)
        gs = S.GraphedStep(None).capture_fn(loss_fn, params, dev)
        try:
            for _ in range(max(1, warmup)):
                out = gs.replay()
            torch.cuda.synchronize()
            graph_check.append(abs(float(out) - eager_loss) / max(abs(eager_loss), 1e-30))
            if sync_group:
                dist.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(steps):
                gs.replay()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if sync_group:
                tms = torch.tensor([ms], device=dev)
                dist.all_reduce(tms, op=dist.ReduceOp.MAX)
                ms = float(tms.item())
            return ms / steps
        finally:
            gs.close()

    def time_it(step, sync_group):
        for _ in range(max(1, warmup)):
            c = step()
        torch.cuda.synchronize()
        if sync_group:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = _lib.launch_count()
        e0.record()
        for _ in range(steps):
            c = step()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if sync_group:
            tms = torch.tensor([ms], device=dev)
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
            ms = float(tms.item())
        return ms / steps, c, _lib.launch_count() - l0

    ms_n, c_n, launches = time_it(make_step(x, b - a, group, a), world > 1)
    ms_graph = graph_err = None
    if world > 1 and os.environ.get("IMP_GIANT_GRAPH", "1") != "0":
        ok = torch.ones(1, device=dev)
        try:
            ms_graph = time_graph(make_loss(x, b - a, group, a), True)
        except Exception as exc:                                  # every rank must agree on whether the number exists
            graph_err = "%s: %s" % (type(exc).__name__, str(exc)[:160])
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) == 0.0:
            ms_graph = None
    agree = 0.0
    if world > 1:
        cs = [torch.empty_like(c_n) for _ in range(world)]
        dist.all_gather(cs, c_n.contiguous())
        agree = max(float((ci - cs[0]).abs().max().item()) for ci in cs)    # merged tokens are identical on every rank
    rec = None
    if rank == 0:
        rec = {"workload": "configs[3]: 1 slide of 120000x512 patches, %d prototypes, rows sharded over %d GPU(s), fwd+bwd incl. modularity" % (P, world),
               "ms_per_step_sharded": ms_n, "n_gpus": world, "max_abs_token_difference_between_ranks": agree,
               "gpu_launches": int(launches), "rows_per_rank_mib": round((b - a) * D_IN * 2 / 2 ** 20, 1)}
        if with_single and world > 1:
            xall = torch.cat([x] + [shard(r) for r in range(1, world)])
            ms_1, c_1, _ = time_it(make_step(xall, n, None, 0), False)
            rec["ms_per_step_1gpu"] = ms_1
            rec["strong_scaling_speedup"] = ms_1 / ms_n
            if ms_graph is not None:
                try:
                    ms_1g = time_graph(make_loss(xall, n, None, 0), False)
                    rec.update({"ms_per_step_sharded_graph": ms_graph, "ms_per_step_1gpu_graph": ms_1g,
                                "strong_scaling_speedup_graph": ms_1g / ms_graph,
                                "graph": "the per-rank step incl. its NCCL collectives replayed from one CUDA graph, both arms",
                                "graph_vs_eager_loss_rel_diff": max(graph_check) if not net.dropout else None})
                except Exception as exc:
                    rec["graph_error"] = "1-GPU arm: %s: %s" % (type(exc).__name__, str(exc)[:160])
            elif graph_err:
                rec["graph_error"] = graph_err
            rec["token_max_abs_diff_sharded_vs_1gpu"] = float((c_1 - c_n).abs().max().item())
            rec["token_rel_err_sharded_vs_1gpu"] = float(((c_1 - c_n).norm() / c_1.norm()).item())
    if world > 1:
        dist.barrier()
    return rec


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import imp_b200
    from imp_b200 import _lib, model as M, ops, step as S

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the IMP hot path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    tf_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))
    peak_src = "measured" if peaks else "fallback"

    sampler = ClockSampler(local) if rank == 0 else None     # started early: nvidia-smi needs ~1 s before its first row
    windows = []                                             # host-clock intervals of the timed training steps
    B, N, P = args.bags, args.patches, args.protos
    torch.manual_seed(1234 + rank)
    net = M.IMPHotPath(n_proto=P, dropout=0.25, seed=0).to(dev)
    runner = S.HotPathStep(net, with_modularity=True).to(dev).train()
    gen = torch.Generator(device=dev).manual_seed(100 + rank)
    x = torch.randn(B * N, D_IN, device=dev, generator=gen).bfloat16()          # 16 MiB per slide
    cu = torch.arange(0, (B + 1) * N, N, dtype=torch.int32, device=dev)
    omic = torch.rand(B, sum(GROUP_SIZES), device=dev, generator=gen)
    cot_p = torch.randn(B, P, 256, device=dev, generator=gen) * 1e-2
    cot_o = torch.randn(B, N_PATHWAYS + 1, 256, device=dev, generator=gen) * 1e-2
    batch = {"x_packed": x, "cu_seqlens": cu, "max_len": N, "omic": omic}
    if args.missing_omics > 0:      # incompletely paired batch: whole-sample masks, imputed inside the encoder kernel
        miss = (torch.rand(B, device=dev, generator=gen) < args.missing_omics)
        batch["insample_without_omic"] = miss[:, None].expand(B, sum(GROUP_SIZES)).to(torch.int32).contiguous()
        net.omic_means = torch.full((sum(GROUP_SIZES),), 0.5, device=dev)
    params = [p for p in runner.parameters()]

    def one_step(b, cp, co, with_mod=True, lengths=None):
        runner.with_modularity = with_mod
        for p in params:
            p.grad = None
        loss = runner(b, cp, co, lengths)
        loss.backward()
        if world > 1:
            S.allreduce_gradients(runner, world)
        return loss

    def timed(fn, steps, warmup, profile=False):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        if profile:
            _lib.profile_collect()
            _lib.profile_enable(True)
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        recs = []
        if profile:
            _lib.profile_enable(False)
            recs = _lib.profile_collect()
        launches = _lib.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, recs, launches, (t0, t1)

    # ---- headline: device-resident inputs, training step incl. modularity ----
    ms, recs, launches, win = timed(lambda: one_step(batch, cot_p, cot_o, True), args.steps, args.warmup, profile=True)
    windows.append(win)
    value = world * B * args.steps / (ms * 1e-3)

    # ---- the same step without the O(N^2) modularity term (eval-mode cost of the streaming path) ----
    ms_s, recs_s, _, _ = timed(lambda: one_step(batch, cot_p, cot_o, False), args.steps, args.warmup, profile=True)
    value_s = world * B * args.steps / (ms_s * 1e-3)

    # ---- the same two steps captured in CUDA graphs (what a deployment replays; the eager numbers above carry
    #      the per-kernel event timing).  world > 1: the NCCL all-reduce stays outside the graph. ----
    eager = {"value": value, "ms_per_step": ms / args.steps, "streaming_only_value": value_s,
             "streaming_only_ms_per_step": ms_s / args.steps}
    graph_note = "training and streaming-only steps replayed from CUDA graphs (step.GraphedStep); 'eager' = same steps launched from Python"
    if not args.no_graph:
        try:
            def graphed(with_mod):
                runner.with_modularity = with_mod
                gs = S.GraphedStep(runner).capture(batch, cot_p, cot_o)

                def fn():
                    gs.replay()
                    if world > 1:
                        S.allreduce_gradients(runner, world)
                m, _, _, win = timed(fn, args.steps, args.warmup)
                if with_mod:
                    windows.append(win)
                    # the sampler reports every 100 ms and the timed region may be shorter than that: keep the same
                    # load running (outside the timing) until a few rows have landed inside a window
                    t_end = time.perf_counter() + 2.0
                    while sampler is not None and world == 1 and sampler.count(windows) < 5 and time.perf_counter() < t_end:
                        t0 = time.perf_counter()
                        for _ in range(4):
                            fn()
                        torch.cuda.synchronize()
                        windows.append((t0, time.perf_counter()))
                gs.close()
                return m
            ms = graphed(True)
            ms_s = graphed(False)
            value = world * B * args.steps / (ms * 1e-3)
            value_s = world * B * args.steps / (ms_s * 1e-3)
        except Exception as exc:                      # keep the eager measurement, say why
            graph_note = "CUDA-graph capture unavailable (%s: %s); eager numbers reported" % (type(exc).__name__, str(exc)[:200])
    else:
        graph_note = "eager launches (--no-graph)"

    clocks = sampler.stop(windows) if sampler else None

    # ---- per-kernel roofline from the event-bracketed launches of the timed region ----
    def per_kernel(records, steps):
        agg = {}
        for name, t in records:
            a = agg.setdefault(name, [0.0, 0])
            a[0] += t; a[1] += 1
        return {k: {"ms_per_step": v[0] / steps, "launches_per_step": v[1] / steps, "avg_ms": v[0] / v[1]} for k, v in agg.items()}

    pk = per_kernel(recs, args.steps)
    pk_s = per_kernel(recs_s, args.steps)
    traffic_tab, traffic_src = ncu_traffic(B)
    rows = B * N
    alg = {   # algorithmic bytes / flops per launch
        "pathnet_fwd": ("hbm", rows * D_IN * 2.0), "pathnet_dw": ("hbm", rows * D_IN * 2.0),
        "pool_fwd": ("hbm", rows * 256 * 2.0), "pool_bwd_dq": ("hbm", rows * 256 * 2.0),
        "pool_bwd_dz": ("hbm", rows * 256 * 2.0 * 2),
        "modularity_degrees_gram": ("tensor", 2.0 * B * N * N * 256), "modularity_sweep": ("tensor", 2.0 * B * N * N * 256),
        # O(N) kernels of the modularity term: bytes they have to move (h read, xh + assignments written / read back)
        "modularity_prep": ("hbm", rows * (256 * 2.0 * 2 + 40 * 4.0)), "modularity_degrees_closed": ("hbm", rows * 256 * 2.0),
        "modularity_finish": ("hbm", rows * (256 * 2.0 + 2 * 40 * 4.0)),
    }
    kernels_out = []
    tot_kernel_ms = sum(v["ms_per_step"] for v in pk.values()) or 1.0
    for name, v in sorted(pk.items(), key=lambda kv: -kv[1]["ms_per_step"]):
        ent = {"kernel": name, "ms_per_step": round(v["ms_per_step"], 4), "share_of_kernel_time": round(v["ms_per_step"] / tot_kernel_ms, 4),
               "launches_per_step": v["launches_per_step"]}
        if name in alg:
            bound, work = alg[name]
            ach = work / (v["avg_ms"] * 1e-3) / (1e9 if bound == "hbm" else 1e12)
            peak = hbm_peak if bound == "hbm" else tf_peak
            ent.update({"bound": bound, "achieved": round(ach, 2), "peak": peak, "unit": "GB/s" if bound == "hbm" else "TFLOP/s",
                        "frac": round(ach / peak, 4)})
            if name == "modularity_degrees_gram" and ach > peak:
                # post-ReLU features: the closed-form degree kernel did the work and this launch returned at its flag test
                ent = {k: ent[k] for k in ("kernel", "ms_per_step", "share_of_kernel_time", "launches_per_step")}
                ent["note"] = "early exit (features non-negative: closed-form degrees in use)"
        kernels_out.append(ent)
    dom = kernels_out[0] if kernels_out else {}
    roofline = {"kernel": dom.get("kernel"), "bound": dom.get("bound", "hbm"), "achieved": dom.get("achieved"),
                "peak": dom.get("peak"), "unit": dom.get("unit"), "frac": dom.get("frac"), "traffic": None,
                "peak_source": peak_src + " (MEASURED_PEAKS.json, sustained bf16 / copy bandwidth)",
                "note": "dominant kernel of the training step is the modularity pair sweep: tcgen05 Gram tiles (the FLOPs counted "
                        "in 'achieved') feed a (min,+) contraction over the 40 token slots plus a tanh/gradient tail on the CUDA "
                        "cores; it is issue-bound there (see 'cuda_core_issue'), not tensor- or HBM-bound"}
    if dom.get("kernel") == "modularity_sweep":
        pairs = float(B) * N * N
        t = pk["modularity_sweep"]["avg_ms"] * 1e-3
        clk = (clocks or {}).get("sm_mhz") or 1965.0
        sm_count = torch.cuda.get_device_properties(dev).multi_processor_count
        roofline["cuda_core_issue"] = {
            "pairs_per_s": pairs / t, "token_pair_evals_per_s": pairs * (P + N_PATHWAYS + 1) / t,
            "sm_cycles_per_32_pairs": t * clk * 1e6 * sm_count * 4 / (pairs / 32),
            "warp_instructions_per_32_pairs": 87, "alu_pipe_cycles_per_32_pairs": 92,
            "what": "cycles of one scheduler sub-partition per warp of 32 pairs; 87 instructions at 1 IPC would be the "
            "issue bound, 92 cycles the ALU-pipe bound (19.5 FMNMX3 at 3.7 cycles + 10 other ALU instructions at 2, "
            "measured pipe rates in profiles/r01_sweep_iterations.md)"}
        roofline["cuda_core_issue"]["frac_of_alu_pipe_bound"] = round(92.0 / roofline["cuda_core_issue"]["sm_cycles_per_32_pairs"], 4)
        roofline["traffic"] = traffic_tab.get("modularity_sweep")
        roofline["traffic_source"] = traffic_src
    # fused streaming path (the metric's 'fused-kernel HBM GB/s'): x read once forward + once backward
    stream_names = ["pathnet_fwd", "pool_fwd", "pool_merge", "pool_bwd_dq", "pool_bwd_dz", "reduce_dq", "reduce_db", "pathnet_dw",
                    "sum_partials", "cast_bf16"]
    stream_ms = sum(pk_s[k]["ms_per_step"] for k in stream_names if k in pk_s)
    stream_traffic, traffic_missing = None, []
    if traffic_tab:      # the five large kernels carry > 99 % of the bytes; merge / reduce kernels move a few MB and are listed if uncaptured
        stream_traffic = sum(traffic_tab[k] * pk_s[k]["launches_per_step"] for k in stream_names if k in pk_s and k in traffic_tab)
        traffic_missing = [k for k in stream_names if k in pk_s and k not in traffic_tab]
    alg_stream = 2.0 * rows * D_IN * 2.0
    roofline_stream = {"bound": "hbm", "achieved": round(alg_stream / (stream_ms * 1e-3) / 1e9, 2) if stream_ms else None,
                       "peak": hbm_peak, "unit": "GB/s", "traffic": stream_traffic, "traffic_source": traffic_src,
                       "traffic_kernels_without_capture": traffic_missing,
                       "kernels": [k for k in stream_names if k in pk_s], "kernel_ms_per_step": round(stream_ms, 4),
                       "algorithmic_bytes_per_step": alg_stream}
    if roofline_stream["achieved"]:
        roofline_stream["frac"] = round(roofline_stream["achieved"] / hbm_peak, 4)

    # ---- e2e: a batch in pinned host memory -> H2D -> (strip/pack) -> step -> D2H loss, every step ----
    # Two host layouts: the native wire format (imp_b200.wire: valid rows only, bf16, cu_seqlens) and
    # the reference batch dict (img (B,N,512) fp32 as data_manager.py:395-403 collates it) as the compatibility path.
    # The copies run on a second stream into a double buffer (batch k+1 uploads while step k computes) and the loss
    # of step k is read after step k+1 has been queued, as a training loop would do.
    def run_e2e(layout):
        Be = args.e2e_bags
        omic_h = torch.rand(Be, sum(GROUP_SIZES)).pin_memory()
        if layout == "packed_bf16":
            from imp_b200 import wire
            feats = torch.empty(Be * N, D_IN, dtype=torch.float32).normal_(generator=torch.Generator().manual_seed(7 + rank))
            host = wire.pack_bags([feats[i * N:(i + 1) * N] for i in range(Be)], pin=True, strip=False)
            del feats
            src, cu_e = host["x_packed"], host["cu_seqlens"].to(dev)
            mk = lambda: torch.empty(Be * N, D_IN, device=dev, dtype=torch.bfloat16)
            as_batch = lambda sl: {"x_packed": sl["img"], "cu_seqlens": cu_e, "max_len": N, "omic": sl["omic"]}
            desc = "imp_b200.wire packed batch: x (B*%d,512) bf16 pinned + cu_seqlens, omic (B,3354) fp32" % N
        else:
            src = torch.empty(Be, N, D_IN, dtype=torch.float32).pin_memory()
            src.normal_(generator=torch.Generator().manual_seed(7 + rank))
            mk = lambda: torch.empty(Be, N, D_IN, device=dev)
            as_batch = lambda sl: {"img": sl["img"], "omic": sl["omic"]}
            desc = "reference batch dict: img (B,%d,512) fp32 pinned, omic (B,3354) fp32" % N
        cp, co = cot_p[:Be].contiguous(), cot_o[:Be].contiguous()
        sink = torch.empty(2, dtype=torch.float32).pin_memory()
        copy_stream = torch.cuda.Stream(device=dev)
        main_stream = torch.cuda.current_stream(dev)
        dbuf = [{"img": mk(), "omic": torch.empty(Be, sum(GROUP_SIZES), device=dev),
                 "ready": torch.cuda.Event(), "free": torch.cuda.Event()} for _ in range(2)]
        state = {"k": 0, "pending": None}

        def upload(slot):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(slot["free"])             # the step that last read this buffer has finished
                slot["img"].copy_(src, non_blocking=True)
                slot["omic"].copy_(omic_h, non_blocking=True)
                slot["ready"].record(copy_stream)

        for sl in dbuf:
            sl["free"].record(main_stream)
        upload(dbuf[0])

        # one CUDA graph per buffer slot (the step reads fixed device addresses), the NCCL all-reduce stays outside
        graphs = [None, None]
        if not args.no_graph:
            try:
                runner.with_modularity = True
                for i, sl in enumerate(dbuf):
                    main_stream.wait_event(dbuf[0]["ready"])
                    sl["img"].copy_(dbuf[0]["img"]) if i else None
                    sl["omic"].copy_(dbuf[0]["omic"]) if i else None
                    graphs[i] = S.GraphedStep(runner).capture(as_batch(sl), cp, co, lengths=None)
            except Exception:
                graphs = [None, None]

        def e2e_step():
            k = state["k"]
            cur, nxt = dbuf[k % 2], dbuf[(k + 1) % 2]
            upload(nxt)                                           # H2D of the next batch overlaps this step
            main_stream.wait_event(cur["ready"])
            if graphs[k % 2] is not None:
                loss = graphs[k % 2].replay()
                if world > 1:
                    S.allreduce_gradients(runner, world)
            else:
                loss = one_step(as_batch(cur), cp, co, True, lengths=None)
            cur["free"].record(main_stream)
            if state["pending"] is not None:                      # D2H of the previous step's loss
                state["pending"].synchronize()
            sink[k % 2:k % 2 + 1].copy_(loss.detach().reshape(1), non_blocking=True)
            ev = torch.cuda.Event(); ev.record(main_stream)
            state["pending"] = ev
            state["k"] = k + 1

        n_e = max(2, args.steps // 2)
        ms_e, _, _, _ = timed(e2e_step, n_e, min(args.warmup, 2))
        for g_ in graphs:
            if g_ is not None:
                g_.close()
        return {"value": world * Be * n_e / (ms_e * 1e-3), "unit": "bags/s",
                "h2d_bytes_per_step": int(src.numel() * src.element_size() + omic_h.numel() * 4), "d2h_bytes_per_step": 4,
                "bags_per_step": Be, "host_layout": desc,
                "overlap": ("H2D double-buffered on a copy stream; loss read back one step behind; "
                            + ("step replayed from a CUDA graph per buffer slot" if graphs[0] is not None else "eager launches"))}

    e2e = None
    if not args.no_e2e:
        e2e = run_e2e("packed_bf16")
        torch.cuda.empty_cache()
        e2e["reference_layout"] = run_e2e("reference_fp32")
        torch.cuda.empty_cache()

    # ---- the reference arithmetic in eager PyTorch ON THIS GPU (BASELINE.md 5 / SURVEY 8(d)): same ATen op sequence
    #      as umeml_gan.py:410,425-434 + ops/utils.py:188-228 (materialises N x N and P x N x N), fp32, TF32 off ----
    gpu_eager = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        gpu_eager = gpu_eager_baseline(dev)

    # ---- the drop-in model end to end: build_model("umeml_gan", cfg) -> forward (7-tuple) -> NLL + KD + modularity ->
    #      backward, i.e. the hot path PLUS the token-level tail (Nystrom layers, bottleneck fusion, classifier) ----
    drop_in = None
    if rank == 0 and world == 1 and not args.no_e2e:
        try:
            drop_in = drop_in_model_step(dev, x, cu, omic, B, N, P)
        except Exception as exc:
            drop_in = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:200])}
        torch.cuda.empty_cache()

    # ---- multi-GPU correctness + the giant-bag strong-scaling record, so that the scaling run carries them ----
    dp_check = giant = None
    if world > 1:
        dp_check = run_dp_check(dev, rank, world)
        giant = giant_measure(dev, rank, world, P, steps=3, warmup=2, with_single=True)

    # ---- CPU baseline (oracle port), rank 0, N = 1 only ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        dt, _ = cpu_port_step(N, P)
        cpu = {"value": 1.0 / dt, "unit": "bags/s", "cores": cores, "kind": "port",
               "sample": "1 slide %dx%d, P=%d, fwd+bwd incl. modularity, oracle/imp_oracle.py on torch CPU fp32 (%.1f s)" % (N, D_IN, P, dt)}

    if rank == 0:
        out = {
            "metric": "wsi_bags_per_s_fwd_bwd", "value": value, "unit": "bags/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": workload_label(N, P), "bags_per_step_per_gpu": B, "patches": N, "prototypes": P, "pathways": N_PATHWAYS,
                       "modularity": True, "dropout": 0.25, "missing_omics_fraction": args.missing_omics,
                       "l2": "inputs larger than L2: %.0f MiB of bf16 features per step per GPU" % (B * N * D_IN * 2 / 2 ** 20),
                       "parallelism": "dp%d" % world},
            "streaming_only": {"value": value_s, "unit": "bags/s", "ms_per_step": ms_s / args.steps,
                               "what": "same step without the O(N^2) modularity term"},
            "launch_mode": graph_note, "eager": eager,
            "roofline": roofline, "roofline_streaming": roofline_stream, "kernels": kernels_out,
            "e2e": e2e, "cpu_baseline": cpu, "gpu_eager_baseline": gpu_eager, "drop_in_model": drop_in, "dp_check": dp_check, "giant": giant,
            "gpu_launches": int(launches), "clocks": clocks,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def run_kmeans(args):
    """BASELINE.json configs[4]: k-means prototype assignment, 2^20 patch features x 512 fp32 -> 32 centroids.
    A step = one assignment pass (distance + argmin) over the resident matrix (2 GiB, far larger than L2)."""
    import torch
    from imp_b200 import _lib, kernels
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    n, d, k = 1 << 20, D_IN, 32
    g = torch.Generator(device=dev).manual_seed(0)
    x = torch.randn(n, d, device=dev, generator=g)
    mu = x[torch.randperm(n, device=dev, generator=g)[:k]].clone()
    for _ in range(max(3, args.warmup)):
        a = kernels.kmeans_assign(x, mu)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        a = kernels.kmeans_assign(x, mu)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / args.steps
    # exactness on a sample against fp64 (the full oracle comparison lives in tests/test_fullsize_gpu.py)
    idx = torch.randperm(n, device=dev, generator=g)[:65536]
    d64 = torch.cdist(x[idx].double(), mu.double()) ** 2
    agree = float((d64.argmin(1) == a[idx].long()).float().mean().item())
    gbs = n * d * 4 / (ms * 1e-3) / 1e9
    print(json.dumps({
        "metric": "kmeans_assign_passes_per_s", "value": 1e3 / ms, "unit": "passes/s", "n_gpus": 1, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[4]: PLIP k-means prototype extraction, 2^20 x 512 fp32 -> 32 centroids, assignment pass",
                   "l2": "inputs larger than L2: 2048 MiB of fp32 features per pass"},
        "roofline": {"kernel": "kmeans_assign", "bound": "hbm", "achieved": round(gbs, 1), "peak": hbm_peak, "unit": "GB/s",
                     "frac": round(gbs / hbm_peak, 4), "traffic": None},
        "agreement_with_fp64_argmin_on_65536_rows": agree, "gpu_launches": int(_lib.launch_count() - l0),
    }))


def run_giant(args):
    """BASELINE.json configs[3]: giant-bag stress, 120 000 patches x 512 in ONE slide, 32 prototypes, sharded by
    rows over the ranks of the box: every pooling block exchanges its (pooled, lse) state once and merges it with
    the log-sum-exp kernel; the modularity term runs as prepare -> exchange -> sweep of the local row blocks.
    A step = forward + backward of that one bag (strong scaling: the bag is fixed, the ranks split it)."""
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    rec = giant_measure(dev, rank, world, args.protos, args.steps, max(1, args.warmup), with_single=False)
    if rank == 0:
        ms = rec["ms_per_step_sharded"]
        print(json.dumps({
            "metric": "giant_bags_per_s_fwd_bwd", "value": 1e3 / ms, "unit": "bags/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(1, args.warmup), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": rec["workload"], "modularity": True, "dropout": 0.25,
                       "l2": "inputs larger than L2: %.0f MiB of bf16 features per rank" % rec["rows_per_rank_mib"],
                       "parallelism": "rows%d" % world},
            "max_abs_token_difference_between_ranks": rec["max_abs_token_difference_between_ranks"],
            "gpu_launches": rec["gpu_launches"],
        }))
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    # Exactly one line on stdout: libraries (NCCL prints its version banner there) write to fd 1 behind Python's
    # back, so fd 1 is pointed at stderr for the run and the JSON line goes to the saved descriptor.
    sys.stdout.flush()
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real_stdout, "w", buffering=1)
    if args.workload == "kmeans":
        run_kmeans(args)
    elif args.workload == "giant":
        run_giant(args)
    elif args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
