#!/usr/bin/env python
"""bench.py -- train triples/s of the R-TuckER hot path (1-N fwd + bwd + RSGD step incl. retraction)
and filtered-eval queries/s, on a synthetic graph of WN18RR shape (BASELINE.json configs[0]/metric).

    python bench.py --gpus 1 --steps K --warmup W            # this repo's CUDA path (one JSON line)
    python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (port, oracle/)
    torchrun ... bench.py --gpus N ...                       # entity-sharded over NCCL (strong scaling)

A step = one optimiser step (fit + step) on one batch of 512 (s,r) queries.
`value`  : triples/s with the batches already resident in HBM (CUDA events around K steps).
`e2e`    : same metric through the public API (model(...) -> FusedLoss -> optimizer.fit/step) with the
           batch coming from pinned HOST memory every step and the loss read back to the host.
`roofline`: the fused score+BCE+backward kernel (algorithmic flops 6*B*N*r2) against the measured bf16
           tensor peak of MEASURED_PEAKS.json (FP32-FFMA variant runs on the CUDA cores; the fraction
           says how far it is from the tensor roofline the tcgen05 variant is judged on).
`cpu_baseline`: the reference's step (oracle/reference_step.py port) timed on this host's cores.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (n_entities, n_relations incl. reverse, rank, n_queries, mean objects per query)
    "wn18rr": dict(N=40943, M=22, rank=(10, 200, 200), Q=103509, mean_obj=1.678, sym=False),
    "fb15k237": dict(N=14541, M=474, rank=(200, 20, 20), Q=149689, mean_obj=3.636, sym=False),
    "wn18rr-sym": dict(N=40943, M=22, rank=(10, 200, 200), Q=103509, mean_obj=1.678, sym=True),
    "synthetic-1m": dict(N=1000000, M=1000, rank=(200, 200, 200), Q=200000, mean_obj=5.0, sym=False),
}
BATCH = 512
LABEL_SMOOTHING = 0.1
MOMENTUM = 0.8
LR = 109.09      # OneCycleLR(max_lr=600, div_factor=5.5) at epoch 1 (train.py:213-215)
REG = 1e-11      # configs/base_config.py:19


def synth_graph(w, seed=1234):
    """Seeded synthetic (s,r)->objects vocabulary of the workload's shape."""
    rng = np.random.default_rng(seed)
    N, M, Q = w["N"], w["M"], w["Q"]
    pop = 1.0 / np.arange(1, N + 1) ** 0.6
    pop /= pop.sum()
    perm = rng.permutation(N)
    keys = set()
    feats = np.zeros((Q, 2), np.int32)
    n = 0
    while n < Q:
        s = perm[rng.choice(N, size=Q, p=pop)]
        r = rng.integers(0, M, size=Q)
        for a, b in zip(s.tolist(), r.tolist()):
            if (a, b) not in keys:
                keys.add((a, b))
                feats[n] = (a, b)
                n += 1
                if n == Q:
                    break
    extra = rng.geometric(1.0 / w["mean_obj"], size=Q)  # >= 1, mean = mean_obj
    extra = np.minimum(extra, 400)
    off = np.zeros(Q + 1, np.int64)
    np.cumsum(extra, out=off[1:])
    idx = perm[rng.choice(N, size=int(off[-1]), p=pop)].astype(np.int32)
    # unique + ascending per query
    lists = [np.unique(idx[off[i]:off[i + 1]]) for i in range(Q)]
    cnt = np.asarray([len(x) for x in lists], np.int64)
    off = np.zeros(Q + 1, np.int64)
    np.cumsum(cnt, out=off[1:])
    return feats, off, np.concatenate(lists).astype(np.int32), cnt


def batch_arrays(feats, off, idx, cnt, items):
    c = cnt[items]
    boff = np.zeros(len(items) + 1, np.int32)
    np.cumsum(c, out=boff[1:])
    gather = np.repeat(off[items] - boff[:-1], c) + np.arange(int(boff[-1]), dtype=np.int64)
    return feats[items], boff, idx[gather]


def init_params(w, seed=20):
    """R_TuckER.init: xavier core + QR-orthonormalised xavier factors on the CPU generator."""
    from rtucker_b200 import asymmetric, symmetric
    torch.manual_seed(seed)
    mod = symmetric if w["sym"] else asymmetric
    model = mod.R_TuckER((w["N"], w["M"]), w["rank"])
    model.init(None)
    return model


class ClockSampler:
    def __init__(self):
        self.rows, self.proc = [], None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", os.environ.get("LOCAL_RANK", "0")],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def cpu_reference_steps(w, graph, n_steps, warmup, threads=None):
    """Times the reference's train step (port) on the host cores.  Dense targets are built outside the
    timed region (the reference builds them in DataLoader worker processes)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import reference_step as RS
    if threads:
        torch.set_num_threads(threads)
    feats, off, idx, cnt = graph
    model = init_params(w)
    if w["sym"]:
        st = RS.ReferenceStepper(model.core.data, model.R.weight.data, model.E.weight.data, None, MOMENTUM)
    else:
        st = RS.ReferenceStepper(model.core.data, model.R.weight.data, model.S.weight.data, model.O.weight.data,
                                 MOMENTUM)
    rng = np.random.default_rng(7)
    order = rng.permutation(len(cnt))
    triples, total = 0, 0.0
    for i in range(warmup + n_steps):
        items = order[i * BATCH:(i + 1) * BATCH]
        f, boff, bidx = batch_arrays(feats, off, idx, cnt, items)
        tg = RS.dense_targets(w["N"], torch.from_numpy(boff.astype(np.int64)), torch.from_numpy(bidx), LABEL_SMOOTHING)
        ft = torch.from_numpy(f.astype(np.int64))
        t0 = time.perf_counter()
        st.train_step(ft[:, 0], ft[:, 1], tg, REG, LR)
        dt = time.perf_counter() - t0
        if i >= warmup:
            total += dt
            triples += int(cnt[items].sum())
    return triples / total, total / n_steps, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = WORKLOADS[args.workload]
    graph = synth_graph(w)
    # torchrun exports OMP_NUM_THREADS=1; the reference arm uses every host core it can get
    tps, sps, cores = cpu_reference_steps(w, graph, args.steps, args.warmup, threads=os.cpu_count())
    line = {
        "impl": "reference", "metric": "train triples/s (1-N fwd+bwd+RSGD step), WN18RR shape",
        "value": tps, "unit": "triples/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": sps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": config_dict(w, args, "reference CPU path"),
        "cpu_baseline": {"value": tps, "unit": "triples/s", "cores": cores, "kind": "port",
                         "sample": f"{args.steps} full train steps of B={BATCH} (oracle/reference_step.py), dense targets built outside the timed region"},
        "e2e": {"value": tps, "unit": "triples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


def ncu_traffic(workload, variant):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu
    --set full capture of the same workload (profiles/ncu_traffic.json); None if never captured."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(path):
        return json.load(open(path)).get("%s:variant%d" % (workload, variant))
    return None


def config_dict(w, args, note):
    return {"workload": f"{args.workload}: synthetic graph of that shape, N={w['N']} entities, M={w['M']} relations, "
                        f"rank={w['rank']}, batch {BATCH}, {'SF-Tucker rgd' if w['sym'] else 'Tucker rsgd'} "
                        f"beta={MOMENTUM}, label_smoothing={LABEL_SMOOTHING}, lr={LR}, reg={REG}",
            "l2": "per-step working set (factors + tangent/momentum buffers) > 126 MB L2; no explicit flush",
            "note": note}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="wn18rr", choices=list(WORKLOADS))
    ap.add_argument("--variant", type=int, default=2,
                    help="fused score kernel: 0 = fp32 FFMA (1e-5 parity), 1 = tcgen05 TF32 (2e-3), "
                         "2 = warp-specialised tcgen05 with scaled fp16 operands (2e-3)")
    ap.add_argument("--cpu-steps", type=int, default=6, help="reference steps timed for cpu_baseline (0 = skip)")
    ap.add_argument("--eval-batches", type=int, default=8)
    ap.add_argument("--no-graphs", action="store_true", help="launch kernels eagerly instead of replaying CUDA graphs")
    ap.add_argument("--no-large-kernel", action="store_true", help="skip timing the fused kernel on the 1M-entity shard")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: rtucker_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    group = None
    if world > 1:
        import torch.distributed as dist
        # NCCL prints its version banner on stdout when the communicator is created; rank 0 must print ONE JSON line,
        # so stdout points at stderr until the communicator exists
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            group = dist.group.WORLD
            dist.all_reduce(torch.zeros(1, device=dev))
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    from rtucker_b200 import asymmetric, lib, symmetric
    from rtucker_b200.engine import SparseTargets
    from rtucker_b200.optim import FusedLoss
    from rtucker_b200.evaluation import rank_batch
    from rtucker_b200.manifold import SFTucker, Tucker

    w = WORKLOADS[args.workload]
    graph = synth_graph(w)
    feats, off, idx, cnt = graph
    model = init_params(w)
    N = w["N"]
    # ---- entity sharding: contiguous row blocks of S/O (or E) per rank ----
    per = (N + world - 1) // world
    n_begin, n_end = rank * per, min(N, (rank + 1) * per)
    with torch.no_grad():
        if w["sym"]:
            model.E.weight.data = model.E.weight.data[n_begin:n_end].contiguous()
        else:
            model.S.weight.data = model.S.weight.data[n_begin:n_end].contiguous()
            model.O.weight.data = model.O.weight.data[n_begin:n_end].contiguous()
    model.to(dev)
    mod = symmetric if w["sym"] else asymmetric
    kw = dict(group=group, n_total=N, n_begin=n_begin, score_variant=args.variant, use_graphs=(world == 1 and not args.no_graphs))
    if w["sym"]:
        opt = mod.RGD([model.core, model.E.weight, model.R.weight], w["rank"], LR, **kw)
    else:
        opt = mod.RSGDwithMomentum([model.core, model.S.weight, model.R.weight, model.O.weight], w["rank"], LR,
                                   MOMENTUM, **kw)
    opt.param_groups[0]["lr"] = LR

    rng = np.random.default_rng(7)
    order = rng.permutation(len(cnt))
    total_steps = args.warmup + args.steps
    host_batches, dev_batches, triples = [], [], []
    for i in range(2 * total_steps):
        items = order[(i * BATCH) % (len(order) - BATCH):][:BATCH]
        f, boff, bidx = batch_arrays(feats, off, idx, cnt, items)
        hb = [torch.from_numpy(np.ascontiguousarray(a)).pin_memory() for a in (f, boff, bidx)]
        host_batches.append(hb)
        triples.append(int(cnt[items].sum()))
        if i < total_steps:
            dev_batches.append([t.to(dev) for t in hb])
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def one_step(fd, od, xd):
        score_fn = model(fd[:, 0], fd[:, 1])
        opt.fit(FusedLoss(score_fn, SparseTargets(od, xd), LABEL_SMOOTHING, REG), None)
        opt.step()

    # ---- device-resident timing (CUDA graphs replay fit + step; inputs already in HBM) ----
    for i in range(args.warmup):
        one_step(*dev_batches[i])
    eng = opt._engine
    launches0 = lib().rt_launch_count()
    clocks = ClockSampler()
    barrier()
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.warmup, total_steps):
        one_step(*dev_batches[i])
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches_eager = int(lib().rt_launch_count() - launches0)
    n_tr = sum(triples[args.warmup:total_steps])
    t_all = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t_all, op=torch.distributed.ReduceOp.MAX)
    ms = float(t_all.item())
    value = n_tr / (ms * 1e-3)

    # ---- per-stage profile pass: same steps launched eagerly with CUDA-event brackets around every stage
    #      (the graph path cannot be bracketed); also counts the kernels of one step ----
    prof_steps = min(args.steps, 10)
    eng.timers = {}
    l0 = lib().rt_launch_count()
    barrier()
    for i in range(prof_steps):
        one_step(*dev_batches[args.warmup + i])
    barrier()
    launches_per_step = int(lib().rt_launch_count() - l0) // prof_steps
    stage_ms = eng.stage_ms()
    eng.timers = None
    launches = launches_per_step * args.steps if launches_eager == 0 else launches_eager

    # ---- end to end: host batches in, loss out, every step ----
    h2d = d2h = 0
    barrier()
    e0.record()
    for i in range(total_steps, total_steps + args.steps):
        hb = host_batches[i]
        db = [t.to(dev, non_blocking=True) for t in hb]
        h2d = sum(t.numel() * t.element_size() for t in hb)
        one_step(*db)
        loss_host = opt.loss.cpu()       # device -> host read of the step's result
        d2h = loss_host.numel() * loss_host.element_size()
    e1.record()
    barrier()
    clock_info = clocks.stop()
    ms_e2e = e0.elapsed_time(e1)
    t_all = torch.tensor([ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        torch.distributed.all_reduce(t_all, op=torch.distributed.ReduceOp.MAX)
    e2e_value = sum(triples[total_steps:total_steps + args.steps]) / (float(t_all.item()) * 1e-3)

    # ---- filtered evaluation (fused score + rank), queries/s ----
    eval_qps = None
    if args.eval_batches > 0:
        point = (SFTucker(model.core.data, [model.R.weight.data], 2, model.E.weight.data) if w["sym"] else
                 Tucker(model.core.data, [model.R.weight.data, model.S.weight.data, model.O.weight.data]))
        ebs = []
        for i in range(args.eval_batches):
            f, boff, bidx = host_batches[i]
            tgt = bidx[boff[:-1].long()]   # first known object of each query as the target
            f3 = torch.cat([f, tgt[:, None]], dim=1).contiguous().to(dev)
            ebs.append((f3, SparseTargets(boff.to(dev), bidx.to(dev))))
        rank_batch(point, ebs[0][0], ebs[0][1], n_begin, group)
        barrier()
        e0.record()
        for f3, flt in ebs:
            rank_batch(point, f3, flt, n_begin, group)
        e1.record()
        barrier()
        eval_qps = BATCH * len(ebs) / (e0.elapsed_time(e1) * 1e-3)

    # ---- the fused score kernel on its own, launched back to back on the step's inputs (CUDA events on the
    #      launching stream; its working set -- O, packed images, dO, per-CTA H partials -- exceeds the L2) ----
    score_timing = None
    if rank == 0 or world > 1:
        from rtucker_b200 import ops as _ops
        fd, od, xd = dev_batches[0]
        O_fac = model.E.weight.data if w["sym"] else model.O.weight.data
        B_, r2_ = BATCH, w["rank"][2]
        # the query rows of a real batch at the current (trained-for-a-few-steps) point, exactly what the step feeds
        # the kernel: q = core x1 R[rel] x2 S[sub]
        S_fac = model.E.weight.data if w["sym"] else model.S.weight.data
        r_rows = _ops.gather_rows(model.R.weight.data, fd[:, 1].contiguous())
        s_rows = _ops.gather_rows(S_fac, fd[:, 0].contiguous(), n_begin)
        if world > 1:
            torch.distributed.all_reduce(s_rows)
        qq = _ops.query_fwd(model.core.data, r_rows, s_rows)
        outs = (torch.empty(1, dtype=torch.float64, device=dev), torch.empty(B_, r2_, device=dev),
                torch.empty_like(O_fac))
        v = args.variant
        if v == 2 and not _ops.score_v3_supported(r2_):
            v = 1
        ws_s = torch.empty(int(lib().rt_score_bce_ws_bytes(B_, O_fac.shape[0], r2_, v)) + 16, dtype=torch.uint8, device=dev)
        def run_score(phases=7):
            _ops.score_bce_fwd_bwd(qq, None if v == 2 else qq, O_fac, od, xd, LABEL_SMOOTHING, n_total=N, b_total=B_,
                                   n_begin=n_begin, variant=v, out=outs, ws=ws_s, o_absmax=1.0 if v == 2 else None,
                                   phases=phases)
        reps = max(10, args.steps)
        def timed(phases):
            for _ in range(3):
                run_score(phases)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(reps):
                run_score(phases)
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / reps
        run_score(7)
        score_timing = {"op_ms": timed(7), "kernel_ms": timed(2) if v == 2 else None, "launches_timed": reps}
        # the same kernel on the 1M-entity / rank-200 shard of BASELINE configs[4] (single GPU only): at the WN18RR
        # size a launch is 12 (tile, chunk) pairs per SM, so prologue, tail and first-touch of the H partials weigh in
        roofline_1m = None
        if world == 1 and v == 2 and args.workload != "synthetic-1m" and not args.no_large_kernel:
            try:
                N1, r1m = 1000000, 200
                O1 = torch.randn(N1, r1m, device=dev) * (1.0 / N1 ** 0.5)
                q1 = torch.randn(B_, r1m, device=dev) * 4.0 * (N1 / r1m) ** 0.5      # logits ~ N(0, 4^2)
                out1 = (torch.empty(1, dtype=torch.float64, device=dev), torch.empty(B_, r1m, device=dev), torch.empty_like(O1))
                ws1 = torch.empty(int(lib().rt_score_bce_ws_bytes(B_, N1, r1m, 2)) + 16, dtype=torch.uint8, device=dev)
                def run1(phases=7):
                    _ops.score_bce_fwd_bwd(q1, None, O1, od, xd, LABEL_SMOOTHING, n_total=N1, b_total=B_, variant=2,
                                           out=out1, ws=ws1, o_absmax=1.0, phases=phases)
                run1(7)
                for _ in range(2):
                    run1(2)
                torch.cuda.synchronize()
                e0.record()
                for _ in range(5):
                    run1(2)
                e1.record()
                torch.cuda.synchronize()
                k_ms = e0.elapsed_time(e1) / 5
                roofline_1m = {"kernel_ms": k_ms, "flops": 6.0 * B_ * N1 * r1m, "launches_timed": 5}
                del O1, out1, ws1
            except RuntimeError as exc:      # out of memory on a smaller part: the number is simply absent
                roofline_1m = {"error": str(exc)[:120]}
        score_timing["large"] = roofline_1m

    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if rank != 0:
        return
    # ---- roofline of the fused score kernel ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peaks = json.load(open(peaks_path))
        peak_tf, peak_src = float(peaks["bf16_tflops_sustained"]), "MEASURED_PEAKS.json bf16_tflops_sustained"
    else:
        peak_tf, peak_src = 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"
    r2 = w["rank"][2]
    flops = 6.0 * BATCH * (n_end - n_begin) * r2
    score_ms = stage_ms.get("score_bce_fwd_bwd")
    if score_timing is not None:
        score_ms = score_timing["kernel_ms"] if score_timing.get("kernel_ms") else score_timing["op_ms"]
    achieved = flops / (score_ms * 1e-3) / 1e12 if score_ms else None
    roofline = {"kernel": "fused 1-N score + BCE + backward, variant %d (%s)" % (
                    args.variant, {0: "score_kernel, fp32 FFMA", 1: "score_tc_kernel, tcgen05 TF32",
                                   2: "score_v3_kernel, warp-specialised tcgen05 kind::f16 + its pack / reduce launches"}[args.variant]),
                "bound": "tensor", "achieved": achieved, "peak": peak_tf, "unit": "TFLOP/s",
                "frac": achieved / peak_tf if achieved else None, "traffic": ncu_traffic(args.workload, args.variant),
                "peak_source": peak_src, "algorithmic_flops_per_launch": flops, "ms_per_launch": score_ms,
                "timing": score_timing,
                "same_kernel_on_1m_entities": (None if not (score_timing and score_timing.get("large") and score_timing["large"].get("kernel_ms")) else {
                    "workload": "BASELINE configs[4] shard: N=1,000,000 entities, r2=200, B=512, one GPU",
                    "ms_per_launch": score_timing["large"]["kernel_ms"],
                    "achieved": score_timing["large"]["flops"] / (score_timing["large"]["kernel_ms"] * 1e-3) / 1e12,
                    "frac": score_timing["large"]["flops"] / (score_timing["large"]["kernel_ms"] * 1e-3) / 1e12 / peak_tf,
                    "unit": "TFLOP/s", "algorithmic_flops_per_launch": score_timing["large"]["flops"]}),
                "timing_note": "ms_per_launch = the fused kernel alone (kernel_ms: %d back-to-back launches between CUDA "
                               "events, on the query rows of a real batch at the current point); op_ms adds its operand "
                               "packing and H-reduction launches; in-step stage time is stage_ms.score_bce_fwd_bwd. The "
                               "epilogue has a warp-uniform short path for logits in (-27.7, 16.6): a launch on inputs with "
                               "many saturated logits is up to 1.6x slower (tools/zdist.py)"
                               % (score_timing["launches_timed"] if score_timing else 0)}

    cpu = None
    if args.cpu_steps > 0 and world == 1:
        tps, sps, cores = cpu_reference_steps(w, graph, args.cpu_steps, 1, threads=os.cpu_count())
        cpu = {"value": tps, "unit": "triples/s", "cores": cores, "kind": "port",
               "sample": f"{args.cpu_steps} full train steps of B={BATCH} after 1 warm-up (oracle/reference_step.py: "
                         f"the reference's autodiff-through-rank-2r step), {sps:.3f} s/step; dense targets built outside the timed region"}

    line = {
        "metric": "train triples/s (1-N fwd+bwd+RSGD step), WN18RR shape", "value": value, "unit": "triples/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "dtype_note": "fused score kernel: fp16 tensor-core operands scaled by powers of two (11 significant bits), fp32 "
                      "accumulation; tall-skinny passes fp32; N-independent stage fp64",
        "config": config_dict(w, args, "entity-sharded over %d GPU(s)" % world),
        "queries_per_s": BATCH * args.steps / (ms * 1e-3),
        "eval_queries_per_s": eval_qps,
        "e2e": {"value": e2e_value, "unit": "triples/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": launches,
        "gpu_launches_note": "kernels of this library executed in the timed region (%d per step; replayed from 2 CUDA graphs per step when graphs are on)" % launches_per_step,
        "cuda_graphs": bool(eng.use_graphs and eng._graphs),
        "roofline": roofline,
        "stage_ms": stage_ms,
        "stage_ms_note": "mean ms per stage over %d EAGERLY launched steps (CUDA events on the launching stream): stages made "
                         "of many short launches include host launch gaps, so the sum exceeds ms_per_step, which is "
                         "measured on the CUDA-graph path" % prof_steps,
        "cpu_baseline": cpu,
        "clocks": clock_info,
    }
    print(json.dumps(line))


if __name__ == "__main__":
    main()
