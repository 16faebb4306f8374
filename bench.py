#!/usr/bin/env python
"""Benchmark of the MV-KPConv hot path on B200 (contract: see the task statement / DESIGN.md §5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

Workload (BASELINE.json configs[1]): the KPConv baseline encoder-decoder of train_ScanNet_baseline
(14 rigid KPConv layers, 24.4 M parameters) forward + backward + SGD step on a stacked batch of 8
synthetic ScanNet-shaped spheres (r = 2 m, first_subsampling_dl = 0.04, K = 15) per GPU.  One
"step" is one pass of the hot path over one batch: the 5-level input pyramid (3 radius-neighbour
calls + 1 grid subsampling per level, on the GPU), the network forward, the loss, the backward
and the optimiser update.  Metric: stacked input points processed per second (whole job).

  value : inputs already resident in HBM when the timed region starts.
  e2e   : the same step fed from pinned HOST buffers (points, features, labels copied H2D every
          step) and ending with the loss read back to the host.
Weak scaling: every rank gets its own batch of 8 spheres (sharded by sphere); the only collective
is the DDP gradient all-reduce (NCCL).

The training step is replayed as a CUDA graph (harness.GraphedTrainStep; --no-graphs launches it eagerly).
Beside the headline the line carries (N = 1): `config0` = BASELINE configs[0] (one rigid 64->128 layer on one
~20k-point sphere: stage-A HBM fraction, contraction tensor fraction, CPU layer beside it), `neighbors` = the
radius-neighbour queries/s half of the metric (GPU vs the unmodified reference C++: one thread and P worker
processes) and `fp32_contraction` = the same step with the strict fp32 contraction.

--impl reference times the reference's CPU path (the unmodified reference C++ in oracle/_ref for
the pyramid + the torch-CPU restatement of KPConv for the network) on the host cores, on the SAME batch of 8
spheres.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "KPConv fwd+bwd points/s"
UNIT = "points/s"
IN_RADIUS = 2.0
FIRST_DL = 0.04
SPHERES_PER_GPU = 8


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


WORKLOAD = ("configs[1]: KPConv baseline encoder-decoder (train_ScanNet_baseline shape, 14 KPConv layers, 24.4M params) "
            "pyramid + fwd + bwd + SGD, one batch of 8 synthetic ScanNet-shaped spheres per GPU")


def workload_config(n_pts, limits):
    """The `config` object of the JSON line -- identical in the B200 arm and the reference arm."""
    return {"workload": WORKLOAD, "spheres_per_gpu": SPHERES_PER_GPU, "points_per_gpu": int(n_pts), "in_radius": IN_RADIUS,
            "first_subsampling_dl": FIRST_DL, "K": 15, "neighborhood_limits": [int(v) for v in limits],
            "parallelism": "one batch of 8 spheres per GPU (N > 1: sphere-sharded, gradient all-reduce)",
            "l2": "per-step working set (saved [N,15*Cin] operands, >1 GB) far exceeds the 126 MB L2; no flush"}


def host_features(pts):
    # in_features_dim = 2: constant 1 + height (train_ScanNet_baseline.py:183)
    return np.concatenate([np.ones((len(pts), 1), np.float32), pts[:, 2:3]], 1).astype(np.float32)


# =================================================================================================
# B200 arm
# =================================================================================================
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), "--query-gpu=" + self.Q,
                                       "--format=csv,noheader,nounits", "-lms", "25"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_batch(rank):
    """Host side of one rank's batch (untimed): 8 seeded synthetic spheres, stacked; first_subsampling_dl on the GPU."""
    import mvkpconv_b200 as mvk
    from mvkpconv_b200 import synthetic
    sub = lambda p, dl: mvk.grid_subsampling(p, sampleDl=dl)
    # weak scaling = the SAME amount of work on every GPU: every rank stacks the same 8 seeded spheres, rotated by its
    # rank (another batch order, other labels), so that the N-GPU job is N copies of the 1-GPU workload and the
    # efficiency the driver computes measures the communication, not a difference between the ranks' batches
    spheres = synthetic.make_spheres(SPHERES_PER_GPU, sub, seed=0, in_radius=IN_RADIUS, first_dl=FIRST_DL)
    spheres = spheres[rank % len(spheres):] + spheres[:rank % len(spheres)]
    pts_h, lens_h = synthetic.stack(spheres)
    feats_h = host_features(pts_h)
    labels_h = np.random.default_rng(rank).integers(0, 20, len(pts_h)).astype(np.int64)
    return spheres, pts_h, lens_h, feats_h, labels_h


def set_contraction(net, contraction):
    """Every module that contracts on the tensor cores (KPConv incl. nested offset convolutions, UnaryBlock)."""
    for m in net.modules():
        if hasattr(m, "contraction"):
            m.contraction = contraction


def run_b200(args):
    import torch.distributed as dist
    import mvkpconv_b200 as mvk
    from mvkpconv_b200 import _lib, harness, pyramid

    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback in the product path)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL prints its version banner to fd 1 when the communicator is created; stdout must carry
        # exactly ONE JSON line, so park fd 1 on stderr while the process group comes up
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    L = _lib.lib()

    # ---------------- synthetic batch of this rank (host side, untimed) ----------------
    spheres, pts_h, lens_h, feats_h, labels_h = make_batch(rank)
    n_pts = len(pts_h)
    pin = lambda a: torch.from_numpy(a).pin_memory()
    pts_p, feats_p, labels_p, lens_p = pin(pts_h), pin(feats_h), pin(labels_h), pin(lens_h)

    cfg = pyramid.baseline_config(in_radius=IN_RADIUS, first_subsampling_dl=FIRST_DL)
    pts_d, lens_d = pts_p.to(dev), lens_p.to(dev)
    cfg.neighborhood_limits = pyramid.calibrate_neighborhood_limits(pts_d, lens_d, cfg)
    feats_d, labels_d = feats_p.to(dev), labels_p.to(dev)

    def build_model(contraction):
        np.random.seed(0)
        torch.manual_seed(0)
        net = harness.KPFCNN(cfg).to(dev)
        set_contraction(net, contraction)
        if world > 1 and args.allreduce != "ddp":
            # identical replicas to start with (DDP would broadcast rank 0's parameters and buffers)
            with torch.no_grad():  # in-place writes that move the version counters (the bf16 weight pairs key on them)
                for t in list(net.parameters()) + list(net.buffers()):
                    dist.broadcast(t, 0)
        opt = torch.optim.SGD(net.parameters(), lr=1e-2, momentum=0.98, weight_decay=1e-3, fused=True)
        return net, opt

    net, opt = build_model(args.contraction)
    model = net
    use_ddp = world > 1 and args.allreduce == "ddp"
    if use_ddp:
        # gradients live inside the all-reduce buckets (no per-step copy); two buckets for ~97 MB of fp32 gradients
        model = torch.nn.parallel.DistributedDataParallel(net, device_ids=[local], gradient_as_bucket_view=True,
                                                          bucket_cap_mb=64)
    grad_params = [p for p in net.parameters() if p.requires_grad]
    averager = harness.OverlappedGradientAverager(net, split_level=3) if (world > 1 and args.allreduce == "overlap") else None

    def allreduce_grads(grads):
        """Gradient averaging right after backward, without DDP's buckets and per-parameter hooks: ONE grouped NCCL
        all-reduce over the per-parameter gradient tensors ('coalesced'), or one all-reduce of a packed copy ('flat')."""
        if len(grads) > 1 and args.allreduce == "flat":  # eager steps of the flat mode: pack, reduce, unpack
            flat = torch.cat([g.reshape(-1) for g in grads])
            dist.all_reduce(flat, op=dist.ReduceOp.AVG)
            torch._foreach_copy_(grads, [v.view_as(g) for v, g in zip(flat.split([g.numel() for g in grads]), grads)])
            return
        with dist._coalescing_manager(device=dev, async_ops=False):
            for g in grads:
                dist.all_reduce(g, op=dist.ReduceOp.AVG)  # the mean inside the collective: no separate division pass

    graphs_on = (not args.no_graphs) and not use_ddp and averager is None
    stepper = harness.GraphedTrainStep(net, opt, grad_clip=100.0,
                                       reduce_grads=allreduce_grads if (world > 1 and not use_ddp and averager is None) else None,
                                       warm=2 if graphs_on else 10 ** 9, flat_grads=args.allreduce != "coalesced")
    queries_per_step = [0]
    host_enqueue_ms = [0.0]
    per_rank = []

    def load_inputs(from_host):
        """(points, lengths, (features, labels)) of one batch; from_host: H2D copies from pinned memory."""
        np.random.seed(1)  # batch_grid_subsampling draws its grid orientations from np.random
        if from_host:
            return (pts_p.to(dev, non_blocking=True), lens_p.to(dev, non_blocking=True),
                    (feats_p.to(dev, non_blocking=True), labels_p.to(dev, non_blocking=True)))
        return pts_d, lens_d, (feats_d, labels_d)

    def train(pyr, f, y, st=None):
        st = st or stepper
        queries_per_step[0] = sum(t.shape[0] for t in pyr.neighbors + pyr.pools + pyr.upsamples)
        if use_ddp or averager is not None:  # comparison modes: eager only
            batch = SimpleNamespace(points=pyr.points, neighbors=pyr.neighbors, pools=pyr.pools,
                                    upsamples=pyr.upsamples, lengths=pyr.lengths, features=f, labels=y)
            loss = net.loss(model(batch), y)
            opt.zero_grad(set_to_none=True)
            loss.backward()
            if averager is not None:
                averager.finish()
            torch.nn.utils.clip_grad_value_(grad_params, 100.0)  # utils/trainer.py:191-193
            opt.step()
            return loss.detach()
        return st(pyr, f, y)

    def step_eager(st=None):
        """One un-pipelined eager step (per-kernel profiling pass): pyramid, then training, one stream."""
        p, ln, (f, y) = load_inputs(False)
        pyr = pyramid.build_pyramid(p, ln, cfg)
        return (st or stepper).eager(pyr, f, y) if not (use_ddp or averager is not None) else train(pyr, f, y)

    prefetch = None if args.no_prefetch else pyramid.PyramidPrefetcher(cfg, dev)
    loss_pin = torch.zeros(2, dtype=torch.float32).pin_memory()  # read-back slots of the e2e arm
    d2h_bytes = loss_pin[0:1].numel() * loss_pin.element_size()

    def run_steps(from_host, steps, st=None):
        """`steps` steps; with the prefetcher the pyramid of batch i+1 is enqueued on the side stream
        after batch i's training launches, so K steps contain K pyramid builds and K training passes.
        from_host: every step's loss is read back (one step late, so the read never drains the queue)."""
        last, prev_ev, loss = None, None, None
        for i in range(steps):
            if prefetch is None:
                p, ln, (f, y) = load_inputs(from_host)
                pyr = pyramid.build_pyramid(p, ln, cfg)
            else:
                pyr, (f, y) = prefetch.take()
            loss = train(pyr, f, y, st)
            if from_host:
                slot = loss_pin[i % 2:i % 2 + 1]
                slot.copy_(loss.detach().reshape(1), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                if prev_ev is not None:
                    prev_ev[0].synchronize()
                    last = float(prev_ev[1])
                prev_ev = (ev, slot)
            if prefetch is not None:
                prefetch.submit(lambda: load_inputs(from_host))
        if from_host:
            prev_ev[0].synchronize()
            return float(prev_ev[1])
        return loss

    def timed(from_host, steps, warmup, sample_clocks=False, st=None):
        st = st or stepper
        if prefetch is not None:
            if prefetch._pending is not None:
                prefetch.take()  # left over from the previous timed region (other input source)
            prefetch.submit(lambda: load_inputs(from_host))
        # one clock sampler per job, on rank 0; started BEFORE the warm-up steps: nvidia-smi's own start-up (driver
        # queries, ~100 ms) must not land inside the timed region (seen as 16-18 ms/step outliers in 20-step runs)
        sampler = ClockSampler(local) if (sample_clocks and rank == 0) else None
        if sampler is not None:
            time.sleep(0.3)
        run_steps(from_host, warmup, st)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0, r0 = L.mvk_launch_count(), st.replays
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host0 = time.perf_counter()
        last = run_steps(from_host, steps, st)
        if prefetch is not None:  # the pyramid enqueued by the last step belongs to the timed region
            torch.cuda.current_stream().wait_stream(prefetch.stream)
        e1.record()
        host_enqueue_ms[0] = 1e3 * (time.perf_counter() - t_host0) / steps  # host time to ENQUEUE a step
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        ms = e0.elapsed_time(e1)
        # kernels launched in the region: eager launches pass through the C ABI (counted there); a graph replay
        # launches the kernels that were counted once when the graph was captured
        launches = (L.mvk_launch_count() - l0) + (st.replays - r0) * (st.launches_per_step or 0)
        if world > 1:
            mine = torch.tensor([ms, float(n_pts)], device=dev)
            allr = [torch.zeros_like(mine) for _ in range(world)]
            dist.all_gather(allr, mine)
            per_rank[:] = [(round(float(a[0]) / steps, 3), int(a[1])) for a in allr]
            ms = max(float(a[0]) for a in allr)  # the job is as slow as its slowest rank
        return ms, launches, clocks, float(last)

    if graphs_on:
        # set-up, not warm-up: two eager steps (optimiser state, weight-operand registry, kernel attributes) and the
        # capture of the step's graph; the W warm-up steps below are replays like the timed ones
        if prefetch is not None:
            prefetch.submit(lambda: load_inputs(False))
        run_steps(False, 3)
        torch.cuda.synchronize()
    ms, launches, clocks, loss_v = timed(False, args.steps, args.warmup, sample_clocks=True)
    enqueue_ms = host_enqueue_ms[0]
    ranks_info = list(per_rank)
    if args.quick:  # profiling runs (ncu): only the device-resident timed region
        if rank == 0:
            print(json.dumps({"quick": True, "ms_per_step": round(ms / args.steps, 3), "points": n_pts,
                              "gpu_launches": int(launches), "host_enqueue_ms_per_step": round(enqueue_ms, 3),
                              "graphs": bool(graphs_on)}), flush=True)
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return
    ms_e2e, _, _, _ = timed(True, args.steps, max(1, args.warmup // 2))
    enqueue_e2e_ms = host_enqueue_ms[0]

    # ---------------- per-kernel timing inside a timed region (roofline) ----------------
    torch.cuda.synchronize()
    with _lib.profile() as records:
        pe0, pe1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        pe0.record()
        for _ in range(args.steps):
            step_eager()
        pe1.record()
        torch.cuda.synchronize()
    prof_ms = pe0.elapsed_time(pe1)
    pk, pk_kind = peaks()
    hbm_bps, tc_fps = pk["hbm_gbs"] * 1e9, pk["bf16_tflops_sustained"] * 1e12

    def nz(x):
        return 1 if (x is not None and getattr(x, "value", x)) else 0

    def work_of(name, a):
        """(family, shape key, algorithmic bytes, flops) of one C-ABI call -- DESIGN.md section 4."""
        if name == "mvk_kpconv_weighted":
            nq, is64, h, cin, ld = a[1], a[5], a[6], a[8], a[14]
            return "stage_a_fwd", f"[cin={cin}]", nq * (h * ((8 if is64 else 4) + 12 + 4 * cin) + 12 + 4 * ld), 0.0
        if name == "mvk_kpconv_weighted_bwd":
            nq, is64, h, cin, ld = a[1], a[5], a[6], a[7], a[14]
            return "stage_a_bwd", f"[cin={cin}]", nq * (h * ((8 if is64 else 4) + 12 + 4 * cin) + 12 + 4 * ld), 0.0
        if name == "mvk_kpconv_fused":
            nq, is64, h, cin, cout, save = a[1], a[5], a[6], a[8], a[12], nz(a[17])
            kd = 15 * cin
            return ("kpconv_fused_fwd", f"[cin={cin},cout={cout}]",
                    nq * (h * ((8 if is64 else 4) + 12 + 4 * cin) + 12 + 4 * cout + (4 * kd if save else 0)) + 4.0 * kd * cout,
                    2.0 * nq * kd * cout * 3)
        if name == "mvk_gemm_bf16x3":
            M, N, K, nv, terms = a[8], a[9], a[10], a[13], a[14]
            ob = 4 if terms == 3 else 2  # bytes per operand element (hi + lo, or hi only)
            return ("contraction", f"[M={M},N={N},K={K},amn={a[2]},bmn={a[6]},split={a[15]}]",
                    ob * (M * K + K * N) + 4.0 * M * nv, 2.0 * M * N * K * terms)
        if name in ("mvk_col_stats", "mvk_bn_batch_stats"):
            return "bn_stats", f"[{a[1]}x{a[2]}]", 4.0 * a[1] * a[2], 0.0
        if name == "mvk_scale_shift_act":
            return "bn_act_fwd", f"[{a[1]}x{a[2]}]", 4.0 * a[1] * a[2] * (2 + nz(a[6]) + nz(a[11])), 0.0
        if name == "mvk_act_bwd_reduce":
            return "bn_act_bwd_reduce", f"[{a[3]}x{a[4]}]", 4.0 * a[3] * a[4] * (2 + nz(a[8])), 0.0
        if name == "mvk_act_bwd_apply":
            n = a[3] * a[4]
            return "bn_act_bwd_apply", f"[{a[3]}x{a[4]}]", 4.0 * n * (3 + nz(a[8]) + nz(a[20])), 0.0
        if name == "mvk_split_bf16":
            return "operand_split", f"[{a[1]}x{a[2]}]", 8.0 * a[1] * a[2], 0.0
        if name in ("mvk_neighbors_query_capped",):
            nq, ns, width, is64 = a[1], a[3], a[10], a[13]
            return "neighbors", "", nq * 12.0 + ns * 12.0 + nq * width * (8 if is64 else 4), 0.0
        if name in ("mvk_neighbors_count", "mvk_neighbors_fill", "mvk_neighbors_fill_i64"):
            return "neighbors", "", a[1] * 12.0 + a[3] * 12.0, 0.0
        if name in ("mvk_pool", "mvk_pool_bwd"):
            return "pools", "", 0.0, 0.0
        return name.replace("mvk_", ""), "", 0.0, 0.0

    agg, fam = {}, {}
    for name, a, s, e in records:
        d = s.elapsed_time(e)
        family, shape, nbytes, flops = work_of(name, a)
        r = agg.setdefault(name + shape, [0.0, 0, name])
        r[0] += d
        r[1] += 1
        f = fam.setdefault(family, {"ms": 0.0, "calls": 0, "bytes": 0.0, "flops": 0.0, "ideal_ms": 0.0})
        f["ms"] += d
        f["calls"] += 1
        f["bytes"] += nbytes
        f["flops"] += flops
        f["ideal_ms"] += 1e3 * max(nbytes / hbm_bps, flops / tc_fps)
    kernel_ms = sum(f["ms"] for f in fam.values())  # all C-ABI calls of the profiled pass (pyramid included)
    by_entry = {}
    for key, (d, c, name) in agg.items():
        r = by_entry.setdefault(name, [0.0, 0])
        r[0] += d
        r[1] += c

    def roof(family, f):
        t = f["ms"] * 1e-3
        hbm_bound = f["bytes"] / hbm_bps >= f["flops"] / tc_fps
        if hbm_bound:
            ach, peak, unit, kind = f["bytes"] / t / 1e9, pk["hbm_gbs"], "GB/s", pk_kind
        else:
            ach, peak, unit, kind = f["flops"] / t / 1e12, pk["bf16_tflops_sustained"], "TFLOP/s", pk_kind + " sustained bf16"
        return {"kernel": family, "bound": "hbm" if hbm_bound else "tensor", "achieved": round(ach, 1), "peak": peak,
                "peak_kind": kind, "unit": unit, "frac": round(ach / peak, 4), "traffic": None, "launches": f["calls"],
                "avg_launch_us": round(1e3 * f["ms"] / f["calls"], 2),
                "share_of_step": round(f["ms"] / kernel_ms, 4),  # share of the libmvk kernel time of a step
                "frac_of_mixed_roofline": round(f["ideal_ms"] / f["ms"], 4)}

    KERNEL_OF = {"stage_a_fwd": "kp_fwd_fast / kp_fwd_tiny (mvk_kpconv_weighted)", "stage_a_bwd": "kp_bwd_fast (mvk_kpconv_weighted_bwd)",
                 "kpconv_fused_fwd": "kp_fused_fwd (mvk_kpconv_fused: gather + influence + tcgen05 contraction)",
                 "contraction": "gemm_tc_kernel (mvk_gemm_bf16x3)", "bn_stats": "col_stats_kernel (mvk_bn_batch_stats)",
                 "bn_act_fwd": "scale_shift_act_kernel (mvk_scale_shift_act)", "bn_act_bwd_reduce": "act_bwd_reduce_kernel",
                 "bn_act_bwd_apply": "act_bwd_apply_kernel", "operand_split": "split_bf16_vec4 (mvk_split_bf16)",
                 "neighbors": "k_query (mvk_neighbors_query_capped)"}
    rooflines = {k: roof(KERNEL_OF.get(k, k), f) for k, f in fam.items() if f["bytes"] > 0 or f["flops"] > 0}
    # measured DRAM / L2 traffic per launch of each family from the committed ncu captures (profiles/traffic.json)
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        tr = {}
    for k, r in rooflines.items():
        r["algorithmic_bytes_per_launch"] = round(fam[k]["bytes"] / fam[k]["calls"], 1)
        if k in tr:
            r["traffic"] = tr[k].get("dram_bytes_per_launch")
            r["l2_bytes_per_launch"] = tr[k].get("l2_bytes_per_launch")
            r["traffic_source"] = tr[k].get("source")
            if tr[k].get("dram_gbs") is not None:  # ncu-measured DRAM bandwidth of the largest launches of the family
                r["ncu_dram_frac"] = round(tr[k]["dram_gbs"] / pk["hbm_gbs"], 4)
    top = max(rooflines.items(), key=lambda kv: fam[kv[0]]["ms"], default=(None, None))
    roofline = top[1]
    breakdown = {n: {"ms_per_step": round(v[0] / args.steps, 3), "calls_per_step": v[1] // args.steps}
                 for n, v in sorted(by_entry.items(), key=lambda kv: -kv[1][0])}
    nb_ms = sum(v[0] for n, v in by_entry.items() if n.startswith("mvk_neighbors"))
    nb_qps = queries_per_step[0] * args.steps / (nb_ms * 1e-3) if nb_ms > 0 else None

    # ---------------- the same step with the strict fp32 contraction (N = 1) ----------------
    fp32_line = None
    if world == 1 and args.contraction != "fp32" and not args.no_extras:
        del stepper.graphs
        stepper.graphs = {}
        net32, opt32 = build_model("fp32")
        st32 = harness.GraphedTrainStep(net32, opt32, grad_clip=100.0, warm=2 if graphs_on else 10 ** 9)
        k32 = max(3, args.steps // 2)
        if prefetch is not None and prefetch._pending is None:
            prefetch.submit(lambda: load_inputs(False))
        run_steps(False, 3, st32)  # set-up: eager warm steps + capture
        ms32, launches32, _, loss32 = timed(False, k32, 3, st=st32)
        fp32_line = {"contraction": "fp32", "dtype": "f32", "value": round(n_pts * k32 / (ms32 * 1e-3), 1), "unit": UNIT,
                     "ms_per_step": round(ms32 / k32, 3), "steps": k32, "gpu_launches": int(launches32), "loss": loss32,
                     "what": "same workload, strict fp32 FFMA contraction (csrc/gemm_simt.cu): the path with network-level "
                             "1e-4 parity (tests/test_gpu_network.py)"}
        del st32, net32, opt32
        torch.cuda.empty_cache()

    total_pts = n_pts  # weak scaling: every rank holds its own 8 spheres (sizes differ slightly by seed)
    if world > 1:
        t = torch.tensor([float(n_pts)], device=dev)
        dist.all_reduce(t)
        total_pts = int(t.item())
    extras = {}
    if world == 1 and not args.no_extras:
        extras["config0"] = bench_config0(dev, pk, pk_kind, cpu=not args.no_cpu_baseline)
        extras["neighbors"] = bench_neighbors(dev, pts_h, lens_h, pk, cpu=not args.no_cpu_baseline)
        stepper.graphs = {}
        torch.cuda.empty_cache()
        try:
            r2 = run_fusion(args, "early", steps=max(3, args.steps // 2), warmup=3, emit=False)
            extras["config2"] = {k: r2[k] for k in ("metric", "value", "unit", "ms_per_step", "e2e", "lifting_ms", "gpu_launches", "config")}
            r4 = run_scene(args, emit=False)
            extras["config4"] = {k: r4[k] for k in ("metric", "value", "unit", "ms_per_step", "scene_points_per_s",
                                                    "neighbor_queries_per_s", "config")}
        except Exception as e:  # noqa: BLE001  (the headline must not die with an extra)
            extras["config2_4_error"] = repr(e)[:300]
    if rank == 0:
        h2d = pts_p.numel() * 4 + feats_p.numel() * 4 + labels_p.numel() * 8 + lens_p.numel() * 4
        dtype = {"bf16x3": "bf16x3", "bf16": "bf16", "fp32": "f32"}[args.contraction]
        line = {
            "metric": METRIC, "value": round(total_pts * args.steps / (ms * 1e-3), 1), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": dtype,
            "dtype_note": {"bf16x3": "contractions on tcgen05 with bf16 hi/lo split operands (3 MMA terms, fp32 accumulate): "
                                     "per-operator error ~5e-6 of max|ref| (bar 1e-4); everything else fp32",
                           "bf16": "plain bf16 tensor-core contractions (stated separately: 2e-2)",
                           "fp32": "strict fp32 FFMA contractions"}[args.contraction],
            "contraction": args.contraction, "data": "synthetic",
            "config": workload_config(n_pts, cfg.neighborhood_limits),
            "impl_notes": {
                "parallelism": (f"sphere-sharded x{world}, gradient all-reduce (NCCL, {args.allreduce})" if world > 1 else "single GPU"),
                "pipeline": ("one stream: pyramid, forward, backward, SGD in sequence" if args.no_prefetch else
                             "pyramid of batch i+1 built on a side stream while batch i trains "
                             "(K timed steps = K pyramids + K training passes)"),
                "launch": ("training step replayed as a CUDA graph (captured once per pyramid shape signature); pyramid eager"
                           if graphs_on else "eager launches")},
            "e2e": {"value": round(total_pts * args.steps / (ms_e2e * 1e-3), 1), "unit": UNIT,
                    "ms_per_step": round(ms_e2e / args.steps, 3), "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h_bytes), "host_enqueue_ms_per_step": round(enqueue_e2e_ms, 3)},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "rooflines": rooflines,
            "rooflines_note": "per C-ABI call, CUDA events on the launching stream, in an eager un-pipelined pass of the same step",
            "neighbor_queries_per_s": round(nb_qps, 1) if nb_qps else None,
            "neighbor_queries_per_step": int(queries_per_step[0]),
            "breakdown_ms": breakdown, "loss": loss_v, "host_enqueue_ms_per_step": round(enqueue_ms, 3),
        }
        if fp32_line:
            line["fp32_contraction"] = fp32_line
        line.update(extras)
        if ranks_info:
            line["per_rank_ms_and_points"] = ranks_info
        if args.detail:
            det = sorted(agg.items(), key=lambda kv: -kv[1][0])[:args.detail]
            line["detail_us_per_call"] = {k: [round(1e3 * v[0] / v[1], 1), v[1] // args.steps] for k, v in det}
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_reference(steps=1, warmup=1, seed=0)
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


# =================================================================================================
# Fusion workloads (BASELINE configs[2], configs[3]) and the whole-scene sweep (configs[4])
# =================================================================================================
FUSION_VIEWS = {"early": 3, "middle": 5, "late": 5}


def run_fusion(args, fusion, steps=None, warmup=None, emit=True):
    """MV-KPConv fusion step on one batch of 8 synthetic spheres with nv synthetic 160x120 views each (early: 3 views,
    configs[2]; middle / late: 5 views, configs[3]): frozen UNet-ResNet34 on the B*nv images, batched lifting
    (unprojection + per-sphere grid 3-NN + FeatureAggregation over the whole stacked batch), fusion KPFCNN forward +
    loss + backward + SGD.  N > 1: sphere-sharded like the baseline, gradient all-reduce after backward."""
    import torch.distributed as dist
    import mvkpconv_b200 as mvk
    from mvkpconv_b200 import _lib, fusion as fu, harness, pyramid, synthetic
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    steps, warmup = steps or args.steps, warmup or args.warmup
    nv, h, w = FUSION_VIEWS[fusion], 120, 160
    L = _lib.lib()
    sub = lambda p, dl: mvk.grid_subsampling(p, sampleDl=dl)
    sph = synthetic.make_fusion_spheres(SPHERES_PER_GPU, sub, seed=0, in_radius=IN_RADIUS, first_dl=FIRST_DL,
                                        n_views=nv, h=h, w=w)
    sph = sph[rank % len(sph):] + sph[:rank % len(sph)]  # same work on every rank (see make_batch)
    pts_h = np.concatenate([s.points for s in sph], 0)
    world_h = np.concatenate([s.world for s in sph], 0)
    lens_h = np.array([len(s.points) for s in sph], np.int32)
    n_pts = len(pts_h)
    rng = np.random.default_rng(rank)
    labels_h = rng.integers(0, 20, n_pts).astype(np.int64)
    images_h = rng.normal(0, 1, (SPHERES_PER_GPU, nv, 3, h, w)).astype(np.float32)
    depths_h = np.stack([s.depths for s in sph], 0)
    poses_h = np.stack([s.poses for s in sph], 0)
    cams_h = np.stack([s.cam for s in sph], 0)
    f3d_h = (np.concatenate([np.ones((n_pts, 1), np.float32), world_h[:, 2:3]], 1) if fusion == "early" else
             np.concatenate([np.ones((n_pts, 1), np.float32), world_h], 1)).astype(np.float32)
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    host = {k: pin(v) for k, v in dict(pts=pts_h, world=world_h, lens=lens_h, labels=labels_h, images=images_h,
                                       depths=depths_h, poses=poses_h, f3d=f3d_h).items()}
    resident = {k: v.to(dev) for k, v in host.items()}
    h2d = sum(v.numel() * v.element_size() for v in host.values())

    cfg = fu.fusion_config(fusion, in_radius=IN_RADIUS, first_subsampling_dl=FIRST_DL)
    cfg.neighborhood_limits = pyramid.calibrate_neighborhood_limits(resident["pts"], resident["lens"], cfg)
    np.random.seed(0)
    torch.manual_seed(0)
    net = fu.FusionKPFCNN(cfg, fusion=fusion, precision_2d=args.precision_2d).to(dev)
    set_contraction(net, args.contraction)
    if world > 1:
        with torch.no_grad():
            for t in list(net.parameters()) + list(net.buffers()):
                dist.broadcast(t, 0)
    params = [p for p in net.parameters() if p.requires_grad]
    opt = torch.optim.SGD(params, lr=1e-2, momentum=0.98, weight_decay=1e-3, fused=True)

    def allreduce_grads(grads):
        with dist._coalescing_manager(device=dev, async_ops=False):
            for g in grads:
                dist.all_reduce(g, op=dist.ReduceOp.AVG)

    graphs_on = not args.no_graphs
    stepper = harness.GraphedTrainStep(net, opt, grad_clip=100.0, reduce_grads=allreduce_grads if world > 1 else None,
                                       warm=2 if graphs_on else 10 ** 9)
    from mvkpconv_b200 import lifting
    kinv = lifting.intrinsics_inverse(cams_h, SPHERES_PER_GPU, nv, dev)  # fixed intrinsics: inverted once

    def one_step(from_host):
        np.random.seed(1)
        src = {k: v.to(dev, non_blocking=True) for k, v in host.items()} if from_host else resident
        pyr = pyramid.build_pyramid(src["pts"], src["lens"], cfg)
        # the numeric half of get_rgbd_data for the whole batch: unprojection + 3-NN on the GPU
        xyz32, _, knn = fu.prepare_lifting(cams_h, src["depths"], src["poses"], src["world"], src["lens"], kinv=kinv)
        extras = dict(images=src["images"], image_xyz=xyz32, knn_global=knn, feat_aggre_points=src["world"],
                      feature_3d=src["f3d"])
        return stepper(pyr, src["f3d"], src["labels"], extras)

    def timed(from_host, k, wu):
        for _ in range(wu):
            one_step(from_host)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        l0, r0 = L.mvk_launch_count(), stepper.replays
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        last = None
        for _ in range(k):
            loss = one_step(from_host)
            if from_host:
                last = loss.item()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = e0.elapsed_time(e1)
        launches = (L.mvk_launch_count() - l0) + (stepper.replays - r0) * (stepper.launches_per_step or 0)
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches, float(loss if last is None else last)

    for _ in range(3):  # set-up: eager warm steps + capture
        one_step(False)
    sampler = ClockSampler(local) if (emit and rank == 0) else None
    if sampler is not None:
        time.sleep(0.3)  # nvidia-smi start-up outside the timed region
        one_step(False)
    ms, launches, loss_v = timed(False, steps, warmup)
    clocks = sampler.stop() if sampler else None
    ms_e2e, _, _ = timed(True, steps, max(1, warmup // 2))
    # lifting alone (unprojection + kNN + 2D net + FeatureAggregation), resident
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    net.eval()
    e0.record()
    for _ in range(5):
        xyz32, _, knn = fu.prepare_lifting(cams_h, resident["depths"], resident["poses"], resident["world"], resident["lens"],
                                           kinv=kinv)
    e1.record()
    torch.cuda.synchronize()
    ms_prep = e0.elapsed_time(e1) / 5
    b = SimpleNamespace(images=resident["images"], image_xyz=xyz32, knn_global=knn, feat_aggre_points=resident["world"])
    with torch.no_grad():
        net.lift(b)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            net.lift(b)
        e1.record()
        torch.cuda.synchronize()
    ms_lift = e0.elapsed_time(e1) / 5
    net.train()
    total_pts = n_pts
    if world > 1:
        t = torch.tensor([float(n_pts)], device=dev)
        dist.all_reduce(t)
        total_pts = int(t.item())
    res = {
        "metric": "MV-KPConv %s-fusion fwd+bwd points/s" % fusion, "value": round(total_pts * steps / (ms * 1e-3), 1),
        "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": round(ms / steps, 3),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": {"bf16x3": "bf16x3", "bf16": "bf16", "fp32": "f32"}[args.contraction], "data": "synthetic",
        "config": {"workload": "configs[%d]: MV-KPConv %s fusion, %d synthetic 160x120 views per sphere + UNet-ResNet34 64-d "
                               "features lifted to the sphere points, fwd+bwd+SGD, 8 spheres/GPU" % (2 if fusion == "early" else 3, fusion, nv),
                   "spheres_per_gpu": SPHERES_PER_GPU, "points_per_gpu": n_pts, "views": nv, "image_hw": [h, w],
                   "neighborhood_limits": [int(v) for v in cfg.neighborhood_limits], "precision_2d": args.precision_2d,
                   "l2": "per-step working set > 1 GB; no flush"},
        "e2e": {"value": round(total_pts * steps / (ms_e2e * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms_e2e / steps, 3),
                "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": 4},
        "gpu_launches": int(launches), "clocks": clocks, "loss": loss_v,
        "lifting_ms": {"unproject_plus_knn": round(ms_prep, 3), "unet_plus_feature_aggregation": round(ms_lift, 3),
                       "pixels": int(SPHERES_PER_GPU * nv * h * w), "points": n_pts},
    }
    if emit and world == 1 and not args.no_cpu_baseline:
        from oracle import cpu_bench
        s0 = sph[0]
        feat2d = np.random.default_rng(0).normal(0, 1, (nv, 64, h, w)).astype(np.float32)
        wts = [np.random.default_rng(i).normal(0, 0.1, sh).astype(np.float32) for i, sh in enumerate([(64, 68), (64, 64), (64, 64)])]
        res["cpu_lifting"] = cpu_bench.lifting_cpu(s0.cam, s0.depths, s0.poses, feat2d, s0.world, wts, iters=1)
    if emit and rank == 0:
        print(json.dumps(res), flush=True)
    if emit and world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return res


def run_scene(args, emit=True):
    """BASELINE configs[4]: inference sweep over a synthetic ~200k-point scene (dl = 0.04): a fixed lattice of
    overlapping r = 2 m spheres, per batch of spheres the 5-level pyramid (grid subsampling + radius search) and the
    KPConv stack (eval mode), votes accumulated per point; spheres sharded round-robin over the ranks, ONE all-reduce
    of the vote table at the end.  Strong scaling: the scene is fixed."""
    import torch.distributed as dist
    import mvkpconv_b200 as mvk
    from mvkpconv_b200 import harness, pyramid, scene, synthetic
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and not dist.is_initialized():
        sys.stdout.flush()
        saved_fd = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved_fd, 1)
            os.close(saved_fd)
    room = synthetic.make_room(seed=7, density=9000.0, size=(12.0, 8.0, 2.6), n_boxes=16)
    pts = mvk.grid_subsampling(room, sampleDl=FIRST_DL)
    cfg = pyramid.baseline_config(in_radius=IN_RADIUS, first_subsampling_dl=FIRST_DL)
    centers = scene.sphere_centers(pts, IN_RADIUS, spacing=IN_RADIUS, z=1.0)
    np.random.seed(0)
    torch.manual_seed(0)
    net = harness.KPFCNN(cfg).to(dev)
    set_contraction(net, args.contraction)
    # neighbourhood limits from the first batch of spheres (the reference calibrates them on its sampler)
    sw0 = scene.SceneSweep(net, cfg, spheres_per_batch=SPHERES_PER_GPU)
    scene_t = torch.from_numpy(pts).to(dev)
    _, centred, lengths = sw0.crop(scene_t, torch.from_numpy(centers[:SPHERES_PER_GPU]).to(dev))
    lengths = lengths[lengths > 0]
    cfg.neighborhood_limits = pyramid.calibrate_neighborhood_limits(centred, lengths, cfg)
    sweep = scene.SceneSweep(net, cfg, spheres_per_batch=SPHERES_PER_GPU)
    steps, warmup = max(1, args.steps // 2), 2
    for _ in range(warmup):
        sweep.run(scene_t, centers, rank, world)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sweep.stats = SimpleNamespace(batches=0, spheres=0, points=0, queries=0)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        probs, counts = sweep.run(scene_t, centers, rank, world)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    st = torch.tensor([ms, float(sweep.stats.points), float(sweep.stats.queries), float(sweep.stats.spheres)], device=dev)
    if world > 1:
        mx = st.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(st)
        ms = float(mx[0])
    sphere_pts, queries, spheres = float(st[1]) / steps, float(st[2]) / steps, float(st[3]) / steps
    res = {"metric": "whole-scene sweep sphere-points/s", "value": round(sphere_pts * steps / (ms * 1e-3), 1), "unit": UNIT,
           "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": round(ms / steps, 3), "higher_is_better": True,
           "scaling": "strong", "vs_baseline": None, "dtype": {"bf16x3": "bf16x3", "bf16": "bf16", "fp32": "f32"}[args.contraction],
           "data": "synthetic",
           "config": {"workload": "configs[4]: whole-scene inference sweep, synthetic 12x8 m scene at dl=0.04, lattice of r=2 m "
                                  "spheres, multi-level grid subsampling + radius search + KPConv stack (eval), sphere-sharded",
                      "scene_points": int(len(pts)), "spheres": int(len(centers)), "sphere_points_per_sweep": int(sphere_pts),
                      "neighbor_queries_per_sweep": int(queries), "covered": float((counts > 0).float().mean()),
                      "l2": "every batch of spheres streams > 1 GB of operands; no flush"},
           "scene_points_per_s": round(len(pts) * steps / (ms * 1e-3), 1),
           "neighbor_queries_per_s": round(queries * steps / (ms * 1e-3), 1), "gpu_launches": None}
    if emit and rank == 0:
        print(json.dumps(res), flush=True)
    if emit and world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return res


# =================================================================================================
# BASELINE configs[0] and the queries/s half of the metric (N = 1 extras of the same JSON line)
# =================================================================================================
def _time_rotating(fn, sets, iters):
    """Average ms of fn(set) over `iters` calls cycling through `sets` (distinct input/output buffers whose combined
    footprint exceeds the 126 MB L2: every call starts with cold caches without a flush kernel in the timing)."""
    for s in sets:
        fn(s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        fn(sets[i % len(sets)])
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def bench_config0(dev, pk, pk_kind, cpu=True):
    """BASELINE configs[0]: ONE rigid KPConv layer 64 -> 128 on ONE synthetic sphere (r = 2 m, dl = 0.04, ~20-40k points,
    K = 15, conv radius 0.10, rows cropped to the p90 width like neighborhood_limits), forward and forward+backward.
    Roofline fractions per SURVEY section 8(d): stage A against HBM with B_gi = H (idx + 12 + 4 Cin) + 12 bytes per point
    (+ 4*15*Cin written when the weighted operand is staged), contraction against the tensor pipe with
    F_c = 2*15*Cin*Cout flop per point (x3 MMA terms executed for bf16x3)."""
    import mvkpconv_b200 as mvk
    from mvkpconv_b200 import _lib, synthetic
    from mvkpconv_b200._lib import check, ptr, stream_ptr
    L = _lib.lib()
    sub = lambda p, dl: mvk.grid_subsampling(p, sampleDl=dl)
    pts = synthetic.make_spheres(1, sub, seed=0, in_radius=IN_RADIUS, first_dl=FIRST_DL)[0]
    n = len(pts)
    lens = np.array([n], np.int32)
    r, cin, cout, K = 2.5 * FIRST_DL, 64, 128, 15
    extent = r * 1.2 / 2.5
    p_d, l_d = torch.from_numpy(pts).to(dev), torch.from_numpy(lens).to(dev)
    full, counts = mvk.batch_neighbors(p_d, p_d, l_d, l_d, r, return_counts=True)
    H = max(1, int(torch.quantile(counts.float().cpu(), 0.9).item()))
    inds = full[:, :H].contiguous().long()
    np.random.seed(0)
    torch.manual_seed(0)
    conv = mvk.KPConv(K, 3, cin, cout, extent, r).to(dev)
    nsets = 6  # 6 x (x 5-10 MB + saved operand ~75-150 MB) >> L2
    sets = [SimpleNamespace(x=torch.randn(n, cin, device=dev), go=torch.randn(n, cout, device=dev)) for _ in range(nsets)]

    def fwd(s):
        with torch.no_grad():
            return conv(p_d, p_d, inds, s.x)

    def fwd_bwd(s):
        x = s.x.requires_grad_(True)
        conv.weights.grad = None
        x.grad = None
        conv(p_d, p_d, inds, x).backward(s.go)

    ms_f = _time_rotating(fwd, sets, 30)
    ms_fb = _time_rotating(fwd_bwd, sets, 30)
    # per-kernel durations (CUDA events around each C-ABI call, eager)
    with _lib.profile() as rec:
        for i in range(12):
            fwd_bwd(sets[i % nsets])
        torch.cuda.synchronize()
    per = {}
    for name, a, s, e in rec:
        per.setdefault(name, []).append(s.elapsed_time(e))
    med = {k: float(np.median(v)) for k, v in per.items()}
    kd = K * cin
    b_gi = H * (8 + 12 + 4 * cin) + 12
    hbm, tc = pk["hbm_gbs"] * 1e9, pk["bf16_tflops_sustained"] * 1e12
    out = {"workload": "configs[0]: single rigid KPConv layer 64->128 on one synthetic ScanNet sphere (r=2 m, dl=0.04, K=15)",
           "points": n, "H": H, "H_uncropped": int(full.shape[1]), "contraction": conv.contraction,
           "fwd": {"points_per_s": round(n / (ms_f * 1e-3), 1), "ms": round(ms_f, 4)},
           "fwd_bwd": {"points_per_s": round(n / (ms_fb * 1e-3), 1), "ms": round(ms_fb, 4)},
           "l2": f"{nsets} rotating input/output sets (footprint > 126 MB L2)",
           "kernel_us": {k: round(1e3 * v, 2) for k, v in med.items()}, "peak_kind": pk_kind}
    if "mvk_kpconv_fused" in med:
        t = med["mvk_kpconv_fused"] * 1e-3
        out["fused_fwd"] = {"B_gi_bytes_per_point": b_gi + 4 * cout, "hbm_frac": round(n * (b_gi + 4 * cout) / t / hbm, 4),
                            "F_c_flop_per_point": 2 * kd * cout, "tensor_frac_useful": round(n * 2.0 * kd * cout / t / tc, 4),
                            "tensor_frac_executed": round(n * 6.0 * kd * cout / t / tc, 4)}
    if "mvk_kpconv_weighted" in med:
        t = med["mvk_kpconv_weighted"] * 1e-3
        out["stage_a_fwd"] = {"B_gi_bytes_per_point": b_gi, "staged_bytes_per_point": 4 * kd,
                              "hbm_frac_B_gi": round(n * b_gi / t / hbm, 4),
                              "hbm_frac_with_staged_write": round(n * (b_gi + 4 * kd) / t / hbm, 4)}
    if "mvk_kpconv_weighted_bwd" in med:
        t = med["mvk_kpconv_weighted_bwd"] * 1e-3
        out["stage_a_bwd"] = {"hbm_frac_B_gi": round(n * b_gi / t / hbm, 4),
                              "hbm_frac_with_staged_read": round(n * (b_gi + 4 * kd) / t / hbm, 4)}
    if "mvk_gemm_bf16x3" in per:
        # forward A.W, dW = A^T dOut, dA = dOut W^T: three launches per fwd+bwd, in call order
        calls = [(a, s.elapsed_time(e)) for name, a, s, e in rec if name == "mvk_gemm_bf16x3"]
        per_step = len(calls) // 12
        names = ["fwd A.W", "dW = A^T.dOut", "dA = dOut.W^T"] if per_step == 3 else ["dW = A^T.dOut", "dA = dOut.W^T"]
        gem = {}
        for j, nm in enumerate(names):
            ts = [calls[i * per_step + j][1] for i in range(12)]
            t = float(np.median(ts)) * 1e-3
            gem[nm] = {"us": round(t * 1e6, 2), "tensor_frac_useful": round(n * 2.0 * kd * cout / t / tc, 4),
                       "tensor_frac_executed_3_terms": round(n * 6.0 * kd * cout / t / tc, 4),
                       "hbm_frac_operands": round((4.0 * n * kd + 4.0 * kd * cout + 4.0 * n * cout) / t / hbm, 4)}
        out["contraction"] = gem
    if cpu:
        from oracle import cpu_bench
        out["cpu"] = cpu_bench.kpconv_layer_cpu(pts, inds.cpu().numpy(), sets[0].x.detach().cpu().numpy(),
                                                conv.kernel_points.detach().cpu().numpy(),
                                                conv.weights.detach().cpu().numpy(), extent)
    return out


def bench_neighbors(dev, pts_h, lens_h, pk, cpu=True):
    """Radius-neighbour queries/s on the stacked batch (8 spheres, ~267k queries = supports, r = 0.10: the level-0
    conv-neighbour call of the pyramid) through the reference-facing `batch_neighbors` (count + fill, uncapped rows:
    the reference's semantics), with inputs resident in HBM and end to end from host numpy arrays; the unmodified
    reference C++ (nanoflann) is timed beside it, one thread per call and P worker processes."""
    import mvkpconv_b200 as mvk
    r = 2.5 * FIRST_DL
    p_d, l_d = torch.from_numpy(pts_h).to(dev), torch.from_numpy(lens_h).to(dev)
    n = len(pts_h)
    w = [0]

    def resident():
        w[0] = mvk.batch_neighbors(p_d, p_d, l_d, l_d, r).shape[1]

    def host():
        mvk.batch_neighbors(pts_h, pts_h, lens_h, lens_h, r)

    def wall(fn, iters):
        fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / iters

    t_res, t_host = wall(resident, 10), wall(host, 5)
    alg = 12.0 * n + 12.0 * n + 4.0 * n * w[0]
    out = {"queries": n, "radius": r, "row_width": int(w[0]),
           "resident": {"queries_per_s": round(n / t_res, 1), "ms": round(1e3 * t_res, 3),
                        "hbm_frac_algorithmic": round(alg / t_res / (pk["hbm_gbs"] * 1e9), 4)},
           "e2e_host_arrays": {"queries_per_s": round(n / t_host, 1), "ms": round(1e3 * t_host, 3),
                               "what": "numpy in -> H2D -> count + fill -> D2H numpy out (the reference's calling convention)"}}
    if cpu:
        import tempfile
        with tempfile.TemporaryDirectory() as td:
            path = os.path.join(td, "nbr.npz")
            np.savez(path, points=pts_h, lengths=lens_h, radius=np.float32(r))
            res = subprocess.run([sys.executable, "-m", "oracle.cpu_bench", "neighbors", path], cwd=ROOT,
                                 capture_output=True, text=True, timeout=600)
        try:
            out["cpu"] = json.loads(res.stdout.strip().splitlines()[-1])
            out["speedup_vs_cpu_single_thread"] = round(out["e2e_host_arrays"]["queries_per_s"] /
                                                        out["cpu"]["single_thread"]["queries_per_s"], 1)
            out["speedup_vs_cpu_worker_processes"] = round(out["e2e_host_arrays"]["queries_per_s"] /
                                                           out["cpu"]["worker_processes"]["queries_per_s"], 1)
        except Exception as e:  # noqa: BLE001
            out["cpu"] = {"error": repr(e), "stderr": res.stderr[-300:]}
    return out


# =================================================================================================
# CPU reference arm / cpu_baseline leg.  The ONLY code in this file that touches oracle/.
# =================================================================================================
def cpu_reference(steps, warmup, seed=0, n_spheres=SPHERES_PER_GPU):
    """The reference's CPU path on the SAME workload as the B200 arm (the batch of 8 seeded spheres of rank 0):
    pyramid by the unmodified reference C++ (oracle/_ref, one thread like the reference's workers), network
    forward + backward + SGD on torch CPU with all host threads through the oracle's restatement of KPConv."""
    from mvkpconv_b200 import harness, pyramid, synthetic
    from oracle import geom, modules

    have_ref = geom.have_ref()
    nbr = geom.ref_batch_neighbors if have_ref else geom.batch_neighbors
    gsub = geom.ref_grid_subsample_batch if have_ref else geom.grid_subsample_batch
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    sub = lambda p, dl: gsub(p, np.array([len(p)], np.int32), sampleDl=dl)[0]
    spheres = synthetic.make_spheres(n_spheres, sub, seed=seed, in_radius=IN_RADIUS, first_dl=FIRST_DL)
    pts, lens = synthetic.stack(spheres)
    cfg = pyramid.baseline_config(in_radius=IN_RADIUS, first_subsampling_dl=FIRST_DL)

    def cpu_subsample(points, lengths, sampleDl=0.1, random_grid_orient=True):
        return gsub(points, lengths, sampleDl=sampleDl)

    gops = SimpleNamespace(batch_neighbors=nbr, batch_grid_subsampling=cpu_subsample)
    cfg.neighborhood_limits = pyramid.calibrate_neighborhood_limits(pts, lens, cfg, ops=gops)
    np.random.seed(0)
    torch.manual_seed(0)
    mops = SimpleNamespace(KPConv=modules.KPConvOracle, max_pool=modules.max_pool, closest_pool=modules.closest_pool)
    net = harness.KPFCNN(cfg, ops=mops)
    opt = torch.optim.SGD(net.parameters(), lr=1e-2, momentum=0.98, weight_decay=1e-3)
    feats = torch.from_numpy(host_features(pts))
    labels = torch.from_numpy(np.random.default_rng(0).integers(0, 20, len(pts)).astype(np.int64))
    t_pyr = t_net = 0.0

    def step():
        nonlocal t_pyr, t_net
        t0 = time.perf_counter()
        pyr = pyramid.build_pyramid(pts, lens, cfg, ops=gops, random_grid_orient=False)
        t1 = time.perf_counter()
        as_t = lambda lst, dt: [torch.from_numpy(np.ascontiguousarray(a)).to(dt) for a in lst]
        batch = SimpleNamespace(points=as_t(pyr.points, torch.float32), neighbors=as_t(pyr.neighbors, torch.int64),
                                pools=as_t(pyr.pools, torch.int64), upsamples=as_t(pyr.upsamples, torch.int64),
                                lengths=pyr.lengths, features=feats, labels=labels)
        out = net(batch)
        loss = net.loss(out, labels)
        opt.zero_grad(set_to_none=True)
        loss.backward()
        torch.nn.utils.clip_grad_value_(net.parameters(), 100.0)
        opt.step()
        t2 = time.perf_counter()
        t_pyr += t1 - t0
        t_net += t2 - t1
        return float(loss.detach())

    for _ in range(warmup):
        step()
    t_pyr = t_net = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = time.perf_counter() - t0
    return {"value": round(len(pts) * steps / dt, 1), "unit": UNIT, "cores": cores, "kind": "port",
            "kind_detail": ("pyramid (radius neighbours, grid subsampling) = the UNMODIFIED reference C++ compiled from "
                            "/root/reference (oracle/_ref/libref.so); " if have_ref else "pyramid = the plain-C restatement; ")
                           + "network = torch-CPU restatement of the reference graph (oracle.modules.KPConvOracle + torch "
                             "Linear / BatchNorm1d / LeakyReLU; reproduces the reference's own KPFCNN class on a golden "
                             "batch, tests/test_fusion_golden.py), not the reference's Python files (absent on the GPU box)",
            "sample": f"the whole batch: {n_spheres} of {SPHERES_PER_GPU} spheres ({len(pts)} points), {steps} step(s) after "
                      f"{warmup} warm-up; pyramid 1 thread (like a reference DataLoader worker), network {cores} threads",
            "ms_per_step": round(1e3 * dt / steps, 1), "ms_pyramid": round(1e3 * t_pyr / steps, 1),
            "ms_network": round(1e3 * t_net / steps, 1), "points": int(len(pts)),
            "neighborhood_limits": [int(v) for v in cfg.neighborhood_limits]}


def run_reference(args):
    rank = int(os.environ.get("RANK", 0))
    if rank != 0:
        return  # rank 0 alone runs the CPU reference; the other ranks exit 0 without work
    world = int(os.environ.get("WORLD_SIZE", args.gpus))
    steps, warmup = args.steps, args.warmup  # one CPU step of the full batch takes ~10 s on 16 cores
    base = cpu_reference(steps=steps, warmup=warmup, seed=0)
    line = {
        "impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": world,
        "steps": steps, "warmup": warmup,
        "ms_per_step": base["ms_per_step"], "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(base["points"], base["neighborhood_limits"]),
        "cpu_baseline": base,
        "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--contraction", default=os.environ.get("MVK_CONTRACTION", "bf16x3"),
                    choices=["bf16x3", "bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graphs", action="store_true", help="launch the training step eagerly instead of replaying a CUDA graph")
    ap.add_argument("--workload", default="baseline", choices=["baseline", "early", "middle", "late", "scene"],
                    help="baseline = BASELINE configs[1] (the headline); early = configs[2]; middle / late = configs[3]; "
                         "scene = configs[4] (whole-scene sweep)")
    ap.add_argument("--precision-2d", default="fp32", choices=["fp32", "bf16"], help="autocast of the frozen 2D network")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the config0 / neighbours / fp32-contraction legs of the N = 1 line")
    ap.add_argument("--allreduce", default="flat", choices=["coalesced", "overlap", "flat", "ddp"],
                    help="N > 1: gradients packed into one buffer + ONE NCCL all-reduce (flat), a grouped all-reduce of the "
                         "per-parameter tensors (coalesced), or torch DDP buckets (ddp)")
    ap.add_argument("--detail", type=int, default=0, help="add the N most expensive (entry point, shape) rows")
    ap.add_argument("--quick", action="store_true", help="device-resident timed region only (for ncu runs)")
    ap.add_argument("--no-prefetch", action="store_true",
                    help="build each batch's pyramid on the training stream right before its forward pass "
                         "(default: pyramid of batch i+1 on a side stream while batch i trains)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        run_reference(args)
    elif args.workload in ("early", "middle", "late"):
        run_fusion(args, args.workload)
    elif args.workload == "scene":
        run_scene(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
