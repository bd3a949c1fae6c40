#!/usr/bin/env python
"""Benchmark of the B200 U-Net hot path (contract: one JSON line on rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload infer|train]

Workload at N=1 (BASELINE.json configs[1], the configuration the metric is quoted on): U-Net inference of ONE synthetic
1920x1080 4-channel G-buffer frame in fp32 mode, Mpix/s = B*H*W / latency.  With N>1 every rank runs its own frame
(frames shard with no collective: weak scaling) and `value` is the sum over ranks / max-over-ranks time.

  value : K steps timed on the device with CUDA events, inputs resident in HBM, L2 flushed before every step
  e2e   : the same step through the reference-facing calls with HOST buffers: the frame pipeline (Unet.open_pipe ->
          nsm_unet_pipe_submit: pinned H2D copy + forward + D2H copy of every frame inside the timed region, copies of
          neighbouring frames overlapping the kernels) and, as e2e.single_call, one synchronous nsm_unet_infer_host per frame
  roofline : the tcgen05 implicit-GEMM kernel, timed live with CUDA events on its launch stream inside the step
  cpu_baseline : the CPU oracle port of the reference (torch CPU ops, all host cores) on a bounded sample
`--impl reference` times that CPU port alone (the reference is pure Python on PyTorch; its tree does not travel to the
GPU box, so the restated port under oracle/ is what runs -- kind "port").
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

# rank 0 prints exactly one JSON line on stdout: keep NCCL's "NCCL version ..." banner (NCCL_DEBUG=VERSION) off it
if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
    os.environ["NCCL_DEBUG"] = "WARN"

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "pcss-unet_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

FLOP_PER_FRAME_1080P = 1550.0e9       # BASELINE.md section 3 (convolutions only, 2*MAC)
H1080, W1080 = 1080, 1920
# one string for both arms (the driver compares config.workload of the two lines)
CFG1_WORKLOAD = "cfg1: U-Net inference, 1 synthetic 1920x1080 G-buffer frame per step, fp32, eval BatchNorm"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "bf16_tflops": d["bf16_tflops"],
                "bf16_tflops_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


def measured_traffic():
    """dram__bytes_read+write of the dominant launch from the latest committed `ncu --set full` capture (profiles/)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_traffic.json")))
    if not files:
        return None
    try:
        d = json.load(open(files[-1]))
        k = next(iter(d))
        return {"kernel_launch": k, "dram_bytes": d[k]["dram_bytes_per_launch"],
                "algorithmic_bytes": d[k]["algorithmic_bytes_per_launch"], "source": os.path.basename(files[-1])}
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [c.strip() for c in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def make_params(seed=42):
    """Weights for the CPU legs only (cpu_baseline / --impl reference): seeded default init + calibrated BN buffers."""
    import oracle
    g = torch.Generator().manual_seed(0)
    P = oracle.init_params(seed)
    oracle.calibrate_bn(P, torch.randn(1, 4, 64, 64, generator=g), generator=g)
    return P


def make_model(precision, dev, seed=42):
    """B200 arm: the drop-in Unet with seeded default init (main.py:73-92) and BN buffers calibrated by one train-mode
    pass of the product path itself (momentum 1.0) -- nothing under oracle/ is touched by the measured arm."""
    from Unetmodel import Unet
    torch.manual_seed(seed)
    net = Unet(precision=precision).to(dev)
    g = torch.Generator().manual_seed(0)
    bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    with torch.no_grad():
        for m in bns:
            m.weight.copy_(torch.empty(m.num_features).uniform_(0.5, 1.5, generator=g))
            m.bias.copy_(torch.empty(m.num_features).uniform_(-0.5, 0.5, generator=g))
            m.momentum = 1.0
        net.train()
        for m in net.modules():
            if isinstance(m, torch.nn.Dropout2d):
                m.p_saved, m.p = m.p, 0.0
        net(torch.randn(1, 4, 64, 64, generator=g).to(dev))
        for m in net.modules():
            if isinstance(m, torch.nn.Dropout2d):
                m.p = m.p_saved
        for m in bns:
            m.momentum = 0.1
    return net.eval()


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the CPU port of the reference path
# ----------------------------------------------------------------------------------------------------------------
def reference_unet():
    """The reference's OWN `Unet` class (baseline/_ref/Unetmodel.py, an unmodified run-time copy of
    /root/reference/Unetmodel.py staged by __graft_entry__.build()), seeded like main.py:73-92, BN buffers calibrated by one
    train-mode pass of the reference itself.  Returns None when baseline/_ref is absent (then the oracle port is timed)."""
    sys.path.insert(0, os.path.join(ROOT, "baseline"))
    try:
        import ref_harness
        if not ref_harness.available():
            return None
        Unet = ref_harness.reference_unet_class()
    except Exception:
        return None
    torch.manual_seed(42)
    net = Unet()
    g = torch.Generator().manual_seed(0)
    bns = [m for m in net.modules() if isinstance(m, torch.nn.BatchNorm2d)]
    with torch.no_grad():
        for m in bns:
            m.weight.copy_(torch.empty(m.num_features).uniform_(0.5, 1.5, generator=g))
            m.bias.copy_(torch.empty(m.num_features).uniform_(-0.5, 0.5, generator=g))
            m.momentum = 1.0
        net.train()
        net(torch.randn(1, 4, 64, 64, generator=g))
        for m in bns:
            m.momentum = 0.1
    return net.eval()


def cpu_forward_timer(H, W, steps, warmup, bf16=False):
    """Seconds per eval forward of one HxW frame on the host cores.  Returns (seconds, kind): kind "reference" = the
    unmodified reference class from baseline/_ref, "port" = the oracle restatement (fallback when _ref is absent)."""
    torch.set_num_threads(os.cpu_count() or 1)
    x = torch.randn(1, 4, H, W, generator=torch.Generator().manual_seed(1))
    net = reference_unet()
    if net is not None:
        kind = "reference"
        fwd = (lambda: net(x)) if not bf16 else None
        if bf16:
            def fwd():
                with torch.autocast("cpu", dtype=torch.bfloat16):
                    return net(x)
    else:
        import oracle
        kind = "port"
        P = make_params()
        fwd = lambda: oracle.unet_forward(x, P, training=False, bf16=bf16)  # noqa: E731
    with torch.no_grad():
        for _ in range(warmup):
            fwd()
        t0 = time.perf_counter()
        for _ in range(steps):
            fwd()
        dt = time.perf_counter() - t0
    return dt / steps, kind


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    full = (args.steps + args.warmup) <= 40
    H, W = (H1080, W1080) if full else (H1080 // 2, W1080)
    sec, kind = cpu_forward_timer(H, W, args.steps, args.warmup)
    mpix = H * W / 1e6 / sec
    what = ("the reference's own Unet class (unmodified copy of Unetmodel.py in baseline/_ref), model.eval(), "
            "torch.no_grad()" if kind == "reference" else "oracle port of Unetmodel.py:90-149 (baseline/_ref absent)")
    sample = (f"{args.steps} timed + {args.warmup} warm-up eval forwards of one {W}x{H} frame, fp32, torch CPU, {what} "
              f"({'full frame' if full else 'half-height frame, Mpix/s is size-normalised'})")
    line = {"impl": "reference", "metric": "U-Net inference Mpix/s @1080p", "value": mpix, "unit": "Mpix/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp32", "data": "synthetic",
            "config": {"workload": CFG1_WORKLOAD},
            "cpu_baseline": {"value": mpix, "unit": "Mpix/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": mpix, "unit": "Mpix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------------------------
# B200 arm
# ----------------------------------------------------------------------------------------------------------------
def run_b200(args):
    import torch.distributed as dist
    import nsm
    from Unetmodel import Unet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    nsm.require_device()
    precision = args.precision
    B, H, W = args.batch, args.height, args.width

    net = make_model(precision, dev)
    g = torch.Generator().manual_seed(100 + rank)
    x_host = torch.randn(B, 4, H, W, generator=g).pin_memory()
    y_host = torch.empty(B, 1, H - H % 2, W - W % 2).pin_memory()
    x = x_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # 256 MiB > 126 MB L2

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.inference_mode():
        for _ in range(max(args.warmup, 3)):
            net(x)
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        # ---- device-resident timing -------------------------------------------------------------------------
        evs = []
        barrier()
        launches0 = nsm.launch_count()
        for _ in range(args.steps):
            flush.zero_()
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            net(x)
            e.record()
            evs.append((s, e))
        barrier()
        launches = nsm.launch_count() - launches0
        dev_ms = sum(s.elapsed_time(e) for s, e in evs)
        # ---- end-to-end timing through the host-buffer entry points --------------------------------------------
        # (a) one synchronous call per frame: H2D + forward + D2H + stream sync inside nsm_unet_infer_host
        for _ in range(2):
            net.infer_host(x_host, y_host)
        barrier()
        t_single = 0.0
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            net.infer_host(x_host, y_host)
            t_single += time.perf_counter() - t0
        # (b) the frame pipeline (nsm_unet_pipe_*): every frame is still copied in from pinned memory and its result
        #     copied back inside the timed region, but the copies of frames k+1 / k-1 overlap the kernels of frame k.
        #     No L2 flush between frames: one frame's intermediates (~1.5 GB) already exceed the 126 MB L2.
        pipe = net.open_pipe(B, H, W)
        xs_host = [x_host] + [torch.randn(B, 4, H, W, generator=torch.Generator().manual_seed(7 + k)).pin_memory()
                              for k in range(2)]
        ys_host = [torch.empty_like(y_host).pin_memory() for _ in range(3)]
        for k in range(3):
            pipe.submit(xs_host[k], ys_host[k])
        pipe.sync()
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            pipe.submit(xs_host[k % 3], ys_host[k % 3])
        pipe.sync()
        t_e2e = time.perf_counter() - t0
        if not torch.equal(ys_host[0], y_host):
            raise RuntimeError("frame pipeline result differs from the synchronous host call")
        pipe.close()
        barrier()
        clocks = sampler.stop()
        # ---- per-kernel timing (CUDA events on the launch stream, same step) ---------------------------------
        nsm.profile_enable(True)
        for _ in range(args.steps):
            flush.zero_()
            net(x)
        torch.cuda.synchronize()
        rows = nsm.profile_read()
        nsm.profile_enable(False)

    t = torch.tensor([dev_ms, t_e2e * 1e3, t_single * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, single_ms = t.tolist()
    # stock PyTorch / cuDNN on the same GPU, same parameters, same frame (before the model is released for training)
    stock = None
    if world == 1 and not args.no_stock:
        stock = stock_pytorch_baseline(net, x)
    # second headline of BASELINE.json ("train samples/s @1/2/4/8 B200"): a short training measurement attached to the
    # same line (all ranks take part: data-parallel step with NCCL gradient all-reduce)
    train = None
    if not args.no_train:
        import copy
        targs = copy.copy(args)
        targs.steps, targs.warmup = min(args.steps, 10), 3
        try:
            del net, x
            torch.cuda.empty_cache()
            full = {}
            for tag, no_pert in (("customLoss", True), ("customLoss+pert_loss", False)):
                targs.no_perturb = no_pert
                tl = measure_train(targs, own_process_group=False)
                if tl is not None:
                    full[tag] = {"value": tl["value"], "unit": tl["unit"], "ms_per_step": tl["ms_per_step"],
                                 "e2e": tl["e2e"]["value"], "stock_pytorch_gpu": tl.get("stock_pytorch_gpu"),
                                 "dp_check": tl.get("dp_check"),
                                 "gemm_tflops": tl["roofline"]["achieved"],
                                 "gemm_frac_of_sustained_peak": tl["roofline"]["frac"],
                                 "workload": tl["config"]["workload"]}
            train = full
        except Exception as ex:   # the inference line must still be printed
            train = {"error": repr(ex)[:300]}
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pix = B * (H - H % 2) * (W - W % 2)
    value = world * pix * args.steps / (dev_ms * 1e-3) / 1e6
    e2e_value = world * pix * args.steps / (e2e_ms * 1e-3) / 1e6

    # dominant kernel: the tcgen05 implicit-GEMM convolution (all its launches inside the step)
    pk = peaks()
    conv = [r for r in rows if r[0].startswith("conv")]
    conv_ms = sum(r[1] for r in conv)
    conv_fl = sum(r[2] for r in conv)
    all_ms = sum(r[1] for r in rows)
    achieved = conv_fl / (conv_ms * 1e-3) / 1e12 if conv_ms > 0 else 0.0
    per_layer = {}
    for name, ms, fl, by in rows:
        d = per_layer.setdefault(name, [0.0, 0.0, 0.0, 0])
        d[0] += ms; d[1] += fl; d[2] += by; d[3] += 1
    # MMA slots (in units of one bf16-rate MMA) per algorithmic MAC: bf16 mode 1; fp32 mode 3 (hi*hi + hi*lo + lo*hi), or 2
    # for the decoder's 3x3 convolutions whose two cross terms are ONE e4m3 MMA at twice the rate (fp16 hi + 8-bit cross)
    # (the fused decoder blocks "convK block (...)" are 95 % 3x3 work in that format)
    x8_names = () if os.environ.get("NSM_NO_X8") else ("conv6.3x3", "conv7.3x3", "conv8.3x3", "conv9.3x3", "conv8 block",
                                                       "conv9 block")
    x8_layers = tuple(k for k in per_layer if k.startswith(x8_names)) if x8_names else ()
    mma_factor = 3 if precision == "fp32" else 1
    issued = 0.0
    layers = {}
    for k, v in per_layer.items():
        ms, fl, by = v[0] / v[3], v[1] / v[3], v[2] / v[3]
        # per-layer roofline: the slower of tensor time (issued MMA work / measured bf16 peak) and HBM time
        f_k = 2 if (precision == "fp32" and k in x8_layers) else mma_factor
        if k.startswith("conv"):
            issued += fl * f_k
        t_alg = fl / (pk["bf16_tflops"] * 1e12) * 1e3          # algorithmic 2*M*K*N FLOPs at the measured bf16 peak
        t_tensor = t_alg * f_k                                  # the MMA slots this mode actually issues
        t_hbm = by / (pk["hbm_gbs"] * 1e9) * 1e3
        bound = "tensor" if t_alg >= t_hbm else "hbm"
        layers[k] = {"ms": ms, "tflops": fl / max(ms, 1e-9) / 1e9, "gbs": by / max(ms, 1e-9) / 1e6, "bound": bound,
                     "roofline_ms": max(t_alg, t_hbm), "frac_of_roofline": max(t_alg, t_hbm) / max(ms, 1e-9),
                     "mma_slots_per_mac": f_k if k.startswith("conv") else None,
                     "frac_of_issued_roofline": max(t_tensor, t_hbm) / max(ms, 1e-9)}
    t_roof = sum(v["roofline_ms"] for v in layers.values())
    planes = 2 if precision == "fp32" else 1
    roofline = {"kernel": f"conv_gemm_kernel<BN,{planes}> + conv_gemm_wide_kernel + upblock_kernel (tcgen05 implicit GEMM; "
                          f"conv8 and conv9..output as fused up-sample/3x3/1x1 blocks; "
                          f"{len(conv) // max(args.steps, 1)} launches/step)",
                "bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_tflops"], "traffic": measured_traffic(),
                "peak_source": f"{pk['source']} cuBLAS bf16 burst (MEASURED_PEAKS.json)",
                "mma_work_factor": mma_factor,
                "mma_work_factor_x8_layers": 2 if (precision == "fp32" and x8_layers) else None,
                "frac_of_issued_mma": issued * args.steps / (conv_ms * 1e-3) / 1e12 / pk["bf16_tflops"]
                if conv_ms > 0 else None,
                "step_roofline_ms": t_roof, "step_frac_of_roofline": t_roof / max(all_ms / max(args.steps, 1), 1e-9),
                "per_layer_note": "frac_of_roofline = slower of (algorithmic FLOPs / measured bf16 peak) and (algorithmic "
                                  "bytes / measured HBM GB/s), divided by the event-timed stage time; "
                                  "frac_of_issued_roofline counts the 2-3 MMA slots per MAC of the fp32 mode instead",
                "note": "achieved = algorithmic 2*M*K*N FLOPs of the conv launches / their CUDA-event time inside "
                        "the step; fp32 mode issues 3 bf16-rate MMA slots per algorithmic MAC (hi*hi+hi*lo+lo*hi), "
                        "2 in the decoder's 3x3 convolutions (cross terms as one e4m3 MMA), so the ceiling of frac "
                        "lies between 1/3 and 1/2; frac_of_issued_mma counts the issued slots",
                "kernel_share_of_step": conv_ms / all_ms if all_ms else None,
                "per_layer": layers}

    cfg_name = ("cfg1" if (B, H, W) == (1, 1080, 1920) else
                f"cfg4 (16 4K frames sharded by frame over {world} GPU(s), {B} per GPU, no collective)"
                if (H, W) == (2160, 3840) and B * world == 16 else
                "cfg4 (per-GPU share: 2 of the 16 4K frames)" if (B, H, W) == (2, 2160, 3840) else "custom shape")
    summary = None
    if isinstance(train, dict) and "error" not in train:
        summary = {k: {"samples_per_s": round(v["value"], 1), "ms_per_step": round(v["ms_per_step"], 3),
                       "e2e_samples_per_s": round(v["e2e"], 1), "n_gpus": world,
                       "dp_check_rel_l2": (v.get("dp_check") or {}).get("allreduced_vs_mean_of_local_rel_l2")}
                   for k, v in train.items()}
    line = {"metric": "U-Net inference Mpix/s @1080p" if H == 1080 else f"U-Net inference Mpix/s @{W}x{H}", "value": value, "unit": "Mpix/s", "n_gpus": world,
            "train_summary": summary,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32 (fp16 hi + fp16-lo / e4m3-cross operand planes, 2-3 MMA slots per MAC, fp32 accumulate; "
                     "output within 1e-4 of the fp32 reference)" if precision == "fp32" else "bf16",
            "data": "synthetic",
            "config": {"workload": CFG1_WORKLOAD if (B, H, W, precision) == (1, 1080, 1920, "fp32") else
                                   f"{cfg_name}: U-Net inference, {B} synthetic {W}x{H} G-buffer frame(s) per GPU "
                                   f"per step, {precision} mode, eval BatchNorm",
                       "frames_per_gpu_per_step": B,
                       "l2": "256 MiB buffer written before every timed step (L2 flush); step working set ~1.5 GB",
                       "sharding": "frames per rank, no collective"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "Mpix/s", "h2d_bytes_per_step": x_host.numel() * 4,
                    "d2h_bytes_per_step": y_host.numel() * 4, "ms_per_step": e2e_ms / args.steps,
                    "path": "Unet.open_pipe -> nsm_unet_pipe_submit per frame (pinned host in/out, copies of "
                            "neighbouring frames overlap the kernels), nsm_unet_pipe_sync at the end",
                    "single_call": {"value": world * pix * args.steps / (single_ms * 1e-3) / 1e6, "unit": "Mpix/s",
                                    "ms_per_step": single_ms / args.steps,
                                    "path": "Unet.infer_host -> nsm_unet_infer_host (H2D + forward + D2H + sync "
                                            "per call, L2 flushed between calls)"}},
            "gpu_launches": launches,
            "roofline": roofline,
            "train": train}
    if stock is not None:
        line["stock_pytorch_gpu"] = stock
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        Hc, Wc = (H, W) if H * W <= H1080 * W1080 else (H1080, W1080)
        sec, kind = cpu_forward_timer(Hc, Wc, 5, 1)
        line["cpu_baseline"] = {"value": Hc * Wc / 1e6 / sec, "unit": "Mpix/s", "cores": cores, "kind": kind,
                                "sample": f"5 timed + 1 warm-up eval forwards of one {Wc}x{Hc} frame, fp32, torch CPU "
                                          f"on {cores} threads, " +
                                          ("the reference's own Unet class (unmodified copy in baseline/_ref)"
                                           if kind == "reference" else "oracle port of Unetmodel.py:90-149")}
    line["summary"] = {"infer_mpix_per_s": round(value, 1), "e2e_mpix_per_s": round(e2e_value, 1), "n_gpus": world,
                       "train": summary}        # repeated at the very end: a truncated tail of the line still holds it
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def stock_pytorch_forward(net, x):
    """The reference's forward graph (Unetmodel.py:90-149) executed by stock PyTorch ops on the drop-in's own parameter
    containers (real nn.Conv2d / nn.BatchNorm2d modules): the cuDNN / ATen baseline on the same GPU."""
    import torch.nn.functional as F

    def match(t, ref):
        return t if t.shape[-2:] == ref.shape[-2:] else F.interpolate(t, size=ref.shape[-2:], mode="bilinear",
                                                                      align_corners=True)
    H, W = x.shape[-2] - x.shape[-2] % 2, x.shape[-1] - x.shape[-1] % 2
    if (H, W) != tuple(x.shape[-2:]):
        x = F.interpolate(x, size=(H, W), mode="bilinear", align_corners=True)
    x = F.pixel_unshuffle(x.float(), 2)
    c2 = net.conv2.conv(x); c3 = net.conv3.conv(F.avg_pool2d(c2, 2)); c4 = net.conv4.conv(F.avg_pool2d(c3, 2))
    c5 = net.conv5.conv(F.avg_pool2d(c4, 2))
    up = lambda t: F.interpolate(t, scale_factor=2, mode="bilinear", align_corners=True)
    c6 = net.conv6.conv(match(up(c5), c4)) + c4
    c7 = net.conv7.conv(match(up(c6), c3)) + c3
    c8 = net.conv8.conv(match(up(c7), c2)) + c2
    c9 = net.conv9.conv(match(up(c8), c2))
    return torch.sigmoid(F.pixel_shuffle(net.conv10(c9), 2))


def stock_pytorch_baseline(net, x, steps=5):
    """ms per frame of stock PyTorch on this GPU: fp32 with TF32 allowed (PyTorch's default conv setting), strict fp32, and
    bf16 autocast with channels_last -- BASELINE.md section 3 item 5.  Measurement only; never on the product path."""
    out = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=x.device)
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    torch.backends.cudnn.benchmark = True
    try:
        for name, tf32, amp in (("fp32_tf32", True, False), ("fp32_strict", False, False), ("bf16_autocast", True, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            m = net.to(memory_format=torch.channels_last)
            xin = x.contiguous(memory_format=torch.channels_last)
            with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=amp):
                for _ in range(3):
                    stock_pytorch_forward(m, xin)
                ms = 0.0
                for _ in range(steps):
                    flush.zero_()
                    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s.record()
                    stock_pytorch_forward(m, xin)
                    e.record()
                    torch.cuda.synchronize()
                    ms += s.elapsed_time(e)
            out[name] = {"ms_per_step": ms / steps, "value": x.shape[0] * x.shape[-2] * x.shape[-1] / (ms / steps) / 1e3,
                         "unit": "Mpix/s"}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = saved
        net.to(memory_format=torch.contiguous_format)
    out["note"] = ("reference forward graph through stock PyTorch ops (cuDNN, channels_last, cudnn.benchmark) on the same "
                   "GPU and parameters; fp32_tf32 = PyTorch's default conv precision (~2e-3 output error), fp32_strict "
                   "= allow_tf32 off, the only stock setting within 1e-4 of the fp32 reference")
    return out


# ----------------------------------------------------------------------------------------------------------------
# training workload (BASELINE.json configs[2]/[3]): batch 32 of 512x512 crops per GPU, bf16, CustomLoss + PerturbationLoss
# ----------------------------------------------------------------------------------------------------------------
TRAIN_FLOP_PER_SAMPLE = 3 * 747680.0 * 512 * 512           # fwd + dgrad + wgrad
PERT_FLOP_PER_SAMPLE = 3 * 747680.0 * 512 * 512            # + 3 no-grad forwards of the perturbation loss


def stock_pytorch_train_baseline(B, H, W, dev, use_pert, steps=3):
    """ms per training step of the same objective through stock PyTorch on this GPU (reference forward graph + autograd,
    bf16 autocast, channels_last, cuDNN benchmark, L1 (+ 3 no-grad perturbed forwards), clip_grad_norm_ + fused AdamW).
    Measurement only; a fresh model, never on the product path."""
    from Unetmodel import Unet
    torch.manual_seed(7)
    net = Unet(dropout_rate=0.2).to(dev).to(memory_format=torch.channels_last).train()
    opt = torch.optim.AdamW(net.parameters(), lr=7e-4, weight_decay=1e-3, fused=True)
    x = torch.randn(B, 4, H, W, device=dev).contiguous(memory_format=torch.channels_last)
    t = torch.rand(B, 1, H, W, device=dev)
    saved = torch.backends.cudnn.benchmark
    torch.backends.cudnn.benchmark = True

    def step():
        opt.zero_grad(set_to_none=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = stock_pytorch_forward(net, x)
            loss = 0.9 * (out.float() - t).abs().mean()
            if use_pert:
                with torch.no_grad():
                    sd = x.float().std(dim=(0, 2, 3), keepdim=True)
                    ys = [stock_pytorch_forward(net, x + torch.randn_like(x) * sd * 0.01) for _ in range(3)]
                loss = loss + 0.1 * sum((out.float() - y.float()).abs().mean() for y in ys) / 3
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt.step()

    try:
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(steps):
            step()
        e.record()
        torch.cuda.synchronize()
        ms = s.elapsed_time(e) / steps
    finally:
        torch.backends.cudnn.benchmark = saved
    return {"ms_per_step": ms, "value": B / (ms * 1e-3), "unit": "samples/s",
            "note": "stock PyTorch autograd on the same GPU: bf16 autocast, channels_last, cudnn.benchmark, fused AdamW"}


def dp_gradient_check(net, crit, sync, x, t, world):
    """Data-parallel parity inside the bench run (every rank): the gradients the overlapped, bucketed NCCL all-reduce leaves
    in p.grad must equal the mean over ranks of the gradients each rank computes alone on its own shard (same Dropout2d
    masks replayed, per-replica BatchNorm).  Returns the global rel-L2 difference (max over ranks)."""
    import torch.distributed as dist
    import nsm_train
    B = x.shape[0]
    gm = torch.Generator().manual_seed(4242 + dist.get_rank())
    masks = []
    for name, cin, _ in nsm_train.BLOCKS:
        p = getattr(net, name).conv[3].p
        masks.append(torch.empty(B, cin, 1, 1).bernoulli_(1 - p, generator=gm).div_(1 - p) if p > 0 else None)

    def grads(with_sync):
        net._grad_sync = sync if with_sync else None
        for p in net.parameters():
            p.grad = None
        nsm_train.replay_masks(net, masks)
        out = net(x)
        was = crit.perturb_weight
        crit.perturb_weight = 0.0                      # the perturbed forwards draw fresh noise: keep the check exact
        loss, _ = crit(net, out, t, x)
        crit.perturb_weight = was
        loss.backward()
        if with_sync:
            sync.finish()
        return torch.cat([p.grad.detach().reshape(-1).double() for p in net.parameters()])

    local_g = grads(False)
    dist.all_reduce(local_g, op=dist.ReduceOp.SUM)
    want = local_g / world
    got = grads(True)
    net._grad_sync = sync
    r = ((got - want).norm() / want.norm()).reshape(1)
    dist.all_reduce(r, op=dist.ReduceOp.MAX)
    for p in net.parameters():
        p.grad = None
    return {"allreduced_vs_mean_of_local_rel_l2": float(r.item()), "ok": bool(r.item() <= 1e-4),
            "what": "p.grad after GradSync (bucketed NCCL all-reduce overlapped with backward) vs all_reduce(SUM)/N of "
                    "the same step's local gradients, all 66 tensors, max over ranks"}


def measure_train(args, own_process_group=True):
    """Times the training step; returns the JSON line (rank 0) or None (other ranks)."""
    import torch.distributed as dist
    import nsm
    from Unetmodel import Unet
    from pert_loss import EnhancedCustomLoss
    from parallel import GradSync

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1 and own_process_group:
        dist.init_process_group("nccl", device_id=dev)
    nsm.require_device()
    precision = args.train_precision or "bf16"
    B = args.train_batch or 32
    H, W = args.train_size or 512, args.train_size or 512
    use_pert = not args.no_perturb

    torch.manual_seed(42)
    net = Unet(dropout_rate=0.2, precision=precision).to(dev).train()
    crit = EnhancedCustomLoss(dev, alpha=0.9, perturb_weight=0.1 if use_pert else 0.0).train()
    from nsm_optim import FusedAdamWClip
    opt = FusedAdamWClip(net.parameters(), lr=7e-4, weight_decay=1e-3, max_norm=1.0)    # main.py:405,955 in two launches
    sync = GradSync(net) if world > 1 else None
    g = torch.Generator().manual_seed(100 + rank)
    x_host = torch.randn(B, 4, H, W, generator=g).pin_memory()
    t_host = torch.rand(B, 1, H, W, generator=g).pin_memory()
    x, t = x_host.to(dev), t_host.to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(xd, td):
        opt.zero_grad(set_to_none=True)
        out = net(xd)
        loss, _ = crit(net, out, td, xd)
        loss.backward()
        if sync is not None:
            sync.finish()
        opt.step()                                  # non-finite scan + clip_grad_norm_(1.0) + AdamW, no host sync
        return loss

    steps = args.steps
    for _ in range(max(args.warmup, 3)):
        step(x, t)
    barrier()
    dp_check = dp_gradient_check(net, crit, sync, x, t, world) if world > 1 else None
    # The whole step (forward, loss, backward, NCCL gradient buckets, clip + AdamW) as ONE CUDA graph launch
    # (nsm_graph.GraphedTrainStep): same kernels, same arithmetic, no Python between them.  --no-graph times the eager loop.
    graphed = None
    if not args.no_graph:
        try:
            from nsm_graph import GraphedTrainStep
            graphed = GraphedTrainStep(net, crit, opt, x, t, sync=sync, warmup=2)
            eager_step = step

            def step(xd, td):                         # noqa: F811
                return graphed(xd if xd is not graphed.x else None, td if td is not graphed.t else None)
            x, t = graphed.x, graphed.t               # the device-resident timing feeds the static buffers in place
        except Exception as ex:                       # capture refused (e.g. a collective that cannot be captured): eager
            graphed = None
            graph_error = repr(ex)[:200]
            torch.cuda.synchronize()
    for _ in range(2):
        step(x, t)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    evs = []
    barrier()
    launches0 = nsm.launch_count()
    for _ in range(args.steps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        step(x, t)
        e.record()
        evs.append((s, e))
    barrier()
    launches = nsm.launch_count() - launches0
    if graphed is not None:
        # replays do not pass through the library's launch counter: count one eager step of the same sequence
        l0 = nsm.launch_count()
        eager_step(x, t)
        torch.cuda.synchronize()
        launches = (nsm.launch_count() - l0) * args.steps
    dev_ms = sum(s.elapsed_time(e) for s, e in evs)
    # end to end like a training loop with a pinned, prefetching loader (setdata_b200.DeviceFeeder's scheme): every step's
    # batch is copied from pinned host memory inside the timed region -- on a side stream, one step ahead -- and every
    # step's loss is read back to the host (asynchronously into pinned memory; all of them are on the host at the end).
    copy_stream = torch.cuda.Stream(device=dev)
    xds = [torch.empty_like(x) for _ in range(2)]
    tds = [torch.empty_like(t) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    losses_host = torch.zeros(args.steps, dtype=torch.float32).pin_memory()

    def prefetch(k):
        slot = k & 1
        with torch.cuda.stream(copy_stream):
            if k >= 2:
                copy_stream.wait_event(consumed[slot])
            xds[slot].copy_(x_host, non_blocking=True)
            tds[slot].copy_(t_host, non_blocking=True)
            ready[slot].record(copy_stream)

    torch.cuda.synchronize()
    t0 = time.perf_counter()
    prefetch(0)
    for k in range(args.steps):
        slot = k & 1
        if k + 1 < args.steps:
            prefetch(k + 1)
        torch.cuda.current_stream().wait_event(ready[slot])
        loss = step(xds[slot], tds[slot])
        consumed[slot].record()
        losses_host[k:k + 1].copy_(loss.detach().reshape(1), non_blocking=True)   # D2H read of the step's loss
    torch.cuda.synchronize()
    t_e2e = time.perf_counter() - t0
    last = float(losses_host[-1])
    barrier()
    clocks = sampler.stop()
    nsm.profile_enable(True)
    (eager_step if graphed is not None else step)(x, t)
    torch.cuda.synchronize()
    rows = nsm.profile_read()
    nsm.profile_enable(False)
    if graphed is not None:
        graphed.check()

    tt = torch.tensor([dev_ms, t_e2e * 1e3], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms = tt.tolist()
    if rank != 0:
        if world > 1 and own_process_group:
            dist.destroy_process_group()
        return None
    value = world * B * args.steps / (dev_ms * 1e-3)
    e2e_value = world * B * args.steps / (e2e_ms * 1e-3)
    pk = peaks()
    fam, other = {}, {}
    for name, ms, fl, by in rows:
        key = name.split()[0]
        if key in ("conv_gemm", "wgrad_gemm"):
            d = fam.setdefault(key, [0.0, 0.0, 0])
            d[0] += ms; d[1] += fl; d[2] += 1
        else:
            d = other.setdefault(key, [0.0, 0.0, 0])
            d[0] += ms; d[1] += by; d[2] += 1
    conv_ms = sum(v[0] for v in fam.values())
    conv_fl = sum(v[1] for v in fam.values())
    achieved = conv_fl / max(conv_ms, 1e-9) / 1e9
    step_ms = dev_ms / args.steps
    roofline = {"kernel": "conv_gemm_kernel + wgrad_gemm_kernel (tcgen05 implicit GEMM; fwd, dgrad, wgrad)",
                "bound": "tensor", "achieved": achieved, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": achieved / pk["bf16_tflops_sustained"], "traffic": None,
                "peak_source": f"{pk['source']} cuBLAS bf16 sustained (MEASURED_PEAKS.json)",
                "kernel_share_of_step": conv_ms / step_ms,
                "families": {k: {"ms": v[0], "tflops": v[1] / max(v[0], 1e-9) / 1e9, "launches": v[2]}
                             for k, v in fam.items()},
                "streaming_kernels": {k: {"ms": v[0], "gbs": v[1] / max(v[0], 1e-9) / 1e6, "launches": v[2]}
                                      for k, v in other.items()},
                "note": "padded thin layers (16/4 -> 64 channels) are counted with their padded FLOPs; streaming_kernels "
                        "are event-timed per launch with their algorithmic bytes"}
    flop = (TRAIN_FLOP_PER_SAMPLE + (PERT_FLOP_PER_SAMPLE if use_pert else 0)) * (H * W) / (512 * 512)
    line = {"metric": "U-Net train samples/s", "value": value, "unit": "samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": precision, "data": "synthetic",
            "config": {"workload": f"cfg2: U-Net training step, batch {B} of {H}x{W} crops per GPU, {precision}, "
                                   f"CustomLoss(alpha 0.9){' + PerturbationLoss(3 copies, weight 0.1)' if use_pert else ''}"
                                   ", Dropout2d, train-mode BatchNorm, fused non-finite scan + grad clip 1.0 + AdamW",
                       "algorithmic_tflop_per_step": flop * B / 1e12,
                       "model_tflops": flop * B * world / (step_ms * 1e-3) / 1e12,
                       "l2": "256 MiB buffer written before every timed step; activations >> L2",
                       "parallelism": f"dp{world}: per-rank batch, NCCL gradient all-reduce overlapped with backward"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "samples/s", "h2d_bytes_per_step": (x_host.numel() + t_host.numel()) * 4,
                    "d2h_bytes_per_step": 4, "ms_per_step": e2e_ms / args.steps, "last_loss": last},
            "gpu_launches": launches,
            "cuda_graph": (graphed is not None),
            "dp_check": dp_check,
            "roofline": roofline}
    if dp_check is not None and not dp_check["ok"]:
        raise RuntimeError(f"data-parallel gradient check failed: {dp_check}")
    if world == 1 and not args.no_stock and precision == "bf16":
        del net, crit, opt
        torch.cuda.empty_cache()
        line["stock_pytorch_gpu"] = stock_pytorch_train_baseline(B, H, W, dev, use_pert)
    if world > 1 and own_process_group:
        dist.destroy_process_group()
    return line


def run_b200_train(args):
    line = measure_train(args)
    if line is not None:
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="infer", choices=["infer", "train", "cfg4"],
                    help="infer: cfg1 (one 1080p frame per GPU, fp32 mode; the headline) + attached cfg2/cfg3 training "
                         "measurement; train: cfg2/cfg3 alone; cfg4: 16 frames of 3840x2160, bf16, sharded by frame over "
                         "the N GPUs")
    ap.add_argument("--precision", default=None, choices=["fp32", "bf16"])
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--height", type=int, default=None)
    ap.add_argument("--width", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-stock", action="store_true", help="skip the stock PyTorch/cuDNN same-GPU baseline")
    ap.add_argument("--no-perturb", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="training: eager Python-driven launches instead of the CUDA graph")
    ap.add_argument("--no-train", action="store_true", help="inference workload: skip the attached training measurement")
    ap.add_argument("--train-precision", default=None, choices=["fp32", "bf16"])
    ap.add_argument("--train-batch", type=int, default=None)
    ap.add_argument("--train-size", type=int, default=None)
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "train":
        args.train_precision = args.train_precision or args.precision
        args.train_batch = args.train_batch or args.batch
        args.train_size = args.train_size or args.height
        run_b200_train(args)
    elif args.workload == "cfg4":
        from parallel import shard_frames
        world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
        args.precision = args.precision or "bf16"
        args.batch, args.height, args.width = len(shard_frames(16, rank, world)), 2160, 3840
        args.no_train = args.no_stock = args.no_cpu_baseline = True
        run_b200(args)
    else:
        args.precision = args.precision or "fp32"
        args.batch, args.height, args.width = args.batch or 1, args.height or H1080, args.width or W1080
        run_b200(args)


if __name__ == "__main__":
    main()
