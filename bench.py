#!/usr/bin/env python
"""bench.py -- MPI frames/s (96 planes, 1024^2) on N B200s, with roofline, end-to-end and CPU-baseline legs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], "FFHQ1024"): per GPU a batch of 4 MPIs, 96 planes, 1024^2 textures, one
1024^2 view per MPI, random RGBA in [0,1), FFHQ geometry, in-envelope poses.  One "step"
renders the batch (4 frames: RGB + depth).  Views are sharded across ranks (weak scaling: 4 frames per GPU) and
the step ends with ONE all-gather of the frames (NCCL).  `value` = frames/s, whole job, inputs resident in HBM.

Keys beyond the base contract: `roofline` (dominant kernel vs measured HBM peak), `cpu_baseline` (the reference's
PyTorch grid_sample+cumprod op sequence, restated in oracle/torch_port.py, timed on this box's host cores),
`e2e` (host buffers -> C-ABI host entry point -> host frames, copies inside the timed region), `train_step`
(forward+backward through the autograd Function).
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
sys.path.insert(0, ROOT)

N_PLANES, RES, BATCH = 96, 1024, 4
WORKLOAD = "FFHQ1024: 96 planes, 1024^2 textures and views, 4 MPIs x 1 view per GPU, forward render"


def algorithmic_bytes_fwd(n, ht, wt, h, w):
    return 16 * n * ht * wt + 16 * h * w            # read every texel once, write RGB+depth


def algorithmic_bytes_bwd(n, ht, wt, h, w):
    return 2 * 16 * n * ht * wt + 16 * h * w        # re-read the MPI, write d rgba once, read upstream grads


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy kernel)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        inside = [r for (ts, r) in self.rows if t0 is not None and t0 <= ts <= t1 + 0.03]
        window = "timed region"
        if not inside:      # timed region shorter than the sampling period: use everything since warm-up started
            inside, window = [r for (_, r) in self.rows], "warm-up + timed region"
        for r in inside:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "window": window}


# ----------------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's torch op sequence on the host cores
# ----------------------------------------------------------------------------------------------------------------
def cpu_reference_frames_per_s(steps, warmup, budget_s):
    """Times oracle/torch_port.render (== MPIRenderer.render arithmetic) on a bounded sample of the workload:
    the top `rows` rows of one 1024^2 frame of one 96-plane MPI.  Returns (frames/s, sample text, cores)."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import torch_port
    from ml_gmpi_b200 import synth
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    case = synth.make_case(n_planes=N_PLANES, tex=RES, img=RES, n_mpi=1, seed=1234, device="cpu")
    dhw = case.dhw[0]

    def run(rows):
        ray = case.ray_dir[:, :, :rows, :].contiguous()
        t0 = time.perf_counter()
        with torch.no_grad():
            torch_port.render(case.rgba, dhw, [ray], [case.eye], [case.z_dir], True)
        return time.perf_counter() - t0

    run(8)                                           # page in
    per_row = run(32) / 32.0
    rows = int(max(8, min(RES, budget_s / max(steps + warmup, 1) / per_row)))
    rows -= rows % 8
    for _ in range(warmup):
        run(rows)
    ts = [run(rows) for _ in range(steps)]
    t = sum(ts) / len(ts)
    fps = (rows / RES) / t
    sample = (f"top {rows} of {RES} rows of 1 frame (1 MPI, {N_PLANES} planes, {RES}^2 texture), forward, no_grad, "
              f"torch {torch.__version__} CPU ops, {steps} steps of {t:.2f} s")
    return fps, sample, cores, t * 1e3


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    fps, sample, cores, ms = cpu_reference_frames_per_s(args.steps, args.warmup, budget_s=args.ref_budget_s)
    line = {
        "impl": "reference", "metric": "MPI frames/s (96 planes, 1024^2)", "value": fps, "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "planes": N_PLANES, "tex": RES, "img": RES, "device": "host CPU"},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-budget-s", type=float, default=150.0, help="wall-clock budget of the --impl reference run")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-train-step", action="store_true")
    ap.add_argument("--planes", type=int, default=N_PLANES)
    ap.add_argument("--res", type=int, default=RES)
    ap.add_argument("--batch", type=int, default=BATCH)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    args.warmup = max(args.warmup, 3)

    import torch
    import torch.distributed as dist
    import ml_gmpi_b200 as g
    from ml_gmpi_b200 import _lib, synth, dist as gdist, host_api

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py (impl ours) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    NP, R, B = args.planes, args.res, args.batch

    case = synth.make_case(n_planes=NP, tex=R, img=R, n_mpi=B, seed=1234 + rank, device=dev)
    flags = torch.zeros(1, dtype=torch.int32, device=dev)
    color = torch.empty((B, 3, R, R), device=dev)
    depth = torch.empty((B, 1, R, R), device=dev)
    frames_all = frames_local = fused = None
    gather_mode = "none"
    if world > 1:
        try:    # all-gather fused into the render epilogue: peer stores into symmetric memory over NVLink
            fused = gdist.FrameGather(B, R, R, dev)
            gather_mode = "fused: render epilogue stores frames into every rank's symmetric-memory buffer (NVLink), 1 device barrier per step"
        except Exception as ex:   # no symmetric memory on this box: render, then ncclAllGather
            if rank == 0:
                print(f"[bench] symmetric memory unavailable ({type(ex).__name__}: {ex}); using ncclAllGather", file=sys.stderr)
            frames_all = torch.empty((world * B, 4, R, R), device=dev)
            frames_local = torch.empty((B, 4, R, R), device=dev)
            gather_mode = "render, then one ncclAllGather of [V,4,H,W] frames"
    stream = torch.cuda.current_stream(dev)
    opts = _lib.OPT_ALIGN_CORNERS | _lib.OPT_CHECK_LAST_PLANE | _lib.OPT_COLOR_MINUS1_1
    launches = [0]

    def render():
        _lib.check(lib.gmpi_mpi_render_fwd(case.rgba.data_ptr(), case.view2mpi.data_ptr(), case.dhw.data_ptr(),
                                           case.ray_dir.data_ptr(), case.eye.data_ptr(), case.z_dir.data_ptr(),
                                           color.data_ptr(), depth.data_ptr(), flags.data_ptr(), B, B, NP, R, R, R, R, opts,
                                           stream.cuda_stream))
        launches[0] += 1

    def render_gather():
        fused.render(case.rgba, case.dhw, case.view2mpi, case.ray_dir, case.eye, case.z_dir, flags, check_last_plane=True,
                     color_minus1_1=True)
        launches[0] += 1

    def step():
        if fused is not None:   # the one collective of the path, fused into the kernel
            render_gather()
            fused.finish()
        else:
            render()
            if world > 1:
                frames_local[:, :3].copy_(color); frames_local[:, 3:].copy_(depth)
                dist.all_gather_into_tensor(frames_all, frames_local)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for _ in range(args.warmup):
        step()
    barrier()
    assert int(flags.item()) == 0, f"render flagged {int(flags.item())} on the synthetic workload"

    # --- timed region: whole step (render [+ all-gather]), CUDA events, max over ranks ---------------------------
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    launches[0] = 0
    t_wall0 = time.time()
    ev[0].record(stream)
    for i in range(args.steps):
        kev[i][0].record(stream)
        if fused is not None:
            render_gather()
            kev[i][1].record(stream)
            fused.finish()
        else:
            render()
            kev[i][1].record(stream)
            if world > 1:
                frames_local[:, :3].copy_(color); frames_local[:, 3:].copy_(depth)
                dist.all_gather_into_tensor(frames_all, frames_local)
    ev[1].record(stream)
    barrier()
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    total_ms = ev[0].elapsed_time(ev[1])
    kernel_ms = sum(a.elapsed_time(b) for a, b in kev) / args.steps
    n_launch = launches[0]
    t = torch.tensor([total_ms, kernel_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = float(t[0]), float(t[1])
    ms_per_step = total_ms / args.steps
    frames_per_s = world * B / (ms_per_step * 1e-3)

    peak, peak_src = measured_peak()
    alg = algorithmic_bytes_fwd(NP, R, R, R, R) * B
    achieved = alg / (kernel_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("fwd_dram_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": lib.gmpi_mpi_render_fwd_variant(NP, R, R, R, R).decode(), "kernel_ms": kernel_ms,
                "algorithmic_bytes_per_launch": alg, "peak_source": peak_src}

    # --- forward+backward (BASELINE configs[2] is fwd+bwd): autograd Function, grad w.r.t. rgba -------------------
    train = None
    if not args.no_train_step:
        rg = case.rgba.requires_grad_(True)
        gcol = torch.randn((B, 3, R, R), device=dev)
        tsteps = max(3, min(args.steps, 5))

        def fb():
            rg.grad = None
            c, d = g.render_views(rg, case.dhw, case.view2mpi, case.ray_dir, case.eye, case.z_dir, color_minus1_1=True)
            (c * gcol).sum().backward()
        fb(); fb()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(tsteps):
            fb()
        e1.record(stream)
        barrier()
        tms = torch.tensor([e0.elapsed_time(e1) / tsteps], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        algfb = (algorithmic_bytes_fwd(NP, R, R, R, R) + algorithmic_bytes_bwd(NP, R, R, R, R)) * B
        train = {"value": world * B / (float(tms[0]) * 1e-3), "unit": "frames/s (forward+backward, d/d rgba)", "ms_per_step": float(tms[0]),
                 "steps": tsteps, "roofline_frac": algfb / (float(tms[0]) * 1e-3) / 1e9 / peak,
                 "note": "includes grad-buffer memset and the torch (c*g).sum() loss kernels"}
        case.rgba.requires_grad_(False)
        rg.grad = None
        del gcol
        torch.cuda.empty_cache()

    # --- end to end: pinned HOST buffers -> C-ABI host entry point -> pinned host frames ---------------------------
    e2e = None
    if not args.no_e2e:
        h_rgba = torch.empty(case.rgba.shape, dtype=torch.float32).pin_memory()
        h_rgba.copy_(case.rgba)
        hc = {k: getattr(case, k).cpu().pin_memory() for k in ("dhw", "view2mpi", "ray_dir", "eye", "z_dir")}
        oc = torch.empty((B, 3, R, R), dtype=torch.float32).pin_memory()
        od = torch.empty((B, 1, R, R), dtype=torch.float32).pin_memory()
        esteps = max(2, min(args.steps, 4))

        def e2e_step():
            return host_api.render_host(h_rgba, hc["dhw"], hc["view2mpi"], hc["ray_dir"], hc["eye"], hc["z_dir"],
                                        check_last_plane=True, color_minus1_1=True, device=local, out_color=oc, out_depth=od)
        e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(esteps):
            _, _, fl = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / esteps
        tt = torch.tensor([dt], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        assert fl == 0
        assert torch.equal(oc.to(dev), color) and torch.equal(od.to(dev), depth), "e2e result differs from the device-resident run"
        h2d = sum(int(x.numel()) * x.element_size() for x in (h_rgba, *hc.values()))
        d2h = (oc.numel() + od.numel()) * 4 + 4
        e2e = {"value": world * B / float(tt[0]), "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "steps": esteps, "ms_per_step": float(tt[0]) * 1e3, "api": "ml_gmpi_b200.host_api.render_host -> gmpi_mpi_render_fwd_host (C ABI)",
               "h2d_gbs": h2d / float(tt[0]) / 1e9}
        del h_rgba, oc, od
        lib.gmpi_mpi_release_host_cache()

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        fps, sample, cores, _ = cpu_reference_frames_per_s(steps=2, warmup=1, budget_s=20.0)
        cpu = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": "MPI frames/s (96 planes, 1024^2)", "value": frames_per_s, "unit": "frames/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD if (NP, R, B) == (N_PLANES, RES, BATCH) else f"{NP} planes, {R}^2, {B} MPIs x 1 view per GPU",
                       "planes": NP, "tex": R, "img": R, "mpis_per_gpu": B, "views_per_gpu": B, "parallelism": f"views sharded x{world}; all-gather of frames: {gather_mode}",
                       "l2": f"inputs {case.rgba.numel() * 4 / 1e9:.2f} GB per GPU >> 126 MB L2 (no flush needed)",
                       "validate": "geometric flags fused in-kernel; range scan off in the timed region"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": n_launch, "roofline": roofline, "cpu_baseline": cpu, "train_step": train,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
