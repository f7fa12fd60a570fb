#!/usr/bin/env python
"""bench.py -- headline benchmark of the compositing hot path (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU algorithm (oracle port) on host cores

A "step" is one pass of the hot path over one frame stack. Default workload = BASELINE.json configs[2] (c3):
`--mode outlier -t abs/0.05/0.2 -l extreme -b first`, 200 frames x 6000x4000 RGB8 (synthetic series S2, seed 42).

N > 1 (one rank per GPU under torchrun): STRONG scaling by default -- the workload's own 6000x4000 image is cut into N
horizontal bands, one per GPU, no data-path collective; the only exchange is the final band gather (NCCL gather of the
composite and mask bands onto rank 0), which is timed inside `e2e` and verified against rank 0's single-GPU result. The
weak-scaling figure (every rank a full-size band of an N-times taller image) is timed in the same run as `weak_scaling`.

value  = pixel-frames/s, stack resident in HBM, CUDA events around K steps, max over ranks.
e2e    = same metric through the C ABI with pinned HOST frames: H2D upload of all frames + kernels + (N > 1: band gather)
         + D2H of composite and mask inside the timed region.
verified = after the timed region, rows of the timed stack's output are compared bit for bit with the CPU oracle on the
         same synthetic frames (and, at N > 1, the gathered image with rank 0's single-GPU result).
other_workloads (N = 1, default run only): the other BASELINE.json configs (c2 darker, c4, c5 video) and the two
         adversarial series, each with its own roofline / cpu_baseline / e2e / verified block.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
T_START = time.perf_counter()

WORKLOADS = {
    # name: (mode, frames, height, width, synthetic kind, description)
    "c3-outlier-abs-extreme": ("outlier", 200, 4000, 6000, 2, "--mode outlier -t abs/0.05/0.2 -l extreme -b first, 200 x 6000x4000 RGB8"),
    "c2-darker": ("darker", 200, 4000, 6000, 2, "--mode darker, 200 x 6000x4000 RGB8"),
    "c2-lighter": ("lighter", 200, 4000, 6000, 2, "--mode lighter, 200 x 6000x4000 RGB8"),
    "c4-outlier-rel-forward": ("outlier-rel", 1000, 2160, 3840, 2, "--mode outlier -t rel/3.0/5.0 -l forward -b first, 1000 x 3840x2160 RGB8"),
    "c1-minimal": ("outlier-c1", 25, 768, 1024, 1, "cmd_examples/minimal: defaults abs/0.05/0.2, extreme, 25 x 1024x768 RGB8 (background first)"),
    # adversarial / noisier series: not configs of BASELINE.json, carried for the worst case
    "a1-iid-uniform": ("outlier", 200, 2048, 2048, 3, "worst case: iid uniform bytes (every pixel through the iterative tier and the per-frame path), "
                       "--mode outlier -t abs/0.05/0.2 -l extreme -b first, 200 x 2048x2048 RGB8"),
    "a4-gauss-noise": ("outlier", 200, 4000, 6000, 4, "gradient + Gaussian-like noise (sigma ~ 4.6, range +-14) + discs: an outlier frame in every pixel, "
                       "--mode outlier -t abs/0.05/0.2 -l extreme -b first, 200 x 6000x4000 RGB8"),
    "c5-video": ("video", 1800, 1064, 1904, 2, "chrono-video: --video-in 0/25/1 over 1800 x 1080p frames cropped to 1904x1064 by shake offsets in [-8,8]^2, "
                 "outlier abs/0.05/0.2 extreme (1824 output frames)"),
}
DEFAULT_WORKLOAD = "c3-outlier-abs-extreme"
OTHER_WORKLOADS = ["c2-darker", "a4-gauss-noise", "a1-iid-uniform", "c4-outlier-rel-forward", "c5-video", "c1-minimal"]
L2_FLUSH_BYTES = 512 << 20  # written between the timed steps of a stack that fits the 126 MB L2
OTHERS_BUDGET_S = 210.0  # no further workload is started once the run is this old


def log(msg):
    print(f"[bench {time.perf_counter() - T_START:6.1f}s] {msg}", file=sys.stderr, flush=True)


def config_of(wl):
    """The `config` block, identical in both arms."""
    mode, n, H, W, kind, desc = WORKLOADS[wl]
    ng = (n + 15) // 16
    gb = ((H * W + 31) // 32) * 3 * ng * 512 / 1e9
    return {"workload": wl, "description": desc, "frames": n, "height": H, "width": W, "channels": 3, "series": f"S{kind} seed 42",
            "l2": ("stack (%.1f GB) larger than L2: no flush needed between timed steps" % gb) if gb * 1e9 > L2_FLUSH_BYTES // 2 else
                  ("stack (%.0f MB) fits L2: every timed step is preceded by an untimed %d MB write and timed on its own" % (gb * 1e3, L2_FLUSH_BYTES >> 20))}


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy read+write)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md, 6.65 TB/s)"


def traffic_from_profiles(kernel_key):
    path = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    try:
        t = json.load(open(path))[kernel_key]
        return t["dram_bytes_read"] + t["dram_bytes_write"]
    except Exception:
        return None


class ClockSampler(threading.Thread):
    """Samples SM clocks and throttle reasons of one GPU during the timed region (NVML)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop_evt = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {"hw_slowdown": nv.nvmlClocksThrottleReasonHwSlowdown, "hw_thermal_slowdown": nv.nvmlClocksThrottleReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": nv.nvmlClocksThrottleReasonSwThermalSlowdown, "sw_power_cap": nv.nvmlClocksThrottleReasonSwPowerCap}
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.004)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {"sm_mhz": (s[len(s) // 2] if s else None), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(s)}


# ------------------------------------------------------------------------------------------------ oracle side (CPU)
def oracle():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib as orc
    return orc


def oracle_call(orc, mode, st, n_threads, indices=None):
    if mode in ("outlier", "outlier-c1", "video"):
        return orc.outlier(st, orc.threshold(True, 0.05, 0.2), 0, 2, n_threads=n_threads, indices=indices)
    if mode == "outlier-rel":
        return orc.outlier(st, orc.threshold(False, 3.0, 5.0), 0, 4, n_threads=n_threads, indices=indices)
    return orc.simple(st, mode == "darker", n_threads=n_threads)


def time_cpu(orc, mode, kind, n, H, W, n_threads, target_s, max_rows=256):
    """Times the oracle (CPU port of the reference algorithm) on a bounded row band of the workload, generated by the
    oracle's own twin of the synthetic series (nothing of the product library is loaded)."""
    probe_rows = min(H, 2 * n_threads if mode != "video" else n_threads)
    r0 = max(0, H // 2 - probe_rows // 2)
    n_gen = n if mode != "video" else 40  # video: a block of 16 windows of 25 frames
    st = orc.synth_frames(kind, 42, n, W, H, rows=probe_rows, row0=r0, n_out=n_gen)

    def call(stack):
        if mode == "video":
            t0 = time.perf_counter()
            for k in range(16):
                oracle_call(orc, mode, stack[k:k + 25], n_threads)
            return time.perf_counter() - t0, 16 * 25
        t0 = time.perf_counter()
        oracle_call(orc, mode, stack, n_threads)
        return time.perf_counter() - t0, n

    dt, frames = call(st)
    rate = frames * probe_rows * W / max(dt, 1e-6)
    rows = int(max(1, min(H, max_rows, rate * target_s / (frames * W))))
    if rows >= n_threads:
        rows -= rows % n_threads
    if rows != probe_rows:
        r0 = max(0, min(H // 2, H - rows))
        st = orc.synth_frames(kind, 42, n, W, H, rows=rows, row0=r0, n_out=n_gen)
        dt, frames = call(st)
    else:
        rows = probe_rows
    return frames * rows * W / dt, rows, r0, dt, st, call


def cpu_baseline_block(orc, mode, kind, n, H, W, target_s, with_reference_threading):
    cores = os.cpu_count() or 1
    rate, rows, r0, dt, _, _ = time_cpu(orc, mode, kind, n, H, W, cores, target_s)
    what = "16 windows of 25 frames" if mode == "video" else f"{n} frames"
    cb = {"value": rate, "unit": "pixel-frames/s", "cores": cores, "kind": "port",
          "sample": f"rows [{r0}, +{rows}) of {H} x {W} px x {what} of the workload ({dt:.1f} s, row-parallel over all host threads), in-memory stack: "
                    "excludes the reference's JPEG decode, deflate/inflate and temp-file I/O (flatters the reference)"}
    if with_reference_threading:  # outlier photo compute is single-threaded in the reference (src/chrono.rs:96,169)
        rate1, rows1, _, dt1, _, _ = time_cpu(orc, mode, kind, n, H, W, 1, target_s * 0.5, max_rows=16)
        cb["value_reference_threading"] = rate1
        cb["cores_reference_threading"] = 1
        cb["sample_reference_threading"] = f"{rows1} rows ({dt1:.1f} s, one thread like src/chrono.rs:96,169)"
    if mode == "video":
        cb["note"] = "the reference runs video frames in parallel over output frames (src/main.rs:260-261); the port runs the rows of each window in parallel: same cores busy"
    return cb


def run_reference(args, wl):
    """--impl reference: the reference's own CPU algorithm (oracle port; the Rust crate cannot be built here: no cargo)
    with all host threads on a bounded sample of the workload. Rank 0 only. Loads nothing of the product library."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    orc = oracle()
    mode, n, H, W, kind, desc = WORKLOADS[wl]
    cores = os.cpu_count() or 1
    rate, rows, r0, dt, st, call = time_cpu(orc, mode, kind, n, H, W, cores, target_s=3.0)
    times, frames = [], n
    for i in range(args.warmup + args.steps):
        t, frames = call(st)
        if i >= args.warmup:
            times.append(t)
    ms = 1e3 * sum(times) / len(times)
    value = frames * rows * W / (ms / 1e3)
    what = "16 windows of 25 frames" if mode == "video" else f"{n} frames"
    sample = f"rows [{r0}, +{rows}) x {W} px x {what} of the workload per step, in-memory stack (no JPEG decode / temp files)"
    print(json.dumps({
        "impl": "reference", "metric": "pixel-frames/s", "value": value, "unit": "pixel-frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if args.gpus > 1 else "weak", "vs_baseline": None,
        "dtype": "u8", "data": "synthetic", "config": config_of(wl),
        "cpu_baseline": {"value": value, "unit": "pixel-frames/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "pixel-frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "CPU port of chrono-photo's algorithm (oracle/chrono_oracle.c), row-parallel over all host threads; the reference itself runs "
                "outlier photos single-threaded (src/chrono.rs:96,169)"}))


# ------------------------------------------------------------------------------------------------ CUDA side
def make_processor(cp, mode, seed=42, pixel_offset=0, **shard_args):
    if mode in ("outlier", "outlier-c1", "video"):
        return cp.OutlierProcessor(cp.Threshold.abs(0.05, 0.2), cp.BackgroundMode.FIRST, cp.OutlierSelectionMode.EXTREME, seed=seed, pixel_offset=pixel_offset,
                                   **shard_args)
    if mode == "outlier-rel":
        return cp.OutlierProcessor(cp.Threshold.rel(3.0, 5.0), cp.BackgroundMode.FIRST, cp.OutlierSelectionMode.ALL_FORWARD, seed=seed, pixel_offset=pixel_offset,
                                   **shard_args)
    return cp.SimpleProcessor(darker=(mode == "darker"))


def verify_rows(H, want=64):
    """Row blocks compared with the oracle: 16-row blocks at the top, across three disc tracks and at the bottom."""
    if H <= want:
        return [(0, H)]
    blk = want // 4
    starts = [0, H * 3 // 16 - blk // 2, H // 2 - blk // 2, H - blk]
    return [(max(0, min(H - blk, s)), blk) for s in starts]


def verify_photo(orc, np, mode, kind, n, H, W, image, mask):
    """image/mask: the full (H, W, 3) result of the timed stack. Bit-exact comparison of row blocks with the oracle."""
    bad, rows_done, where = 0, 0, []
    for r0, r in verify_rows(H):
        st = orc.synth_frames(kind, 42, n, W, H, rows=r, row0=r0)
        res = oracle_call(orc, mode, st, os.cpu_count() or 1)
        if mode.startswith("outlier"):
            oimg, omsk, _ = res
            d = (oimg != image[r0:r0 + r]).any(axis=2)
            if mask is not None:
                d |= (omsk != mask[r0:r0 + r]).any(axis=2)
        else:
            d = (res != image[r0:r0 + r]).any(axis=2)
        bad += int(d.sum())
        for y, x in np.argwhere(d)[:4]:
            where.append({"y": int(r0 + y), "x": int(x), "gpu": image[r0 + y, x].tolist() + ([int(mask[r0 + y, x, 0])] if mask is not None else []),
                          "oracle": (oimg if mode.startswith("outlier") else res)[y, x].tolist() + ([int(omsk[y, x, 0])] if mask is not None else [])})
        rows_done += r
    return rows_done, bad, where


def bind_to_gpu_cpus(index):
    """One rank per GPU: the process (and with it the first-touch placement of its pinned host frames) is bound to the CPUs NVML
    reports as local to the GPU, so that every rank's H2D stream crosses its own root complex instead of the socket interconnect.
    Returns the number of CPUs bound to, or None when NVML gives no answer."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (int(words[i // 64]) >> (i % 64)) & 1]
        allowed = os.sched_getaffinity(0)
        cpus = [c for c in cpus if c in allowed]
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


class Env:
    """torch / torch.distributed plumbing of one bench process."""
    cpu_affinity = None

    def __init__(self, args):
        import torch
        self.torch = torch
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py needs a CUDA device: the compositing path has no CPU fallback")
        self.multi_proc = self.world > 1
        self.dist = None
        if self.multi_proc:
            import torch.distributed as dist
            self.dist = dist
            torch.cuda.set_device(self.local_rank)
            self.cpu_affinity = bind_to_gpu_cpus(self.local_rank)
            # stdout carries the one JSON line only: NCCL prints its version banner with a plain printf when the communicator is
            # created, so file descriptor 1 points at stderr until the first collective is through
            sys.stdout.flush()
            saved_stdout = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=torch.device("cuda", self.local_rank))
                dist.barrier()
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved_stdout, 1)
                os.close(saved_stdout)
            self.devices, self.shard_rank, self.shard_world = [self.local_rank], self.rank, self.world
        else:
            self.devices, self.shard_rank, self.shard_world = list(range(args.gpus)), 0, 1  # one process drives all GPUs through one context
            torch.cuda.set_device(0)
        self.n_gpus = self.world if self.multi_proc else args.gpus

    def barrier(self):
        if self.multi_proc:
            self.dist.barrier()
        for d in self.devices:
            self.torch.cuda.synchronize(d)

    def max_over_ranks(self, x):
        if not self.multi_proc:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def min_over_ranks(self, x):
        if not self.multi_proc:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MIN)
        return float(t.item())


def run_photo(env, wl, steps, warmup, e2e_steps, do_e2e, do_cpu, cpu_target_s, do_verify, scaling, primary, shared=None):
    """One photo workload (outlier / darker / lighter). Returns the dict of its JSON block. `shared`: dict carrying the
    resident S2 stack / pinned host copy between workloads of the same series and shape (c3 -> c2)."""
    import numpy as np
    import chrono_photo_b200 as cp
    from chrono_photo_b200 import _lib
    from chrono_photo_b200.sharding import InterleavedShard, deinterleave, interleave_block_rows, shard_rows
    torch, dist = env.torch, env.dist
    mode, n, H_img, W, kind, desc = WORKLOADS[wl]
    is_outlier = mode.startswith("outlier")
    n_gpus = env.n_gpus
    weak = scaling == "weak" and env.shard_world > 1
    H = H_img * env.shard_world if weak else H_img
    row0, rows = (env.shard_rank * H_img, H_img) if weak else shard_rows(H_img, env.shard_rank, env.shard_world)
    # strong scaling: GPU g owns the row blocks with index = g mod G (objects are compact in the image: contiguous bands leave the
    # GPU whose band holds the most object pixels defining the step); contiguous bands when the image does not divide
    def strong_shard():
        b = interleave_block_rows(H_img, env.shard_world) if (env.shard_world > 1 and os.environ.get("CHB_BENCH_CONTIGUOUS", "0") != "1") else None
        return InterleavedShard(H_img, W, env.shard_rank, env.shard_world, b) if b else None
    ishard = None if weak else strong_shard()
    if ishard:
        row0, rows = ishard.row0, ishard.rows
    fill_args = ishard.fill_args() if ishard else dict(row0_global=row0, full_height=H)
    proc_args = ishard.processor_args() if ishard else dict(pixel_offset=row0 * W)
    key = (kind, n, rows, W, row0, H, ishard.block_rows if ishard else 0)
    ctx = shared.get("ctx") if shared else None
    own_ctx = ctx is None
    if own_ctx:
        ctx = cp.Context(env.devices)
        if len(env.devices) == 1:
            ctx.set_stream(0, torch.cuda.current_stream().cuda_stream)  # torch events then bracket the launches
        if shared is not None:
            shared["ctx"] = ctx
    stack = shared.get(("stack", key)) if shared else None
    if stack is None:
        if shared:  # a stack of another series / shape is not needed any more
            for k in [k for k in shared if isinstance(k, tuple) and k[0] in ("stack", "host")]:
                v = shared.pop(k)
                if k[0] == "stack":
                    v.close()
        stack = cp.FrameStack(ctx, W, rows, 3, n)
        if weak:  # every band is its own full-size series of the workload's recipe, seeds 42 + rank
            stack.fill_synthetic(kind, seed=42 + env.shard_rank, row0_global=0, full_height=H_img)
        else:
            stack.fill_synthetic(kind, seed=42, **fill_args)
        if shared is not None:
            shared[("stack", key)] = stack
    proc = make_processor(cp, mode, **(proc_args if is_outlier else {}))

    def timed_steps(st, pr):
        for _ in range(warmup):
            pr.process_device(st)
        env.barrier()
        _lib.lib().chb_launch_count_reset()
        if len(env.devices) == 1 and st.device_bytes(0) < L2_FLUSH_BYTES // 2:  # the stack would stay in L2: flush it between steps
            flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda")
            ms_total = 0.0
            for _ in range(steps):
                flush.fill_(1)
                torch.cuda.synchronize()
                ms_total += pr.process_device(st)  # device time of the call: CUDA events on the launching stream, inside the library
            del flush
            env.barrier()
            return env.max_over_ranks(ms_total / steps), int(_lib.lib().chb_launch_count())
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_wall0 = time.perf_counter()
        ev0.record()
        if is_outlier:  # K launches back to back, one wait at the end: no host round trip inside the timed region
            for _ in range(steps):
                pr.enqueue_device(st)
            st.wait()
        else:
            for _ in range(steps):
                pr.process_device(st)
        ev1.record()
        env.barrier()
        t_wall = time.perf_counter() - t_wall0
        n_launch = int(_lib.lib().chb_launch_count())
        ms_total = ev0.elapsed_time(ev1) if len(env.devices) == 1 else t_wall * 1e3
        return env.max_over_ranks(ms_total / steps), n_launch

    # ---- kernel-only: stack resident in HBM
    sampler = ClockSampler(env.devices[0])
    sampler.start()
    ms_step, launches = timed_steps(stack, proc)
    clocks = sampler.stop()
    launch_ms, main_ms = [], []
    flush = torch.empty(L2_FLUSH_BYTES, dtype=torch.uint8, device="cuda") if (len(env.devices) == 1 and stack.device_bytes(0) < L2_FLUSH_BYTES // 2) else None
    for _ in range(min(5, steps)):  # per-call device time (CUDA events on the launching stream, inside the library)
        if flush is not None:  # a stack that fits L2 is evicted first
            flush.fill_(1)
            torch.cuda.synchronize()
        launch_ms.append(proc.process_device(stack))
        main_ms.append(float(_lib.lib().chb_last_main_kernel_ms()) if is_outlier else launch_ms[-1])
    del flush
    # (the median of the five calls: one call that falls on a clock ramp after the idle gap must not move the roofline figure)
    kernel_ms = env.max_over_ranks(sorted(launch_ms)[len(launch_ms) // 2])
    main_kernel_ms = env.max_over_ranks(sorted(main_ms)[len(main_ms) // 2])
    if is_outlier and len(env.devices) == 1 and ms_step < kernel_ms:
        # K back-to-back calls of the timed region took ms_step each, gaps included: a call cannot be longer than that
        main_kernel_ms *= ms_step / kernel_ms
        kernel_ms = ms_step
    total_pf = float(n) * H * W
    value = total_pf / (ms_step / 1e3)

    # ---- roofline (algorithmic bytes of this rank's shard), whole call and dominant kernel
    alg_bytes = rows * W * 3 * (n + (2 if is_outlier else 1))  # read the stack once + composite (+ mask) write
    peak, peak_src = peaks()
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9
    kernel_name = "outlier_kernel" if is_outlier else "simple_int_kernel"
    traffic = traffic_from_profiles(kernel_name + ":" + wl) if n_gpus == 1 else None
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": (kernel_name + " + compact_hard_kernel + outlier_hard_kernel|outlier_hist_kernel [+ outlier_exact_kernel where the dense pass does not apply] (one call)") if is_outlier else kernel_name,
                "algorithmic_bytes_per_launch": alg_bytes, "avg_launch_ms": kernel_ms, "peak_source": peak_src,
                "slow_path_pixels_per_launch": int(_lib.lib().chb_last_slow_pixels()) if is_outlier else 0}
    if is_outlier:
        dom = alg_bytes / (main_kernel_ms / 1e3) / 1e9
        roofline["traffic_kernel"] = kernel_name
        roofline["tier_kernels_ms"] = kernel_ms - main_kernel_ms
        roofline["dominant_kernel"] = {"kernel": kernel_name, "avg_launch_ms": main_kernel_ms, "achieved": dom, "frac": dom / peak, "traffic": traffic,
                                       "note": "streaming kernel alone (CUDA events around it inside the library): reads the whole stack, finishes the certified "
                                               "pixels and the tiles it classifies frame by frame in place"}

    # ---- the other scaling regime, timed in the same run
    other_scaling = None
    if env.shard_world > 1 and primary:
        if weak:  # the workload's own image cut into N bands
            o_sh = strong_shard()
            o_row0, o_rows = (o_sh.row0, o_sh.rows) if o_sh else shard_rows(H_img, env.shard_rank, env.shard_world)
            o_stack = cp.FrameStack(ctx, W, o_rows, 3, n)
            o_stack.fill_synthetic(kind, seed=42, **(o_sh.fill_args() if o_sh else dict(row0_global=o_row0, full_height=H_img)))
            o_pargs = (o_sh.processor_args() if o_sh else dict(pixel_offset=o_row0 * W)) if is_outlier else {}
            o_ms, _ = timed_steps(o_stack, make_processor(cp, mode, **o_pargs))
            other_scaling = ("strong_scaling", {"value": float(n) * H_img * W / (o_ms / 1e3), "unit": "pixel-frames/s", "ms_per_step": o_ms,
                                                "image": f"{W}x{H_img} cut into {n_gpus} shards of {o_rows} rows" + (f" (interleaved blocks of {o_sh.block_rows} rows)" if o_sh else "")})
        else:  # every rank a full-size band of an N-times taller image
            o_stack = cp.FrameStack(ctx, W, H_img, 3, n)
            o_stack.fill_synthetic(kind, seed=42 + env.shard_rank, row0_global=0, full_height=H_img)
            o_ms, _ = timed_steps(o_stack, make_processor(cp, mode, pixel_offset=env.shard_rank * H_img * W))
            other_scaling = ("weak_scaling", {"value": float(n) * H_img * env.shard_world * W / (o_ms / 1e3), "unit": "pixel-frames/s", "ms_per_step": o_ms,
                                              "image": f"{W}x{H_img * n_gpus}: every GPU owns one full-size band ({H_img} rows) of the workload"})
        o_stack.close()

    # ---- gather helpers (N > 1 under torchrun): device bands -> rank 0 over NCCL -> pinned host image
    frame_bytes = rows * W * 3
    d_img = d_msk = g_img = g_msk = None
    equal_bands = ishard is not None or all(shard_rows(H_img, r, env.shard_world)[1] == rows for r in range(env.shard_world))

    def assemble(parts):  # rank 0: the gathered shards in rank order -> the image
        return deinterleave(torch.stack(parts), ishard.block_rows) if ishard else torch.cat(parts)

    if env.multi_proc and not weak and equal_bands:
        d_img = torch.empty((rows, W, 3), dtype=torch.uint8, device="cuda")
        d_msk = torch.empty((rows, W, 3), dtype=torch.uint8, device="cuda") if is_outlier else None
        if env.rank == 0:
            g_img = [torch.empty_like(d_img) for _ in range(env.world)]
            g_msk = [torch.empty_like(d_img) for _ in range(env.world)] if is_outlier else None

    def gather_last(st):
        """The last call's bands -> rank 0 (device). Returns device ms spent in the gather (max over ranks)."""
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        cp.fetch_last_device(st, d_img.data_ptr(), d_msk.data_ptr() if d_msk is not None else None)
        ev0.record()
        dist.gather(d_img, g_img, dst=0)
        if d_msk is not None:
            dist.gather(d_msk, g_msk, dst=0)
        ev1.record()
        torch.cuda.synchronize()
        return ev0.elapsed_time(ev1)

    # ---- verification of the timed stack's output
    verified = None
    if do_verify and not weak:
        full_img = full_msk = None
        vinfo = {}
        if env.multi_proc and d_img is not None:
            proc.process_device(stack)
            gather_last(stack)  # warm-up of the communicator's gather path
            gms = env.max_over_ranks(gather_last(stack))
            vinfo["gather_ms"] = gms
            if env.rank == 0:
                full_img = assemble(g_img).cpu().numpy()
                full_msk = assemble(g_msk).cpu().numpy() if g_msk is not None else None
                # rank 0's own single-GPU result of the whole image
                s1 = cp.FrameStack(ctx, W, H_img, 3, n)
                s1.fill_synthetic(kind, seed=42, row0_global=0, full_height=H_img)
                p1 = make_processor(cp, mode)
                if is_outlier:
                    img1, msk1 = p1.process(s1)
                    diff = int((img1 != full_img).any(axis=2).sum()) + int((msk1 != full_msk).any(axis=2).sum())
                else:
                    diff = int((p1.process(s1) != full_img).any(axis=2).sum())
                s1.close()
                vinfo["gathered_vs_single_gpu_pixels_differing"] = diff
        elif not env.multi_proc:
            if is_outlier:
                full_img, full_msk = proc.process(stack)
            else:
                full_img = proc.process(stack)
        if env.rank == 0 and full_img is not None:
            orc = oracle()
            vrows, bad, where = verify_photo(orc, np, mode, kind, n, H_img, W, full_img, full_msk)
            if where:
                vinfo["first_differences"] = where
            verified = dict(vinfo, rows=vrows, pixels_differing_from_oracle=bad, configs=[wl],
                            ok=(bad == 0 and vinfo.get("gathered_vs_single_gpu_pixels_differing", 0) == 0),
                            what="composite" + (" + mask" if is_outlier else "") + " of the timed stack, bit for bit against the CPU oracle on the same synthetic frames")
            if not verified["ok"]:
                log(f"VERIFICATION FAILED for {wl}: {verified}")

    # ---- end to end through the C ABI with pinned host frames
    e2e = None
    if do_e2e:
        import psutil
        e_stack, e_rows, e_row0, e_H, e_pf, e_note = stack, rows, row0, H, float(n) * H * W, None
        avail = env.min_over_ranks(float(psutil.virtual_memory().available))
        ranks_here = env.world if env.multi_proc else 1
        need = 1.3 * ranks_here * n * rows * W * 3
        if avail < need:
            e2e = {"value": None, "note": f"skipped: {need / 1e9:.0f} GB of pinned host frames needed, {avail / 1e9:.0f} GB of host RAM available"}
        else:
            fb = e_rows * W * 3
            host = shared.get(("host", key)) if shared else None
            if host is None:
                host = torch.empty((n, e_rows, W, 3), dtype=torch.uint8, pin_memory=True)
                for f in range(n):
                    e_stack.download_raw(f, host.data_ptr() + f * fb, W * 3)  # host copy of the synthetic series
                if shared is not None:
                    shared[("host", key)] = host
            base, pitch = host.data_ptr(), W * 3
            gather = env.multi_proc and d_img is not None
            full_rows = H_img if gather else e_rows
            out_t = torch.empty((full_rows, W, 3), dtype=torch.uint8, pin_memory=True)
            msk_t = torch.empty((full_rows, W, 3), dtype=torch.uint8, pin_memory=True)
            out_np, msk_np = out_t.numpy(), msk_t.numpy()
            stack2 = cp.FrameStack(ctx, W, e_rows, 3, n)
            gather_ms = []

            def e2e_step():
                for f in range(n):
                    stack2.upload_raw(f, base + f * fb, pitch, pinned=True)
                if gather:  # kernels, then the band gather over NCCL, then rank 0's D2H of the whole image
                    if is_outlier:
                        proc.process_device(stack2)
                    else:
                        proc.process_device(stack2)
                    gather_ms.append(gather_last(stack2))
                    if env.rank == 0:
                        out_t.copy_(assemble(g_img), non_blocking=True)
                        if g_msk is not None:
                            msk_t.copy_(assemble(g_msk), non_blocking=True)
                        torch.cuda.synchronize()
                elif is_outlier:
                    proc.process(stack2, out=out_np, mask_out=msk_np)
                else:
                    _lib.check(_lib.lib().chb_simple(stack2._h, proc._params(), None, 0, out_np.ctypes.data))

            e2e_step()  # warm-up (allocations, first touch)
            env.barrier()
            gather_ms.clear()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                e2e_step()
            env.barrier()
            e2e_s = env.max_over_ranks((time.perf_counter() - t0) / e2e_steps)
            d2h = H_img * W * 3 * (2 if is_outlier else 1) if gather else fb * (2 if is_outlier else 1) * ranks_here
            e2e = {"value": e_pf / e2e_s, "unit": "pixel-frames/s", "h2d_bytes_per_step": n * fb * ranks_here, "d2h_bytes_per_step": d2h,
                   "ms_per_step": e2e_s * 1e3, "steps": e2e_steps,
                   "path": "pinned host frames -> chb_stack_upload_pinned x %d -> %s -> %s" % (
                       n, "chb_outlier_device" if (gather and is_outlier) else ("chb_simple_device" if gather else ("chb_outlier" if is_outlier else "chb_simple")),
                       "NCCL gather of the bands onto rank 0 -> D2H of the whole image" if gather else "D2H of composite" + (" + mask" if is_outlier else ""))}
            if gather:
                e2e["gather_ms"] = env.max_over_ranks(sum(gather_ms) / max(1, len(gather_ms)))
            # the same step from PAGEABLE frames uploaded by 8 host threads (what decode threads hand over: chb_stack_upload stages
            # every frame through a pool of pinned slots, the memcpy into the slot runs in the calling thread)
            if primary and not env.multi_proc and psutil.virtual_memory().available > 1.5 * n * fb:
                from concurrent.futures import ThreadPoolExecutor
                pageable = host.numpy().copy()
                pbase = pageable.ctypes.data
                n_thr = 8

                def pageable_step():
                    with ThreadPoolExecutor(n_thr) as ex:
                        list(ex.map(lambda f: stack2.upload_raw(f, pbase + f * fb, pitch, pinned=False), range(n)))
                    if is_outlier:
                        proc.process(stack2, out=out_np, mask_out=msk_np)
                    else:
                        _lib.check(_lib.lib().chb_simple(stack2._h, proc._params(), None, 0, out_np.ctypes.data))

                pageable_step()
                t0 = time.perf_counter()
                pageable_step()
                dt = time.perf_counter() - t0
                e2e["pageable"] = {"value": e_pf / dt, "unit": "pixel-frames/s", "ms_per_step": dt * 1e3, "uploader_threads": n_thr,
                                   "h2d_gbs": n * fb / dt / 1e9, "path": "pageable host frames -> chb_stack_upload from 8 threads -> kernels -> D2H"}
                del pageable
                # JPEG in: the frames cross PCIe compressed and are decoded on the device (chb_stack_upload_jpeg: nvJPEG, Huffman
                # stage in the calling thread) -- 32 frames (two frame groups) from 8 decode threads, encoded outside the timed region
                try:
                    n_j = min(32, n)
                    jpegs = [cp.encode_jpeg(ctx, host[f].numpy(), quality=90) for f in range(n_j)]
                    stack3 = cp.FrameStack(ctx, W, e_rows, 3, n_j)

                    def jpeg_step():
                        with ThreadPoolExecutor(n_thr) as ex:
                            list(ex.map(lambda f: stack3.upload_jpeg(f, jpegs[f]), range(n_j)))
                        stack3.sync()

                    jpeg_step()
                    t0 = time.perf_counter()
                    jpeg_step()
                    dt = time.perf_counter() - t0
                    e2e["jpeg_ingest"] = {"value": float(n_j) * e_rows * W / dt, "unit": "pixel-frames/s", "frames": n_j, "decode_threads": n_thr,
                                          "ms_per_frame": dt * 1e3 / n_j, "compressed_mb_per_frame": sum(len(j) for j in jpegs) / n_j / 1e6,
                                          "path": "JPEG bytes (quality 90, 4:2:0) -> chb_stack_upload_jpeg from 8 threads (nvJPEG decode on the device) -> re-layout"}
                    stack3.close()
                    del jpegs
                except Exception as ex:  # the JPEG leg must not take the line with it
                    e2e["jpeg_ingest"] = {"error": f"{type(ex).__name__}: {ex}"}
            if e_note:
                e2e["note"] = e_note
            stack2.close()
            if shared is None:
                host = None

    # ---- CPU baseline beside it (rank 0, N = 1): oracle port on a bounded sample of the same workload
    cpu_baseline = None
    if env.rank == 0 and n_gpus == 1 and do_cpu:
        cpu_baseline = cpu_baseline_block(oracle(), mode, kind, n, H_img, W, cpu_target_s, with_reference_threading=is_outlier and primary)

    cfg = config_of(wl)
    cfg_run = {"sharding": (f"{n_gpus} bands of {rows} rows: a {W}x{H} image, every GPU owns one full-size band of the workload" if weak
                            else (f"{W}x{H_img} image cut into {n_gpus} shards of {rows} rows: GPU g owns the blocks of {ishard.block_rows} rows with index = g mod {n_gpus}"
                                  if ishard else f"{W}x{H_img} image cut into {n_gpus} band(s) of {rows} rows")),
               "launcher": "torchrun" if env.multi_proc else "single-process", "stack_gb_per_gpu": stack.device_bytes(0) / 1e9,
               "cpus_bound_per_rank": env.cpu_affinity}
    out = {"metric": "pixel-frames/s", "value": value, "unit": "pixel-frames/s", "n_gpus": n_gpus, "steps": steps, "warmup": warmup,
           "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak" if (weak or n_gpus == 1) else "strong", "vs_baseline": None,
           "dtype": "u8", "data": "synthetic", "config": cfg, "run": cfg_run,
           "hbm_gbs": achieved * n_gpus if n_gpus > 1 else achieved,
           "roofline": roofline, "cpu_baseline": cpu_baseline, "e2e": e2e, "verified": verified, "gpu_launches": launches, "clocks": clocks}
    if other_scaling:
        out[other_scaling[0]] = other_scaling[1]
    if shared is None:
        stack.close()
        if own_ctx:
            ctx.close()
    return out


def run_video(env, wl, steps, warmup, do_e2e, do_cpu, cpu_target_s, do_verify, shared=None):
    """Config 5: sliding-window compositing over a resident clip; a step = the whole video (all windows)."""
    import numpy as np
    import chrono_photo_b200 as cp
    from chrono_photo_b200 import _lib
    torch = env.torch
    mode, n, H, W, kind, desc = WORKLOADS[wl]
    if shared:  # free the photo stacks first
        for k in [k for k in shared if isinstance(k, tuple) and k[0] in ("stack", "host")]:
            v = shared.pop(k)
            if k[0] == "stack":
                v.close()
    ctx = shared.get("ctx") if shared else None
    own_ctx = ctx is None
    if own_ctx:
        ctx = cp.Context([env.devices[0]])
        ctx.set_stream(0, torch.cuda.current_stream().cuda_stream)
    stack = cp.FrameStack(ctx, W, H, 3, n)
    stack.fill_synthetic(kind, seed=42)
    proc = make_processor(cp, "outlier")
    wins = cp.video_windows(n, cp.FrameRange(0, 25, 1), cp.FrameRange.empty())
    total_pf = float(sum(len(idx) for _, idx in wins)) * H * W
    alg_bytes = sum((len(idx) + 2) for _, idx in wins) * H * W * 3
    runs = proc.video_runs(wins)  # maximal runs of equal-length windows sliding by one frame -> one chb_outlier_video call each

    def step():
        ms = 0.0
        for pos, count in runs:
            idx = wins[pos][1]
            if count > 1:
                ms += proc.process_video_run_device(stack, idx[0], len(idx), count)
            else:
                ms += proc.process_device(stack, idx)
        return ms

    n_warm, n_steps = max(1, warmup // 3), max(1, steps // 5)
    for _ in range(n_warm):
        step()
    torch.cuda.synchronize()
    _lib.lib().chb_launch_count_reset()
    sampler = ClockSampler(env.devices[0])
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    kms = [step() for _ in range(n_steps)]
    ev1.record()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    launches = int(_lib.lib().chb_launch_count())
    ms_step = ev0.elapsed_time(ev1) / n_steps
    kernel_ms = sum(kms) / len(kms)
    peak, peak_src = peaks()
    achieved = alg_bytes / (kernel_ms / 1e3) / 1e9

    # ---- verification: one block of 16 windows on a row band against the oracle
    verified = None
    if do_verify:
        orc = oracle()
        w0, r0, r = 400, H // 2 - 16, 32
        imgs, masks, _ = proc.process_video_run(stack, w0, 25, 16)
        sub = orc.synth_frames(kind, 42, n, W, H, rows=r, row0=r0, f0=w0, n_out=40)
        bad = 0
        for k in range(16):
            oimg, omsk, _ = oracle_call(orc, "video", sub[k:k + 25], os.cpu_count() or 1)
            bad += int((oimg != imgs[k][r0:r0 + r]).any(axis=2).sum()) + int((omsk != masks[k][r0:r0 + r]).any(axis=2).sum())
        verified = {"rows": r, "windows": 16, "pixels_differing_from_oracle": bad, "ok": bad == 0, "configs": [wl],
                    "what": f"composite + mask of windows {w0}..{w0 + 15} of the timed clip, rows [{r0}, +{r}), bit for bit against the CPU oracle"}
        del imgs, masks

    # ---- end to end: shaky 1920x1080 host frames -> upload through the Crop origins -> every window -> host frames
    e2e = None
    if do_e2e:
        import psutil
        FW, FH = 1920, 1080
        need = 1.2 * n * FW * FH * 3 + 2e9
        if psutil.virtual_memory().available < need:
            e2e = {"value": None, "note": f"skipped: {need / 1e9:.0f} GB of pinned host memory needed"}
        else:
            rng = np.random.default_rng(43)
            offs = rng.integers(-8, 9, size=(n, 2)).astype(np.int32)
            offs[0] = (0, 0); offs[1] = (-8, -8); offs[2] = (8, 8)  # the crop then is exactly 1904x1064 (Crop::create, src/shake.rs:136-176)
            origins, cw, ch = cp.crop_create(offs, FW, FH)
            assert (cw, ch) == (W, H), (cw, ch)
            host = torch.empty((n, FH, FW, 3), dtype=torch.uint8, pin_memory=True)
            fbytes, pitch = FH * FW * 3, FW * 3
            for f in range(n):  # the resident synthetic clip written back into shaky full-size host frames
                ox, oy = int(origins[f][0]), int(origins[f][1])
                stack.download_raw(f, host.data_ptr() + f * fbytes + oy * pitch + ox * 3, pitch)
            stack2 = cp.FrameStack(ctx, W, H, 3, n)
            chunk = 128
            out_t = torch.empty((chunk, H, W, 3), dtype=torch.uint8, pin_memory=True)
            out_np = out_t.numpy()

            def e2e_step():
                for f in range(n):
                    stack2.upload_raw(f, host.data_ptr() + f * fbytes, pitch, crop_xy=(int(origins[f][0]), int(origins[f][1])), pinned=True)
                for pos, count in runs:
                    idx = wins[pos][1]
                    if count > 1:
                        for c0 in range(0, count, chunk):
                            c = min(chunk, count - c0)
                            proc.process_video_run(stack2, idx[0] + c0, len(idx), c, want_mask=False, out=out_np[:c])
                    else:
                        proc.process(stack2, idx, want_mask=False, out=out_np[0])

            e2e_step()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e_step()
            torch.cuda.synchronize()
            e2e_s = time.perf_counter() - t0
            e2e = {"value": total_pf / e2e_s, "unit": "pixel-frames/s", "h2d_bytes_per_step": n * H * W * 3, "d2h_bytes_per_step": len(wins) * H * W * 3,
                   "ms_per_step": e2e_s * 1e3, "steps": 1,
                   "path": f"pinned {FW}x{FH} host frames -> chb_stack_upload_pinned with the Crop::create origin of every frame (shake offsets in [-8,8]^2) -> "
                           "chb_outlier_video runs / chb_outlier -> D2H of every output frame (masks are not requested: create_video discards them)"}
            stack2.close()
            host = None

    cpu_baseline = cpu_baseline_block(oracle(), "video", kind, n, H, W, cpu_target_s, False) if do_cpu else None
    out = {"metric": "pixel-frames/s", "value": total_pf / (ms_step / 1e3), "unit": "pixel-frames/s", "n_gpus": 1, "steps": n_steps, "warmup": n_warm,
           "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
           "config": dict(config_of(wl), windows=len(wins)),
           "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": None, "kernel": "video_kernel",
                        "algorithmic_bytes_per_window": alg_bytes / len(wins), "avg_ms_per_window": kernel_ms / len(wins), "peak_source": peak_src,
                        "runs": [[len(wins[p][1]), c] for p, c in runs if c > 1], "single_window_launches": sum(1 for _, c in runs if c == 1),
                        "note": "algorithmic bytes = what the reference's per-frame loop reads and writes (every window re-read); the sliding kernel "
                                "loads each frame group once per 16 windows; kernel time = sum over the launches"},
           "cpu_baseline": cpu_baseline, "e2e": e2e, "verified": verified, "gpu_launches": launches, "clocks": clocks}
    stack.close()
    if own_ctx:
        ctx.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=None, choices=sorted(WORKLOADS))
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-others", action="store_true", help="default run at N = 1: skip the other_workloads blocks")
    ap.add_argument("--scaling", default="strong", choices=["weak", "strong"],
                    help="N > 1: strong (default) = the workload's image row-sharded H/N; weak = every rank owns a full-size band of an N-times taller image")
    args = ap.parse_args()
    default_run = args.workload is None
    wl = args.workload or DEFAULT_WORKLOAD
    if args.impl == "reference":
        return run_reference(args, wl)
    env = Env(args)
    shared = {} if (default_run and env.n_gpus == 1 and not args.no_others) else None
    if WORKLOADS[wl][0] == "video":
        line = run_video(env, wl, args.steps, args.warmup, not args.no_e2e, not args.no_cpu and env.rank == 0, 8.0, not args.no_verify)
    else:
        line = run_photo(env, wl, args.steps, args.warmup, args.e2e_steps, not args.no_e2e, not args.no_cpu, 10.0, not args.no_verify,
                         args.scaling, primary=True, shared=shared)
    if shared is not None:
        others, skipped = [], []
        for owl in OTHER_WORKLOADS:
            if time.perf_counter() - T_START > OTHERS_BUDGET_S:
                skipped.append(owl)
                continue
            log(f"other workload {owl}")
            try:
                if WORKLOADS[owl][0] == "video":
                    o = run_video(env, owl, 5, 3, not args.no_e2e, not args.no_cpu, 3.0, not args.no_verify, shared=shared)
                else:
                    o = run_photo(env, owl, max(3, args.steps // 2), 3, 1, not args.no_e2e, not args.no_cpu, 3.0, not args.no_verify, "strong",
                                  primary=False, shared=shared)
                others.append({k: o[k] for k in ("config", "value", "unit", "ms_per_step", "steps", "warmup", "roofline", "cpu_baseline", "e2e", "verified",
                                                  "gpu_launches")})
            except Exception as e:  # a failing side workload must not take the headline line with it
                others.append({"config": config_of(owl), "error": f"{type(e).__name__}: {e}"})
        line["other_workloads"] = others
        if skipped:
            line["other_workloads_skipped"] = {"workloads": skipped, "why": f"run older than {OTHERS_BUDGET_S:.0f} s"}
        configs, rows = [], 0
        for o in [line] + others:
            v = o.get("verified")
            if v and v.get("ok"):
                configs += v["configs"]
                rows += v["rows"]
        if line.get("verified"):
            line["verified"] = dict(line["verified"], configs=configs, rows_total=rows)
        for k in [k for k in shared if isinstance(k, tuple) and k[0] == "stack"]:
            shared[k].close()
        if "ctx" in shared:
            shared["ctx"].close()
    if env.rank == 0:
        print(json.dumps(line))
    if env.multi_proc:
        env.dist.destroy_process_group()


if __name__ == "__main__":
    main()
