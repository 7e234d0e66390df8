#!/usr/bin/env python
"""bench.py -- limg encode/decode hot path on B200 (contract in the task statement, section 4).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

A step = one blocked encode (compact stream out: area table + three code planes) followed by one standalone decode of
that stream, on one synthetic frame of the workload (default: BASELINE.json configs[1], one 3840x2160 RGB photo-like
frame). With N > 1 (torchrun, one rank per GPU) every rank processes its own frame of the same shape (independent units,
no data-path collective, "weak" scaling); value = frames * pixels of all ranks / max-over-ranks device time.

  value      device-resident throughput (inputs in HBM before the timed region), CUDA events on the codec's stream
  e2e        same step through the host-buffer C ABI entry points with pinned host buffers: H2D of the frame, D2H of the
             stream, H2D of the stream, D2H of the decoded frame -- all inside the timed region
  roofline   encode path at SURVEY.md 8(d)'s 7 algorithmic bytes per pixel against MEASURED_PEAKS.json's HBM copy bandwidth,
             per-kernel-phase durations from CUDA events on the launching stream (limgcu phase timing), dominant kernel named
  cpu_baseline  the reference itself (oracle/_ref/libref.so, kind "reference") or the C port (oracle/, kind "port") on the
             box's host cores, bounded sample, rank 0 at N == 1 only

--impl reference times the reference's own CPU implementation (limg_blocked_encode3d_test, all host threads through its thread
pool) on the same workload and prints the same line with "impl": "reference".
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

import numpy as np  # noqa: E402

METRIC = "encode+decode throughput"
UNIT = "Mpixel/s"

WORKLOADS = {
    # name: (width, height, has_alpha, generator kind)
    "c2_4k_photo": (3840, 2160, False, "photo"),
    "c4_4k_flatui": (3840, 2160, False, "flatui"),
    "c5_1080p_photo": (1920, 1080, False, "photo"),
    "c1_512_gradient": (512, 512, False, "gradient"),
    "c3_8k_rgba": (7680, 4320, True, "photo"),
}


def make_frame(workload: str, index: int) -> np.ndarray:
    from limg_b200 import synth
    w, h, alpha, kind = WORKLOADS[workload]
    if kind == "photo":
        seed = {"c2_4k_photo": 1, "c5_1080p_photo": 3, "c3_8k_rgba": 4}[workload] + index
        return synth.photo_like(w, h, seed, 4 if alpha else 3)
    if kind == "flatui":
        return synth.flat_ui(w, h, 2 + index)
    return synth.gradient_noise(w, h, 1234 + index)


class stdout_to_stderr:
    """NCCL prints its version banner on stdout when the communicator is created; rank 0's stdout must carry the JSON line only."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


def init_distributed(dev):
    import torch
    import torch.distributed as dist
    with stdout_to_stderr():
        dist.init_process_group("nccl", device_id=dev)
        dist.barrier()  # creates the communicator (and prints the banner) now
        torch.cuda.synchronize(dev)


PHASE_KERNELS = {"pass1": "k_pass1", "predicate_windows": "k_pred_records + k_pred_window + k_plan_seeds", "merge_scan": "k_merge_cta (+ k_merge_verify, k_prepare_*)",
                 "area_encode": "k_encode_small + k_encode_large", "dither_scan": "k_dither_scan + k_dither_states", "finalize": "k_finalize"}


def ncu_traffic(workload: str, phase: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the phase's main kernel, from the ncu --set full capture recorded in
    profiles/ncu_traffic.json ({workload: {phase: {"bytes": ..., "source": "profiles/<file>"}}}); None when no capture of this build exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)[workload][phase]
    except Exception:
        return None


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock + throttle reasons during the timed region (B200_PROFILING.md's clocks line), sampled in-process through NVML
    every few ms (the timed region of a short run is shorter than nvidia-smi's start-up); nvidia-smi -lms is the fallback."""

    FIELDS = "index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []
        self.sm, self.mx, self.reasons = [], [], set()
        self.stop_flag = threading.Event()
        self.thread = None
        self.source = None

    def _nvml_handle(self):
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        index = self.gpu_index
        if visible:
            entry = visible.split(",")[self.gpu_index].strip()
            if entry.isdigit():
                index = int(entry)
            else:
                return pynvml, pynvml.nvmlDeviceGetHandleByUUID(entry)
        return pynvml, pynvml.nvmlDeviceGetHandleByIndex(index)

    def _poll_nvml(self, pynvml, handle):
        while True:
            try:
                self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(handle, pynvml.NVML_CLOCK_SM)))
                mask = int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(handle))
                for name, bit in self.REASONS:
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            if self.stop_flag.wait(0.004):
                break

    def start(self):
        try:
            pynvml, handle = self._nvml_handle()
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(handle, pynvml.NVML_CLOCK_SM)))
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll_nvml, args=(pynvml, handle), daemon=True)
            self.thread.start()
            return
        except Exception:
            self.source = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if self.source is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml and nvidia-smi unavailable"], "samples": 0}
        if self.source == "nvml":
            self.stop_flag.set()
            self.thread.join(timeout=2)
        else:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except Exception:
                self.proc.kill()
            for line in self.lines:
                p = [x.strip() for x in line.split(",")]
                if len(p) < 9:
                    continue
                try:
                    self.sm.append(float(p[1])); self.mx.append(float(p[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)
        sm, mx = self.sm, self.mx
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(self.reasons), "samples": len(sm), "source": self.source}


def common_config(workload: str) -> dict:
    """The keys both arms print under "config" (identical for `ours` and `--impl reference`); everything arm-specific goes under "notes"."""
    w, h, alpha, _ = WORKLOADS[workload]
    return {"workload": workload, "width": w, "height": h, "channels": 4 if alpha else 3, "error_factor": 100, "fast_bit_crushing": True}


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_baseline(workload: str, threads: int, budget_s: float = 20.0) -> dict:
    """Times the reference (or the C port) on the host. Sample: whole frames of the workload, as many as fit ~budget_s."""
    from oracle import ref
    w, h, alpha, _ = WORKLOADS[workload]
    img = make_frame(workload, 0)
    mpx = w * h / 1e6
    if ref.available():
        ref.set_modes(True, False)
        t = ref.time_blocked(img, alpha, 100, True, threads, 1, False)  # one frame, also the warm-up
        reps = max(1, min(8, int(budget_s / max(t, 1e-3)) - 1))
        t = ref.time_blocked(img, alpha, 100, True, threads, reps, False)
        out = {"value": mpx / t, "unit": UNIT, "cores": threads, "kind": "reference",
               "sample": "%d x limg_blocked_encode3d_test (encode + in-encoder decode) of one %dx%d frame, %d-thread limg_thread_pool (only pass 1 of the reference "
                         "uses the pool; the merge, refit and bit-crush search run on one core), LCG dither" % (reps, w, h, threads),
               "seconds_per_frame": t}
        try:
            out["frame_parallel"] = cpu_frame_parallel(workload, max(1, min(threads, 32)))
        except Exception as e:
            out["frame_parallel"] = {"value": None, "unit": UNIT, "processes": 0, "sample": repr(e)}
        # the reference's only fully threaded path, limg_encode3d_test_perf (every 8x8 block its own area, no planes written)
        tp = ref.time_blocked(img, alpha, 100, True, threads, 1, True)
        tp = ref.time_blocked(img, alpha, 100, True, threads, max(1, min(8, int(5.0 / max(tp, 1e-3)))), True)
        t1 = ref.time_blocked(img, alpha, 100, True, 0, 1, True)
        out["unmerged_perf_path"] = {"value": mpx / tp, "unit": UNIT, "cores": threads, "single_thread_value": mpx / t1,
                                     "sample": "limg_encode3d_test_perf of the same frame, %d-thread pool and pool-less" % threads}
        return out
    from oracle import oracle as lo
    crop = img[: min(h, 1080), : min(w, 1920)]
    t0 = time.time()
    lo.blocked_encode3d(np.ascontiguousarray(crop), alpha, 100, True)
    t = time.time() - t0
    return {"value": crop.size / 1e6 / t, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": "1 x oracle lo_blocked_encode3d on a %dx%d crop (scalar C port, 1 thread)" % (crop.shape[1], crop.shape[0]), "seconds_per_frame": t}


def _frame_parallel_worker(job):
    """one reference process: limg_blocked_encode3d_test of one frame, pool-less (the reference's merge path uses one core after pass 1 anyway)"""
    workload, index = job
    from oracle import ref
    w, h, alpha, _ = WORKLOADS[workload]
    img = make_frame(workload, index % 4)
    ref.set_modes(True, False)
    t0 = time.time()
    ref.time_blocked(img, alpha, 100, True, 0, 1, False)
    return t0, time.time()


def cpu_frame_parallel(workload: str, procs: int) -> dict:
    """What the CPU box does on a BATCH of frames: `procs` reference processes side by side, one frame each (the reference's own thread pool
    only helps pass 1; frames are independent). Aggregate Mpixel/s from the first start to the last end."""
    import multiprocessing as mp
    from oracle import ref
    if not ref.available():
        return {"value": None, "unit": UNIT, "processes": 0, "sample": "oracle/_ref/libref.so not present"}
    w, h, alpha, _ = WORKLOADS[workload]
    with mp.get_context("spawn").Pool(procs) as pool:
        pool.map(_frame_parallel_worker, [(workload, i) for i in range(procs)])          # start-up (imports, page faults) outside the sample
        spans = pool.map(_frame_parallel_worker, [(workload, i) for i in range(procs)], chunksize=1)
    dt = max(e for _, e in spans) - min(b for b, _ in spans)
    return {"value": procs * w * h / 1e6 / dt, "unit": UNIT, "processes": procs,
            "sample": "%d processes x 1 frame of %dx%d through limg_blocked_encode3d_test (pool-less each), wall clock first start to last end" % (procs, w, h)}


def run_reference(args, rank: int, world: int):
    """--impl reference: the reference's CPU path on the same workload; rank 0 only."""
    if rank != 0:
        return
    from oracle import ref
    w, h, alpha, _ = WORKLOADS[args.workload]
    mpx = w * h / 1e6
    threads = host_cores()
    img = make_frame(args.workload, 0)
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
            "config": common_config(args.workload),
            "notes": {"step": "limg_blocked_encode3d_test: encode + in-encoder decode of one frame, all 13 API planes written (the reference has no standalone decoder, no stream)",
                      "dither": "lcg"}}
    if ref.available():
        ref.set_modes(True, False)
        ref.time_blocked(img, alpha, 100, True, threads, max(1, min(args.warmup, 1)), False)
        times = [ref.time_blocked(img, alpha, 100, True, threads, 1, False) for _ in range(args.steps)]
        kind, cores = "reference", threads
        sample = "%d steps, each one full %dx%d frame through limg_blocked_encode3d_test with a %d-thread limg_thread_pool (oracle/_ref/libref.so, LCG dither)" % (args.steps, w, h, threads)
    else:
        from oracle import oracle as lo
        crop = np.ascontiguousarray(img[: min(h, 1080), : min(w, 1920)])
        mpx = crop.size / 1e6
        times = []
        for _ in range(max(1, min(args.steps, 3))):
            t0 = time.time(); lo.blocked_encode3d(crop, alpha, 100, True); times.append(time.time() - t0)
        kind, cores = "port", 1
        sample = "%d steps, each a %dx%d crop through the scalar C port (oracle/liblimg_oracle.so); libref.so not present" % (len(times), crop.shape[1], crop.shape[0])
    t = sum(times) / len(times)
    value = mpx / t
    line.update({"value": value, "ms_per_step": t * 1e3,
                 "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
                 "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    print(json.dumps(line), flush=True)


DROPIN_PLANES_U32 = ("pDecoded", "pShiftABCX", "pColAMin", "pColAMax", "pColBMin", "pColBMax", "pColCMin", "pColCMax", "pBlockIndex")
DROPIN_PLANES_U8 = ("pFactorsA", "pFactorsB", "pFactorsC", "pBitsPerPixel")
DROPIN_ORDER = ("pDecoded", "pFactorsA", "pFactorsB", "pFactorsC", "pBlockError", "pBitsPerPixel", "pShiftABCX", "pColAMin", "pColAMax", "pColBMin", "pColBMax", "pColCMin",
                "pColCMax", "pBlockIndex")  # limg_blocked_encode3d_info, limg.h:39-44


def e2e_dropin(lib, frame: np.ndarray, alpha: bool, device: int, steps: int) -> dict:
    """What a user of the reference gets after swapping the library: the reference's own C++ entry point limg_blocked_encode3d_test
    (limg.h:46, main.cpp:255) through the drop-in (limg_api.cpp), host source in, ALL 13 host planes out (40 B/px device-to-host; pBlockError is
    written by neither implementation), wall clock per blocking call. Both dither generators: the drop-in picks like the reference does (AES
    rounds on a host with AES-NI: walked on the host, the call synchronises in the middle), LIMGCU_DITHER=lcg selects the device-side LCG."""
    import ctypes as C
    import torch
    h, w = frame.shape
    fn = getattr(lib, "_Z26limg_blocked_encode3d_testPKjmmbP26limg_blocked_encode3d_infojP16limg_thread_poolb")
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_bool, C.c_void_p, C.c_uint32, C.c_void_p, C.c_bool]
    set_device = getattr(lib, "_Z20limg_b200_set_devicei")
    set_device.restype = C.c_int
    set_dither = getattr(lib, "_Z25limg_b200_set_dither_modei")
    set_dither.restype = C.c_int
    lib.limgcu_host_has_aesni.restype = C.c_int

    class Info(C.Structure):
        _fields_ = [(k, C.c_void_p) for k in DROPIN_ORDER]

    src = torch.from_numpy(frame.view(np.int32)).pin_memory()
    planes = {k: torch.empty((h, w), dtype=torch.int32).pin_memory() for k in DROPIN_PLANES_U32}
    planes.update({k: torch.empty((h, w), dtype=torch.uint8).pin_memory() for k in DROPIN_PLANES_U8})
    info = Info(**{k: planes[k].data_ptr() for k in planes})
    assert set_device(device) == 0
    out = {"entry_point": "limg_blocked_encode3d_test (C++ symbol of include/limg_dropin.h), 13 host planes, pinned host buffers",
           "h2d_bytes_per_step": 4 * h * w, "d2h_bytes_per_step": (4 * len(DROPIN_PLANES_U32) + len(DROPIN_PLANES_U8)) * h * w,
           "host_has_aesni": bool(lib.limgcu_host_has_aesni())}
    for name, mode in (("lcg", 0), ("aes", 1)):
        set_dither(mode)
        for _ in range(2):
            assert fn(src.data_ptr(), w, h, alpha, C.byref(info), 100, None, True) == 0
        t0 = time.perf_counter()
        for _ in range(steps):
            assert fn(src.data_ptr(), w, h, alpha, C.byref(info), 100, None, True) == 0
        dt = (time.perf_counter() - t0) / steps
        out[name] = {"ms_per_call": dt * 1e3, "value": h * w / 1e6 / dt, "unit": "Mpixel/s (encode + in-encoder decode, as the reference's step)"}
    picked = "aes" if set_dither(-1) == 1 else "lcg"
    out["default_on_this_host"] = picked
    out["value"] = out[picked]["value"]
    out["unit"] = "Mpixel/s"
    return out


def north_star_8k_rgb(codec, stream, dev, d_flush, peak: float, iters: int = 5) -> dict:
    """The configuration BASELINE.json's targets are quoted on (8K RGB: encode >= 50 %, decode >= 70 % of the HBM roofline), measured
    next to the headline workload: device-resident, CUDA events on the codec's stream, L2 flushed between iterations, medians."""
    import torch
    from limg_b200 import AREA_DTYPE, synth
    w, h = 7680, 4320
    frame = synth.photo_like(w, h, 1, 3)
    bx, by = w // 8, h // 8
    d_src = torch.from_numpy(frame.view(np.int32)).to(dev)
    d_codes = [torch.empty((h, w), dtype=torch.uint8, device=dev) for _ in range(3)]
    d_areas = torch.empty(bx * by * AREA_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_map = torch.empty(bx * by, dtype=torch.int32, device=dev)
    d_count = torch.zeros(1, dtype=torch.int32, device=dev)
    d_dec = torch.empty((h, w), dtype=torch.int32, device=dev)
    st = {"areas": d_areas.data_ptr(), "area_count": d_count.data_ptr(), "block_to_area": d_map.data_ptr(),
          "codesA": d_codes[0].data_ptr(), "codesB": d_codes[1].data_ptr(), "codesC": d_codes[2].data_ptr()}
    ev = []
    with torch.cuda.stream(stream):
        for i in range(iters + 2):
            d_flush.fill_(i & 0xFF)
            e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
            e[0].record(stream)
            codec.blocked_encode3d_device(d_src.data_ptr(), w, h, False, 100, True, False, st, None)
            e[1].record(stream)
            codec.decode_device(d_areas.data_ptr(), d_map.data_ptr(), d_codes[0].data_ptr(), d_codes[1].data_ptr(), d_codes[2].data_ptr(), w, h, False, d_dec.data_ptr())
            e[2].record(stream)
            ev.append(e)
    codec.sync()
    enc = statistics.median(e[0].elapsed_time(e[1]) for e in ev[2:])
    dec = statistics.median(e[1].elapsed_time(e[2]) for e in ev[2:])
    psnr, _, _ = codec.compare_device(d_src.data_ptr(), d_dec.data_ptr(), w, h, False)
    gbs = lambda ms: 7.0 * w * h / (ms * 1e-3) / 1e9
    return {"workload": "7680x4320 RGB photo-like synthetic (seed 1), one frame", "encode_ms": enc, "decode_ms": dec,
            "encode_mpixel_s": w * h / 1e6 / (enc * 1e-3), "decode_mpixel_s": w * h / 1e6 / (dec * 1e-3),
            "encode_hbm_frac": gbs(enc) / peak, "decode_hbm_frac": gbs(dec) / peak, "peak_gbs": peak,
            "decode_traffic": (ncu_traffic("north_star_8k_rgb", "decode") or {}).get("bytes"), "decode_traffic_source": (ncu_traffic("north_star_8k_rgb", "decode") or {}).get("source"),
            "decode_algorithmic_bytes": 7 * w * h, "bytes_per_px": 7, "psnr_db": psnr, "iters": iters}


def batch_1080p(local_rank: int, dev, lanes: int = 8, frames: int = 32, steps: int = 3) -> dict:
    """BASELINE.json config 5 on one GPU, a short form of --mode batch: `frames` 1920x1080 frames per step on `lanes` contexts (batch_core)."""
    import torch
    r = batch_core(local_rank, dev, frames, 0, lanes, steps, 3, lambda: torch.cuda.synchronize(dev))
    px = r["pixels"] * steps / 1e6
    return {"workload": "%d x 1920x1080 RGB photo-like frames per step, %d lanes (contexts) on one GPU" % (frames, r["lanes"]), "lanes": r["lanes"], "ms_per_step": r["dev_ms"] / steps,
            "value": px / (r["dev_ms"] / 1e3), "unit": UNIT,
            "e2e": {"value": px / (r["e2e_ms"] / 1e3), "unit": UNIT, "frames": frames,
                    "path": "limgcu_host_encode_container + limgcu_host_decode_container per frame, pinned host buffers, one host thread per lane"},
            "round_trip_equals_device_path": bool(r["ok"])}


def batch_core(local_rank: int, dev, n_frames: int, pool_seed: int, lanes: int, steps: int, warmup: int, barrier) -> dict:
    """One rank's part of a batch of 1920x1080 RGB frames: `lanes` contexts (limgcu_create each: own streams and scratch; every frame's area scan is
    one thread-block cluster, the throughput kernels of the other frames fill the rest of the GPU), frame k on lane k % lanes.
      device-resident: frames in HBM, streams out to HBM; CUDA events from a common start to the last lane's end, per step
      end to end:      limgcu_host_encode_container + limgcu_host_decode_container per frame with PINNED host buffers, one host thread per lane
                       (the copies of one lane overlap the kernels of the others), wall clock"""
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    import torch
    from limg_b200 import AREA_DTYPE, Codec, synth
    w, h, alpha = 1920, 1080, False
    bx, by = w // 8, (h + 7) // 8
    pool_n = min(16, max(1, n_frames))
    pool = [synth.frame(pool_seed + i) for i in range(pool_n)]  # distinct frames; the rank's k-th frame is pool[k % 16]
    lanes = max(1, min(lanes, max(1, n_frames)))
    codecs = [Codec(local_rank) for _ in range(lanes)]
    main = torch.cuda.current_stream(dev)
    d_pool = [torch.from_numpy(f.view(np.int32)).to(dev) for f in pool]
    state = []
    for c in codecs:
        codes = [torch.empty((h, w), dtype=torch.uint8, device=dev) for _ in range(3)]
        t = {"codec": c, "stream": torch.cuda.ExternalStream(c.stream, device=dev), "codes": codes,
             "areas": torch.empty(bx * by * AREA_DTYPE.itemsize, dtype=torch.uint8, device=dev), "map": torch.empty(bx * by, dtype=torch.int32, device=dev),
             "count": torch.zeros(1, dtype=torch.int32, device=dev), "dec": torch.empty((h, w), dtype=torch.int32, device=dev)}
        t["st"] = {"areas": t["areas"].data_ptr(), "area_count": t["count"].data_ptr(), "block_to_area": t["map"].data_ptr(),
                   "codesA": codes[0].data_ptr(), "codesB": codes[1].data_ptr(), "codesC": codes[2].data_ptr()}
        state.append(t)

    def step_device():
        """enqueue the rank's frames round-robin over the lanes; returns the start event and the lanes' end events"""
        start = torch.cuda.Event(enable_timing=True)
        start.record(main)
        for t in state:
            t["stream"].wait_event(start)
        for k in range(n_frames):
            t = state[k % lanes]
            src = d_pool[k % pool_n]
            t["codec"].blocked_encode3d_device(src.data_ptr(), w, h, alpha, 100, True, False, t["st"], None)
            t["codec"].decode_device(t["areas"].data_ptr(), t["map"].data_ptr(), t["codes"][0].data_ptr(), t["codes"][1].data_ptr(), t["codes"][2].data_ptr(), w, h, alpha, t["dec"].data_ptr())
        ends = []
        for t in state:
            e = torch.cuda.Event(enable_timing=True)
            e.record(t["stream"])
            ends.append(e)
        return start, ends

    for _ in range(warmup):
        step_device()
    torch.cuda.synchronize(dev)
    sampler = ClockSampler(local_rank)
    launches0 = sum(c.launch_count() for c in codecs)
    barrier()
    sampler.start()
    evs = [step_device() for _ in range(steps)]
    torch.cuda.synchronize(dev)
    barrier()
    clocks = sampler.stop()
    launches = sum(c.launch_count() for c in codecs) - launches0
    dev_ms = sum(max(s.elapsed_time(e) for e in ends) for s, ends in evs)
    for c in codecs:
        c.status()  # a truncated scan would be a wrong batch

    # ---- end to end: pinned host buffers, one host thread per lane ----------------------------------------------------
    lib = codecs[0].lib
    cap = int(lib.limgcu_container_bound(w, h, 0))
    h_pool = [torch.from_numpy(f.view(np.int32)).pin_memory() for f in pool]
    h_cont = [torch.empty(cap, dtype=torch.uint8).pin_memory() for _ in range(lanes)]
    h_out = [torch.empty((h, w), dtype=torch.int32).pin_memory() for _ in range(lanes)]
    sizes = [0] * lanes

    def lane_e2e(j):
        n = C.c_size_t(0)
        c = codecs[j]
        for k in range(j, n_frames, lanes):
            rc = lib.limgcu_host_encode_container(c.h, h_pool[k % pool_n].data_ptr(), w, h, 0, 100, 1, h_cont[j].data_ptr(), cap, C.byref(n))
            assert rc == 0, (rc, lib.limgcu_last_error(c.h))
            rc = lib.limgcu_host_decode_container(c.h, h_cont[j].data_ptr(), n.value, h_out[j].data_ptr(), w * h)
            assert rc == 0, (rc, lib.limgcu_last_error(c.h))
            sizes[j] += n.value

    threads = ThreadPoolExecutor(max_workers=lanes)

    def step_e2e():
        for f in [threads.submit(lane_e2e, j) for j in range(lanes)]:
            f.result()

    step_e2e()
    for j in range(lanes):
        sizes[j] = 0
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    threads.shutdown()
    cont_bytes = sum(sizes) / max(steps, 1)  # container bytes of the rank's frames per step
    # the last frame of lane 0, decoded through the host path, equals the device path's reconstruction of the same frame
    ok = True
    if n_frames:
        k_last = ((n_frames - 1) // lanes) * lanes
        chk = Codec(local_rank)
        want = chk.encode_stream(pool[k_last % pool_n], alpha, 100, True, decoded=True)["decoded"]
        ok = bool(np.array_equal(h_out[0].numpy().view(np.uint32), want))
        chk.close()
    for c in codecs:
        c.close()
    return {"dev_ms": dev_ms, "e2e_ms": e2e_s * 1e3, "launches": launches, "container_bytes": cont_bytes, "ok": ok, "pixels": n_frames * w * h, "clocks": clocks,
            "lanes": lanes, "pool": pool_n}


def run_batch(args, rank: int, local_rank: int, world: int):
    """--mode batch (BASELINE.json config 5): a batch of --frames 1920x1080 RGB frames, frame i on rank i % world (independent units, no data-path
    collective), --lanes contexts per GPU. One step = the whole batch, encode + decode; value = device-resident, e2e = pinned host buffers
    through the container entry points (batch_core); both as the max over ranks."""
    import torch
    import torch.distributed as dist
    from limg_b200 import shard
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        init_distributed(dev)
    w, h = 1920, 1080
    mine = shard.frames_for_rank(args.frames, rank, world)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    r = batch_core(local_rank, dev, len(mine), rank * 16, args.lanes, args.steps, args.warmup, barrier)
    t = torch.tensor([r["dev_ms"], r["e2e_ms"]], dtype=torch.float64, device=dev)
    acc = torch.tensor([float(r["pixels"]), float(r["launches"]), float(r["container_bytes"]), float(r["ok"])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        mn = acc[3:].clone()
        dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        dist.all_reduce(acc[:3], op=dist.ReduceOp.SUM)
        acc[3] = mn[0]
    if rank == 0:
        dev_ms, e2e_ms = [float(x) for x in t.tolist()]
        px_job, launches_all, cont_all = float(acc[0].item()), int(acc[1].item()), float(acc[2].item())
        line = {"metric": METRIC, "value": px_job * args.steps / 1e6 / (dev_ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
                "config": {"workload": "c5_batch_1080p", "width": w, "height": h, "channels": 3, "error_factor": 100, "fast_bit_crushing": True, "frames": args.frames},
                "notes": {"dither": "lcg", "lanes_per_gpu": r["lanes"], "frames_on_rank_0": len(mine), "distinct_frames_per_rank": r["pool"],
                          "parallelism": "frame i on rank i %% %d, %d contexts per GPU, no data-path collective" % (world, r["lanes"]),
                          "step": "the whole batch: limgcu_blocked_encode3d (stream out) + limgcu_decode per frame",
                          "l2": "every step streams %d MB per rank through a 126 MB L2" % (len(mine) * w * h * 11 // 1000000)},
                "clocks": r["clocks"], "gpu_launches": launches_all,
                "e2e": {"value": px_job * args.steps / 1e6 / (e2e_ms / 1e3), "unit": UNIT, "ms_per_step": e2e_ms / args.steps,
                        "h2d_bytes_per_step": int(px_job * 4 + cont_all), "d2h_bytes_per_step": int(cont_all + px_job * 4),
                        "path": "limgcu_host_encode_container + limgcu_host_decode_container per frame, pinned host buffers, one host thread per lane",
                        "container_bits_per_pixel": 8.0 * cont_all / max(px_job, 1.0), "round_trip_equals_device_path": bool(acc[3].item() == 1.0)}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_rowband_exact(args, rank: int, local_rank: int, world: int):
    """--mode rowband_exact: ONE image, row bands over the ranks, the SAME stream as a single-GPU encode of the whole image.
    Every rank starts a step with only its band of the source resident; the step is: in-place all-gather of the source bands, pass 1 per band,
    all-gather of the table, the redundant scan, the per-area encode of the rank's areas, all-reduce of the results, finalize + decode of the
    rank's rows; every collective is issued on the codec's stream. Timed with host clocks around device-synchronised steps (the work alternates between the codec's
    stream and NCCL's), max over ranks."""
    import torch
    import torch.distributed as dist
    from limg_b200 import Codec, shard
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        init_distributed(dev)
    w, h, alpha, _ = WORKLOADS[args.workload]
    codec = Codec(local_rank)
    frame = make_frame(args.workload, 0)
    y0, y1 = shard.row_bands(h, world)[rank]
    band = torch.from_numpy(np.ascontiguousarray(frame[y0:y1]).view(np.int32)).to(dev)
    chunk_rows = shard.padded_block_rows(h, world) // world * 8  # pixel rows of every rank's (padded) chunk: equal chunks, in-place all-gather
    d_src_padded = torch.zeros((world * chunk_rows, w), dtype=torch.int32, device=dev)
    d_src = d_src_padded[:h]
    d_dec = torch.zeros((h, w), dtype=torch.int32, device=dev)
    bx = (w + 7) // 8
    cstream = torch.cuda.ExternalStream(codec.stream, device=dev)
    torch.cuda.synchronize(dev)
    rb = shard.RowBandExact(codec, d_src, w, h, alpha, rank, world)  # buffers allocated and zeroed once, reused by every step

    def step():
        with torch.cuda.stream(cstream):  # everything of a step is ordered on the codec's stream; the only host synchronisation is finalize()'s
            mine = d_src_padded[rank * chunk_rows:(rank + 1) * chunk_rows]
            mine[: y1 - y0].copy_(band)
            if world > 1:
                dist.all_gather_into_tensor(d_src_padded.view(-1), mine.view(-1))
        r = shard.encode_rowbands_exact(codec, d_src, w, h, alpha, rank, world, band=rb)
        if y1 > y0:
            off = y0 * w
            codec.decode_device(r.areas.data_ptr(), r.block_to_area.data_ptr() + (y0 // 8) * bx * 4, r.codes[0].data_ptr() + off, r.codes[1].data_ptr() + off, r.codes[2].data_ptr() + off,
                                w, y1 - y0, alpha, d_dec.data_ptr() + off * 4)
        codec.sync()
        return r

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(args.warmup):
        r = step()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        r = step()
    barrier()
    dt = time.perf_counter() - t0
    clocks = sampler.stop()

    # parity self-check outside the timed region: the whole image through the monolithic path on this rank
    ref = Codec(local_rank)
    st = ref.encode_stream(frame, alpha, 100, True, decoded=True)
    same_table = r.area_table().tobytes() == st["areas"].tobytes()
    same_codes = all(np.array_equal(r.codes[k][y0:y1].cpu().numpy(), st[n][y0:y1]) for k, n in enumerate(("codesA", "codesB", "codesC")))
    same_dec = np.array_equal(d_dec[y0:y1].cpu().numpy().view(np.uint32), st["decoded"][y0:y1])
    ok = torch.tensor([float(same_table and same_codes and same_dec), dt], dtype=torch.float64, device=dev)
    if world > 1:
        t = ok.clone()
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        ok[0] = t[0]
        dist.all_reduce(ok[1:], op=dist.ReduceOp.MAX)
    if rank == 0:
        dt = float(ok[1].item())
        print(json.dumps({"metric": METRIC, "value": w * h * args.steps / 1e6 / dt, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                          "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32+i32", "data": "synthetic",
                          "config": {"workload": args.workload, "width": w, "height": h, "channels": 4 if alpha else 3, "error_factor": 100, "fast_bit_crushing": True, "dither": "lcg",
                                     "parallelism": "whole-image-exact row bands: %d ranks, NCCL all-gather of source and pass-1 table, all-reduce of the per-area results, on the codec's stream" % world,
                                     "step": "source all-gather + limgcu_pass1 / limgcu_merge / limgcu_encode_areas / limgcu_finalize_rows + limgcu_decode of the rank's rows",
                                     "timing": "host clock around device-synchronised steps, max over ranks"},
                          "identical_to_single_gpu_encode": bool(ok[0].item() == 1.0), "areas": int(r.count.item()), "clocks": clocks}), flush=True)
    codec.close(); ref.close()
    if world > 1:
        dist.destroy_process_group()


def run_ours(args, rank: int, local_rank: int, world: int):
    import torch
    import torch.distributed as dist
    from limg_b200 import AREA_DTYPE, Codec

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product path has no CPU fallback")

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        init_distributed(dev)

    w, h, alpha, _ = WORKLOADS[args.workload]
    npx = w * h
    bx, by = (w + 7) // 8, (h + 7) // 8
    codec = Codec(local_rank)
    stream = torch.cuda.ExternalStream(codec.stream, device=dev)

    if args.mode == "rowband":
        # one image cut into bands of whole block rows, one per rank: areas do not cross bands, every band restarts the
        # dither chain -- by construction the reference run per band (tests/test_shard_gloo.py)
        from limg_b200 import shard
        full_h = h
        y0, y1 = shard.row_bands(full_h, world)[rank]
        if y1 <= y0:
            raise SystemExit("bench.py: more ranks than block-row bands")
        frame = np.ascontiguousarray(make_frame(args.workload, 0)[y0:y1])
        h = y1 - y0
        npx = w * h
        by = (h + 7) // 8
    else:
        frame = make_frame(args.workload, rank)  # every rank its own frame (independent unit)
    d_src = torch.from_numpy(frame.view(np.int32)).to(dev)
    d_codes = [torch.empty((h, w), dtype=torch.uint8, device=dev) for _ in range(3)]
    d_areas = torch.empty(bx * by * AREA_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    d_map = torch.empty(bx * by, dtype=torch.int32, device=dev)
    d_count = torch.zeros(1, dtype=torch.int32, device=dev)
    d_dec = torch.empty((h, w), dtype=torch.int32, device=dev)
    d_flush = torch.empty(512 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    st = {"areas": d_areas.data_ptr(), "area_count": d_count.data_ptr(), "block_to_area": d_map.data_ptr(),
          "codesA": d_codes[0].data_ptr(), "codesB": d_codes[1].data_ptr(), "codesC": d_codes[2].data_ptr()}

    def step_device():
        codec.blocked_encode3d_device(d_src.data_ptr(), w, h, alpha, 100, True, False, st, None)
        codec.decode_device(d_areas.data_ptr(), d_map.data_ptr(), d_codes[0].data_ptr(), d_codes[1].data_ptr(), d_codes[2].data_ptr(), w, h, alpha, d_dec.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---- device-resident timing -------------------------------------------------------------------------------
    for _ in range(args.warmup):
        step_device()
    codec.sync()

    sampler = ClockSampler(local_rank)
    launches0 = codec.launch_count()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    sampler.start()
    with torch.cuda.stream(stream):
        for i in range(args.steps):
            d_flush.fill_(i & 0xFF)  # L2 flush between timed iterations (not timed)
            ev[i][0].record(stream)
            codec.blocked_encode3d_device(d_src.data_ptr(), w, h, alpha, 100, True, False, st, None)
            ev[i][1].record(stream)
            codec.decode_device(d_areas.data_ptr(), d_map.data_ptr(), d_codes[0].data_ptr(), d_codes[1].data_ptr(), d_codes[2].data_ptr(), w, h, alpha, d_dec.data_ptr())
            ev[i][2].record(stream)
    codec.sync()
    barrier()
    clocks = sampler.stop()
    launches = codec.launch_count() - launches0

    enc_ms = [e[0].elapsed_time(e[1]) for e in ev]
    dec_ms = [e[1].elapsed_time(e[2]) for e in ev]
    total_ms = sum(enc_ms) + sum(dec_ms)

    # ---- per-kernel-phase durations (CUDA events on the codec's stream, same workload, separate passes) ------------
    codec.enable_phase_timing(True)
    phase_acc = {}
    reps = max(3, min(args.steps, 10))
    for i in range(reps):
        d_flush.fill_(i & 0xFF)
        codec.blocked_encode3d_device(d_src.data_ptr(), w, h, alpha, 100, True, False, st, None)
        for k, v in codec.phase_ms().items():
            phase_acc[k] = phase_acc.get(k, 0.0) + v / reps
    codec.enable_phase_timing(False)
    counters = codec.debug_counters()

    # the non-merged encoder (limg_encode3d_test's path): every 8x8 block its own area, no area scan
    ev_nm = []
    with torch.cuda.stream(stream):
        for i in range(5):
            d_flush.fill_(i & 0xFF)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream)
            codec.blocked_encode3d_device(d_src.data_ptr(), w, h, alpha, 100, True, True, st, None)
            e1.record(stream)
            ev_nm.append((e0, e1))
    codec.sync()
    unmerged_ms = statistics.median(a.elapsed_time(b) for a, b in ev_nm[1:])
    codec.blocked_encode3d_device(d_src.data_ptr(), w, h, alpha, 100, True, False, st, None)  # leave the merged stream in the buffers
    codec.decode_device(d_areas.data_ptr(), d_map.data_ptr(), d_codes[0].data_ptr(), d_codes[1].data_ptr(), d_codes[2].data_ptr(), w, h, alpha, d_dec.data_ptr())
    codec.sync()

    # ---- end to end through the host-buffer C ABI, pinned host memory ----------------------------------------------
    h_src = torch.from_numpy(frame.view(np.int32)).pin_memory()
    h_codes = [torch.empty((h, w), dtype=torch.uint8).pin_memory() for _ in range(3)]
    h_areas = torch.empty(bx * by * AREA_DTYPE.itemsize, dtype=torch.uint8).pin_memory()
    h_dec = torch.empty((h, w), dtype=torch.int32).pin_memory()
    import ctypes as C
    n_areas = C.c_uint32(0)
    lib = codec.lib

    def step_e2e():
        rc = lib.limgcu_host_encode_stream(codec.h, h_src.data_ptr(), w, h, int(alpha), 100, 1, h_areas.data_ptr(), C.byref(n_areas),
                                           h_codes[0].data_ptr(), h_codes[1].data_ptr(), h_codes[2].data_ptr(), None)
        assert rc == 0, rc
        rc = lib.limgcu_host_decode(codec.h, h_areas.data_ptr(), n_areas.value, h_codes[0].data_ptr(), h_codes[1].data_ptr(), h_codes[2].data_ptr(), w, h, int(alpha), h_dec.data_ptr())
        assert rc == 0, rc

    for _ in range(max(1, args.warmup)):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    area_bytes = n_areas.value * AREA_DTYPE.itemsize
    h2d = npx * 4 + area_bytes + 3 * npx
    d2h = area_bytes + 3 * npx + 4 + npx * 4

    # row-band mode: by construction the result is the reference run per band; checked against the compiled reference where it is present
    band_parity = -1.0
    if args.mode == "rowband":
        try:
            from oracle import ref
            from tools.dump_rsqrt_lut import committed_table, host_table
            if ref.available() and np.array_equal(host_table(), committed_table()):
                ref.set_modes(True, False)
                want = ref.blocked_encode3d(frame, alpha, 100, True)["pDecoded"]
                band_parity = float(np.array_equal(d_dec.cpu().numpy().view(np.uint32), want))
        except Exception:
            band_parity = -1.0

    # ---- max over ranks -----------------------------------------------------------------------------------------------
    t = torch.tensor([total_ms, sum(enc_ms), sum(dec_ms), e2e_s * 1e3], dtype=torch.float64, device=dev)
    px_all = torch.tensor([float(npx)], dtype=torch.float64, device=dev)
    parity_all = torch.tensor([band_parity], dtype=torch.float64, device=dev)
    fails_all = torch.tensor([float(counters[24])], dtype=torch.float64, device=dev)  # tries of the scan that failed their verification (last encode)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(px_all, op=dist.ReduceOp.SUM)
        dist.all_reduce(parity_all, op=dist.ReduceOp.MIN)
        dist.all_reduce(fails_all, op=dist.ReduceOp.MAX)
    total_ms, enc_total, dec_total, e2e_ms = [float(x) for x in t.tolist()]
    px_job = float(px_all.item())  # pixels all ranks process per step

    # self-check of the timed data against the round trip (cheap, outside the timed region)
    psnr, _, _ = codec.compare_device(d_src.data_ptr(), d_dec.data_ptr(), w, h, alpha)

    if rank == 0:
        peak, peak_src = measured_peaks()
        mpx_total = px_job * args.steps / 1e6
        value = mpx_total / (total_ms / 1e3)
        enc_step_ms = enc_total / args.steps
        dec_step_ms = dec_total / args.steps
        enc_gbs = 7.0 * npx / (enc_step_ms * 1e-3) / 1e9
        dec_gbs = 7.0 * npx / (dec_step_ms * 1e-3) / 1e9
        dominant = max(phase_acc, key=phase_acc.get)
        dom_ms = phase_acc[dominant]
        dom_gbs = 7.0 * npx / (dom_ms * 1e-3) / 1e9
        traffic = ncu_traffic(args.workload, dominant)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.mode == "rowband" else "weak", "vs_baseline": None,
            "dtype": "f32+i32", "data": "synthetic",
            "config": common_config(args.workload),
            "notes": {"dither": "lcg", "frames_per_rank_per_step": 1,
                      "parallelism": ("row bands of one image, one per GPU (%d rows on rank 0)" % h) if args.mode == "rowband" else ("independent frames, one per GPU" if world > 1 else "single GPU"),
                      "step": "limgcu_blocked_encode3d (stream out: area table + three code planes) + limgcu_decode", "l2": "flushed between timed steps (512 MiB fill, untimed)",
                      "warmup": "at least 3 warm-up steps are always run"},
            "encode_mpixel_s": px_job / 1e6 / (enc_step_ms * 1e-3), "decode_mpixel_s": px_job / 1e6 / (dec_step_ms * 1e-3),
            "encode_ms": enc_step_ms, "decode_ms": dec_step_ms, "psnr_db": psnr,
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": px_job * args.steps / 1e6 / (e2e_ms / 1e3), "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "path": "limgcu_host_encode_stream + limgcu_host_decode, pinned host buffers"},
            # the DOMINANT kernel phase of the encode (CUDA events around its kernels on the codec's stream): 7 algorithmic B/px of the frame over its duration
            "roofline": {"bound": "hbm", "achieved": dom_gbs, "peak": peak, "unit": "GB/s", "frac": dom_gbs / peak,
                         "traffic": traffic["bytes"] if traffic else None, "traffic_source": traffic["source"] if traffic else None,
                         "kernel": PHASE_KERNELS.get(dominant, dominant), "phase": dominant, "ms": dom_ms,
                         "share_of_encode": dom_ms / max(sum(phase_acc.values()), 1e-9), "algorithmic_bytes": 7 * npx, "peak_source": peak_src,
                         "phase_ms": {k: round(v, 4) for k, v in phase_acc.items()},
                         "note": "merge_scan = the row-pipelined greedy area scan + its verification (latency bound: a dependency chain, not HBM bound); predicate_windows = "
                                 "the main-stream part of the predicate precompute (the speculative match bitmaps run on a second stream concurrently with the scan)"},
            "roofline_encode_path": {"bound": "hbm", "achieved": enc_gbs, "peak": peak, "unit": "GB/s", "frac": enc_gbs / peak,
                                     "kernel": "all kernels of limgcu_blocked_encode3d, 7 algorithmic B/px"},
            "unmerged_encode": {"encode_ms": unmerged_ms, "encode_mpixel_s": npx / 1e6 / (unmerged_ms * 1e-3), "hbm_frac": 7.0 * npx / (unmerged_ms * 1e-3) / 1e9 / peak,
                                "what": "limgcu_blocked_encode3d with LIMGCU_FLAG_NO_MERGE (the path of limg_encode3d_test / _perf: every 8x8 block its own area), per GPU"},
            "merge": {"failed_first_tries": int(counters[24]), "failed_first_tries_max_over_ranks": int(fails_all.item()), "areas": int(counters[1]), "merged_rectangles": int(counters[0])},
            "roofline_decode": {"bound": "hbm", "achieved": dec_gbs, "peak": peak, "unit": "GB/s", "frac": dec_gbs / peak, "traffic": None,
                                "kernel": "k_decode_tile, 7 algorithmic B/px"},
        }
        if args.mode == "rowband":
            v = float(parity_all.item())
            line["bands_equal_the_reference_run_per_band"] = None if v < 0 else bool(v == 1.0)  # None: libref.so absent or another RSQRTPS table on this host
        if world == 1 and args.workload == "c2_4k_photo":
            try:
                line["north_star_8k_rgb"] = north_star_8k_rgb(codec, stream, dev, d_flush, peak)
            except Exception as e:  # an extra, never required for the headline line
                line["north_star_8k_rgb"] = {"error": repr(e)}
        if world == 1 and args.workload == "c2_4k_photo":
            try:
                line["batch_1080p"] = batch_1080p(local_rank, dev)
            except Exception as e:
                line["batch_1080p"] = {"error": repr(e)}
        try:
            line["e2e_dropin"] = e2e_dropin(codec.lib, frame, alpha, local_rank, max(3, min(args.steps, 10)))
        except Exception as e:
            line["e2e_dropin"] = {"error": repr(e)}
        if world == 1 and not args.no_cpu_baseline:
            try:
                line["cpu_baseline"] = cpu_baseline(args.workload, host_cores())
            except Exception as e:  # the baseline is reported, never required for the GPU numbers
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": 0, "kind": "unavailable", "sample": repr(e)}
        print(json.dumps(line), flush=True)

    codec.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c2_4k_photo", choices=sorted(WORKLOADS))
    ap.add_argument("--mode", default="frames", choices=["frames", "rowband", "rowband_exact", "batch"],
                    help="frames: one frame per rank (weak scaling); rowband: one image, one independent row band per rank; rowband_exact: row bands with the whole-image result; "
                         "batch: --frames 1080p frames sharded over the ranks, --lanes contexts per GPU (BASELINE.json config 5)")
    ap.add_argument("--frames", type=int, default=1024, help="--mode batch: frames of the whole batch (all ranks)")
    ap.add_argument("--lanes", type=int, default=8, help="--mode batch: contexts (frames in flight) per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
    elif args.mode == "rowband_exact":
        run_rowband_exact(args, rank, local_rank, world)
    elif args.mode == "batch":
        run_batch(args, rank, local_rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
