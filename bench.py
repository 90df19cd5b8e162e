#!/usr/bin/env python
"""bench.py -- the reference's headline workload on B200 (BASELINE.json): the scripted flythrough at 1920x1024
through the full image-warping pipeline (reprojection, 2x2 hole gather, hole raycast, 8x4 tile refresh, cache
copy, small-gap filter, colorize), reported as frames/s, plus the full-raycast primary-ray rate.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One JSON line on stdout (rank 0).  A "step" is one frame.  Keys beyond the base contract:
  full_raycast_mrays_per_s   primary rays/s of a full-screen raycast (config 1 of BASELINE.json)
  roofline                   dominant kernel of the warped frame (traversal kernels: issue roof, not HBM) ; roofline_frame: the
                             whole frame's SURVEY 8(d) bytes against the measured HBM peak
  cpu_baseline               the reference's own kernel.cl (oracle/_ref) or its C restatement on the host cores, every kernel
                             work-group-parallel; cpu_baseline_fast_math: its -ffast-math twin (src/ocl.h:47)
  parity                     GPU output compared bit for bit with the serial reference (oracle/_ref) on the measured configurations
  e2e                        the same frames through the C ABI with the finished frame read back to host memory
Scene: data/Imrodh.rle4 if present, else the stand-in S1 (procedural, written as .rle4 and loaded through the
.rle4 loader) -- the reference's only scene is not in the mount.
N > 1: view-parallel batch (one independent camera path per GPU, no inter-GPU traffic), weak scaling.
"""
import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RES_X, RES_Y = 1920, 1024
HOLE = 0xFFFFFF00


def flythrough_pose(f, cam_id=0):
    """20 world units/s at 60 fps forward-diagonal walk with a slow pan and pitch wobble;
    cam_id offsets the start so that view-parallel ranks render different paths.  rot.x = +0.6 where the survey wrote -0.6:
    with the reference's matrix conventions a negative pitch looks at the sky (98.5 % of the rays leave the world at once);
    +0.6 looks down at the scene as the reference's screenshot does (oracle/frame.py:flythrough_pose is the same path)."""
    pos = (1.0 + f * 0.2357 + 13.0 * cam_id, 50.0, 1.0 + f * 0.2357 + 7.0 * cam_id)
    rot = (0.6 + 0.1 * math.sin(2.0 * math.pi * f / 128.0), 0.8 + 0.005 * f + 0.4 * cam_id, 0.0)
    return pos, rot


def scene_path():
    p = os.path.join(ROOT, "data", "Imrodh.rle4")
    if os.path.exists(p):
        return p, "Imrodh.rle4"
    return os.path.join("/tmp", "svo_b200_standin_S1.rle4"), "stand-in S1 (procedural floor plate + 6 blobs, depth 11, via .rle4)"


def make_scene(path):
    """The stand-in scene as a .rle4 file, written by the plain host program svo_make_scene (scene_io.cpp, no CUDA, not the
    library) so that both arms read the same file and the reference arm never maps libsvo_b200.so."""
    if os.path.exists(path):
        return
    exe = os.path.join(ROOT, "sparse-voxel-octree-raycasting_b200", "svo_make_scene")
    subprocess.run([exe, path], check=True, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)


def bench_config(scene_name, steps, warmup):
    """config of BOTH arms (b200 and --impl reference), key for key: BASELINE.json config 2."""
    return {"workload": f"scripted flythrough (SURVEY 8(d) config 2, 256 frames defined), {RES_X}x{RES_Y}, full warping pipeline (reproject 2 buffers, "
                        f"2x2 hole gather, hole raycast, 8x4 tile refresh, cache copy, gap filter, colorize); frames {warmup}..{warmup + steps - 1} timed, "
                        f"frames 0..{warmup - 1} warm-up (0 and 1 are full raycasts, src/raycast.h:150-154)",
            "scene": scene_name, "resolution": f"{RES_X}x{RES_Y}", "timed_frames": [warmup, warmup + steps - 1],
            "camera": "pos = (1,50,1) + f*(0.2357,0,0.2357), rot = (+0.6 + 0.1 sin(2 pi f/128), 0.8 + 0.005 f, 0)",
            "l2": "no flush between frames: a frame reads what the previous one wrote, as in the real pipeline; the frame's buffers + octree (~140 MB) exceed the 126 MB L2"}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        try:                                   # NVML in-process: a sample every ~20 ms (NVML queries take driver locks: sampling faster stalls the launches being timed)
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            bits = (("hw_slowdown", nv.nvmlClocksThrottleReasonHwSlowdown), ("hw_thermal_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown),
                    ("sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwThermalSlowdown), ("sw_power_cap", nv.nvmlClocksThrottleReasonSwPowerCap))
            while not self.stop_flag:
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append([str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(mx), str(nv.nvmlDeviceGetPowerUsage(h) / 1000.0)]
                                    + ["Active" if r & b else "Not Active" for _, b in bits])
                time.sleep(0.02)
            return
        except Exception:
            pass
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([s.strip() for s in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.samples[0][1]) if self.samples else None,
                "samples": len(self.samples), "reasons": sorted(reasons)}


def pin_to_gpu_numa(index, world=1):
    """Run this process on the CPUs next to its GPU: the NUMA node the GPU hangs off (sysfs local_cpulist of its PCI function)
    when that names a proper subset of the machine; otherwise (one-node boxes report every CPU for every GPU, which pins
    nothing and leaves N ranks' launch threads to migrate over each other) the rank's own 1/world slice of the allowed CPUs.
    Returns a description of the CPU set used, or None (left alone)."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
        cpus, how = None, None
        try:
            import pynvml as nv
            nv.nvmlInit()
            bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(index)).busId
            bus = bus.decode() if isinstance(bus, bytes) else bus
            bus = bus.lower()
            if len(bus.split(":")[0]) == 8:                   # NVML prints an 8-digit domain, sysfs a 4-digit one
                bus = bus[4:]
            txt = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
            loc = set()
            for part in txt.split(","):
                a, _, b = part.partition("-")
                loc.update(range(int(a), int(b or a) + 1))
            loc &= set(allowed)
            if loc and len(loc) < len(allowed):
                cpus, how = loc, f"numa {txt}"
        except Exception:
            pass
        if cpus is None and world > 1 and len(allowed) >= 2 * world:
            per = len(allowed) // world
            cpus = set(allowed[index * per:(index + 1) * per])
            how = f"slice {min(cpus)}-{max(cpus)} of {len(allowed)} cpus (no NUMA locality reported)"
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return how
    except Exception:
        return None


def ncu_record(kernel):
    """The committed ncu figures of `kernel` (profiles/ncu_traffic.json, written by tools/ncu_summary.py from the --set full
    capture of this same command): dram bytes, warp instructions, lanes per instruction, issue activity per launch."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return json.load(open(p))[kernel]
    except Exception:
        return {}


def ncu_traffic(kernel):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`; None if there is no capture."""
    v = ncu_record(kernel).get("dram_bytes_per_launch")
    return float(v) if v is not None else None


def issue_roofline(kernel, rays, launch_ms, sms, sm_mhz):
    """Roof of a traversal kernel: it is bound by instruction issue (full screen) or by the dependent chain of its slowest
    ray (sparse launches), not by HBM.  achieved = Mrays/s; peak = the rate at which the same instruction stream would
    run with every issue slot of every SM sub-partition busy and all 32 lanes active: issue slots/s x 32 / thread
    instructions per ray, from the committed ncu capture (warp instructions x lanes per instruction / rays)."""
    rec = ncu_record(kernel)
    achieved = rays / (launch_ms * 1e-3) / 1e6
    out = {"kernel": kernel, "bound": "issue", "achieved": achieved, "unit": "Mrays/s", "avg_launch_ms": launch_ms, "rays_per_launch": rays,
           "traffic": ncu_traffic(kernel), "issue_active_pct_ncu": rec.get("issue_active_pct"), "thr_per_inst_ncu": rec.get("thr_per_inst")}
    if rec.get("warp_inst_per_launch") and rec.get("thr_per_inst") and (rec.get("rays_per_launch") or rays):
        thread_inst_per_ray = rec["warp_inst_per_launch"] * rec["thr_per_inst"] / (rec.get("rays_per_launch") or rays)
        peak = sms * 4 * sm_mhz * 1e6 * 32.0 / thread_inst_per_ray / 1e6
        out.update(peak=peak, frac=achieved / peak, thread_inst_per_ray_ncu=thread_inst_per_ray,
                   peak_is=f"{sms} SMs x 4 issue slots x {sm_mhz:.0f} MHz x 32 lanes / thread instructions per ray")
    else:
        out.update(peak=None, frac=None)
    return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own kernel.cl through oracle/_ref (or the C restatement) on the host cores
# --------------------------------------------------------------------------------------------------------------
def opencl_probe():
    """north_star asks for the reference kernel.cl on the host through a CPU OpenCL runtime (PoCL).  This image has none (SURVEY
    F3); the box is re-probed at run time so that the report says what it found: platform names, or why there are none."""
    import ctypes
    try:
        cl = ctypes.CDLL("libOpenCL.so.1")
    except OSError:
        return "no libOpenCL.so.1"
    try:
        n = ctypes.c_uint(0)
        rc_ = cl.clGetPlatformIDs(0, None, ctypes.byref(n))
        if rc_ != 0 or n.value == 0:
            return f"ICD loader present, 0 platforms (clGetPlatformIDs -> {rc_})"
        ids = (ctypes.c_void_p * n.value)()
        cl.clGetPlatformIDs(n.value, ids, None)
        names = []
        for pid in ids:
            buf = ctypes.create_string_buffer(256)
            cl.clGetPlatformInfo(ctypes.c_void_p(pid), 0x0902, 256, buf, None)      # CL_PLATFORM_NAME
            names.append(buf.value.decode(errors="replace"))
        return "platforms: " + ", ".join(names) + " (not used: the baseline runs kernel.cl through the C++ shim, oracle/_ref)"
    except Exception as e:                                                          # a broken ICD must not take the bench down
        return f"probe failed: {type(e).__name__}"


def host_threads():
    """Host threads the CPU legs use: the CPUs this process may run on (torchrun exports OMP_NUM_THREADS=1, which would make
    omp_get_max_threads() say 1; the kernels take an explicit num_threads)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return os.cpu_count() or 1


def oracle_octree(path, depth=11):
    """(lib, kind, octree, root) of the CPU side, built by the CPU side's own loader: the reference's RLE4::load -> set_voxel ->
    convert_tree_blocks when oracle/_ref exists (src/raycast.h:13-46), else the C restatement."""
    from oracle import binding
    kind = "reference" if binding.have_ref() else "port"
    lib = binding.get("ref" if kind == "reference" else "orc")
    octree, root = lib.build_octree_rle4(path)
    return lib, kind, octree, root


def cpu_arm(lib, kind, octree, root, steps, warmup, budget_s, threads=None):
    """Frames 0.. of the same flythrough at 1920x1024 on the host cores; fps over the frames after `warmup`.  Every kernel runs
    work-group-parallel on all host threads, the way an OpenCL CPU runtime runs the reference (raycast_proj with its payload
    race: this leg feeds the clock, not the parity check)."""
    from oracle import frame as ofr
    cores = threads or host_threads()
    F = ofr.OracleFrame(lib, octree, root, RES_X, RES_Y, threads=cores, timed=True)
    times = []
    t_start = time.perf_counter()
    f = 0
    while f < warmup + steps:
        pos, rot = flythrough_pose(f)
        t0 = time.perf_counter()
        F.draw(pos, rot)
        dt = time.perf_counter() - t0
        if f >= warmup:
            times.append(dt)
        f += 1
        if budget_s and time.perf_counter() - t_start > budget_s and len(times) >= 3:
            break
    fps = len(times) / sum(times)
    t_full = []
    for _ in range(2):                                        # full-screen raycast_fine_2 (BASELINE.json config 1), second run timed
        t0 = time.perf_counter()
        oracle_full_raycast(lib, octree, root, RES_X, RES_Y, flythrough_pose(0))
        t_full.append(time.perf_counter() - t0)
    rays_full = RES_X * RES_Y / t_full[-1] / 1e6
    what = {"reference": "reference kernel.cl via oracle/_ref", "port": "C restatement oracle/svo_oracle.c", "reference-fast-math": "reference kernel.cl via oracle/_ref built -O3 -ffast-math (mirrors -cl-fast-relaxed-math, src/ocl.h:47)"}[kind]
    return dict(value=fps, unit="frames/s", cores=cores, kind=kind, frames_timed=len(times),
                ms_per_step=1000.0 * sum(times) / len(times), full_raycast_mrays_per_s=rays_full, opencl_cpu_runtime=opencl_probe(),
                sample=f"frames {warmup}..{warmup + len(times) - 1} of the same 1920x1024 flythrough (frames 0..{warmup - 1} untimed warm-up), "
                       f"{what}, OpenMP over work-groups for every kernel (raycast_proj with the reference's own atomics and payload race, "
                       f"as an OpenCL CPU runtime would run it; sumids is one work item)")


def cpu_arm_fast_math(octree, root, warmup, budget_s=12.0):
    """The -ffast-math twin of the reference arm (the reference builds its kernels -cl-fast-relaxed-math, src/ocl.h:47);
    None when oracle/_ref/libsvo_ref_fast.so was not built."""
    from oracle import binding
    if not binding.have_ref_fast():
        return None
    r = cpu_arm(binding.get("ref_fast"), "reference-fast-math", octree, root, 16, warmup, budget_s)
    return {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "full_raycast_mrays_per_s")}


# --------------------------------------------------------------------------------------------------------------
# parity of the measured configurations: GPU output against the serial reference, bit for bit
# --------------------------------------------------------------------------------------------------------------
def _mism(a, b):
    a, b = np.asarray(a).ravel(), np.asarray(b).ravel()
    if a.shape != b.shape:
        return int(max(a.size, b.size))
    return int(np.count_nonzero(a.view(np.uint32) != b.view(np.uint32)))


def parity_flythrough(svo, lib, octree_cpu, root_cpu, params, nframes, mode):
    """Frames 0..nframes-1 of the bench flythrough at 1920x1024: the colorized image, the hole index buffer and idbuf_size of
    every frame, and after the last one every colour / coordinate buffer, GPU (fused frame, observed after each frame) against
    the serial reference."""
    from oracle import frame as ofr
    rc = svo.raycast
    n, nb = RES_X * RES_Y, (RES_X // 16) * (RES_Y // 16)
    O = ofr.OracleFrame(lib, octree_cpu, root_cpu, RES_X, RES_Y, threads=host_threads())
    rc.reset_frames()
    # the coordinate buffers still hold the timed passes' frames; pixels no kernel writes in this pass (stale positions under
    # hole words, w of never-reprojected pixels) are compared too, so both sides start from zeroed memory
    svo.ocl.ocl_memset(rc.S.mem_backbuffer, 0, 0, 16 * n * 4)
    out = {"frames": nframes, "image_words": 0, "id_words": 0, "idbuf_size": 0, "buffer_words": 0, "position_words": 0, "hole_pixels": []}
    for f in range(nframes):
        O.draw(*flythrough_pose(f))
        rc.draw_prepared(params[f], sync=True)
        out["image_words"] += _mism(rc.read_frame(RES_X, RES_Y), O.tex)
        gsz = rc.idbuf_size()
        out["idbuf_size"] += int(gsz != O.idbuf_size)
        out["hole_pixels"].append(int(O.idbuf_size))
        idb = rc.S.mem_idbuffer.to_numpy(np.uint32, 2 * nb + max(gsz, O.idbuf_size))
        out["id_words"] += _mism(idb[nb:2 * nb + O.idbuf_size], O.idbuf[nb:2 * nb + O.idbuf_size]) + _mism(idb[1:nb], O.idbuf[1:nb])
    screen, back, _ = rc.read_buffers(RES_X, RES_Y)
    if mode == "fused":                                   # exact mode: every buffer as the reference leaves it
        out["buffer_words"] = _mism(screen, O.screen[:4 * n])
        out["position_words"] = _mism(back, O.back[:16 * n])          # xyz and w of all four coordinate buffers, bit for bit
    else:                                                 # ping-pong: the slot rendered into holds the reference's frame
        slot = rc.last_slot()
        out["buffer_words"] = _mism(screen[slot * n:(slot + 1) * n], O.screen[2 * n:3 * n])
        g, o = back.reshape(4, n, 4)[slot], O.back[:16 * n].reshape(4, n, 4)[2]
        valid = O.screen[2 * n:3 * n] != HOLE
        out["position_words"] = int(np.count_nonzero((g[:, :3].view(np.uint32) != o[:, :3].view(np.uint32)) & valid[:, None]))
    out["mismatching_words"] = out["image_words"] + out["id_words"] + out["idbuf_size"] + out["buffer_words"] + out["position_words"]
    return out


def oracle_full_raycast(lib, octree_cpu, root_cpu, rx, ry, pose, depth=11):
    """One full raycast (raycast_fine_2 over the whole screen) + colorize on the CPU: (screen words, colorized image)."""
    from oracle import frame as ofr
    n = rx * ry
    screen = np.full(4 * n + 64, HOLE, dtype=np.uint32)
    back = np.zeros(16 * n + 64, dtype=np.float32)
    tex = np.zeros(n, dtype=np.uint32)
    cam = ofr.camera_args(*pose, depth)
    t = host_threads()
    lib.raycast_fine_2(screen, back, octree_cpu, root_cpu, rx, ry, 0, 0, 0, cam["v0"], *cam["cols"], threads=t, gx=rx, gy=ry)
    lib.raycast_colorize(screen, tex, rx, ry, threads=t)
    return screen[:n], back[:4 * n], tex


def octree_words_per_ray(octree, root, frames=(0, 40)):
    """L of SURVEY.md 8(d): octree words the reference algorithm loads per ray, counted by the instrumented C oracle
    on the tile-refresh rays of a few frames of the workload (bounded sample)."""
    from oracle import binding, frame as ofr
    orc = binding.get("orc")
    import ctypes as C
    orc.lib.orc_stats_reset()
    n = RES_X * RES_Y
    screen = np.full(4 * n + 64, HOLE, dtype=np.uint32)
    back = np.zeros(16 * n + 64, dtype=np.float32)
    for f in frames:
        cam = ofr.camera_args(*flythrough_pose(f))
        add_x, add_y = (RES_X // 8) * (f & 7), (RES_Y // 4) * ((f >> 3) & 3)
        orc.raycast_fine_2(screen, back, octree, root, RES_X, RES_Y, f, add_x, add_y, cam["v0"], *cam["cols"], threads=host_threads())
    r, i, l = C.c_uint64(), C.c_uint64(), C.c_uint64()
    orc.lib.orc_stats(C.byref(r), C.byref(i), C.byref(l))
    return l.value / max(1, r.value), i.value / max(1, r.value)


def band_bench(svo, octree, root, rank, world, local_rank, dist, torch, args, frames=48, cpu=None):
    """BASELINE.json config 3: 3840x2160 on screen bands (stripes dealt round-robin to the ranks, octree replicated): full
    raycasts (every ray stores its colorized pixel into rank 0's frame over NVLink; end-of-frame flag barrier off the
    critical path) and the warped pipeline (reprojection by peer atomics).  Strong scaling: one image, all ranks.  The
    stripe height is a layout parameter chosen per workload: fine stripes balance the full raycast (2160 rows do not
    divide evenly into 64-row stripes over 8 ranks), coarse ones keep the warped pipeline's halo and id lists short.
    cpu = (lib, octree, root) of the CPU reference on rank 0: the rank-assembled output is compared with it bit for bit
    (one full-raycast image; image, id lists and every rank's own rows of the warped pipeline after frame 5)."""
    import zlib
    RX, RY = 3840, 2160
    n = RX * RY
    ocl, rc, B = svo.ocl, svo.raycast, svo.bands

    def params(f):
        rc.set_camera(*flythrough_pose(f))
        return rc.prepare_params(RX, RY, f)

    P = [params(f) for f in range(4 + frames)]
    PARITY_FRAMES = 6
    parity = {}

    def gather(obj):
        if dist is None:
            return [obj]
        out = [None] * world if rank == 0 else None
        dist.gather_object(obj, out, dst=0)
        return out

    def run(stripe_rows, what):
        db = B.DistributedBand(octree, root, RX, RY, local_rank, stripe_rows=stripe_rows, dist=dist)
        sr_eff = db.band.lay["SR"]

        def fence():
            db.band.sync()
            if dist is not None:
                dist.barrier()
                torch.cuda.synchronize()

        def timed(fn, items):
            fence()
            ocl.event_record(6)
            for it in items:
                fn(it)
            db.band.join()                            # the last frame's barrier belongs to the timed region
            ocl.event_record(7)
            ms = ocl.event_elapsed_ms(6, 7)
            fence()
            t = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
            if dist is not None:
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t[0])

        def my_rows_crc():
            """crc32 of the rows this rank owns in buffer 0 (colour words, positions): (rank, crc_colour, crc_xyz)"""
            rows = B.owned_rows(rank, world, RY, stripe_rows)
            scr = db.band.read(B.SCREEN, np.uint32, n).reshape(RY, RX)[rows]
            bck = db.band.read(B.BACK, np.float32, 4 * n).reshape(RY, RX * 4)[rows]
            return rank, zlib.crc32(np.ascontiguousarray(scr).tobytes()), zlib.crc32(np.ascontiguousarray(bck).tobytes())

        if what == "raycast":
            # SVO_FRAME_PINGPONG: consecutive full raycasts alternate between buffers 0 / 2 and between two streams, so one
            # frame's tail (a few long rays) overlaps the next frame's bulk
            R = []
            for f in range(4 + frames):
                rc.set_camera(*flythrough_pose(f))
                q = rc.prepare_params(RX, RY, f)
                q.flags |= ocl.FRAME_PINGPONG
                R.append(q)
            for p in R[:4]:
                db.band.raycast(p)
            fence()
            if cpu is not None or world > 1:             # parity: the writer's assembled image of frame 3 against the CPU reference
                img = db.band.read(B.TEX, np.uint32, n) if rank == 0 else None
                if rank == 0 and cpu is not None:
                    _, _, tex = oracle_full_raycast(cpu[0], cpu[1], cpu[2], RX, RY, flythrough_pose(3))
                    parity["bands_raycast_3840x2160"] = {"frames": 1, "ranks": world, "image_words": _mism(img, tex), "mismatching_words": _mism(img, tex)}
                fence()
            ms = timed(db.band.raycast, R[4:4 + frames])
        else:
            if cpu is not None or world > 1:
                # parity: frames 0..5 observed at the end -- the writer's image, every rank's id list, every rank's own rows
                for p in P[:PARITY_FRAMES]:
                    db.band.frame(p)
                fence()
                counts, offsets, ids = db.band.idbuf()
                mine = gather((rank, counts, ids, my_rows_crc()))
                img = db.band.read(B.TEX, np.uint32, n) if rank == 0 else None
                if rank == 0 and cpu is not None:
                    from oracle import frame as ofr
                    O = ofr.OracleFrame(cpu[0], cpu[1], cpu[2], RX, RY, threads=host_threads())
                    for f in range(PARITY_FRAMES):
                        O.draw(*flythrough_pose(f))
                    nb = O.nblocks
                    exp_off = O.idbuf[nb:2 * nb].astype(np.int64)
                    exp_cnt = np.diff(np.append(exp_off, O.idbuf_size))
                    id_bad, row_bad = 0, 0
                    for r, cnt, rid, crc in mine:
                        blocks = B.owned_blocks(r, world, RX, RY, stripe_rows)
                        exp_ids = np.concatenate([O.idbuf[2 * nb + exp_off[g]:2 * nb + exp_off[g] + exp_cnt[g]] for g in blocks]) if len(blocks) else np.zeros(0, np.uint32)
                        id_bad += _mism(cnt, exp_cnt[blocks].astype(np.uint32)) + _mism(rid, exp_ids)
                        rows = B.owned_rows(r, world, RY, stripe_rows)
                        e_s = zlib.crc32(np.ascontiguousarray(O.screen[:n].reshape(RY, RX)[rows]).tobytes())
                        e_b = zlib.crc32(np.ascontiguousarray(O.back[:4 * n].reshape(RY, RX * 4)[rows]).tobytes())
                        row_bad += int(crc[1] != e_s) + int(crc[2] != e_b)
                    iw = _mism(img, O.tex)
                    parity["bands_warped_3840x2160"] = {"frames": PARITY_FRAMES, "ranks": world, "image_words": iw, "id_words": id_bad, "hole_pixels_last_frame": int(O.idbuf_size),
                                                        "rank_row_sets_differing": row_bad, "mismatching_words": iw + id_bad + row_bad}
                fence()
            for p in P[:4]:                               # frames 0 and 1 are full raycasts through the hole path
                db.band.frame(p)
            ms = timed(db.band.frame, P[4:4 + frames])
        db.close()
        return ms, sr_eff

    ray_ms, ray_sr = run(args.ray_stripe_rows, "raycast")
    warp_ms, warp_sr = run(args.stripe_rows, "frame")
    return {"ranks": world, "stripe_rows": warp_sr, "ray_stripe_rows": ray_sr, "scaling": "strong", "frames": frames,
            "full_raycast_mrays_per_s": frames * RX * RY / (ray_ms * 1e-3) / 1e6, "full_raycast_ms": ray_ms / frames,
            "warped_fps": frames / (warp_ms * 1e-3), "warped_ms": warp_ms / frames,
            "exchange": "peer atomicMin / peer gather / halo rows / colorized pixels over NVLink, flag barriers in peer memory; no NCCL"}, parity


def builder_bench(svo, path, local_rank):
    """the compact octree of the bench scene built by the host builder and by the device builder
    (svo_octree_build_device: the node pool is built on the GPU and stays there); the two arrays must be identical."""
    ocl = svo.ocl
    t0 = time.perf_counter()
    vox = svo.scene.rle4_load(path)
    load_s = time.perf_counter() - t0
    x, y, z, c = vox.arrays()
    vox.free()
    t0 = time.perf_counter()
    words, root, _ = svo.scene.build_octree(x, y, z, c, depth=11)
    host_s = time.perf_counter() - t0
    ocl.ocl_init(local_rank)
    svo.scene.build_octree_device(x[:1000], y[:1000], z[:1000], c[:1000], depth=11)[0].free()      # CUDA / CUB warm-up
    t0 = time.perf_counter()
    mem, droot, st = svo.scene.build_octree_device(x, y, z, c, depth=11)
    dev_s = time.perf_counter() - t0
    same = bool(droot == root and mem.size == words.nbytes and np.array_equal(mem.to_numpy(), words))
    mem.free()
    # file -> device octree, the slab stream decoded on the GPU (no host voxel stream at all)
    t0 = time.perf_counter()
    mem, droot2, _ = svo.scene.octree_init_device(path, depth=11)
    file_dev_s = time.perf_counter() - t0
    same2 = bool(droot2 == root and mem.size == words.nbytes and np.array_equal(mem.to_numpy(), words))
    mem.free()
    ocl.ocl_exit()
    return {"voxels": int(len(x)), "host_rle4_load_s": round(load_s, 3), "host_builder_s": round(host_s, 3), "device_builder_s": round(dev_s, 3),
            "file_to_device_octree_s": round(file_dev_s, 3), "identical": same and same2, "host_threads": os.cpu_count(),
            "includes": "device_builder: H2D of the voxel stream (16 B/voxel) + sort + emit + normal nodes; file_to_device_octree: file read + "
                        "column walk on the host, slab decode + the same builder on the GPU (compare with host_rle4_load_s + host_builder_s)"}


def terrain14_bench(svo, args, frames=64, with_parity=True):
    """BASELINE.json config 4: fBm fractal terrain at OCTREE_DEPTH 14 (4096^2 columns of a 16384^3 world, 3-voxel shell),
    the flythrough at 8x translation and 4x rotation speed (high hole fraction), 1920x1024, fused frame."""
    rc, ocl = svo.raycast, svo.ocl
    t0 = time.perf_counter()
    vox = svo.scene.generate(kind=2, depth=14, size=4096, nblobs=0, seed=0x5EED)
    octree, root, stats = svo.scene.build_octree_voxels(vox, depth=14)
    vox.free()
    build_s = time.perf_counter() - t0
    rc.raycast_init(octree, root, max_w=RES_X, max_h=RES_Y, depth=14, mode=args.mode)
    n = RES_X * RES_Y

    def pose(f):
        return ((8.0 + f * 8 * 0.2357, 440.0, 8.0 + f * 8 * 0.2357),
                (0.6 + 0.1 * math.sin(2.0 * math.pi * f * 4 / 128.0), 0.8 + 4 * 0.005 * f, 0.0))

    P = []
    for f in range(4 + frames):
        rc.set_camera(*pose(f))
        P.append(rc.prepare_params(RES_X, RES_Y, f))
    for p in P[:4]:
        rc.draw_prepared(p, sync=False)
    ocl.ocl_end_all_kernels()
    holes = []
    ocl.event_record(8)
    for p in P[4:]:
        rc.draw_prepared(p, sync=False)
    ocl.event_record(9)
    ms = ocl.event_elapsed_ms(8, 9)
    ocl.ocl_end_all_kernels()
    rc.reset_frames()
    for p in P[:4 + 16]:                               # hole fraction of a few frames (needs a read-back each: untimed)
        rc.draw_prepared(p, sync=True)
        if p.frame >= 4:
            holes.append(rc.idbuf_size() / n)
    rc.set_camera(*pose(8))
    ray_ms = rc.full_raycast_ms(RES_X, RES_Y)
    parity = None
    if with_parity:
        # frames 0..5 of this path against the CPU oracle at depth 14 (the C restatement: oracle/_ref is the reference's
        # compile-time OCTREE_DEPTH 11): colorized image, id list and idbuf_size of every frame
        from oracle import binding, frame as ofr
        orc = binding.get("orc")
        orc.lib.orc_set_depth(14)
        try:
            O = ofr.OracleFrame(orc, octree, root, RES_X, RES_Y, threads=host_threads(), depth=14)
            nb = O.nblocks
            rc.reset_frames()
            bad = {"image_words": 0, "id_words": 0, "idbuf_size": 0}
            for f in range(6):
                O.draw(*pose(f))
                rc.draw_prepared(P[f], sync=True)
                bad["image_words"] += _mism(rc.read_frame(RES_X, RES_Y), O.tex)
                gsz = rc.idbuf_size()
                bad["idbuf_size"] += int(gsz != O.idbuf_size)
                idb = rc.S.mem_idbuffer.to_numpy(np.uint32, 2 * nb + max(gsz, O.idbuf_size))
                bad["id_words"] += _mism(idb[nb:2 * nb + O.idbuf_size], O.idbuf[nb:2 * nb + O.idbuf_size])
        finally:
            orc.lib.orc_set_depth(11)
        parity = dict(frames=6, oracle="C restatement (depth 14)", hole_pixels_last_frame=int(O.idbuf_size), **bad, mismatching_words=sum(bad.values()))
    rc.raycast_exit()
    return {"warped_fps": frames / (ms * 1e-3), "ms_per_frame": ms / frames, "hole_fraction_mean": float(np.mean(holes)),
            "full_raycast_mrays_per_s": n / (ray_ms * 1e-3) / 1e6, "octree_mb": round(octree.nbytes / 2 ** 20, 1),
            "voxels": stats["num_voxels"], "depth": 14, "host_build_s": round(build_s, 1),
            "config": "fBm terrain 4096x4096 columns at depth 14, camera 8x translation / 4x rotation of the config-2 path"}, parity


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=252)
    ap.add_argument("--warmup", type=int, default=4)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true", help="skip the comparison of the GPU output with the CPU reference (development aid)")
    ap.add_argument("--parity-frames", type=int, default=12, help="frames 0.. of the flythrough compared with the serial reference")
    ap.add_argument("--e2e-depth", type=int, default=3, choices=(2, 3, 4), help="frames in flight of the e2e leg (host buffers / colorize targets)")
    ap.add_argument("--e2e-pack", action="store_true", help="development aid: e2e frames packed to RGB24 by a pass on the copy stream "
                    "(svo_present_rgb24_async) instead of by the producing kernels (SVO_FRAME_TEX_RGB24)")
    ap.add_argument("--debug-switches", default="", help="development aid: schedule A/B switches for svo_debug_set, e.g. no_split_resolve=1,holes_smax=32")
    ap.add_argument("--profile-frames", type=int, default=32)
    ap.add_argument("--stripe-rows", type=int, default=64, help="screen-band stripe height of the 3840x2160 band measurements (0 = contiguous bands)")
    ap.add_argument("--ray-stripe-rows", type=int, default=16, help="stripe height of the banded full raycast at 3840x2160")
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary configs (bands at 3840x2160, 64-camera batch, depth-14 terrain)")
    ap.add_argument("--distinct-paths", action="store_true",
                    help="view-parallel ranks render different camera paths (default: every rank renders the config-2 flythrough, so the work per GPU is fixed as N grows)")
    ap.add_argument("--bands-only", action="store_true", help="only the 3840x2160 screen-band measurements (development aid)")
    ap.add_argument("--mode", default="fused", choices=["fused", "pingpong"],
                    help="fused: every buffer as the reference leaves it (cache copy kept); pingpong: SVO_FRAME_PINGPONG, no cache copy")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))

    from __graft_entry__ import load_package
    path, scene_name = scene_path()

    if args.impl == "reference":
        # The reference's own CPU path, stock: its loader and octree builder (RLE4::load -> convert_tree_blocks) and its
        # kernel.cl through oracle/_ref; none of this repo's kernels, builder or library (the scene file comes from the
        # plain host program svo_make_scene when data/Imrodh.rle4 is absent).
        if rank != 0:
            return 0
        make_scene(path)
        lib, kind, octree, root = oracle_octree(path)
        r = cpu_arm(lib, kind, octree, root, args.steps, args.warmup, budget_s=150.0)
        line = {"impl": "reference", "metric": "warped_pipeline_fps_1920x1024", "value": r["value"], "unit": "frames/s",
                "n_gpus": args.gpus, "steps": r["frames_timed"], "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
                "config": bench_config(scene_name, args.steps, args.warmup),
                "full_raycast_mrays_per_s": r["full_raycast_mrays_per_s"],
                "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "opencl_cpu_runtime")},
                "e2e": {"value": r["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        fm = cpu_arm_fast_math(octree, root, args.warmup)
        if fm:
            line["cpu_baseline_fast_math"] = fm
        print(json.dumps(line))
        return 0

    host_cpus = pin_to_gpu_numa(local_rank, world)
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local_rank)
    host_group = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # The NCCL communicator is created lazily (no device_id): creating it enables peer access between the GPUs, and from
        # then on every kernel boundary of every process costs more -- the same 252 frames take 0.144 instead of 0.128 ms per
        # frame (tools/gpu_replica_probe.sh: two independent processes on two GPUs run at full speed, the same two under an
        # initialised NCCL communicator lose 11 %).  The view-parallel workload has no collective, so the barriers that bracket
        # its timed regions run over a gloo group of the same ranks and NCCL first runs when the device times are reduced.
        dist.init_process_group("nccl")
        host_group = dist.new_group(backend="gloo") if os.environ.get("SVO_BENCH_BARRIER", "gloo") == "gloo" else None
    svo = load_package()
    for item in filter(None, args.debug_switches.split(",")):
        name, _, value = item.partition("=")
        svo.ocl.debug_set(name, int(value or 1))
    if rank == 0:
        make_scene(path)
    if world > 1:
        dist.barrier(group=host_group) if host_group is not None else dist.barrier()
    octree, root, stats = svo.scene.octree_init(path)          # .rle4 loader -> direct compact-octree builder
    rc, ocl = svo.raycast, svo.ocl
    if args.bands_only:
        rc.S.mode = args.mode                                     # prepare_params() without raycast_init()
        cpu = None
        if rank == 0 and not args.no_parity:
            lib, _, o_cpu, r_cpu = oracle_octree(path)
            cpu = (lib, o_cpu, r_cpu)
        out, bpar = band_bench(svo, octree, root, rank, world, local_rank, dist if world > 1 else None, torch, args, cpu=cpu)
        if rank == 0:
            print(json.dumps({"bands_3840x2160": out, "parity": bpar}))
        if world > 1:
            dist.destroy_process_group()
        return 0
    rc.raycast_init(octree, root, max_w=RES_X, max_h=RES_Y, device=local_rank, mode=args.mode)
    n = RES_X * RES_Y

    def params(f):
        rc.set_camera(*flythrough_pose(f, cam_id=rank if args.distinct_paths else 0))
        return rc.prepare_params(RES_X, RES_Y, f)

    total = args.warmup + args.steps
    P = [params(f) for f in range(total)]

    def sync_all():
        ocl.ocl_end_all_kernels()
        if world > 1:
            dist.barrier(group=host_group) if host_group is not None else dist.barrier()
            torch.cuda.synchronize()

    # ---- warm-up (frames 0 and 1 are full raycasts by construction, src/raycast.h:150-154) ----
    for f in range(args.warmup):
        rc.draw_prepared(P[f], sync=False)
    sync_all()
    # ---- timed: K frames back to back, device time on the launching stream ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = svo.launch_count()
    ocl.event_record(0)
    t_enq = time.perf_counter()
    for f in range(args.warmup, total):
        rc.draw_prepared(P[f], sync=False)
    ocl.event_record(1)
    host_enqueue_ms = (time.perf_counter() - t_enq) * 1000.0 / args.steps   # CPU time to issue one frame (launch-bound check)
    ms = ocl.event_elapsed_ms(0, 1)
    sync_all()
    launches = svo.launch_count() - launches0
    hole_frac = rc.idbuf_size() / n
    scr_all, _, _ = rc.read_buffers(RES_X, RES_Y)
    cache_slot = rc.last_slot() if args.mode == "pingpong" else 2
    valid_src = int(np.count_nonzero(scr_all[cache_slot * n:(cache_slot + 1) * n] != HOLE))
    nsrc_px = n if args.mode == "pingpong" else 2 * n
    resid_px = n - valid_src                                       # pixels neither reprojected nor traced (gap-filter input)
    del scr_all

    # ---- e2e: same frames again (fresh cache state), each finished frame read back into pinned host memory ----
    DEPTH = args.e2e_depth                                           # frames in flight (default 3): render f, copy f-1, consume f-2
    host_frames = [ocl.host_alloc(n * 3) for _ in range(DEPTH)]
    rc.prepare_present(DEPTH)                                        # no allocation inside the timed loop
    E2E_PASSES = 3        # the loop waits on the host every frame and is sensitive to whatever else the box does: median of three passes
    e2e_pass_s = []
    for e2e_pass in range(E2E_PASSES):
        rc.reset_frames()
        for f in range(args.warmup):
            rc.draw_prepared(P[f], sync=True)
        if e2e_pass == 0:
            # PCIe warm-up: 0.3 s of untimed read-backs of the last warm-up frame through every slot (link training, staging
            # buffers, first-touch of the pinned pages all happen here, not in the timed loop)
            t_w, i = time.perf_counter(), 0
            while i < 64 or (time.perf_counter() - t_w < 0.3 and i < 4096):
                ocl.present_rgb24_async(host_frames[i % DEPTH], rc.S.present_tex[i % DEPTH], n, i % DEPTH)
                if i >= DEPTH - 1:
                    ocl.present_wait((i - (DEPTH - 1)) % DEPTH)
                i += 1
            for j in range(i - (DEPTH - 1), i):
                ocl.present_wait(j % DEPTH)
        sync_all()
        t0 = time.perf_counter()
        for f in range(args.warmup, total):
            # host -> device: the frame's camera block (kernel arguments); device -> host: the finished colorized frame as R,G,B
            # bytes (what the headless writer puts into a PPM), read back on the copy stream while the next frames render.  The
            # caller owns frame f-2 after its present_wait.
            k = rc.draw_present(P[f], host_frames, rgb24="pack" if args.e2e_pack else True)
            if f - (DEPTH - 1) >= args.warmup:
                ocl.present_wait((f - (DEPTH - 1)) % DEPTH)
        for f in range(max(args.warmup, total - (DEPTH - 1)), total):
            ocl.present_wait(f % DEPTH)
        e2e_pass_s.append(time.perf_counter() - t0)
    host_frame = host_frames[k]
    sampler.stop_flag = True
    sampler.join()
    checksum = int(np.frombuffer(host_frame, dtype=np.uint8, count=n * 3).sum(dtype=np.uint64))
    # the box's plain device -> pinned-host copy rate, to read the e2e figure against (it moves n*3 bytes per frame over this link)
    try:
        dsrc = torch.empty(64 << 20, dtype=torch.uint8, device=f"cuda:{local_rank}")
        hdst = torch.empty(64 << 20, dtype=torch.uint8, pin_memory=True)
        hdst.copy_(dsrc, non_blocking=True)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(4):
            hdst.copy_(dsrc, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        d2h_gbs = 4 * (64 << 20) / (e0.elapsed_time(e1) * 1e-3) / 1e9
        del dsrc, hdst
    except Exception:
        d2h_gbs = None

    # ---- full-screen primary rays (BASELINE.json config 1): raycast_fine_2 over the whole screen ----
    ray_ms = []
    for f in (0, 40, 80, 120, 160, 200):
        rc.set_camera(*flythrough_pose(f, cam_id=rank if args.distinct_paths else 0))
        ray_ms.append(rc.full_raycast_ms(RES_X, RES_Y))
    mrays = n / (np.median(ray_ms) * 1e-3) / 1e6

    # ---- per-kernel breakdown on the stream (events around every launch), separate pass ----
    rc.reset_frames()
    for f in range(args.warmup):
        rc.draw_prepared(P[f], sync=True)
    ocl.profile_enable(True)
    ocl.profile_reset()
    pf = min(args.profile_frames, args.steps)
    for f in range(args.warmup, args.warmup + pf):
        rc.draw_prepared(P[f], sync=False)
    ocl.ocl_end_all_kernels()
    prof = ocl.profile_all()
    ocl.profile_enable(False)

    # ---- BASELINE.json config 5: 64 cameras view-parallel (full raycast each), cameras dealt round-robin to the ranks ----
    cams64 = [flythrough_pose(4 * i) for i in range(64)][rank::world]
    rc.full_raycast_batch(RES_X, RES_Y, cams64[:2])                 # (svo_raycast_batch: buffers 0 / 2 and two streams alternate)
    sync_all()
    ocl.event_record(4)
    rc.full_raycast_batch(RES_X, RES_Y, cams64)
    ocl.event_record(5)
    cams_ms = ocl.event_elapsed_ms(4, 5)
    sync_all()

    # ---- max over ranks ----
    t = torch.tensor([ms, 0.0, cams_ms] + [x * 1000.0 for x in e2e_pass_s], dtype=torch.float64, device=f"cuda:{local_rank}")
    mr = torch.tensor([mrays], dtype=torch.float64, device=f"cuda:{local_rank}")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(mr, op=dist.ReduceOp.SUM)
    if world > 1:                                                    # every rank's own device time: is the max one slow rank or all of them?
        mine = torch.tensor([ms], dtype=torch.float64, device=f"cuda:{local_rank}")
        allms = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allms, mine)
        ms_ranks = [float(v[0]) for v in allms]
    else:
        ms_ranks = [ms]
    e2e_passes_ms = sorted(float(v) for v in t[3:])                  # per pass: max over ranks
    ms_max, e2e_ms_max, cams_ms_max = float(t[0]), e2e_passes_ms[len(e2e_passes_ms) // 2], float(t[2])
    fps = world * args.steps / (ms_max * 1e-3)
    e2e_fps = world * args.steps / (e2e_ms_max * 1e-3)

    # ---- parity of the measured configuration: the same frames, observed, against the serial reference (rank 0) ----
    parity, cpu = {}, None
    if rank == 0 and not args.no_parity:
        lib, cpu_kind, o_cpu, r_cpu = oracle_octree(path)
        cpu = (lib, o_cpu, r_cpu)
        parity["oracle"] = ("reference kernel.cl + octree.h + Rle4.cpp via oracle/_ref, serial work-item order" if cpu_kind == "reference"
                            else "C restatement oracle/svo_oracle.c (oracle/_ref not built on this box)")
        parity["octree"] = {"words": int(len(o_cpu)), "mismatching_words": _mism(octree, o_cpu) + int(root != r_cpu),
                            "what": "compact octree of the product's loader + builder vs the reference's RLE4::load -> convert_tree_blocks"}
        parity["flythrough_1920x1024"] = parity_flythrough(svo, lib, o_cpu, r_cpu, P, min(args.parity_frames, total), args.mode)

    if rank == 0:
        peak, peak_src = measured_peak()
        clocks = sampler.summary()
        sm_mhz = clocks.get("sm_mhz") or 1965.0
        sms = torch.cuda.get_device_properties(local_rank).multi_processor_count
        # algorithmic bytes per launch; traversal: (4L+16) per ray (+4 index word for hole rays)
        L, iters = octree_words_per_ray(octree, root)
        per_frame_ms = {k: v[0] / pf for k, v in prof.items()}
        dom = max(per_frame_ms, key=per_frame_ms.get)
        tile_rays = (RES_X // 8) * (RES_Y // 4)
        # V = non-hole cache pixels the reprojection reads, W = pixels it fills, R = residual holes (counted on the last frame)
        V, W, H = valid_src, n - int(hole_frac * n) - resid_px, hole_frac * n
        alg = {"k_memcpy": 40.0 * n / 2, "k_memset": 4.0 * n, "k_colorize": 8.0 * n, "k_fillhole2": 4.0 * n,
               "k_proj_scatter": 4.0 * n + 16.0 * n * 0.5, "k_proj_resolve": 8.0 * n + 36.0 * n,
               "k_counthole": 4.0 * n, "k_writeids": 4.0 * n, "k_sumids": 8.0 * (n // 256),
               "k_raycast_fine_2": (4.0 * L + 16.0) * tile_rays, "k_raycast_holes": (4.0 * L + 20.0) * H,
               # fused frame
               # exact mode: the pass also carries the previous frame's cache copy (20N in, 20N out) and reads buffer 0 only
               "k_proj_scatter2": (40.0 * n if args.mode == "fused" else 4.0 * nsrc_px + 16.0 * V) + 8.0 * V,   # + one 8-byte key RMW each
               # keys in, gather, dest out, hole words, re-arm, ids, colorized word per pixel
               "k_resolve_gather": 8.0 * n + 20.0 * W + 20.0 * W + 4.0 * (n - W) + 8.0 * W + 4.0 * H + 4.0 * n,
               "k_hole_ids": 8.0 * n + 4.0 * H,                                 # keys in, ids out
               "k_rays_tile": (4.0 * L + 16.0) * tile_rays, "k_rays_holes": (4.0 * L + 20.0) * H,
               "k_copy_colorize": 40.0 * n,                                       # only when the host observes the buffers
               "k_fill_list": 28.0 * resid_px, "k_apply_patches": 12.0 * resid_px}
        rays_of = {"k_rays_holes": H, "k_raycast_holes": H, "k_rays_tile": tile_rays, "k_raycast_fine_2": tile_rays}
        d_ms, d_cnt = prof[dom]
        avg_ms = d_ms / max(1, d_cnt)
        if dom in rays_of:
            # a traversal kernel dominates the frame: HBM is not its roof (0.2 MB of DRAM traffic per launch); the algorithmic
            # bytes and the HBM fraction are kept next to the issue roof for completeness
            roof = issue_roofline(dom, rays_of[dom], avg_ms, sms, sm_mhz)
            hbm_ach = alg.get(dom, 0.0) / (avg_ms * 1e-3) / 1e9
            roof.update(algorithmic_bytes_per_launch=alg.get(dom), hbm_achieved_gbs=hbm_ach, hbm_frac=hbm_ach / peak, hbm_peak_gbs=peak,
                        octree_words_per_ray=L, iterations_per_ray=iters,
                        note="latency bound: the launch lasts as long as its slowest warp (profiles/r2_ray_latency.md); achieved Mrays/s of such a sparse launch is far below any throughput roof by construction")
        else:
            achieved = alg.get(dom, 0.0) / (avg_ms * 1e-3) / 1e9
            roof = {"kernel": dom, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": ncu_traffic(dom), "algorithmic_bytes_per_launch": alg.get(dom), "avg_launch_ms": avg_ms}
        roof["peak_source"] = peak_src
        config = bench_config(scene_name, args.steps, args.warmup)
        line = {"metric": "warped_pipeline_fps_1920x1024", "value": fps, "unit": "frames/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_max / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32+u32", "data": "synthetic",
                "config": config,
                "details": {"octree_mb": round(octree.nbytes / 2 ** 20, 1), "voxels": stats["num_voxels"], "mode": args.mode,
                            "parallelism": "1 GPU" if world == 1 else f"view-parallel x{world} ({'a different camera path' if args.distinct_paths else 'the config-2 flythrough'} on every GPU, octree replicated, no communication)",
                            "hole_fraction_last_frame": hole_frac, "ms_per_step_by_rank": [round(v / args.steps, 5) for v in ms_ranks]},
                "full_raycast_mrays_per_s": float(mr[0]), "full_raycast_ms": float(np.median(ray_ms)),
                "e2e": {"value": e2e_fps, "unit": "frames/s", "h2d_bytes_per_step": 120, "d2h_bytes_per_step": n * 3, "frame_format": "rgb24",
                        "ms_per_step": e2e_ms_max / args.steps, "frame_checksum": checksum, "frames_in_flight": DEPTH, "host_cpus": host_cpus,
                        "passes_fps": [round(world * args.steps / (v * 1e-3), 1) for v in e2e_passes_ms], "value_is": "median of the passes",
                        "d2h_link_gbs": d2h_gbs, "d2h_used_gbs": e2e_fps / world * n * 3 / 1e9,
                        # what the box's device -> pinned-host path allows when every rank reads back at once (the GPUs share
                        # PCIe switches: ~55 GB/s per GPU alone, 15-20 GB/s each when 8 copy together)
                        "link_ceiling_fps": (d2h_gbs * 1e9 / (n * 3) * world) if d2h_gbs else None},
                "gpu_launches": launches, "host_enqueue_ms_per_frame": host_enqueue_ms, "clocks": clocks,
                "kernel_ms_per_frame": {k: round(v, 5) for k, v in sorted(per_frame_ms.items(), key=lambda kv: -kv[1])},
                "roofline": roof}
        # the whole frame against the HBM roof: SURVEY 8(d)'s bytes (the warp passes at 80-100 B/pixel + the rays' node words)
        frame_bytes = (4 + 4 + 16.0 * V / n + 20.0 * W / n + 4 + 4.0 * H / n + 40 + 4 + 8) * n + (4.0 * L + 16.0) * (H + tile_rays)
        fb_ach = frame_bytes / (ms_max / args.steps * 1e-3) / 1e9
        line["roofline_frame"] = {"kernel": "whole frame (all launches)", "bound": "hbm", "achieved": fb_ach, "peak": peak, "unit": "GB/s", "frac": fb_ach / peak,
                                  "algorithmic_bytes_per_frame": frame_bytes, "ms_per_frame": ms_max / args.steps,
                                  "note": "SURVEY 8(d): clear 4N + reprojection 4N+16V+20W + hole gather 4N+4H + cache copy 40N + gap filter 4N + colorize 8N + (4L+16) per ray; the frame is a chain of latencies, not bandwidth bound"}
        # the frame's dominant bandwidth kernel
        stream_k = [k for k in ("k_copy_colorize", "k_resolve_gather", "k_proj_scatter2", "k_memcpy", "k_proj_resolve") if k in prof]
        if stream_k:
            sk = max(stream_k, key=lambda k: per_frame_ms[k])
            s_ms = prof[sk][0] / max(1, prof[sk][1])
            s_ach = alg[sk] / (s_ms * 1e-3) / 1e9
            line["roofline_streaming"] = {"kernel": sk, "bound": "hbm", "achieved": s_ach, "peak": peak, "unit": "GB/s", "frac": s_ach / peak,
                                          "traffic": ncu_traffic(sk), "algorithmic_bytes_per_launch": alg[sk], "avg_launch_ms": s_ms}
        # and the other half of the metric, the full-screen traversal kernel: issue bound
        fr = issue_roofline("k_raycast_fine_2", n, float(np.median(ray_ms)), sms, sm_mhz)
        fr_bytes = (4.0 * L + 16.0) * n
        fr.update(algorithmic_bytes_per_launch=fr_bytes, hbm_achieved_gbs=fr_bytes / (float(np.median(ray_ms)) * 1e-3) / 1e9, hbm_peak_gbs=peak)
        line["roofline_full_raycast"] = fr
        if not args.no_cpu_baseline and world == 1:
            if cpu is None:
                lib, cpu_kind, o_cpu, r_cpu = oracle_octree(path)
            r = cpu_arm(lib, cpu_kind, o_cpu, r_cpu, steps=24, warmup=args.warmup, budget_s=25.0)
            line["cpu_baseline"] = {k: r[k] for k in ("value", "unit", "cores", "kind", "sample", "opencl_cpu_runtime")}
            line["cpu_baseline"]["full_raycast_mrays_per_s"] = r["full_raycast_mrays_per_s"]
            fm = cpu_arm_fast_math(o_cpu, r_cpu, args.warmup)
            if fm:
                line["cpu_baseline_fast_math"] = fm
    rc.raycast_exit()
    extras = {}
    if not args.no_extras:
        extras["view_parallel_64_cameras"] = {"grays_per_s": 64 * n / (cams_ms_max * 1e-3) / 1e9, "ms": cams_ms_max,
                                              "config": "64 poses of the flythrough (every 4th frame), full raycast 1920x1024 each, dealt round-robin to the ranks, no communication; per rank one svo_raycast_batch (buffers 0 / 2 and two streams alternate, consecutive cameras overlap)"}
        extras["bands_3840x2160"], bpar = band_bench(svo, octree, root, rank, world, local_rank, dist if world > 1 else None, torch, args, cpu=cpu)
        parity.update(bpar)
        if world == 1:
            extras["terrain_depth14"], tpar = terrain14_bench(svo, args, with_parity=not args.no_parity)
            if tpar:
                parity["terrain_depth14"] = tpar
            extras["octree_build"] = builder_bench(svo, path, local_rank)
    if rank == 0:
        line.update(extras)
        if parity:
            checks = [v for v in parity.values() if isinstance(v, dict) and "mismatching_words" in v]
            parity["frames_checked"] = int(sum(v.get("frames", 0) for v in checks))
            parity["mismatching_words"] = int(sum(v["mismatching_words"] for v in checks))
            line["parity"] = parity
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
