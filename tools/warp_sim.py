#!/usr/bin/env python
"""Lane-occupancy model of the full-screen traversal kernel (VERDICT r1 task 4), run on the CPU.

The instrumented C oracle logs, for every primary ray of a full-screen raycast, how many descents it makes before each
step (orc_set_trip_log).  From those per-ray sequences this tool replays the GPU kernel's while-while loop warp by warp
(8x4 footprints) and counts issued warp instructions and active lanes for

  * static   - today's kernel: one ray per lane, the warp lasts as long as its longest ray;
  * refill-T - persistent warps over a strip of footprints: whenever fewer than T lanes hold a live ray (and rays are left
               in the strip) the idle lanes finalise their ray and set up the next one (partial-lane set-up/finalise).

Costs are SASS instruction counts per loop part (tools/sass_loop.py): descent iteration, step, set-up, finalise.
Analysis only: reads the oracle, never the product.
"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

C_DESC, C_STEP, C_SETUP, C_FINAL, C_REFILL_CHECK = 50, 85, 160, 60, 4


def ray_log(rx, ry, frame, stride=160):
    import bench
    from oracle import binding, frame as ofr
    path, _ = bench.scene_path()
    bench.make_scene(path)
    orc = binding.get("orc")
    octree, root = orc.build_octree_rle4(path)
    n = rx * ry
    log = np.zeros((n, stride), dtype=np.uint8)
    orc.lib.orc_set_trip_log.argtypes = [C.c_void_p, C.c_int]
    orc.lib.orc_set_trip_log(log.ctypes.data, stride)
    screen = np.zeros(4 * n + 64, dtype=np.uint32)
    back = np.zeros(16 * n + 64, dtype=np.float32)
    cam = ofr.camera_args(*bench.flythrough_pose(frame))
    orc.raycast_fine_2(screen, back, octree, root, rx, ry, 0, 0, 0, cam["v0"], *cam["cols"], threads=bench.host_threads(), gx=rx, gy=ry)
    orc.lib.orc_set_trip_log(None, 0)
    return log


def decode(log):
    """per ray: list of descents per trip; the last trip has no step when the ray ended in the descent loop (hit)."""
    last = (log & 0x80) != 0
    ntrips = last.argmax(axis=1) + 1
    desc = (log & 0x7f).astype(np.int32)
    return desc, ntrips


def sim_static(desc, ntrips, rx, ry, fw=8, fh=4):
    """today's kernel.  Returns (warp instructions, thread instructions)."""
    n = rx * ry
    d = desc.reshape(ry // fh, fh, rx // fw, fw, -1).transpose(0, 2, 1, 3, 4).reshape(-1, 32, desc.shape[1])   # [warp][lane][trip]
    nt = ntrips.reshape(ry // fh, fh, rx // fw, fw).transpose(0, 2, 1, 3).reshape(-1, 32)
    T = desc.shape[1]
    trip = np.arange(T)[None, None, :]
    alive = trip < nt[:, :, None]                       # lane runs trip t
    d = np.where(alive, d, 0)
    # descent loop of trip t: iterations = max over lanes of (descents + 1 failing test, folded into the cost), lanes active in
    # iteration j = lanes with descents > j
    dmax = d.max(axis=1)                                # [warp][trip]
    warp_alive = alive.any(axis=1)
    desc_warp = dmax.sum()
    desc_thread = d.sum()
    # step phase: every lane alive in trip t and not ending inside this trip's descent loop; a ray that ends by a hit has its
    # last entry WITHOUT a step; a ray that ends by leaving does step.  The log cannot tell them apart: count the last entry as a step
    # for misses only -> approximate with "last entry has a step when its descent count is 0".
    lastt = trip == (nt[:, :, None] - 1)
    steps = alive & ~(lastt & (d > 0))
    step_warp = steps.any(axis=1).sum()
    step_thread = steps.sum()
    nw = d.shape[0]
    wi = desc_warp * C_DESC + step_warp * C_STEP + nw * (C_SETUP + C_FINAL) + warp_alive.sum() * 0
    ti = desc_thread * C_DESC + step_thread * C_STEP + n * (C_SETUP + C_FINAL)
    return wi, ti


def sim_refill(desc, ntrips, rx, ry, thresh, strip, fw=8, fh=4, converged_setup=False):
    """persistent warp over `strip` consecutive footprints (x direction); python loop, run it on a sample of strips."""
    fy, fx = ry // fh, rx // fw
    d5 = desc.reshape(fy, fh, fx, fw, -1).transpose(0, 2, 1, 3, 4).reshape(fy, fx, 32, -1)
    n4 = ntrips.reshape(fy, fh, fx, fw).transpose(0, 2, 1, 3).reshape(fy, fx, 32)
    wi = ti = 0
    rng = np.random.default_rng(1)
    rows = rng.choice(fy, size=min(fy, 48), replace=False)
    nrays = 0
    for r in rows:
        for s0 in range(0, fx, strip):
            q_d = d5[r, s0:s0 + strip].reshape(-1, d5.shape[-1])
            q_n = n4[r, s0:s0 + strip].reshape(-1)
            nq = len(q_n)
            nrays += nq
            cur = np.arange(32)                     # ray index per lane
            t = np.zeros(32, dtype=np.int64)        # trip of the lane's ray
            live = np.ones(32, dtype=bool)
            nxt = 32
            wi += C_SETUP; ti += 32 * C_SETUP
            while True:
                nl = live.sum()
                if nl == 0 and nxt >= nq:
                    break
                if (nl < thresh and nxt < nq) or nl == 0:
                    idle = np.flatnonzero(~live)
                    take = min(len(idle), nq - nxt)
                    if converged_setup:
                        # set-up / finalise done 32 rays at a time through shared memory: full-lane cost, plus a small swap
                        wi += 24; ti += 24 * len(idle)
                        ti += take * (C_SETUP + C_FINAL); wi += (C_SETUP + C_FINAL) * take / 32.0
                    else:
                        wi += C_SETUP + C_FINAL; ti += take * (C_SETUP + C_FINAL)
                    for k in range(take):
                        cur[idle[k]] = nxt + k; t[idle[k]] = 0; live[idle[k]] = True
                    nxt += take
                    continue
                dd = np.where(live, q_d[cur, np.minimum(t, q_d.shape[1] - 1)], 0)
                lastt = live & (t == q_n[cur] - 1)
                dm = dd.max()
                wi += dm * C_DESC + C_REFILL_CHECK; ti += dd.sum() * C_DESC
                st = live & ~(lastt & (dd > 0))
                if st.any():
                    wi += C_STEP; ti += st.sum() * C_STEP
                t += 1
                live &= ~lastt
            # finalise of the last rays
            wi += C_FINAL; ti += 32 * C_FINAL
    return wi, ti, nrays


def sim_policies(desc, ntrips, rx, ry, nwarps=4000, fw=8, fh=4):
    """Which phase a warp runs when its lanes disagree.  Every ray is a string of actions (D = descent iteration, S = step);
    per warp iteration the policy picks the action(s) to issue, lanes whose next action matches advance.  Vectorised over a
    random sample of warps.  Returns {policy: (warp instructions per ray, lanes per instruction)}."""
    d = desc.reshape(ry // fh, fh, rx // fw, fw, -1).transpose(0, 2, 1, 3, 4).reshape(-1, 32, desc.shape[1])
    nt = ntrips.reshape(ry // fh, fh, rx // fw, fw).transpose(0, 2, 1, 3).reshape(-1, 32)
    sel = np.random.default_rng(0).choice(d.shape[0], min(nwarps, d.shape[0]), replace=False)
    d, nt = d[sel], nt[sel]
    W, T, L = d.shape[0], d.shape[2], 640
    trip = np.arange(T)[None, None, :]
    alive = trip < nt[:, :, None]
    d = np.where(alive, d, 0)
    has_step = alive & ~((trip == (nt[:, :, None] - 1)) & (d > 0))
    cs = np.cumsum(d, axis=2)
    pos_s = cs + np.cumsum(has_step, axis=2) - has_step                 # position of trip k's step in the action string
    length = cs[:, :, -1] + has_step.sum(axis=2)
    act = np.full((W, 32, L), 2, np.int8)                               # 0 = D, 1 = S, 2 = done
    act[np.arange(L)[None, None, :] < length[:, :, None]] = 0
    w_i, l_i, t_i = np.nonzero(has_step)
    p = pos_s[w_i, l_i, t_i]
    ok = p < L
    act[w_i[ok], l_i[ok], p[ok]] = 1
    wI, lI = np.arange(W)[:, None], np.arange(32)[None, :]
    out = {}
    for policy in ("descents first (today)", "steps first", "if-if", "majority", "descend while >= 12 lanes"):
        ptr = np.zeros((W, 32), np.int64)
        wi = np.zeros(W)
        ti = np.zeros(W)
        while True:
            a = act[wI, lI, np.minimum(ptr, L - 1)]
            wd, wsx = a == 0, a == 1
            nd, ns = wd.sum(1), wsx.sum(1)
            if not ((nd + ns) > 0).any():
                break
            if policy.startswith("descents"):
                do_d = nd > 0; do_s = (nd == 0) & (ns > 0)
            elif policy.startswith("steps"):
                do_s = ns > 0; do_d = (ns == 0) & (nd > 0)
            elif policy == "if-if":
                do_d = nd > 0; do_s = ns > 0
            elif policy == "majority":
                do_d = (nd > 0) & (nd >= ns); do_s = (ns > 0) & ~do_d
            else:
                do_d = (nd >= 12) | ((ns == 0) & (nd > 0)); do_s = ~do_d & (ns > 0)
            wi += do_d * C_DESC + do_s * C_STEP
            ti += do_d * nd * C_DESC + do_s * ns * C_STEP
            ptr += (wd & do_d[:, None]) | (wsx & do_s[:, None])
        n = W * 32
        wi_t, ti_t = wi.sum() + W * (C_SETUP + C_FINAL), ti.sum() + n * (C_SETUP + C_FINAL)
        out[policy] = (wi_t / n, ti_t / wi_t)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--res", default="1920x1024")
    ap.add_argument("--frame", type=int, default=40)
    a = ap.parse_args()
    rx, ry = map(int, a.res.split("x"))
    log = ray_log(rx, ry, a.frame)
    desc, ntrips = decode(log)
    n = rx * ry
    print(f"rays {n}, trips/ray {ntrips.mean():.1f} (max {ntrips.max()}), descents/ray {np.where(np.arange(desc.shape[1])[None,:] < ntrips[:,None], desc, 0).sum()/n:.1f}")
    wi, ti = sim_static(desc, ntrips, rx, ry)
    print(f"static   : {wi/n:7.1f} warp inst/ray, {ti/wi:5.1f} lanes/inst")
    base = wi / n
    for pol, (w, l) in sim_policies(desc, ntrips, rx, ry).items():
        print(f"policy {pol:28s}: {w:7.1f} warp inst/ray, {l:5.1f} lanes/inst")
    for conv in (False, True):
        for strip in (8, 30):
            for thresh in (8, 16, 24, 28):
                wi, ti, nr = sim_refill(desc, ntrips, rx, ry, thresh, strip, converged_setup=conv)
                print(f"refill T={thresh:2d} strip={strip:2d} {'smem-staged setup' if conv else 'partial-lane setup'}: {wi/nr:7.1f} warp inst/ray, {ti/wi:5.1f} lanes/inst, x{base/(wi/nr):.2f}")


if __name__ == "__main__":
    main()
