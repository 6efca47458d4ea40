#!/usr/bin/env python3
"""bench.py — volumetric ReSTIR hot-path benchmark (contract: see the task brief / DESIGN.md §6).

  python bench.py --gpus N --steps K --warmup W            # our CUDA path through the C ABI (libvrs.so)
  python bench.py --impl reference ...                     # the reference math on the box's host cores (oracle port)

A "step" is one frame of the workload.  The default is BASELINE.json configs[3] (the north-star target): bunny_cloud
stand-in at 3840x2160, 10k point lights, RIS M=32, visibility + temporal + spatial reuse (k=5, 2 iterations), camera on a
6 deg/frame orbit.  configs[1] (smoke.vdb 1080p temporal) is `--workload smoke_1080p_temporal`.  N > 1 (torchrun) splits
the screen into horizontal bands, grid replicated, halo rows exchanged over NVLink (strong scaling: total work fixed).
"""
import argparse
import ctypes as C
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

WORKLOADS = {
    # name: (asset, W, H, lights, M, flags, spatial k, iterations)
    "smoke_1080p_temporal": dict(asset="smoke", W=1920, H=1080, lights=64, M=32, flags=1 | 2, k=5, iters=0,
                                 desc="configs[1]: smoke.vdb 1920x1080, 64 point lights, RIS M=32 + temporal reuse, 6deg/frame orbit"),
    "smoke_1080p_full": dict(asset="smoke", W=1920, H=1080, lights=64, M=32, flags=1 | 2 | 4, k=5, iters=2,
                             desc="smoke.vdb 1920x1080, 64 lights, full spatiotemporal (k=5, 2 iterations)"),
    "smoke_4k_full": dict(asset="smoke", W=3840, H=2160, lights=10000, M=32, flags=1 | 2 | 4, k=5, iters=2,
                          desc="configs[3] shape on smoke.vdb: 3840x2160, 10k lights, full spatiotemporal"),
    # the assets of configs[2..4] are missing blobs in the reference checkout: deterministic procedural stand-ins
    "explosion_1080p_full": dict(asset="proc:explosion:296", W=1920, H=1080, lights=-1001, M=32, flags=1 | 2 | 4, k=5, iters=2,
                                 desc="configs[2] on a procedural stand-in for explosion.vdb (~8M voxels): 1920x1080, <=1001 emissive-voxel lights, full spatiotemporal k=5 x2"),
    "bunny_4k_full": dict(asset="proc:bunny_cloud:288", W=3840, H=2160, lights=10000, M=32, flags=1 | 2 | 4, k=5, iters=2,
                          desc="configs[3] on a procedural stand-in for bunny_cloud.vdb (~1.2M voxels): 3840x2160, 10k lights, full spatiotemporal k=5 x2"),
    # configs[4] (meant for --gpus 8; 33 Mpixel: 8 GB of planes + 2.8 GB of queues when run on one GPU)
    "composite_8k_full": dict(asset="proc:fire_torus:384", W=7680, H=4320, lights=100000, M=32, flags=1 | 2 | 4, k=5, iters=2,
                              desc="configs[4] on a procedural stand-in for the fire.vdb + torus_knot_helix.vdb composite (one merged grid): 7680x4320, 100k lights, full spatiotemporal k=5 x2"),
}
DEFAULT_WORKLOAD = "bunny_4k_full"
ORBIT_DEG, ORBIT_RADIUS = 6.0, 1.25          # camera orbit of every workload: degrees per frame, radius in bbox half-diagonals
UNBIASED_FLAGS = 2 | 4 | 16 | 32             # temporal + spatial + FINAL_VISIBILITY + FINALIZE_W (tests/test_unbiased.py)

# SURVEY.md §8d algorithmic (compulsory) bytes per pixel per frame, by pass and by the kernel that moves them
BYTES_PER_PX = {"initial": 96, "temporal": 96, "spatial_iter": 128, "shade": 128}
KERNEL_BYTES_PER_PX = {"k_ris": 96, "k_finish": 96, "k_spatial": 128, "k_shade": 128}


def data_desc(wl):
    if wl["asset"].startswith("proc:"):
        return "synthetic: camera orbit + generated lights over a deterministic procedural stand-in grid (%s; the reference's asset is a missing blob)" % wl["asset"]
    return "synthetic camera orbit + generated lights over the reference's %s.vdb grid (assets/%s.vrsg)" % (wl["asset"], wl["asset"])


def config_of(wl, name):
    """Identical for both arms (the driver compares the dicts)."""
    return {"workload": wl["desc"], "name": name, "resolution": [wl["W"], wl["H"]], "M": wl["M"], "lights": wl["lights"], "flags": wl["flags"],
            "spatial_neighbors": wl["k"], "spatial_iterations": wl["iters"], "orbit_deg_per_frame": ORBIT_DEG,
            "l2": "per-frame working set %.0f MB > 126 MB L2 (inputs larger than L2, no flush)" % (wl["W"] * wl["H"] * 240 / 1e6)}


def asset_path(V, name):
    """assets/<name>.vrsg, or a procedural stand-in generated (host-only, deterministic) into the scratch directory.
    V = the loaded product module, or None to generate in a child process (the reference arm never maps libvrs.so)."""
    if not name.startswith("proc:"):
        return os.path.join(ROOT, "assets", name + ".vrsg")
    _, kind, res = name.split(":")
    d = os.path.join(os.environ.get("TMPDIR", "/tmp"), "vrs_assets")
    os.makedirs(d, exist_ok=True)
    path = os.path.join(d, "%s_%s.vrsg" % (kind, res))
    if not os.path.exists(path):
        tmp = path + ".%d.tmp" % os.getpid()
        if V is not None:
            V.write_procedural_vrsg(kind, int(res), tmp)
        else:
            subprocess.check_call([sys.executable, "-c", "import sys; sys.path.insert(0, %r); import vrs_pkg; vrs_pkg.load().write_procedural_vrsg(%r, %d, %r)"
                                   % (ROOT, kind, int(res), tmp)])
        os.replace(tmp, path)
    return path


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def orbit_eye(center, radius, height, angle_deg):
    a = math.radians(angle_deg)
    return (center[0] + radius * math.cos(a), center[1] + height, center[2] + radius * math.sin(a))


class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def scene_lights(gen_lights, wl, lo, hi, emissive):
    """Lights of a workload from the grid's world bbox: generatePointLights semantics inside the box, or the emissive-voxel
    rule (callable `emissive`).  Returns (lights, bbox centre, bbox half-diagonal)."""
    ctr = [(a + b) * 0.5 for a, b in zip(lo, hi)]
    ext = [(b - a) * 0.5 for a, b in zip(lo, hi)]
    if wl["lights"] < 0:
        lights = emissive(-wl["lights"])
    else:
        lights = gen_lights([c - e for c, e in zip(ctr, ext)], [c + e for c, e in zip(ctr, ext)], False, wl["lights"])
    return lights, ctr, math.sqrt(sum(e * e for e in ext))


def build_scene_inputs(V, wl, R):
    gi = R.gridInfo()
    # emissive-voxel lights (Renderer.cpp:1615-1637); "hot" = raw value above 85 % of the maximum
    return scene_lights(V.generate_point_lights, wl, list(gi.world_bbox_min), list(gi.world_bbox_max),
                        lambda n: R.collectEmissiveLights(0.85 * gi.max_density, n))


def measured_reach_rows(V, wl, path, device, frames=61, q=4):
    """How far (rows) the temporal reprojection of actual hit points moves on this orbit, measured: the orbit rendered at 1/q
    resolution with temporal reuse on, the library's temporal_reach_rows counter scaled back up, plus 15 % and 8 rows.  Tighter
    than the bounding-box bound of temporal_halo_rows (hits sit inside the box), which lets cost-balanced bands get thinner."""
    W, H = wl["W"], wl["H"]
    P = V.Renderer(max(W // q, 16), max(H // q, 16), spatial_iterations=0, device=device)
    P.loadVDB(path)
    lights, ctr, diag = build_scene_inputs(V, wl, P)
    P.createRestirLights(lights[:1])
    P.m_restirUniforms.initialLightSampleCount, P.m_restirUniforms.flags = 1, 2
    P.CameraManip.setLookat(orbit_eye(ctr, ORBIT_RADIUS * diag, 0.0, 0.0), ctr)
    P.createRestirUniformBuffer()
    for f in range(frames):
        P.CameraManip.setLookat(orbit_eye(ctr, ORBIT_RADIUS * diag, 0.0, ORBIT_DEG * f), ctr)
        P.renderFrame(clock=f)
    rows = P.counters().temporal_reach_rows
    P.destroy()
    return int(math.ceil((rows + 1) * q * 1.15)) + 8


def temporal_halo_rows(V, wl, lo, hi, ctr, diag, frames=60, minimum=32):
    """Rows a band must keep above / below itself so that the temporal reprojection of any point of the grid's bounding box
    stays inside them on this camera orbit: max |row(prevVP p) - row(curVP p)| over a lattice of the box, plus a margin.
    (The library counts violations: vrs_get_counters().temporal_out_of_halo must stay 0.)"""
    W, H = wl["W"], wl["H"]
    ax = [np.linspace(lo[a], hi[a], 9) for a in range(3)]
    P = np.stack(np.meshgrid(*ax, indexing="ij"), -1).reshape(-1, 3)
    P4 = np.concatenate([P, np.ones((len(P), 1))], 1)
    aspect = float(np.float32(W) / np.float32(H))
    proj = V.perspectiveVK(60.0, aspect, 0.1, 1000.0).reshape(4, 4).T.astype(np.float64)
    rows, worst = None, 0.0
    for f in range(frames + 1):
        view = V.look_at(orbit_eye(ctr, ORBIT_RADIUS * diag, 0.0, ORBIT_DEG * f), ctr).reshape(4, 4).T.astype(np.float64)
        q = P4 @ (proj @ view).T
        y = (q[:, 1] / q[:, 3] + 1.0) * 0.5 * H
        if rows is not None:
            worst = max(worst, float(np.abs(y - rows).max()))
        rows = y
    need = int(math.ceil(worst * 1.1)) + 4
    return max(minimum, (need + 7) // 8 * 8)


PROBE_ANGLES = tuple(float(a) for a in range(0, 360, 30))       # orbit positions the band balancer looks at


def row_costs(V, wl, path, device):
    """Per-row cost model at several orbit positions: a quarter-resolution probe frame per position (rendered by every rank on
    its own GPU, bit-identical everywhere) gives the per-row count of volume-hitting pixels; cost(row) = hits + 2.5 % of the
    pixels.  Returns an array [angle][row] at full resolution."""
    W, H = wl["W"], wl["H"]
    q = 4
    w4, h4 = max(W // q, 16), max(H // q, 16)
    P = V.Renderer(w4, h4, spatial_iterations=0, device=device)
    P.loadVDB(path)
    lights, ctr, diag = build_scene_inputs(V, wl, P)
    P.createRestirLights(lights[:1])
    P.m_restirUniforms.initialLightSampleCount, P.m_restirUniforms.flags = 1, 0
    out = []
    for ang in PROBE_ANGLES:
        P.CameraManip.setLookat(orbit_eye(ctr, ORBIT_RADIUS * diag, 0.0, ang), ctr)
        P.createRestirUniformBuffer()
        P.renderFrame(clock=int(ang))
        hits = P.readGBuffer()["worldPos"][..., 3].sum(1).astype(np.float64)
        cost = np.repeat(hits, q)[:H] * q + 0.025 * W         # per full-resolution row (ncu: ~1.6 ns per hit, ~0.04 ns per pixel)
        if len(cost) < H:
            cost = np.concatenate([cost, np.full(H - len(cost), cost[-1])])
        out.append(cost)
    P.destroy()
    return np.stack(out)


def split_rows(cost, world, min_rows=32):
    """Band edges for `world` ranks, every band at least `min_rows` tall.  `cost` is [row] or [position][row]: the bands first
    get equal shares of the summed cost, then (several positions) the interior edges move to minimise the sum over the
    positions of the slowest band's cost — frames advance in lock step with the slowest band, so what counts is the maximum
    per frame, not the average load."""
    cost = np.atleast_2d(np.asarray(cost, np.float64))
    H = cost.shape[1]
    tot = cost.sum(0)
    cum = np.concatenate([[0.0], np.cumsum(tot)])
    edges = [0]
    for r in range(1, world):
        y = int(np.searchsorted(cum, cum[-1] * r / world))
        edges.append(y)
    edges.append(H)
    for r in range(1, world):                                     # enforce the minimum band height front to back, then back to front
        edges[r] = max(edges[r], edges[r - 1] + min_rows)
    for r in range(world - 1, 0, -1):
        edges[r] = min(edges[r], edges[r + 1] - min_rows)
    if cost.shape[0] > 1 and world > 1:
        cums = np.concatenate([np.zeros((cost.shape[0], 1)), np.cumsum(cost, 1)], 1)

        def objective(e):
            return float(np.max(cums[:, e[1:]] - cums[:, e[:-1]], axis=1).sum())

        best = objective(np.array(edges))
        step = max(H // (8 * world), 1)
        while step >= 1:
            improved = False
            for r in range(1, world):
                for d in (-step, step):
                    e = list(edges)
                    e[r] += d
                    if e[r] - e[r - 1] < min_rows or e[r + 1] - e[r] < min_rows:
                        continue
                    v = objective(np.array(e))
                    if v < best * (1.0 - 1e-9):
                        best, edges, improved = v, e, True
            if not improved:
                step //= 2
    return [(edges[r], edges[r + 1]) for r in range(world)]


def calibrated_bands(V, wl, path, world, rank, device, dist, halo, min_rows, rounds=5):
    """Start from the hit-count model, then correct it with measurements: every rank renders its band (no exchange, timing
    only) at the probe positions of the orbit, the per-band times are all-gathered and turned into a per-band correction of
    the row costs.  Which rows a rank renders never changes a pixel (tests/test_gpu_multi.py); only the load balance does."""
    import torch
    cost = row_costs(V, wl, path, device)
    bands = split_rows(cost, world, min_rows)
    for _ in range(rounds):
        band = bands[rank]
        R = V.Renderer(wl["W"], wl["H"], spatial_iterations=wl["iters"], band=band, halo_rows=halo, device=device)
        R.loadVDB(path)
        lights, ctr, diag = build_scene_inputs(V, wl, R)
        R.createRestirLights(lights)
        u = R.m_restirUniforms
        u.initialLightSampleCount, u.spatialNeighbors, u.flags = wl["M"], wl["k"], wl["flags"]
        R.CameraManip.setLookat(orbit_eye(ctr, ORBIT_RADIUS * diag, 0.0, 0.0), ctr)
        R.createRestirUniformBuffer()
        R.setPassTiming(True)
        times = np.zeros(len(PROBE_ANGLES))
        for rep in range(2):                                      # first sweep warms up (graphs, caches)
            for i, ang in enumerate(PROBE_ANGLES):
                R.CameraManip.setLookat(orbit_eye(ctr, ORBIT_RADIUS * diag, 0.0, ang), ctr)
                R.renderFrame(clock=i)
                if rep == 1:
                    times[i] = R.timings().frame_ms
        R.destroy()
        t = torch.tensor(times, device="cuda", dtype=torch.float64)
        allt = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        meas = np.stack([x.cpu().numpy() for x in allt])                                    # [rank][position]
        model = np.stack([cost[:, b[0]:b[1]].sum(1) for b in bands])                        # [rank][position]
        corr = meas.sum(1) / np.maximum(model.sum(1), 1e-30)
        corr = corr / corr.mean()
        for (y0, y1), c in zip(bands, corr):
            cost[:, y0:y1] *= c ** 0.75                                                     # damped
        bands = split_rows(cost, world, min_rows)
    return bands


def load_traffic(workload):
    """ncu-measured DRAM bytes and warp instructions per kernel launch for this workload (profiles/traffic.json), or {}."""
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if not os.path.exists(tp):
        return {}
    try:
        return json.load(open(tp)).get("workloads", {}).get(workload, {})
    except Exception:
        return {}


def issue_roofline(traffic, fps, clocks, world):
    """The frame is instruction-issue bound, not HBM bound: besides the contract's HBM roofline, report how
    much of the SMs' issue capacity the frame uses.  Warp instructions per frame come from the committed ncu launch list of
    this workload (profiles/traffic.json), the frame rate and the SM clock are measured live.  None when the inputs do not
    apply (no capture for the workload, several GPUs, no clock sample)."""
    try:
        if not traffic or world != 1 or not clocks or not clocks.get("sm_mhz"):
            return None
        winst = sum(float(k_["warp_inst_M"]) * float(k_.get("launches_per_frame", 1)) for k_ in traffic.values()) * 1e6
        if winst <= 0:
            return None
        peak_issue = 148 * 4 * float(clocks["sm_mhz"]) * 1e6            # 4 warp schedulers per SM, one instruction per clock each
        return {"warp_instructions_per_frame": int(winst), "achieved_ginst_s": round(winst * fps / 1e9, 1),
                "peak_ginst_s": round(peak_issue / 1e9, 1), "frac": round(winst * fps / peak_issue, 4),
                "source": "profiles/traffic.json (ncu smsp__inst_executed.sum per kernel) x live frames/s; peak = 148 SMs x 4 schedulers x measured SM clock"}
    except Exception:
        return None


def run_ours(args):
    import torch
    import vrs_pkg
    V = vrs_pkg.load()
    wl = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit("bench.py: --gpus %d but WORLD_SIZE=%d (launch N>1 with torch.distributed.run)" % (args.gpus, world))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    W, H = wl["W"], wl["H"]
    path = asset_path(V, wl["asset"]) if rank == 0 or world == 1 else None
    if dist is not None:
        dist.barrier()       # the stand-in asset (if any) is on disk for every rank
        path = asset_path(V, wl["asset"])
    halo, reach = 32, 32
    if world > 1:
        P = V.Renderer(16, 16, spatial_iterations=0, device=local)
        P.loadVDB(path)
        gi = P.gridInfo()
        lo, hi = list(gi.world_bbox_min), list(gi.world_bbox_max)
        _, ctr0, diag0 = scene_lights(lambda *a: None, dict(wl, lights=1), lo, hi, None)
        P.destroy()
        # rows the temporal reprojection can move on this orbit.  Peer memory: such pixels are read in place from the adjacent
        # band (which therefore must be at least that tall), halo rows only serve spatial reuse (radius 30).  NCCL: they must
        # be shipped, the halo is that tall.
        reach = temporal_halo_rows(V, wl, lo, hi, ctr0, diag0) if wl["flags"] & 2 else 32
        if wl["flags"] & 2:
            reach = max(32, min(reach, (measured_reach_rows(V, wl, path, local) + 7) // 8 * 8))     # same on every rank (deterministic render)
        halo = reach if args.exchange == "nccl" else 32
    bands = calibrated_bands(V, wl, path, world, rank, local, dist, halo, reach) if world > 1 else [(0, H)]
    band = bands[rank] if world > 1 else None
    R = V.Renderer(W, H, spatial_iterations=wl["iters"], band=band, halo_rows=halo, device=local)
    R.loadVDB(path)
    lights, ctr, diag = build_scene_inputs(V, wl, R)
    wl = dict(wl, lights=len(lights))
    R.createRestirLights(lights)
    if world > 1:
        if args.exchange == "nccl":
            uid = [V.comm_unique_id() if rank == 0 else None]
            dist.broadcast_object_list(uid, src=0)
            R.commInit(uid[0], rank, world)
        else:                                  # peer memory over NVLink: all-gather the CUDA-IPC handle blobs
            blobs = [None] * world
            dist.all_gather_object(blobs, R.peerExport())
            R.peerConnect(rank, world, blobs)
    u = R.m_restirUniforms
    u.initialLightSampleCount, u.spatialNeighbors, u.flags = wl["M"], wl["k"], wl["flags"]
    radius = ORBIT_RADIUS * diag
    R.CameraManip.setLookat(orbit_eye(ctr, radius, 0.0, 0.0), ctr)
    R.createRestirUniformBuffer()

    stream = torch.cuda.ExternalStream(R.stream())
    L = V.lib()
    # Host inputs of every frame (camera uniforms, push constants) are produced by the Renderer methods that mirror the
    # reference's updateUniformBuffer / updateRestirUniformBuffer / updateFrame, a block of frames ahead of their use (outside
    # the timed regions); a step then is the one C-ABI call.
    inputs = []
    frame_no = [0]
    flags_now = [wl["flags"]]

    def extend_inputs(n):
        while len(inputs) < frame_no[0] + n:
            f = len(inputs)
            R.CameraManip.setLookat(orbit_eye(ctr, radius, 0.0, ORBIT_DEG * f), ctr)
            R.m_restirUniforms.flags = flags_now[0]
            R.updateUniformBuffer(); R.updateRestirUniformBuffer(); R.updateFrame()
            inputs.append((V.GlobalUniforms.from_buffer_copy(R.m_globalUniforms), V.RestirUniforms.from_buffer_copy(R.m_restirUniforms),
                           V.PushConstantRestir.from_buffer_copy(R.m_pcRestirPost)))
            if R.m_pcRestirPost.frame > 10:
                R.m_pcRestirPost.initialize = 0

    def step():
        gu, ru, pc = inputs[frame_no[0]]
        s_ = L.vrs_render_frame(R._ctx, C.byref(gu), C.byref(ru), C.byref(pc), frame_no[0])
        if s_:
            raise SystemExit("vrs_render_frame failed: %s" % L.vrs_last_error(R._ctx).decode())
        frame_no[0] += 1

    def barrier():
        R.synchronize()
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_block(n):
        """n frames bracketed by CUDA events on the launching stream; returns ms per frame (this rank)."""
        extend_inputs(n)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(n):
            step()
        e1.record(stream)
        barrier()
        return e0.elapsed_time(e1) / n

    def max_over_ranks(v):
        if dist is None:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0])

    # ---- device-resident throughput: K frames per repetition, CUDA events on the launching stream, max over ranks per
    # repetition; repetitions until the timed region adds up to >= 0.5 s; the median repetition is reported.
    R.setPassTiming(False)
    extend_inputs(args.warmup)
    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    reps = []
    while True:
        reps.append(max_over_ranks(timed_block(args.steps)))
        if sum(reps) * args.steps >= 500.0 or len(reps) >= 25:
            break
    t_wall = time.perf_counter() - t_wall0
    ms_step = float(np.median(reps))
    clocks = sampler.stop() if rank == 0 else None
    cnt = R.counters()
    hits_band, ooh = cnt.hits, cnt.temporal_out_of_halo

    # ---- per-pass event times of a few more frames (events recorded inside vrs_render_frame)
    pass_ms = {"initial": 0.0, "spatial": 0.0, "shade": 0.0, "exchange": 0.0}
    launches = 0
    R.setPassTiming(True)
    probe = int(round(360.0 / ORBIT_DEG))          # one whole orbit: the average does not depend on where the timed blocks ended
    extend_inputs(probe + 2)
    barrier()
    for i in range(probe + 2):
        step()
        t = R.timings()
        if i < 2:
            continue                      # ranks re-align after the barrier
        pass_ms["initial"] += t.initial_ms; pass_ms["spatial"] += t.spatial_ms; pass_ms["shade"] += t.shade_ms; pass_ms["exchange"] += t.exchange_ms
        launches = t.launches
    for k_ in pass_ms:
        pass_ms[k_] /= probe
    R.setPassTiming(False)
    barrier()

    # ---- per-kernel event times (frames launched kernel by kernel, one event after each kernel on the context stream)
    kernel_ms = {}
    R.setKernelTiming(True)
    extend_inputs(probe + 2)
    barrier()
    for i in range(probe + 2):                    # one whole orbit again
        step()
        if i >= 2:
            for name, ms in R.kernelTimes():
                kernel_ms[name] = kernel_ms.get(name, 0.0) + ms / float(probe)
    R.setKernelTiming(False)
    barrier()
    kernel_ms_ranks = None
    if dist is not None:            # per rank: k_halo_wait is the time a band waits for its neighbours (skew), k_halo_push the transfer
        allk = [None] * world
        dist.all_gather_object(allk, kernel_ms)
        kernel_ms_ranks = {n_: [round(k_.get(n_, 0.0), 5) for k_ in allk] for n_ in kernel_ms}
        kernel_ms = {n_: max(v_) for n_, v_ in kernel_ms_ranks.items()}

    # ---- end to end through the public API: host uniforms in, host frame buffer out, every step.
    # The result a caller of the reference gets per frame is the presented 8-bit image (restir_post.frag:104 ->
    # swapchain); here it is presented headlessly into pinned host memory, the copy of frame i overlapping frame i+1.
    pinned = [torch.empty((R.rows, W, 4), dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    host_frames = [t.numpy() for t in pinned]
    extend_inputs(2)
    for i in range(2):
        step(); R.presentAsync(host_frames[i & 1])
    R.presentWait()
    e2e_reps = []
    for _ in range(len(reps)):                     # the same number of K-step blocks as the device-resident measurement (same orbit sectors)
        extend_inputs(args.steps)
        barrier()
        t0 = time.perf_counter()
        for i in range(args.steps):
            step()
            R.presentAsync(host_frames[i & 1])
        R.presentWait()
        barrier()
        e2e_reps.append(max_over_ranks(1000.0 * (time.perf_counter() - t0) / args.steps))
    e2e_ms = float(np.median(e2e_reps))
    # the same with the full RGBA32F accumulation buffer read back synchronously (debug / parity read path)
    pinned_f = torch.empty((R.rows, W, 4), dtype=torch.float32, pin_memory=True)
    hf = pinned_f.numpy()
    extend_inputs(5)
    t0 = time.perf_counter()
    for _ in range(5):
        step(); R.readFrame(hf)
    barrier()
    e2e_f32_ms = 1000.0 * (time.perf_counter() - t0) / 5

    # ---- the unbiased configuration (FINALIZE_W + FINAL_VISIBILITY instead of the reference's selection-time w and
    # reused visibility): same frame, second ratio-tracking ray in the shade pass
    unbiased_ms = None
    if not args.no_unbiased:
        ub_flags = UNBIASED_FLAGS if wl["flags"] & 4 else (UNBIASED_FLAGS & ~4)
        flags_now[0] = ub_flags
        del inputs[frame_no[0]:]
        extend_inputs(16)
        for _ in range(16):                     # every buffer-rotation phase of the new flag set has its graphs before the clock starts
            step()
        unbiased_ms = float(np.median([max_over_ranks(timed_block(args.steps)) for _ in range(3)]))
        flags_now[0] = wl["flags"]
        del inputs[frame_no[0]:]

    if dist is not None:
        tt = torch.tensor([pass_ms["initial"], pass_ms["spatial"], pass_ms["shade"], pass_ms["exchange"], float(hits_band), float(ooh)], device="cuda", dtype=torch.float64)
        per_rank = [torch.zeros_like(tt) for _ in range(world)]
        dist.all_gather(per_rank, tt)
        per_rank_initial = [round(float(t_[0]), 4) for t_ in per_rank]
        hits_total = sum(float(t_[4]) for t_ in per_rank)
        ooh = int(sum(float(t_[5]) for t_ in per_rank))
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        pass_ms = {"initial": float(tt[0]), "spatial": float(tt[1]), "shade": float(tt[2]), "exchange": float(tt[3])}
    else:
        hits_total = float(hits_band)
    if rank == 0:
        fps = 1000.0 / ms_step
        px = W * H
        peak, peak_src = measured_peak()
        temporal = bool(wl["flags"] & 2)
        frame_bytes = px * (BYTES_PER_PX["initial"] + (BYTES_PER_PX["temporal"] if temporal else 0) + BYTES_PER_PX["shade"] +
                            (BYTES_PER_PX["spatial_iter"] * wl["iters"] if wl["flags"] & 4 else 0))
        init_bytes = px * (R.rows / H) * (BYTES_PER_PX["initial"] + (BYTES_PER_PX["temporal"] if temporal else 0))
        init_ach = init_bytes / (pass_ms["initial"] * 1e-3) / 1e9 if pass_ms["initial"] > 0 else 0.0
        traffic = load_traffic(args.workload)
        # dominant kernel by live event time; algorithmic bytes per launch = SURVEY §8d bytes/px of the part of the pass that
        # kernel moves x pixels.  At N > 1 the per-kernel probe is skipped: the initial pass as a whole is used.
        per_kernel = {}
        for name, ms in kernel_ms.items():
            e = {"ms": round(ms, 5)}
            if name in KERNEL_BYTES_PER_PX and ms > 0:
                nl = wl["iters"] if name == "k_spatial" else 1
                e["launches_per_frame"] = nl
                e["algorithmic_gbs"] = round(KERNEL_BYTES_PER_PX[name] * px * nl / (ms * 1e-3) / 1e9, 1)
            tk = traffic.get(name)
            if tk and ms > 0:
                e["dram_bytes_per_launch"] = tk.get("dram_bytes_per_launch")
                e["dram_gbs"] = round(tk["dram_bytes_per_launch"] * tk.get("launches_per_frame", 1) / (ms * 1e-3) / 1e9, 1)
            per_kernel[name] = e
        if kernel_ms and world == 1:
            dom = max((n_ for n_ in kernel_ms if n_ in KERNEL_BYTES_PER_PX), key=lambda n_: kernel_ms[n_])
            nl = wl["iters"] if dom == "k_spatial" else 1
            dom_ms = kernel_ms[dom] / nl
            dom_bytes = KERNEL_BYTES_PER_PX[dom] * px
            roof = {"kernel": dom + (" (RIS stage of the initial pass: G-buffer 64 B + reservoir 32 B written per pixel, M=%d candidates each)" % wl["M"] if dom == "k_ris" else ""),
                    "bound": "hbm", "achieved": round(dom_bytes / (dom_ms * 1e-3) / 1e9, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(dom_bytes / (dom_ms * 1e-3) / 1e9 / peak, 4),
                    "traffic": (traffic.get(dom) or {}).get("dram_bytes_per_launch"), "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": int(dom_bytes), "launch_ms": round(dom_ms, 5),
                    "note": "instruction-issue bound, not HBM bound (DESIGN.md §4): see per_kernel for the streaming kernels and profiles/ for issue-slot utilisation"}
        else:
            roof = {"kernel": "initial pass (all kernels of restir.rgen main on a volume)", "bound": "hbm", "achieved": round(init_ach, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(init_ach / peak, 4), "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_launch": int(init_bytes)}
        hit_fraction = hits_total / px
        line = {
            "metric": "ReSTIR frames/s", "value": round(fps, 3), "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(ms_step, 5), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": data_desc(wl),
            "config": config_of(wl, args.workload),
            "partition": ("bands minimising the slowest band's cost over 12 orbit positions (hit-count model corrected by 5 timed calibration rounds) x%d %s, halo %d rows, temporal reach %d rows (%s)"
                          % (world, [b[1] - b[0] for b in bands], halo, reach,
                             "shipped as halo rows" if args.exchange == "nccl" else "read in place from the adjacent band over NVLink")) if world > 1 else "single GPU",
            "repetitions_ms_per_step": [round(r_, 5) for r_ in reps], "timed_region_s": round(sum(reps) * args.steps / 1e3, 3),
            "orbit_mean_ms_per_step": round(float(np.mean(reps)), 5),
            "mpixels_per_s": round(px * fps / 1e6, 1), "mpixel_samples_per_s": round(px * wl["M"] * fps / 1e6, 1),
            "hit_fraction": round(hit_fraction, 4), "hit_pixel_samples_per_s_M": round(hits_total * wl["M"] * fps / 1e6, 1),
            "frame_hbm_gbs": round(frame_bytes * fps / 1e9, 1), "frame_hbm_frac": round(frame_bytes * fps / 1e9 / peak, 4),
            "pass_ms": {k_: round(v, 5) for k_, v in pass_ms.items()},
            "initial_pass_roofline": {"achieved": round(init_ach, 1), "frac": round(init_ach / peak, 4), "algorithmic_bytes_per_launch": int(init_bytes)},
            "roofline": roof, "per_kernel": per_kernel,
            "e2e": {"value": round(1000.0 / e2e_ms, 3), "unit": "frames/s", "h2d_bytes_per_step": 192 + 320 + 20,
                    "d2h_bytes_per_step": int(R.rows * W * 4), "ms_per_step": round(e2e_ms, 4),
                    "result": "presented RGBA8 frame (vrs_present_async, double-buffered, pinned host memory); wall clock around K-step blocks, median of %d blocks" % len(e2e_reps),
                    "repetitions_ms_per_step": [round(r_, 5) for r_ in e2e_reps],
                    "rgba32f_sync_readback_ms_per_step": round(e2e_f32_ms, 4)},
            "unbiased": None if unbiased_ms is None else {"value": round(1000.0 / unbiased_ms, 3), "unit": "frames/s", "ms_per_step": round(unbiased_ms, 5),
                                                          "flags": ub_flags,
                                                          "what": "same workload with FINALIZE_W + FINAL_VISIBILITY (the configuration tests/test_unbiased.py proves unbiased)"},
            "temporal_out_of_halo": int(ooh),
            "gpu_launches": int(launches * args.steps * len(reps)), "clocks": clocks, "wall_s": round(t_wall, 3),
            "per_rank_initial_ms": per_rank_initial if world > 1 else None,
            "per_rank_kernel_ms": kernel_ms_ranks,
            "issue_roofline": issue_roofline(traffic, fps, clocks, world),
        }
        if not args.no_cpu_baseline and world == 1:
            line["cpu_baseline"] = cpu_baseline(args.workload, seconds=args.cpu_seconds)
        print(json.dumps(line), flush=True)
    R.destroy()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def oracle_setup(workload, lights_override=None):
    """The oracle's own scene for a workload: its own numpy .vrsg reader, its own light generator, no libvrs.so in this
    process (procedural stand-in assets are generated by a child process).  `lights_override` (parity tests): use these
    lights instead of generating them."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import grid_py
    import oracle as O
    import vdb_py
    O.build()
    O.set_num_threads(os.cpu_count() or 1)          # launchers pin OMP_NUM_THREADS=1 (torch.distributed.run): use the box
    wl = WORKLOADS[workload]
    g = grid_py.read_vrsg(asset_path(None, wl["asset"]))
    raw, vmin, _ = grid_py.dense_raw(g)
    dens = vdb_py.density_from_raw(raw, g.level_set, g.background)
    probe = O.OracleScene(dens, vmin, g.voxel_size, g.translation, np.ones((1, 8), np.float32))
    lo, hi = probe.world_bbox()

    def emissive(n):      # emissive-voxel lights, Renderer.cpp:1615-1637 (first voxels in tree order above the threshold)
        thr = np.float32(0.85) * np.float32(g.leaf_value.max())
        sel = np.argwhere(g.leaf_mask & (g.leaf_value > thr))[:n]
        off = sel[:, 1]
        ijk = g.leaf_origin[sel[:, 0]] + np.stack([off >> 6, (off >> 3) & 7, off & 7], 1)
        lights = np.zeros((len(ijk), 8), np.float32)
        for a in range(3):
            lights[:, a] = np.float32(probe.c.A) * ijk[:, a].astype(np.float32) + np.float32(probe.c.B[a])
        lights[:, 3] = 1.0
        lights[:, 4:7] = [0.6, 0.2, 0.1]
        lights[:, 7] = np.float32(0.2126) * np.float32(0.6) + np.float32(0.7152) * np.float32(0.2) + np.float32(0.0722) * np.float32(0.1)
        return lights

    lights, ctr, diag = scene_lights(O.generate_point_lights, wl, lo, hi, emissive)
    if lights_override is not None:
        lights = np.ascontiguousarray(lights_override, np.float32).reshape(-1, 8)
    wl = dict(wl, lights=len(lights))
    scene = O.OracleScene(dens, vmin, g.voxel_size, g.translation, lights)
    return O, wl, scene, ctr, diag


def oracle_frames(O, wl, scene, ctr, diag, frames, first=0, renderer=None, y_rows=None):
    """Render `frames` frames of the workload's orbit on the oracle; y_rows = (y0, y1) restricts every pass to a band of
    rows (a bounded sample of the frame: the per-row work is what the full frame does for those rows)."""
    W, H = wl["W"], wl["H"]
    OR = renderer or O.OracleRenderer(scene, W, H, spatial_iterations=wl["iters"])
    radius = ORBIT_RADIUS * diag
    prev = O.Camera(orbit_eye(ctr, radius, 0.0, ORBIT_DEG * (first - 1)), ctr) if first > 0 else None
    times = []
    for f in range(first, first + frames):
        cam = O.Camera(orbit_eye(ctr, radius, 0.0, ORBIT_DEG * f), ctr)
        gu = O.global_uniforms(cam, W, H)
        ru = O.restir_uniforms(cam, prev, W, H, wl["lights"], M=wl["M"], flags=wl["flags"], k=wl["k"])
        pc = O.PushConstant(0, 0, 0, 0, 1)
        t0 = time.perf_counter()
        if y_rows is None:
            OR.render(gu, ru, pc, f)
        else:
            OR.render(gu, ru, pc, f, y0=y_rows[0], y1=y_rows[1])
        times.append(time.perf_counter() - t0)
        prev = cam
    return OR, times


def reference_arm(workload, steps, warmup, budget_s):
    """The reference math on the host cores.  Each step is one full frame of the workload when the whole run fits the time
    budget; otherwise every step renders the same centred band of rows (all passes, all M candidates, all reuse) and the
    frame rate is scaled by the band's share of the frame's hit pixels — the rows differ in cost only through how many of
    their pixels hit the volume (stated in `sample`)."""
    O, wl, scene, ctr, diag = oracle_setup(workload)
    W, H = wl["W"], wl["H"]
    OR, t_first = oracle_frames(O, wl, scene, ctr, diag, 1)
    hit_rows = OR.gbuffer()["worldPos"][..., 3].sum(1).astype(np.float64) + 0.025 * W       # same per-row cost model as the band balancer
    est_frame = t_first[0]
    full = est_frame * (steps + warmup) <= budget_s
    if full:
        rows, share = None, 1.0
    else:
        want = max(budget_s / (steps + warmup) / est_frame, 0.02)                           # share of the frame's cost per step
        cum = np.cumsum(hit_rows) / hit_rows.sum()
        mid = int(np.searchsorted(cum, 0.5))
        y0 = y1 = mid
        while (cum[min(y1, H - 1)] - (cum[y0 - 1] if y0 > 0 else 0.0)) < want and (y0 > 0 or y1 < H):
            y0, y1 = max(0, y0 - 4), min(H, y1 + 4)
        rows = (y0, max(y1, y0 + 8))
        share = float(hit_rows[rows[0]:rows[1]].sum() / hit_rows.sum())
    if warmup > 1:
        oracle_frames(O, wl, scene, ctr, diag, warmup - 1, first=1, renderer=OR, y_rows=rows)
    _, times = oracle_frames(O, wl, scene, ctr, diag, steps, first=max(1, warmup), renderer=OR, y_rows=rows)
    ms = 1000.0 * sum(times) / len(times) / share
    sample = ("each step = one full %dx%d frame of the workload" % (W, H)) if full else \
             ("each step = rows %d..%d of the %dx%d frame (all passes), %.1f %% of the frame's cost by the hit-pixel row model; frame time = band time / that share"
              % (rows[0], rows[1], W, H, 100.0 * share))
    return O, wl, ms, sample


def cpu_baseline(workload, seconds=15.0):
    O, wl, ms, sample = reference_arm(workload, steps=4, warmup=1, budget_s=seconds)
    return {"value": round(1000.0 / ms, 4), "unit": "frames/s", "cores": O.num_threads(), "kind": "port",
            "sample": sample + "; oracle/liboracle.so (the reference shader math restated in C++, OpenMP over rows, -O2 -ffp-contract=off)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    O, wl, ms, sample = reference_arm(args.workload, steps=args.steps, warmup=max(1, args.warmup), budget_s=args.reference_seconds)
    fps = 1000.0 / ms
    W, H = wl["W"], wl["H"]
    line = {
        "impl": "reference", "metric": "ReSTIR frames/s", "value": round(fps, 4), "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms, 3), "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
        "data": data_desc(wl),
        "config": config_of(wl, args.workload),
        "cpu_baseline": {"value": round(fps, 4), "unit": "frames/s", "cores": O.num_threads(), "kind": "port",
                         "sample": sample + " on the host cores (oracle/liboracle.so: the reference shader math restated in C++, OpenMP over rows; "
                                            "the reference's own Vulkan-RT binary cannot run here)"},
        "e2e": {"value": round(fps, 4), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "mpixel_samples_per_s": round(W * H * wl["M"] * fps / 1e6, 2),
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"], help="multi-GPU halo exchange: NVLink peer-memory kernel or NCCL send/recv")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-unbiased", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=15.0, help="time budget of the cpu_baseline leg")
    ap.add_argument("--reference-seconds", type=float, default=150.0, help="time budget of --impl reference (all steps)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
