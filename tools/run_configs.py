"""Times the BASELINE.json configs beyond the bench default (GPU).  Prints one JSON object per config."""
import json, math, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import godotoceanwaves_b200 as gow
from bench import synth_params

def run(N, C, frames, label, regen=False):
    g = gow.WaveGenerator(); g.map_size = N; g.init_gpu(max(2, C))
    p = [synth_params(gow.WaveCascadeParameters, c) for c in range(C)]
    for _ in range(3):
        g.update_all(0.02, p)
    g.synchronize()
    g.timer_start()
    for f in range(frames):
        if regen:                       # cfg5: wind/fetch sweep, spectrum regenerated every step
            U = 5.0 + 25.0 * ((f * 7) % 26) / 25.0
            F = 10.0 ** (3.0 * ((f * 5) % 31) / 30.0)
            for q in p:
                q.wind_speed = U; q.fetch_length = F
        g.update_all(0.02, p)
    ms = g.timer_stop()
    texels = C * N * N * frames
    bytes_per_texel = 56.0 if regen else 40.0
    out = {"config": label, "map_size": N, "cascades": C, "frames": frames, "ms_total": ms, "us_per_frame": 1e3 * ms / frames,
           "cascades_per_s": C * frames / (ms * 1e-3), "gtexels_per_s": texels / (ms * 1e-3) / 1e9,
           "algorithmic_GBps": bytes_per_texel * texels / (ms * 1e-3) / 1e9}
    g.free()
    print(json.dumps(out), flush=True)

run(256, 4, 2000, "cfg2 latency: one 256x256x4 set per launch (launch/latency-bound, L2-resident)")
run(512, 4, 1000, "cfg3: 512x512x4, 1000-frame foam accumulate/decay loop")
run(1024, 8, 200, "cfg4 (1 GPU): 1024x1024x8")
run(256, 4, 300, "cfg5: 256x256x4 wind/fetch sweep, spectrum regenerated every step", regen=True)
run(128, 1, 2000, "cfg1 shape on GPU: 128x128x1")


def run_query(N, C, n_points, reps, label):
    """Map-query op (water.gdshader sampling contract) on device-resident points: kernel time only."""
    import ctypes
    import numpy as np
    import torch
    from godotoceanwaves_b200.native import load_library, check
    g = gow.WaveGenerator(); g.map_size = N; g.init_gpu(max(2, C))
    p = [synth_params(gow.WaveCascadeParameters, c) for c in range(C)]
    for _ in range(2):
        g.update_all(0.02, p)
    scales = gow.WaveGenerator.map_scales(p)
    dev = torch.device("cuda", 0)
    pts = (torch.rand(n_points, 2, device=dev) * 1000.0 - 500.0).contiguous()
    disp = torch.empty(n_points, 3, device=dev)
    grad = torch.empty(n_points, 3, device=dev)
    torch.cuda.synchronize()
    lib = load_library()
    call = lambda: check(lib.ocean_sample_maps_device(g.context, n_points, pts.data_ptr(), C, scales.ctypes.data, disp.data_ptr(), grad.data_ptr()))
    for _ in range(3):
        call()
    g.synchronize()
    g.timer_start()
    for _ in range(reps):
        call()
    ms = g.timer_stop()
    per = ms / reps
    gathered = n_points * C * 24 * 8           # 4 + 16 + 4 texel reads of 8 B per cascade and point
    out = {"config": label, "map_size": N, "cascades": C, "points": n_points, "us_per_call": 1e3 * per,
           "mpoints_per_s": n_points / (per * 1e-3) / 1e6, "gathered_GBps": gathered / (per * 1e-3) / 1e9,
           "io_GBps": n_points * 32 / (per * 1e-3) / 1e9}
    g.free()
    print(json.dumps(out), flush=True)


if "--query" in sys.argv or os.environ.get("OCEAN_RUN_QUERY", "1") != "0":
    run_query(256, 4, 1 << 20, 50, "query op: 2^20 random points x 4 cascades of 256x256 (maps L2-resident)")
    run_query(1024, 8, 1 << 20, 20, "query op: 2^20 random points x 8 cascades of 1024x1024 (maps 128 MiB)")
