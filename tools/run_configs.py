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
