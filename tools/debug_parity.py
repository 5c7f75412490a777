"""Debug helper (GPU): per-channel comparison of the CUDA path with the oracle."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import godotoceanwaves_b200 as gow
from oracle import pyoracle as po
from conftest import demo_params

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
C = int(sys.argv[2]) if len(sys.argv) > 2 else 1
pg = [demo_params(gow.WaveCascadeParameters, c) for c in range(C)]
pc = [demo_params(po.CascadeParams, c) for c in range(C)]
g = gow.WaveGenerator(); g.map_size = N; g.init_gpu(max(2, C)); g.enable_f32_taps(True)
o = po.OracleWaveGenerator(N)
g.update_all(0.02, pg); o.update_all(0.02, pc)
for c in range(C):
    d32, n32 = g.f32_maps_to_host(c)
    od, on = o.displacement_f32[c], o.normal_f32[c]
    for name, a, b in (("disp", d32, od), ("normal", n32, on)):
        for ch in range(4):
            diff = a[..., ch] != b[..., ch]
            rel = np.abs(a[..., ch].astype(np.float64) - b[..., ch]).max() / max(1e-30, np.abs(b[..., ch]).max())
            ys, xs = np.nonzero(diff)
            print(f"cascade {c} {name}[{ch}]: mismatches {diff.sum()}/{diff.size} rel {rel:.3e}",
                  "rows", np.unique(ys)[:8], "cols", np.unique(xs)[:8])
    # identify what the CUDA normal channels correlate with
    h1 = o.fft_buffer[c, 1]   # final buffer [layer][y][x][2]
    ys, xs = np.meshgrid(np.arange(N), np.arange(N), indexing="ij")
    sign = (1.0 - 2.0 * ((xs ^ ys) & 1)).astype(np.float32)
    fields = {f"L{l}.{'re' if k == 0 else 'im'}": h1[l, ..., k] * sign for l in range(4) for k in range(2)}
    for ch in range(3):
        best = sorted(((float(np.abs(np.corrcoef(n32[..., ch].ravel(), v.ravel())[0, 1])), k) for k, v in fields.items()), reverse=True)[:3]
        print("normal ch", ch, "correlates with", best)
g.free()
