"""Debug helper (GPU): multi-frame foam parity, prints mismatch counts per frame."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import godotoceanwaves_b200 as gow
from godotoceanwaves_b200 import native
if os.environ.get("OCEAN_LIB"):
    native._LIB_PATH = os.environ["OCEAN_LIB"]
from oracle import pyoracle as po
from conftest import demo_params
N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
C = int(sys.argv[2]) if len(sys.argv) > 2 else 3
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 4
pg = [demo_params(gow.WaveCascadeParameters, c) for c in range(C)]
pc = [demo_params(po.CascadeParams, c) for c in range(C)]
g = gow.WaveGenerator(); g.map_size = N; g.init_gpu(max(2, C))
o = po.OracleWaveGenerator(N)
for f in range(frames):
    g.update_all(0.02, pg); o.update_all(0.02, pc)
    d, n = g.maps_to_host(0, C)
    for c in range(C):
        bad = (n[c].view(np.uint16) != o.normal_map[c])
        print(f"frame {f} cascade {c}: normal mismatches per channel {bad.reshape(-1,4).sum(0)}, disp mismatches {(d[c].view(np.uint16) != o.displacement_map[c]).sum()}")
        if bad[..., 3].any():
            ys, xs = np.nonzero(bad[..., 3]); print("   foam bad rows", np.unique(ys)[:10], "cols", np.unique(xs)[:10], "got", n[c][ys[0], xs[0], 3], "want", o.normal_half()[c][ys[0], xs[0], 3])
g.free()
