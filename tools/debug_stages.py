"""Debug helper (GPU): stage-by-stage comparison of the CUDA path with the oracle (spectrum, row pass, maps)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import godotoceanwaves_b200 as gow
from oracle import pyoracle as po
from conftest import demo_params

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
C = int(sys.argv[2]) if len(sys.argv) > 2 else 1
frames = int(sys.argv[3]) if len(sys.argv) > 3 else 1
pg = [demo_params(gow.WaveCascadeParameters, c) for c in range(C)]
pc = [demo_params(po.CascadeParams, c) for c in range(C)]
g = gow.WaveGenerator(); g.map_size = N; g.init_gpu(max(2, C)); g.enable_f32_taps(True)
o = po.OracleWaveGenerator(N)
for f in range(frames):
    g.update_all(0.02, pg); o.update_all(0.02, pc)
def where(diff):
    idx = np.argwhere(diff)
    if len(idx) == 0: return ""
    rows = np.unique(idx[:, -3] if idx.shape[1] >= 3 else idx[:, 0]); cols = np.unique(idx[:, -2] if idx.shape[1] >= 3 else idx[:, 1])
    return f"rows {rows[:10]}{'...' if len(rows) > 10 else ''} ({len(rows)}) cols {cols[:10]}{'...' if len(cols) > 10 else ''} ({len(cols)})"
for c in range(C):
    sp = g.spectrum_to_host(c)
    d = sp != o.spectrum[c]
    print(f"cascade {c}: spectrum mismatches {d.sum()}/{d.size}", where(d))
    rp = g.rowpass_to_host(c)                       # [4][N][N][2]
    ref = np.ascontiguousarray(np.swapaxes(o.fft_buffer[c, 0], 1, 2))
    for l in range(4):
        d = rp[l] != ref[l]
        rel = np.abs(rp[l].astype(np.float64) - ref[l]).max() / max(1e-30, np.abs(ref[l]).max())
        print(f"  rowpass layer {l}: mismatches {d.sum()}/{d.size} rel {rel:.3e}", where(d))
    d32, n32 = g.f32_maps_to_host(c)
    for name, a, b in (("disp", d32, o.displacement_f32[c]), ("normal", n32, o.normal_f32[c])):
        for ch in range(4):
            d = a[..., ch] != b[..., ch]
            rel = np.abs(a[..., ch].astype(np.float64) - b[..., ch]).max() / max(1e-30, np.abs(b[..., ch]).max())
            ys, xs = np.nonzero(d)
            print(f"  {name}[{ch}]: mismatches {d.sum()}/{d.size} rel {rel:.3e} rows {np.unique(ys)[:8]} cols {np.unique(xs)[:8]}")
g.enable_f32_taps(False)
g2 = gow.WaveGenerator(); g2.map_size = N; g2.init_gpu(max(2, C))
pg2 = [demo_params(gow.WaveCascadeParameters, c) for c in range(C)]
o2 = po.OracleWaveGenerator(N); pc2 = [demo_params(po.CascadeParams, c) for c in range(C)]
for f in range(frames):
    g2.update_all(0.02, pg2); o2.update_all(0.02, pc2)
dd, nn = g2.maps_to_host(0, C)
print("taps off: displacement texture mismatches", int((dd.view(np.uint16) != o2.displacement_map[:C]).sum()), "normal", int((nn.view(np.uint16) != o2.normal_map[:C]).sum()))
g.free(); g2.free()
