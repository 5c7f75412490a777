#!/bin/bash
# compute-sanitizer passes over a small but queue-exercising workload (run under gpurun): 256x256 x 12 cascades, 3 updates + 5 fused frames.
cat > /tmp/san.py <<'PY'
import sys, numpy as np
sys.path.insert(0, ".")
import godotoceanwaves_b200 as gow
from bench import synth_params
g = gow.WaveGenerator(); g.map_size = 256; g.init_gpu(12)
p = [synth_params(gow.WaveCascadeParameters, c) for c in range(12)]
for _ in range(3):
    g.update_all(0.02, p)
g.update_frames(0.02, p, 5)        # fused frames: alternating scratch halves, column pass waiting for the previous frame's
d, n = g.maps_to_host()
pts = np.random.default_rng(0).uniform(-100, 100, (4096, 2)).astype(np.float32)
g.sample(pts, gow.WaveGenerator.map_scales(p))
print("checksum", int(d.view(np.uint16).astype(np.uint64).sum()), int(n.view(np.uint16).astype(np.uint64).sum()))
g.free()
PY
for tool in memcheck racecheck synccheck; do
  echo "== $tool"; timeout 600 compute-sanitizer --tool $tool python /tmp/san.py 2>&1 | grep -i "checksum\|ERROR SUMMARY\|hazard\|error" | head -8
done
