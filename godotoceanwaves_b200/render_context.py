"""Mirror of the one RenderingContext helper the hot path's callers rely on for numerics:
``create_push_constant`` (assets/render_context.gd:122-135).  The rest of RenderingContext is a
Vulkan resource wrapper (RIDs, descriptor sets, pipelines) that libocean.so replaces wholesale
(see INTEGRATION.md); it has no counterpart here by design."""
from __future__ import annotations

import math
import struct


class RenderingContext:
    @staticmethod
    def create_push_constant(data) -> bytes:
        """Packs ints as s32 and floats as binary32 (round-to-nearest), zero-padded to a multiple of
        16 bytes; at most 128 bytes (render_context.gd:122-135)."""
        packed_size = len(data) * 4
        assert packed_size <= 128, "Push constant size must be at most 128 bytes!"
        padding = math.ceil(packed_size / 16.0) * 16 - packed_size
        out = bytearray(packed_size + (padding if padding > 0 else 0))
        for i, v in enumerate(data):
            if isinstance(v, (bool, int)):
                struct.pack_into("<i", out, i * 4, int(v))
            else:
                struct.pack_into("<f", out, i * 4, float(v))
        return bytes(out)
