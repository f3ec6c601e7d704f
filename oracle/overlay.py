"""numpy restatement of the overlay kernel's item semantics (oracle; test infrastructure - only tests/ may import this).

Follows ``ai-camera_b200/csrc/overlay.cu`` item by item and is itself pinned to OpenCV: ``tests/test_overlay.py`` checks
type 0 / type 1 against ``cv2.rectangle`` (the calls of ``/root/reference/src/utils/visualization.py:43,50-56,97,104-110``)
over random, clipped, degenerate and swapped rectangles, and whole overlays against the reference's own ``draw_tracks`` /
``draw_info_panel`` / ``draw_fps``."""
import numpy as np


def draw_items(frame, items, atlas):
    """frame uint8 [H, W, 3] (modified in place); items: (type, x1, y1, x2, y2, color, slot) in order; atlas uint8
    [slots, slot_h, slot_w, 8]."""
    H, W = frame.shape[:2]
    for (typ, x1, y1, x2, y2, color, slot) in items:
        col = np.array([color & 255, (color >> 8) & 255, (color >> 16) & 255], np.uint8)
        if typ == 2:
            dw, dh = min(x2, atlas.shape[2]), min(y2, atlas.shape[1])
            for dy in range(dh):
                y = y1 + dy
                if y < 0 or y >= H:
                    continue
                xs0, xs1 = max(0, -x1), min(dw, W - x1)
                if xs0 >= xs1:
                    continue
                d = atlas[slot, dy, xs0:xs1].astype(np.int32)
                dst = frame[y, x1 + xs0:x1 + xs1].astype(np.int32)
                frame[y, x1 + xs0:x1 + xs1] = (d[:, 0:3] + (dst * d[:, 4:7] + 127) // 255).astype(np.uint8)
            continue
        xa, xb, ya, yb = min(x1, x2), max(x1, x2), min(y1, y2), max(y1, y2)
        if typ == 1:
            bx0, bx1, by0, by1 = max(xa, 0), min(xb, W - 1), max(ya, 0), min(yb, H - 1)
            if bx0 <= bx1 and by0 <= by1:
                frame[by0:by1 + 1, bx0:bx1 + 1] = col
            continue
        # thickness-2 outline: the 3-pixel band around the rectangle minus its four outer corner pixels
        for y in range(max(ya - 1, 0), min(yb + 1, H - 1) + 1):
            for x in range(max(xa - 1, 0), min(xb + 1, W - 1) + 1):
                inner = xa + 2 <= x <= xb - 2 and ya + 2 <= y <= yb - 2
                corner = x in (xa - 1, xb + 1) and y in (ya - 1, yb + 1)
                if not inner and not corner:
                    frame[y, x] = col
    return frame
