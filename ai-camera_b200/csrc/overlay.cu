// N3: track / detection overlay drawn on DEVICE frames.
//
// Replaces the drawing loop of /root/reference/src/utils/visualization.py (draw_tracks :72-124, draw_detections :9-69,
// draw_fps :127-167, draw_info_panel :170-227), which runs cv2.rectangle / cv2.putText on a host copy of every frame
// (src/aicamera_tracker.py:211-225).  Here the frames stay where the detector and the tracker read them:
//
//   * the box outline (cv2.rectangle(.., thickness 2)) and filled rectangles are rasterised analytically - the pixel set
//     OpenCV produces for thickness 2 is the 3-pixel band around the rectangle minus its four outer corner pixels
//     (thick lines are filled quads plus radius-1 end caps; checked against cv2 4.13 in tests/test_overlay.py);
//   * anti-aliased text is a DECAL: ai-camera_b200/visualization.py renders a label once with cv2 itself, on a black and
//     on a white canvas, and keeps per pixel and channel the affine map (A, B) with out = B + dst * A / 255 that
//     reproduces both; labels are cached by their text ("ID:17 person"), so steady state uploads nothing.  Inside the
//     label's filled background A = 0 (the decal is exactly what cv2 draws); on the few anti-aliased fringe pixels
//     outside it the affine map is within one grey level of cv2's integer blend.
//
// One CTA per frame walks the frame's items IN ORDER (a block barrier between items), threads over the pixels of an
// item: later items overwrite earlier ones exactly as sequential cv2 calls do.
#include "common.cuh"

namespace aicam {

extern void count_launch();

namespace {

constexpr int OVERLAY_THREADS = 512;

__global__ void __launch_bounds__(OVERLAY_THREADS) overlay_kernel(uint8_t* __restrict__ frames, int h, int w,
                                                                  const aicam_overlay_item* __restrict__ items,
                                                                  const int* __restrict__ item_start,
                                                                  const uint8_t* __restrict__ atlas, int slot_w, int slot_h) {
  const int n = blockIdx.x;
  uint8_t* f = frames + static_cast<long long>(n) * h * w * 3;
  const int i0 = item_start[n], i1 = item_start[n + 1];
  for (int i = i0; i < i1; ++i) {
    const aicam_overlay_item it = items[i];
    const uint8_t cb = it.color & 255, cg = (it.color >> 8) & 255, cr = (it.color >> 16) & 255;
    if (it.type == 2) {
      // decal: it.x1, it.y1 = frame position of the decal's first pixel, it.x2, it.y2 = its width / height
      const uint8_t* d = atlas + static_cast<long long>(it.slot) * slot_h * slot_w * 8;
      const int dw = min(it.x2, slot_w), dh = min(it.y2, slot_h);
      for (int p = threadIdx.x; p < dw * dh; p += OVERLAY_THREADS) {
        const int dy = p / dw, dx = p - dy * dw;
        const int x = it.x1 + dx, y = it.y1 + dy;
        if (x < 0 || x >= w || y < 0 || y >= h) continue;
        const uint2 ab = *reinterpret_cast<const uint2*>(d + (static_cast<long long>(dy) * slot_w + dx) * 8);
        if (ab.y == 0x00ffffffu && ab.x == 0u) continue;  // pass-through pixel
        uint8_t* o = f + (static_cast<long long>(y) * w + x) * 3;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const int B = (ab.x >> (8 * c)) & 255, A = (ab.y >> (8 * c)) & 255;
          o[c] = static_cast<uint8_t>(B + (o[c] * A + 127) / 255);
        }
      }
    } else {
      const int xa = min(it.x1, it.x2), xb = max(it.x1, it.x2), ya = min(it.y1, it.y2), yb = max(it.y1, it.y2);
      // filled: the inclusive rectangle; outline: its 3-pixel band
      const int g = it.type == 0 ? 1 : 0;
      const int bx0 = max(xa - g, 0), bx1 = min(xb + g, w - 1), by0 = max(ya - g, 0), by1 = min(yb + g, h - 1);
      if (bx0 <= bx1 && by0 <= by1) {
        const int bw = bx1 - bx0 + 1;
        const long long total = static_cast<long long>(bw) * (by1 - by0 + 1);
        if (it.type == 1) {
          for (long long p = threadIdx.x; p < total; p += OVERLAY_THREADS) {
            const int y = by0 + static_cast<int>(p / bw), x = bx0 + static_cast<int>(p % bw);
            uint8_t* o = f + (static_cast<long long>(y) * w + x) * 3;
            o[0] = cb; o[1] = cg; o[2] = cr;
          }
        } else {
          // rows of the top and bottom bands whole, the other rows only their left and right bands (6 pixels)
          const int rows = by1 - by0 + 1;
          for (int ry = threadIdx.x / 32; ry < rows; ry += OVERLAY_THREADS / 32) {
            const int y = by0 + ry;
            const bool band = y <= ya + 1 || y >= yb - 1;
            const int lane = threadIdx.x & 31;
            const int count = band ? bw : 6;
            for (int k = lane; k < count; k += 32) {
              const int x = band ? bx0 + k : (k < 3 ? xa - 1 + k : xb - 4 + k);
              if (x < bx0 || x > bx1) continue;
              if (!band && k >= 3 && x <= xa + 1) continue;  // narrow rectangle: the two bands overlap, written once
              const bool corner = (x == xa - 1 || x == xb + 1) && (y == ya - 1 || y == yb + 1);
              if (corner) continue;
              uint8_t* o = f + (static_cast<long long>(y) * w + x) * 3;
              o[0] = cb; o[1] = cg; o[2] = cr;
            }
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace
}  // namespace aicam

using namespace aicam;

extern "C" int aicam_overlay_draw(uint8_t* frames_bgr, int batch, int h, int w, const aicam_overlay_item* items,
                                  const int32_t* item_start, const uint8_t* atlas, int slot_w, int slot_h, void* stream) {
  if (!frames_bgr || !items || !item_start || batch < 0 || h <= 0 || w <= 0)
    return fail(AICAM_ERR_INVALID_ARG, "overlay_draw: bad arguments");
  if (slot_w < 0 || slot_h < 0 || (!atlas && slot_w * slot_h != 0)) return fail(AICAM_ERR_INVALID_ARG, "overlay_draw: bad atlas");
  if (reinterpret_cast<uintptr_t>(atlas) % 8) return fail(AICAM_ERR_INVALID_ARG, "overlay_draw: the atlas must be 8-byte aligned");
  if (batch == 0) return AICAM_OK;
  overlay_kernel<<<batch, OVERLAY_THREADS, 0, static_cast<cudaStream_t>(stream)>>>(frames_bgr, h, w, items, item_start, atlas, slot_w, slot_h);
  count_launch();
  return last_launch("overlay_kernel");
}
