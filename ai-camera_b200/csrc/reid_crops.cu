// K5: detection filter + ROI crop / resize / normalise of every tracked detection of every
// stream into one batched ReID input tensor.
//
// Replaces DeepSORT.update steps 1-2 (confidence/class filter,
// /root/reference/src/tracker/deepsort_tracker.py:88-101), _extract_image_crops (:143-159,
// int() truncation + clamp; empty rectangle -> no feature), preprocess_reid_input
// (src/utils/image_processing.py:105-138: cv2.resize INTER_LINEAR to 64x128, BGR->RGB,
// (x/255 - mean)/std, CHW) and the concat + H2D of src/tracker/reid_model.py:83-101.
//
// The resize is OpenCV's fixed-point bilinear (see preprocess.cu), evaluated per crop: the
// coefficient tables depend on the crop size, so each CTA derives them in shared memory with
// the same double -> float -> 11-bit rounding steps (explicit _rn intrinsics; the file is also
// compiled with --fmad=false), then every thread interpolates its pixels.  Result: the uint8
// resized crop, and therefore the float tensor, equals the reference's bit for bit.
// HBM-bound: reads ~crop bytes, writes 128*64*8 B (NHWC4 bf16) or 128*64*16 B (NHWC8, what the fused ReID stem reads) per crop.
#include "common.cuh"
#include "frame_src.cuh"

namespace aicam {

extern void count_launch();

namespace {

constexpr int RH = AICAM_REID_H, RW = AICAM_REID_W;

// One block, one WARP per frame: stable filter of the frame's detections (ballot compaction, lanes over
// detections), crop rectangles, and (via a block-wide exclusive scan over frames) deterministic crop rows.
__global__ void __launch_bounds__(1024) filter_kernel(const float* __restrict__ boxes, const float* __restrict__ scores,
                                                      const int* __restrict__ labels, const int* __restrict__ num_dets,
                                                      int batch, int stride_k, int h, int w, float min_conf,
                                                      unsigned long long mask_lo, unsigned long long mask_hi,
                                                      int max_crops, int* __restrict__ det_index,
                                                      int* __restrict__ det_count, int* __restrict__ crop_slot,
                                                      int* __restrict__ crop_rect, int* __restrict__ crop_count) {
  __shared__ int s_scan[32];  // valid crops per frame of the current group of 32 frames, then their exclusive scan
  __shared__ int s_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  if (threadIdx.x == 0) s_base = 0;
  __syncthreads();
  for (int b0 = 0; b0 < batch; b0 += 32) {
    const int b = b0 + warp;
    int nvalid = 0, nkeep = 0;
    if (b < batch) {
      const int n = min(num_dets[b], stride_k);
      const long long fo = static_cast<long long>(b) * stride_k;
      for (int i0 = 0; i0 < n; i0 += 32) {
        const int i = i0 + lane;
        bool keep = false, ok = false;
        if (i < n) {
          const int cls = labels[fo + i];
          const bool tracked = cls >= 0 && cls < 128 && (((cls < 64 ? mask_lo >> cls : mask_hi >> (cls - 64)) & 1ull) != 0);
          keep = scores[fo + i] >= min_conf && tracked;
          if (keep) {
            const float4 bx = reinterpret_cast<const float4*>(boxes)[fo + i];
            const int x1 = max(0, static_cast<int>(bx.x)), y1 = max(0, static_cast<int>(bx.y));
            const int x2 = min(w, static_cast<int>(bx.z)), y2 = min(h, static_cast<int>(bx.w));
            ok = x1 < x2 && y1 < y2;
          }
        }
        const unsigned km = __ballot_sync(0xffffffffu, keep), vm = __ballot_sync(0xffffffffu, ok);
        if (keep) {
          const int k = nkeep + __popc(km & lt);
          det_index[fo + k] = i;
          // provisional: local crop ordinal or -1; rebased after the scan
          crop_slot[fo + k] = ok ? nvalid + __popc(vm & lt) : -1;
        }
        nkeep += __popc(km);
        nvalid += __popc(vm);
      }
      if (lane == 0) det_count[b] = nkeep;
    }
    if (lane == 0) s_scan[warp] = nvalid;
    __syncthreads();
    if (warp == 0) {  // exclusive scan of the group's 32 counts
      const int v = s_scan[lane];
      int incl = v;
      for (int o = 1; o < 32; o <<= 1) {
        const int u = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += u;
      }
      s_scan[lane] = s_base + incl - v;
      __syncwarp();
      if (lane == 31) s_base += incl;
    }
    __syncthreads();
    if (b < batch) {
      const int base = s_scan[warp];
      const long long fo = static_cast<long long>(b) * stride_k;
      for (int k = lane; k < nkeep; k += 32) {  // (my own writes of det_index / crop_slot: same warp, ordered by the barriers)
        const int local = crop_slot[fo + k];
        if (local < 0) continue;
        const int slot = base + local;
        if (slot < max_crops) {
          const float4 bx = reinterpret_cast<const float4*>(boxes)[fo + det_index[fo + k]];
          crop_slot[fo + k] = slot;
          crop_rect[slot * 5 + 0] = b;
          crop_rect[slot * 5 + 1] = max(0, static_cast<int>(bx.x));
          crop_rect[slot * 5 + 2] = max(0, static_cast<int>(bx.y));
          crop_rect[slot * 5 + 3] = min(w, static_cast<int>(bx.z));
          crop_rect[slot * 5 + 4] = min(h, static_cast<int>(bx.w));
        } else {
          crop_slot[fo + k] = -1;
        }
      }
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    crop_count[0] = min(s_base, max_crops);
    // high-water mark of the crops WANTED (un-clamped): > max_crops means detections lost their feature to the
    // capacity (the reference always extracts one); sticky until the caller zeroes it
    if (s_base > crop_count[1]) crop_count[1] = s_base;
  }
}

// cv2 linear-resize tables for one axis, one entry per thread
__device__ __forceinline__ void axis_entry(int src, int dst, int d, bool horizontal, int* i0, int* i1, int* w0, int* w1) {
  const double scale = __ddiv_rn(1.0, __ddiv_rn(static_cast<double>(dst), static_cast<double>(src)));
  float f = __double2float_rn(__dsub_rn(__dmul_rn(static_cast<double>(d) + 0.5, scale), 0.5));
  int s = __float2int_rd(f);
  f = __fsub_rn(f, static_cast<float>(s));
  if (horizontal) {
    if (s < 0) { f = 0.0f; s = 0; }
    if (s >= src - 1) { f = 0.0f; s = src - 1; }
    *i0 = s;
    *i1 = min(s + 1, src - 1);
  } else {
    *i0 = min(max(s, 0), src - 1);
    *i1 = min(max(s + 1, 0), src - 1);
  }
  *w0 = __float2int_rn(__fmul_rn(__fsub_rn(1.0f, f), 2048.0f));
  *w1 = __float2int_rn(__fmul_rn(f, 2048.0f));
}

template <int FORMAT, int SRC>
__global__ void __launch_bounds__(256) crop_kernel(const uint8_t* __restrict__ frames, int h, int w,
                                                   const int* __restrict__ crop_rect,
                                                   const int* __restrict__ crop_count, void* __restrict__ out) {
  const int slot = blockIdx.x;
  if (slot >= *crop_count) return;
  __shared__ int tx[4][RW];
  __shared__ int ty[4][RH];
  const int b = crop_rect[slot * 5], x1 = crop_rect[slot * 5 + 1], y1 = crop_rect[slot * 5 + 2];
  const int cw = crop_rect[slot * 5 + 3] - x1, ch = crop_rect[slot * 5 + 4] - y1;
  const int mode = (cw == RW && ch == RH) ? 2 : ((cw == 2 * RW && ch == 2 * RH) ? 1 : 0);
  if (mode == 0) {
    const int t = threadIdx.x;
    if (t < RW) axis_entry(cw, RW, t, true, &tx[0][t], &tx[1][t], &tx[2][t], &tx[3][t]);
    else if (t < RW + RH) axis_entry(ch, RH, t - RW, false, &ty[0][t - RW], &ty[1][t - RW], &ty[2][t - RW], &ty[3][t - RW]);
  }
  __syncthreads();
  const FrameSrc<SRC> src{frames + b * FrameSrc<SRC>::frame_bytes(h, w), h, w};
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  for (int p = threadIdx.x; p < RH * RW; p += blockDim.x) {
    const int oy = p / RW, ox = p - oy * RW;
    int v[3];
    if (mode == 2) {
      src.pix(y1 + oy, x1 + ox, v);
    } else if (mode == 1) {
      int p00[3], p01[3], p10[3], p11[3];
      src.pix(y1 + 2 * oy, x1 + 2 * ox, p00); src.pix(y1 + 2 * oy, x1 + 2 * ox + 1, p01);
      src.pix(y1 + 2 * oy + 1, x1 + 2 * ox, p10); src.pix(y1 + 2 * oy + 1, x1 + 2 * ox + 1, p11);
#pragma unroll
      for (int c = 0; c < 3; ++c) v[c] = (p00[c] + p01[c] + p10[c] + p11[c] + 2) >> 2;
    } else {
      const int sx0 = x1 + tx[0][ox], sx1 = x1 + tx[1][ox], a0 = tx[2][ox], a1 = tx[3][ox];
      const int sy0 = y1 + ty[0][oy], sy1 = y1 + ty[1][oy], b0 = ty[2][oy], b1 = ty[3][oy];
      // taps with a zero weight are not fetched
      int p00[3] = {0, 0, 0}, p01[3] = {0, 0, 0}, p10[3] = {0, 0, 0}, p11[3] = {0, 0, 0};
      if (b0 != 0) {
        src.pix(sy0, sx0, p00);
        if (a1 != 0) src.pix(sy0, sx1, p01);
      }
      if (b1 != 0) {
        src.pix(sy1, sx0, p10);
        if (a1 != 0) src.pix(sy1, sx1, p11);
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const int h0 = p00[c] * a0 + p01[c] * a1, h1 = p10[c] * a0 + p11[c] * a1;
        const int o = (((b0 * (h0 >> 4)) >> 16) + ((b1 * (h1 >> 4)) >> 16) + 2) >> 2;
        v[c] = min(max(o, 0), 255);
      }
    }
    float rgb[3];
#pragma unroll
    for (int c = 0; c < 3; ++c)  // BGR -> RGB, (x/255 - mean)/std in float32
      rgb[c] = __fdiv_rn(__fsub_rn(__fdiv_rn(static_cast<float>(v[2 - c]), 255.0f), mean[c]), stdv[c]);
    if (FORMAT == 0) {
      float* o = static_cast<float*>(out) + static_cast<long long>(slot) * 3 * RH * RW + p;
      o[0] = rgb[0]; o[RH * RW] = rgb[1]; o[2 * RH * RW] = rgb[2];
    } else if (FORMAT == 1) {
      uint2 q;
      q.x = pack_bf16x2(rgb[0], rgb[1]);
      q.y = pack_bf16x2(rgb[2], 0.0f);
      reinterpret_cast<uint2*>(out)[static_cast<long long>(slot) * RH * RW + p] = q;
    } else {  // NHWC8: one pixel = one 16-byte K chunk of the fused ReID stem (stem_pool.cu)
      reinterpret_cast<uint4*>(out)[static_cast<long long>(slot) * RH * RW + p] =
          make_uint4(pack_bf16x2(rgb[0], rgb[1]), pack_bf16x2(rgb[2], 0.0f), 0u, 0u);
    }
  }
}

}  // namespace
}  // namespace aicam

using namespace aicam;

namespace aicam {
namespace {
template <int SRC>
int reid_crops_impl(const uint8_t* frames, int batch, int h, int w, const float* boxes, const float* scores,
                    const int32_t* labels, const int32_t* num_dets, int stride_k, float min_confidence,
                    uint64_t class_mask_lo, uint64_t class_mask_hi, int format, int max_crops,
                    int32_t* det_index, int32_t* det_count, int32_t* crop_slot, int32_t* crop_rect,
                    void* crops, int32_t* crop_count, void* stream) {
  if (!boxes || !scores || !labels || !num_dets || !det_index || !det_count || !crop_slot || !crop_rect || !crop_count)
    return fail(AICAM_ERR_INVALID_ARG, "reid_crops: null argument");
  if (batch < 0 || stride_k <= 0 || h <= 0 || w <= 0 || max_crops < 0 || (format < 0 || format > 2))
    return fail(AICAM_ERR_INVALID_ARG, "reid_crops: bad shape arguments");
  if (reinterpret_cast<uintptr_t>(boxes) % 16) return fail(AICAM_ERR_INVALID_ARG, "reid_crops: boxes must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  filter_kernel<<<1, 1024, 0, st>>>(boxes, scores, labels, num_dets, batch, stride_k, h, w, min_confidence,
                                    class_mask_lo, class_mask_hi, max_crops, det_index, det_count, crop_slot, crop_rect,
                                    crop_count);
  count_launch();
  if (int rc = last_launch("filter_kernel")) return rc;
  if (max_crops == 0 || !frames || !crops) return AICAM_OK;  // filter only
  if (format == 0)
    crop_kernel<0, SRC><<<max_crops, 256, 0, st>>>(frames, h, w, crop_rect, crop_count, crops);
  else if (format == 1)
    crop_kernel<1, SRC><<<max_crops, 256, 0, st>>>(frames, h, w, crop_rect, crop_count, crops);
  else
    crop_kernel<2, SRC><<<max_crops, 256, 0, st>>>(frames, h, w, crop_rect, crop_count, crops);
  count_launch();
  return last_launch("crop_kernel");
}
}  // namespace
}  // namespace aicam

extern "C" int aicam_reid_crops(const uint8_t* frames, int batch, int h, int w, const float* boxes, const float* scores,
                                const int32_t* labels, const int32_t* num_dets, int stride_k, float min_confidence,
                                uint64_t class_mask_lo, uint64_t class_mask_hi, int format, int max_crops,
                                int32_t* det_index, int32_t* det_count, int32_t* crop_slot, int32_t* crop_rect,
                                void* crops, int32_t* crop_count, void* stream) {
  return reid_crops_impl<0>(frames, batch, h, w, boxes, scores, labels, num_dets, stride_k, min_confidence, class_mask_lo,
                            class_mask_hi, format, max_crops, det_index, det_count, crop_slot, crop_rect, crops, crop_count, stream);
}

extern "C" int aicam_reid_crops_nv12(const uint8_t* frames_nv12, int batch, int h, int w, const float* boxes, const float* scores,
                                     const int32_t* labels, const int32_t* num_dets, int stride_k, float min_confidence,
                                     uint64_t class_mask_lo, uint64_t class_mask_hi, int format, int max_crops,
                                     int32_t* det_index, int32_t* det_count, int32_t* crop_slot, int32_t* crop_rect,
                                     void* crops, int32_t* crop_count, void* stream) {
  if (h % 2 || w % 2) return fail(AICAM_ERR_INVALID_ARG, "reid_crops_nv12: NV12 frames have even height and width");
  return reid_crops_impl<1>(frames_nv12, batch, h, w, boxes, scores, labels, num_dets, stride_k, min_confidence, class_mask_lo,
                            class_mask_hi, format, max_crops, det_index, det_count, crop_slot, crop_rect, crops, crop_count, stream);
}

// NV12 -> packed BGR (cv2.cvtColor(.., COLOR_YUV2BGR_NV12)): the frames the NV12 entry points see, as the reference's
// decoder would have handed them out; for parity tests and for callers that keep a BGR copy.
namespace aicam {
namespace {
__global__ void __launch_bounds__(256) nv12_to_bgr_kernel(const uint8_t* __restrict__ nv12, int h, int w, uint8_t* __restrict__ bgr,
                                                          long long total) {
  const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;  // one thread per 2 horizontal pixels
  if (idx >= total) return;
  const int xp = static_cast<int>(idx % (w / 2));
  const long long t = idx / (w / 2);
  const int y = static_cast<int>(t % h);
  const long long n = t / h;
  const uint8_t* f = nv12 + n * FrameSrc<1>::frame_bytes(h, w);
  const uint8_t* c = f + static_cast<long long>(h) * w + static_cast<long long>(y >> 1) * w + 2 * xp;
  const int u = __ldg(c), v = __ldg(c + 1);
  int p0[3], p1[3];
  yuv_to_bgr_601(__ldg(f + static_cast<long long>(y) * w + 2 * xp), u, v, p0);
  yuv_to_bgr_601(__ldg(f + static_cast<long long>(y) * w + 2 * xp + 1), u, v, p1);
  uint8_t* o = bgr + ((n * h + y) * w + 2 * xp) * 3;
  o[0] = p0[0]; o[1] = p0[1]; o[2] = p0[2]; o[3] = p1[0]; o[4] = p1[1]; o[5] = p1[2];
}
}  // namespace
}  // namespace aicam

extern "C" int aicam_nv12_to_bgr(const uint8_t* frames_nv12, int batch, int h, int w, uint8_t* frames_bgr, void* stream) {
  if (!frames_nv12 || !frames_bgr || batch < 0 || h <= 0 || w <= 0 || h % 2 || w % 2)
    return fail(AICAM_ERR_INVALID_ARG, "nv12_to_bgr: bad arguments (NV12 frames have even height and width)");
  if (batch == 0) return AICAM_OK;
  const long long total = static_cast<long long>(batch) * h * (w / 2);
  nv12_to_bgr_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(frames_nv12, h, w, frames_bgr, total);
  count_launch();
  return last_launch("nv12_to_bgr_kernel");
}
