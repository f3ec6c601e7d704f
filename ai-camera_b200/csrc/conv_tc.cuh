// Implicit-GEMM convolution on the 5th-gen tensor cores (tcgen05 + TMEM), sm_100a.
// See conv_tc.cu for the design notes.
#pragma once
#include "common.cuh"

namespace aicam {

// Packed weights: bf16 [q_pad][cout][8] where a "chunk" q holds 8 consecutive K elements
// (16 bytes), K index = tap * cin_pad + c.  q_pad is even (one tcgen05.mma consumes K = 16).
struct PackedConv {
  __nv_bfloat16* w = nullptr;  // device
  float* bias = nullptr;       // device, [cout]
  float* bias_host = nullptr;  // host copy, [cout padded to 16]: the TMA-epilogue kernels take the bias as a kernel argument
                               // (constant bank) - a shared-memory copy cost a third of their LDS wavefronts
  int cin = 0;        // real input channels
  int cin_pad = 0;    // 4 (stem: 3 -> 4) or a multiple of 8
  int cout = 0, ksize = 1, stride = 1;
  int q = 0, q_pad = 0;
  int s2d_c0 = 0;  // != 0: a 3x3 stride-2 layer packed as a 2x2 window over 2x2 input blocks of s2d_c0 channels (conv_win.cu)
  // Layers wider than one n-tile keep a second copy [cout_pad / nt_block][q_pad][nt_block][8]: the weights one
  // (K slab, n-tile) stage needs are then one contiguous run = one bulk copy instead of one per K chunk
  __nv_bfloat16* w_nt = nullptr;
  int nt_block = 0;
  // 3x3 stride-1 layers with 64 / 128 output channels: [rank][K / 8][cout / 2][8], the halves of the weight columns the
  // two CTAs of a cta_group::2 pair keep (conv_pair.cu); made by pack_pair_weights for layers that run zero-bordered
  __nv_bfloat16* w_pair = nullptr;
};

struct ConvLaunch {
  const __nv_bfloat16* in;
  long long in_img_stride;  // elements between images
  int in_cstride, in_coff;  // channels per pixel in the buffer, first channel used
  int batch, h, w;          // input spatial size
  int ho, wo;
  void* out;
  long long out_img_stride;
  int out_cstride, out_coff;
  int out_f32;
  const __nv_bfloat16* res;
  long long res_img_stride;
  int res_cstride, res_coff;
  int res_mode;  // 0 none, 1 act(conv)+res, 2 act(conv+res)
  int act;       // 0 none, 1 SiLU, 2 ReLU
  const int* batch_dev = nullptr;  // optional device-side image count (<= batch)
  int out_s2d = 0;                 // store the output space-to-depth: [ho/2][wo/2][2x2 sub-pixel][out_cstride] (window kernel only)
  long long* trace = nullptr;      // debug: clock64 stamps of CTA 0
  // "Padded" tensors carry a one-pixel zero border in memory: [batch][h + 2][w + 2][cstride], interior at (1, 1);
  // h / w / ho / wo stay the logical sizes and the image strides describe the padded images.  The border is
  // never written, so a 3x3 stride-1 layer can run over the flat raster of ALL padded pixels (window kernel,
  // operand mode 4) without per-image halo handling.  The residual shares the output's layout.
  //   pad = 1  symmetric border: [h + 2][w + 2], interior at (1, 1);
  //   pad = 2  SHARED border: [h + 1][w + 1], interior at (0, 0).  The zero column x = w is the right border of its row AND
  //            the left border of the next row (the raster wraps), the zero row y = h the bottom border of its image AND
  //            the top border of the next image; what lies above the first image are the TMA unit's out-of-bounds zeros.
  //            Same flat-raster arithmetic, fewer wasted positions: h w / ((h + 1)(w + 1)) useful instead of
  //            h w / ((h + 2)(w + 2)) - 71 % instead of 53 % on the 8 x 4 maps of ReID layer 4.
  int in_pad = 0, out_pad = 0;
  // Detect-head 1x1 layers (window kernel, generic epilogue): instead of storing the fp32 head rows, the epilogue decodes them
  // where they are staged in shared memory (same arithmetic as decode_kernel, detect_post.cu) and writes the dense per-anchor
  // arrays the NMS kernel reads.  1: box branch (64 DFL logits -> xyxy), 2: class branch (nc logits -> score, label).
  int decode = 0;
  int dec_anchor_base = 0, dec_anchors = 0;  // first anchor of this Detect level inside an image's anchors, anchors per image
  float* dec_boxes = nullptr;                // [batch][anchors][4]
  float* dec_scores = nullptr;               // [batch][anchors]
  int* dec_labels = nullptr;                 // [batch][anchors]
};

inline int pad_lo(int pad) { return pad == 1 ? 1 : 0; }                    // interior offset (rows and columns)
inline int pad_ext(int pad) { return pad == 1 ? 2 : (pad == 2 ? 1 : 0); }  // extra rows / columns an image carries

// host: pack OIHW fp32 weights (host) into the device layout above
int pack_conv_weights(const float* w_oihw, const float* bias, int cout, int cin, int ksize, int stride,
                      PackedConv* out);
// 3x3 / stride 2 / pad 1 layer re-expressed over its space-to-depth input [h/2][w/2][2x2 sub-pixel][c0_pad]:
// a 2x2 stride-1 window (pad 1 top / left) with K = tap2 * (4 c0_pad) + (sy * 2 + sx) * c0_pad + c; cin <= c0_pad in {4, 16}
int pack_conv_weights_s2d(const float* w_oihw, const float* bias, int cout, int cin, int c0_pad, PackedConv* out);
int pack_pair_weights(PackedConv* p);
void free_packed_conv(PackedConv* p);
int launch_conv(const PackedConv& pc, const ConvLaunch& L, cudaStream_t stream);

}  // namespace aicam
