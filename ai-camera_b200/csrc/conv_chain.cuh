// Chains of stride-1 convolutions fused into one tcgen05 kernel (conv_chain.cu): the intermediate activations of a
// 3x3 -> 3x3 (-> 1x1) chain never leave shared memory.
#pragma once
#include "conv_tc.cuh"

namespace aicam {

constexpr int CHAIN_MAX_STAGES = 5;

// One stage: a 1x1 or 3x3 stride-1 convolution (+ bias, activation, optional residual) whose input channels are the
// concatenation of up to two channel ranges of earlier buffers.  Buffer 0 is the chain's input patch, buffer i >= 1 the
// output of stage i - 1 (shared memory only); the LAST stage's output is the only thing stored to global memory.
struct ChainStageSpec {
  const PackedConv* pc = nullptr;  // packed weights [K / 8][cout_pad][8], K = tap * cin + (source-concatenated channel)
  int act = 0;                     // 0 none, 1 SiLU, 2 ReLU
  int nsrc = 1;
  int src_buf[2] = {0, 0};
  int src_coff[2] = {0, 0};        // first channel inside that buffer (multiple of 16)
  int src_c[2] = {0, 0};           // channels taken (multiple of 16); their sum is the stage's cin
  int res_buf = -1;                // residual read from an earlier buffer (same pixel), channels [res_coff, res_coff + cout)
  int res_coff = 0;
  int res_mode = 0;                // 1: act(conv) + res, 2: act(conv + res)
};

struct ChainSpec {
  int nstages = 0;
  ChainStageSpec st[CHAIN_MAX_STAGES];
  const __nv_bfloat16* in = nullptr;  // NHWC bf16, dense images
  long long in_img_stride = 0;
  int in_cstride = 0, in_coff = 0, in_c = 0;  // channels per pixel in memory, first channel loaded, channels loaded (buffer 0)
  int batch = 0, h = 0, w = 0;
  void* out = nullptr;                // NHWC bf16 or fp32 (channel slice [out_coff, out_coff + cout_last) of out_cstride)
  long long out_img_stride = 0;
  int out_cstride = 0, out_coff = 0, out_f32 = 0;
  long long* trace = nullptr;         // debug: event trace of CTA 0 (aicam_conv_chain_bench)
};

// 1: launched, 0: this chain / geometry is not eligible (the caller launches the layers one by one), < 0: error
int try_launch_conv_chain(const ChainSpec& spec, cudaStream_t stream);

}  // namespace aicam
