// Fused chains of stride-1 convolutions on tcgen05 tensor cores, sm_100a.
//
// The layers the reference hands to its TensorRT engine (/root/reference/src/trt_utils/trt_engine.py:191, called at
// src/detector/yolo_detector.py:97) include many short chains over the SAME pixels: the Bottleneck pair of a C2f block
// (3x3 -> 3x3 + shortcut, followed by the 1x1 over the concatenation) and the Detect-head branches (3x3 -> 3x3 -> 1x1).
// Launched one layer at a time (conv_win.cu) every link costs a launch, a grid-wide drain and a round trip of the
// intermediate activation through HBM - and for 16/32-channel tensors through a TMA unit that moves one 32/64-byte
// pixel per request.  This kernel runs a whole chain per spatial tile and keeps the intermediates in shared memory:
//
//   * one TMA box load brings the input patch of a tile: (th + 2 m) x (tw + 2 m) pixels, m = the chain's halo;
//   * every buffer (the patch, each stage's output) lives in shared memory in the SAME raster (row pitch RW = tw + 2 m)
//     in the canonical K-major swizzled UMMA layout, split in <= 64-channel slabs.  As in conv_win.cu a 3x3 tap (dy, dx)
//     is the same buffer read through a descriptor whose start address is advanced by dy RW + dx rows, so a stage is
//     taps x (K / 16) MMAs per 128-row chunk with no data movement at all;
//   * the epilogue of a stage (TMEM -> registers -> bias / residual from an earlier buffer / activation -> bf16) writes
//     the next stage's A operand straight into its swizzled buffer; positions outside the image are written as zeros
//     (= the next layer's padding).  Only the last stage's rows go to a compact staging tile and out through TMA stores;
//   * halo positions are recomputed per tile (the price of not synchronising with the neighbours).
//
// Work items are 128-row chunks of a stage, in a fixed order per tile.  Chunk g (running index) owns TMEM slot g % nslot
// and is finished by epilogue warpgroup g % 4, so up to four chunks are in their epilogue while the MMA warp issues the
// next ones; mbarriers per (buffer, chunk) tell the MMA warp when the rows a chunk reads have been written.
//
// Warp roles (640 threads): 0-15 epilogue (4 warpgroups; TMEM lane quarter = warp % 4), 16 MMA issuer + TMEM owner,
// 17 patch producer, 18 weight producer (resident weights, or a ring of per-tap stages for deep layers), 19 store warp.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "conv_chain.cuh"
#include "tc_ptx.cuh"

namespace aicam {

extern void count_launch();
bool profile_begin(cudaStream_t st, size_t* slot);
void profile_end(cudaStream_t st, size_t slot);

namespace {

using namespace ptx;

constexpr int NWG = 4;
constexpr int CH_THREADS = (4 * NWG + 4) * 32;
constexpr int MAX_BUFS = CHAIN_MAX_STAGES;  // buffer 0 + the outputs of all stages but the last
constexpr int MAX_CHUNKS = 10;
constexpr int MAX_SLOTS = 16;
constexpr int MAX_RING = 4;
constexpr int MAX_PIECES = 6;
constexpr size_t SMEM_LIMIT = 227 * 1024;
constexpr uint32_t OFF_DATA = 2048;
constexpr size_t RESIDENT_LIMIT = 72 * 1024;
constexpr uint32_t RING_SLOT_MAX = 32 * 1024;

// mbarrier offsets from the start of shared memory
constexpr uint32_t BAR_P0_FULL = 0, BAR_P0_EMPTY = 16, BAR_B_FULL = 32, BAR_B_EMPTY = 64, BAR_W_FULL = 96, BAR_TILE_DONE = 104,
                   BAR_STAGE_FREE = 112, BAR_ACC_FULL = 128, BAR_ACC_EMPTY = 256, BAR_READY = 384, OFF_TMEM_PTR = 1024;

struct ChainBufD {
  uint32_t off;          // byte offset of slab 0 (1024-aligned)
  uint32_t slab_stride;  // bytes between slabs (1024-aligned)
  uint32_t row_bytes;    // 32 / 64 / 128 = slab channels x 2
  uint32_t xor_mask;     // swizzle: 16-byte chunk index ^= (row offset >> 7) & mask
  uint32_t a_hi;         // high word of the UMMA descriptor (SBO, version, swizzle mode)
  int c16_shift;         // log2(16-channel groups per slab)
  int nslabs, nchunks;   // (nchunks: 128-row chunks its producer stage writes)
};
struct ChainSrcD {
  int buf, c16_0, nc16;  // 16-channel groups [c16_0, c16_0 + nc16) of that buffer
  int shift;             // rows to add to the output row for tap (0, 0)
  int wk8;               // first K chunk (8 channels) of this source inside a tap's weight block
};
struct ChainStageD {
  int ksize, taps, nsrc;
  ChainSrcD src[2];
  int k8_per_tap;
  int n_pad;             // MMA N = output channels rounded to 16
  int act;
  int res_buf, res_c16_0, res_shift, res_mode;
  int dst_buf;           // shared-memory buffer this stage writes, -1 for the last stage
  int margin;            // rows q of this stage are pixels (Y0 - margin + q / RW, X0 - margin + q % RW)
  int nchunks, group, chunk0;
  int bias4_off;
  int gs;                // 16-channel groups per weight stage (streamed)
  int p0_last;           // the last MMAs of this stage are the last tensor-core reads of the patch
  uint32_t idesc;
  uint32_t w_smem_off;   // resident: byte offset inside the weight area
  const __nv_bfloat16* w_gmem;
};
struct ChainArgs {
  int nstages, nbufs;
  ChainStageD st[CHAIN_MAX_STAGES];
  ChainBufD buf[MAX_BUFS];
  int h, w, rw, tw, th, tiles_y, tiles_per_img, batch;
  float inv_rw;
  int m0, p0_nbuf;
  uint32_t p0_stride, p0_tx_bytes;
  int p0_res_readers;    // an epilogue reads its residual from the patch: the patch is free only after those reads
  int resident;
  uint32_t w_off, ring_off, ring_slot_bytes;
  int ring_n;
  int ns, nslot, chunks_per_tile;
  uint32_t stage_off;
  int npieces, out_es;
  int piece_start[MAX_PIECES], piece_bytes[MAX_PIECES], piece_map[MAX_PIECES], piece_c0[MAX_PIECES];
  uint32_t piece_off[MAX_PIECES];
  float4 bias4[128];
};
struct ChainMaps {
  CUtensorMap in;
  CUtensorMap out[3];  // 128- / 64- / 32-byte channel pieces of an output row
};

__device__ __forceinline__ float silu_fast(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}
__device__ __forceinline__ float activate(float x, int act) {
  if (act == 1) return silu_fast(x);
  if (act == 2) return fmaxf(x, 0.0f);
  return x;
}
// (see conv_win.cu: the whole warp executes, the elected lane performs)
__device__ __forceinline__ void mma_issue(bool leader, uint32_t tmem_d, uint32_t a_lo, uint32_t a_hi, uint32_t b_lo,
                                          uint32_t b_hi, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t.reg .b64 da, db;\n\t"
      "setp.ne.b32 p, %6, 0;\n\t"
      "setp.ne.b32 q, %7, 0;\n\t"
      "mov.b64 da, {%1, %2};\n\t"
      "mov.b64 db, {%3, %4};\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %5, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(a_hi), "r"(b_lo), "r"(b_hi), "r"(idesc), "r"(accumulate), "r"(static_cast<uint32_t>(leader))
      : "memory");
}
__device__ __forceinline__ void tc_commit_if(bool leader, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "setp.ne.b32 q, %1, 0;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar),
      "r"(static_cast<uint32_t>(leader))
      : "memory");
}
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

struct TileXY {
  int n, x0, y0;
};
__device__ __forceinline__ TileXY tile_xy(const ChainArgs& a, int tile) {
  TileXY t;
  t.n = tile / a.tiles_per_img;
  const int r = tile - t.n * a.tiles_per_img;
  const int strip = r / a.tiles_y;
  t.x0 = strip * a.tw;
  t.y0 = (r - strip * a.tiles_y) * a.th;
  return t;
}

__global__ void __launch_bounds__(CH_THREADS, 1) conv_chain_kernel(const __grid_constant__ ChainArgs a, const __grid_constant__ ChainMaps maps) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + OFF_TMEM_PTR);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_tiles = a.batch * a.tiles_per_img;
  pdl_trigger();
  if (static_cast<int>(blockIdx.x) >= total_tiles) return;

  if (threadIdx.x == 0) {
    for (int s = 0; s < 2; ++s) {
      mbar_init(sbase + BAR_P0_FULL + 8 * s, 1);
      mbar_init(sbase + BAR_P0_EMPTY + 8 * s, 1 + (a.p0_res_readers ? 128 * NWG : 0));
    }
    for (int s = 0; s < MAX_RING; ++s) {
      mbar_init(sbase + BAR_B_FULL + 8 * s, 1);
      mbar_init(sbase + BAR_B_EMPTY + 8 * s, 1);
    }
    mbar_init(sbase + BAR_W_FULL, 1);
    mbar_init(sbase + BAR_TILE_DONE, 128 * NWG);
    mbar_init(sbase + BAR_STAGE_FREE, 1);
    for (int s = 0; s < MAX_SLOTS; ++s) {
      mbar_init(sbase + BAR_ACC_FULL + 8 * s, 1);
      mbar_init(sbase + BAR_ACC_EMPTY + 8 * s, 128);
    }
    for (int s = 0; s < MAX_BUFS * MAX_CHUNKS; ++s) mbar_init(sbase + BAR_READY + 8 * s, 128);
    mbar_init_fence();
  }
  if (warp == 4 * NWG) tc_alloc(smem_u32(tmem_ptr_smem), 512);
  if (warp == 4 * NWG + 1 && lane == 0) tma_prefetch_desc(&maps.in);
  if (warp == 4 * NWG + 3 && lane == 0) tma_prefetch_desc(&maps.out[0]);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp < 4 * NWG) {
    // ================================================================== epilogue (4 warpgroups, whole chunks each)
    const int wg = warp >> 2, wq = warp & 3;
    const int rloc = wq * 32 + lane;
    const uint32_t tlane = tmem_base + (static_cast<uint32_t>(wq * 32) << 16);
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const TileXY t = tile_xy(a, tile);
      const int p0buf = a.p0_nbuf == 2 ? (it & 1) : 0;
      // the store of the previous tile has finished reading the staging tile (which may alias a buffer written below)
      mbar_wait(sbase + BAR_STAGE_FREE, (it & 1) ^ 1);
      // residual rows are read from the patch with ordinary loads: observe the TMA's completion barrier myself
      if (a.p0_res_readers) mbar_wait(sbase + BAR_P0_FULL + 8 * p0buf, a.p0_nbuf == 2 ? ((it >> 1) & 1) : (it & 1));
      for (int s = 0; s < a.nstages; ++s) {
        const ChainStageD& S = a.st[s];
        const bool last = S.dst_buf < 0;
        const int ngroups = S.n_pad >> 4;
        for (int k = 0; k < S.nchunks; ++k) {
          const int gidx = it * a.chunks_per_tile + S.chunk0 + k;
          if ((gidx & (NWG - 1)) != wg) continue;  // another warpgroup's chunk
          const int slot = gidx % a.nslot;
          const uint32_t ph = static_cast<uint32_t>(gidx / a.nslot) & 1u;
          mbar_wait(sbase + BAR_ACC_FULL + 8 * slot, ph);
          tc_fence_after();
          const int q = k * 128 + rloc;
          const int yl = __float2int_rd((static_cast<float>(q) + 0.5f) * a.inv_rw);
          const int xl = q - yl * a.rw;
          const int py = t.y0 - S.margin + yl, px = t.x0 - S.margin + xl;
          const bool inside = py >= 0 && py < a.h && px >= 0 && px < a.w;
          // destination row in the next stage's buffer
          uint32_t d_row = 0, d_xor = 0, d_stride = 0;
          int d_shift = 0;
          if (!last) {
            const ChainBufD& D = a.buf[S.dst_buf];
            const uint32_t ro = static_cast<uint32_t>(q) * D.row_bytes;
            d_row = D.off + ro;
            d_xor = ((ro >> 7) & D.xor_mask) << 4;
            d_stride = D.slab_stride;
            d_shift = D.c16_shift;
          }
          // residual row
          uint32_t r_row = 0, r_xor = 0, r_stride = 0;
          int r_shift = 0;
          if (S.res_mode) {
            const ChainBufD& R = a.buf[S.res_buf];
            const uint32_t ro = static_cast<uint32_t>(q + S.res_shift) * R.row_bytes;
            r_row = R.off + (S.res_buf == 0 ? p0buf * a.p0_stride : 0u) + ro;
            r_xor = ((ro >> 7) & R.xor_mask) << 4;
            r_stride = R.slab_stride;
            r_shift = R.c16_shift;
          }
          const bool store_out = last && yl < a.th && xl < a.tw && inside;
          const uint32_t crow = static_cast<uint32_t>(yl * a.tw + xl);
          const uint32_t taddr = tlane + static_cast<uint32_t>(slot * a.ns);
          for (int g = 0; g < ngroups; g += 2) {
            uint32_t v[2][16];
            const bool two = g + 1 < ngroups;  // warp-uniform
            __syncwarp();
            tc_ld16_nowait(taddr + g * 16, v[0]);
            if (two) tc_ld16_nowait(taddr + g * 16 + 16, v[1]);
            tc_ld_wait();
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              if (hh == 1 && !two) break;
              const int cg = g + hh;
              float x[16];
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float4 b4 = a.bias4[S.bias4_off + cg * 4 + i];
                x[4 * i] = __uint_as_float(v[hh][4 * i]) + b4.x;
                x[4 * i + 1] = __uint_as_float(v[hh][4 * i + 1]) + b4.y;
                x[4 * i + 2] = __uint_as_float(v[hh][4 * i + 2]) + b4.z;
                x[4 * i + 3] = __uint_as_float(v[hh][4 * i + 3]) + b4.w;
              }
              if (S.res_mode) {
                const int rc = S.res_c16_0 + cg;
                const uint32_t base = r_row + static_cast<uint32_t>(rc >> r_shift) * r_stride;
                const uint32_t o = static_cast<uint32_t>(rc & ((1 << r_shift) - 1)) * 32;
                const uint4 q0 = *reinterpret_cast<const uint4*>(smem + base + (o ^ r_xor));
                const uint4 q1 = *reinterpret_cast<const uint4*>(smem + base + ((o + 16) ^ r_xor));
                const uint32_t rw_[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
                if (S.res_mode == 2) {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    x[2 * i] = activate(x[2 * i] + bf16_lo(rw_[i]), S.act);
                    x[2 * i + 1] = activate(x[2 * i + 1] + bf16_hi(rw_[i]), S.act);
                  }
                } else {
#pragma unroll
                  for (int i = 0; i < 8; ++i) {
                    x[2 * i] = activate(x[2 * i], S.act) + bf16_lo(rw_[i]);
                    x[2 * i + 1] = activate(x[2 * i + 1], S.act) + bf16_hi(rw_[i]);
                  }
                }
              } else {
#pragma unroll
                for (int i = 0; i < 16; ++i) x[i] = activate(x[i], S.act);
              }
              if (!last) {
                uint4 o0 = make_uint4(0u, 0u, 0u, 0u), o1 = o0;
                if (inside) {  // positions outside the image are the next layer's zero padding
                  o0.x = pack_bf16x2(x[0], x[1]);   o0.y = pack_bf16x2(x[2], x[3]);
                  o0.z = pack_bf16x2(x[4], x[5]);   o0.w = pack_bf16x2(x[6], x[7]);
                  o1.x = pack_bf16x2(x[8], x[9]);   o1.y = pack_bf16x2(x[10], x[11]);
                  o1.z = pack_bf16x2(x[12], x[13]); o1.w = pack_bf16x2(x[14], x[15]);
                }
                const uint32_t base = d_row + static_cast<uint32_t>(cg >> d_shift) * d_stride;
                const uint32_t o = static_cast<uint32_t>(cg & ((1 << d_shift) - 1)) * 32;
                *reinterpret_cast<uint4*>(smem + base + (o ^ d_xor)) = o0;
                *reinterpret_cast<uint4*>(smem + base + ((o + 16) ^ d_xor)) = o1;
              } else if (store_out) {
                // byte offset of this 16-channel group inside an output row -> its channel piece
                const int ob = cg * 16 * a.out_es;
                int p = 0;
#pragma unroll
                for (int pp = 1; pp < MAX_PIECES; ++pp)
                  if (pp < a.npieces && a.piece_start[pp] <= ob) p = pp;
                const uint32_t pb = static_cast<uint32_t>(a.piece_bytes[p]);
                const uint32_t ro = crow * pb;
                const uint32_t sx = ((ro >> 7) & (pb == 128 ? 7u : (pb == 64 ? 3u : 1u))) << 4;
                const uint32_t base = a.stage_off + a.piece_off[p] + ro;
                const uint32_t oi = static_cast<uint32_t>(ob - a.piece_start[p]);
                if (a.out_es == 4) {
#pragma unroll
                  for (int i = 0; i < 4; ++i)
                    *reinterpret_cast<float4*>(smem + base + ((oi + 16 * i) ^ sx)) = make_float4(x[4 * i], x[4 * i + 1], x[4 * i + 2], x[4 * i + 3]);
                } else {
                  uint4 o0, o1;
                  o0.x = pack_bf16x2(x[0], x[1]);   o0.y = pack_bf16x2(x[2], x[3]);
                  o0.z = pack_bf16x2(x[4], x[5]);   o0.w = pack_bf16x2(x[6], x[7]);
                  o1.x = pack_bf16x2(x[8], x[9]);   o1.y = pack_bf16x2(x[10], x[11]);
                  o1.z = pack_bf16x2(x[12], x[13]); o1.w = pack_bf16x2(x[14], x[15]);
                  *reinterpret_cast<uint4*>(smem + base + (oi ^ sx)) = o0;
                  *reinterpret_cast<uint4*>(smem + base + ((oi + 16) ^ sx)) = o1;
                }
              }
            }
          }
          // the accumulator slot has been read: the MMA warp may reuse it
          tc_fence_before();
          mbar_arrive(sbase + BAR_ACC_EMPTY + 8 * slot);
          if (!last) {  // my row of the next stage's operand is written: publish it to the tensor core (async proxy)
            fence_proxy_async();
            mbar_arrive(sbase + BAR_READY + 8 * (S.dst_buf * MAX_CHUNKS + k));
          }
        }
      }
      fence_proxy_async();  // staging writes -> visible to the TMA unit
      mbar_arrive(sbase + BAR_TILE_DONE);
      if (a.p0_res_readers) mbar_arrive(sbase + BAR_P0_EMPTY + 8 * p0buf);
    }
  } else if (warp == 4 * NWG) {
    // ================================================================== MMA issuer (whole warp convergent, one lane issues)
    const bool leader = elect_one();
    const uint32_t tmem0 = __shfl_sync(0xffffffffu, tmem_base, 0);
    const uint32_t b_hi = (128u >> 4) | (1u << 14);  // SBO = 128 B, no swizzle
    uint32_t rs = 0, rp = 0;                         // weight ring slot / parity
    if (a.resident) mbar_wait(sbase + BAR_W_FULL, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int p0buf = a.p0_nbuf == 2 ? (it & 1) : 0;
      const uint32_t p0par = a.p0_nbuf == 2 ? ((it >> 1) & 1) : (it & 1);
      mbar_wait(sbase + BAR_P0_FULL + 8 * p0buf, p0par);
      tc_fence_after();
      int waited[MAX_BUFS];
#pragma unroll
      for (int b = 0; b < MAX_BUFS; ++b) waited[b] = 0;
      for (int s = 0; s < a.nstages; ++s) {
        const ChainStageD& S = a.st[s];
        const uint32_t lbo = static_cast<uint32_t>(S.n_pad);
        for (int g0 = 0; g0 < S.nchunks; g0 += S.group) {
          const int g1 = min(g0 + S.group, S.nchunks);
          const int gbase = it * a.chunks_per_tile + S.chunk0;
          for (int k = g0; k < g1; ++k) {  // the accumulator slots of the group have been drained
            const int gi = gbase + k;
            mbar_wait(sbase + BAR_ACC_EMPTY + 8 * (gi % a.nslot), (static_cast<uint32_t>(gi / a.nslot) & 1u) ^ 1u);
          }
          // every row this group reads from a stage-written buffer has been published
          for (int j = 0; j < S.nsrc; ++j) {
            const int b = S.src[j].buf;
            if (b == 0) continue;
            const int reach = g1 * 128 + S.src[j].shift + (S.ksize - 1) * (a.rw + 1);
            const int need = min(a.buf[b].nchunks, (reach + 127) >> 7);
#pragma unroll
            for (int bb = 1; bb < MAX_BUFS; ++bb)
              if (bb == b)
                while (waited[bb] < need) {
                  mbar_wait(sbase + BAR_READY + 8 * (bb * MAX_CHUNKS + waited[bb]), it & 1);
                  ++waited[bb];
                }
          }
          tc_fence_after();
          for (int tap = 0; tap < S.taps; ++tap) {
            const int dy = tap / S.ksize, dx = tap - dy * S.ksize;
            for (int j = 0; j < S.nsrc; ++j) {
              const ChainSrcD& src = S.src[j];
              const ChainBufD& B = a.buf[src.buf];
              const uint32_t bufbase = sbase + B.off + (src.buf == 0 ? p0buf * a.p0_stride : 0u);
              const int row_tap = src.shift + dy * a.rw + dx;
              for (int sub0 = 0; sub0 < src.nc16; sub0 += S.gs) {
                const int sub1 = min(sub0 + S.gs, src.nc16);
                uint32_t b_base;
                if (a.resident) {
                  b_base = sbase + a.w_off + S.w_smem_off + static_cast<uint32_t>(tap * S.k8_per_tap + src.wk8 + sub0 * 2) * lbo * 16;
                } else {
                  mbar_wait(sbase + BAR_B_FULL + 8 * rs, rp);
                  tc_fence_after();
                  b_base = sbase + a.ring_off + rs * a.ring_slot_bytes;
                }
                const uint32_t b_lo0 = (b_base >> 4) | (lbo << 16);
                const bool first_k = tap == 0 && j == 0 && sub0 == 0;
                for (int k = g0; k < g1; ++k) {
                  const uint32_t d_tmem = tmem0 + static_cast<uint32_t>(((gbase + k) % a.nslot) * a.ns);
                  const uint32_t row = static_cast<uint32_t>(k * 128 + row_tap);
                  for (int c = sub0; c < sub1; ++c) {
                    const int c16 = src.c16_0 + c;
                    const uint32_t a_addr = bufbase + static_cast<uint32_t>(c16 >> B.c16_shift) * B.slab_stride + row * B.row_bytes;
                    const uint32_t a_lo = (a_addr >> 4) + static_cast<uint32_t>(c16 & ((1 << B.c16_shift) - 1)) * 2 + (1u << 16);
                    mma_issue(leader, d_tmem, a_lo, B.a_hi, b_lo0 + static_cast<uint32_t>(c - sub0) * 2 * lbo, b_hi, S.idesc,
                              (first_k && c == sub0) ? 0u : 1u);
                  }
                }
                if (!a.resident) {
                  tc_commit_if(leader, sbase + BAR_B_EMPTY + 8 * rs);
                  if (++rs == static_cast<uint32_t>(a.ring_n)) { rs = 0; rp ^= 1; }
                }
              }
            }
          }
          for (int k = g0; k < g1; ++k) tc_commit_if(leader, sbase + BAR_ACC_FULL + 8 * ((gbase + k) % a.nslot));
        }
        if (S.p0_last) tc_commit_if(leader, sbase + BAR_P0_EMPTY + 8 * p0buf);
      }
    }
    tc_fence_before();
  } else if (warp == 4 * NWG + 1) {
    // ================================================================== patch producer
    if (lane == 0) {
      pdl_wait();  // the patch is the previous layer's output
      const ChainBufD& B = a.buf[0];
      int it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
        const TileXY t = tile_xy(a, tile);
        const int p0buf = a.p0_nbuf == 2 ? (it & 1) : 0;
        const uint32_t par = a.p0_nbuf == 2 ? ((it >> 1) & 1) : (it & 1);
        mbar_wait(sbase + BAR_P0_EMPTY + 8 * p0buf, par ^ 1);
        const uint32_t bar = sbase + BAR_P0_FULL + 8 * p0buf;
        mbar_arrive_expect_tx(bar, a.p0_tx_bytes);
        const int slab_ch = static_cast<int>(B.row_bytes >> 1);
        for (int sl = 0; sl < B.nslabs; ++sl)
          tma_load_4d(sbase + B.off + p0buf * a.p0_stride + sl * B.slab_stride, &maps.in, bar, sl * slab_ch, t.x0 - a.m0, t.y0 - a.m0, t.n);
      }
    }
  } else if (warp == 4 * NWG + 2) {
    // ================================================================== weight producer
    if (lane == 0) {
      if (a.resident) {
        uint32_t total = 0;
        for (int s = 0; s < a.nstages; ++s) total += static_cast<uint32_t>(a.st[s].taps * a.st[s].k8_per_tap * a.st[s].n_pad * 16);
        mbar_arrive_expect_tx(sbase + BAR_W_FULL, total);
        for (int s = 0; s < a.nstages; ++s) {
          const ChainStageD& S = a.st[s];
          const uint32_t bytes = static_cast<uint32_t>(S.taps * S.k8_per_tap * S.n_pad * 16);
          const uint8_t* srcp = reinterpret_cast<const uint8_t*>(S.w_gmem);
          for (uint32_t off = 0; off < bytes; off += 16384)
            bulk_g2s(sbase + a.w_off + S.w_smem_off + off, srcp + off, min(16384u, bytes - off), sbase + BAR_W_FULL);
        }
      } else {
        uint32_t rs = 0, rp = 1;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          for (int s = 0; s < a.nstages; ++s) {
            const ChainStageD& S = a.st[s];
            for (int g0 = 0; g0 < S.nchunks; g0 += S.group)
              for (int tap = 0; tap < S.taps; ++tap)
                for (int j = 0; j < S.nsrc; ++j)
                  for (int sub0 = 0; sub0 < S.src[j].nc16; sub0 += S.gs) {
                    const int sub1 = min(sub0 + S.gs, S.src[j].nc16);
                    const uint32_t bytes = static_cast<uint32_t>((sub1 - sub0) * 2 * S.n_pad * 16);
                    mbar_wait(sbase + BAR_B_EMPTY + 8 * rs, rp);
                    const uint32_t bar = sbase + BAR_B_FULL + 8 * rs;
                    mbar_arrive_expect_tx(bar, bytes);
                    const long long k8 = static_cast<long long>(tap) * S.k8_per_tap + S.src[j].wk8 + sub0 * 2;
                    bulk_g2s(sbase + a.ring_off + rs * a.ring_slot_bytes, S.w_gmem + k8 * S.n_pad * 8, bytes, bar);
                    if (++rs == static_cast<uint32_t>(a.ring_n)) { rs = 0; rp ^= 1; }
                  }
          }
        }
      }
    }
  } else {
    // ================================================================== store warp
    pdl_wait();  // output stores: only after the previous layer has completed
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const TileXY t = tile_xy(a, tile);
      mbar_wait(sbase + BAR_TILE_DONE, it & 1);
      if (lane == 0) {
        for (int p = 0; p < a.npieces; ++p)
          tma_store_4d(&maps.out[a.piece_map[p]], sbase + a.stage_off + a.piece_off[p], a.piece_c0[p], t.x0, t.y0, t.n);
        bulk_commit();
        bulk_wait_read_all();
        mbar_arrive(sbase + BAR_STAGE_FREE);
      }
      __syncwarp();
    }
    if (lane == 0) bulk_wait_all();
  }
  __syncthreads();
  if (warp == 4 * NWG) {
    tc_fence_after();
    tc_dealloc(tmem_base, 512);
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

inline uint32_t up1024(size_t v) { return static_cast<uint32_t>((v + 1023) / 1024 * 1024); }
inline int slab_of(int c) { return c % 64 == 0 ? 64 : (c % 32 == 0 ? 32 : 16); }

struct Plan {
  bool ok = false;
  double cost = 0.0;
  int tw = 0, th = 0, strips = 0, rw = 0;
  int nchunks[CHAIN_MAX_STAGES] = {0};
  uint32_t rows[MAX_BUFS] = {0};
  int p0_nbuf = 1;
  int alias_buf = -1;
  size_t smem = 0;
};

}  // namespace

int try_launch_conv_chain(const ChainSpec& sp, cudaStream_t stream) {
  static const bool disabled = getenv("AICAM_NO_CHAIN") != nullptr;
  static const bool debug = getenv("AICAM_CHAIN_DEBUG") != nullptr;
  if (disabled || encode_tiled() == nullptr) return 0;
  const int ns_ = sp.nstages;
  if (ns_ < 2 || ns_ > CHAIN_MAX_STAGES || sp.batch <= 0) return 0;
  if (sp.in_c % 16 != 0 || sp.in_cstride % 8 != 0 || sp.in_coff % 8 != 0 || reinterpret_cast<uintptr_t>(sp.in) % 16 != 0) return 0;
  if (sp.in_img_stride != static_cast<long long>(sp.h) * sp.w * sp.in_cstride) return 0;
  const int es = sp.out_f32 ? 4 : 2;
  if ((static_cast<long long>(sp.out_cstride) * es) % 16 != 0 || (static_cast<long long>(sp.out_coff) * es) % 16 != 0 ||
      (sp.out_img_stride * es) % 16 != 0 || reinterpret_cast<uintptr_t>(sp.out) % 16 != 0)
    return 0;

  // ---- buffers: channels, margins, consumers
  int buf_c[MAX_BUFS] = {0}, buf_margin[MAX_BUFS], st_margin[CHAIN_MAX_STAGES] = {0};
  buf_c[0] = sp.in_c;
  size_t wbytes_total = 0;
  int bias_ch = 0;
  for (int s = 0; s < ns_; ++s) {
    const ChainStageSpec& S = sp.st[s];
    if (!S.pc || S.pc->stride != 1 || (S.pc->ksize != 1 && S.pc->ksize != 3) || S.pc->s2d_c0) return 0;
    if (S.pc->cout % 16 != 0 || S.pc->cout > 256 || S.pc->cin % 16 != 0 || S.pc->cin_pad != S.pc->cin || !S.pc->bias_host) return 0;
    if (S.nsrc < 1 || S.nsrc > 2) return 0;
    int csum = 0;
    for (int j = 0; j < S.nsrc; ++j) {
      if (S.src_buf[j] < 0 || S.src_buf[j] > s || S.src_coff[j] % 16 != 0 || S.src_c[j] % 16 != 0 || S.src_c[j] <= 0) return 0;
      csum += S.src_c[j];
    }
    if (csum != S.pc->cin) return 0;
    if (S.res_mode && (S.res_buf < 0 || S.res_buf > s || S.res_coff % 16 != 0)) return 0;
    if (s + 1 < ns_) buf_c[s + 1] = S.pc->cout;
    wbytes_total += static_cast<size_t>(S.pc->ksize) * S.pc->ksize * S.pc->cin * S.pc->cout * 2;
    bias_ch += S.pc->cout;
  }
  if (bias_ch > 512) return 0;
  for (int s = 0; s < ns_; ++s)  // channel ranges must lie inside their buffers
    for (int j = 0; j < sp.st[s].nsrc; ++j)
      if (sp.st[s].src_coff[j] + sp.st[s].src_c[j] > buf_c[sp.st[s].src_buf[j]]) return 0;
  for (int s = 0; s < ns_; ++s)
    if (sp.st[s].res_mode && sp.st[s].res_coff + sp.st[s].pc->cout > buf_c[sp.st[s].res_buf]) return 0;
  for (int b = 0; b < MAX_BUFS; ++b) buf_margin[b] = -1;
  for (int s = ns_ - 1; s >= 0; --s) {
    const ChainStageSpec& S = sp.st[s];
    st_margin[s] = s == ns_ - 1 ? 0 : buf_margin[s + 1];
    if (st_margin[s] < 0) return 0;  // a stage nobody reads
    const int p = S.pc->ksize / 2;
    for (int j = 0; j < S.nsrc; ++j) buf_margin[S.src_buf[j]] = std::max(buf_margin[S.src_buf[j]], st_margin[s] + p);
    if (S.res_mode) buf_margin[S.res_buf] = std::max(buf_margin[S.res_buf], st_margin[s]);
  }
  const int m0 = buf_margin[0];
  if (m0 < 0) return 0;
  int ns_cols = 16;
  for (int s = 0; s < ns_; ++s) ns_cols = std::max(ns_cols, sp.st[s].pc->cout);
  const int nslot = std::min(MAX_SLOTS, 512 / ns_cols);
  if (nslot < 2) return 0;
  const bool resident = wbytes_total <= RESIDENT_LIMIT;
  // weight stages of the streamed form: 16-channel groups per stage, per stage of the chain
  int gs[CHAIN_MAX_STAGES];
  uint32_t ring_slot = 0;
  for (int s = 0; s < ns_; ++s) {
    const ChainStageSpec& S = sp.st[s];
    int maxc16 = 0;
    for (int j = 0; j < S.nsrc; ++j) maxc16 = std::max(maxc16, S.src_c[j] / 16);
    gs[s] = resident ? maxc16 : std::max(1, std::min(maxc16, static_cast<int>(RING_SLOT_MAX / (32 * S.pc->cout))));
    if (!resident) ring_slot = std::max(ring_slot, static_cast<uint32_t>(gs[s]) * 32u * S.pc->cout);
  }
  const int ring_n = resident ? 0 : 3;
  const size_t wt_bytes = resident ? up1024(wbytes_total) : static_cast<size_t>(ring_n) * up1024(ring_slot);
  const int last_c = sp.st[ns_ - 1].pc->cout;
  const int out_row_bytes = last_c * es;
  // which buffer may double as the staging tile: one whose readers are all done when the last stage's first epilogue starts
  bool alias_ok[MAX_BUFS];
  for (int b = 0; b < ns_; ++b) {
    alias_ok[b] = b >= 1;
    for (int s = 0; s < ns_; ++s) {
      const ChainStageSpec& S = sp.st[s];
      for (int j = 0; j < S.nsrc; ++j)
        if (S.src_buf[j] == b && s == ns_ - 1) alias_ok[b] = false;
      if (S.res_mode && S.res_buf == b && s >= ns_ - 2) alias_ok[b] = false;
    }
  }

  // ---- choose the tile: strips x rows, cheapest estimated time
  const int num_sms = current_num_sms();
  Plan best;
  for (int strips = 1; strips <= 8; ++strips) {
    const int tw = (sp.w + strips - 1) / strips;
    if (strips > 1 && tw < 8) break;
    const int rw = tw + 2 * m0;
    if (rw > 256) continue;
    for (int th = 1; th <= std::min(sp.h, 64); ++th) {
      Plan p;
      p.tw = tw; p.th = th; p.strips = strips; p.rw = rw;
      const int bh0 = th + 2 * m0;
      if (bh0 > 256) break;
      bool fits = true;
      double mma = 0.0, epi = 0.0;
      int chunks_total = 0;
      for (int s = 0; s < ns_ && fits; ++s) {
        const int m = st_margin[s];
        const int rows_needed = (th + 2 * m - 1) * rw + tw + 2 * m;
        p.nchunks[s] = (rows_needed + 127) / 128;
        if (p.nchunks[s] > MAX_CHUNKS) fits = false;
        if (!resident && p.nchunks[s] > nslot) fits = false;  // a streamed stage keeps all its accumulators live
        const PackedConv* pc = sp.st[s].pc;
        const int n = pc->cout;
        mma += static_cast<double>(p.nchunks[s]) * (pc->ksize * pc->ksize * pc->cin / 16) * std::max(n / 2.0, 32.0 + n / 4.0);
        epi += static_cast<double>(p.nchunks[s]) * (n * 35.0 + 600.0) / NWG;
        chunks_total += p.nchunks[s];
      }
      if (!fits) break;
      // rows every buffer must hold: what its producer writes and the furthest row a consumer's descriptors reach
      for (int b = 0; b < ns_; ++b) p.rows[b] = b == 0 ? static_cast<uint32_t>(bh0 * rw) : static_cast<uint32_t>(p.nchunks[b - 1] * 128);
      for (int s = 0; s < ns_; ++s) {
        const ChainStageSpec& S = sp.st[s];
        const int pp = S.pc->ksize / 2;
        for (int j = 0; j < S.nsrc; ++j) {
          const int b = S.src_buf[j];
          const int shift = (buf_margin[b] - st_margin[s] - pp) * (rw + 1);
          p.rows[b] = std::max(p.rows[b], static_cast<uint32_t>(p.nchunks[s] * 128 + shift + 2 * pp * (rw + 1) + 1));
        }
        if (S.res_mode) {
          const int b = S.res_buf;
          p.rows[b] = std::max(p.rows[b], static_cast<uint32_t>(p.nchunks[s] * 128 + (buf_margin[b] - st_margin[s]) * (rw + 1) + 1));
        }
      }
      size_t buf_bytes[MAX_BUFS] = {0};
      for (int b = 0; b < ns_; ++b) {
        const int sl = slab_of(buf_c[b]);
        buf_bytes[b] = static_cast<size_t>(buf_c[b] / sl) * up1024(static_cast<size_t>(p.rows[b]) * sl * 2);
      }
      // staging: compact [th * tw rows] per channel piece
      size_t stage_bytes = 0;
      {
        int rem = out_row_bytes;
        while (rem > 0) {
          const int pb = rem >= 128 ? 128 : (rem >= 64 ? 64 : 32);
          stage_bytes += up1024(static_cast<size_t>(th) * tw * pb);
          rem -= pb;
        }
      }
      p.alias_buf = -1;
      for (int b = 1; b < ns_; ++b)
        if (alias_ok[b] && buf_bytes[b] >= stage_bytes && (p.alias_buf < 0 || buf_bytes[b] > buf_bytes[p.alias_buf])) p.alias_buf = b;
      size_t fixed = OFF_DATA + wt_bytes + (p.alias_buf >= 0 ? 0 : stage_bytes);
      for (int b = 1; b < ns_; ++b) fixed += buf_bytes[b];
      if (fixed + buf_bytes[0] > SMEM_LIMIT) break;  // (taller tiles only need more)
      p.p0_nbuf = fixed + 2 * buf_bytes[0] <= SMEM_LIMIT ? 2 : 1;
      p.smem = fixed + p.p0_nbuf * buf_bytes[0];
      const int tiles_y = (sp.h + th - 1) / th;
      const long long tiles = static_cast<long long>(sp.batch) * strips * tiles_y;
      // TMA moves one pixel row (32 / 64 / 128 bytes) per request, ~3.4 cycles each (measured on the 16-channel layers)
      const int sl0 = slab_of(buf_c[0]);
      int out_pieces = 0;
      for (int rem = out_row_bytes; rem > 0; rem -= (rem >= 128 ? 128 : (rem >= 64 ? 64 : 32))) ++out_pieces;
      const double tma = (static_cast<double>(bh0) * rw * (buf_c[0] / sl0) + static_cast<double>(th) * tw * out_pieces) * 3.4;
      double per_tile = std::max(mma, std::max(epi, tma)) + 0.25 * (mma + epi) + 1500.0;
      if (p.p0_nbuf == 1) per_tile += 0.5 * tma + 1500.0;  // the patch load is only partly hidden
      const long long rounds = (tiles + num_sms - 1) / num_sms;
      p.cost = static_cast<double>(rounds) * per_tile;
      p.ok = true;
      (void)chunks_total;
      if (!best.ok || p.cost < best.cost) best = p;
    }
  }
  if (!best.ok) return 0;

  // ---- kernel arguments
  ChainArgs a;
  std::memset(&a, 0, sizeof(a));
  a.nstages = ns_; a.nbufs = ns_;
  a.h = sp.h; a.w = sp.w; a.rw = best.rw; a.tw = best.tw; a.th = best.th;
  a.tiles_y = (sp.h + best.th - 1) / best.th; a.tiles_per_img = best.strips * a.tiles_y; a.batch = sp.batch;
  a.inv_rw = 1.0f / static_cast<float>(best.rw);
  a.m0 = m0; a.p0_nbuf = best.p0_nbuf;
  a.resident = resident ? 1 : 0;
  a.ns = ns_cols; a.nslot = nslot;
  a.out_es = es;
  uint32_t off = OFF_DATA;
  for (int b = 0; b < ns_; ++b) {
    ChainBufD& B = a.buf[b];
    const int sl = slab_of(buf_c[b]);
    B.row_bytes = sl * 2;
    B.xor_mask = sl == 64 ? 7u : (sl == 32 ? 3u : 1u);
    const uint32_t ltype = sl == 64 ? 2u : (sl == 32 ? 4u : 6u);
    B.a_hi = ((8 * B.row_bytes) >> 4) | (1u << 14) | (ltype << 29);
    B.c16_shift = sl == 64 ? 2 : (sl == 32 ? 1 : 0);
    B.nslabs = buf_c[b] / sl;
    B.nchunks = b == 0 ? 0 : best.nchunks[b - 1];
    B.slab_stride = up1024(static_cast<size_t>(best.rows[b]) * sl * 2);
    B.off = off;
    const uint32_t bytes = B.nslabs * B.slab_stride;
    if (b == 0) {
      a.p0_stride = bytes;
      a.p0_tx_bytes = static_cast<uint32_t>(B.nslabs) * (best.th + 2 * m0) * best.rw * B.row_bytes;
      off += bytes * best.p0_nbuf;
    } else {
      off += bytes;
    }
  }
  if (resident) {
    a.w_off = off;
    uint32_t wo = 0;
    for (int s = 0; s < ns_; ++s) {
      a.st[s].w_smem_off = wo;
      wo += static_cast<uint32_t>(sp.st[s].pc->ksize * sp.st[s].pc->ksize * sp.st[s].pc->cin * sp.st[s].pc->cout * 2);
    }
    off += up1024(wbytes_total);
  } else {
    a.ring_off = off;
    a.ring_slot_bytes = up1024(ring_slot);
    a.ring_n = ring_n;
    off += ring_n * a.ring_slot_bytes;
  }
  a.stage_off = best.alias_buf >= 0 ? a.buf[best.alias_buf].off : off;
  {
    int rem = out_row_bytes, start = 0;
    uint32_t poff = 0;
    a.npieces = 0;
    while (rem > 0) {
      if (a.npieces >= MAX_PIECES) return 0;
      const int pb = rem >= 128 ? 128 : (rem >= 64 ? 64 : 32);
      const int p = a.npieces++;
      a.piece_start[p] = start; a.piece_bytes[p] = pb; a.piece_off[p] = poff;
      a.piece_map[p] = pb == 128 ? 0 : (pb == 64 ? 1 : 2);
      a.piece_c0[p] = start / es;
      poff += up1024(static_cast<size_t>(best.th) * best.tw * pb);
      start += pb; rem -= pb;
    }
    if (best.alias_buf < 0) off += poff;
  }
  const size_t smem = off;
  if (smem > SMEM_LIMIT) return fail(AICAM_ERR_INVALID_ARG, "conv_chain: internal error, shared-memory plan exceeds the limit");
  int chunk0 = 0, bias_off = 0, last_p0_stage = -1;
  for (int s = 0; s < ns_; ++s) {
    const ChainStageSpec& S = sp.st[s];
    ChainStageD& D = a.st[s];
    D.ksize = S.pc->ksize; D.taps = D.ksize * D.ksize; D.nsrc = S.nsrc;
    const int pp = D.ksize / 2;
    int wk8 = 0;
    for (int j = 0; j < S.nsrc; ++j) {
      D.src[j].buf = S.src_buf[j];
      D.src[j].c16_0 = S.src_coff[j] / 16;
      D.src[j].nc16 = S.src_c[j] / 16;
      D.src[j].shift = (buf_margin[S.src_buf[j]] - st_margin[s] - pp) * (best.rw + 1);
      D.src[j].wk8 = wk8;
      wk8 += S.src_c[j] / 8;
      if (S.src_buf[j] == 0) last_p0_stage = s;
    }
    D.k8_per_tap = S.pc->cin / 8;
    D.n_pad = S.pc->cout;
    D.act = S.act;
    D.res_mode = S.res_mode;
    D.res_buf = S.res_mode ? S.res_buf : 0;
    D.res_c16_0 = S.res_coff / 16;
    D.res_shift = S.res_mode ? (buf_margin[S.res_buf] - st_margin[s]) * (best.rw + 1) : 0;
    if (S.res_mode && S.res_buf == 0) a.p0_res_readers = 1;
    D.dst_buf = s + 1 < ns_ ? s + 1 : -1;
    D.margin = st_margin[s];
    D.nchunks = best.nchunks[s];
    D.group = resident ? 1 : D.nchunks;
    D.chunk0 = chunk0;
    chunk0 += D.nchunks;
    D.bias4_off = bias_off / 4;
    std::memcpy(reinterpret_cast<float*>(a.bias4) + bias_off, S.pc->bias_host, sizeof(float) * S.pc->cout);
    bias_off += S.pc->cout;
    D.gs = gs[s];
    D.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(D.n_pad >> 3) << 17) | (static_cast<uint32_t>(128 >> 4) << 24);
    D.w_gmem = S.pc->w;
  }
  a.chunks_per_tile = chunk0;
  if (last_p0_stage < 0) return 0;
  a.st[last_p0_stage].p0_last = 1;

  // ---- tensor maps
  alignas(64) ChainMaps maps;
  std::memset(&maps, 0, sizeof(maps));
  {
    const int sl = slab_of(buf_c[0]);
    const CUtensorMapSwizzle sw = sl == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (sl == 32 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(sp.in_c), static_cast<cuuint64_t>(sp.w), static_cast<cuuint64_t>(sp.h),
                                static_cast<cuuint64_t>(sp.batch)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(sp.in_cstride) * 2, static_cast<cuuint64_t>(sp.w) * sp.in_cstride * 2,
                                   static_cast<cuuint64_t>(sp.in_img_stride) * 2};
    const cuuint32_t box[4] = {static_cast<cuuint32_t>(sl), static_cast<cuuint32_t>(best.rw), static_cast<cuuint32_t>(best.th + 2 * m0), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    void* base = const_cast<__nv_bfloat16*>(sp.in) + sp.in_coff;
    const CUresult cr = encode_tiled()(&maps.in, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                       sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(AICAM_ERR_CUDA, "conv_chain: cuTensorMapEncodeTiled (input) failed with " + std::to_string(static_cast<int>(cr)));
  }
  bool map_used[3] = {false, false, false};
  for (int p = 0; p < a.npieces; ++p) map_used[a.piece_map[p]] = true;
  for (int m = 0; m < 3; ++m) {
    if (!map_used[m]) continue;
    const int pb = m == 0 ? 128 : (m == 1 ? 64 : 32);
    const CUtensorMapSwizzle sw = m == 0 ? CU_TENSOR_MAP_SWIZZLE_128B : (m == 1 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
    const cuuint64_t dims[4] = {static_cast<cuuint64_t>(last_c), static_cast<cuuint64_t>(sp.w), static_cast<cuuint64_t>(sp.h),
                                static_cast<cuuint64_t>(sp.batch)};
    const cuuint64_t strides[3] = {static_cast<cuuint64_t>(sp.out_cstride) * es, static_cast<cuuint64_t>(sp.w) * sp.out_cstride * es,
                                   static_cast<cuuint64_t>(sp.out_img_stride) * es};
    const cuuint32_t box[4] = {static_cast<cuuint32_t>(pb / es), static_cast<cuuint32_t>(best.tw), static_cast<cuuint32_t>(best.th), 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    void* base = static_cast<uint8_t*>(sp.out) + static_cast<size_t>(sp.out_coff) * es;
    const CUresult cr = encode_tiled()(&maps.out[m], sp.out_f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims,
                                       strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, sw, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(AICAM_ERR_CUDA, "conv_chain: cuTensorMapEncodeTiled (output) failed with " + std::to_string(static_cast<int>(cr)));
  }
  if (!map_used[0]) maps.out[0] = maps.out[map_used[1] ? 1 : 2];  // (the store warp prefetches map 0)

  const long long total_tiles = static_cast<long long>(sp.batch) * a.tiles_per_img;
  if (debug) {
    fprintf(stderr, "conv_chain: %dx%d x%d stages %d cin %d tile %dx%d (rw %d, strips %d) chunks", sp.h, sp.w, sp.batch, ns_, sp.in_c, best.th, best.tw,
            best.rw, best.strips);
    for (int s = 0; s < ns_; ++s) fprintf(stderr, " %d", best.nchunks[s]);
    fprintf(stderr, " resident %d p0x%d alias %d nslot %d smem %zu tiles %lld cost %.0f\n", a.resident, a.p0_nbuf, best.alias_buf, nslot, smem, total_tiles, best.cost);
  }
  if (int rc = ensure_dynamic_smem(conv_chain_kernel, SMEM_LIMIT)) return rc;
  dim3 grid(static_cast<unsigned>(std::min<long long>(total_tiles, num_sms)));
  size_t slot = 0;
  const bool prof = profile_begin(stream, &slot);
  {
    static const bool no_pdl = getenv("AICAM_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = dim3(CH_THREADS); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
    const cudaError_t le = cudaLaunchKernelEx(&cfg, conv_chain_kernel, a, maps);
    if (le != cudaSuccess) return fail(AICAM_ERR_CUDA, std::string("conv_chain: launch failed: ") + cudaGetErrorString(le));
  }
  if (prof) profile_end(stream, slot);
  count_launch();
  const int rc = last_launch("conv_chain_kernel");
  return rc ? rc : 1;
}

}  // namespace aicam

using namespace aicam;

extern "C" int aicam_conv_chain(const aicam_chain_desc* d, const void* in_nhwc, const float* const* weights_oihw, const float* const* bias,
                                void* out_nhwc, void* stream) {
  if (!d || !in_nhwc || !weights_oihw || !bias || !out_nhwc) return fail(AICAM_ERR_INVALID_ARG, "conv_chain: null argument");
  if (d->nstages < 2 || d->nstages > CHAIN_MAX_STAGES) return fail(AICAM_ERR_INVALID_ARG, "conv_chain: 2 to 5 stages");
  std::vector<PackedConv> pcs(d->nstages);
  ChainSpec sp;
  sp.nstages = d->nstages;
  int rc = AICAM_OK;
  int made = 0;
  for (int s = 0; s < d->nstages && !rc; ++s) {
    const aicam_chain_stage& S = d->st[s];
    rc = pack_conv_weights(weights_oihw[s], bias[s], S.cout, S.cin, S.ksize, 1, &pcs[s]);
    if (rc) break;
    ++made;
    ChainStageSpec& T = sp.st[s];
    T.pc = &pcs[s];
    T.act = S.act;
    T.nsrc = S.nsrc;
    for (int j = 0; j < 2; ++j) { T.src_buf[j] = S.src_buf[j]; T.src_coff[j] = S.src_coff[j]; T.src_c[j] = S.src_c[j]; }
    T.res_buf = S.res_mode ? S.res_buf : -1; T.res_coff = S.res_coff; T.res_mode = S.res_mode;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!rc) {
    const int cl = d->st[d->nstages - 1].cout;
    sp.in = static_cast<const __nv_bfloat16*>(in_nhwc);
    sp.in_cstride = d->in_c; sp.in_coff = 0; sp.in_c = d->in_c;
    sp.in_img_stride = static_cast<long long>(d->h) * d->w * d->in_c;
    sp.batch = d->batch; sp.h = d->h; sp.w = d->w;
    sp.out = out_nhwc; sp.out_cstride = cl; sp.out_coff = 0; sp.out_f32 = d->out_f32;
    sp.out_img_stride = static_cast<long long>(d->h) * d->w * cl;
    const int lrc = try_launch_conv_chain(sp, st);
    if (lrc < 0) rc = lrc;
    else if (lrc == 0) rc = fail(AICAM_ERR_UNSUPPORTED, "conv_chain: this chain / geometry is not eligible for the fused kernel");
  }
  const cudaError_t se = cudaStreamSynchronize(st);
  for (int s = 0; s < made; ++s) free_packed_conv(&pcs[s]);
  if (rc) return rc;
  if (se != cudaSuccess) return fail(AICAM_ERR_CUDA, std::string("conv_chain: ") + cudaGetErrorString(se));
  return AICAM_OK;
}
