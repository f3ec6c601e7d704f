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
constexpr int MAX_OPS = 288;               // MMA instructions of one 128-row chunk, summed over the stages
constexpr int MAX_GROUPS = 48;             // MMA groups (chunks issued together) per tile
constexpr uint32_t OFF_ST = 1040;         // shared-memory copies of the stage / buffer descriptors (indexed constant-bank
constexpr uint32_t OFF_BUF = 1840;        //   loads miss the constant cache once the parameters exceed a few KB)
constexpr uint32_t OFF_OPS = 2048;        // the op table (16 bytes per op)
constexpr uint32_t OFF_GROUPS = OFF_OPS + MAX_OPS * 16;  // the group table (32 bytes per group)
constexpr uint32_t OFF_DATA = 8192;
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
// One tcgen05.mma of a chunk, precomputed on the host (the issue loop is one 16-byte shared-memory load + a few adds per MMA):
//   A descriptor low word = sbase / 16 + a_off + chunk * a_step (+ the patch's second buffer), high word a_hi;
//   B descriptor low word = b_off (+ the ring slot's address / 16 when the weights are streamed).
struct ChainOpD {
  uint32_t a_off;    // ((buffer offset + slab offset + tap / shift rows) >> 4) + K offset inside the slab + LBO flag
  uint32_t a_step;   // bits 0-15: 128 rows in 16-byte units; bit 31: reads the patch (buffer 0); bits 16-27: ops in the weight stage
                     // this op starts (0: not the first op of a stage)
  uint32_t a_hi;
  uint32_t b_off;    // (byte offset >> 4) | N << 16
};
struct ChainStageD {
  int op0, nops;
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
  long long* trace;
  int ngroups;
  float4 bias4[128];
  ChainOpD ops[MAX_OPS];
  // MMA groups of one tile in issue order, 8 words each: op0 | nops << 16;  first chunk | chunks << 8 | flags << 24 (bit 0: the
  // patch is free after this group);  idesc;  number of READY barriers to wait for;  then their indices, one byte each
  uint32_t groups[MAX_GROUPS * 8];
};
static_assert(sizeof(ChainStageD) * CHAIN_MAX_STAGES <= OFF_BUF - OFF_ST, "stage descriptors do not fit their shared-memory area");
static_assert(sizeof(ChainBufD) * MAX_BUFS <= OFF_OPS - OFF_BUF, "buffer descriptors do not fit their shared-memory area");
static_assert(OFF_GROUPS + MAX_GROUPS * 32 <= OFF_DATA, "tables overlap the data area");
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

// debug trace of CTA 0 (AICAM_CONV_TRACE in aicam_conv_chain_bench): six tracing threads (MMA lane 0, thread 0 of each epilogue
// warpgroup, store-warp lane 0) own 600 (id, clock) slots each; the event count is kept in a register (tr_n)
#define CH_TRACE(kind, id)                                                                   \
  do {                                                                                       \
    if (tr_role >= 0 && tr_n < 600) {                                                        \
      a.trace[(tr_role * 600 + tr_n) * 2] = (static_cast<long long>(kind) << 32) | static_cast<long long>(id); \
      a.trace[(tr_role * 600 + tr_n) * 2 + 1] = clock64();                                   \
      ++tr_n;                                                                                \
    }                                                                                        \
  } while (0)

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
  int tr_role = -1, tr_n = 0;
  if (a.trace && blockIdx.x == 0) {
    if (threadIdx.x < 128 * NWG && (threadIdx.x & 127) == 0) tr_role = 1 + (threadIdx.x >> 7);
    else if (threadIdx.x == 128 * NWG) tr_role = 0;
    else if (threadIdx.x == 128 * NWG + 96) tr_role = 5;
  }
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
  {  // descriptors and tables: constant bank -> shared memory
    const int nops_total = a.st[a.nstages - 1].op0 + a.st[a.nstages - 1].nops;
    uint4* dst = reinterpret_cast<uint4*>(smem + OFF_OPS);
    for (int i = threadIdx.x; i < nops_total; i += CH_THREADS) dst[i] = make_uint4(a.ops[i].a_off, a.ops[i].a_step, a.ops[i].a_hi, a.ops[i].b_off);
    uint32_t* gdst = reinterpret_cast<uint32_t*>(smem + OFF_GROUPS);
    for (int i = threadIdx.x; i < a.ngroups * 8; i += CH_THREADS) gdst[i] = a.groups[i];
    uint32_t* sdst = reinterpret_cast<uint32_t*>(smem + OFF_ST);
    const uint32_t* ssrc = reinterpret_cast<const uint32_t*>(a.st);
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(ChainStageD) * CHAIN_MAX_STAGES / 4); i += CH_THREADS) sdst[i] = ssrc[i];
    uint32_t* bdst = reinterpret_cast<uint32_t*>(smem + OFF_BUF);
    const uint32_t* bsrc = reinterpret_cast<const uint32_t*>(a.buf);
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(ChainBufD) * MAX_BUFS / 4); i += CH_THREADS) bdst[i] = bsrc[i];
  }
  const ChainStageD* st_s = reinterpret_cast<const ChainStageD*>(smem + OFF_ST);
  const ChainBufD* buf_s = reinterpret_cast<const ChainBufD*>(smem + OFF_BUF);
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
    int e_slot = 0, e_wg = 0;  // accumulator slot / owning warpgroup of the running chunk index
    uint32_t e_ph = 0;
    const int nslot = a.nslot;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const TileXY t = tile_xy(a, tile);
      const int p0buf = a.p0_nbuf == 2 ? (it & 1) : 0;
      // the store of the previous tile has finished reading the staging tile (which may alias a buffer written below)
      mbar_wait(sbase + BAR_STAGE_FREE, (it & 1) ^ 1);
      // residual rows are read from the patch with ordinary loads: observe the TMA's completion barrier myself
      if (a.p0_res_readers) mbar_wait(sbase + BAR_P0_FULL + 8 * p0buf, a.p0_nbuf == 2 ? ((it >> 1) & 1) : (it & 1));
      for (int s = 0; s < a.nstages; ++s) {
        const ChainStageD& S = st_s[s];
        const bool last = S.dst_buf < 0;
        const int ngroups = S.n_pad >> 4;
        for (int k = 0; k < S.nchunks; ++k) {
          const int slot = e_slot;
          const uint32_t ph = e_ph;
          const bool mine = e_wg == wg;
          if (++e_slot == nslot) { e_slot = 0; e_ph ^= 1; }
          e_wg = (e_wg + 1) & (NWG - 1);
          if (!mine) continue;  // another warpgroup's chunk
          mbar_wait(sbase + BAR_ACC_FULL + 8 * slot, ph);
          tc_fence_after();
          CH_TRACE(4, (it << 16) | (s << 8) | k);
          const int q = k * 128 + rloc;
          const int yl = __float2int_rd((static_cast<float>(q) + 0.5f) * a.inv_rw);
          const int xl = q - yl * a.rw;
          const int py = t.y0 - S.margin + yl, px = t.x0 - S.margin + xl;
          const bool inside = py >= 0 && py < a.h && px >= 0 && px < a.w;
          // destination row in the next stage's buffer
          uint32_t d_row = 0, d_xor = 0, d_stride = 0;
          int d_shift = 0;
          if (!last) {
            const ChainBufD& D = buf_s[S.dst_buf];
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
            const ChainBufD& R = buf_s[S.res_buf];
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
          CH_TRACE(5, (it << 16) | (s << 8) | k);
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
    const int nslot = a.nslot, resident = a.resident, ngroups = a.ngroups;
    const uint32_t ns = static_cast<uint32_t>(a.ns), sbase16 = sbase >> 4;
    int slot0 = 0;       // accumulator slot of the next chunk to issue; its "drained" parity
    uint32_t sph = 1;
    if (resident) mbar_wait(sbase + BAR_W_FULL, 0);
    int it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
      const int p0buf = a.p0_nbuf == 2 ? (it & 1) : 0;
      const uint32_t p0par = a.p0_nbuf == 2 ? ((it >> 1) & 1) : (it & 1);
      mbar_wait(sbase + BAR_P0_FULL + 8 * p0buf, p0par);
      tc_fence_after();
      CH_TRACE(8, it << 16);
      const uint32_t p0add = (static_cast<uint32_t>(p0buf) * a.p0_stride) >> 4;
      const uint4* gtab = reinterpret_cast<const uint4*>(smem + OFF_GROUPS);
      for (int gi = 0; gi < ngroups; ++gi) {
        const uint4 gd = gtab[2 * gi], gw = gtab[2 * gi + 1];
        const int nops = static_cast<int>(gd.x >> 16), k0 = static_cast<int>(gd.y & 255u), nk = static_cast<int>((gd.y >> 8) & 255u);
        const uint32_t idesc = gd.z;
        CH_TRACE(1, (it << 16) | gi);
        {  // the accumulator slots of the group have been drained
          int sl = slot0;
          uint32_t ph = sph;
          for (int k = 0; k < nk; ++k) {
            mbar_wait(sbase + BAR_ACC_EMPTY + 8 * sl, ph);
            if (++sl == nslot) { sl = 0; ph ^= 1; }
          }
        }
        // every row this group reads from a stage-written buffer has been published
        for (int w = 0; w < static_cast<int>(gd.w); ++w) {
          const uint32_t word = w < 4 ? gw.x : (w < 8 ? gw.y : (w < 12 ? gw.z : gw.w));
          mbar_wait(sbase + BAR_READY + 8 * ((word >> (8 * (w & 3))) & 255u), it & 1);
        }
        tc_fence_after();
        CH_TRACE(2, (it << 16) | gi);
        {
          const uint4* ops = reinterpret_cast<const uint4*>(smem + OFF_OPS) + (gd.x & 0xffffu);
          int i = 0;
          while (i < nops) {
            const int run = static_cast<int>((ops[i].y >> 16) & 0xfffu);
            uint32_t b_base16 = sbase16;  // resident weights: b_off is relative to the start of shared memory
            if (!resident) {
              mbar_wait(sbase + BAR_B_FULL + 8 * rs, rp);
              tc_fence_after();
              b_base16 = (sbase + a.ring_off + rs * a.ring_slot_bytes) >> 4;
            }
            int sl = slot0;
            for (int k = 0; k < nk; ++k) {
              const uint32_t d_tmem = tmem0 + static_cast<uint32_t>(sl) * ns;
              const uint32_t kk = static_cast<uint32_t>(k0 + k);
#pragma unroll 4
              for (int r = 0; r < run; ++r) {
                const uint4 op = ops[i + r];
                const uint32_t a_lo = sbase16 + op.x + kk * (op.y & 0xffffu) + ((op.y >> 31) ? p0add : 0u);
                mma_issue(leader, d_tmem, a_lo, op.z, b_base16 + op.w, b_hi, idesc, (i + r) != 0 ? 1u : 0u);
              }
              if (++sl == nslot) sl = 0;
            }
            if (!resident) {
              tc_commit_if(leader, sbase + BAR_B_EMPTY + 8 * rs);
              if (++rs == static_cast<uint32_t>(a.ring_n)) { rs = 0; rp ^= 1; }
            }
            i += run;
          }
        }
        for (int k = 0; k < nk; ++k) {
          tc_commit_if(leader, sbase + BAR_ACC_FULL + 8 * slot0);
          if (++slot0 == nslot) { slot0 = 0; sph ^= 1; }
        }
        if (gd.y >> 24) tc_commit_if(leader, sbase + BAR_P0_EMPTY + 8 * p0buf);
        CH_TRACE(3, (it << 16) | gi);
      }
    }
    tc_fence_before();
  } else if (warp == 4 * NWG + 1) {
    // ================================================================== patch producer
    if (lane == 0) {
      pdl_wait();  // the patch is the previous layer's output
      const ChainBufD& B = buf_s[0];
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
          const ChainStageD& S = st_s[s];
          const uint32_t bytes = static_cast<uint32_t>(S.taps * S.k8_per_tap * S.n_pad * 16);
          const uint8_t* srcp = reinterpret_cast<const uint8_t*>(S.w_gmem);
          for (uint32_t off = 0; off < bytes; off += 16384)
            bulk_g2s(sbase + a.w_off + S.w_smem_off + off, srcp + off, min(16384u, bytes - off), sbase + BAR_W_FULL);
        }
      } else {
        uint32_t rs = 0, rp = 1;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
          for (int s = 0; s < a.nstages; ++s) {
            const ChainStageD& S = st_s[s];
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
      CH_TRACE(6, it << 16);
      if (lane == 0) {
        for (int p = 0; p < a.npieces; ++p)
          tma_store_4d(&maps.out[a.piece_map[p]], sbase + a.stage_off + a.piece_off[p], a.piece_c0[p], t.x0, t.y0, t.n);
        bulk_commit();
        bulk_wait_read_all();
        mbar_arrive(sbase + BAR_STAGE_FREE);
      }
      __syncwarp();
      CH_TRACE(7, it << 16);
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
  a.trace = sp.trace;
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
    // chunks issued per commit: every commit exposes ~700 cycles of tensor-pipe latency to the issuing warp (measured), so a
    // stage is issued as few groups as the accumulator slots allow (AICAM_CHAIN_GROUP overrides it for resident weights)
    static const int group_env = getenv("AICAM_CHAIN_GROUP") ? atoi(getenv("AICAM_CHAIN_GROUP")) : 0;
    D.group = resident ? std::max(1, std::min(group_env > 0 ? group_env : D.nchunks, nslot)) : D.nchunks;
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
  // ---- the op table: every MMA of one chunk, stage by stage, in issue order (tap, source, 16-channel group); a weight
  // stage (streamed form) is a run of ops that share one ring slot
  {
    int nops = 0;
    for (int s = 0; s < ns_; ++s) {
      ChainStageD& D = a.st[s];
      D.op0 = nops;
      const uint32_t lbo = static_cast<uint32_t>(D.n_pad);
      for (int tap = 0; tap < D.taps; ++tap) {
        const int dy = tap / D.ksize, dx = tap % D.ksize;
        for (int j = 0; j < D.nsrc; ++j) {
          const ChainSrcD& src = D.src[j];
          const ChainBufD& B = a.buf[src.buf];
          for (int c = 0; c < src.nc16; ++c) {
            if (nops >= MAX_OPS) return 0;
            const int c16 = src.c16_0 + c;
            const uint32_t addr = B.off + static_cast<uint32_t>(c16 >> B.c16_shift) * B.slab_stride +
                                  static_cast<uint32_t>(src.shift + dy * best.rw + dx) * B.row_bytes;
            ChainOpD& op = a.ops[nops];
            op.a_off = (addr >> 4) + static_cast<uint32_t>(c16 & ((1 << B.c16_shift) - 1)) * 2 + (1u << 16);
            op.a_step = (128 * B.row_bytes) >> 4;
            if (src.buf == 0) op.a_step |= 1u << 31;
            op.a_hi = B.a_hi;
            const int sub0 = (c / D.gs) * D.gs;  // first group of this op's weight stage
            if (resident) {
              const uint32_t wb = a.w_off + D.w_smem_off + static_cast<uint32_t>(tap * D.k8_per_tap + src.wk8 + c * 2) * lbo * 16;
              op.b_off = (wb >> 4) | (lbo << 16);
              if (nops == D.op0) op.a_step |= static_cast<uint32_t>(D.taps * D.k8_per_tap / 2) << 16;  // one run: the whole stage
            } else {
              op.b_off = ((static_cast<uint32_t>(c - sub0) * 2 * lbo * 16) >> 4) | (lbo << 16);
              if (c == sub0) op.a_step |= static_cast<uint32_t>(std::min(D.gs, src.nc16 - sub0)) << 16;
            }
            ++nops;
          }
        }
      }
      D.nops = nops - D.op0;
      if (D.nops > 0xfff) return 0;
    }
  }
  // ---- the group table: which READY barriers each group waits for is static (the consumer's reach into each source
  // buffer, minus what earlier groups of the tile have already waited for)
  {
    int waited[MAX_BUFS] = {0};
    int ng = 0;
    for (int s = 0; s < ns_; ++s) {
      const ChainStageD& D = a.st[s];
      for (int g0 = 0; g0 < D.nchunks; g0 += D.group) {
        if (ng >= MAX_GROUPS) return 0;
        const int g1 = std::min(g0 + D.group, D.nchunks);
        uint32_t* G = a.groups + ng * 8;
        std::memset(G, 0, 32);
        G[0] = static_cast<uint32_t>(D.op0) | (static_cast<uint32_t>(D.nops) << 16);
        G[1] = static_cast<uint32_t>(g0) | (static_cast<uint32_t>(g1 - g0) << 8) | ((D.p0_last && g1 == D.nchunks) ? (1u << 24) : 0u);
        G[2] = D.idesc;
        int nwait = 0;
        for (int j = 0; j < D.nsrc; ++j) {
          const int b = D.src[j].buf;
          if (b == 0) continue;
          const int reach = g1 * 128 + D.src[j].shift + (D.ksize - 1) * (best.rw + 1);
          const int need = std::min(a.buf[b].nchunks, (reach + 127) >> 7);
          for (; waited[b] < need; ++waited[b]) {
            if (nwait >= 16) return 0;
            G[4 + nwait / 4] |= static_cast<uint32_t>(b * MAX_CHUNKS + waited[b]) << (8 * (nwait % 4));
            ++nwait;
          }
        }
        G[3] = static_cast<uint32_t>(nwait);
        ++ng;
      }
    }
    a.ngroups = ng;
  }

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

// Micro-benchmark of one fused chain on device-resident pseudo-random data (weights included): `iters` launches bracketed
// by CUDA events.  AICAM_CONV_TRACE=1 prints the event trace of CTA 0 of one extra launch.
extern "C" int aicam_conv_chain_bench(const aicam_chain_desc* d, int iters, double* mean_ms, void* stream) {
  if (!d || !mean_ms || iters <= 0) return fail(AICAM_ERR_INVALID_ARG, "conv_chain_bench: bad arguments");
  if (d->nstages < 2 || d->nstages > CHAIN_MAX_STAGES) return fail(AICAM_ERR_INVALID_ARG, "conv_chain_bench: 2 to 5 stages");
  std::vector<PackedConv> pcs(d->nstages);
  ChainSpec sp;
  sp.nstages = d->nstages;
  int rc = AICAM_OK, made = 0;
  uint32_t seed = 12345u;
  for (int s = 0; s < d->nstages && !rc; ++s) {
    const aicam_chain_stage& S = d->st[s];
    std::vector<float> w(static_cast<size_t>(S.cout) * S.cin * S.ksize * S.ksize), b(S.cout, 0.1f);
    for (auto& v : w) { seed = seed * 1664525u + 1013904223u; v = (static_cast<int>(seed >> 16) % 2001 - 1000) * 1e-4f; }
    rc = pack_conv_weights(w.data(), b.data(), S.cout, S.cin, S.ksize, 1, &pcs[s]);
    if (rc) break;
    ++made;
    ChainStageSpec& T = sp.st[s];
    T.pc = &pcs[s];
    T.act = S.act; T.nsrc = S.nsrc;
    for (int j = 0; j < 2; ++j) { T.src_buf[j] = S.src_buf[j]; T.src_coff[j] = S.src_coff[j]; T.src_c[j] = S.src_c[j]; }
    T.res_buf = S.res_mode ? S.res_buf : -1; T.res_coff = S.res_coff; T.res_mode = S.res_mode;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  __nv_bfloat16* in = nullptr;
  void* out = nullptr;
  float ms = 0.0f;
  if (!rc) {
    const int cl = d->st[d->nstages - 1].cout;
    const size_t in_elems = static_cast<size_t>(d->batch) * d->h * d->w * d->in_c, out_elems = static_cast<size_t>(d->batch) * d->h * d->w * cl;
    if (cudaMalloc(&in, in_elems * 2) != cudaSuccess || cudaMalloc(&out, out_elems * (d->out_f32 ? 4 : 2)) != cudaSuccess) rc = fail(AICAM_ERR_CUDA, "conv_chain_bench: cudaMalloc failed");
    if (!rc) {
      cudaMemset(in, 0x3c, in_elems * 2);
      sp.in = in; sp.in_cstride = d->in_c; sp.in_coff = 0; sp.in_c = d->in_c;
      sp.in_img_stride = static_cast<long long>(d->h) * d->w * d->in_c;
      sp.batch = d->batch; sp.h = d->h; sp.w = d->w;
      sp.out = out; sp.out_cstride = cl; sp.out_coff = 0; sp.out_f32 = d->out_f32;
      sp.out_img_stride = static_cast<long long>(d->h) * d->w * cl;
      for (int i = 0; i < 3 && !rc; ++i) {
        const int lrc = try_launch_conv_chain(sp, st);
        if (lrc < 0) rc = lrc;
        else if (lrc == 0) rc = fail(AICAM_ERR_UNSUPPORTED, "conv_chain_bench: this chain / geometry is not eligible for the fused kernel");
      }
      if (!rc && getenv("AICAM_CONV_TRACE")) {
        long long* tr = nullptr;
        const size_t tr_words = 6 * 600 * 2;
        cudaMalloc(&tr, 8 * tr_words);
        cudaMemset(tr, 0, 8 * tr_words);
        sp.trace = tr;
        try_launch_conv_chain(sp, st);
        cudaStreamSynchronize(st);
        sp.trace = nullptr;
        std::vector<long long> h(tr_words);
        cudaMemcpy(h.data(), tr, 8 * tr_words, cudaMemcpyDeviceToHost);
        cudaFree(tr);
        struct Ev { long long t, id; int role; };
        std::vector<Ev> ev;
        for (int r = 0; r < 6; ++r)
          for (int i = 0; i < 600; ++i)
            if (h[(r * 600 + i) * 2 + 1]) ev.push_back({h[(r * 600 + i) * 2 + 1], h[(r * 600 + i) * 2], r});
        std::sort(ev.begin(), ev.end(), [](const Ev& x, const Ev& y) { return x.t < y.t; });
        static const char* names[] = {"?", "mma_begin", "mma_deps_ok", "mma_issued", "epi_acc_full", "epi_done", "store_tile_done", "store_read_done", "mma_p0_full"};
        static const char* roles[] = {"MMA", "WG0", "WG1", "WG2", "WG3", "STORE"};
        printf("chain trace of CTA 0 (%zu events): cycles since the first event, role, kind, tile.stage.chunk\n", ev.size());
        const int max_ev = getenv("AICAM_TRACE_EVENTS") ? atoi(getenv("AICAM_TRACE_EVENTS")) : 240;
        for (size_t i = 0; i < ev.size() && static_cast<int>(i) < max_ev; ++i) {
          const long long id = ev[i].id & 0xffffffffll, kind = ev[i].id >> 32;
          printf("  %8lld %-5s %-15s %lld.%lld.%lld\n", ev[i].t - ev[0].t, roles[ev[i].role], names[kind < 9 ? kind : 0], id >> 16, (id >> 8) & 255, id & 255);
        }
        fflush(stdout);
      }
      if (!rc) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, st);
        for (int i = 0; i < iters && !rc; ++i) {
          const int lrc = try_launch_conv_chain(sp, st);
          if (lrc < 0) rc = lrc;
        }
        cudaEventRecord(e1, st);
        const cudaError_t se = cudaStreamSynchronize(st);
        cudaEventElapsedTime(&ms, e0, e1);
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        if (!rc && se != cudaSuccess) rc = fail(AICAM_ERR_CUDA, std::string("conv_chain_bench: ") + cudaGetErrorString(se));
      }
    }
  }
  if (in) cudaFree(in);
  if (out) cudaFree(out);
  for (int s = 0; s < made; ++s) free_packed_conv(&pcs[s]);
  if (rc) return rc;
  *mean_ms = ms / iters;
  return AICAM_OK;
}
