// K4 / K5 -- SBR and parametric-stereo signal path (frame-parallel pipeline over tiles of frames, see below):
//   32-band QMF analysis     sbr/AnalysisFilterbank.java:9-73
//   HF generation            sbr/HFGeneration.java:17-245       (covariance LPC + patching)
//   HF adjustment            sbr/HFAdjustment.java:20-415       (envelope estimate, gains, limiter, assembly)
//   parametric stereo        ps/PSImpl.java:685-707, ps/Filterbank.java
//   64-band QMF synthesis    sbr/SynthesisFilterbank64.java:9-79
//   Math.round / clamp / interleave as S/SampleBuffer.java:168-209
// Every floating-point operation is the binary32 operation of the Java code on the same operands in the same order
// (this file is compiled with --fmad=false), so the PCM is bit-identical to the reference's, like the AAC-LC path.
#pragma once
#include "jaadb_types.cuh"
#include "k2_filterbank.cuh"
#include "sbr_types.cuh"
#include "generated/jaad_dct32.h"   // DCT4_32 / DST4_32 operation lists, qmf32_pre_twiddle (tools/extract_dct32.py)

namespace jaadb {

// tables addressed with compile-time indices by the unrolled DCT / polyphase code (uploaded per device at engine start)
__constant__ float c_sbr_dct4[192];
__constant__ float c_sbr_w_real[16];
__constant__ float c_sbr_w_imag[16];
__constant__ float c_sbr_qmf_c[640];
__constant__ int c_sbr_bit_rev[32];

// DCT.fft_dif (sbr/DCT.java:135-345), fully unrolled: Real / Imag live in registers
__device__ __forceinline__ void sbr_fft_dif(float (&Real)[32], float (&Imag)[32]) {
  float w_real, w_imag, p1r, p1i, p2r, p2i;
#pragma unroll
  for (int i = 0; i < 16; i++) {
    p1r = Real[i]; p1i = Imag[i];
    p2r = Real[i + 16]; p2i = Imag[i + 16];
    w_real = c_sbr_w_real[i]; w_imag = c_sbr_w_imag[i];
    p1r -= p2r; p1i -= p2i;
    Real[i] += p2r; Imag[i] += p2i;
    Real[i + 16] = ((p1r * w_real) - (p1i * w_imag));
    Imag[i + 16] = ((p1r * w_imag) + (p1i * w_real));
  }
#pragma unroll
  for (int j = 0; j < 8; j++) {
    w_real = c_sbr_w_real[2 * j]; w_imag = c_sbr_w_imag[2 * j];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int i = j + 16 * half, i2 = i + 8;
      p1r = Real[i]; p1i = Imag[i]; p2r = Real[i2]; p2i = Imag[i2];
      p1r -= p2r; p1i -= p2i;
      Real[i] += p2r; Imag[i] += p2i;
      Real[i2] = ((p1r * w_real) - (p1i * w_imag));
      Imag[i2] = ((p1r * w_imag) + (p1i * w_real));
    }
  }
#pragma unroll
  for (int i = 0; i < 32; i += 8) {
    const int i2 = i + 4;
    p1r = Real[i]; p1i = Imag[i]; p2r = Real[i2]; p2i = Imag[i2];
    Real[i] += p2r; Imag[i] += p2i;
    Real[i2] = p1r - p2r; Imag[i2] = p1i - p2i;
  }
  w_real = c_sbr_w_real[4];
#pragma unroll
  for (int i = 1; i < 32; i += 8) {
    const int i2 = i + 4;
    p1r = Real[i]; p1i = Imag[i]; p2r = Real[i2]; p2i = Imag[i2];
    p1r -= p2r; p1i -= p2i;
    Real[i] += p2r; Imag[i] += p2i;
    Real[i2] = (p1r + p1i) * w_real;
    Imag[i2] = (p1i - p1r) * w_real;
  }
#pragma unroll
  for (int i = 2; i < 32; i += 8) {
    const int i2 = i + 4;
    p1r = Real[i]; p1i = Imag[i]; p2r = Real[i2]; p2i = Imag[i2];
    Real[i] += p2r; Imag[i] += p2i;
    Real[i2] = p1i - p2i;
    Imag[i2] = p2r - p1r;
  }
  w_real = c_sbr_w_real[12];
#pragma unroll
  for (int i = 3; i < 32; i += 8) {
    const int i2 = i + 4;
    p1r = Real[i]; p1i = Imag[i]; p2r = Real[i2]; p2i = Imag[i2];
    p1r -= p2r; p1i -= p2i;
    Real[i] += p2r; Imag[i] += p2i;
    Real[i2] = (p1r - p1i) * w_real;
    Imag[i2] = (p1r + p1i) * w_real;
  }
#pragma unroll
  for (int i = 0; i < 32; i += 4) {
    const int i2 = i + 2;
    p1r = Real[i]; p1i = Imag[i]; p2r = Real[i2]; p2i = Imag[i2];
    Real[i] += p2r; Imag[i] += p2i;
    Real[i2] = p1r - p2r; Imag[i2] = p1i - p2i;
  }
#pragma unroll
  for (int i = 1; i < 32; i += 4) {
    const int i2 = i + 2;
    p1r = Real[i]; p1i = Imag[i]; p2r = Real[i2]; p2i = Imag[i2];
    Real[i] += p2r; Imag[i] += p2i;
    Real[i2] = p1i - p2i;
    Imag[i2] = p2r - p1r;
  }
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    const int i2 = i + 1;
    p1r = Real[i]; p1i = Imag[i]; p2r = Real[i2]; p2i = Imag[i2];
    Real[i] += p2r; Imag[i] += p2i;
    Real[i2] = p1r - p2r; Imag[i2] = p1i - p2i;
  }
}

// DCT.dct4_kernel (sbr/DCT.java:347-391)
__device__ __forceinline__ void sbr_dct4_kernel(float (&in_real)[32], float (&in_imag)[32], float (&out_real)[32], float (&out_imag)[32]) {
#pragma unroll
  for (int i = 0; i < 32; i++) {
    const float x_re = in_real[i], x_im = in_imag[i];
    const float tmp = (x_re + x_im) * c_sbr_dct4[i];
    in_real[i] = (x_im * c_sbr_dct4[i + 64]) + tmp;
    in_imag[i] = (x_re * c_sbr_dct4[i + 32]) + tmp;
  }
  sbr_fft_dif(in_real, in_imag);
  constexpr int rev[32] = {0, 16, 8, 24, 4, 20, 12, 28, 2, 18, 10, 26, 6, 22, 14, 30, 1, 17, 9, 25, 5, 21, 13, 29, 3, 19, 11, 27, 7, 23, 15, 31};
#pragma unroll
  for (int i = 0; i < 32; i++) {
    if (i == 16) {
      out_imag[16] = (in_imag[1] - in_real[1]) * c_sbr_dct4[16 + 3 * 32];
      out_real[16] = (in_real[1] + in_imag[1]) * c_sbr_dct4[16 + 3 * 32];
    } else {
      const float x_re = in_real[rev[i]], x_im = in_imag[rev[i]];
      const float tmp = (x_re + x_im) * c_sbr_dct4[i + 3 * 32];
      out_real[i] = (x_im * c_sbr_dct4[i + 5 * 32]) + tmp;
      out_imag[i] = (x_re * c_sbr_dct4[i + 4 * 32]) + tmp;
    }
  }
}

// ps/PSTables.java:26-45 (20-band configuration): group borders and the parameter band of each group
__device__ __forceinline__ int ps_group_border(int gr) {
  constexpr int gb[23] = {6, 7, 0, 1, 2, 3, 9, 8, 10, 11, 3, 4, 5, 6, 7, 8, 9, 11, 14, 18, 23, 35, 64};
  int v = 64;
#pragma unroll
  for (int i = 0; i < 23; ++i) if (i == gr) v = gb[i];
  return v;
}
__device__ __forceinline__ int ps_bk(int gr) { return gr == 0 ? 1 : (gr == 1 ? 0 : gr - 2); }   // map_group2bk20 & ~NEGATE_IPD_MASK

constexpr int kK4Threads = 64;
// Shared-memory row strides are odd so that the one-thread-per-time-slot phases (the DCTs), where the 32 lanes of a warp
// address the same column of 32 different rows, spread over all 32 banks.
constexpr int kXsStride = 129;              // one staged Xsbr slot: 128 + 1
constexpr int kVbStride = 129;              // one synthesis v-vector: 128 + 1
#define INB(i) inbuf[(i) + ((i) >> 5)]      // analysis input: one pad float per 32

// One SBR channel of one stream inside a batch.
struct K4RunDev {
  int32_t stream_slot;
  uint32_t first, count;    // into run_frames
  uint32_t sbr_base;        // first SbrFrameDev pair of the element's run
  uint8_t chan;             // channel inside the element (0/1)
  uint8_t ch_slot;          // channel slot inside the stream (core PCM / persistent state index)
  uint8_t out_ch;           // output channel index
  uint8_t n_out;            // output channels of the stream
  uint8_t dup;              // mono element: copy the result to the second output channel (SBR1.process)
  uint8_t ds;               // down-sampled SBR (SBR.isSBRDownSampled): 32-band synthesis, 1024 output samples per frame
  uint8_t pad[2];
  uint32_t ps_base;         // first PsFrameDev of the run (SBR+PS streams)
};

// =====================================================================================================================
// Frame-parallel SBR pipeline.  The QMF banks are FIR structures: analysis slot l of a frame needs
// the 320 newest core samples, synthesis slot l the ten newest v-vectors, so every (channel, frame) pair can run on its
// own warp / CTA once the few truly recursive pieces are out of the way.  The Xsbr matrix of SBR.java lives in global
// memory for a TILE of frames (ft consecutive frames of every run):
//
//   xg[run][8 + 32 * ft rows][64 bands][re, im]      row 32 * o + r = row r of Xsbr for the o-th processed frame
//                                                    of the tile (rows 32..39 of a frame ARE rows 0..7 of the next one,
//                                                    which is what SBR.sbr_save_matrix copies)
//
//   K4a  k4a_analysis_kernel   warp per (channel, frame): 32-band analysis, one time slot per lane       -> xg low band
//   K4b  k4b_hf_kernel         warp per channel, frames of the tile in order (chirp factors, smoothing ring, noise and
//                              sine phase are recursive): HF generation + HF adjustment, one band per lane -> xg high band
//   K4c  k4c_synthesis_kernel  CTA per (channel, frame): 64-band synthesis; the nine v-vectors a frame inherits are
//                              recomputed from the previous frame's rows (or come from the carried state at a tile start)
//   k4_commit_kernel           flips the double-buffered v-vector state
// Frames that do not run the SBR tool (mode 0, failed frames) take no rows: K3 numbers the processed frames (ord) and
// links them (back / fwd).
constexpr int kXgRow = 128;   // floats per row of xg

struct K4Tile {
  uint32_t lo, ft;    // frames [lo, lo + ft) of every run
  uint32_t rows;      // rows per run in xg: 8 + 32 * ft
};

__device__ __forceinline__ const SbrFrameDev* k4_frame(const SbrFrameDev* sframes, const K4RunDev& run, uint32_t it) {
  return sframes + ((size_t)run.sbr_base + it) * 2 + run.chan;
}
__device__ __forceinline__ float2 ldg2(const float* p) { return __ldcg(reinterpret_cast<const float2*>(p)); }
__device__ __forceinline__ void st2(float* p, float a, float b) { *reinterpret_cast<float2*>(p) = make_float2(a, b); }

// ---- K4a: 32-band QMF analysis (sbr/AnalysisFilterbank.java:9-73)
constexpr int kK4aWarps = 4;
constexpr int kK4aStage = 32 * 65;   // floats per warp: the 1312 input samples first, then the [32 slots][64 + 1] output
constexpr size_t k4a_smem_bytes() { return sizeof(float) * kK4aStage * kK4aWarps; }

__global__ void __launch_bounds__(32 * kK4aWarps)
k4a_analysis_kernel(const K4RunDev* __restrict__ runs, uint32_t n_runs, const RunFrameDev* __restrict__ run_frames,
                    const SbrFrameDev* __restrict__ sframes, const float* __restrict__ core, const SbrChanDev* __restrict__ chans,
                    float* __restrict__ xg, K4Tile tile) {
  extern __shared__ __align__(16) float k4a_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t w = blockIdx.x * kK4aWarps + warp;
  if (w >= n_runs * tile.ft) return;
  const uint32_t r = w / tile.ft, it = tile.lo + w % tile.ft;
  const K4RunDev run = runs[r];
  if (it >= run.count) return;
  const SbrFrameDev* fp = k4_frame(sframes, run, it);
  const int mode = fp->mode;
  if (fp->frame_status != 0 || mode == 0) return;
  const uint32_t o = fp->ord - k4_frame(sframes, run, tile.lo)->ord;
  const int kx = mode == 2 ? fp->kx : 32;
  const uint32_t back = fp->back;
  float* X = xg + ((size_t)r * tile.rows + 32 * (size_t)o) * kXgRow;
  const SbrChanDev* st = chans + (size_t)run.stream_slot * kSbrChansPerStream + run.ch_slot;
  if (o == 0) {
    // the rows the previous tile (or batch) left behind
    const float4* s4 = reinterpret_cast<const float4*>(&st->xsbr[0][0][0]);
    float4* d4 = reinterpret_cast<float4*>(X);
    for (int i = lane; i < kSbrHfGen * 32; i += 32) d4[i] = s4[i];
  }
  float* inbuf = k4a_smem + warp * kK4aStage;
  {
    // all 17 loads of the lane are issued before the first one is consumed
    const float* hist = back ? core + ((size_t)run_frames[run.first + it - back].ics_base + run.ch_slot) * 1024 + 736 : st->ana_hist;
    const float4* cs = reinterpret_cast<const float4*>(core + ((size_t)run_frames[run.first + it].ics_base + run.ch_slot) * 1024);
    float h[9];
    float4 c[8];
#pragma unroll
    for (int u = 0; u < 9; ++u) h[u] = hist[lane + 32 * u];
#pragma unroll
    for (int u = 0; u < 8; ++u) c[u] = __ldg(cs + lane + 32 * u);
#pragma unroll
    for (int u = 0; u < 9; ++u) INB(lane + 32 * u) = h[u];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int i = lane + 32 * u;
      INB(288 + 4 * i) = c[u].x; INB(288 + 4 * i + 1) = c[u].y; INB(288 + 4 * i + 2) = c[u].z; INB(288 + 4 * i + 3) = c[u].w;
    }
  }
  __syncwarp();
  const int xi = 288 + 32 * lane + 31;   // newest sample of the slot; sample xi - j = v[v_index + j] of the reference
  float in_real[32], in_imag[32], out_real[32], out_imag[32];
#pragma unroll
  for (int n = 0; n < 64; ++n) {
    const float u = (INB(xi - n) * c_sbr_qmf_c[2 * n]) + (INB(xi - (n + 64)) * c_sbr_qmf_c[2 * (n + 64)]) +
                    (INB(xi - (n + 128)) * c_sbr_qmf_c[2 * (n + 128)]) + (INB(xi - (n + 192)) * c_sbr_qmf_c[2 * (n + 192)]) +
                    (INB(xi - (n + 256)) * c_sbr_qmf_c[2 * (n + 256)]);
    // reordering of AnalysisFilterbank.java:40-47
    if (n == 0) in_real[0] = u;
    else if (n == 1) in_imag[31] = u;
    else if (n <= 31) in_imag[32 - n] = u;
    else if (n == 32) in_imag[0] = u;
    else if (n == 33) in_real[31] = -u;
    else in_real[64 - n] = -u;
  }
  sbr_dct4_kernel(in_real, in_imag, out_real, out_imag);
  __syncwarp();   // every lane is done with the input samples: the region becomes the output stage
  float* stage = inbuf + lane * 65;
#pragma unroll
  for (int n = 0; n < 16; n++) {
    if (2 * n + 1 < kx) {
      stage[4 * n] = 2.0f * out_real[n];
      stage[4 * n + 1] = 2.0f * out_imag[n];
      stage[4 * n + 2] = -2.0f * out_imag[31 - n];
      stage[4 * n + 3] = -2.0f * out_real[31 - n];
    } else {
      if (2 * n < kx) { stage[4 * n] = 2.0f * out_real[n]; stage[4 * n + 1] = 2.0f * out_imag[n]; }
      else { stage[4 * n] = 0; stage[4 * n + 1] = 0; }
      stage[4 * n + 2] = 0;
      stage[4 * n + 3] = 0;
    }
  }
  __syncwarp();
  // rows 8..39 of the frame: the low band, and zeros above it (what the reference's fresh / shifted matrix holds there)
  for (int row = 0; row < 32; ++row) {
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (lane < 16) {
      const float* s = inbuf + row * 65 + 4 * lane;
      v = make_float4(s[0], s[1], s[2], s[3]);
    }
    reinterpret_cast<float4*>(X + (size_t)(kSbrHfGen + row) * kXgRow)[lane] = v;
  }
}

// ---- K4b: HF generation (sbr/HFGeneration.java:17-245) + HF adjustment (sbr/HFAdjustment.java:20-415)
#ifndef K4B_WARPS
#define K4B_WARPS 10
#endif
constexpr int kK4bWarps = K4B_WARPS;
#ifndef K4B_MIN_BLOCKS
#define K4B_MIN_BLOCKS 2
#endif
#ifndef K4B_PREFETCH
#define K4B_PREFETCH 3   // bit 0: this frame's generated band before the gains, bit 1: the next frame's low band before the assembly
#endif
// K4B_ALIGN: the warps of a CTA pass the phases of a frame together (1: generation, envelope estimate, gains, assembly; 2: in
// front of the two long ones, generation and assembly, only -- measured 57.1 / 141.5 ms against 57.7 / 141.6 ms for the SBR
// stage of configs 3 / 4; 3: generation, gains, assembly: 57.3 / 141.8 ms).  They
// are independent channels; the barriers are there for the instruction caches only (87 KB of SASS, 53 KB of it on the
// path of a regular frame, every phase a loop of 4 .. 10 KB: warps spread over the phases evict each other's loops).
#ifndef K4B_ALIGN
#define K4B_ALIGN 2
#endif
#if K4B_ALIGN == 1
#define K4B_PHASE() __syncthreads()
#define K4B_PHASE_MINOR() __syncthreads()
#define K4B_SKIP_FRAME() do { __syncthreads(); __syncthreads(); __syncthreads(); __syncthreads(); } while (0)
#elif K4B_ALIGN == 2   // only in front of the two long phases (generation, assembly)
#define K4B_PHASE() __syncthreads()
#define K4B_PHASE_MINOR() do {} while (0)
#define K4B_SKIP_FRAME() do { __syncthreads(); __syncthreads(); } while (0)
#elif K4B_ALIGN == 3   // generation, gains, assembly
#define K4B_PHASE() __syncthreads()
#define K4B_PHASE_MINOR() do {} while (0)
#define K4B_SKIP_FRAME() do { __syncthreads(); __syncthreads(); __syncthreads(); } while (0)
#else
#define K4B_PHASE() do {} while (0)
#define K4B_PHASE_MINOR() do {} while (0)
#define K4B_SKIP_FRAME() do {} while (0)
#endif
constexpr int kK4bMaxNL = 32;
constexpr int kK4bPwCols = 50;   // >= SBR.MAX_M, even
struct K4bGain {
  float G[kSbrMaxLE][64], Q[kSbrMaxLE][64], S[kSbrMaxLE][64];   // G_lim_boost, Q_M_lim_boost, S_M_boost (contiguous)
  float gmax[kSbrMaxLE][kK4bMaxNL], acc1[kSbrMaxLE][kK4bMaxNL], boost[kSbrMaxLE][kK4bMaxNL];
  uint8_t rb[2][64];             // band of f_table_res[res] that holds k = m + kx
  uint8_t nb[64], lb[64];        // noise-floor band, limiter band
  uint8_t sflag[kSbrMaxLE][64];  // S_index_mapped != 0
};
// Of a frame's record only the part behind the dequantised envelopes is staged in shared memory; E_orig / Q_div / Q_div2
// (a handful of reads per band in calculate_gain) come straight from the record in global memory, and Channel.E_curr lives
// in the channel's state (SbrChanDev::E_curr; the warp that owns the run is its only reader and writer).
constexpr int kK4bRecHead = (int)offsetof(SbrFrameDev, f_table_res);   // bytes of the record that stay in global memory
constexpr int kK4bRecTail = (int)sizeof(SbrFrameDev) - kK4bRecHead;
static_assert(kK4bRecHead % 16 == 0 && kK4bRecTail % 16 == 0 && kK4bRecTail <= 32 * 16, "one uint4 per lane moves the staged part");
struct __align__(16) K4bSmem {
  union {
    K4bGain g;                   // calculate_gain's working set
    float pw[38][kK4bPwCols];    // before that: |X|^2 of the generated band samples, [slot][m] (estimate_current_envelope)
  };
  uint8_t rec_tail[kK4bRecTail]; // SbrFrameDev from f_table_res on; `fp` below points kK4bRecHead bytes in front of it
  float eband[64];               // envelope energy per frequency band (bs_interpol_freq == 0)
  float bw[8];
  float ringG[5][64], ringQ[5][64];   // G_temp_prev / Q_temp_prev: the smoothing ring, by ring position
  uint8_t seq_i[192], seq_l[192];     // hf_assembly's slot order: for every envelope l, slots t_E[l] .. t_E[l + 1] - 1
};
static_assert(offsetof(K4bSmem, rec_tail) >= (size_t)kK4bRecHead && offsetof(K4bSmem, rec_tail) % 16 == 0, "fp stays inside the warp's block");
constexpr size_t k4b_smem_bytes() { return sizeof(K4bSmem) * kK4bWarps; }

__global__ void __launch_bounds__(32 * kK4bWarps, K4B_MIN_BLOCKS)
k4b_hf_kernel(const K4RunDev* __restrict__ runs, uint32_t n_runs, const RunFrameDev* __restrict__ run_frames,
              const SbrFrameDev* __restrict__ sframes, const float* __restrict__ core, SbrChanDev* __restrict__ chans,
              float* xg, SbrTablesDev T, K4Tile tile) {
  extern __shared__ __align__(16) uint8_t k4b_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t r = blockIdx.x * kK4bWarps + warp;
#if K4B_ALIGN
  // (a warp without a run, or whose run ends before this tile, keeps the CTA's barriers company)
  if (r >= n_runs || tile.lo >= runs[r].count) {
    for (uint32_t k = 0; k < tile.ft; ++k) K4B_SKIP_FRAME();
    return;
  }
#else
  if (r >= n_runs) return;
#endif
  K4bSmem& W = reinterpret_cast<K4bSmem*>(k4b_raw)[warp];
  const K4RunDev run = runs[r];
  if (tile.lo >= run.count) return;
  const uint32_t hi = min(run.count, tile.lo + tile.ft);
  SbrChanDev* st = chans + (size_t)run.stream_slot * kSbrChansPerStream + run.ch_slot;
  float* Xrun = xg + (size_t)r * tile.rows * kXgRow;

  // ---- recursive state in: the smoothing ring, chirp factors (lanes < 8), phases
  for (int i = lane; i < 5 * 64; i += 32) { (&W.ringG[0][0])[i] = (&st->G_temp_prev[0][0])[i]; (&W.ringQ[0][0])[i] = (&st->Q_temp_prev[0][0])[i]; }
  float* const E_curr = &st->E_curr[0][0];   // [l * 64 + m], (see SbrChanDev::E_curr)
  int ring_index = st->GQ_ringbuf_index;
  int index_noise_prev = st->index_noise_prev, psi_is_prev = st->psi_is_prev;
  float bw_prev = 0.f;
  int invf_prev = 0;
  if (lane < 8) { bw_prev = st->bwArray_prev[lane]; invf_prev = st->bs_invf_mode_prev[lane]; }

  uint32_t n_done = 0;   // processed frames of the tile so far
  int64_t last_it = -1;
  // fp-> reaches the staged fields; the three float tables in front of them must be read through `gfp`
  const SbrFrameDev* fp = reinterpret_cast<const SbrFrameDev*>(W.rec_tail - kK4bRecHead);
  // the frame records are fetched one frame ahead
  constexpr int kRecVec = kK4bRecTail / 16;
  uint4 rec = make_uint4(0u, 0u, 0u, 0u);
  auto fetch_record = [&](uint32_t it2) {
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const uint8_t*>(k4_frame(sframes, run, it2)) + kK4bRecHead);
    if (lane < kRecVec) rec = __ldg(src + lane);
  };
  fetch_record(tile.lo);
#if K4B_ALIGN
  const uint32_t it_end = tile.lo + tile.ft;   // the same trip count for every warp of the CTA
#else
  const uint32_t it_end = hi;
#endif
  for (uint32_t it = tile.lo; it < it_end; ++it) {
    if (K4B_ALIGN && it >= hi) { K4B_SKIP_FRAME(); continue; }
    __syncwarp();
    if (lane < kRecVec) reinterpret_cast<uint4*>(W.rec_tail)[lane] = rec;
    JAADB_ASSERT(reinterpret_cast<const uint8_t*>(fp) >= reinterpret_cast<const uint8_t*>(&W) && n_done <= tile.ft);
    const SbrFrameDev* gfp = k4_frame(sframes, run, it);
    if (it + 1 < hi) fetch_record(it + 1);
    __syncwarp();
    const int mode = fp->mode;
    if (fp->frame_status != 0 || mode == 0) { K4B_SKIP_FRAME(); continue; }
    float* X = Xrun + 32 * (size_t)n_done * kXgRow;   // row 0 of this frame's Xsbr
    ++n_done;
    last_it = it;
    if (mode != 2) { K4B_SKIP_FRAME(); continue; }
    K4B_PHASE();
    const int kx = fp->kx, M = fp->M, L_E = fp->L_E;
    // (borders beyond the matrix can only come from a damaged grid, where the reference dies with an index error; the
    // clamps keep the accesses inside the tile)
    const int first_slot = min((int)fp->t_E[0], 38), last_slot = min((int)fp->t_E[L_E], 38);
    bool grid_sorted = true;
    for (int l = 0; l < L_E; ++l) grid_sorted = grid_sorted && fp->t_E[l + 1] >= fp->t_E[l] && fp->t_E[l + 1] <= 38;
    // the frequency-band tables are strictly increasing from kx to at most kx + M (anything else takes the literal walks)
    bool res_regular;
    {
      bool bad = fp->N_low < 1 || fp->N_high < 1 || fp->f_table_res[0][0] != kx || fp->f_table_res[1][0] != kx ||
                 fp->f_table_res[0][fp->N_low] > kx + M || fp->f_table_res[1][fp->N_high] > kx + M || M > kK4bPwCols;
      for (int i = lane; i < 64; i += 32) {
        if (i < fp->N_low && fp->f_table_res[0][i + 1] <= fp->f_table_res[0][i]) bad = true;
        if (i < fp->N_high && fp->f_table_res[1][i + 1] <= fp->f_table_res[1][i]) bad = true;
      }
      res_regular = !__any_sync(0xFFFFFFFFu, bad);
    }
#ifdef K4B_FORCE_LITERAL   // test builds: always take the literal walks (the paths damaged band tables fall back to)
    const bool pw_ok = false;
#else
    const bool pw_ok = grid_sorted && res_regular;
#endif
    // calc_chirp_factors (HFGeneration.java:230-245): lane i < N_Q
    if (lane < 8) {
      float bw = 0.f;
      if (lane < fp->N_Q) {
        const int mode_i = fp->bs_invf_mode[lane];
        switch (mode_i) {
          case 1: bw = (invf_prev == 0) ? 0.6f : 0.75f; break;
          case 2: bw = 0.9f; break;
          case 3: bw = 0.98f; break;
          default: bw = (invf_prev == 1) ? 0.6f : 0.0f; break;
        }
        if (bw < bw_prev) bw = (bw * 0.75f) + (bw_prev * 0.25f);
        else bw = (bw * 0.90625f) + (bw_prev * 0.09375f);
        if (bw < 0.015625f) bw = 0.0f;
        if (bw >= 0.99609375f) bw = 0.99609375f;
        bw_prev = bw;
        invf_prev = mode_i;
      }
      W.bw[lane] = bw;
    }
    __syncwarp();
    // ---- HF generation: one lane per generated band (band x of the concatenated patches).  Loops run in blocks of four
    // slots whose loads are issued together (the next block's while the current one is worked on), rolled to stay small.
    // |X|^2 of every generated sample is left in W.pw for the envelope estimate.
    uint32_t not_generated = 0;   // bit pass: band lane + 32 * pass < M was not produced by any patch
#pragma unroll 1
    for (int x0 = 0; kx + x0 < 64; x0 += 32) {
      const int k = kx + x0 + lane;
      int i = 0, x = x0 + lane;
      while (i < fp->noPatches && x >= fp->patchNoSubbands[i]) { x -= fp->patchNoSubbands[i]; ++i; }
      if (k >= 64 || i >= fp->noPatches) {
        if (k < kx + M) not_generated |= 1u << (x0 >> 5);
        continue;
      }
      const int p = fp->patchStartSubband[i] + x;
      const float bw = W.bw[fp->table_map_k_to_g[k]];
      const float bw2 = bw * bw;
      const int offset = kSbrHfAdj;
      const float* src = X + 2 * p;     // the source band is never written by this kernel
      float* dst = X + 2 * k;
      auto ld_src = [&](int row) -> float2 { return __ldg(reinterpret_cast<const float2*>(src + min(row, 39) * kXgRow)); };
      float a0_r = 0, a1_r = 0, a0_i = 0, a1_i = 0;
      if (bw2 > 0) {
        // calc_prediction_coef / auto_correlation (:100-204), len = numTimeSlotsRate + 6
        float r01r = 0, r01i = 0, r02r = 0, r02i = 0, r11r = 0;
        float temp1_r, temp1_i, temp2_r, temp2_i, temp3_r, temp3_i, temp4_r, temp4_i, temp5_r, temp5_i;
        const float rel = 1.0f / (1 + 1e-6f);
        float2 v = ld_src(offset - 2);
        temp2_r = v.x; temp2_i = v.y;
        v = ld_src(offset - 1);
        temp3_r = v.x; temp3_i = v.y;
        temp4_r = temp2_r; temp4_i = temp2_i; temp5_r = temp3_r; temp5_i = temp3_i;
        temp1_r = 0; temp1_i = 0;
        float2 nxt[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) nxt[u] = ld_src(offset + u);
#pragma unroll 1
        for (int j0 = offset; j0 < kSbrSlots + 6 + offset; j0 += 4) {
          float2 cur[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { cur[u] = nxt[u]; nxt[u] = ld_src(j0 + 4 + u); }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            if (j0 + u < kSbrSlots + 6 + offset) {
              temp1_r = temp2_r; temp1_i = temp2_i;
              temp2_r = temp3_r; temp2_i = temp3_i;
              temp3_r = cur[u].x; temp3_i = cur[u].y;
              r01r += temp3_r * temp2_r + temp3_i * temp2_i;
              r01i += temp3_i * temp2_r - temp3_r * temp2_i;
              r02r += temp3_r * temp1_r + temp3_i * temp1_i;
              r02i += temp3_i * temp1_r - temp3_r * temp1_i;
              r11r += temp2_r * temp2_r + temp2_i * temp2_i;
            }
          }
        }
        const float r12r = r01r - (temp3_r * temp2_r + temp3_i * temp2_i) + (temp5_r * temp4_r + temp5_i * temp4_i);
        const float r12i = r01i - (temp3_i * temp2_r - temp3_r * temp2_i) + (temp5_i * temp4_r - temp5_r * temp4_i);
        const float r22r = r11r - (temp2_r * temp2_r + temp2_i * temp2_i) + (temp4_r * temp4_r + temp4_i * temp4_i);
        const float det = (r11r * r22r) - (rel * ((r12r * r12r) + (r12i * r12i)));
        float al0r, al0i, al1r, al1i;
        if (det == 0) { al1r = 0; al1i = 0; }
        else {
          const float tmp = 1.0f / det;
          al1r = ((r01r * r12r) - (r01i * r12i) - (r02r * r11r)) * tmp;
          al1i = ((r01i * r12r) + (r01r * r12i) - (r02i * r11r)) * tmp;
        }
        if (r11r == 0) { al0r = 0; al0i = 0; }
        else {
          const float tmp = 1.0f / r11r;
          al0r = -(r01r + (al1r * r12r) + (al1i * r12i)) * tmp;
          al0i = -(r01i + (al1i * r12r) - (al1r * r12i)) * tmp;
        }
        if (((al0r * al0r) + (al0i * al0i) >= 16.0f) || ((al1r * al1r) + (al1i * al1i) >= 16.0f)) { al0r = 0; al0i = 0; al1r = 0; al1i = 0; }
        a0_r = (al0r * bw); a1_r = (al1r * bw2); a0_i = (al0i * bw); a1_i = (al1i * bw2);
      }
      // patch the band (:60-97)
      const bool keep_pw = k - kx < kK4bPwCols;
      float2 t2 = ld_src(first_slot - 2 + offset), t3 = ld_src(first_slot - 1 + offset);
      float2 nxt[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) nxt[u] = ld_src(first_slot + u + offset);
#pragma unroll 1
      for (int l0 = first_slot; l0 < last_slot; l0 += 4) {
        float2 cur[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { cur[u] = nxt[u]; nxt[u] = ld_src(l0 + 4 + u + offset); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int l = l0 + u;
          if (l < last_slot) {
            const float2 t1 = t2;
            t2 = t3;
            t3 = cur[u];
            float o0, o1;
            if (bw2 > 0) {
              o0 = t3.x + ((a0_r * t2.x) - (a0_i * t2.y) + (a1_r * t1.x) - (a1_i * t1.y));
              o1 = t3.y + ((a0_i * t2.x) + (a0_r * t2.y) + (a1_i * t1.x) + (a1_r * t1.y));
            } else { o0 = t3.x; o1 = t3.y; }
            st2(dst + (l + offset) * kXgRow, o0, o1);
            if (keep_pw) W.pw[l][k - kx] = (o0 * o0) + (o1 * o1);
          }
        }
      }
    }
    __syncwarp();

    // ---- HF adjustment (HFAdjustment.java)
    K4B_PHASE_MINOR();
    // estimate_current_envelope (:78-131).  With a sorted grid every sample it reads was produced just above and its
    // |X|^2 sits in W.pw; the sums below add the same terms in the same order as the reference.
    if (pw_ok) {
      if (__any_sync(0xFFFFFFFFu, not_generated != 0)) {
        // bands inside [kx, kx + M) that no patch covers keep whatever the matrix holds there
        for (int m = lane; m < M; m += 32)
          if ((not_generated >> (m >> 5)) & 1u)
            for (int sl = first_slot; sl < last_slot; ++sl) {
              const float2 v = ldg2(X + (size_t)(sl + kSbrHfAdj) * kXgRow + 2 * (m + kx));
              W.pw[sl][m] = (v.x * v.x) + (v.y * v.y);
            }
        __syncwarp();
      }
      for (int l = 0; l < L_E; l++) {
        const int l_i = fp->t_E[l], u_i = fp->t_E[l + 1];
        if (fp->interpol_freq) {
          float div = (float)(u_i - l_i);
          if (div == 0) div = 1;
          for (int m = lane; m < M; m += 32) {
            float nrg = 0;
            for (int sl = l_i; sl < u_i; ++sl) nrg += W.pw[sl][m];
            E_curr[(l) * 64 + (m)] = nrg / div;
          }
        } else {
          // one sum per frequency band of the envelope's resolution, taken by the lane of the band's first m
          const int res = fp->f[l], nb = res ? fp->N_high : fp->N_low;
          int pb[2];
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int m = lane + 32 * q;
            int pp = 0;
            while (pp + 1 < nb && fp->f_table_res[res][pp + 1] <= m + kx) ++pp;
            pb[q] = pp;
          }
          const int up0 = __shfl_up_sync(0xFFFFFFFFu, pb[0], 1), up1 = __shfl_up_sync(0xFFFFFFFFu, pb[1], 1);
          const int last0 = __shfl_sync(0xFFFFFFFFu, pb[0], 31);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int m = lane + 32 * q;
            const int before = q == 0 ? (lane == 0 ? -1 : up0) : (lane == 0 ? last0 : up1);   // the band of m - 1
            if (m < M && before != pb[q]) {
              const int k_l = fp->f_table_res[res][pb[q]], k_h = fp->f_table_res[res][pb[q] + 1];
              float div = (float)((u_i - l_i) * (k_h - k_l));
              if (div == 0) div = 1;
              float nrg = 0;
              for (int sl = l_i; sl < u_i; ++sl)
                for (int j = k_l; j < k_h; j++) nrg += W.pw[sl][j - kx];
              W.eband[pb[q]] = nrg / div;
            }
          }
          __syncwarp();
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int m = lane + 32 * q;
            if (m < M) E_curr[(l) * 64 + (m)] = W.eband[pb[q]];
          }
          __syncwarp();
        }
      }
    } else {
      for (int m = lane; m < M; m += 32) {
        for (int l = 0; l < L_E; l++) {
          const int l_i = fp->t_E[l], u_i = min((int)fp->t_E[l + 1], 38);
          float nrg = 0, div;
          if (fp->interpol_freq) {
            div = (float)(u_i - l_i);
            if (div == 0) div = 1;
            const float* col = X + 2 * (m + kx);
            for (int i = l_i + kSbrHfAdj; i < u_i + kSbrHfAdj; i++) {
              const float2 v = ldg2(col + i * kXgRow);
              nrg += (v.x * v.x) + (v.y * v.y);
            }
          } else {
            // the band of the envelope's resolution that holds k = m + kx
            const int res = fp->f[l], nb = res ? fp->N_high : fp->N_low;
            int pp = 0;
            while (pp + 1 < nb && fp->f_table_res[res][pp + 1] <= m + kx) ++pp;
            const int k_l = fp->f_table_res[res][pp], k_h = fp->f_table_res[res][pp + 1];
            div = (float)((u_i - l_i) * (k_h - k_l));
            if (div == 0) div = 1;
            for (int i = l_i + kSbrHfAdj; i < u_i + kSbrHfAdj; i++)
              for (int j = k_l; j < k_h; j++) {
                const float2 v = ldg2(X + i * kXgRow + 2 * j);
                nrg += (v.x * v.x) + (v.y * v.y);
              }
          }
          E_curr[(l) * 64 + (m)] = nrg / div;
        }
      }
    }
    __syncwarp();
    // `new HFAdjustment()` per call: the boost arrays start from zero; the limiter table does not always reach M
    // (FBT.limiter_frequency_table sorts a shrinking prefix) and bands it leaves out keep gain 0.
    for (int i = lane; i < 3 * kSbrMaxLE * 64; i += 32) (&W.g.G[0][0])[i] = 0.f;
    __syncwarp();

#if K4B_PREFETCH
    // The tile's matrices are far larger than L2, and the throughput kernels of the other half of the batch stream through
    // it all the time: what this warp generated a phase ago is back in DRAM when the assembly walks it (long_scoreboard on
    // that load: 13 % of the stall samples), and so is the next frame's low band, which the analysis kernel wrote before
    // this kernel started.  Both are asked into L2 one phase ahead of their use.
    if (K4B_PREFETCH & 1) {
      const int r0 = first_slot + kSbrHfAdj, nr = last_slot - first_slot;
      const char* base = reinterpret_cast<const char*>(X + (size_t)r0 * kXgRow) + ((kx * 8) & ~127);
      const int nl = (((kx + M) * 8 + 127) >> 7) - ((kx * 8) >> 7);
      JAADB_ASSERT(nr >= 0 && r0 + nr <= 40 && nl >= 0 && ((kx * 8) & ~127) + nl * 128 <= (int)(kXgRow * sizeof(float)));
      for (int i = lane; i < nr * nl; i += 32) {
        const int row = i / nl, ln = i - row * nl;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)row * kXgRow * sizeof(float) + ln * 128));
      }
    }
#endif
#if K4B_ALIGN == 2
    K4B_PHASE_MINOR();
#else
    K4B_PHASE();
#endif
    // calculate_gain (:242-415)
    const float EPS = 1e-12f;
    const int l_A = fp->l_A;
    const bool flag_prev = fp->add_harmonic_flag_prev != 0;
    const int N_L = fp->N_L;
    float limg;
    switch (fp->limiter_gains) { case 0: limg = 0.5f; break; case 1: limg = 1.0f; break; case 2: limg = 2.0f; break; default: limg = 1e10f; break; }
    auto get_S_mapped = [&](int l, int res, int current_band) -> int {   // :46-76
      if (res == SBR_HI_RES) {
        if ((l >= l_A) || (fp->bs_add_harmonic_prev[current_band] != 0 && flag_prev)) return fp->bs_add_harmonic[current_band];
      } else {
        const int odd = (fp->N_high & 1) ? 1 : 0;
        const int lb = 2 * current_band - odd, ub = 2 * (current_band + 1) - odd;
        for (int b = max(lb, 0); b < ub && b < 64; b++)
          if ((l >= l_A) || (fp->bs_add_harmonic_prev[b] != 0 && flag_prev)) { if (fp->bs_add_harmonic[b] == 1) return 1; }
      }
      return 0;
    };
    auto t_noise_band = [&](int l) -> int {   // current_t_noise_band advances once per envelope whose end passes t_Q
      int tb = 0;
      for (int ll = 0; ll <= l; ++ll)
        if (fp->t_E[ll + 1] > fp->t_Q[tb + 1]) tb++;
      return tb;
    };
    // The reference walks m with running band counters; when every band table is strictly increasing and starts where
    // it should, the counters equal plain lookups and the work splits over (envelope, band).  Anything else takes the
    // literal walk below.
    bool regular;
    {
      bool bad = N_L > kK4bMaxNL || N_L < 1 || !res_regular;
      for (int i = lane; i < 64; i += 32) {
        if (i < fp->N_Q && i < 7 && fp->f_table_noise[i + 1] <= fp->f_table_noise[i]) bad = true;
        if (i < N_L && i < 63 && fp->f_table_lim[i + 1] <= fp->f_table_lim[i]) bad = true;
      }
      if (fp->f_table_noise[0] != kx || fp->f_table_lim[0] != 0) bad = true;
      if (fp->N_Q > 7 || fp->N_Q < 1) bad = true;
      if (N_L >= 1 && N_L <= kK4bMaxNL && fp->f_table_lim[N_L] > M) bad = true;
      regular = !__any_sync(0xFFFFFFFFu, bad);
#ifdef K4B_FORCE_LITERAL
      regular = false;
#endif
    }
    if (regular) {
      const int mcov = fp->f_table_lim[N_L];
      for (int m = lane; m < mcov; m += 32) {
        const int k = m + kx;
        int p = 0;
        while (p + 1 < fp->N_low && fp->f_table_res[0][p + 1] <= k) ++p;
        W.g.rb[0][m] = (uint8_t)p;
        p = 0;
        while (p + 1 < fp->N_high && fp->f_table_res[1][p + 1] <= k) ++p;
        W.g.rb[1][m] = (uint8_t)p;
        p = 0;
        while (p + 1 < fp->N_Q && fp->f_table_noise[p + 1] <= k) ++p;
        W.g.nb[m] = (uint8_t)p;
        p = 0;
        while (p + 1 < N_L && fp->f_table_lim[p + 1] <= m) ++p;
        W.g.lb[m] = (uint8_t)p;
      }
      __syncwarp();
      // limiter-band sums, in band order as the reference adds them
      for (int task = lane; task < L_E * N_L; task += 32) {
        const int l = task / N_L, kb = task - l * N_L, res = fp->f[l];
        float acc1 = 0, acc2 = 0;
        for (int m = fp->f_table_lim[kb]; m < fp->f_table_lim[kb + 1]; m++) {
          acc1 += __ldg(&gfp->E_orig[l][W.g.rb[res][m]]);
          acc2 += E_curr[(l) * 64 + (m)];
        }
        float G_max = ((EPS + acc1) / (EPS + acc2)) * limg;
        G_max = fminf(G_max, 1e10f);
        W.g.gmax[l][kb] = G_max;
        W.g.acc1[l][kb] = acc1;
      }
      __syncwarp();
      for (int m = lane; m < mcov; m += 32) {
        const int k = m + kx;
        const int hb = W.g.rb[1][m];
        const bool centre = k == ((fp->f_table_res[SBR_HI_RES][hb + 1] + fp->f_table_res[SBR_HI_RES][hb]) >> 1);
        for (int l = 0; l < L_E; ++l) {
          const int res = fp->f[l];
          const int rb = W.g.rb[res][m];
          const int S_mapped = get_S_mapped(l, res, rb);
          int S_index_mapped = 0;
          if (((l >= l_A) || (fp->bs_add_harmonic_prev[hb] != 0 && flag_prev)) && centre) S_index_mapped = fp->bs_add_harmonic[hb];
          const int tb = t_noise_band(l);
          const float delta = (l == l_A || l == fp->prevEnvIsShort) ? 0.f : 1.f;
          const float Q_div = __ldg(&gfp->Q_div[tb][W.g.nb[m]]);
          const float Q_div2 = __ldg(&gfp->Q_div2[tb][W.g.nb[m]]);
          const float E_o = __ldg(&gfp->E_orig[l][rb]);
          const float Q_M = E_o * Q_div2;
          const float S_M = (S_index_mapped == 0) ? 0.f : E_o * Q_div;
          float G = E_o / (1.0f + E_curr[(l) * 64 + (m)]);
          if ((S_mapped == 0) && (delta == 1)) G *= Q_div;
          else if (S_mapped == 1) G *= Q_div2;
          const float G_max = W.g.gmax[l][W.g.lb[m]];
          float Q_M_lim, G_lim;
          if (G_max > G) { Q_M_lim = Q_M; G_lim = G; }
          else { Q_M_lim = Q_M * G_max / G; G_lim = G_max; }
          W.g.G[l][m] = G_lim;
          W.g.Q[l][m] = Q_M_lim;
          W.g.S[l][m] = S_M;
          W.g.sflag[l][m] = (uint8_t)(S_index_mapped != 0);
        }
      }
      __syncwarp();
      for (int task = lane; task < L_E * N_L; task += 32) {
        const int l = task / N_L, kb = task - l * N_L;
        float den = 0;
        for (int m = fp->f_table_lim[kb]; m < fp->f_table_lim[kb + 1]; m++) {
          const bool sf = W.g.sflag[l][m] != 0;
          if (sf) den += W.g.S[l][m];
          den += E_curr[(l) * 64 + (m)] * W.g.G[l][m];
          if (!sf && (l != l_A)) den += W.g.Q[l][m];
        }
        float G_boost = (W.g.acc1[l][kb] + EPS) / (den + EPS);
        G_boost = fminf(G_boost, 2.51188643f);
        W.g.boost[l][kb] = G_boost;
      }
      __syncwarp();
      for (int m = lane; m < mcov; m += 32) {
        for (int l = 0; l < L_E; ++l) {
          const float G_boost = W.g.boost[l][W.g.lb[m]];
          // (float) Math.sqrt(float product): the double square root of a binary32 value, rounded to binary32, is the
          // correctly rounded binary32 square root
          W.g.G[l][m] = __fsqrt_rn(W.g.G[l][m] * G_boost);
          W.g.Q[l][m] = __fsqrt_rn(W.g.Q[l][m] * G_boost);
          const float sm = W.g.S[l][m];
          W.g.S[l][m] = (sm != 0) ? __fsqrt_rn(sm * G_boost) : 0.f;
        }
      }
    } else if (lane < L_E) {
      // the reference's own loop, one lane per envelope
      const int l = lane;
      const int res = fp->f[l];
      const int current_t_noise_band = t_noise_band(l);
      int current_f_noise_band = 0, current_res_band = 0, current_res_band2 = 0, current_hi_res_band = 0;
      const float delta = (l == l_A || l == fp->prevEnvIsShort) ? 0.f : 1.f;
      int S_mapped = get_S_mapped(l, res, current_res_band2);
      for (int k = 0; k < N_L; k++) {
        float den = 0, acc1 = 0, acc2 = 0;
        const int ml1 = fp->f_table_lim[k], ml2 = fp->f_table_lim[k + 1];
        for (int m = ml1; m < ml2; m++) {
          if ((m + kx) == fp->f_table_res[res][current_res_band + 1]) current_res_band++;
          acc1 += __ldg(&gfp->E_orig[l][current_res_band]);
          acc2 += E_curr[(l) * 64 + (m)];
        }
        float G_max = ((EPS + acc1) / (EPS + acc2)) * limg;
        G_max = fminf(G_max, 1e10f);
        for (int m = ml1; m < ml2; m++) {
          if ((m + kx) == fp->f_table_noise[current_f_noise_band + 1]) current_f_noise_band++;
          if ((m + kx) == fp->f_table_res[res][current_res_band2 + 1]) {
            current_res_band2++;
            S_mapped = get_S_mapped(l, res, current_res_band2);
          }
          if ((m + kx) == fp->f_table_res[SBR_HI_RES][current_hi_res_band + 1]) current_hi_res_band++;
          int S_index_mapped = 0;
          if ((l >= l_A) || (fp->bs_add_harmonic_prev[current_hi_res_band] != 0 && flag_prev)) {
            if ((m + kx) == (fp->f_table_res[SBR_HI_RES][current_hi_res_band + 1] + fp->f_table_res[SBR_HI_RES][current_hi_res_band]) >> 1)
              S_index_mapped = fp->bs_add_harmonic[current_hi_res_band];
          }
          const float Q_div = __ldg(&gfp->Q_div[current_t_noise_band][current_f_noise_band]);
          const float Q_div2 = __ldg(&gfp->Q_div2[current_t_noise_band][current_f_noise_band]);
          const float E_o = __ldg(&gfp->E_orig[l][current_res_band2]);
          const float Q_M = E_o * Q_div2;
          float S_M;
          if (S_index_mapped == 0) S_M = 0;
          else { S_M = E_o * Q_div; den += S_M; }
          float G = E_o / (1.0f + E_curr[(l) * 64 + (m)]);
          if ((S_mapped == 0) && (delta == 1)) G *= Q_div;
          else if (S_mapped == 1) G *= Q_div2;
          float Q_M_lim, G_lim;
          if (G_max > G) { Q_M_lim = Q_M; G_lim = G; }
          else { Q_M_lim = Q_M * G_max / G; G_lim = G_max; }
          den += E_curr[(l) * 64 + (m)] * G_lim;
          if ((S_index_mapped == 0) && (l != l_A)) den += Q_M_lim;
          W.g.G[l][m] = G_lim;
          W.g.Q[l][m] = Q_M_lim;
          W.g.S[l][m] = S_M;
        }
        float G_boost = (acc1 + EPS) / (den + EPS);
        G_boost = fminf(G_boost, 2.51188643f);
        for (int m = ml1; m < ml2; m++) {
          W.g.G[l][m] = __fsqrt_rn(W.g.G[l][m] * G_boost);
          W.g.Q[l][m] = __fsqrt_rn(W.g.Q[l][m] * G_boost);
          const float sm = W.g.S[l][m];
          W.g.S[l][m] = (sm != 0) ? __fsqrt_rn(sm * G_boost) : 0.f;
        }
      }
    }
    __syncwarp();

#if K4B_PREFETCH
    if ((K4B_PREFETCH & 2) && n_done < tile.ft) {   // the next frame's 32 new rows, bands below kx
      const char* nx = reinterpret_cast<const char*>(X + (size_t)(32 + kSbrHfGen) * kXgRow);
      const int nl = min((kx * 8 + 127) >> 7, 4);
      for (int i = lane; i < 32 * nl; i += 32) {
        const int row = i / nl, ln = i - row * nl;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(nx + (size_t)row * kXgRow * sizeof(float) + ln * 128));
      }
    }
#endif
    K4B_PHASE();
    // hf_assembly (:133-240): lane m walks the slots of its band.  The reference's 5-entry ring (written at
    // GQ_ringbuf_index, read oldest to newest) is kept age-ordered in registers while a band is walked: A[0] = oldest ...
    // A[4] = newest, and goes back to its ring positions afterwards.
    {
      const int fIndexNoise = fp->reset ? 0 : index_noise_prev;
      const int ri0 = fp->reset ? 4 : ring_index;
      int slots_total = 0;
      // the slot order of the reference's loops, laid out once per frame
      int n_seq = 0;
      for (int l = 0; l < L_E; ++l) {
        const int b0 = fp->t_E[l], b1 = min((int)fp->t_E[l + 1], 38);
        for (int i = b0 + lane; i < b1; i += 32) { W.seq_i[n_seq + i - b0] = (uint8_t)i; W.seq_l[n_seq + i - b0] = (uint8_t)l; }
        n_seq += max(0, b1 - b0);
      }
      for (int i = n_seq + lane; i < 192; i += 32) W.seq_i[i] = 0;   // (prefetches past the end read row tHFAdj)
      __syncwarp();
#pragma unroll 1
      for (int m = lane; m < M; m += 32) {
        // System.arraycopy(.., 0, .., 0, sbr.M) on a reset: positions 0..3 take the first envelope's values (bands < M only)
        if (fp->reset) {
          const float g0 = W.g.G[0][m], q0 = W.g.Q[0][m];
#pragma unroll
          for (int n = 0; n < 4; ++n) { W.ringG[n][m] = g0; W.ringQ[n][m] = q0; }
        }
        float Ag[5], Aq[5];
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          int pos = ri0 + j;
          if (pos >= 5) pos -= 5;
          Ag[j] = W.ringG[pos][m];
          Aq[j] = W.ringQ[pos][m];
        }
        float* col = X + 2 * (m + kx);
        const int rev = (((m + kx) & 1) != 0 ? -1 : 1);
        int l_cached = -1;
        float g_new = 0, q_new = 0, s_m = 0, rs_m = 0;
        bool no_noise = false;
        int h_SL = 0;
        // four slots at a time; the next four are loaded while these are worked on
        float2 nxt[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) nxt[u] = ldg2(col + (size_t)(W.seq_i[min(u, 191)] + kSbrHfAdj) * kXgRow);
#pragma unroll 1
        for (int n0 = 0; n0 < n_seq; n0 += 4) {
          float2 cur[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) { cur[u] = nxt[u]; nxt[u] = ldg2(col + (size_t)(W.seq_i[min(n0 + 4 + u, 191)] + kSbrHfAdj) * kXgRow); }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int n = n0 + u;
            if (n < n_seq) {
              const int i = W.seq_i[n], l = W.seq_l[n];
              // (a damaged grid may visit a slot twice within the prefetch distance: read it when it is due)
              if (!grid_sorted) cur[u] = ldg2(col + (size_t)(i + kSbrHfAdj) * kXgRow);
              if (l != l_cached) {
                l_cached = l;
                no_noise = (l == fp->l_A || l == fp->prevEnvIsShort);
                h_SL = (fp->smoothing_mode || no_noise) ? 0 : 4;
                g_new = W.g.G[l][m]; q_new = W.g.Q[l][m]; s_m = W.g.S[l][m];
                rs_m = (float)rev * s_m;
              }
              // the slot's entry: the oldest leaves, the envelope's value comes in as the newest
              Ag[0] = Ag[1]; Ag[1] = Ag[2]; Ag[2] = Ag[3]; Ag[3] = Ag[4]; Ag[4] = g_new;
              Aq[0] = Aq[1]; Aq[1] = Aq[2]; Aq[2] = Aq[3]; Aq[3] = Aq[4]; Aq[4] = q_new;
              float G_filt = 0, Q_filt = 0;
              if (h_SL != 0) {
#pragma unroll
                for (int z = 0; z <= 4; z++) {
                  const float h = z == 0 ? 0.03183050093751f : z == 1 ? 0.11516383427084f : z == 2 ? 0.21816949906249f
                                : z == 3 ? 0.30150283239582f : 0.33333333333333f;
                  G_filt += (Ag[z] * h);
                  Q_filt += (Aq[z] * h);
                }
              } else { G_filt = g_new; Q_filt = q_new; }
              Q_filt = (s_m != 0 || no_noise) ? 0 : Q_filt;
              const int ni = (fIndexNoise + n * M + m + 1) & 511;
              const float2 nz = __ldg(reinterpret_cast<const float2*>(T.noise_table) + ni);
              float x0 = G_filt * cur[u].x + (Q_filt * nz.x);
              float x1 = G_filt * cur[u].y + (Q_filt * nz.y);
              const int fs = (psi_is_prev + n) & 3;   // fIndexSine
              x0 += s_m * (fs == 0 ? 1.f : (fs == 2 ? -1.f : 0.f));
              x1 += rs_m * (fs == 1 ? 1.f : (fs == 3 ? -1.f : 0.f));
              st2(col + (size_t)(i + kSbrHfAdj) * kXgRow, x0, x1);
            }
          }
        }
        const int slots_done = n_seq;
        // back to ring positions: after slots_done writes the next write position is ri0 + slots_done
        int pos = (ri0 + slots_done) % 5;
#pragma unroll
        for (int j = 0; j < 5; ++j) {
          W.ringG[pos][m] = Ag[j];
          W.ringQ[pos][m] = Aq[j];
          if (++pos >= 5) pos = 0;
        }
      }
      slots_total = n_seq;
      ring_index = (ri0 + slots_total) % 5;
      index_noise_prev = (fIndexNoise + slots_total * M) & 511;
      psi_is_prev = (psi_is_prev + slots_total) & 3;
    }
  }
  __syncwarp();

  // ---- recursive state out
  for (int i = lane; i < 5 * 64; i += 32) { (&st->G_temp_prev[0][0])[i] = (&W.ringG[0][0])[i]; (&st->Q_temp_prev[0][0])[i] = (&W.ringQ[0][0])[i]; }
  if (lane < 8) { st->bwArray_prev[lane] = bw_prev; st->bs_invf_mode_prev[lane] = (uint8_t)invf_prev; }
  if (lane == 0) { st->GQ_ringbuf_index = ring_index; st->index_noise_prev = index_noise_prev; st->psi_is_prev = psi_is_prev; }
  if (last_it >= 0) {
    // what the next tile's analysis starts from: the last 288 core samples and rows 32..39 of the last processed frame
    const float* cs = core + ((size_t)run_frames[run.first + last_it].ics_base + run.ch_slot) * 1024 + 736;
    for (int i = lane; i < 288; i += 32) st->ana_hist[i] = cs[i];
    const float4* s4 = reinterpret_cast<const float4*>(Xrun + 32 * (size_t)n_done * kXgRow);
    float4* d4 = reinterpret_cast<float4*>(&st->xsbr[0][0][0]);
    for (int i = lane; i < kSbrHfGen * 32; i += 32) d4[i] = __ldcg(s4 + i);
  }
}

// ---- K4c: 64-band QMF synthesis (sbr/SynthesisFilterbank64.java:9-79) + PCM pack (S/SampleBuffer.java:168-209)
// One CTA takes kK4cG consecutive frames of one channel (bank): their 32 * G time slots plus the nine slots the first of
// them inherits, one slot per thread for the two DCT-IVs, so the re-computation of inherited slots costs one partial warp
// per G frames.  A slot's 64 complex QMF samples are staged in shared memory as [Re X(0..63) | Im X(63..0)] and replaced
// in place by the slot's 128-entry v-vector.
constexpr int kK4cG = 3;
constexpr int kK4cThreads = 32 * (kK4cG + 1);
constexpr int kK4cRows = 9 + 32 * kK4cG;   // the nine inherited slots, then the frames' slots
constexpr int kK4cFloats = kK4cRows * kVbStride + 3;
static_assert(kK4cFloats % 4 == 0, "16-byte aligned regions");

struct K4cFrame {        // what the CTA knows about one frame of its group
  uint8_t* dst;          // the frame's PCM
  const float* cs;       // core PCM of the channel (SBR.upsample)
  const float* src;      // QMF matrix source: row of slot 0 (xg: row 2 of the frame; xps: row 0)
  uint32_t it;
  int16_t row0;          // first v row of the frame (synthesised frames)
  uint8_t kind;          // 0: nothing to write, 1: SBR.upsample, 2: synthesise
  uint8_t dup, pair_store, from_ps, last_in_tile, mode;
  uint8_t lim_lo, lim_hi, first_slot;
};
constexpr size_t k4c_smem_bytes() { return sizeof(float) * kK4cFloats + sizeof(K4cFrame) * kK4cG + 16; }

// The two DCT-IVs of one slot (SynthesisFilterbank64.java:27-60), in place.  ONE copy of the unrolled DCT serves both (the
// kernel is instruction-fetch bound otherwise): pass 0 transforms the real parts and parks (o1r[n], o1i[31 - n]) in
// cells (2n, 2n + 1), which pass 1 consumes exactly when it writes v[2n], v[2n + 1] there.
__device__ __forceinline__ void sbr_synth_slot(float* __restrict__ row) {
  const float scale = 1.f / 64.f;
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    const float* b = row + 64 * pass;
    float in_r[32], in_i[32], o_r[32], o_i[32];
    in_i[31] = scale * b[1];
    in_r[0] = scale * b[0];
#pragma unroll
    for (int k = 1; k < 31; k++) { in_i[31 - k] = scale * b[2 * k + 1]; in_r[k] = scale * b[2 * k]; }
    in_i[0] = scale * b[63];
    in_r[31] = scale * b[62];
    sbr_dct4_kernel(in_r, in_i, o_r, o_i);
    if (pass == 0) {
#pragma unroll
      for (int n = 0; n < 32; n++) { row[2 * n] = o_r[n]; row[2 * n + 1] = o_i[31 - n]; }
    } else {
#pragma unroll
      for (int n = 0; n < 32; n++) {
        const float o1r = row[2 * n], o1i = row[2 * n + 1];
        row[2 * n] = o_r[n] - o1r;
        row[127 - 2 * n] = o_r[n] + o1r;
        row[2 * n + 1] = o_i[31 - n] + o1i;
        row[127 - (2 * n + 1)] = o_i[31 - n] - o1i;
      }
    }
  }
}

// Down-sampled bank (sbr/SynthesisFilterbank32.java:57-77): the staged row holds the pre-twiddled, scaled x1[0..31] | x2[0..31];
// DCT4_32 / DST4_32 are the reference's operation lists (generated/jaad_dct32.h), the row is replaced by the slot's 64-entry
// v-vector.
#define JD_ADD(d, a, b) d = a + b;
#define JD_SUB(d, a, b) d = a - b;
#define JD_MUL(d, c, a) d = (c * a);
__device__ __forceinline__ void sbr_dct4_32(float (&x)[32]) {
  JAAD_DCT4_32_TEMPS
  JAAD_DCT4_32_OPS
}
__device__ __forceinline__ void sbr_dst4_32(float (&x)[32]) {
  JAAD_DST4_32_TEMPS
  JAAD_DST4_32_OPS
}
#undef JD_ADD
#undef JD_SUB
#undef JD_MUL
__device__ __forceinline__ void sbr_synth_slot32(float* __restrict__ row) {
  float x[32];
#pragma unroll
  for (int k = 0; k < 32; ++k) x[k] = row[k];
  sbr_dct4_32(x);
#pragma unroll
  for (int k = 0; k < 32; ++k) { const float t = row[32 + k]; row[32 + k] = x[k]; x[k] = t; }   // park x1', fetch x2
  sbr_dst4_32(x);
  float x1[32];
#pragma unroll
  for (int n = 0; n < 32; ++n) x1[n] = row[32 + n];
#pragma unroll
  for (int n = 0; n < 32; ++n) {   // :69-72
    row[n] = -x1[n] + x[n];
    row[63 - n] = x1[n] + x[n];
  }
}

// X[l][k] = Xsbr[l + tHFAdj][k] below kx + M of the slot's frame, 0 above (Channel.process_channel, SBR.java:604-645)
__device__ __forceinline__ int k4_x_limit(const SbrFrameDev* fp, int mode, int l) {
  if (mode == 2) return (l < fp->t_E[0]) ? (fp->kx_prev + fp->M_prev) : (fp->kx + fp->M);
  return 32;
}

// PS: the runs are the SCEs of SBR+PS streams, blockIdx.y is the synthesis bank (SBR1.processPS, :121-122): 0 = left, fed
// by K5's left matrix in frames that carry ps_data and by Xsbr (output duplicated) otherwise; 1 = right, which only exists
// -- and only moves its history -- in frames that carry ps_data.
template <int PCM_FORMAT, bool PS, bool DS>
__global__ void __launch_bounds__(kK4cThreads)
k4c_synthesis_kernel(const K4RunDev* __restrict__ runs, uint32_t run0, const RunFrameDev* __restrict__ run_frames,
                     const SbrFrameDev* __restrict__ sframes, const float* __restrict__ core, SbrChanDev* __restrict__ chans,
                     const float* __restrict__ xg, uint8_t* __restrict__ pcm, const uint64_t* __restrict__ pcm_off,
                     uint32_t* __restrict__ pcm_bytes_out, SbrTablesDev T, K4Tile tile,
                     const PsFrameDev* __restrict__ ps_frames, PsChanDev* __restrict__ ps_chans, const float* __restrict__ xps) {
  extern __shared__ __align__(16) float k4c_smem[];
  float* vb = k4c_smem;   // [kK4cRows][kVbStride]: rows 0..8 = the inherited slots, then 32 rows per synthesised frame
  K4cFrame* info = reinterpret_cast<K4cFrame*>(k4c_smem + kK4cFloats);
  __shared__ int s_nsyn, s_halo;   // synthesised frames of the group; 0: none, 1: v-vectors from the carried state, 2: recompute
  __shared__ const float* s_halo_src;
  __shared__ int s_halo_lim[9];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const uint32_t n_groups = (tile.ft + kK4cG - 1) / kK4cG;
  const uint32_t rl = blockIdx.x / n_groups, r = run0 + rl, it0 = tile.lo + (blockIdx.x % n_groups) * kK4cG;
  const int bank = PS ? (int)blockIdx.y : 0;
  const K4RunDev run = runs[r];
  if (it0 >= run.count || (run.ds != 0) != DS) return;   // the two banks are two instantiations over the same grid
  constexpr int kOutLen = DS ? 1024 : 2048;   // samples per frame and channel
  constexpr int kVLen = DS ? 64 : 128;        // entries of one v-vector
  const uint32_t it_end = min(min(run.count, tile.lo + tile.ft), it0 + kK4cG);
  const int nfr = (int)(it_end - it0);
  const int n_out = run.n_out;
  const int out_ch = PS ? bank : run.out_ch;
  SbrChanDev* st = chans + (size_t)run.stream_slot * kSbrChansPerStream + run.ch_slot;
  PsChanDev* pst = PS ? ps_chans + run.stream_slot : nullptr;
  const int sel = (PS && bank == 1) ? pst->v_sel : st->v_sel;
  float* v_in = (PS && bank == 1) ? &pst->syn_v_right[sel][0][0] : &st->syn_v[sel][0][0];
  float* v_out = (PS && bank == 1) ? &pst->syn_v_right[sel ^ 1][0][0] : &st->syn_v[sel ^ 1][0][0];
  // K5's matrices of frame `fr` of this run: [left, right][32][kXgRow]
  auto ps_matrix = [&](uint32_t fr) -> const float* { return xps + (((size_t)rl * tile.ft + (fr - tile.lo)) * 2 + bank) * 32 * kXgRow; };
  const uint32_t ord_lo = k4_frame(sframes, run, tile.lo)->ord;

  // ---- what to do with each frame of the group (thread j), then the row plan (thread 0)
  if (t < nfr) {
    const uint32_t it = it0 + t;
    const SbrFrameDev* fp = k4_frame(sframes, run, it);
    const RunFrameDev rf = run_frames[run.first + it];
    const int mode = fp->mode;
    const bool ok = fp->frame_status == 0;
    const bool use_ps = PS && ok && mode != 0 && ps_frames[run.ps_base + it].use_ps != 0;
    K4cFrame f;
    f.dst = pcm + pcm_off[rf.frame];
    f.cs = core + ((size_t)rf.ics_base + run.ch_slot) * 1024;
    f.it = it;
    f.mode = (uint8_t)mode;
    f.row0 = 0;
    f.dup = (run.dup != 0 && !use_ps) ? 1 : 0;   // mono element: SBR1.process copies the channel unless PS makes the second one
    f.pair_store = (f.dup && out_ch == 0 && n_out == 2 && (reinterpret_cast<uintptr_t>(f.dst) & 3u) == 0) ? 1 : 0;
    f.from_ps = use_ps ? 1 : 0;
    f.kind = 0;
    f.last_in_tile = 0;
    f.lim_lo = f.lim_hi = 64;
    f.first_slot = 0;
    f.src = nullptr;
    if (PS && bank == 1) {
      if (use_ps) f.kind = 2;
    } else {
      if (ok) f.kind = mode == 0 ? 1 : 2;
      if (out_ch == 0) pcm_bytes_out[rf.frame] = ok ? (uint32_t)(kOutLen * n_out * (PCM_FORMAT == 2 ? 4 : 2)) : 0u;
    }
    if (f.kind == 2) {
      const uint32_t fwd = (PS && bank == 1) ? fp->fwd_ps : fp->fwd;
      f.last_in_tile = (fwd == 0 || it + fwd >= tile.lo + tile.ft) ? 1 : 0;
      if (use_ps) f.src = ps_matrix(it);
      else {
        f.src = xg + ((size_t)r * tile.rows + 32 * (size_t)(fp->ord - ord_lo) + kSbrHfAdj) * kXgRow;
        f.lim_lo = (uint8_t)k4_x_limit(fp, mode, 0);
        f.lim_hi = (uint8_t)k4_x_limit(fp, mode, 31);
        f.first_slot = mode == 2 ? fp->t_E[0] : 0;
      }
    }
    info[t] = f;
  }
  __syncthreads();
  // a frame that failed yields no PCM: its slot of the output is zero-filled (first output channel's CTA, left bank)
  if (!(PS && bank == 1) && out_ch == 0) {
    for (int j = 0; j < nfr; ++j) {
      if (info[j].kind != 0) continue;
      uint32_t* d = reinterpret_cast<uint32_t*>(info[j].dst);
      for (int i = t; i < kOutLen * n_out * (PCM_FORMAT == 2 ? 4 : 2) / 4; i += kK4cThreads) d[i] = 0u;
    }
  }
  if (t == 0) {
    int n = 0, first = -1;
    for (int j = 0; j < nfr; ++j)
      if (info[j].kind == 2) { if (first < 0) first = j; info[j].row0 = (int16_t)(9 + 32 * n); ++n; }
    s_nsyn = n;
    s_halo = 0;
    if (n) {
      // the frame whose last nine slots the first synthesised frame inherits, and whether it lies inside the tile
      const uint32_t it = info[first].it;
      const SbrFrameDev* fp = k4_frame(sframes, run, it);
      const uint32_t back = (PS && bank == 1) ? fp->back_ps : fp->back;
      const bool in_tile = (PS && bank == 1) ? (back != 0 && it - back >= tile.lo) : (fp->ord > ord_lo);
      s_halo = in_tile ? 2 : 1;
      if (in_tile) {
        const uint32_t q = it - back;
        const bool q_ps = PS && (bank == 1 || ps_frames[run.ps_base + q].use_ps != 0);
        if (q_ps) {
          s_halo_src = ps_matrix(q) + 23 * kXgRow;   // slots 23..31 of the matrix K5 made of that frame
          for (int h = 0; h < 9; ++h) s_halo_lim[h] = 64;
        } else {
          // slots 23..31 of the previous processed frame: its rows 25..33
          const SbrFrameDev* fq = k4_frame(sframes, run, q);
          s_halo_src = xg + ((size_t)r * tile.rows + 32 * (size_t)(fq->ord - ord_lo) + kSbrHfAdj + 23) * kXgRow;
          for (int h = 0; h < 9; ++h) s_halo_lim[h] = k4_x_limit(fq, fq->mode, 23 + h);
        }
      }
    }
  }
  __syncthreads();
  const int nsyn = s_nsyn, halo = s_halo;

  // ---- stage the QMF rows (band limit applied): half a CTA per row, thread = band
  if (nsyn) {
    const int band = t & 63, rsel = t >> 6;   // kK4cThreads / 64 rows at a time
    constexpr int kRowsPerIter = kK4cThreads / 64;
    // one QMF sample into its staged row: [Re X(0..63) | Im X(63..0)] for the 64-band bank; the down-sampled bank keeps
    // bands 0..31 only, pre-twiddled and scaled (SynthesisFilterbank32.java:57-64), as [x1(0..31) | x2(0..31)]
    float tw0 = 0.f, tw1 = 0.f;
    if (DS && band < 32) { tw0 = __ldg(T.qmf32_tw + 2 * band); tw1 = __ldg(T.qmf32_tw + 2 * band + 1); }
    auto stage = [&](float* row, float2 x) {
      if (DS) {
        if (band < 32) {
          float x1 = (x.x * tw0) - (x.y * tw1);
          float x2 = (x.y * tw0) + (x.x * tw1);
          x1 *= 1.f / 64.f;
          x2 *= 1.f / 64.f;
          row[band] = x1;
          row[32 + band] = x2;
        }
      } else {
        row[band] = x.x;
        row[127 - band] = x.y;
      }
    };
    for (int j = 0; j < nfr; ++j) {
      if (info[j].kind != 2) continue;
      const float* src = info[j].src;
      const int row0 = info[j].row0, lim_lo = info[j].lim_lo, lim_hi = info[j].lim_hi, fs = info[j].first_slot;
      // all loads of the thread are issued before the first one is consumed
      float2 v[32 / kRowsPerIter];
#pragma unroll
      for (int u = 0; u < 32 / kRowsPerIter; ++u) {
        const int l = rsel + u * kRowsPerIter;
        v[u] = make_float2(0.f, 0.f);
        if (band < (l < fs ? lim_lo : lim_hi)) v[u] = __ldg(reinterpret_cast<const float2*>(src + (size_t)l * kXgRow) + band);
      }
#pragma unroll
      for (int u = 0; u < 32 / kRowsPerIter; ++u) {
        const int l = rsel + u * kRowsPerIter;
        stage(vb + (row0 + l) * kVbStride, v[u]);
      }
    }
    if (halo == 2) {
      const float* src = s_halo_src;
      for (int h = rsel; h < 9; h += kRowsPerIter) {
        float2 v = make_float2(0.f, 0.f);
        if (band < s_halo_lim[h]) v = __ldg(reinterpret_cast<const float2*>(src + (size_t)h * kXgRow) + band);
        stage(vb + h * kVbStride, v);
      }
    } else {
      // carried v-vectors: row 8 is the newest (slot -1), row 0 the oldest (slot -9); [0] of the state = newest
      for (int i = t; i < 9 * 128; i += kK4cThreads)
        if ((i % 128) < kVLen) vb[(8 - i / 128) * kVbStride + (i % 128)] = v_in[i];
    }
  }
  __syncthreads();
  // ---- the DCTs: one slot per thread (the frames' slots first, then the nine inherited ones)
  {
    const int n_main = 32 * nsyn;
    int srow = -1;
    if (t < n_main) srow = 9 + t;
    else if (halo == 2 && t - n_main < 9) srow = t - n_main;
    if (srow >= 0) { if (DS) sbr_synth_slot32(vb + srow * kVbStride); else sbr_synth_slot(vb + srow * kVbStride); }
  }
  __syncthreads();
  // ---- window + output: a warp per time slot, lanes take output samples lane and lane + 32
  auto put_sample = [&](const K4cFrame& f, int i, float v) {
    if (PCM_FORMAT == 2) {
      float* d = reinterpret_cast<float*>(f.dst);
      d[(size_t)out_ch * kOutLen + i] = v;
      if (f.dup) d[(size_t)(out_ch + 1) * kOutLen + i] = v;
    } else {
      uint32_t u = pcm_round16(v);
      if (PCM_FORMAT == 1) u = __byte_perm(u, 0, 0x4401);
      uint16_t* d = reinterpret_cast<uint16_t*>(f.dst);
      if (f.pair_store) reinterpret_cast<uint32_t*>(d)[i] = u | (u << 16);
      else {
        d[(size_t)i * n_out + out_ch] = (uint16_t)u;
        if (f.dup) d[(size_t)i * n_out + out_ch + 1] = (uint16_t)u;
      }
    }
  };
  if (nsyn) {
    constexpr int kHalves = DS ? 1 : 2;   // output samples per lane and slot
    float qc[kHalves][10];
#pragma unroll
    for (int hf = 0; hf < kHalves; ++hf)
#pragma unroll
      for (int j = 0; j < 10; ++j)   // (lane-dependent index: not the constant bank); the 32-band bank takes every other tap (:80-89)
        qc[hf][j] = DS ? __ldg(T.qmf_c + 2 * lane + 64 * j) : __ldg(T.qmf_c + lane + 32 * hf + 64 * j);
    for (int j = 0; j < nfr; ++j) {
      if (info[j].kind != 2) continue;
      const K4cFrame f = info[j];
      for (int l = warp; l < 32; l += kK4cThreads / 32) {
        const int cur = f.row0 + l;
#pragma unroll
        for (int hf = 0; hf < kHalves; ++hf) {
          const int k = lane + 32 * hf;
          float ov = (vb[cur * kVbStride + k] * qc[hf][0]);
#pragma unroll
          for (int jj = 1; jj < 10; ++jj) ov = ov + (vb[(cur - jj) * kVbStride + k + (kVLen / 2) * (jj & 1)] * qc[hf][jj]);
          put_sample(f, (kVLen / 2) * l + k, ov);
        }
      }
      // ---- the bank's last frame of the tile hands its nine newest v-vectors to the next tile
      if (f.last_in_tile) {
        for (int i = t; i < 9 * 128; i += kK4cThreads)
          if ((i % 128) < kVLen) v_out[i] = vb[(f.row0 + 31 - i / 128) * kVbStride + (i % 128)];
        if (t == 0) { if (PS && bank == 1) pst->v_flip = 1; else st->v_flip = 1; }
      }
    }
  }
  // ---- frames without valid SBR data: SBR.upsample (sbr/SBR.java:302-309; sample 1 keeps the core value)
  for (int j = 0; j < nfr; ++j) {
    if (info[j].kind != 1) continue;
    const K4cFrame f = info[j];
    // (down-sampled streams keep the core's length: SCE/CPE.process leave the core PCM as it is)
    for (int i = t; i < kOutLen; i += kK4cThreads) put_sample(f, i, DS ? f.cs[i] : (i < 2 ? f.cs[i] : f.cs[i >> 1]));
  }
}

__global__ void k4_commit_kernel(const K4RunDev* __restrict__ runs, uint32_t n_runs, uint32_t n_plain, SbrChanDev* __restrict__ chans,
                                 PsChanDev* __restrict__ ps_chans) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_runs) return;
  SbrChanDev* st = chans + (size_t)runs[r].stream_slot * kSbrChansPerStream + runs[r].ch_slot;
  if (st->v_flip) { st->v_sel ^= 1; st->v_flip = 0; }
  if (r >= n_plain) {
    PsChanDev* pst = ps_chans + runs[r].stream_slot;
    if (pst->v_flip) { pst->v_sel ^= 1; pst->v_flip = 0; }
  }
}

// ---- K5: parametric stereo (ps/PSImpl.java:685-707) for the frames of an SBR+PS stream that carry ps_data: hybrid
// analysis, transient detection, all-pass decorrelation, H-matrix mixing, hybrid synthesis.  The decorrelator and the
// transient detector are recursive in time, so one CTA walks the frames of the tile in order with the delay lines in
// shared memory for the whole tile (96 threads: one per time slot in the filterbank phases; one per band in the
// decorrelator: threads 3..63 = QMF bands 3..63, threads 64..73 = the ten hybrid sub-bands).  Input is the finished
// Xsbr matrix in xg, output the left / right QMF matrices the synthesis kernel reads:
//   xps[ps run][frame of the tile][left, right][32 slots][64 bands][re, im]
constexpr int kK5Threads = 96;
constexpr int kK5PhaseThread = 80;  // an otherwise idle thread: the sequential IPD/OPD phase bookkeeping
constexpr int kK5Bands = 71;        // 10 hybrid sub-bands + QMF bands 3..63
constexpr int kK5AllPass = 30;      // of which run the all-pass chain: the hybrid ones and QMF bands 3..22
constexpr int kK5Floats = 2 * 32 * 24 + 32 * 20 + 3 * 88 + kK5Bands * 28 + kK5AllPass * 30 + 22 * 8 + 72 + 22 * 5 * 4 + 80;
#ifndef K5_MIN_BLOCKS
#define K5_MIN_BLOCKS 6
#endif
// QMF band 3..63 -> its parameter band minus 8 (groups 10..21 of the 20-band configuration, ps/PSTables.java group_border20)
__host__ __device__ constexpr int k5_qmf_pb(int band) {
  return band < 9 ? band - 3 : (band < 11 ? 6 : (band < 14 ? 7 : (band < 18 ? 8 : (band < 23 ? 9 : (band < 35 ? 10 : 11)))));
}
static_assert(kK5Floats % 4 == 0, "PsFrameDev must land 16-byte aligned");
constexpr size_t k5_smem_bytes() { return sizeof(float) * kK5Floats + sizeof(PsFrameDev); }

__global__ void __launch_bounds__(kK5Threads, K5_MIN_BLOCKS)
k5_ps_kernel(const K4RunDev* __restrict__ runs, uint32_t run0, const SbrFrameDev* __restrict__ sframes,
             const PsFrameDev* __restrict__ ps_frames, PsChanDev* __restrict__ ps_chans, const float* __restrict__ xg,
             float* __restrict__ xps, SbrTablesDev T, K4Tile tile) {
  extern __shared__ __align__(16) float k5_smem[];
  // The frame's two 32 x 64 QMF matrices are NOT staged in shared memory (they were: 33 KB of the 58 KB a CTA needed, three
  // CTAs = nine warps per SM): the decorrelator's band threads read X_left from the tile workspace (L2) slot by slot and
  // write both outputs straight to xps.  The only whole-matrix quantity another thread needs is |X|^2 for the transient
  // detector's band energies: the second warp sums them per slot straight from the slot's row in the workspace (512
  // contiguous bytes per lane) while the first runs the hybrid analysis -- an 8 KB energy plane in between (band threads
  // write, slot threads read) held the kernel at six CTAs per SM.
  float* hyl = k5_smem;                                  // [32][12][2]
  float* hyr = hyl + 32 * 24;                            // [32][12][2]
  float* pg = hyr + 32 * 24;                             // [32][20] band energies, then transient ratios
  float* hwork = pg + 32 * 20;                           // [3][44][2] hybrid analysis input
  float* dly = hwork + 3 * 88;                           // [71][14][2] input delay of every decorrelator band
  float* ser = dly + kK5Bands * 28;                      // [30][3 links][5][2] all-pass delay lines
  float* hprev = ser + kK5AllPass * 30;                  // [22][8] mixing matrix of the previous envelope, per group (re x4, im x4)
  float* hybuf = hprev + 22 * 8;                         // [3][12][2] hybrid analysis history
  float* phases = hybuf + 72;                            // [22][5][4] IPD/OPD: phaseLeft, phaseRight per (group, envelope)
  float* pdprev = phases + 22 * 5 * 4;                   // [20][2][2] PDData.prev
  PsFrameDev* pp = reinterpret_cast<PsFrameDev*>(pdprev + 80);
#define HYL(n, k, c) hyl[((n) * 12 + (k)) * 2 + (c)]
#define HYR(n, k, c) hyr[((n) * 12 + (k)) * 2 + (c)]
  const int t = threadIdx.x;
  const uint32_t r = run0 + blockIdx.x;
  const K4RunDev run = runs[r];
  if (tile.lo >= run.count) return;
  const uint32_t hi = min(run.count, tile.lo + tile.ft);
  PsChanDev* pst = ps_chans + run.stream_slot;
  const float* Xrun = xg + (size_t)r * tile.rows * kXgRow;
  const uint32_t ord_lo = k4_frame(sframes, run, tile.lo)->ord;

  // ---- the band this thread decorrelates, and its state
  const bool hyb = t >= 64;
  const bool band_task = (t >= 3 && t < 64) || (t >= 64 && t < 74);
  int gr = 0, sb = 0;                                    // group; hybrid sub-band or QMF band
  if (band_task) {
    if (hyb) { gr = t - 64; sb = ps_group_border(gr); }
    else { sb = t; gr = 10; while (ps_group_border(gr + 1) <= sb) ++gr; }
  }
  const int bslot = hyb ? gr : 10 + (sb - 3);
  const bool delay_band = !hyb && sb > 22;
  const bool allpass = band_task && !delay_band;
  float* my_dly = dly + bslot * 28;
  float* my_ser = ser + min(bslot, kK5AllPass - 1) * 30;
  const int bk = ps_bk(gr);
  int td = pst->saved_delay, s0 = pst->delay_buf_index_ser[0], s1 = pst->delay_buf_index_ser[1], s2 = pst->delay_buf_index_ser[2];
  int di = 0;
  const int dD = sb < 35 ? 14 : 1;
  if (band_task) {
    const float* gd = hyb ? &pst->delay_sub[sb][0][0] : &pst->delay_qmf[sb][0][0];
    const int nd = hyb ? 4 : 28;
    for (int i = 0; i < nd; ++i) my_dly[i] = gd[i];
    if (allpass) {
      const float* gs = hyb ? &pst->delay_sub_ser[sb][0][0][0] : &pst->delay_qmf_ser[sb][0][0][0];
      for (int i = 0; i < 30; ++i) my_ser[i] = gs[i];
    }
    if (delay_band) di = pst->delay_buf_index_delay[sb];
  }
  if (t < 22)
    for (int i = 0; i < 8; ++i) hprev[t * 8 + i] = pst->h_prev[t][i];
  if (t < 72) hybuf[t] = (&pst->hyb_buffer[0][0][0])[t];
  if (t < 80) pdprev[t] = (&pst->pd_prev[0][0][0])[t];
  int phase_hist = pst->phase_hist;             // (uniform: every thread counts the pairs)
  float peak = 0, pprev = 0, smooth_prev = 0;   // transient detector of parameter band t
  if (t < 20) { peak = pst->P_PeakDecayNrg[t]; pprev = pst->P_prev[t]; smooth_prev = pst->P_SmoothPeakDecayDiffNrg_prev[t]; }
  // per-band constants of the decorrelator (ps/PSImpl.java:266-396)
  float gf0 = 0, gf1 = 0, gf2 = 0, phi0 = 0, phi1 = 0, q00 = 0, q01 = 0, q10 = 0, q11 = 0, q20 = 0, q21 = 0;
  if (allpass) {
    float g_DecaySlope;
    if (hyb || sb <= 3) g_DecaySlope = 1.0f;
    else {
      const int decay = 3 - sb;
      g_DecaySlope = (decay <= -20) ? 0.f : 1.0f + 0.05f * (float)decay;
    }
    gf0 = g_DecaySlope * __ldg(T.ps_filter_a); gf1 = g_DecaySlope * __ldg(T.ps_filter_a + 1); gf2 = g_DecaySlope * __ldg(T.ps_filter_a + 2);
    phi0 = __ldg((hyb ? T.ps_phi_sub : T.ps_phi_qmf) + 2 * sb); phi1 = __ldg((hyb ? T.ps_phi_sub : T.ps_phi_qmf) + 2 * sb + 1);
    const float* qf = (hyb ? T.ps_q_sub : T.ps_q_qmf) + sb * 6;
    q00 = __ldg(qf); q01 = __ldg(qf + 1); q10 = __ldg(qf + 2); q11 = __ldg(qf + 3); q20 = __ldg(qf + 4); q21 = __ldg(qf + 5);
  }
  __syncthreads();

  for (uint32_t it = tile.lo; it < hi; ++it) {
    const SbrFrameDev* fp = k4_frame(sframes, run, it);
    const int mode = fp->mode;
    if (fp->frame_status != 0 || mode == 0) continue;
    const PsFrameDev* pf = ps_frames + run.ps_base + it;
    if (pf->use_ps == 0) continue;
    __syncthreads();
    if (t < (int)(sizeof(PsFrameDev) / 16)) reinterpret_cast<uint4*>(pp)[t] = __ldg(reinterpret_cast<const uint4*>(pf) + t);
    const float* X = Xrun + 32 * (size_t)(fp->ord - ord_lo) * kXgRow;
    // X_left = the band-limited copy of Xsbr (SBR1.processPS): band t of slot l is X[l + 2][t] below the limit, zero above.
    // Hybrid analysis input: QMF bands 0..2 of slots 6..37
    const int lim_lo = k4_x_limit(fp, mode, 0), lim_hi = k4_x_limit(fp, mode, 31), fs = mode == 2 ? fp->t_E[0] : 0;
    float* outl = xps + ((size_t)(blockIdx.x * tile.ft + (it - tile.lo)) * 2) * 32 * kXgRow;
    float* outr = outl + 32 * kXgRow;
    // Hybrid analysis input: work[0..11] = history, work[12 + n] = X[n + 6][band]: rows 6..31 come from X_left, rows 32..37
    // straight from Xsbr (:108-113) -- for bands 0..2 both are the unmodified analysis output
    for (int i = t; i < 3 * 44; i += kK5Threads) {
      const int band = i / 44, j = i % 44;
      float2 v;
      if (j < 12) v = make_float2(hybuf[(band * 12 + j) * 2], hybuf[(band * 12 + j) * 2 + 1]);
      else v = __ldg(reinterpret_cast<const float2*>(X + (size_t)(j - 12 + 6 + kSbrHfAdj) * kXgRow) + band);
      hwork[(band * 44 + j) * 2] = v.x;
      hwork[(band * 44 + j) * 2 + 1] = v.y;
    }
    __syncthreads();
    const int num_env = pp->num_env;
    for (int u = t - 64; u >= 0 && u < 36; u += 32) { const int band = u / 12, j = u % 12; hybuf[(band * 12 + j) * 2] = hwork[(band * 44 + 32 + j) * 2]; hybuf[(band * 12 + j) * 2 + 1] = hwork[(band * 44 + 32 + j) * 2 + 1]; }
    // ---- hybrid analysis (ps/Filterbank.java:18-68): thread n
    if (t < 32) {
      const int i = t;
      {
        // Filter8 (ps/Filter8.java:53-122) on QMF band 0
        const float* b = hwork;
        float f[7];
#pragma unroll
        for (int z = 0; z < 7; ++z) f[z] = __ldg(T.ps_p8 + z);
#define BR(k) b[((k) + i) * 2]
#define BI(k) b[((k) + i) * 2 + 1]
        auto dct3 = [](float (&y)[4], const float (&x)[4]) {   // DCT3_4_unscaled (:124-138)
          const float f0 = (x[2] * 0.7071067811865476f);
          const float f1 = x[0] - f0;
          const float f2 = x[0] + f0;
          const float f3 = x[1] + x[3];
          const float f4 = (x[1] * 1.3065629648763766f);
          const float f5 = (f3 * (-0.9238795325112866f));
          const float f6 = (x[3] * (-0.5411961001461967f));
          const float f7 = f4 + f5;
          const float f8 = f6 - f5;
          y[3] = f2 - f8; y[0] = f2 + f8; y[2] = f1 - f7; y[1] = f1 + f7;
        };
        float re1[4], im1[4], re2[4], im2[4], x[4], y[4];
        re1[0] = (f[6] * BR(6));
        re1[1] = (f[5] * (BR(5) + BR(7)));
        re1[2] = -(f[0] * (BR(0) + BR(12))) + (f[4] * (BR(4) + BR(8)));
        re1[3] = -(f[1] * (BR(1) + BR(11))) + (f[3] * (BR(3) + BR(9)));
        im1[0] = (f[5] * (BI(7) - BI(5)));
        im1[1] = (f[0] * (BI(12) - BI(0))) + (f[4] * (BI(8) - BI(4)));
        im1[2] = (f[1] * (BI(11) - BI(1))) + (f[3] * (BI(9) - BI(3)));
        im1[3] = (f[2] * (BI(10) - BI(2)));
#pragma unroll
        for (int n = 0; n < 4; n++) x[n] = re1[n] - im1[3 - n];
        dct3(y, x);
        HYL(i, 7, 0) = y[0]; HYL(i, 5, 0) = y[2]; HYL(i, 3, 0) = y[3]; HYL(i, 1, 0) = y[1];
#pragma unroll
        for (int n = 0; n < 4; n++) x[n] = re1[n] + im1[3 - n];
        dct3(y, x);
        HYL(i, 6, 0) = y[1]; HYL(i, 4, 0) = y[3]; HYL(i, 2, 0) = y[2]; HYL(i, 0, 0) = y[0];
        im2[0] = (f[6] * BI(6));
        im2[1] = (f[5] * (BI(5) + BI(7)));
        im2[2] = -(f[0] * (BI(0) + BI(12))) + (f[4] * (BI(4) + BI(8)));
        im2[3] = -(f[1] * (BI(1) + BI(11))) + (f[3] * (BI(3) + BI(9)));
        re2[0] = (f[5] * (BR(7) - BR(5)));
        re2[1] = (f[0] * (BR(12) - BR(0))) + (f[4] * (BR(8) - BR(4)));
        re2[2] = (f[1] * (BR(11) - BR(1))) + (f[3] * (BR(9) - BR(3)));
        re2[3] = (f[2] * (BR(10) - BR(2)));
#pragma unroll
        for (int n = 0; n < 4; n++) x[n] = im2[n] + re2[3 - n];
        dct3(y, x);
        HYL(i, 7, 1) = y[0]; HYL(i, 5, 1) = y[2]; HYL(i, 3, 1) = y[3]; HYL(i, 1, 1) = y[1];
#pragma unroll
        for (int n = 0; n < 4; n++) x[n] = im2[n] - re2[3 - n];
        dct3(y, x);
        HYL(i, 6, 1) = y[1]; HYL(i, 4, 1) = y[3]; HYL(i, 2, 1) = y[2]; HYL(i, 0, 1) = y[0];
#undef BR
#undef BI
      }
      for (int band = 1; band < 3; ++band) {
        // Filter2 (ps/Filter2.java:40-68) on QMF bands 1 and 2
        const float* b = hwork + band * 88;
        float f[7];
#pragma unroll
        for (int z = 0; z < 7; ++z) f[z] = __ldg(T.ps_p2 + z);
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const float r0 = (f[0] * (b[(0 + i) * 2 + c] + b[(12 + i) * 2 + c]));
          const float r1 = (f[1] * (b[(1 + i) * 2 + c] + b[(11 + i) * 2 + c]));
          const float r2 = (f[2] * (b[(2 + i) * 2 + c] + b[(10 + i) * 2 + c]));
          const float r3 = (f[3] * (b[(3 + i) * 2 + c] + b[(9 + i) * 2 + c]));
          const float r4 = (f[4] * (b[(4 + i) * 2 + c] + b[(8 + i) * 2 + c]));
          const float r5 = (f[5] * (b[(5 + i) * 2 + c] + b[(7 + i) * 2 + c]));
          const float r6 = (f[6] * b[(6 + i) * 2 + c]);
          HYL(i, 8 + 2 * (band - 1), c) = r0 + r1 + r2 + r3 + r4 + r5 + r6;
          HYL(i, 9 + 2 * (band - 1), c) = r0 - r1 + r2 - r3 + r4 - r5 + r6;
        }
      }
      // group hybrid channels (:56-66)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        HYL(i, 3, c) += HYL(i, 4, c);
        HYL(i, 4, c) = 0;
        HYL(i, 2, c) += HYL(i, 5, c);
        HYL(i, 5, c) = 0;
      }
#pragma unroll
      for (int k = 0; k < 12; ++k) { HYR(i, k, 0) = 0.f; HYR(i, k, 1) = 0.f; }
      // ---- energy per parameter band (ps_decorrelate, :213-234): groups in order, sub-bands in order.  Parameter bands
      // 0..7 belong to the hybrid sub-bands (groups 0..9), 8..19 to the QMF bands (groups 10..21, the next warp).
      {
        constexpr int gb[10] = {6, 7, 0, 1, 2, 3, 9, 8, 10, 11};
        float P[8];
#pragma unroll
        for (int b2 = 0; b2 < 8; ++b2) P[b2] = 0.f;
#pragma unroll
        for (int g2 = 0; g2 < 10; ++g2) {
          const int pb = g2 == 0 ? 1 : (g2 == 1 ? 0 : g2 - 2);
          const float re = HYL(i, gb[g2], 0), im = HYL(i, gb[g2], 1);
          P[pb] += (re * re) + (im * im);
        }
#pragma unroll
        for (int b2 = 0; b2 < 8; ++b2) pg[i * 20 + b2] = P[b2];
      }
    } else if (t < 64) {
      // X_left = the band-limited copy of Xsbr (SBR1.processPS): band b of slot l is X[l + 2][b] below the limit, zero above
      const int i = t - 32;
      const float4* row = reinterpret_cast<const float4*>(X + (size_t)(i + kSbrHfAdj) * kXgRow);   // [j] = bands 2j, 2j + 1
      JAADB_ASSERT(fp->ord - ord_lo < tile.ft && (size_t)(32 * (fp->ord - ord_lo) + i + kSbrHfAdj) < tile.rows);
      const int lim = i < fs ? lim_lo : lim_hi;
      float P[12];
#pragma unroll
      for (int b2 = 0; b2 < 12; ++b2) P[b2] = 0.f;
#pragma unroll
      for (int j0 = 0; j0 < 32; j0 += 8) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (j0 + u > 0) v[u] = __ldg(row + j0 + u);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int band = 2 * (j0 + u) + h;
            if (band >= 3) {
              const float re = h ? v[u].z : v[u].x, im = h ? v[u].w : v[u].y;
              P[k5_qmf_pb(band)] += band < lim ? (re * re) + (im * im) : 0.f;
            }
          }
        }
      }
#pragma unroll
      for (int b2 = 0; b2 < 12; ++b2) pg[i * 20 + 8 + b2] = P[b2];
    }
    __syncthreads();
    // ---- transient reduction ratio (:236-264): thread bk, sequential in time
    if (t < 20) {
      for (int n = 0; n < 32; ++n) {
        const float Pn = pg[n * 20 + t];
        const float gamma = 1.5f;
        peak = (peak * 0.76592833836465f);
        if (peak < Pn) peak = Pn;
        float sm = smooth_prev;
        sm += ((peak - Pn - smooth_prev) * 0.25f);
        smooth_prev = sm;
        float nrg = pprev;
        nrg += ((Pn - pprev) * 0.25f);
        pprev = nrg;
        pg[n * 20 + t] = ((sm * gamma) <= nrg) ? 1.0f : (nrg / (sm * gamma));
      }
    }
    // ---- IPD/OPD phase rotation parameters (ps_mix_phase, :488-560) alongside the transient detector.  Quirks kept
    // (SURVEY A-15): opd_index is read from ipd -- so ipd.prev and opd.prev always hold the same values (pdprev) -- and the
    // value "before previous" comes from opd.prev for both.
    const int nr_ipdopd_par = pp->nr_ipdopd_par;
    if (nr_ipdopd_par) {
      // The reference walks (group, envelope) pairs in group order and flips phase_hist once per pair; the groups it visits
      // are 0 .. nr_ipdopd_par + 1, so pair (g, env) sees phase_hist + g * num_env + env.  A parameter band's history is only
      // touched by its own pairs (bands 0 and 1 have two groups each: 1, 2 and 0, 3), so the bands run on one thread each
      // -- threads that have no decorrelator band -- with their pairs in the reference's order.
      const int b2 = t - (kK5Threads - 20);
      if (b2 >= 0 && b2 < nr_ipdopd_par) {
        const int ga = b2 == 0 ? 1 : (b2 == 1 ? 0 : b2 + 2), gb = b2 == 0 ? 2 : (b2 == 1 ? 3 : -1);
        for (int k = 0; k < 2; ++k) {
          const int g2 = k ? gb : ga;
          if (g2 < 0) break;
          for (int env = 0; env < num_env; ++env) {
            const int ph0 = (phase_hist + g2 * num_env + env) & 1;
            float* pv = pdprev + (b2 * 2 + ph0) * 2;
            float tl0 = (pv[0] * 0.25f), tl1 = (pv[1] * 0.25f), tr0 = (pv[0] * 0.25f), tr1 = (pv[1] * 0.25f);
            const int ix = min(abs((int)pp->ipd[env][b2]), 8);
            const float c0 = __ldg(T.ps_ipdopd_cos + ix), s0v = __ldg(T.ps_ipdopd_sin + ix);
            pv[0] = c0; pv[1] = s0v;
            tl0 += c0; tl1 += s0v; tr0 += c0; tr1 += s0v;
            const float* pb = pdprev + (b2 * 2 + (ph0 ^ 1)) * 2;
            tl0 += (pb[0] * 0.5f); tl1 += (pb[1] * 0.5f); tr0 += (pb[0] * 0.5f); tr1 += (pb[1] * 0.5f);
            // magnitude_c (:402-404): (float) Math.sqrt of a float sum of squares -- a correctly rounded float sqrt is the same value
            const float xy = __fsqrt_rn((tr0 * tr0) + (tr1 * tr1)), pq = __fsqrt_rn((tl0 * tl0) + (tl1 * tl1));
            float pl0 = 0.f, pl1 = 0.f, pr0 = 0.f, pr1 = 0.f;
            if (xy != 0.f) { pl0 = __fdiv_rn(tr0, xy); pl1 = __fdiv_rn(tr1, xy); }
            const float xypq = (xy * pq);
            if (xypq != 0.f) {
              const float tmp1 = (tr0 * tl0) + (tr1 * tl1), tmp2 = (tr1 * tl0) - (tr0 * tl1);
              pr0 = __fdiv_rn(tmp1, xypq); pr1 = __fdiv_rn(tmp2, xypq);
            }
            float* ph = phases + (g2 * 5 + env) * 4;
            ph[0] = pl0; ph[1] = pl1; ph[2] = pr0; ph[3] = pr1;
          }
        }
      }
      phase_hist = (phase_hist + min(22, nr_ipdopd_par + 2) * num_env) & 1;   // (every thread keeps the count)
    }
    // the mixing matrices the groups ended the previous envelope with (read by every band of the group before any of
    // them stores the new ones)
    float hp11 = 1.f, hp12 = 0.f, hp21 = 0.f, hp22 = 0.f, hq11 = 0.f, hq12 = 0.f, hq21 = 0.f, hq22 = 0.f;
    if (band_task) {
      hp11 = hprev[gr * 8]; hp12 = hprev[gr * 8 + 1]; hp21 = hprev[gr * 8 + 2]; hp22 = hprev[gr * 8 + 3];
      hq11 = hprev[gr * 8 + 4]; hq12 = hprev[gr * 8 + 5]; hq21 = hprev[gr * 8 + 6]; hq22 = hprev[gr * 8 + 7];
    }
    __syncthreads();
    // ---- decorrelation (:266-396) + mixing (ps_mix_phase, :406-681), one band per thread
    if (band_task) {
      const bool fine = pp->iid_mode >= 3;
      const int num_steps = fine ? 15 : 7;
      const float* sf_iid = T.ps_sf_iid[fine];
      float H11 = hp11, H12 = hp12, H21 = hp21, H22 = hp22;
      float dH11 = 0, dH12 = 0, dH21 = 0, dH22 = 0;
      // imaginary parts: only for the parameter bands the IPD/OPD extension covers (bk < nr_ipdopd_par)
      const bool rot = bk < nr_ipdopd_par;
      const bool bkm = bk != 0;   // FBType.bkm tests `& ~NEGATE_IPD_MASK` (FBType.java:71-73, A-13): true for every bk != 0
      float G11 = 0, G12 = 0, G21 = 0, G22 = 0, dG11 = 0, dG12 = 0, dG21 = 0, dG22 = 0;
      int env = -1, env_end = 0;
      // X_left[n][sb] of a QMF band comes from the tile workspace, one slot ahead of its use
      const float2* xcol = reinterpret_cast<const float2*>(X + (size_t)kSbrHfAdj * kXgRow) + sb;
      float2 x_next = hyb ? make_float2(0.f, 0.f) : __ldg(xcol);
      for (int n = 0; n < 32; ++n) {
        const float2 x_cur = x_next;
        if (!hyb && n < 31) x_next = __ldg(xcol + (size_t)(n + 1) * (kXgRow / 2));
        if (n == env_end) {
          // next envelope: target H from the IID / ICC indices (:424-478), linear interpolation over its length
          do { ++env; env_end = pp->border[env + 1]; } while (env + 1 < num_env && env_end <= n);
          int iid_index = pp->iid[env][bk];
          const int iid_sign = iid_index < 0 ? -1 : 1;
          iid_index = min(abs(iid_index), num_steps);
          const int icc_index = min(max((int)pp->icc[env][bk], 0), 7);
          float h11, h12, h21, h22;
          if (pp->icc_mode < 3) {
            const float c_1 = __ldg(sf_iid + num_steps + iid_index), c_2 = __ldg(sf_iid + num_steps - iid_index);
            const float cosa = __ldg(T.ps_cos_alphas + icc_index), sina = __ldg(T.ps_sin_alphas + icc_index);
            const float cosb = __ldg(T.ps_cos_betas[fine] + iid_index * 8 + icc_index);
            const float sinb = __ldg(T.ps_sin_betas[fine] + iid_index * 8 + icc_index) * (float)iid_sign;
            const float ab1 = (cosb * cosa), ab2 = (sinb * sina), ab3 = (sinb * cosa), ab4 = (cosb * sina);
            h11 = (c_2 * (ab1 - ab2));
            h12 = (c_1 * (ab1 + ab2));
            h21 = (c_2 * (ab3 + ab4));
            h22 = (c_1 * (ab3 - ab4));
          } else {
            const float cosa = __ldg(T.ps_sincos_alphas_b[fine] + (num_steps + iid_index) * 8 + icc_index);
            const float sina = __ldg(T.ps_sincos_alphas_b[fine] + (2 * num_steps - (num_steps + iid_index)) * 8 + icc_index);
            const float cosg = __ldg(T.ps_cos_gammas[fine] + iid_index * 8 + icc_index);
            const float sing = __ldg(T.ps_sin_gammas[fine] + iid_index * 8 + icc_index);
            h11 = (1.4142135623731f * (cosa * cosg));
            h12 = (1.4142135623731f * (sina * cosg));
            h21 = (1.4142135623731f * (-cosa * sing));
            h22 = (1.4142135623731f * (sina * sing));
          }
          float g11 = 0, g12 = 0, g21 = 0, g22 = 0;
          if (rot) {
            const float* ph = phases + (gr * 5 + env) * 4;   // phaseLeft, phaseRight (:562-571)
            g11 = (h11 * ph[1]); g12 = (h12 * ph[3]); g21 = (h21 * ph[1]); g22 = (h22 * ph[3]);
            h11 = (h11 * ph[0]); h12 = (h12 * ph[2]); h21 = (h21 * ph[0]); h22 = (h22 * ph[2]);
          }
          const float L = (float)(pp->border[env + 1] - pp->border[env]);
          dH11 = (h11 - hp11) / L; dH12 = (h12 - hp12) / L; dH21 = (h21 - hp21) / L; dH22 = (h22 - hp22) / L;
          H11 = hp11; H12 = hp12; H21 = hp21; H22 = hp22;
          hp11 = h11; hp12 = h12; hp21 = h21; hp22 = h22;
          if (rot) {
            dG11 = (g11 - hq11) / L; dG12 = (g12 - hq12) / L; dG21 = (g21 - hq21) / L; dG22 = (g22 - hq22) / L;
            G11 = hq11; G12 = hq12; G21 = hq21; G22 = hq22;
            if (bkm) { dG11 = -dG11; dG12 = -dG12; dG21 = -dG21; dG22 = -dG22; G11 = -G11; G12 = -G12; G21 = -G21; G22 = -G22; }
            hq11 = g11; hq12 = g12; hq21 = g21; hq22 = g22;
          }
        }
        // -- decorrelate
        const int lim = n < fs ? lim_lo : lim_hi;
        const float re = hyb ? HYL(n, sb, 0) : (sb < lim ? x_cur.x : 0.f);
        const float im = hyb ? HYL(n, sb, 1) : (sb < lim ? x_cur.y : 0.f);
        float r0Re, r0Im;
        if (delay_band) {
          float* d = my_dly + 2 * di;
          r0Re = d[0]; r0Im = d[1];
          d[0] = re; d[1] = im;
        } else {
          float* d = my_dly + 2 * td;
          float tmp0Re = d[0], tmp0Im = d[1];
          d[0] = re; d[1] = im;
          r0Re = (tmp0Re * phi0) + (tmp0Im * phi1);
          r0Im = (tmp0Im * phi0) - (tmp0Re * phi1);
#pragma unroll
          for (int m = 0; m < 3; ++m) {
            const float qa = m == 0 ? q00 : (m == 1 ? q10 : q20), qb = m == 0 ? q01 : (m == 1 ? q11 : q21);
            const float gf = m == 0 ? gf0 : (m == 1 ? gf1 : gf2);
            float* dl = my_ser + (m * 5 + (m == 0 ? s0 : (m == 1 ? s1 : s2))) * 2;
            tmp0Re = dl[0]; tmp0Im = dl[1];
            float tmpRe = (tmp0Re * qa) + (tmp0Im * qb);
            float tmpIm = (tmp0Im * qa) - (tmp0Re * qb);
            tmpRe -= gf * r0Re;
            tmpIm -= gf * r0Im;
            dl[0] = r0Re + (gf * tmpRe);
            dl[1] = r0Im + (gf * tmpIm);
            r0Re = tmpRe;
            r0Im = tmpIm;
          }
        }
        const float G = pg[n * 20 + bk];
        const float rRe = (G * r0Re), rIm = (G * r0Im);
        if (++td >= 2) td = 0;
        if (delay_band) { if (++di >= dD) di = 0; }
        if (++s0 >= 3) s0 = 0;
        if (++s1 >= 4) s1 = 0;
        if (++s2 >= 5) s2 = 0;
        // -- mix
        H11 += dH11; H12 += dH12; H21 += dH21; H22 += dH22;
        float lRe = (H11 * re) + (H21 * rRe), lIm = (H11 * im) + (H21 * rIm);
        float oRe = (H12 * re) + (H22 * rRe), oIm = (H12 * im) + (H22 * rIm);
        if (rot) {
          // apply rotation (:650-656)
          G11 += dG11; G12 += dG12; G21 += dG21; G22 += dG22;
          lRe -= (G11 * im) + (G21 * rIm);
          lIm += (G11 * re) + (G21 * rRe);
          oRe -= (G12 * im) + (G22 * rIm);
          oIm += (G12 * re) + (G22 * rRe);
        }
        if (hyb) { HYL(n, sb, 0) = lRe; HYL(n, sb, 1) = lIm; HYR(n, sb, 0) = oRe; HYR(n, sb, 1) = oIm; }
        else {
          reinterpret_cast<float2*>(outl + (size_t)n * kXgRow)[sb] = make_float2(lRe, lIm);
          reinterpret_cast<float2*>(outr + (size_t)n * kXgRow)[sb] = make_float2(oRe, oIm);
        }
      }
      if (sb == ps_group_border(gr)) {
        hprev[gr * 8] = hp11; hprev[gr * 8 + 1] = hp12; hprev[gr * 8 + 2] = hp21; hprev[gr * 8 + 3] = hp22;
        hprev[gr * 8 + 4] = hq11; hprev[gr * 8 + 5] = hq12; hprev[gr * 8 + 6] = hq21; hprev[gr * 8 + 7] = hq22;
      }
    } else {
      // every thread keeps the (uniform) ring positions in step: 32 slots per frame
      td = (td + 32) % 2; s0 = (s0 + 32) % 3; s1 = (s1 + 32) % 4; s2 = (s2 + 32) % 5;
    }
    __syncthreads();
    // ---- hybrid synthesis (ps/Filterbank.java:70-86) for both channels: thread n; QMF bands 0..2 of the two matrices go out
    if (t < 32) {
      float a0[2] = {0, 0}, b0[2] = {0, 0}, a1[2] = {0, 0}, b1[2] = {0, 0}, a2[2] = {0, 0}, b2[2] = {0, 0};
#pragma unroll
      for (int c = 0; c < 2; ++c) {
#pragma unroll
        for (int k = 0; k < 8; ++k) { a0[c] += HYL(t, k, c); b0[c] += HYR(t, k, c); }
        a1[c] += HYL(t, 8, c); a1[c] += HYL(t, 9, c); b1[c] += HYR(t, 8, c); b1[c] += HYR(t, 9, c);
        a2[c] += HYL(t, 10, c); a2[c] += HYL(t, 11, c); b2[c] += HYR(t, 10, c); b2[c] += HYR(t, 11, c);
      }
      float2* ol = reinterpret_cast<float2*>(outl + (size_t)t * kXgRow);
      float2* orr = reinterpret_cast<float2*>(outr + (size_t)t * kXgRow);
      ol[0] = make_float2(a0[0], a0[1]); ol[1] = make_float2(a1[0], a1[1]); ol[2] = make_float2(a2[0], a2[1]);
      orr[0] = make_float2(b0[0], b0[1]); orr[1] = make_float2(b1[0], b1[1]); orr[2] = make_float2(b2[0], b2[1]);
    }
  }
  __syncthreads();
  // ---- state out
  if (band_task) {
    float* gd = hyb ? &pst->delay_sub[sb][0][0] : &pst->delay_qmf[sb][0][0];
    const int nd = hyb ? 4 : 28;
    for (int i = 0; i < nd; ++i) gd[i] = my_dly[i];
    if (allpass) {
      float* gs = hyb ? &pst->delay_sub_ser[sb][0][0][0] : &pst->delay_qmf_ser[sb][0][0][0];
      for (int i = 0; i < 30; ++i) gs[i] = my_ser[i];
    }
    if (delay_band) pst->delay_buf_index_delay[sb] = di;
  }
  if (t == 3) { pst->saved_delay = td; pst->delay_buf_index_ser[0] = s0; pst->delay_buf_index_ser[1] = s1; pst->delay_buf_index_ser[2] = s2; }
  if (t < 22)
    for (int i = 0; i < 8; ++i) pst->h_prev[t][i] = hprev[t * 8 + i];
  if (t < 72) (&pst->hyb_buffer[0][0][0])[t] = hybuf[t];
  if (t < 80) (&pst->pd_prev[0][0][0])[t] = pdprev[t];
  if (t == kK5PhaseThread) pst->phase_hist = phase_hist;
  if (t < 20) { pst->P_PeakDecayNrg[t] = peak; pst->P_prev[t] = pprev; pst->P_SmoothPeakDecayDiffNrg_prev[t] = smooth_prev; }
#undef HYL
#undef HYR
}

}  // namespace jaadb
