// K4 -- SBR signal path.  One CTA (64 threads) owns one SBR channel of one stream and walks that stream's frames of the
// batch in order, with the channel's persistent state (QMF analysis history, the last 8 slots of Xsbr, 9 synthesis
// v-vectors, the gain smoothing ring) resident in shared memory for the whole run:
//   32-band QMF analysis     sbr/AnalysisFilterbank.java:9-73   (one thread per time slot, DCT-IV in registers)
//   HF generation            sbr/HFGeneration.java:17-245       (one thread per high band: covariance LPC + patching)
//   HF adjustment            sbr/HFAdjustment.java:20-415       (envelope estimate per band, gains per envelope, assembly per band)
//   64-band QMF synthesis    sbr/SynthesisFilterbank64.java:9-79 (one thread per slot for the two DCT-IVs, 64 threads window)
//   Math.round / clamp / interleave as S/SampleBuffer.java:168-209
// Every floating-point operation is the binary32 operation of the Java code on the same operands in the same order
// (this file is compiled with --fmad=false), so the PCM is bit-identical to the reference's, like the AAC-LC path.
#pragma once
#include "jaadb_types.cuh"
#include "k2_filterbank.cuh"
#include "sbr_types.cuh"

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

constexpr int kK4Threads = 64;
// shared-memory carve (floats)
// Row strides are odd so that the one-thread-per-time-slot phases (analysis, synthesis DCTs), where the 32 lanes of a
// warp address the same column of 32 different rows, spread over all 32 banks.
constexpr int kXsStride = 129;              // one Xsbr slot: 64 bands x (re, im) + 1
constexpr int kVbStride = 129;              // one synthesis v-vector: 128 + 1
constexpr int kK4Xs = 40 * kXsStride;       // Xsbr
constexpr int kK4In = 288 + 1024 + 41 + 3;  // analysis input: history + this frame's core PCM, one pad float per 32 (+3: 16 B)
constexpr int kK4V = (9 + 32) * kVbStride + 3;  // synthesis v-vectors: 9 carried + 32 new (+3: keeps what follows 16 B aligned)
static_assert(kK4Xs % 4 == 0 && kK4In % 4 == 0 && kK4V % 4 == 0, "K4 shared-memory regions must stay 16-byte aligned");
#define XS(l, k, c) xs[(l) * kXsStride + (k) * 2 + (c)]
#define VB(s, r) vb[(s) * kVbStride + (r)]
#define INB(i) inbuf[(i) + ((i) >> 5)]
constexpr int kK4Adj = 3 * kSbrMaxLE * 64 + kSbrMaxLE * 64;   // G_lim_boost, Q_M_lim_boost, S_M_boost, E_curr
__host__ __device__ constexpr size_t k4_smem_bytes() {
  return sizeof(float) * (kK4Xs + kK4In + kK4V + kK4Adj + 16) + sizeof(SbrFrameDev);
}

struct K4RunDev {
  int32_t stream_slot;
  uint32_t first, count;    // into run_frames
  uint32_t sbr_base;        // first SbrFrameDev pair of the element's run
  uint8_t chan;             // channel inside the element (0/1)
  uint8_t ch_slot;          // channel slot inside the stream (core PCM / persistent state index)
  uint8_t out_ch;           // output channel index
  uint8_t n_out;            // output channels of the stream
  uint8_t dup;              // mono element: copy the result to the second output channel (SBR1.process)
  uint8_t pad[3];
};

template <int PCM_FORMAT>
__global__ void __launch_bounds__(kK4Threads)
k4_sbr_process_kernel(const K4RunDev* __restrict__ runs, const RunFrameDev* __restrict__ run_frames,
                      const SbrFrameDev* __restrict__ sframes, const float* __restrict__ core, SbrChanDev* __restrict__ chans,
                      uint8_t* __restrict__ pcm, const uint64_t* __restrict__ pcm_off, uint32_t* __restrict__ pcm_bytes_out,
                      SbrTablesDev T) {
  extern __shared__ __align__(16) float k4_smem[];
  float* xs = k4_smem;
  float* inbuf = k4_smem + kK4Xs;
  float* vb = inbuf + kK4In;
  float* adj = inbuf + kK4In + kK4V;
  float (*G_lim_boost)[64] = reinterpret_cast<float (*)[64]>(adj);
  float (*Q_M_lim_boost)[64] = reinterpret_cast<float (*)[64]>(adj + kSbrMaxLE * 64);
  float (*S_M_boost)[64] = reinterpret_cast<float (*)[64]>(adj + 2 * kSbrMaxLE * 64);
  float (*E_curr)[64] = reinterpret_cast<float (*)[64]>(adj + 3 * kSbrMaxLE * 64);   // [envelope][m]
  float* s_bw = adj + kK4Adj;                                                           // bwArray[8] (+8 spare)
  SbrFrameDev* fp = reinterpret_cast<SbrFrameDev*>(s_bw + 16);

  const K4RunDev run = runs[blockIdx.x];
  const int t = threadIdx.x;
  SbrChanDev* st = chans + (size_t)run.stream_slot * kSbrChansPerStream + run.ch_slot;

  // ---- persistent state in
  for (int i = t; i < 288; i += kK4Threads) INB(i) = st->ana_hist[i];
  for (int i = t; i < kSbrHfGen * 128; i += kK4Threads) XS(i >> 7, 0, i & 127) = (&st->xsbr[0][0][0])[i];
  for (int i = t; i < (40 - kSbrHfGen) * 128; i += kK4Threads) XS(kSbrHfGen + (i >> 7), 0, i & 127) = 0.f;
  // carried v-vectors: row 8 is the newest (slot -1), row 0 the oldest (slot -9); syn_v[0] = newest
  for (int i = t; i < 9 * 128; i += kK4Threads) VB(8 - i / 128, i % 128) = st->syn_v[i / 128][i % 128];
  float Gt[5], Qt[5];   // smoothing ring of band m = t
#pragma unroll
  for (int n = 0; n < 5; ++n) { Gt[n] = st->G_temp_prev[n][t]; Qt[n] = st->Q_temp_prev[n][t]; }
  int ring_index = st->GQ_ringbuf_index;
  int index_noise_prev = st->index_noise_prev, psi_is_prev = st->psi_is_prev;
  float bw_prev = 0.f;
  int invf_prev = 0;
  if (t < 8) { bw_prev = st->bwArray_prev[t]; invf_prev = st->bs_invf_mode_prev[t]; }
  float qc[10];
#pragma unroll
  for (int j = 0; j < 10; ++j) qc[j] = T.qmf_c[t + 64 * j];
  __syncthreads();

  for (uint32_t it = 0; it < run.count; ++it) {
    const RunFrameDev rf = run_frames[run.first + it];
    const uint32_t f = rf.frame;
    // frame record + core PCM -> shared
    {
      const uint4* src = reinterpret_cast<const uint4*>(sframes + ((size_t)run.sbr_base + it) * 2 + run.chan);
      uint4* dst = reinterpret_cast<uint4*>(fp);
      for (int i = t; i < (int)(sizeof(SbrFrameDev) / 16); i += kK4Threads) dst[i] = src[i];
      const float4* cs = reinterpret_cast<const float4*>(core + ((size_t)rf.ics_base + run.ch_slot) * 1024);
      for (int i = t; i < 256; i += kK4Threads) {
        const float4 v = cs[i];
        INB(288 + 4 * i) = v.x; INB(288 + 4 * i + 1) = v.y; INB(288 + 4 * i + 2) = v.z; INB(288 + 4 * i + 3) = v.w;
      }
    }
    __syncthreads();
    const int mode = fp->mode;
    uint8_t* dst = pcm + pcm_off[f];
    const int n_out = run.n_out;
    if (fp->frame_status != 0) {
      if (t == 0 && run.out_ch == 0) pcm_bytes_out[f] = 0;
      __syncthreads();
      continue;
    }
    auto put_sample = [&](int i, float v) {
      if (PCM_FORMAT == 2) {
        float* d = reinterpret_cast<float*>(dst);
        d[(size_t)run.out_ch * 2048 + i] = v;
        if (run.dup) d[(size_t)(run.out_ch + 1) * 2048 + i] = v;
      } else {
        uint32_t u = (uint32_t)pcm_round(v) & 0xFFFFu;
        if (PCM_FORMAT == 1) u = __byte_perm(u, 0, 0x4401);
        uint16_t* d = reinterpret_cast<uint16_t*>(dst);
        d[(size_t)i * n_out + run.out_ch] = (uint16_t)u;
        if (run.dup) d[(size_t)i * n_out + run.out_ch + 1] = (uint16_t)u;
      }
    };
    if (t == 0 && run.out_ch == 0) pcm_bytes_out[f] = (uint32_t)(2048 * n_out * (PCM_FORMAT == 2 ? 4 : 2));

    if (mode == 0) {
      // no valid SBR data in this frame: SBR.upsample (sbr/SBR.java:302-309; sample 1 keeps the core value)
      for (int i = t; i < 2048; i += kK4Threads) put_sample(i, i < 2 ? INB(288 + i) : INB(288 + (i >> 1)));
      __syncthreads();
      continue;
    }

    const int kx = mode == 2 ? fp->kx : 32;
    // ---- 32-band QMF analysis: thread l < 32 computes time slot l
    if (t < 32) {
      const int xi = 288 + 32 * t + 31;   // newest sample of the slot; sample xi - j = v[v_index + j] of the reference
      float in_real[32], in_imag[32], out_real[32], out_imag[32];
#pragma unroll
      for (int n = 0; n < 64; ++n) {
        const float u = (INB(xi - n) * c_sbr_qmf_c[2 * n]) + (INB(xi - (n + 64)) * c_sbr_qmf_c[2 * (n + 64)]) +
                        (INB(xi - (n + 128)) * c_sbr_qmf_c[2 * (n + 128)]) + (INB(xi - (n + 192)) * c_sbr_qmf_c[2 * (n + 192)]) +
                        (INB(xi - (n + 256)) * c_sbr_qmf_c[2 * (n + 256)]);
        // reordering of AnalysisFilterbank.java:40-47
        if (n == 0) in_real[0] = u;
        else if (n == 1) in_imag[31] = u;
        else if (n <= 31) in_imag[32 - n] = u;          // in_imag[31-(n-1)] = u[n]
        else if (n == 32) in_imag[0] = u;
        else if (n == 33) in_real[31] = -u;
        else in_real[64 - n] = -u;                      // in_real[m] = -u[64-m], m = 1..30
      }
      sbr_dct4_kernel(in_real, in_imag, out_real, out_imag);
      const int xrow = t + kSbrHfGen;
#pragma unroll
      for (int n = 0; n < 16; n++) {
        if (2 * n + 1 < kx) {
          XS(xrow, 2 * n, 0) = 2.0f * out_real[n];
          XS(xrow, 2 * n, 1) = 2.0f * out_imag[n];
          XS(xrow, 2 * n + 1, 0) = -2.0f * out_imag[31 - n];
          XS(xrow, 2 * n + 1, 1) = -2.0f * out_real[31 - n];
        } else {
          if (2 * n < kx) { XS(xrow, 2 * n, 0) = 2.0f * out_real[n]; XS(xrow, 2 * n, 1) = 2.0f * out_imag[n]; }
          else { XS(xrow, 2 * n, 0) = 0; XS(xrow, 2 * n, 1) = 0; }
          XS(xrow, 2 * n + 1, 0) = 0;
          XS(xrow, 2 * n + 1, 1) = 0;
        }
      }
    }
    __syncthreads();

    int first_slot = 0;   // t_E[0]
    if (mode == 2) {
      const int L_E = fp->L_E, M = fp->M;
      first_slot = fp->t_E[0];
      const int last_slot = fp->t_E[L_E];
      // ---- HF generation (HFGeneration.java)
      // calc_chirp_factors (:230-245): thread i < N_Q
      if (t < 8) {
        float bw = 0.f;
        if (t < fp->N_Q) {
          const int mode_i = fp->bs_invf_mode[t];
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
        s_bw[t] = bw;
      }
      __syncthreads();
      // one thread per generated band: band x of the concatenated patches
      if (kx + t < 64) {
        int i = 0, x = t, k = kx + t;
        while (i < fp->noPatches && x >= fp->patchNoSubbands[i]) { x -= fp->patchNoSubbands[i]; ++i; }
        if (i < fp->noPatches) {
          const int p = fp->patchStartSubband[i] + x;
          const int g = fp->table_map_k_to_g[k];
          const float bw = s_bw[g];
          const float bw2 = bw * bw;
          const int offset = kSbrHfAdj;
          if (bw2 > 0) {
            // calc_prediction_coef / auto_correlation (:100-204), len = numTimeSlotsRate + 6
            float r01r = 0, r01i = 0, r02r = 0, r02i = 0, r11r = 0;
            float temp1_r, temp1_i, temp2_r, temp2_i, temp3_r, temp3_i, temp4_r, temp4_i, temp5_r, temp5_i;
            const float rel = 1.0f / (1 + 1e-6f);
            temp2_r = XS(offset - 2, p, 0); temp2_i = XS(offset - 2, p, 1);
            temp3_r = XS(offset - 1, p, 0); temp3_i = XS(offset - 1, p, 1);
            temp4_r = temp2_r; temp4_i = temp2_i; temp5_r = temp3_r; temp5_i = temp3_i;
            temp1_r = 0; temp1_i = 0;
            for (int j = offset; j < kSbrSlots + 6 + offset; j++) {
              temp1_r = temp2_r; temp1_i = temp2_i;
              temp2_r = temp3_r; temp2_i = temp3_i;
              temp3_r = XS(j, p, 0); temp3_i = XS(j, p, 1);
              r01r += temp3_r * temp2_r + temp3_i * temp2_i;
              r01i += temp3_i * temp2_r - temp3_r * temp2_i;
              r02r += temp3_r * temp1_r + temp3_i * temp1_i;
              r02i += temp3_i * temp1_r - temp3_r * temp1_i;
              r11r += temp2_r * temp2_r + temp2_i * temp2_i;
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
            const float a0_r = (al0r * bw), a1_r = (al1r * bw2), a0_i = (al0i * bw), a1_i = (al1i * bw2);
            temp2_r = XS(first_slot - 2 + offset, p, 0); temp3_r = XS(first_slot - 1 + offset, p, 0);
            temp2_i = XS(first_slot - 2 + offset, p, 1); temp3_i = XS(first_slot - 1 + offset, p, 1);
            for (int l = first_slot; l < last_slot; l++) {
              temp1_r = temp2_r; temp2_r = temp3_r; temp3_r = XS(l + offset, p, 0);
              temp1_i = temp2_i; temp2_i = temp3_i; temp3_i = XS(l + offset, p, 1);
              XS(l + offset, k, 0) = temp3_r + ((a0_r * temp2_r) - (a0_i * temp2_i) + (a1_r * temp1_r) - (a1_i * temp1_i));
              XS(l + offset, k, 1) = temp3_i + ((a0_i * temp2_r) + (a0_r * temp2_i) + (a1_i * temp1_r) + (a1_r * temp1_i));
            }
          } else {
            for (int l = first_slot; l < last_slot; l++) {
              XS(l + offset, k, 0) = XS(l + offset, p, 0);
              XS(l + offset, k, 1) = XS(l + offset, p, 1);
            }
          }
        }
      }
      __syncthreads();

      // ---- HF adjustment (HFAdjustment.java)
      // `new HFAdjustment()` per call: the boost arrays start from zero.  The limiter table does not always reach M
      // (FBT.limiter_frequency_table sorts a shrinking prefix), and bands it leaves out keep gain 0.
      for (int i = t; i < 3 * kSbrMaxLE * 64; i += kK4Threads) adj[i] = 0.f;
      // estimate_current_envelope (:78-131): thread m
      if (t < M) {
        for (int l = 0; l < L_E; l++) {
          const int l_i = fp->t_E[l], u_i = fp->t_E[l + 1];
          float nrg = 0, div;
          if (fp->interpol_freq) {
            div = (float)(u_i - l_i);
            if (div == 0) div = 1;
            for (int i = l_i + kSbrHfAdj; i < u_i + kSbrHfAdj; i++)
              nrg += (XS(i, t + kx, 0) * XS(i, t + kx, 0)) + (XS(i, t + kx, 1) * XS(i, t + kx, 1));
          } else {
            // the band of the envelope's resolution that holds k = t + kx
            const int res = fp->f[l], nb = res ? fp->N_high : fp->N_low;
            int p = 0;
            while (p + 1 < nb && fp->f_table_res[res][p + 1] <= t + kx) ++p;
            const int k_l = fp->f_table_res[res][p], k_h = fp->f_table_res[res][p + 1];
            div = (float)((u_i - l_i) * (k_h - k_l));
            if (div == 0) div = 1;
            for (int i = l_i + kSbrHfAdj; i < u_i + kSbrHfAdj; i++)
              for (int j = k_l; j < k_h; j++) nrg += (XS(i, j, 0) * XS(i, j, 0)) + (XS(i, j, 1) * XS(i, j, 1));
          }
          E_curr[l][t] = nrg / div;
        }
      }
      __syncthreads();
      // calculate_gain (:242-415): thread l < L_E runs its envelope
      if (t < L_E) {
        const int l = t;
        const float EPS = 1e-12f;
        const int l_A = fp->l_A;
        const int res = fp->f[l];
        const bool flag_prev = fp->add_harmonic_flag_prev != 0;
        auto get_S_mapped = [&](int current_band) -> int {   // :46-76
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
        // the noise-floor time band of envelope l: current_t_noise_band advances once per envelope whose end passes t_Q
        int current_t_noise_band = 0;
        for (int ll = 0; ll <= l; ++ll)
          if (fp->t_E[ll + 1] > fp->t_Q[current_t_noise_band + 1]) current_t_noise_band++;
        int current_f_noise_band = 0, current_res_band = 0, current_res_band2 = 0, current_hi_res_band = 0;
        const float delta = (l == l_A || l == fp->prevEnvIsShort) ? 0.f : 1.f;
        int S_mapped = get_S_mapped(current_res_band2);
        float limg;
        switch (fp->limiter_gains) { case 0: limg = 0.5f; break; case 1: limg = 1.0f; break; case 2: limg = 2.0f; break; default: limg = 1e10f; break; }
        for (int k = 0; k < fp->N_L; k++) {
          float den = 0, acc1 = 0, acc2 = 0;
          const int ml1 = fp->f_table_lim[k], ml2 = fp->f_table_lim[k + 1];
          for (int m = ml1; m < ml2; m++) {
            if ((m + kx) == fp->f_table_res[res][current_res_band + 1]) current_res_band++;
            acc1 += fp->E_orig[l][current_res_band];
            acc2 += E_curr[l][m];
          }
          float G_max = ((EPS + acc1) / (EPS + acc2)) * limg;
          G_max = fminf(G_max, 1e10f);
          for (int m = ml1; m < ml2; m++) {
            if ((m + kx) == fp->f_table_noise[current_f_noise_band + 1]) current_f_noise_band++;
            if ((m + kx) == fp->f_table_res[res][current_res_band2 + 1]) {
              current_res_band2++;
              S_mapped = get_S_mapped(current_res_band2);
            }
            if ((m + kx) == fp->f_table_res[SBR_HI_RES][current_hi_res_band + 1]) current_hi_res_band++;
            int S_index_mapped = 0;
            if ((l >= l_A) || (fp->bs_add_harmonic_prev[current_hi_res_band] != 0 && flag_prev)) {
              if ((m + kx) == (fp->f_table_res[SBR_HI_RES][current_hi_res_band + 1] + fp->f_table_res[SBR_HI_RES][current_hi_res_band]) >> 1)
                S_index_mapped = fp->bs_add_harmonic[current_hi_res_band];
            }
            const float Q_div = fp->Q_div[current_t_noise_band][current_f_noise_band];
            const float Q_div2 = fp->Q_div2[current_t_noise_band][current_f_noise_band];
            const float E_o = fp->E_orig[l][current_res_band2];
            const float Q_M = E_o * Q_div2;
            float S_M;
            if (S_index_mapped == 0) S_M = 0;
            else { S_M = E_o * Q_div; den += S_M; }
            float G = E_o / (1.0f + E_curr[l][m]);
            if ((S_mapped == 0) && (delta == 1)) G *= Q_div;
            else if (S_mapped == 1) G *= Q_div2;
            float Q_M_lim, G_lim;
            if (G_max > G) { Q_M_lim = Q_M; G_lim = G; }
            else { Q_M_lim = Q_M * G_max / G; G_lim = G_max; }
            den += E_curr[l][m] * G_lim;
            if ((S_index_mapped == 0) && (l != l_A)) den += Q_M_lim;
            // park the un-boosted values; the boost needs the whole limiter band's `den`
            G_lim_boost[l][m] = G_lim;
            Q_M_lim_boost[l][m] = Q_M_lim;
            S_M_boost[l][m] = S_M;
          }
          float G_boost = (acc1 + EPS) / (den + EPS);
          G_boost = fminf(G_boost, 2.51188643f);
          for (int m = ml1; m < ml2; m++) {
            // (float) Math.sqrt(float product): the double square root of a binary32 value, rounded to binary32, is the
            // correctly rounded binary32 square root
            G_lim_boost[l][m] = __fsqrt_rn(G_lim_boost[l][m] * G_boost);
            Q_M_lim_boost[l][m] = __fsqrt_rn(Q_M_lim_boost[l][m] * G_boost);
            const float sm = S_M_boost[l][m];
            S_M_boost[l][m] = (sm != 0) ? __fsqrt_rn(sm * G_boost) : 0.f;
          }
        }
      }
      __syncthreads();
      // hf_assembly (:133-240): thread m walks the slots; the 5-entry smoothing ring of its band lives in registers
      {
        const bool active = t < M;
        int fIndexNoise = fp->reset ? 0 : index_noise_prev;
        int fIndexSine = psi_is_prev;
        bool assembly_reset = fp->reset != 0;
        int slots_done = 0;
        for (int l = 0; l < L_E; l++) {
          const bool no_noise = (l == fp->l_A || l == fp->prevEnvIsShort);
          int h_SL = fp->smoothing_mode ? 0 : 4;
          h_SL = no_noise ? 0 : h_SL;
          const float g_new = active ? G_lim_boost[l][t] : 0.f, q_new = active ? Q_M_lim_boost[l][t] : 0.f;
          const float s_m = active ? S_M_boost[l][t] : 0.f;
          // System.arraycopy(.., 0, .., 0, sbr.M): ring entries of bands >= M keep their old contents
          if (assembly_reset) {
            if (active) {
#pragma unroll
              for (int n = 0; n < 4; ++n) { Gt[n] = g_new; Qt[n] = q_new; }
            }
            ring_index = 4;
            assembly_reset = false;
          }
          for (int i = fp->t_E[l]; i < fp->t_E[l + 1]; i++) {
            if (active) {
#pragma unroll
              for (int n = 0; n < 5; ++n) if (n == ring_index) { Gt[n] = g_new; Qt[n] = q_new; }
            }
            float G_filt = 0, Q_filt = 0;
            if (h_SL != 0) {
              int ri = ring_index;
#pragma unroll
              for (int n = 0; n <= 4; n++) {
                const float h = n == 0 ? 0.03183050093751f : n == 1 ? 0.11516383427084f : n == 2 ? 0.21816949906249f
                              : n == 3 ? 0.30150283239582f : 0.33333333333333f;
                ri++;
                if (ri >= 5) ri -= 5;
                float gv = Gt[0], qv = Qt[0];
#pragma unroll
                for (int z = 1; z < 5; ++z) if (z == ri) { gv = Gt[z]; qv = Qt[z]; }
                G_filt += (gv * h);
                Q_filt += (qv * h);
              }
            } else {
#pragma unroll
              for (int z = 0; z < 5; ++z) if (z == ring_index) { G_filt = Gt[z]; Q_filt = Qt[z]; }
            }
            Q_filt = (s_m != 0 || no_noise) ? 0 : Q_filt;
            if (active) {
              const int ni = (fIndexNoise + slots_done * M + t + 1) & 511;
              float* x = &XS(i + kSbrHfAdj, t + kx, 0);
              x[0] = G_filt * x[0] + (Q_filt * __ldg(T.noise_table + 2 * ni));
              x[1] = G_filt * x[1] + (Q_filt * __ldg(T.noise_table + 2 * ni + 1));
              const int rev = (((t + kx) & 1) != 0 ? -1 : 1);
              const int phi_re = fIndexSine == 0 ? 1 : (fIndexSine == 2 ? -1 : 0);
              const int phi_im = fIndexSine == 1 ? 1 : (fIndexSine == 3 ? -1 : 0);
              x[0] += s_m * (float)phi_re;
              x[1] += (float)rev * s_m * (float)phi_im;
            }
            ++slots_done;
            fIndexSine = (fIndexSine + 1) & 3;
            ring_index++;
            if (ring_index >= 5) ring_index = 0;
          }
        }
        index_noise_prev = (fIndexNoise + slots_done * M) & 511;
        psi_is_prev = fIndexSine;
      }
      __syncthreads();
    }

    // ---- 64-band QMF synthesis.  X[l][k] = Xsbr[l + tHFAdj][k] below kx + M of the slot's frame, 0 above
    // (Channel.process_channel, :604-645)
    if (t < 32) {
      const int l = t;
      int lim;
      if (mode == 2) lim = (l < first_slot) ? (fp->kx_prev + fp->M_prev) : (fp->kx + fp->M);
      else lim = 32;
      const float scale = 1.f / 64.f;
      auto Xr = [&](int k) -> float { return k < lim ? XS(l + kSbrHfAdj, k, 0) : 0.f; };
      auto Xi = [&](int k) -> float { return k < lim ? XS(l + kSbrHfAdj, k, 1) : 0.f; };
      float in_r[32], in_i[32], o1r[32], o1i[32], o2r[32], o2i[32];
      in_i[31] = scale * Xr(1);
      in_r[0] = scale * Xr(0);
#pragma unroll
      for (int k = 1; k < 31; k++) { in_i[31 - k] = scale * Xr(2 * k + 1); in_r[k] = scale * Xr(2 * k); }
      in_i[0] = scale * Xr(63);
      in_r[31] = scale * Xr(62);
      sbr_dct4_kernel(in_r, in_i, o1r, o1i);
      in_i[31] = scale * Xi(63 - 1);
      in_r[0] = scale * Xi(63 - 0);
#pragma unroll
      for (int k = 1; k < 31; k++) { in_i[31 - k] = scale * Xi(63 - (2 * k + 1)); in_r[k] = scale * Xi(63 - (2 * k)); }
      in_i[0] = scale * Xi(63 - 63);
      in_r[31] = scale * Xi(63 - 62);
      sbr_dct4_kernel(in_r, in_i, o2r, o2i);
      float* v = &VB(9 + l, 0);
#pragma unroll
      for (int n = 0; n < 32; n++) {
        v[2 * n] = o2r[n] - o1r[n];
        v[127 - 2 * n] = o2r[n] + o1r[n];
        v[2 * n + 1] = o2i[31 - n] + o1i[31 - n];
        v[127 - (2 * n + 1)] = o2i[31 - n] - o1i[31 - n];
      }
    }
    __syncthreads();
    // window + output: thread k, all 32 slots
    for (int l = 0; l < 32; ++l) {
      const int cur = 9 + l;
      float o = (VB(cur, t) * qc[0]);
#pragma unroll
      for (int j = 1; j < 10; ++j) o = o + (VB(cur - j, t + 64 * (j & 1)) * qc[j]);
      put_sample(64 * l + t, o);
    }
    __syncthreads();
    // ---- carry: analysis history, the last 8 Xsbr slots (sbr_save_matrix), the last 9 v-vectors
    for (int i = t; i < 288; i += kK4Threads) INB(i) = INB(1024 + i);   // disjoint ranges
    {
      float keep[kSbrHfGen * 2];
#pragma unroll
      for (int i = 0; i < kSbrHfGen; ++i) { keep[2 * i] = XS(i + kSbrSlots, t, 0); keep[2 * i + 1] = XS(i + kSbrSlots, t, 1); }
      float vk[18];
#pragma unroll
      for (int s = 0; s < 9; ++s) { vk[2 * s] = VB(32 + s, t); vk[2 * s + 1] = VB(32 + s, t + 64); }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < kSbrHfGen; ++i) { XS(i, t, 0) = keep[2 * i]; XS(i, t, 1) = keep[2 * i + 1]; }
      for (int i = kSbrHfGen; i < 40; ++i) { XS(i, t, 0) = 0.f; XS(i, t, 1) = 0.f; }
#pragma unroll
      for (int s = 0; s < 9; ++s) { VB(s, t) = vk[2 * s]; VB(s, t + 64) = vk[2 * s + 1]; }
    }
    __syncthreads();
  }

  // ---- persistent state out
  for (int i = t; i < 288; i += kK4Threads) st->ana_hist[i] = INB(i);
  for (int i = t; i < kSbrHfGen * 128; i += kK4Threads) (&st->xsbr[0][0][0])[i] = XS(i >> 7, 0, i & 127);
  for (int i = t; i < 9 * 128; i += kK4Threads) st->syn_v[i / 128][i % 128] = VB(8 - i / 128, i % 128);
#pragma unroll
  for (int n = 0; n < 5; ++n) { st->G_temp_prev[n][t] = Gt[n]; st->Q_temp_prev[n][t] = Qt[n]; }
  if (t < 8) { st->bwArray_prev[t] = bw_prev; st->bs_invf_mode_prev[t] = (uint8_t)invf_prev; }
  if (t == 0) { st->GQ_ringbuf_index = ring_index; st->index_noise_prev = index_noise_prev; st->psi_is_prev = psi_is_prev; }
}

}  // namespace jaadb
