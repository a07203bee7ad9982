// K2 -- spectral processing.  One CTA owns one stream and walks that stream's
// frames of the batch in decode order, 64 threads per channel:
//   dequantisation (|q|^(4/3) LUT x 2^((sf-100)/4) LUT, ICStream.java:264-269)
//   -> M/S (tools/MS.java:17-41) -> intensity stereo (tools/IS.java:17-53)
//   -> IMDCT as an N/4-point complex FFT in registers + shared memory
//      (filterbank/MDCT.java:36-81, FFT.java:48-135)
//   -> sine/KBD windowing + overlap-add (filterbank/FilterBank.java:39-123)
//   -> Math.round / clamp / interleave to int16 (S/SampleBuffer.java:168-209).
// The overlap buffers of the stream stay in shared memory for the whole run and
// touch HBM once per call, window_shape[PREVIOUS] is carried in a register.
//
// Bit-exactness: every floating-point operation is the same binary32 operation,
// on the same operands, as in the Java code (separate multiply and add, no FMA:
// this file must be compiled with --fmad=false).  The FFT keeps JAAD's butterfly
// graph -- bit reversal, one radix-4 stage without twiddles, then radix-2 stages --
// and only changes which thread evaluates which butterfly, so the float PCM is
// bit-identical to JAAD's, not merely within tolerance.
#pragma once
#include "jaadb_types.cuh"

namespace jaadb {

constexpr int kThreadsPerChannel = 64;
// shared memory per channel (floats): spectrum / FFT exchange (1024 + pad), overlap (1024)
constexpr int kSpecStride = 1024 + 32;  // one pad float per 32 keeps the bit-reversed gather conflict-free
constexpr int kXchgStride = 8 * 72;     // 8 blocks of 8x8 complex, rows padded to 9 (re and im planes)

__device__ __forceinline__ int spec_addr(int i) { return i + (i >> 5); }
__device__ __forceinline__ int brev3(int j) { return ((j & 1) << 2) | (j & 2) | ((j >> 2) & 1); }

struct Cplx { float re, im; };

// bottom radix-4 round of FFT.process, inverse direction (FFT.java:68-106)
__device__ __forceinline__ void radix4_inv(Cplx& x0, Cplx& x1, Cplx& x2, Cplx& x3) {
  float aRe = x0.re + x1.re, aIm = x0.im + x1.im;
  float bRe = x2.re + x3.re, bIm = x2.im + x3.im;
  float cRe = x0.re - x1.re, cIm = x0.im - x1.im;
  float dRe = x2.re - x3.re, dIm = x2.im - x3.im;
  x0.re = aRe + bRe; x0.im = aIm + bIm;
  x2.re = aRe - bRe; x2.im = aIm - bIm;
  float e1Re = cRe - dIm, e1Im = cIm + dRe;
  float e2Re = cRe + dIm, e2Im = cIm - dRe;
  x1.re = e1Re; x1.im = e1Im;
  x3.re = e2Re; x3.im = e2Im;
}

// one radix-2 butterfly of FFT.process (FFT.java:113-134)
__device__ __forceinline__ void bfly(Cplx& v0, Cplx& v1, float rootRe, float rootIm) {
  float zRe = v1.re * rootRe - v1.im * rootIm;
  float zIm = v1.re * rootIm + v1.im * rootRe;
  v1.re = v0.re - zRe;
  v1.im = v0.im - zIm;
  v0.re = v0.re + zRe;
  v0.im = v0.im + zIm;
}

// MDCT.process reorder (MDCT.java:60-80): time sample m (0 <= m < N) from the post-twiddled buffer.
// buf is stored as planes re[n], im[n].
__device__ __forceinline__ float mdct_out(const float* __restrict__ re, const float* __restrict__ im, int N4, int N8, int m) {
  const int quarter = m / N4, mm = m - quarter * N4;
  const int h = mm >> 1;
  if (mm & 1) {
    switch (quarter) {
      case 0: return -re[N8 - 1 - h];
      case 1: return -im[N4 - 1 - h];
      case 2: return -im[N8 - 1 - h];
      default: return re[N4 - 1 - h];
    }
  } else {
    switch (quarter) {
      case 0: return im[N8 + h];
      case 1: return re[h];
      case 2: return re[N8 + h];
      default: return -im[h];
    }
  }
}

// Java Math.round(float) + SampleBuffer clamp (S/SampleBuffer.java:193-205)
__device__ __forceinline__ int pcm_round(float x) {
  if (x != x) return 0;
  x = fminf(fmaxf(x, -40000.f), 40000.f);
  float f = floorf(x);
  int r = (int)f + ((x - f) >= 0.5f ? 1 : 0);
  return min(max(r, -32768), 32767);
}

struct ChanCtx {
  const IcsSide* side;      // shared-memory copy
  const int16_t* q;         // global
};

// dequantised value of coefficient i (window de-interleaved index, as ICStream.iqData) of one channel before the
// stereo tools; cb_out gets the band's codebook.  K1 leaves q in bitstream order: group g starts at
// 128 * first_window(g), band sfb at glen * swb[sfb], glen windows x width coefficients, window-major.
// win_info[w] = group | first window of the group << 8 | group length << 16.
__device__ __forceinline__ float dequant_at(const IcsSide* __restrict__ s, const int16_t* __restrict__ q,
                                            const TablesDev& T, const uint8_t* __restrict__ sfb_of,
                                            const int16_t* __restrict__ swb_short, int i,
                                            const uint32_t* __restrict__ win_info, int& cb_out, int& idx_out) {
  int sfb, g = 0, qpos = i;
  const bool sh = s->window_sequence == 2;
  uint32_t wi = 0;
  if (sh) { sfb = sfb_of[i & 127]; wi = win_info[i >> 7]; g = (int)(wi & 255u); }
  else sfb = sfb_of[i];
  cb_out = 0;
  idx_out = 0;
  if (sfb >= s->max_sfb) return 0.f;
  const int idx = g * s->max_sfb + sfb;
  idx_out = idx;
  const int cb = s->sfb_cb[idx];
  cb_out = cb;
  if (cb == 0 || cb > 11) return 0.f;
  if (sh) {
    const int gstart = (int)((wi >> 8) & 255u), glen = (int)(wi >> 16);
    const int lo = swb_short[sfb], width = swb_short[sfb + 1] - lo;
    qpos = 128 * gstart + glen * lo + ((i >> 7) - gstart) * width + ((i & 127) - lo);
  }
  const int v = q[qpos];
  const float sf = __ldg(T.sf + s->sf_idx[idx]);
  const float m = __ldg(T.iq + (v < 0 ? -v : v));
  // iqData = (v>0) ? IQ[v] : -IQ[-v]; iqData *= scaleFactors[idx]   (ICStream.java:266-267)
  return (v > 0 ? m : -m) * sf;
}

template <int PCM_FORMAT>
__global__ void k2_filterbank_kernel(const RunDev* __restrict__ runs, const uint32_t* __restrict__ run_frames,
                                     const FrameDev* __restrict__ frames, const FrameSide* __restrict__ fside,
                                     const IcsSide* __restrict__ iside, const int16_t* __restrict__ qall,
                                     float* __restrict__ overlap_all, StreamState* __restrict__ sstate,
                                     uint8_t* __restrict__ pcm, const uint64_t* __restrict__ pcm_off,
                                     uint32_t* __restrict__ pcm_bytes_out, float* __restrict__ spec_tap,
                                     TablesDev T, const LayoutDev* __restrict__ layouts, int nch) {
  extern __shared__ __align__(16) float smem[];
  // carve: per channel [spec kSpecStride][overlap 1024][xre kXchgStride][xim kXchgStride]; then sides, pcm staging
  const int per_ch = kSpecStride + 1024 + 2 * kXchgStride;
  float* s_fft_tw = smem;                         // fft512 re/im (inverse) [256][2] + fft64 [32][2]
  float* s_ch = s_fft_tw + 2 * 256 + 2 * 32;
  IcsSide* s_side = reinterpret_cast<IcsSide*>(s_ch + nch * per_ch);
  uint32_t* s_wgroup = reinterpret_cast<uint32_t*>(s_side + nch);   // [nch][8] window -> group info
  int16_t* s_pcm = reinterpret_cast<int16_t*>(s_wgroup + 8 * kMaxChannels);  // [1024][out_ch] (s16 formats)

  const RunDev run = runs[blockIdx.x];
  const LayoutDev lay = layouts[run.layout];
  const int tid = threadIdx.x;
  const int c = tid / kThreadsPerChannel;         // channel slot of this thread
  const int t = tid - c * kThreadsPerChannel;
  const int nthreads = blockDim.x;
  const int out_ch = run.mono_dup ? 2 : nch;
  const int sf_index = run.sf_index;

  float* my_spec = s_ch + c * per_ch;
  float* my_ovl = my_spec + kSpecStride;
  float* my_xre = my_ovl + 1024;
  float* my_xim = my_xre + kXchgStride;

  // twiddles used by the radix-2 stages: roots[k*m] with k*m < length/2
  for (int i = tid; i < 256; i += nthreads) {
    s_fft_tw[2 * i] = T.fft512[3 * i];
    s_fft_tw[2 * i + 1] = T.fft512[3 * i + 1];
  }
  for (int i = tid; i < 32; i += nthreads) {
    s_fft_tw[512 + 2 * i] = T.fft64[2 * i];
    s_fft_tw[512 + 2 * i + 1] = T.fft64[2 * i + 1];
  }
  const float* tw512 = s_fft_tw;
  const float* tw64 = s_fft_tw + 512;

  // persistent state in: overlap + current window shapes
  float* g_ovl = overlap_all + ((size_t)run.stream_slot * kMaxChannels + c) * 1024;
  for (int i = t; i < 1024; i += kThreadsPerChannel) my_ovl[i] = g_ovl[i];
  int shape_cur = sstate[run.stream_slot].window_shape[c];
  __syncthreads();

  // element of this thread's channel
  int el_first = c, el_nch = 1;
  for (int e = 0; e < lay.n_elements; ++e) {
    int f0 = lay.el_first_ch[e];
    int n = lay.el_type[e] == EL_CPE ? 2 : 1;
    if (c >= f0 && c < f0 + n) { el_first = f0; el_nch = n; }
  }

  for (uint32_t it = 0; it < run.count; ++it) {
    const uint32_t f = run_frames[run.first + it];
    const FrameDev fr = frames[f];
    const int status = fside[f].status;
    // side info -> shared
    {
      const uint32_t* src = reinterpret_cast<const uint32_t*>(iside + fr.ics_base);
      uint32_t* dst = reinterpret_cast<uint32_t*>(s_side);
      const int nwords = nch * (int)(sizeof(IcsSide) / 4);
      for (int i = tid; i < nwords; i += nthreads) dst[i] = src[i];
    }
    __syncthreads();
    const IcsSide* sd = s_side + c;
    // window-shape bookkeeping of ICSInfo.decode / setCommonData (ICSInfo.java:90-91,196-197)
    int shape_prev = shape_cur;
    if (sd->info_decoded) { shape_prev = shape_cur; shape_cur = sd->window_shape; }
    if (t < 8) {
      // window -> group map for short frames
      int w = t, g = 0, acc = 0, gstart = 0;
      for (int k = 0; k < 8; ++k) { acc += sd->group_len[k]; if (w >= acc) { g = k + 1; gstart = acc; } }
      g = min(g, 7);
      s_wgroup[c * 8 + t] = (uint32_t)g | ((uint32_t)gstart << 8) | ((uint32_t)sd->group_len[g] << 16);
    }
    __syncthreads();
    if (status != 0) {
      if (tid == 0) pcm_bytes_out[f] = 0;
      __syncthreads();
      continue;  // the frame produced no PCM; overlap untouched (Decoder.java:96-98)
    }

    // ---- phase 1: dequantise + M/S + IS into my_spec (threads of an element cover all its channels)
    {
      const int ws = sd->window_sequence;
      const int el_threads = el_nch * kThreadsPerChannel;
      const int et = tid - el_first * kThreadsPerChannel;
      const int16_t* qL = qall + ((size_t)fr.ics_base + el_first) * 1024;
      const IcsSide* sL = s_side + el_first;
      float* specL = s_ch + el_first * per_ch;
      const int16_t* swb_sh = T.swb_short + sf_index * 17;
      if (el_nch == 1) {
        const uint8_t* sfb_of = (ws == 2) ? T.sfb_of_short + sf_index * 128 : T.sfb_of_long + sf_index * 1024;
        for (int i = et; i < 1024; i += el_threads) {
          int cb, idx;
          float v = dequant_at(sL, qL, T, sfb_of, swb_sh, i, s_wgroup + el_first * 8, cb, idx);
          specL[spec_addr(i)] = v;
          if (spec_tap) spec_tap[((size_t)fr.ics_base + el_first) * 1024 + i] = v;
        }
      } else {
        const IcsSide* sR = sL + 1;
        const int16_t* qR = qL + 1024;
        float* specR = specL + per_ch;
        const uint8_t* sfb_ofL = (sL->window_sequence == 2) ? T.sfb_of_short + sf_index * 128 : T.sfb_of_long + sf_index * 1024;
        const uint8_t* sfb_ofR = (sR->window_sequence == 2) ? T.sfb_of_short + sf_index * 128 : T.sfb_of_long + sf_index * 1024;
        const bool ms_on = sL->common_window && sL->ms_mask != 0;   // CPE.java:159-160
        const bool ms_present = sL->ms_mask != 0;                   // CPE.isMSMaskPresent
        for (int i = et; i < 1024; i += el_threads) {
          int cbL, idxL, cbR, idxR;
          float l = dequant_at(sL, qL, T, sfb_ofL, swb_sh, i, s_wgroup + el_first * 8, cbL, idxL);
          float r = dequant_at(sR, qR, T, sfb_ofR, swb_sh, i, s_wgroup + (el_first + 1) * 8, cbR, idxR);
          // MS.process: both codebooks < NOISE_HCB, band flagged (MS.java:28-36)
          if (ms_on && cbL < 13 && cbR < 13) {
            // idxL == idxR here (common window); bands above max_sfb have cb 0 but are never flagged
            const bool band_in = (sL->window_sequence == 2 ? sfb_ofL[i & 127] : sfb_ofL[i]) < sL->max_sfb;
            if (band_in && ((sL->ms_used[idxL >> 3] >> (idxL & 7)) & 1)) {
              float tt = l - r;
              l = l + r;
              r = tt;
            }
          }
          // IS.process: right channel bands with codebook 14/15 (IS.java:29-44)
          if (cbR == 15 || cbR == 14) {
            int sgn = cbR == 15 ? 1 : -1;
            if (ms_present) sgn *= ((sL->ms_used[idxR >> 3] >> (idxR & 7)) & 1) ? -1 : 1;
            const unsigned si = sR->sf_idx[idxR];
            float scale = __ldg(T.sf + si);
            if (sgn < 0) scale = -scale;
            r = l * scale;
          }
          specL[spec_addr(i)] = l;
          specR[spec_addr(i)] = r;
          if (spec_tap) {
            spec_tap[((size_t)fr.ics_base + el_first) * 1024 + i] = l;
            spec_tap[((size_t)fr.ics_base + el_first + 1) * 1024 + i] = r;
          }
        }
      }
    }
    __syncthreads();

    // ---- phase 2: IMDCT of this channel (thread t owns points 8t..8t+7 of the bit-reversed input)
    const int ws = sd->window_sequence;
    const bool is_short = ws == 2;
    Cplx a[8];
    {
      if (!is_short) {
        // pre-IFFT complex multiplication (MDCT.java:39-42), gathered in bit-reversed order (FFT.java:51-61)
        const int kb = (int)(__brev((unsigned)t) >> 26);  // bitrev6(t)
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = kb + 64 * brev3(j);
          const float x0 = my_spec[spec_addr(2 * k)];
          const float x1 = my_spec[spec_addr(1023 - 2 * k)];
          const float cs = __ldg(T.mdct_long + 2 * k), sn = __ldg(T.mdct_long + 2 * k + 1);
          a[j].im = (x0 * cs) + (x1 * sn);
          a[j].re = (x1 * cs) - (x0 * sn);
        }
      } else {
        const int w = t >> 3;                                  // short window handled by this thread
        const int kb = brev3(t & 7);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = kb + 8 * brev3(j);                     // bitrev6(8*(t&7)+j)
          const float x0 = my_spec[spec_addr(128 * w + 2 * k)];
          const float x1 = my_spec[spec_addr(128 * w + 127 - 2 * k)];
          const float cs = __ldg(T.mdct_short + 2 * k), sn = __ldg(T.mdct_short + 2 * k + 1);
          a[j].im = (x0 * cs) + (x1 * sn);
          a[j].re = (x1 * cs) - (x0 * sn);
        }
      }
      // stage A: radix-4 on (0..3), (4..7), then radix-2 stage i=4 with roots[k*m], m = length/8
      radix4_inv(a[0], a[1], a[2], a[3]);
      radix4_inv(a[4], a[5], a[6], a[7]);
      const float* tw = is_short ? tw64 : tw512;
      const int m4 = is_short ? 8 : 64;
#pragma unroll
      for (int k = 0; k < 4; ++k) bfly(a[k], a[k + 4], tw[2 * k * m4], tw[2 * k * m4 + 1]);
      // exchange 1: 8x8 transposes inside each 64-point block
      {
        const int blk = t >> 3, row = t & 7;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          my_xre[blk * 72 + row * 9 + j] = a[j].re;
          my_xim[blk * 72 + row * 9 + j] = a[j].im;
        }
      }
    }
    __syncthreads();
    {
      const int blk = t >> 3, col = t & 7;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[j].re = my_xre[blk * 72 + j * 9 + col];
        a[j].im = my_xim[blk * 72 + j * 9 + col];
      }
      // thread holds n = 64*blk + col + 8*j.  stage B: i = 8, 16, 32
      const float* tw = is_short ? tw64 : tw512;
      const int mB = is_short ? 4 : 32;   // m for i=8: length/16
#pragma unroll
      for (int j = 0; j < 8; j += 2) {    // i=8: pairs (j, j+1); k = n & 7 = col
        const int k = col;
        bfly(a[j], a[j + 1], tw[2 * k * mB], tw[2 * k * mB + 1]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {       // i=16: pairs (j, j+2), j&2==0; k = col + 8*(j&1)
        if (j & 2) continue;
        const int k = col + 8 * (j & 1);
        bfly(a[j], a[j + 2], tw[2 * k * (mB >> 1)], tw[2 * k * (mB >> 1) + 1]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {       // i=32: pairs (j, j+4); k = col + 8*j
        const int k = col + 8 * j;
        bfly(a[j], a[j + 4], tw[2 * k * (mB >> 2)], tw[2 * k * (mB >> 2) + 1]);
      }
    }
    __syncthreads();  // everyone finished reading exchange 1
    float* bre = my_spec;              // reuse: post-twiddled buffer as planes re[512] | im[512]
    float* bim = my_spec + 512;
    if (!is_short) {
      // exchange 2 through the same planes: write n = 64*blk + col + 8*j, read n = t + 64*j
      const int blk = t >> 3, col = t & 7;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = 64 * blk + col + 8 * j;
        my_xre[n + (n >> 5)] = a[j].re;   // kXchgStride = 576 >= 512 + 16
        my_xim[n + (n >> 5)] = a[j].im;
      }
    }
    __syncthreads();  // unconditional: channels of one CTA may mix long and short windows
    if (!is_short) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = t + 64 * j;
        a[j].re = my_xre[n + (n >> 5)];
        a[j].im = my_xim[n + (n >> 5)];
      }
      // stage C: i = 64, 128, 256 on local index j
#pragma unroll
      for (int j = 0; j < 8; j += 2) {    // i=64: k = n & 63 = t, m = 4
        bfly(a[j], a[j + 1], tw512[2 * t * 4], tw512[2 * t * 4 + 1]);
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {       // i=128: k = t + 64*(j&1), m = 2
        if (j & 2) continue;
        const int k = t + 64 * (j & 1);
        bfly(a[j], a[j + 2], tw512[2 * k * 2], tw512[2 * k * 2 + 1]);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {       // i=256: k = t + 64*j, m = 1
        const int k = t + 64 * j;
        bfly(a[j], a[j + 4], tw512[2 * k], tw512[2 * k + 1]);
      }
      // post-IFFT complex multiplication (MDCT.java:48-53)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = t + 64 * j;
        const float cs = __ldg(T.mdct_long + 2 * n), sn = __ldg(T.mdct_long + 2 * n + 1);
        const float t0 = a[j].re, t1 = a[j].im;
        bim[n] = (t1 * cs) + (t0 * sn);
        bre[n] = (t0 * cs) - (t1 * sn);
      }
    } else {
      const int blk = t >> 3, col = t & 7;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int nl = col + 8 * j;         // index inside the 64-point FFT of window blk
        const float cs = __ldg(T.mdct_short + 2 * nl), sn = __ldg(T.mdct_short + 2 * nl + 1);
        const float t0 = a[j].re, t1 = a[j].im;
        bim[64 * blk + nl] = (t1 * cs) + (t0 * sn);
        bre[64 * blk + nl] = (t0 * cs) - (t1 * sn);
      }
    }
    __syncthreads();

    // ---- phase 3: windowing + overlap-add (FilterBank.java:39-123) + PCM
    {
      const float* LWp = T.win_long[shape_prev];
      const float* LW = T.win_long[shape_cur];
      const float* SWp = T.win_short[shape_prev];
      const float* SW = T.win_short[shape_cur];
      float outv[16], ovlv[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int i = t + 64 * j;
        const float ov = my_ovl[i];
        float o, nv;
        if (ws == 0) {
          o = ov + (mdct_out(bre, bim, 512, 256, i) * __ldg(LWp + i));
          nv = mdct_out(bre, bim, 512, 256, 1024 + i) * __ldg(LW + 1023 - i);
        } else if (ws == 1) {
          o = ov + (mdct_out(bre, bim, 512, 256, i) * __ldg(LWp + i));
          if (i < 448) nv = mdct_out(bre, bim, 512, 256, 1024 + i);
          else if (i < 576) nv = mdct_out(bre, bim, 512, 256, 1024 + i) * __ldg(SW + 127 - (i - 448));
          else nv = 0.f;
        } else if (ws == 3) {
          if (i < 448) o = ov;
          else if (i < 576) o = ov + (mdct_out(bre, bim, 512, 256, i) * __ldg(SWp + (i - 448)));
          else o = ov + mdct_out(bre, bim, 512, 256, i);
          nv = mdct_out(bre, bim, 512, 256, 1024 + i) * __ldg(LW + 1023 - i);
        } else {
          // EIGHT_SHORT: window w occupies b[256w .. 256w+255]; its samples come from FFT block w
          // out[448 + 128*s + r]
          if (i < 448) o = ov;
          else {
            const int s = (i - 448) >> 7, r = (i - 448) & 127;
            if (s == 0) {
              o = ov + (mdct_out(bre, bim, 64, 32, r) * __ldg(SWp + r));
            } else {
              // second half of window s-1 + first half of window s (s = 1..4; s==4 only for r < 64)
              const float a2 = mdct_out(bre + 64 * (s - 1), bim + 64 * (s - 1), 64, 32, 128 + r) * __ldg(SW + 127 - r);
              const float b2 = mdct_out(bre + 64 * s, bim + 64 * s, 64, 32, r) * __ldg(SW + r);
              o = (ov + a2) + b2;
            }
          }
          // new overlap
          if (i >= 576) nv = 0.f;
          else {
            // overlap[i]: i in [0,64): window 3 second half (r = 64+i) + window 4 first half
            //             i = 64 + 128*u + r: window 4+u second half + window 5+u first half (u = 0..2)
            //             i in [448,576): window 7 second half only
            if (i < 64) {
              const int r = 64 + i;
              nv = (mdct_out(bre + 64 * 3, bim + 64 * 3, 64, 32, 128 + r) * __ldg(SW + 127 - r)) +
                   (mdct_out(bre + 64 * 4, bim + 64 * 4, 64, 32, r) * __ldg(SW + r));
            } else if (i < 448) {
              const int u = (i - 64) >> 7, r = (i - 64) & 127;
              nv = (mdct_out(bre + 64 * (4 + u), bim + 64 * (4 + u), 64, 32, 128 + r) * __ldg(SW + 127 - r)) +
                   (mdct_out(bre + 64 * (5 + u), bim + 64 * (5 + u), 64, 32, r) * __ldg(SW + r));
            } else {
              const int r = i - 448;
              nv = mdct_out(bre + 64 * 7, bim + 64 * 7, 64, 32, 128 + r) * __ldg(SW + 127 - r);
            }
          }
        }
        outv[j] = o;
        ovlv[j] = nv;
      }
      uint8_t* dst = pcm + pcm_off[f];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const int i = t + 64 * j;
        my_ovl[i] = ovlv[j];
        if (PCM_FORMAT == 2) {
          float* d = reinterpret_cast<float*>(dst);
          d[(size_t)c * 1024 + i] = outv[j];
          if (run.mono_dup) d[1024 + i] = outv[j];
        } else {
          const int v = pcm_round(outv[j]);
          uint16_t u = (uint16_t)(int16_t)v;
          if (PCM_FORMAT == 1) u = (uint16_t)((u >> 8) | (u << 8));
          if (run.mono_dup) { s_pcm[2 * i] = (int16_t)u; s_pcm[2 * i + 1] = (int16_t)u; }
          else s_pcm[i * out_ch + c] = (int16_t)u;
        }
      }
      __syncthreads();
      if (PCM_FORMAT != 2) {
        // coalesced copy-out of the interleaved frame
        const int nwords = 1024 * out_ch / 2;   // 32-bit words
        const uint32_t* src = reinterpret_cast<const uint32_t*>(s_pcm);
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        for (int i = tid; i < nwords; i += nthreads) d[i] = src[i];
      }
      if (tid == 0) pcm_bytes_out[f] = (uint32_t)(1024 * out_ch * (PCM_FORMAT == 2 ? 4 : 2));
    }
    __syncthreads();
  }

  // persistent state out
  for (int i = t; i < 1024; i += kThreadsPerChannel) g_ovl[i] = my_ovl[i];
  if (t == 0) sstate[run.stream_slot].window_shape[c] = (uint8_t)shape_cur;
}

}  // namespace jaadb
