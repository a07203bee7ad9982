// K2 -- spectral processing.  One CTA owns one stream and walks that stream's
// frames of the batch in decode order, 64 threads per channel.  The next frame's
// descriptor, side information and quantised coefficients are fetched into
// registers while the current frame is in its FFT, so HBM latency stays off the
// per-frame critical path:
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

// Java Math.round(float) + SampleBuffer clamp (S/SampleBuffer.java:193-205).
// floor + exact fractional test is Math.round for every float: |x| >= 2^23 has no fraction, +-inf give a NaN
// difference (test false) and saturate in the conversion, NaN converts to 0 -- the same values Java produces.
__device__ __forceinline__ int pcm_round(float x) {
  const float f = floorf(x);
  const int r = __float2int_rz(f) + ((x - f) >= 0.5f ? 1 : 0);
  return min(max(r, -32768), 32767);
}

// Per-thread constants of one 4-coefficient sub-chunk (all SWB offsets are multiples of 4, so the four
// coefficients i0..i0+3 always share a scalefactor band):
//   [5:0] sfb if the window is long (63: beyond the table)   [9:6] sfb if short (15: beyond the table)
//   [17:10] swb_short[sfb]   [23:18] width of that short band   [26:24] short window index (i0 >> 7)
__device__ __forceinline__ uint32_t subchunk_consts(const TablesDev& T, int sf_index, int i0) {
  int sl = T.sfb_of_long[sf_index * 1024 + i0];
  if (sl > 62) sl = 63;
  int ss = T.sfb_of_short[sf_index * 128 + (i0 & 127)];
  int lo = 0, width = 0;
  if (ss > 14) ss = 15;
  else { lo = T.swb_short[sf_index * 17 + ss]; width = T.swb_short[sf_index * 17 + ss + 1] - lo; }
  return (uint32_t)sl | ((uint32_t)ss << 6) | ((uint32_t)lo << 10) | ((uint32_t)width << 18) | ((uint32_t)(i0 >> 7) << 24);
}

// Dequantises coefficients i0..i0+3 of one channel (ICStream.java:264-269): v = +-IQ_TABLE[|q|] * scaleFactors[idx].
// `qpre` holds q[i0..i0+3] as fetched at the natural position, which is where K1 put them for long windows; for
// EIGHT_SHORT the bitstream position is computed from the grouping and the four values are re-read.
// cb_out / idx_out: the band's codebook and its (group, sfb) index; cb_out = 0 for bands at or above max_sfb.
__device__ __forceinline__ void dequant4(const IcsSide* __restrict__ s, const int16_t* __restrict__ q, uint2 qpre,
                                         uint32_t pk, int i0, const TablesDev& T, float v[4], int& cb_out, int& idx_out) {
  v[0] = v[1] = v[2] = v[3] = 0.f;
  cb_out = 0;
  idx_out = 0;
  const int max_sfb = s->max_sfb;
  int sfb, idx;
  if (s->window_sequence == 2) {
    sfb = (int)((pk >> 6) & 15u);
    if (sfb >= max_sfb) return;
    // window -> group: first window and length of the group that holds this window
    const int w = (int)((pk >> 24) & 7u);
    int g = 0, gstart = 0, glen = s->group_len[0];
    for (int k = 0, acc = 0; k < 7; ++k) {
      acc += s->group_len[k];
      if (w >= acc && k + 1 < s->num_groups) { g = k + 1; gstart = acc; glen = s->group_len[k + 1]; }
    }
    idx = g * max_sfb + sfb;
    const int cb = s->sfb_cb[idx];
    cb_out = cb;
    idx_out = idx;
    if (cb == 0 || cb > 11) return;
    const int lo = (int)((pk >> 10) & 255u), width = (int)((pk >> 18) & 63u);
    const int qpos = 128 * gstart + glen * lo + (w - gstart) * width + ((i0 & 127) - lo);
    qpre = *reinterpret_cast<const uint2*>(q + qpos);
  } else {
    sfb = (int)(pk & 63u);
    if (sfb >= max_sfb) return;
    idx = sfb;
    const int cb = s->sfb_cb[idx];
    cb_out = cb;
    idx_out = idx;
    if (cb == 0 || cb > 11) return;
  }
  const float sf = __ldg(T.sf + s->sf_idx[idx]);
  const int q0 = (int)(int16_t)(qpre.x & 0xFFFFu), q1 = (int)qpre.x >> 16;
  const int q2 = (int)(int16_t)(qpre.y & 0xFFFFu), q3 = (int)qpre.y >> 16;
  const float m0 = __ldg(T.iq + abs(q0)), m1 = __ldg(T.iq + abs(q1));
  const float m2 = __ldg(T.iq + abs(q2)), m3 = __ldg(T.iq + abs(q3));
  // iqData = (v>0) ? IQ[v] : -IQ[-v]; iqData *= scaleFactors[idx]   (ICStream.java:266-267)
  v[0] = (q0 > 0 ? m0 : -m0) * sf;
  v[1] = (q1 > 0 ? m1 : -m1) * sf;
  v[2] = (q2 > 0 ? m2 : -m2) * sf;
  v[3] = (q3 > 0 ? m3 : -m3) * sf;
}

__device__ __forceinline__ void channel_barrier(int c) {
  // the 64 threads (two warps) of one channel
  asm volatile("bar.sync %0, 64;" ::"r"(c + 1) : "memory");
}

template <int PCM_FORMAT, int MAX_THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(MAX_THREADS, MIN_BLOCKS)
k2_filterbank_kernel(const RunDev* __restrict__ runs, const RunFrameDev* __restrict__ run_frames,
                     FrameSide* __restrict__ fside, const IcsSide* __restrict__ iside,
                     const int16_t* __restrict__ qall, float* __restrict__ overlap_all, StreamState* __restrict__ sstate,
                     uint8_t* __restrict__ pcm, const uint64_t* __restrict__ pcm_off,
                     uint32_t* __restrict__ pcm_bytes_out, float* __restrict__ spec_tap, float* __restrict__ core, TablesDev T,
                     const LayoutDev* __restrict__ layouts, int nch) {
  extern __shared__ __align__(16) float smem[];
  // carve: twiddles; per channel [spec kSpecStride][overlap 1024][xre kXchgStride][xim kXchgStride]; sides; pcm staging
  const int per_ch = kSpecStride + 1024 + 2 * kXchgStride;
  float* s_fft_tw = smem;                         // fft512 re/im (inverse) [256][2] + fft64 [32][2]
  float* s_ch = s_fft_tw + 2 * 256 + 2 * 32;
  IcsSide* s_side = reinterpret_cast<IcsSide*>(s_ch + nch * per_ch);
  int16_t* s_pcm = reinterpret_cast<int16_t*>(s_side + nch);  // [1024][out_ch] (s16 formats)

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
  // windows and MDCT twiddles stay in global memory (L1-resident, 13.8 KB): staging them in shared memory was measured
  // to give nothing and costs a resident CTA per SM
  const float2* mdct_long2 = reinterpret_cast<const float2*>(T.mdct_long);
  const float2* mdct_short2 = reinterpret_cast<const float2*>(T.mdct_short);

  // persistent state in: overlap + current window shapes
  float* g_ovl = overlap_all + ((size_t)run.stream_slot * kMaxChannels + c) * 1024;
  for (int i = t; i < 256; i += kThreadsPerChannel)
    reinterpret_cast<float4*>(my_ovl)[i] = reinterpret_cast<const float4*>(g_ovl)[i];
  int shape_cur = sstate[run.stream_slot].window_shape[c];

  // element of this thread's channel
  int el_first = c, el_nch = 1, my_el = 0;
  for (int e = 0; e < lay.n_elements; ++e) {
    int f0 = lay.el_first_ch[e];
    int n = lay.el_type[e] == EL_CPE ? 2 : 1;
    if (c >= f0 && c < f0 + n) { el_first = f0; el_nch = n; my_el = e; }
  }
  // dequantisation work split: a CPE thread owns coefficients 8*et .. 8*et+7 of L and of R (M/S and IS are
  // element-wise across the pair); an SCE/LFE thread owns 16*et .. 16*et+15 of its channel.
  const int et = tid - el_first * kThreadsPerChannel;
  const int chA = el_first, chB = el_first + (el_nch == 2 ? 1 : 0);
  const int iA = (el_nch == 2) ? 8 * et : 16 * et;
  const int iB = (el_nch == 2) ? iA : iA + 8;
  const uint32_t pkA0 = subchunk_consts(T, sf_index, iA), pkA1 = subchunk_consts(T, sf_index, iA + 4);
  const uint32_t pkB0 = subchunk_consts(T, sf_index, iB), pkB1 = subchunk_consts(T, sf_index, iB + 4);

  // ---- software pipeline: descriptor, status, side information and q of the next frame live in registers
  RunFrameDev cur = run_frames[run.first];
  // status and the element instance tags of the frame: {status, tags | n_elements << 16 | n_started << 24}
  uint2 fsw = *reinterpret_cast<const uint2*>(fside + cur.frame);
  uint32_t exp_tags = sstate[run.stream_slot].tags, exp_mask = 0;
  {
    const uint32_t v = sstate[run.stream_slot].tags_valid;
    for (int i = 0; i < 4; ++i) exp_mask |= ((v >> i) & 1u) ? (0xFu << (4 * i)) : 0u;
  }
  uint4 side_pf = make_uint4(0, 0, 0, 0);
  const int side_vecs = nch * (int)(sizeof(IcsSide) / 16);
  if (tid < side_vecs) side_pf = reinterpret_cast<const uint4*>(iside + cur.ics_base)[tid];
  uint4 qA = *reinterpret_cast<const uint4*>(qall + ((size_t)cur.ics_base + chA) * 1024 + iA);
  uint4 qB = *reinterpret_cast<const uint4*>(qall + ((size_t)cur.ics_base + chB) * 1024 + iB);
  uint64_t poff_pf = pcm_off[cur.frame];
  __syncthreads();

  for (uint32_t it = 0; it < run.count; ++it) {
    const uint32_t f = cur.frame;
    const uint32_t ics_base = cur.ics_base;
    const uint64_t poff = poff_pf;
    const bool have_next = it + 1 < run.count;
    RunFrameDev nxt = cur;
    if (have_next) nxt = run_frames[run.first + it + 1];
    // side info -> shared
    if (tid < side_vecs) reinterpret_cast<uint4*>(s_side)[tid] = side_pf;
    __syncthreads();
    const IcsSide* sd = s_side + c;
    // Element objects are per (type, instance tag) in JAAD (StreamState::tags).  An element that shows another tag, or
    // that the layout does not have, addresses objects this stream does not own: it leaves them alone, the frame is
    // reported as JAADB_ST_LAYOUT -- and the stream's own elements of that frame still go through the filterbank when
    // the frame was parsed to its end, as they do in JAAD (SyntacticElements.process runs after the whole parse).
    int frame_status = (int)fsw.x;
    bool el_live;      // this thread's element belongs to the stream and was parsed completely
    bool shape_ok;     // ... belongs to the stream (window-shape bookkeeping happens in failing frames too)
    {
      // nibble i of `started` / `exp_mask` = element i has shown its tag in this frame / earlier
      const uint32_t tags = fsw.y & 0xFFFFu, started = (1u << (4 * min(fsw.y >> 24, 4u))) - 1u;
      const uint32_t diff = (tags ^ exp_tags) & exp_mask & started;
      const uint32_t fresh = started & ~exp_mask;
      exp_tags |= tags & fresh;
      exp_mask |= fresh;
      const int n_good = (int)((fsw.y >> 16) & 0xFFu);
      shape_ok = my_el >= 4 || ((diff >> (4 * my_el)) & 15u) == 0;
      el_live = shape_ok && my_el < n_good;
      if (diff != 0 && frame_status == 0) {
        frame_status = JAADB_ST_LAYOUT;
        if (tid == 0) fside[f].status = JAADB_ST_LAYOUT;
      }
    }
    const bool emit = frame_status == 0;                                        // the frame yields PCM
    const bool parsed = emit || frame_status == JAADB_ST_LAYOUT;                // JAAD reached SyntacticElements.process
    // (SBR streams: the SBR stages only run for frames that yield PCM, so the core coder's state waits for them too)
    const bool run_ch = parsed && el_live && (emit || !run.sbr);
    // window-shape bookkeeping of ICSInfo.decode / setCommonData (ICSInfo.java:90-91,196-197)
    int shape_prev = shape_cur;
    if (sd->info_decoded && shape_ok) { shape_prev = shape_cur; shape_cur = sd->window_shape; }
    const int ws = sd->window_sequence;

    if (run_ch) {
      // ---- phase 1: dequantise + M/S + IS into the element's spectra
      const int16_t* qL = qall + ((size_t)ics_base + chA) * 1024;
      const int16_t* qR = qall + ((size_t)ics_base + chB) * 1024;
      const IcsSide* sL = s_side + chA;
      const IcsSide* sR = s_side + chB;
      float* specA = s_ch + chA * per_ch;
      float* specB = s_ch + chB * per_ch;
      float a[8], b[8];
      int cbA0, cbA1, cbB0, cbB1, idxA0, idxA1, idxB0, idxB1;
      dequant4(sL, qL, make_uint2(qA.x, qA.y), pkA0, iA, T, a, cbA0, idxA0);
      dequant4(sL, qL, make_uint2(qA.z, qA.w), pkA1, iA + 4, T, a + 4, cbA1, idxA1);
      dequant4(sR, qR, make_uint2(qB.x, qB.y), pkB0, iB, T, b, cbB0, idxB0);
      dequant4(sR, qR, make_uint2(qB.z, qB.w), pkB1, iB + 4, T, b + 4, cbB1, idxB1);
      if (el_nch == 2) {
        const bool ms_on = sL->common_window && sL->ms_mask != 0;   // CPE.java:159-160
        const bool ms_present = sL->ms_mask != 0;                   // CPE.isMSMaskPresent
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int cbL = h ? cbA1 : cbA0, cbR = h ? cbB1 : cbB0;
          const int idxL = h ? idxA1 : idxA0, idxR = h ? idxB1 : idxB0;
          // MS.process: both codebooks < NOISE_HCB, band flagged (MS.java:28-36).  Bands at or above max_sfb have
          // cb_out 0 and idx 0: they hold zeros, for which the butterfly is the identity up to the sign of zero --
          // JAAD never touches them, so they are excluded through the left channel's band test.
          const bool in_band = ((sL->window_sequence == 2) ? (int)(((h ? pkA1 : pkA0) >> 6) & 15u) : (int)((h ? pkA1 : pkA0) & 63u)) < sL->max_sfb;
          if (ms_on && cbL < 13 && cbR < 13 && in_band && ((sL->ms_used[idxL >> 3] >> (idxL & 7)) & 1)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float l = a[4 * h + j], r = b[4 * h + j];
              a[4 * h + j] = l + r;
              b[4 * h + j] = l - r;
            }
          }
          // IS.process: right channel bands with codebook 14/15 (IS.java:29-44)
          if (cbR == 15 || cbR == 14) {
            int sgn = cbR == 15 ? 1 : -1;
            if (ms_present) sgn *= ((sL->ms_used[idxR >> 3] >> (idxR & 7)) & 1) ? -1 : 1;
            float scale = __ldg(T.sf + sR->sf_idx[idxR]);
            if (sgn < 0) scale = -scale;
#pragma unroll
            for (int j = 0; j < 4; ++j) b[4 * h + j] = a[4 * h + j] * scale;
          }
        }
      }
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        specA[spec_addr(iA + j)] = a[j];
        specB[spec_addr(iB + j)] = b[j];
      }
      if (spec_tap) {
        float* tA = spec_tap + ((size_t)ics_base + chA) * 1024 + iA;
        float* tB = spec_tap + ((size_t)ics_base + chB) * 1024 + iB;
#pragma unroll
        for (int j = 0; j < 8; ++j) { tA[j] = a[j]; tB[j] = b[j]; }
      }
    }
    // ---- prefetch the next frame while this one is transformed
    if (have_next) {
      fsw = *reinterpret_cast<const uint2*>(fside + nxt.frame);
      if (tid < side_vecs) side_pf = reinterpret_cast<const uint4*>(iside + nxt.ics_base)[tid];
      qA = *reinterpret_cast<const uint4*>(qall + ((size_t)nxt.ics_base + chA) * 1024 + iA);
      qB = *reinterpret_cast<const uint4*>(qall + ((size_t)nxt.ics_base + chB) * 1024 + iB);
      poff_pf = pcm_off[nxt.frame];
    }
    cur = nxt;
    __syncthreads();
    if (!parsed) {
      if (tid == 0 && !run.sbr) pcm_bytes_out[f] = 0;
      continue;  // the frame produced no PCM; overlap untouched (Decoder.java:96-98)
    }

    // ---- phase 2: IMDCT of this channel (thread t owns points 8t..8t+7 of the bit-reversed input)
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
          const float2 cs = __ldg(mdct_long2 + k);
          a[j].im = (x0 * cs.x) + (x1 * cs.y);
          a[j].re = (x1 * cs.x) - (x0 * cs.y);
        }
      } else {
        const int w = t >> 3;                                  // short window handled by this thread
        const int kb = brev3(t & 7);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int k = kb + 8 * brev3(j);                     // bitrev6(8*(t&7)+j)
          const float x0 = my_spec[spec_addr(128 * w + 2 * k)];
          const float x1 = my_spec[spec_addr(128 * w + 127 - 2 * k)];
          const float2 cs = __ldg(mdct_short2 + k);
          a[j].im = (x0 * cs.x) + (x1 * cs.y);
          a[j].re = (x1 * cs.x) - (x0 * cs.y);
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
    channel_barrier(c);
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
    // The spectrum was consumed before the channel barrier above, so its storage now carries exchange 2
    // (planes re | im, 528 floats each); the post-twiddled buffer then goes where exchange 1 was.
    float* x2re = my_spec;
    float* x2im = my_spec + 528;
    float* bre = my_xre;               // post-twiddled buffer as planes re[512] | im[512]
    float* bim = my_xim;
    if (!is_short) {
      // exchange 2: write n = 64*blk + col + 8*j, read n = t + 64*j
      const int blk = t >> 3, col = t & 7;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = 64 * blk + col + 8 * j;
        x2re[n + (n >> 5)] = a[j].re;
        x2im[n + (n >> 5)] = a[j].im;
      }
    }
    channel_barrier(c);   // also: every thread of the channel is done reading exchange 1
    if (!is_short) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int n = t + 64 * j;
        a[j].re = x2re[n + (n >> 5)];
        a[j].im = x2im[n + (n >> 5)];
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
        const float2 cs = __ldg(mdct_long2 + n);
        const float t0 = a[j].re, t1 = a[j].im;
        bim[n] = (t1 * cs.x) + (t0 * cs.y);
        bre[n] = (t0 * cs.x) - (t1 * cs.y);
      }
    } else {
      const int blk = t >> 3, col = t & 7;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int nl = col + 8 * j;         // index inside the 64-point FFT of window blk
        const float2 cs = __ldg(mdct_short2 + nl);
        const float t0 = a[j].re, t1 = a[j].im;
        bim[64 * blk + nl] = (t1 * cs.x) + (t0 * cs.y);
        bre[64 * blk + nl] = (t0 * cs.x) - (t1 * cs.y);
      }
    }
    channel_barrier(c);

    // ---- phase 3: windowing + overlap-add (FilterBank.java:39-123) + PCM.  Thread t owns the sample pairs
    // i = 2t + 128j, i+1 (j = 0..7): the MDCT reorder (MDCT.java:60-80) then reads mirrored positions of the two
    // planes without any per-lane case split, and the overlap goes back as float2.
    {
      const float* __restrict__ LWp = T.win_long[shape_prev];
      const float* __restrict__ LW = T.win_long[shape_cur];
      const float* __restrict__ SWp = T.win_short[shape_prev];
      const float* __restrict__ SW = T.win_short[shape_cur];
      uint8_t* dst = pcm + poff;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const int i = 2 * t + 128 * j;
        const float2 ov = *reinterpret_cast<const float2*>(my_ovl + i);
        float o0, o1, n0, n1;
        if (!is_short) {
          // x1* = first half of the IMDCT output at i, i+1; x2* = second half (index 1024+i, 1024+i+1)
          float x10, x11, x20, x21;
          if (j < 4) {
            const int H = t + 64 * j;
            x10 = bim[256 + H]; x11 = -bre[255 - H];
            x20 = bre[256 + H]; x21 = -bim[255 - H];
          } else {
            const int H = t + 64 * (j - 4);
            x10 = bre[H]; x11 = -bim[511 - H];
            x20 = -bim[H]; x21 = bre[511 - H];
          }
          if (ws == 3) {
            // LONG_STOP
            if (i < 448) { o0 = ov.x; o1 = ov.y; }
            else if (i < 576) {
              const float2 w = __ldg(reinterpret_cast<const float2*>(SWp + (i - 448)));
              o0 = ov.x + (x10 * w.x); o1 = ov.y + (x11 * w.y);
            } else { o0 = ov.x + x10; o1 = ov.y + x11; }
          } else {
            const float2 w = __ldg(reinterpret_cast<const float2*>(LWp + i));
            o0 = ov.x + (x10 * w.x); o1 = ov.y + (x11 * w.y);
          }
          if (ws == 1) {
            // LONG_START
            if (i < 448) { n0 = x20; n1 = x21; }
            else if (i < 576) {
              const float2 w = __ldg(reinterpret_cast<const float2*>(SW + 126 - (i - 448)));
              n0 = x20 * w.y; n1 = x21 * w.x;
            } else { n0 = 0.f; n1 = 0.f; }
          } else {
            const float2 w = __ldg(reinterpret_cast<const float2*>(LW + 1022 - i));
            n0 = x20 * w.y; n1 = x21 * w.x;
          }
        } else {
          // EIGHT_SHORT: window w occupies b[256w .. 256w+255]; its samples come from FFT block w
          float oo[2], nn[2];
#pragma unroll
          for (int p = 0; p < 2; ++p) {
            const int ii = i + p;
            const float ovp = p ? ov.y : ov.x;
            float o, nv;
            if (ii < 448) o = ovp;
            else {
              const int s = (ii - 448) >> 7, r = (ii - 448) & 127;
              if (s == 0) {
                o = ovp + (mdct_out(bre, bim, 64, 32, r) * __ldg(SWp + r));
              } else {
                // second half of window s-1 + first half of window s (s = 1..4; s==4 only for r < 64)
                const float a2 = mdct_out(bre + 64 * (s - 1), bim + 64 * (s - 1), 64, 32, 128 + r) * __ldg(SW + 127 - r);
                const float b2 = mdct_out(bre + 64 * s, bim + 64 * s, 64, 32, r) * __ldg(SW + r);
                o = (ovp + a2) + b2;
              }
            }
            if (ii >= 576) nv = 0.f;
            else if (ii < 64) {
              // overlap[i], i in [0,64): window 3 second half (r = 64+i) + window 4 first half
              const int r = 64 + ii;
              nv = (mdct_out(bre + 64 * 3, bim + 64 * 3, 64, 32, 128 + r) * __ldg(SW + 127 - r)) +
                   (mdct_out(bre + 64 * 4, bim + 64 * 4, 64, 32, r) * __ldg(SW + r));
            } else if (ii < 448) {
              // i = 64 + 128*u + r: window 4+u second half + window 5+u first half (u = 0..2)
              const int u = (ii - 64) >> 7, r = (ii - 64) & 127;
              nv = (mdct_out(bre + 64 * (4 + u), bim + 64 * (4 + u), 64, 32, 128 + r) * __ldg(SW + 127 - r)) +
                   (mdct_out(bre + 64 * (5 + u), bim + 64 * (5 + u), 64, 32, r) * __ldg(SW + r));
            } else {
              // i in [448,576): window 7 second half only
              const int r = ii - 448;
              nv = mdct_out(bre + 64 * 7, bim + 64 * 7, 64, 32, 128 + r) * __ldg(SW + 127 - r);
            }
            oo[p] = o;
            nn[p] = nv;
          }
          o0 = oo[0]; o1 = oo[1]; n0 = nn[0]; n1 = nn[1];
        }
        if (run_ch) *reinterpret_cast<float2*>(my_ovl + i) = make_float2(n0, n1);
        if (!emit) {
          // a frame JAAD decodes against element objects this stream does not own: state only, no PCM
        } else if (run.sbr) {
          // core-coder output of an SBR stream: K4 continues from here
          *reinterpret_cast<float2*>(core + ((size_t)ics_base + c) * 1024 + i) = make_float2(o0, o1);
        } else if (PCM_FORMAT == 2) {
          float* d = reinterpret_cast<float*>(dst);
          *reinterpret_cast<float2*>(d + (size_t)c * 1024 + i) = make_float2(o0, o1);
          if (run.mono_dup) *reinterpret_cast<float2*>(d + 1024 + i) = make_float2(o0, o1);
        } else {
          uint32_t u0 = (uint32_t)pcm_round(o0) & 0xFFFFu, u1 = (uint32_t)pcm_round(o1) & 0xFFFFu;
          if (PCM_FORMAT == 1) { u0 = __byte_perm(u0, 0, 0x4401); u1 = __byte_perm(u1, 0, 0x4401); }
          if (run.mono_dup) {
            *reinterpret_cast<uint2*>(s_pcm + 2 * i) = make_uint2(u0 | (u0 << 16), u1 | (u1 << 16));
          } else {
            s_pcm[i * out_ch + c] = (int16_t)u0;
            s_pcm[(i + 1) * out_ch + c] = (int16_t)u1;
          }
        }
      }
      __syncthreads();
      if (PCM_FORMAT != 2 && !run.sbr && emit) {
        // coalesced copy-out of the interleaved frame (pcm offsets are 4-byte aligned; 16 B when the caller packs)
        const int nwords = 1024 * out_ch / 2;   // 32-bit words
        const uint32_t* src = reinterpret_cast<const uint32_t*>(s_pcm);
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
          for (int i = tid; i < nwords / 4; i += nthreads)
            reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(src)[i];
        } else {
          for (int i = tid; i < nwords; i += nthreads) d[i] = src[i];
        }
      }
      if (tid == 0 && !run.sbr) pcm_bytes_out[f] = emit ? (uint32_t)(1024 * out_ch * (PCM_FORMAT == 2 ? 4 : 2)) : 0u;
    }
  }

  // persistent state out (the last frame's phase 3 wrote my_ovl; the loop's trailing barrier ordered it)
  __syncthreads();
  for (int i = t; i < 256; i += kThreadsPerChannel)
    reinterpret_cast<float4*>(g_ovl)[i] = reinterpret_cast<const float4*>(my_ovl)[i];
  if (t == 0) sstate[run.stream_slot].window_shape[c] = (uint8_t)shape_cur;
  if (tid == 0) {
    uint32_t v = 0;
    for (int i = 0; i < 4; ++i) v |= ((exp_mask >> (4 * i)) & 1u) << i;
    sstate[run.stream_slot].tags = (uint16_t)exp_tags;
    sstate[run.stream_slot].tags_valid = (uint8_t)v;
  }
}

}  // namespace jaadb
