// K2 -- spectral processing.
//
// k2_prepass_kernel: one thread per stream walks that stream's frames of the batch in decode order and resolves
//   everything that depends on earlier frames -- window_shape[PREVIOUS] (ICSInfo.java:90-91,196-197), the element
//   instance tags the stream owns (Element.java:36-38), which channels reach the filterbank, the state of the PNS
//   generator at the start of each frame (ICStream.java:26,247) -- into one 16-byte K2FrameDev per frame.
// k2_filterbank_kernel: one CTA per segment of a stream (the whole run when there are plenty of streams, a few
//   frames when there are few), 64 threads per channel:
//   dequantisation (|q|^(4/3) LUT x 2^((sf-100)/4) LUT, ICStream.java:264-269) + PNS (ICStream.java:241-257)
//   -> M/S (tools/MS.java:17-41) -> intensity stereo (tools/IS.java:17-53)
//   -> [JAADB_TNS_ISO: the all-pole filter of ISO/IEC 14496-3 4.6.9.3; JAAD's TNS.process is a stub]
//   -> IMDCT as an N/4-point complex FFT in registers + shared memory
//      (filterbank/MDCT.java:36-81, FFT.java:48-135)
//   -> sine/KBD windowing + overlap-add (filterbank/FilterBank.java:39-123)
//   -> Math.round / clamp / interleave to int16 (S/SampleBuffer.java:168-209).
// The next frame's quantised coefficients and side information come in by a TMA bulk copy (cp.async.bulk + mbarrier)
// into a shared-memory stage while the current frame is transformed; the overlap buffers of the stream stay in shared
// memory for the whole segment.  The overlap a frame leaves behind is a function of that frame alone (every window
// sequence assigns all 1024 entries), so a segment that does not start its run first re-runs the frame(s) before it
// without output to get the overlap it starts from: the same float operations in the same order, still bit-exact.
//
// Bit-exactness: every floating-point operation is the same binary32 operation, on the same operands, as in the Java
// code (separate multiply and add, no FMA: this file must be compiled with --fmad=false).  The FFT keeps JAAD's
// butterfly graph -- bit reversal, one radix-4 stage without twiddles, then radix-2 stages -- and only changes which
// thread evaluates which butterfly, so the float PCM is bit-identical to JAAD's, not merely within tolerance.
#pragma once
#include "jaadb_types.cuh"

namespace jaadb {

constexpr int kThreadsPerChannel = 64;
// shared memory per channel (floats): spectrum (1024 + 1 pad per 32) / FFT exchange 2 (two planes of 576) / packed PCM,
// overlap (1024), FFT exchange 1 + post-twiddled buffer (two planes of 576)
constexpr int kSpecStride = 1136;   // spectrum: 1024 + 1024 / 32 padding words; as exchange 2: planes re [0, 568) | im [568, 1136)
constexpr int kXchgStride = 8 * 72;     // 8 blocks of 8x8 complex, rows padded to 9; 8 short windows of 64 + 8
constexpr int kK2ChFloats = kSpecStride + 1024 + 2 * kXchgStride;
constexpr int kK2StageBytesPerCh = 2048 + (int)sizeof(IcsSide);   // q[1024] int16 + IcsSide of the next frame (+ its K2FrameDev, once)

__host__ __device__ constexpr size_t k2_smem_bytes(int nch, int out_ch, bool planar_pcm) {
  return sizeof(float) * (2 * 256 + 2 * 32 + (size_t)nch * kK2ChFloats) + (size_t)nch * kK2StageBytesPerCh + sizeof(K2FrameDev) + 16 +
         (planar_pcm ? sizeof(uint32_t) * 512 * (size_t)nch : sizeof(int16_t) * 1024 * (size_t)out_ch);
}

// ISO TNS tables (JAADB_TNS_ISO): tools/TNSTables.java:10-25 in TNS_TABLES order {0_3, 0_4, 1_3, 1_4}, and
// SampleFrequency.getMaximalTNS_SFB (SampleFrequency.java:15-26) as [sf_index][long, short]
__constant__ float c_tns_coef[36];
__constant__ uint8_t c_tns_max_sfb[24];

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

// Java Math.round(float) + SampleBuffer clamp (S/SampleBuffer.java:193-205) in two instructions.
// Math.round(x) = floor(x + 0.5) evaluated exactly.  fma.rm(x, 1, 0.5) is the exact sum rounded toward -inf: the largest
// float <= x + 0.5, which is never below floor(x + 0.5) (an integer that is a float itself whenever a fraction exists), so
// its floor is the floor of the exact sum.  cvt.rmi.sat.s16 takes that floor and saturates to [-32768, 32767] -- Java
// saturates at the int range first and SampleBuffer clamps to 16 bits, the same result -- and converts NaN to 0 like Java.
__device__ __forceinline__ uint32_t pcm_round16(float x) {
  float y;
  asm("fma.rm.f32 %0, %1, 0f3F800000, 0f3F000000;" : "=f"(y) : "f"(x));
  int16_t r;
  asm("cvt.rmi.sat.s16.f32 %0, %1;" : "=h"(r) : "f"(y));
  return (uint32_t)(uint16_t)r;
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

// Where short window w sits in the grouping of an EIGHT_SHORT channel: [3:0] group, [7:4] first window of the group,
// [11:8] windows in the group.  Branch-free from the group-start mask (bit j: window j starts a group).
__device__ __forceinline__ uint32_t short_group_of(const IcsSide* __restrict__ s, int w) {
  uint32_t starts = 1u, acc = 0;
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    acc += s->group_len[k];
    if (k + 1 < s->num_groups) starts |= 1u << acc;
  }
  const uint32_t below = starts & ((2u << w) - 1u);
  const int gstart = 31 - __clz((int)below);
  const int g = __popc(below) - 1;
  const int next = w + 1 + (__ffs((int)((starts | 0x100u) >> (w + 1))) - 1);
  return (uint32_t)g | ((uint32_t)gstart << 4) | ((uint32_t)(next - gstart) << 8);
}

// Dequantises coefficients i0..i0+3 of one channel (ICStream.java:264-269): v = +-IQ_TABLE[|q|] * scaleFactors[idx].
// `q` is the channel's quantised spectrum as K1 wrote it (bitstream order, see k1_parse.cuh) in the shared-memory stage:
// long windows sit at their natural position, for EIGHT_SHORT the position follows from the grouping (`grp`,
// short_group_of of the chunk's window).  cb_out / idx_out: the band's codebook and its (group, sfb) index; cb_out = 0
// for bands at or above max_sfb.
__device__ __forceinline__ void dequant4(const IcsSide* __restrict__ s, const int16_t* __restrict__ q, uint32_t grp,
                                         uint32_t pk, int i0, const TablesDev& T, float v[4], int& cb_out, int& idx_out) {
  v[0] = v[1] = v[2] = v[3] = 0.f;
  cb_out = 0;
  idx_out = 0;
  const int max_sfb = s->max_sfb;
  int sfb, idx, qpos = i0;
  if (s->window_sequence == 2) {
    sfb = (int)((pk >> 6) & 15u);
    if (sfb >= max_sfb) return;
    const int w = (int)((pk >> 24) & 7u);
    const int g = (int)(grp & 15u), gstart = (int)((grp >> 4) & 15u), glen = (int)((grp >> 8) & 15u);
    idx = g * max_sfb + sfb;
    const int lo = (int)((pk >> 10) & 255u), width = (int)((pk >> 18) & 63u);
    qpos = 128 * gstart + glen * lo + (w - gstart) * width + ((i0 & 127) - lo);
  } else {
    sfb = (int)(pk & 63u);
    if (sfb >= max_sfb) return;
    idx = sfb;
  }
  const int cb = s->sfb_cb[idx];
  cb_out = cb;
  idx_out = idx;
  if (cb == 0 || cb > 11) return;
  const uint2 qv = *reinterpret_cast<const uint2*>(q + qpos);
  const float sf = __ldg(T.sf + s->sf_idx[idx]);
  const int q0 = (int)(int16_t)(qv.x & 0xFFFFu), q1 = (int)qv.x >> 16;
  const int q2 = (int)(int16_t)(qv.y & 0xFFFFu), q3 = (int)qv.y >> 16;
  const float m0 = __ldg(T.iq + abs(q0)), m1 = __ldg(T.iq + abs(q1));
  const float m2 = __ldg(T.iq + abs(q2)), m3 = __ldg(T.iq + abs(q3));
  // iqData = (v>0) ? IQ[v] : -IQ[-v]; iqData *= scaleFactors[idx]   (ICStream.java:266-267)
  v[0] = (q0 > 0 ? m0 : -m0) * sf;
  v[1] = (q1 > 0 ? m1 : -m1) * sf;
  v[2] = (q2 > 0 ? m2 : -m2) * sf;
  v[3] = (q3 > 0 ? m3 : -m3) * sf;
}

// PNS (ICStream.java:241-257) for coefficients i0..i0+3 of a band with codebook 13.  While it parses, JAAD draws
// width values per window of the band from the generator, sums their squares in float in index order and scales the
// window's values by (float)(scaleFactors[idx] / Math.sqrt(energy)).  The draws are regenerated here: `state` is the
// generator at the start of this channel's parse, the band's draw offset is the number of values the channel's
// earlier noise bands took.  Rare path, kept out of line.
__device__ __noinline__ float4 pns_fill4(const IcsSide* __restrict__ s, uint32_t grp, uint32_t pk, int i0, int idx, uint32_t state,
                                         const int16_t* __restrict__ swb, const float* __restrict__ sf_table) {
  const bool sh = s->window_sequence == 2;
  const int max_sfb = s->max_sfb;
  // draws of the bands before idx, in the order ICStream.decodeSpectralData walks them
  uint32_t before = 0;
  for (int j = 0, g = 0, sfb = 0; j < idx; ++j) {
    if (s->sfb_cb[j] == 13) before += (uint32_t)(s->group_len[g] * (swb[sfb + 1] - swb[sfb]));
    if (++sfb == max_sfb) { sfb = 0; ++g; }
  }
  int width, k0;
  if (sh) {
    const int w = (int)((pk >> 24) & 7u), gstart = (int)((grp >> 4) & 15u);
    width = (int)((pk >> 18) & 63u);
    before += (uint32_t)((w - gstart) * width);
    k0 = (i0 & 127) - (int)((pk >> 10) & 255u);
  } else {
    const int sfb = (int)(pk & 63u);
    width = swb[sfb + 1] - swb[sfb];
    k0 = i0 - swb[sfb];
  }
  uint32_t r = pns_jump(state, before);
  float energy = 0.f;
  float x0 = 0.f, x1 = 0.f, x2 = 0.f, x3 = 0.f;
  for (int k = 0; k < width; ++k) {
    r = 1664525u * r + 1013904223u;
    const float f = (float)(int32_t)r;
    energy += f * f;
    const int d = k - k0;
    x0 = d == 0 ? f : x0; x1 = d == 1 ? f : x1; x2 = d == 2 ? f : x2; x3 = d == 3 ? f : x3;
  }
  const float sf = -__ldg(sf_table + (s->sf_idx[idx] & 0x3FFF));   // scaleFactors[idx] of a noise band is -2^(e/4) (ICStream.java:206)
  const float scale = (float)((double)sf / sqrt((double)energy));
  return make_float4(x0 * scale, x1 * scale, x2 * scale, x3 * scale);
}

// ISO/IEC 14496-3 4.6.9.3 (tns_decode_frame, tns_decode_coef, tns_ar_filter) for filter `filt` of window `w` of one channel,
// in place on the channel's spectrum in shared memory.  The filter parameters are read again from the frame (K1 checked
// them: TNS.java:35-61).  Same operations in the same order as the test oracle's ISO restatement.  Rare path, out of line.
__device__ __noinline__ void tns_iso_filter(float* __restrict__ spec, const IcsSide* __restrict__ s, const uint32_t* __restrict__ words,
                                            const int16_t* __restrict__ swb, int swb_count, int max_tns, int w, int filt) {
  const bool sh = s->window_sequence == 2;
  if (w >= (sh ? 8 : 1)) return;
  uint32_t pos = s->tns_bit_off;
  auto get = [&](int n) -> uint32_t {
    const uint32_t wi = pos >> 5;
    const uint32_t a = __byte_perm(__ldg(words + wi), 0, 0x0123), b = __byte_perm(__ldg(words + wi + 1), 0, 0x0123);
    const uint32_t v = __funnelshift_l(b, a, pos & 31u) >> (32 - n);
    pos += n;
    return v;
  };
  const int b0 = sh ? 1 : 2, b1 = sh ? 4 : 6, b2 = sh ? 3 : 5;
  int top = 0, bottom = 0, order = 0, direction = 0, table = 0, coef_len = 0;
  uint32_t coef_pos = 0;
  bool found = false;
  for (int ww = 0; ww <= w && !found; ++ww) {
    const int nf = (int)get(b0);
    if (!nf) continue;
    const int coef_res = (int)get(1);
    bottom = swb_count;
    for (int f = 0; f < nf; ++f) {
      const int length = (int)get(b1);
      const int ord = (int)get(b2);
      top = bottom;
      bottom = max(top - length, 0);
      int dir = 0, compress = 0;
      if (ord) { dir = (int)get(1); compress = (int)get(1); }
      if (ww == w && f == filt) {
        found = true;
        order = ord; direction = dir; table = 2 * compress + coef_res; coef_len = coef_res + 3 - compress; coef_pos = pos;
        break;
      }
      pos += (uint32_t)(ord * (coef_res + 3 - compress));
    }
  }
  if (!found || order == 0) return;
  // tns_decode_coef: TNSTables holds -sin(..), 4.6.9.3 has tmp2 = sin(coef / iqfac)
  const int tab_off = table == 0 ? 0 : table == 1 ? 8 : table == 2 ? 24 : 28;
  float lpc[21], b[21];
  lpc[0] = 1.0f;
  pos = coef_pos;
  for (int m = 1; m <= order; ++m) {
    const float t = -c_tns_coef[tab_off + (int)get(coef_len)];
    for (int i = 1; i < m; ++i) b[i] = lpc[i] + (t * lpc[m - i]);
    for (int i = 1; i < m; ++i) lpc[i] = b[i];
    lpc[m] = t;
  }
  const int start = swb[min(min(bottom, max_tns), (int)s->max_sfb)];
  const int end = swb[min(min(top, max_tns), (int)s->max_sfb)];
  const int size = end - start;
  if (size <= 0) return;
  int p = w * 128 + start, inc = 1;
  if (direction) { inc = -1; p = w * 128 + end - 1; }
  float state[20];
#pragma unroll
  for (int j = 0; j < 20; ++j) state[j] = 0.f;
  for (int i = 0; i < size; ++i, p += inc) {
    float y = spec[spec_addr(p)];
    for (int j = 0; j < order; ++j) y = y - (state[j] * lpc[j + 1]);
    for (int j = order - 1; j > 0; --j) state[j] = state[j - 1];
    state[0] = y;
    spec[spec_addr(p)] = y;
  }
}

// The 64 threads (two warps) of one channel.  The barrier number is an immediate: with a register operand ptxas reserves all
// 16 named barriers for the CTA, and an SM only has 64 -- four resident CTAs, whatever registers and shared memory allow
// (launch__occupancy_limit_barriers in the round-1 profiles).
template <int MAX_CH>
__device__ __forceinline__ void channel_barrier(int c) {
  if (MAX_CH <= 2) {
    if (c == 0) asm volatile("bar.sync 1, 64;" ::: "memory");
    else asm volatile("bar.sync 2, 64;" ::: "memory");
  } else {
    switch (c) {
      case 0: asm volatile("bar.sync 1, 64;" ::: "memory"); break;
      case 1: asm volatile("bar.sync 2, 64;" ::: "memory"); break;
      case 2: asm volatile("bar.sync 3, 64;" ::: "memory"); break;
      case 3: asm volatile("bar.sync 4, 64;" ::: "memory"); break;
      case 4: asm volatile("bar.sync 5, 64;" ::: "memory"); break;
      case 5: asm volatile("bar.sync 6, 64;" ::: "memory"); break;
      case 6: asm volatile("bar.sync 7, 64;" ::: "memory"); break;
      default: asm volatile("bar.sync 8, 64;" ::: "memory"); break;
    }
  }
}

// ---- mbarrier + TMA bulk copy (global -> shared::cta), one transaction barrier per CTA --------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the TMA unit (async proxy) sees the initialised barrier
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
               "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---------------------------------------------------------------------------------------------------------------------
// Pre-pass: the sequential part of a run, 16 bytes out per frame.  One warp per run, 32 frames per step: each lane loads
// its frame's status words and channel headers, and the three things that depend on earlier frames are resolved with
// warp votes / scans instead of a serial walk -- the instance tag an element showed first (first lane that starts it),
// windowShape[CURRENT] of each channel (last lane below that decoded an ics_info), the PNS generator (prefix sum of the
// draw counts, then an O(log n) jump).
// ---------------------------------------------------------------------------------------------------------------------
constexpr int kK2PreWarps = 4;

__global__ void __launch_bounds__(32 * kK2PreWarps)
k2_prepass_kernel(const RunDev* __restrict__ runs, uint32_t n_runs, const RunFrameDev* __restrict__ run_frames,
                  FrameSide* __restrict__ fside, const IcsSide* __restrict__ iside, StreamState* __restrict__ sstate,
                  const LayoutDev* __restrict__ layouts, const uint64_t* __restrict__ pcm_off, K2FrameDev* __restrict__ out,
                  uint32_t* __restrict__ pcm_bytes_out, int bytes_per_sample, int tns_iso) {
  const uint32_t r = blockIdx.x * kK2PreWarps + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n_runs) return;
  const RunDev run = runs[r];
  const LayoutDev lay = layouts[run.layout];
  const int nch = lay.n_channels;
  StreamState st = sstate[run.stream_slot];
  // element of each channel slot, 3 bits each
  uint32_t el_of = 0;
  for (int e = 0; e < lay.n_elements; ++e) {
    const int f0 = lay.el_first_ch[e], n = lay.el_type[e] == EL_CPE ? 2 : 1;
    for (int c = f0; c < f0 + n; ++c) el_of |= (uint32_t)e << (3 * c);
  }
  // carried along the run (the same in every lane)
  uint32_t shape_cur = 0;
  for (int c = 0; c < nch; ++c) shape_cur |= (uint32_t)(st.window_shape[c] & 1u) << c;
  uint32_t exp_tags = st.tags, exp_mask = 0;
  for (int i = 0; i < 4; ++i) exp_mask |= ((st.tags_valid >> i) & 1u) ? (0xFu << (4 * i)) : 0u;
  uint32_t pns = st.pns_state;
  const uint32_t frame_bytes = (uint32_t)(1024 * (run.mono_dup ? 2 : nch) * bytes_per_sample);
  const uint32_t lt_mask = (1u << lane) - 1u;

  for (uint32_t base = 0; base < run.count; base += 32) {
    const uint32_t it = base + (uint32_t)lane;
    const bool valid = it < run.count;
    RunFrameDev rf{0u, 0u};
    uint4 fsw = make_uint4(0, 0, 0, 0);   // status, tags | n_elements << 16 | n_started << 24, sbr_bit_off[2]
    uint32_t draws = 0, notes = 0;
    uint64_t poff = 0;
    if (valid) {
      rf = run_frames[run.first + it];
      fsw = *reinterpret_cast<const uint4*>(fside + rf.frame);
      const uint2 dn = *reinterpret_cast<const uint2*>(&fside[rf.frame].pns_draws);   // pns_draws, notes
      draws = dn.x;
      notes = dn.y;
      poff = pcm_off[rf.frame];
    }
    const uint32_t next_ics = (it + 1 < run.count) ? run_frames[run.first + it + 1].ics_base : 0u;
    // Element objects are per (type, instance tag) in JAAD (StreamState::tags).  An element that shows another tag than
    // the stream's first one for that element, or that the layout does not have, addresses objects this stream does not
    // own: it leaves them alone, the frame is reported as JAADB_ST_LAYOUT -- and the stream's own elements of that frame
    // still go through the filterbank when the frame was parsed to its end, as they do in JAAD
    // (SyntacticElements.process runs after the whole parse).
    const uint32_t tags = fsw.y & 0xFFFFu;
    const int n_started = (int)min(fsw.y >> 24, 4u), n_good = (int)((fsw.y >> 16) & 0xFFu);
    uint32_t diff = 0;   // nibble e set: element e shows a tag that is not the stream's
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const bool started = valid && e < n_started;
      const uint32_t who = __ballot_sync(0xFFFFFFFFu, started);
      const bool known = ((exp_mask >> (4 * e)) & 1u) != 0;
      const uint32_t tag = (tags >> (4 * e)) & 15u;
      const int first = who ? __ffs((int)who) - 1 : 0;
      const uint32_t first_tag = __shfl_sync(0xFFFFFFFFu, tag, first);
      const uint32_t expect = known ? ((exp_tags >> (4 * e)) & 15u) : first_tag;
      if (started && (known || lane > first) && tag != expect) diff |= 0xFu << (4 * e);
      if (!known && who) { exp_tags |= first_tag << (4 * e); exp_mask |= 0xFu << (4 * e); }
    }
    int frame_status = (int)fsw.x;
    if (valid && diff != 0 && frame_status == 0) {
      frame_status = JAADB_ST_LAYOUT;
      fside[rf.frame].status = JAADB_ST_LAYOUT;
    }
    const bool emit = frame_status == 0;                                        // the frame yields PCM
    const bool parsed = emit || frame_status == JAADB_ST_LAYOUT;                // JAAD reached SyntacticElements.process
    uint32_t flags = (emit ? kK2Emit : 0u) | (parsed ? kK2Parsed : 0u);
    for (int c = 0; c < nch; ++c) {
      uint32_t h0 = 0, tw = 0;
      if (valid) {
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(iside + rf.ics_base + c);
        h0 = sw[0];    // present | info_decoded << 8 | window_sequence << 16 | window_shape << 24
        tw = sw[98];   // tns_present | has_pns << 8 | pns_base << 16
      }
      const int my_el = (int)((el_of >> (3 * c)) & 7u);
      const bool shape_ok = my_el >= 4 || ((diff >> (4 * my_el)) & 15u) == 0;   // the element belongs to the stream
      const bool el_live = shape_ok && my_el < n_good;                         // ... and was parsed completely
      // window-shape bookkeeping of ICSInfo.decode / setCommonData (ICSInfo.java:90-91,196-197): in failing frames too
      bool upd = valid && ((h0 >> 8) & 0xFFu) != 0 && shape_ok;
      uint32_t sbit = (h0 >> 24) & 1u;
      // A frame that ended in an error can have decoded an element into THIS element's object from another place of the
      // frame (SyntacticElements.java:39-41: one object per type and tag): an element of the same type at another position
      // of the layout that shows this element's tag (K1 parsed it into that position's slots), or an element outside the
      // layout (K1's notes, note_dup_shape).  Its ics_info moved this object's window shape like the element's own does;
      // the later one in the bitstream wins.  Rare (damaged element ids / tags): the lanes diverge, no votes inside.
      if (valid && fsw.x != 0u && fsw.x != (uint32_t)JAADB_ST_LAYOUT && (diff | notes) != 0u && my_el < 4 && ((exp_mask >> (4 * my_el)) & 1u)) {
        const uint32_t my_type = lay.el_type[my_el], my_tag = (exp_tags >> (4 * my_el)) & 15u;
        const int k = c - lay.el_first_ch[my_el];
        for (int e = 0; e < n_started; ++e) {
          if (e == my_el || !((diff >> (4 * e)) & 15u) || lay.el_type[e] != my_type || ((tags >> (4 * e)) & 15u) != my_tag) continue;
          const uint32_t hs = *reinterpret_cast<const uint32_t*>(iside + rf.ics_base + lay.el_first_ch[e] + k);
          if (((hs >> 8) & 0xFFu) != 0 && (!upd || e > my_el)) { upd = true; sbit = (hs >> 24) & 1u; }
        }
        for (int n = 0; n < 3; ++n) {
          const uint32_t nt = notes >> (9 * n);
          if ((nt & 256u) && (nt & 3u) == my_type && ((nt >> 2) & 15u) == my_tag && (int)((nt >> 6) & 1u) == k) { upd = true; sbit = (nt >> 7) & 1u; }
        }
      }
      const uint32_t U = __ballot_sync(0xFFFFFFFFu, upd), S = __ballot_sync(0xFFFFFFFFu, upd && sbit);
      const uint32_t below = U & lt_mask;
      const uint32_t prev = below ? ((S >> (31 - __clz((int)below))) & 1u) : ((shape_cur >> c) & 1u);
      const uint32_t cur = upd ? sbit : prev;
      flags |= prev << (8 + c) | cur << (16 + c);
      if (U) shape_cur = (shape_cur & ~(1u << c)) | (((S >> (31 - __clz((int)U))) & 1u) << c);
      // (SBR streams: the SBR stages only run for frames that yield PCM, so the core coder's state waits for them too)
      if (valid && parsed && el_live && (emit || !run.sbr)) {
        flags |= 1u << c;
        if (tns_iso && (tw & 0xFFu)) flags |= kK2Tns;
        if (tw & 0xFF00u) flags |= kK2Pns;
      }
    }
    // PNS generator at the start of each frame: prefix sum of the draws, then jump
    uint32_t incl = draws;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
      if (lane >= d) incl += v;
    }
    const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
    if (valid) {
      uint4* o = reinterpret_cast<uint4*>(out + run.first + it);
      o[0] = make_uint4(rf.frame, rf.ics_base, flags, total ? pns_jump(pns, incl - draws) : pns);
      o[1] = make_uint4((uint32_t)poff, (uint32_t)(poff >> 32), next_ics, 0u);
      if (!run.sbr) pcm_bytes_out[rf.frame] = emit ? frame_bytes : 0u;
    }
    if (total) pns = pns_jump(pns, total);
  }
  if (lane == 0) {
    for (int c = 0; c < nch; ++c) st.window_shape[c] = (uint8_t)((shape_cur >> c) & 1u);
    uint32_t v = 0;
    for (int i = 0; i < 4; ++i) v |= ((exp_mask >> (4 * i)) & 1u) << i;
    st.tags = (uint16_t)exp_tags;
    st.tags_valid = (uint8_t)v;
    st.pns_state = pns;
    sstate[run.stream_slot] = st;
  }
}

// Final overlap of segmented runs: the last segment of a run leaves it in a staging buffer (other segments of the run may
// still be reading the persistent one), this kernel moves it into the stream's state.
__global__ void k2_commit_kernel(const RunDev* __restrict__ runs, uint32_t n_runs, int nch, const float* __restrict__ stage,
                                 float* __restrict__ overlap_all) {
  const uint32_t r = blockIdx.x;
  if (r >= n_runs) return;
  const float4* src = reinterpret_cast<const float4*>(stage + (size_t)r * kMaxChannels * 1024);   // (stage: this group's first run)
  float4* dst = reinterpret_cast<float4*>(overlap_all + (size_t)runs[r].stream_slot * kMaxChannels * 1024);
  for (int i = threadIdx.x; i < nch * 256; i += blockDim.x) dst[i] = src[i];
}

struct K2Args {
  const K2SegDev* segs;
  const RunDev* runs;
  const K2FrameDev* k2frames;
  const FrameDev* frames;       // JAADB_TNS_ISO only: where a frame's bytes are
  const uint8_t* blob;          // "
  const IcsSide* iside;
  const int16_t* qall;
  float* overlap_all;
  float* overlap_stage;         // segmented runs: [run][kMaxChannels][1024], moved into overlap_all by k2_commit_kernel; else null
  uint8_t* pcm;
  float* spec_tap;
  float* core;
  const LayoutDev* layouts;
  int nch;
};

template <int PCM_FORMAT, int MAX_THREADS, int MIN_BLOCKS>
__global__ void __launch_bounds__(MAX_THREADS, MIN_BLOCKS)
k2_filterbank_kernel(const K2Args A, const TablesDev T) {
  // the 128-thread instantiation serves one- and two-channel streams: always two output channels (mono is duplicated,
  // SyntacticElements.java:244-245), PCM packed per channel and interleaved on the way out
  constexpr bool kPlanarPcm = MAX_THREADS == 128;
  extern __shared__ __align__(16) float smem[];
  const int nch = A.nch;
  // carve: twiddles; per channel [spec][overlap][xre][xim]; stage q[nch][1024], side[nch]; mbarrier; [pcm staging]
  float* s_fft_tw = smem;                         // fft512 re/im (inverse) [256][2] + fft64 [32][2]
  float* s_ch = s_fft_tw + 2 * 256 + 2 * 32;
  int16_t* s_q = reinterpret_cast<int16_t*>(s_ch + nch * kK2ChFloats);
  IcsSide* s_side = reinterpret_cast<IcsSide*>(s_q + nch * 1024);
  K2FrameDev* s_kf = reinterpret_cast<K2FrameDev*>(s_side + nch);   // the staged frame's record (16-byte aligned: 400 B per side)
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(s_kf + 1);
  int16_t* s_pcm = reinterpret_cast<int16_t*>(s_bar + 2);  // s16 formats: [1024][out_ch] interleaved (more than two channels)
                                                           // or, kPlanarPcm, [nch][512] words = sample pairs (i, i+1) of a channel

  const K2SegDev seg = A.segs[blockIdx.x];
  const RunDev run = A.runs[seg.run];
  const LayoutDev lay = A.layouts[run.layout];
  const int tid = threadIdx.x;
  const int c = tid / kThreadsPerChannel;         // channel slot of this thread
  const int t = tid - c * kThreadsPerChannel;
  const int nthreads = blockDim.x;
  const int out_ch = run.mono_dup ? 2 : nch;
  const int sf_index = run.sf_index;
  const K2FrameDev* __restrict__ kf = A.k2frames + run.first;

  float* my_spec = s_ch + c * kK2ChFloats;
  float* my_ovl = my_spec + kSpecStride;
  float* my_xre = my_ovl + 1024;
  float* my_xim = my_xre + kXchgStride;

  // Where to start.  The overlap a channel holds after a frame depends on that frame alone, so a segment in the middle
  // of a run first re-runs (without output) the frames before it, back to where every channel has been through the
  // filterbank once; channels that never were keep the persistent overlap.
  uint32_t it0 = seg.first;
  if (it0 > 0) {
    uint32_t have = 0;
    const uint32_t all = (1u << nch) - 1u;
    while (it0 > 0 && have != all) { --it0; have |= kf[it0].flags & 0xFFu; }
  }
  const uint32_t it_end = seg.first + seg.count;
  const uint32_t stage_bytes = (uint32_t)nch * 2048u, side_bytes = (uint32_t)nch * (uint32_t)sizeof(IcsSide);
  constexpr uint32_t kf_bytes = (uint32_t)sizeof(K2FrameDev);
  if (tid == 0) {
    mbar_init(s_bar, 1);
    const uint32_t ics0 = kf[it0].ics_base;
    mbar_expect_tx(s_bar, stage_bytes + side_bytes + kf_bytes);
    tma_bulk_g2s(s_q, A.qall + (size_t)ics0 * 1024, stage_bytes, s_bar);
    tma_bulk_g2s(s_side, A.iside + ics0, side_bytes, s_bar);
    tma_bulk_g2s(s_kf, kf + it0, kf_bytes, s_bar);
  }

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
  const float2* mdct_long_g = reinterpret_cast<const float2*>(T.mdct_long_gather);
  const float2* mdct_short_g = reinterpret_cast<const float2*>(T.mdct_short_gather);

  // persistent state in: overlap
  float* g_ovl = A.overlap_all + ((size_t)run.stream_slot * kMaxChannels + c) * 1024;
  for (int i = t; i < 256; i += kThreadsPerChannel)
    reinterpret_cast<float4*>(my_ovl)[i] = reinterpret_cast<const float4*>(g_ovl)[i];

  // element of this thread's channel
  int el_first = c, el_nch = 1;
  for (int e = 0; e < lay.n_elements; ++e) {
    int f0 = lay.el_first_ch[e];
    int n = lay.el_type[e] == EL_CPE ? 2 : 1;
    if (c >= f0 && c < f0 + n) { el_first = f0; el_nch = n; }
  }
  // dequantisation work split: a CPE thread owns coefficients 8*et .. 8*et+7 of L and of R (M/S and IS are
  // element-wise across the pair); an SCE/LFE thread owns 16*et .. 16*et+15 of its channel.
  const int et = tid - el_first * kThreadsPerChannel;
  const int chA = el_first, chB = el_first + (el_nch == 2 ? 1 : 0);
  const int iA = (el_nch == 2) ? 8 * et : 16 * et;
  const int iB = (el_nch == 2) ? iA : iA + 8;
  const uint32_t pkA0 = subchunk_consts(T, sf_index, iA), pkA1 = subchunk_consts(T, sf_index, iA + 4);
  const uint32_t pkB0 = subchunk_consts(T, sf_index, iB), pkB1 = subchunk_consts(T, sf_index, iB + 4);

  __syncthreads();   // twiddles, overlap and the barrier's initialisation are visible

  for (uint32_t it = it0; it < it_end; ++it) {
    // the frame's record, quantised coefficients and side information have landed in the stage
    mbar_wait(s_bar, (it - it0) & 1u);
    const uint4 kf0 = reinterpret_cast<const uint4*>(s_kf)[0], kf1 = reinterpret_cast<const uint4*>(s_kf)[1];
    const uint32_t f = kf0.x;
    const uint32_t ics_base = kf0.y;
    const uint32_t flags = kf0.z;
    const uint32_t pns_state = kf0.w;
    const uint64_t poff = (uint64_t)kf1.x | ((uint64_t)kf1.y << 32);
    const uint32_t next_ics_base = kf1.z;
    const bool have_next = it + 1 < it_end;
    const bool warm = it < seg.first;                         // re-run for the overlap only
    const bool emit = (flags & kK2Emit) != 0 && !warm;        // the frame yields PCM
    const bool parsed = (flags & kK2Parsed) != 0;             // JAAD reached SyntacticElements.process
    const bool run_ch = ((flags >> c) & 1u) != 0;             // this thread's channel goes through the filterbank
    const int shape_prev = (int)((flags >> (8 + c)) & 1u), shape_cur = (int)((flags >> (16 + c)) & 1u);
    const int ws = s_side[c].window_sequence;

    if (run_ch) {   // (the channels of an element run together)
      // ---- phase 1: dequantise + M/S + IS into the element's spectra
      const int16_t* qL = s_q + chA * 1024;
      const int16_t* qR = s_q + chB * 1024;
      const IcsSide* sL = s_side + chA;
      const IcsSide* sR = s_side + chB;
      float* specA = s_ch + chA * kK2ChFloats;
      float* specB = s_ch + chB * kK2ChFloats;
      // (the eight coefficients of a channel lie in one short window)
      const uint32_t grpA = sL->window_sequence == 2 ? short_group_of(sL, (int)((pkA0 >> 24) & 7u)) : 0u;
      const uint32_t grpB = sR->window_sequence == 2 ? short_group_of(sR, (int)((pkB0 >> 24) & 7u)) : 0u;
      const bool ms_on = el_nch == 2 && sL->common_window && sL->ms_mask != 0;   // CPE.java:159-160
      const bool ms_present = sL->ms_mask != 0;                                   // CPE.isMSMaskPresent
      // two rounds of four coefficients per channel, as a real loop: the kernel's instruction footprint per frame is what
      // the instruction caches see (32 KB of L1.5), and this phase was a third of it when unrolled
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const uint32_t pkA = h ? pkA1 : pkA0, pkB = h ? pkB1 : pkB0;
        const int i0A = iA + 4 * h, i0B = iB + 4 * h;
        float a[4], b[4];
        int cbL, cbR, idxL, idxR;
        dequant4(sL, qL, grpA, pkA, i0A, T, a, cbL, idxL);
        dequant4(sR, qR, grpB, pkB, i0B, T, b, cbR, idxR);
        if (flags & kK2Pns) {
          // generator state when this channel's parse began: the frame's + what the earlier channels took
          if (cbL == 13) {
            const int16_t* swb_ = sL->window_sequence == 2 ? (T.swb_short + sf_index * 17) : (T.swb_long + sf_index * 53);
            const float4 n_ = pns_fill4(sL, grpA, pkA, i0A, idxL, pns_jump(pns_state, sL->pns_base), swb_, T.sf);
            a[0] = n_.x; a[1] = n_.y; a[2] = n_.z; a[3] = n_.w;
          }
          if (cbR == 13) {
            const int16_t* swb_ = sR->window_sequence == 2 ? (T.swb_short + sf_index * 17) : (T.swb_long + sf_index * 53);
            const float4 n_ = pns_fill4(sR, grpB, pkB, i0B, idxR, pns_jump(pns_state, sR->pns_base), swb_, T.sf);
            b[0] = n_.x; b[1] = n_.y; b[2] = n_.z; b[3] = n_.w;
          }
        }
        if (el_nch == 2) {
          // MS.process: both codebooks < NOISE_HCB, band flagged (MS.java:28-36).  Bands at or above max_sfb have
          // cb_out 0 and idx 0: they hold zeros, for which the butterfly is the identity up to the sign of zero --
          // JAAD never touches them, so they are excluded through the left channel's band test.
          const bool in_band = ((sL->window_sequence == 2) ? (int)((pkA >> 6) & 15u) : (int)(pkA & 63u)) < sL->max_sfb;
          if (ms_on && cbL < 13 && cbR < 13 && in_band && ((sL->ms_used[idxL >> 3] >> (idxL & 7)) & 1)) {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float l = a[j], r = b[j];
              a[j] = l + r;
              b[j] = l - r;
            }
          }
          // IS.process: right channel bands with codebook 14/15 (IS.java:29-44)
          if (cbR == 15 || cbR == 14) {
            int sgn = cbR == 15 ? 1 : -1;
            if (ms_present) sgn *= ((sL->ms_used[idxR >> 3] >> (idxR & 7)) & 1) ? -1 : 1;
            float scale = __ldg(T.sf + sR->sf_idx[idxR]);
            if (sgn < 0) scale = -scale;
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = a[j] * scale;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          specA[spec_addr(i0A + j)] = a[j];
          specB[spec_addr(i0B + j)] = b[j];
        }
        if (A.spec_tap && !warm && !(flags & kK2Tns)) {
          float* tA = A.spec_tap + ((size_t)ics_base + chA) * 1024 + i0A;
          float* tB = A.spec_tap + ((size_t)ics_base + chB) * 1024 + i0B;
#pragma unroll
          for (int j = 0; j < 4; ++j) { tA[j] = a[j]; tB[j] = b[j]; }
        }
      }
    }
    if (flags & kK2Tns) {
      // ISO TNS between the stereo tools and the filterbank: one thread per (window, filter), in place on the spectrum
      __syncthreads();
      if (run_ch && s_side[c].tns_present && t < 32) {
        const uint64_t addr = reinterpret_cast<uint64_t>(A.blob) + A.frames[f].blob_off;
        const bool sh = ws == 2;
        tns_iso_filter(my_spec, s_side + c, reinterpret_cast<const uint32_t*>(addr - (addr & 3u)),
                       sh ? (T.swb_short + sf_index * 17) : (T.swb_long + sf_index * 53),
                       sh ? T.swb_short_count[sf_index] : T.swb_long_count[sf_index], c_tns_max_sfb[sf_index * 2 + (sh ? 1 : 0)], t >> 2, t & 3);
      }
      if (A.spec_tap && !warm) {
        __syncthreads();
        if (run_ch)
          for (int i = t; i < 1024; i += kThreadsPerChannel) A.spec_tap[((size_t)ics_base + c) * 1024 + i] = my_spec[spec_addr(i)];
      }
    }
    __syncthreads();   // spectra complete; every thread is done with the stage
    // ---- the next frame comes in while this one is transformed
    if (tid == 0 && have_next) {
      mbar_expect_tx(s_bar, stage_bytes + side_bytes + kf_bytes);
      tma_bulk_g2s(s_q, A.qall + (size_t)next_ics_base * 1024, stage_bytes, s_bar);
      tma_bulk_g2s(s_side, A.iside + next_ics_base, side_bytes, s_bar);
      tma_bulk_g2s(s_kf, kf + it + 1, kf_bytes, s_bar);
    }
    if (!parsed) {
      // the frame produced no PCM; overlap untouched (Decoder.java:96-98).  Its slot of the output is zero-filled.
      if (!warm && !run.sbr) {
        uint32_t* d = reinterpret_cast<uint32_t*>(A.pcm + poff);
        for (int i = tid; i < 512 * out_ch * (PCM_FORMAT == 2 ? 2 : 1); i += nthreads) d[i] = 0u;
      }
      continue;
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
          const float2 cs = __ldg(mdct_long_g + j * 64 + t);   // = mdct_long2[k]
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
          const float2 cs = __ldg(mdct_short_g + j * 8 + (t & 7));   // = mdct_short2[k]
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
    channel_barrier<MAX_THREADS / kThreadsPerChannel>(c);
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
    // The spectrum was consumed before the channel barrier above, so its storage now carries exchange 2 (planes re | im,
    // 8 blocks of 72 floats each: 8 pad floats per 64 keep both the strided writes and the contiguous reads conflict-free;
    // the last block's pad is never touched, so the planes sit 568 apart -- those 64 bytes per channel are what lets a
    // sixth two-channel CTA fit the SM); the post-twiddled buffer then goes where exchange 1 was.
    float* x2re = my_spec;
    float* x2im = my_spec + 568;
    float* bre = my_xre;               // post-twiddled buffer as planes re[512] | im[512] (short: 8 windows of 64 at stride 72)
    float* bim = my_xim;
    if (!is_short) {
      // exchange 2: write n = 64*blk + col + 8*j, read n = t + 64*j
      const int blk = t >> 3, col = t & 7;
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        x2re[72 * blk + col + 8 * j] = a[j].re;
        x2im[72 * blk + col + 8 * j] = a[j].im;
      }
    }
    channel_barrier<MAX_THREADS / kThreadsPerChannel>(c);   // also: every thread of the channel is done reading exchange 1
    if (!is_short) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        a[j].re = x2re[t + 72 * j];
        a[j].im = x2im[t + 72 * j];
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
        bim[72 * blk + nl] = (t1 * cs.x) + (t0 * cs.y);
        bre[72 * blk + nl] = (t0 * cs.x) - (t1 * cs.y);
      }
    }
    channel_barrier<MAX_THREADS / kThreadsPerChannel>(c);

    // ---- phase 3: windowing + overlap-add (FilterBank.java:39-123) + PCM.  Thread t owns the sample pairs
    // i = 2t + 128j, i+1 (j = 0..7): the MDCT reorder (MDCT.java:60-80) then reads mirrored positions of the two
    // planes without any per-lane case split, and the overlap goes back as float2.
    {
      const float* __restrict__ LWp = T.win_long[shape_prev];
      const float* __restrict__ LW = T.win_long[shape_cur];
      const float* __restrict__ SWp = T.win_short[shape_prev];
      const float* __restrict__ SW = T.win_short[shape_cur];
      uint8_t* dst = A.pcm + poff;
      uint32_t* my_pk = reinterpret_cast<uint32_t*>(s_pcm) + 512 * c;   // kPlanarPcm: sample pairs (i, i+1) of this channel, word i/2
      // (requesting all of a frame's window values ahead of this loop was measured: the 32 extra registers cost more -- spills at
      //  96 registers, instruction-cache misses at 128 -- than the L1 latency they hide)
#ifndef K2_P3_UNROLL
#define K2_P3_UNROLL 8
#endif
      constexpr int kP3Unroll = K2_P3_UNROLL;
      // long windows: the values of round j + 1 are requested before round j computes (the loads sat right in front of their
      // use behind the sequence branches: the largest long-scoreboard stall of the kernel)
      const bool pre_rise = !is_short && ws != 3, pre_fall = !is_short && ws != 1;
      float2 w_rise_n = make_float2(0.f, 0.f), w_fall_n = make_float2(0.f, 0.f);
      if (pre_rise) w_rise_n = __ldg(reinterpret_cast<const float2*>(LWp + 2 * t));
      if (pre_fall) w_fall_n = __ldg(reinterpret_cast<const float2*>(LW + 1022 - 2 * t));
#pragma unroll kP3Unroll
      for (int j = 0; j < 8; ++j) {
        const int i = 2 * t + 128 * j;
        const float2 w_rise = w_rise_n, w_fall = w_fall_n;
        if (j < 7) {
          if (pre_rise) w_rise_n = __ldg(reinterpret_cast<const float2*>(LWp + i + 128));
          if (pre_fall) w_fall_n = __ldg(reinterpret_cast<const float2*>(LW + 1022 - i - 128));
        }
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
            const float2 w = w_rise;
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
            const float2 w = w_fall;
            n0 = x20 * w.y; n1 = x21 * w.x;
          }
        } else {
          // EIGHT_SHORT: window w occupies b[256w .. 256w+255]; its samples come from FFT block w (planes at stride 72)
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
                const float a2 = mdct_out(bre + 72 * (s - 1), bim + 72 * (s - 1), 64, 32, 128 + r) * __ldg(SW + 127 - r);
                const float b2 = mdct_out(bre + 72 * s, bim + 72 * s, 64, 32, r) * __ldg(SW + r);
                o = (ovp + a2) + b2;
              }
            }
            if (ii >= 576) nv = 0.f;
            else if (ii < 64) {
              // overlap[i], i in [0,64): window 3 second half (r = 64+i) + window 4 first half
              const int r = 64 + ii;
              nv = (mdct_out(bre + 72 * 3, bim + 72 * 3, 64, 32, 128 + r) * __ldg(SW + 127 - r)) +
                   (mdct_out(bre + 72 * 4, bim + 72 * 4, 64, 32, r) * __ldg(SW + r));
            } else if (ii < 448) {
              // i = 64 + 128*u + r: window 4+u second half + window 5+u first half (u = 0..2)
              const int u = (ii - 64) >> 7, r = (ii - 64) & 127;
              nv = (mdct_out(bre + 72 * (4 + u), bim + 72 * (4 + u), 64, 32, 128 + r) * __ldg(SW + 127 - r)) +
                   (mdct_out(bre + 72 * (5 + u), bim + 72 * (5 + u), 64, 32, r) * __ldg(SW + r));
            } else {
              // i in [448,576): window 7 second half only
              const int r = ii - 448;
              nv = mdct_out(bre + 72 * 7, bim + 72 * 7, 64, 32, 128 + r) * __ldg(SW + 127 - r);
            }
            oo[p] = o;
            nn[p] = nv;
          }
          o0 = oo[0]; o1 = oo[1]; n0 = nn[0]; n1 = nn[1];
        }
        if (run_ch) *reinterpret_cast<float2*>(my_ovl + i) = make_float2(n0, n1);
        if (!emit) {
          // a frame JAAD decodes against element objects this stream does not own, or a re-run: state only, no PCM
        } else if (run.sbr) {
          // core-coder output of an SBR stream: K4 continues from here
          *reinterpret_cast<float2*>(A.core + ((size_t)ics_base + c) * 1024 + i) = make_float2(o0, o1);
        } else if (PCM_FORMAT == 2) {
          float* d = reinterpret_cast<float*>(dst);
          *reinterpret_cast<float2*>(d + (size_t)c * 1024 + i) = make_float2(o0, o1);
          if (run.mono_dup) *reinterpret_cast<float2*>(d + 1024 + i) = make_float2(o0, o1);
        } else {
          uint32_t pr = pcm_round16(o0) | (pcm_round16(o1) << 16);
          if (PCM_FORMAT == 1) pr = __byte_perm(pr, 0, 0x2301);
          if (kPlanarPcm) {
            my_pk[t + 64 * j] = pr;   // (the copy-out of the previous frame finished before this frame's phase-1 barrier)
          } else {
            s_pcm[i * out_ch + c] = (int16_t)(pr & 0xFFFFu);
            s_pcm[(i + 1) * out_ch + c] = (int16_t)(pr >> 16);
          }
        }
      }
      __syncthreads();
      if (PCM_FORMAT != 2 && !run.sbr && (flags & kK2Emit) && !warm) {
        // coalesced copy-out of the interleaved frame (pcm offsets are 4-byte aligned; 16 B when the caller packs)
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        if (kPlanarPcm) {
          // samples i..i+3 of both channels = words i/2, i/2+1 of the two planes (one plane twice when mono is duplicated)
          const uint2* pl = reinterpret_cast<const uint2*>(s_pcm);
          const uint2* prr = pl + (nch == 2 ? 256 : 0);
          const bool al = (reinterpret_cast<uintptr_t>(dst) & 15u) == 0;
          for (int i = tid; i < 256; i += nthreads) {
            const uint2 l = pl[i], r = prr[i];
            const uint4 o = make_uint4(__byte_perm(l.x, r.x, 0x5410), __byte_perm(l.x, r.x, 0x7632), __byte_perm(l.y, r.y, 0x5410),
                                       __byte_perm(l.y, r.y, 0x7632));
            if (al) reinterpret_cast<uint4*>(d)[i] = o;
            else { d[4 * i] = o.x; d[4 * i + 1] = o.y; d[4 * i + 2] = o.z; d[4 * i + 3] = o.w; }
          }
        } else {
          const int nwords = 1024 * out_ch / 2;   // 32-bit words
          const uint32_t* src = reinterpret_cast<const uint32_t*>(s_pcm);
          if ((reinterpret_cast<uintptr_t>(dst) & 15u) == 0) {
            for (int i = tid; i < nwords / 4; i += nthreads)
              reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(src)[i];
          } else {
            for (int i = tid; i < nwords; i += nthreads) d[i] = src[i];
          }
        }
      } else if (!(flags & kK2Emit) && !warm && !run.sbr) {
        uint32_t* d = reinterpret_cast<uint32_t*>(dst);
        for (int i = tid; i < 512 * out_ch * (PCM_FORMAT == 2 ? 2 : 1); i += nthreads) d[i] = 0u;   // no PCM: the slot is zero-filled
      }
    }
  }

  // persistent state out (the last frame's phase 3 wrote my_ovl; the loop's trailing barrier ordered it).  Only the
  // run's last segment holds the final overlap.
  __syncthreads();
  if (it_end == run.count) {
    float* o = A.overlap_stage ? A.overlap_stage + ((size_t)seg.run * kMaxChannels + c) * 1024 : g_ovl;
    for (int i = t; i < 256; i += kThreadsPerChannel)
      reinterpret_cast<float4*>(o)[i] = reinterpret_cast<const float4*>(my_ovl)[i];
  }
}

}  // namespace jaadb
