// K1 -- noiseless decode.  One thread parses one AAC frame (raw_data_block):
// element loop, ics_info, section data, scalefactors, pulse/TNS side info and the
// spectral Huffman codewords, with the codebook LUTs staged in shared memory.
// 32 frames advance per warp; frames are independent at this stage (the only
// cross-frame state of AAC-LC -- window_shape[PREVIOUS] and the IMDCT overlap --
// is resolved in K2), so a batch of T frames of one stream parses in parallel.
//
// Output per channel-frame: quantised coefficients q[1024] (int16, windows
// de-interleaved like ICStream.decodeSpectralData) and an IcsSide record.
// Reference behaviour followed (paths relative to aac/src/main/java/net/sourceforge/jaad/aac/):
//   syntax/SyntacticElements.java:57-203   element loop, FIL, DSE
//   syntax/CPE.java:85-123                 common_window, ms_mask
//   syntax/ICSInfo.java:86-119,193-211     ics_info
//   syntax/ICStream.java:60-275            section / scalefactor / pulse / spectral data
//   huffman/Huffman.java:15-84             codeword, sign bits, escape
//   tools/TNS.java:35-61                   TNS side info (checked here; K2 reads the coefficients in JAADB_TNS_ISO mode)
#pragma once
#include "jaadb_types.cuh"

namespace jaadb {

#define JAADB_ST_OK 0
#define JAADB_ST_EOS 1
#define JAADB_ST_INVALID_CODEBOOK 2
#define JAADB_ST_TOO_MANY_BANDS 3
#define JAADB_ST_SF_RANGE 4
#define JAADB_ST_PULSE_SHORT 5
#define JAADB_ST_PULSE_RANGE 6
#define JAADB_ST_MS_RESERVED 7
#define JAADB_ST_TNS_ORDER 8
#define JAADB_ST_LTP_PROFILE 9
#define JAADB_ST_UNSUPPORTED_ELEMENT 10
#define JAADB_ST_LAYOUT 11
#define JAADB_ST_PROFILE 12
#define JAADB_ST_ARRAY_BOUNDS 13

// MSB-first bit reader over 32-bit words of the blob (ByteArrayBitStream.java semantics:
// a read that would pass the end of the frame is an EOSException; here the reader runs on
// and the caller tests overrun() at element/error boundaries, which yields the same status).
struct BitReader {
  const uint32_t* __restrict__ words;  // 4-byte aligned, <= frame start
  uint32_t pos;                        // bit position relative to words
  uint32_t end;                        // first bit after the frame
  uint32_t last_word;                  // index of the last word that holds frame bits
  uint32_t widx;
  uint32_t w0, w1;                     // words widx, widx + 1 (big-endian order restored)
  uint32_t w2;                         // word widx + 2 as loaded: requested a word ahead of its first use and not touched
                                       // (no byte swap) until then, so the in-order pipe never waits on the request

  __device__ __forceinline__ uint32_t load_raw(uint32_t i) const { return (i <= last_word) ? __ldg(words + i) : 0u; }
  __device__ __forceinline__ uint32_t load(uint32_t i) const { return __byte_perm(load_raw(i), 0, 0x0123); }
  __device__ __forceinline__ void init(const uint8_t* blob, uint64_t off, uint32_t nbytes) {
    uint64_t addr = reinterpret_cast<uint64_t>(blob) + off;
    uint32_t mis = (uint32_t)(addr & 3u);
    words = reinterpret_cast<const uint32_t*>(addr - mis);
    pos = mis * 8u;
    end = pos + nbytes * 8u;
    last_word = (end - 1u) >> 5;
    widx = pos >> 5;
    w0 = load(widx);
    w1 = load(widx + 1);
    w2 = load_raw(widx + 2);
#pragma unroll
    for (uint32_t k = 8; k <= 64; k += 8)
      if (widx + k <= last_word) asm volatile("prefetch.global.L2 [%0];" ::"l"(words + widx + k));
  }
  // the next 32 bits, left aligned
  __device__ __forceinline__ uint32_t peek() {
    uint32_t wi = pos >> 5;
    if (wi != widx) {
      // sequential walk: the word that becomes w1 was requested a word ago, so its L2 latency (the frames of a warp's 32
      // lanes thrash L1) is behind the decoding of the previous 32 bits instead of in front of the next funnel shift
      if (wi == widx + 1) { w0 = w1; w1 = __byte_perm(w2, 0, 0x0123); }
      else { w0 = load(wi); w1 = load(wi + 1); }
      w2 = load_raw(wi + 2);
      widx = wi;
      // one 32-byte sector, 256 bytes ahead of the read position, per sector consumed: the compressed frame
      // streams through L2 ahead of the serial Huffman walk instead of stalling it on HBM
      if ((wi & 7u) == 0u && wi + 64u <= last_word) asm volatile("prefetch.global.L2 [%0];" ::"l"(words + wi + 64u));
    }
    return __funnelshift_l(w1, w0, pos & 31u);
  }
  __device__ __forceinline__ uint32_t read(int n) {  // 1 <= n <= 32
    uint32_t v = peek() >> (32 - n);
    pos += n;
    return v;
  }
  __device__ __forceinline__ uint32_t read1() { return read(1); }
  __device__ __forceinline__ void skip(uint32_t n) { pos += n; }
  __device__ __forceinline__ bool overrun() const { return pos > end; }
  __device__ __forceinline__ uint32_t bits_left() const { return end > pos ? end - pos : 0u; }
};

__device__ __forceinline__ uint32_t huff_lookup(const uint32_t* lut, uint32_t base, int first_bits, uint32_t w) {
  uint32_t e = lut[base + (w >> (32 - first_bits))];
  if (e & 0x100u) {
    uint32_t x = e & 31u;
    e = lut[(e >> 16) + ((w << first_bits) >> (32 - x))];
  }
  return e;
}

struct IcsInfoRegs {
  int ws, shape, max_sfb, ngroups;
  int shape_ok;          // the window_shape bit lay inside the frame: ICSInfo.decode got as far as storing it
  uint32_t glen_packed;  // 8 x 4 bits
  __device__ __forceinline__ int glen(int g) const { return (glen_packed >> (4 * g)) & 15; }
};

// ICSInfo.decode (ICSInfo.java:86-119).  Returns a status.
__device__ __forceinline__ int parse_ics_info(BitReader& br, IcsInfoRegs& in) {
  br.skip(1);
  in.ws = (int)br.read(2);
  in.shape = (int)br.read(1);
  // windowShape[CURRENT] = in.readBit() (ICSInfo.java:91) throws before the assignment when the frame ends here; the
  // copy into windowShape[PREVIOUS] that precedes it is undone by the next frame's own copy
  in.shape_ok = br.overrun() ? 0 : 1;
  in.ngroups = 1;
  in.glen_packed = 1u;
  if (in.ws == 2) {
    in.max_sfb = (int)br.read(4);
    uint32_t grouping = br.read(7);
    int g = 0;
    for (int i = 6; i >= 0; --i) {
      if ((grouping >> i) & 1u) in.glen_packed += 1u << (4 * g);
      else { ++g; in.glen_packed |= 1u << (4 * g); }
    }
    in.ngroups = g + 1;
  } else {
    in.max_sfb = (int)br.read(6);
    if (br.read1()) return JAADB_ST_LTP_PROFILE;  // predictor_data_present on an LC stream (ICSInfo.java:121-141)
  }
  return JAADB_ST_OK;
}

__device__ __forceinline__ void store_ics_header(IcsSide* s, const IcsInfoRegs& in, int present, int info_decoded,
                                                 int ms_mask, int common) {
  uint32_t h0 = (uint32_t)present | ((uint32_t)(info_decoded & in.shape_ok) << 8) | ((uint32_t)in.ws << 16) | ((uint32_t)in.shape << 24);
  uint32_t h1 = (uint32_t)in.max_sfb | ((uint32_t)in.ngroups << 8) | ((uint32_t)ms_mask << 16) | ((uint32_t)common << 24);
  uint32_t g0 = 0, g1 = 0;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    g0 |= (uint32_t)in.glen(g) << (8 * g);
    g1 |= (uint32_t)in.glen(g + 4) << (8 * g);
  }
  *reinterpret_cast<uint4*>(s) = make_uint4(h0, h1, g0, g1);
}

// ---------------------------------------------------------------------------------------------------------
// Warp-convergent frame parser.
//
// One thread owns one frame, but the 32 frames of a warp advance through the SAME instruction stream: every
// loop below runs until a warp vote says no lane needs another iteration, and a lane that has nothing to do in
// an iteration idles under a predicate.  There is no early return anywhere -- a lane that hits an error (or has
// no frame at all) records its status, goes inactive and keeps voting -- so the warp never splits into
// independently scheduled fragments.  The spectral loop, where ~90 % of the work is, executes one Huffman
// codeword per lane per iteration regardless of window sequence, group, band or codebook.
//
// q is written in BITSTREAM order: group g starts at 128 * (first window of g); inside the group, scalefactor
// band sfb starts at glen * swb[sfb] and holds glen windows x width coefficients, window-major
// (ICStream.java:258-271 walks the data in exactly this order).  K2 undoes the interleave when it dequantises.
// Bands with codebook 0 / 13 / 14 / 15 are not written and never read.
// ---------------------------------------------------------------------------------------------------------
constexpr unsigned kFullMask = 0xFFFFFFFFu;
constexpr int kSwbTableEntries = 12 * 53 + 12 * 17 + 4;   // int16 entries (+4 keeps what follows 8-byte aligned)
constexpr int kK1Threads = 256;
// frames per warp (log2) for a batch of n frames on a GPU that holds `resident_warps` K1 warps at a time
inline uint32_t k1_lanes_log2(uint32_t n_frames, uint32_t resident_warps) {
  uint32_t l = 0;
  while (l < 5 && ((uint64_t)resident_warps << l) < n_frames) ++l;
  return l;
}
__host__ __device__ constexpr size_t k1_smem_bytes(uint32_t lut_entries) {
  return (size_t)lut_entries * 4 + kSwbTableEntries * 2 + (size_t)(kMaxSfbEntries + 8) * kK1Threads;
}

__device__ __forceinline__ void fail(int& status, bool& flag, int code) { status = code; flag = false; }

// JAAD keeps ONE object per (element type, instance tag) (SyntacticElements.java:39-41).  A damaged frame can show an element
// the stream's layout does not have at that position -- a lost bit in an element id or a tag -- and that element can name
// one of the stream's OWN objects: the one an earlier element of this frame was decoded into (its type and tag a second
// time), or one whose own element the frame has not reached yet (a 5.1 frame whose leading SCE id reads CPE, tag 0).  JAAD
// decodes the element into that object, so its ics_info moves the object's window shapes (ICSInfo.java:90-91) before the
// frame dies further on.  Such a frame is never transformed, so all that survives it is windowShape[CURRENT] = the LAST shape
// read.  The updates are collected here as notes of nine bits -- [1:0] element type, [5:2] instance tag, [6] channel of the
// element, [7] shape, [8] valid; up to three -- and, if the frame does end in an error, (a) written over the header of the
// slot when the object is one this frame decoded before, which the frame itself tells, and (b) handed to the pre-pass in
// FrameSide::notes, which knows the tags of the stream's elements and resolves the others (k2_prepass_kernel).
__device__ __forceinline__ void note_dup_shape(uint32_t& pend, int note_id, int k, const IcsInfoRegs& in) {
  if (note_id < 0 || !in.shape_ok) return;
  int n = 0;
  while (n < 3 && ((pend >> (9 * n + 8)) & 1u)) ++n;
  if (n < 3) pend |= ((uint32_t)note_id | ((uint32_t)(k & 1) << 6) | ((uint32_t)(in.shape & 1) << 7) | 256u) << (9 * n);
}

// Out-of-line readers for the two rare side paths below (dynamic_range_info, ISO pulses): n <= 25 bits at an absolute bit
// position of the frame's word array, straight from global memory -- nothing of the hot reader's register state is touched,
// and the code stays out of the parse loops' instruction stream.
__device__ __noinline__ uint32_t slow_bits(const uint32_t* __restrict__ words, uint32_t last_word, uint32_t pos, int n) {
  const uint32_t wi = pos >> 5;
  const uint32_t a = wi <= last_word ? __byte_perm(__ldg(words + wi), 0, 0x0123) : 0u;
  const uint32_t b = wi + 1 <= last_word ? __byte_perm(__ldg(words + wi + 1), 0, 0x0123) : 0u;
  return __funnelshift_l(b, a, pos & 31u) >> (32 - n);
}

// dynamic_range_info (DRC.decode, syntax/DRC.java:31-83) inside a fill element whose sub-stream is [pos - 4, sub_end): JAAD
// parses it into an object nobody reads (SyntacticElements.java:216-224).  The parse can still end the frame: a read past
// the sub-stream is an EOSException, a second group of excluded-channel flags runs over `new boolean[7]` (DRC.java:27,72-82;
// the flag is read before the store is checked, so an over-read there is still the EOS).  Returns a status.
__device__ __noinline__ int drc_parse(const uint32_t* __restrict__ words, uint32_t last_word, uint32_t pos, uint32_t sub_end) {
  int st = 0, bands = 1;
  uint32_t v = 0;
#define DRC_READ(n) do { if (!st) { if (sub_end - pos < (uint32_t)(n)) st = JAADB_ST_EOS; else { v = slow_bits(words, last_word, pos, n); pos += (n); } } } while (0)
  DRC_READ(1);
  if (!st && v) DRC_READ(8);                                  // pce_tag_present: tag(4) + reserved(4)
  DRC_READ(1);
  if (!st && v) {                                             // excluded_chns_present
    DRC_READ(7);
    DRC_READ(1);
    if (!st && v) { DRC_READ(1); if (!st) st = JAADB_ST_ARRAY_BOUNDS; }
  }
  DRC_READ(1);
  if (!st && v) {                                             // drc_bands_present: increment(4) + interpolation(4)
    DRC_READ(8);
    if (!st) bands += (int)(v >> 4);
    for (int i = 0; i < bands; ++i) DRC_READ(8);
  }
  DRC_READ(1);
  if (!st && v) DRC_READ(8);                                  // prog_ref_level(7) + reserved(1)
  for (int i = 0; i < bands; ++i) DRC_READ(8);                // dyn_rng_sgn(1) + dyn_rng_ctl(7)
#undef DRC_READ
  return st;
}

// JAADB_FLAG_PULSE_ISO: the pulses JAAD parses and never applies ("TODO: apply pulse data", ICStream.java:17), added to the
// quantised coefficients as ISO/IEC 14496-3 4.6.3.3 says: quant[k] += amp if quant[k] > 0, else -= amp.  Long windows only,
// so q is in natural order.  Only coefficients of bands with spectral data take a pulse (codebooks 1..11 below max_sfb); a
// magnitude past IQ_TABLE fails the frame like an escape value of that size.  pos: where pulse_data starts (validated by the
// first parse).  Returns a status.
__device__ __noinline__ int pulse_apply(const uint32_t* __restrict__ words, uint32_t last_word, uint32_t pos,
                                        const int16_t* __restrict__ swb, int max_sfb, const uint8_t* __restrict__ cb_lane,
                                        int cb_stride, int16_t* __restrict__ q) {
  const int count = (int)slow_bits(words, last_word, pos, 2) + 1;
  int sfb = (int)slow_bits(words, last_word, pos + 2, 6);
  pos += 8;
  int off = swb[sfb];
  for (int i = 0; i < count; ++i) {
    const uint32_t v9 = slow_bits(words, last_word, pos, 9);
    pos += 9;
    off += (int)(v9 >> 4);
    const int amp = (int)(v9 & 15u);
    while (sfb < max_sfb && swb[sfb + 1] <= off) ++sfb;
    if (sfb < max_sfb) {
      const int hcb = cb_lane[sfb * cb_stride];
      if (hcb >= 1 && hcb <= 11) {
        int v = q[off];
        v = v > 0 ? v + amp : v - amp;
        if (v > 8190 || v < -8190) return JAADB_ST_ARRAY_BOUNDS;
        q[off] = (int16_t)v;
      }
    }
  }
  return 0;
}

// individual_channel_stream (ICStream.decode, ICStream.java:60-111) for the lanes with go == true.
// `in` holds the shared ics_info when common_window is set.  Returns with status updated.
__device__ __forceinline__ void parse_ics_warp(bool go, BitReader& br, int& status, const uint32_t* __restrict__ lut,
                                               const TablesDev& T, int sf_index, bool common, IcsInfoRegs& in,
                                               IcsSide* side, int16_t* __restrict__ q, int ms_mask,
                                               uint8_t* __restrict__ cb_lane, const int16_t* __restrict__ s_swb, bool discard,
                                               int note_id, int note_k, uint32_t& dup_pend, uint32_t& pns_draws, bool pulse_iso) {
  // discard: the element is not part of the stream's layout (see the element loop): it is parsed for its errors and its
  // length only, nothing is stored
  // codebook per (group, sfb) of this lane's ICS: shared memory, one byte column per thread
#define CB(i) cb_lane[(i) * kK1Threads]
  int global_gain = 0;
  if (go) {
    global_gain = (int)br.read(8);
    if (!common) {
      int st = parse_ics_info(br, in);
      // window_shape bookkeeping happens before predictor data is looked at (ICSInfo.java:90-91)
      if (!discard) store_ics_header(side, in, 0, 1, 0, 0);
      else note_dup_shape(dup_pend, note_id, note_k, in);
      if (st) fail(status, go, st);
    }
  }
  __syncwarp();
  const bool is_short = in.ws == 2;
  const int max_sfb = in.max_sfb;
  const int ngroups = in.ngroups;
  const int nbands = go ? ngroups * max_sfb : 0;

  // ---- section_data (ICStream.java:113-146), one section per iteration
  {
    const int bits = is_short ? 3 : 5;
    const uint32_t esc = (1u << bits) - 1u;
    int g = 0, k = 0;
    bool sec = go && max_sfb > 0;
    while (__any_sync(kFullMask, sec)) {
      if (sec) {
        const uint32_t c = br.read(4);
        if (c == 12) fail(status, sec, JAADB_ST_INVALID_CODEBOOK);
        else {
          int end = k;
          uint32_t incr;
          do {
            incr = br.read(bits);
            end += (int)incr;
          } while (incr == esc && !br.overrun());
          if (br.overrun()) fail(status, sec, JAADB_ST_EOS);
          else if (end > max_sfb) fail(status, sec, JAADB_ST_TOO_MANY_BANDS);
          else {
            for (int idx = g * max_sfb + k; k < end; ++k, ++idx) CB(idx) = (uint8_t)c;
            if (k == max_sfb) { k = 0; if (++g == ngroups) sec = false; }
          }
        }
      }
    }
    if (status) go = false;
  }

  // ---- scale_factor_data (ICStream.java:172-220), one band per iteration
  const uint32_t pns_base = pns_draws;   // what the frame's earlier channels took from the PNS generator
  bool noise_flag = true;                // (false afterwards: the channel has noise bands)
  {
    int off0 = global_gain, off1 = global_gain - 90, off2 = 0;
    const uint32_t sfbase = T.book_base[0];
    uint16_t* sfo = side->sf_idx;
    int idx = 0;
    bool sfa = go && nbands > 0;
    while (__any_sync(kFullMask, sfa)) {
      if (sfa) {
        const int c = CB(idx);
        uint32_t out = 0xFFFFu;
        if (c != 0) {
          if (c == 13 && noise_flag) {
            off1 += (int)br.read(9) - 256;
            noise_flag = false;
            out = (uint32_t)(min(max(off1, -100), 155) + 200) | 0x4000u;
          } else {
            const uint32_t w = br.peek();
            const uint32_t e = huff_lookup(lut, sfbase, kHuffSfFirstBits, w);
            br.skip(e & 31u);
            const int delta = (int)(e >> 16) - 60;
            if (c >= 14) {
              off2 += delta;
              out = (uint32_t)(200 - min(max(off2, -155), 100));
            } else if (c == 13) {
              off1 += delta;
              out = (uint32_t)(min(max(off1, -100), 155) + 200) | 0x4000u;
            } else {
              off0 += delta;
              if (off0 > 255) fail(status, sfa, JAADB_ST_SF_RANGE);
              else if (off0 + 100 < 0) fail(status, sfa, JAADB_ST_ARRAY_BOUNDS);
              out = (uint32_t)(off0 + 100);
            }
          }
        }
        if (sfa) {
          if (!discard) sfo[idx] = (uint16_t)out;
          if (++idx == nbands) sfa = false;
        }
      }
    }
    if (status) go = false;
  }

  // ---- pulse_data (ICStream.java:76-83,148-170) and tns_data (TNS.java:35-61): parsed, never applied by JAAD;
  //      gain_control_data: SSR only.  Rare and short: the lanes diverge here and rejoin at the __syncwarp.
  // JAADB_FLAG_PULSE_ISO: where pulse_data starts (it is read again after the spectral data), kept in the four spare rows of
  // the lane's codebook column instead of a register that would live across the spectral loop
  if (pulse_iso) { CB(124) = 0; CB(125) = 0; CB(126) = 0; CB(127) = 0; }
  if (go) {
    if (br.read1()) {
      if (is_short) fail(status, go, JAADB_ST_PULSE_SHORT);
      else {
        if (pulse_iso) {
          CB(124) = (uint8_t)br.pos; CB(125) = (uint8_t)(br.pos >> 8); CB(126) = (uint8_t)(br.pos >> 16); CB(127) = (uint8_t)(br.pos >> 24);
        }
        const int count = (int)br.read(2) + 1;
        const int start = (int)br.read(6);
        const int swb_count = T.swb_long_count[sf_index];
        if (start >= swb_count) fail(status, go, JAADB_ST_PULSE_RANGE);
        else {
          int off = T.swb_long[sf_index * 53 + start];
          off += (int)br.read(5);
          br.skip(4);
          for (int i = 1; i < count && go; ++i) {
            off += (int)br.read(5);
            if (off > 1023) fail(status, go, JAADB_ST_PULSE_RANGE);
            else br.skip(4);
          }
        }
      }
    }
  }
  if (go) {
    const uint32_t tns_present = br.read1();
    const uint32_t tns_off = br.pos;
    if (tns_present) {
      const int nwin = is_short ? 8 : 1;
      const int b0 = is_short ? 1 : 2, b1 = is_short ? 4 : 6, b2 = is_short ? 3 : 5;
      for (int w = 0; w < nwin && go; ++w) {
        const int nfilt = (int)br.read(b0);
        if (nfilt) {
          const int coef_res = (int)br.read1();
          for (int f = 0; f < nfilt && go; ++f) {
            br.skip(b1);
            const int order = (int)br.read(b2);
            if (order > 20) fail(status, go, JAADB_ST_TNS_ORDER);
            else if (order) {
              br.skip(1);
              const int compress = (int)br.read1();
              br.skip((uint32_t)(order * (coef_res + 3 - compress)));
            }
          }
        }
        if (go && br.overrun()) fail(status, go, JAADB_ST_EOS);
      }
    }
    if (go) {
      // tns_present | has_pns | pns_base, tns_bit_off (IcsSide): K2 reads tns_data again in JAADB_TNS_ISO mode
      if (!discard)
        *reinterpret_cast<uint2*>(&side->tns_present) = make_uint2(tns_present | (noise_flag ? 0u : 0x100u) | (pns_base << 16), tns_off);
      if (br.read1()) fail(status, go, JAADB_ST_UNSUPPORTED_ELEMENT);  // gain control: outside the engine's scope
    }
  }
  // section table out (K2 needs it for dequantisation / stereo tools)
  if (go && !discard) {
    uint32_t* dst = reinterpret_cast<uint32_t*>(side->sfb_cb);
    for (int i = 0; i < (nbands + 3) / 4; ++i)
      dst[i] = (uint32_t)CB(4 * i) | ((uint32_t)CB(4 * i + 1) << 8) | ((uint32_t)CB(4 * i + 2) << 16) | ((uint32_t)CB(4 * i + 3) << 24);
  }
  __syncwarp();

  // ---- spectral_data (ICStream.java:222-275, Huffman.java:56-84): one codeword per lane per iteration
  {
    const int16_t* __restrict__ swb = is_short ? (s_swb + 12 * 53 + sf_index * 17) : (s_swb + sf_index * 53);
    const int swb_count = is_short ? T.swb_short_count[sf_index] : T.swb_long_count[sf_index];
    int idx = 0, sfb = 0, g = 0, gbase = 0, glen = in.glen(0);
    int rem = 0, pos = 0, hcb = 0;
    uint32_t base = 0;
    bool sp = go && nbands > 0;
    while (__any_sync(kFullMask, sp)) {
      if (sp && rem == 0) {
        // open the next scalefactor band
        if (idx == nbands) sp = false;
        else {
          hcb = CB(idx);
          if (sfb > swb_count) fail(status, sp, JAADB_ST_ARRAY_BOUNDS);          // offsets[sfb+1] past the table
          else if (hcb == 0 || hcb >= 14) {
            if (sfb == swb_count) fail(status, sp, JAADB_ST_ARRAY_BOUNDS);       // Arrays.fill with a negative range
          } else if (hcb == 13) {
            // PNS (ICStream.java:241-257): the band takes glen * width values from the generator while it is parsed; K2
            // regenerates them from the draw offset (a band at the table's end has a negative width: nothing happens)
            if (sfb < swb_count) pns_draws += (uint32_t)(glen * (swb[sfb + 1] - swb[sfb]));
          } else if (sfb < swb_count) {                                            // (== swb_count: negative width, body never runs)
            const int lo = swb[sfb], width = swb[sfb + 1] - lo;
            pos = gbase + glen * lo;
            rem = (glen * width) >> (hcb < 5 ? 2 : 1);
            base = T.book_base[hcb];
          }
          ++idx;
          if (++sfb == max_sfb) {
            sfb = 0;
            gbase += glen << 7;
            ++g;
            glen = in.glen(g & 7);
          }
        }
      }
      if (sp && rem > 0) {
        const uint32_t bits = br.peek();
        const uint32_t e = huff_lookup(lut, base, kHuffFirstBits, bits);
        const uint32_t len = e & 31u, ns = (e >> 5) & 7u, pay = e >> 16;
        const bool quad = hcb < 5;
        int v0, v1, v2, v3;
        if (quad) {
          v0 = ((int)(pay << 28)) >> 28; v1 = ((int)(pay << 24)) >> 28;
          v2 = ((int)(pay << 20)) >> 28; v3 = ((int)(pay << 16)) >> 28;
        } else {
          v0 = ((int)(pay << 24)) >> 24; v1 = ((int)(pay << 16)) >> 24;
          v2 = 0; v3 = 0;
        }
        if (ns) {
          const uint32_t sb = (bits << len) >> (32 - ns);
          int i = (int)ns;
          if (v0) { --i; if ((sb >> i) & 1u) v0 = -v0; }
          if (v1) { --i; if ((sb >> i) & 1u) v1 = -v1; }
          if (v2) { --i; if ((sb >> i) & 1u) v2 = -v2; }
          if (v3) { --i; if ((sb >> i) & 1u) v3 = -v3; }
        }
        br.skip(len + ns);
        if (hcb == 11) {
          // getEscape (Huffman.java:39-49): N ones, a zero, then 4+N bits; value = bits | 1<<(4+N)
          if (v0 == 16 || v0 == -16) {
            const uint32_t eb = br.peek();
            const int n1 = __clz((int)~eb);
            if (n1 > 8) { br.skip(n1 + 1 + 4 + n1); fail(status, sp, JAADB_ST_ARRAY_BOUNDS); }
            else {
              const int i = 4 + n1;
              const int mag = (int)((eb << (n1 + 1)) >> (32 - i)) | (1 << i);
              br.skip(n1 + 1 + i);
              v0 = v0 < 0 ? -mag : mag;
            }
          }
          if (sp && (v1 == 16 || v1 == -16)) {
            const uint32_t eb = br.peek();
            const int n1 = __clz((int)~eb);
            if (n1 > 8) { br.skip(n1 + 1 + 4 + n1); fail(status, sp, JAADB_ST_ARRAY_BOUNDS); }
            else {
              const int i = 4 + n1;
              const int mag = (int)((eb << (n1 + 1)) >> (32 - i)) | (1 << i);
              br.skip(n1 + 1 + i);
              v1 = v1 < 0 ? -mag : mag;
            }
          }
          if (sp && (v0 > 8190 || v0 < -8190 || v1 > 8190 || v1 < -8190)) fail(status, sp, JAADB_ST_ARRAY_BOUNDS);  // IQ_TABLE has 8191 entries
        }
        if (sp) {
          const uint32_t lo32 = ((uint32_t)v0 & 0xFFFFu) | ((uint32_t)v1 << 16);
          if (quad) {
            if (!discard) *reinterpret_cast<uint2*>(q + pos) = make_uint2(lo32, ((uint32_t)v2 & 0xFFFFu) | ((uint32_t)v3 << 16));
            pos += 4;
          } else {
            if (!discard) *reinterpret_cast<uint32_t*>(q + pos) = lo32;
            pos += 2;
          }
          --rem;
          // JAAD tests for the end of the frame on every read; here once per band is enough (the reader
          // returns zeros past the end and any status found there is replaced by EOS below)
          if (rem == 0 && br.overrun()) fail(status, sp, JAADB_ST_EOS);
        }
      }
    }
    if (status) go = false;
  }
  // ---- JAADB_FLAG_PULSE_ISO (pulse_apply above).  Rare and short: the lanes diverge and rejoin below.  (Elements outside
  //      the stream's layout store no coefficients: nothing to do.)
  if (pulse_iso && go && !discard) {
    const uint32_t pulse_pos = (uint32_t)CB(124) | ((uint32_t)CB(125) << 8) | ((uint32_t)CB(126) << 16) | ((uint32_t)CB(127) << 24);
    if (pulse_pos != 0u) {
      const int st = pulse_apply(br.words, br.last_word, pulse_pos, s_swb + sf_index * 53, max_sfb, cb_lane, kK1Threads, q);
      if (st) fail(status, go, st);
    }
  }
  if (go && !discard) store_ics_header(side, in, 1, 1, ms_mask, common ? 1 : 0);
  __syncwarp();
#undef CB
}

__global__ void __launch_bounds__(kK1Threads, 4)   // 64 registers: four CTAs per SM (the shared-memory limit)
k1_parse_kernel(const uint8_t* __restrict__ blob, const FrameDev* __restrict__ frames, uint32_t n_frames,
                FrameSide* __restrict__ fside, IcsSide* __restrict__ iside_all, int16_t* __restrict__ q_all, TablesDev T,
                const LayoutDev* __restrict__ layouts, int pulse_iso, uint32_t lanes_log2) {
  extern __shared__ uint32_t s_lut[];
  // shared memory: Huffman LUTs | SWB offset tables (long [12][53], short [12][17]) | per-lane codebook columns
  int16_t* s_swb = reinterpret_cast<int16_t*>(s_lut + T.huff_lut_entries);
  uint8_t* s_cb = reinterpret_cast<uint8_t*>(s_swb + kSwbTableEntries);
  for (uint32_t i = threadIdx.x; i < T.huff_lut_entries; i += blockDim.x) s_lut[i] = T.huff_lut[i];
  for (int i = threadIdx.x; i < 12 * 53; i += blockDim.x) s_swb[i] = T.swb_long[i];
  for (int i = threadIdx.x; i < 12 * 17; i += blockDim.x) s_swb[12 * 53 + i] = T.swb_short[i];
  __syncthreads();
  // Frames per warp: 32 for a batch that fills the GPU; fewer -- down to ONE -- for a small one (a tick of a live batch:
  // a few thousand frames).  A warp's 32 frames advance in lock-step, so its time is the longest frame's with every branch
  // any lane takes; a small batch cannot use the lanes for throughput anyway and gets the latency of a frame parsed alone
  // (k1_lanes_log2: as many warps as the GPU holds in one wave before the warps start to fill up).
  const uint32_t gt = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t f = ((gt >> 5) << lanes_log2) + (threadIdx.x & 31u);
  const bool valid = (threadIdx.x & 31u) < (1u << lanes_log2) && f < n_frames;
  FrameDev fr;
  fr.blob_off = 0; fr.nbytes = 0; fr.stream_slot = 0; fr.ics_base = 0; fr.sf_index = 0; fr.layout = 0; fr.profile_ok = 1; fr.flags = 0;
  if (valid) fr = frames[f];
  const LayoutDev lay = layouts[fr.layout];
  IcsSide* const iside = iside_all + fr.ics_base;
  int16_t* const qbase = q_all + (size_t)fr.ics_base * 1024;
  const int sf_index = fr.sf_index;

  FrameSide fs;
  fs.n_started = 0;
  fs.tags = 0;
  fs.n_elements = 0;
  fs.sbr_bit_off[0] = fs.sbr_bit_off[1] = 0;
  fs.sbr_bits[0] = fs.sbr_bits[1] = 0;
  fs.notes = 0;
  uint32_t pns_draws = 0;
  int status = JAADB_ST_OK;
  bool active = valid;
  BitReader br;
  br.init(blob, fr.blob_off, valid ? fr.nbytes : 0u);
  const uint32_t start = br.pos;
  if (valid) {
    // every channel slot starts out "absent"
    for (int c = 0; c < lay.n_channels; ++c) *reinterpret_cast<uint4*>(iside + c) = make_uint4(0, 0, 0, 0);
    if (fr.nbytes < 4) fail(status, active, JAADB_ST_EOS);  // ADIFHeader.isPresent peeks 32 bits (transport/ADIFHeader.java:18)
    else if (br.peek() == 0x41444946u) fail(status, active, JAADB_ST_UNSUPPORTED_ELEMENT);  // 'ADIF'
    else if (!fr.profile_ok) fail(status, active, JAADB_ST_PROFILE);
  }

  int el = 0, n_good = 0;
  bool layout_bad = false;
  int n_in_layout = 0;            // elements parsed while the frame still followed the stream's layout
  int note_id = -1;               // the current element is outside the layout: its type | tag << 2 (note_dup_shape)
  uint32_t dup_pend = 0;          // note_dup_shape
  bool pend_r = false;            // the right channel of the current CPE is next
  int ch0 = 0, ms_mask = 0;
  bool common = false;
  IcsInfoRegs in, in_r;
  in.ws = 0; in.shape = 0; in.shape_ok = 1; in.max_sfb = 0; in.ngroups = 1; in.glen_packed = 1;
  in_r = in;

  // one syntactic element (or one channel of a CPE) per iteration
  while (__any_sync(kFullMask, active)) {
    bool go = false;   // this lane parses an individual_channel_stream in this iteration
    int ch = 0;
    bool is_cpe_left = false;
    if (active) {
      if (pend_r) {
        pend_r = false;
        go = true;
        ch = min(ch0 + 1, lay.n_channels - 1);
        in = in_r;
      } else if (br.overrun()) {
        fail(status, active, JAADB_ST_EOS);
      } else {
        const int type = (int)br.read(3);
        if (type == EL_END) {
          active = false;
        } else if (type == EL_SCE || type == EL_LFE || type == EL_CPE) {
          const uint32_t tag = br.read(4);
          // An element the stream's channel layout does not have: JAAD decodes it all the same (and usually dies of
          // something else further on), so the parse goes on -- into channel slots that exist -- and the frame ends as
          // JAADB_ST_LAYOUT only if nothing else stops it first.
          if (el >= lay.n_elements || lay.el_type[el] != type) layout_bad = true;
          if (!layout_bad) n_in_layout = el + 1;
          note_id = layout_bad ? (int)((uint32_t)type | (tag << 2)) : -1;
          {
            ch0 = layout_bad ? 0 : lay.el_first_ch[el];
            ch = ch0;
            if (el < 4) { fs.tags |= (uint16_t)(tag << (4 * el)); fs.n_started = (uint8_t)(el + 1); }
            in.ws = 0; in.shape = 0; in.shape_ok = 1; in.max_sfb = 0; in.ngroups = 1; in.glen_packed = 1;
            common = false;
            ms_mask = 0;
            go = true;
            if (type == EL_CPE) {
              // CPE.decode (CPE.java:85-123)
              is_cpe_left = true;
              common = br.read1() != 0;
              if (common) {
                const int st = parse_ics_info(br, in);
                if (!layout_bad) store_ics_header(iside + ch0, in, 0, 1, 0, 1);
                else note_dup_shape(dup_pend, note_id, 0, in);
                if (st) { fail(status, active, st); go = false; }  // thrown inside infoL.decode: R's setCommonData never ran (CPE.java:95-96)
                // the frame ended inside infoL.decode: JAAD's EOSException comes before setCommonData, so R keeps its window
                // shape (the reader here would run on over zeros, and the channels' final headers would update R after all)
                else if (br.overrun()) { fail(status, active, JAADB_ST_EOS); go = false; }
                else {
                  // setCommonData updates R's window shape too
                  if (!layout_bad) store_ics_header(iside + ch0 + 1, in, 0, 1, 0, 1);
                  else note_dup_shape(dup_pend, note_id, 1, in);
                  ms_mask = (int)br.read(2);
                  uint32_t msv[4] = {0u, 0u, 0u, 0u};
                  if (ms_mask == 1) {
                    const int n = in.ngroups * in.max_sfb;
                    for (int i = 0; i < 4; ++i) {
                      const int take = min(32, n - 32 * i);
                      if (take > 0) msv[i] = __brev(br.read(take) << (32 - take));
                    }
                  } else if (ms_mask == 2) {
                    msv[0] = msv[1] = msv[2] = msv[3] = 0xFFFFFFFFu;
                  } else if (ms_mask != 0) { fail(status, active, JAADB_ST_MS_RESERVED); go = false; }
                  if (go && !layout_bad) {
                    uint32_t* ms = reinterpret_cast<uint32_t*>((iside + ch0)->ms_used);
                    ms[0] = msv[0]; ms[1] = msv[1]; ms[2] = msv[2]; ms[3] = msv[3];
                  }
                }
              }
              in_r = in;
            }
          }
        } else if (type == EL_DSE) {
          // DSE.decode (syntax/DSE.java:54-66)
          br.skip(4);
          const bool align = br.read1() != 0;
          uint32_t count = br.read(8);
          if (count == 255) count += br.read(8);
          if (align) br.pos = start + (((br.pos - start) + 7u) & ~7u);
          br.skip(8 * count);
        } else if (type == EL_FIL) {
          // decodeFIL (SyntacticElements.java:169-203)
          int count = (int)br.read(4);
          if (count == 15) count += (int)br.read(8) - 1;
          if (count > 0) {
            if (br.bits_left() < (uint32_t)(8 * count) || br.overrun()) { br.skip(8 * count); fail(status, active, JAADB_ST_EOS); }
            else {
              const uint32_t ext = br.peek() >> 28;
              if (ext == 11) {
                // dynamic range info: parsed and dropped, like JAAD does (drc_parse above)
                const uint32_t sub_end = br.pos + 8u * (uint32_t)count;
                const int st = drc_parse(br.words, br.last_word, br.pos + 4u, sub_end);
                br.pos = sub_end;
                if (st) fail(status, active, st);
              } else {
                if ((ext == 13 || ext == 14) && el > 0 && el <= 2 && !layout_bad) {
                  fs.sbr_bit_off[el - 1] = br.pos;   // relative to the aligned word base of the frame
                  fs.sbr_bits[el - 1] = 8u * (uint32_t)count;
                }
                br.skip(8 * count);
              }
            }
          }
        } else {
          fail(status, active, JAADB_ST_UNSUPPORTED_ELEMENT);  // CCE / PCE
        }
      }
    }
    __syncwarp();
    parse_ics_warp(go, br, status, s_lut, T, sf_index, common, in, iside + ch, qbase + ch * 1024, ms_mask,
                   s_cb + threadIdx.x, s_swb, layout_bad, layout_bad ? note_id : -1, ch - ch0, dup_pend,
                   pns_draws, pulse_iso != 0);
    if (go) {
      if (status) active = false;
      else if (is_cpe_left) pend_r = true;
      else { ++el; if (!layout_bad) n_good = el; }
    }
  }
  fs.n_elements = (uint8_t)n_good;   // elements of the stream's layout that were decoded completely
  if (valid) {
    if (br.overrun()) status = JAADB_ST_EOS;
    if (status == JAADB_ST_OK && (layout_bad || el != lay.n_elements)) status = JAADB_ST_LAYOUT;
    if (dup_pend && status != JAADB_ST_OK && status != JAADB_ST_LAYOUT) {
      for (int n = 0; n < 3 && ((dup_pend >> (9 * n + 8)) & 1u); ++n) {
        const uint32_t nt = dup_pend >> (9 * n);
        for (int j = 0; j < n_in_layout && j < 4; ++j)
          if (lay.el_type[j] == (nt & 3u) && ((uint32_t)(fs.tags >> (4 * j)) & 15u) == ((nt >> 2) & 15u)) {
            IcsSide* const s = iside + lay.el_first_ch[j] + ((nt >> 6) & 1u);
            s->window_shape = (uint8_t)((nt >> 7) & 1u);
            s->info_decoded = 1;
          }
      }
      fs.notes = dup_pend;
    }
    fs.status = status;
    fs.pns_draws = pns_draws;
    fside[f] = fs;
  }
}

}  // namespace jaadb
