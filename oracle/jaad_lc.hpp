// ORACLE -- TEST INFRASTRUCTURE ONLY (see jaad_bits.hpp for the full notice).
//
// CPU restatement of JAAD's AAC-LC decode path: AudioSpecificConfig, the
// raw_data_block element loop, ics_info / section / scalefactor / spectral
// Huffman with fused dequantisation, M/S, intensity stereo, TNS (parse only,
// as in JAAD), the IMDCT filterbank and the SampleBuffer int16 packing.
// Same operation order, same float/double promotion points, same quirks.
// Every function names the reference file:line it follows (paths relative to
// aac/src/main/java/net/sourceforge/jaad/aac/ unless they start with S/ =
// src/main/java/net/sourceforge/jaad/).
#pragma once
#include <algorithm>
#include <cmath>
#include <functional>
#include <map>
#include <memory>

#include "../jaadec_b200/csrc/generated/jaad_tables_host.h"
#include "jaad_bits.hpp"

namespace jaad {

namespace T = ::jaad_tables;

// ---------------------------------------------------------------------------
// enums / config  (Profile.java, SampleFrequency.java, ChannelConfiguration.java)
// ---------------------------------------------------------------------------
static const int SF_FREQ[12] = {96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000};

// Profile.forInt (Profile.java:51-57) + isDecodingSupported / isErrorResilientProfile
struct Profile {
  int num = -1;  // -1 == UNKNOWN
  static Profile forInt(int i) {
    static const int ALL[30] = {1, 2, 3, 4, 5, 6, 7, -1, -1, -1, 11, -1, -1, -1, -1, -1, 17, 18,
                                19, 20, 21, 22, 23, -1, -1, -1, -1, -1, 29, -1};
    Profile p;
    p.num = (i >= 1 && i <= 30) ? ALL[i - 1] : -1;
    return p;
  }
  bool supported() const {
    switch (num) { case 1: case 2: case 4: case 5: case 17: case 19: case 29: return true; default: return false; }
  }
  bool errorResilient() const { return num > 16; }
};

// SampleRate: a nominal table index plus the actual frequency (SampleRate.java,
// SampleFrequency.java:42-66).  index<0 == SF_NONE.
struct SampleRate {
  int index = -1;
  int frequency = 0;
  bool none() const { return index < 0; }
  static SampleRate forIndex(int i) {
    SampleRate r;
    if (i >= 0 && i < 12) { r.index = i; r.frequency = SF_FREQ[i]; }
    return r;
  }
  // SampleFrequency.nominalFrequency (SampleFrequency.java:68-93)
  static SampleRate forFrequency(int freq) {
    int result = -1;
    float dev = INFINITY;
    for (int i = 0; i < 12; ++i) {
      float d = ((float)freq - (float)SF_FREQ[i]) / (float)SF_FREQ[i];
      if (d == 0) { result = i; break; }
      if (d < dev) { result = i; dev = d; }
      if (SF_FREQ[i] < freq) break;
    }
    SampleRate r;
    r.index = result;
    r.frequency = freq;
    return r;
  }
  // SampleFrequency.duplicated (SampleFrequency.java:136-138) / anonymous SampleRate :59-62
  SampleRate duplicated() const {
    if (index < 3) return SampleRate();
    SampleRate r;
    r.index = index - 3;
    r.frequency = (frequency == SF_FREQ[index]) ? SF_FREQ[index - 3] : 2 * frequency;
    return r;
  }
  bool same(const SampleRate& o) const { return index == o.index && frequency == o.frequency; }
  // SampleRate.decode (SampleRate.java:28-38)
  static SampleRate decode(BitStream& in) {
    int index = in.readBits(4);
    if (index != 0x0f) {
      if (index >= 12) throw AACException(ST_ARRAY_BOUNDS, "sample frequency index out of range");
      return forIndex(index);
    }
    return forFrequency(in.readBits(24));
  }
};

// ChannelConfiguration.forInt (ChannelConfiguration.java:26-33): returns channel count (ordinal)
inline int channelConfigForInt(int i) {
  if (i >= 7) ++i;
  if (i > 8) throw AACException(ST_ARRAY_BOUNDS, "channel configuration out of range");
  return i;  // ordinal == channel count; 7 is INVALID_SEVEN
}

// DecoderConfig.java
struct DecoderConfig {
  Profile profile = Profile::forInt(1), extProfile;
  SampleRate sampleFrequency;
  int channelConfiguration = -1;  // ordinal; -1 unsupported
  bool frameLengthFlag = false, dependsOnCoreCoder = false, extensionFlag = false;
  int coreCoderDelay = 0;
  bool sbrEnabled = true;
  bool sbrPresent = false;
  bool hasOutputFrequency = false;
  SampleRate outputFrequency;
  bool psEnabled = true, psPresent = false;
  bool sectionDataResilience = false, scalefactorResilience = false, spectralDataResilience = false;
  // Not JAAD fields.  (1) JAAD's PNS generator is ONE static int for the whole JVM (ICStream.java:26), so its output depends
  // on every decoder that ran before; the oracle (and the engine) give each Decoder its own generator with JAAD's seed,
  // which is exactly "this stream decoded alone in a fresh JVM".  (2) tnsMode 0 = JAAD (TNS.process is a stub,
  // tools/TNS.java:63-68), 1 = the ISO/IEC 14496-3 4.6.9 all-pole filter (TNS::process below).
  // (3) pulseMode 0 = JAAD (pulse_data parsed, never applied: "TODO: apply pulse data", ICStream.java:17), 1 = the pulses of
  // ISO/IEC 14496-3 4.6.3.3 are added to the quantised coefficients before the inverse quantisation.
  mutable int32_t pnsState = 0x1F2E3D4C;
  int tnsMode = 0;
  int pulseMode = 0;

  int getFrameLength() const { return frameLengthFlag ? 960 : 1024; }
  bool isUpSampled() const { return hasOutputFrequency && !outputFrequency.same(sampleFrequency); }
  int getSampleLength() const { return (isUpSampled() ? 2 : 1) * getFrameLength(); }  // :83-86
  SampleRate getOutputFrequency() const { return hasOutputFrequency ? outputFrequency : sampleFrequency; }
  int getChannelCount() const {  // :108-115
    if (sbrEnabled && channelConfiguration == 1) return 2;
    return channelConfiguration;
  }
  bool setSBRPresent() {  // :124-135
    sbrPresent = true;
    if (!hasOutputFrequency) {
      SampleRate d = sampleFrequency.duplicated();
      if (d.none()) return false;
      outputFrequency = d;
      hasOutputFrequency = true;
    }
    return isUpSampled();
  }

  // DecoderConfig.create(AudioDecoderInfo) (:164-166, :59-64)
  static DecoderConfig fromInfo(int profileNum, int sfIndex, int chanCfg) {
    DecoderConfig c;
    c.profile = Profile::forInt(profileNum);
    c.sampleFrequency = SampleRate::forIndex(sfIndex);
    c.channelConfiguration = channelConfigForInt(chanCfg);
    return c;
  }

  static Profile readProfile(BitStream& in) {  // :256-261
    int i = in.readBits(5);
    if (i == 31) i = 32 + in.readBits(6);
    return Profile::forInt(i);
  }

  // DecoderConfig.decode (:175-254)
  void decode(BitStream& in) {
    profile = readProfile(in);
    sampleFrequency = SampleRate::decode(in);
    outputFrequency = sampleFrequency;
    hasOutputFrequency = true;
    channelConfiguration = channelConfigForInt(in.readBits(4));
    switch (profile.num) {
      case 29:
        psPresent = true;  // falls through
      case 5: {
        SampleRate frequency = SampleRate::decode(in);
        extProfile = profile;
        profile = readProfile(in);
        if (sbrEnabled) outputFrequency = frequency;
        break;
      }
      case 1: case 2: case 3: case 4: case 17: case 19: case 23:
        frameLengthFlag = in.readBool();
        if (frameLengthFlag) throw AACException(ST_CONFIG, "config uses 960-sample frames, not yet supported");
        dependsOnCoreCoder = in.readBool();
        coreCoderDelay = dependsOnCoreCoder ? in.readBits(14) : 0;
        extensionFlag = in.readBool();
        if (extensionFlag) {
          if (profile.errorResilient()) {
            sectionDataResilience = in.readBool();
            scalefactorResilience = in.readBool();
            spectralDataResilience = in.readBool();
          }
          in.skipBit();
        }
        if (channelConfiguration == 0)
          throw AACException(ST_UNSUPPORTED_ELEMENT, "PCE in AudioSpecificConfig is outside the engine's scope");
        if (sbrEnabled && in.getBitsLeft() > 10) readSyncExtension(in);
        break;
      default:
        throw AACException(ST_CONFIG, "profile not supported");
    }
  }

  void readSyncExtension(BitStream& in) {  // :268-291
    int extensionType = in.readBits(11);
    if (extensionType == 0x2B7) {
      extProfile = Profile::forInt(in.readBits(5));
      if (extProfile.num == 5 || extProfile.num == 22) {
        sbrPresent = in.readBool();
        if (sbrPresent) { outputFrequency = SampleRate::decode(in); hasOutputFrequency = true; }
        if (extProfile.num == 5) {
          if (in.getBitsLeft() > 12) {
            extensionType = in.readBits(11);
            if (extensionType == 0x548) psPresent = in.readBool();
          }
        } else {
          channelConfigForInt(in.readBits(4));
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------
// Huffman  (huffman/Huffman.java)
// ---------------------------------------------------------------------------
struct HuffBook { const int32_t* rows; int nrows; int width; };
inline HuffBook spectralBook(int cb) {  // Codebooks.CODEBOOKS[cb-1]
  switch (cb) {
    case 1: return {T::HCB1, 81, 6};   case 2: return {T::HCB2, 81, 6};
    case 3: return {T::HCB3, 81, 6};   case 4: return {T::HCB4, 81, 6};
    case 5: return {T::HCB5, 81, 4};   case 6: return {T::HCB6, 81, 4};
    case 7: return {T::HCB7, 64, 4};   case 8: return {T::HCB8, 64, 4};
    case 9: return {T::HCB9, 169, 4};  case 10: return {T::HCB10, 169, 4};
    default: return {T::HCB11, 289, 4};
  }
}

// Huffman.findOffset (Huffman.java:15-28): linear scan of {len, code, ...} rows.
inline int findOffset(BitStream& in, const HuffBook& b) {
  int off = 0;
  int len = b.rows[0];
  int cw = in.readBits(len);
  while (cw != b.rows[off * b.width + 1]) {
    off++;
    if (off >= b.nrows) throw AACException(ST_ARRAY_BOUNDS, "huffman codeword not in table");
    int j = b.rows[off * b.width] - len;
    len = b.rows[off * b.width];
    cw = (int)((uint32_t)cw << (j & 31));
    cw |= in.readBits(j);
  }
  return off;
}

inline int decodeScaleFactor(BitStream& in) {  // Huffman.java:51-54
  HuffBook b{T::HCB_SF, 121, 3};
  return b.rows[findOffset(in, b) * 3 + 2];
}

inline void signValues(BitStream& in, int* data, int off, int len) {  // Huffman.java:30-37
  for (int i = off; i < off + len; i++)
    if (data[i] != 0)
      if (in.readBool()) data[i] = -data[i];
}

inline int getEscape(BitStream& in, int s) {  // Huffman.java:39-49
  bool neg = s < 0;
  int i = 4;
  while (in.readBool()) i++;
  if (i > 32) throw AACException(ST_ARRAY_BOUNDS, "escape prefix too long");
  int j = in.readBits(i) | (int)(1u << (i & 31));
  return neg ? -j : j;
}

inline void decodeSpectralData(BitStream& in, int cb, int* data, int off) {  // Huffman.java:56-84
  static const bool UNSIGNED[11] = {false, false, true, true, false, false, true, true, true, true, true};
  HuffBook b = spectralBook(cb);
  int offset = findOffset(in, b);
  const int32_t* row = b.rows + offset * b.width;
  data[off] = row[2];
  data[off + 1] = row[3];
  if (cb < 5) { data[off + 2] = row[4]; data[off + 3] = row[5]; }
  if (cb < 11) {
    if (UNSIGNED[cb - 1]) signValues(in, data, off, cb < 5 ? 4 : 2);
  } else if (cb == 11 || cb > 15) {
    signValues(in, data, off, cb < 5 ? 4 : 2);
    if (std::abs(data[off]) == 16) data[off] = getEscape(in, data[off]);
    if (std::abs(data[off + 1]) == 16) data[off + 1] = getEscape(in, data[off + 1]);
  } else {
    throw AACException(ST_INVALID_CODEBOOK, "Huffman: unknown spectral codebook");
  }
}

// ---------------------------------------------------------------------------
// ICSInfo (syntax/ICSInfo.java)
// ---------------------------------------------------------------------------
enum WindowSequence { ONLY_LONG_SEQUENCE = 0, LONG_START_SEQUENCE = 1, EIGHT_SHORT_SEQUENCE = 2, LONG_STOP_SEQUENCE = 3 };

struct ICSInfo {
  static const int PREVIOUS = 0, CURRENT = 1;
  const DecoderConfig* config;
  int sfIndex;
  WindowSequence windowSequence = ONLY_LONG_SEQUENCE;
  int windowShape[2] = {0, 0};
  int maxSFB = 0;
  bool predictionDataPresent = false;
  int windowCount = 0, windowGroupCount = 0;
  int windowGroupLength[8] = {0};
  int swbCount = 0;
  const int16_t* swbOffsets = nullptr;

  explicit ICSInfo(const DecoderConfig& c) : config(&c), sfIndex(c.sampleFrequency.index) {}

  bool isEightShortFrame() const { return windowSequence == EIGHT_SHORT_SEQUENCE; }

  void decode(BitStream& in, bool /*commonWindow*/) {  // ICSInfo.java:86-119
    in.skipBit();
    windowSequence = (WindowSequence)in.readBits(2);
    windowShape[PREVIOUS] = windowShape[CURRENT];
    windowShape[CURRENT] = in.readBit();
    windowGroupCount = 1;
    windowGroupLength[0] = 1;
    if (windowSequence == EIGHT_SHORT_SEQUENCE) {
      maxSFB = in.readBits(4);
      for (int i = 0; i < 7; i++) {
        if (in.readBool()) windowGroupLength[windowGroupCount - 1]++;
        else { windowGroupCount++; windowGroupLength[windowGroupCount - 1] = 1; }
      }
      windowCount = 8;
      swbOffsets = T::SWB_OFFSET_SHORT + 17 * sfIndex;
      swbCount = T::SWB_SHORT_WINDOW_COUNT[sfIndex];
      predictionDataPresent = false;
    } else {
      maxSFB = in.readBits(6);
      windowCount = 1;
      swbOffsets = T::SWB_OFFSET_LONG + 53 * sfIndex;
      swbCount = T::SWB_LONG_WINDOW_COUNT[sfIndex];
      predictionDataPresent = in.readBool();
      if (predictionDataPresent) {
        // readPredictionData (:121-141): Main/LTP profiles are outside the engine's scope;
        // every other profile throws in JAAD.
        if (config->profile.num == 1 || config->profile.num == 4 || config->profile.num == 19)
          throw AACException(ST_UNSUPPORTED_ELEMENT, "Main/LTP prediction is outside the engine's scope");
        throw AACException(ST_LTP_PROFILE, "unexpected profile for LTP");
      }
    }
  }

  void setCommonData(const ICSInfo& info) {  // ICSInfo.java:193-211
    windowSequence = info.windowSequence;
    windowShape[PREVIOUS] = windowShape[CURRENT];
    windowShape[CURRENT] = info.windowShape[CURRENT];
    maxSFB = info.maxSFB;
    predictionDataPresent = info.predictionDataPresent;
    windowCount = info.windowCount;
    windowGroupCount = info.windowGroupCount;
    std::copy(info.windowGroupLength, info.windowGroupLength + 8, windowGroupLength);
    swbCount = info.swbCount;
    swbOffsets = info.swbOffsets;
  }
};

// ---------------------------------------------------------------------------
// FFT / MDCT / FilterBank (filterbank/FFT.java, MDCT.java, FilterBank.java)
// ---------------------------------------------------------------------------
struct FFT {
  int length;
  const float* roots;  // [length][cols]
  int cols;
  std::vector<float> rev;
  explicit FFT(int len) : length(len) {
    if (len == 512) { roots = JT(FFT_TABLE_512); cols = 3; }
    else if (len == 64) { roots = JT(FFT_TABLE_64); cols = 2; }
    else throw AACException(ST_CONFIG, "unexpected FFT length");
    rev.resize(2 * len);
  }
  // FFT.process (FFT.java:48-135), in = [length][2] interleaved
  void process(float* in, bool forward) {
    int ii = 0;
    for (int i = 0; i < length; i++) {
      rev[2 * i] = in[2 * ii];
      rev[2 * i + 1] = in[2 * ii + 1];
      int k = length >> 1;
      while (ii >= k && k > 0) { ii -= k; k >>= 1; }
      ii += k;
    }
    for (int i = 0; i < 2 * length; i++) in[i] = rev[i];

    for (int i = 0; i < length; i += 4) {
      float* p = in + 2 * i;
      float aRe = p[0] + p[2], aIm = p[1] + p[3];
      float bRe = p[4] + p[6], bIm = p[5] + p[7];
      float cRe = p[0] - p[2], cIm = p[1] - p[3];
      float dRe = p[4] - p[6], dIm = p[5] - p[7];
      p[0] = aRe + bRe; p[1] = aIm + bIm;
      p[4] = aRe - bRe; p[5] = aIm - bIm;
      float e1Re = cRe - dIm, e1Im = cIm + dRe;
      float e2Re = cRe + dIm, e2Im = cIm - dRe;
      if (forward) { p[2] = e2Re; p[3] = e2Im; p[6] = e1Re; p[7] = e1Im; }
      else { p[2] = e1Re; p[3] = e1Im; p[6] = e2Re; p[7] = e2Im; }
    }
    const int imOff = forward ? 2 : 1;
    for (int i = 4; i < length; i <<= 1) {
      const int shift = i << 1;
      const int m = length / shift;
      for (int j = 0; j < length; j += shift) {
        for (int k = 0; k < i; k++) {
          int km = k * m;
          float rootRe = roots[km * cols];
          float rootIm = roots[km * cols + imOff];
          float* v0 = in + 2 * (j + k);
          float* v1 = in + 2 * (i + k + j);
          float zRe = v1[0] * rootRe - v1[1] * rootIm;
          float zIm = v1[0] * rootIm + v1[1] * rootRe;
          v1[0] = v0[0] - zRe;
          v1[1] = v0[1] - zIm;
          v0[0] = v0[0] + zRe;
          v0[1] = v0[1] + zIm;
        }
      }
    }
  }
};

struct MDCT {
  int N, N2, N4, N8;
  const float* sincos;  // [N4][2]
  FFT fft;
  std::vector<float> buf;  // [N4][2]
  explicit MDCT(int length) : N(length), N2(length >> 1), N4(length >> 2), N8(length >> 3), fft(length >> 2) {
    sincos = (length == 2048) ? JT(MDCT_TABLE_2048) : JT(MDCT_TABLE_128);
    buf.resize(2 * N4);
  }
  // MDCT.process (MDCT.java:36-81)
  void process(const float* in, int inOff, float* out, int outOff) {
    float* b = buf.data();
    for (int k = 0; k < N4; k++) {
      b[2 * k + 1] = (in[inOff + 2 * k] * sincos[2 * k]) + (in[inOff + N2 - 1 - 2 * k] * sincos[2 * k + 1]);
      b[2 * k] = (in[inOff + N2 - 1 - 2 * k] * sincos[2 * k]) - (in[inOff + 2 * k] * sincos[2 * k + 1]);
    }
    fft.process(b, false);
    for (int k = 0; k < N4; k++) {
      float t0 = b[2 * k], t1 = b[2 * k + 1];
      b[2 * k + 1] = (t1 * sincos[2 * k]) + (t0 * sincos[2 * k + 1]);
      b[2 * k] = (t0 * sincos[2 * k]) - (t1 * sincos[2 * k + 1]);
    }
#define RE(i) b[2 * (i)]
#define IM(i) b[2 * (i) + 1]
    for (int k = 0; k < N8; k += 2) {
      out[outOff + 2 * k] = IM(N8 + k);
      out[outOff + 2 + 2 * k] = IM(N8 + 1 + k);
      out[outOff + 1 + 2 * k] = -RE(N8 - 1 - k);
      out[outOff + 3 + 2 * k] = -RE(N8 - 2 - k);
      out[outOff + N4 + 2 * k] = RE(k);
      out[outOff + N4 + 2 + 2 * k] = RE(1 + k);
      out[outOff + N4 + 1 + 2 * k] = -IM(N4 - 1 - k);
      out[outOff + N4 + 3 + 2 * k] = -IM(N4 - 2 - k);
      out[outOff + N2 + 2 * k] = RE(N8 + k);
      out[outOff + N2 + 2 + 2 * k] = RE(N8 + 1 + k);
      out[outOff + N2 + 1 + 2 * k] = -IM(N8 - 1 - k);
      out[outOff + N2 + 3 + 2 * k] = -IM(N8 - 2 - k);
      out[outOff + N2 + N4 + 2 * k] = -IM(k);
      out[outOff + N2 + N4 + 2 + 2 * k] = -IM(1 + k);
      out[outOff + N2 + N4 + 1 + 2 * k] = RE(N4 - 1 - k);
      out[outOff + N2 + N4 + 3 + 2 * k] = RE(N4 - 2 - k);
    }
#undef RE
#undef IM
  }
};

struct FilterBank {
  static const int length = 1024, shortLen = 128, mid = 448, trans = 64;
  MDCT mdctShort, mdctLong;
  std::vector<float> buf;
  const float* LONG_WINDOWS[2];
  const float* SHORT_WINDOWS[2];
  FilterBank() : mdctShort(256), mdctLong(2048), buf(2048) {
    LONG_WINDOWS[0] = JT(SINE_1024); LONG_WINDOWS[1] = JT(KBD_1024);
    SHORT_WINDOWS[0] = JT(SINE_128); SHORT_WINDOWS[1] = JT(KBD_128);
  }
  // FilterBank.process (FilterBank.java:39-123)
  void process(WindowSequence ws, int windowShape, int windowShapePrev, const float* in, float* out, float* overlap) {
    float* b = buf.data();
    const float* LW = LONG_WINDOWS[windowShape];
    const float* LWp = LONG_WINDOWS[windowShapePrev];
    const float* SW = SHORT_WINDOWS[windowShape];
    const float* SWp = SHORT_WINDOWS[windowShapePrev];
    switch (ws) {
      case ONLY_LONG_SEQUENCE:
        mdctLong.process(in, 0, b, 0);
        for (int i = 0; i < length; i++) out[i] = overlap[i] + (b[i] * LWp[i]);
        for (int i = 0; i < length; i++) overlap[i] = b[length + i] * LW[length - 1 - i];
        break;
      case LONG_START_SEQUENCE:
        mdctLong.process(in, 0, b, 0);
        for (int i = 0; i < length; i++) out[i] = overlap[i] + (b[i] * LWp[i]);
        for (int i = 0; i < mid; i++) overlap[i] = b[length + i];
        for (int i = 0; i < shortLen; i++) overlap[mid + i] = b[length + mid + i] * SW[shortLen - i - 1];
        for (int i = 0; i < mid; i++) overlap[mid + shortLen + i] = 0;
        break;
      case EIGHT_SHORT_SEQUENCE:
        for (int i = 0; i < 8; i++) mdctShort.process(in, i * shortLen, b, 2 * i * shortLen);
        for (int i = 0; i < mid; i++) out[i] = overlap[i];
        for (int i = 0; i < shortLen; i++) {
          out[mid + i] = overlap[mid + i] + (b[i] * SWp[i]);
          out[mid + 1 * shortLen + i] = overlap[mid + shortLen * 1 + i] + (b[shortLen * 1 + i] * SW[shortLen - 1 - i]) + (b[shortLen * 2 + i] * SW[i]);
          out[mid + 2 * shortLen + i] = overlap[mid + shortLen * 2 + i] + (b[shortLen * 3 + i] * SW[shortLen - 1 - i]) + (b[shortLen * 4 + i] * SW[i]);
          out[mid + 3 * shortLen + i] = overlap[mid + shortLen * 3 + i] + (b[shortLen * 5 + i] * SW[shortLen - 1 - i]) + (b[shortLen * 6 + i] * SW[i]);
          if (i < trans)
            out[mid + 4 * shortLen + i] = overlap[mid + shortLen * 4 + i] + (b[shortLen * 7 + i] * SW[shortLen - 1 - i]) + (b[shortLen * 8 + i] * SW[i]);
        }
        for (int i = 0; i < shortLen; i++) {
          if (i >= trans)
            overlap[mid + 4 * shortLen + i - length] = (b[shortLen * 7 + i] * SW[shortLen - 1 - i]) + (b[shortLen * 8 + i] * SW[i]);
          overlap[mid + 5 * shortLen + i - length] = (b[shortLen * 9 + i] * SW[shortLen - 1 - i]) + (b[shortLen * 10 + i] * SW[i]);
          overlap[mid + 6 * shortLen + i - length] = (b[shortLen * 11 + i] * SW[shortLen - 1 - i]) + (b[shortLen * 12 + i] * SW[i]);
          overlap[mid + 7 * shortLen + i - length] = (b[shortLen * 13 + i] * SW[shortLen - 1 - i]) + (b[shortLen * 14 + i] * SW[i]);
          overlap[mid + 8 * shortLen + i - length] = (b[shortLen * 15 + i] * SW[shortLen - 1 - i]);
        }
        for (int i = 0; i < mid; i++) overlap[mid + shortLen + i] = 0;
        break;
      case LONG_STOP_SEQUENCE:
        mdctLong.process(in, 0, b, 0);
        for (int i = 0; i < mid; i++) out[i] = overlap[i];
        for (int i = 0; i < shortLen; i++) out[mid + i] = overlap[mid + i] + (b[mid + i] * SWp[i]);
        for (int i = 0; i < mid; i++) out[mid + shortLen + i] = overlap[mid + shortLen + i] + b[mid + shortLen + i];
        for (int i = 0; i < length; i++) overlap[i] = b[length + i] * LW[length - 1 - i];
        break;
    }
  }
};

// ---------------------------------------------------------------------------
// TNS: parsed as JAAD does (tools/TNS.java:35-61).  JAAD never applies it (TNS.process is a stub, :63-68); process()
// below is the tool as ISO/IEC 14496-3 4.6.9.3 defines it (tns_decode_frame / tns_decode_coef / tns_ar_filter) and runs
// only in tnsMode 1.  Arithmetic: binary32, one rounding per operation, in the order written here (the engine's kernel
// does the same operations in the same order, so the two agree bit for bit; the float64 direct form in
// tests/test_oracle_cpu.py bounds the distance to the ideal filter).
// ---------------------------------------------------------------------------
// SampleFrequency.getMaximalTNS_SFB (SampleFrequency.java:15-26, second constructor array): {long, short}
static const int TNS_MAX_SFB[12][2] = {{31, 9}, {31, 9}, {34, 10}, {40, 14}, {42, 14}, {51, 14},
                                       {46, 14}, {46, 14}, {42, 14}, {42, 14}, {42, 14}, {39, 14}};

struct TNS {
  int nFilt[8] = {0};
  int length[8][4] = {{0}}, order[8][4] = {{0}};
  bool direction[8][4] = {{false}};
  float coef[8][4][20] = {{{0}}};

  // 4.6.9.3: spec is the channel's spectrum after M/S and intensity stereo, windows of 128 for EIGHT_SHORT
  void process(const ICSInfo& info, float* spec) const {
    const bool sh = info.isEightShortFrame();
    const int maxTns = TNS_MAX_SFB[info.sfIndex][sh ? 1 : 0];
    for (int w = 0; w < info.windowCount; w++) {
      int bottom = info.swbCount;
      for (int f = 0; f < nFilt[w]; f++) {
        const int top = bottom;
        bottom = std::max(top - length[w][f], 0);
        const int ord = std::min(order[w][f], 20);
        if (!ord) continue;
        // tns_decode_coef.  TNSTables holds the NEGATED sin-mapped values (tools/TNSTables.java:10-25 has -sin(..) where
        // 4.6.9.3 has tmp2 = sin(coef / iqfac)), so tmp2 = -coef, an exact operation.
        float lpc[21], b[21];
        lpc[0] = 1.0f;
        for (int m = 1; m <= ord; m++) {
          const float t = -coef[w][f][m - 1];
          for (int i = 1; i < m; i++) b[i] = lpc[i] + (t * lpc[m - i]);
          for (int i = 1; i < m; i++) lpc[i] = b[i];
          lpc[m] = t;
        }
        const int start = info.swbOffsets[std::min(std::min(bottom, maxTns), info.maxSFB)];
        const int end = info.swbOffsets[std::min(std::min(top, maxTns), info.maxSFB)];
        const int size = end - start;
        if (size <= 0) continue;
        int pos = w * 128 + start, inc = 1;
        if (direction[w][f]) { inc = -1; pos = w * 128 + end - 1; }
        // tns_ar_filter: y(n) = x(n) - lpc[1] y(n-1) - ... - lpc[order] y(n-order), zero initial state
        float state[20];
        for (int j = 0; j < 20; j++) state[j] = 0.f;
        for (int i = 0; i < size; i++, pos += inc) {
          float y = spec[pos];
          for (int j = 0; j < ord; j++) y = y - (state[j] * lpc[j + 1]);
          for (int j = ord - 1; j > 0; j--) state[j] = state[j - 1];
          state[0] = y;
          spec[pos] = y;
        }
      }
    }
  }

  void decode(BitStream& in, const ICSInfo& info) {
    static const int SHORT_BITS[3] = {1, 4, 3}, LONG_BITS[3] = {2, 6, 5};
    const int* bits = info.isEightShortFrame() ? SHORT_BITS : LONG_BITS;
    const float* TNS_TABLES[4] = {JT(TNS_COEF_0_3), JT(TNS_COEF_0_4), JT(TNS_COEF_1_3), JT(TNS_COEF_1_4)};
    for (int w = 0; w < info.windowCount; w++) {
      if ((nFilt[w] = in.readBits(bits[0])) != 0) {
        int coefRes = in.readBit();
        for (int filt = 0; filt < nFilt[w]; filt++) {
          length[w][filt] = in.readBits(bits[1]);
          if ((order[w][filt] = in.readBits(bits[2])) > 20) throw AACException(ST_TNS_ORDER, "TNS filter out of range");
          else if (order[w][filt] != 0) {
            direction[w][filt] = in.readBool();
            int coefCompress = in.readBit();
            int coefLen = coefRes + 3 - coefCompress;
            int tmp = 2 * coefCompress + coefRes;
            for (int i = 0; i < order[w][filt]; i++) coef[w][filt][i] = TNS_TABLES[tmp][in.readBits(coefLen)];
          }
        }
      }
    }
  }
};

// ---------------------------------------------------------------------------
// ICStream (syntax/ICStream.java)
// ---------------------------------------------------------------------------
struct ICStream {
  static const int MAX_SECTIONS = 120;
  ICSInfo info;
  int sfbCB[MAX_SECTIONS], sectEnd[MAX_SECTIONS];
  float iqData[1024];
  float scaleFactors[MAX_SECTIONS];
  int globalGain = 0;
  bool pulseDataPresent = false, tnsDataPresent = false, gainControlPresent = false;
  TNS tns;
  float overlap[1024];
  // taps for the parity tests (not part of JAAD): raw integers behind the floats
  int16_t q[1024];        // quantised coefficients in iqData layout
  int16_t sfIndex[MAX_SECTIONS];  // SCALEFACTOR_TABLE index (-1: 0.0f); noise bands store index|0x4000
  // pulse_data as parsed (JAAD keeps pulseOffset[] / pulseAmp[] the same way, ICStream.java:39-41, and never reads them again)
  int pulseCount = 0, pulseOffset[4] = {0, 0, 0, 0}, pulseAmp[4] = {0, 0, 0, 0};
  bool infoDecoded = false;

  explicit ICStream(const DecoderConfig& c) : info(c) {
    std::fill(sfbCB, sfbCB + MAX_SECTIONS, 0);
    std::fill(sectEnd, sectEnd + MAX_SECTIONS, 0);
    std::fill(iqData, iqData + 1024, 0.f);
    std::fill(scaleFactors, scaleFactors + MAX_SECTIONS, 0.f);
    std::fill(overlap, overlap + 1024, 0.f);
    std::fill(q, q + 1024, (int16_t)0);
    std::fill(sfIndex, sfIndex + MAX_SECTIONS, (int16_t)-1);
  }

  // ICStream.decode (:60-111)
  void decode(BitStream& in, bool commonWindow, const DecoderConfig& conf) {
    globalGain = in.readBits(8);
    if (!commonWindow) { info.decode(in, commonWindow); infoDecoded = true; }
    decodeSectionData(in);
    decodeScaleFactors(in);
    pulseDataPresent = in.readBool();
    if (pulseDataPresent) {
      if (info.isEightShortFrame()) throw AACException(ST_PULSE_SHORT, "pulse data not allowed for short frames");
      decodePulseData(in);
    }
    tnsDataPresent = in.readBool();
    if (tnsDataPresent && !conf.profile.errorResilient()) tns.decode(in, info);
    gainControlPresent = in.readBool();
    if (gainControlPresent)
      throw AACException(ST_UNSUPPORTED_ELEMENT, "gain control (SSR) is outside the engine's scope");
    decodeSpectralData(in, conf);
  }

  // ICStream.processTNS (:313-316) -> TNS.process
  void processTNS(const DecoderConfig& conf) {
    if (tnsDataPresent && conf.tnsMode == 1) tns.process(info, iqData);
  }

  void decodeSectionData(BitStream& in) {  // :113-146
    std::fill(sfbCB, sfbCB + MAX_SECTIONS, 0);
    std::fill(sectEnd, sectEnd + MAX_SECTIONS, 0);
    const int bits = info.isEightShortFrame() ? 3 : 5;
    const int escVal = (1 << bits) - 1;
    const int windowGroupCount = info.windowGroupCount;
    const int maxSFB = info.maxSFB;
    int idx = 0;
    for (int g = 0; g < windowGroupCount; g++) {
      for (int k = 0; k < maxSFB;) {
        int end = k;
        int cb = in.readBits(4);
        if (cb == 12) throw AACException(ST_INVALID_CODEBOOK, "invalid huffman codebook: 12");
        int incr;
        do { incr = in.readBits(bits); end += incr; } while (incr == escVal);
        if (end > maxSFB) throw AACException(ST_TOO_MANY_BANDS, "too many bands");
        for (; k < end; k++, idx++) {
          if (idx >= MAX_SECTIONS) throw AACException(ST_ARRAY_BOUNDS, "section index out of bounds");
          sfbCB[idx] = cb;
          sectEnd[idx] = end;
        }
      }
    }
  }

  void decodePulseData(BitStream& in) {  // :148-170  (parsed; applied in pulseMode 1 only)
    pulseCount = in.readBits(2) + 1;
    int pulseStartSWB = in.readBits(6);
    if (pulseStartSWB >= info.swbCount) throw AACException(ST_PULSE_RANGE, "pulse SWB out of range");
    int off = info.swbOffsets[pulseStartSWB];
    off += in.readBits(5);
    pulseOffset[0] = off;
    pulseAmp[0] = in.readBits(4);
    for (int i = 1; i < pulseCount; i++) {
      off = in.readBits(5) + off;
      if (off > 1023) throw AACException(ST_PULSE_RANGE, "pulse offset out of range");
      pulseOffset[i] = off;
      pulseAmp[i] = in.readBits(4);
    }
  }

  // Not in JAAD (pulseMode 1).  ISO/IEC 14496-3 4.6.3.3: "if (quant[k] > 0) quant[k] += pulse_amp else quant[k] -= pulse_amp",
  // ahead of the inverse quantisation; here after JAAD's fused decode + dequantisation, redoing that one coefficient with the
  // same two operations (IQ_TABLE look-up, one multiplication by the band's scalefactor).  Only coefficients of bands that
  // carry spectral data take a pulse (codebooks 1..11 below max_sfb): the others have no quantised value and no scalefactor
  // of their own (FFmpeg's decoder draws the same line).  A magnitude past IQ_TABLE's 8191 entries fails like any other.
  void applyPulses() {
    const float* IQ = JT(IQ_TABLE);
    for (int i = 0; i < pulseCount; i++) {
      const int pos = pulseOffset[i];
      int sfb = 0;
      while (sfb < info.maxSFB && info.swbOffsets[sfb + 1] <= pos) sfb++;
      if (sfb >= info.maxSFB) continue;
      const int hcb = sfbCB[sfb];
      if (hcb < 1 || hcb > 11) continue;
      const int v = q[pos] > 0 ? q[pos] + pulseAmp[i] : q[pos] - pulseAmp[i];
      const int a = v > 0 ? v : -v;
      if (a > 8190) throw AACException(ST_ARRAY_BOUNDS, "IQ table index out of range");
      iqData[pos] = (v > 0) ? IQ[v] : -IQ[-v];
      iqData[pos] *= scaleFactors[sfb];
      q[pos] = (int16_t)v;
    }
  }

  void decodeScaleFactors(BitStream& in) {  // :172-220
    const float* TAB = JT(SCALEFACTOR_TABLE);
    const int windowGroups = info.windowGroupCount;
    const int maxSFB = info.maxSFB;
    int offset[3] = {globalGain, globalGain - 90, 0};
    bool noiseFlag = true;
    std::fill(sfIndex, sfIndex + MAX_SECTIONS, (int16_t)-1);  // tap only
    for (int g = 0, idx = 0; g < windowGroups; g++) {
      for (int sfb = 0; sfb < maxSFB;) {
        int end = sectEnd[idx];
        switch (sfbCB[idx]) {
          case 0:
            for (; sfb < end; sfb++, idx++) { scaleFactors[idx] = 0; sfIndex[idx] = -1; }
            break;
          case 15: case 14:
            for (; sfb < end; sfb++, idx++) {
              offset[2] += decodeScaleFactor(in) - 60;
              int tmp = std::min(std::max(offset[2], -155), 100);
              scaleFactors[idx] = TAB[-tmp + 200];
              sfIndex[idx] = (int16_t)(-tmp + 200);
            }
            break;
          case 13:
            for (; sfb < end; sfb++, idx++) {
              if (noiseFlag) { offset[1] += in.readBits(9) - 256; noiseFlag = false; }
              else offset[1] += decodeScaleFactor(in) - 60;
              int tmp = std::min(std::max(offset[1], -100), 155);
              scaleFactors[idx] = -TAB[tmp + 200];
              sfIndex[idx] = (int16_t)((tmp + 200) | 0x4000);
            }
            break;
          default:
            for (; sfb < end; sfb++, idx++) {
              offset[0] += decodeScaleFactor(in) - 60;
              if (offset[0] > 255) throw AACException(ST_SF_RANGE, "scalefactor out of range");
              if (offset[0] + 100 < 0) throw AACException(ST_ARRAY_BOUNDS, "scalefactor index negative");
              scaleFactors[idx] = TAB[offset[0] - 100 + 200];
              sfIndex[idx] = (int16_t)(offset[0] + 100);
            }
            break;
        }
      }
    }
  }

  // PNS random generator (ICStream.java:26,247): static in JAAD, per Decoder here -- see DecoderConfig::pnsState
  void decodeSpectralData(BitStream& in, const DecoderConfig& conf) {  // :222-275
    const float* IQ = JT(IQ_TABLE);
    std::fill(iqData, iqData + 1024, 0.f);
    std::fill(q, q + 1024, (int16_t)0);
    const int maxSFB = info.maxSFB;
    const int windowGroups = info.windowGroupCount;
    const int16_t* offsets = info.swbOffsets;
    int buf[4];
    for (int g = 0, idx = 0, groupOff = 0; g < windowGroups; g++) {
      int groupLen = info.windowGroupLength[g];
      for (int sfb = 0; sfb < maxSFB; sfb++, idx++) {
        int hcb = sfbCB[idx];
        int off = groupOff + offsets[sfb];
        int width = offsets[sfb + 1] - offsets[sfb];
        if (hcb == 0 || hcb == 15 || hcb == 14) {
          for (int w = 0; w < groupLen; w++, off += 128) {
            if (off < 0 || off > off + width || off + width > 1024) throw AACException(ST_ARRAY_BOUNDS, "band out of range");
            std::fill(iqData + off, iqData + off + width, 0.f);
          }
        } else if (hcb == 13) {
          for (int w = 0; w < groupLen; w++, off += 128) {
            // (a band at the end of the offset table has a negative width: the Java loops do not run, nothing is thrown)
            if (width > 0 && (off < 0 || off + width > 1024)) throw AACException(ST_ARRAY_BOUNDS, "band out of range");
            if (width <= 0) continue;
            float energy = 0;
            for (int k = 0; k < width; k++) {
              int32_t& rs = conf.pnsState;
              rs = (int32_t)(1664525u * (uint32_t)rs + 1013904223u);
              iqData[off + k] = (float)rs;
              energy += iqData[off + k] * iqData[off + k];
            }
            const float scale = (float)((double)scaleFactors[idx] / std::sqrt((double)energy));
            for (int k = 0; k < width; k++) iqData[off + k] *= scale;
          }
        } else {
          for (int w = 0; w < groupLen; w++, off += 128) {
            int num = (hcb >= 5) ? 2 : 4;
            for (int k = 0; k < width; k += num) {
              decodeSpectralData_(in, hcb, buf);
              for (int j = 0; j < num; j++) {
                int pos = off + k + j;
                if (pos < 0 || pos >= 1024) throw AACException(ST_ARRAY_BOUNDS, "coefficient index out of range");
                int a = buf[j] > 0 ? buf[j] : -buf[j];
                if (a > 8190) throw AACException(ST_ARRAY_BOUNDS, "IQ table index out of range");
                iqData[pos] = (buf[j] > 0) ? IQ[buf[j]] : -IQ[-buf[j]];
                iqData[pos] *= scaleFactors[idx];
                q[pos] = (int16_t)buf[j];
              }
            }
          }
        }
      }
      groupOff += groupLen << 7;
    }
    if (conf.pulseMode == 1 && pulseDataPresent) applyPulses();
  }
  static void decodeSpectralData_(BitStream& in, int hcb, int* buf) { ::jaad::decodeSpectralData(in, hcb, buf, 0); }

  // ICStream.process (:308-311)
  void process(float* data, FilterBank& fb) {
    fb.process(info.windowSequence, info.windowShape[ICSInfo::CURRENT], info.windowShape[ICSInfo::PREVIOUS], iqData, data, overlap);
  }
};

// ---------------------------------------------------------------------------
// SBR hook (sbr/SBR.java); the implementation lives in jaad_sbr.hpp
// ---------------------------------------------------------------------------
struct SBRBase {
  bool valid = false;
  bool downSampled = false;   // SBR.isSBRDownSampled (sbr/SBR.java:29-33): 32-band synthesis, output stays at the core's length
  virtual ~SBRBase() {}
  void invalidate() { valid = false; }
  bool isValid() const { return valid; }
  virtual void decode(BitStream& ld, bool crc) = 0;
  virtual void process(float* left, float* right) = 0;
};
// Factory set by jaad_sbr.hpp (stereo=false -> SBR1, true -> SBR2).
using SBRFactory = SBRBase* (*)(DecoderConfig& config, bool stereo);
inline SBRFactory& sbrFactory() { static SBRFactory f = nullptr; return f; }

inline void sbrUpsample(float* data, int len) {  // SBR.upsample (sbr/SBR.java:302-309)
  for (int i = len / 2 - 1; i > 0; --i) {
    float v = data[i];
    data[2 * i] = v;
    data[2 * i + 1] = v;
  }
}

// ---------------------------------------------------------------------------
// Elements (syntax/ChannelElement.java, SCE.java, CPE.java, LFE.java)
// ---------------------------------------------------------------------------
enum ElementType { EL_SCE = 0, EL_CPE = 1, EL_CCE = 2, EL_LFE = 3, EL_DSE = 4, EL_PCE = 5, EL_FIL = 6, EL_END = 7 };

struct ChannelElement {
  DecoderConfig* config;
  int type, tag;
  std::unique_ptr<SBRBase> sbr;
  std::vector<float> dataL, dataR;
  ChannelElement(DecoderConfig& c, int ty, int tg) : config(&c), type(ty), tag(tg) {}
  virtual ~ChannelElement() {}
  virtual void decode(BitStream& in) { if (sbr) sbr->invalidate(); }  // ChannelElement.java:58-61
  virtual bool sbrAllowed() const { return true; }
  virtual bool stereoSBR() const = 0;
  void decodeSBR(BitStream& in, bool crc) {  // ChannelElement.java:65-76
    if (!config->sbrEnabled) return;
    if (!sbr && sbrAllowed() && sbrFactory()) sbr.reset(sbrFactory()(*config, stereoSBR()));
    if (sbr) sbr->decode(in, crc);
  }
  bool isSBRPresent() const { return sbr && sbr->isValid(); }
  float* getDataL() { if (dataL.empty()) dataL.assign(config->getSampleLength(), 0.f); return dataL.data(); }
  float* getDataR() { if (dataR.empty()) dataR.assign(config->getSampleLength(), 0.f); return dataR.data(); }
  virtual void process(FilterBank& fb, std::vector<std::pair<float*, int>>& target) = 0;
};

struct SCE : ChannelElement {
  ICStream ics;
  SCE(DecoderConfig& c, int ty, int tg) : ChannelElement(c, ty, tg), ics(c) {}
  bool sbrAllowed() const override { return type != EL_LFE; }  // LFE.openSBR returns null (LFE.java:48-50)
  bool stereoSBR() const override { return false; }
  void decode(BitStream& in) override {  // SCE.java:67-70
    ChannelElement::decode(in);
    ics.infoDecoded = false;
    ics.decode(in, false, *config);
  }
  void process(FilterBank& fb, std::vector<std::pair<float*, int>>& target) override {  // SCE.java:90-133
    float* dL = getDataL();
    ics.processTNS(*config);
    ics.process(dL, fb);
    target.push_back({dL, (int)dataL.size()});
    if (isSBRPresent() && config->sbrEnabled) {
      float* dR = getDataR();
      // the element's buffers were created before the stream switched to SBR output: the Java code runs into an
      // ArrayIndexOutOfBoundsException inside the synthesis filterbank (memory safety matters more here than the exact spot)
      if (!sbr->downSampled && ((int)dataL.size() < 2 * config->getFrameLength() || (!dataR.empty() && (int)dataR.size() < 2 * config->getFrameLength())))
        throw AACException(ST_ARRAY_BOUNDS, "SBR output does not fit the element's buffers");
      sbr->process(dL, dR);
      target.push_back({dR, (int)dataR.size()});
    } else if ((int)dataL.size() != config->getFrameLength()) {
      sbrUpsample(dL, (int)dataL.size());
    }
  }
};

struct CPE : ChannelElement {
  int msMask = 0;
  bool msUsed[128];
  bool commonWindow = false;
  ICStream icsL, icsR;
  CPE(DecoderConfig& c, int tg) : ChannelElement(c, EL_CPE, tg), icsL(c), icsR(c) { std::fill(msUsed, msUsed + 128, false); }
  bool stereoSBR() const override { return true; }
  bool isMSMaskPresent() const { return msMask != 0; }

  void decode(BitStream& in) override {  // CPE.java:85-123
    ChannelElement::decode(in);
    icsL.infoDecoded = icsR.infoDecoded = false;
    commonWindow = in.readBool();
    if (commonWindow) {
      icsL.info.decode(in, commonWindow);
      icsL.infoDecoded = true;
      icsR.info.setCommonData(icsL.info);
      icsR.infoDecoded = true;
      msMask = in.readBits(2);
      if (msMask == 1) {
        const int n = icsL.info.windowGroupCount * icsL.info.maxSFB;
        for (int idx = 0; idx < n; idx++) {
          if (idx >= 128) throw AACException(ST_ARRAY_BOUNDS, "ms_used index out of bounds");
          msUsed[idx] = in.readBool();
        }
      } else if (msMask == 2) std::fill(msUsed, msUsed + 128, true);
      else if (msMask == 0) std::fill(msUsed, msUsed + 128, false);
      else throw AACException(ST_MS_RESERVED, "reserved MS mask type used");
    } else {
      msMask = 0;
      std::fill(msUsed, msUsed + 128, false);
    }
    icsL.decode(in, commonWindow, *config);
    icsR.decode(in, commonWindow, *config);
  }

  void processMS() {  // tools/MS.java:17-41
    const ICSInfo& info = icsL.info;
    const int16_t* offsets = info.swbOffsets;
    float* specL = icsL.iqData;
    float* specR = icsR.iqData;
    for (int g = 0, idx = 0, groupOff = 0; g < info.windowGroupCount; g++) {
      for (int i = 0; i < info.maxSFB; i++, idx++) {
        if (msUsed[idx] && icsL.sfbCB[idx] < 13 && icsR.sfbCB[idx] < 13) {
          for (int w = 0; w < info.windowGroupLength[g]; w++) {
            const int off = groupOff + w * 128 + offsets[i];
            for (int j = 0; j < offsets[i + 1] - offsets[i]; j++) {
              float t = specL[off + j] - specR[off + j];
              specL[off + j] += specR[off + j];
              specR[off + j] = t;
            }
          }
        }
      }
      groupOff += info.windowGroupLength[g] * 128;
    }
  }

  void processIS() {  // tools/IS.java:17-53
    const ICSInfo& info = icsR.info;
    const int16_t* offsets = info.swbOffsets;
    float* specL = icsL.iqData;
    float* specR = icsR.iqData;
    int idx = 0, groupOff = 0;
    for (int g = 0; g < info.windowGroupCount; g++) {
      for (int i = 0; i < info.maxSFB;) {
        if (icsR.sfbCB[idx] == 15 || icsR.sfbCB[idx] == 14) {
          int end = icsR.sectEnd[idx];
          for (; i < end; i++, idx++) {
            int c = icsR.sfbCB[idx] == 15 ? 1 : -1;
            if (isMSMaskPresent()) c *= msUsed[idx] ? -1 : 1;
            float scale = (float)c * icsR.scaleFactors[idx];
            for (int w = 0; w < info.windowGroupLength[g]; w++) {
              int off = groupOff + w * 128 + offsets[i];
              for (int j = 0; j < offsets[i + 1] - offsets[i]; j++) specR[off + j] = specL[off + j] * scale;
            }
          }
        } else {
          int end = icsR.sectEnd[idx];
          idx += end - i;
          i = end;
        }
      }
      groupOff += info.windowGroupLength[g] * 128;
    }
  }

  void process(FilterBank& fb, std::vector<std::pair<float*, int>>& target) override {  // CPE.java:149-208
    float* dL = getDataL();
    float* dR = getDataR();
    if (commonWindow & isMSMaskPresent()) processMS();
    processIS();
    // TNS.process is a stub in JAAD (tools/TNS.java:63-68): nothing happens in tnsMode 0
    icsL.processTNS(*config);
    icsR.processTNS(*config);
    icsL.process(dL, fb);
    icsR.process(dR, fb);
    if (isSBRPresent() && config->sbrEnabled) {
      // the element's buffers were created before the stream switched to SBR output: the Java code runs into an
      // ArrayIndexOutOfBoundsException inside the synthesis filterbank (memory safety matters more here than the exact spot)
      if (!sbr->downSampled && ((int)dataL.size() < 2 * config->getFrameLength() || (!dataR.empty() && (int)dataR.size() < 2 * config->getFrameLength())))
        throw AACException(ST_ARRAY_BOUNDS, "SBR output does not fit the element's buffers");
      sbr->process(dL, dR);
    } else if ((int)dataL.size() != config->getFrameLength()) {
      sbrUpsample(dL, (int)dataL.size());
      sbrUpsample(dR, (int)dataR.size());
    }
    target.push_back({dL, (int)dataL.size()});
    target.push_back({dR, (int)dataR.size()});
  }
};

// ---------------------------------------------------------------------------
// SyntacticElements (syntax/SyntacticElements.java) + Decoder (Decoder.java)
// ---------------------------------------------------------------------------
struct SyntacticElements {
  DecoderConfig* config;
  FilterBank filterBank;
  std::map<int, std::unique_ptr<ChannelElement>> elements;  // key = type + 8*tag (Element.java:36-38)
  std::vector<ChannelElement*> audioElements;
  std::vector<std::pair<float*, int>> channels;

  explicit SyntacticElements(DecoderConfig& c) : config(&c) {}
  void startNewFrame() { audioElements.clear(); channels.clear(); }

  ChannelElement* getElement(int type, int tag) {
    int key = type + 8 * tag;
    auto it = elements.find(key);
    if (it != elements.end()) return it->second.get();
    ChannelElement* e = (type == EL_CPE) ? (ChannelElement*)new CPE(*config, tag) : (ChannelElement*)new SCE(*config, type, tag);
    elements[key].reset(e);
    return e;
  }

  void decodeChannelElement(int type, BitStream& in) {  // :134-159
    int id = in.readBits(4);
    ChannelElement* e = getElement(type, id);
    e->decode(in);
    audioElements.push_back(e);
  }

  void decodeDSE(BitStream& in) {  // syntax/DSE.java:54-66
    in.readBits(4);  // tag
    const bool byteAlign = in.readBool();
    int count = in.readBits(8);
    if (count == 255) count += in.readBits(8);
    if (byteAlign) in.byteAlign();
    for (int i = 0; i < count; i++) in.readBits(8);
  }

  // DRC.decode (syntax/DRC.java:31-83), reached from decodeFIL for extension type 11 (EXT_DYNAMIC_RANGE).  JAAD parses
  // dynamic_range_info into a DRC object nobody reads; what matters here is where the parse can end a frame: a read past the
  // fill element's sub-stream (EOSException), and `excludeMask = new boolean[7]` (DRC.java:27) taking a second group of seven
  // excluded-channel flags (DRC.java:72-82: ArrayIndexOutOfBoundsException; the flag is read before the store is checked,
  // JLS 15.26.1, so an over-read there is still the EOS).
  static void decodeDynamicRangeInfo(BitStream& in) {
    int bandCount = 1;
    if (in.readBool()) { in.readBits(4); in.readBits(4); }
    if (in.readBool()) {
      int exclChs = 0;
      do {
        for (int i = 0; i < 7; i++) {
          in.readBool();
          if (exclChs >= 7) throw AACException(ST_ARRAY_BOUNDS, "excludeMask index out of bounds");
          exclChs++;
        }
      } while (exclChs < 57 && in.readBool());
    }
    if (in.readBool()) {
      int bandsIncrement = in.readBits(4);
      in.readBits(4);
      bandCount += bandsIncrement;
      for (int i = 0; i < bandCount; i++) in.readBits(8);
    }
    if (in.readBool()) { in.readBits(7); in.readBits(1); }
    for (int i = 0; i < bandCount; i++) { in.readBool(); in.readBits(7); }
  }

  void decodeFIL(BitStream& in0) {  // :169-203
    int count = in0.readBits(4);
    if (count == 15) count += in0.readBits(8) - 1;
    if (count == 0) return;
    BitStream in = in0.readSubStream(8 * count);
    int type = in.readBits(4);
    switch (type) {
      case 11: decodeDynamicRangeInfo(in); break;   // :219-224: "decoded but unused"
      case 13: case 14: {
        sbrPayloadSeen = true;
        ChannelElement* prev = audioElements.empty() ? nullptr : audioElements.back();
        if (prev) prev->decodeSBR(in, type == 14);
        break;
      }
      default: break;
    }
  }

  // tap for the corrupted-stream tools (not JAAD): the frame's parse reached a fill element that claims an SBR payload
  bool sbrPayloadSeen = false;

  void decode(BitStream& in) {  // :57-132 (non error-resilient branch)
    sbrPayloadSeen = false;
    if (config->profile.errorResilient())
      throw AACException(ST_UNSUPPORTED_ELEMENT, "error resilient syntax is outside the engine's scope");
    for (;;) {
      int type = in.readBits(3);
      if (type == EL_END) break;
      switch (type) {
        case EL_SCE: case EL_CPE: case EL_LFE: decodeChannelElement(type, in); break;
        case EL_DSE: decodeDSE(in); break;
        case EL_FIL: decodeFIL(in); break;
        default: throw AACException(ST_UNSUPPORTED_ELEMENT, "CCE/PCE elements are outside the engine's scope");
      }
    }
    in.byteAlign();
  }

  void process() {  // :235-248
    channels.clear();
    for (ChannelElement* e : audioElements) e->process(filterBank, channels);
    if (channels.size() == 1 && config->getChannelCount() > 1) channels.push_back(channels[0]);
  }
};

// Java Math.round(float): nearest int, ties toward +inf, saturating; NaN -> 0.
inline int javaRound(float x) {
  if (x != x) return 0;
  double r = std::floor((double)x + 0.5);  // exact in double for every float
  if (r >= 2147483647.0) return 2147483647;
  if (r <= -2147483648.0) return (-2147483647 - 1);
  return (int)r;
}

struct FrameOutput {
  int status = ST_OK;
  int channels = 0, sampleLength = 0, sampleRate = 0;
  std::vector<std::pair<float*, int>> planes;  // per channel: (data, length) -- valid until the next decode
};

struct Decoder {
  DecoderConfig config;
  SyntacticElements syn;
  int frames = 0;
  std::string lastError;

  explicit Decoder(const DecoderConfig& c) : config(c), syn(config) {}

  // Decoder.decodeFrame (Decoder.java:89-101) + decode0 (:103-121).  AACException is
  // reported as a status instead of propagating; EOS is swallowed as in JAAD.
  FrameOutput decodeFrame(const uint8_t* data, size_t n) {
    FrameOutput out;
    try {
      BitStream in(data, n);
      if ((uint32_t)in.peekBits(32) == 0x41444946u)
        throw AACException(ST_UNSUPPORTED_ELEMENT, "ADIF header is outside the engine's scope");
      if (!config.profile.supported()) throw AACException(ST_PROFILE, "unsupported profile");
      syn.startNewFrame();
      syn.decode(in);
      syn.process();
      out.planes = syn.channels;
      out.channels = (int)syn.channels.size();
      out.sampleLength = config.getSampleLength();
      out.sampleRate = config.getOutputFrequency().frequency;
    } catch (const AACException& e) {
      out.status = e.code;
      lastError = e.what();
    }
    ++frames;
    return out;
  }
};

// SampleBuffer.accept (S/SampleBuffer.java:168-209): interleave + round + clamp.
inline void sampleBufferAccept(const FrameOutput& f, int16_t* dst, bool bigEndian) {
  uint8_t* p = reinterpret_cast<uint8_t*>(dst);
  for (int is = 0; is < f.sampleLength; ++is) {
    for (const auto& pl : f.planes) {
      int k = (int)((long long)pl.second * is / f.sampleLength);
      int pulse = javaRound(pl.first[k]);
      int v = pulse > 32767 ? 32767 : (pulse < -32768 ? -32768 : pulse);
      uint16_t u = (uint16_t)(int16_t)v;
      if (bigEndian) { p[0] = (uint8_t)(u >> 8); p[1] = (uint8_t)u; }
      else { p[0] = (uint8_t)u; p[1] = (uint8_t)(u >> 8); }
      p += 2;
    }
  }
}

}  // namespace jaad
