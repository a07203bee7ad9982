// ORACLE -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the JAAD reference
// decoder (pucgenie/JAADec), used solely as the parity checker by tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
// Nothing under jaadec_b200/ (the product) may include, link or call this.
//
// Parity status: the reference ships no golden vectors and no JVM exists in the
// build image, so this restatement is "parity unpinned" against real JAAD
// output; it is pinned instead by (1) the generator's own ground truth for the
// integer stage, (2) a float64 direct-form IMDCT check, see tests/ and DESIGN.md.
//
// Bit reader: follows aac/.../syntax/ByteArrayBitStream.java operation for
// operation (32-bit cache, leading size%4 bytes pre-loaded, EOS checks first).
#pragma once
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace jaad {

// Mirrors AACException / EOSException (aac/.../AACException.java:8, EOSException.java).
// `code` is the status word the C ABI reports (include/jaadb200.h, JAADB_ST_*).
struct AACException : std::runtime_error {
  int code;
  AACException(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};
struct EOSException : AACException {
  explicit EOSException(const std::string& m) : AACException(1, m) {}
};

enum Status : int {
  ST_OK = 0,
  ST_EOS = 1,                 // EOSException (swallowed by Decoder.decodeFrame)
  ST_INVALID_CODEBOOK = 2,    // "invalid huffman codebook: 12"        ICStream.java:129
  ST_TOO_MANY_BANDS = 3,      // "too many bands"                      ICStream.java:138
  ST_SF_RANGE = 4,            // "scalefactor out of range"            ICStream.java:213
  ST_PULSE_SHORT = 5,         // "pulse data not allowed for short"    ICStream.java:79
  ST_PULSE_RANGE = 6,         // pulse SWB/offset out of range         ICStream.java:152,166
  ST_MS_RESERVED = 7,         // "reserved MS mask type used"          CPE.java:114
  ST_TNS_ORDER = 8,           // "TNS filter out of range"             TNS.java:47
  ST_LTP_PROFILE = 9,         // "unexpected profile for LTP"          ICSInfo.java:139
  ST_UNSUPPORTED_ELEMENT = 10,// CCE/PCE or gain control (outside the engine's scope)
  ST_LAYOUT = 11,             // element sequence differs from the stream's channel layout
  ST_PROFILE = 12,            // "unsupported profile"                 Decoder.java:110
  ST_ARRAY_BOUNDS = 13,       // Java ArrayIndexOutOfBounds (bad codeword, IQ index > 8190, sf < 0 ...)
  ST_SBR = 14,                // AACException raised inside the SBR tool
  ST_CONFIG = 15              // bad AudioSpecificConfig
};

class BitStream {
 public:
  BitStream() = default;
  BitStream(const uint8_t* d, size_t n) { setData(d, n); }

  // ByteArrayBitStream.java:68-93
  void setData(const uint8_t* data, size_t size) {
    reset();
    length_ = 8 * (int)size;
    int shift = (int)(size % 4);
    bitsCached_ = 8 * shift;
    for (int i = 0; i < shift; ++i) cache_ = (cache_ << 8) | data[i];
    buf_ = data + shift;
    bufLen_ = (int)size - shift;
  }
  void reset() { pos_ = 0; length_ = 0; bitsCached_ = 0; cache_ = 0; position_ = 0; }

  int getPosition() const { return position_; }
  int getBitsLeft() const { return length_ - position_; }

  // ByteArrayBitStream.java:121-126 + sub-stream ctor :44-58
  BitStream readSubStream(int n) {
    if (getBitsLeft() < n) throw EOSException("stream overrun");
    BitStream s;
    s.length_ = position_ + n;
    s.buf_ = buf_; s.bufLen_ = bufLen_; s.pos_ = pos_;
    s.cache_ = cache_; s.bitsCached_ = bitsCached_; s.position_ = position_;
    skipBits(n);
    return s;
  }

  void byteAlign() {  // :128-133
    int toFlush = bitsCached_ & 7;
    if (toFlush > 0) skipBits(toFlush);
  }

  int readBits(int n) {  // :169-190
    if (getBitsLeft() < n) throw EOSException("stream overrun");
    uint32_t result;
    if (bitsCached_ >= n) {
      bitsCached_ -= n;
      result = (sar(cache_, bitsCached_)) & mask(n);
      position_ += n;
    } else {
      position_ += n;
      uint32_t c = cache_ & mask(bitsCached_);
      int left = n - bitsCached_;
      cache_ = readCache(false);
      bitsCached_ = 32 - left;
      result = (sar(cache_, bitsCached_) & mask(left)) | shl(c, left);
    }
    return (int)result;
  }
  int readBit() {  // :192-210
    if (getBitsLeft() < 1) throw EOSException("stream overrun");
    int i;
    if (bitsCached_ > 0) {
      bitsCached_--;
      i = (int)(sar(cache_, bitsCached_) & 1u);
    } else {
      cache_ = readCache(false);
      bitsCached_ = 31;
      i = (int)(sar(cache_, bitsCached_) & 1u);
    }
    position_++;
    return i;
  }
  bool readBool() { return (readBit() & 1) != 0; }

  int peekBits(int n) {  // :217-234
    if (getBitsLeft() < n) throw EOSException("stream overrun");
    uint32_t ret;
    if (bitsCached_ >= n) {
      ret = sar(cache_, bitsCached_ - n) & mask(n);
    } else {
      uint32_t c = cache_ & mask(bitsCached_);
      n -= bitsCached_;
      ret = (sar(readCache(true), 32 - n) & mask(n)) | shl(c, n);
    }
    return (int)ret;
  }

  void skipBits(int n) {  // :254-279
    if (getBitsLeft() < n) throw EOSException("stream overrun");
    position_ += n;
    if (n <= bitsCached_) {
      bitsCached_ -= n;
    } else {
      n -= bitsCached_;
      while (n >= 32) { n -= 32; readCache(false); }
      if (n > 0) { cache_ = readCache(false); bitsCached_ = 32 - n; }
      else { cache_ = 0; bitsCached_ = 0; }
    }
  }
  void skipBit() {  // :281-295
    if (getBitsLeft() < 1) throw EOSException("end of stream");
    position_++;
    if (bitsCached_ > 0) bitsCached_--;
    else { cache_ = readCache(false); bitsCached_ = 31; }
  }

 private:
  // Java `>>` on int is arithmetic; results are masked afterwards, so the sign
  // fill never shows, except for n==32 masks which Java special-cases (maskBits).
  static uint32_t sar(uint32_t v, int s) { return (uint32_t)((int32_t)v >> (s & 31)); }
  static uint32_t shl(uint32_t v, int s) { return v << (s & 31); }   // Java shifts mask the count to 5 bits
  static uint32_t mask(int n) { return n == 32 ? 0xFFFFFFFFu : ((1u << (n & 31)) - 1u); }

  uint32_t readCache(bool peek) {  // :154-167
    if (pos_ > bufLen_ - 4) throw EOSException("end of stream");
    uint32_t i = ((uint32_t)buf_[pos_] << 24) | ((uint32_t)buf_[pos_ + 1] << 16) |
                 ((uint32_t)buf_[pos_ + 2] << 8) | (uint32_t)buf_[pos_ + 3];
    if (!peek) pos_ += 4;
    return i;
  }

  const uint8_t* buf_ = nullptr;
  int bufLen_ = 0;
  int length_ = 0;
  int pos_ = 0;
  uint32_t cache_ = 0;
  int bitsCached_ = 0;
  int position_ = 0;
};

}  // namespace jaad
