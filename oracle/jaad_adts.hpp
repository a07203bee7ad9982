// ORACLE -- TEST INFRASTRUCTURE ONLY (see jaad_bits.hpp for the full notice).
//
// Restatement of JAAD's ADTS demultiplexer:
//   S/adts/ADTSDemultiplexer.java:25-58  (sync search, frame copy)
//   S/adts/ADTSFrame.java:68-117         (7-byte header bit twiddling, CRC skip)
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>

namespace jaad {

struct ADTSFrameInfo {
  size_t payloadOffset;  // first byte after the header (and CRC, when present)
  int payloadBytes;      // ADTSFrame.getFrameLength()
  int profile;           // Profile.forInt argument (2 bits + 1)
  int sfIndex;
  int channelConfiguration;
  bool protectionAbsent;
};

class ADTSDemultiplexer {
 public:
  ADTSDemultiplexer(const uint8_t* d, size_t n) : d_(d), n_(n) {}

  // findNextFrame + new ADTSFrame(din) + the payload copy loop of readNextFrame.
  // Returns false at end of input / when no sync is found within 6144 bytes.
  bool next(ADTSFrameInfo& f) {
    bool found = false;
    int left = 6144;  // MAXIMUM_FRAME_SIZE
    while (!found && left > 0) {
      if (pos_ >= n_) return false;
      int i = d_[pos_++];
      left--;
      if (i == 0xFF) {
        if (pos_ >= n_) return false;
        i = d_[pos_];  // read + unread
        if ((i & 0xF6) == 0xF0) found = true;
      }
    }
    if (!found) return false;
    if (pos_ + 6 > n_) return false;
    // ADTSFrame.readHeader (ADTSFrame.java:68-113); pos_ is at the byte after 0xFF
    int b1 = d_[pos_], b2 = d_[pos_ + 1], b3 = d_[pos_ + 2];
    int s = (d_[pos_ + 3] << 8) | d_[pos_ + 4];
    int b6 = d_[pos_ + 5];
    pos_ += 6;
    f.protectionAbsent = (b1 & 1) == 1;
    f.profile = ((b2 & 0xC0) >> 6) + 1;
    f.sfIndex = (b2 & 0x3C) >> 2;
    f.channelConfiguration = ((b2 & 1) << 2) | ((b3 & 0xC0) >> 6);
    int frameLength = ((b3 & 3) << 11) | ((s & 0xFFE0) >> 5);
    int rawDataBlockCount = b6 & 3;
    if (!f.protectionAbsent) pos_ += 2;  // crcCheck
    if (rawDataBlockCount != 0 && !f.protectionAbsent) pos_ += 2 * rawDataBlockCount + 2 + 2 * rawDataBlockCount;
    f.payloadBytes = frameLength - (f.protectionAbsent ? 7 : 9);
    if (f.payloadBytes < 0 || pos_ + (size_t)f.payloadBytes > n_) return false;  // EOFException in Java
    f.payloadOffset = pos_;
    pos_ += f.payloadBytes;
    return true;
  }

 private:
  const uint8_t* d_;
  size_t n_;
  size_t pos_ = 0;
};

}  // namespace jaad
