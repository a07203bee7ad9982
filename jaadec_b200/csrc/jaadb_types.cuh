// Device-side data layout shared by the kernels and the host runtime.
//
// HBM layout (all arrays are engine- or batch-owned, SoA across frames/streams):
//   blob            uint8  [blob_bytes + 16]        compressed frames (caller's blob, padded)
//   frames          FrameDev [n_frames]             per frame: where it is and which stream it belongs to
//   frame_side      FrameSide [n_frames]            K1 out: status word per frame
//   ics_side        IcsSide [n_ics]                 K1 out: window flags, sections, scalefactors per channel-frame
//   q               int16  [n_ics][1024]            K1 out: quantised coefficients (window de-interleaved)
//   runs / run_frames                               per stream: its frames of this batch in decode order
//   overlap         float  [max_streams][8][1024]   persistent IMDCT overlap per channel slot (ICStream.java:47)
//   stream_state    StreamState [max_streams]       persistent window_shape[CURRENT] per channel slot
//   pcm             int16/float                     K2 out
#pragma once
#include <cstdint>

namespace jaadb {

constexpr int kMaxChannels = 8;      // channel slots per stream (7.1)
constexpr int kMaxElements = 5;      // SCE,CPE,CPE,CPE,LFE
constexpr int kMaxSfbEntries = 120;  // ICStream.MAX_SECTIONS

enum ElementType : int { EL_SCE = 0, EL_CPE = 1, EL_CCE = 2, EL_LFE = 3, EL_DSE = 4, EL_PCE = 5, EL_FIL = 6, EL_END = 7 };

// Channel layouts by channel_configuration (SyntacticElements.java:89-127 lists the same orders).
struct LayoutDev {
  uint8_t n_elements;
  uint8_t n_channels;
  uint8_t el_type[kMaxElements];
  uint8_t el_first_ch[kMaxElements];
};

struct FrameDev {
  uint64_t blob_off;
  uint32_t nbytes;
  int32_t stream_slot;
  uint32_t ics_base;      // index of this frame's first channel slot in ics_side / q
  uint8_t sf_index;
  uint8_t layout;         // channel_configuration (1..7)
  uint8_t profile_ok;
  uint8_t flags;
};

struct __align__(8) FrameSide {
  int32_t status;
  uint16_t tags;          // 4-bit instance tag of up to 4 elements... (element i in bits 4i..4i+3)
  uint8_t n_elements;     // SCE / CPE / LFE elements completely decoded
  uint8_t n_started;      // ... whose element_instance_tag was read (a failing element counts here only)
  uint32_t sbr_bit_off[2];  // bit offset of an SBR FIL payload following element 0/1 (0 = none)
  uint32_t sbr_bits[2];
  uint32_t pns_draws;     // values the frame's parse took from the PNS generator (ICStream.java:241-257), also in frames
                          // that failed later on: the generator has moved by then
  uint32_t notes;         // window-shape updates of elements outside the layout that name objects of the stream, for the
                          // pre-pass to resolve (note_dup_shape, k1_parse.cuh): 3 x 9 bits, only in frames that end in an error
};
static_assert(sizeof(FrameSide) == 32, "FrameSide layout");

// Side information of one individual_channel_stream (one channel of one frame).
struct __align__(16) IcsSide {
  uint8_t present;          // spectral data completely decoded
  uint8_t info_decoded;     // ics_info was read (window_shape bookkeeping, ICSInfo.java:90-91)
  uint8_t window_sequence;
  uint8_t window_shape;
  uint8_t max_sfb;
  uint8_t num_groups;
  uint8_t ms_mask;          // CPE only: ms_mask_present
  uint8_t common_window;    // CPE only
  uint8_t group_len[8];
  uint8_t ms_used[16];      // CPE only: bit (g*max_sfb+sfb)
  uint8_t sfb_cb[kMaxSfbEntries];
  uint16_t sf_idx[kMaxSfbEntries];  // SCALEFACTOR_TABLE index, 0xFFFF: scalefactor is 0.0f
  uint8_t tns_present;
  uint8_t has_pns;          // some band uses codebook 13
  uint16_t pns_base;        // generator values the frame's earlier channels took (draw offset of this channel's first noise band)
  uint32_t tns_bit_off;     // position of tns_data (TNS.java:35-61) relative to the frame's aligned word base, like
                            // FrameSide::sbr_bit_off; read again by K2 in JAADB_TNS_ISO mode
};
static_assert(sizeof(IcsSide) == 400, "IcsSide layout");

struct StreamState {
  uint8_t window_shape[kMaxChannels];  // windowShape[CURRENT] of each channel slot
  // JAAD keeps one element object per (type, element_instance_tag) (A/syntax/SyntacticElements.java, Element.java:36-38);
  // the engine keeps the objects of the tags a stream uses first.  A frame that carries another tag addresses objects
  // the engine does not have: it is reported as JAADB_ST_LAYOUT and leaves the stream's state alone (in JAAD it would
  // decode against fresh objects and also leave these alone).
  uint16_t tags;                       // expected instance tag of element i in bits 4i..4i+3
  uint8_t tags_valid;                  // bit i: element i has been seen
  uint8_t pad;
  // PNS generator (ICStream.java:26,247).  JAAD has ONE static generator per JVM; the engine keeps one per stream, seeded like
  // JAAD's, i.e. every stream decodes as it would alone in a fresh JVM.
  uint32_t pns_state;
};
static_assert(sizeof(StreamState) == 16, "StreamState layout");
constexpr uint32_t kPnsSeed = 0x1F2E3D4Cu;

// n steps of the generator s -> 1664525 s + 1013904223 (mod 2^32) at once: f^2(x) = a^2 x + (a + 1) c
__host__ __device__ inline uint32_t pns_jump(uint32_t s, uint32_t n) {
  uint32_t a = 1664525u, c = 1013904223u;
  while (n) {
    if (n & 1u) s = a * s + c;
    c = (a + 1u) * c;
    a = a * a;
    n >>= 1;
  }
  return s;
}

struct RunDev {
  int32_t stream_slot;
  uint32_t first;      // index into run_frames
  uint32_t count;
  uint8_t layout;
  uint8_t sf_index;
  uint8_t mono_dup;    // duplicate the single channel (SyntacticElements.java:244-245)
  uint8_t sbr;         // SBR stream: K2 hands the core PCM (float, 1024 per channel) to K4 instead of packing output
};

// One frame of a run, in decode order.
struct RunFrameDev {
  uint32_t frame;      // index into frames / frame_side / pcm_off
  uint32_t ics_base;   // that frame's first channel slot in ics_side / q
};

// The same frame as K2 sees it: everything that depends on the stream's earlier frames is resolved by the pre-pass
// (k2_prepass_kernel), so the filterbank kernel can start anywhere in a run.
struct __align__(32) K2FrameDev {
  uint32_t frame;
  uint32_t ics_base;
  // [7:0] channel slot c goes through the filterbank   [15:8] windowShape[PREVIOUS] of slot c   [23:16] windowShape[CURRENT]
  // [24] the frame yields PCM   [25] JAAD reached SyntacticElements.process   [26] ISO TNS to apply   [27] noise bands present
  uint32_t flags;
  uint32_t pns_state;  // PNS generator state when the frame's parse starts
  uint64_t pcm_off;    // where the frame's PCM goes (a copy of pcm_off[frame])
  uint32_t next_ics_base;  // ics_base of the run's next frame: K2 asks the TMA unit for that frame's coefficients while this
                           // record is all it holds
  uint32_t pad;
};
static_assert(sizeof(K2FrameDev) == 32, "K2FrameDev layout");
constexpr uint32_t kK2Emit = 1u << 24, kK2Parsed = 1u << 25, kK2Tns = 1u << 26, kK2Pns = 1u << 27;

// A piece of a run for one K2 CTA: frames [first, first + count) of run `run` (positions inside the run).  One segment per
// run when there are plenty of streams; several when there are few (the IMDCT overlap a segment starts from is recomputed
// from the frames before it, see k2_filterbank_kernel).
struct K2SegDev {
  uint32_t run;
  uint32_t first;
  uint32_t count;
};

// Huffman LUT entry (uint32):
//   leaf: [4:0] code length, [7:5] number of sign bits that follow, [8]=0, [31:16] payload
//         payload quads: 4 x 4-bit two's complement; pairs: 2 x 8-bit two's complement; sf book: value
//   link: [4:0] extra index bits, [8]=1, [31:16] sub-table offset (entries, from the LUT base)
constexpr int kHuffFirstBits = 8;
constexpr int kHuffSfFirstBits = 9;

struct TablesDev {
  const uint32_t* huff_lut;      // all books
  uint32_t huff_lut_entries;
  uint32_t book_base[12];        // [0] = scalefactor book, [1..11] spectral
  const float* iq;               // [8191]
  const float* sf;               // [428]
  const int16_t* swb_long;       // [12][53]
  const int16_t* swb_short;      // [12][17]
  const uint8_t* swb_long_count; // [12]
  const uint8_t* swb_short_count;
  const uint8_t* sfb_of_long;    // [12][1024] coefficient -> sfb
  const uint8_t* sfb_of_short;   // [12][128]
  const float* mdct_long;        // [512][2]
  const float* mdct_short;       // [64][2]
  // the same twiddles in the order K2's pre-IFFT gather reads them (bit-reversed, FFT.java:51-61): entry [j * 64 + t] of
  // the long table = mdct_long[bitrev6(t) + 64 * brev3(j)], entry [j * 8 + u] of the short one = mdct_short[brev3(u) + 8 * brev3(j)]
  // -- the 64 threads of a channel then read consecutive 8-byte entries instead of one 32-byte sector each
  const float* mdct_long_gather;   // [8][64][2]
  const float* mdct_short_gather;  // [8][8][2]
  const float* fft512;           // [512][3]
  const float* fft64;            // [64][2]
  const float* win_long[2];      // sine, kbd [1024]
  const float* win_short[2];     // [128]
};

}  // namespace jaadb
