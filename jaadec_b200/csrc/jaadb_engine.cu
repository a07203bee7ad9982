// Host runtime + C ABI of the B200 batched AAC decode engine (include/jaadb200.h).
//
// One engine = one GPU = one CUDA stream.  The stream table lives on the host,
// per-stream persistent decode state (IMDCT overlap, window shapes) lives in HBM.
// A batch indexes the caller's frames by stream, uploads the descriptors once and
// can then be decoded any number of times with everything resident on the device.
//
// There is deliberately no CPU decode path in this file: if CUDA is unavailable
// every entry point fails with JAADB_E_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <chrono>
#include <cstring>
#include <string>
#include <thread>
#include <utility>
#include <vector>

#include "../../include/jaadb200.h"
#include "generated/jaad_tables_host.h"
#include "k1_parse.cuh"
#include "k2_filterbank.cuh"
#include "k3_sbr_parse.cuh"
#include "k4_sbr_process.cuh"

// resident CTAs per SM the one- / two-channel filterbank kernel is compiled for.  Measured on B200 (config 2, round 2
// session D): 5 (96 registers, 12 bytes of spills) 18.9 ms; 4 (128 registers) 19.6 ms; 6 did not fit the shared memory.
// Session AB, after the kernel had lost its register prefetch and the exchange-2 planes had moved 8 floats closer (37 840
// bytes per CTA: six fit with 48 bytes to spare): 6 (80 registers, 0 .. 12 bytes of spills) 18.0 ms against 18.2 ms with 5;
// issue slots 64 %, but `no_instructions` is now the top stall (17.6 %): 24 warps spread over 80 KB of SASS.
#ifndef K2_STEREO_MIN_BLOCKS
#define K2_STEREO_MIN_BLOCKS 6
#endif
// shared-memory carve-out preference of the one- / two-channel filterbank kernel in percent (-1: the driver's choice).  The
// kernel's table look-ups (IQ, windows, twiddles) live in what is left of the 256 KB for L1.
#ifndef K2_CARVEOUT
#define K2_CARVEOUT -1
#endif

namespace T = ::jaad_tables;
using namespace jaadb;

namespace {

#define CUDA_TRY(e, expr)                                                                      \
  do {                                                                                         \
    cudaError_t _err = (expr);                                                                 \
    if (_err != cudaSuccess) {                                                                 \
      (e)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_err));                    \
      return JAADB_E_CUDA;                                                                     \
    }                                                                                          \
  } while (0)

const int kSfFreq[12] = {96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000};

struct StreamHost {
  bool open = false;
  int profile = 0, sf_index = 0, chan_cfg = 0;
  int n_slots = 0;        // coded channels (channel slots)
  int out_channels = 0;   // DecoderConfig.getChannelCount()
  int sample_rate = 0, sample_length = 1024;
  int sbr = 0;
  int sbr_sr_index = 0;   // SBR.sample_rate: index of the OUTPUT rate (A/sbr/SBR.java:102)
  bool sbr_ds = false;    // SBR.isSBRDownSampled: the output rate could not be doubled (A/sbr/SBR.java:100, A/DecoderConfig.java:124-139)
  bool profile_ok = true;
};

// Huffman LUT builder: two-level tables from the {len, code, values} rows (huffman/Codebooks.java).
struct LutBuilder {
  std::vector<uint32_t> lut;
  uint32_t base[12];

  static uint32_t leaf(int len, int nsign, uint32_t payload) { return (uint32_t)len | ((uint32_t)nsign << 5) | (payload << 16); }

  void add_book(int book, const int32_t* rows, int nrows, int width, int first_bits, bool uns) {
    const uint32_t b0 = (uint32_t)lut.size();
    base[book] = b0;
    lut.resize(b0 + (1u << first_bits), 0xFFFFFFFFu);
    struct Sub { uint32_t prefix; int bits; uint32_t off; };
    std::vector<Sub> subs;
    // pass 1: sub-table sizes
    for (int r = 0; r < nrows; ++r) {
      int len = rows[r * width];
      uint32_t code = (uint32_t)rows[r * width + 1];
      if (len > first_bits) {
        uint32_t pre = code >> (len - first_bits);
        auto it = std::find_if(subs.begin(), subs.end(), [&](const Sub& s) { return s.prefix == pre; });
        if (it == subs.end()) subs.push_back({pre, len - first_bits, 0});
        else it->bits = std::max(it->bits, len - first_bits);
      }
    }
    for (auto& s : subs) {
      s.off = (uint32_t)lut.size();
      lut.resize(lut.size() + (1u << s.bits), 0xFFFFFFFFu);
      lut[b0 + s.prefix] = (uint32_t)s.bits | 0x100u | (s.off << 16);
    }
    // pass 2: leaves
    for (int r = 0; r < nrows; ++r) {
      int len = rows[r * width];
      uint32_t code = (uint32_t)rows[r * width + 1];
      uint32_t payload;
      int nsign = 0;
      if (width == 3) {
        payload = (uint32_t)rows[r * width + 2];
      } else if (width == 6) {
        payload = 0;
        for (int j = 0; j < 4; ++j) {
          int v = rows[r * width + 2 + j];
          payload |= ((uint32_t)v & 15u) << (4 * j);
          if (uns && v != 0) ++nsign;
        }
      } else {
        payload = 0;
        for (int j = 0; j < 2; ++j) {
          int v = rows[r * width + 2 + j];
          payload |= ((uint32_t)v & 255u) << (8 * j);
          if (uns && v != 0) ++nsign;
        }
      }
      uint32_t e = leaf(len, nsign, payload);
      if (len <= first_bits) {
        uint32_t lo = code << (first_bits - len);
        for (uint32_t i = 0; i < (1u << (first_bits - len)); ++i) lut[b0 + lo + i] = e;
      } else {
        uint32_t pre = code >> (len - first_bits);
        auto it = std::find_if(subs.begin(), subs.end(), [&](const Sub& s) { return s.prefix == pre; });
        int rem = len - first_bits;
        uint32_t suffix = code & ((1u << rem) - 1u);
        uint32_t lo = suffix << (it->bits - rem);
        for (uint32_t i = 0; i < (1u << (it->bits - rem)); ++i) lut[it->off + lo + i] = e;
      }
    }
  }
};

LayoutDev make_layout(int chan_cfg) {
  LayoutDev l;
  memset(&l, 0, sizeof l);
  auto add = [&](int type) {
    l.el_type[l.n_elements] = (uint8_t)type;
    l.el_first_ch[l.n_elements] = l.n_channels;
    l.n_channels += (type == EL_CPE) ? 2 : 1;
    l.n_elements++;
  };
  switch (chan_cfg) {
    case 1: add(EL_SCE); break;
    case 2: add(EL_CPE); break;
    case 3: add(EL_SCE); add(EL_CPE); break;
    case 4: add(EL_SCE); add(EL_CPE); add(EL_SCE); break;
    case 5: add(EL_SCE); add(EL_CPE); add(EL_CPE); break;
    case 6: add(EL_SCE); add(EL_CPE); add(EL_CPE); add(EL_LFE); break;
    case 7: add(EL_SCE); add(EL_CPE); add(EL_CPE); add(EL_CPE); add(EL_LFE); break;
    default: break;
  }
  return l;
}

constexpr uint32_t kK2SegCtas = 148 * 6;   // K2 CTAs worth aiming for when a batch has few streams (SMs x resident CTAs)

struct FrameIndex {
  struct Group { int nch; uint32_t first_run, n_runs, first_seg, n_segs; bool segmented; };
  std::vector<FrameDev> frames;          // in the caller's order
  std::vector<RunDev> runs;              // grouped by channel-slot count so every K2 launch has one block size
  std::vector<RunFrameDev> run_frames;   // per run, its frames in the caller's order
  std::vector<Group> groups;
  std::vector<K2SegDev> segs;            // K2's CTAs: per group, the runs cut into segments (one per run when streams abound)
  uint32_t n_ics = 0;
  // SBR streams: one parse run per element (K3), one process run per channel (K4)
  std::vector<SbrRunDev> sbr_runs;
  std::vector<K4RunDev> k4_runs;      // plain SBR channels first, then the SBR+PS ones
  uint32_t n_sbr_frames = 0, n_k4_plain = 0, n_ps_frames = 0, k4_max_count = 0;
  uint32_t k4_banks = 0;              // bit 0: some run uses the 64-band synthesis bank, bit 1: some run the down-sampled one
  // when set, frames / run_frames are written here (pinned staging of the one-call path) instead of the vectors
  FrameDev* frames_out = nullptr;
  RunFrameDev* run_frames_out = nullptr;
};

template <typename Tp>
struct DevBuf {
  Tp* p = nullptr;
  size_t cap = 0;  // elements
  cudaError_t ensure(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&p), n * sizeof(Tp));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

}  // namespace

struct jaadb_engine {
  jaadb_options opts;
  cudaStream_t stream = nullptr;
  std::string error;
  std::vector<StreamHost> streams;
  std::vector<int32_t> free_slots;
  std::vector<uint32_t> scratch_count, scratch_run_of, scratch_fill, scratch_size;
  std::vector<jaadb_frame_desc> scratch_frames, scratch_frames_sm;   // jaadb_decode_containers: frame-major table, stream-major scratch
  FrameIndex scratch_ix;
  // device tables
  std::vector<void*> table_allocs;
  TablesDev tables;
  LayoutDev* d_layouts = nullptr;
  LayoutDev layouts[8];
  uint32_t lut_entries = 0;
  // persistent stream state
  float* d_overlap = nullptr;
  StreamState* d_sstate = nullptr;
  // SBR: tables + persistent state (allocated when the first SBR stream is opened)
  SbrTablesDev sbr_tables;
  SbrConstTables sbr_const;
  bool sbr_ready = false;
  std::vector<SbrElemDev> fresh_sbr_elem;   // a freshly constructed SBR object per element (init_sbr)
  std::vector<PsChanDev> fresh_ps;          // a freshly constructed PSImpl
  SbrElemDev* d_sbr_elem = nullptr;   // [max_streams][2]
  SbrChanDev* d_sbr_chan = nullptr;   // [max_streams][kSbrChansPerStream]
  DevBuf<float> d_xg;                 // K4 tile workspace: the Xsbr matrices of one tile of frames (k4_sbr_process.cuh)
  DevBuf<float> d_xps;                // K5 -> K4c: left / right QMF matrices of the tile's parametric-stereo frames
  PsChanDev* d_ps_chan = nullptr;     // [max_streams]
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  // the SBR / PS pipeline runs one half of a batch's runs on `stream` and the other on `k4_stream` (launch_decode)
  static constexpr int kK4MaxParts = 8;
  cudaStream_t k4_stream[kK4MaxParts - 1] = {};
  cudaEvent_t k4_fork = nullptr, k4_join[kK4MaxParts - 1] = {};
  int k4_parts = 2;        // parts the K4 pipeline is cut into (launch_decode); 0: one part per full wave of K4b warps
  int sm_count = 148;
  int k1_lanes_force = -1;   // JAADB_K1_LANES_LOG2 (testing): frames per warp of the parse kernel, log2; -1: by batch size

  // workspace of the one-call path (jaadb_decode): grow-only, so a steady stream of calls allocates nothing.
  // The call is cut into chunks of consecutive frames; chunk k's PCM goes out over PCIe on copy_stream while
  // chunk k+1 is parsed and transformed on `stream` (two PCM buffers, ping-pong).
  struct Workspace {
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t k_done[2] = {nullptr, nullptr}, d2h_done[2] = {nullptr, nullptr};
    DevBuf<uint8_t> blob, pcm[2];
    DevBuf<FrameDev> frames;
    DevBuf<FrameSide> fside;
    DevBuf<IcsSide> iside;
    DevBuf<int16_t> q;
    DevBuf<RunDev> runs;
    DevBuf<RunFrameDev> run_frames;
    DevBuf<K2FrameDev> k2frames;
    DevBuf<K2SegDev> segs;
    DevBuf<float> ovl_stage;
    DevBuf<uint32_t> pcm_bytes;
    DevBuf<uint64_t> pcm_off;
    DevBuf<SbrRunDev> sbr_runs;
    DevBuf<K4RunDev> k4_runs;
    DevBuf<SbrFrameDev> sbr_frames;
    DevBuf<PsFrameDev> ps_frames;
    DevBuf<float> core;
    FrameSide* h_fside = nullptr;      // pinned
    uint64_t* h_off = nullptr;         // pinned: PCM placement of every frame of a call (uploaded while the host goes on)
    size_t h_off_cap = 0;
    uint32_t* h_pcm_bytes = nullptr;   // pinned
    size_t h_cap = 0;
    // pinned descriptor staging, double buffered (chunk k uses slot k & 1)
    cudaEvent_t desc_done[2] = {nullptr, nullptr};
    FrameDev* h_frames[2] = {nullptr, nullptr};
    RunFrameDev* h_run_frames[2] = {nullptr, nullptr};
    RunDev* h_runs[2] = {nullptr, nullptr};
    K2SegDev* h_segs[2] = {nullptr, nullptr};
    SbrRunDev* h_sbr_runs[2] = {nullptr, nullptr};
    K4RunDev* h_k4_runs[2] = {nullptr, nullptr};
    size_t h_chunk_cap = 0, h_runs_cap = 0;
  } ws;

  void set_error(const std::string& s) { error = s; }

  template <typename Tp>
  int upload(const Tp* host, size_t n, const Tp** out) {
    void* d = nullptr;
    CUDA_TRY(this, cudaMalloc(&d, n * sizeof(Tp)));
    table_allocs.push_back(d);
    CUDA_TRY(this, cudaMemcpy(d, host, n * sizeof(Tp), cudaMemcpyHostToDevice));
    *out = reinterpret_cast<const Tp*>(d);
    return 0;
  }
};

struct jaadb_batch {
  jaadb_engine* e = nullptr;
  uint32_t n_frames = 0, n_ics = 0;
  uint64_t blob_bytes = 0, pcm_bytes = 0;
  std::vector<FrameDev> frames;
  std::vector<RunDev> runs;           // grouped by channel count, see run_groups
  std::vector<RunFrameDev> run_frames;
  std::vector<uint64_t> pcm_off;
  std::vector<uint32_t> frame_pcm_size;  // expected size per frame
  typedef FrameIndex::Group Group;
  std::vector<Group> groups;
  // device
  DevBuf<uint8_t> d_blob, d_pcm;
  DevBuf<FrameDev> d_frames;
  DevBuf<FrameSide> d_fside;
  DevBuf<IcsSide> d_iside;
  DevBuf<int16_t> d_q;
  DevBuf<RunDev> d_runs;
  DevBuf<RunFrameDev> d_run_frames;
  DevBuf<K2FrameDev> d_k2frames;
  DevBuf<K2SegDev> d_segs;
  DevBuf<float> d_ovl_stage;
  std::vector<K2SegDev> segs;
  DevBuf<uint32_t> d_pcm_bytes;
  DevBuf<uint64_t> d_pcm_off;
  DevBuf<float> d_spec_tap;
  std::vector<SbrRunDev> sbr_runs;
  std::vector<K4RunDev> k4_runs;
  uint32_t n_sbr_frames = 0, n_k4_plain = 0, n_ps_frames = 0, k4_max_count = 0, k4_banks = 0;
  DevBuf<PsFrameDev> d_ps_frames;
  DevBuf<SbrRunDev> d_sbr_runs;
  DevBuf<K4RunDev> d_k4_runs;
  DevBuf<SbrFrameDev> d_sbr_frames;
  DevBuf<float> d_core;
  std::vector<FrameSide> h_fside;
  std::vector<uint32_t> h_pcm_bytes;
  jaadb_timings timings;
  bool decoded = false;
};

namespace {

int init_tables(jaadb_engine* e) {
  TablesDev& D = e->tables;
  memset(&D, 0, sizeof D);
  LutBuilder lb;
  lb.add_book(0, T::HCB_SF, 121, 3, kHuffSfFirstBits, false);
  static const int32_t* rows[12] = {nullptr, T::HCB1, T::HCB2, T::HCB3, T::HCB4, T::HCB5, T::HCB6,
                                    T::HCB7, T::HCB8, T::HCB9, T::HCB10, T::HCB11};
  static const int nrows[12] = {0, 81, 81, 81, 81, 81, 81, 64, 64, 169, 169, 289};
  static const bool uns[12] = {false, false, false, true, true, false, false, true, true, true, true, true};
  for (int b = 1; b <= 11; ++b) lb.add_book(b, rows[b], nrows[b], b < 5 ? 6 : 4, kHuffFirstBits, uns[b]);
  if (lb.lut.size() > 0xFFFF) { e->set_error("huffman LUT too large"); return JAADB_E_INVALID; }
  for (uint32_t v : lb.lut)
    if (v == 0xFFFFFFFFu) { e->set_error("huffman LUT has holes (codebook not complete)"); return JAADB_E_INVALID; }
  for (int b = 0; b < 12; ++b) D.book_base[b] = lb.base[b];
  D.huff_lut_entries = (uint32_t)lb.lut.size();
  e->lut_entries = D.huff_lut_entries;
  int rc;
  if ((rc = e->upload(lb.lut.data(), lb.lut.size(), &D.huff_lut))) return rc;
  if ((rc = e->upload(JT(IQ_TABLE), T::IQ_TABLE_N, &D.iq))) return rc;
  if ((rc = e->upload(JT(SCALEFACTOR_TABLE), T::SCALEFACTOR_TABLE_N, &D.sf))) return rc;
  if ((rc = e->upload(T::SWB_OFFSET_LONG, T::SWB_OFFSET_LONG_N, &D.swb_long))) return rc;
  if ((rc = e->upload(T::SWB_OFFSET_SHORT, T::SWB_OFFSET_SHORT_N, &D.swb_short))) return rc;
  uint8_t lc[12], sc[12];
  std::vector<uint8_t> sfbl(12 * 1024, 255), sfbs(12 * 128, 255);
  for (int s = 0; s < 12; ++s) {
    lc[s] = (uint8_t)T::SWB_LONG_WINDOW_COUNT[s];
    sc[s] = (uint8_t)T::SWB_SHORT_WINDOW_COUNT[s];
    for (int b = 0; b < lc[s]; ++b)
      for (int k = T::SWB_OFFSET_LONG[s * 53 + b]; k < T::SWB_OFFSET_LONG[s * 53 + b + 1]; ++k) sfbl[s * 1024 + k] = (uint8_t)b;
    for (int b = 0; b < sc[s]; ++b)
      for (int k = T::SWB_OFFSET_SHORT[s * 17 + b]; k < T::SWB_OFFSET_SHORT[s * 17 + b + 1]; ++k) sfbs[s * 128 + k] = (uint8_t)b;
  }
  if ((rc = e->upload(lc, 12, &D.swb_long_count))) return rc;
  if ((rc = e->upload(sc, 12, &D.swb_short_count))) return rc;
  if ((rc = e->upload(sfbl.data(), sfbl.size(), &D.sfb_of_long))) return rc;
  if ((rc = e->upload(sfbs.data(), sfbs.size(), &D.sfb_of_short))) return rc;
  if ((rc = e->upload(JT(MDCT_TABLE_2048), T::MDCT_TABLE_2048_N, &D.mdct_long))) return rc;
  if ((rc = e->upload(JT(MDCT_TABLE_128), T::MDCT_TABLE_128_N, &D.mdct_short))) return rc;
  {
    // the pre-IFFT twiddles in the order the kernel's bit-reversed gather reads them (TablesDev::mdct_long_gather)
    auto brev3 = [](int j) { return ((j & 1) << 2) | (j & 2) | ((j >> 2) & 1); };
    auto brev6 = [](int t) { int r = 0; for (int b = 0; b < 6; ++b) r |= ((t >> b) & 1) << (5 - b); return r; };
    const float* ml = JT(MDCT_TABLE_2048);
    const float* ms = JT(MDCT_TABLE_128);
    std::vector<float> gl(8 * 64 * 2), gs(8 * 8 * 2);
    for (int j = 0; j < 8; ++j) {
      for (int t = 0; t < 64; ++t) {
        const int k = brev6(t) + 64 * brev3(j);
        gl[(j * 64 + t) * 2] = ml[2 * k];
        gl[(j * 64 + t) * 2 + 1] = ml[2 * k + 1];
      }
      for (int u = 0; u < 8; ++u) {
        const int k = brev3(u) + 8 * brev3(j);
        gs[(j * 8 + u) * 2] = ms[2 * k];
        gs[(j * 8 + u) * 2 + 1] = ms[2 * k + 1];
      }
    }
    if ((rc = e->upload(gl.data(), gl.size(), &D.mdct_long_gather))) return rc;
    if ((rc = e->upload(gs.data(), gs.size(), &D.mdct_short_gather))) return rc;
  }
  if ((rc = e->upload(JT(FFT_TABLE_512), T::FFT_TABLE_512_N, &D.fft512))) return rc;
  if ((rc = e->upload(JT(FFT_TABLE_64), T::FFT_TABLE_64_N, &D.fft64))) return rc;
  if ((rc = e->upload(JT(SINE_1024), 1024, &D.win_long[0]))) return rc;
  if ((rc = e->upload(JT(KBD_1024), 1024, &D.win_long[1]))) return rc;
  if ((rc = e->upload(JT(SINE_128), 128, &D.win_short[0]))) return rc;
  if ((rc = e->upload(JT(KBD_128), 128, &D.win_short[1]))) return rc;
  {
    // ISO TNS (JAADB_TNS_ISO): tools/TNSTables.java:10-25 in TNS_TABLES order, SampleFrequency.java:15-26 {long, short}
    float tns[36];
    memcpy(tns, JT(TNS_COEF_0_3), 8 * 4);
    memcpy(tns + 8, JT(TNS_COEF_0_4), 16 * 4);
    memcpy(tns + 24, JT(TNS_COEF_1_3), 4 * 4);
    memcpy(tns + 28, JT(TNS_COEF_1_4), 8 * 4);
    static const uint8_t max_sfb[24] = {31, 9, 31, 9, 34, 10, 40, 14, 42, 14, 51, 14, 46, 14, 46, 14, 42, 14, 42, 14, 42, 14, 39, 14};
    CUDA_TRY(e, cudaMemcpyToSymbol(c_tns_coef, tns, sizeof tns));
    CUDA_TRY(e, cudaMemcpyToSymbol(c_tns_max_sfb, max_sfb, sizeof max_sfb));
  }
  for (int c = 0; c < 8; ++c) e->layouts[c] = make_layout(c);
  const LayoutDev* dl = nullptr;
  if ((rc = e->upload(e->layouts, 8, &dl))) return rc;
  e->d_layouts = const_cast<LayoutDev*>(dl);
  return 0;
}

// SBR tables + persistent state, set up when the first SBR stream is opened.
int init_sbr(jaadb_engine* e) {
  if (e->sbr_ready) return 0;
  SbrTablesDev& D = e->sbr_tables;
  memset(&D, 0, sizeof D);
  int rc;
  static const int16_t* huff[10] = {T::SBR_T_HUFFMAN_ENV_1_5DB, T::SBR_F_HUFFMAN_ENV_1_5DB, T::SBR_T_HUFFMAN_ENV_BAL_1_5DB,
                                    T::SBR_F_HUFFMAN_ENV_BAL_1_5DB, T::SBR_T_HUFFMAN_ENV_3_0DB, T::SBR_F_HUFFMAN_ENV_3_0DB,
                                    T::SBR_T_HUFFMAN_ENV_BAL_3_0DB, T::SBR_F_HUFFMAN_ENV_BAL_3_0DB, T::SBR_T_HUFFMAN_NOISE_3_0DB,
                                    T::SBR_T_HUFFMAN_NOISE_BAL_3_0DB};
  static const int huff_n[10] = {T::SBR_T_HUFFMAN_ENV_1_5DB_N, T::SBR_F_HUFFMAN_ENV_1_5DB_N, T::SBR_T_HUFFMAN_ENV_BAL_1_5DB_N,
                                 T::SBR_F_HUFFMAN_ENV_BAL_1_5DB_N, T::SBR_T_HUFFMAN_ENV_3_0DB_N, T::SBR_F_HUFFMAN_ENV_3_0DB_N,
                                 T::SBR_T_HUFFMAN_ENV_BAL_3_0DB_N, T::SBR_F_HUFFMAN_ENV_BAL_3_0DB_N, T::SBR_T_HUFFMAN_NOISE_3_0DB_N,
                                 T::SBR_T_HUFFMAN_NOISE_BAL_3_0DB_N};
  for (int i = 0; i < 10; ++i)
    if ((rc = e->upload(huff[i], huff_n[i], &D.huff[i]))) return rc;
  // first-eight-bits tables of the Huffman trees (huff_decode): [23:16] code length + [15:0] value for a leaf within eight
  // bits, else bit 31 + the node the bit-serial walk continues from
  auto build_huff_lut = [](const int16_t* tree, int bias, uint32_t* lut) {
    for (uint32_t w = 0; w < 256; ++w) {
      int index = 0, len = 0;
      while (index >= 0 && len < 8) { index = tree[index * 2 + ((w >> (7 - len)) & 1u)]; ++len; }
      lut[w] = index < 0 ? (((uint32_t)len << 16) | (uint32_t)(uint16_t)(int16_t)(index + bias)) : (0x80000000u | (uint32_t)index);
    }
  };
  {
    std::vector<uint32_t> lut(10 * 256);
    for (int i = 0; i < 10; ++i) build_huff_lut(huff[i], 64, lut.data() + 256 * i);
    if ((rc = e->upload(lut.data(), lut.size(), &D.huff_lut))) return rc;
  }
  if ((rc = e->upload(JT(SBR_E_DEQ_TAB), T::SBR_E_DEQ_TAB_N, &D.e_deq))) return rc;
  if ((rc = e->upload(JT(SBR_Q_DIV_TAB), T::SBR_Q_DIV_TAB_N, &D.q_div))) return rc;
  if ((rc = e->upload(JT(SBR_Q_DIV2_TAB), T::SBR_Q_DIV2_TAB_N, &D.q_div2))) return rc;
  if ((rc = e->upload(JT(SBR_Q_DIV_TAB_LEFT), T::SBR_Q_DIV_TAB_LEFT_N, &D.q_div_left))) return rc;
  if ((rc = e->upload(JT(SBR_Q_DIV_TAB_RIGHT), T::SBR_Q_DIV_TAB_RIGHT_N, &D.q_div_right))) return rc;
  if ((rc = e->upload(JT(SBR_Q_DIV2_TAB_LEFT), T::SBR_Q_DIV2_TAB_LEFT_N, &D.q_div2_left))) return rc;
  if ((rc = e->upload(JT(SBR_Q_DIV2_TAB_RIGHT), T::SBR_Q_DIV2_TAB_RIGHT_N, &D.q_div2_right))) return rc;
  if ((rc = e->upload(JT(SBR_E_PAN_TAB), T::SBR_E_PAN_TAB_N, &D.e_pan))) return rc;
  if ((rc = e->upload(JT(SBR_QMF_C), T::SBR_QMF_C_N, &D.qmf_c))) return rc;
  if ((rc = e->upload(reinterpret_cast<const float*>(JAAD_QMF32_PRE_TWIDDLE_BITS), 64, &D.qmf32_tw))) return rc;
  if ((rc = e->upload(JT(SBR_DCT4_64_TAB), T::SBR_DCT4_64_TAB_N, &D.dct4_tab))) return rc;
  if ((rc = e->upload(JT(SBR_W_ARRAY_REAL), 16, &D.w_real))) return rc;
  if ((rc = e->upload(JT(SBR_W_ARRAY_IMAG), 16, &D.w_imag))) return rc;
  if ((rc = e->upload(JT(SBR_NOISE_TABLE), T::SBR_NOISE_TABLE_N, &D.noise_table))) return rc;
  // parametric stereo
  {
    static const int16_t* ph[10] = {T::PS_F_HUFF_IID_DEF, T::PS_T_HUFF_IID_DEF, T::PS_F_HUFF_IID_FINE, T::PS_T_HUFF_IID_FINE,
                                    T::PS_F_HUFF_ICC, T::PS_T_HUFF_ICC, T::PS_F_HUFF_IPD, T::PS_T_HUFF_IPD, T::PS_F_HUFF_OPD, T::PS_T_HUFF_OPD};
    static const int ph_n[10] = {T::PS_F_HUFF_IID_DEF_N, T::PS_T_HUFF_IID_DEF_N, T::PS_F_HUFF_IID_FINE_N, T::PS_T_HUFF_IID_FINE_N,
                                 T::PS_F_HUFF_ICC_N, T::PS_T_HUFF_ICC_N, T::PS_F_HUFF_IPD_N, T::PS_T_HUFF_IPD_N, T::PS_F_HUFF_OPD_N, T::PS_T_HUFF_OPD_N};
    for (int i = 0; i < 10; ++i)
      if ((rc = e->upload(ph[i], ph_n[i], &D.ps_huff[i]))) return rc;
    {
      std::vector<uint32_t> lut(10 * 256);
      for (int i = 0; i < 10; ++i) build_huff_lut(ph[i], 31, lut.data() + 256 * i);
      if ((rc = e->upload(lut.data(), lut.size(), &D.ps_huff_lut))) return rc;
    }
    if ((rc = e->upload(JT(PS_IPDOPD_COS_TAB), 9, &D.ps_ipdopd_cos))) return rc;
    if ((rc = e->upload(JT(PS_IPDOPD_SIN_TAB), 9, &D.ps_ipdopd_sin))) return rc;
    if ((rc = e->upload(JT(PS_FILTER_A), 3, &D.ps_filter_a))) return rc;
    if ((rc = e->upload(JT(PS_PHI_FRACT_QMF), T::PS_PHI_FRACT_QMF_N, &D.ps_phi_qmf))) return rc;
    if ((rc = e->upload(JT(PS_PHI_FRACT_SUBQMF20), T::PS_PHI_FRACT_SUBQMF20_N, &D.ps_phi_sub))) return rc;
    if ((rc = e->upload(JT(PS_Q_FRACT_ALLPASS_QMF), T::PS_Q_FRACT_ALLPASS_QMF_N, &D.ps_q_qmf))) return rc;
    if ((rc = e->upload(JT(PS_Q_FRACT_ALLPASS_SUBQMF20), T::PS_Q_FRACT_ALLPASS_SUBQMF20_N, &D.ps_q_sub))) return rc;
    if ((rc = e->upload(JT(PS_COS_ALPHAS), 8, &D.ps_cos_alphas))) return rc;
    if ((rc = e->upload(JT(PS_SIN_ALPHAS), 8, &D.ps_sin_alphas))) return rc;
    if ((rc = e->upload(JT(PS_COS_BETAS_NORMAL), T::PS_COS_BETAS_NORMAL_N, &D.ps_cos_betas[0]))) return rc;
    if ((rc = e->upload(JT(PS_COS_BETAS_FINE), T::PS_COS_BETAS_FINE_N, &D.ps_cos_betas[1]))) return rc;
    if ((rc = e->upload(JT(PS_SIN_BETAS_NORMAL), T::PS_SIN_BETAS_NORMAL_N, &D.ps_sin_betas[0]))) return rc;
    if ((rc = e->upload(JT(PS_SIN_BETAS_FINE), T::PS_SIN_BETAS_FINE_N, &D.ps_sin_betas[1]))) return rc;
    // IIDMode hands (sin_gammas, cos_gammas) to constructor parameters named (cos_gammas, sin_gammas)
    // (ps/IIDMode.java:16-28 vs ps/IIDTables.java:17-21): what the mixing code calls cos_gammas is the sin table
    if ((rc = e->upload(JT(PS_SIN_GAMMAS_NORMAL), T::PS_SIN_GAMMAS_NORMAL_N, &D.ps_cos_gammas[0]))) return rc;
    if ((rc = e->upload(JT(PS_SIN_GAMMAS_FINE), T::PS_SIN_GAMMAS_FINE_N, &D.ps_cos_gammas[1]))) return rc;
    if ((rc = e->upload(JT(PS_COS_GAMMAS_NORMAL), T::PS_COS_GAMMAS_NORMAL_N, &D.ps_sin_gammas[0]))) return rc;
    if ((rc = e->upload(JT(PS_COS_GAMMAS_FINE), T::PS_COS_GAMMAS_FINE_N, &D.ps_sin_gammas[1]))) return rc;
    if ((rc = e->upload(JT(PS_SINCOS_ALPHAS_B_NORMAL), T::PS_SINCOS_ALPHAS_B_NORMAL_N, &D.ps_sincos_alphas_b[0]))) return rc;
    if ((rc = e->upload(JT(PS_SINCOS_ALPHAS_B_FINE), T::PS_SINCOS_ALPHAS_B_FINE_N, &D.ps_sincos_alphas_b[1]))) return rc;
    if ((rc = e->upload(JT(PS_SF_IID_NORMAL), T::PS_SF_IID_NORMAL_N, &D.ps_sf_iid[0]))) return rc;
    if ((rc = e->upload(JT(PS_SF_IID_FINE), T::PS_SF_IID_FINE_N, &D.ps_sf_iid[1]))) return rc;
    if ((rc = e->upload(JT(PS_P8_13_20), 7, &D.ps_p8))) return rc;
    if ((rc = e->upload(JT(PS_P2_13_20), 7, &D.ps_p2))) return rc;
  }
  // FBT.find_bands / find_initial_power (sbr/FBT.java:135-145) for every argument the band-table code can pass, evaluated
  // here on the host in double precision exactly as the Java expressions are
  {
    std::vector<uint8_t> fb((size_t)2 * 7 * 65 * 65, 0);
    for (int warp = 0; warp < 2; ++warp)
      for (int bands = 0; bands <= 6; ++bands)
        for (int a0 = 1; a0 <= 64; ++a0)
          for (int a1 = 1; a1 <= 64; ++a1) {
            float div = (float)std::log(2.0);
            if (warp != 0) div *= 1.3f;
            int v = (int)(bands * std::log((double)((float)a1 / (float)a0)) / div + 0.5);
            fb[(((size_t)warp * 7 + bands) * 65 + a0) * 65 + a1] = (uint8_t)std::min(std::max(v, 0), 255);
          }
    if ((rc = e->upload(fb.data(), fb.size(), &D.find_bands))) return rc;
    std::vector<float> ip((size_t)64 * 65 * 65, 1.0f);
    for (int bands = 1; bands < 64; ++bands)
      for (int a0 = 1; a0 <= 64; ++a0)
        for (int a1 = 1; a1 <= 64; ++a1)
          ip[((size_t)bands * 65 + a0) * 65 + a1] = (float)std::pow((double)((float)a1 / (float)a0), (double)(1.0f / (float)bands));
    if ((rc = e->upload(ip.data(), ip.size(), &D.init_power))) return rc;
  }
  SbrConstTables& K = e->sbr_const;
  if ((rc = e->upload(T::SBR_START_MIN_TABLE, 12, &K.start_min))) return rc;
  if ((rc = e->upload(T::SBR_OFFSET_INDEX_TABLE, 12, &K.offset_index))) return rc;
  if ((rc = e->upload(T::SBR_OFFSET, T::SBR_OFFSET_N, &K.offset))) return rc;
  if ((rc = e->upload(T::SBR_STOP_MIN_TABLE, 12, &K.stop_min))) return rc;
  if ((rc = e->upload(T::SBR_STOP_OFFSET_TABLE, T::SBR_STOP_OFFSET_TABLE_N, &K.stop_offset))) return rc;
  if ((rc = e->upload(T::SBR_GOAL_SB_TAB, 12, &K.goal_sb))) return rc;
  if ((rc = e->upload(JT(SBR_LIMITER_BANDS_COMPARE), 3, &K.limiter_cmp))) return rc;
  CUDA_TRY(e, cudaMemcpyToSymbol(c_sbr_dct4, JT(SBR_DCT4_64_TAB), sizeof(float) * 192));
  CUDA_TRY(e, cudaMemcpyToSymbol(c_sbr_w_real, JT(SBR_W_ARRAY_REAL), sizeof(float) * 16));
  CUDA_TRY(e, cudaMemcpyToSymbol(c_sbr_w_imag, JT(SBR_W_ARRAY_IMAG), sizeof(float) * 16));
  CUDA_TRY(e, cudaMemcpyToSymbol(c_sbr_qmf_c, JT(SBR_QMF_C), sizeof(float) * 640));
  CUDA_TRY(e, cudaMemcpyToSymbol(c_sbr_bit_rev, T::SBR_BIT_REV_TAB, sizeof(int) * 32));
  const size_t ns = e->streams.size();
  CUDA_TRY(e, cudaMalloc(reinterpret_cast<void**>(&e->d_sbr_elem), sizeof(SbrElemDev) * ns * 2));
  CUDA_TRY(e, cudaMalloc(reinterpret_cast<void**>(&e->d_sbr_chan), sizeof(SbrChanDev) * ns * kSbrChansPerStream));
  CUDA_TRY(e, cudaMalloc(reinterpret_cast<void**>(&e->d_ps_chan), sizeof(PsChanDev) * ns));
  cudaFuncSetAttribute(k3_sbr_parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k3_smem_bytes());
#ifdef K3_CARVEOUT
  cudaFuncSetAttribute(k3_sbr_parse_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, K3_CARVEOUT);
#endif
  cudaFuncSetAttribute(k4a_analysis_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k4a_smem_bytes());
  cudaFuncSetAttribute(k4b_hf_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k4b_smem_bytes());
  cudaFuncSetAttribute(k4b_hf_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);   // two CTAs of ten warps, 110 KB each
#define K4C_ATTR(FMT)                                                                                                        \
  cudaFuncSetAttribute(k4c_synthesis_kernel<FMT, false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k4c_smem_bytes()); \
  cudaFuncSetAttribute(k4c_synthesis_kernel<FMT, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k4c_smem_bytes()); \
  cudaFuncSetAttribute(k4c_synthesis_kernel<FMT, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k4c_smem_bytes()); \
  cudaFuncSetAttribute(k4c_synthesis_kernel<FMT, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k4c_smem_bytes())
  K4C_ATTR(0); K4C_ATTR(1); K4C_ATTR(2);
#undef K4C_ATTR
  cudaFuncSetAttribute(k5_ps_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k5_smem_bytes());
  // a fresh SBR object per element: everything zero except Channel.prevEnvIsShort = -1 (Channel.java:51) and the null modes
  e->fresh_sbr_elem.resize(2);
  memset(e->fresh_sbr_elem.data(), 0, sizeof(SbrElemDev) * 2);
  for (auto& el : e->fresh_sbr_elem) {
    el.ch[0].prevEnvIsShort = -1; el.ch[1].prevEnvIsShort = -1;
    el.ps.iid.mode = -1; el.ps.icc.mode = -1; el.ps.ipd.mode = -1; el.ps.opd.mode = -1;
  }
  // a fresh PSImpl (ps/PSImpl.java:64-94): everything zero except h11_prev = 1 and -- the constructor sets
  // h12_prev[i][1] where it means h12_prev[i][0] (SURVEY A-14) -- the imaginary part of h12_prev = 1
  e->fresh_ps.resize(1);
  memset(e->fresh_ps.data(), 0, sizeof(PsChanDev));
  for (auto& h : e->fresh_ps[0].h_prev) { h[0] = 1.0f; h[5] = 1.0f; }
  e->sbr_ready = true;
  return 0;
}

// Minimal MSB-first reader for the AudioSpecificConfig (host side only).
struct HostBits {
  const uint8_t* d; uint32_t n, pos = 0;
  bool ok = true;
  uint32_t left() const { return 8 * n - pos; }
  uint32_t read(int k) {
    if (left() < (uint32_t)k) { ok = false; pos = 8 * n; return 0; }
    uint32_t v = 0;
    for (int i = 0; i < k; ++i, ++pos) v = (v << 1) | ((d[pos >> 3] >> (7 - (pos & 7))) & 1u);
    return v;
  }
};

bool profile_supported(int aot) {  // Profile.isDecodingSupported (Profile.java:5-21)
  switch (aot) { case 1: case 2: case 4: case 5: case 17: case 19: case 29: return true; default: return false; }
}

int profile_for_int(int i) {  // Profile.forInt (Profile.java:23-57); -1 = UNKNOWN
  static const int ALL[30] = {1, 2, 3, 4, 5, 6, 7, -1, -1, -1, 11, -1, -1, -1, -1, -1, 17, 18,
                              19, 20, 21, 22, 23, -1, -1, -1, -1, -1, 29, -1};
  return (i >= 1 && i <= 30) ? ALL[i - 1] : -1;
}

int nominal_index(int freq) {  // SampleFrequency.nominalFrequency (SampleFrequency.java:68-93)
  int result = -1;
  float dev = INFINITY;
  for (int i = 0; i < 12; ++i) {
    float d = ((float)freq - (float)kSfFreq[i]) / (float)kSfFreq[i];
    if (d == 0) return i;
    if (d < dev) { result = i; dev = d; }
    if (kSfFreq[i] < freq) break;
  }
  return result;
}

int finish_open(jaadb_engine* e, StreamHost& s, int32_t* stream_id) {
  if (s.chan_cfg < 1 || s.chan_cfg > 7) { e->set_error("unsupported channel configuration"); return JAADB_E_CONFIG; }
  if (s.sf_index < 0 || s.sf_index > 11) { e->set_error("unsupported sampling frequency index"); return JAADB_E_CONFIG; }
  const LayoutDev& l = e->layouts[s.chan_cfg];
  s.n_slots = l.n_channels;
  s.out_channels = (s.chan_cfg == 1) ? 2 : l.n_channels;  // DecoderConfig.getChannelCount (DecoderConfig.java:108-115)
  s.profile_ok = profile_supported(s.profile);
  if (s.sbr > 1 && s.chan_cfg != 1) { e->set_error("parametric stereo needs a mono core"); return JAADB_E_CONFIG; }
  if (s.sbr) {
    if (s.chan_cfg > 2) { e->set_error("SBR is implemented for mono and stereo streams only"); return JAADB_E_CONFIG; }
    s.sbr_ds = s.sample_length != 2048;
    if (s.sbr_ds) s.sbr_sr_index = s.sf_index;
    if (s.sbr_sr_index < 0 || s.sbr_sr_index > 11) { e->set_error("SBR core sampling rate too high"); return JAADB_E_CONFIG; }
    int rc = init_sbr(e);
    if (rc) return rc;
    s.out_channels = 2;   // SCE: SBR1.process fills a second channel (PS or a copy), CPE: two channels
  }
  if (e->free_slots.empty()) { e->set_error("stream table full"); return JAADB_E_CAPACITY; }
  const int32_t slot = e->free_slots.back();
  // fresh decoder state for the slot; the slot is only taken once all of it is in place
  cudaError_t ce = cudaSuccess;
  auto chk = [&](cudaError_t x) { if (ce == cudaSuccess) ce = x; };
  if (s.sbr) {
    // engine-owned templates (built once in init_sbr, read-only afterwards: engines on different threads share nothing)
    chk(cudaMemcpyAsync(e->d_sbr_elem + (size_t)slot * 2, e->fresh_sbr_elem.data(), sizeof(SbrElemDev) * 2, cudaMemcpyHostToDevice, e->stream));
    chk(cudaMemsetAsync(e->d_sbr_chan + (size_t)slot * kSbrChansPerStream, 0, sizeof(SbrChanDev) * kSbrChansPerStream, e->stream));
    if (s.sbr > 1) chk(cudaMemcpyAsync(e->d_ps_chan + slot, e->fresh_ps.data(), sizeof(PsChanDev), cudaMemcpyHostToDevice, e->stream));
  }
  chk(cudaMemsetAsync(e->d_overlap + (size_t)slot * kMaxChannels * 1024, 0, sizeof(float) * kMaxChannels * 1024, e->stream));
  StreamState fresh_state;
  memset(&fresh_state, 0, sizeof fresh_state);
  fresh_state.pns_state = kPnsSeed;   // ICStream.randomState (ICStream.java:26), one generator per stream
  chk(cudaMemcpyAsync(e->d_sstate + slot, &fresh_state, sizeof fresh_state, cudaMemcpyHostToDevice, e->stream));
  chk(cudaStreamSynchronize(e->stream));
  if (ce != cudaSuccess) {
    e->set_error(std::string("stream open: ") + cudaGetErrorString(ce));
    return JAADB_E_CUDA;   // the slot stays in the free list
  }
  e->free_slots.pop_back();
  s.open = true;
  e->streams[slot] = s;
  *stream_id = slot;
  return JAADB_OK;
}

uint32_t frame_pcm_bytes(const jaadb_engine* e, const StreamHost& s) {
  const uint32_t per = (e->opts.pcm_format == JAADB_PCM_F32_PLANAR) ? 4u : 2u;
  return (uint32_t)s.out_channels * (uint32_t)s.sample_length * per;
}


// ---- host-side indexing shared by the staged and the one-call paths -----------------------------------------

// PCM placement: the caller's offsets, or frames packed back to back in array order.
int layout_pcm(jaadb_engine* e, const jaadb_frame_desc* fd, uint32_t n, const uint64_t* pcm_offsets,
               std::vector<uint64_t>& off, std::vector<uint32_t>& size, uint64_t* total) {
  off.resize(n);
  size.resize(n);
  uint64_t pos = 0, end = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const int32_t sid = fd[i].stream_id;
    if (sid < 0 || sid >= (int32_t)e->streams.size() || !e->streams[sid].open) {
      e->set_error("frame refers to an unknown stream");
      return JAADB_E_NOSTREAM;
    }
    const uint32_t sz = frame_pcm_bytes(e, e->streams[sid]);
    size[i] = sz;
    if (pcm_offsets) {
      if (pcm_offsets[i] & 3u) { e->set_error("pcm offsets must be 4-byte aligned"); return JAADB_E_INVALID; }
      if (pcm_offsets[i] > UINT64_MAX - sz) { e->set_error("pcm offset out of range"); return JAADB_E_INVALID; }
      off[i] = pcm_offsets[i];
    } else {
      off[i] = pos;
      pos += sz;
    }
    end = std::max(end, off[i] + sz);
  }
  *total = end;
  return JAADB_OK;
}

// Frames [0, n) of fd -> device descriptors + per-stream runs.  Frame indices in the result are relative to fd.
int index_frames(jaadb_engine* e, const jaadb_frame_desc* fd, uint32_t n, uint64_t blob_bytes, FrameIndex& ix) {
  if (!ix.frames_out) ix.frames.resize(n);
  FrameDev* const fout = ix.frames_out ? ix.frames_out : ix.frames.data();
  ix.runs.clear();
  ix.groups.clear();
  ix.sbr_runs.clear();
  ix.k4_runs.clear();
  ix.n_sbr_frames = 0;
  ix.n_k4_plain = 0;
  ix.k4_max_count = 0;
  ix.k4_banks = 0;
  ix.n_ps_frames = 0;
  // per-stream frame counts (counting sort keeps array order inside each stream)
  std::vector<uint32_t>& count = e->scratch_count;
  count.assign(e->streams.size(), 0);
  uint32_t ics = 0;
  for (uint32_t i = 0; i < n; ++i) {
    const jaadb_frame_desc& d = fd[i];
    if (d.stream_id < 0 || d.stream_id >= (int32_t)e->streams.size() || !e->streams[d.stream_id].open) {
      e->set_error("frame refers to an unknown stream");
      return JAADB_E_NOSTREAM;
    }
    // (overflow-safe: an offset close to 2^64 must not wrap past the test; 2^29 bytes keeps the bit positions in 32 bits)
    if (d.nbytes > blob_bytes || d.offset > blob_bytes - d.nbytes || d.nbytes >= (1u << 29)) {
      e->set_error("frame exceeds the blob");
      return JAADB_E_INVALID;
    }
    const StreamHost& s = e->streams[d.stream_id];
    FrameDev& f = fout[i];
    f.blob_off = d.offset;
    f.nbytes = d.nbytes;
    f.stream_slot = d.stream_id;
    f.ics_base = ics;
    f.sf_index = (uint8_t)s.sf_index;
    f.layout = (uint8_t)s.chan_cfg;
    f.profile_ok = s.profile_ok ? 1 : 0;
    f.flags = 0;
    ics += (uint32_t)s.n_slots;
    count[d.stream_id]++;
  }
  ix.n_ics = ics;
  std::vector<uint32_t>& run_of = e->scratch_run_of;
  run_of.assign(e->streams.size(), 0xFFFFFFFFu);
  for (int nch = 1; nch <= kMaxChannels; ++nch) {
    FrameIndex::Group g{nch, (uint32_t)ix.runs.size(), 0, 0, 0, false};
    for (size_t s = 0; s < e->streams.size(); ++s) {
      if (!count[s] || e->streams[s].n_slots != nch) continue;
      RunDev r;
      memset(&r, 0, sizeof r);
      r.stream_slot = (int32_t)s;
      r.count = count[s];
      r.layout = (uint8_t)e->streams[s].chan_cfg;
      r.sf_index = (uint8_t)e->streams[s].sf_index;
      r.mono_dup = (e->streams[s].chan_cfg == 1) ? 1 : 0;
      r.sbr = e->streams[s].sbr ? 1 : 0;
      run_of[s] = (uint32_t)ix.runs.size();
      ix.runs.push_back(r);
      g.n_runs++;
    }
    if (g.n_runs) ix.groups.push_back(g);
  }
  uint32_t acc = 0;
  for (auto& r : ix.runs) { r.first = acc; acc += r.count; }
  // K2's CTAs.  With plenty of streams one CTA walks a whole run; with few, runs are cut into segments of G frames so that
  // the GPU fills up (each segment re-runs the frame before it for the overlap it starts from: 1/G extra work).
  ix.segs.clear();
  for (auto& g : ix.groups) {
    uint64_t total = 0;
    uint32_t longest = 0;
    for (uint32_t r = g.first_run; r < g.first_run + g.n_runs; ++r) { total += ix.runs[r].count; longest = std::max(longest, ix.runs[r].count); }
    uint32_t G = 0;
    if (e->opts.k2_segment_frames) G = e->opts.k2_segment_frames;
    else if (g.n_runs < 2 * kK2SegCtas) G = (uint32_t)std::max<uint64_t>(2, (total + 2 * kK2SegCtas - 1) / (2 * kK2SegCtas));   // two waves of segments
    if (G >= longest) G = 0;
    g.first_seg = (uint32_t)ix.segs.size();
    g.segmented = G != 0;
    for (uint32_t r = g.first_run; r < g.first_run + g.n_runs; ++r) {
      const uint32_t cnt = ix.runs[r].count, step = G ? G : cnt;
      for (uint32_t f0 = 0; f0 < cnt; f0 += step) ix.segs.push_back(K2SegDev{r, f0, std::min(step, cnt - f0)});
    }
    g.n_segs = (uint32_t)ix.segs.size() - g.first_seg;
  }
  for (int pass = 0; pass < 2; ++pass) {   // plain SBR first, SBR+PS second: K4 launches them as two grids
    for (const auto& r : ix.runs) {
      if (!r.sbr) continue;
      const StreamHost& sh = e->streams[r.stream_slot];
      const bool with_ps = sh.sbr > 1;
      if (with_ps != (pass == 1)) continue;
      const bool stereo = sh.chan_cfg == 2;
      SbrRunDev sr;
      memset(&sr, 0, sizeof sr);
      sr.stream_slot = r.stream_slot;
      sr.first = r.first;
      sr.count = r.count;
      sr.sbr_base = ix.n_sbr_frames;
      sr.element = 0;
      sr.stereo = stereo ? 1 : 0;
      sr.sr_index = (uint8_t)sh.sbr_sr_index;
      sr.first_ch = 0;
      sr.ps = with_ps ? 1 : 0;
      sr.ps_base = ix.n_ps_frames;
      ix.sbr_runs.push_back(sr);
      for (int c = 0; c < (stereo ? 2 : 1); ++c) {
        K4RunDev kr;
        memset(&kr, 0, sizeof kr);
        kr.stream_slot = r.stream_slot;
        kr.first = r.first;
        kr.count = r.count;
        kr.sbr_base = sr.sbr_base;
        kr.chan = (uint8_t)c;
        kr.ch_slot = (uint8_t)c;
        kr.out_ch = (uint8_t)c;
        kr.n_out = 2;
        kr.dup = stereo ? 0 : 1;
        kr.ds = sh.sbr_ds ? 1 : 0;
        ix.k4_banks |= sh.sbr_ds ? 2u : 1u;
        kr.ps_base = sr.ps_base;
        ix.k4_runs.push_back(kr);
        ix.k4_max_count = std::max(ix.k4_max_count, r.count);
      }
      ix.n_sbr_frames += r.count;
      if (with_ps) ix.n_ps_frames += r.count;
    }
    if (pass == 0) ix.n_k4_plain = (uint32_t)ix.k4_runs.size();
  }
  if (!ix.run_frames_out) ix.run_frames.resize(n);
  RunFrameDev* const rout = ix.run_frames_out ? ix.run_frames_out : ix.run_frames.data();
  std::vector<uint32_t>& fill = e->scratch_fill;
  fill.assign(ix.runs.size(), 0);
  for (uint32_t i = 0; i < n; ++i) {
    const uint32_t r = run_of[fd[i].stream_id];
    rout[ix.runs[r].first + fill[r]++] = RunFrameDev{i, fout[i].ics_base};
  }
  return JAADB_OK;
}

// Device buffers of one decode pass.
struct DecodeBufs {
  const uint8_t* blob;
  const FrameDev* frames;
  FrameSide* fside;
  IcsSide* iside;
  int16_t* q;
  const RunDev* runs;
  uint32_t n_runs;
  const RunFrameDev* run_frames;
  K2FrameDev* k2frames;
  const K2SegDev* segs;
  float* ovl_stage;
  uint8_t* pcm;
  const uint64_t* pcm_off;
  uint32_t* pcm_bytes;
  float* tap;
  // SBR
  const SbrRunDev* sbr_runs;
  const K4RunDev* k4_runs;
  SbrFrameDev* sbr_frames;
  float* core;
  PsFrameDev* ps_frames;
  uint32_t n_k4_plain;
  uint32_t k4_max_count;   // longest plain-SBR run
  uint32_t k4_banks;       // FrameIndex::k4_banks
};

// Streams of three to six channels (up to 5.1: 384 threads, 107 KB of shared memory): two CTAs per SM need <= 80 registers
#ifndef K2_MC_MIN_BLOCKS
#define K2_MC_MIN_BLOCKS 2
#endif
#ifndef K4_TILE_BYTES
#define K4_TILE_BYTES (16384ull << 20)   // measured: 1.5 GB -> 8 GB took the SBR stages of configs 3 / 4 from 75 / 199 ms to 72 / 176 ms, 8 -> 16 GB 65.4 / 154.0 -> 64.9 / 152.2 ms
#endif
constexpr uint64_t kK4TileBytes = K4_TILE_BYTES;   // upper bound of the K4 tile workspace (Xsbr matrices of one tile)

// K1 (+ K3) + K2 (+ K4) over an indexed set of frames, everything already on the device.
cudaError_t launch_decode(jaadb_engine* e, const FrameIndex::Group* groups, size_t n_groups, uint32_t n_frames, uint32_t n_sbr_runs,
                   uint32_t n_k4_runs, const DecodeBufs& B, cudaEvent_t after_k1, cudaEvent_t after_k2, uint32_t* launches) {
  {
    const int threads = kK1Threads;
    // four CTAs of eight warps per SM are resident (64 registers, 56 KB of shared memory each)
    const uint32_t lanes_log2 = e->k1_lanes_force >= 0 ? (uint32_t)e->k1_lanes_force
                                                       : k1_lanes_log2(n_frames, (uint32_t)e->sm_count * 4u * (uint32_t)(kK1Threads / 32));
    const uint32_t per_block = (uint32_t)(kK1Threads / 32) << lanes_log2;
    const int blocks = (int)((n_frames + per_block - 1) / per_block);
    k1_parse_kernel<<<blocks, threads, k1_smem_bytes(e->lut_entries), e->stream>>>(B.blob, B.frames, n_frames, B.fside, B.iside, B.q,
                                                                                   e->tables, e->d_layouts,
                                                                                   (e->opts.flags & JAADB_FLAG_PULSE_ISO) ? 1 : 0,
                                                                                   lanes_log2);
    ++*launches;
  }
  if (n_sbr_runs) {
    // SBR payload parse before the filterbank: an exception inside SBR.decode fails the whole frame
    const int blocks = (int)((n_sbr_runs + kK3WarpsPerBlock - 1) / kK3WarpsPerBlock);
    k3_sbr_parse_kernel<<<blocks, 32 * kK3WarpsPerBlock, k3_smem_bytes(), e->stream>>>(
        B.blob, B.frames, B.fside, B.sbr_runs, n_sbr_runs, B.run_frames, e->d_sbr_elem, B.sbr_frames, B.ps_frames, e->sbr_tables, e->sbr_const);
    ++*launches;
  }
  if (after_k1) cudaEventRecord(after_k1, e->stream);
  {
    // everything that depends on a stream's earlier frames, once per run
    const int bps = e->opts.pcm_format == JAADB_PCM_F32_PLANAR ? 4 : 2;
    k2_prepass_kernel<<<(B.n_runs + kK2PreWarps - 1) / kK2PreWarps, 32 * kK2PreWarps, 0, e->stream>>>(B.runs, B.n_runs, B.run_frames, B.fside, B.iside, e->d_sstate,
                                                                     e->d_layouts, B.pcm_off, B.k2frames, B.pcm_bytes, bps,
                                                                     e->opts.tns_mode == JAADB_TNS_ISO ? 1 : 0);
    ++*launches;
  }
  for (size_t gi = 0; gi < n_groups; ++gi) {
    const FrameIndex::Group& g = groups[gi];
    const int threads = g.nch * kThreadsPerChannel;
    const int out_ch = (g.nch == 1) ? 2 : g.nch;
    const size_t smem = k2_smem_bytes(g.nch, out_ch, threads <= 128);
    K2Args A;
    A.segs = B.segs + g.first_seg;
    A.runs = B.runs;
    A.k2frames = B.k2frames;
    A.frames = B.frames;
    A.blob = B.blob;
    A.iside = B.iside;
    A.qall = B.q;
    A.overlap_all = e->d_overlap;
    A.overlap_stage = g.segmented ? B.ovl_stage : nullptr;
    A.pcm = B.pcm;
    A.spec_tap = B.tap;
    A.core = B.core;
    A.layouts = e->d_layouts;
    A.nch = g.nch;
#define LAUNCH_K2(FMT)                                                                                              \
  do {                                                                                                              \
    if (threads <= 128) k2_filterbank_kernel<FMT, 128, K2_STEREO_MIN_BLOCKS><<<g.n_segs, threads, smem, e->stream>>>(A, e->tables); \
    else if (threads <= 384 && K2_MC_MIN_BLOCKS > 1) k2_filterbank_kernel<FMT, 384, K2_MC_MIN_BLOCKS><<<g.n_segs, threads, smem, e->stream>>>(A, e->tables); \
    else k2_filterbank_kernel<FMT, 512, 1><<<g.n_segs, threads, smem, e->stream>>>(A, e->tables);                   \
  } while (0)
    if (e->opts.pcm_format == JAADB_PCM_S16LE) LAUNCH_K2(0);
    else if (e->opts.pcm_format == JAADB_PCM_S16BE) LAUNCH_K2(1);
    else LAUNCH_K2(2);
#undef LAUNCH_K2
    ++*launches;
    if (g.segmented) {
      k2_commit_kernel<<<g.n_runs, 256, 0, e->stream>>>(B.runs + g.first_run, g.n_runs, g.nch, B.ovl_stage + (size_t)g.first_run * kMaxChannels * 1024,
                                                        e->d_overlap);
      ++*launches;
    }
  }
  if (after_k2) cudaEventRecord(after_k2, e->stream);
  if (n_k4_runs) {
    const uint32_t n_plain = B.n_k4_plain, n_ps = n_k4_runs - B.n_k4_plain;
    // frame-parallel pipeline over tiles of ft frames per run; the tile's matrices stay within kK4TileBytes
    const uint64_t per_frame = ((uint64_t)n_k4_runs * 32 + (uint64_t)n_ps * 64) * kXgRow * sizeof(float);
    uint32_t ft = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(B.k4_max_count, kK4TileBytes / per_frame));
    if (ft > kK4cG) ft -= ft % kK4cG;   // the synthesis kernel takes kK4cG frames per CTA
    if (e->opts.sbr_tile_frames) ft = std::min(ft, e->opts.sbr_tile_frames);
    const uint32_t rows = 8 + 32 * ft;
    cudaError_t err = e->d_xg.ensure((size_t)n_k4_runs * rows * kXgRow);
    if (err != cudaSuccess) return err;
    if (n_ps) {
      err = e->d_xps.ensure((size_t)n_ps * ft * 64 * kXgRow);
      if (err != cudaSuccess) return err;
    }
    // The runs are independent, and the pipeline mixes a latency-bound kernel (K4b: one warp per channel walking the
    // tile's frames, a quarter of the SM's warp slots, whole waves of equal length) with throughput kernels (K4a, K4c).
    // So the runs go down the pipeline in two halves on two CUDA streams: one half's K4b shares the SMs with the other
    // half's analysis / synthesis, and the part-filled last wave of a K4b launch is filled from the other stream.
    struct Part { uint32_t p0, pn, q0, qn; cudaStream_t st; };   // plain runs [p0, p0 + pn), SBR+PS runs [q0, q0 + qn)
    Part parts[jaadb_engine::kK4MaxParts];
    // (0 = one part per full wave of K4b warps: 8192 channel runs are 2.77 waves of 2960, i.e. three parts.  Measured after
    // K4b went to five CTAs per SM: 66.9 / 156.3 ms for configs 3 / 4 against 65.8 / 154.9 ms with two parts, so two stays.
    // Also measured and dropped: sending K3 and K2 down in the same parts, so that part p + 1 is parsed underneath part p's
    // QMF pipeline -- K3 starves next to the K4 kernels and the last part's pipeline waits for it: 105.7 / 195.1 ms with two
    // parts, worse with more.)
    int want = e->k4_parts;
    if (want <= 0) {
      const uint32_t wave = (uint32_t)e->sm_count * K4B_MIN_BLOCKS * kK4bWarps;
      want = (int)std::min<uint32_t>(std::max<uint32_t>((n_k4_runs + wave / 2) / wave, 1u), (uint32_t)jaadb_engine::kK4MaxParts);
    }
    const int n_parts = (n_k4_runs >= 8u * (uint32_t)want) ? want : 1;
    for (int i = 0; i < n_parts; ++i) {
      // (cut between elements: a CPE's two channel runs are neighbours)
      const uint32_t pa = (uint32_t)((uint64_t)n_plain * i / n_parts) & ~1u, pb = (i + 1 == n_parts) ? n_plain : ((uint32_t)((uint64_t)n_plain * (i + 1) / n_parts) & ~1u);
      const uint32_t qa = (uint32_t)((uint64_t)n_ps * i / n_parts), qb = (uint32_t)((uint64_t)n_ps * (i + 1) / n_parts);
      parts[i] = {pa, pb - pa, n_plain + qa, qb - qa, i == 0 ? e->stream : e->k4_stream[i - 1]};
    }
    if (n_parts > 1) {
      if ((err = cudaEventRecord(e->k4_fork, e->stream)) != cudaSuccess) return err;
      for (int i = 1; i < n_parts; ++i)
        if ((err = cudaStreamWaitEvent(parts[i].st, e->k4_fork, 0)) != cudaSuccess) return err;
    }
    const size_t xg_run = (size_t)rows * kXgRow, xps_run = (size_t)ft * 64 * kXgRow;
    for (uint32_t lo = 0; lo < B.k4_max_count; lo += ft) {
      const K4Tile tile{lo, ft, rows};
      const uint32_t n_groups = (ft + kK4cG - 1) / kK4cG;
      for (int pi = 0; pi < n_parts; ++pi) {
        const Part& P = parts[pi];
        const cudaStream_t st = P.st;
        // analysis + HF generation / adjustment over a contiguous range of runs (sub-ranges see their own slice of xg)
        auto front = [&](uint32_t r0, uint32_t n) {
          if (!n) return;
          const uint32_t n_cf = n * ft;
          k4a_analysis_kernel<<<(n_cf + kK4aWarps - 1) / kK4aWarps, 32 * kK4aWarps, k4a_smem_bytes(), st>>>(
              B.k4_runs + r0, n, B.run_frames, B.sbr_frames, B.core, e->d_sbr_chan, e->d_xg.p + r0 * xg_run, tile);
          k4b_hf_kernel<<<(n + kK4bWarps - 1) / kK4bWarps, 32 * kK4bWarps, k4b_smem_bytes(), st>>>(
              B.k4_runs + r0, n, B.run_frames, B.sbr_frames, B.core, e->d_sbr_chan, e->d_xg.p + r0 * xg_run, e->sbr_tables, tile);
          *launches += 2;
        };
        front(P.p0, P.pn);
        front(P.q0, P.qn);
        float* const xps_part = P.qn ? e->d_xps.p + (size_t)(P.q0 - n_plain) * xps_run : nullptr;
        if (P.qn) {
          k5_ps_kernel<<<P.qn, kK5Threads, k5_smem_bytes(), st>>>(B.k4_runs, P.q0, B.sbr_frames, B.ps_frames, e->d_ps_chan,
                                                                   e->d_xg.p, xps_part, e->sbr_tables, tile);
          ++*launches;
        }
#define K4C_ARGS(R0) B.k4_runs, R0, B.run_frames, B.sbr_frames, B.core, e->d_sbr_chan, e->d_xg.p, B.pcm, B.pcm_off, B.pcm_bytes, \
                     e->sbr_tables, tile, B.ps_frames, e->d_ps_chan, xps_part
#define LAUNCH_K4C_BANK(FMT, DS)                                                                                                \
  do {                                                                                                                         \
    if (P.pn) { k4c_synthesis_kernel<FMT, false, DS><<<P.pn * n_groups, kK4cThreads, k4c_smem_bytes(), st>>>(K4C_ARGS(P.p0)); ++*launches; } \
    if (P.qn) {                                                                                                                \
      k4c_synthesis_kernel<FMT, true, DS><<<dim3(P.qn * n_groups, 2), kK4cThreads, k4c_smem_bytes(), st>>>(K4C_ARGS(P.q0));     \
      ++*launches;                                                                                                             \
    }                                                                                                                          \
  } while (0)
  // the two synthesis banks are two instantiations over the same grid; each leaves the other's runs alone
#define LAUNCH_K4C(FMT)                                                                                                        \
  do {                                                                                                                         \
    if (B.k4_banks & 1u) LAUNCH_K4C_BANK(FMT, false);                                                                          \
    if (B.k4_banks & 2u) LAUNCH_K4C_BANK(FMT, true);                                                                           \
  } while (0)
        if (e->opts.pcm_format == JAADB_PCM_S16LE) LAUNCH_K4C(0);
        else if (e->opts.pcm_format == JAADB_PCM_S16BE) LAUNCH_K4C(1);
        else LAUNCH_K4C(2);
#undef LAUNCH_K4C
#undef LAUNCH_K4C_BANK
#undef K4C_ARGS
        if (P.pn) { k4_commit_kernel<<<(P.pn + 255) / 256, 256, 0, st>>>(B.k4_runs + P.p0, P.pn, P.pn, e->d_sbr_chan, e->d_ps_chan); ++*launches; }
        if (P.qn) { k4_commit_kernel<<<(P.qn + 255) / 256, 256, 0, st>>>(B.k4_runs + P.q0, P.qn, 0u, e->d_sbr_chan, e->d_ps_chan); ++*launches; }
      }
    }
    for (int i = 1; i < n_parts; ++i) {
      if ((err = cudaEventRecord(e->k4_join[i - 1], parts[i].st)) != cudaSuccess) return err;
      if ((err = cudaStreamWaitEvent(e->stream, e->k4_join[i - 1], 0)) != cudaSuccess) return err;
    }
  }
  return cudaSuccess;
}

}  // namespace

extern "C" {

int jaadb_abi_version(void) { return JAADB_ABI_VERSION; }

const char* jaadb_status_string(int32_t st) {
  switch (st) {
    case JAADB_ST_OK: return "ok";
    case JAADB_ST_EOS: return "unexpected end of frame";
    case JAADB_ST_INVALID_CODEBOOK: return "invalid huffman codebook: 12";
    case JAADB_ST_TOO_MANY_BANDS: return "too many bands";
    case JAADB_ST_SF_RANGE: return "scalefactor out of range";
    case JAADB_ST_PULSE_SHORT: return "pulse data not allowed for short frames";
    case JAADB_ST_PULSE_RANGE: return "pulse data out of range";
    case JAADB_ST_MS_RESERVED: return "reserved MS mask type used";
    case JAADB_ST_TNS_ORDER: return "TNS filter out of range";
    case JAADB_ST_LTP_PROFILE: return "unexpected profile for LTP";
    case JAADB_ST_UNSUPPORTED_ELEMENT: return "syntax element outside the engine's scope";
    case JAADB_ST_LAYOUT: return "element sequence differs from the stream's channel layout";
    case JAADB_ST_PROFILE: return "unsupported profile";
    case JAADB_ST_ARRAY_BOUNDS: return "table index out of bounds";
    case JAADB_ST_SBR: return "SBR error";
    case JAADB_ST_CONFIG: return "bad configuration";
    default: return "unknown status";
  }
}

const char* jaadb_last_error(const jaadb_engine* e) { return e ? e->error.c_str() : "null engine"; }

int jaadb_engine_create(const jaadb_options* opts, jaadb_engine** out) {
  if (!opts || !out) return JAADB_E_INVALID;
  *out = nullptr;
  if (opts->pcm_format < 0 || opts->pcm_format > 2 || (opts->tns_mode != JAADB_TNS_JAAD && opts->tns_mode != JAADB_TNS_ISO) || opts->max_streams == 0)
    return JAADB_E_INVALID;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || opts->device < 0 || opts->device >= ndev) return JAADB_E_CUDA;
  jaadb_engine* e = new jaadb_engine();
  e->opts = *opts;
  auto fail = [&](int rc) { jaadb_engine_destroy(e); return rc; };
  if (cudaSetDevice(opts->device) != cudaSuccess) return fail(JAADB_E_CUDA);
  if (cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking) != cudaSuccess) return fail(JAADB_E_CUDA);
  for (auto& ev : e->ev)
    if (cudaEventCreate(&ev) != cudaSuccess) return fail(JAADB_E_CUDA);
  if (const char* v = getenv("JAADB200_K4_PARTS")) e->k4_parts = std::min(std::max(atoi(v), 1), (int)jaadb_engine::kK4MaxParts);   // tuning experiments only
  if (cudaEventCreateWithFlags(&e->k4_fork, cudaEventDisableTiming) != cudaSuccess) return fail(JAADB_E_CUDA);
  cudaDeviceGetAttribute(&e->sm_count, cudaDevAttrMultiProcessorCount, opts->device);
  if (e->sm_count <= 0) e->sm_count = 148;
  if (const char* v = getenv("JAADB_K1_LANES_LOG2")) { const int l = atoi(v); if (l >= 0 && l <= 5) e->k1_lanes_force = l; }
  for (int i = 0; i + 1 < jaadb_engine::kK4MaxParts; ++i) {
    if (cudaStreamCreateWithFlags(&e->k4_stream[i], cudaStreamNonBlocking) != cudaSuccess) return fail(JAADB_E_CUDA);
    if (cudaEventCreateWithFlags(&e->k4_join[i], cudaEventDisableTiming) != cudaSuccess) return fail(JAADB_E_CUDA);
  }
  int rc = init_tables(e);
  if (rc) return fail(rc);
  e->streams.resize(opts->max_streams);
  e->free_slots.reserve(opts->max_streams);
  for (int32_t i = (int32_t)opts->max_streams - 1; i >= 0; --i) e->free_slots.push_back(i);
  if (cudaMalloc(reinterpret_cast<void**>(&e->d_overlap), sizeof(float) * (size_t)opts->max_streams * kMaxChannels * 1024) != cudaSuccess)
    return fail(JAADB_E_NOMEM);
  if (cudaMalloc(reinterpret_cast<void**>(&e->d_sstate), sizeof(StreamState) * (size_t)opts->max_streams) != cudaSuccess)
    return fail(JAADB_E_NOMEM);
  // opt in to the shared-memory sizes the kernels need
  cudaFuncSetAttribute(k1_parse_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k1_smem_bytes(e->lut_entries));
  const int k2max = (int)k2_smem_bytes(kMaxChannels, kMaxChannels, false), k2max2 = (int)k2_smem_bytes(2, 2, true);
  // (the one- / two-channel instantiation wants as many resident CTAs as its registers allow: all of the SM's shared memory)
#define K2_ATTR(FMT)                                                                                                     \
  cudaFuncSetAttribute(k2_filterbank_kernel<FMT, 512, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, k2max);             \
  cudaFuncSetAttribute(k2_filterbank_kernel<FMT, 384, K2_MC_MIN_BLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, k2max); \
  cudaFuncSetAttribute(k2_filterbank_kernel<FMT, 384, K2_MC_MIN_BLOCKS>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); \
  cudaFuncSetAttribute(k2_filterbank_kernel<FMT, 128, K2_STEREO_MIN_BLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, k2max2); \
  if (K2_CARVEOUT >= 0) cudaFuncSetAttribute(k2_filterbank_kernel<FMT, 128, K2_STEREO_MIN_BLOCKS>, cudaFuncAttributePreferredSharedMemoryCarveout, K2_CARVEOUT)
  K2_ATTR(0); K2_ATTR(1); K2_ATTR(2);
#undef K2_ATTR
  if (cudaStreamSynchronize(e->stream) != cudaSuccess) return fail(JAADB_E_CUDA);
  *out = e;
  return JAADB_OK;
}

void jaadb_engine_destroy(jaadb_engine* e) {
  if (!e) return;
  cudaSetDevice(e->opts.device);
  if (e->stream) cudaStreamSynchronize(e->stream);
  for (void* p : e->table_allocs) cudaFree(p);
  if (e->d_overlap) cudaFree(e->d_overlap);
  if (e->d_sstate) cudaFree(e->d_sstate);
  if (e->d_sbr_elem) cudaFree(e->d_sbr_elem);
  if (e->d_sbr_chan) cudaFree(e->d_sbr_chan);
  e->d_xg.release();
  e->d_xps.release();
  if (e->d_ps_chan) cudaFree(e->d_ps_chan);
  for (auto& ev : e->ev)
    if (ev) cudaEventDestroy(ev);
  for (auto& st : e->k4_stream)
    if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  if (e->k4_fork) cudaEventDestroy(e->k4_fork);
  for (auto& ev : e->k4_join)
    if (ev) cudaEventDestroy(ev);
  auto& W = e->ws;
  if (W.copy_stream) { cudaStreamSynchronize(W.copy_stream); cudaStreamDestroy(W.copy_stream); }
  for (int i = 0; i < 2; ++i) {
    if (W.k_done[i]) cudaEventDestroy(W.k_done[i]);
    if (W.d2h_done[i]) cudaEventDestroy(W.d2h_done[i]);
    if (W.desc_done[i]) cudaEventDestroy(W.desc_done[i]);
    if (W.h_frames[i]) cudaFreeHost(W.h_frames[i]);
    if (W.h_run_frames[i]) cudaFreeHost(W.h_run_frames[i]);
    if (W.h_runs[i]) cudaFreeHost(W.h_runs[i]);
    if (W.h_segs[i]) cudaFreeHost(W.h_segs[i]);
    if (W.h_sbr_runs[i]) cudaFreeHost(W.h_sbr_runs[i]);
    if (W.h_k4_runs[i]) cudaFreeHost(W.h_k4_runs[i]);
    W.pcm[i].release();
  }
  W.blob.release(); W.frames.release(); W.fside.release(); W.iside.release(); W.q.release(); W.runs.release();
  W.run_frames.release(); W.pcm_bytes.release(); W.pcm_off.release(); W.k2frames.release(); W.segs.release(); W.ovl_stage.release();
  W.sbr_runs.release(); W.k4_runs.release(); W.sbr_frames.release(); W.core.release(); W.ps_frames.release();
  if (W.h_fside) cudaFreeHost(W.h_fside);
  if (W.h_off) cudaFreeHost(W.h_off);
  if (W.h_pcm_bytes) cudaFreeHost(W.h_pcm_bytes);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
}

int jaadb_stream_open_adts(jaadb_engine* e, int32_t profile, int32_t sf_index, int32_t channel_config,
                           int32_t expect_sbr, int32_t* stream_id) {
  if (!e || !stream_id) return JAADB_E_INVALID;
  cudaSetDevice(e->opts.device);
  StreamHost s;
  s.profile = profile_for_int(profile);   // ADTSFrame.getProfile -> Profile.forInt (S/adts/ADTSFrame.java:119-121)
  s.sf_index = sf_index;
  int cc = channel_config >= 7 ? channel_config + 1 : channel_config;  // ChannelConfiguration.forInt (:26-33)
  if (cc > 8 || cc == 7) { e->set_error("invalid channel configuration"); return JAADB_E_CONFIG; }
  s.chan_cfg = (cc == 8) ? 7 : cc;  // layout index 7 = 7.1 (8 channels)
  if (sf_index < 0 || sf_index > 11) { e->set_error("unsupported sampling frequency index"); return JAADB_E_CONFIG; }
  s.sample_rate = kSfFreq[sf_index];
  s.sample_length = 1024;
  if (expect_sbr < 0) expect_sbr = (sf_index >= 6) ? 1 : 0;
  s.sbr = expect_sbr;
  if (s.sbr) {
    // DecoderConfig.setSBRPresent (A/DecoderConfig.java:124-135): ADTS streams have no output frequency yet
    // (rates above 48 kHz cannot be doubled: SampleRate.duplicated() is SF_NONE and the stream runs the down-sampled bank)
    if (sf_index >= 3) { s.sample_rate = kSfFreq[sf_index - 3]; s.sample_length = 2048; s.sbr_sr_index = sf_index - 3; }
  }
  return finish_open(e, s, stream_id);
}

int jaadb_stream_open_asc(jaadb_engine* e, const uint8_t* asc, uint32_t n, int32_t* stream_id) {
  return jaadb_stream_open_asc_sbr(e, asc, n, 0, stream_id);
}

int jaadb_stream_open_asc_sbr(jaadb_engine* e, const uint8_t* asc, uint32_t n, int32_t expect_sbr, int32_t* stream_id) {
  if (!e || !asc || !stream_id) return JAADB_E_INVALID;
  cudaSetDevice(e->opts.device);
  // DecoderConfig.decode (A/DecoderConfig.java:175-254)
  HostBits in{asc, n};
  auto read_profile = [&]() { int i = (int)in.read(5); if (i == 31) i = 32 + (int)in.read(6); return profile_for_int(i); };
  auto read_rate = [&](int& index, int& freq) {
    int idx = (int)in.read(4);
    if (idx != 15) { if (idx >= 12) return false; index = idx; freq = kSfFreq[idx]; return true; }
    freq = (int)in.read(24);
    index = nominal_index(freq);
    return index >= 0;
  };
  StreamHost s;
  int profile = read_profile();
  int sf_index = -1, freq = 0;
  if (!read_rate(sf_index, freq)) { e->set_error("bad sampling frequency in AudioSpecificConfig"); return JAADB_E_CONFIG; }
  int out_freq = freq, out_index = sf_index;
  int cc = (int)in.read(4);
  if (cc >= 7) ++cc;
  if (cc > 8 || cc == 7 || cc == 0) { e->set_error("unsupported channel configuration in AudioSpecificConfig"); return JAADB_E_CONFIG; }
  int sbr = 0;
  switch (profile) {
    case 29: sbr = 2;  // fall through: PS implies SBR
    case 5: {
      if (!sbr) sbr = 1;
      int ext_index, ext_freq;
      if (!read_rate(ext_index, ext_freq)) { e->set_error("bad extension sampling frequency"); return JAADB_E_CONFIG; }
      profile = read_profile();
      out_freq = ext_freq;
      out_index = ext_index;
      break;
    }
    case 1: case 2: case 3: case 4: case 17: case 19: case 23: {
      if (in.read(1)) { e->set_error("config uses 960-sample frames, not yet supported"); return JAADB_E_CONFIG; }
      if (in.read(1)) in.read(14);
      if (in.read(1)) {
        if (profile > 16) in.read(3);
        in.read(1);
      }
      if (in.left() > 10) {
        // readSyncExtension (A/DecoderConfig.java:268-291)
        if (in.read(11) == 0x2B7) {
          int ext = profile_for_int((int)in.read(5));
          if (ext == 5 || ext == 22) {
            bool present = in.read(1) != 0;
            if (present) {
              int ext_index, ext_freq;
              if (!read_rate(ext_index, ext_freq)) { e->set_error("bad extension sampling frequency"); return JAADB_E_CONFIG; }
              out_freq = ext_freq;
              out_index = ext_index;
              sbr = 1;
            }
            if (ext == 5 && in.left() > 12 && in.read(11) == 0x548 && in.read(1)) sbr = 2;
          }
        }
      }
      break;
    }
    default:
      e->set_error("profile not supported");
      return JAADB_E_CONFIG;
  }
  if (!in.ok) { e->set_error("AudioSpecificConfig truncated"); return JAADB_E_CONFIG; }
  s.profile = profile;
  s.sf_index = sf_index;
  s.chan_cfg = (cc == 8) ? 7 : cc;
  s.sample_rate = out_freq;
  // outputFrequency was set from the ASC: sample length doubles only if it differs (A/DecoderConfig.java:83-86)
  s.sample_length = (out_freq != freq) ? 2048 : 1024;
  // Implicit signalling: the ASC says nothing, the first SBR payload creates the tool (A/syntax/ChannelElement.java:65-76).
  // outputFrequency is already set by then, so it is not doubled and the stream runs the down-sampled bank (SURVEY A-20).
  s.sbr = std::max(sbr, std::max(expect_sbr, 0));
  s.sbr_sr_index = out_index;
  return finish_open(e, s, stream_id);
}

int jaadb_stream_close(jaadb_engine* e, int32_t id) {
  if (!e) return JAADB_E_INVALID;
  if (id < 0 || id >= (int32_t)e->streams.size() || !e->streams[id].open) return JAADB_E_NOSTREAM;
  e->streams[id].open = false;
  e->free_slots.push_back(id);
  return JAADB_OK;
}

int jaadb_stream_get_info(const jaadb_engine* e, int32_t id, jaadb_stream_info* info) {
  if (!e || !info) return JAADB_E_INVALID;
  if (id < 0 || id >= (int32_t)e->streams.size() || !e->streams[id].open) return JAADB_E_NOSTREAM;
  const StreamHost& s = e->streams[id];
  info->profile = s.profile;
  info->sf_index = s.sf_index;
  info->channel_config = s.chan_cfg;
  info->channels = s.out_channels;
  info->sample_rate = s.sample_rate;
  info->sample_length = s.sample_length;
  info->sbr = s.sbr;
  info->reserved = 0;
  return JAADB_OK;
}

int jaadb_batch_create(jaadb_engine* e, const jaadb_frame_desc* fd, uint32_t n, uint64_t blob_bytes,
                       const uint64_t* pcm_offsets, jaadb_batch** out) {
  if (!e || !out || (n && !fd)) return JAADB_E_INVALID;
  *out = nullptr;
  cudaSetDevice(e->opts.device);
  jaadb_batch* b = new jaadb_batch();
  b->e = e;
  b->n_frames = n;
  b->blob_bytes = blob_bytes;
  auto fail = [&](int rc) { jaadb_batch_destroy(b); return rc; };
  int rc = layout_pcm(e, fd, n, pcm_offsets, b->pcm_off, b->frame_pcm_size, &b->pcm_bytes);
  if (rc) return fail(rc);
  FrameIndex ix;
  rc = index_frames(e, fd, n, blob_bytes, ix);
  if (rc) return fail(rc);
  b->frames.swap(ix.frames);
  b->runs.swap(ix.runs);
  b->run_frames.swap(ix.run_frames);
  b->groups = ix.groups;
  b->segs.swap(ix.segs);
  b->n_ics = ix.n_ics;
  b->sbr_runs.swap(ix.sbr_runs);
  b->k4_runs.swap(ix.k4_runs);
  b->n_sbr_frames = ix.n_sbr_frames;
  b->n_k4_plain = ix.n_k4_plain;
  b->k4_max_count = ix.k4_max_count;
  b->k4_banks = ix.k4_banks;
  b->n_ps_frames = ix.n_ps_frames;
  const uint32_t ics = ix.n_ics;
  // device side
  cudaError_t ce = cudaSuccess;
  auto chk = [&](cudaError_t x) { if (ce == cudaSuccess) ce = x; };
  chk(b->d_blob.ensure(blob_bytes + 64));
  chk(b->d_pcm.ensure(std::max<uint64_t>(b->pcm_bytes, 16)));
  chk(b->d_frames.ensure(std::max<uint32_t>(n, 1)));
  chk(b->d_fside.ensure(std::max<uint32_t>(n, 1)));
  chk(b->d_iside.ensure(std::max<uint32_t>(ics, 1)));
  chk(b->d_q.ensure(std::max<size_t>((size_t)ics * 1024, 16)));
  chk(b->d_runs.ensure(std::max<size_t>(b->runs.size(), 1)));
  chk(b->d_run_frames.ensure(std::max<uint32_t>(n, 1)));
  chk(b->d_k2frames.ensure(std::max<uint32_t>(n, 1)));
  chk(b->d_segs.ensure(std::max<size_t>(b->segs.size(), 1)));
  {
    bool segmented = false;
    for (const auto& g : b->groups) segmented = segmented || g.segmented;
    if (segmented) chk(b->d_ovl_stage.ensure(b->runs.size() * kMaxChannels * 1024));
  }
  chk(b->d_pcm_bytes.ensure(std::max<uint32_t>(n, 1)));
  chk(b->d_pcm_off.ensure(std::max<uint32_t>(n, 1)));
  if (e->opts.flags & JAADB_FLAG_DEBUG_TAPS) chk(b->d_spec_tap.ensure(std::max<size_t>((size_t)ics * 1024, 16)));
  if (b->n_sbr_frames) {
    chk(b->d_sbr_runs.ensure(b->sbr_runs.size()));
    chk(b->d_k4_runs.ensure(b->k4_runs.size()));
    chk(b->d_sbr_frames.ensure((size_t)b->n_sbr_frames * 2));
    chk(b->d_core.ensure((size_t)ics * 1024));
    if (b->n_ps_frames) chk(b->d_ps_frames.ensure(b->n_ps_frames));
  }
  if (ce != cudaSuccess) { e->set_error(std::string("batch allocation: ") + cudaGetErrorString(ce)); return fail(JAADB_E_NOMEM); }
  if (n) {
    chk(cudaMemcpyAsync(b->d_frames.p, b->frames.data(), sizeof(FrameDev) * n, cudaMemcpyHostToDevice, e->stream));
    chk(cudaMemcpyAsync(b->d_runs.p, b->runs.data(), sizeof(RunDev) * b->runs.size(), cudaMemcpyHostToDevice, e->stream));
    chk(cudaMemcpyAsync(b->d_run_frames.p, b->run_frames.data(), sizeof(RunFrameDev) * n, cudaMemcpyHostToDevice, e->stream));
    chk(cudaMemcpyAsync(b->d_segs.p, b->segs.data(), sizeof(K2SegDev) * b->segs.size(), cudaMemcpyHostToDevice, e->stream));
    chk(cudaMemcpyAsync(b->d_pcm_off.p, b->pcm_off.data(), sizeof(uint64_t) * n, cudaMemcpyHostToDevice, e->stream));
    chk(cudaMemsetAsync(b->d_blob.p + blob_bytes, 0, 64, e->stream));
    if (b->n_sbr_frames) {
      chk(cudaMemcpyAsync(b->d_sbr_runs.p, b->sbr_runs.data(), sizeof(SbrRunDev) * b->sbr_runs.size(), cudaMemcpyHostToDevice, e->stream));
      chk(cudaMemcpyAsync(b->d_k4_runs.p, b->k4_runs.data(), sizeof(K4RunDev) * b->k4_runs.size(), cudaMemcpyHostToDevice, e->stream));
    }
    chk(cudaStreamSynchronize(e->stream));
  }
  if (ce != cudaSuccess) { e->set_error(std::string("batch upload: ") + cudaGetErrorString(ce)); return fail(JAADB_E_CUDA); }
  memset(&b->timings, 0, sizeof b->timings);
  *out = b;
  return JAADB_OK;
}

uint64_t jaadb_batch_pcm_bytes(const jaadb_batch* b) { return b ? b->pcm_bytes : 0; }

int jaadb_batch_upload(jaadb_batch* b, const uint8_t* blob, uint64_t blob_bytes) {
  if (!b || (!blob && blob_bytes) || blob_bytes != b->blob_bytes) return JAADB_E_INVALID;
  jaadb_engine* e = b->e;
  cudaSetDevice(e->opts.device);
  if (blob_bytes) CUDA_TRY(e, cudaMemcpyAsync(b->d_blob.p, blob, blob_bytes, cudaMemcpyDefault, e->stream));   // host or device source
  return JAADB_OK;
}

int jaadb_batch_decode(jaadb_batch* b) {
  if (!b) return JAADB_E_INVALID;
  jaadb_engine* e = b->e;
  cudaSetDevice(e->opts.device);
  const bool prof = (e->opts.flags & JAADB_FLAG_PROFILE) != 0;
  uint32_t launches = 0;
  if (b->n_frames == 0) { b->decoded = true; return JAADB_OK; }
  if (prof) CUDA_TRY(e, cudaEventRecord(e->ev[0], e->stream));
  DecodeBufs B{b->d_blob.p, b->d_frames.p, b->d_fside.p, b->d_iside.p, b->d_q.p, b->d_runs.p, (uint32_t)b->runs.size(), b->d_run_frames.p,
               b->d_k2frames.p, b->d_segs.p, b->d_ovl_stage.p, b->d_pcm.p,
               b->d_pcm_off.p, b->d_pcm_bytes.p, (e->opts.flags & JAADB_FLAG_DEBUG_TAPS) ? b->d_spec_tap.p : nullptr,
               b->d_sbr_runs.p, b->d_k4_runs.p, b->d_sbr_frames.p, b->d_core.p, b->d_ps_frames.p, b->n_k4_plain,
               b->k4_max_count, b->k4_banks};
  CUDA_TRY(e, launch_decode(e, b->groups.data(), b->groups.size(), b->n_frames, (uint32_t)b->sbr_runs.size(), (uint32_t)b->k4_runs.size(), B,
                prof ? e->ev[1] : nullptr, prof ? e->ev[3] : nullptr, &launches));
  if (prof) CUDA_TRY(e, cudaEventRecord(e->ev[2], e->stream));
  CUDA_TRY(e, cudaGetLastError());
  b->timings.launches = launches;
  b->decoded = true;
  return JAADB_OK;
}

int jaadb_batch_sync(jaadb_batch* b) {
  if (!b) return JAADB_E_INVALID;
  CUDA_TRY(b->e, cudaStreamSynchronize(b->e->stream));
  return JAADB_OK;
}

int jaadb_batch_timings(jaadb_batch* b, jaadb_timings* t) {
  if (!b || !t) return JAADB_E_INVALID;
  jaadb_engine* e = b->e;
  if (!(e->opts.flags & JAADB_FLAG_PROFILE) || !b->decoded) return JAADB_E_INVALID;
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  if (b->n_frames) {
    CUDA_TRY(e, cudaEventElapsedTime(&b->timings.parse_ms, e->ev[0], e->ev[1]));
    CUDA_TRY(e, cudaEventElapsedTime(&b->timings.filterbank_ms, e->ev[1], e->ev[3]));
    CUDA_TRY(e, cudaEventElapsedTime(&b->timings.sbr_ms, e->ev[3], e->ev[2]));
    CUDA_TRY(e, cudaEventElapsedTime(&b->timings.total_ms, e->ev[0], e->ev[2]));
  }
  *t = b->timings;
  return JAADB_OK;
}

int jaadb_batch_download(jaadb_batch* b, void* pcm_out, uint64_t cap, jaadb_frame_result* results) {
  if (!b || !b->decoded) return JAADB_E_INVALID;
  jaadb_engine* e = b->e;
  cudaSetDevice(e->opts.device);
  if (pcm_out) {
    if (cap < b->pcm_bytes) { e->set_error("pcm buffer too small"); return JAADB_E_CAPACITY; }
    if (b->pcm_bytes) CUDA_TRY(e, cudaMemcpyAsync(pcm_out, b->d_pcm.p, b->pcm_bytes, cudaMemcpyDeviceToHost, e->stream));
  }
  if (results && b->n_frames) {
    b->h_fside.resize(b->n_frames);
    b->h_pcm_bytes.resize(b->n_frames);
    CUDA_TRY(e, cudaMemcpyAsync(b->h_fside.data(), b->d_fside.p, sizeof(FrameSide) * b->n_frames, cudaMemcpyDeviceToHost, e->stream));
    CUDA_TRY(e, cudaMemcpyAsync(b->h_pcm_bytes.data(), b->d_pcm_bytes.p, sizeof(uint32_t) * b->n_frames, cudaMemcpyDeviceToHost, e->stream));
  }
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  if (results) {
    for (uint32_t i = 0; i < b->n_frames; ++i) {
      const StreamHost& s = e->streams[b->frames[i].stream_slot];
      jaadb_frame_result& r = results[i];
      r.status = b->h_fside[i].status;
      r.pcm_bytes = b->h_pcm_bytes[i];
      r.channels = r.status ? 0 : (uint16_t)s.out_channels;
      r.sample_length = r.status ? 0 : (uint16_t)s.sample_length;
      r.sample_rate = (uint32_t)s.sample_rate;
    }
  }
  return JAADB_OK;
}

void jaadb_batch_destroy(jaadb_batch* b) {
  if (!b) return;
  cudaSetDevice(b->e->opts.device);
  cudaStreamSynchronize(b->e->stream);
  b->d_blob.release(); b->d_pcm.release(); b->d_frames.release(); b->d_fside.release(); b->d_iside.release();
  b->d_q.release(); b->d_runs.release(); b->d_run_frames.release(); b->d_pcm_bytes.release(); b->d_pcm_off.release();
  b->d_k2frames.release(); b->d_segs.release(); b->d_ovl_stage.release();
  b->d_spec_tap.release();
  b->d_sbr_runs.release(); b->d_k4_runs.release(); b->d_sbr_frames.release(); b->d_core.release(); b->d_ps_frames.release();
  delete b;
}

// One-call decode, pipelined.  The frame array is cut into chunks of consecutive frames; per chunk: index on the
// host, upload descriptors, K1, K2 into one of two device PCM buffers, then the chunk's PCM byte range and result
// words go back over PCIe on a second stream while the next chunk is decoded.  Frames of a stream stay in array
// order across chunks because chunks run in order on one stream and the overlap state lives in HBM in between.
}  // extern "C"

// container_index.cpp
int64_t jaadb_internal_index_interleaved(int kind, const uint8_t* blob, const uint64_t* begin, uint32_t n_streams, const int32_t* stream_ids,
                                         std::vector<jaadb_frame_desc>& scratch, std::vector<jaadb_frame_desc>& frames, uint32_t threads);

namespace {

// the compressed bytes go on the bus (or across HBM) first, so that host-side work runs while they travel
int start_blob_upload(jaadb_engine* e, const uint8_t* blob, uint64_t blob_bytes) {
  auto& W = e->ws;
  const cudaError_t be = W.blob.ensure(blob_bytes + 64);
  if (be != cudaSuccess) { e->set_error(std::string("workspace allocation: ") + cudaGetErrorString(be)); return JAADB_E_NOMEM; }
  // (host or device source: a device blob is copied once inside HBM, which gives it the padding the bit readers rely on)
  if (blob_bytes) CUDA_TRY(e, cudaMemcpyAsync(W.blob.p, blob, blob_bytes, cudaMemcpyDefault, e->stream));
  CUDA_TRY(e, cudaMemsetAsync(W.blob.p + blob_bytes, 0, 64, e->stream));
  return JAADB_OK;
}

int decode_impl(jaadb_engine* e, const uint8_t* blob, uint64_t blob_bytes, const jaadb_frame_desc* frames, uint32_t n_frames,
                void* pcm_out, uint64_t pcm_capacity, const uint64_t* pcm_offsets, jaadb_frame_result* results, bool blob_in_flight);

}  // namespace

extern "C" {

int jaadb_decode(jaadb_engine* e, const uint8_t* blob, uint64_t blob_bytes, const jaadb_frame_desc* frames,
                 uint32_t n_frames, void* pcm_out, uint64_t pcm_capacity, const uint64_t* pcm_offsets,
                 jaadb_frame_result* results) {
  if (!e || (n_frames && !frames) || (!blob && blob_bytes)) return JAADB_E_INVALID;
  cudaSetDevice(e->opts.device);
  if (n_frames == 0) return JAADB_OK;
  return decode_impl(e, blob, blob_bytes, frames, n_frames, pcm_out, pcm_capacity, pcm_offsets, results, false);
}

int64_t jaadb_decode_containers(jaadb_engine* e, int32_t kind, const uint8_t* blob, const uint64_t* stream_begin, uint32_t n_streams,
                                const int32_t* stream_ids, void* pcm_out, uint64_t pcm_capacity, jaadb_frame_result* results,
                                uint64_t max_frames, jaadb_frame_desc* frames_out, uint32_t threads) {
  if (!e || !blob || !stream_begin || (kind != JAADB_CONTAINER_ADTS && kind != JAADB_CONTAINER_MP4)) return JAADB_E_INVALID;
  cudaSetDevice(e->opts.device);
  if (n_streams == 0) return 0;
  if (stream_begin[0] != 0) { e->set_error("stream_begin[0] must be 0: the blob starts with the first container"); return JAADB_E_INVALID; }
  const uint64_t blob_bytes = stream_begin[n_streams];
  // the containers travel to the GPU while the host threads index them (both read the same bytes)
  {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, blob) == cudaSuccess && pa.type == cudaMemoryTypeDevice) {
      e->set_error("jaadb_decode_containers indexes on the host: the containers must be in host memory");
      return JAADB_E_INVALID;
    }
    cudaGetLastError();
  }
  static const bool trace = getenv("JAADB200_TRACE") != nullptr;
  const auto t_call = std::chrono::steady_clock::now();
  int rc = start_blob_upload(e, blob, blob_bytes);
  if (rc) return rc;
  const int64_t n = jaadb_internal_index_interleaved(kind, blob, stream_begin, n_streams, stream_ids, e->scratch_frames_sm, e->scratch_frames, threads);
  if (n < 0 || (uint64_t)n > 0xFFFFFFFFull) { cudaStreamSynchronize(e->stream); e->set_error("container indexing failed"); return n < 0 ? n : JAADB_E_CAPACITY; }
  if ((results || frames_out) && (uint64_t)n > max_frames) {
    cudaStreamSynchronize(e->stream);
    e->set_error("the containers hold more frames than max_frames");
    return JAADB_E_CAPACITY;
  }
  if (n == 0) { cudaStreamSynchronize(e->stream); return 0; }
  if (trace) fprintf(stderr, "[jaadb] containers indexed (%lld frames) at %.2f ms\n", (long long)n,
                     std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count());
  // the caller's copy of the frame table (30 MB for two million frames) is made next to the decode, not in front of it
  std::thread copy_out;
  struct Joiner { std::thread& t; ~Joiner() { if (t.joinable()) t.join(); } } joiner{copy_out};   // (also on an exception)
  if (frames_out) {
    try {
      copy_out = std::thread([&] { memcpy(frames_out, e->scratch_frames.data(), (size_t)n * sizeof(jaadb_frame_desc)); });
    } catch (...) {
      memcpy(frames_out, e->scratch_frames.data(), (size_t)n * sizeof(jaadb_frame_desc));
    }
  }
  rc = decode_impl(e, blob, blob_bytes, e->scratch_frames.data(), (uint32_t)n, pcm_out, pcm_capacity, nullptr, results, true);
  return rc ? rc : n;
}

}  // extern "C"

namespace {

// One-call decode, pipelined (see jaadb_decode below for the contract).
int decode_impl(jaadb_engine* e, const uint8_t* blob, uint64_t blob_bytes, const jaadb_frame_desc* frames, uint32_t n_frames,
                void* pcm_out, uint64_t pcm_capacity, const uint64_t* pcm_offsets, jaadb_frame_result* results, bool blob_in_flight) {
  auto& W = e->ws;
  // JAADB200_TRACE=1: host-side timeline of the call on stderr (tuning aid)
  static const bool trace = getenv("JAADB200_TRACE") != nullptr;
  const auto t_call = std::chrono::steady_clock::now();
  auto ms_now = [&]() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_call).count(); };
  // The compressed frames go first (one copy: frames of a chunk may sit anywhere in the caller's blob), so that the
  // host-side layout work below runs while they are on the bus.
  if (!blob_in_flight) {
    const int urc = start_blob_upload(e, blob, blob_bytes);
    if (urc) return urc;
  }
  // A pcm_out in device memory (of this engine's GPU) is written by the kernels directly: no staging buffers, nothing on
  // PCIe but the descriptors in and the per-frame results out.
  bool out_dev = false;
  if (pcm_out) {
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, pcm_out) == cudaSuccess) out_dev = pa.type == cudaMemoryTypeDevice || pa.type == cudaMemoryTypeManaged;
    else cudaGetLastError();
    if (out_dev && (reinterpret_cast<uintptr_t>(pcm_out) & 3u)) { cudaStreamSynchronize(e->stream); e->set_error("device pcm_out must be 4-byte aligned"); return JAADB_E_INVALID; }
  }
  // (pinned: from a pageable vector the upload below would be a blocking copy queued behind the containers' 15 ms on the
  // bus -- the host reached the first chunk 4 ms later than the GPU could have started it)
  if (W.h_off_cap < n_frames) {
    if (W.h_off) cudaFreeHost(W.h_off);
    W.h_off = nullptr;
    W.h_off_cap = 0;
    if (cudaHostAlloc(reinterpret_cast<void**>(&W.h_off), sizeof(uint64_t) * (size_t)n_frames, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      cudaStreamSynchronize(e->stream);
      e->set_error("workspace allocation (pinned PCM offsets)");
      return JAADB_E_NOMEM;
    }
    W.h_off_cap = n_frames;
  }
  uint64_t* const off = W.h_off;
  std::vector<uint32_t>& size = e->scratch_size;
  // One pass over the frame table: PCM placement (the caller's offsets, or frames packed back to back in array order), the
  // checks of every descriptor -- before the first kernel: a bad one in a later chunk must not leave the streams of the
  // earlier chunks half-way through the call -- and what the workspace sizing needs.  (This runs while the compressed bytes
  // are on the bus; it used to be five passes.)
  uint64_t pcm_total = 0;
  int max_slots = 1;
  bool any_sbr = false;
  int rc = JAADB_OK;
  {
    if (size.size() < n_frames) size.resize(n_frames);
    const size_t n_streams = e->streams.size();
    const uint32_t per = (e->opts.pcm_format == JAADB_PCM_F32_PLANAR) ? 4u : 2u;
    // slices of the table on host threads; packed placement is a prefix sum, so: sizes + checks per slice, slice totals,
    // then offsets per slice
    const uint32_t n_thr = n_frames < (1u << 16) ? 1u : std::min<uint32_t>(16u, std::max(1u, std::thread::hardware_concurrency()));
    struct Slice { uint64_t bytes = 0, end = 0; int rc = JAADB_OK, slots = 1; bool sbr = false; };
    std::vector<Slice> sl(n_thr);
    auto bounds = [&](uint32_t t) { return std::make_pair((uint32_t)((uint64_t)n_frames * t / n_thr), (uint32_t)((uint64_t)n_frames * (t + 1) / n_thr)); };
    auto pass1 = [&](uint32_t t) {
      Slice& S = sl[t];
      const auto [lo, hi] = bounds(t);
      for (uint32_t i = lo; i < hi; ++i) {
        const jaadb_frame_desc& d = frames[i];
        if (d.stream_id < 0 || (size_t)d.stream_id >= n_streams || !e->streams[d.stream_id].open) { S.rc = JAADB_E_NOSTREAM; return; }
        if (d.nbytes > blob_bytes || d.offset > blob_bytes - d.nbytes || d.nbytes >= (1u << 29)) { S.rc = JAADB_E_INVALID; return; }
        const StreamHost& sh = e->streams[d.stream_id];
        const uint32_t sz = (uint32_t)sh.out_channels * (uint32_t)sh.sample_length * per;
        size[i] = sz;
        S.bytes += sz;
        if (pcm_offsets) {
          if ((pcm_offsets[i] & 3u) || pcm_offsets[i] > UINT64_MAX - sz) { S.rc = JAADB_E_INVALID; return; }
          off[i] = pcm_offsets[i];
          S.end = std::max(S.end, off[i] + sz);
        }
        S.slots = std::max(S.slots, sh.n_slots);
        S.sbr = S.sbr || sh.sbr != 0;
      }
    };
    auto pass2 = [&](uint32_t t, uint64_t pos) {
      const auto [lo, hi] = bounds(t);
      for (uint32_t i = lo; i < hi; ++i) { off[i] = pos; pos += size[i]; }
    };
    auto run = [&](auto&& fn) {
      if (n_thr == 1) { fn(0u); return; }
      std::vector<std::thread> pool;
      pool.reserve(n_thr);
      uint32_t started = 0;
      try { for (; started + 1 < n_thr; ++started) pool.emplace_back(fn, started + 1); } catch (...) {}
      fn(0u);
      for (uint32_t t = started + 1; t < n_thr; ++t) fn(t);   // (threads the host refused)
      for (auto& th : pool) th.join();
    };
    run(pass1);
    uint64_t end = 0;
    std::vector<uint64_t> start(n_thr + 1, 0);
    for (uint32_t t = 0; t < n_thr; ++t) {
      if (sl[t].rc != JAADB_OK && rc == JAADB_OK) rc = sl[t].rc;
      start[t + 1] = start[t] + sl[t].bytes;
      end = std::max(end, sl[t].end);
      max_slots = std::max(max_slots, sl[t].slots);
      any_sbr = any_sbr || sl[t].sbr;
    }
    if (rc == JAADB_E_NOSTREAM) e->set_error("frame refers to an unknown stream");
    else if (rc == JAADB_E_INVALID) e->set_error("a frame exceeds the blob, or a pcm offset is misaligned or out of range");
    if (rc == JAADB_OK && !pcm_offsets) {
      run([&](uint32_t t) { pass2(t, start[t]); });
      end = start[n_thr];
    }
    pcm_total = end;
  }
  if (rc) { cudaStreamSynchronize(e->stream); return rc; }
  if (pcm_out && pcm_capacity < pcm_total) { cudaStreamSynchronize(e->stream); e->set_error("pcm buffer too small"); return JAADB_E_CAPACITY; }

  // chunking: ~128 Ki frames per chunk, unless the caller's PCM placement is not monotonic over chunks
  uint32_t chunk = e->opts.chunk_frames ? e->opts.chunk_frames : 131072u;
  if (out_dev && !e->opts.chunk_frames) {
    // nothing to overlap with when the PCM stays in HBM: chunks only bound the workspace (quantised coefficients and side
    // information, 2448 bytes per channel-frame), 16 GB of it
    // (and 256 Ki frames keep the host's per-chunk indexing -- about as long as the chunk's kernels -- off the critical path)
    chunk = (uint32_t)std::min<uint64_t>(262144ull, (16ull << 30) / ((uint64_t)max_slots * 2448ull));
  }
  if (n_frames <= chunk + chunk / 2) chunk = n_frames;
  struct Range { uint32_t i0, i1; uint64_t lo, hi; };
  std::vector<Range> ranges;
  // the first chunks are short (from 1/8 of a chunk, growing by a quarter each): the PCM download -- the long pole of the
  // call -- starts after a fraction of a chunk's kernel time instead of a whole one, and stays fed while the chunks grow
  // (a chunk's kernels take up to 3/4 of the time of its download, so faster growth would starve the copy engine)
  const bool ramp = !e->opts.chunk_frames && n_frames >= 4 * chunk;
  uint32_t step = ramp ? chunk / 8 : chunk;
  for (uint32_t i0 = 0; i0 < n_frames;) {
    Range r{i0, std::min(n_frames, i0 + step), ~0ull, 0};
    if (!pcm_offsets) { r.lo = off[r.i0]; r.hi = off[r.i1 - 1] + size[r.i1 - 1]; }   // packed back to back in array order
    else for (uint32_t i = r.i0; i < r.i1; ++i) { r.lo = std::min(r.lo, off[i]); r.hi = std::max(r.hi, off[i] + size[i]); }
    ranges.push_back(r);
    i0 = r.i1;
    step = std::min(chunk, step + step / 4);
  }
  bool monotonic = true;
  for (size_t k = 1; k < ranges.size(); ++k) monotonic = monotonic && ranges[k].lo >= ranges[k - 1].hi;
  if (!monotonic && !out_dev) { ranges.assign(1, Range{0, n_frames, 0, pcm_total}); chunk = n_frames; }   // (device output needs no byte ranges)
  uint64_t max_pcm = 16;
  for (const auto& r : ranges) max_pcm = std::max(max_pcm, r.hi - r.lo);

  // workspace
  if (!W.copy_stream) {
    CUDA_TRY(e, cudaStreamCreateWithFlags(&W.copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
      CUDA_TRY(e, cudaEventCreateWithFlags(&W.k_done[i], cudaEventDisableTiming));
      CUDA_TRY(e, cudaEventCreateWithFlags(&W.d2h_done[i], cudaEventDisableTiming));
      CUDA_TRY(e, cudaEventCreateWithFlags(&W.desc_done[i], cudaEventDisableTiming));
    }
  }
  const size_t max_ics = (size_t)chunk * max_slots;   // worst case channel slots per frame in this call
  cudaError_t ce = cudaSuccess;
  auto chk = [&](cudaError_t x) { if (ce == cudaSuccess) ce = x; };
  if (!out_dev) {
    chk(W.pcm[0].ensure(max_pcm));
    if (ranges.size() > 1) chk(W.pcm[1].ensure(max_pcm));
  }
  chk(W.frames.ensure(chunk));
  chk(W.fside.ensure(n_frames));
  chk(W.pcm_bytes.ensure(n_frames));
  chk(W.iside.ensure(max_ics));
  chk(W.q.ensure(max_ics * 1024));
  chk(W.runs.ensure(e->streams.size()));
  chk(W.run_frames.ensure(chunk));
  chk(W.k2frames.ensure(chunk));
  chk(W.segs.ensure((size_t)chunk + e->streams.size()));
  chk(W.pcm_off.ensure(n_frames));
  if (any_sbr) {
    chk(W.sbr_runs.ensure(e->streams.size()));
    chk(W.k4_runs.ensure(e->streams.size() * 2));
    chk(W.sbr_frames.ensure((size_t)chunk * 2));
    chk(W.ps_frames.ensure((size_t)chunk));
    chk(W.core.ensure(max_ics * 1024));
  }
  if (ce == cudaSuccess && (W.h_chunk_cap < chunk || W.h_runs_cap < e->streams.size())) {
    for (int i = 0; i < 2; ++i) {
      if (W.h_frames[i]) cudaFreeHost(W.h_frames[i]);
      if (W.h_run_frames[i]) cudaFreeHost(W.h_run_frames[i]);
      if (W.h_runs[i]) cudaFreeHost(W.h_runs[i]);
      if (W.h_segs[i]) cudaFreeHost(W.h_segs[i]);
      if (W.h_sbr_runs[i]) cudaFreeHost(W.h_sbr_runs[i]);
      if (W.h_k4_runs[i]) cudaFreeHost(W.h_k4_runs[i]);
      W.h_frames[i] = nullptr; W.h_run_frames[i] = nullptr; W.h_runs[i] = nullptr; W.h_segs[i] = nullptr; W.h_sbr_runs[i] = nullptr; W.h_k4_runs[i] = nullptr;
    }
    W.h_chunk_cap = W.h_runs_cap = 0;
    const size_t cc = std::max<size_t>(chunk, W.h_chunk_cap), rr = e->streams.size();
    for (int i = 0; i < 2; ++i) {
      chk(cudaHostAlloc(reinterpret_cast<void**>(&W.h_frames[i]), sizeof(FrameDev) * cc, cudaHostAllocDefault));
      chk(cudaHostAlloc(reinterpret_cast<void**>(&W.h_run_frames[i]), sizeof(RunFrameDev) * cc, cudaHostAllocDefault));
      chk(cudaHostAlloc(reinterpret_cast<void**>(&W.h_runs[i]), sizeof(RunDev) * rr, cudaHostAllocDefault));
      chk(cudaHostAlloc(reinterpret_cast<void**>(&W.h_segs[i]), sizeof(K2SegDev) * (cc + rr), cudaHostAllocDefault));
      chk(cudaHostAlloc(reinterpret_cast<void**>(&W.h_sbr_runs[i]), sizeof(SbrRunDev) * rr, cudaHostAllocDefault));
      chk(cudaHostAlloc(reinterpret_cast<void**>(&W.h_k4_runs[i]), sizeof(K4RunDev) * rr * 2, cudaHostAllocDefault));
    }
    if (ce == cudaSuccess) { W.h_chunk_cap = cc; W.h_runs_cap = rr; }
  }
  if (ce == cudaSuccess && W.h_cap < n_frames) {
    if (W.h_fside) cudaFreeHost(W.h_fside);
    if (W.h_pcm_bytes) cudaFreeHost(W.h_pcm_bytes);
    W.h_fside = nullptr; W.h_pcm_bytes = nullptr; W.h_cap = 0;
    chk(cudaHostAlloc(reinterpret_cast<void**>(&W.h_fside), sizeof(FrameSide) * (size_t)n_frames, cudaHostAllocDefault));
    chk(cudaHostAlloc(reinterpret_cast<void**>(&W.h_pcm_bytes), sizeof(uint32_t) * (size_t)n_frames, cudaHostAllocDefault));
    if (ce == cudaSuccess) W.h_cap = n_frames;
  }
  if (ce != cudaSuccess) {
    cudaStreamSynchronize(e->stream);
    e->set_error(std::string("workspace allocation: ") + cudaGetErrorString(ce));
    return JAADB_E_NOMEM;
  }

  // PCM placement of every frame
  CUDA_TRY(e, cudaMemcpyAsync(W.pcm_off.p, off, sizeof(uint64_t) * n_frames, cudaMemcpyHostToDevice, e->stream));

  FrameIndex& ix = e->scratch_ix;
  uint32_t launches = 0;
  // per-frame results of frames [i0, i1) from the pinned copies of the device-side records
  auto convert_results = [&](uint32_t i0, uint32_t i1) {
    for (uint32_t i = i0; i < i1; ++i) {
      const StreamHost& s = e->streams[frames[i].stream_id];
      jaadb_frame_result& r = results[i];
      r.status = W.h_fside[i].status;
      r.pcm_bytes = W.h_pcm_bytes[i];
      r.channels = r.status ? 0 : (uint16_t)s.out_channels;
      r.sample_length = r.status ? 0 : (uint16_t)s.sample_length;
      r.sample_rate = (uint32_t)s.sample_rate;
    }
  };
  // an error inside the pipelined loop: nothing may still be writing the caller's buffers when the call returns
#define CUDA_TRY_SYNC(e, expr)                                                                 \
  do {                                                                                         \
    cudaError_t _err = (expr);                                                                 \
    if (_err != cudaSuccess) {                                                                 \
      (e)->set_error(std::string(#expr) + ": " + cudaGetErrorString(_err));                    \
      cudaStreamSynchronize((e)->stream);                                                      \
      cudaStreamSynchronize(W.copy_stream);                                                    \
      return JAADB_E_CUDA;                                                                     \
    }                                                                                          \
  } while (0)
  for (size_t k = 0; k < ranges.size(); ++k) {
    const Range& r = ranges[k];
    const uint32_t n = r.i1 - r.i0;
    const int pb = (int)(k & 1);
    if (k >= 2) CUDA_TRY_SYNC(e, cudaEventSynchronize(W.desc_done[pb]));   // staging slot pb has been consumed
    const double t_wait = ms_now();
    ix.frames_out = W.h_frames[pb];
    ix.run_frames_out = W.h_run_frames[pb];
    rc = index_frames(e, frames + r.i0, n, blob_bytes, ix);
    const double t_idx = ms_now();
    if (rc) { cudaStreamSynchronize(e->stream); cudaStreamSynchronize(W.copy_stream); return rc; }
    memcpy(W.h_runs[pb], ix.runs.data(), sizeof(RunDev) * ix.runs.size());
    memcpy(W.h_segs[pb], ix.segs.data(), sizeof(K2SegDev) * ix.segs.size());
    {
      bool segmented = false;
      for (const auto& g : ix.groups) segmented = segmented || g.segmented;
      if (segmented) {
        const cudaError_t se = W.ovl_stage.ensure(ix.runs.size() * kMaxChannels * 1024);
        if (se != cudaSuccess) { cudaStreamSynchronize(e->stream); cudaStreamSynchronize(W.copy_stream); e->set_error("workspace allocation"); return JAADB_E_NOMEM; }
      }
    }
    // the device descriptor buffers are still being read by the previous chunk's kernels: stream order protects them
    CUDA_TRY_SYNC(e, cudaMemcpyAsync(W.frames.p, W.h_frames[pb], sizeof(FrameDev) * n, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY_SYNC(e, cudaMemcpyAsync(W.runs.p, W.h_runs[pb], sizeof(RunDev) * ix.runs.size(), cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY_SYNC(e, cudaMemcpyAsync(W.run_frames.p, W.h_run_frames[pb], sizeof(RunFrameDev) * n, cudaMemcpyHostToDevice, e->stream));
    CUDA_TRY_SYNC(e, cudaMemcpyAsync(W.segs.p, W.h_segs[pb], sizeof(K2SegDev) * ix.segs.size(), cudaMemcpyHostToDevice, e->stream));
    if (k >= 2 && !out_dev) CUDA_TRY_SYNC(e, cudaStreamWaitEvent(e->stream, W.d2h_done[pb], 0));   // PCM buffer pb is free again
    if (!ix.sbr_runs.empty()) {
      // (pinned, double buffered like the other descriptors: the host goes on to index the next chunk)
      memcpy(W.h_sbr_runs[pb], ix.sbr_runs.data(), sizeof(SbrRunDev) * ix.sbr_runs.size());
      memcpy(W.h_k4_runs[pb], ix.k4_runs.data(), sizeof(K4RunDev) * ix.k4_runs.size());
      CUDA_TRY_SYNC(e, cudaMemcpyAsync(W.sbr_runs.p, W.h_sbr_runs[pb], sizeof(SbrRunDev) * ix.sbr_runs.size(), cudaMemcpyHostToDevice, e->stream));
      CUDA_TRY_SYNC(e, cudaMemcpyAsync(W.k4_runs.p, W.h_k4_runs[pb], sizeof(K4RunDev) * ix.k4_runs.size(), cudaMemcpyHostToDevice, e->stream));
    }
    CUDA_TRY_SYNC(e, cudaEventRecord(W.desc_done[pb], e->stream));   // staging slot pb is consumed once the copies above are done
    DecodeBufs B{W.blob.p, W.frames.p, W.fside.p + r.i0, W.iside.p, W.q.p, W.runs.p, (uint32_t)ix.runs.size(), W.run_frames.p,
                 W.k2frames.p, W.segs.p, W.ovl_stage.p, out_dev ? static_cast<uint8_t*>(pcm_out) : W.pcm[pb].p - r.lo,
                 W.pcm_off.p + r.i0, W.pcm_bytes.p + r.i0, nullptr, W.sbr_runs.p, W.k4_runs.p, W.sbr_frames.p, W.core.p, W.ps_frames.p,
                 ix.n_k4_plain, ix.k4_max_count, ix.k4_banks};
    CUDA_TRY_SYNC(e, launch_decode(e, ix.groups.data(), ix.groups.size(), n, (uint32_t)ix.sbr_runs.size(), (uint32_t)ix.k4_runs.size(), B,
                              nullptr, nullptr, &launches));
    CUDA_TRY_SYNC(e, cudaGetLastError());
    CUDA_TRY_SYNC(e, cudaEventRecord(W.k_done[pb], e->stream));
    CUDA_TRY_SYNC(e, cudaStreamWaitEvent(W.copy_stream, W.k_done[pb], 0));
    if (pcm_out && !out_dev && r.hi > r.lo)
      CUDA_TRY_SYNC(e, cudaMemcpyAsync(static_cast<uint8_t*>(pcm_out) + r.lo, W.pcm[pb].p, r.hi - r.lo, cudaMemcpyDeviceToHost, W.copy_stream));
    if (results) {
      CUDA_TRY_SYNC(e, cudaMemcpyAsync(W.h_fside + r.i0, W.fside.p + r.i0, sizeof(FrameSide) * n, cudaMemcpyDeviceToHost, W.copy_stream));
      CUDA_TRY_SYNC(e, cudaMemcpyAsync(W.h_pcm_bytes + r.i0, W.pcm_bytes.p + r.i0, sizeof(uint32_t) * n, cudaMemcpyDeviceToHost, W.copy_stream));
    }
    // Chunk k is on its way; only now does the host wait for chunk k - 2 to be back (d2h_done[pb] still stands for it) and
    // convert its results, while the GPU works on chunks k - 1 and k.  (Waiting before indexing chunk k left the compute
    // stream idle for the host's indexing time once per chunk.)
    if (k >= 2 && results) {
      CUDA_TRY_SYNC(e, cudaEventSynchronize(W.d2h_done[pb]));
      convert_results(ranges[k - 2].i0, ranges[k - 2].i1);
    }
    CUDA_TRY_SYNC(e, cudaEventRecord(W.d2h_done[pb], W.copy_stream));
    if (trace) fprintf(stderr, "[jaadb] chunk %zu: %u frames, waited until %.2f ms, indexed by %.2f, launched by %.2f\n", k, n, t_wait, t_idx, ms_now());
  }
#undef CUDA_TRY_SYNC
  if (trace) fprintf(stderr, "[jaadb] all chunks submitted at %.2f ms\n", ms_now());
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  if (trace) fprintf(stderr, "[jaadb] kernels done at %.2f ms\n", ms_now());
  CUDA_TRY(e, cudaStreamSynchronize(W.copy_stream));
  if (trace) fprintf(stderr, "[jaadb] downloads done at %.2f ms\n", ms_now());
  if (results) convert_results(ranges.size() >= 2 ? ranges[ranges.size() - 2].i0 : 0, n_frames);
  return JAADB_OK;
}

}  // namespace

extern "C" {

int jaadb_batch_tap(jaadb_batch* b, uint32_t frame, uint32_t ch, int16_t* q, int16_t* sfidx, uint8_t* sfbcb, float* spec,
                    int32_t* info, uint8_t* ms_used128) {
  if (!b || !b->decoded || frame >= b->n_frames) return JAADB_E_INVALID;
  jaadb_engine* e = b->e;
  cudaSetDevice(e->opts.device);
  const StreamHost& s = e->streams[b->frames[frame].stream_slot];
  if ((int)ch >= s.n_slots) return JAADB_E_INVALID;
  const uint32_t ics = b->frames[frame].ics_base + ch;
  CUDA_TRY(e, cudaStreamSynchronize(e->stream));
  IcsSide side;
  CUDA_TRY(e, cudaMemcpy(&side, b->d_iside.p + ics, sizeof side, cudaMemcpyDeviceToHost));
  const int nb = side.num_groups * side.max_sfb;
  if (info) {
    info[0] = side.present; info[1] = side.window_sequence; info[2] = side.window_shape; info[3] = side.info_decoded;
    info[4] = side.max_sfb; info[5] = side.num_groups;
    for (int i = 0; i < 8; ++i) info[6 + i] = i < side.num_groups ? side.group_len[i] : 0;
    info[14] = side.ms_mask; info[15] = side.common_window;
  }
  if (sfbcb) { memset(sfbcb, 0, 120); for (int i = 0; i < nb && i < 120; ++i) sfbcb[i] = side.sfb_cb[i]; }
  if (sfidx) {
    for (int i = 0; i < 120; ++i) sfidx[i] = -1;
    for (int i = 0; i < nb && i < 120; ++i) sfidx[i] = side.sf_idx[i] == 0xFFFF ? (int16_t)-1 : (int16_t)side.sf_idx[i];
  }
  if (ms_used128) for (int i = 0; i < 128; ++i) ms_used128[i] = (side.ms_used[i >> 3] >> (i & 7)) & 1;
  if (q) {
    std::vector<int16_t> raw(1024);
    CUDA_TRY(e, cudaMemcpy(raw.data(), b->d_q.p + (size_t)ics * 1024, 2048, cudaMemcpyDeviceToHost));
    // present the coefficients the way ICStream.decodeSpectralData leaves them: zero outside coded bands
    memset(q, 0, 2048);
    const bool sh = side.window_sequence == 2;
    const int16_t* swb = sh ? T::SWB_OFFSET_SHORT + 17 * s.sf_index : T::SWB_OFFSET_LONG + 53 * s.sf_index;
    const int swbc = sh ? T::SWB_SHORT_WINDOW_COUNT[s.sf_index] : T::SWB_LONG_WINDOW_COUNT[s.sf_index];
    // K1 leaves q in bitstream order (group base = 128 * first window; band at glen * swb[sfb]; window-major inside)
    int goff = 0, idx = 0;
    for (int g = 0; g < side.num_groups; ++g) {
      const int glen = side.group_len[g];
      for (int sfb = 0; sfb < side.max_sfb; ++sfb, ++idx) {
        int cb = side.sfb_cb[idx];
        if (cb < 1 || cb > 11 || sfb >= swbc) continue;
        const int lo = swb[sfb], width = swb[sfb + 1] - lo;
        for (int w = 0; w < glen; ++w)
          for (int k = 0; k < width; ++k) q[goff + w * 128 + lo + k] = raw[goff + glen * lo + w * width + k];
      }
      goff += glen * 128;
    }
  }
  if (spec) {
    if (!b->d_spec_tap.p) return JAADB_E_INVALID;
    CUDA_TRY(e, cudaMemcpy(spec, b->d_spec_tap.p + (size_t)ics * 1024, 4096, cudaMemcpyDeviceToHost));
  }
  return JAADB_OK;
}

int jaadb_batch_tap_sbr(jaadb_batch* b, uint32_t frame, uint32_t ch, void* out, uint32_t out_bytes) {
  if (!b || !b->decoded || frame >= b->n_frames || ch > 1 || !out) return JAADB_E_INVALID;
  jaadb_engine* e = b->e;
  cudaSetDevice(e->opts.device);
  const int32_t slot = b->frames[frame].stream_slot;
  for (const SbrRunDev& r : b->sbr_runs) {
    if (r.stream_slot != slot) continue;
    for (uint32_t it = 0; it < r.count; ++it) {
      if (b->run_frames[r.first + it].frame != frame) continue;
      if (out_bytes < sizeof(SbrFrameDev)) return JAADB_E_CAPACITY;
      CUDA_TRY(e, cudaStreamSynchronize(e->stream));
      CUDA_TRY(e, cudaMemcpy(out, b->d_sbr_frames.p + ((size_t)r.sbr_base + it) * 2 + ch, sizeof(SbrFrameDev), cudaMemcpyDeviceToHost));
      return (int)sizeof(SbrFrameDev);
    }
  }
  return 0;
}

int jaadb_batch_tap_ps(jaadb_batch* b, uint32_t frame, void* out, uint32_t out_bytes) {
  if (!b || !b->decoded || frame >= b->n_frames || !out) return JAADB_E_INVALID;
  jaadb_engine* e = b->e;
  cudaSetDevice(e->opts.device);
  const int32_t slot = b->frames[frame].stream_slot;
  for (const SbrRunDev& r : b->sbr_runs) {
    if (r.stream_slot != slot || !r.ps) continue;
    for (uint32_t it = 0; it < r.count; ++it) {
      if (b->run_frames[r.first + it].frame != frame) continue;
      if (out_bytes < sizeof(PsFrameDev)) return JAADB_E_CAPACITY;
      CUDA_TRY(e, cudaStreamSynchronize(e->stream));
      CUDA_TRY(e, cudaMemcpy(out, b->d_ps_frames.p + (size_t)r.ps_base + it, sizeof(PsFrameDev), cudaMemcpyDeviceToHost));
      return (int)sizeof(PsFrameDev);
    }
  }
  return 0;
}

namespace {
// decodes `frame` with the scratch stream `sid` (opened expecting SBR, and PS on a mono core) and reports what it carries
int probe_with_stream(jaadb_engine* e, int32_t sid, bool mono, const uint8_t* frame, uint32_t nbytes, int32_t* expect_sbr) {
  jaadb_frame_desc fd{0, nbytes, sid};
  jaadb_batch* b = nullptr;
  int rc = jaadb_batch_create(e, &fd, 1, nbytes, nullptr, &b);
  if (rc == JAADB_OK) rc = jaadb_batch_upload(b, frame, nbytes);
  if (rc == JAADB_OK) rc = jaadb_batch_decode(b);
  if (rc == JAADB_OK) rc = jaadb_batch_sync(b);
  if (rc == JAADB_OK) {
    FrameSide fs;
    memset(&fs, 0, sizeof fs);
    if (cudaMemcpy(&fs, b->d_fside.p, sizeof fs, cudaMemcpyDeviceToHost) != cudaSuccess) rc = JAADB_E_CUDA;
    else if (fs.sbr_bits[0] != 0) {
      *expect_sbr = 1;
      if (mono && b->d_ps_frames.p) {
        PsFrameDev pf;
        memset(&pf, 0, sizeof pf);
        if (cudaMemcpy(&pf, b->d_ps_frames.p, sizeof pf, cudaMemcpyDeviceToHost) != cudaSuccess) rc = JAADB_E_CUDA;
        else if (pf.use_ps) *expect_sbr = 2;
      }
    }
  }
  if (b) jaadb_batch_destroy(b);
  jaadb_stream_close(e, sid);
  return rc;
}
}  // namespace

int jaadb_probe_sbr(jaadb_engine* e, int32_t profile, int32_t sf_index, int32_t channel_config, const uint8_t* frame,
                    uint32_t nbytes, int32_t* expect_sbr) {
  if (!e || !frame || !expect_sbr) return JAADB_E_INVALID;
  *expect_sbr = 0;
  // SBR is only implemented for one SCE or one CPE (jaadb_stream_open_*)
  if (channel_config < 1 || channel_config > 2 || sf_index < 0 || sf_index > 11) return JAADB_OK;
  int32_t sid = -1;
  int rc = jaadb_stream_open_adts(e, profile, sf_index, channel_config, channel_config == 1 ? 2 : 1, &sid);
  if (rc) return rc;
  return probe_with_stream(e, sid, channel_config == 1, frame, nbytes, expect_sbr);
}

int jaadb_probe_sbr_asc(jaadb_engine* e, const uint8_t* asc, uint32_t asc_bytes, const uint8_t* frame, uint32_t nbytes,
                        int32_t* expect_sbr) {
  if (!e || !asc || !frame || !expect_sbr) return JAADB_E_INVALID;
  *expect_sbr = 0;
  int32_t sid = -1;
  // the scratch stream expects everything a mono / stereo stream could carry; other layouts have no SBR in this engine
  int rc = jaadb_stream_open_asc_sbr(e, asc, asc_bytes, 2, &sid);
  if (rc == JAADB_E_CONFIG) rc = jaadb_stream_open_asc_sbr(e, asc, asc_bytes, 1, &sid);
  if (rc == JAADB_E_CONFIG) {
    rc = jaadb_stream_open_asc(e, asc, asc_bytes, &sid);   // a valid ASC the SBR tool does not apply to: no SBR
    if (rc == JAADB_OK) jaadb_stream_close(e, sid);
    return rc;
  }
  if (rc) return rc;
  const bool mono = e->streams[sid].chan_cfg == 1;
  return probe_with_stream(e, sid, mono, frame, nbytes, expect_sbr);
}

}  // extern "C"
