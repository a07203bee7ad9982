// Device-side data layout of the SBR tool (HE-AAC v1), shared by K3 (payload parse), K4 (QMF / HF generation /
// HF adjustment) and the host runtime.
//
// HBM layout (per engine, persistent, one entry per stream x SBR element):
//   sbr_elem   SbrElemDev   [max_streams][2]   everything sbr/SBR.java + Channel.java keep between frames that the
//                                              PARSE needs: headers, band tables, patches, E/Q and their *_prev copies,
//                                              grid of the last frame, delta flags, harmonics
//   sbr_chan   SbrChanDev   [max_streams][2]   what the PROCESS keeps per channel: QMF analysis history (288), the last
//                                              8 slots of Xsbr, 9 synthesis v-vectors, the G/Q smoothing ring, chirp
//                                              factors, noise / sine phase indices
// Per batch:
//   sbr_frame  SbrFrameDev  [n_sbr_frames][2]  K3 out / K4 in: per channel of every SBR element frame, the dequantised
//                                              envelopes + the band tables in force for that frame (they can change
//                                              in the middle of a batch when a header arrives)
//   core       float        [n_ics][1024]      K2 out / K4 in: core-coder PCM of SBR streams
#pragma once
#include <cstdint>

namespace jaadb {

constexpr int kSbrMaxM = 49;        // SBR.MAX_M
constexpr int kSbrMaxLE = 5;        // SBR.MAX_L_E
constexpr int kSbrSlots = 32;       // numTimeSlotsRate
constexpr int kSbrHfGen = 8, kSbrHfAdj = 2;
constexpr int kSbrChansPerStream = 2;   // SBR is supported for mono and stereo streams (one SCE or one CPE)

struct SbrHeaderDev {  // sbr/Header.java
  uint8_t present;
  uint8_t amp_res, start_freq, stop_freq, xover_band, freq_scale, alter_scale, noise_bands;
  uint8_t limiter_bands, limiter_gains, interpol_freq, smoothing_mode;
};

// Parse-side state of one channel (sbr/Channel.java fields the bitstream syntax reads or updates)
struct SbrChanParse {
  int16_t E[64][kSbrMaxLE];
  int16_t Q[64][2];
  int16_t E_prev[64], Q_prev[64];
  uint8_t bs_add_harmonic[64], bs_add_harmonic_prev[64];
  uint8_t t_E[6], t_Q[3], f[6];
  uint8_t bs_df_env[9], bs_df_noise[3], bs_invf_mode[5];
  uint8_t bs_rel_bord[9], bs_rel_bord_0[9], bs_rel_bord_1[9];
  uint8_t amp_res, L_E, L_E_prev, L_Q, f_prev, frame_class, bs_pointer, bs_num_rel_0, bs_num_rel_1;
  uint8_t abs_bord_lead, abs_bord_trail;
  uint8_t add_harmonic_flag, add_harmonic_flag_prev;
  int8_t l_A, prevEnvIsShort;
  uint8_t pad[3];
};

struct SbrElemDev {
  SbrHeaderDev hdr, hdr_saved;
  uint8_t opened;        // an SBR payload has been seen (ChannelElement.sbr != null)
  uint8_t reset, valid, bs_coupling;
  uint8_t k0, kx, M, N_master, N_high, N_low, N_Q, kx_prev, M_prev;
  uint8_t n[2];
  uint8_t N_L[4];
  uint8_t noPatches;
  uint8_t f_master[64];
  uint8_t f_table_res[2][64];
  uint8_t f_table_noise[64];
  int8_t f_table_lim[4][64];
  uint8_t table_map_k_to_g[64];
  uint8_t patchNoSubbands[64];
  int8_t patchStartSubband[64];
  uint8_t pad[3];
  SbrChanParse ch[2];
};

// One channel of one SBR element frame, as K4 consumes it.
struct __align__(16) SbrFrameDev {
  float E_orig[kSbrMaxLE][64];      // [envelope][band of the envelope's resolution]
  float Q_div[2][8], Q_div2[2][8];  // [noise floor][noise band]
  uint8_t f_table_res[2][64];
  uint8_t f_table_noise[8];
  int8_t f_table_lim[64];           // for the header's bs_limiter_bands
  uint8_t table_map_k_to_g[64];
  uint8_t bs_add_harmonic[64], bs_add_harmonic_prev[64];
  uint8_t patchNoSubbands[8];
  int8_t patchStartSubband[8];
  uint8_t t_E[6], t_Q[3], f[6], bs_invf_mode[5];
  uint8_t mode;                     // 0: no SBR data this frame (upsample), 1: SBR without header (analysis + synthesis
                                    // of the low band only), 2: full process
  uint8_t reset, L_E, L_Q, kx, M, N_high, N_low, N_Q, N_L, kx_prev, M_prev, noPatches;
  uint8_t limiter_gains, interpol_freq, smoothing_mode;
  uint8_t add_harmonic_flag_prev;
  int8_t l_A, prevEnvIsShort;
  uint8_t frame_status;             // the frame's final status != 0: nothing is processed
  uint8_t pad[16];
};
static_assert(sizeof(SbrFrameDev) == 1872, "SbrFrameDev layout (mirrored by jaadec_b200/engine.py SBR_FRAME_DTYPE)");

// Process-side persistent state of one SBR channel.
struct __align__(16) SbrChanDev {
  float ana_hist[288];              // the last 288 core samples (QMF analysis ring of sbr/AnalysisFilterbank.java)
  float xsbr[kSbrHfGen][64][2];     // Xsbr rows 0..7 (the last 8 slots of the previous frame, SBR.sbr_save_matrix)
  float syn_v[9][128];              // the 9 most recent synthesis v-vectors ([0] = newest; sbr/SynthesisFilterbank64.java)
  float G_temp_prev[5][64], Q_temp_prev[5][64];
  float bwArray_prev[8];
  uint8_t bs_invf_mode_prev[8];
  int32_t GQ_ringbuf_index, index_noise_prev, psi_is_prev;
  int32_t pad;
};

// Read-only tables of the SBR tool (engine-owned device memory).
struct SbrTablesDev {
  const int16_t* huff[10];          // t_env15, f_env15, t_bal15, f_bal15, t_env30, f_env30, t_bal30, f_bal30, t_noise30, t_nbal30
  const float* e_deq;               // [64]
  const float* q_div;               // [31]
  const float* q_div2;              // [31]
  const float* q_div_left;          // [31][13]
  const float* q_div_right;
  const float* q_div2_left;
  const float* q_div2_right;
  const float* e_pan;               // [25]
  const uint8_t* find_bands;        // [2 warp][7 bands][65 a0][65 a1]   FBT.find_bands, evaluated on the host
  const float* init_power;          // [64 bands][65 a0][65 a1]          FBT.find_initial_power, evaluated on the host
  const float* qmf_c;               // [640]
  const float* dct4_tab;            // [192]
  const float* w_real;              // [16]
  const float* w_imag;              // [16]
  const float* noise_table;         // [512][2]
};

// One SBR element stream inside a batch: K3 walks `count` frames starting at run_frames[first].
struct SbrRunDev {
  int32_t stream_slot;
  uint32_t first, count;   // into the batch's run_frames (same order as the K2 run of the stream)
  uint32_t sbr_base;       // index of this run's first SbrFrameDev pair
  uint8_t element;         // SBR element index inside the stream (0 or 1)
  uint8_t stereo;          // CPE
  uint8_t sr_index;        // output sampling-frequency index (FBT tables)
  uint8_t first_ch;        // channel slot of the element's first channel
};

}  // namespace jaadb
