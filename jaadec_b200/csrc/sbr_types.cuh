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
  // Channel.E_orig / Q_div / Q_div2 (sbr/Channel.java): float arrays NoiseEnvelope.dequantChannel fills for the bands and
  // envelopes of the current frame only and nothing ever clears.  HFAdjustment.calculate_gain walks them with band counters
  // that run past the current tables when the limiter table is older than the band tables (a header change in a frame
  // that failed: tables rebuilt, patches / limiter bands not) -- and then reads what an earlier frame left there.
  float E_orig[kSbrMaxLE][64];
  float Q_div[2][8], Q_div2[2][8];
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

// Parametric stereo, parse side (ps/EnvData.java + Envelope.java for IID and ICC; ps/PSImpl.java:21-36)
struct PsParamDev {
  int8_t mode;            // -1: disabled (null)
  uint8_t dt[5];
  int8_t first[34];       // indices of the last envelope of the previous frame
  int8_t index[5][34];
};
// IPD / OPD (ps/PDData.java, PDMode.java): 17 entries, stride 1, indices modulo 8
struct PsPdDev {
  int8_t mode;            // PDMode id = the IID mode's id, -1: null
  uint8_t dt[5];
  int8_t first[17];
  int8_t index[5][17];
};
struct PsParseDev {
  PsParamDev iid, icc;
  PsPdDev ipd, opd;       // the IPD/OPD extension (ps/Extension.java, ExtData.java)
  uint8_t opened;         // SBR1.ps != null
  uint8_t var_borders, num_env, data_available, header_read;
  uint8_t ext_enabled;    // Extension.enabled (enable_ext of the PS header)
  uint8_t ext_has_data;   // Extension.data != null: created when a header first enables the extension, never dropped
  uint8_t ext_data_enabled;   // ExtData.enabled (enable_ipdopd of the last ps_extension read)
  uint8_t border_position[6];
};

// One frame of PS parameters as K4 consumes them (after PSImpl.ps_data_decode).
struct __align__(16) PsFrameDev {
  uint8_t use_ps;         // SBR1.isPSUsed() for this frame
  uint8_t num_env;
  uint8_t border[6];
  int8_t iid_mode, icc_mode;   // as EnvData.mode() resolves them (IID null -> 0, ICC null -> 1)
  int8_t iid[5][20], icc[5][20];
  uint8_t nr_ipdopd_par;  // Extension.nr_par(): parameter bands below it get the IPD/OPD phase rotation (0: extension off)
  uint8_t enable_ipdopd;  // ExtData.enabled of this frame (parity tap only; the rotation runs without it, as in JAAD)
  uint8_t pad[12];
  int8_t ipd[5][17];      // IPD indices as PSImpl.ps_mix_phase reads them (opd_index is read from the same array, A-15)
  uint8_t pad2[11];
};
static_assert(sizeof(PsFrameDev) == 320, "PsFrameDev layout (mirrored by jaadec_b200/engine.py PS_FRAME_DTYPE)");

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
  uint8_t tag_valid, tag;   // element_instance_tag of the element object this state belongs to (StreamState::tags)
  uint8_t dequant;          // this frame's sbr_data got as far as NoiseEnvelope.dequantChannel / unmap (K3, per frame)
  SbrChanParse ch[2];
  PsParseDev ps;          // mono element of an SBR+PS stream
  uint8_t pad2[2];
};
static_assert(sizeof(SbrElemDev) % 4 == 0, "SbrElemDev is copied word-wise");

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
  // frame-parallel K4: a frame is "processed" when frame_status == 0 and mode != 0 (it runs the QMF banks and moves the
  // channel's state); the others leave the state alone
  uint32_t ord;                     // processed frames of this run (in this batch) before this one
  uint32_t back;                    // distance (in run frames) to the previous processed frame, 0: none in this batch
  uint32_t fwd;                     // distance to the next processed frame, 0: none in this batch
  uint32_t back_ps, fwd_ps;         // the same over frames that run the parametric-stereo tool (use_ps)
  uint8_t pad[12];
};
static_assert(sizeof(SbrFrameDev) == 1888, "SbrFrameDev layout (mirrored by jaadec_b200/engine.py SBR_FRAME_DTYPE)");

// Process-side persistent state of the parametric-stereo tool of one stream (ps/PSImpl.java:39-62, ps/Filterbank.java).
// Delay lines are stored per band so that the thread that owns a band touches one contiguous piece.
struct __align__(16) PsChanDev {
  float hyb_buffer[3][12][2];        // hybrid analysis history of QMF bands 0..2
  float delay_qmf[64][14][2];        // [band][slot]: 14-slot delay (bands > 22) or the 2-slot all-pass input delay
  float delay_qmf_ser[64][3][5][2];  // [band][link][slot]
  float delay_sub[12][2][2];         // hybrid sub-bands
  float delay_sub_ser[12][3][5][2];
  float P_PeakDecayNrg[20], P_prev[20], P_SmoothPeakDecayDiffNrg_prev[20];
  float h_prev[22][8];               // h11, h12, h21, h22 per group: real parts, then imaginary parts (IPD/OPD rotation)
  // PDData.prev (ps/PDData.java:13) of ipd AND opd: ps_mix_phase stores the same value into both at the same place
  // (opd_index is read from ipd, ps/PSImpl.java:503-504), so the two arrays are always equal and one copy serves
  float pd_prev[20][2][2];
  int32_t phase_hist;
  int32_t pad3[3];
  float syn_v_right[2][9][128];      // right channel: the 9 most recent synthesis v-vectors ([.][0] = newest), double
                                     // buffered like SbrChanDev::syn_v
  int32_t saved_delay, delay_buf_index_ser[3];
  int32_t delay_buf_index_delay[64];
  int32_t v_sel, v_flip;
  int32_t pad[2];
};
static_assert(sizeof(PsChanDev) % 16 == 0, "PsChanDev alignment");

// Process-side persistent state of one SBR channel.
struct __align__(16) SbrChanDev {
  float ana_hist[288];              // the last 288 core samples (QMF analysis ring of sbr/AnalysisFilterbank.java)
  float xsbr[kSbrHfGen][64][2];     // Xsbr rows 0..7 (the last 8 slots of the previous frame, SBR.sbr_save_matrix)
  // the 9 most recent synthesis v-vectors ([0] = newest; sbr/SynthesisFilterbank64.java), double buffered: the synthesis
  // kernel reads [v_sel] for the first frame of a tile while another CTA of the same launch stores the tile's last nine
  // into [v_sel ^ 1]; k4_commit_kernel flips v_sel afterwards
  float syn_v[2][9][128];
  float G_temp_prev[5][64], Q_temp_prev[5][64];
  // Channel.E_curr (sbr/Channel.java): a scratch array in JAAD, but one that is never cleared -- a limiter table that
  // outlived a header change (the frame that carried the header failed, so the tables were rebuilt for the new kx / M while
  // the patches and the limiter bands were not) makes calculate_gain read entries at or above M that an earlier frame wrote
  float E_curr[kSbrMaxLE][64];
  float bwArray_prev[8];
  uint8_t bs_invf_mode_prev[8];
  int32_t GQ_ringbuf_index, index_noise_prev, psi_is_prev;
  int32_t v_sel, v_flip;
  int32_t pad[3];
};
static_assert(sizeof(SbrChanDev) % 16 == 0, "SbrChanDev is copied with 16-byte accesses");

// Read-only tables of the SBR tool (engine-owned device memory).
struct SbrTablesDev {
  const int16_t* huff[10];          // t_env15, f_env15, t_bal15, f_bal15, t_env30, f_env30, t_bal30, f_bal30, t_noise30, t_nbal30
  const uint32_t* huff_lut;         // [10][256] first-eight-bits tables of the same trees (k3_sbr_parse.cuh huff_decode)
  const uint32_t* ps_huff_lut;      // [10][256]
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
  const float* qmf32_tw;            // [32][2]  SynthesisFilterbank32.qmf32_pre_twiddle
  const float* dct4_tab;            // [192]
  const float* w_real;              // [16]
  const float* w_imag;              // [16]
  const float* noise_table;         // [512][2]
  // parametric stereo
  const int16_t* ps_huff[10];       // f_iid_def, t_iid_def, f_iid_fine, t_iid_fine, f_icc, t_icc, f_ipd, t_ipd, f_opd, t_opd
  const float* ps_ipdopd_cos;       // [9]
  const float* ps_ipdopd_sin;       // [9]
  const float* ps_filter_a;         // [3]
  const float* ps_phi_qmf;          // [64][2]
  const float* ps_phi_sub;          // [12][2]
  const float* ps_q_qmf;            // [64][3][2]
  const float* ps_q_sub;            // [12][3][2]
  const float* ps_cos_alphas;       // [8]
  const float* ps_sin_alphas;
  const float* ps_cos_betas[2];     // normal [8][8], fine [16][8]
  const float* ps_sin_betas[2];
  const float* ps_cos_gammas[2];    // as IIDTables names them (the reference swaps sin/cos here, IIDMode.java:16-28)
  const float* ps_sin_gammas[2];
  const float* ps_sincos_alphas_b[2];  // [15][8], [31][8]
  const float* ps_sf_iid[2];        // [15], [31]
  const float* ps_p8;               // [7] Filter8 prototype
  const float* ps_p2;               // [7] Filter2 prototype
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
  uint8_t ps;              // the element may carry parametric stereo (mono SBR+PS stream)
  uint8_t pad[3];
  uint32_t ps_base;        // index of this run's first PsFrameDev
};

}  // namespace jaadb
