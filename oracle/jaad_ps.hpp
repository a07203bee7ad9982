// ORACLE -- TEST INFRASTRUCTURE ONLY (see jaad_bits.hpp for the full notice).
//
// CPU restatement of JAAD's parametric-stereo tool (aac/src/main/java/net/sourceforge/jaad/aac/ps/): ps_data parse
// (IID / ICC modes, envelopes, Huffman delta coding), hybrid analysis / synthesis filterbank, all-pass decorrelator
// with transient ducking and the H-matrix mixing.  Same operation order and the same quirks (SURVEY.md A-12..A-16,
// A-31); tables come verbatim from the reference through tools/extract_tables_sbr.py.
// The IPD/OPD extension (Extension / ExtData / PDData / PDMode and the phase rotation in ps_mix_phase) is restated with
// its quirks.  The 34-band configuration is dead code in the reference (FBType.max returns the
// smaller type, ps/FBType.java:17-19), so only the 20-band hybrid structure exists here.
// File:line references are relative to aac/src/main/java/net/sourceforge/jaad/aac/ps/.
#pragma once
#include "jaad_sbr.hpp"

namespace jaad {
namespace ps {

namespace T = ::jaad_tables;
using sbr::Cpx;

static const int MAX_PS_ENVELOPES = 5, NO_ALLPASS_LINKS = 3, NEGATE_IPD_MASK = 0x1000;
static const int NUM_GROUPS = 10 + 12, NUM_HYBRID_GROUPS = 10, NR_PAR_BANDS = 20, DECAY_CUTOFF = 3;   // FBType.T20 (:29-33)
static const int group_border20[23] = {6, 7, 0, 1, 2, 3, 9, 8, 10, 11, 3, 4, 5, 6, 7, 8, 9, 11, 14, 18, 23, 35, 64};   // PSTables.java:26-31
static const int map_group2bk20[22] = {NEGATE_IPD_MASK | 1, NEGATE_IPD_MASK | 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 14, 15, 16, 17, 18, 19};

inline int bk_of(int gr) { return map_group2bk20[gr] & ~NEGATE_IPD_MASK; }                                    // FBType.bk (:63-65)
inline int maxsb_of(int gr) { return (gr < NUM_HYBRID_GROUPS) ? group_border20[gr] + 1 : group_border20[gr + 1]; }  // FBType.maxsb (:67-69)

inline int huffRead(BitStream& ld, const int16_t* t) {  // Huffman.java:264-276
  int index = 0;
  while (index >= 0) {
    int bit = ld.readBit();
    index = t[index * 2 + bit];
  }
  return index + 31;
}

// EnvData / Envelope / ICData (+ IIDMode / ICCMode): one parameter set (IID or ICC)
struct ParamData {
  bool isIcc;
  int first[34];
  int index[MAX_PS_ENVELOPES][34];
  bool dt[MAX_PS_ENVELOPES];
  int mode = -1;   // -1 = null
  explicit ParamData(bool icc) : isIcc(icc) { memset(first, 0, sizeof first); memset(index, 0, sizeof index); memset(dt, 0, sizeof dt); }

  static int nrPar(int id) { static const int n[6] = {10, 20, 34, 10, 20, 34}; return n[id]; }
  static int stride(int id) { return (id % 3) == 0 ? 2 : 0; }                      // ICMode.stride (:19-21)
  int numSteps() const { int id = mode < 0 ? 0 : mode; return id < 3 ? 7 : 15; }   // IIDMode tables (:16-28)
  int clip(int idx) const {
    if (isIcc) return std::min(std::max(idx, 0), 7);                                // ICCMode.clip (:25-29)
    const int c = numSteps();
    return std::min(std::max(idx, -c), c);                                         // IIDMode.clip (:42-45)
  }
  const int16_t* table(bool dtFlag) const {
    if (isIcc) return dtFlag ? T::PS_T_HUFF_ICC : T::PS_F_HUFF_ICC;
    if (mode < 3) return dtFlag ? T::PS_T_HUFF_IID_DEF : T::PS_F_HUFF_IID_DEF;
    return dtFlag ? T::PS_T_HUFF_IID_FINE : T::PS_F_HUFF_IID_FINE;
  }
  const int* prevOf(int env) const { return env == 0 ? first : index[env - 1]; }

  void readMode(BitStream& ld) {  // ICData.readMode (:21-32)
    bool enabled = ld.readBool();
    if (enabled) {
      int id = ld.readBits(3);
      if (id > 5) throw AACException(ST_ARRAY_BOUNDS, "PS mode index out of bounds");
      mode = id;
    } else mode = -1;
  }
  void readData(BitStream& ld, int num_env) {  // EnvData.readData (:44-49) + Envelope.read (:24-30)
    if (mode < 0) return;
    for (int n = 0; n < num_env; n++) {
      dt[n] = ld.readBool();
      const int16_t* h = table(dt[n]);
      for (int i = 0; i < nrPar(mode); i++) index[n][i] = huffRead(ld, h);
    }
  }
  void resetEnv(int env) { dt[env] = false; memset(index[env], 0, sizeof index[env]); }
  void restoreEnv(int env) { memcpy(index[env], prevOf(env), sizeof index[env]); }
  void decodeEnv(int env) {  // Envelope.decode (:45-74)
    if (mode < 0) { resetEnv(env); return; }
    const int st = stride(mode), nr = nrPar(mode);
    int* ix = index[env];
    const int* prev = prevOf(env);
    if (dt[env]) {
      for (int i = 0; i < nr; i++) {
        int p = prev[i * st];
        ix[i] = clip(p + ix[i]);
      }
    } else {
      int p = ix[0];
      for (int i = 1; i < nr; i++) {
        p = clip(p + ix[i]);
        ix[i] = p;
      }
    }
    if (st > 1)
      for (int i = st * nr - 1; i > 0; --i) ix[i] = ix[i / st];
  }
  void decode(int num_env) {  // EnvData.decode (:51-62)
    if (num_env == 0) {
      if (mode >= 0) restoreEnv(0);
      else resetEnv(0);
    } else {
      for (int env = 0; env < num_env; env++) decodeEnv(env);
    }
  }
  void update(int num_env) {  // :64-69
    if (num_env == 0) memset(first, 0, sizeof first);
    else memcpy(first, index[num_env - 1], sizeof first);
  }
};

// PDData (PDData.java, PDMode.java): one phase parameter set (IPD or OPD), 17 entries, stride 1, indices modulo 8
struct PdData {
  bool isOpd;
  int first[17];
  int index[MAX_PS_ENVELOPES][17];
  bool dt[MAX_PS_ENVELOPES];
  int mode = -1;             // PDMode id = the IID mode's id (PDData.setMode, :19-21); -1 = null
  float prev[20][2][2];      // PDData.prev (:13)
  explicit PdData(bool opd) : isOpd(opd) {
    memset(first, 0, sizeof first); memset(index, 0, sizeof index); memset(dt, 0, sizeof dt); memset(prev, 0, sizeof prev);
  }
  static int nrPar(int id) { static const int n[6] = {5, 11, 17, 5, 11, 17}; return n[id]; }   // PDMode.java:50-57
  const int16_t* table(bool dtFlag) const {
    if (isOpd) return dtFlag ? T::PS_T_HUFF_OPD : T::PS_F_HUFF_OPD;
    return dtFlag ? T::PS_T_HUFF_IPD : T::PS_F_HUFF_IPD;
  }
  const int* prevOf(int env) const { return env == 0 ? first : index[env - 1]; }
  void readData(BitStream& ld, int num_env) {  // EnvData.readData (:44-49)
    if (mode < 0) return;
    for (int n = 0; n < num_env; n++) {
      dt[n] = ld.readBool();
      const int16_t* h = table(dt[n]);
      for (int i = 0; i < nrPar(mode); i++) index[n][i] = huffRead(ld, h);
    }
  }
  void decodeEnv(int env) {  // Envelope.decode (:45-74) with PDMode.stride() = 1, clip = idx & 7
    if (mode < 0) { dt[env] = false; memset(index[env], 0, sizeof index[env]); return; }
    const int nr = nrPar(mode);
    int* ix = index[env];
    const int* pv = prevOf(env);
    if (dt[env]) {
      for (int i = 0; i < nr; i++) ix[i] = (pv[i] + ix[i]) & 7;
    } else {
      int p = ix[0];
      for (int i = 1; i < nr; i++) { p = (p + ix[i]) & 7; ix[i] = p; }
    }
  }
  void decode(int num_env) {  // EnvData.decode (:51-62)
    if (num_env == 0) {
      if (mode >= 0) memcpy(index[0], first, sizeof first);
      else { dt[0] = false; memset(index[0], 0, sizeof index[0]); }
    } else {
      for (int env = 0; env < num_env; env++) decodeEnv(env);
    }
  }
  void update(int num_env) {  // :64-69
    if (num_env == 0) memset(first, 0, sizeof first);
    else memcpy(first, index[num_env - 1], sizeof first);
  }
};

// Extension.java + ExtData.java: the IPD/OPD extension of ps_data
struct PsExtension {
  bool enabled = false;       // Extension.enabled: enable_ext of the PS header
  bool has_data = false;      // Extension.data != null (created the first time a header enables the extension, never dropped)
  bool data_enabled = false;  // ExtData.enabled: enable_ipdopd of the last ps_extension(0) read
  PdData ipd{false}, opd{true};

  void readMode(BitStream& ld, int iidMode) {  // Extension.readMode (:31-38)
    enabled = ld.readBool();
    if (enabled) has_data = true;
    if (has_data) ipd.mode = opd.mode = enabled ? iidMode : -1;   // ExtData.setMode(enabled ? parent.mode : null)
  }
  void readData(BitStream& ld0, int num_env) {  // Extension.readData (:40-59)
    if (!enabled) return;
    int cnt = ld0.readBits(4);
    if (cnt == 15) cnt += ld0.readBits(8);
    BitStream ld = ld0.readSubStream(8 * cnt);
    while (ld.getBitsLeft() > 7) {
      const int ps_extension_id = ld.readBits(2);
      if (ps_extension_id == 0 && has_data) {
        // ExtData.readData (ExtData.java:17-25)
        data_enabled = ld.readBool();
        if (data_enabled) { ipd.readData(ld, num_env); opd.readData(ld, num_env); }
        ld.readBit();
      }
    }
  }
  void decode(int num_env) { if (enabled && has_data && data_enabled) { ipd.decode(num_env); opd.decode(num_env); } }   // :61-64, ExtData :27-32
  void update(int num_env) { if (enabled && has_data) { ipd.update(num_env); opd.update(num_env); } }                  // :66-69, ExtData :34-37
  void restore(int num_env) { update(num_env); }   // ExtData.restore calls update (ExtData.java:39-42)
  // Extension.nr_par (:81-86) -> ExtData.nr_par (ExtData.java:49-54).  -1: JAAD dereferences the null PDMode of an extension
  // that was enabled while IID is off (NullPointerException)
  int nr_par() const {
    if (!(enabled && has_data)) return 0;
    if (ipd.mode < 0) return -1;
    return std::max(PdData::nrPar(ipd.mode), 11);
  }
};

// Filterbank.java + Filter2.java + Filter8.java
struct HybridFilterbank {
  static const int len = 32;
  float buffer[5][12][2];
  HybridFilterbank() { memset(buffer, 0, sizeof buffer); }

  static void DCT3_4_unscaled(float* y, const float* x) {  // Filter8.java:124-138
    float f0 = (x[2] * 0.7071067811865476f);
    float f1 = x[0] - f0;
    float f2 = x[0] + f0;
    float f3 = x[1] + x[3];
    float f4 = (x[1] * 1.3065629648763766f);
    float f5 = (f3 * (-0.9238795325112866f));
    float f6 = (x[3] * (-0.5411961001461967f));
    float f7 = f4 + f5;
    float f8 = f6 - f5;
    y[3] = f2 - f8;
    y[0] = f2 + f8;
    y[2] = f1 - f7;
    y[1] = f1 + f7;
  }
  static void filter8(const float (*b)[2], float (*result)[12][2]) {  // Filter8.java:53-122
    const float* filter = JT(PS_P8_13_20);
    float input_re1[4], input_re2[4], input_im1[4], input_im2[4], x[4], y[4];
    for (int i = 0; i < len; i++) {
      input_re1[0] = (filter[6] * b[6 + i][0]);
      input_re1[1] = (filter[5] * (b[5 + i][0] + b[7 + i][0]));
      input_re1[2] = -(filter[0] * (b[0 + i][0] + b[12 + i][0])) + (filter[4] * (b[4 + i][0] + b[8 + i][0]));
      input_re1[3] = -(filter[1] * (b[1 + i][0] + b[11 + i][0])) + (filter[3] * (b[3 + i][0] + b[9 + i][0]));
      input_im1[0] = (filter[5] * (b[7 + i][1] - b[5 + i][1]));
      input_im1[1] = (filter[0] * (b[12 + i][1] - b[0 + i][1])) + (filter[4] * (b[8 + i][1] - b[4 + i][1]));
      input_im1[2] = (filter[1] * (b[11 + i][1] - b[1 + i][1])) + (filter[3] * (b[9 + i][1] - b[3 + i][1]));
      input_im1[3] = (filter[2] * (b[10 + i][1] - b[2 + i][1]));
      for (int n = 0; n < 4; n++) x[n] = input_re1[n] - input_im1[3 - n];
      DCT3_4_unscaled(y, x);
      result[i][7][0] = y[0]; result[i][5][0] = y[2]; result[i][3][0] = y[3]; result[i][1][0] = y[1];
      for (int n = 0; n < 4; n++) x[n] = input_re1[n] + input_im1[3 - n];
      DCT3_4_unscaled(y, x);
      result[i][6][0] = y[1]; result[i][4][0] = y[3]; result[i][2][0] = y[2]; result[i][0][0] = y[0];
      input_im2[0] = (filter[6] * b[6 + i][1]);
      input_im2[1] = (filter[5] * (b[5 + i][1] + b[7 + i][1]));
      input_im2[2] = -(filter[0] * (b[0 + i][1] + b[12 + i][1])) + (filter[4] * (b[4 + i][1] + b[8 + i][1]));
      input_im2[3] = -(filter[1] * (b[1 + i][1] + b[11 + i][1])) + (filter[3] * (b[3 + i][1] + b[9 + i][1]));
      input_re2[0] = (filter[5] * (b[7 + i][0] - b[5 + i][0]));
      input_re2[1] = (filter[0] * (b[12 + i][0] - b[0 + i][0])) + (filter[4] * (b[8 + i][0] - b[4 + i][0]));
      input_re2[2] = (filter[1] * (b[11 + i][0] - b[1 + i][0])) + (filter[3] * (b[9 + i][0] - b[3 + i][0]));
      input_re2[3] = (filter[2] * (b[10 + i][0] - b[2 + i][0]));
      for (int n = 0; n < 4; n++) x[n] = input_im2[n] + input_re2[3 - n];
      DCT3_4_unscaled(y, x);
      result[i][7][1] = y[0]; result[i][5][1] = y[2]; result[i][3][1] = y[3]; result[i][1][1] = y[1];
      for (int n = 0; n < 4; n++) x[n] = input_im2[n] - input_re2[3 - n];
      DCT3_4_unscaled(y, x);
      result[i][6][1] = y[1]; result[i][4][1] = y[3]; result[i][2][1] = y[2]; result[i][0][1] = y[0];
    }
  }
  static void filter2(const float (*b)[2], float (*result)[12][2]) {  // Filter2.java:40-68
    const float* filter = JT(PS_P2_13_20);
    for (int i = 0; i < len; i++) {
      float r0 = (filter[0] * (b[0 + i][0] + b[12 + i][0]));
      float r1 = (filter[1] * (b[1 + i][0] + b[11 + i][0]));
      float r2 = (filter[2] * (b[2 + i][0] + b[10 + i][0]));
      float r3 = (filter[3] * (b[3 + i][0] + b[9 + i][0]));
      float r4 = (filter[4] * (b[4 + i][0] + b[8 + i][0]));
      float r5 = (filter[5] * (b[5 + i][0] + b[7 + i][0]));
      float r6 = (filter[6] * b[6 + i][0]);
      float i0 = (filter[0] * (b[0 + i][1] + b[12 + i][1]));
      float i1 = (filter[1] * (b[1 + i][1] + b[11 + i][1]));
      float i2 = (filter[2] * (b[2 + i][1] + b[10 + i][1]));
      float i3 = (filter[3] * (b[3 + i][1] + b[9 + i][1]));
      float i4 = (filter[4] * (b[4 + i][1] + b[8 + i][1]));
      float i5 = (filter[5] * (b[5 + i][1] + b[7 + i][1]));
      float i6 = (filter[6] * b[6 + i][1]);
      result[i][0][0] = r0 + r1 + r2 + r3 + r4 + r5 + r6;
      result[i][0][1] = i0 + i1 + i2 + i3 + i4 + i5 + i6;
      result[i][1][0] = r0 - r1 + r2 - r3 + r4 - r5 + r6;
      result[i][1][1] = i0 - i1 + i2 - i3 + i4 - i5 + i6;
    }
  }

  void hybrid_analysis(Cpx (*X)[64], float (*X_hybrid)[32][2]) {  // Filterbank.java:18-68
    float work[len + 12][2];
    float temp[len][12][2];
    for (int band = 0, offset = 0; band < DECAY_CUTOFF; band++) {
      for (int i = 0; i < 12; i++) { work[i][0] = buffer[band][i][0]; work[i][1] = buffer[band][i][1]; }
      for (int n = 0; n < len; n++) { work[12 + n][0] = X[n + 6][band][0]; work[12 + n][1] = X[n + 6][band][1]; }
      for (int i = 0; i < 12; i++) { buffer[band][i][0] = work[len + i][0]; buffer[band][i][1] = work[len + i][1]; }
      const int resolution = band == 0 ? 8 : 2;
      if (band == 0) filter8(work, temp);
      else filter2(work, temp);
      for (int n = 0; n < len; n++)
        for (int k = 0; k < resolution; k++) { X_hybrid[n][offset + k][0] = temp[n][k][0]; X_hybrid[n][offset + k][1] = temp[n][k][1]; }
      offset += resolution;
    }
    for (int n = 0; n < len; n++) {
      X_hybrid[n][3][0] += X_hybrid[n][4][0];
      X_hybrid[n][3][1] += X_hybrid[n][4][1];
      X_hybrid[n][4][0] = 0;
      X_hybrid[n][4][1] = 0;
      X_hybrid[n][2][0] += X_hybrid[n][5][0];
      X_hybrid[n][2][1] += X_hybrid[n][5][1];
      X_hybrid[n][5][0] = 0;
      X_hybrid[n][5][1] = 0;
    }
  }
  void hybrid_synthesis(Cpx (*X)[64], float (*X_hybrid)[32][2]) {  // :70-86
    for (int band = 0, offset = 0; band < DECAY_CUTOFF; band++) {
      const int resolution = band == 0 ? 8 : 2;
      for (int n = 0; n < len; n++) {
        X[n][band][0] = 0;
        X[n][band][1] = 0;
        for (int k = 0; k < resolution; k++) {
          X[n][band][0] += X_hybrid[n][offset + k][0];
          X[n][band][1] += X_hybrid[n][offset + k][1];
        }
      }
      offset += resolution;
    }
  }
};

// PSImpl.java
struct PSImpl : sbr::PSBase {
  ParamData iid{false}, icc{true};
  PsExtension ext;
  bool var_borders = false;
  int num_env = 0;
  int border_position[MAX_PS_ENVELOPES + 1] = {0};
  bool ps_data_available = false, header_read = false;
  HybridFilterbank fb;
  static const int NR_ALLPASS_BANDS = 22, SHORT_DELAY_BAND = 35;
  int saved_delay = 0;
  int delay_buf_index_ser[NO_ALLPASS_LINKS] = {0, 0, 0};
  int num_sample_delay_ser[NO_ALLPASS_LINKS] = {3, 4, 5};   // delay_length_d (PSTables.java:56)
  int delay_D[64];
  int delay_buf_index_delay[64];
  float delay_Qmf[14][64][2];
  float delay_SubQmf[2][32][2];
  float delay_Qmf_ser[NO_ALLPASS_LINKS][5][64][2];
  float delay_SubQmf_ser[NO_ALLPASS_LINKS][5][32][2];
  float P_PeakDecayNrg[34], P_prev[34], P_SmoothPeakDecayDiffNrg_prev[34];
  float h11_prev[50][2], h12_prev[50][2], h21_prev[50][2], h22_prev[50][2];
  int phase_hist = 0;

  PSImpl() {  // :64-94
    memset(delay_buf_index_delay, 0, sizeof delay_buf_index_delay);
    for (int i = 0; i < 64; i++) delay_D[i] = i < SHORT_DELAY_BAND ? 14 : 1;
    memset(delay_Qmf, 0, sizeof delay_Qmf); memset(delay_SubQmf, 0, sizeof delay_SubQmf);
    memset(delay_Qmf_ser, 0, sizeof delay_Qmf_ser); memset(delay_SubQmf_ser, 0, sizeof delay_SubQmf_ser);
    memset(P_PeakDecayNrg, 0, sizeof P_PeakDecayNrg); memset(P_prev, 0, sizeof P_prev);
    memset(P_SmoothPeakDecayDiffNrg_prev, 0, sizeof P_SmoothPeakDecayDiffNrg_prev);
    memset(h11_prev, 0, sizeof h11_prev); memset(h12_prev, 0, sizeof h12_prev);
    memset(h21_prev, 0, sizeof h21_prev); memset(h22_prev, 0, sizeof h22_prev);
    for (int i = 0; i < 50; i++) { h11_prev[i][0] = 1; h12_prev[i][1] = 1; }   // the constructor sets these two, twice (A-14)
  }

  bool isDataAvailable() const override { return ps_data_available; }

  void decode(BitStream& ld) override {  // :103-135
    if (ld.readBool()) {
      header_read = true;
      iid.readMode(ld);
      icc.readMode(ld);
      ext.readMode(ld, iid.mode);
    }
    var_borders = ld.readBit() != 0;
    int tmp = ld.readBits(2);
    static const int num_env_tab[2][4] = {{0, 1, 2, 4}, {1, 2, 3, 4}};
    num_env = num_env_tab[var_borders ? 1 : 0][tmp];
    if (var_borders)
      for (int n = 1; n < num_env + 1; n++) border_position[n] = ld.readBits(5) + 1;
    iid.readData(ld, num_env);
    icc.readData(ld, num_env);
    ext.readData(ld, num_env);
    ps_data_available = true;
  }

  void ps_data_decode() {  // :137-199
    if (!ps_data_available) num_env = 0;
    iid.decode(num_env);
    icc.decode(num_env);
    ext.decode(num_env);
    if (num_env == 0) num_env = 1;
    iid.update(num_env);
    icc.update(num_env);
    ext.update(num_env);
    ps_data_available = false;
    const int L = HybridFilterbank::len;
    if (!var_borders) {
      border_position[0] = 0;
      for (int env = 1; env < num_env; env++) border_position[env] = (env * L) / num_env;
      border_position[num_env] = L;
    } else {
      border_position[0] = 0;
      if (border_position[num_env] < L) {
        iid.restoreEnv(num_env);
        icc.restoreEnv(num_env);
        ext.restore(num_env);
        ++num_env;
        border_position[num_env] = L;
      }
      int bpl = border_position[0];
      for (int env = 1; env < num_env; env++) {
        int bp = border_position[env];
        int mx = L - (num_env - env);
        bpl = std::min(std::max(bp, bpl + 1), mx);   // Utils.clip
        if (bpl != bp) border_position[env] = bpl;
      }
    }
  }

  void ps_decorrelate(Cpx (*X_left)[64], Cpx (*X_right)[64], float (*X_hybrid_left)[32][2], float (*X_hybrid_right)[32][2]) {  // :202-400
    static const float ALPHA_DECAY = 0.76592833836465f, ALPHA_SMOOTH = 0.25f, DECAY_SLOPE = 0.05f;
    const float* filter_a = JT(PS_FILTER_A);
    float P[32][34], G_TransientRatio[32][34];
    memset(P, 0, sizeof P);
    memset(G_TransientRatio, 0, sizeof G_TransientRatio);
    for (int gr = 0; gr < NUM_GROUPS; gr++) {
      const int bk = bk_of(gr), maxsb = maxsb_of(gr);
      const bool hyb = gr < NUM_HYBRID_GROUPS;
      for (int n = border_position[0]; n < border_position[num_env]; n++) {
        for (int sb = group_border20[gr]; sb < maxsb; sb++) {
          const float re = hyb ? X_hybrid_left[n][sb][0] : X_left[n][sb][0];
          const float im = hyb ? X_hybrid_left[n][sb][1] : X_left[n][sb][1];
          P[n][bk] += (re * re) + (im * im);
        }
      }
    }
    for (int bk = 0; bk < NR_PAR_BANDS; bk++) {
      for (int n = border_position[0]; n < border_position[num_env]; n++) {
        const float gamma = 1.5f;
        P_PeakDecayNrg[bk] = (P_PeakDecayNrg[bk] * ALPHA_DECAY);
        if (P_PeakDecayNrg[bk] < P[n][bk]) P_PeakDecayNrg[bk] = P[n][bk];
        float P_SmoothPeakDecayDiffNrg = P_SmoothPeakDecayDiffNrg_prev[bk];
        P_SmoothPeakDecayDiffNrg += ((P_PeakDecayNrg[bk] - P[n][bk] - P_SmoothPeakDecayDiffNrg_prev[bk]) * ALPHA_SMOOTH);
        P_SmoothPeakDecayDiffNrg_prev[bk] = P_SmoothPeakDecayDiffNrg;
        float nrg = P_prev[bk];
        nrg += ((P[n][bk] - P_prev[bk]) * ALPHA_SMOOTH);
        P_prev[bk] = nrg;
        if ((P_SmoothPeakDecayDiffNrg * gamma) <= nrg) G_TransientRatio[n][bk] = 1.0f;
        else G_TransientRatio[n][bk] = (nrg / (P_SmoothPeakDecayDiffNrg * gamma));
      }
    }
    int temp_delay = 0;
    int temp_delay_ser[NO_ALLPASS_LINKS] = {0, 0, 0};
    float g_DecaySlope_filt[NO_ALLPASS_LINKS];
    const float (*PhiQmf)[2] = reinterpret_cast<const float (*)[2]>(JT(PS_PHI_FRACT_QMF));
    const float (*PhiSub)[2] = reinterpret_cast<const float (*)[2]>(JT(PS_PHI_FRACT_SUBQMF20));
    const float (*QQmf)[3][2] = reinterpret_cast<const float (*)[3][2]>(JT(PS_Q_FRACT_ALLPASS_QMF));
    const float (*QSub)[3][2] = reinterpret_cast<const float (*)[3][2]>(JT(PS_Q_FRACT_ALLPASS_SUBQMF20));
    for (int gr = 0; gr < NUM_GROUPS; gr++) {
      const int maxsb = maxsb_of(gr);
      const bool hyb = gr < NUM_HYBRID_GROUPS;
      for (int sb = group_border20[gr]; sb < maxsb; sb++) {
        float g_DecaySlope;
        if (hyb || sb <= DECAY_CUTOFF) g_DecaySlope = 1.0f;
        else {
          int decay = DECAY_CUTOFF - sb;
          if (decay <= -20) g_DecaySlope = 0;
          else g_DecaySlope = 1.0f + DECAY_SLOPE * decay;
        }
        for (int m = 0; m < NO_ALLPASS_LINKS; m++) g_DecaySlope_filt[m] = g_DecaySlope * filter_a[m];
        temp_delay = saved_delay;
        for (int n = 0; n < NO_ALLPASS_LINKS; n++) temp_delay_ser[n] = delay_buf_index_ser[n];
        for (int n = border_position[0]; n < border_position[num_env]; n++) {
          float r0Re, r0Im;
          const float re = hyb ? X_hybrid_left[n][sb][0] : X_left[n][sb][0];
          const float im = hyb ? X_hybrid_left[n][sb][1] : X_left[n][sb][1];
          if (sb > NR_ALLPASS_BANDS && !hyb) {
            float* delay = delay_Qmf[delay_buf_index_delay[sb]][sb];
            r0Re = delay[0];
            r0Im = delay[1];
            delay[0] = re;
            delay[1] = im;
          } else {
            float* delayQmf = hyb ? delay_SubQmf[temp_delay][sb] : delay_Qmf[temp_delay][sb];
            const float* Phi_Fract = hyb ? PhiSub[sb] : PhiQmf[sb];
            float tmp0Re = delayQmf[0], tmp0Im = delayQmf[1];
            delayQmf[0] = re;
            delayQmf[1] = im;
            r0Re = (tmp0Re * Phi_Fract[0]) + (tmp0Im * Phi_Fract[1]);
            r0Im = (tmp0Im * Phi_Fract[0]) - (tmp0Re * Phi_Fract[1]);
            for (int m = 0; m < NO_ALLPASS_LINKS; m++) {
              const float* qFractAllpass = hyb ? QSub[sb][m] : QQmf[sb][m];
              float* delay = hyb ? delay_SubQmf_ser[m][temp_delay_ser[m]][sb] : delay_Qmf_ser[m][temp_delay_ser[m]][sb];
              tmp0Re = delay[0];
              tmp0Im = delay[1];
              float tmpRe = (tmp0Re * qFractAllpass[0]) + (tmp0Im * qFractAllpass[1]);
              float tmpIm = (tmp0Im * qFractAllpass[0]) - (tmp0Re * qFractAllpass[1]);
              tmpRe -= g_DecaySlope_filt[m] * r0Re;
              tmpIm -= g_DecaySlope_filt[m] * r0Im;
              delay[0] = r0Re + (g_DecaySlope_filt[m] * tmpRe);
              delay[1] = r0Im + (g_DecaySlope_filt[m] * tmpIm);
              r0Re = tmpRe;
              r0Im = tmpIm;
            }
          }
          const int bk = bk_of(gr);
          if (hyb) { X_hybrid_right[n][sb][0] = (G_TransientRatio[n][bk] * r0Re); X_hybrid_right[n][sb][1] = (G_TransientRatio[n][bk] * r0Im); }
          else { X_right[n][sb][0] = (G_TransientRatio[n][bk] * r0Re); X_right[n][sb][1] = (G_TransientRatio[n][bk] * r0Im); }
          if (++temp_delay >= 2) temp_delay = 0;
          if (sb > NR_ALLPASS_BANDS && !hyb) {
            if (++delay_buf_index_delay[sb] >= delay_D[sb]) delay_buf_index_delay[sb] = 0;
          }
          for (int m = 0; m < NO_ALLPASS_LINKS; m++)
            if (++temp_delay_ser[m] >= num_sample_delay_ser[m]) temp_delay_ser[m] = 0;
        }
      }
    }
    saved_delay = temp_delay;
    for (int m = 0; m < NO_ALLPASS_LINKS; m++) delay_buf_index_ser[m] = temp_delay_ser[m];
  }

  void ps_mix_phase(Cpx (*X_left)[64], Cpx (*X_right)[64], float (*X_hybrid_left)[32][2], float (*X_hybrid_right)[32][2]) {  // :406-681
    static const float COEF_SQRT2 = 1.4142135623731f;
    const int iidMode = iid.mode < 0 ? 0 : iid.mode;   // EnvData.mode(): null -> mode(0)
    const int iccMode = icc.mode < 0 ? 1 : icc.mode;   // ICCData.mode(): null -> mode(1)
    const bool fine = iidMode >= 3;
    const int num_steps = fine ? 15 : 7;
    const float* sf_iid = fine ? JT(PS_SF_IID_FINE) : JT(PS_SF_IID_NORMAL);
    const float* cos_betas = fine ? JT(PS_COS_BETAS_FINE) : JT(PS_COS_BETAS_NORMAL);
    const float* sin_betas = fine ? JT(PS_SIN_BETAS_FINE) : JT(PS_SIN_BETAS_NORMAL);
    // IIDMode hands (sin_gammas, cos_gammas) to constructor parameters named (cos_gammas, sin_gammas) (IIDMode.java:16-28)
    const float* tab_cos_gammas = fine ? JT(PS_SIN_GAMMAS_FINE) : JT(PS_SIN_GAMMAS_NORMAL);
    const float* tab_sin_gammas = fine ? JT(PS_COS_GAMMAS_FINE) : JT(PS_COS_GAMMAS_NORMAL);
    const float* sincos_alphas_b = fine ? JT(PS_SINCOS_ALPHAS_B_FINE) : JT(PS_SINCOS_ALPHAS_B_NORMAL);
    const float* cos_alphas = JT(PS_COS_ALPHAS);
    const float* sin_alphas = JT(PS_SIN_ALPHAS);
    const float* ipdopd_cos_tab = JT(PS_IPDOPD_COS_TAB);
    const float* ipdopd_sin_tab = JT(PS_IPDOPD_SIN_TAB);
    float h11, h12, h21, h22, H11, H12, H21, H22, dH11, dH12, dH21, dH22;
    float h11i = 0, h12i = 0, h21i = 0, h22i = 0, H11i = 0, H12i = 0, H21i = 0, H22i = 0, dH11i = 0, dH12i = 0, dH21i = 0, dH22i = 0;
    const int nr_ipdopd_par = ext.nr_par();
    for (int gr = 0; gr < NUM_GROUPS; gr++) {
      const int bk = bk_of(gr);
      const bool rot = bk < nr_ipdopd_par;
      // FBType.bkm tests `& ~NEGATE_IPD_MASK` instead of `& NEGATE_IPD_MASK` (FBType.java:71-73): true for every bk != 0
      const bool bkm = (map_group2bk20[gr] & ~NEGATE_IPD_MASK) != 0;
      const bool hyb = gr < NUM_HYBRID_GROUPS;
      const int maxsb = hyb ? group_border20[gr] + 1 : group_border20[gr + 1];
      for (int env = 0; env < num_env; env++) {
        int iid_index = iid.index[env][bk];
        const int iid_sign = iid_index < 0 ? -1 : 1;
        iid_index = std::abs(iid_index);
        const int icc_index = icc.index[env][bk];
        if (iid_index > num_steps || icc_index < 0 || icc_index > 7)
          throw AACException(ST_ARRAY_BOUNDS, "PS parameter index out of bounds");
        if (iccMode < 3) {
          const float c_1 = sf_iid[num_steps + iid_index];
          const float c_2 = sf_iid[num_steps - iid_index];
          const float cosa = cos_alphas[icc_index];
          const float sina = sin_alphas[icc_index];
          const float cosb = cos_betas[iid_index * 8 + icc_index];
          const float sinb = sin_betas[iid_index * 8 + icc_index] * (float)iid_sign;
          const float ab1 = (cosb * cosa), ab2 = (sinb * sina), ab3 = (sinb * cosa), ab4 = (cosb * sina);
          h11 = (c_2 * (ab1 - ab2));
          h12 = (c_1 * (ab1 + ab2));
          h21 = (c_2 * (ab3 + ab4));
          h22 = (c_1 * (ab3 - ab4));
        } else {
          const float cosa = sincos_alphas_b[(num_steps + iid_index) * 8 + icc_index];
          const float sina = sincos_alphas_b[(2 * num_steps - (num_steps + iid_index)) * 8 + icc_index];
          const float cosg = tab_cos_gammas[iid_index * 8 + icc_index];
          const float sing = tab_sin_gammas[iid_index * 8 + icc_index];
          h11 = (COEF_SQRT2 * (cosa * cosg));
          h12 = (COEF_SQRT2 * (sina * cosg));
          h21 = (COEF_SQRT2 * (-cosa * sing));
          h22 = (COEF_SQRT2 * (sina * sing));
        }
        if (rot) {
          // phase rotation parameters (:488-560), quirks kept: opd_index is read from ipd, the value "before previous" comes
          // from opd.prev for both, phase_hist moves once per (group, envelope)
          float* ipd_prev = ext.ipd.prev[bk][phase_hist];
          float* opd_prev = ext.opd.prev[bk][phase_hist];
          float tempLeft[2], tempRight[2], phaseLeft[2], phaseRight[2];
          tempLeft[0] = (ipd_prev[0] * 0.25f);
          tempLeft[1] = (ipd_prev[1] * 0.25f);
          tempRight[0] = (opd_prev[0] * 0.25f);
          tempRight[1] = (opd_prev[1] * 0.25f);
          const int ipd_index = std::abs(ext.ipd.index[env][bk]);
          const int opd_index = std::abs(ext.ipd.index[env][bk]);
          if (ipd_index > 8) throw AACException(ST_ARRAY_BOUNDS, "PS phase index out of bounds");
          ipd_prev[0] = ipdopd_cos_tab[ipd_index];
          ipd_prev[1] = ipdopd_sin_tab[ipd_index];
          opd_prev[0] = ipdopd_cos_tab[opd_index];
          opd_prev[1] = ipdopd_sin_tab[opd_index];
          tempLeft[0] += ipd_prev[0];
          tempLeft[1] += ipd_prev[1];
          tempRight[0] += opd_prev[0];
          tempRight[1] += opd_prev[1];
          ++phase_hist;
          phase_hist %= 2;
          ipd_prev = ext.opd.prev[bk][phase_hist];
          opd_prev = ext.opd.prev[bk][phase_hist];
          tempLeft[0] += (ipd_prev[0] * 0.5f);
          tempLeft[1] += (ipd_prev[1] * 0.5f);
          tempRight[0] += (opd_prev[0] * 0.5f);
          tempRight[1] += (opd_prev[1] * 0.5f);
          const float xy = (float)std::sqrt((double)(tempRight[0] * tempRight[0] + tempRight[1] * tempRight[1]));   // magnitude_c (:402-404)
          const float pq = (float)std::sqrt((double)(tempLeft[0] * tempLeft[0] + tempLeft[1] * tempLeft[1]));
          if (xy != 0) { phaseLeft[0] = (tempRight[0] / xy); phaseLeft[1] = (tempRight[1] / xy); }
          else { phaseLeft[0] = 0; phaseLeft[1] = 0; }
          const float xypq = (xy * pq);
          if (xypq != 0) {
            const float tmp1 = (tempRight[0] * tempLeft[0]) + (tempRight[1] * tempLeft[1]);
            const float tmp2 = (tempRight[1] * tempLeft[0]) - (tempRight[0] * tempLeft[1]);
            phaseRight[0] = (tmp1 / xypq);
            phaseRight[1] = (tmp2 / xypq);
          } else { phaseRight[0] = 0; phaseRight[1] = 0; }
          h11i = (h11 * phaseLeft[1]);
          h12i = (h12 * phaseRight[1]);
          h21i = (h21 * phaseLeft[1]);
          h22i = (h22 * phaseRight[1]);
          h11 = (h11 * phaseLeft[0]);
          h12 = (h12 * phaseRight[0]);
          h21 = (h21 * phaseLeft[0]);
          h22 = (h22 * phaseRight[0]);
        }
        const float L = (float)(border_position[env + 1] - border_position[env]);
        dH11 = (h11 - h11_prev[gr][0]) / L;
        dH12 = (h12 - h12_prev[gr][0]) / L;
        dH21 = (h21 - h21_prev[gr][0]) / L;
        dH22 = (h22 - h22_prev[gr][0]) / L;
        H11 = h11_prev[gr][0]; H12 = h12_prev[gr][0]; H21 = h21_prev[gr][0]; H22 = h22_prev[gr][0];
        h11_prev[gr][0] = h11; h12_prev[gr][0] = h12; h21_prev[gr][0] = h21; h22_prev[gr][0] = h22;
        if (rot) {
          dH11i = (h11i - h11_prev[gr][1]) / L;
          dH12i = (h12i - h12_prev[gr][1]) / L;
          dH21i = (h21i - h21_prev[gr][1]) / L;
          dH22i = (h22i - h22_prev[gr][1]) / L;
          H11i = h11_prev[gr][1]; H12i = h12_prev[gr][1]; H21i = h21_prev[gr][1]; H22i = h22_prev[gr][1];
          if (bkm) {
            dH11i = -dH11i; dH12i = -dH12i; dH21i = -dH21i; dH22i = -dH22i;
            H11i = -H11i; H12i = -H12i; H21i = -H21i; H22i = -H22i;
          }
          h11_prev[gr][1] = h11i; h12_prev[gr][1] = h12i; h21_prev[gr][1] = h21i; h22_prev[gr][1] = h22i;
        }
        for (int n = border_position[env]; n < border_position[env + 1]; n++) {
          H11 += dH11; H12 += dH12; H21 += dH21; H22 += dH22;
          if (rot) { H11i += dH11i; H12i += dH12i; H21i += dH21i; H22i += dH22i; }
          for (int sb = group_border20[gr]; sb < maxsb; sb++) {
            float* l = hyb ? X_hybrid_left[n][sb] : X_left[n][sb];
            float* r = hyb ? X_hybrid_right[n][sb] : X_right[n][sb];
            const float inL0 = l[0], inL1 = l[1], inR0 = r[0], inR1 = r[1];
            float tL0 = (H11 * inL0) + (H21 * inR0);
            float tL1 = (H11 * inL1) + (H21 * inR1);
            float tR0 = (H12 * inL0) + (H22 * inR0);
            float tR1 = (H12 * inL1) + (H22 * inR1);
            if (rot) {
              tL0 -= (H11i * inL1) + (H21i * inR1);
              tL1 += (H11i * inL0) + (H21i * inR0);
              tR0 -= (H12i * inL1) + (H22i * inR1);
              tR1 += (H12i * inL0) + (H22i * inR0);
            }
            l[0] = tL0; l[1] = tL1; r[0] = tR0; r[1] = tR1;
          }
        }
      }
    }
  }

  void process(Cpx (*X_left)[64], Cpx (*X_right)[64]) override {  // :685-707
    static thread_local float X_hybrid_left[32][32][2], X_hybrid_right[32][32][2];
    memset(X_hybrid_left, 0, sizeof X_hybrid_left);
    memset(X_hybrid_right, 0, sizeof X_hybrid_right);
    ps_data_decode();
    // JAAD runs into a NullPointerException inside ps_mix_phase (ExtData.nr_par on a null mode) when a header enables the
    // extension while IID is off; reported here before the frame's processing starts
    if (ext.nr_par() < 0) throw AACException(ST_ARRAY_BOUNDS, "PS extension enabled without IID (NullPointerException in JAAD)");
    fb.hybrid_analysis(X_left, X_hybrid_left);
    ps_decorrelate(X_left, X_right, X_hybrid_left, X_hybrid_right);
    ps_mix_phase(X_left, X_right, X_hybrid_left, X_hybrid_right);
    fb.hybrid_synthesis(X_left, X_hybrid_left);
    fb.hybrid_synthesis(X_right, X_hybrid_right);
  }
};

inline sbr::PSBase* makePS(int) { return new PSImpl(); }
struct Registrar { Registrar() { sbr::psFactory() = &makePS; } };
static Registrar g_registrar;

}  // namespace ps
}  // namespace jaad
