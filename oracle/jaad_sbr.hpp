// ORACLE -- TEST INFRASTRUCTURE ONLY (see jaad_bits.hpp for the full notice).
//
// CPU restatement of JAAD's SBR tool (aac/src/main/java/net/sourceforge/jaad/aac/sbr/):
// payload parse (header, grid, dtdf, invf, envelope / noise Huffman), frequency band tables,
// 32-band QMF analysis, HF generation, HF adjustment, the 64-band QMF synthesis and the 32-band (down-sampled) one.
// Same operation order and the same float/double promotion points as the Java code; tables
// come verbatim from the reference through tools/extract_tables_sbr.py.
// File:line references are relative to aac/src/main/java/net/sourceforge/jaad/aac/sbr/.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <vector>

#include "jaad_lc.hpp"
#include "../jaadec_b200/csrc/generated/jaad_dct32.h"   // operation lists of DCT4_32 / DST4_32 (tools/extract_dct32.py)

namespace jaad {
namespace sbr {

namespace T = ::jaad_tables;

static const int MAX_NTSR = 32, MAX_M = 49, MAX_L_E = 5;
static const int NO_TIME_SLOTS = 16, RATE = 2, T_HFGEN = 8, T_HFADJ = 2;
static const int MAX_NTSRHFG = 40;
enum { FIXFIX = 0, FIXVAR = 1, VARFIX = 2, VARVAR = 3 };
enum { LO_RES = 0, HI_RES = 1 };

typedef float Cpx[2];

// Header.java:12-78
struct Header {
  bool bs_amp_res = true;
  int bs_start_freq = 5, bs_stop_freq = 0, bs_xover_band = 0, bs_freq_scale = 2;
  bool bs_alter_scale = true;
  int bs_noise_bands = 2, bs_limiter_bands = 2, bs_limiter_gains = 2;
  bool bs_interpol_freq = false, bs_smoothing_mode = false;

  void decode(BitStream& ld) {
    bs_amp_res = ld.readBool();
    bs_start_freq = ld.readBits(4);
    bs_stop_freq = ld.readBits(4);
    bs_xover_band = ld.readBits(3);
    ld.readBits(2);
    bool extra1 = ld.readBool();
    bool extra2 = ld.readBool();
    if (extra1) {
      bs_freq_scale = ld.readBits(2);
      bs_alter_scale = ld.readBool();
      bs_noise_bands = ld.readBits(2);
    } else { bs_freq_scale = 2; bs_alter_scale = true; bs_noise_bands = 2; }
    if (extra2) {
      bs_limiter_bands = ld.readBits(2);
      bs_limiter_gains = ld.readBits(2);
      bs_interpol_freq = ld.readBool();
      bs_smoothing_mode = ld.readBool();
    } else { bs_limiter_bands = 2; bs_limiter_gains = 2; bs_interpol_freq = true; bs_smoothing_mode = true; }
  }
  bool differs(const Header* prev) const {
    return prev == nullptr || bs_start_freq != prev->bs_start_freq || bs_stop_freq != prev->bs_stop_freq ||
           bs_freq_scale != prev->bs_freq_scale || bs_alter_scale != prev->bs_alter_scale ||
           bs_xover_band != prev->bs_xover_band || bs_noise_bands != prev->bs_noise_bands;
  }
};

// DCT.java:135-391
struct DCT {
  static void fft_dif(float* Real, float* Imag) {
    const float* wr = JT(SBR_W_ARRAY_REAL);
    const float* wi = JT(SBR_W_ARRAY_IMAG);
    float w_real, w_imag, p1r, p1i, p2r, p2i;
    for (int i = 0; i < 16; i++) {
      p1r = Real[i]; p1i = Imag[i];
      int i2 = i + 16;
      p2r = Real[i2]; p2i = Imag[i2];
      w_real = wr[i]; w_imag = wi[i];
      p1r -= p2r; p1i -= p2i;
      Real[i] += p2r; Imag[i] += p2i;
      Real[i2] = ((p1r * w_real) - (p1i * w_imag));
      Imag[i2] = ((p1r * w_imag) + (p1i * w_real));
    }
    for (int j = 0, w_index = 0; j < 8; j++, w_index += 2) {
      w_real = wr[w_index]; w_imag = wi[w_index];
      for (int half = 0; half < 2; ++half) {
        int i = j + 16 * half;
        p1r = Real[i]; p1i = Imag[i];
        int i2 = i + 8;
        p2r = Real[i2]; p2i = Imag[i2];
        p1r -= p2r; p1i -= p2i;
        Real[i] += p2r; Imag[i] += p2i;
        Real[i2] = ((p1r * w_real) - (p1i * w_imag));
        Imag[i2] = ((p1r * w_imag) + (p1i * w_real));
      }
    }
    for (int i = 0; i < 32; i += 8) {
      int i2 = i + 4;
      p1r = Real[i]; p1i = Imag[i]; p2r = Real[i2]; p2i = Imag[i2];
      Real[i] += p2r; Imag[i] += p2i;
      Real[i2] = p1r - p2r; Imag[i2] = p1i - p2i;
    }
    w_real = wr[4];
    for (int i = 1; i < 32; i += 8) {
      int i2 = i + 4;
      p1r = Real[i]; p1i = Imag[i]; p2r = Real[i2]; p2i = Imag[i2];
      p1r -= p2r; p1i -= p2i;
      Real[i] += p2r; Imag[i] += p2i;
      Real[i2] = (p1r + p1i) * w_real;
      Imag[i2] = (p1i - p1r) * w_real;
    }
    for (int i = 2; i < 32; i += 8) {
      int i2 = i + 4;
      p1r = Real[i]; p1i = Imag[i]; p2r = Real[i2]; p2i = Imag[i2];
      Real[i] += p2r; Imag[i] += p2i;
      Real[i2] = p1i - p2i;
      Imag[i2] = p2r - p1r;
    }
    w_real = wr[12];
    for (int i = 3; i < 32; i += 8) {
      int i2 = i + 4;
      p1r = Real[i]; p1i = Imag[i]; p2r = Real[i2]; p2i = Imag[i2];
      p1r -= p2r; p1i -= p2i;
      Real[i] += p2r; Imag[i] += p2i;
      Real[i2] = (p1r - p1i) * w_real;
      Imag[i2] = (p1r + p1i) * w_real;
    }
    for (int i = 0; i < 32; i += 4) {
      int i2 = i + 2;
      p1r = Real[i]; p1i = Imag[i]; p2r = Real[i2]; p2i = Imag[i2];
      Real[i] += p2r; Imag[i] += p2i;
      Real[i2] = p1r - p2r; Imag[i2] = p1i - p2i;
    }
    for (int i = 1; i < 32; i += 4) {
      int i2 = i + 2;
      p1r = Real[i]; p1i = Imag[i]; p2r = Real[i2]; p2i = Imag[i2];
      Real[i] += p2r; Imag[i] += p2i;
      Real[i2] = p1i - p2i;
      Imag[i2] = p2r - p1r;
    }
    for (int i = 0; i < 32; i += 2) {
      int i2 = i + 1;
      p1r = Real[i]; p1i = Imag[i]; p2r = Real[i2]; p2i = Imag[i2];
      Real[i] += p2r; Imag[i] += p2i;
      Real[i2] = p1r - p2r; Imag[i2] = p1i - p2i;
    }
  }

  static void dct4_kernel(float* in_real, float* in_imag, float* out_real, float* out_imag) {
    const float* tab = JT(SBR_DCT4_64_TAB);
    const int32_t* rev = T::SBR_BIT_REV_TAB;
    for (int i = 0; i < 32; i++) {
      float x_re = in_real[i], x_im = in_imag[i];
      float tmp = (x_re + x_im) * tab[i];
      in_real[i] = (x_im * tab[i + 64]) + tmp;
      in_imag[i] = (x_re * tab[i + 32]) + tmp;
    }
    fft_dif(in_real, in_imag);
    for (int i = 0; i < 16; i++) {
      int i_rev = rev[i];
      float x_re = in_real[i_rev], x_im = in_imag[i_rev];
      float tmp = (x_re + x_im) * tab[i + 3 * 32];
      out_real[i] = (x_im * tab[i + 5 * 32]) + tmp;
      out_imag[i] = (x_re * tab[i + 4 * 32]) + tmp;
    }
    out_imag[16] = (in_imag[1] - in_real[1]) * tab[16 + 3 * 32];
    out_real[16] = (in_real[1] + in_imag[1]) * tab[16 + 3 * 32];
    for (int i = 17; i < 32; i++) {
      int i_rev = rev[i];
      float x_re = in_real[i_rev], x_im = in_imag[i_rev];
      float tmp = (x_re + x_im) * tab[i + 3 * 32];
      out_real[i] = (x_im * tab[i + 5 * 32]) + tmp;
      out_imag[i] = (x_re * tab[i + 4 * 32]) + tmp;
    }
  }
};

// AnalysisFilterbank.java:9-73 (ring buffer from Filterbank.java:330-336)
struct AnalysisFilterbank {
  std::vector<float> v;
  int v_index = 0;
  AnalysisFilterbank() : v(2 * 32 * 20, 0.f) {}

  void sbr_qmf_analysis_32(int numTimeSlotsRate, const float* input, Cpx (*X)[64], int offset, int kx) {
    const float* qmf_c = JT(SBR_QMF_C);
    float u[64], in_real[32], in_imag[32], out_real[32], out_imag[32];
    int in = 0;
    for (int l = 0; l < numTimeSlotsRate; l++) {
      for (int n = 32 - 1; n >= 0; n--) v[v_index + n] = v[v_index + n + 320] = input[in++];
      for (int n = 0; n < 64; n++) {
        u[n] = (v[v_index + n] * qmf_c[2 * n]) + (v[v_index + n + 64] * qmf_c[2 * (n + 64)]) +
               (v[v_index + n + 128] * qmf_c[2 * (n + 128)]) + (v[v_index + n + 192] * qmf_c[2 * (n + 192)]) +
               (v[v_index + n + 256] * qmf_c[2 * (n + 256)]);
      }
      v_index -= 32;
      if (v_index < 0) v_index = (320 - 32);
      in_imag[31] = u[1];
      in_real[0] = u[0];
      for (int n = 1; n < 31; n++) {
        in_imag[31 - n] = u[n + 1];
        in_real[n] = -u[64 - n];
      }
      in_imag[0] = u[32];
      in_real[31] = -u[33];
      DCT::dct4_kernel(in_real, in_imag, out_real, out_imag);
      for (int n = 0; n < 16; n++) {
        if (2 * n + 1 < kx) {
          X[l + offset][2 * n][0] = 2.0f * out_real[n];
          X[l + offset][2 * n][1] = 2.0f * out_imag[n];
          X[l + offset][2 * n + 1][0] = -2.0f * out_imag[31 - n];
          X[l + offset][2 * n + 1][1] = -2.0f * out_real[31 - n];
        } else {
          if (2 * n < kx) {
            X[l + offset][2 * n][0] = 2.0f * out_real[n];
            X[l + offset][2 * n][1] = 2.0f * out_imag[n];
          } else {
            X[l + offset][2 * n][0] = 0;
            X[l + offset][2 * n][1] = 0;
          }
          X[l + offset][2 * n + 1][0] = 0;
          X[l + offset][2 * n + 1][1] = 0;
        }
      }
    }
  }
};

// SynthesisFilterbank64.java:9-79
struct SynthesisFilterbank64 {
  std::vector<float> v;
  int v_index = 0;
  SynthesisFilterbank64() : v(2 * 64 * 20, 0.f) {}

  void synthesis(int numTimeSlotsRate, Cpx (*X)[64], float* output) {
    const float* qmf_c = JT(SBR_QMF_C);
    float in_real1[32], in_imag1[32], out_real1[32], out_imag1[32];
    float in_real2[32], in_imag2[32], out_real2[32], out_imag2[32];
    const float scale = 1.f / 64.f;
    int out = 0;
    for (int l = 0; l < numTimeSlotsRate; l++) {
      Cpx* pX = X[l];
      in_imag1[31] = scale * pX[1][0];
      in_real1[0] = scale * pX[0][0];
      in_imag2[31] = scale * pX[63 - 1][1];
      in_real2[0] = scale * pX[63 - 0][1];
      for (int k = 1; k < 31; k++) {
        in_imag1[31 - k] = scale * pX[2 * k + 1][0];
        in_real1[k] = scale * pX[2 * k][0];
        in_imag2[31 - k] = scale * pX[63 - (2 * k + 1)][1];
        in_real2[k] = scale * pX[63 - (2 * k)][1];
      }
      in_imag1[0] = scale * pX[63][0];
      in_real1[31] = scale * pX[62][0];
      in_imag2[0] = scale * pX[63 - 63][1];
      in_real2[31] = scale * pX[63 - 62][1];
      DCT::dct4_kernel(in_real1, in_imag1, out_real1, out_imag1);
      DCT::dct4_kernel(in_real2, in_imag2, out_real2, out_imag2);
      int p1 = v_index, p3 = p1 + 1280;
      for (int n = 0; n < 32; n++) {
        v[p1 + 2 * n] = v[p3 + 2 * n] = out_real2[n] - out_real1[n];
        v[p1 + 127 - 2 * n] = v[p3 + 127 - 2 * n] = out_real2[n] + out_real1[n];
        v[p1 + 2 * n + 1] = v[p3 + 2 * n + 1] = out_imag2[31 - n] + out_imag1[31 - n];
        v[p1 + 127 - (2 * n + 1)] = v[p3 + 127 - (2 * n + 1)] = out_imag2[31 - n] - out_imag1[31 - n];
      }
      p1 = v_index;
      for (int k = 0; k < 64; k++) {
        output[out++] = (v[p1 + k + 0] * qmf_c[k + 0]) + (v[p1 + k + 192] * qmf_c[k + 64]) +
                        (v[p1 + k + 256] * qmf_c[k + 128]) + (v[p1 + k + (256 + 192)] * qmf_c[k + 192]) +
                        (v[p1 + k + 512] * qmf_c[k + 256]) + (v[p1 + k + (512 + 192)] * qmf_c[k + 320]) +
                        (v[p1 + k + 768] * qmf_c[k + 384]) + (v[p1 + k + (768 + 192)] * qmf_c[k + 448]) +
                        (v[p1 + k + 1024] * qmf_c[k + 512]) + (v[p1 + k + (1024 + 192)] * qmf_c[k + 576]);
      }
      v_index -= 128;
      if (v_index < 0) v_index = (1280 - 128);
    }
  }
};

// SynthesisFilterbank32.java:40-93 (down-sampled SBR: the low 32 QMF bands only, 32 output samples per slot)
struct SynthesisFilterbank32 {
  std::vector<float> v;
  int v_index = 0;
  SynthesisFilterbank32() : v(2 * 32 * 20, 0.f) {}

#define JD_ADD(d, a, b) d = a + b;
#define JD_SUB(d, a, b) d = a - b;
#define JD_MUL(d, c, a) d = (c * a);
  static void DCT4_32(float* x) {  // :95-534, called in place
    JAAD_DCT4_32_TEMPS
    JAAD_DCT4_32_OPS
  }
  static void DST4_32(float* x) {  // :536-940, called in place
    JAAD_DST4_32_TEMPS
    JAAD_DST4_32_OPS
  }
#undef JD_ADD
#undef JD_SUB
#undef JD_MUL

  void synthesis(int numTimeSlotsRate, Cpx (*X)[64], float* output) {
    const float* qmf_c = JT(SBR_QMF_C);
    float tw[64];
    memcpy(tw, JAAD_QMF32_PRE_TWIDDLE_BITS, sizeof tw);
    float x1[32], x2[32];
    const float scale = 1.f / 64.f;
    int out = 0;
    for (int l = 0; l < numTimeSlotsRate; l++) {
      for (int k = 0; k < 32; k++) {
        x1[k] = (X[l][k][0] * tw[2 * k]) - (X[l][k][1] * tw[2 * k + 1]);
        x2[k] = (X[l][k][1] * tw[2 * k]) + (X[l][k][0] * tw[2 * k + 1]);
        x1[k] *= scale;
        x2[k] *= scale;
      }
      DCT4_32(x1);
      DST4_32(x2);
      for (int n = 0; n < 32; n++) {
        v[v_index + n] = v[v_index + 640 + n] = -x1[n] + x2[n];
        v[v_index + 63 - n] = v[v_index + 640 + 63 - n] = x1[n] + x2[n];
      }
      for (int k = 0; k < 32; k++) {
        output[out++] = (v[v_index + k] * qmf_c[2 * k]) + (v[v_index + 96 + k] * qmf_c[64 + 2 * k]) +
                        (v[v_index + 128 + k] * qmf_c[128 + 2 * k]) + (v[v_index + 224 + k] * qmf_c[192 + 2 * k]) +
                        (v[v_index + 256 + k] * qmf_c[256 + 2 * k]) + (v[v_index + 352 + k] * qmf_c[320 + 2 * k]) +
                        (v[v_index + 384 + k] * qmf_c[384 + 2 * k]) + (v[v_index + 480 + k] * qmf_c[448 + 2 * k]) +
                        (v[v_index + 512 + k] * qmf_c[512 + 2 * k]) + (v[v_index + 608 + k] * qmf_c[576 + 2 * k]);
      }
      v_index -= 64;
      if (v_index < 0) v_index = (640 - 64);
    }
  }
};

// SBR.openFilterbank (SBR.java:35-37): the bank an SBR element opens depends on whether the output rate could be doubled
struct SynthesisFilterbank {
  std::unique_ptr<SynthesisFilterbank64> f64;
  std::unique_ptr<SynthesisFilterbank32> f32;
  explicit SynthesisFilterbank(bool downSampled) {
    if (downSampled) f32.reset(new SynthesisFilterbank32());
    else f64.reset(new SynthesisFilterbank64());
  }
  void synthesis(int numTimeSlotsRate, Cpx (*X)[64], float* output) {
    if (f32) f32->synthesis(numTimeSlotsRate, X, output);
    else f64->synthesis(numTimeSlotsRate, X, output);
  }
};

struct SBR;

// Channel.java
struct Channel {
  SBR* sbr;
  bool amp_res = false;
  int abs_bord_lead = 0, abs_bord_trail = 0, n_rel_lead = 0, n_rel_trail = 0;
  int L_E = 0, L_E_prev = 0, L_Q = 0;
  int t_E[MAX_L_E + 1] = {0}, t_Q[3] = {0}, f[MAX_L_E + 1] = {0};
  int f_prev = 0;
  float G_temp_prev[5][64], Q_temp_prev[5][64];
  int GQ_ringbuf_index = 0;
  int E[64][MAX_L_E];
  int E_prev[64];
  float E_orig[64][MAX_L_E], E_curr[64][MAX_L_E];
  int Q[64][2];
  float Q_div[64][2], Q_div2[64][2];
  int Q_prev[64];
  int l_A = 0;
  int bs_invf_mode[MAX_L_E] = {0}, bs_invf_mode_prev[MAX_L_E] = {0};
  float bwArray[64], bwArray_prev[64];
  int bs_add_harmonic[64], bs_add_harmonic_prev[64];
  int index_noise_prev = 0, psi_is_prev = 0, prevEnvIsShort = -1;
  AnalysisFilterbank qmfa;
  std::vector<float> XsbrStore;
  Cpx (*Xsbr)[64];
  int bs_frame_class = FIXFIX;
  int bs_rel_bord[9] = {0}, bs_rel_bord_0[9] = {0}, bs_rel_bord_1[9] = {0};
  int bs_pointer = 0, bs_num_rel_0 = 0, bs_num_rel_1 = 0;
  int bs_df_env[9] = {0}, bs_df_noise[3] = {0};
  bool bs_add_harmonic_flag = false, bs_add_harmonic_flag_prev = false;
  int eTmp[6] = {0};

  explicit Channel(SBR* s) : sbr(s), XsbrStore((size_t)MAX_NTSRHFG * 64 * 2, 0.f) {
    Xsbr = reinterpret_cast<Cpx(*)[64]>(XsbrStore.data());
    memset(G_temp_prev, 0, sizeof G_temp_prev); memset(Q_temp_prev, 0, sizeof Q_temp_prev);
    memset(E, 0, sizeof E); memset(E_prev, 0, sizeof E_prev); memset(E_orig, 0, sizeof E_orig);
    memset(E_curr, 0, sizeof E_curr); memset(Q, 0, sizeof Q); memset(Q_div, 0, sizeof Q_div);
    memset(Q_div2, 0, sizeof Q_div2); memset(Q_prev, 0, sizeof Q_prev); memset(bwArray, 0, sizeof bwArray);
    memset(bwArray_prev, 0, sizeof bwArray_prev); memset(bs_add_harmonic, 0, sizeof bs_add_harmonic);
    memset(bs_add_harmonic_prev, 0, sizeof bs_add_harmonic_prev);
  }

  void sbr_dtdf(BitStream& ld) {  // :85-94
    for (int i = 0; i < L_E; i++) bs_df_env[i] = ld.readBit();
    for (int i = 0; i < L_Q; i++) bs_df_noise[i] = ld.readBit();
  }
  void invf_mode(BitStream& ld);
  void couple(const Channel& oc, int N_Q) {  // :103-122
    bs_frame_class = oc.bs_frame_class;
    L_E = oc.L_E;
    L_Q = oc.L_Q;
    bs_pointer = oc.bs_pointer;
    for (int n = 0; n <= oc.L_E; n++) { t_E[n] = oc.t_E[n]; f[n] = oc.f[n]; }
    for (int n = 0; n <= oc.L_Q; n++) t_Q[n] = oc.t_Q[n];
    for (int n = 0; n < N_Q; n++) bs_invf_mode[n] = oc.bs_invf_mode[n];
  }
  static int decodeHuffman(BitStream& ld, const int16_t* t_huff) {  // :280-289
    int index = 0;
    while (index >= 0) {
      int bit = ld.readBit();
      index = t_huff[index * 2 + bit];
    }
    return index + 64;
  }
  void sbr_envelope(BitStream& ld, bool coupled);
  void extract_envelope_data();
  void sbr_noise(BitStream& ld, bool coupled);
  void extract_noise_floor_data();
  int sbr_grid(BitStream& ld);
  int envelope_time_border_vector();
  void noise_floor_time_border_vector() {  // :529-541
    t_Q[0] = t_E[0];
    if (L_E == 1) { t_Q[1] = t_E[1]; t_Q[2] = 0; }
    else { int index = middleBorder(); t_Q[1] = t_E[index]; t_Q[2] = t_E[L_E]; }
  }
  int middleBorder() const {  // :543-568
    int retval = 0;
    switch (bs_frame_class) {
      case FIXFIX: retval = L_E / 2; break;
      case VARFIX:
        if (bs_pointer == 0) retval = 1;
        else if (bs_pointer == 1) retval = L_E - 1;
        else retval = bs_pointer - 1;
        break;
      case FIXVAR: case VARVAR:
        if (bs_pointer > 1) retval = L_E + 1 - bs_pointer;
        else retval = L_E - 1;
        break;
    }
    return (retval > 0) ? retval : 0;
  }
  void process_channel(float* channel_buf, Cpx (*X)[64], bool reset);
};

// SBR.java (+ FBT.java, HFGeneration.java, HFAdjustment.java, NoiseEnvelope.java as members / friends)
struct SBR : SBRBase {
  DecoderConfig* config;
  int sr_index, sr_freq;  // sample_rate = output frequency (nominal)
  int rate = 2;
  int k0 = 0, kx = 0, M = 0, N_master = 0, N_high = 0, N_low = 0, N_Q = 0;
  int N_L[4] = {0}, n[2] = {0};
  int f_master[64] = {0};
  int f_table_res[2][64];
  int f_table_noise[64] = {0};
  int f_table_lim[4][64];
  int table_map_k_to_g[64] = {0};
  int kx_prev = 0, bsco = 0, bsco_prev = 0, M_prev = 0;
  bool reset = false;
  int frame = 0;
  int noPatches = 0;
  int patchNoSubbands[64] = {0}, patchStartSubband[64] = {0};
  const int numTimeSlotsRate = RATE * NO_TIME_SLOTS, numTimeSlots = NO_TIME_SLOTS;
  int tHFGen = T_HFGEN, tHFAdj = T_HFADJ;
  int bs_sbr_crc_bits = -1;
  std::unique_ptr<Header> hdr, hdr_saved;
  int bs_samplerate_mode = 1;

  explicit SBR(DecoderConfig& c) : config(&c) {  // :101-123
    downSampled = !c.setSBRPresent();
    SampleRate out = c.getOutputFrequency();
    sr_index = out.index;
    sr_freq = SF_FREQ[out.index];
    memset(f_table_res, 0, sizeof f_table_res);
    memset(f_table_lim, 0, sizeof f_table_lim);
  }

  // ---- FBT.java ------------------------------------------------------------------------------------------
  static int qmf_start_channel(int bs_start_freq, int bs_samplerate_mode, int sr_index) {  // :29-41
    int startMin = T::SBR_START_MIN_TABLE[sr_index];
    int offsetIndex = T::SBR_OFFSET_INDEX_TABLE[sr_index];
    if (bs_samplerate_mode != 0) return startMin + T::SBR_OFFSET[offsetIndex * 16 + bs_start_freq];
    return startMin + T::SBR_OFFSET[6 * 16 + bs_start_freq];
  }
  static int qmf_stop_channel(int bs_stop_freq, int sr_index, int k0) {  // :62-76
    if (bs_stop_freq == 15) return std::min(64, k0 * 3);
    if (bs_stop_freq == 14) return std::min(64, k0 * 2);
    int stopMin = T::SBR_STOP_MIN_TABLE[sr_index];
    return std::min(64, stopMin + T::SBR_STOP_OFFSET_TABLE[sr_index * 14 + std::min(bs_stop_freq, 13)]);
  }
  int master_frequency_table_fs0(int k0_, int k2, bool bs_alter_scale) {  // :82-129
    int vDk[64] = {0};
    if (k2 <= k0_) { N_master = 0; return 1; }
    int dk = bs_alter_scale ? 2 : 1;
    int nrBands = bs_alter_scale ? (((k2 - k0_ + 2) >> 2) << 1) : (((k2 - k0_) >> 1) << 1);
    nrBands = std::min(nrBands, 63);
    if (nrBands <= 0) return 1;
    int k2Achieved = k0_ + nrBands * dk;
    int k2Diff = k2 - k2Achieved;
    for (int k = 0; k < nrBands; k++) vDk[k] = dk;
    if (k2Diff != 0) {
      int incr = (k2Diff > 0) ? -1 : 1;
      int k = ((k2Diff > 0) ? (nrBands - 1) : 0);
      while (k2Diff != 0) {
        if (k < 0 || k >= 64) throw AACException(ST_SBR, "master table index out of bounds");
        vDk[k] -= incr;
        k += incr;
        k2Diff += incr;
      }
    }
    f_master[0] = k0_;
    for (int k = 1; k <= nrBands; k++) f_master[k] = (f_master[k - 1] + vDk[k - 1]);
    N_master = std::min(nrBands, 64);
    return 0;
  }
  static int find_bands(int warp, int bands, int a0, int a1) {  // :135-141
    float div = (float)std::log(2.0);
    if (warp != 0) div *= 1.3f;
    return (int)(bands * std::log((double)((float)a1 / (float)a0)) / div + 0.5);
  }
  static float find_initial_power(int bands, int a0, int a1) {  // :143-145
    return (float)std::pow((double)((float)a1 / (float)a0), (double)(1.0f / (float)bands));
  }
  int master_frequency_table(int k0_, int k2, int bs_freq_scale, bool bs_alter_scale) {  // :150-260
    int vDk0[64] = {0}, vDk1[64] = {0}, vk0[64] = {0}, vk1[64] = {0};
    static const int temp1[3] = {6, 5, 4};
    if (k2 <= k0_) { N_master = 0; return 1; }
    int bands = temp1[bs_freq_scale - 1];
    bool twoRegions;
    int k1;
    if ((double)((float)k2 / (float)k0_) > 2.2449) { twoRegions = true; k1 = k0_ << 1; }
    else { twoRegions = false; k1 = k2; }
    int nrBand0 = (2 * find_bands(0, bands, k0_, k1));
    nrBand0 = std::min(nrBand0, 63);
    if (nrBand0 <= 0) return 1;
    float q = find_initial_power(nrBand0, k0_, k1);
    float qk = (float)k0_;
    int A_1 = (int)(qk + 0.5f);
    for (int k = 0; k <= nrBand0; k++) {
      int A_0 = A_1;
      qk *= q;
      A_1 = (int)(qk + 0.5f);
      vDk0[k] = A_1 - A_0;
    }
    std::sort(vDk0, vDk0 + nrBand0);
    vk0[0] = k0_;
    for (int k = 1; k <= nrBand0; k++) {
      vk0[k] = vk0[k - 1] + vDk0[k - 1];
      if (vDk0[k - 1] == 0) return 1;
    }
    if (!twoRegions) {
      for (int k = 0; k <= nrBand0; k++) f_master[k] = vk0[k];
      N_master = std::min(nrBand0, 64);
      return 0;
    }
    int nrBand1 = (2 * find_bands(1, bands, k1, k2));
    nrBand1 = std::min(nrBand1, 63);
    q = find_initial_power(nrBand1, k1, k2);
    qk = (float)k1;
    A_1 = (int)(qk + 0.5f);
    for (int k = 0; k <= nrBand1 - 1; k++) {
      int A_0 = A_1;
      qk *= q;
      A_1 = (int)(qk + 0.5f);
      vDk1[k] = A_1 - A_0;
    }
    if (vDk1[0] < vDk0[nrBand0 - 1]) {
      std::sort(vDk1, vDk1 + nrBand1 + 1);
      int change = vDk0[nrBand0 - 1] - vDk1[0];
      vDk1[0] = vDk0[nrBand0 - 1];
      if (nrBand1 < 1) throw AACException(ST_SBR, "master table index out of bounds");
      vDk1[nrBand1 - 1] = vDk1[nrBand1 - 1] - change;
    }
    std::sort(vDk1, vDk1 + std::max(nrBand1, 0));
    vk1[0] = k1;
    for (int k = 1; k <= nrBand1; k++) {
      vk1[k] = vk1[k - 1] + vDk1[k - 1];
      if (vDk1[k - 1] == 0) return 1;
    }
    N_master = std::min(nrBand0 + nrBand1, 64);
    for (int k = 0; k <= nrBand0; k++) f_master[k] = vk0[k];
    for (int k = nrBand0 + 1; k <= N_master; k++) {
      if (k >= 64) throw AACException(ST_SBR, "master table index out of bounds");
      f_master[k] = vk1[k - nrBand0];
    }
    return 0;
  }
  int derived_frequency_table(int bs_xover_band, int k2) {  // :263-320
    if (N_master <= bs_xover_band) return 1;
    N_high = N_master - bs_xover_band;
    N_low = (N_high >> 1) + (N_high - ((N_high >> 1) << 1));
    n[0] = N_low;
    n[1] = N_high;
    for (int k = 0; k <= N_high; k++) f_table_res[HI_RES][k] = f_master[k + bs_xover_band];
    M = f_table_res[HI_RES][N_high] - f_table_res[HI_RES][0];
    kx = f_table_res[HI_RES][0];
    if (kx > 32) return 1;
    if (kx + M > 64) return 1;
    int minus = ((N_high & 1) != 0) ? 1 : 0;
    for (int i = 0, k = 0; k <= N_low; k++) {
      if (k > 0) i = 2 * k - minus;
      f_table_res[LO_RES][k] = f_table_res[HI_RES][i];
    }
    N_Q = 0;
    if (hdr->bs_noise_bands == 0) N_Q = 1;
    else {
      N_Q = (std::max(1, find_bands(0, hdr->bs_noise_bands, kx, k2)));
      N_Q = std::min(5, N_Q);
    }
    for (int i = 0, k = 0; k <= N_Q; k++) {
      if (k > 0) i += (N_low - i) / (N_Q + 1 - k);
      f_table_noise[k] = f_table_res[LO_RES][i];
    }
    for (int k = 0; k < 64; k++) {
      for (int g = 0; g < N_Q; g++) {
        if ((f_table_noise[g] <= k) && (k < f_table_noise[g + 1])) { table_map_k_to_g[k] = g; break; }
      }
    }
    return 0;
  }
  void limiter_frequency_table() {  // :329-416
    const float* cmp = JT(SBR_LIMITER_BANDS_COMPARE);
    f_table_lim[0][0] = f_table_res[LO_RES][0] - kx;
    f_table_lim[0][1] = f_table_res[LO_RES][N_low] - kx;
    N_L[0] = 1;
    for (int s = 1; s < 4; s++) {
      int limTable[100] = {0};
      int patchBorders[64] = {0};
      patchBorders[0] = kx;
      for (int k = 1; k <= noPatches; k++) patchBorders[k] = patchBorders[k - 1] + patchNoSubbands[k - 1];
      for (int k = 0; k <= N_low; k++) limTable[k] = f_table_res[LO_RES][k];
      for (int k = 1; k < noPatches; k++) limTable[k + N_low] = patchBorders[k];
      std::sort(limTable, limTable + noPatches + N_low);
      int k = 1;
      int nrLim = noPatches + N_low - 1;
      if (nrLim < 0) return;
      while (k <= nrLim) {
        float nOctaves;
        if (limTable[k - 1] != 0) nOctaves = (float)limTable[k] / (float)limTable[k - 1];
        else nOctaves = 0;
        if (nOctaves < cmp[s - 1]) {
          if (limTable[k] != limTable[k - 1]) {
            bool found = false, found2 = false;
            for (int i = 0; i <= noPatches; i++) if (limTable[k] == patchBorders[i]) found = true;
            if (found) {
              found2 = false;
              for (int i = 0; i <= noPatches; i++) if (limTable[k - 1] == patchBorders[i]) found2 = true;
              if (found2) { k++; continue; }
              limTable[k - 1] = f_table_res[LO_RES][N_low];
              std::sort(limTable, limTable + noPatches + N_low);
              nrLim--;
              continue;
            }
          }
          limTable[k] = f_table_res[LO_RES][N_low];
          std::sort(limTable, limTable + nrLim);
          nrLim--;
        } else {
          k++;
        }
      }
      N_L[s] = nrLim;
      for (int l = 0; l <= nrLim; l++) f_table_lim[s][l] = limTable[l] - kx;
    }
  }

  int calc_sbr_tables(Header* h) {  // SBR.java:125-158
    int result = 0;
    k0 = qmf_start_channel(h->bs_start_freq, bs_samplerate_mode, sr_index);
    int k2 = qmf_stop_channel(h->bs_stop_freq, sr_index, k0);
    if (sr_freq >= 48000) { if ((k2 - k0) > 32) result += 1; }
    else if (sr_freq <= 32000) { if ((k2 - k0) > 48) result += 1; }
    else { if ((k2 - k0) > 45) result += 1; }
    if (h->bs_freq_scale == 0) result += master_frequency_table_fs0(k0, k2, h->bs_alter_scale);
    else result += master_frequency_table(k0, k2, h->bs_freq_scale, h->bs_alter_scale);
    result += derived_frequency_table(h->bs_xover_band, k2);
    return (result > 0) ? 1 : 0;
  }

  Header* swapHeaders() {  // :192-204
    std::unique_ptr<Header> h = std::move(hdr_saved);
    hdr_saved = std::move(hdr);
    if (!h) h.reset(new Header());
    hdr = std::move(h);
    return hdr.get();
  }
  bool readHeader(BitStream& ld) {  // :214-223
    bool bs_header_flag = ld.readBool();
    if (bs_header_flag) {
      Header* h = swapHeaders();
      h->decode(ld);
      return h->differs(hdr_saved.get());
    }
    return false;
  }

  void decode(BitStream& ld, bool crc) override {  // :161-185
    if (crc) bs_sbr_crc_bits = ld.readBits(10);
    else bs_sbr_crc_bits = -1;
    reset = readHeader(ld);
    if (reset) {
      int rt = calc_sbr_tables(hdr.get());
      if (rt > 0) calc_sbr_tables(swapHeaders());
    }
    if (hdr) {
      int result = sbr_data(ld);
      valid = (result == 0);
    } else valid = true;
  }

  virtual int sbr_data(BitStream& ld) = 0;
  virtual void sbr_extension(BitStream& ld, int bs_extension_id) { (void)ld; (void)bs_extension_id; }

  void readExtendedData(BitStream& ld0) {  // :229-242
    bool bs_extended_data = ld0.readBool();
    if (bs_extended_data) {
      int cnt = ld0.readBits(4);
      if (cnt == 15) cnt += ld0.readBits(8);
      BitStream ld = ld0.readSubStream(8 * cnt);
      while (ld.getBitsLeft() > 7) {
        int bs_extension_id = ld.readBits(2);
        sbr_extension(ld, bs_extension_id);
      }
    }
  }
  void sinusoidal_coding(BitStream& ld, Channel& ch) {  // :249-254
    for (int i = 0; i < N_high; i++) ch.bs_add_harmonic[i] = ld.readBit();
  }
  void sbr_save_prev_data(Channel& ch) {  // :256-284
    kx_prev = kx;
    M_prev = M;
    bsco_prev = bsco;
    ch.L_E_prev = ch.L_E;
    if (ch.L_E <= 0) throw AACException(ST_SBR, "L_E<0");
    ch.f_prev = ch.f[ch.L_E - 1];
    for (int i = 0; i < MAX_M; i++) {
      ch.E_prev[i] = ch.E[i][ch.L_E - 1];
      ch.Q_prev[i] = ch.Q[i][ch.L_Q - 1];
    }
    for (int i = 0; i < MAX_M; i++) ch.bs_add_harmonic_prev[i] = ch.bs_add_harmonic[i];
    ch.bs_add_harmonic_flag_prev = ch.bs_add_harmonic_flag;
    if (ch.l_A == ch.L_E) ch.prevEnvIsShort = 0;
    else ch.prevEnvIsShort = -1;
  }
  void sbr_save_matrix(Channel& ch) {  // :286-300
    for (int i = 0; i < tHFGen; i++)
      for (int j = 0; j < 64; j++) {
        ch.Xsbr[i][j][0] = ch.Xsbr[i + numTimeSlotsRate][j][0];
        ch.Xsbr[i][j][1] = ch.Xsbr[i + numTimeSlotsRate][j][1];
      }
    for (int i = tHFGen; i < MAX_NTSRHFG; i++)
      for (int j = 0; j < 64; j++) { ch.Xsbr[i][j][0] = 0; ch.Xsbr[i][j][1] = 0; }
  }

  // ---- NoiseEnvelope.java --------------------------------------------------------------------------------
  static float calc_Q_div(const Channel& ch, int m, int l) {  // :207-213
    if (ch.Q[m][l] < 0 || ch.Q[m][l] > 30) return 0;
    return JT(SBR_Q_DIV_TAB)[ch.Q[m][l]];
  }
  static float calc_Q_div2(const Channel& ch, int m, int l) {  // :240-246
    if (ch.Q[m][l] < 0 || ch.Q[m][l] > 30) return 0;
    return JT(SBR_Q_DIV2_TAB)[ch.Q[m][l]];
  }
  void dequantChannel(Channel& ch) {  // :250-281
    const int amp = (ch.amp_res) ? 0 : 1;
    for (int l = 0; l < ch.L_E; l++) {
      for (int k = 0; k < n[ch.f[l]]; k++) {
        int exp = (ch.E[k][l] >> amp);
        if ((exp < 0) || (exp >= 64)) ch.E_orig[k][l] = 0;
        else {
          ch.E_orig[k][l] = JT(SBR_E_DEQ_TAB)[exp];
          if (amp != 0 && (ch.E[k][l] & 1) != 0) ch.E_orig[k][l] = (ch.E_orig[k][l] * 1.414213562f);
        }
      }
    }
    for (int l = 0; l < ch.L_Q; l++)
      for (int k = 0; k < N_Q; k++) {
        ch.Q_div[k][l] = calc_Q_div(ch, k, l);
        ch.Q_div2[k][l] = calc_Q_div2(ch, k, l);
      }
  }

  // ---- HFGeneration.java ---------------------------------------------------------------------------------
  struct acorr_coef { float r01[2], r02[2], r11[2], r12[2], r22[2], det; };
  void auto_correlation(acorr_coef& ac, Cpx (*buffer)[64], int bd, int len) {  // :100-166
    float r01r = 0, r01i = 0, r02r = 0, r02i = 0, r11r = 0;
    float temp1_r, temp1_i, temp2_r, temp2_i, temp3_r, temp3_i, temp4_r, temp4_i, temp5_r, temp5_i;
    const float rel = 1.0f / (1 + 1e-6f);
    const int offset = tHFAdj;
    temp2_r = buffer[offset - 2][bd][0];
    temp2_i = buffer[offset - 2][bd][1];
    temp3_r = buffer[offset - 1][bd][0];
    temp3_i = buffer[offset - 1][bd][1];
    temp4_r = temp2_r; temp4_i = temp2_i; temp5_r = temp3_r; temp5_i = temp3_i;
    for (int j = offset; j < len + offset; j++) {
      temp1_r = temp2_r; temp1_i = temp2_i;
      temp2_r = temp3_r; temp2_i = temp3_i;
      temp3_r = buffer[j][bd][0];
      temp3_i = buffer[j][bd][1];
      r01r += temp3_r * temp2_r + temp3_i * temp2_i;
      r01i += temp3_i * temp2_r - temp3_r * temp2_i;
      r02r += temp3_r * temp1_r + temp3_i * temp1_i;
      r02i += temp3_i * temp1_r - temp3_r * temp1_i;
      r11r += temp2_r * temp2_r + temp2_i * temp2_i;
    }
    ac.r12[0] = r01r - (temp3_r * temp2_r + temp3_i * temp2_i) + (temp5_r * temp4_r + temp5_i * temp4_i);
    ac.r12[1] = r01i - (temp3_i * temp2_r - temp3_r * temp2_i) + (temp5_i * temp4_r - temp5_r * temp4_i);
    ac.r22[0] = r11r - (temp2_r * temp2_r + temp2_i * temp2_i) + (temp4_r * temp4_r + temp4_i * temp4_i);
    ac.r01[0] = r01r; ac.r01[1] = r01i;
    ac.r02[0] = r02r; ac.r02[1] = r02i;
    ac.r11[0] = r11r;
    ac.det = (ac.r11[0] * ac.r22[0]) - (rel * ((ac.r12[0] * ac.r12[0]) + (ac.r12[1] * ac.r12[1])));
  }
  void calc_prediction_coef(Cpx (*Xlow)[64], Cpx* alpha_0, Cpx* alpha_1, int k) {  // :168-204
    float tmp;
    acorr_coef ac;
    auto_correlation(ac, Xlow, k, numTimeSlotsRate + 6);
    if (ac.det == 0) { alpha_1[k][0] = 0; alpha_1[k][1] = 0; }
    else {
      tmp = 1.0f / ac.det;
      alpha_1[k][0] = ((ac.r01[0] * ac.r12[0]) - (ac.r01[1] * ac.r12[1]) - (ac.r02[0] * ac.r11[0])) * tmp;
      alpha_1[k][1] = ((ac.r01[1] * ac.r12[0]) + (ac.r01[0] * ac.r12[1]) - (ac.r02[1] * ac.r11[0])) * tmp;
    }
    if (ac.r11[0] == 0) { alpha_0[k][0] = 0; alpha_0[k][1] = 0; }
    else {
      tmp = 1.0f / ac.r11[0];
      alpha_0[k][0] = -(ac.r01[0] + (alpha_1[k][0] * ac.r12[0]) + (alpha_1[k][1] * ac.r12[1])) * tmp;
      alpha_0[k][1] = -(ac.r01[1] + (alpha_1[k][1] * ac.r12[0]) - (alpha_1[k][0] * ac.r12[1])) * tmp;
    }
    if (((alpha_0[k][0] * alpha_0[k][0]) + (alpha_0[k][1] * alpha_0[k][1]) >= 16.0f) ||
        ((alpha_1[k][0] * alpha_1[k][0]) + (alpha_1[k][1] * alpha_1[k][1]) >= 16.0f)) {
      alpha_0[k][0] = 0; alpha_0[k][1] = 0; alpha_1[k][0] = 0; alpha_1[k][1] = 0;
    }
  }
  static float mapNewBw(int invf_mode, int invf_mode_prev) {  // :207-227
    switch (invf_mode) {
      case 1: return (invf_mode_prev == 0) ? 0.6f : 0.75f;
      case 2: return 0.9f;
      case 3: return 0.98f;
      default: return (invf_mode_prev == 1) ? 0.6f : 0.0f;
    }
  }
  void calc_chirp_factors(Channel& ch) {  // :230-245
    for (int i = 0; i < N_Q; i++) {
      ch.bwArray[i] = mapNewBw(ch.bs_invf_mode[i], ch.bs_invf_mode_prev[i]);
      if (ch.bwArray[i] < ch.bwArray_prev[i]) ch.bwArray[i] = (ch.bwArray[i] * 0.75f) + (ch.bwArray_prev[i] * 0.25f);
      else ch.bwArray[i] = (ch.bwArray[i] * 0.90625f) + (ch.bwArray_prev[i] * 0.09375f);
      if (ch.bwArray[i] < 0.015625f) ch.bwArray[i] = 0.0f;
      if (ch.bwArray[i] >= 0.99609375f) ch.bwArray[i] = 0.99609375f;
      ch.bwArray_prev[i] = ch.bwArray[i];
      ch.bs_invf_mode_prev[i] = ch.bs_invf_mode[i];
    }
  }
  void patch_construction() {  // :247-309
    int msb = k0;
    int usb = kx;
    int goalSb = T::SBR_GOAL_SB_TAB[sr_index];
    noPatches = 0;
    int k = 0;
    if (goalSb < (kx + M)) {
      for (int i = 0; f_master[i] < goalSb; i++) k = i + 1;
    } else k = N_master;
    if (N_master == 0) { noPatches = 0; patchNoSubbands[0] = 0; patchStartSubband[0] = 0; return; }
    int sb;
    int guard = 0;
    do {
      int j = k + 1;
      int odd;
      do {
        j--;
        if (j < 0) throw AACException(ST_SBR, "patch construction ran out of bands");
        sb = f_master[j];
        odd = (sb - 2 + k0) % 2;
      } while (sb > (k0 - 1 + msb - odd));
      patchNoSubbands[noPatches] = std::max(sb - usb, 0);
      patchStartSubband[noPatches] = k0 - odd - patchNoSubbands[noPatches];
      if (patchNoSubbands[noPatches] > 0) { usb = sb; msb = sb; noPatches++; }
      else msb = kx;
      if (f_master[k] - sb < 3) k = N_master;
      if (++guard > 1000 || noPatches >= 63) throw AACException(ST_SBR, "patch construction does not terminate");
    } while (sb != (kx + M));
    if ((patchNoSubbands[noPatches - 1] < 3) && (noPatches > 1)) noPatches--;
    noPatches = std::min(noPatches, 5);
  }
  void hf_generation(Cpx (*Xlow)[64], Cpx (*Xhigh)[64], Channel& ch, bool reset_) {  // :17-98
    Cpx alpha_0[64], alpha_1[64];
    memset(alpha_0, 0, sizeof alpha_0);
    memset(alpha_1, 0, sizeof alpha_1);
    const int offset = tHFAdj;
    const int first = ch.t_E[0];
    const int last = ch.t_E[ch.L_E];
    calc_chirp_factors(ch);
    if (reset_) patch_construction();
    for (int i = 0; i < noPatches; i++) {
      for (int x = 0; x < patchNoSubbands[i]; x++) {
        int k = kx + x;
        for (int q = 0; q < i; q++) k += patchNoSubbands[q];
        int p = patchStartSubband[i] + x;
        if (k < 0 || k >= 64 || p < 0 || p >= 64) throw AACException(ST_SBR, "patch band out of range");
        int g = table_map_k_to_g[k];
        float bw = ch.bwArray[g];
        float bw2 = bw * bw;
        if (bw2 > 0) {
          float temp1_r, temp2_r, temp3_r, temp1_i, temp2_i, temp3_i;
          calc_prediction_coef(Xlow, alpha_0, alpha_1, p);
          float a0_r = (alpha_0[p][0] * bw);
          float a1_r = (alpha_1[p][0] * bw2);
          float a0_i = (alpha_0[p][1] * bw);
          float a1_i = (alpha_1[p][1] * bw2);
          temp2_r = (Xlow[first - 2 + offset][p][0]);
          temp3_r = (Xlow[first - 1 + offset][p][0]);
          temp2_i = (Xlow[first - 2 + offset][p][1]);
          temp3_i = (Xlow[first - 1 + offset][p][1]);
          for (int l = first; l < last; l++) {
            temp1_r = temp2_r; temp2_r = temp3_r; temp3_r = (Xlow[l + offset][p][0]);
            temp1_i = temp2_i; temp2_i = temp3_i; temp3_i = (Xlow[l + offset][p][1]);
            Xhigh[l + offset][k][0] = temp3_r + ((a0_r * temp2_r) - (a0_i * temp2_i) + (a1_r * temp1_r) - (a1_i * temp1_i));
            Xhigh[l + offset][k][1] = temp3_i + ((a0_i * temp2_r) + (a0_r * temp2_i) + (a1_i * temp1_r) + (a1_r * temp1_i));
          }
        } else {
          for (int l = first; l < last; l++) {
            Xhigh[l + offset][k][0] = Xlow[l + offset][p][0];
            Xhigh[l + offset][k][1] = Xlow[l + offset][p][1];
          }
        }
      }
    }
    if (reset) limiter_frequency_table();
  }

  // ---- HFAdjustment.java ---------------------------------------------------------------------------------
  struct Adj {
    float G_lim_boost[MAX_L_E][MAX_M], Q_M_lim_boost[MAX_L_E][MAX_M], S_M_boost[MAX_L_E][MAX_M];
    Adj() { memset(this, 0, sizeof *this); }
  };
  int get_S_mapped(const Channel& ch, int l, int current_band) const {  // :46-76
    if (ch.f[l] == HI_RES) {
      if ((l >= ch.l_A) || (ch.bs_add_harmonic_prev[current_band] != 0 && ch.bs_add_harmonic_flag_prev))
        return ch.bs_add_harmonic[current_band];
    } else {
      int lb = 2 * current_band - ((N_high & 1) != 0 ? 1 : 0);
      int ub = 2 * (current_band + 1) - ((N_high & 1) != 0 ? 1 : 0);
      for (int b = lb; b < ub; b++) {
        if (b < 0 || b >= 64) throw AACException(ST_ARRAY_BOUNDS, "bs_add_harmonic index out of bounds (Java ArrayIndexOutOfBoundsException)");
        if ((l >= ch.l_A) || (ch.bs_add_harmonic_prev[b] != 0 && ch.bs_add_harmonic_flag_prev)) {
          if (ch.bs_add_harmonic[b] == 1) return 1;
        }
      }
    }
    return 0;
  }
  void estimate_current_envelope(Cpx (*Xs)[64], Channel& ch) {  // :78-131
    float nrg, div;
    if (hdr->bs_interpol_freq) {
      for (int l = 0; l < ch.L_E; l++) {
        int l_i = ch.t_E[l], u_i = ch.t_E[l + 1];
        div = (float)(u_i - l_i);
        if (div == 0) div = 1;
        for (int m = 0; m < M; m++) {
          nrg = 0;
          for (int i = l_i + tHFAdj; i < u_i + tHFAdj; i++)
            nrg += (Xs[i][m + kx][0] * Xs[i][m + kx][0]) + (Xs[i][m + kx][1] * Xs[i][m + kx][1]);
          ch.E_curr[m][l] = nrg / div;
        }
      }
    } else {
      for (int l = 0; l < ch.L_E; l++) {
        for (int p = 0; p < n[ch.f[l]]; p++) {
          int k_l = f_table_res[ch.f[l]][p], k_h = f_table_res[ch.f[l]][p + 1];
          for (int k = k_l; k < k_h; k++) {
            nrg = 0;
            int l_i = ch.t_E[l], u_i = ch.t_E[l + 1];
            div = (float)((u_i - l_i) * (k_h - k_l));
            if (div == 0) div = 1;
            for (int i = l_i + tHFAdj; i < u_i + tHFAdj; i++)
              for (int j = k_l; j < k_h; j++) nrg += (Xs[i][j][0] * Xs[i][j][0]) + (Xs[i][j][1] * Xs[i][j][1]);
            ch.E_curr[k - kx][l] = nrg / div;
          }
        }
      }
    }
  }
  void calculate_gain(Adj& adj, Channel& ch) {  // :242-415
    const float EPS = 1e-12f;
    const float* limGain = JT(SBR_LIM_GAIN);
    int current_t_noise_band = 0;
    int S_mapped;
    float Q_M_lim[MAX_M], G_lim[MAX_M], S_M[MAX_M];
    memset(Q_M_lim, 0, sizeof Q_M_lim); memset(G_lim, 0, sizeof G_lim); memset(S_M, 0, sizeof S_M);
    float G_boost;
    for (int l = 0; l < ch.L_E; l++) {
      int current_f_noise_band = 0, current_res_band = 0, current_res_band2 = 0, current_hi_res_band = 0;
      float delta = (l == ch.l_A || l == ch.prevEnvIsShort) ? 0 : 1;
      S_mapped = get_S_mapped(ch, l, current_res_band2);
      if (ch.t_E[l + 1] > ch.t_Q[current_t_noise_band + 1]) current_t_noise_band++;
      for (int k = 0; k < N_L[hdr->bs_limiter_bands]; k++) {
        float G_max, den = 0, acc1 = 0, acc2 = 0;
        int ml1 = f_table_lim[hdr->bs_limiter_bands][k];
        int ml2 = f_table_lim[hdr->bs_limiter_bands][k + 1];
        if (ml1 < 0 || ml2 > MAX_M) throw AACException(ST_SBR, "limiter band out of range");
        for (int m = ml1; m < ml2; m++) {
          if ((m + kx) == f_table_res[ch.f[l]][current_res_band + 1]) current_res_band++;
          acc1 += ch.E_orig[current_res_band][l];
          acc2 += ch.E_curr[m][l];
        }
        G_max = ((EPS + acc1) / (EPS + acc2)) * limGain[hdr->bs_limiter_gains];
        G_max = std::min(G_max, 1e10f);
        for (int m = ml1; m < ml2; m++) {
          float Q_M, G, Q_div, Q_div2;
          int S_index_mapped;
          if ((m + kx) == f_table_noise[current_f_noise_band + 1]) current_f_noise_band++;
          if ((m + kx) == f_table_res[ch.f[l]][current_res_band2 + 1]) {
            current_res_band2++;
            S_mapped = get_S_mapped(ch, l, current_res_band2);
          }
          if ((m + kx) == f_table_res[HI_RES][current_hi_res_band + 1]) current_hi_res_band++;
          S_index_mapped = 0;
          if ((l >= ch.l_A) || (ch.bs_add_harmonic_prev[current_hi_res_band] != 0 && ch.bs_add_harmonic_flag_prev)) {
            if ((m + kx) == (f_table_res[HI_RES][current_hi_res_band + 1] + f_table_res[HI_RES][current_hi_res_band]) >> 1)
              S_index_mapped = ch.bs_add_harmonic[current_hi_res_band];
          }
          Q_div = ch.Q_div[current_f_noise_band][current_t_noise_band];
          Q_div2 = ch.Q_div2[current_f_noise_band][current_t_noise_band];
          Q_M = ch.E_orig[current_res_band2][l] * Q_div2;
          if (S_index_mapped == 0) S_M[m] = 0;
          else {
            S_M[m] = ch.E_orig[current_res_band2][l] * Q_div;
            den += S_M[m];
          }
          G = ch.E_orig[current_res_band2][l] / (1.0f + ch.E_curr[m][l]);
          if ((S_mapped == 0) && (delta == 1)) G *= Q_div;
          else if (S_mapped == 1) G *= Q_div2;
          if (G_max > G) { Q_M_lim[m] = Q_M; G_lim[m] = G; }
          else { Q_M_lim[m] = Q_M * G_max / G; G_lim[m] = G_max; }
          den += ch.E_curr[m][l] * G_lim[m];
          if ((S_index_mapped == 0) && (l != ch.l_A)) den += Q_M_lim[m];
        }
        G_boost = (acc1 + EPS) / (den + EPS);
        G_boost = std::min(G_boost, 2.51188643f);
        for (int m = ml1; m < ml2; m++) {
          adj.G_lim_boost[l][m] = (float)std::sqrt((double)(G_lim[m] * G_boost));
          adj.Q_M_lim_boost[l][m] = (float)std::sqrt((double)(Q_M_lim[m] * G_boost));
          if (S_M[m] != 0) adj.S_M_boost[l][m] = (float)std::sqrt((double)(S_M[m] * G_boost));
          else adj.S_M_boost[l][m] = 0;
        }
      }
    }
  }
  void hf_assembly(Adj& adj, Cpx (*Xs)[64], Channel& ch) {  // :133-240
    static const int phi_re[4] = {1, 0, -1, 0};
    static const int phi_im[4] = {0, 1, 0, -1};
    const float* h_smooth = JT(SBR_H_SMOOTH);
    const float* NOISE = JT(SBR_NOISE_TABLE);
    int fIndexNoise = 0, fIndexSine = 0;
    bool assembly_reset = false;
    float G_filt, Q_filt;
    int h_SL;
    if (reset) { assembly_reset = true; fIndexNoise = 0; }
    else fIndexNoise = ch.index_noise_prev;
    fIndexSine = ch.psi_is_prev;
    for (int l = 0; l < ch.L_E; l++) {
      bool no_noise = (l == ch.l_A || l == ch.prevEnvIsShort);
      h_SL = (hdr->bs_smoothing_mode) ? 0 : 4;
      h_SL = (no_noise ? 0 : h_SL);
      if (assembly_reset) {
        for (int nn = 0; nn < 4; nn++) {
          memcpy(ch.G_temp_prev[nn], adj.G_lim_boost[l], sizeof(float) * M);
          memcpy(ch.Q_temp_prev[nn], adj.Q_M_lim_boost[l], sizeof(float) * M);
        }
        ch.GQ_ringbuf_index = 4;
        assembly_reset = false;
      }
      for (int i = ch.t_E[l]; i < ch.t_E[l + 1]; i++) {
        memcpy(ch.G_temp_prev[ch.GQ_ringbuf_index], adj.G_lim_boost[l], sizeof(float) * M);
        memcpy(ch.Q_temp_prev[ch.GQ_ringbuf_index], adj.Q_M_lim_boost[l], sizeof(float) * M);
        for (int m = 0; m < M; m++) {
          float psi[2];
          G_filt = 0;
          Q_filt = 0;
          if (h_SL != 0) {
            int ri = ch.GQ_ringbuf_index;
            for (int nn = 0; nn <= 4; nn++) {
              float curr_h_smooth = h_smooth[nn];
              ri++;
              if (ri >= 5) ri -= 5;
              G_filt += (ch.G_temp_prev[ri][m] * curr_h_smooth);
              Q_filt += (ch.Q_temp_prev[ri][m] * curr_h_smooth);
            }
          } else {
            G_filt = ch.G_temp_prev[ch.GQ_ringbuf_index][m];
            Q_filt = ch.Q_temp_prev[ch.GQ_ringbuf_index][m];
          }
          Q_filt = (adj.S_M_boost[l][m] != 0 || no_noise) ? 0 : Q_filt;
          fIndexNoise = (fIndexNoise + 1) & 511;
          float* x = Xs[i + tHFAdj][m + kx];
          x[0] = G_filt * x[0] + (Q_filt * NOISE[fIndexNoise * 2 + 0]);
          x[1] = G_filt * x[1] + (Q_filt * NOISE[fIndexNoise * 2 + 1]);
          {
            int rev = (((m + kx) & 1) != 0 ? -1 : 1);
            psi[0] = adj.S_M_boost[l][m] * (float)phi_re[fIndexSine];
            x[0] += psi[0];
            psi[1] = (float)rev * adj.S_M_boost[l][m] * (float)phi_im[fIndexSine];
            x[1] += psi[1];
          }
        }
        fIndexSine = (fIndexSine + 1) & 3;
        ch.GQ_ringbuf_index++;
        if (ch.GQ_ringbuf_index >= 5) ch.GQ_ringbuf_index = 0;
      }
    }
    ch.index_noise_prev = fIndexNoise;
    ch.psi_is_prev = fIndexSine;
  }
  void hf_adjustment(Cpx (*Xs)[64], Channel& ch) {  // :20-44
    Adj adj;
    if (ch.bs_frame_class == FIXFIX) ch.l_A = -1;
    else if (ch.bs_frame_class == VARFIX) {
      if (ch.bs_pointer > 1) ch.l_A = ch.bs_pointer - 1;
      else ch.l_A = -1;
    } else {
      if (ch.bs_pointer == 0) ch.l_A = -1;
      else ch.l_A = ch.L_E + 1 - ch.bs_pointer;
    }
    estimate_current_envelope(Xs, ch);
    calculate_gain(adj, ch);
    hf_assembly(adj, Xs, ch);
  }
};

// ---- Channel methods that need SBR ---------------------------------------------------------------------
inline void Channel::invf_mode(BitStream& ld) {  // :97-101
  for (int i = 0; i < sbr->N_Q; i++) bs_invf_mode[i] = ld.readBits(2);
}

inline void Channel::sbr_envelope(BitStream& ld, bool coupled) {  // :126-190
  int delta = 0;
  const int16_t *t_huff, *f_huff;
  if ((L_E == 1) && (bs_frame_class == FIXFIX)) amp_res = false;
  else amp_res = sbr->hdr->bs_amp_res;
  if (coupled) {
    delta = 1;
    if (amp_res) { t_huff = T::SBR_T_HUFFMAN_ENV_BAL_3_0DB; f_huff = T::SBR_F_HUFFMAN_ENV_BAL_3_0DB; }
    else { t_huff = T::SBR_T_HUFFMAN_ENV_BAL_1_5DB; f_huff = T::SBR_F_HUFFMAN_ENV_BAL_1_5DB; }
  } else {
    delta = 0;
    if (amp_res) { t_huff = T::SBR_T_HUFFMAN_ENV_3_0DB; f_huff = T::SBR_F_HUFFMAN_ENV_3_0DB; }
    else { t_huff = T::SBR_T_HUFFMAN_ENV_1_5DB; f_huff = T::SBR_F_HUFFMAN_ENV_1_5DB; }
  }
  for (int env = 0; env < L_E; env++) {
    if (bs_df_env[env] == 0) {
      if (coupled) {
        if (amp_res) E[0][env] = ld.readBits(5) << delta;
        else E[0][env] = ld.readBits(6) << delta;
      } else {
        if (amp_res) E[0][env] = ld.readBits(6) << delta;
        else E[0][env] = ld.readBits(7) << delta;
      }
      if (getenv("JO_DBG")) fprintf(stderr, "env %d start %d coupled %d amp %d nb %d pos %d\n", env, E[0][env], (int)coupled, (int)amp_res, sbr->n[f[env]], ld.getPosition());
      for (int band = 1; band < sbr->n[f[env]]; band++) { E[band][env] = (decodeHuffman(ld, f_huff) << delta); if (getenv("JO_DBG")) fprintf(stderr, " d%d", E[band][env]); }
    } else {
      for (int band = 0; band < sbr->n[f[env]]; band++) E[band][env] = (decodeHuffman(ld, t_huff) << delta);
    }
  }
  extract_envelope_data();
}

inline void Channel::extract_envelope_data() {  // :192-240
  for (int l = 0; l < L_E; l++) {
    if (bs_df_env[l] == 0) {
      for (int k = 1; k < sbr->n[f[l]]; k++) {
        E[k][l] = E[k - 1][l] + E[k][l];
        if (E[k][l] < 0) E[k][l] = 0;
      }
    } else {
      int g = (l == 0) ? f_prev : f[l - 1];
      if (f[l] == g) {
        for (int k = 0; k < sbr->n[f[l]]; k++) {
          int prev = l == 0 ? E_prev[k] : E[k][l - 1];
          E[k][l] = prev + E[k][l];
        }
      } else if ((g == 1) && (f[l] == 0)) {
        for (int k = 0; k < sbr->n[f[l]]; k++) {
          for (int i = 0; i < sbr->N_high; i++) {
            if (sbr->f_table_res[HI_RES][i] == sbr->f_table_res[LO_RES][k]) {
              int prev = l == 0 ? E_prev[i] : E[i][l - 1];
              E[k][l] = prev + E[k][l];
            }
          }
        }
      } else if ((g == 0) && (f[l] == 1)) {
        for (int k = 0; k < sbr->n[f[l]]; k++) {
          for (int i = 0; i < sbr->N_low; i++) {
            if ((sbr->f_table_res[LO_RES][i] <= sbr->f_table_res[HI_RES][k]) &&
                (sbr->f_table_res[HI_RES][k] < sbr->f_table_res[LO_RES][i + 1])) {
              int prev = l == 0 ? E_prev[i] : E[i][l - 1];
              E[k][l] = prev + E[k][l];
            }
          }
        }
      }
    }
  }
}

inline void Channel::sbr_noise(BitStream& ld, bool coupled) {  // :243-278
  int delta = 0;
  const int16_t *t_huff, *f_huff;
  if (coupled) { delta = 1; t_huff = T::SBR_T_HUFFMAN_NOISE_BAL_3_0DB; f_huff = T::SBR_F_HUFFMAN_ENV_BAL_3_0DB; }
  else { delta = 0; t_huff = T::SBR_T_HUFFMAN_NOISE_3_0DB; f_huff = T::SBR_F_HUFFMAN_ENV_3_0DB; }
  for (int noise = 0; noise < L_Q; noise++) {
    if (bs_df_noise[noise] == 0) {
      Q[0][noise] = ld.readBits(5) << delta;
      for (int band = 1; band < sbr->N_Q; band++) Q[band][noise] = (decodeHuffman(ld, f_huff) << delta);
    } else {
      for (int band = 0; band < sbr->N_Q; band++) Q[band][noise] = (decodeHuffman(ld, t_huff) << delta);
    }
  }
  extract_noise_floor_data();
}

inline void Channel::extract_noise_floor_data() {  // :291-312
  for (int l = 0; l < L_Q; l++) {
    if (bs_df_noise[l] == 0) {
      for (int k = 1; k < sbr->N_Q; k++) Q[k][l] = Q[k][l] + Q[k - 1][l];
    } else {
      if (l == 0) { for (int k = 0; k < sbr->N_Q; k++) Q[k][l] = Q_prev[k] + Q[k][0]; }
      else { for (int k = 0; k < sbr->N_Q; k++) Q[k][l] = Q[k][l - 1] + Q[k][l]; }
    }
  }
}

static inline int sbr_log2(int val) {  // :442-447
  static const int log2tab[10] = {0, 0, 1, 2, 2, 3, 3, 3, 3, 4};
  return (val < 10 && val >= 0) ? log2tab[val] : 0;
}

inline int Channel::sbr_grid(BitStream& ld) {  // :315-437
  int result;
  int saved_L_E = L_E, saved_L_Q = L_Q, saved_frame_class = bs_frame_class;
  bs_frame_class = ld.readBits(2);
  switch (bs_frame_class) {
    case FIXFIX: {
      int i = ld.readBits(2);
      int bs_num_env = std::min(1 << i, 5);
      i = ld.readBit();
      for (int env = 0; env < bs_num_env; env++) f[env] = i;
      L_E = std::min(bs_num_env, 4);
      abs_bord_lead = 0;
      abs_bord_trail = sbr->numTimeSlots;
      n_rel_lead = bs_num_env - 1;
      n_rel_trail = 0;
      break;
    }
    case FIXVAR: {
      int bs_abs_bord = ld.readBits(2) + sbr->numTimeSlots;
      int bs_num_env = ld.readBits(2) + 1;
      for (int rel = 0; rel < bs_num_env - 1; rel++) bs_rel_bord[rel] = 2 * ld.readBits(2) + 2;
      int i = sbr_log2(bs_num_env + 1);
      bs_pointer = ld.readBits(i);
      for (int env = 0; env < bs_num_env; env++) f[bs_num_env - env - 1] = ld.readBit();
      L_E = std::min(bs_num_env, 4);
      abs_bord_lead = 0;
      abs_bord_trail = bs_abs_bord;
      n_rel_lead = 0;
      n_rel_trail = bs_num_env - 1;
      break;
    }
    case VARFIX: {
      int bs_abs_bord = ld.readBits(2);
      int bs_num_env = ld.readBits(2) + 1;
      for (int rel = 0; rel < bs_num_env - 1; rel++) bs_rel_bord[rel] = 2 * ld.readBits(2) + 2;
      int i = sbr_log2(bs_num_env + 1);
      bs_pointer = ld.readBits(i);
      for (int env = 0; env < bs_num_env; env++) f[env] = ld.readBit();
      L_E = std::min(bs_num_env, 4);
      abs_bord_lead = bs_abs_bord;
      abs_bord_trail = sbr->numTimeSlots;
      n_rel_lead = bs_num_env - 1;
      n_rel_trail = 0;
      break;
    }
    case VARVAR: {
      int bs_abs_bord = ld.readBits(2);
      int bs_abs_bord_1 = ld.readBits(2) + sbr->numTimeSlots;
      bs_num_rel_0 = ld.readBits(2);
      bs_num_rel_1 = ld.readBits(2);
      int bs_num_env = std::min(5, bs_num_rel_0 + bs_num_rel_1 + 1);
      for (int rel = 0; rel < bs_num_rel_0; rel++) bs_rel_bord_0[rel] = 2 * ld.readBits(2) + 2;
      for (int rel = 0; rel < bs_num_rel_1; rel++) bs_rel_bord_1[rel] = 2 * ld.readBits(2) + 2;
      int i = sbr_log2(bs_num_rel_0 + bs_num_rel_1 + 2);
      bs_pointer = ld.readBits(i);
      for (int env = 0; env < bs_num_env; env++) f[env] = ld.readBit();
      L_E = std::min(bs_num_env, 5);
      abs_bord_lead = bs_abs_bord;
      abs_bord_trail = bs_abs_bord_1;
      n_rel_lead = bs_num_rel_0;
      n_rel_trail = bs_num_rel_1;
      break;
    }
  }
  if (L_E <= 0) return 1;
  if (L_E > 1) L_Q = 2;
  else L_Q = 1;
  if ((result = envelope_time_border_vector()) > 0) {
    bs_frame_class = saved_frame_class;
    L_E = saved_L_E;
    L_Q = saved_L_Q;
    return result;
  }
  noise_floor_time_border_vector();
  return 0;
}

inline int Channel::envelope_time_border_vector() {  // :452-527
  const int rate = sbr->rate;
  eTmp[0] = rate * abs_bord_lead;
  eTmp[L_E] = rate * abs_bord_trail;
  switch (bs_frame_class) {
    case FIXFIX:
      switch (L_E) {
        case 4: {
          int temp = (sbr->numTimeSlots / 4);
          eTmp[3] = rate * 3 * temp;
          eTmp[2] = rate * 2 * temp;
          eTmp[1] = rate * temp;
          break;
        }
        case 2: eTmp[1] = rate * (sbr->numTimeSlots / 2); break;
        default: break;
      }
      break;
    case FIXVAR:
      if (L_E > 1) {
        int i = L_E;
        int border = abs_bord_trail;
        for (int l = 0; l < (L_E - 1); l++) {
          if (border < bs_rel_bord[l]) return 1;
          border -= bs_rel_bord[l];
          eTmp[--i] = rate * border;
        }
      }
      break;
    case VARFIX:
      if (L_E > 1) {
        int i = 1;
        int border = abs_bord_lead;
        for (int l = 0; l < (L_E - 1); l++) {
          border += bs_rel_bord[l];
          if (rate * border + sbr->tHFAdj > sbr->numTimeSlotsRate + sbr->tHFGen) return 1;
          eTmp[i++] = rate * border;
        }
      }
      break;
    case VARVAR:
      if (bs_num_rel_0 != 0) {
        int i = 1;
        int border = abs_bord_lead;
        for (int l = 0; l < bs_num_rel_0; l++) {
          border += bs_rel_bord_0[l];
          if (rate * border + sbr->tHFAdj > sbr->numTimeSlotsRate + sbr->tHFGen) return 1;
          if (i > 5) throw AACException(ST_SBR, "time border index out of bounds");
          eTmp[i++] = rate * border;
        }
      }
      if (bs_num_rel_1 != 0) {
        int i = L_E;
        int border = abs_bord_trail;
        for (int l = 0; l < bs_num_rel_1; l++) {
          if (border < bs_rel_bord_1[l]) return 1;
          border -= bs_rel_bord_1[l];
          if (i < 1) throw AACException(ST_SBR, "time border index out of bounds");
          eTmp[--i] = rate * border;
        }
      }
      break;
  }
  memcpy(t_E, eTmp, sizeof(int) * 6);
  return 0;
}

inline void Channel::process_channel(float* channel_buf, Cpx (*X)[64], bool reset) {  // :586-647
  sbr->bsco = 0;
  const bool dont_process = !sbr->hdr;
  qmfa.sbr_qmf_analysis_32(sbr->numTimeSlotsRate, channel_buf, Xsbr, sbr->tHFGen, dont_process ? 32 : sbr->kx);
  if (!dont_process) {
    sbr->hf_generation(Xsbr, Xsbr, *this, reset);
    sbr->hf_adjustment(Xsbr, *this);
  }
  if (dont_process) {
    for (int l = 0; l < sbr->numTimeSlotsRate; l++) {
      for (int k = 0; k < 32; k++) {
        X[l][k][0] = Xsbr[l + sbr->tHFAdj][k][0];
        X[l][k][1] = Xsbr[l + sbr->tHFAdj][k][1];
      }
      for (int k = 32; k < 64; k++) { X[l][k][0] = 0; X[l][k][1] = 0; }
    }
  } else {
    for (int l = 0; l < sbr->numTimeSlotsRate; l++) {
      int kx_band, M_band, bsco_band;
      if (l < t_E[0]) { kx_band = sbr->kx_prev; M_band = sbr->M_prev; bsco_band = sbr->bsco_prev; }
      else { kx_band = sbr->kx; M_band = sbr->M; bsco_band = sbr->bsco; }
      for (int k = 0; k < kx_band + bsco_band; k++) {
        X[l][k][0] = Xsbr[l + sbr->tHFAdj][k][0];
        X[l][k][1] = Xsbr[l + sbr->tHFAdj][k][1];
      }
      for (int k = kx_band + bsco_band; k < kx_band + M_band; k++) {
        X[l][k][0] = Xsbr[l + sbr->tHFAdj][k][0];
        X[l][k][1] = Xsbr[l + sbr->tHFAdj][k][1];
      }
      for (int k = std::max(kx_band + bsco_band, kx_band + M_band); k < 64; k++) { X[l][k][0] = 0; X[l][k][1] = 0; }
    }
  }
}

// Hook for the parametric-stereo tool (ps/PSImpl.java); set by jaad_ps.hpp when present.
struct PSBase {
  virtual ~PSBase() {}
  virtual void decode(BitStream& ld) = 0;
  virtual bool isDataAvailable() const = 0;
  virtual void process(Cpx (*X_left)[64], Cpx (*X_right)[64]) = 0;
};
using PSFactory = PSBase* (*)(int numTimeSlotsRate);
inline PSFactory& psFactory() { static PSFactory f = nullptr; return f; }

// SBR1.java
struct SBR1 : SBR {
  Channel ch0;
  SynthesisFilterbank qmfs0;
  std::unique_ptr<SynthesisFilterbank> qmfs1;
  std::unique_ptr<PSBase> ps;
  explicit SBR1(DecoderConfig& c) : SBR(c), ch0(this), qmfs0(downSampled) {}

  int sbr_data(BitStream& ld) override {  // :34-60
    int result;
    if (ld.readBool()) ld.readBits(4);
    if ((result = ch0.sbr_grid(ld)) > 0) return result;
    ch0.sbr_dtdf(ld);
    ch0.invf_mode(ld);
    ch0.sbr_envelope(ld, false);
    ch0.sbr_noise(ld, false);
    dequantChannel(ch0);
    std::fill(ch0.bs_add_harmonic, ch0.bs_add_harmonic + 64, 0);
    ch0.bs_add_harmonic_flag = ld.readBool();
    if (ch0.bs_add_harmonic_flag) sinusoidal_coding(ld, ch0);
    readExtendedData(ld);
    return 0;
  }
  void sbr_extension(BitStream& ld, int bs_extension_id) override {  // :62-73
    if (bs_extension_id == 2 && config->psEnabled) {
      if (!ps) {
        if (!psFactory()) throw AACException(ST_UNSUPPORTED_ELEMENT, "parametric stereo is not built into this oracle");
        ps.reset(psFactory()(numTimeSlotsRate));
        qmfs1.reset(new SynthesisFilterbank(downSampled));
      }
      ps->decode(ld);
    }
  }
  bool isPSUsed() const { return ps && ps->isDataAvailable(); }
  void process(float* left, float* right) override {  // :75-134
    if (isPSUsed()) {
      std::vector<float> xl((size_t)(MAX_NTSR + 6) * 64 * 2, 0.f), xr((size_t)(MAX_NTSR + 6) * 64 * 2, 0.f);
      Cpx (*X_left)[64] = reinterpret_cast<Cpx(*)[64]>(xl.data());
      Cpx (*X_right)[64] = reinterpret_cast<Cpx(*)[64]>(xr.data());
      ch0.process_channel(left, X_left, reset);
      for (int l = numTimeSlotsRate; l < numTimeSlotsRate + 6; l++)
        for (int k = 0; k < 5; k++) {
          X_left[l][k][0] = ch0.Xsbr[tHFAdj + l][k][0];
          X_left[l][k][1] = ch0.Xsbr[tHFAdj + l][k][1];
        }
      ps->process(X_left, X_right);
      qmfs0.synthesis(numTimeSlotsRate, X_left, left);
      qmfs1->synthesis(numTimeSlotsRate, X_right, right);
      if (hdr) sbr_save_prev_data(ch0);
      sbr_save_matrix(ch0);
      frame++;
    } else {
      std::vector<float> xs((size_t)MAX_NTSR * 64 * 2, 0.f);
      Cpx (*X)[64] = reinterpret_cast<Cpx(*)[64]>(xs.data());
      ch0.process_channel(left, X, reset);
      qmfs0.synthesis(numTimeSlotsRate, X, left);
      if (hdr) sbr_save_prev_data(ch0);
      sbr_save_matrix(ch0);
      frame++;
      memcpy(right, left, sizeof(float) * (downSampled ? 1024 : 2048));   // right_chan.length (:80)
    }
  }
};

// SBR2.java
struct SBR2 : SBR {
  Channel ch0, ch1;
  bool bs_coupling = false;
  SynthesisFilterbank qmfs0, qmfs1;
  explicit SBR2(DecoderConfig& c) : SBR(c), ch0(this), ch1(this), qmfs0(downSampled), qmfs1(downSampled) {}

  float calc_Q_div_c(const Channel& ch, int m, int l) const {  // NoiseEnvelope.java:186-205
    if (bs_coupling) {
      const int ch0q = ch0.Q[m][l], ch1q = ch1.Q[m][l];
      if ((ch0q < 0 || ch0q > 30) || (ch1q < 0 || ch1q > 24)) return 0;
      return ((&ch == &ch0) ? JT(SBR_Q_DIV_TAB_LEFT) : JT(SBR_Q_DIV_TAB_RIGHT))[ch0q * 13 + (ch1q >> 1)];
    }
    return calc_Q_div(ch, m, l);
  }
  float calc_Q_div2_c(const Channel& ch, int m, int l) const {  // :215-238
    if (bs_coupling) {
      const int ch0q = ch0.Q[m][l], ch1q = ch1.Q[m][l];
      if ((ch0q < 0 || ch0q > 30) || (ch1q < 0 || ch1q > 24)) return 0;
      return ((&ch == &ch0) ? JT(SBR_Q_DIV2_TAB_LEFT) : JT(SBR_Q_DIV2_TAB_RIGHT))[ch0q * 13 + (ch1q >> 1)];
    }
    return calc_Q_div2(ch, m, l);
  }
  void unmap() {  // NoiseEnvelope.java:299-345
    const int amp0 = (ch0.amp_res) ? 0 : 1;
    const int amp1 = (ch1.amp_res) ? 0 : 1;
    for (int l = 0; l < ch0.L_E; l++) {
      for (int k = 0; k < n[ch0.f[l]]; k++) {
        int ch0E = ch0.E[k][l];
        int exp0 = (ch0E >> amp0) + 1;
        int exp1 = (ch1.E[k][l] >> amp1);
        if ((exp0 < 0) || (exp0 >= 64) || (exp1 < 0) || (exp1 > 24)) {
          ch1.E_orig[k][l] = 0;
          ch0.E_orig[k][l] = 0;
        } else {
          float tmp = JT(SBR_E_DEQ_TAB)[exp0];
          if (amp0 != 0 && (ch0E & 1) != 0) tmp = (float)((double)tmp * 1.414213562);  // `tmp *= 1.414213562` with a double literal
          ch0.E_orig[k][l] = (tmp * JT(SBR_E_PAN_TAB)[exp1]);
          ch1.E_orig[k][l] = (tmp * JT(SBR_E_PAN_TAB)[24 - exp1]);
        }
      }
    }
    for (int l = 0; l < ch0.L_Q; l++)
      for (int k = 0; k < N_Q; k++) {
        ch0.Q_div[k][l] = calc_Q_div_c(ch0, k, l);
        ch1.Q_div[k][l] = calc_Q_div_c(ch1, k, l);
        ch0.Q_div2[k][l] = calc_Q_div2_c(ch0, k, l);
        ch1.Q_div2[k][l] = calc_Q_div2_c(ch1, k, l);
      }
  }

  int sbr_data(BitStream& ld) override {  // :35-135
    int result;
    if (ld.readBool()) { ld.readBits(4); ld.readBits(4); }
    bs_coupling = ld.readBool();
    if (bs_coupling) {
      if ((result = ch0.sbr_grid(ld)) > 0) return result;
      ch0.sbr_dtdf(ld);
      ch1.sbr_dtdf(ld);
      ch0.invf_mode(ld);
      ch1.couple(ch0, N_Q);
      ch0.sbr_envelope(ld, false);
      ch0.sbr_noise(ld, false);
      ch1.sbr_envelope(ld, bs_coupling);
      ch1.sbr_noise(ld, bs_coupling);
      std::fill(ch0.bs_add_harmonic, ch0.bs_add_harmonic + 64, 0);
      std::fill(ch1.bs_add_harmonic, ch1.bs_add_harmonic + 64, 0);
      ch0.bs_add_harmonic_flag = ld.readBool();
      if (ch0.bs_add_harmonic_flag) sinusoidal_coding(ld, ch0);
      ch1.bs_add_harmonic_flag = ld.readBool();
      if (ch1.bs_add_harmonic_flag) sinusoidal_coding(ld, ch1);
    } else {
      int saved_t_E[6] = {0}, saved_t_Q[3] = {0};
      int saved_L_E = ch0.L_E, saved_L_Q = ch0.L_Q, saved_frame_class = ch0.bs_frame_class;
      for (int i = 0; i < saved_L_E; i++) saved_t_E[i] = ch0.t_E[i];
      for (int i = 0; i < saved_L_Q; i++) saved_t_Q[i] = ch0.t_Q[i];
      if ((result = ch0.sbr_grid(ld)) > 0) return result;
      if ((result = ch1.sbr_grid(ld)) > 0) {
        ch0.bs_frame_class = saved_frame_class;
        ch0.L_E = saved_L_E;
        ch0.L_Q = saved_L_Q;
        for (int i = 0; i < 6; i++) ch0.t_E[i] = saved_t_E[i];
        for (int i = 0; i < 3; i++) ch0.t_Q[i] = saved_t_Q[i];
        return result;
      }
      ch0.sbr_dtdf(ld);
      ch1.sbr_dtdf(ld);
      ch0.invf_mode(ld);
      ch1.invf_mode(ld);
      ch0.sbr_envelope(ld, false);
      ch1.sbr_envelope(ld, false);
      ch0.sbr_noise(ld, bs_coupling);
      ch1.sbr_noise(ld, bs_coupling);
      std::fill(ch0.bs_add_harmonic, ch0.bs_add_harmonic + 64, 0);
      std::fill(ch1.bs_add_harmonic, ch1.bs_add_harmonic + 64, 0);
      ch0.bs_add_harmonic_flag = ld.readBool();
      if (ch0.bs_add_harmonic_flag) sinusoidal_coding(ld, ch0);
      ch1.bs_add_harmonic_flag = ld.readBool();
      if (ch1.bs_add_harmonic_flag) sinusoidal_coding(ld, ch1);
    }
    if (!bs_coupling) { dequantChannel(ch0); dequantChannel(ch1); }
    else unmap();
    readExtendedData(ld);
    return 0;
  }

  void process(float* left, float* right) override {  // :137-157
    std::vector<float> xs((size_t)MAX_NTSR * 64 * 2, 0.f);
    Cpx (*X)[64] = reinterpret_cast<Cpx(*)[64]>(xs.data());
    ch0.process_channel(left, X, reset);
    qmfs0.synthesis(numTimeSlotsRate, X, left);
    ch1.process_channel(right, X, false);
    qmfs1.synthesis(numTimeSlotsRate, X, right);
    if (hdr) { sbr_save_prev_data(ch0); sbr_save_prev_data(ch1); }
    sbr_save_matrix(ch0);
    sbr_save_matrix(ch1);
    frame++;
  }
};

inline SBRBase* makeSBR(DecoderConfig& c, bool stereo) {
  if (stereo) return new SBR2(c);
  return new SBR1(c);
}
struct Registrar { Registrar() { sbrFactory() = &makeSBR; } };
static Registrar g_registrar;

}  // namespace sbr
}  // namespace jaad
