// K3 -- SBR payload parse.  One warp owns one SBR element of one stream and walks that stream's frames of the batch in
// order (delta-time coding, header state and the grid-restore rules make the parse sequential per element); lane 0 runs
// the bit-serial syntax on a shared-memory copy of the element's persistent state, all 32 lanes build the per-frame
// record K4 consumes (dequantised envelopes + the band tables in force for that frame).
//
// Reference behaviour followed (paths relative to aac/src/main/java/net/sourceforge/jaad/aac/sbr/):
//   SBR.java:125-300        decode, readHeader/swapHeaders, calc_sbr_tables, extended data, save_prev_data
//   Header.java:24-78       header syntax and the reset test
//   SBR1.java:34-73, SBR2.java:35-135   sbr_data for SCE / CPE (incl. the coupled-branch flag-count quirk)
//   Channel.java:85-583     grid, dtdf, invf, envelope, noise, Huffman tree walk, delta decoding, time borders
//   FBT.java:29-416         start/stop channel, master / derived / limiter band tables
//   HFGeneration.java:247-309  patch construction
//   NoiseEnvelope.java:186-345 dequantisation, coupled un-mapping
//   HFAdjustment.java:24-38 l_A
// A read past the end of the payload is an EOSException in JAAD, which fails the whole frame: every read here is
// checked before anything is stored, so the persistent state is left exactly as the exception would leave it.
#pragma once
#include "jaadb_types.cuh"
#include "sbr_types.cuh"

namespace jaadb {

#define JAADB_ST_SBR 14

enum { SBR_FIXFIX = 0, SBR_FIXVAR = 1, SBR_VARFIX = 2, SBR_VARVAR = 3 };
enum { SBR_LO_RES = 0, SBR_HI_RES = 1 };

// The Huffman tables (20 KB of 8-bit front tables, the trees behind them) are the only global data the parse touches again
// and again; with 200 KB of the SM's L1 carved out as shared memory for the element states they compete for what is left
// with data that is read once (payload words, frame descriptors, the state copy-in).  So: keep the former, do not
// allocate the latter.
// JAADB_BOUNDS_ASSERT: debug builds only (tools/build_variants.sh): device-side asserts on the index ranges of the
// shared-memory windows below; the product build compiles them out.
#ifdef JAADB_BOUNDS_ASSERT
#include <cassert>
#define JAADB_ASSERT(x) assert(x)
#else
#define JAADB_ASSERT(x) do {} while (0)
#endif
#ifndef K3_L1_HINTS
#define K3_L1_HINTS 1
#endif
__device__ __forceinline__ uint32_t k3_ld_keep(const uint32_t* p) {
#if K3_L1_HINTS
  uint32_t v;
  asm("ld.global.nc.L1::evict_last.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ int k3_ld_keep(const int16_t* p) {
#if K3_L1_HINTS
  int v;
  asm("ld.global.nc.L1::evict_last.s16 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ uint32_t k3_ld_stream(const uint32_t* p) {
#if K3_L1_HINTS
  uint32_t v;
  asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
#else
  return __ldg(p);
#endif
}
__device__ __forceinline__ uint32_t k3_ld_stream_rw(const uint32_t* p) {   // data this kernel also writes: no .nc
#if K3_L1_HINTS
  uint32_t v;
  asm volatile("ld.global.L1::no_allocate.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
#else
  return *p;
#endif
}

// Bit reader over one SBR payload.  The warp stages the payload's words in shared memory, already in big-endian bit
// order (k3_sbr_parse_kernel), so a read is two shared loads and a funnel shift: no 64-bit address arithmetic, no byte
// swap and no trip to L1 per field.  `pos` / `end` count bits from bit 31 of words[0].
struct SbrBits {
  const uint32_t* words;
  uint32_t pos, end;
  __device__ __forceinline__ uint32_t word(uint32_t i) const { return words[i]; }
  __device__ __forceinline__ uint32_t left() const { return end > pos ? end - pos : 0u; }
  __device__ __forceinline__ uint32_t peek32() const {
    const uint32_t wi = pos >> 5;
    JAADB_ASSERT(wi + 1 < 72u);   // kK3StageWords
    return __funnelshift_l(words[wi + 1], words[wi], pos & 31u);
  }
  // n <= 25
  __device__ __forceinline__ bool get(int n, int& v) {
    if (n == 0) { v = 0; return true; }
    if (pos + (uint32_t)n > end) return false;
    v = (int)(peek32() >> (32 - n));
    pos += n;
    return true;
  }
};

// every syntax function returns 0, or a JAADB_ST_* code; EOS aborts immediately
#define SBR_RD(var, n) do { if (!ld.get((n), (var))) return JAADB_ST_EOS; } while (0)
#define SBR_TRY(expr) do { int _st = (expr); if (_st) return _st; } while (0)

struct SbrCtx {
  SbrElemDev* S;
  const SbrTablesDev* T;
  const int32_t* start_min, *offset_index, *offset, *stop_min, *stop_offset, *goal_sb;
  const float* limiter_cmp;
  int sr_index;
};

__device__ inline void sbr_sort(int* a, int lo, int hi) {  // Arrays.sort(a, lo, hi)
  for (int i = lo + 1; i < hi; ++i) {
    int v = a[i], j = i - 1;
    while (j >= lo && a[j] > v) { a[j + 1] = a[j]; --j; }
    a[j + 1] = v;
  }
}

// ---- Header.java -------------------------------------------------------------------------------------------
__device__ inline int sbr_header_decode(SbrBits& ld, SbrHeaderDev& h) {
  int v, e1, e2;
  SBR_RD(v, 1); h.amp_res = (uint8_t)v;
  SBR_RD(v, 4); h.start_freq = (uint8_t)v;
  SBR_RD(v, 4); h.stop_freq = (uint8_t)v;
  SBR_RD(v, 3); h.xover_band = (uint8_t)v;
  SBR_RD(v, 2);
  SBR_RD(e1, 1);
  SBR_RD(e2, 1);
  if (e1) {
    SBR_RD(v, 2); h.freq_scale = (uint8_t)v;
    SBR_RD(v, 1); h.alter_scale = (uint8_t)v;
    SBR_RD(v, 2); h.noise_bands = (uint8_t)v;
  } else { h.freq_scale = 2; h.alter_scale = 1; h.noise_bands = 2; }
  if (e2) {
    SBR_RD(v, 2); h.limiter_bands = (uint8_t)v;
    SBR_RD(v, 2); h.limiter_gains = (uint8_t)v;
    SBR_RD(v, 1); h.interpol_freq = (uint8_t)v;
    SBR_RD(v, 1); h.smoothing_mode = (uint8_t)v;
  } else { h.limiter_bands = 2; h.limiter_gains = 2; h.interpol_freq = 1; h.smoothing_mode = 1; }
  return 0;
}

__device__ inline bool sbr_header_differs(const SbrHeaderDev& a, const SbrHeaderDev& prev) {
  return !prev.present || a.start_freq != prev.start_freq || a.stop_freq != prev.stop_freq || a.freq_scale != prev.freq_scale ||
         a.alter_scale != prev.alter_scale || a.xover_band != prev.xover_band || a.noise_bands != prev.noise_bands;
}

// a freshly constructed Header (Header.java:14-22)
__device__ inline SbrHeaderDev sbr_header_new() {
  SbrHeaderDev h;
  h.present = 1; h.amp_res = 1; h.start_freq = 5; h.stop_freq = 0; h.xover_band = 0; h.freq_scale = 2; h.alter_scale = 1;
  h.noise_bands = 2; h.limiter_bands = 2; h.limiter_gains = 2; h.interpol_freq = 0; h.smoothing_mode = 0;
  return h;
}

// SBR.swapHeaders (:192-204)
__device__ inline void sbr_swap_headers(SbrElemDev& S) {
  SbrHeaderDev h = S.hdr_saved;
  S.hdr_saved = S.hdr;
  if (!h.present) h = sbr_header_new();
  S.hdr = h;
}

// ---- FBT.java ----------------------------------------------------------------------------------------------
__device__ inline int sbr_find_bands(const SbrCtx& C, int warp, int bands, int a0, int a1) {
  if (a0 < 1 || a0 > 64 || a1 < 0 || a1 > 64 || bands < 0 || bands > 6) return 0;
  return C.T->find_bands[((warp * 7 + bands) * 65 + a0) * 65 + a1];
}
__device__ inline float sbr_initial_power(const SbrCtx& C, int bands, int a0, int a1) {
  return C.T->init_power[(bands * 65 + a0) * 65 + a1];
}

__device__ inline int sbr_master_table_fs0(const SbrCtx& C, int k0, int k2, bool alter) {  // :82-129
  SbrElemDev& S = *C.S;
  int vDk[64];
  if (k2 <= k0) { S.N_master = 0; return 1; }
  const int dk = alter ? 2 : 1;
  int nrBands = alter ? (((k2 - k0 + 2) >> 2) << 1) : (((k2 - k0) >> 1) << 1);
  nrBands = min(nrBands, 63);
  if (nrBands <= 0) return 1;
  int k2Diff = k2 - (k0 + nrBands * dk);
  for (int k = 0; k < nrBands; k++) vDk[k] = dk;
  if (k2Diff != 0) {
    const int incr = (k2Diff > 0) ? -1 : 1;
    int k = (k2Diff > 0) ? (nrBands - 1) : 0;
    while (k2Diff != 0) {
      if (k < 0 || k >= 64) return -JAADB_ST_SBR;
      vDk[k] -= incr;
      k += incr;
      k2Diff += incr;
    }
  }
  int fm = k0;
  S.f_master[0] = (uint8_t)fm;
  for (int k = 1; k <= nrBands; k++) { fm += vDk[k - 1]; S.f_master[k] = (uint8_t)fm; }
  S.N_master = (uint8_t)min(nrBands, 64);
  return 0;
}

__device__ inline int sbr_master_table(const SbrCtx& C, int k0, int k2, int freq_scale, bool alter) {  // :150-260
  (void)alter;
  SbrElemDev& S = *C.S;
  int vDk0[64], vDk1[64], vk0[64], vk1[64];
  for (int i = 0; i < 64; ++i) { vDk0[i] = 0; vDk1[i] = 0; vk0[i] = 0; vk1[i] = 0; }
  if (k2 <= k0) { S.N_master = 0; return 1; }
  const int bands = freq_scale == 1 ? 6 : (freq_scale == 2 ? 5 : 4);
  bool twoRegions;
  int k1;
  if ((double)((float)k2 / (float)k0) > 2.2449) { twoRegions = true; k1 = k0 << 1; }
  else { twoRegions = false; k1 = k2; }
  int nrBand0 = 2 * sbr_find_bands(C, 0, bands, k0, k1);
  nrBand0 = min(nrBand0, 63);
  if (nrBand0 <= 0) return 1;
  float q = sbr_initial_power(C, nrBand0, k0, k1);
  float qk = (float)k0;
  int A_1 = (int)(qk + 0.5f);
  for (int k = 0; k <= nrBand0; k++) {
    const int A_0 = A_1;
    qk *= q;
    A_1 = (int)(qk + 0.5f);
    vDk0[k] = A_1 - A_0;
  }
  sbr_sort(vDk0, 0, nrBand0);
  vk0[0] = k0;
  for (int k = 1; k <= nrBand0; k++) {
    vk0[k] = vk0[k - 1] + vDk0[k - 1];
    if (vDk0[k - 1] == 0) return 1;
  }
  if (!twoRegions) {
    for (int k = 0; k <= nrBand0; k++) S.f_master[k] = (uint8_t)vk0[k];
    S.N_master = (uint8_t)min(nrBand0, 64);
    return 0;
  }
  int nrBand1 = 2 * sbr_find_bands(C, 1, bands, k1, k2);
  nrBand1 = min(nrBand1, 63);
  if (nrBand1 < 1) return -JAADB_ST_SBR;   // Math.pow(x, 1/0) / negative sort ranges: outside the decodable subset
  q = sbr_initial_power(C, nrBand1, k1, k2);
  qk = (float)k1;
  A_1 = (int)(qk + 0.5f);
  for (int k = 0; k <= nrBand1 - 1; k++) {
    const int A_0 = A_1;
    qk *= q;
    A_1 = (int)(qk + 0.5f);
    vDk1[k] = A_1 - A_0;
  }
  if (vDk1[0] < vDk0[nrBand0 - 1]) {
    sbr_sort(vDk1, 0, nrBand1 + 1);
    const int change = vDk0[nrBand0 - 1] - vDk1[0];
    vDk1[0] = vDk0[nrBand0 - 1];
    vDk1[nrBand1 - 1] = vDk1[nrBand1 - 1] - change;
  }
  sbr_sort(vDk1, 0, nrBand1);
  vk1[0] = k1;
  for (int k = 1; k <= nrBand1; k++) {
    vk1[k] = vk1[k - 1] + vDk1[k - 1];
    if (vDk1[k - 1] == 0) return 1;
  }
  const int nm = min(nrBand0 + nrBand1, 64);
  S.N_master = (uint8_t)nm;
  for (int k = 0; k <= nrBand0; k++) S.f_master[k] = (uint8_t)vk0[k];
  for (int k = nrBand0 + 1; k <= nm; k++) {
    if (k >= 64) return -JAADB_ST_SBR;
    S.f_master[k] = (uint8_t)vk1[k - nrBand0];
  }
  return 0;
}

__device__ inline int sbr_derived_table(const SbrCtx& C, int xover, int k2) {  // :263-320
  SbrElemDev& S = *C.S;
  if (S.N_master <= xover) return 1;
  const int N_high = S.N_master - xover;
  const int N_low = (N_high >> 1) + (N_high - ((N_high >> 1) << 1));
  S.N_high = (uint8_t)N_high;
  S.N_low = (uint8_t)N_low;
  S.n[0] = (uint8_t)N_low;
  S.n[1] = (uint8_t)N_high;
  for (int k = 0; k <= N_high; k++) S.f_table_res[SBR_HI_RES][k] = S.f_master[k + xover];
  const int M = S.f_table_res[SBR_HI_RES][N_high] - S.f_table_res[SBR_HI_RES][0];
  const int kx = S.f_table_res[SBR_HI_RES][0];
  S.M = (uint8_t)M;
  S.kx = (uint8_t)kx;
  if (kx > 32) return 1;
  if (kx + M > 64) return 1;
  const int minus = (N_high & 1) ? 1 : 0;
  for (int i = 0, k = 0; k <= N_low; k++) {
    if (k > 0) i = 2 * k - minus;
    S.f_table_res[SBR_LO_RES][k] = S.f_table_res[SBR_HI_RES][i];
  }
  int N_Q;
  if (S.hdr.noise_bands == 0) N_Q = 1;
  else {
    N_Q = max(1, sbr_find_bands(C, 0, S.hdr.noise_bands, kx, k2));
    N_Q = min(5, N_Q);
  }
  S.N_Q = (uint8_t)N_Q;
  for (int i = 0, k = 0; k <= N_Q; k++) {
    if (k > 0) i += (N_low - i) / (N_Q + 1 - k);
    S.f_table_noise[k] = S.f_table_res[SBR_LO_RES][i];
  }
  for (int k = 0; k < 64; k++)
    for (int g = 0; g < N_Q; g++)
      if ((S.f_table_noise[g] <= k) && (k < S.f_table_noise[g + 1])) { S.table_map_k_to_g[k] = (uint8_t)g; break; }
  return 0;
}

// SBR.calc_sbr_tables (:125-158).  Negative return: an internal index ran out of bounds (a Java exception).
__device__ inline int sbr_calc_tables(const SbrCtx& C) {
  SbrElemDev& S = *C.S;
  const SbrHeaderDev& h = S.hdr;
  int result = 0;
  const int si = C.sr_index;
  const int k0 = C.start_min[si] + C.offset[C.offset_index[si] * 16 + h.start_freq];   // bs_samplerate_mode = 1
  int k2;
  if (h.stop_freq == 15) k2 = min(64, k0 * 3);
  else if (h.stop_freq == 14) k2 = min(64, k0 * 2);
  else k2 = min(64, C.stop_min[si] + C.stop_offset[si * 14 + min((int)h.stop_freq, 13)]);
  S.k0 = (uint8_t)k0;
  const int freq = si == 0 ? 96000 : si == 1 ? 88200 : si == 2 ? 64000 : si == 3 ? 48000 : si == 4 ? 44100 : si == 5 ? 32000 :
                   si == 6 ? 24000 : si == 7 ? 22050 : si == 8 ? 16000 : si == 9 ? 12000 : si == 10 ? 11025 : 8000;
  if (freq >= 48000) { if ((k2 - k0) > 32) result += 1; }
  else if (freq <= 32000) { if ((k2 - k0) > 48) result += 1; }
  else { if ((k2 - k0) > 45) result += 1; }
  int r;
  if (h.freq_scale == 0) r = sbr_master_table_fs0(C, k0, k2, h.alter_scale != 0);
  else r = sbr_master_table(C, k0, k2, h.freq_scale, h.alter_scale != 0);
  if (r < 0) return r;
  result += r;
  r = sbr_derived_table(C, h.xover_band, k2);
  if (r < 0) return r;
  result += r;
  return result > 0 ? 1 : 0;
}

// HFGeneration.patch_construction (:247-309)
__device__ inline int sbr_patch_construction(const SbrCtx& C) {
  SbrElemDev& S = *C.S;
  int msb = S.k0, usb = S.kx;
  const int goalSb = C.goal_sb[C.sr_index];
  int noPatches = 0;
  int k = 0;
  if (goalSb < (S.kx + S.M)) { for (int i = 0; i < 64 && S.f_master[i] < goalSb; i++) k = i + 1; }
  else k = S.N_master;
  if (S.N_master == 0) { S.noPatches = 0; S.patchNoSubbands[0] = 0; S.patchStartSubband[0] = 0; return 0; }
  int sb, guard = 0;
  do {
    int j = k + 1, odd;
    do {
      j--;
      if (j < 0 || j >= 64) return JAADB_ST_SBR;
      sb = S.f_master[j];
      odd = (sb - 2 + S.k0) % 2;
    } while (sb > (S.k0 - 1 + msb - odd));
    const int nsub = max(sb - usb, 0);
    S.patchNoSubbands[noPatches] = (uint8_t)nsub;
    S.patchStartSubband[noPatches] = (int8_t)(S.k0 - odd - nsub);
    if (nsub > 0) { usb = sb; msb = sb; noPatches++; }
    else msb = S.kx;
    if (k >= 0 && k < 64 && S.f_master[k] - sb < 3) k = S.N_master;
    if (++guard > 1000 || noPatches >= 63) return JAADB_ST_SBR;
  } while (sb != (S.kx + S.M));
  if ((S.patchNoSubbands[noPatches - 1] < 3) && (noPatches > 1)) noPatches--;
  S.noPatches = (uint8_t)min(noPatches, 5);
  return 0;
}

// FBT.limiter_frequency_table (:329-416)
__device__ inline void sbr_limiter_table(const SbrCtx& C) {
  SbrElemDev& S = *C.S;
  const int N_low = S.N_low, kx = S.kx, noPatches = S.noPatches;
  S.f_table_lim[0][0] = (int8_t)(S.f_table_res[SBR_LO_RES][0] - kx);
  S.f_table_lim[0][1] = (int8_t)(S.f_table_res[SBR_LO_RES][N_low] - kx);
  S.N_L[0] = 1;
  for (int s = 1; s < 4; s++) {
    int limTable[100];
    int patchBorders[64];
    for (int i = 0; i < 100; ++i) limTable[i] = 0;
    for (int i = 0; i < 64; ++i) patchBorders[i] = 0;
    patchBorders[0] = kx;
    for (int k = 1; k <= noPatches; k++) patchBorders[k] = patchBorders[k - 1] + S.patchNoSubbands[k - 1];
    for (int k = 0; k <= N_low; k++) limTable[k] = S.f_table_res[SBR_LO_RES][k];
    for (int k = 1; k < noPatches; k++) limTable[k + N_low] = patchBorders[k];
    sbr_sort(limTable, 0, noPatches + N_low);
    int k = 1;
    int nrLim = noPatches + N_low - 1;
    if (nrLim < 0) return;
    while (k <= nrLim) {
      float nOctaves;
      if (limTable[k - 1] != 0) nOctaves = (float)limTable[k] / (float)limTable[k - 1];
      else nOctaves = 0;
      if (nOctaves < C.limiter_cmp[s - 1]) {
        if (limTable[k] != limTable[k - 1]) {
          bool found = false, found2 = false;
          for (int i = 0; i <= noPatches; i++) if (limTable[k] == patchBorders[i]) found = true;
          if (found) {
            for (int i = 0; i <= noPatches; i++) if (limTable[k - 1] == patchBorders[i]) found2 = true;
            if (found2) { k++; continue; }
            limTable[k - 1] = S.f_table_res[SBR_LO_RES][N_low];
            sbr_sort(limTable, 0, noPatches + N_low);
            nrLim--;
            continue;
          }
        }
        limTable[k] = S.f_table_res[SBR_LO_RES][N_low];
        sbr_sort(limTable, 0, nrLim);
        nrLim--;
      } else {
        k++;
      }
    }
    S.N_L[s] = (uint8_t)nrLim;
    for (int l = 0; l <= nrLim; l++) S.f_table_lim[s][l] = (int8_t)(limTable[l] - kx);
  }
}

// ---- Channel.java ------------------------------------------------------------------------------------------
__device__ inline int sbr_log2(int val) {
  return (val < 10 && val >= 0) ? ((0x4333322100ull >> (4 * val)) & 15) : 0;   // {0,0,1,2,2,3,3,3,3,4}
}

__device__ inline int sbr_middle_border(const SbrChanParse& c) {  // :543-568
  int retval = 0;
  switch (c.frame_class) {
    case SBR_FIXFIX: retval = c.L_E / 2; break;
    case SBR_VARFIX:
      if (c.bs_pointer == 0) retval = 1;
      else if (c.bs_pointer == 1) retval = c.L_E - 1;
      else retval = c.bs_pointer - 1;
      break;
    default:
      if (c.bs_pointer > 1) retval = c.L_E + 1 - c.bs_pointer;
      else retval = c.L_E - 1;
      break;
  }
  return retval > 0 ? retval : 0;
}

__device__ inline int sbr_time_border_vector(SbrChanParse& c) {  // :452-527; 1 = invalid grid
  const int rate = 2, numTimeSlots = 16, numTimeSlotsRate = 32, tHFAdj = kSbrHfAdj, tHFGen = kSbrHfGen;
  int eTmp[6];
  for (int i = 0; i < 6; ++i) eTmp[i] = c.t_E[i];   // `eTmp` is a Channel field in the reference: it keeps old entries
  // (every entry that is read below is written first, so the carried values never reach t_E beyond index L_E; they are
  // copied along with System.arraycopy(eTmp, 0, t_E, 0, 6) and kept here for the same reason)
  eTmp[0] = rate * c.abs_bord_lead;
  eTmp[c.L_E] = rate * c.abs_bord_trail;
  switch (c.frame_class) {
    case SBR_FIXFIX:
      if (c.L_E == 4) { const int temp = numTimeSlots / 4; eTmp[3] = rate * 3 * temp; eTmp[2] = rate * 2 * temp; eTmp[1] = rate * temp; }
      else if (c.L_E == 2) eTmp[1] = rate * (numTimeSlots / 2);
      break;
    case SBR_FIXVAR:
      if (c.L_E > 1) {
        int i = c.L_E, border = c.abs_bord_trail;
        for (int l = 0; l < (c.L_E - 1); l++) {
          if (border < c.bs_rel_bord[l]) return 1;
          border -= c.bs_rel_bord[l];
          eTmp[--i] = rate * border;
        }
      }
      break;
    case SBR_VARFIX:
      if (c.L_E > 1) {
        int i = 1, border = c.abs_bord_lead;
        for (int l = 0; l < (c.L_E - 1); l++) {
          border += c.bs_rel_bord[l];
          if (rate * border + tHFAdj > numTimeSlotsRate + tHFGen) return 1;
          eTmp[i++] = rate * border;
        }
      }
      break;
    default:
      if (c.bs_num_rel_0 != 0) {
        int i = 1, border = c.abs_bord_lead;
        for (int l = 0; l < c.bs_num_rel_0; l++) {
          border += c.bs_rel_bord_0[l];
          if (rate * border + tHFAdj > numTimeSlotsRate + tHFGen) return 1;
          if (i > 5) return -JAADB_ST_SBR;
          eTmp[i++] = rate * border;
        }
      }
      if (c.bs_num_rel_1 != 0) {
        int i = c.L_E, border = c.abs_bord_trail;
        for (int l = 0; l < c.bs_num_rel_1; l++) {
          if (border < c.bs_rel_bord_1[l]) return 1;
          border -= c.bs_rel_bord_1[l];
          if (i < 1) return -JAADB_ST_SBR;
          eTmp[--i] = rate * border;
        }
      }
      break;
  }
  for (int i = 0; i < 6; ++i) c.t_E[i] = (uint8_t)eTmp[i];
  return 0;
}

// Channel.sbr_grid (:315-437).  Returns 0, 1 (invalid grid: sbr_data gives up, valid = false) or a negative status.
__device__ inline int sbr_grid(SbrBits& ld, SbrChanParse& c, int& eos) {
  const int numTimeSlots = 16;
  eos = 0;
#define GRID_RD(var, n) do { if (!ld.get((n), (var))) { eos = 1; return -JAADB_ST_EOS; } } while (0)
  const int saved_L_E = c.L_E, saved_L_Q = c.L_Q, saved_class = c.frame_class;
  int v;
  GRID_RD(v, 2);
  c.frame_class = (uint8_t)v;
  switch (c.frame_class) {
    case SBR_FIXFIX: {
      int i;
      GRID_RD(i, 2);
      const int bs_num_env = min(1 << i, 5);
      GRID_RD(i, 1);
      for (int env = 0; env < bs_num_env; env++) c.f[env] = (uint8_t)i;
      c.L_E = (uint8_t)min(bs_num_env, 4);
      c.abs_bord_lead = 0;
      c.abs_bord_trail = numTimeSlots;
      break;
    }
    case SBR_FIXVAR: {
      int bs_abs_bord, bs_num_env;
      GRID_RD(bs_abs_bord, 2);
      bs_abs_bord += numTimeSlots;
      GRID_RD(bs_num_env, 2);
      bs_num_env += 1;
      for (int rel = 0; rel < bs_num_env - 1; rel++) { GRID_RD(v, 2); c.bs_rel_bord[rel] = (uint8_t)(2 * v + 2); }
      GRID_RD(v, sbr_log2(bs_num_env + 1));
      c.bs_pointer = (uint8_t)v;
      for (int env = 0; env < bs_num_env; env++) { GRID_RD(v, 1); c.f[bs_num_env - env - 1] = (uint8_t)v; }
      c.L_E = (uint8_t)min(bs_num_env, 4);
      c.abs_bord_lead = 0;
      c.abs_bord_trail = (uint8_t)bs_abs_bord;
      break;
    }
    case SBR_VARFIX: {
      int bs_abs_bord, bs_num_env;
      GRID_RD(bs_abs_bord, 2);
      GRID_RD(bs_num_env, 2);
      bs_num_env += 1;
      for (int rel = 0; rel < bs_num_env - 1; rel++) { GRID_RD(v, 2); c.bs_rel_bord[rel] = (uint8_t)(2 * v + 2); }
      GRID_RD(v, sbr_log2(bs_num_env + 1));
      c.bs_pointer = (uint8_t)v;
      for (int env = 0; env < bs_num_env; env++) { GRID_RD(v, 1); c.f[env] = (uint8_t)v; }
      c.L_E = (uint8_t)min(bs_num_env, 4);
      c.abs_bord_lead = (uint8_t)bs_abs_bord;
      c.abs_bord_trail = numTimeSlots;
      break;
    }
    default: {
      int bs_abs_bord, bs_abs_bord_1, n0, n1;
      GRID_RD(bs_abs_bord, 2);
      GRID_RD(bs_abs_bord_1, 2);
      bs_abs_bord_1 += numTimeSlots;
      GRID_RD(n0, 2);
      c.bs_num_rel_0 = (uint8_t)n0;
      GRID_RD(n1, 2);
      c.bs_num_rel_1 = (uint8_t)n1;
      const int bs_num_env = min(5, n0 + n1 + 1);
      for (int rel = 0; rel < n0; rel++) { GRID_RD(v, 2); c.bs_rel_bord_0[rel] = (uint8_t)(2 * v + 2); }
      for (int rel = 0; rel < n1; rel++) { GRID_RD(v, 2); c.bs_rel_bord_1[rel] = (uint8_t)(2 * v + 2); }
      GRID_RD(v, sbr_log2(n0 + n1 + 2));
      c.bs_pointer = (uint8_t)v;
      for (int env = 0; env < bs_num_env; env++) { GRID_RD(v, 1); c.f[env] = (uint8_t)v; }
      c.L_E = (uint8_t)min(bs_num_env, 5);
      c.abs_bord_lead = (uint8_t)bs_abs_bord;
      c.abs_bord_trail = (uint8_t)bs_abs_bord_1;
      break;
    }
  }
#undef GRID_RD
  if (c.L_E <= 0) return 1;
  c.L_Q = (c.L_E > 1) ? 2 : 1;
  const int r = sbr_time_border_vector(c);
  if (r != 0) {
    if (r > 0) { c.frame_class = (uint8_t)saved_class; c.L_E = (uint8_t)saved_L_E; c.L_Q = (uint8_t)saved_L_Q; }
    return r;
  }
  // noise_floor_time_border_vector (:529-541)
  c.t_Q[0] = c.t_E[0];
  if (c.L_E == 1) { c.t_Q[1] = c.t_E[1]; c.t_Q[2] = 0; }
  else { const int index = sbr_middle_border(c); c.t_Q[1] = c.t_E[index]; c.t_Q[2] = c.t_E[c.L_E]; }
  return 0;
}

__device__ inline int sbr_dtdf(SbrBits& ld, SbrChanParse& c) {  // :85-94
  int v;
  for (int i = 0; i < c.L_E; i++) { SBR_RD(v, 1); c.bs_df_env[i] = (uint8_t)v; }
  for (int i = 0; i < c.L_Q; i++) { SBR_RD(v, 1); c.bs_df_noise[i] = (uint8_t)v; }
  return 0;
}

__device__ inline int sbr_invf_mode(SbrBits& ld, SbrChanParse& c, int N_Q) {  // :97-101
  int v;
  for (int i = 0; i < N_Q; i++) { SBR_RD(v, 2); c.bs_invf_mode[i] = (uint8_t)v; }
  return 0;
}

// Channel.decodeHuffman (:280-289) / ps Huffman.read (ps/Huffman.java:264-276): a binary tree walked one bit per node.  The
// walk through the bounds-checked reader was 45 % of this kernel's instructions (round-1 profile), and the codes are short,
// so the first eight bits go through a 256-entry table built from the same tree on the host (build_huff_lut in
// jaadb_engine.cu): a leaf within eight bits -> [23:16] code length, [15:0] the decoded value; otherwise bit 31 and the node
// the walk continues from.  With fewer than eight bits left in the payload the bit-serial walk runs from the root, so the
// end-of-stream behaviour is the reference's bit for bit.
__device__ __forceinline__ int huff_decode(SbrBits& ld, const int16_t* __restrict__ t, const uint32_t* __restrict__ lut, int bias, int& out) {
  int index = 0;
  if (ld.left() >= 8u) {
    const uint32_t e = k3_ld_keep(lut + (ld.peek32() >> 24));
    if (!(e & 0x80000000u)) {
      ld.pos += e >> 16;
      out = (int)(int16_t)(e & 0xFFFFu);
      return 0;
    }
    ld.pos += 8;
    index = (int)(e & 0xFFFFu);
  }
  while (index >= 0) {
    int bit;
    SBR_RD(bit, 1);
    index = k3_ld_keep(t + index * 2 + bit);
  }
  out = index + bias;
  return 0;
}

__device__ __forceinline__ int sbr_huff(SbrBits& ld, const SbrTablesDev& T, int table, int& out) {
  return huff_decode(ld, T.huff[table], T.huff_lut + 256 * table, 64, out);
}

__device__ inline void sbr_extract_envelope(const SbrElemDev& S, SbrChanParse& c) {  // :192-240
  for (int l = 0; l < c.L_E; l++) {
    const int nb = S.n[c.f[l]];
    if (c.bs_df_env[l] == 0) {
      for (int k = 1; k < nb; k++) {
        int v = c.E[k - 1][l] + c.E[k][l];
        if (v < 0) v = 0;
        c.E[k][l] = (int16_t)v;
      }
    } else {
      const int g = (l == 0) ? c.f_prev : c.f[l - 1];
      if (c.f[l] == g) {
        for (int k = 0; k < nb; k++) {
          const int prev = l == 0 ? c.E_prev[k] : c.E[k][l - 1];
          c.E[k][l] = (int16_t)(prev + c.E[k][l]);
        }
      } else if ((g == 1) && (c.f[l] == 0)) {
        for (int k = 0; k < nb; k++)
          for (int i = 0; i < S.N_high; i++)
            if (S.f_table_res[SBR_HI_RES][i] == S.f_table_res[SBR_LO_RES][k]) {
              const int prev = l == 0 ? c.E_prev[i] : c.E[i][l - 1];
              c.E[k][l] = (int16_t)(prev + c.E[k][l]);
            }
      } else if ((g == 0) && (c.f[l] == 1)) {
        for (int k = 0; k < nb; k++)
          for (int i = 0; i < S.N_low; i++)
            if ((S.f_table_res[SBR_LO_RES][i] <= S.f_table_res[SBR_HI_RES][k]) &&
                (S.f_table_res[SBR_HI_RES][k] < S.f_table_res[SBR_LO_RES][i + 1])) {
              const int prev = l == 0 ? c.E_prev[i] : c.E[i][l - 1];
              c.E[k][l] = (int16_t)(prev + c.E[k][l]);
            }
      }
    }
  }
}

__device__ inline int sbr_envelope(SbrBits& ld, const SbrCtx& C, SbrChanParse& c, bool coupled) {  // :126-190
  const SbrElemDev& S = *C.S;
  if ((c.L_E == 1) && (c.frame_class == SBR_FIXFIX)) c.amp_res = 0;
  else c.amp_res = S.hdr.amp_res;
  const int delta = coupled ? 1 : 0;
  int t_huff, f_huff;   // table numbers (SbrTablesDev::huff)
  if (coupled) { t_huff = c.amp_res ? 6 : 2; f_huff = c.amp_res ? 7 : 3; }
  else { t_huff = c.amp_res ? 4 : 0; f_huff = c.amp_res ? 5 : 1; }
  for (int env = 0; env < c.L_E; env++) {
    const int nb = S.n[c.f[env]];
    int v, band = 0, table = t_huff;
    if (c.bs_df_env[env] == 0) {
      const int bits = coupled ? (c.amp_res ? 5 : 6) : (c.amp_res ? 6 : 7);
      SBR_RD(v, bits);
      c.E[0][env] = (int16_t)(v << delta);
      band = 1;
      table = f_huff;
    }
    const int16_t* tree = C.T->huff[table];
    const uint32_t* lut = C.T->huff_lut + 256 * table;
#pragma unroll 1
    for (; band < nb; band++) { SBR_TRY(huff_decode(ld, tree, lut, 64, v)); c.E[band][env] = (int16_t)(v << delta); }
  }
  sbr_extract_envelope(S, c);
  return 0;
}

__device__ inline int sbr_noise(SbrBits& ld, const SbrCtx& C, SbrChanParse& c, bool coupled) {  // :243-312
  const SbrElemDev& S = *C.S;
  const int delta = coupled ? 1 : 0;
  const int t_huff = coupled ? 9 : 8, f_huff = coupled ? 7 : 5;
  for (int noise = 0; noise < c.L_Q; noise++) {
    int v, band = 0, table = t_huff;
    if (c.bs_df_noise[noise] == 0) {
      SBR_RD(v, 5);
      c.Q[0][noise] = (int16_t)(v << delta);
      band = 1;
      table = f_huff;
    }
    const int16_t* tree = C.T->huff[table];
    const uint32_t* lut = C.T->huff_lut + 256 * table;
#pragma unroll 1
    for (; band < S.N_Q; band++) { SBR_TRY(huff_decode(ld, tree, lut, 64, v)); c.Q[band][noise] = (int16_t)(v << delta); }
  }
  for (int l = 0; l < c.L_Q; l++) {
    if (c.bs_df_noise[l] == 0) {
      for (int k = 1; k < S.N_Q; k++) c.Q[k][l] = (int16_t)(c.Q[k][l] + c.Q[k - 1][l]);
    } else if (l == 0) {
      for (int k = 0; k < S.N_Q; k++) c.Q[k][l] = (int16_t)(c.Q_prev[k] + c.Q[k][0]);
    } else {
      for (int k = 0; k < S.N_Q; k++) c.Q[k][l] = (int16_t)(c.Q[k][l - 1] + c.Q[k][l]);
    }
  }
  return 0;
}

__device__ inline int sbr_sinusoidal(SbrBits& ld, SbrChanParse& c, int N_high) {  // SBR.java:249-254
  int v;
  for (int i = 0; i < N_high; i++) { SBR_RD(v, 1); c.bs_add_harmonic[i] = (uint8_t)v; }
  return 0;
}

__device__ inline int sbr_harmonics(SbrBits& ld, SbrChanParse& c, int N_high) {
  int v;
  SBR_RD(v, 1);
  c.add_harmonic_flag = (uint8_t)v;
  if (v) SBR_TRY(sbr_sinusoidal(ld, c, N_high));
  return 0;
}

// ---- parametric stereo payload (ps/PSImpl.java:103-135, ICData.java:21-32, EnvData.java:44-49, Envelope.java:24-30) --
__device__ inline int ps_nr_par(int id) { return id % 3 == 0 ? 10 : (id % 3 == 1 ? 20 : 34); }
__device__ inline int ps_stride(int id) { return (id % 3) == 0 ? 2 : 0; }   // ICMode.stride

__device__ __forceinline__ int ps_huff(SbrBits& ld, const SbrTablesDev& T, int table, int& out) {  // ps/Huffman.java:264-276
  return huff_decode(ld, T.ps_huff[table], T.ps_huff_lut + 256 * table, 31, out);
}

__device__ inline int ps_read_mode(SbrBits& ld, PsParamDev& p) {
  int en, id;
  SBR_RD(en, 1);
  if (en) {
    SBR_RD(id, 3);
    if (id > 5) return JAADB_ST_ARRAY_BOUNDS;   // IID_MODES[id] / ICC_MODES[id]
    p.mode = (int8_t)id;
  } else p.mode = -1;
  return 0;
}

__device__ inline int ps_read_data(SbrBits& ld, const SbrTablesDev& T, PsParamDev& p, bool icc, int num_env) {
  if (p.mode < 0) return 0;
  const int nr = ps_nr_par(p.mode);
  for (int n = 0; n < num_env; n++) {
    int dt, v;
    SBR_RD(dt, 1);
    p.dt[n] = (uint8_t)dt;
    const int h = icc ? (dt ? 5 : 4) : (p.mode < 3 ? (dt ? 1 : 0) : (dt ? 3 : 2));
    for (int i = 0; i < nr; i++) { SBR_TRY(ps_huff(ld, T, h, v)); p.index[n][i] = (int8_t)v; }
  }
  return 0;
}

__device__ inline int ps_pd_nr_par(int id) { return id % 3 == 0 ? 5 : (id % 3 == 1 ? 11 : 17); }   // PDMode.java:50-57

// EnvData.readData for IPD / OPD
__device__ inline int ps_read_pd(SbrBits& ld, const SbrTablesDev& T, PsPdDev& p, bool opd, int num_env) {
  if (p.mode < 0) return 0;
  const int nr = ps_pd_nr_par(p.mode);
  for (int n = 0; n < num_env; n++) {
    int dt, v;
    SBR_RD(dt, 1);
    p.dt[n] = (uint8_t)dt;
    const int h = (opd ? 8 : 6) + (dt ? 1 : 0);
    for (int i = 0; i < nr; i++) { SBR_TRY(ps_huff(ld, T, h, v)); p.index[n][i] = (int8_t)v; }
  }
  return 0;
}

// PSImpl.decode
__device__ inline int ps_decode(SbrBits& ld, const SbrTablesDev& T, PsParseDev& P) {
  int v;
  SBR_RD(v, 1);
  if (v) {
    P.header_read = 1;
#pragma unroll 1
    for (int k = 0; k < 2; ++k) SBR_TRY(ps_read_mode(ld, k ? P.icc : P.iid));
    // Extension.readMode (ps/Extension.java:31-38)
    SBR_RD(v, 1);
    P.ext_enabled = (uint8_t)v;
    if (v) P.ext_has_data = 1;
    if (P.ext_has_data) P.ipd.mode = P.opd.mode = v ? P.iid.mode : (int8_t)-1;   // ExtData.setMode(enabled ? parent.mode : null)
  }
  SBR_RD(v, 1);
  P.var_borders = (uint8_t)v;
  int tmp;
  SBR_RD(tmp, 2);
  const int num_env = P.var_borders ? tmp + 1 : (tmp == 3 ? 4 : tmp);   // num_env_tab (PSTables.java:16-19)
  P.num_env = (uint8_t)num_env;
  if (P.var_borders)
    for (int n = 1; n < num_env + 1; n++) { SBR_RD(v, 5); P.border_position[n] = (uint8_t)(v + 1); }
#pragma unroll 1
  for (int k = 0; k < 2; ++k) SBR_TRY(ps_read_data(ld, T, k ? P.icc : P.iid, k == 1, num_env));
  if (P.ext_enabled) {
    // Extension.readData (ps/Extension.java:40-59): cnt bytes in a sub-stream; extension id 0 is ExtData.readData
    // (ps/ExtData.java:17-25), every other id only consumes its two bits
    int cnt;
    SBR_RD(cnt, 4);
    if (cnt == 15) { SBR_RD(v, 8); cnt += v; }
    if (ld.left() < (uint32_t)(8 * cnt)) return JAADB_ST_EOS;
    SbrBits sub = ld;
    sub.end = ld.pos + 8 * cnt;
    while (sub.left() > 7) {
      int id;
      if (!sub.get(2, id)) return JAADB_ST_EOS;
      if (id == 0 && P.ext_has_data) {
        if (!sub.get(1, v)) return JAADB_ST_EOS;
        P.ext_data_enabled = (uint8_t)v;
        if (v) {
#pragma unroll 1
          for (int k = 0; k < 2; ++k) SBR_TRY(ps_read_pd(sub, T, k ? P.opd : P.ipd, k == 1, num_env));
        }
        if (!sub.get(1, v)) return JAADB_ST_EOS;
      }
    }
    ld.pos += 8 * cnt;
  }
  P.data_available = 1;
  return 0;
}

// ps_data_decode and what it calls run on the whole warp: every lane reads the same fields of the element's state in
// shared memory, so the control flow is uniform; element-wise steps (time-differential rows, the stride expansion, the
// copies) take one lane per parameter, the frequency-differential running sums stay on lane 0.  Rows are finished with a
// __syncwarp() before anything reads them.

// Envelope.decode (:45-74) for one envelope of one parameter set
__device__ inline void ps_decode_env(PsParamDev& p, bool icc, int env, int lane) {
  int8_t* ix = p.index[env];
  if (p.mode < 0) {
    if (lane == 0) p.dt[env] = 0;
    for (int i = lane; i < 34; i += 32) ix[i] = 0;
    __syncwarp();
    return;
  }
  const int st = ps_stride(p.mode), nr = ps_nr_par(p.mode);
  const int lim = icc ? 7 : (p.mode < 3 ? 7 : 15);
  const int lo = icc ? 0 : -lim;
  const int8_t* prev = env == 0 ? p.first : p.index[env - 1];
  if (p.dt[env]) {
    for (int i = lane; i < nr; i += 32) ix[i] = (int8_t)min(max(prev[i * st] + ix[i], lo), lim);
  } else if (lane == 0) {
    int pc = ix[0];
#pragma unroll 1
    for (int i = 1; i < nr; i++) { pc = min(max(pc + ix[i], lo), lim); ix[i] = (int8_t)pc; }
  }
  __syncwarp();
  if (st > 1) {
    // for (i = st * nr - 1; i > 0; --i) ix[i] = ix[i / st]: every entry takes the value its source had before the loop
    // (the sources lie below the entries still to be written); st > 1 only comes with nr = 10, i.e. 20 entries
    const int n = st * nr;
    int8_t v = 0;
    if (lane < n) v = ix[lane / st];
    __syncwarp();
    if (lane > 0 && lane < n) ix[lane] = v;
    __syncwarp();
  }
}

// What ends a frame inside JAAD's HF generation / adjustment although its payload parsed (lane 0):
//  * hf_generation walks the patches of every processed frame (HFGeneration.java:61-70): a source or target band outside the
//    64 QMF bands -- patches that outlived a header change in a frame that failed -- ends the frame (the oracle's status);
//  * calculate_gain opens every envelope with get_S_mapped(ch, l, 0) (HFAdjustment.java:46-76,262): for a low-resolution
//    envelope it walks bs_add_harmonic from 2 * band - (N_high & 1), i.e. from index -1 when N_high is odd -- an
//    ArrayIndexOutOfBoundsException in JAAD (FAAD2's C reads the byte in front of the array).  Encoders do pair an odd
//    N_high with low-resolution envelopes; JAAD fails every such frame, and so does the engine.
__device__ __forceinline__ int sbr_process_errors(const SbrElemDev& S, int nch) {
  int k0 = S.kx;
  for (int i = 0; i < S.noPatches; ++i) {
    const int nsb = S.patchNoSubbands[i], p0 = S.patchStartSubband[i];
    if (nsb > 0 && (k0 < 0 || k0 + nsb > 64 || p0 < 0 || p0 + nsb > 64)) return JAADB_ST_SBR;
    k0 += nsb;
  }
  if (S.N_high & 1)
    for (int c = 0; c < nch; ++c)
      for (int l = 0; l < S.ch[c].L_E; ++l)
        if (S.ch[c].f[l] == SBR_LO_RES) return JAADB_ST_ARRAY_BOUNDS;
  return 0;
}

// PSImpl.ps_data_decode (:137-199) -> the frame record K4 mixes with.  Returns Extension.nr_par as the record holds it.
__device__ inline int ps_data_decode(PsParseDev& P, PsFrameDev& o, int lane) {
  int num_env = P.data_available ? P.num_env : 0;
  PsParamDev* ps[2] = {&P.iid, &P.icc};
  for (int k = 0; k < 2; ++k) {
    PsParamDev& p = *ps[k];
    if (num_env == 0) {
      const bool on = p.mode >= 0;
      if (!on && lane == 0) p.dt[0] = 0;
      for (int i = lane; i < 34; i += 32) p.index[0][i] = on ? p.first[i] : (int8_t)0;
      __syncwarp();
    } else {
      for (int env = 0; env < num_env; env++) ps_decode_env(p, k == 1, env, lane);
    }
  }
  // Extension.decode / ExtData.decode (ps/Extension.java:61-64, ExtData.java:27-32) with PDMode: stride 1, clip = idx & 7
  const bool ext_live = P.ext_enabled && P.ext_has_data;
  PsPdDev* pd[2] = {&P.ipd, &P.opd};
  if (ext_live && P.ext_data_enabled) {
    for (int k = 0; k < 2; ++k) {
      PsPdDev& p = *pd[k];
      const bool on = p.mode >= 0;
      if (num_env == 0) {
        if (!on && lane == 0) p.dt[0] = 0;
        if (lane < 17) p.index[0][lane] = on ? p.first[lane] : (int8_t)0;
        __syncwarp();
      } else {
        for (int env = 0; env < num_env; env++) {
          int8_t* ix = p.index[env];
          if (!on) {
            if (lane == 0) p.dt[env] = 0;
            if (lane < 17) ix[lane] = 0;
          } else {
            const int nr = ps_pd_nr_par(p.mode);
            const int8_t* prev = env == 0 ? p.first : p.index[env - 1];
            if (p.dt[env]) {
              if (lane < nr) ix[lane] = (int8_t)((prev[lane] + ix[lane]) & 7);
            } else if (lane == 0) {
              int pc = ix[0];
#pragma unroll 1
              for (int i = 1; i < nr; i++) { pc = (pc + ix[i]) & 7; ix[i] = (int8_t)pc; }
            }
          }
          __syncwarp();
        }
      }
    }
  }
  if (num_env == 0) num_env = 1;
  for (int k = 0; k < 2; ++k)
    for (int i = lane; i < 34; i += 32) ps[k]->first[i] = ps[k]->index[num_env - 1][i];
  // ExtData.update runs whether or not the frame carried phase data (ps/ExtData.java:34-37); ExtData.restore in the
  // variable-border branch below calls update as well (:39-42), i.e. changes nothing more
  if (ext_live && lane < 17)
    for (int k = 0; k < 2; ++k) pd[k]->first[lane] = pd[k]->index[num_env - 1][lane];
  const int L = 32;
  const bool var_borders = P.var_borders != 0;
  const bool restore = var_borders && P.border_position[num_env] < L;
  __syncwarp();   // (every lane has read data_available / border_position[num_env] before lane 0 moves them)
  if (restore)
    for (int k = 0; k < 2; ++k)
      for (int i = lane; i < 34; i += 32) ps[k]->index[num_env][i] = ps[k]->index[num_env - 1][i];   // Envelope.restore
  if (lane == 0) {
    P.data_available = 0;
    if (!var_borders) {
      P.border_position[0] = 0;
#pragma unroll 1
      for (int env = 1; env < num_env; env++) P.border_position[env] = (uint8_t)((env * L) / num_env);
      P.border_position[num_env] = L;
    } else {
      P.border_position[0] = 0;
      int ne = num_env;
      if (restore) { ++ne; P.border_position[ne] = L; }
      int bpl = P.border_position[0];
#pragma unroll 1
      for (int env = 1; env < ne; env++) {
        const int bp = P.border_position[env];
        const int mx = L - (ne - env);
        bpl = min(max(bp, bpl + 1), mx);
        if (bpl != bp) P.border_position[env] = (uint8_t)bpl;
      }
    }
  }
  if (restore) ++num_env;
  // Extension.nr_par (ps/Extension.java:81-86, ExtData.java:49-54).  255: the extension is on while IID is off -- JAAD
  // dereferences the null PDMode (NullPointerException); the caller fails the frame.  254: ps_index_oob
  const int nr_ipdopd = !ext_live ? 0 : (P.ipd.mode < 0 ? 255 : max(ps_pd_nr_par(P.ipd.mode), 11));
  __syncwarp();
  if (lane == 0) {
    P.num_env = (uint8_t)num_env;
    o.use_ps = 1;
    o.num_env = (uint8_t)num_env;
    o.iid_mode = P.iid.mode < 0 ? 0 : P.iid.mode;
    o.icc_mode = P.icc.mode < 0 ? 1 : P.icc.mode;
    o.nr_ipdopd_par = (uint8_t)nr_ipdopd;
    o.enable_ipdopd = P.ext_data_enabled;
  }
  if (lane < 6) o.border[lane] = P.border_position[lane];
  // (ps_mix_phase indexes its tables with |iid| <= num_steps and 0 <= icc <= 7 for every parameter band of every envelope,
  // PSImpl.java:424-478; the delta decoding clips what it adds up, but not a frequency-differential row's first value, and a
  // header that goes from the fine to the coarse IID grid leaves the carried row where it was: an index past the tables is an
  // ArrayIndexOutOfBoundsException in JAAD -- 254: the caller fails the frame)
  const int num_steps = P.iid.mode >= 3 ? 15 : 7;
  bool oob = false;
  for (int idx = lane; idx < 5 * 20; idx += 32) {
    const int env = idx / 20, i = idx - env * 20;
    const int a = P.iid.index[env][i], b = P.icc.index[env][i];
    o.iid[env][i] = (int8_t)a;
    o.icc[env][i] = (int8_t)b;
    oob |= env < num_env && (a > num_steps || a < -num_steps || (unsigned)b > 7u);
  }
  for (int idx = lane; idx < 5 * 17; idx += 32) {
    const int env = idx / 17, i = idx - env * 17;
    o.ipd[env][i] = P.ipd.index[env][i];
  }
#ifndef K3_NO_PSCHK
  if (nr_ipdopd != 255 && __any_sync(0xFFFFFFFFu, oob)) return 254;
#endif
  return nr_ipdopd;
}

// SBR.readExtendedData (:229-242).  Extension payloads (parametric stereo = id 2) are skipped in this build: the
// engine refuses SBR+PS streams at open, and for plain SBR streams the reference's sbr_extension is a no-op that
// re-reads 2-bit ids until fewer than 8 bits remain -- nothing observable.
__device__ inline int sbr_extended_data(SbrBits& ld, const SbrTablesDev& T, PsParseDev* ps) {
  int v;
  SBR_RD(v, 1);
  if (v) {
    int cnt;
    SBR_RD(cnt, 4);
    if (cnt == 15) { SBR_RD(v, 8); cnt += v; }
    if (ld.left() < (uint32_t)(8 * cnt)) return JAADB_ST_EOS;
    if (ps) {
      // SBR1.sbr_extension (:62-73): extension id 2 is parametric stereo; everything else only consumes its 2-bit id
      SbrBits sub = ld;
      sub.end = ld.pos + 8 * cnt;
      while (sub.left() > 7) {
        int id;
        if (!sub.get(2, id)) return JAADB_ST_EOS;
        if (id == 2) {
          ps->opened = 1;
          const int st = ps_decode(sub, T, *ps);
          if (st) return st;
        }
      }
    }
    ld.pos += 8 * cnt;
  }
  return 0;
}

// SBR1.sbr_data (:34-60) / SBR2.sbr_data (:35-135).  `valid` as SBR.decode sets it.
// The three element layouts (single channel, coupled pair, independent pair) are step lists over ONE expansion of each
// syntax function: with every call written out, the kernel carried five copies of the envelope / noise readers and three
// of the grid (230 KB of SASS, two of each on the path of a stereo frame), and it is instruction-fetch bound.
enum { K3_END = 0, K3_GRID, K3_DTDF, K3_INVF, K3_COUPLE, K3_ENV, K3_NOISE, K3_ZERO, K3_HARM, K3_DEQ, K3_EXT, K3_SAVE };
#define K3S(op, ch, flag) (uint8_t)((op) | ((ch) << 4) | ((flag) << 5))
__constant__ uint8_t c_k3_steps[3][20] = {
    // SBR1.sbr_data: dequantChannel comes before the harmonics; the extension may carry parametric stereo (flag)
    {K3S(K3_GRID, 0, 0), K3S(K3_DTDF, 0, 0), K3S(K3_INVF, 0, 0), K3S(K3_ENV, 0, 0), K3S(K3_NOISE, 0, 0), K3S(K3_DEQ, 0, 0),
     K3S(K3_ZERO, 0, 0), K3S(K3_HARM, 0, 0), K3S(K3_EXT, 0, 1), K3_END},
    // SBR2.sbr_data, bs_coupling: the second channel's dt/df flags are read with its OLD L_E / L_Q (the grid is copied
    // over only afterwards, Channel.couple), and its envelopes / noise floors use the balance tables (flag)
    {K3S(K3_GRID, 0, 0), K3S(K3_DTDF, 0, 0), K3S(K3_DTDF, 1, 0), K3S(K3_INVF, 0, 0), K3S(K3_COUPLE, 0, 0), K3S(K3_ENV, 0, 0),
     K3S(K3_NOISE, 0, 0), K3S(K3_ENV, 1, 1), K3S(K3_NOISE, 1, 1), K3S(K3_ZERO, 0, 0), K3S(K3_ZERO, 1, 0), K3S(K3_HARM, 0, 0),
     K3S(K3_HARM, 1, 0), K3S(K3_DEQ, 0, 0), K3S(K3_EXT, 0, 0), K3_END},
    // SBR2.sbr_data, independent channels: a second grid that fails puts the first channel's grid back (flag)
    {K3S(K3_SAVE, 0, 0), K3S(K3_GRID, 0, 0), K3S(K3_GRID, 1, 1), K3S(K3_DTDF, 0, 0), K3S(K3_DTDF, 1, 0), K3S(K3_INVF, 0, 0),
     K3S(K3_INVF, 1, 0), K3S(K3_ENV, 0, 0), K3S(K3_ENV, 1, 0), K3S(K3_NOISE, 0, 0), K3S(K3_NOISE, 1, 0), K3S(K3_ZERO, 0, 0),
     K3S(K3_ZERO, 1, 0), K3S(K3_HARM, 0, 0), K3S(K3_HARM, 1, 0), K3S(K3_DEQ, 0, 0), K3S(K3_EXT, 0, 0), K3_END}};
#undef K3S

__device__ inline int sbr_data(SbrBits& ld, const SbrCtx& C, bool stereo, bool with_ps, int& result) {
  SbrElemDev& S = *C.S;
  int v, eos;
  result = 0;
  SBR_RD(v, 1);
  if (v) { SBR_RD(v, 4); if (stereo) SBR_RD(v, 4); }
  int layout = 0;
  if (stereo) {
    SBR_RD(v, 1);
    S.bs_coupling = (uint8_t)v;
    layout = v ? 1 : 2;
  }
  uint8_t saved_t_E[6] = {0, 0, 0, 0, 0, 0}, saved_t_Q[3] = {0, 0, 0};
  int saved_L_E = 0, saved_L_Q = 0, saved_class = 0;
  const uint8_t* steps = c_k3_steps[layout];
#pragma unroll 1
  for (int pc = 0;; ++pc) {
    const int step = steps[pc];
    const int op = step & 15, flag = step >> 5;
    SbrChanParse& c = S.ch[(step >> 4) & 1];
    if (op == K3_END) break;
    switch (op) {
      case K3_SAVE: {
        saved_L_E = c.L_E; saved_L_Q = c.L_Q; saved_class = c.frame_class;
        for (int i = 0; i < saved_L_E && i < 6; i++) saved_t_E[i] = c.t_E[i];
        for (int i = 0; i < saved_L_Q && i < 3; i++) saved_t_Q[i] = c.t_Q[i];
        break;
      }
      case K3_GRID: {
        const int r = sbr_grid(ld, c, eos);
        if (r < 0) return -r;
        if (r > 0) {
          if (flag) {
            SbrChanParse& c0 = S.ch[0];
            c0.frame_class = (uint8_t)saved_class;
            c0.L_E = (uint8_t)saved_L_E;
            c0.L_Q = (uint8_t)saved_L_Q;
            for (int i = 0; i < 6; i++) c0.t_E[i] = saved_t_E[i];
            for (int i = 0; i < 3; i++) c0.t_Q[i] = saved_t_Q[i];
          }
          result = r;
          return 0;
        }
        break;
      }
      case K3_DTDF: SBR_TRY(sbr_dtdf(ld, c)); break;
      case K3_INVF: SBR_TRY(sbr_invf_mode(ld, c, S.N_Q)); break;
      case K3_COUPLE: {
        // Channel.couple (:103-122)
        const SbrChanParse& c0 = S.ch[0];
        SbrChanParse& c1 = S.ch[1];
        c1.frame_class = c0.frame_class;
        c1.L_E = c0.L_E;
        c1.L_Q = c0.L_Q;
        c1.bs_pointer = c0.bs_pointer;
        for (int i = 0; i <= c0.L_E; i++) { c1.t_E[i] = c0.t_E[i]; c1.f[i] = c0.f[i]; }
        for (int i = 0; i <= c0.L_Q; i++) c1.t_Q[i] = c0.t_Q[i];
        for (int i = 0; i < S.N_Q; i++) c1.bs_invf_mode[i] = c0.bs_invf_mode[i];
        break;
      }
      case K3_ENV: SBR_TRY(sbr_envelope(ld, C, c, flag != 0)); break;
      case K3_NOISE: SBR_TRY(sbr_noise(ld, C, c, flag != 0)); break;
      case K3_ZERO: {
        uint32_t* h = reinterpret_cast<uint32_t*>(c.bs_add_harmonic);
#pragma unroll 1
        for (int i = 0; i < 16; ++i) h[i] = 0u;
        break;
      }
      case K3_HARM: SBR_TRY(sbr_harmonics(ld, c, S.N_high)); break;
      case K3_DEQ:
        // NoiseEnvelope.dequantChannel (x 2, or unmap: SBR2.java:128-133) happens here in the reference; the float tables
        // are evaluated by all lanes after the syntax (same inputs: E, Q, amp_res, f, n) -- also when the rest of the
        // payload fails
        S.dequant = 1;
        break;
      default:   // K3_EXT
        SBR_TRY(sbr_extended_data(ld, *C.T, (flag && with_ps) ? &S.ps : nullptr));
        break;
    }
  }
  return 0;
}

// SBR.decode (:161-185).  Returns a frame status (0 = fine).
__device__ inline int sbr_decode(SbrBits& ld, const SbrCtx& C, bool stereo, bool with_ps, bool crc) {
  SbrElemDev& S = *C.S;
  int v;
  if (crc) SBR_RD(v, 10);
  // readHeader (:214-223)
  SBR_RD(v, 1);
  bool reset = false;
  if (v) {
    sbr_swap_headers(S);
    SBR_TRY(sbr_header_decode(ld, S.hdr));
    S.hdr.present = 1;
    reset = sbr_header_differs(S.hdr, S.hdr_saved);
  }
  S.reset = reset ? 1 : 0;
  if (reset) {
    int rt = sbr_calc_tables(C);
    if (rt < 0) return -rt;
    if (rt > 0) {
      sbr_swap_headers(S);
      rt = sbr_calc_tables(C);
      if (rt < 0) return -rt;
    }
  }
  if (S.hdr.present) {
    int result = 0;
    SBR_TRY(sbr_data(ld, C, stereo, with_ps, result));
    S.valid = (result == 0) ? 1 : 0;
  } else {
    S.valid = 1;
  }
  return 0;
}

// ---- NoiseEnvelope.java: dequantisation of one (band, envelope) / (band, noise floor) entry ---------------------
__device__ inline float sbr_e_orig(const SbrTablesDev& T, const SbrElemDev& S, bool stereo, int ch, int k, int l) {
  if (stereo && S.bs_coupling) {
    // unmap (:299-345)
    const int amp0 = S.ch[0].amp_res ? 0 : 1, amp1 = S.ch[1].amp_res ? 0 : 1;
    const int ch0E = S.ch[0].E[k][l];
    const int exp0 = (ch0E >> amp0) + 1;
    const int exp1 = (S.ch[1].E[k][l] >> amp1);
    if ((exp0 < 0) || (exp0 >= 64) || (exp1 < 0) || (exp1 > 24)) return 0.f;
    float tmp = T.e_deq[exp0];
    if (amp0 != 0 && (ch0E & 1) != 0) tmp = (float)((double)tmp * 1.414213562);   // `tmp *= 1.414213562` (double literal)
    return tmp * T.e_pan[ch == 0 ? exp1 : 24 - exp1];
  }
  // dequantChannel (:250-281)
  const SbrChanParse& c = S.ch[ch];
  const int amp = c.amp_res ? 0 : 1;
  const int e = c.E[k][l];
  const int exp = e >> amp;
  if ((exp < 0) || (exp >= 64)) return 0.f;
  float v = T.e_deq[exp];
  if (amp != 0 && (e & 1) != 0) v = v * 1.414213562f;
  return v;
}

__device__ inline void sbr_q_div(const SbrTablesDev& T, const SbrElemDev& S, bool stereo, int ch, int k, int l, float& qd, float& qd2) {
  if (stereo && S.bs_coupling) {
    const int q0 = S.ch[0].Q[k][l], q1 = S.ch[1].Q[k][l];
    if ((q0 < 0 || q0 > 30) || (q1 < 0 || q1 > 24)) { qd = 0.f; qd2 = 0.f; return; }
    qd = (ch == 0 ? T.q_div_left : T.q_div_right)[q0 * 13 + (q1 >> 1)];
    qd2 = (ch == 0 ? T.q_div2_left : T.q_div2_right)[q0 * 13 + (q1 >> 1)];
    return;
  }
  const int q = S.ch[ch].Q[k][l];
  if (q < 0 || q > 30) { qd = 0.f; qd2 = 0.f; return; }
  qd = T.q_div[q];
  qd2 = T.q_div2[q];
}

struct SbrConstTables {
  const int32_t *start_min, *offset_index, *offset, *stop_min, *stop_offset, *goal_sb;
  const float* limiter_cmp;
};

constexpr int kK3WarpsPerBlock = 4;
// a fill element carries at most 269 payload bytes (count 15 + 255 - 1): 68 words, one more for a start inside a word and
// one for the word the funnel shift reads past the end
constexpr int kK3StageWords = 72;
constexpr size_t k3_smem_bytes() { return (sizeof(SbrElemDev) + 4 * kK3StageWords) * kK3WarpsPerBlock; }

__global__ void __launch_bounds__(32 * kK3WarpsPerBlock)
k3_sbr_parse_kernel(const uint8_t* __restrict__ blob, const FrameDev* __restrict__ frames, FrameSide* __restrict__ fside,
                    const SbrRunDev* __restrict__ runs, uint32_t n_runs, const RunFrameDev* __restrict__ run_frames,
                    SbrElemDev* __restrict__ elems, SbrFrameDev* __restrict__ out, PsFrameDev* __restrict__ ps_out,
                    SbrTablesDev T, SbrConstTables K) {
  extern __shared__ __align__(16) uint8_t k3_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t r = blockIdx.x * kK3WarpsPerBlock + warp;
  if (r >= n_runs) return;
  const SbrRunDev run = runs[r];
  SbrElemDev* S = reinterpret_cast<SbrElemDev*>(k3_smem) + warp;
  uint32_t* stage = reinterpret_cast<uint32_t*>(k3_smem + sizeof(SbrElemDev) * kK3WarpsPerBlock) + warp * kK3StageWords;
  SbrElemDev* G = elems + (size_t)run.stream_slot * 2 + run.element;
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(G);
    uint32_t* dst = reinterpret_cast<uint32_t*>(S);
    for (int i = lane; i < (int)(sizeof(SbrElemDev) / 4); i += 32) dst[i] = k3_ld_stream_rw(src + i);
  }
  __syncwarp();
  SbrCtx C;
  C.S = S; C.T = &T; C.start_min = K.start_min; C.offset_index = K.offset_index; C.offset = K.offset; C.stop_min = K.stop_min;
  C.stop_offset = K.stop_offset; C.goal_sb = K.goal_sb; C.limiter_cmp = K.limiter_cmp; C.sr_index = run.sr_index;
  const bool stereo = run.stereo != 0;
  const int nch = stereo ? 2 : 1;

  // bookkeeping for the frame-parallel K4: ordinal of / distances between the frames that run the QMF banks (and the PS tool)
  uint32_t n_proc = 0;
  int64_t last_proc = -1, last_ps = -1;
  for (uint32_t it = 0; it < run.count; ++it) {
    const uint32_t f = run_frames[run.first + it].frame;
    int mode = 0;          // what K4 does with the frame (SbrFrameDev.mode)
    int frame_status = 0;
    int dequant = 0;       // NoiseEnvelope.dequantChannel / unmap ran for this frame (lane 0 decides, all lanes do it)
    // ---- the frame's payload for this element, staged by the warp (SbrBits); a frame without one stages nothing
    const FrameSide fs = fside[f];
    const uint32_t pay_bits = (run.element < 2) ? fs.sbr_bits[run.element] : 0u;
    const uint32_t pay_off = (run.element < 2) ? fs.sbr_bit_off[run.element] : 0u;
    if (pay_bits) {
      const uint64_t addr = reinterpret_cast<uint64_t>(blob) + frames[f].blob_off;
      const uint32_t* gw = reinterpret_cast<const uint32_t*>(addr - (addr & 3u)) + (pay_off >> 5);
      const uint32_t nw = min((((pay_off & 31u) + pay_bits + 31u) >> 5) + 1u, (uint32_t)kK3StageWords);
      for (uint32_t i = lane; i < nw; i += 32) stage[i] = __byte_perm(k3_ld_stream(gw + i), 0, 0x0123);
    }
    __syncwarp();
    if (lane == 0) {
      frame_status = fs.status;
      S->dequant = 0;
      // JAAD's element objects (and their SBR) are per instance tag: a frame that carries another tag does not touch
      // this one (StreamState::tags; K2 applies the same rule to the core coder's state)
      bool foreign = false;
      if (run.element < fs.n_started && run.element < 4) {
        const uint8_t tag = (uint8_t)((fs.tags >> (4 * run.element)) & 15u);
        if (S->tag_valid) foreign = tag != S->tag;
        else { S->tag = tag; S->tag_valid = 1; }
      }
      if (foreign && frame_status == 0) { frame_status = JAADB_ST_LAYOUT; fside[f].status = JAADB_ST_LAYOUT; }
      // ChannelElement.decode invalidates the element's SBR at the start of every frame (ChannelElement.java:58-61);
      // the element was reached iff K1 counted it
      if (!foreign && S->opened && run.element < fs.n_elements) S->valid = 0;
      const uint32_t nbits = foreign ? 0u : pay_bits;
      if (nbits) {
        // the FIL payload was seen by K1 (also in frames that failed later on): decodeSBR runs
        S->opened = 1;
        SbrBits ld;
        ld.words = stage;
        ld.pos = pay_off & 31u;
        // (a span beyond the staging buffer cannot come from a fill element; it would end as an early end of stream)
        ld.end = min(ld.pos + nbits, 32u * (uint32_t)(kK3StageWords - 1));
        int ext = 0;
        ld.get(4, ext);
        const int st = sbr_decode(ld, C, stereo, run.ps != 0, ext == 14);
        if (st != 0 && frame_status != JAADB_ST_LAYOUT) {
          // an exception inside SBR.decode fails the whole frame (EOS is swallowed by decodeFrame: no output).  It also
          // wins over an error K1 met: JAAD parses the payload when it reaches the fill element (decodeFIL,
          // SyntacticElements.java:169-203), and K1, which stops at its first error, only records a payload it got past
          // -- so whatever K1 reported happened later in the frame and JAAD never gets there.
          frame_status = st;
          fside[f].status = st;
        }
      }
      if (frame_status == 0 && S->opened && S->valid) {
        mode = S->hdr.present ? 2 : 1;
        if (mode == 2) {
          // hf_adjustment's l_A (HFAdjustment.java:24-38) for each channel
          for (int c = 0; c < nch; ++c) {
            SbrChanParse& cp = S->ch[c];
            if (cp.frame_class == SBR_FIXFIX) cp.l_A = -1;
            else if (cp.frame_class == SBR_VARFIX) cp.l_A = (cp.bs_pointer > 1) ? (int8_t)(cp.bs_pointer - 1) : (int8_t)-1;
            else cp.l_A = (cp.bs_pointer == 0) ? (int8_t)-1 : (int8_t)(cp.L_E + 1 - cp.bs_pointer);
          }
          if (S->reset) {
            // the reference builds patches and limiter bands inside the first hf_generation after a reset
            const int pst = sbr_patch_construction(C);
            if (pst != 0) { frame_status = pst; fside[f].status = pst; mode = 0; }
            else sbr_limiter_table(C);
          }
        }
      }
      dequant = S->dequant;
    }
    dequant = __shfl_sync(0xFFFFFFFFu, dequant, 0);
    mode = __shfl_sync(0xFFFFFFFFu, mode, 0);
    frame_status = __shfl_sync(0xFFFFFFFFu, frame_status, 0);
    __syncwarp();
#ifndef K3_NO_PROCCHK
    if (mode == 2 && frame_status == 0) {
      // (every lane: the element's tables are in shared memory, the answer is the same in all of them)
      const int xst = sbr_process_errors(*S, nch);
      if (xst != 0) {
        frame_status = xst;
        mode = 0;
        if (lane == 0) fside[f].status = xst;
      }
    }
#endif
    int use_ps = 0;
    if (run.ps) {
      // SBR1.process: parametric stereo runs iff this frame brought ps_data (SBR1.isPSUsed); PSImpl.ps_data_decode
      PsFrameDev* po = ps_out + run.ps_base + it;
      if (mode != 0 && S->ps.opened && S->ps.data_available) {
        const int nr_par = ps_data_decode(S->ps, *po, lane);
        use_ps = 1;
        if (nr_par >= 254) {
          // JAAD dies of a NullPointerException (255) or of an index past its tables (254) inside ps_mix_phase (see
          // ps_data_decode): the frame fails
          use_ps = 0;
          __syncwarp();
          if (lane == 0) {
            po->use_ps = 0;
            if (frame_status == 0) fside[f].status = JAADB_ST_ARRAY_BOUNDS;
          }
          if (frame_status == 0) frame_status = JAADB_ST_ARRAY_BOUNDS;
          mode = 0;
        }
      } else if (lane == 0) po->use_ps = 0;
    }
    const bool processed = frame_status == 0 && mode != 0;
    __syncwarp();
    // ---- frame records (all lanes)
    for (int c = 0; c < nch; ++c) {
      SbrFrameDev* o = out + ((size_t)run.sbr_base + it) * 2 + c;
      SbrChanParse& cp = S->ch[c];
      if (dequant || mode == 2) {
        // NoiseEnvelope.dequantChannel / unmap (when this frame's sbr_data got there): the bands and envelopes of this
        // frame; everything else keeps what an earlier frame put there (SbrChanParse::E_orig).  The record of a frame that
        // runs the tool is a copy of the arrays as they stand now.
        for (int i = lane; i < kSbrMaxLE * 64; i += 32) {
          const int l = i >> 6, k = i & 63;
          float v = cp.E_orig[l][k];
          if (dequant && l < cp.L_E && k < S->n[cp.f[l]]) { v = sbr_e_orig(T, *S, stereo, c, k, l); cp.E_orig[l][k] = v; }
          if (mode == 2) o->E_orig[l][k] = v;
        }
        if (lane < 16) {
          const int l = lane >> 3, k = lane & 7;
          float qd = cp.Q_div[l][k], qd2 = cp.Q_div2[l][k];
          if (dequant && l < cp.L_Q && k < S->N_Q) { sbr_q_div(T, *S, stereo, c, k, l, qd, qd2); cp.Q_div[l][k] = qd; cp.Q_div2[l][k] = qd2; }
          if (mode == 2) { o->Q_div[l][k] = qd; o->Q_div2[l][k] = qd2; }
        }
      }
      if (mode == 2) {
        for (int i = lane; i < 64; i += 32) {
          o->f_table_res[0][i] = S->f_table_res[0][i];
          o->f_table_res[1][i] = S->f_table_res[1][i];
          o->f_table_lim[i] = S->f_table_lim[S->hdr.limiter_bands][i];
          o->table_map_k_to_g[i] = S->table_map_k_to_g[i];
          o->bs_add_harmonic[i] = cp.bs_add_harmonic[i];
          o->bs_add_harmonic_prev[i] = cp.bs_add_harmonic_prev[i];
        }
        if (lane < 8) {
          o->f_table_noise[lane] = S->f_table_noise[lane];
          o->patchNoSubbands[lane] = S->patchNoSubbands[lane];
          o->patchStartSubband[lane] = S->patchStartSubband[lane];
        }
        if (lane < 6) { o->t_E[lane] = cp.t_E[lane]; o->f[lane] = cp.f[lane]; }
        if (lane < 3) o->t_Q[lane] = cp.t_Q[lane];
        if (lane < 5) o->bs_invf_mode[lane] = cp.bs_invf_mode[lane];
      }
      if (lane == 0) {
        o->mode = (uint8_t)mode;
        o->frame_status = (uint8_t)frame_status;
        // SBR2.process passes reset = false for the second channel's patch construction; the patches are built above once
        o->reset = S->reset;
        o->L_E = cp.L_E; o->L_Q = cp.L_Q; o->kx = S->kx; o->M = S->M; o->N_high = S->N_high; o->N_low = S->N_low; o->N_Q = S->N_Q;
        o->N_L = S->N_L[S->hdr.limiter_bands]; o->kx_prev = S->kx_prev; o->M_prev = S->M_prev; o->noPatches = S->noPatches;
        o->limiter_gains = S->hdr.limiter_gains; o->interpol_freq = S->hdr.interpol_freq; o->smoothing_mode = S->hdr.smoothing_mode;
        o->add_harmonic_flag_prev = cp.add_harmonic_flag_prev;
        o->l_A = cp.l_A; o->prevEnvIsShort = cp.prevEnvIsShort;
        o->ord = n_proc;
        o->back = last_proc < 0 ? 0u : (uint32_t)(it - last_proc);
        o->fwd = 0;
        o->back_ps = last_ps < 0 ? 0u : (uint32_t)(it - last_ps);
        o->fwd_ps = 0;
        if (processed && last_proc >= 0) out[((size_t)run.sbr_base + last_proc) * 2 + c].fwd = (uint32_t)(it - last_proc);
        if (processed && use_ps && last_ps >= 0) out[((size_t)run.sbr_base + last_ps) * 2 + c].fwd_ps = (uint32_t)(it - last_ps);
      }
    }
    if (processed) { last_proc = it; ++n_proc; if (use_ps) last_ps = it; }
    __syncwarp();
    // ---- what SBR.process leaves behind for the next frame's parse (sbr_save_prev_data, SBR.java:256-284)
    if (mode == 2) {
      for (int c = 0; c < nch; ++c) {
        SbrChanParse& cp = S->ch[c];
        const int le = cp.L_E - 1, lq = cp.L_Q - 1;
        for (int i = lane; i < kSbrMaxM; i += 32) {
          cp.E_prev[i] = cp.E[i][le];
          cp.Q_prev[i] = cp.Q[i][lq];
          cp.bs_add_harmonic_prev[i] = cp.bs_add_harmonic[i];
        }
        __syncwarp();
        if (lane == 0) {
          S->kx_prev = S->kx;
          S->M_prev = S->M;
          cp.L_E_prev = cp.L_E;
          cp.f_prev = cp.f[cp.L_E - 1];
          cp.add_harmonic_flag_prev = cp.add_harmonic_flag;
          cp.prevEnvIsShort = (cp.l_A == cp.L_E) ? 0 : -1;
        }
      }
    }
    __syncwarp();
  }
  {
    const uint32_t* src = reinterpret_cast<const uint32_t*>(S);
    uint32_t* dst = reinterpret_cast<uint32_t*>(G);
    for (int i = lane; i < (int)(sizeof(SbrElemDev) / 4); i += 32) dst[i] = src[i];
  }
}

}  // namespace jaadb
