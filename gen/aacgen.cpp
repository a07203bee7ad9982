// Synthetic AAC bitstream generator (test + benchmark input).
//
// Emits random *valid* raw_data_blocks restricted to the subset JAAD decodes
// safely (no codebook 12/13, no pulse, no gain control, no prediction, abs(q) <= 8190),
// wrapped in ADTS headers or as bare frames, together with the ground truth the
// generator knows by construction (window flags, sections, scalefactors,
// quantised coefficients).  The ground truth is independent of every decoder,
// which is what pins the integer stage of the parity tests.
//
// Deterministic: one xoshiro256** stream per (seed, stream_id); no libc rand, no
// floating-point in any decision that changes the bitstream except log2/pow on
// exactly representable inputs guarded by integer clamps.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../jaadec_b200/csrc/generated/jaad_tables_host.h"

namespace T = ::jaad_tables;
static const int T_SF_FREQ[12] = {96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000};

namespace {

struct Rng {
  uint64_t s[4];
  static uint64_t splitmix(uint64_t& x) {
    uint64_t z = (x += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
  explicit Rng(uint64_t seed) { for (auto& v : s) v = splitmix(seed); }
  static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  int range(int lo, int hi) { return lo + (int)(next() % (uint64_t)(hi - lo + 1)); }  // inclusive
  double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
  bool chance(double p) { return unit() < p; }
};

struct BitWriter {
  std::vector<uint8_t> buf;
  uint64_t acc = 0;
  int nacc = 0;
  size_t bits = 0;
  void put(uint32_t v, int n) {
    if (n == 0) return;
    bits += n;
    acc = (acc << n) | (v & (n == 32 ? 0xFFFFFFFFu : ((1u << n) - 1u)));
    nacc += n;
    while (nacc >= 8) { buf.push_back((uint8_t)(acc >> (nacc - 8))); nacc -= 8; }
  }
  void align() { if (nacc) put(0, 8 - nacc); }
  void append(const BitWriter& o) {
    for (uint8_t b : o.buf) put(b, 8);
    if (o.nacc) put((uint32_t)(o.acc & ((1u << o.nacc) - 1u)), o.nacc);
  }
};

// ---- Huffman encode tables built from the {len, code, values} rows -------------------
struct HuffEnc {
  int dim, lav, base;
  bool uns;
  std::vector<uint32_t> code;
  std::vector<uint8_t> len;
  int key(const int* v) const {
    int k = 0;
    for (int i = 0; i < dim; ++i) k = k * base + (uns ? std::abs(v[i]) : v[i] + lav);
    return k;
  }
};
HuffEnc g_enc[12];
uint32_t g_sfCode[121];
uint8_t g_sfLen[121];
double g_cbRms[12];
bool g_init = false;

void initTables() {
  if (g_init) return;
  static const int32_t* rows[12] = {nullptr, T::HCB1, T::HCB2, T::HCB3, T::HCB4, T::HCB5, T::HCB6,
                                    T::HCB7, T::HCB8, T::HCB9, T::HCB10, T::HCB11};
  static const int nrows[12] = {0, 81, 81, 81, 81, 81, 81, 64, 64, 169, 169, 289};
  static const int lav[12] = {0, 1, 1, 2, 2, 4, 4, 7, 7, 12, 12, 16};
  static const bool uns[12] = {false, false, false, true, true, false, false, true, true, true, true, true};
  for (int cb = 1; cb <= 11; ++cb) {
    HuffEnc& e = g_enc[cb];
    e.dim = cb < 5 ? 4 : 2;
    e.lav = lav[cb];
    e.uns = uns[cb];
    e.base = e.uns ? e.lav + 1 : 2 * e.lav + 1;
    int n = 1;
    for (int i = 0; i < e.dim; ++i) n *= e.base;
    e.code.assign(n, 0);
    e.len.assign(n, 0);
    int w = e.dim + 2;
    for (int r = 0; r < nrows[cb]; ++r) {
      const int32_t* row = rows[cb] + r * w;
      int v[4];
      for (int i = 0; i < e.dim; ++i) v[i] = row[2 + i];
      int k = e.key(v);
      e.code[k] = (uint32_t)row[1];
      e.len[k] = (uint8_t)row[0];
    }
    // rms of |q|^(4/3) under the generator's own value distribution (used for scalefactor targets)
    double s = 0;
    int cnt = 0;
    for (int v = -e.lav; v <= e.lav; ++v) { s += std::pow(std::pow(std::abs((double)v), 4.0 / 3.0), 2); ++cnt; }
    g_cbRms[cb] = std::sqrt(s / cnt);
  }
  g_cbRms[11] = 60.0;
  for (int r = 0; r < 121; ++r) {
    int val = T::HCB_SF[r * 3 + 2];
    g_sfCode[val] = (uint32_t)T::HCB_SF[r * 3 + 1];
    g_sfLen[val] = (uint8_t)T::HCB_SF[r * 3];
  }
  g_init = true;
}

// ---- per-ICS plan ------------------------------------------------------------------------
struct IcsPlan {
  int ws = 0, shape = 0, maxSfb = 0, ngroups = 1;
  int glen[8] = {1, 0, 0, 0, 0, 0, 0, 0};
  int cb[120];       // per (group, sfb)
  int sf[120];       // spectrum: scalefactor; IS: intensity position (running value)
  int16_t q[1024];   // decoded layout
  int globalGain = 0;
  bool tns = false;
  BitWriter tnsBits;
  int32_t tnsTruth[600];   // present, then per window {n_filt, coef_res, 3 x {length, order, direction, compress, coef[20] (signed)}}
  const int16_t* swb = nullptr;
  int nswb = 0;
  IcsPlan() { memset(cb, 0, sizeof cb); memset(sf, 0, sizeof sf); memset(q, 0, sizeof q); memset(tnsTruth, 0, sizeof tnsTruth); }
};

struct Ctx {
  Rng rng;
  int sfIndex;
  double pPns = 0;      // > 0: perceptual noise substitution (codebook 13) on that share of the bands
  double psExt = 0;     // > 0: PS headers enable the IPD/OPD extension with this probability
  bool psIso = false;   // PS modes restricted to the ones JAAD decodes like ISO/IEC 14496-3 (aacgen_ps.inc)
  bool tnsMild = false; // TNS filters an ISO decoder can apply without blowing up: order <= 12 / 7, small reflection coefficients
  double pPulse = 0;    // > 0: probability that a long-window ICS carries pulse_data (ISO/IEC 14496-3 4.6.3.3)
  bool pulseWild = false; // pulses may also land past max_sfb (no encoder does that; decoders differ: FFmpeg uses stale band data)
  explicit Ctx(uint64_t seed, int sfi) : rng(seed), sfIndex(sfi) {}
};

int bandBits(const IcsPlan& p, int g, int sfb, int groupOff) {
  int cb = p.cb[g * p.maxSfb + sfb];
  if (cb == 0 || cb > 11) return 0;
  const HuffEnc& e = g_enc[cb];
  int bits = 0;
  int width = p.swb[sfb + 1] - p.swb[sfb];
  for (int w = 0; w < p.glen[g]; ++w) {
    int off = groupOff + w * 128 + p.swb[sfb];
    for (int k = 0; k < width; k += e.dim) {
      int v[4] = {0, 0, 0, 0};
      for (int j = 0; j < e.dim; ++j) {
        int x = p.q[off + k + j];
        v[j] = (cb == 11 && std::abs(x) > 15) ? (x < 0 ? -16 : 16) : x;
      }
      bits += e.len[e.key(v)];
      if (e.uns) for (int j = 0; j < e.dim; ++j) bits += v[j] != 0;
      if (cb == 11)
        for (int j = 0; j < 2; ++j) {
          int a = std::abs((int)p.q[off + k + j]);
          if (a >= 16) { int n = 31 - __builtin_clz(a); bits += (n - 4) + 1 + n; }
        }
    }
  }
  return bits;
}

void writeBand(BitWriter& bw, const IcsPlan& p, int g, int sfb, int groupOff) {
  int cb = p.cb[g * p.maxSfb + sfb];
  if (cb == 0 || cb > 11) return;
  const HuffEnc& e = g_enc[cb];
  int width = p.swb[sfb + 1] - p.swb[sfb];
  for (int w = 0; w < p.glen[g]; ++w) {
    int off = groupOff + w * 128 + p.swb[sfb];
    for (int k = 0; k < width; k += e.dim) {
      int v[4] = {0, 0, 0, 0};
      for (int j = 0; j < e.dim; ++j) {
        int x = p.q[off + k + j];
        v[j] = (cb == 11 && std::abs(x) > 15) ? (x < 0 ? -16 : 16) : x;
      }
      int key = e.key(v);
      bw.put(e.code[key], e.len[key]);
      if (e.uns)
        for (int j = 0; j < e.dim; ++j)
          if (v[j] != 0) bw.put(v[j] < 0 ? 1 : 0, 1);
      if (cb == 11)
        for (int j = 0; j < 2; ++j) {
          int a = std::abs((int)p.q[off + k + j]);
          if (a >= 16) {
            int n = 31 - __builtin_clz(a);  // a in [2^n, 2^(n+1)), n in 4..12
            for (int i = 4; i < n; ++i) bw.put(1, 1);
            bw.put(0, 1);
            bw.put((uint32_t)(a - (1 << n)), n);
          }
        }
    }
  }
}

// choose window/grouping and max_sfb
void planInfo(Ctx& c, IcsPlan& p, int ws, int shape, bool lfe) {
  p.ws = ws;
  p.shape = shape;
  if (ws == 2) {
    p.swb = T::SWB_OFFSET_SHORT + 17 * c.sfIndex;
    p.nswb = T::SWB_SHORT_WINDOW_COUNT[c.sfIndex];
    p.ngroups = 1;
    p.glen[0] = 1;
    for (int i = 1; i < 8; ++i) p.glen[i] = 0;
    for (int i = 0; i < 7; ++i) {
      if (c.rng.chance(0.6)) p.glen[p.ngroups - 1]++;
      else { p.ngroups++; p.glen[p.ngroups - 1] = 1; }
    }
    p.maxSfb = c.rng.range(p.nswb / 2, p.nswb);
  } else {
    p.swb = T::SWB_OFFSET_LONG + 53 * c.sfIndex;
    p.nswb = T::SWB_LONG_WINDOW_COUNT[c.sfIndex];
    p.ngroups = 1;
    p.glen[0] = 1;
    for (int i = 1; i < 8; ++i) p.glen[i] = 0;
    p.maxSfb = lfe ? c.rng.range(4, 12) : c.rng.range(p.nswb / 2, p.nswb);
  }
}

// random codebooks + coefficients; isAllowed -> may use intensity codebooks 14/15
void planSpectrum(Ctx& c, IcsPlan& p, bool isAllowed, int budgetBits) {
  memset(p.q, 0, sizeof p.q);
  int n = p.ngroups * p.maxSfb;
  int prev = c.rng.range(0, 11);
  for (int i = 0; i < n; ++i) {
    int cb = c.rng.chance(0.45) ? prev : c.rng.range(0, 11);
    if (isAllowed && c.rng.chance(0.05)) cb = c.rng.chance(0.5) ? 14 : 15;
    if (c.pPns > 0 && c.rng.chance(c.pPns)) cb = 13;   // (no draw when the knob is off: older seeds keep their streams)
    p.cb[i] = cb;
    if (cb <= 11) prev = cb;
  }
  int groupOff = 0;
  for (int g = 0; g < p.ngroups; ++g) {
    for (int sfb = 0; sfb < p.maxSfb; ++sfb) {
      int cb = p.cb[g * p.maxSfb + sfb];
      if (cb == 0 || cb > 11) continue;
      const HuffEnc& e = g_enc[cb];
      int width = p.swb[sfb + 1] - p.swb[sfb];
      for (int w = 0; w < p.glen[g]; ++w) {
        int off = groupOff + w * 128 + p.swb[sfb];
        for (int k = 0; k < width; ++k) {
          int v;
          if (cb == 11) {
            if (c.rng.chance(0.08)) {
              // abs(q) log-uniform in [16, 8190]
              double l = 4.0 + c.rng.unit() * (std::log2(8190.0) - 4.0);
              v = (int)std::floor(std::exp2(l));
              v = std::min(std::max(v, 16), 8190);
            } else v = c.rng.range(0, 15);
            if (c.rng.chance(0.5)) v = -v;
          } else {
            v = c.rng.range(-e.lav, e.lav);
          }
          p.q[off + k] = (int16_t)v;
        }
      }
    }
    groupOff += p.glen[g] * 128;
  }
  // fit the bit budget by silencing bands from the top of the spectrum downwards
  std::vector<int> bits(n);
  int total = 0;
  groupOff = 0;
  for (int g = 0; g < p.ngroups; ++g) {
    for (int sfb = 0; sfb < p.maxSfb; ++sfb) { bits[g * p.maxSfb + sfb] = bandBits(p, g, sfb, groupOff); total += bits[g * p.maxSfb + sfb]; }
    groupOff += p.glen[g] * 128;
  }
  for (int sfb = p.maxSfb - 1; sfb >= 0 && total > budgetBits; --sfb) {
    int goff = 0;
    for (int g = 0; g < p.ngroups; ++g) {
      int i = g * p.maxSfb + sfb;
      if (p.cb[i] >= 1 && p.cb[i] <= 11) {
        total -= bits[i];
        int width = p.swb[sfb + 1] - p.swb[sfb];
        for (int w = 0; w < p.glen[g]; ++w) memset(p.q + goff + w * 128 + p.swb[sfb], 0, width * sizeof(int16_t));
        p.cb[i] = 0;
      }
      goff += p.glen[g] * 128;
    }
  }
  // sometimes trim max_sfb down to the last used band (exercises different max_sfb values)
  if (c.rng.chance(0.5)) {
    int last = 0;
    for (int g = 0; g < p.ngroups; ++g)
      for (int sfb = 0; sfb < p.maxSfb; ++sfb)
        if (p.cb[g * p.maxSfb + sfb] != 0) last = std::max(last, sfb + 1);
    if (last < p.maxSfb) {
      int tmp[120];
      for (int g = 0; g < p.ngroups; ++g)
        for (int sfb = 0; sfb < last; ++sfb) tmp[g * last + sfb] = p.cb[g * p.maxSfb + sfb];
      p.maxSfb = last;
      memcpy(p.cb, tmp, sizeof(int) * p.ngroups * last);
    }
  }
}

// energy of the dequantised spectrum for unit scalefactor base; returns the per-window maximum sum of squares
double planScalefactors(Ctx& c, IcsPlan& p, double targetRms, double extraGain) {
  const float* IQ = JT(IQ_TABLE);
  int n = p.ngroups * p.maxSfb;
  // relative scalefactors: equalise band energies, add jitter, bound consecutive deltas
  int rel[120];
  int prevSf = 0, prevIs = 0, prevNoise = -1000;
  bool first = true;
  for (int i = 0; i < n; ++i) {
    int cb = p.cb[i];
    if (cb == 0) { rel[i] = 0; continue; }
    if (cb == 13) {
      // noise energy: the band's L2 norm is 2^(sf/4) (ICStream.java:252); a few hundred per coefficient
      const int lo = p.ws == 2 ? 4 : 24, hi = p.ws == 2 ? 36 : 56;
      int v = c.rng.range(lo, hi);
      if (prevNoise > -1000) v = std::min(std::max(v, prevNoise - 60), prevNoise + 60);
      rel[i] = v;
      prevNoise = v;
      continue;
    }
    if (cb >= 14) {
      int pos = prevIs + c.rng.range(-6, 6);
      pos = std::min(std::max(pos, -4), 24);
      rel[i] = pos;
      prevIs = pos;
      continue;
    }
    int r = (int)std::lround(-4.0 * std::log2(g_cbRms[cb])) + c.rng.range(-8, 8);
    if (!first) r = std::min(std::max(r, prevSf - 60), prevSf + 60);
    first = false;
    rel[i] = r;
    prevSf = r;
  }
  // energy with base 0 -> choose base
  double maxE = 0;
  int groupOff = 0;
  double winE[8] = {0};
  int wbase = 0;
  for (int g = 0; g < p.ngroups; ++g) {
    for (int sfb = 0; sfb < p.maxSfb; ++sfb) {
      int i = g * p.maxSfb + sfb;
      int cb = p.cb[i];
      if (cb == 0 || cb > 11) continue;
      double sc = std::exp2(rel[i] / 4.0);
      int width = p.swb[sfb + 1] - p.swb[sfb];
      for (int w = 0; w < p.glen[g]; ++w) {
        int off = groupOff + w * 128 + p.swb[sfb];
        double e = 0;
        for (int k = 0; k < width; ++k) { double v = IQ[std::abs((int)p.q[off + k])] * sc; e += v * v; }
        winE[p.ws == 2 ? wbase + w : 0] += e;
      }
    }
    groupOff += p.glen[g] * 128;
    wbase += p.glen[g];
  }
  for (int w = 0; w < 8; ++w) maxE = std::max(maxE, winE[w]);
  double halfN = (p.ws == 2) ? 128.0 : 1024.0;   // x_rms = sqrt(E/2) / (N/2)
  int base = 100;
  if (maxE > 0) {
    double rms0 = std::sqrt(maxE / 2.0) / halfN * extraGain;
    base = 100 + (int)std::floor(4.0 * std::log2(targetRms / rms0));
  }
  // keep every spectral scalefactor inside [0, 255]
  int lo = 1 << 30, hi = -(1 << 30);
  for (int i = 0; i < n; ++i)
    if (p.cb[i] >= 1 && p.cb[i] <= 11) { lo = std::min(lo, rel[i]); hi = std::max(hi, rel[i]); }
  if (lo <= hi) {
    if (base + hi > 255) base = 255 - hi;
    if (base + lo < 0) base = -lo;
  }
  for (int i = 0; i < n; ++i) {
    int cb = p.cb[i];
    if (cb == 0) p.sf[i] = 0;
    else if (cb >= 13) p.sf[i] = rel[i];
    else p.sf[i] = std::min(std::max(base + rel[i], 0), 255);
  }
  // global_gain: the first spectral band's scalefactor plus a small offset the first delta undoes
  int firstSf = -1;
  for (int i = 0; i < n; ++i)
    if (p.cb[i] >= 1 && p.cb[i] <= 11) { firstSf = p.sf[i]; break; }
  if (firstSf < 0) p.globalGain = c.rng.range(0, 255);
  else p.globalGain = std::min(std::max(firstSf + c.rng.range(-20, 20), 0), 255);
  return maxE;
}

void planTns(Ctx& c, IcsPlan& p) {
  p.tns = true;
  BitWriter& bw = p.tnsBits;
  bool sh = p.ws == 2;
  int nwin = sh ? 8 : 1;
  int32_t* tt = p.tnsTruth;
  tt[0] = 1;
  for (int w = 0; w < nwin; ++w) {
    int32_t* tw = tt + 1 + 74 * w;
    int nf = sh ? c.rng.range(0, 1) : c.rng.range(0, 3);
    bw.put(nf, sh ? 1 : 2);
    tw[0] = nf;
    if (!nf) continue;
    int coefRes = c.rng.range(0, 1);
    bw.put(coefRes, 1);
    tw[1] = coefRes;
    for (int f = 0; f < nf; ++f) {
      int32_t* tf = tw + 2 + 24 * f;
      tf[0] = c.rng.range(0, sh ? 15 : 63);
      bw.put(tf[0], sh ? 4 : 6);
      int order = sh ? c.rng.range(0, 7) : c.rng.range(0, c.tnsMild ? 12 : 20);
      bw.put(order, sh ? 3 : 5);
      tf[1] = order;
      if (order) {
        tf[2] = c.rng.range(0, 1);
        bw.put(tf[2], 1);
        int cc = c.rng.range(0, 1);
        bw.put(cc, 1);
        tf[3] = cc;
        int len = coefRes + 3 - cc;
        for (int i = 0; i < order; ++i) {
          // mild: reflection coefficients up to ~0.43 (index +-1 at 3-bit, +-2 at 4-bit resolution)
          const int lim = coefRes ? 2 : 1;
          uint32_t raw = c.tnsMild ? ((uint32_t)c.rng.range(-std::min(lim, 1 << (len - 1)), std::min(lim, (1 << (len - 1)) - 1)) & ((1u << len) - 1u))
                                   : (uint32_t)c.rng.range(0, (1 << len) - 1);
          bw.put(raw, len);
          tf[4 + i] = ((int32_t)(raw << (32 - len))) >> (32 - len);   // the signed index of 14496-3 4.6.9.3
        }
      }
    }
  }
}

void writeIcsInfo(BitWriter& bw, const IcsPlan& p) {
  bw.put(0, 1);
  bw.put(p.ws, 2);
  bw.put(p.shape, 1);
  if (p.ws == 2) {
    bw.put(p.maxSfb, 4);
    int bits = 0, nb = 0;
    for (int g = 0; g < p.ngroups; ++g) {
      for (int i = 1; i < p.glen[g]; ++i) { bits = (bits << 1) | 1; nb++; }
      if (g + 1 < p.ngroups) { bits = (bits << 1); nb++; }
    }
    bw.put(bits, 7);
    (void)nb;
  } else {
    bw.put(p.maxSfb, 6);
    bw.put(0, 1);  // predictor_data_present
  }
}

// pulse_data (ISO/IEC 14496-3 4.4.2.7 table 4.7, 4.6.3.3): up to four pulses; the decoder adds each amplitude to the magnitude
// of the quantised coefficient at its position.  IcsPlan::q stays what an ISO decoder must end up with (the ground truth);
// the coefficient that goes into the bitstream is the one with the pulse taken off.  Pulses that land in bands without
// spectral data (codebooks 0, 13, 14, 15, or past max_sfb) carry any amplitude: a decoder has to leave those alone.
struct PulsePlan {
  int count = 0, startSfb = 0;
  int offset[4] = {0, 0, 0, 0};   // the 5-bit increments
  int amp[4] = {0, 0, 0, 0};
  int pos[4] = {0, 0, 0, 0};
  int16_t qtx[4] = {0, 0, 0, 0};  // transmitted coefficient at pos (valid when coded)
  bool coded[4] = {false, false, false, false};
};

void planPulses(Ctx& c, const IcsPlan& p, PulsePlan& pp) {
  if (p.ws == 2 || p.maxSfb == 0 || !(c.pPulse > 0) || !c.rng.chance(c.pPulse)) return;
  pp.startSfb = c.rng.range(0, std::min(p.maxSfb, p.nswb) - 1);
  const int want = c.rng.range(1, 4);
  int pos = p.swb[pp.startSfb];
  for (int i = 0; i < want; ++i) {
    const int inc = c.rng.range(0, 31);
    if (pos + inc > 1023 || (!c.pulseWild && pos + inc >= p.swb[p.maxSfb])) break;
    pos += inc;
    int amp = c.rng.range(0, 15);
    int sfb = 0;
    while (sfb < p.maxSfb && p.swb[sfb + 1] <= pos) ++sfb;
    const int cb = sfb < p.maxSfb ? p.cb[sfb] : 0;
    bool coded = cb >= 1 && cb <= 11;
    for (int j = 0; j < pp.count; ++j) if (pp.pos[j] == pos) coded = false, amp = 0;  // (inc 0: one pulse per coefficient is enough)
    int16_t qtx = 0;
    if (coded) {
      const int qf = p.q[pos];
      amp = qf > 0 ? std::min(amp, qf - 1) : std::min(amp, -qf);
      qtx = (int16_t)(qf > 0 ? qf - amp : qf + amp);
    }
    pp.offset[i] = inc; pp.amp[i] = amp; pp.pos[i] = pos; pp.qtx[i] = qtx; pp.coded[i] = coded;
    pp.count = i + 1;
  }
}

void writeIcs(Ctx& c, BitWriter& bw, const IcsPlan& p, bool commonWindow) {
  bw.put(p.globalGain, 8);
  if (!commonWindow) writeIcsInfo(bw, p);
  // section_data: maximal runs, randomly split now and then
  int sectBits = p.ws == 2 ? 3 : 5, esc = (1 << sectBits) - 1;
  for (int g = 0; g < p.ngroups; ++g) {
    int k = 0;
    while (k < p.maxSfb) {
      int cb = p.cb[g * p.maxSfb + k];
      int end = k + 1;
      while (end < p.maxSfb && p.cb[g * p.maxSfb + end] == cb) ++end;
      if (end - k > 1 && c.rng.chance(0.15)) end = k + c.rng.range(1, end - k);
      int len = end - k;
      bw.put(cb, 4);
      while (len >= esc) { bw.put(esc, sectBits); len -= esc; }
      bw.put(len, sectBits);
      k = end;
    }
  }
  // scale_factor_data
  int cur = p.globalGain, curIs = 0, curNoise = p.globalGain - 90;
  bool noiseFirst = true;
  for (int i = 0; i < p.ngroups * p.maxSfb; ++i) {
    int cb = p.cb[i];
    if (cb == 0) continue;
    if (cb == 13) {
      // ICStream.java:199-208: the first noise energy is 9 bits (offset 256), the others are coded differentially
      if (noiseFirst) { bw.put((uint32_t)(p.sf[i] - curNoise + 256), 9); noiseFirst = false; }
      else { int d = p.sf[i] - curNoise; bw.put(g_sfCode[d + 60], g_sfLen[d + 60]); }
      curNoise = p.sf[i];
      continue;
    }
    if (cb >= 14) { int d = p.sf[i] - curIs; bw.put(g_sfCode[d + 60], g_sfLen[d + 60]); curIs = p.sf[i]; }
    else { int d = p.sf[i] - cur; bw.put(g_sfCode[d + 60], g_sfLen[d + 60]); cur = p.sf[i]; }
  }
  PulsePlan pp;
  planPulses(c, p, pp);
  bw.put(pp.count ? 1 : 0, 1);  // pulse_data_present
  if (pp.count) {
    bw.put(pp.count - 1, 2);
    bw.put(pp.startSfb, 6);
    for (int i = 0; i < pp.count; ++i) { bw.put(pp.offset[i], 5); bw.put(pp.amp[i], 4); }
  }
  bw.put(p.tns ? 1 : 0, 1);
  if (p.tns) bw.append(p.tnsBits);
  bw.put(0, 1);  // gain_control_data_present
  IcsPlan& mp = const_cast<IcsPlan&>(p);
  int16_t keep[4];
  for (int i = 0; i < pp.count; ++i) if (pp.coded[i]) { keep[i] = mp.q[pp.pos[i]]; mp.q[pp.pos[i]] = pp.qtx[i]; }
  int groupOff = 0;
  for (int g = 0; g < p.ngroups; ++g) {
    for (int sfb = 0; sfb < p.maxSfb; ++sfb) writeBand(bw, p, g, sfb, groupOff);
    groupOff += p.glen[g] * 128;
  }
  for (int i = 0; i < pp.count; ++i) if (pp.coded[i]) mp.q[pp.pos[i]] = keep[i];
}

#include "aacgen_sbr.inc"
#include "aacgen_ps.inc"

}  // namespace

extern "C" {

struct jg_config {
  int32_t sf_index;       // 3 = 48 kHz, 4 = 44.1 kHz ...
  int32_t chan_cfg;       // 1 mono (SCE), 2 stereo (CPE), 6 = 5.1 (SCE,CPE,CPE,LFE)
  int32_t n_frames;
  int32_t target_bytes;   // per frame, payload
  int32_t long_only;      // 1: ONLY_LONG windows only (config 1)
  int32_t adts;           // 1: prefix every frame with a 7-byte ADTS header
  float p_transient;      // probability of starting a LONG_START -> EIGHT_SHORT.. -> LONG_STOP run
  float p_common_window;
  float p_tns;
  float p_is;             // >0 enables intensity codebooks on the right channel of common-window CPEs
  int32_t ms_mode;        // 0: never, 1: random ms_mask_present in {0,1,2}
  int32_t sbr_mode;       // 0 none (LC); 1 SBR; 2 SBR+PS  (filled in by aacgen_sbr.inc when present)
  float target_rms;       // PCM rms target (default 2500)
  int32_t sbr_quirk;      // 1: also emit coupled SBR frames only the reference parses (aacgen_sbr.inc, SbrChanState)
  int32_t sbr_downsampled; // 1: SBR band tables for the CORE rate (what JAAD uses when the stream is opened from an ASC that
                          //    does not signal SBR: outputFrequency stays at the core rate, A/DecoderConfig.java:180, A/sbr/SBR.java:100-102)
  float p_pns;            // > 0: share of the bands coded as perceptual noise (codebook 13)
  int32_t tns_mild;       // 1: TNS filters an ISO decoder can apply (orders <= 12 / 7, small coefficients)
  float ps_ext;           // > 0: probability that a PS header enables the IPD/OPD extension (ps/Extension.java)
  int32_t ps_iso;         // 1: only the PS modes on which JAAD and ISO/IEC 14496-3 agree (cross-checks against other decoders)
  float p_pulse;          // > 0: probability that a long-window ICS carries pulse_data; the truth's q is what an ISO decoder
                          //      reconstructs (JAAD parses the pulses and never applies them, A/syntax/ICStream.java:17)
  int32_t pulse_wild;     // 1: pulses may also land past max_sfb
  float p_drc;            // > 0: probability that a frame ends with a dynamic_range_info fill element (+ sometimes padding)
};

// Ground truth per ICS (element order, L before R); arrays may be NULL.
struct jg_truth {
  int16_t* q;        // [n_frames][n_ics][1024]
  int16_t* sfidx;    // [n_frames][n_ics][120]  SCALEFACTOR_TABLE index, -1 for zero bands
  uint8_t* sfbcb;    // [n_frames][n_ics][120]
  int32_t* info;     // [n_frames][n_ics][16]: present, ws, shape, (unused), maxSfb, ngroups, glen[8], msMask, common
  uint8_t* msused;   // [n_frames][n_elements][128]
  int32_t* sbr;      // [n_frames][n_ics][480] SBR streams only: L_E, L_Q, frame class, pointer, t_E[6], f[6], amp_res,
                     //   coupling, (pad to 32), E[5][64], Q[2][64] as a decoder reconstructs them (aacgen_sbr.inc)
  int32_t* ps;       // [n_frames][jg_ps_truth_ints()] SBR+PS streams only: num_env, border_position[6], pad, iid[5][34], icc[5][34],
                     //   ipd[5][17], nr_ipdopd_par, enable_ipdopd (aacgen_ps.inc)
  int32_t* tns;      // [n_frames][n_ics][600]: tns_data_present, then per window {n_filt, coef_res, 3 x {length, order,
                     //   direction, coef_compress, coef[20] as signed indices}} (IcsPlan::tnsTruth)
};

int jg_sbr_truth_ints(void) { return kSbrTruthInts; }
int jg_ps_truth_ints(void) { return kPsTruthInts; }

int jg_ics_per_frame(int chan_cfg) { return chan_cfg == 6 ? 6 : chan_cfg; }
int jg_elements_per_frame(int chan_cfg) { return chan_cfg == 6 ? 4 : 1; }

// Generates one stream.  Returns bytes written to out (or -1 if cap is too small).
// frame_offsets/frame_sizes describe the payload (after the ADTS header when adts=1).
int64_t jg_generate(const jg_config* cfg, uint64_t seed, uint8_t* out, int64_t cap, int64_t* frame_offsets,
                    int32_t* frame_sizes, const jg_truth* truth) {
  initTables();
  Ctx c(seed, cfg->sf_index);
  c.pPns = cfg->p_pns;
  c.tnsMild = cfg->tns_mild != 0;
  c.psExt = cfg->ps_ext;
  c.psIso = cfg->ps_iso != 0;
  c.pPulse = cfg->p_pulse;
  c.pulseWild = cfg->pulse_wild != 0;
  Rng drcRng(seed ^ 0xD2C0D2C0D2C0D2C0ull);   // its own generator: the audio elements are the ones of the stream without p_drc
  const int nIcs = jg_ics_per_frame(cfg->chan_cfg);
  const int nEl = jg_elements_per_frame(cfg->chan_cfg);
  // element layout
  struct El { int type, tag, nch; bool lfe; };
  std::vector<El> els;
  if (cfg->chan_cfg == 1) els = {{0, 0, 1, false}};
  else if (cfg->chan_cfg == 2) els = {{1, 0, 2, false}};
  else if (cfg->chan_cfg == 6) els = {{0, 0, 1, false}, {1, 0, 2, false}, {1, 1, 2, false}, {3, 0, 1, true}};
  else return -2;
  std::vector<int> wsState(els.size(), 0);  // 0 long, 1 start sent -> shorts, 2 in shorts
  std::vector<SbrElemState> sbrState(els.size());
  const int sbrSrIndex = cfg->sbr_downsampled ? cfg->sf_index : cfg->sf_index - 3;
  const int sbrSrFreq = sbrSrIndex >= 0 ? T_SF_FREQ[sbrSrIndex] : 0;
  if (cfg->sbr_mode && (cfg->sf_index < 3 || (cfg->sbr_mode > 1 && cfg->chan_cfg != 1))) return -3;
  PsState psState;
  const double targetRms = cfg->target_rms > 0 ? cfg->target_rms : 2500.0;
  int64_t pos = 0;
  for (int f = 0; f < cfg->n_frames; ++f) {
    BitWriter bw;
    int icsIdx = 0;
    int payloadBudget = cfg->target_bytes * 8 - 3 - 8;
    if (cfg->sbr_mode) payloadBudget -= (cfg->chan_cfg == 2 ? 60 : (cfg->sbr_mode > 1 ? 70 : 36)) * 8;   // room for the SBR fill element
    for (size_t ei = 0; ei < els.size(); ++ei) {
      const El& el = els[ei];
      // window sequence state machine (per element): ONLY_LONG -> LONG_START -> EIGHT_SHORT+ -> LONG_STOP -> ONLY_LONG
      int ws;
      int& st = wsState[ei];
      if (cfg->long_only || el.lfe) ws = 0;
      else if (st == 0) { if (c.rng.chance(cfg->p_transient)) { ws = 1; st = 1; } else ws = 0; }
      else if (st == 1) { ws = 2; st = 2; }
      else { if (c.rng.chance(0.5)) ws = 2; else { ws = 3; st = 0; } }
      double share = el.lfe ? 0.03 : (double)el.nch / (nIcs - (cfg->chan_cfg == 6 ? 0.7 : 0));
      int elBudget = (int)(payloadBudget * share) - 60 * el.nch;
      bw.put(el.type, 3);
      bw.put(el.tag, 4);
      if (el.nch == 1) {
        IcsPlan p;
        planInfo(c, p, ws, c.rng.range(0, 1), el.lfe);
        planSpectrum(c, p, false, std::max(elBudget, 40));
        planScalefactors(c, p, targetRms, 1.0);
        if (c.rng.chance(cfg->p_tns)) planTns(c, p);
        writeIcs(c, bw, p, false);
        if (truth) {
          size_t o = (size_t)f * nIcs + icsIdx;
          if (truth->q) memcpy(truth->q + o * 1024, p.q, 2048);
          if (truth->sfbcb) { memset(truth->sfbcb + o * 120, 0, 120); for (int i = 0; i < p.ngroups * p.maxSfb; ++i) truth->sfbcb[o * 120 + i] = (uint8_t)p.cb[i]; }
          if (truth->sfidx) {
            for (int i = 0; i < 120; ++i) truth->sfidx[o * 120 + i] = -1;
            for (int i = 0; i < p.ngroups * p.maxSfb; ++i)
              truth->sfidx[o * 120 + i] = p.cb[i] == 0 ? -1 : p.cb[i] == 13 ? (int16_t)((std::min(std::max(p.sf[i], -100), 155) + 200) | 0x4000)
                                                                              : (int16_t)(p.sf[i] + 100);
          }
          if (truth->tns) memcpy(truth->tns + o * 600, p.tnsTruth, sizeof p.tnsTruth);
          if (truth->info) {
            int32_t* in = truth->info + o * 16;
            in[0] = 1; in[1] = p.ws; in[2] = p.shape; in[3] = 0; in[4] = p.maxSfb; in[5] = p.ngroups;
            for (int i = 0; i < 8; ++i) in[6 + i] = i < p.ngroups ? p.glen[i] : 0;
            in[14] = 0; in[15] = 0;
          }
        }
        if (cfg->sbr_mode && !el.lfe) {
          BitWriter psBits;
          if (cfg->sbr_mode > 1) psBuild(c, psBits, psState, f, (truth && truth->ps) ? truth->ps + (size_t)f * kPsTruthInts : nullptr);
          sbrEmitFill(c, bw, sbrState[ei], false, f, sbrSrIndex, sbrSrFreq,
                      (truth && truth->sbr) ? truth->sbr + ((size_t)f * nIcs + icsIdx) * kSbrTruthInts : nullptr,
                      cfg->sbr_mode > 1 ? &psBits : nullptr, cfg->sbr_quirk != 0);
        }
        icsIdx += 1;
      } else {
        bool common = c.rng.chance(cfg->p_common_window);
        IcsPlan L, R;
        int shapeL = c.rng.range(0, 1);
        planInfo(c, L, ws, shapeL, false);
        if (common) {
          R.ws = L.ws; R.shape = L.shape; R.maxSfb = L.maxSfb; R.ngroups = L.ngroups;
          memcpy(R.glen, L.glen, sizeof L.glen);
          R.swb = L.swb; R.nswb = L.nswb;
        } else {
          planInfo(c, R, ws, c.rng.range(0, 1), false);
        }
        int msMask = 0;
        uint8_t ms[128];
        memset(ms, 0, sizeof ms);
        planSpectrum(c, L, false, std::max(elBudget / 2, 40));
        if (common) {
          // L may have trimmed max_sfb; R must follow the shared ics_info
          R.maxSfb = L.maxSfb;
        }
        {
          int keep = R.maxSfb;
          planSpectrum(c, R, common && cfg->p_is > 0, std::max(elBudget / 2, 40));
          if (common && R.maxSfb != keep) {
            // undo a max_sfb trim on R: re-expand codebooks with zeros
            int tmp[120];
            memset(tmp, 0, sizeof tmp);
            for (int g = 0; g < R.ngroups; ++g)
              for (int s = 0; s < R.maxSfb; ++s) tmp[g * keep + s] = R.cb[g * R.maxSfb + s];
            R.maxSfb = keep;
            memcpy(R.cb, tmp, sizeof tmp);
          }
        }
        if (common && cfg->ms_mode) {
          double u = c.rng.unit();
          msMask = u < 0.4 ? 0 : (u < 0.8 ? 1 : 2);
          if (msMask == 1) for (int i = 0; i < L.ngroups * L.maxSfb; ++i) ms[i] = (uint8_t)c.rng.range(0, 1);
          if (msMask == 2) memset(ms, 1, sizeof ms);
        }
        double g = (msMask ? 1.5 : 1.0);
        planScalefactors(c, L, targetRms, g);
        planScalefactors(c, R, targetRms, g);
        if (c.rng.chance(cfg->p_tns)) planTns(c, L);
        if (c.rng.chance(cfg->p_tns)) planTns(c, R);
        bw.put(common ? 1 : 0, 1);
        if (common) {
          writeIcsInfo(bw, L);
          bw.put(msMask, 2);
          if (msMask == 1) for (int i = 0; i < L.ngroups * L.maxSfb; ++i) bw.put(ms[i], 1);
        }
        writeIcs(c, bw, L, common);
        writeIcs(c, bw, R, common);
        if (truth) {
          const IcsPlan* ps[2] = {&L, &R};
          for (int ch = 0; ch < 2; ++ch) {
            const IcsPlan& p = *ps[ch];
            size_t o = (size_t)f * nIcs + icsIdx + ch;
            if (truth->q) memcpy(truth->q + o * 1024, p.q, 2048);
            if (truth->sfbcb) { memset(truth->sfbcb + o * 120, 0, 120); for (int i = 0; i < p.ngroups * p.maxSfb; ++i) truth->sfbcb[o * 120 + i] = (uint8_t)p.cb[i]; }
            if (truth->sfidx) {
              for (int i = 0; i < 120; ++i) truth->sfidx[o * 120 + i] = -1;
              for (int i = 0; i < p.ngroups * p.maxSfb; ++i) {
                int cb = p.cb[i];
                int16_t v = -1;
                if (cb >= 14) v = (int16_t)(200 - std::min(std::max(p.sf[i], -155), 100));
                else if (cb == 13) v = (int16_t)((std::min(std::max(p.sf[i], -100), 155) + 200) | 0x4000);
                else if (cb != 0) v = (int16_t)(p.sf[i] + 100);
                truth->sfidx[o * 120 + i] = v;
              }
            }
            if (truth->tns) memcpy(truth->tns + o * 600, p.tnsTruth, sizeof p.tnsTruth);
            if (truth->info) {
              int32_t* in = truth->info + o * 16;
              in[0] = 1; in[1] = p.ws; in[2] = p.shape; in[3] = 0; in[4] = p.maxSfb; in[5] = p.ngroups;
              for (int i = 0; i < 8; ++i) in[6 + i] = i < p.ngroups ? p.glen[i] : 0;
              in[14] = msMask; in[15] = common ? 1 : 0;
            }
          }
          if (truth->msused) memcpy(truth->msused + ((size_t)f * nEl + ei) * 128, ms, 128);
        }
        if (cfg->sbr_mode)
          sbrEmitFill(c, bw, sbrState[ei], true, f, sbrSrIndex, sbrSrFreq,
                      (truth && truth->sbr) ? truth->sbr + ((size_t)f * nIcs + icsIdx) * kSbrTruthInts : nullptr, nullptr,
                      cfg->sbr_quirk != 0);
        icsIdx += 2;
      }
    }
    // fill elements real encoders add after the audio elements: dynamic_range_info (extension type 11; JAAD parses it into
    // an object nobody reads, syntax/DRC.java) and plain padding (types 0 / 1)
    if (cfg->p_drc > 0 && drcRng.chance(cfg->p_drc)) {
      BitWriter d;
      d.put(11, 4);
      const bool pce = drcRng.chance(0.3), excl = drcRng.chance(0.3), bands = drcRng.chance(0.4), ref = drcRng.chance(0.5);
      int nb = 1;
      d.put(pce, 1);
      if (pce) { d.put(drcRng.range(0, 15), 4); d.put(0, 4); }
      d.put(excl, 1);
      if (excl) { for (int i = 0; i < 7; ++i) d.put(drcRng.range(0, 1), 1); d.put(0, 1); }   // (JAAD cannot take a second group)
      d.put(bands, 1);
      if (bands) {
        const int inc = drcRng.range(0, 6);
        d.put(inc, 4); d.put(drcRng.range(0, 15), 4);
        nb += inc;
        for (int i = 0; i < nb; ++i) d.put(drcRng.range(0, 255), 8);
      }
      d.put(ref, 1);
      if (ref) { d.put(drcRng.range(0, 127), 7); d.put(0, 1); }
      for (int i = 0; i < nb; ++i) { d.put(drcRng.range(0, 1), 1); d.put(drcRng.range(0, 127), 7); }
      d.align();
      const int cnt = (int)d.buf.size();
      bw.put(6, 3);
      if (cnt >= 15) { bw.put(15, 4); bw.put(cnt - 14, 8); } else bw.put(cnt, 4);
      for (uint8_t b : d.buf) bw.put(b, 8);
      if (drcRng.chance(0.5)) {
        const int pad = drcRng.range(0, 20);                      // fill_element: count bytes, type 0 (FILL) or 1 (FILL_DATA)
        bw.put(6, 3);
        if (pad >= 15) { bw.put(15, 4); bw.put(pad - 14, 8); } else bw.put(pad, 4);
        if (pad > 0) { bw.put(drcRng.range(0, 1), 4); bw.put(0, 4); for (int i = 1; i < pad; ++i) bw.put(0xA5, 8); }
      }
    }
    bw.put(7, 3);  // END
    bw.align();
    int payload = (int)bw.buf.size();
    int hdr = cfg->adts ? 7 : 0;
    if (pos + hdr + payload > cap) return -1;
    if (cfg->adts) {
      int flen = payload + 7;
      uint8_t* h = out + pos;
      h[0] = 0xFF;
      h[1] = 0xF1;                                                  // MPEG-4, layer 0, protection_absent
      h[2] = (uint8_t)((1 << 6) | (cfg->sf_index << 2) | ((cfg->chan_cfg >> 2) & 1));  // profile field 1 = LC
      h[3] = (uint8_t)(((cfg->chan_cfg & 3) << 6) | ((flen >> 11) & 3));
      h[4] = (uint8_t)((flen >> 3) & 0xFF);
      h[5] = (uint8_t)(((flen & 7) << 5) | 0x10);                   // buffer fullness 0x400
      h[6] = 0x00;
    }
    memcpy(out + pos + hdr, bw.buf.data(), payload);
    if (frame_offsets) frame_offsets[f] = pos + hdr;
    if (frame_sizes) frame_sizes[f] = payload;
    pos += hdr + payload;
  }
  return pos;
}

// Many streams in parallel: stream s uses seed = base_seed + s and writes at out + s*stride.
// frame tables are [n_streams][n_frames]; offsets are absolute within `out`.
int64_t jg_generate_many(const jg_config* cfg, uint64_t base_seed, int n_streams, uint8_t* out, int64_t stride,
                         int64_t* frame_offsets, int32_t* frame_sizes, int64_t* stream_bytes, int n_threads) {
  initTables();
  std::vector<std::thread> th;
  std::vector<int64_t> rc(n_streams, 0);
  auto work = [&](int t) {
    for (int s = t; s < n_streams; s += n_threads) {
      int64_t* fo = frame_offsets + (size_t)s * cfg->n_frames;
      int64_t n = jg_generate(cfg, base_seed + (uint64_t)s, out + (size_t)s * stride, stride, fo,
                              frame_sizes + (size_t)s * cfg->n_frames, nullptr);
      rc[s] = n;
      if (n >= 0) for (int f = 0; f < cfg->n_frames; ++f) fo[f] += (int64_t)s * stride;
      if (stream_bytes) stream_bytes[s] = n;
    }
  };
  for (int t = 0; t < n_threads; ++t) th.emplace_back(work, t);
  for (auto& t : th) t.join();
  for (int s = 0; s < n_streams; ++s) if (rc[s] < 0) return rc[s];
  return 0;
}

}  // extern "C"
