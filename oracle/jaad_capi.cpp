// ORACLE -- TEST INFRASTRUCTURE ONLY (see jaad_bits.hpp for the full notice).
//
// Plain-C entry points over the restatement, loaded with ctypes by tests/,
// __graft_entry__.smoke() and bench.py's CPU-baseline legs.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "jaad_adts.hpp"
#include "jaad_lc.hpp"
#ifdef JAAD_ORACLE_WITH_SBR
#include "jaad_sbr.hpp"
#include "jaad_ps.hpp"
#endif

using namespace jaad;

namespace {
struct Handle {
  std::unique_ptr<Decoder> dec;
  FrameOutput last;
};
}  // namespace

extern "C" {

int jo_has_sbr() {
#ifdef JAAD_ORACLE_WITH_SBR
  return 1;
#else
  return 0;
#endif
}

// Decoder.create(AudioDecoderInfo)  (Decoder.java:45-48)
void* jo_open_adts(int profile, int sf_index, int chan_cfg, int* status) {
  try {
    DecoderConfig c = DecoderConfig::fromInfo(profile, sf_index, chan_cfg);
    auto* h = new Handle();
    h->dec.reset(new Decoder(c));
    if (status) *status = 0;
    return h;
  } catch (const AACException& e) {
    if (status) *status = e.code;
    return nullptr;
  }
}

// Decoder.create(byte[] audioSpecificConfig)  (Decoder.java:36-43)
void* jo_open_asc(const uint8_t* asc, int n, int* status) {
  try {
    DecoderConfig c;
    BitStream in(asc, (size_t)n);
    c.decode(in);
    auto* h = new Handle();
    h->dec.reset(new Decoder(c));
    if (status) *status = 0;
    return h;
  } catch (const AACException& e) {
    if (status) *status = e.code;
    return nullptr;
  }
}

void jo_close(void* hv) { delete static_cast<Handle*>(hv); }

// 0 = JAAD (TNS parsed, never applied), 1 = ISO/IEC 14496-3 4.6.9 filter (DecoderConfig::tnsMode)
void jo_set_tns_mode(void* hv, int mode) { static_cast<Handle*>(hv)->dec->config.tnsMode = mode; }
// 0 = JAAD (pulse_data parsed, never applied), 1 = ISO/IEC 14496-3 4.6.3.3 (DecoderConfig::pulseMode)
void jo_set_pulse_mode(void* hv, int mode) { static_cast<Handle*>(hv)->dec->config.pulseMode = mode; }

// meta[0..3] = status, channels, sampleLength, sampleRate.
// pcm_f32: planar [channels][sampleLength] (may be NULL); pcm_s16: interleaved (may be NULL).
int jo_decode_frame(void* hv, const uint8_t* data, int n, float* pcm_f32, int16_t* pcm_s16, int big_endian, int* meta) {
  Handle* h = static_cast<Handle*>(hv);
  h->last = h->dec->decodeFrame(data, (size_t)n);
  if (h->last.status != ST_OK && getenv("JO_DBG_EXC")) fprintf(stderr, "oracle: %s\n", h->dec->lastError.c_str());
  const FrameOutput& f = h->last;
  if (meta) { meta[0] = f.status; meta[1] = f.channels; meta[2] = f.sampleLength; meta[3] = f.sampleRate; }
  if (f.status != ST_OK) return f.status;
  if (pcm_f32) {
    for (int c = 0; c < f.channels; ++c)
      for (int is = 0; is < f.sampleLength; ++is) {
        int k = (int)((long long)f.planes[c].second * is / f.sampleLength);
        pcm_f32[(size_t)c * f.sampleLength + is] = f.planes[c].first[k];
      }
  }
  if (pcm_s16) sampleBufferAccept(f, pcm_s16, big_endian != 0);
  return 0;
}

// Parity taps of the frame just decoded: element `el` in bitstream order, channel ch (0 = L / SCE, 1 = R).
// info[0..15] = present, windowSequence, shapeCur, shapePrev, maxSFB, groups, glen[8], msMask, commonWindow
int jo_tap_ics(void* hv, int el, int ch, int16_t* q1024, int16_t* sfidx120, uint8_t* sfbcb120, float* spec1024, int* info) {
  Handle* h = static_cast<Handle*>(hv);
  auto& ae = h->dec->syn.audioElements;
  if (el < 0 || el >= (int)ae.size()) return -1;
  ChannelElement* e = ae[el];
  ICStream* ics;
  int msMask = 0, common = 0;
  if (e->type == EL_CPE) {
    CPE* c = static_cast<CPE*>(e);
    ics = ch ? &c->icsR : &c->icsL;
    msMask = c->msMask;
    common = c->commonWindow;
  } else {
    if (ch) return -1;
    ics = &static_cast<SCE*>(e)->ics;
  }
  if (q1024) memcpy(q1024, ics->q, sizeof(ics->q));
  if (sfidx120) memcpy(sfidx120, ics->sfIndex, sizeof(ics->sfIndex));
  if (sfbcb120) for (int i = 0; i < 120; ++i) sfbcb120[i] = (uint8_t)ics->sfbCB[i];
  if (spec1024) memcpy(spec1024, ics->iqData, sizeof(ics->iqData));
  if (info) {
    info[0] = 1;
    info[1] = ics->info.windowSequence;
    info[2] = ics->info.windowShape[1];
    info[3] = ics->info.windowShape[0];
    info[4] = ics->info.maxSFB;
    info[5] = ics->info.windowGroupCount;
    for (int i = 0; i < 8; ++i) info[6 + i] = i < ics->info.windowGroupCount ? ics->info.windowGroupLength[i] : 0;
    info[14] = msMask;
    info[15] = common;
  }
  return e->type;
}

// 1 when the parse of the frame just decoded reached a fill element with an SBR payload (extension type 13 / 14)
int jo_saw_sbr_payload(void* hv) { return static_cast<Handle*>(hv)->dec->syn.sbrPayloadSeen ? 1 : 0; }

int jo_tap_msused(void* hv, int el, uint8_t* ms128) {
  Handle* h = static_cast<Handle*>(hv);
  auto& ae = h->dec->syn.audioElements;
  if (el < 0 || el >= (int)ae.size() || ae[el]->type != EL_CPE) return -1;
  CPE* c = static_cast<CPE*>(ae[el]);
  for (int i = 0; i < 128; ++i) ms128[i] = c->msUsed[i];
  return 0;
}

// SBR parity tap of the frame just decoded: element `el`, channel ch.  Same layout as the generator's truth
// (gen/aacgen_sbr.inc): L_E, L_Q, frame class, pointer, t_E[6], f[6], amp_res, coupling, pad to 32, E[5][64], Q[2][64];
// then (from out[480]) kx, M, N_high, N_low, N_Q, k0, N_master, noPatches, reset, and E_orig bits [5][64], Q_div bits [2][64].
int jo_tap_sbr(void* hv, int el, int ch, int32_t* out) {
#ifdef JAAD_ORACLE_WITH_SBR
  Handle* h = static_cast<Handle*>(hv);
  auto& ae = h->dec->syn.audioElements;
  if (el < 0 || el >= (int)ae.size() || !ae[el]->sbr) return -1;
  sbr::SBR* s = static_cast<sbr::SBR*>(ae[el]->sbr.get());
  sbr::Channel* c;
  int coupling = 0;
  if (ae[el]->type == EL_CPE) {
    sbr::SBR2* s2 = static_cast<sbr::SBR2*>(s);
    c = ch ? &s2->ch1 : &s2->ch0;
    coupling = s2->bs_coupling ? 1 : 0;
  } else {
    if (ch) return -1;
    c = &static_cast<sbr::SBR1*>(s)->ch0;
  }
  memset(out, 0, sizeof(int32_t) * 480);
  out[0] = c->L_E; out[1] = c->L_Q; out[2] = c->bs_frame_class; out[3] = c->bs_pointer;
  for (int i = 0; i < 6; ++i) { out[4 + i] = c->t_E[i]; out[10 + i] = c->f[i]; }
  out[16] = c->amp_res ? 1 : 0;
  out[17] = coupling;
  for (int l = 0; l < 5; ++l) for (int k = 0; k < 64; ++k) out[32 + l * 64 + k] = c->E[k][l];
  for (int l = 0; l < 2; ++l) for (int k = 0; k < 64; ++k) out[32 + 320 + l * 64 + k] = c->Q[k][l];
  int32_t* x = out + 480;
  x[0] = s->kx; x[1] = s->M; x[2] = s->N_high; x[3] = s->N_low; x[4] = s->N_Q; x[5] = s->k0; x[6] = s->N_master;
  x[7] = s->noPatches; x[8] = s->reset ? 1 : 0; x[9] = s->isValid() ? 1 : 0;
  if (s->hdr) { x[10] = s->hdr->bs_limiter_bands; x[11] = s->hdr->bs_limiter_gains; x[12] = s->hdr->bs_interpol_freq; x[13] = s->hdr->bs_smoothing_mode;
                x[14] = s->hdr->bs_freq_scale; x[15] = s->hdr->bs_noise_bands; }
  for (int l = 0; l < 5; ++l) for (int k = 0; k < 64; ++k) memcpy(&x[16 + l * 64 + k], &c->E_orig[k][l], 4);
  for (int l = 0; l < 2; ++l) for (int k = 0; k < 64; ++k) memcpy(&x[16 + 320 + l * 64 + k], &c->Q_div[k][l], 4);
  return 0;
#else
  (void)hv; (void)el; (void)ch; (void)out;
  return -1;
#endif
}

// PS parity tap of the frame just decoded (SCE element `el`): num_env, border_position[6], pad, iid[5][34], icc[5][34],
// then (from out[348]) iid mode, icc mode (-1 = off).  Returns -1 if the element carries no PS.
int jo_tap_ps(void* hv, int el, int32_t* out) {
#ifdef JAAD_ORACLE_WITH_SBR
  Handle* h = static_cast<Handle*>(hv);
  auto& ae = h->dec->syn.audioElements;
  if (el < 0 || el >= (int)ae.size() || !ae[el]->sbr || ae[el]->type == EL_CPE) return -1;
  sbr::SBR1* s = static_cast<sbr::SBR1*>(ae[el]->sbr.get());
  if (!s->ps) return -1;
  ps::PSImpl* p = static_cast<ps::PSImpl*>(s->ps.get());
  memset(out, 0, sizeof(int32_t) * 440);
  // [350..434]: ipd.index[5][17]; [435]: Extension.nr_par(); [436]: ExtData.enabled
  for (int env = 0; env < 5; ++env)
    for (int i = 0; i < 17; ++i) out[350 + env * 17 + i] = p->ext.ipd.index[env][i];
  out[435] = p->ext.nr_par();
  out[436] = p->ext.data_enabled ? 1 : 0;
  out[0] = p->num_env;
  for (int i = 0; i < 6; ++i) out[1 + i] = p->border_position[i];
  for (int env = 0; env < 5; ++env)
    for (int i = 0; i < 34; ++i) { out[8 + env * 34 + i] = p->iid.index[env][i]; out[8 + 170 + env * 34 + i] = p->icc.index[env][i]; }
  out[348] = p->iid.mode;
  out[349] = p->icc.mode;
  return 0;
#else
  (void)hv; (void)el; (void)out;
  return -1;
#endif
}

// QMF bank on its own (for the float64 direct-form check in tests/): n_frames x 1024 input samples -> 32-band analysis
// (X_out [n_frames*32][32][2]) -> the low 32 bands through the 64-band synthesis (pcm_out n_frames x 2048).
int jo_qmf_roundtrip(const float* in, int n_frames, float* X_out, float* pcm_out) {
#ifdef JAAD_ORACLE_WITH_SBR
  sbr::AnalysisFilterbank qa;
  sbr::SynthesisFilterbank64 qs;
  std::vector<float> xs((size_t)40 * 64 * 2, 0.f), X((size_t)32 * 64 * 2, 0.f);
  sbr::Cpx (*Xs)[64] = reinterpret_cast<sbr::Cpx(*)[64]>(xs.data());
  sbr::Cpx (*Xo)[64] = reinterpret_cast<sbr::Cpx(*)[64]>(X.data());
  for (int f = 0; f < n_frames; ++f) {
    qa.sbr_qmf_analysis_32(32, in + (size_t)f * 1024, Xs, 0, 32);
    for (int l = 0; l < 32; ++l)
      for (int k = 0; k < 64; ++k) {
        Xo[l][k][0] = k < 32 ? Xs[l][k][0] : 0.f;
        Xo[l][k][1] = k < 32 ? Xs[l][k][1] : 0.f;
        if (k < 32) { X_out[(((size_t)f * 32 + l) * 32 + k) * 2] = Xs[l][k][0]; X_out[(((size_t)f * 32 + l) * 32 + k) * 2 + 1] = Xs[l][k][1]; }
      }
    qs.synthesis(32, Xo, pcm_out + (size_t)f * 2048);
  }
  return 0;
#else
  (void)in; (void)n_frames; (void)X_out; (void)pcm_out;
  return -1;
#endif
}

// The same through the 32-band (down-sampled) synthesis bank: pcm_out n_frames x 1024.
int jo_qmf_roundtrip32(const float* in, int n_frames, float* pcm_out) {
#ifdef JAAD_ORACLE_WITH_SBR
  sbr::AnalysisFilterbank qa;
  sbr::SynthesisFilterbank32 qs;
  std::vector<float> xs((size_t)40 * 64 * 2, 0.f);
  sbr::Cpx (*Xs)[64] = reinterpret_cast<sbr::Cpx(*)[64]>(xs.data());
  for (int f = 0; f < n_frames; ++f) {
    qa.sbr_qmf_analysis_32(32, in + (size_t)f * 1024, Xs, 0, 32);
    qs.synthesis(32, Xs, pcm_out + (size_t)f * 1024);
  }
  return 0;
#else
  (void)in; (void)n_frames; (void)pcm_out;
  return -1;
#endif
}

// ADTS index: payload offsets/sizes of up to `max` frames; hdr[0..2] = profile, sf_index, chan_cfg of the first frame.
int jo_adts_index(const uint8_t* data, size_t n, int64_t* offsets, int32_t* sizes, int max, int* hdr) {
  ADTSDemultiplexer dm(data, n);
  ADTSFrameInfo f;
  int cnt = 0;
  while (cnt < max && dm.next(f)) {
    if (cnt == 0 && hdr) { hdr[0] = f.profile; hdr[1] = f.sfIndex; hdr[2] = f.channelConfiguration; }
    offsets[cnt] = (int64_t)f.payloadOffset;
    sizes[cnt] = f.payloadBytes;
    ++cnt;
  }
  return cnt;
}

// CPU baseline: decode `n_streams` independent streams, one Decoder per stream, `n_threads` workers taking
// whole streams round-robin (the "one thread per stream" model of S/Main.java:82-111).  Stream s owns frames
// [first[s], first[s+1]) of (offsets, sizes) into blob.  kind: 0 = ADTS-style open (hdr = profile, sf, chan),
// 1 = ASC open.  PCM is packed to int16 (SampleBuffer) into a per-thread scratch buffer; when pcm_out != NULL
// stream s writes its frames consecutively at pcm_out + pcm_first[s] (bytes).  Returns wall seconds; fills
// decoded_samples (per-channel output samples summed over streams) and n_errors.
double jo_decode_streams(int n_streams, const uint8_t* blob, const int64_t* first, const int64_t* offsets,
                         const int32_t* sizes, int kind, const int* hdr, const uint8_t* asc, int asc_len,
                         int n_threads, uint8_t* pcm_out, const int64_t* pcm_first, int big_endian,
                         int64_t* decoded_samples, int64_t* n_errors) {
  std::atomic<int> nextStream(0);
  std::atomic<long long> samples(0), errors(0);
  auto worker = [&]() {
    std::vector<int16_t> scratch(8 * 2048);
    for (;;) {
      int s = nextStream.fetch_add(1);
      if (s >= n_streams) break;
      std::unique_ptr<Decoder> dec;
      try {
        DecoderConfig c;
        if (kind == 0) c = DecoderConfig::fromInfo(hdr[0], hdr[1], hdr[2]);
        else { BitStream in(asc, (size_t)asc_len); c.decode(in); }
        dec.reset(new Decoder(c));
      } catch (const AACException&) { errors++; continue; }
      uint8_t* dst = pcm_out ? pcm_out + pcm_first[s] : nullptr;
      long long local = 0;
      for (int64_t f = first[s]; f < first[s + 1]; ++f) {
        FrameOutput o = dec->decodeFrame(blob + offsets[f], (size_t)sizes[f]);
        if (o.status != ST_OK) { errors++; continue; }
        int16_t* p = dst ? reinterpret_cast<int16_t*>(dst) : scratch.data();
        sampleBufferAccept(o, p, big_endian != 0);
        if (dst) dst += (size_t)o.channels * o.sampleLength * 2;
        local += o.sampleLength;
      }
      samples += local;
    }
  };
  auto t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int i = 0; i < n_threads; ++i) th.emplace_back(worker);
  for (auto& t : th) t.join();
  auto t1 = std::chrono::steady_clock::now();
  if (decoded_samples) *decoded_samples = samples.load();
  if (n_errors) *n_errors = errors.load();
  return std::chrono::duration<double>(t1 - t0).count();
}

}  // extern "C"
