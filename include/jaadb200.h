/* jaadb200.h -- C ABI of the B200 batched AAC decode engine.
 *
 * Drop-in boundary for JAAD's per-frame decode path.  Each entry point names
 * the reference interface it replaces (paths relative to /root/reference):
 *   A/ = aac/src/main/java/net/sourceforge/jaad/aac/
 *   S/ = src/main/java/net/sourceforge/jaad/
 *
 * Plain pointers and sizes only; no exceptions cross this boundary.  Engine
 * level calls return 0 on success or a negative JAADB_E_* code; per-frame
 * decode problems are reported in jaadb_frame_result.status (JAADB_ST_*),
 * mirroring JAAD's AACException messages, and never abort the batch.
 *
 * Threading: one caller thread per engine (JAAD's Decoder is not thread-safe
 * either, A/Decoder.java).  One engine drives one GPU; streams shard across
 * engines/processes by stream id with no collective.
 */
#ifndef JAADB200_H
#define JAADB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JAADB_ABI_VERSION 1

/* engine-level error codes */
#define JAADB_OK 0
#define JAADB_E_INVALID (-1)     /* bad argument */
#define JAADB_E_CUDA (-2)        /* CUDA runtime failure (see jaadb_last_error) */
#define JAADB_E_NOMEM (-3)
#define JAADB_E_CONFIG (-4)      /* unsupported / malformed stream configuration */
#define JAADB_E_NOSTREAM (-5)    /* unknown stream id */
#define JAADB_E_CAPACITY (-6)    /* max_streams exceeded / output buffer too small */

/* per-frame status words (mirror A/AACException.java call sites) */
#define JAADB_ST_OK 0
#define JAADB_ST_EOS 1                /* EOSException, swallowed by Decoder.decodeFrame (A/Decoder.java:96-98) */
#define JAADB_ST_INVALID_CODEBOOK 2   /* "invalid huffman codebook: 12"      A/syntax/ICStream.java:129 */
#define JAADB_ST_TOO_MANY_BANDS 3     /* "too many bands"                    A/syntax/ICStream.java:138 */
#define JAADB_ST_SF_RANGE 4           /* "scalefactor out of range"          A/syntax/ICStream.java:213 */
#define JAADB_ST_PULSE_SHORT 5        /* "pulse data not allowed for short"  A/syntax/ICStream.java:79 */
#define JAADB_ST_PULSE_RANGE 6        /* pulse SWB / offset out of range     A/syntax/ICStream.java:152,166 */
#define JAADB_ST_MS_RESERVED 7        /* "reserved MS mask type used"        A/syntax/CPE.java:114 */
#define JAADB_ST_TNS_ORDER 8          /* "TNS filter out of range"           A/tools/TNS.java:47 */
#define JAADB_ST_LTP_PROFILE 9        /* "unexpected profile for LTP"        A/syntax/ICSInfo.java:139 */
#define JAADB_ST_UNSUPPORTED_ELEMENT 10 /* CCE / PCE / SSR gain control / Main+LTP prediction / DRC / ADIF */
#define JAADB_ST_LAYOUT 11            /* the frame addresses element objects the stream does not own: an element the channel
                                         layout does not have, or another element_instance_tag than the stream used before
                                         (JAAD keeps one object per type and tag, A/syntax/Element.java:36-38, and would decode
                                         such a frame against fresh ones); the stream's state is left as JAAD leaves it */
#define JAADB_ST_PROFILE 12           /* "unsupported profile"               A/Decoder.java:110 */
#define JAADB_ST_ARRAY_BOUNDS 13      /* a Java ArrayIndexOutOfBoundsException (IQ index > 8190, sf index < 0, and inside the
                                       * SBR / PS tools: get_S_mapped with an odd N_high and a low-resolution envelope,
                                       * A/sbr/HFAdjustment.java:64-75; PS indices past the tables, A/ps/PSImpl.java:424-478) */
#define JAADB_ST_SBR 14               /* AACException raised inside the SBR tool */
#define JAADB_ST_CONFIG 15

/* PCM formats.  S16 interleaved is bit-identical to S/SampleBuffer.java:168-209
 * (big-endian is SampleBuffer's default order, little-endian the AudioFormat one);
 * F32 planar is the float[] list handed to A/Receiver.java:14. */
#define JAADB_PCM_S16LE 0
#define JAADB_PCM_S16BE 1
#define JAADB_PCM_F32_PLANAR 2

/* TNS modes.  JAAD parses TNS data and never applies it (A/tools/TNS.java:63-68);
 * JAADB_TNS_JAAD reproduces that and is the parity mode. */
#define JAADB_TNS_JAAD 0
/* JAADB_TNS_ISO applies the all-pole filter of ISO/IEC 14496-3 4.6.9.3 (tns_decode_frame / tns_decode_coef /
 * tns_ar_filter) between the stereo tools and the filterbank, as A/syntax/CPE.java:177-178 and A/syntax/SCE.java:104 place
 * the call JAAD never implemented.  Checked against the oracle's ISO restatement (bit-identical) and a float64 direct form. */
#define JAADB_TNS_ISO 1

#define JAADB_FLAG_PROFILE 1u   /* record per-kernel CUDA-event timings (jaadb_batch_timings) */
#define JAADB_FLAG_DEBUG_TAPS 2u /* keep the dequantised spectra for jaadb_batch_tap (parity tests) */
/* JAAD parses pulse_data and never applies it ("TODO: apply pulse data", A/syntax/ICStream.java:17,148-170); that is the
 * default and the parity mode.  With JAADB_FLAG_PULSE_ISO the pulses are added to the quantised coefficients ahead of the
 * inverse quantisation (ISO/IEC 14496-3 4.6.3.3); bands without spectral data (codebooks 0, 13, 14, 15, past max_sfb) are
 * left alone, a magnitude past JAAD's IQ_TABLE (8190) fails the frame with JAADB_ST_ARRAY_BOUNDS.  Checked bit for bit against
 * the oracle's pulseMode 1, which agrees with FFmpeg's decoder to > 90 dB. */
#define JAADB_FLAG_PULSE_ISO 4u

typedef struct jaadb_engine jaadb_engine;
typedef struct jaadb_batch jaadb_batch;

typedef struct jaadb_options {
  int32_t device;        /* CUDA device ordinal */
  uint32_t max_streams;  /* capacity of the stream table */
  int32_t pcm_format;    /* JAADB_PCM_* */
  int32_t tns_mode;      /* JAADB_TNS_* */
  uint32_t flags;        /* JAADB_FLAG_* */
  uint32_t chunk_frames; /* jaadb_decode pipelines chunks of this many consecutive frames (0: default, 131072) */
  uint32_t sbr_tile_frames; /* SBR stages work on tiles of this many frames per stream (0: sized to the workspace budget) */
  uint32_t k2_segment_frames; /* filterbank kernel: cut every stream's frames into segments of this many frames, one CTA
                                 each (0: decided from the number of streams; tests use it to force the segmented path) */
} jaadb_options;

/* One AAC frame (an ADTS payload or an MP4 sample) inside the caller's blob.
 * Replaces the BitStream argument of Decoder.decodeFrame (A/Decoder.java:89). */
typedef struct jaadb_frame_desc {
  uint64_t offset;     /* byte offset of the raw_data_block in the blob */
  uint32_t nbytes;
  int32_t stream_id;
} jaadb_frame_desc;

/* What SampleBuffer reports after accept() (S/SampleBuffer.java:72-110). */
typedef struct jaadb_frame_result {
  int32_t status;          /* JAADB_ST_* ; non-zero => no PCM for this frame, stream continues */
  uint16_t channels;
  uint16_t sample_length;  /* per channel: 1024, or 2048 with SBR up-sampling (A/DecoderConfig.java:83-86) */
  uint32_t sample_rate;
  uint32_t pcm_bytes;      /* bytes written at this frame's pcm offset */
} jaadb_frame_result;

typedef struct jaadb_stream_info {
  int32_t profile;         /* audio object type of the core coder */
  int32_t sf_index;        /* core sampling-frequency index */
  int32_t channel_config;
  int32_t channels;        /* A/DecoderConfig.java:108-115 (mono reports 2) */
  int32_t sample_rate;     /* output rate */
  int32_t sample_length;   /* per frame and channel */
  int32_t sbr;             /* 1 if the stream is expected to carry SBR (decided at open) */
  int32_t reserved;
} jaadb_stream_info;

typedef struct jaadb_timings {
  float parse_ms;       /* K1: noiseless decode */
  float filterbank_ms;  /* K2: dequant + stereo + IMDCT + overlap-add + PCM pack */
  float sbr_ms;         /* K3..: SBR / PS stages (0 for AAC-LC) */
  float total_ms;       /* first launch to last launch of the decode */
  uint32_t launches;    /* kernels launched by the last jaadb_batch_decode */
  uint32_t reserved[3];
} jaadb_timings;

int jaadb_abi_version(void);
const char* jaadb_status_string(int32_t status);
const char* jaadb_last_error(const jaadb_engine* e);

/* ---- engine -------------------------------------------------------------- */
int jaadb_engine_create(const jaadb_options* opts, jaadb_engine** out);
void jaadb_engine_destroy(jaadb_engine* e);

/* ---- streams -------------------------------------------------------------- */
/* Decoder.create(byte[] audioSpecificConfig)   A/Decoder.java:36-43, A/DecoderConfig.java:175-254 */
int jaadb_stream_open_asc(jaadb_engine* e, const uint8_t* asc, uint32_t asc_bytes, int32_t* stream_id);
/* The same for a stream whose ASC does not signal SBR but whose frames carry it (`expect_sbr` as below, e.g. from
 * jaadb_probe_sbr): JAAD creates the SBR tool when the first payload arrives; the ASC has fixed the output rate by then
 * (A/DecoderConfig.java:180), so the rate is NOT doubled and the stream runs the down-sampled SBR tool -- 32-band synthesis
 * (A/sbr/SynthesisFilterbank32.java), 1024 samples per frame at the core rate (A/sbr/SBR.java:35-37,100, SURVEY A-20).
 * An ASC with explicit signalling keeps what it says; expect_sbr = 2 then only adds parametric stereo. */
int jaadb_stream_open_asc_sbr(jaadb_engine* e, const uint8_t* asc, uint32_t asc_bytes, int32_t expect_sbr, int32_t* stream_id);
/* Decoder.create(AudioDecoderInfo) from an ADTS header   A/Decoder.java:45-48, S/adts/ADTSFrame.java:119-129
 * `expect_sbr`: JAAD switches a stream to 2048-sample output when the first SBR
 * payload arrives (A/sbr/SBR.java:98-101); the batched engine needs that decision at
 * open time (0 = plain AAC-LC, 1 = SBR, 2 = SBR+PS). */
int jaadb_stream_open_adts(jaadb_engine* e, int32_t profile, int32_t sf_index, int32_t channel_config,
                           int32_t expect_sbr, int32_t* stream_id);
/* What `expect_sbr` should be for a stream whose header does not say: JAAD learns that an LC-signalled stream is HE-AAC
 * when the first SBR payload arrives (implicit signalling) and switches the decoder on the spot (A/sbr/SBR.java:98-101,
 * A/syntax/ChannelElement.java:65-76); the batched engine needs the decision when the stream is opened.  Parses `frame`
 * (the stream's first raw_data_block) on the GPU with a scratch stream and reports what it carries:
 * 0 = no SBR payload, 1 = SBR, 2 = SBR + parametric stereo.  Needs one free slot of the stream table while it runs. */
int jaadb_probe_sbr(jaadb_engine* e, int32_t profile, int32_t sf_index, int32_t channel_config, const uint8_t* frame,
                    uint32_t nbytes, int32_t* expect_sbr);
/* The same for a stream described by an AudioSpecificConfig (MP4): the SBR band tables of a stream JAAD opens from an ASC
 * follow the ASC's output rate (A/sbr/SBR.java:102), so the probe has to parse with those. */
int jaadb_probe_sbr_asc(jaadb_engine* e, const uint8_t* asc, uint32_t asc_bytes, const uint8_t* frame, uint32_t nbytes,
                        int32_t* expect_sbr);
int jaadb_stream_close(jaadb_engine* e, int32_t stream_id);
int jaadb_stream_get_info(const jaadb_engine* e, int32_t stream_id, jaadb_stream_info* info);

/* ---- one-call decode ------------------------------------------------------
 * Batched Decoder.decodeFrame (A/Decoder.java:89-121) + SampleBuffer.accept
 * (S/SampleBuffer.java:168-209).  Frames of one stream are applied in array
 * order; different streams are independent.  pcm_offsets[i] is the byte offset
 * of frame i's PCM inside pcm_out (NULL: frames are packed back to back in
 * array order, every frame taking its stream's full frame size).
 * `blob` and `pcm_out` may each be host memory (pinned or pageable) or memory of
 * the engine's GPU (cudaMalloc / managed; detected with cudaPointerGetAttributes).
 * A device pcm_out (4-byte aligned) is written by the kernels in place: the PCM
 * never crosses PCIe, which is the mode for GPU-side consumers (resamplers,
 * feature extraction, encoders).  `frames`, `pcm_offsets` and `results` are host
 * arrays.  A frame that fails (results[i].status != 0) has pcm_bytes 0 and its
 * slot zero-filled.                                                          */
int jaadb_decode(jaadb_engine* e, const uint8_t* blob, uint64_t blob_bytes, const jaadb_frame_desc* frames,
                 uint32_t n_frames, void* pcm_out, uint64_t pcm_capacity, const uint64_t* pcm_offsets,
                 jaadb_frame_result* results);

/* ---- staged decode: the same work with the phases exposed ----------------
 * create (index + upload descriptors) -> upload (H2D blob) -> decode (kernels
 * only, everything resident in HBM) -> download (D2H PCM + results).         */
int jaadb_batch_create(jaadb_engine* e, const jaadb_frame_desc* frames, uint32_t n_frames, uint64_t blob_bytes,
                       const uint64_t* pcm_offsets, jaadb_batch** out);
uint64_t jaadb_batch_pcm_bytes(const jaadb_batch* b);
int jaadb_batch_upload(jaadb_batch* b, const uint8_t* blob, uint64_t blob_bytes);
int jaadb_batch_decode(jaadb_batch* b);          /* asynchronous on the engine's stream */
int jaadb_batch_sync(jaadb_batch* b);
int jaadb_batch_download(jaadb_batch* b, void* pcm_out, uint64_t pcm_capacity, jaadb_frame_result* results);
int jaadb_batch_timings(jaadb_batch* b, jaadb_timings* t);   /* needs JAADB_FLAG_PROFILE */
void jaadb_batch_destroy(jaadb_batch* b);

/* Parity taps (JAADB_FLAG_DEBUG_TAPS): integer and float intermediates of frame
 * `frame`, channel slot `ch` of the last decode.  Any output pointer may be NULL.
 *   q[1024]      quantised coefficients, de-interleaved as A/syntax/ICStream.java:258-271
 *   sfidx[120]   SCALEFACTOR_TABLE index per (group, sfb); -1 where the scalefactor is 0.0f
 *   sfbcb[120]   codebook per (group, sfb)
 *   spec[1024]   dequantised spectrum after M/S and intensity stereo (input of the filterbank)
 *   info[16]     present, window_sequence, window_shape, info_decoded, max_sfb, groups, group_len[8], ms_mask, common_window */
int jaadb_batch_tap(jaadb_batch* b, uint32_t frame, uint32_t ch, int16_t* q, int16_t* sfidx, uint8_t* sfbcb,
                    float* spec, int32_t* info, uint8_t* ms_used128);

/* SBR parity tap: the per-frame SBR record (dequantised envelopes, band tables, grid) the parse kernel produced for
 * frame `frame`, channel `ch` of the last decode, as raw bytes of the engine's internal layout (jaadec_b200/csrc/
 * sbr_types.cuh, SbrFrameDev); returns its size, 0 if the frame's stream carries no SBR, or a negative JAADB_E_* code. */
int jaadb_batch_tap_sbr(jaadb_batch* b, uint32_t frame, uint32_t ch, void* out, uint32_t out_bytes);
/* PS parity tap: the parametric-stereo parameters of frame `frame` after PSImpl.ps_data_decode (A/ps/PSImpl.java:129-199) as
 * the mixing stage uses them (PsFrameDev, 320 bytes: use_ps, num_env, border_position[6], iid / icc mode, iid[5][20],
 * icc[5][20], Extension.nr_par, ExtData.enabled, ipd[5][17]); returns its size, 0 if the frame's stream carries no parametric stereo, or a negative JAADB_E_* code. */
int jaadb_batch_tap_ps(jaadb_batch* b, uint32_t frame, void* out, uint32_t out_bytes);

/* ---- container indexers (host side; frames for jaadb_decode) --------------
 * The caller keeps whole ADTS streams / MP4 files in one blob; these calls produce the frame table.  A NULL `frames`
 * (or a too small max_frames) only counts.  Return value: number of frames found, or a negative JAADB_E_* code.   */
typedef struct jaadb_adts_info {   /* fields of the first frame's header, S/adts/ADTSFrame.java:68-100 */
  int32_t profile;         /* audio object type (2-bit field + 1) */
  int32_t sf_index;
  int32_t channel_config;
  int32_t sample_rate;
  uint64_t n_frames;
} jaadb_adts_info;

typedef struct jaadb_mp4_track {   /* the first AAC audio track of a movie, M/api/AudioTrack.java, M/api/Track.java */
  uint8_t asc[64];         /* DecoderSpecificInfo = AudioSpecificConfig for jaadb_stream_open_asc (M/api/Track.java:155-172) */
  uint32_t asc_bytes;
  int32_t track_id;
  uint32_t timescale;      /* mdhd */
  uint32_t channel_count;  /* AudioSampleEntry */
  uint32_t sample_size_bits;
  uint32_t sample_rate;
  uint32_t object_type;    /* DecoderConfigDescriptor objectTypeIndication (0x40 = MPEG-4 audio) */
  uint32_t max_bitrate, avg_bitrate;
  uint32_t reserved;
  uint64_t duration;       /* mdhd, in timescale units */
  uint64_t n_frames;
} jaadb_mp4_track;

/* ADTSDemultiplexer: sync search + header + payload span of every frame   S/adts/ADTSDemultiplexer.java:26-74.
 * frames[i].offset = blob_offset + position of the raw_data_block inside `data`. */
int64_t jaadb_adts_index(const uint8_t* data, uint64_t nbytes, uint64_t blob_offset, int32_t stream_id,
                         jaadb_frame_desc* frames, uint64_t max_frames, jaadb_adts_info* info);
/* Stream s occupies blob[stream_begin[s], stream_begin[s+1]); indexed on `threads` host threads (0: all cores).
 * frames are stream-major, first_frame[n_streams+1] receives each stream's first row; stream_ids NULL: 0..n-1. */
int64_t jaadb_adts_index_many(const uint8_t* blob, const uint64_t* stream_begin, uint32_t n_streams,
                              const int32_t* stream_ids, jaadb_frame_desc* frames, uint64_t max_frames,
                              uint64_t* first_frame, jaadb_adts_info* infos, uint32_t threads);
/* MP4Container + Movie.getTracks(AAC).get(0) + Track.parseSampleTable: stsz/stco|co64/stsc/stts -> frames in
 * decoding-time order, esds -> AudioSpecificConfig      M/api/Track.java:90-172, M/boxes/BoxFactory.java:319-363 */
int64_t jaadb_mp4_index(const uint8_t* file, uint64_t nbytes, uint64_t blob_offset, int32_t stream_id,
                        jaadb_frame_desc* frames, uint64_t max_frames, jaadb_mp4_track* track);
int64_t jaadb_mp4_index_many(const uint8_t* blob, const uint64_t* file_begin, uint32_t n_files, const int32_t* stream_ids,
                             jaadb_frame_desc* frames, uint64_t max_frames, uint64_t* first_frame,
                             jaadb_mp4_track* tracks, uint32_t threads);

/* ---- containers in, PCM out ------------------------------------------------
 * S/Main.java:49-111 for a whole batch in one call: n_streams ADTS streams (S/adts/ADTSDemultiplexer.java) or MP4 files
 * (M/MP4Container.java, M/api/Track.java) stored back to back in `blob` (host memory; container s = blob[stream_begin[s],
 * stream_begin[s+1]), stream_begin[0] = 0) are indexed on `threads` host threads (0: all cores) WHILE the bytes are already on
 * their way to the GPU, the frames are put in frame-major order (jaadb_frames_interleave) and decoded like jaadb_decode
 * does, PCM packed in that order.  stream_ids[s] (NULL: 0..n-1) names the open stream container s belongs to.
 * Returns the number of frames decoded, or a negative JAADB_E_* code.  `results` / `frames_out` (either may be NULL) receive
 * one row per frame and must hold max_frames rows (JAADB_E_CAPACITY if the containers hold more); pcm_out may be host or
 * device memory as for jaadb_decode.                                          */
#define JAADB_CONTAINER_ADTS 0
#define JAADB_CONTAINER_MP4 1
int64_t jaadb_decode_containers(jaadb_engine* e, int32_t kind, const uint8_t* blob, const uint64_t* stream_begin, uint32_t n_streams,
                                const int32_t* stream_ids, void* pcm_out, uint64_t pcm_capacity, jaadb_frame_result* results,
                                uint64_t max_frames, jaadb_frame_desc* frames_out, uint32_t threads);

/* Reorders the stream-major table of the *_index_many calls frame-major: frame 0 of every stream, frame 1 of every stream,
 * ... -- the order a live batch of concurrent streams arrives in, and the one that gives every chunk of jaadb_decode all
 * streams to work on.  Per-stream order is kept, so both orders decode identically.  `out` has first_frame[n_streams] rows.
 * Returns the number of rows, or a negative JAADB_E_* code.  (S/Main.java reads one file at a time; this is the batch glue.) */
int64_t jaadb_frames_interleave(const jaadb_frame_desc* frames, const uint64_t* first_frame, uint32_t n_streams,
                                jaadb_frame_desc* out, uint32_t threads);

#ifdef __cplusplus
}
#endif
#endif /* JAADB200_H */
