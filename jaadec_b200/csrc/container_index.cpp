// Container indexers of the batched decode engine: turn ADTS byte streams and MP4 files into the jaadb_frame_desc
// tables jaadb_decode consumes.  Host code: the sync search and the sample tables are a serial chain per stream
// (every frame's position depends on the previous frame's length), so the parallelism is across streams, on host
// threads, while the GPU decodes the previous batch.
//
// Behaviour follows JAAD (paths relative to /root/reference):
//   S/ = src/main/java/net/sourceforge/jaad/      M/ = mp4/src/main/java/net/sourceforge/jaad/mp4/
//   ADTS: S/adts/ADTSDemultiplexer.java:26-74 (sync search over at most 6144 bytes, unread of the byte after 0xFF),
//         S/adts/ADTSFrame.java:44-100 (header fields, CRC skips, payload length)
//   MP4:  M/boxes/BoxFactory.java:319-363 (box header, 64-bit size, uuid), M/api/Track.java:90-152 (sample table ->
//         frames, stable sort by timestamp), M/api/Track.java:155-172 + M/od/*.java (esds -> DecoderSpecificInfo),
//         M/boxes/impl/sampleentries/AudioSampleEntry.java:18-31
#include <algorithm>
#include <atomic>
#include <cstring>
#include <memory>
#include <new>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/jaadb200.h"

namespace {

const int kSampleRates[16] = {96000, 88200, 64000, 48000, 44100, 32000, 24000, 22050, 16000, 12000, 11025, 8000, 0, 0, 0, 0};

// ---------------------------------------------------------------- ADTS
struct AdtsWalker {
  const uint8_t* d;
  uint64_t n, pos = 0;
  // One ADTSDemultiplexer.readNextFrame(): false at end of input, when no sync word turns up within 6144 bytes, or
  // when the frame's payload runs past the end (EOFException in JAAD).
  bool next(uint64_t& payload_off, uint32_t& payload_bytes, jaadb_adts_info* hdr) {
    bool found = false;
    int left = 6144;
    while (!found && left > 0) {
      if (pos >= n) return false;
      int b = d[pos++];
      --left;
      if (b == 0xFF) {
        if (pos >= n) return false;
        if ((d[pos] & 0xF6) == 0xF0) found = true;   // the byte is pushed back and read again as header byte 1
      }
    }
    if (!found || pos + 6 > n) return false;
    const uint8_t* h = d + pos;
    pos += 6;
    bool protection_absent = (h[0] & 1) != 0;
    int frame_length = ((h[2] & 3) << 11) | ((((h[3] << 8) | h[4]) & 0xFFE0) >> 5);
    int blocks = h[5] & 3;
    if (!protection_absent) pos += 2;
    if (blocks != 0 && !protection_absent) pos += 2 * blocks + 2 + 2 * blocks;
    int payload = frame_length - (protection_absent ? 7 : 9);
    if (payload < 0 || pos + (uint64_t)payload > n) return false;
    if (hdr) {
      hdr->profile = ((h[1] & 0xC0) >> 6) + 1;
      hdr->sf_index = (h[1] & 0x3C) >> 2;
      hdr->channel_config = ((h[1] & 1) << 2) | ((h[2] & 0xC0) >> 6);
      hdr->sample_rate = kSampleRates[hdr->sf_index];
    }
    payload_off = pos;
    payload_bytes = (uint32_t)payload;
    pos += payload;
    return true;
  }
};

int64_t adts_index(const uint8_t* data, uint64_t nbytes, uint64_t blob_offset, int32_t stream_id, jaadb_frame_desc* frames,
                   uint64_t max_frames, jaadb_adts_info* info) {
  AdtsWalker w{data, nbytes};
  uint64_t off;
  uint32_t len;
  int64_t cnt = 0;
  jaadb_adts_info first;
  std::memset(&first, 0, sizeof first);
  while (w.next(off, len, cnt == 0 ? &first : nullptr)) {
    if (frames && (uint64_t)cnt < max_frames) {
      frames[cnt].offset = blob_offset + off;
      frames[cnt].nbytes = len;
      frames[cnt].stream_id = stream_id;
    }
    ++cnt;
  }
  if (info) {
    *info = first;
    info->n_frames = (uint64_t)cnt;
  }
  return cnt;
}

// ---------------------------------------------------------------- MP4
struct Reader {
  const uint8_t* d;
  uint64_t n;
  bool ok = true;
  uint64_t be(uint64_t p, int bytes) {
    if (p + bytes > n || p + bytes < p) { ok = false; return 0; }
    uint64_t v = 0;
    for (int i = 0; i < bytes; ++i) v = (v << 8) | d[p + i];
    return v;
  }
};

struct Box {
  uint32_t type;
  uint64_t start, body, end;   // header start, first body byte, one past the end
};

constexpr uint32_t fourcc(const char (&s)[5]) {
  return ((uint32_t)(uint8_t)s[0] << 24) | ((uint32_t)(uint8_t)s[1] << 16) | ((uint32_t)(uint8_t)s[2] << 8) | (uint8_t)s[3];
}

// BoxFactory.parseBox: 32-bit size + type, size 1 -> 64-bit size, 'uuid' -> 16 more bytes, a child may not exceed its parent.
bool read_box(Reader& r, uint64_t at, uint64_t limit, Box& b) {
  if (at + 8 > limit) return false;
  uint64_t size = r.be(at, 4);
  b.type = (uint32_t)r.be(at + 4, 4);
  b.start = at;
  uint64_t p = at + 8;
  if (size == 1) {
    if (p + 8 > limit) return false;
    size = r.be(p, 8);
    p += 8;
  } else if (size == 0) {
    size = limit - at;   // box extends to the end of its container
  }
  if (b.type == fourcc("uuid")) p += 16;
  if (!r.ok || size < p - at || size > limit - at) return false;
  b.body = p;
  b.end = at + size;
  return true;
}

bool find_child(Reader& r, uint64_t from, uint64_t to, uint32_t type, Box& out) {
  Box b;
  for (uint64_t p = from; read_box(r, p, to, b); p = b.end)
    if (b.type == type) { out = b; return true; }
  return false;
}

// Descriptor.createDescriptor (M/od/Descriptor.java:27-61): tag, 7-bit continued size.
struct Desc {
  int tag;
  uint64_t body, end;
};
bool read_desc(Reader& r, uint64_t at, uint64_t limit, Desc& d) {
  if (at + 2 > limit) return false;
  d.tag = (int)r.be(at, 1);
  uint64_t p = at + 1, size = 0;
  int b;
  do {
    if (p >= limit) return false;
    b = (int)r.be(p++, 1);
    size = (size << 7) | (uint64_t)(b & 0x7F);
  } while (b & 0x80);
  d.body = p;
  d.end = std::min(limit, p + size);
  return r.ok;
}

// esds -> ES_Descriptor(3) -> DecoderConfigDescriptor(4) -> DecoderSpecificInfo(5)   (Track.findDecoderSpecificInfo)
bool find_asc(Reader& r, const Box& esds, jaadb_mp4_track* t) {
  uint64_t p = esds.body + 4;   // FullBox version + flags
  Desc es;
  if (!read_desc(r, p, esds.end, es) || es.tag != 3) return false;
  uint64_t q = es.body + 2;
  int flags = (int)r.be(q, 1);
  q += 1;
  if (flags & 0x80) q += 2;                               // dependsOn_ES_ID
  if (flags & 0x40) q += 1 + r.be(q, 1);                  // URL
  // (JAAD ignores the OCR flag, M/od/ESDescriptor.java:24-40)
  Desc c;
  bool got = false;
  for (; read_desc(r, q, es.end, c); q = c.end) {
    if (c.tag != 4) continue;
    t->object_type = (uint32_t)r.be(c.body, 1);
    t->max_bitrate = (uint32_t)r.be(c.body + 5, 4);
    t->avg_bitrate = (uint32_t)r.be(c.body + 9, 4);
    Desc s;
    for (uint64_t u = c.body + 13; read_desc(r, u, c.end, s); u = s.end) {
      if (s.tag != 5) continue;
      uint64_t len = s.end - s.body;
      if (len > sizeof t->asc) return false;
      std::memcpy(t->asc, r.d + s.body, len);
      t->asc_bytes = (uint32_t)len;
      got = true;                                         // the last one wins, as in JAAD's loop
    }
  }
  return got && r.ok;
}

struct Sample {
  uint64_t offset, time;
  uint32_t size;
};

// Track.parseSampleTable
int parse_stbl(Reader& r, const Box& stbl, std::vector<Sample>& out) {
  Box stsz, stco, stsc, stts;
  if (!find_child(r, stbl.body, stbl.end, fourcc("stsz"), stsz)) return JAADB_E_CONFIG;
  bool large = false;
  if (!find_child(r, stbl.body, stbl.end, fourcc("stco"), stco)) {
    if (!find_child(r, stbl.body, stbl.end, fourcc("co64"), stco)) return JAADB_E_CONFIG;
    large = true;
  }
  if (!find_child(r, stbl.body, stbl.end, fourcc("stsc"), stsc)) return JAADB_E_CONFIG;
  if (!find_child(r, stbl.body, stbl.end, fourcc("stts"), stts)) return JAADB_E_CONFIG;

  uint64_t fixed = r.be(stsz.body + 4, 4), n_samples = r.be(stsz.body + 8, 4);
  if (!r.ok || (fixed == 0 && stsz.body + 12 + 4 * n_samples > stsz.end)) return JAADB_E_CONFIG;
  // A sample table cannot describe more samples than the file has bytes for: with a fixed sample size the count is not
  // backed by a table, so a crafted header must not size the allocations below (JAAD would die of OutOfMemoryError)
  if (n_samples > r.n / std::max<uint64_t>(fixed, 1)) return JAADB_E_CONFIG;
  auto sample_size = [&](uint64_t i) -> uint32_t { return fixed ? (uint32_t)fixed : (uint32_t)r.be(stsz.body + 12 + 4 * i, 4); };

  uint64_t n_chunks = r.be(stco.body + 4, 4);
  if (!r.ok || stco.body + 8 + (large ? 8 : 4) * n_chunks > stco.end) return JAADB_E_CONFIG;
  auto chunk_offset = [&](uint64_t j) -> uint64_t { return large ? r.be(stco.body + 8 + 8 * j, 8) : r.be(stco.body + 8 + 4 * j, 4); };

  uint64_t n_runs = r.be(stsc.body + 4, 4);
  if (!r.ok || stsc.body + 8 + 12 * n_runs > stsc.end) return JAADB_E_CONFIG;

  // decoding times (stts): timeOffsets[] of Track.java:113-124
  uint64_t n_tt = r.be(stts.body + 4, 4);
  if (!r.ok || stts.body + 8 + 8 * n_tt > stts.end) return JAADB_E_CONFIG;
  std::vector<uint64_t> times(n_samples, 0);
  {
    uint64_t t = 0, k = 0;
    for (uint64_t i = 0; i < n_tt; ++i) {
      uint64_t cnt = r.be(stts.body + 8 + 8 * i, 4), delta = r.be(stts.body + 12 + 8 * i, 4);
      if (cnt > n_samples - k) return JAADB_E_CONFIG;   // ArrayIndexOutOfBoundsException in JAAD
      for (uint64_t j = 0; j < cnt; ++j) {
        times[k++] = t;
        t += delta;
      }
    }
  }

  out.clear();
  out.reserve(n_samples);
  uint64_t cur = 0;
  for (uint64_t i = 0; i < n_runs; ++i) {
    uint64_t first = r.be(stsc.body + 8 + 12 * i, 4), per = r.be(stsc.body + 12 + 12 * i, 4);
    uint64_t last = i + 1 < n_runs ? r.be(stsc.body + 8 + 12 * (i + 1), 4) - 1 : n_chunks;
    if (first == 0 || last > n_chunks) return JAADB_E_CONFIG;
    for (uint64_t j = first - 1; j < last; ++j) {
      uint64_t off = chunk_offset(j);
      if (per > n_samples - cur) return JAADB_E_CONFIG;
      for (uint64_t k = 0; k < per; ++k) {
        uint32_t sz = sample_size(cur);
        out.push_back(Sample{off, times[cur], sz});
        off += sz;
        ++cur;
      }
    }
  }
  if (!r.ok) return JAADB_E_CONFIG;
  // "frames need not to be time-ordered: sort by timestamp" (Collections.sort is stable)
  if (!std::is_sorted(out.begin(), out.end(), [](const Sample& a, const Sample& b) { return a.time < b.time; }))
    std::stable_sort(out.begin(), out.end(), [](const Sample& a, const Sample& b) { return a.time < b.time; });
  return JAADB_OK;
}

int64_t mp4_index(const uint8_t* file, uint64_t nbytes, uint64_t blob_offset, int32_t stream_id, jaadb_frame_desc* frames,
                  uint64_t max_frames, jaadb_mp4_track* track) {
  Reader r{file, nbytes};
  jaadb_mp4_track t;
  std::memset(&t, 0, sizeof t);
  Box moov;
  if (!find_child(r, 0, nbytes, fourcc("moov"), moov)) return JAADB_E_CONFIG;
  Box trak;
  for (uint64_t p = moov.body; read_box(r, p, moov.end, trak); p = trak.end) {
    if (trak.type != fourcc("trak")) continue;
    Box tkhd, mdia, mdhd, hdlr, minf, stbl, stsd;
    if (!find_child(r, trak.body, trak.end, fourcc("mdia"), mdia)) continue;
    if (!find_child(r, mdia.body, mdia.end, fourcc("hdlr"), hdlr)) continue;
    if ((uint32_t)r.be(hdlr.body + 8, 4) != fourcc("soun")) continue;       // Movie.java:60-70: handler type selects AudioTrack
    if (!find_child(r, mdia.body, mdia.end, fourcc("minf"), minf)) continue;
    if (!find_child(r, minf.body, minf.end, fourcc("stbl"), stbl)) continue;
    if (!find_child(r, stbl.body, stbl.end, fourcc("stsd"), stsd)) continue;
    // first sample entry; AAC tracks carry 'mp4a' (AudioTrack.AudioCodec.AAC)
    Box entry;
    if (!read_box(r, stsd.body + 8, stsd.end, entry) || entry.type != fourcc("mp4a")) continue;
    uint64_t e = entry.body;
    t.channel_count = (uint32_t)r.be(e + 16, 2);
    t.sample_size_bits = (uint32_t)r.be(e + 18, 2);
    t.sample_rate = (uint32_t)r.be(e + 24, 2);
    Box esds;
    if (!find_child(r, e + 28, entry.end, fourcc("esds"), esds)) continue;
    if (!find_asc(r, esds, &t)) return JAADB_E_CONFIG;
    if (find_child(r, trak.body, trak.end, fourcc("tkhd"), tkhd)) {
      int v = (int)r.be(tkhd.body, 1);
      t.track_id = (int32_t)r.be(tkhd.body + 4 + (v == 1 ? 16 : 8), 4);
    }
    if (find_child(r, mdia.body, mdia.end, fourcc("mdhd"), mdhd)) {
      int v = (int)r.be(mdhd.body, 1);
      t.timescale = (uint32_t)r.be(mdhd.body + 4 + (v == 1 ? 16 : 8), 4);
      t.duration = r.be(mdhd.body + 4 + (v == 1 ? 20 : 12), v == 1 ? 8 : 4);
    }
    std::vector<Sample> samples;
    int rc = parse_stbl(r, stbl, samples);
    if (rc != JAADB_OK) return rc;
    int64_t cnt = 0;
    for (const Sample& s : samples) {
      if (s.size > nbytes || s.offset > nbytes - s.size) break;   // EOFException while reading the frame: the stream ends here
      if (frames && (uint64_t)cnt < max_frames) {
        frames[cnt].offset = blob_offset + s.offset;
        frames[cnt].nbytes = s.size;
        frames[cnt].stream_id = stream_id;
      }
      ++cnt;
    }
    t.n_frames = (uint64_t)cnt;
    if (track) *track = t;
    return cnt;
  }
  return JAADB_E_CONFIG;   // "movie does not contain any AAC track" (S/Main.java:58-61)
}

template <typename F>
void parallel_streams(uint32_t n_streams, uint32_t threads, F&& fn) {
  if (threads == 0) threads = std::max(1u, std::thread::hardware_concurrency());
  threads = std::min(threads, std::max(1u, n_streams));
  std::atomic<uint32_t> next(0);
  // (fn never throws: the per-stream indexers catch everything themselves -- an exception leaving a std::thread is std::terminate)
  auto worker = [&]() {
    for (;;) {
      uint32_t s = next.fetch_add(1);
      if (s >= n_streams) break;
      fn(s);
    }
  };
  if (threads == 1) { worker(); return; }
  std::vector<std::thread> pool;
  pool.reserve(threads);
  try {
    for (uint32_t i = 0; i < threads; ++i) pool.emplace_back(worker);
  } catch (...) {
    // the host refused another thread: the ones that exist (and this one) finish the work
  }
  worker();
  for (auto& th : pool) th.join();
}

// Shared driver of the two *_index_many calls.  One pass over the containers (the sync search / sample tables are the
// expensive part: a cache miss per frame): every stream's frames go into a per-stream vector, then a prefix sum places
// them in the caller's table.
// ADTS streams, eight at a time per host thread: every frame header is a cache miss (frames are a few hundred bytes apart in
// blobs far larger than the caches), and a stream's next header is only known once the current one is parsed -- so one
// walker has one miss in flight.  Stepping eight independent walkers in turn, each prefetching its next header before the
// others take their step, keeps eight in flight (measured: 125 ns -> ~25 ns per frame and thread).
constexpr uint32_t kAdtsGroup = 8;
void adts_index_group(const uint8_t* blob, const uint64_t* stream_begin, const int32_t* stream_ids, uint32_t s0, uint32_t s1,
                      jaadb_frame_desc* stage, const uint64_t* at, int64_t* count, jaadb_adts_info* local) {
  AdtsWalker w[kAdtsGroup];
  bool alive[kAdtsGroup];
  const uint32_t n = s1 - s0;
  for (uint32_t j = 0; j < n; ++j) {
    w[j] = AdtsWalker{blob + stream_begin[s0 + j], stream_begin[s0 + j + 1] - stream_begin[s0 + j]};
    alive[j] = true;
    count[s0 + j] = 0;
    std::memset(&local[s0 + j], 0, sizeof(jaadb_adts_info));
  }
  for (bool any = true; any;) {
    any = false;
    for (uint32_t j = 0; j < n; ++j) {
      if (!alive[j]) continue;
      const uint32_t s = s0 + j;
      uint64_t off;
      uint32_t len;
      if (!w[j].next(off, len, count[s] == 0 ? &local[s] : nullptr)) { alive[j] = false; continue; }
      any = true;
      if (stage && (uint64_t)count[s] < at[s + 1] - at[s]) {
        jaadb_frame_desc& f = stage[at[s] + (uint64_t)count[s]];
        f.offset = stream_begin[s] + off;
        f.nbytes = len;
        f.stream_id = stream_ids ? stream_ids[s] : (int32_t)s;
      }
      ++count[s];
      if (w[j].pos + 8 <= w[j].n) __builtin_prefetch(w[j].d + w[j].pos);
    }
  }
  for (uint32_t j = 0; j < n; ++j) local[s0 + j].n_frames = (uint64_t)count[s0 + j];
}

template <typename Info, typename One>
int64_t index_many(const uint8_t* blob, const uint64_t* stream_begin, uint32_t n_streams, const int32_t* stream_ids,
                          jaadb_frame_desc* frames, uint64_t max_frames, uint64_t* first_frame, Info* infos,
                          uint32_t threads, One one) {
  if (!blob || !stream_begin) return JAADB_E_INVALID;
  for (uint32_t s = 0; s < n_streams; ++s)
    if (stream_begin[s + 1] < stream_begin[s]) return JAADB_E_INVALID;
  // no exception crosses the C ABI or leaves a worker thread: allocation failures inside one stream's indexer fail that stream
  auto one_safe = [&one](const uint8_t* d, uint64_t n, uint64_t off, int32_t id, jaadb_frame_desc* fr, uint64_t mx, Info* info) -> int64_t {
    try { return one(d, n, off, id, fr, mx, info); }
    catch (const std::bad_alloc&) { return JAADB_E_NOMEM; }
    catch (...) { return JAADB_E_INVALID; }
  };
  std::vector<int64_t> count(n_streams, 0);
  std::vector<Info> local(n_streams);
  // staging rows: stream s may write cap[s] = bytes / 64 + 16 rows at stage + at[s] (one uninitialised allocation: only the
  // rows that are written are ever touched); a stream with more frames than that is indexed again into its own table
  std::vector<uint64_t> at(n_streams + 1, 0);
  std::unique_ptr<jaadb_frame_desc[]> stage;
  std::vector<std::vector<jaadb_frame_desc>> big(frames ? n_streams : 0);
  if (frames) {
    for (uint32_t s = 0; s < n_streams; ++s) at[s + 1] = at[s] + std::min<uint64_t>((stream_begin[s + 1] - stream_begin[s]) / 64 + 16, 1u << 22);
    stage.reset(new jaadb_frame_desc[at[n_streams]]);
  }
  auto redo_big = [&](uint32_t s) {   // more frames than the staging guess: once more, into a table of the right size
    const uint64_t bytes = stream_begin[s + 1] - stream_begin[s];
    try { big[s].resize((size_t)count[s]); one_safe(blob + stream_begin[s], bytes, stream_begin[s], stream_ids ? stream_ids[s] : (int32_t)s, big[s].data(), (uint64_t)count[s], nullptr); }
    catch (...) { count[s] = JAADB_E_NOMEM; }
  };
  if constexpr (std::is_same<Info, jaadb_adts_info>::value) {
    const uint32_t n_groups = (n_streams + kAdtsGroup - 1) / kAdtsGroup;
    parallel_streams(n_groups, threads, [&](uint32_t g) {
      const uint32_t s0 = g * kAdtsGroup, s1 = std::min(n_streams, s0 + kAdtsGroup);
      adts_index_group(blob, stream_begin, stream_ids, s0, s1, stage.get(), at.data(), count.data(), local.data());
      for (uint32_t s = s0; s < s1; ++s)
        if (frames && count[s] > (int64_t)(at[s + 1] - at[s])) redo_big(s);
    });
  } else {
    parallel_streams(n_streams, threads, [&](uint32_t s) {
      const uint64_t bytes = stream_begin[s + 1] - stream_begin[s];
      const uint64_t cap = frames ? at[s + 1] - at[s] : 0;
      count[s] = one_safe(blob + stream_begin[s], bytes, stream_begin[s], stream_ids ? stream_ids[s] : (int32_t)s,
                          frames ? stage.get() + at[s] : nullptr, cap, &local[s]);
      if (frames && count[s] > (int64_t)cap) redo_big(s);
    });
  }
  std::vector<uint64_t> first(n_streams + 1, 0);
  for (uint32_t s = 0; s < n_streams; ++s) first[s + 1] = first[s] + (uint64_t)std::max<int64_t>(count[s], 0);
  if (first_frame) std::memcpy(first_frame, first.data(), (n_streams + 1) * sizeof(uint64_t));
  if (infos)
    for (uint32_t s = 0; s < n_streams; ++s) {
      infos[s] = local[s];
      if (count[s] < 0) infos[s].n_frames = 0;
    }
  if (frames && first[n_streams] <= max_frames)
    parallel_streams(n_streams, threads, [&](uint32_t s) {
      if (count[s] > 0)
        std::memcpy(frames + first[s], big[s].empty() ? stage.get() + at[s] : big[s].data(), (size_t)count[s] * sizeof(jaadb_frame_desc));
    });
  return (int64_t)first[n_streams];
}

// Stream-major frame table -> frame-major ("tick") order: frame 0 of every stream, frame 1 of every stream, ...; a stream
// that has ended simply drops out.  Rows of the result are independent, so they are filled on host threads.
int64_t frames_interleave(const jaadb_frame_desc* in, const uint64_t* first_frame, uint32_t n_streams, jaadb_frame_desc* out,
                          uint32_t threads) {
  uint64_t longest = 0;
  for (uint32_t s = 0; s < n_streams; ++s) {
    if (first_frame[s + 1] < first_frame[s]) return JAADB_E_INVALID;
    longest = std::max(longest, first_frame[s + 1] - first_frame[s]);
  }
  // row_start[f] = frames in rows before f = sum over streams of min(len, f)
  std::vector<uint64_t> alive(longest + 1, 0), row_start(longest + 1, 0);
  for (uint32_t s = 0; s < n_streams; ++s) alive[first_frame[s + 1] - first_frame[s]]++;   // histogram of lengths
  uint64_t ge = n_streams;   // streams with len > f
  for (uint64_t f = 0; f < longest; ++f) {
    ge -= alive[f];          // streams of length exactly f are gone in row f
    row_start[f + 1] = row_start[f] + ge;
  }
  const uint32_t n_rows = (uint32_t)std::min<uint64_t>(longest, 0xFFFFFFFFu);
  parallel_streams(n_rows, threads, [&](uint32_t f) {
    uint64_t pos = row_start[f];
    for (uint32_t s = 0; s < n_streams; ++s)
      if (first_frame[s + 1] - first_frame[s] > f) out[pos++] = in[first_frame[s] + f];
  });
  return (int64_t)row_start[longest];
}

}  // namespace

// Used by jaadb_decode_containers (jaadb_engine.cu): containers -> frame table in frame-major order, `frames` resized to fit.
// kind 0: ADTS streams, 1: MP4 files.  Returns the number of frames or a negative JAADB_E_* code; never throws.
int64_t jaadb_internal_index_interleaved(int kind, const uint8_t* blob, const uint64_t* begin, uint32_t n_streams, const int32_t* stream_ids,
                                         std::vector<jaadb_frame_desc>& scratch, std::vector<jaadb_frame_desc>& frames, uint32_t threads) {
  try {
    std::vector<uint64_t> first(n_streams + 1, 0);
    // a first call that only counts would cost a second pass over the containers: guess, and grow if the guess was short
    uint64_t guess = scratch.size() ? scratch.size() : (begin[n_streams] - begin[0]) / 256 + 1024;
    for (int attempt = 0; attempt < 2; ++attempt) {
      scratch.resize((size_t)guess);
      const int64_t n = kind == 1
          ? index_many(blob, begin, n_streams, stream_ids, scratch.data(), (uint64_t)scratch.size(), first.data(), (jaadb_mp4_track*)nullptr, threads, mp4_index)
          : index_many(blob, begin, n_streams, stream_ids, scratch.data(), (uint64_t)scratch.size(), first.data(), (jaadb_adts_info*)nullptr, threads, adts_index);
      if (n < 0) return n;
      if ((uint64_t)n <= scratch.size()) {
        frames.resize((size_t)n);
        if (n == 0) return 0;
        return frames_interleave(scratch.data(), first.data(), n_streams, frames.data(), threads);
      }
      guess = (uint64_t)n;
    }
    return JAADB_E_INVALID;
  } catch (const std::bad_alloc&) { return JAADB_E_NOMEM; }
  catch (...) { return JAADB_E_INVALID; }
}

extern "C" {

int64_t jaadb_frames_interleave(const jaadb_frame_desc* frames, const uint64_t* first_frame, uint32_t n_streams,
                                jaadb_frame_desc* out, uint32_t threads) {
  if (!frames || !first_frame || !out) return JAADB_E_INVALID;
  try { return frames_interleave(frames, first_frame, n_streams, out, threads); }
  catch (const std::bad_alloc&) { return JAADB_E_NOMEM; }
  catch (...) { return JAADB_E_INVALID; }
}

int64_t jaadb_adts_index(const uint8_t* data, uint64_t nbytes, uint64_t blob_offset, int32_t stream_id,
                         jaadb_frame_desc* frames, uint64_t max_frames, jaadb_adts_info* info) {
  if (!data && nbytes) return JAADB_E_INVALID;
  try { return adts_index(data, nbytes, blob_offset, stream_id, frames, max_frames, info); }
  catch (const std::bad_alloc&) { return JAADB_E_NOMEM; }
  catch (...) { return JAADB_E_INVALID; }
}

int64_t jaadb_mp4_index(const uint8_t* file, uint64_t nbytes, uint64_t blob_offset, int32_t stream_id,
                        jaadb_frame_desc* frames, uint64_t max_frames, jaadb_mp4_track* track) {
  if (!file) return JAADB_E_INVALID;
  try { return mp4_index(file, nbytes, blob_offset, stream_id, frames, max_frames, track); }
  catch (const std::bad_alloc&) { return JAADB_E_NOMEM; }
  catch (...) { return JAADB_E_INVALID; }
}

int64_t jaadb_adts_index_many(const uint8_t* blob, const uint64_t* stream_begin, uint32_t n_streams,
                              const int32_t* stream_ids, jaadb_frame_desc* frames, uint64_t max_frames,
                              uint64_t* first_frame, jaadb_adts_info* infos, uint32_t threads) {
  try { return index_many(blob, stream_begin, n_streams, stream_ids, frames, max_frames, first_frame, infos, threads, adts_index); }
  catch (const std::bad_alloc&) { return JAADB_E_NOMEM; }
  catch (...) { return JAADB_E_INVALID; }
}

int64_t jaadb_mp4_index_many(const uint8_t* blob, const uint64_t* file_begin, uint32_t n_files, const int32_t* stream_ids,
                             jaadb_frame_desc* frames, uint64_t max_frames, uint64_t* first_frame,
                             jaadb_mp4_track* tracks, uint32_t threads) {
  try { return index_many(blob, file_begin, n_files, stream_ids, frames, max_frames, first_frame, tracks, threads, mp4_index); }
  catch (const std::bad_alloc&) { return JAADB_E_NOMEM; }
  catch (...) { return JAADB_E_INVALID; }
}

}  // extern "C"
