package net.sourceforge.jaad.b200;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.*;

/**
 * Panama FFM (Java 22+, the reference builds at source level 21 with preview or 22) binding of
 * include/jaadb200.h.  One NativeEngine drives one GPU.  SOURCE ONLY: the build image has no JVM;
 * tests drive the same C ABI from Python ctypes (jaadec_b200/_lib.py).
 *
 * Replaces, for the decode path only, the objects JAAD creates in
 * net.sourceforge.jaad.aac.Decoder.create(...) (aac/.../Decoder.java:36-54).
 */
public final class NativeEngine implements AutoCloseable {

	public static final int PCM_S16LE = 0, PCM_S16BE = 1, PCM_F32_PLANAR = 2;

	/** jaadb_frame_desc { uint64 offset; uint32 nbytes; int32 stream_id; } */
	public static final StructLayout FRAME_DESC = MemoryLayout.structLayout(
			JAVA_LONG.withName("offset"), JAVA_INT.withName("nbytes"), JAVA_INT.withName("stream_id"));
	/** jaadb_frame_result { int32 status; uint16 channels; uint16 sample_length; uint32 sample_rate; uint32 pcm_bytes; } */
	public static final StructLayout FRAME_RESULT = MemoryLayout.structLayout(
			JAVA_INT.withName("status"), JAVA_SHORT.withName("channels"), JAVA_SHORT.withName("sample_length"),
			JAVA_INT.withName("sample_rate"), JAVA_INT.withName("pcm_bytes"));
	/** jaadb_options.flags (include/jaadb200.h): per-kernel timings, parity-test taps, ISO pulse application
	 *  (JAAD parses pulse_data and never applies it, syntax/ICStream.java:17; the default keeps that). */
	public static final int FLAG_PROFILE = 1, FLAG_DEBUG_TAPS = 2, FLAG_PULSE_ISO = 4;
	/** jaadb_options { int32 device; uint32 max_streams; int32 pcm_format; int32 tns_mode; uint32 flags; uint32 chunk_frames;
	 *  uint32 sbr_tile_frames; uint32 k2_segment_frames; } -- the three tuning knobs stay 0 (defaults) here */
	static final StructLayout OPTIONS = MemoryLayout.structLayout(
			JAVA_INT.withName("device"), JAVA_INT.withName("max_streams"), JAVA_INT.withName("pcm_format"),
			JAVA_INT.withName("tns_mode"), JAVA_INT.withName("flags"), JAVA_INT.withName("chunk_frames"),
			JAVA_INT.withName("sbr_tile_frames"), JAVA_INT.withName("k2_segment_frames"));

	private static final Linker LINKER = Linker.nativeLinker();
	private static final SymbolLookup LIB = SymbolLookup.libraryLookup(
			System.getProperty("jaadb200.library", "libjaadb200.so"), Arena.global());

	private static MethodHandle h(String name, FunctionDescriptor fd) {
		return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
	}

	private static final MethodHandle ENGINE_CREATE = h("jaadb_engine_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
	private static final MethodHandle ENGINE_DESTROY = h("jaadb_engine_destroy", FunctionDescriptor.ofVoid(ADDRESS));
	private static final MethodHandle LAST_ERROR = h("jaadb_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
	private static final MethodHandle STATUS_STRING = h("jaadb_status_string", FunctionDescriptor.of(ADDRESS, JAVA_INT));
	private static final MethodHandle OPEN_ASC = h("jaadb_stream_open_asc", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS));
	private static final MethodHandle OPEN_ASC_SBR = h("jaadb_stream_open_asc_sbr", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));
	private static final MethodHandle OPEN_ADTS = h("jaadb_stream_open_adts", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS));
	private static final MethodHandle PROBE_SBR = h("jaadb_probe_sbr", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
	private static final MethodHandle PROBE_SBR_ASC = h("jaadb_probe_sbr_asc", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
	private static final MethodHandle STREAM_CLOSE = h("jaadb_stream_close", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT));
	private static final MethodHandle DECODE = h("jaadb_decode", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS));

	/** jaadb_adts_info { int32 profile, sf_index, channel_config, sample_rate; uint64 n_frames; } */
	public static final StructLayout ADTS_INFO = MemoryLayout.structLayout(
			JAVA_INT.withName("profile"), JAVA_INT.withName("sf_index"), JAVA_INT.withName("channel_config"),
			JAVA_INT.withName("sample_rate"), JAVA_LONG.withName("n_frames"));
	/** jaadb_mp4_track { uint8 asc[64]; uint32 asc_bytes; int32 track_id; uint32 timescale, channel_count, sample_size_bits,
	 *  sample_rate, object_type, max_bitrate, avg_bitrate, reserved; uint64 duration, n_frames; } */
	public static final StructLayout MP4_TRACK = MemoryLayout.structLayout(
			MemoryLayout.sequenceLayout(64, JAVA_BYTE).withName("asc"), JAVA_INT.withName("asc_bytes"), JAVA_INT.withName("track_id"),
			JAVA_INT.withName("timescale"), JAVA_INT.withName("channel_count"), JAVA_INT.withName("sample_size_bits"),
			JAVA_INT.withName("sample_rate"), JAVA_INT.withName("object_type"), JAVA_INT.withName("max_bitrate"),
			JAVA_INT.withName("avg_bitrate"), JAVA_INT.withName("reserved"), JAVA_LONG.withName("duration"),
			JAVA_LONG.withName("n_frames"));
	// int64 jaadb_{adts,mp4}_index(const uint8* data, uint64 nbytes, uint64 blob_offset, int32 stream_id,
	//                              jaadb_frame_desc* frames, uint64 max_frames, info* out)
	private static final FunctionDescriptor INDEX_FD =
			FunctionDescriptor.of(JAVA_LONG, ADDRESS, JAVA_LONG, JAVA_LONG, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS);
	private static final MethodHandle ADTS_INDEX = h("jaadb_adts_index", INDEX_FD);
	private static final MethodHandle MP4_INDEX = h("jaadb_mp4_index", INDEX_FD);

	/**
	 * The frame table of an ADTS stream held in native memory: what new ADTSDemultiplexer(in) + the readNextFrame() loop
	 * of Main.decodeAAC produce (src/.../adts/ADTSDemultiplexer.java:26-74), without touching the payload bytes.
	 * `frames` receives up to maxFrames jaadb_frame_desc rows (may be NULL to count), `info` an ADTS_INFO.
	 */
	public static long adtsIndex(MemorySegment data, long blobOffset, int streamId, MemorySegment frames, long maxFrames,
								 MemorySegment info) {
		try {
			return (long) ADTS_INDEX.invokeExact(data, data.byteSize(), blobOffset, streamId,
					frames == null ? MemorySegment.NULL : frames, maxFrames, info == null ? MemorySegment.NULL : info);
		} catch (Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	/**
	 * The sample table and AudioSpecificConfig of the first AAC track of an MP4 file held in native memory: what
	 * new MP4Container(in).getMovie().getTracks(AudioTrack.AudioCodec.AAC).get(0) + getDecoderSpecificInfo() + the
	 * readNextFrame() loop of Main.decodeMP4 produce (mp4/.../api/Track.java:90-172).  Negative: no AAC track / malformed.
	 */
	public static long mp4Index(MemorySegment file, long blobOffset, int streamId, MemorySegment frames, long maxFrames,
								MemorySegment track) {
		try {
			return (long) MP4_INDEX.invokeExact(file, file.byteSize(), blobOffset, streamId,
					frames == null ? MemorySegment.NULL : frames, maxFrames, track == null ? MemorySegment.NULL : track);
		} catch (Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	private final MemorySegment engine;
	public final int pcmFormat;

	public NativeEngine(int device, int maxStreams, int pcmFormat) {
		this.pcmFormat = pcmFormat;
		try (Arena a = Arena.ofConfined()) {
			MemorySegment o = a.allocate(OPTIONS);
			o.set(JAVA_INT, 0, device);
			o.set(JAVA_INT, 4, maxStreams);
			o.set(JAVA_INT, 8, pcmFormat);
			MemorySegment out = a.allocate(ADDRESS);
			int rc = (int) ENGINE_CREATE.invokeExact(o, out);
			if (rc != 0)
				throw new IllegalStateException("jaadb_engine_create failed: " + rc + " (no CUDA device? there is no CPU fallback)");
			engine = out.get(ADDRESS, 0);
		} catch (RuntimeException | Error e) {
			throw e;
		} catch (Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	String lastError() {
		try {
			return ((MemorySegment) LAST_ERROR.invokeExact(engine)).reinterpret(4096).getString(0);
		} catch (Throwable t) {
			return t.toString();
		}
	}

	public static String statusString(int status) {
		try {
			return ((MemorySegment) STATUS_STRING.invokeExact(status)).reinterpret(256).getString(0);
		} catch (Throwable t) {
			return "status " + status;
		}
	}

	/** Decoder.create(byte[] audioSpecificConfig), aac/.../Decoder.java:36-43. */
	public int openAsc(byte[] asc) {
		try (Arena a = Arena.ofConfined()) {
			MemorySegment buf = a.allocateFrom(JAVA_BYTE, asc);
			MemorySegment id = a.allocate(JAVA_INT);
			int rc = (int) OPEN_ASC.invokeExact(engine, buf, asc.length, id);
			if (rc != 0) throw new IllegalArgumentException("jaadb_stream_open_asc: " + lastError());
			return id.get(JAVA_INT, 0);
		} catch (RuntimeException e) {
			throw e;
		} catch (Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	/** Decoder.create(byte[] asc) for a track whose frames carry SBR (2: + PS) the ASC does not signal: JAAD's down-sampled SBR tool. */
	public int openAsc(byte[] asc, int expectSbr) {
		try (Arena a = Arena.ofConfined()) {
			MemorySegment buf = a.allocateFrom(JAVA_BYTE, asc);
			MemorySegment id = a.allocate(JAVA_INT);
			int rc = (int) OPEN_ASC_SBR.invokeExact(engine, buf, asc.length, expectSbr, id);
			if (rc != 0) throw new IllegalArgumentException("jaadb_stream_open_asc_sbr: " + lastError());
			return id.get(JAVA_INT, 0);
		} catch (RuntimeException e) {
			throw e;
		} catch (Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	/** Decoder.create(AudioDecoderInfo) from an ADTS header, aac/.../Decoder.java:45-48. */
	public int openAdts(int profileIndex, int sfIndex, int channelConfig, int expectSbr) {
		try (Arena a = Arena.ofConfined()) {
			MemorySegment id = a.allocate(JAVA_INT);
			int rc = (int) OPEN_ADTS.invokeExact(engine, profileIndex, sfIndex, channelConfig, expectSbr, id);
			if (rc != 0) throw new IllegalArgumentException("jaadb_stream_open_adts: " + lastError());
			return id.get(JAVA_INT, 0);
		} catch (RuntimeException e) {
			throw e;
		} catch (Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	/**
	 * What expectSbr should be for an LC-signalled ADTS stream, judged by its first raw_data_block: JAAD creates the SBR
	 * (and PS) tool when the first payload arrives (aac/.../syntax/ChannelElement.java:65-76); the batched engine decides
	 * when the stream is opened.  0 = plain AAC-LC, 1 = SBR, 2 = SBR + parametric stereo.
	 */
	public int probeSbr(int profileIndex, int sfIndex, int channelConfig, byte[] firstFrame) {
		try (Arena a = Arena.ofConfined()) {
			MemorySegment buf = a.allocateFrom(JAVA_BYTE, firstFrame);
			MemorySegment out = a.allocate(JAVA_INT);
			int rc = (int) PROBE_SBR.invokeExact(engine, profileIndex, sfIndex, channelConfig, buf, firstFrame.length, out);
			if (rc != 0) throw new IllegalArgumentException("jaadb_probe_sbr: " + lastError());
			return out.get(JAVA_INT, 0);
		} catch (RuntimeException e) {
			throw e;
		} catch (Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	/** The same for an MP4 track described by its AudioSpecificConfig; feeds {@link #openAsc(byte[], int)}. */
	public int probeSbrAsc(byte[] asc, byte[] firstSample) {
		try (Arena a = Arena.ofConfined()) {
			MemorySegment cfg = a.allocateFrom(JAVA_BYTE, asc);
			MemorySegment buf = a.allocateFrom(JAVA_BYTE, firstSample);
			MemorySegment out = a.allocate(JAVA_INT);
			int rc = (int) PROBE_SBR_ASC.invokeExact(engine, cfg, asc.length, buf, firstSample.length, out);
			if (rc != 0) throw new IllegalArgumentException("jaadb_probe_sbr_asc: " + lastError());
			return out.get(JAVA_INT, 0);
		} catch (RuntimeException e) {
			throw e;
		} catch (Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	public void closeStream(int id) {
		try {
			int rc = (int) STREAM_CLOSE.invokeExact(engine, id);
			if (rc != 0) throw new IllegalArgumentException("jaadb_stream_close: " + rc);
		} catch (RuntimeException e) {
			throw e;
		} catch (Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	/**
	 * Batched decodeFrame (aac/.../Decoder.java:89-121) + SampleBuffer.accept (src/.../SampleBuffer.java:168-209).
	 * All segments are caller-owned native memory (use pinned/direct buffers); frames of one stream are applied in
	 * array order. Returns 0, per-frame status words land in `results`.
	 */
	public void decode(MemorySegment blob, MemorySegment frames, int nFrames, MemorySegment pcmOut, MemorySegment pcmOffsets,
					   MemorySegment results) {
		try {
			int rc = (int) DECODE.invokeExact(engine, blob, blob.byteSize(), frames, nFrames, pcmOut, pcmOut.byteSize(),
					pcmOffsets == null ? MemorySegment.NULL : pcmOffsets, results);
			if (rc != 0) throw new IllegalStateException("jaadb_decode: " + rc + " " + lastError());
		} catch (RuntimeException e) {
			throw e;
		} catch (Throwable t) {
			throw new IllegalStateException(t);
		}
	}

	@Override
	public void close() {
		try {
			ENGINE_DESTROY.invokeExact(engine);
		} catch (Throwable t) {
			throw new IllegalStateException(t);
		}
	}
}
