package net.sourceforge.jaad.b200;

import net.sourceforge.jaad.SampleBuffer;
import net.sourceforge.jaad.aac.AACException;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.nio.ByteBuffer;

import static java.lang.foreign.ValueLayout.*;

/**
 * Batch-size-1 wrapper with the call shape of net.sourceforge.jaad.aac.Decoder
 * (create / decodeFrame), so Main.decodeAAC / decodeMP4 (src/.../Main.java:49-111) keep their loops:
 *
 *   Decoder dec = Decoder.create(asc);              ->  B200Decoder dec = B200Decoder.create(engine, asc);
 *   dec.decodeFrame(bitStream, sampleBuffer);       ->  dec.decodeFrame(frameBytes, sampleBuffer);
 *
 * Error behaviour mirrors Decoder.decodeFrame (aac/.../Decoder.java:89-101): an EOS frame is swallowed and yields no
 * PCM; every other status is rethrown as AACException with JAAD's message text.
 * SOURCE ONLY (no JVM in the build image).  For throughput use BatchDecoder.
 */
public final class B200Decoder implements AutoCloseable {

	private final NativeEngine engine;
	private final int streamId;
	private final Arena arena = Arena.ofConfined();
	private final MemorySegment blob = arena.allocate(6144 * 8);      // ADTSDemultiplexer.MAXIMUM_FRAME_SIZE per channel x 8
	private final MemorySegment desc = arena.allocate(NativeEngine.FRAME_DESC);
	private final MemorySegment result = arena.allocate(NativeEngine.FRAME_RESULT);
	private final MemorySegment pcm = arena.allocate(8 * 2048 * 4);
	public int frames = 0;

	private B200Decoder(NativeEngine engine, int streamId) {
		this.engine = engine;
		this.streamId = streamId;
	}

	public static B200Decoder create(NativeEngine engine, byte[] audioSpecificConfig) {
		return new B200Decoder(engine, engine.openAsc(audioSpecificConfig));
	}

	/**
	 * Decoder.create(byte[]) for a track whose first sample is at hand: JAAD decides about implicit SBR / PS when the first
	 * payload arrives (aac/.../syntax/ChannelElement.java:65-76), this engine when the stream is opened -- so the first sample
	 * is probed.  An ASC that does not signal SBR then runs JAAD's down-sampled SBR tool (1024 samples per frame).
	 */
	public static B200Decoder create(NativeEngine engine, byte[] audioSpecificConfig, byte[] firstSample) {
		return new B200Decoder(engine, engine.openAsc(audioSpecificConfig, engine.probeSbrAsc(audioSpecificConfig, firstSample)));
	}

	/** profileIndex/sfIndex/channelConfig exactly as ADTSFrame reports them (src/.../adts/ADTSFrame.java:119-129). */
	public static B200Decoder create(NativeEngine engine, int profileIndex, int sfIndex, int channelConfig, int expectSbr) {
		return new B200Decoder(engine, engine.openAdts(profileIndex, sfIndex, channelConfig, expectSbr));
	}

	public void decodeFrame(ByteBuffer frame, SampleBuffer out) {
		int n = frame.remaining();
		MemorySegment.copy(MemorySegment.ofBuffer(frame), 0, blob, 0, n);
		desc.set(JAVA_LONG, 0, 0L);
		desc.set(JAVA_INT, 8, n);
		desc.set(JAVA_INT, 12, streamId);
		try {
			engine.decode(blob.asSlice(0, Math.max(n, 1)), desc, 1, pcm, null, result);
		} finally {
			++frames;
		}
		int status = result.get(JAVA_INT, 0);
		if (status == 1) return;                       // EOSException: swallowed, no output for this frame
		if (status != 0) throw new AACException(NativeEngine.statusString(status));
		int bytes = result.get(JAVA_INT, 12);
		ByteBuffer bb = out.getBB();                   // same container SampleBuffer.accept fills (SampleBuffer.java:168-209)
		bb.clear();
		bb.put(pcm.asSlice(0, bytes).asByteBuffer());
		bb.flip();
		out.setFormat(result.get(JAVA_INT, 8), Short.toUnsignedInt(result.get(JAVA_SHORT, 4)), 16);   // see INTEGRATION.md: 3-line setter added to SampleBuffer
	}

	@Override
	public void close() {
		engine.closeStream(streamId);
		arena.close();
	}
}
