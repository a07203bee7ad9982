package net.sourceforge.jaad.b200;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.util.List;

import static java.lang.foreign.ValueLayout.*;

/**
 * The throughput path: Java keeps the container demux (ADTSDemultiplexer / MP4Container give frame boundaries), collects
 * {offset, size, stream} for many streams into one native blob and hands the whole batch to the GPU in one call.
 * SOURCE ONLY (no JVM in the build image).
 */
public final class BatchDecoder implements AutoCloseable {

	public record Frame(long offset, int nbytes, int streamId) {}

	private final NativeEngine engine;

	public BatchDecoder(NativeEngine engine) {
		this.engine = engine;
	}

	/** @return per-frame status words; PCM of frame i starts at pcmOffsets[i] (or packed back to back when null). */
	public int[] decode(MemorySegment blob, List<Frame> frames, MemorySegment pcmOut, long[] pcmOffsets) {
		int n = frames.size();
		// the descriptor / result tables live exactly as long as the call (a confined arena per call: nothing accumulates)
		try (Arena arena = Arena.ofConfined()) {
			MemorySegment desc = arena.allocate(NativeEngine.FRAME_DESC, n);
			for (int i = 0; i < n; i++) {
				Frame f = frames.get(i);
				desc.set(JAVA_LONG, 16L * i, f.offset());
				desc.set(JAVA_INT, 16L * i + 8, f.nbytes());
				desc.set(JAVA_INT, 16L * i + 12, f.streamId());
			}
			MemorySegment offs = pcmOffsets == null ? MemorySegment.NULL : arena.allocateFrom(JAVA_LONG, pcmOffsets);
			MemorySegment res = arena.allocate(NativeEngine.FRAME_RESULT, n);
			engine.decode(blob, desc, n, pcmOut, offs, res);
			int[] status = new int[n];
			for (int i = 0; i < n; i++) status[i] = res.get(JAVA_INT, 16L * i);
			return status;
		}
	}

	@Override
	public void close() {
	}
}
