import net.sourceforge.jaad.aac.Decoder;
import net.sourceforge.jaad.aac.Receiver;
import net.sourceforge.jaad.aac.syntax.ByteArrayBitStream;
import net.sourceforge.jaad.adts.ADTSDemultiplexer;
import net.sourceforge.jaad.mp4.MP4Container;
import net.sourceforge.jaad.mp4.MP4Input;
import net.sourceforge.jaad.mp4.api.AudioTrack;
import net.sourceforge.jaad.mp4.api.Movie;
import net.sourceforge.jaad.mp4.api.Track;

import java.io.*;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.Collection;
import java.util.List;

/**
 * Dumps what real JAAD produces for an ADTS (.aac) or MP4 (.mp4 / .m4a) file, frame by frame, so that the committed golden
 * vectors (tests/golden/*.npz) can be pinned against the reference itself on a host with a JVM -- the one thing this
 * repository's build image cannot do.  It walks the file exactly like net.sourceforge.jaad.Main does (decodeAAC, Main.java:
 * 82-111; decodeMP4, Main.java:49-80) but hands decodeFrame a Receiver that records the float[] channels JAAD delivers
 * (the quantity the GPU tests compare bit for bit) instead of a SampleBuffer.
 *
 * Output, little-endian, one record per call of decodeFrame:
 *   int32 status (0 = samples follow, 1 = the frame produced nothing: EOSException swallowed by decodeFrame,
 *                 2 = an exception left decodeFrame: its message is printed to stderr)
 *   int32 channels, int32 sampleLength, int32 sampleRate, then channels x sampleLength float32 (planar)
 *
 * Usage:  java -cp <jaad classes or jar>:. JaadDump in.aac|in.mp4 out.dump
 * (tools/jaad_verify/run.sh compiles and runs it over every golden stream and compares.)
 */
public class JaadDump {

	static final class Sink implements Receiver {
		final DataOutputStream out;
		boolean called;

		Sink(DataOutputStream out) { this.out = out; }

		@Override
		public void accept(Collection<float[]> samples, int sampleLength, int sampleRate) {
			called = true;
			try {
				ByteBuffer bb = ByteBuffer.allocate(16 + 4 * sampleLength * samples.size()).order(ByteOrder.LITTLE_ENDIAN);
				bb.putInt(0).putInt(samples.size()).putInt(sampleLength).putInt(sampleRate);
				for (float[] ch : samples) {
					// SampleBuffer.accept resamples by index when a channel's array is not sampleLength long (SampleBuffer.java:188-206)
					for (int i = 0; i < sampleLength; i++) bb.putFloat(ch[(int) ((long) ch.length * i / sampleLength)]);
				}
				out.write(bb.array());
			} catch (IOException e) {
				throw new UncheckedIOException(e);
			}
		}
	}

	static void header(DataOutputStream out, int status) throws IOException {
		ByteBuffer bb = ByteBuffer.allocate(16).order(ByteOrder.LITTLE_ENDIAN);
		bb.putInt(status).putInt(0).putInt(0).putInt(0);
		out.write(bb.array());
	}

	static void frame(Decoder dec, ByteArrayBitStream bits, Sink sink) throws IOException {
		sink.called = false;
		try {
			dec.decodeFrame(bits, sink);
			if (!sink.called) header(sink.out, 1);
		} catch (RuntimeException e) {
			System.err.println("decodeFrame: " + e);
			if (!sink.called) header(sink.out, 2);
		}
	}

	public static void main(String[] args) throws Exception {
		final String in = args[0];
		try (DataOutputStream out = new DataOutputStream(new BufferedOutputStream(new FileOutputStream(args[1])))) {
			final Sink sink = new Sink(out);
			final ByteArrayBitStream bits = new ByteArrayBitStream();
			if (in.endsWith(".mp4") || in.endsWith(".m4a")) {
				final MP4Container cont = new MP4Container(MP4Input.open(new RandomAccessFile(in, "r")));
				final Movie movie = cont.getMovie();
				final List<Track> tracks = movie.getTracks(AudioTrack.AudioCodec.AAC);
				if (tracks.isEmpty()) throw new Exception("movie does not contain any AAC track");
				final AudioTrack track = (AudioTrack) tracks.get(0);
				final Decoder dec = Decoder.create(track.getDecoderSpecificInfo().getData());
				while (track.hasMoreFrames()) {
					bits.setData(track.readNextFrame().getData());
					frame(dec, bits, sink);
				}
			} else {
				final ADTSDemultiplexer adts = new ADTSDemultiplexer(new FileInputStream(in));
				final Decoder dec = Decoder.create(adts.getDecoderInfo());
				final ByteBuffer cbb = ByteBuffer.allocateDirect(ADTSDemultiplexer.MAXIMUM_FRAME_SIZE);
				while (true) {
					try {
						adts.readNextFrame(cbb);
					} catch (IOException eof) {
						break;
					}
					cbb.flip();
					bits.setData(cbb);
					cbb.clear();
					frame(dec, bits, sink);
				}
			}
		}
	}
}
