import net.sourceforge.jaad.aac.Decoder;
import net.sourceforge.jaad.aac.Receiver;
import net.sourceforge.jaad.aac.syntax.ByteArrayBitStream;
import net.sourceforge.jaad.adts.ADTSDemultiplexer;

import java.io.*;
import java.nio.ByteBuffer;
import java.nio.ByteOrder;
import java.util.Collection;

/**
 * Dumps what real JAAD produces for an ADTS file so that tests/golden can be re-pinned on a host with a JVM:
 * per frame, the float[] channels handed to Receiver.accept (little-endian float32, planar) -- the quantity the
 * GPU tests compare bit for bit.  Usage: java -cp jaad.jar:. JaadDump in.aac out.f32
 * SOURCE ONLY: there is no JVM in the build image (DESIGN.md section 5).
 */
public class JaadDump {
	public static void main(String[] args) throws IOException {
		final ADTSDemultiplexer adts = new ADTSDemultiplexer(new FileInputStream(args[0]));
		final Decoder dec = Decoder.create(adts.getDecoderInfo());
		final var cbb = ByteBuffer.allocateDirect(ADTSDemultiplexer.MAXIMUM_FRAME_SIZE);
		final var bitStream = new ByteArrayBitStream();
		try (DataOutputStream out = new DataOutputStream(new BufferedOutputStream(new FileOutputStream(args[1])))) {
			Receiver sink = (Collection<float[]> samples, int sampleLength, int sampleRate) -> {
				try {
					ByteBuffer bb = ByteBuffer.allocate(4 * sampleLength * samples.size()).order(ByteOrder.LITTLE_ENDIAN);
					for (float[] ch : samples)
						for (int i = 0; i < sampleLength; i++) bb.putFloat(ch[sampleLength * i / sampleLength]);
					out.write(bb.array());
				} catch (IOException e) {
					throw new UncheckedIOException(e);
				}
			};
			while (true) {
				try {
					adts.readNextFrame(cbb);
				} catch (IOException eof) {
					break;
				}
				cbb.flip();
				bitStream.setData(cbb);
				cbb.clear();
				dec.decodeFrame(bitStream, sink);
			}
		}
	}
}
