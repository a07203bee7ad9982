import net.sourceforge.jaad.SampleBuffer;
import net.sourceforge.jaad.aac.Decoder;
import net.sourceforge.jaad.aac.syntax.ByteArrayBitStream;
import net.sourceforge.jaad.adts.ADTSDemultiplexer;

import java.io.ByteArrayInputStream;
import java.nio.ByteBuffer;
import java.nio.file.Files;
import java.nio.file.Path;
import java.util.concurrent.*;
import java.util.concurrent.atomic.AtomicLong;

/**
 * CPU baseline on a host with a JVM: Main.decodeAAC's loop (src/.../Main.java:82-111), one Decoder per stream, one thread
 * per core.  Usage: java JaadBench dir-with-adts-files [threads].  Prints decoded audio-seconds per wall second.
 * SOURCE ONLY (no JVM in the build image; bench.py times the C++ restatement instead and says so).
 */
public class JaadBench {
	public static void main(String[] args) throws Exception {
		Path[] files = Files.list(Path.of(args[0])).filter(p -> p.toString().endsWith(".aac")).toArray(Path[]::new);
		int threads = args.length > 1 ? Integer.parseInt(args[1]) : Runtime.getRuntime().availableProcessors();
		byte[][] data = new byte[files.length][];
		for (int i = 0; i < files.length; i++) data[i] = Files.readAllBytes(files[i]);
		ExecutorService pool = Executors.newFixedThreadPool(threads);
		AtomicLong microAudio = new AtomicLong();
		long t0 = System.nanoTime();
		CountDownLatch done = new CountDownLatch(files.length);
		for (byte[] d : data)
			pool.submit(() -> {
				try {
					ADTSDemultiplexer adts = new ADTSDemultiplexer(new ByteArrayInputStream(d));
					Decoder dec = Decoder.create(adts.getDecoderInfo());
					SampleBuffer buf = new SampleBuffer(dec.getConfig().getSampleLength() * 2 * 8);
					ByteBuffer cbb = ByteBuffer.allocateDirect(ADTSDemultiplexer.MAXIMUM_FRAME_SIZE);
					ByteArrayBitStream bs = new ByteArrayBitStream();
					long samples = 0;
					try {
						while (true) {
							adts.readNextFrame(cbb);
							cbb.flip();
							bs.setData(cbb);
							cbb.clear();
							dec.decodeFrame(bs, buf);
							samples += dec.getConfig().getSampleLength();
						}
					} catch (java.io.IOException eof) { /* end of stream */ }
					microAudio.addAndGet(samples * 1_000_000L / adts.getSampleFrequency());
				} catch (Exception e) {
					e.printStackTrace();
				} finally {
					done.countDown();
				}
			});
		done.await();
		double wall = (System.nanoTime() - t0) * 1e-9;
		pool.shutdown();
		System.out.printf("{\"impl\": \"jaad-jvm\", \"threads\": %d, \"streams\": %d, \"value\": %.1f, \"unit\": \"audio-s/s\"}%n",
				threads, files.length, microAudio.get() * 1e-6 / wall);
	}
}
