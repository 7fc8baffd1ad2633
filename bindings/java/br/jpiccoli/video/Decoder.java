package br.jpiccoli.video;

import java.io.BufferedInputStream;
import java.io.ByteArrayOutputStream;
import java.io.IOException;
import java.io.InputStream;
import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.nio.channels.FileChannel;
import java.nio.file.Files;
import java.nio.file.Path;
import java.nio.file.StandardOpenOption;
import java.util.zip.InflaterInputStream;

/**
 * Drop-in replacement of this package's Decoder: same command line and file formats; the ExpGolombReader loop,
 * de-quantisation, InverseDCT.run() and the byte conversion (Decoder.java:61-117) are one call into libdct3d.so.
 *
 * <pre>java br.jpiccoli.video.Decoder &lt;input&gt; &lt;output&gt; &lt;width&gt; &lt;height&gt; &lt;frames&gt;</pre>
 */
public final class Decoder {

    private static final int CUBE = 8;

    public static void main(String[] args) throws IOException {
        if (args.length < 5) {
            System.out.println("Usage: java Decoder <input file> <output file> <frame width> <frame height> <number of frames to decode>");
            System.exit(-1);
        }
        int width = Integer.parseInt(args[2]), height = Integer.parseInt(args[3]);
        int frames = Integer.parseInt(args[4]);
        frames -= frames % CUBE;                                           // Decoder.java:35-36
        Dct3d.Precision precision = "32".equals(System.getProperty("dct3d.precision")) ? Dct3d.Precision.FLOAT : Dct3d.Precision.DOUBLE;

        System.out.println("Inflating");
        byte[] stream;
        try (InputStream z = new InflaterInputStream(new BufferedInputStream(Files.newInputStream(Path.of(args[0])), 1 << 20))) {
            ByteArrayOutputStream all = new ByteArrayOutputStream(1 << 24);
            z.transferTo(all);
            stream = all.toByteArray();
        }

        System.out.println("Decoding on the GPU");
        try (Arena arena = Arena.ofConfined();
             Dct3d gpu = new Dct3d(Integer.getInteger("dct3d.device", 0), width, height, CUBE, precision);
             FileChannel ch = FileChannel.open(Path.of(args[1]), StandardOpenOption.CREATE, StandardOpenOption.WRITE,
                                               StandardOpenOption.TRUNCATE_EXISTING)) {
            MemorySegment in = arena.allocate(Math.max(1, stream.length));
            MemorySegment.copy(stream, 0, in, java.lang.foreign.ValueLayout.JAVA_BYTE, 0, stream.length);
            MemorySegment pixels = gpu.decode(in.asSlice(0, stream.length), frames, arena);
            long left = pixels.byteSize(), at = 0;
            while (left > 0) {                                             // ByteBuffer views are limited to 2^31-1 bytes
                long n = Math.min(left, 1L << 30);
                ch.write(pixels.asSlice(at, n).asByteBuffer());
                at += n;
                left -= n;
            }
        }
        System.out.println("Complete!");
    }
}
