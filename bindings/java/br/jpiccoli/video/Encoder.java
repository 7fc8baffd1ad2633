package br.jpiccoli.video;

import java.io.BufferedOutputStream;
import java.io.IOException;
import java.io.OutputStream;
import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.nio.ByteBuffer;
import java.nio.channels.FileChannel;
import java.nio.file.Files;
import java.nio.file.Path;
import java.nio.file.StandardOpenOption;
import java.util.zip.Deflater;
import java.util.zip.DeflaterOutputStream;

/**
 * Drop-in replacement of this package's Encoder: same command line, same raw-grayscale input and same output file
 * (zlib of the Exp-Golomb stream), with everything between the file read and the Deflater done by libdct3d.so.
 *
 * <pre>java br.jpiccoli.video.Encoder &lt;input&gt; &lt;output&gt; &lt;width&gt; &lt;height&gt; [frames]</pre>
 *
 * Kept from the original: the argument rules (frames defaults to the file length, trailing frames that do not fill a
 * cube are dropped, Encoder.java:16-40) and the default-level Deflater (:114-125).  Gone: the double[] copies of the
 * clip, DCT.run(), the quantisation loop, CubeUtils and ExpGolombWriter (:51-111), which are one call now.  The input is
 * memory-mapped, so clips beyond 2^31 samples work.  System property dct3d.precision=32 selects the float path (the
 * C/OpenCL flavour's arithmetic, about 8x faster; a handful of +-1 coefficient differences per clip).
 */
public final class Encoder {

    private static final int CUBE = 8;

    public static void main(String[] args) throws IOException {
        if (args.length < 4) {
            System.out.println("Usage: java Encoder <input file> <output file> <frame width> <frame height> <number of frames to encode>");
            System.exit(-1);
        }
        Path in = Path.of(args[0]), out = Path.of(args[1]);
        int width = Integer.parseInt(args[2]), height = Integer.parseInt(args[3]);
        long frameBytes = (long) width * height;
        int frames = args.length > 4 ? Integer.parseInt(args[4]) : (int) (Files.size(in) / frameBytes);
        frames -= frames % CUBE;
        Dct3d.Precision precision = "32".equals(System.getProperty("dct3d.precision")) ? Dct3d.Precision.FLOAT : Dct3d.Precision.DOUBLE;

        System.out.println("Starting");
        try (Arena arena = Arena.ofConfined();
             FileChannel ch = FileChannel.open(in, StandardOpenOption.READ);
             Dct3d gpu = new Dct3d(Integer.getInteger("dct3d.device", 0), width, height, CUBE, precision)) {
            MemorySegment clip = ch.map(FileChannel.MapMode.READ_ONLY, 0, frameBytes * frames, arena);
            System.out.println("Encoding on the GPU");
            MemorySegment stream = gpu.encode(clip, frames, arena);

            System.out.println("Compressing the resulting data");
            try (OutputStream file = new BufferedOutputStream(Files.newOutputStream(out));
                 DeflaterOutputStream z = new DeflaterOutputStream(file, new Deflater(), 1 << 16)) {
                ByteBuffer bytes = stream.asByteBuffer();
                byte[] piece = new byte[1 << 20];
                while (bytes.hasRemaining()) {
                    int n = Math.min(piece.length, bytes.remaining());
                    bytes.get(piece, 0, n);
                    z.write(piece, 0, n);
                }
            }
        }
        System.out.println("Finished. Frames encoded: " + frames);
    }
}
