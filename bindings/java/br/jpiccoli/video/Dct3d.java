package br.jpiccoli.video;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;

/**
 * java.lang.foreign (JDK 22+) binding of libdct3d.so, the CUDA implementation of this codec's hot path
 * (include/dct3d.h).  No JNI glue: every method is one downcall into the C ABI.
 *
 * <p>What each method stands in for in the original sources of this package:
 * <ul>
 * <li>{@link #encode} -- Encoder.java:51-111: pixel conversion, {@code new DCT(...).run()}, the quantisation loop,
 *     CubeUtils.diagonalSlices and the ExpGolombWriter loop.  Returns the Exp-Golomb bytes, {@code bufferPosition + 1}
 *     of them, ready for the Deflater of Encoder.java:114-125.</li>
 * <li>{@link #decode} -- Decoder.java:61-117: the ExpGolombReader loop, de-quantisation,
 *     {@code new InverseDCT(...).run()} and the byte conversion.</li>
 * <li>{@link #forward} / {@link #inverse} -- {@code new DCT(in, out, w, h, c, c, c).run()} and
 *     {@code new InverseDCT(...).run()} alone (dct/Transform.java:44-104), double[] in and out.</li>
 * </ul>
 *
 * <p>Run with {@code java --enable-native-access=ALL-UNNAMED -Ddct3d.lib=/path/to/libdct3d.so ...}.  The library has
 * no CPU fallback: without a CUDA device the constructor throws.  One instance per thread.
 *
 * <p>This file is compiled by nobody in the build image (it has no JDK); it is the binding a maintainer adds, kept as a
 * source file so that it can be compiled as it is (INTEGRATION.md section 2).
 */
public final class Dct3d implements AutoCloseable {

    /** Arithmetic of the fused calls: FLOAT is the C/OpenCL flavour's (and the fast path), DOUBLE the Java flavour's. */
    public enum Precision { FLOAT, DOUBLE }

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB =
            SymbolLookup.libraryLookup(System.getProperty("dct3d.lib", "libdct3d.so"), Arena.global());

    private static MethodHandle fn(String name, FunctionDescriptor d) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), d);
    }

    private static final MethodHandle CREATE = fn("dct3d_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT));
    private static final MethodHandle DESTROY = fn("dct3d_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle LAST_ERROR = fn("dct3d_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    private static final MethodHandle SET_OPTION = fn("dct3d_set_option", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG));
    private static final MethodHandle HOST_ALLOC = fn("dct3d_host_alloc", FunctionDescriptor.of(ADDRESS, JAVA_LONG));
    private static final MethodHandle HOST_FREE = fn("dct3d_host_free", FunctionDescriptor.ofVoid(ADDRESS));
    private static final MethodHandle ENCODE = fn("dct3d_encode_u8", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS));
    private static final MethodHandle DECODE = fn("dct3d_decode_u8", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_INT, ADDRESS));
    private static final MethodHandle FORWARD_F64 = fn("dct3d_forward_f64", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
    private static final MethodHandle INVERSE_F64 = fn("dct3d_inverse_f64", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
    private static final MethodHandle DEVICE_COUNT = fn("dct3d_device_count", FunctionDescriptor.of(JAVA_INT));

    private final MemorySegment ctx;
    private final int width, height, cube;

    public Dct3d(int device, int width, int height, int cube, Precision precision) {
        this.width = width;
        this.height = height;
        this.cube = cube;
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(ADDRESS);
            check((int) CREATE.invokeExact(out, device, width, height, cube), MemorySegment.NULL);
            ctx = out.get(ADDRESS, 0);
            // Encoder.java:82 rounds with Math.round on doubles: precision 64 + rounding 0 reproduces it without
            // rounding-tie differences; the float path differs by a handful of +-1 coefficients per clip
            option(a, "precision", precision == Precision.DOUBLE ? 64 : 32);
            option(a, "rounding", 0);
        } catch (RuntimeException | Error e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    public static int deviceCount() {
        try {
            return (int) DEVICE_COUNT.invokeExact();
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    private void option(Arena a, String key, long value) throws Throwable {
        check((int) SET_OPTION.invokeExact(ctx, a.allocateFrom(key), value), ctx);
    }

    private static void check(int rc, MemorySegment c) throws Throwable {
        if (rc == 0) return;
        MemorySegment msg = (MemorySegment) LAST_ERROR.invokeExact(c);
        throw new IllegalStateException("dct3d error " + rc + ": " + msg.reinterpret(512).getString(0));
    }

    /** Page-locked staging memory (cudaHostAlloc): copies to the GPU run at PCIe speed from it. */
    private static MemorySegment pinned(long bytes, Arena owner) throws Throwable {
        MemorySegment p = (MemorySegment) HOST_ALLOC.invokeExact(bytes);
        if (p.equals(MemorySegment.NULL)) throw new OutOfMemoryError("dct3d_host_alloc(" + bytes + ")");
        return p.reinterpret(bytes, owner, seg -> {
            try {
                HOST_FREE.invokeExact(seg);
            } catch (Throwable t) {
                throw new IllegalStateException(t);
            }
        });
    }

    /**
     * Raw frames (frame-major unsigned bytes, Encoder.java:47-56) to the Exp-Golomb stream.  {@code frames} may be any
     * segment, e.g. a memory-mapped input file of more than 2^31 samples (the original reads into one byte[]).
     */
    public MemorySegment encode(MemorySegment frames, int nframes, Arena owner) {
        try (Arena a = Arena.ofConfined()) {
            long samples = (long) width * height * (nframes - nframes % cube);
            long cap = samples / 2 + 4096;
            for (;;) {
                MemorySegment out = pinned(cap, owner);
                MemorySegment nbits = a.allocate(JAVA_LONG), nbytes = a.allocate(JAVA_LONG);
                int rc = (int) ENCODE.invokeExact(ctx, frames, nframes, out, cap, nbits, nbytes);
                if (rc == -3 && cap < 4 * samples + 4096) {        // DCT3D_E_OVERFLOW: very dense content, try again
                    cap = Math.min(cap * 4, 4 * samples + 4096);
                    continue;
                }
                check(rc, ctx);
                return out.asSlice(0, nbytes.get(JAVA_LONG, 0));
            }
        } catch (RuntimeException | Error e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    /** byte[] convenience form of {@link #encode(MemorySegment, int, Arena)}. */
    public byte[] encode(byte[] frames, int nframes) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment in = pinned(Math.max(1, frames.length), a);
            MemorySegment.copy(frames, 0, in, JAVA_BYTE, 0, frames.length);
            return encode(in, nframes, a).toArray(JAVA_BYTE);
        } catch (RuntimeException | Error e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    /** Exp-Golomb stream (inflated, Decoder.java:41-59) to raw frames. */
    public MemorySegment decode(MemorySegment stream, int nframes, Arena owner) {
        try {
            MemorySegment out = pinned(Math.max(1, (long) width * height * (nframes - nframes % cube)), owner);
            check((int) DECODE.invokeExact(ctx, stream, stream.byteSize(), nframes, out), ctx);
            return out;
        } catch (RuntimeException | Error e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    public byte[] decode(byte[] stream, int nframes) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment in = pinned(Math.max(1, stream.length), a);
            MemorySegment.copy(stream, 0, in, JAVA_BYTE, 0, stream.length);
            return decode(in.asSlice(0, stream.length), nframes, a).toArray(JAVA_BYTE);
        } catch (RuntimeException | Error e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    /** {@code new DCT(pixels, coeff, width, height, cube, cube, cube).run()}: fp64, planar in and out. */
    public void forward(double[] pixels, double[] coeff, int nframes) {
        transform(FORWARD_F64, pixels, coeff, nframes);
    }

    /** {@code new InverseDCT(coeff, pixels, ...).run()}: clamps to [0, 255] like InverseDCT.java:74-80. */
    public void inverse(double[] coeff, double[] pixels, int nframes) {
        transform(INVERSE_F64, coeff, pixels, nframes);
    }

    private void transform(MethodHandle call, double[] in, double[] out, int nframes) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment src = a.allocateFrom(JAVA_DOUBLE, in), dst = a.allocate(JAVA_DOUBLE, out.length);
            check((int) call.invokeExact(ctx, src, dst, nframes), ctx);
            MemorySegment.copy(dst, JAVA_DOUBLE, 0, out, 0, out.length);
        } catch (RuntimeException | Error e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }

    @Override
    public void close() {
        try {
            DESTROY.invokeExact(ctx);
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
    }
}
