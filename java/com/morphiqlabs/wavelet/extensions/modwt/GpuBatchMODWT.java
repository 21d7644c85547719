package com.morphiqlabs.wavelet.extensions.modwt;

import com.morphiqlabs.wavelet.api.DiscreteWavelet;
import com.morphiqlabs.wavelet.gpu.PinnedArena;
import com.morphiqlabs.wavelet.gpu.VwNative;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;

/**
 * The batch facade on the B200 engine: what {@link BatchMODWT#multiLevelAoS} and {@link BatchMODWT#inverseMultiLevelAoS}
 * (extensions/modwt/BatchMODWT.java:90-111,151-178) dispatch to -- the benchmark path of BASELINE config #2 -- plus the
 * single-level pair (:62-76,122-139).  Same AoS arguments, same result records, PERIODIC like the facade, and the facade's
 * own validation ({@code validateAoS}, :201-212) stays where it is; the engine additionally refuses {@code L_J > N}
 * (SURVEY D10: the reference indexes out of bounds there).
 *
 * <pre>
 *   // BatchMODWT.multiLevelAoS, first lines:
 *   validateAoS(signals);
 *   GpuBatchMODWT gpu = GpuBatchMODWT.ifAvailable();
 *   if (gpu != null) return gpu.multiLevelAoS(wavelet, signals, levels);
 * </pre>
 * One pooled pinned block per thread stages the batch ({@link PinnedArena}); large batches are pipelined over two streams
 * inside the library (H2D of chunk c+1, kernels of chunk c, D2H of chunk c-1 overlap).
 */
public final class GpuBatchMODWT implements AutoCloseable {
    private static final ThreadLocal<GpuBatchMODWT> CURRENT = new ThreadLocal<>();
    private static volatile boolean unavailable = "off".equalsIgnoreCase(System.getProperty("vectorwave.gpu", "on"));
    private final MemorySegment ctx;

    private GpuBatchMODWT(MemorySegment ctx) { this.ctx = ctx; }

    public static GpuBatchMODWT ifAvailable() {
        if (unavailable) return null;
        GpuBatchMODWT g = CURRENT.get();
        if (g != null) return g;
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(ValueLayout.ADDRESS);
            int rc = (int) VwNative.vw_init.invokeExact(Integer.getInteger("vectorwave.gpu.device", -1), out);
            if (rc != VwNative.VW_OK) { unavailable = true; return null; }
            g = new GpuBatchMODWT(out.get(ValueLayout.ADDRESS, 0));
            CURRENT.set(g);
            return g;
        } catch (Throwable t) {
            unavailable = true;
            return null;
        }
    }

    private static double[] scaled(double[] f, boolean haarSingleLevel) {
        // BatchSIMDMODWT.java:90-93: the single-level Haar kernel uses the literal +-0.5 instead of (1/sqrt2)*(1/sqrt2)
        double[] out = new double[f.length];
        double s = 1.0 / Math.sqrt(2.0);
        for (int i = 0; i < f.length; i++) out[i] = haarSingleLevel ? Math.copySign(0.5, f[i]) : f[i] * s;
        return out;
    }

    private void check(int rc) {
        if (rc == VwNative.VW_OK) return;
        String msg = VwNative.lastError(ctx);
        if (rc == 1) throw new NullPointerException(msg);
        if (rc == 5 || rc == 7 || rc == 104 || rc == 400) throw new IllegalArgumentException(msg);   // the facade throws IAE for shape errors (:201-212)
        throw new IllegalStateException("libvwmodwt: " + rc + ": " + msg);
    }

    /** BatchMODWT.multiLevelAoS (:90-111): details [levels][batch][length], final approximation [batch][length]. */
    public BatchMODWT.MultiLevelResult multiLevelAoS(DiscreteWavelet wavelet, double[][] signals, int levels) {
        if (levels < 1) throw new IllegalArgumentException("levels must be >= 1");
        int b = signals.length, n = signals[0].length;
        double[][][] detail = new double[levels][b][];
        double[][] approx = new double[b][];
        forward(wavelet, signals, levels, false, detail, approx);
        return new BatchMODWT.MultiLevelResult(detail, approx);
    }

    /** BatchMODWT.singleLevelAoS (:62-76). */
    public BatchMODWT.SingleLevelResult singleLevelAoS(DiscreteWavelet wavelet, double[][] signals) {
        int b = signals.length;
        double[][][] detail = new double[1][b][];
        double[][] approx = new double[b][];
        forward(wavelet, signals, 1, wavelet.lowPassDecomposition().length == 2, detail, approx);
        return new BatchMODWT.SingleLevelResult(approx, detail[0]);
    }

    private void forward(DiscreteWavelet wavelet, double[][] signals, int levels, boolean haarSingle, double[][][] detail, double[][] approx) {
        int b = signals.length, n = signals[0].length;
        double[] hs = scaled(wavelet.lowPassDecomposition(), haarSingle), gs = scaled(wavelet.highPassDecomposition(), haarSingle);
        long bn = (long) b * n;
        PinnedArena pin = PinnedArena.current();
        pin.reset();
        MemorySegment[] seg = pin.takeAll(8L * bn, 8L * bn * levels, 8L * bn);
        try (Arena a = Arena.ofConfined()) {
            for (int i = 0; i < b; i++) MemorySegment.copy(signals[i], 0, seg[0], ValueLayout.JAVA_DOUBLE, 8L * i * n, n);
            MemorySegment hseg = a.allocateFrom(ValueLayout.JAVA_DOUBLE, hs), gseg = a.allocateFrom(ValueLayout.JAVA_DOUBLE, gs);
            check((int) VwNative.vw_modwt_forward.invokeExact(ctx, seg[0], (long) b, (long) n, (long) n, hseg, gseg, hs.length, levels,
                    0 /* PERIODIC */, seg[1], (long) n, bn, seg[2], (long) n, 0));
            for (int j = 0; j < levels; j++)
                for (int i = 0; i < b; i++) detail[j][i] = VwNative.copyOut(seg[1], j * bn + (long) i * n, n);
            for (int i = 0; i < b; i++) approx[i] = VwNative.copyOut(seg[2], (long) i * n, n);
        } catch (Throwable t) {
            throw VwNative.rethrow(t);
        }
    }

    /** BatchMODWT.inverseMultiLevelAoS (:151-178): PERIODIC cascade, all H taps then all G taps per level. */
    public double[][] inverseMultiLevelAoS(DiscreteWavelet wavelet, double[][][] detailPerLevel, double[][] finalApprox) {
        int levels = detailPerLevel.length, b = finalApprox.length, n = finalApprox[0].length;
        return inverse(wavelet, detailPerLevel, finalApprox, levels, b, n, VwNative.ORDER_SPLIT, false);
    }

    /** BatchMODWT.inverseSingleLevelAoS (:122-139): pair-added products like MODWTTransform.inverse. */
    public double[][] inverseSingleLevelAoS(DiscreteWavelet wavelet, double[][] approx, double[][] detail) {
        return inverse(wavelet, new double[][][]{detail}, approx, 1, approx.length, approx[0].length, VwNative.ORDER_PAIR,
                wavelet.lowPassReconstruction().length == 2);
    }

    private double[][] inverse(DiscreteWavelet wavelet, double[][][] detail, double[][] approx, int levels, int b, int n, int order,
                               boolean haarSingle) {
        double[] hs = scaled(wavelet.lowPassReconstruction(), haarSingle), gs = scaled(wavelet.highPassReconstruction(), haarSingle);
        long bn = (long) b * n;
        PinnedArena pin = PinnedArena.current();
        pin.reset();
        MemorySegment[] seg = pin.takeAll(8L * bn * levels, 8L * bn, 8L * bn);
        try (Arena a = Arena.ofConfined()) {
            for (int j = 0; j < levels; j++)
                for (int i = 0; i < b; i++) MemorySegment.copy(detail[j][i], 0, seg[0], ValueLayout.JAVA_DOUBLE, 8L * (j * bn + (long) i * n), n);
            for (int i = 0; i < b; i++) MemorySegment.copy(approx[i], 0, seg[1], ValueLayout.JAVA_DOUBLE, 8L * i * n, n);
            MemorySegment hseg = a.allocateFrom(ValueLayout.JAVA_DOUBLE, hs), gseg = a.allocateFrom(ValueLayout.JAVA_DOUBLE, gs);
            check((int) VwNative.vw_modwt_inverse.invokeExact(ctx, seg[0], (long) n, bn, seg[1], (long) n, (long) b, (long) n, hseg, gseg,
                    hs.length, levels, 0 /* PERIODIC */, MemorySegment.NULL, order, levels >= 64 ? -1L : (1L << levels) - 1, 1, seg[2],
                    (long) n, 0));
            double[][] out = new double[b][];
            for (int i = 0; i < b; i++) out[i] = VwNative.copyOut(seg[2], (long) i * n, n);
            return out;
        } catch (Throwable t) {
            throw VwNative.rethrow(t);
        }
    }

    @Override
    public void close() {
        try {
            VwNative.vw_destroy.invokeExact(ctx);
        } catch (Throwable t) {
            throw VwNative.rethrow(t);
        } finally {
            CURRENT.remove();
        }
    }
}
