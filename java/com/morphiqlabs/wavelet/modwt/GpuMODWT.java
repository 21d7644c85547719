package com.morphiqlabs.wavelet.modwt;

import com.morphiqlabs.wavelet.api.BoundaryMode;
import com.morphiqlabs.wavelet.api.Wavelet;
import com.morphiqlabs.wavelet.exception.ErrorCode;
import com.morphiqlabs.wavelet.exception.InvalidArgumentException;
import com.morphiqlabs.wavelet.exception.InvalidSignalException;
import com.morphiqlabs.wavelet.gpu.PinnedArena;
import com.morphiqlabs.wavelet.gpu.VwNative;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;
import java.util.Objects;

/**
 * The dispatch target of the MODWT / SWT facades when a B200 is present: same arguments, same result classes, same
 * exceptions as the scalar code it stands in for.
 *
 * <p>It lives in {@code com.morphiqlabs.wavelet.modwt} on purpose: the level cap, the SYMMETRIC alignment table and the
 * result containers are the reference's own package-private classes ({@code MultiLevelMODWTResultImpl},
 * {@code SymmetricAlignmentStrategy}), so their quirks stay theirs -- nothing of that policy is restated in native code.
 * Only {@code computeTauJ} (private static in MultiLevelMODWTTransform.java:795-806) is repeated here; making it
 * package-private there removes the copy.</p>
 *
 * <p>What each method replaces (vectorwave-core/src/main/java/com/morphiqlabs/wavelet):
 * <ul>
 *   <li>{@link #decompose} / {@link #decomposeMutable}: MultiLevelMODWTTransform.decompose / decomposeMutable
 *       (modwt/MultiLevelMODWTTransform.java:209-255,284-330) and VectorWaveSwtAdapter.forward (swt/VectorWaveSwtAdapter.java:198-204)</li>
 *   <li>{@link #reconstruct}, {@link #reconstructFromLevel}, {@link #reconstructLevels}: :339-349, :361-386, :398-446;
 *       VectorWaveSwtAdapter.inverse (:435-474)</li>
 *   <li>{@link #denoise}: VectorWaveSwtAdapter.denoise (:532-562) incl. the universal threshold (:505-520,627-645)</li>
 *   <li>{@link #decomposeResident}: the same decomposition kept in HBM (see {@link GpuResidentResult})</li>
 * </ul>
 * The two-line dispatch a facade needs, e.g. at the top of MultiLevelMODWTTransform.decompose(double[], int):
 * <pre>
 *   GpuMODWT gpu = GpuMODWT.ifAvailable();
 *   if (gpu != null) return gpu.decompose(signal, levels, wavelet, boundaryMode, calculateMaxLevels(signal.length));
 * </pre>
 * Source only: the build image has no JDK ({@code javac --release 22} where one exists).  The C++ twin of this class,
 * include/vw_modwt.hpp, IS compiled and checked against the oracle (tests/test_cpp_host.py).</p>
 */
public final class GpuMODWT implements AutoCloseable {
    private static final ThreadLocal<GpuMODWT> CURRENT = new ThreadLocal<>();
    private static volatile boolean unavailable = "off".equalsIgnoreCase(System.getProperty("vectorwave.gpu", "on"));

    final MemorySegment ctx;

    private GpuMODWT(MemorySegment ctx) {
        this.ctx = ctx;
    }

    /** One engine context per calling thread (the ctx is the only mutable state; transforms stay immutable and shareable).
     *  {@code null} when {@code -Dvectorwave.gpu=off}, the library is missing or no CUDA device exists: the caller keeps its
     *  scalar path -- the engine itself has no CPU fallback. */
    public static GpuMODWT ifAvailable() {
        if (unavailable) return null;
        GpuMODWT g = CURRENT.get();
        if (g != null) return g;
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(ValueLayout.ADDRESS);
            int rc = (int) VwNative.vw_init.invokeExact(Integer.getInteger("vectorwave.gpu.device", -1), out);
            if (rc != VwNative.VW_OK) {
                unavailable = true;
                return null;
            }
            g = new GpuMODWT(out.get(ValueLayout.ADDRESS, 0));
            CURRENT.set(g);
            return g;
        } catch (Throwable t) {       // UnsatisfiedLinkError included
            unavailable = true;
            return null;
        }
    }

    // ---- host policy: the reference's own tables ------------------------------------------------------------------
    static int modeOf(BoundaryMode m) {
        return switch (m) {
            case PERIODIC -> 0;
            case ZERO_PADDING -> 1;
            case SYMMETRIC -> 2;
            default -> 3;      // rejected by the engine with CFG_UNSUPPORTED_BOUNDARY_MODE, as MODWTTransform.java:96-110 does
        };
    }

    static double[] scaled(double[] f) {
        double s = 1.0 / Math.sqrt(2.0);                    // ScalarOps.java:909-916: one rounding per tap
        double[] out = new double[f.length];
        for (int i = 0; i < f.length; i++) out[i] = f[i] * s;
        return out;
    }

    /** MultiLevelMODWTTransform.computeTauJ (:795-806), verbatim semantics. */
    static int computeTauJ(int baseFilterLength, int level) {
        int lm1 = baseFilterLength - 1;
        if (level <= 1) return Math.max(0, lm1 / 2);
        long tau = ((long) lm1 * (1L << (level - 1))) / 2L;
        return tau > Integer.MAX_VALUE ? Integer.MAX_VALUE : (int) tau;
    }

    /** Per-level (sigma_h, tau_h, sigma_g, tau_g) of the multi-level SYMMETRIC synthesis (MultiLevelMODWTTransform.java:602-642)
     *  from SymmetricAlignmentStrategy.decide -- the reference class itself; NULL for the other modes (t + k*d at every level). */
    static MemorySegment alignmentTable(Arena a, Wavelet wavelet, BoundaryMode mode, int levels) {
        if (mode != BoundaryMode.SYMMETRIC) return MemorySegment.NULL;
        int lh = wavelet.lowPassReconstruction().length, lg = wavelet.highPassReconstruction().length;
        MemorySegment t = a.allocate(ValueLayout.JAVA_INT, 4L * levels);
        for (int j = 1; j <= levels; j++) {
            SymmetricAlignmentStrategy.Decision d = SymmetricAlignmentStrategy.decide(wavelet, j);
            t.setAtIndex(ValueLayout.JAVA_INT, 4L * (j - 1), d.approxPlus ? 1 : -1);
            t.setAtIndex(ValueLayout.JAVA_INT, 4L * (j - 1) + 1, computeTauJ(lh, j) + d.deltaApprox);
            t.setAtIndex(ValueLayout.JAVA_INT, 4L * (j - 1) + 2, d.detailPlus ? 1 : -1);
            t.setAtIndex(ValueLayout.JAVA_INT, 4L * (j - 1) + 3, computeTauJ(lg, j) + d.deltaDetail);
        }
        return t;
    }

    /** PERIODIC and SYMMETRIC sum all H taps, then all G taps (:578-589,602-642); ZERO_PADDING adds the pair per tap (:591-601). */
    static int orderOf(BoundaryMode mode) {
        return mode == BoundaryMode.ZERO_PADDING ? VwNative.ORDER_PAIR : VwNative.ORDER_SPLIT;
    }

    /** vw_status -> the reference's exceptions (exception/ErrorCode.java:24-118); same table as vectorwave_b200/_native.py. */
    void check(int rc) {
        if (rc == VwNative.VW_OK) return;
        String msg = VwNative.lastError(ctx);
        switch (rc) {
            case 1 -> throw new NullPointerException(msg);
            case 3 -> throw new InvalidSignalException(ErrorCode.VAL_NON_FINITE_VALUES, msg);
            case 5 -> throw new InvalidArgumentException(ErrorCode.VAL_TOO_LARGE, msg);
            case 6 -> throw new InvalidSignalException(ErrorCode.VAL_EMPTY, msg);
            case 103 -> throw new InvalidArgumentException(ErrorCode.CFG_UNSUPPORTED_BOUNDARY_MODE, msg);
            case 104 -> throw new InvalidArgumentException(ErrorCode.CFG_INVALID_DECOMPOSITION_LEVEL, msg);
            case 7, 400 -> throw new IllegalArgumentException(msg);
            default -> throw new IllegalStateException("libvwmodwt: " + rc + ": " + msg);
        }
    }

    private static void checkLevels(int levels, int maxLevels) {
        if (levels < 1 || levels > maxLevels)                              // MultiLevelMODWTTransform.java:226-239
            throw new InvalidArgumentException(ErrorCode.CFG_INVALID_DECOMPOSITION_LEVEL,
                    "Invalid number of decomposition levels: " + levels + " (maximum " + maxLevels + ")");
    }

    // ---- analysis ---------------------------------------------------------------------------------------------------
    /** @param maxLevels the caller's {@code calculateMaxLevels(signal.length)} (:455-501, cap 9): policy stays in the facade */
    public MultiLevelMODWTResult decompose(double[] signal, int levels, Wavelet wavelet, BoundaryMode mode, int maxLevels) {
        Objects.requireNonNull(signal, "signal cannot be null");
        checkLevels(levels, maxLevels);
        int n = signal.length;
        MultiLevelMODWTResultImpl result = new MultiLevelMODWTResultImpl(n, levels);
        double[][] out = forwardToArrays(signal, levels, wavelet, mode);
        for (int j = 1; j <= levels; j++) result.setDetailCoeffsAtLevel(j, out[j - 1]);
        result.setApproximationCoeffs(out[levels]);
        return result;
    }

    public MutableMultiLevelMODWTResult decomposeMutable(double[] signal, int levels, Wavelet wavelet, BoundaryMode mode, int maxLevels) {
        Objects.requireNonNull(signal, "signal cannot be null");
        checkLevels(levels, maxLevels);
        MutableMultiLevelMODWTResultImpl result = new MutableMultiLevelMODWTResultImpl(signal.length, levels);
        double[][] out = forwardToArrays(signal, levels, wavelet, mode);
        for (int j = 1; j <= levels; j++) result.setDetailCoeffs(j, out[j - 1]);
        result.setApproximationCoeffs(out[levels]);
        return result;
    }

    /** rows 0..levels-1: W_1..W_J; row levels: V_J */
    private double[][] forwardToArrays(double[] signal, int levels, Wavelet wavelet, BoundaryMode mode) {
        int n = signal.length;
        if (n == 0) throw new InvalidSignalException(ErrorCode.VAL_EMPTY, "Signal cannot be empty for multi-level MODWT");
        double[] hs = scaled(wavelet.lowPassDecomposition()), gs = scaled(wavelet.highPassDecomposition());
        PinnedArena pin = PinnedArena.current();
        pin.reset();
        MemorySegment[] seg = pin.takeAll(8L * n, 8L * n * levels, 8L * n);
        MemorySegment x = seg[0], w = seg[1], v = seg[2];
        try (Arena a = Arena.ofConfined()) {
            MemorySegment.copy(signal, 0, x, ValueLayout.JAVA_DOUBLE, 0, n);
            MemorySegment hseg = a.allocateFrom(ValueLayout.JAVA_DOUBLE, hs), gseg = a.allocateFrom(ValueLayout.JAVA_DOUBLE, gs);
            check((int) VwNative.vw_modwt_forward.invokeExact(ctx, x, 1L, (long) n, (long) n, hseg, gseg, hs.length, levels,
                    modeOf(mode), w, (long) n, (long) n, v, (long) n, VwNative.FLAG_CHECK_FINITE));   // validateFiniteValues on the device
            double[][] out = new double[levels + 1][];
            for (int j = 0; j < levels; j++) out[j] = VwNative.copyOut(w, (long) j * n, n);
            out[levels] = VwNative.copyOut(v, 0, n);
            return out;
        } catch (Throwable t) {
            throw VwNative.rethrow(t);
        }
    }

    // ---- synthesis --------------------------------------------------------------------------------------------------
    public double[] reconstruct(MultiLevelMODWTResult result, Wavelet wavelet, BoundaryMode mode) {
        Objects.requireNonNull(result, "result cannot be null");
        int j = result.getLevels();
        return inverse(result, wavelet, mode, j >= 64 ? -1L : (1L << j) - 1, true);
    }

    /** details finer than startLevel are zeros (MultiLevelMODWTTransform.java:361-386) */
    public double[] reconstructFromLevel(MultiLevelMODWTResult result, int startLevel, Wavelet wavelet, BoundaryMode mode) {
        Objects.requireNonNull(result, "result cannot be null");
        if (startLevel < 1 || startLevel > result.getLevels())
            throw new InvalidArgumentException("Invalid start level: " + startLevel + ". Must be between 1 and " + result.getLevels());
        long mask = ((1L << result.getLevels()) - 1) & ~((1L << (startLevel - 1)) - 1);
        return inverse(result, wavelet, mode, mask, true);
    }

    /** details outside [minLevel, maxLevel] are zeros; the approximation only when the coarsest level is included (:398-446) */
    public double[] reconstructLevels(MultiLevelMODWTResult result, int minLevel, int maxLevel, Wavelet wavelet, BoundaryMode mode) {
        Objects.requireNonNull(result, "result cannot be null");
        int j = result.getLevels();
        if (minLevel < 1 || maxLevel > j || minLevel > maxLevel)
            throw new InvalidArgumentException(ErrorCode.CFG_INVALID_DECOMPOSITION_LEVEL,
                    "Invalid level range for partial reconstruction: [" + minLevel + ", " + maxLevel + "] of " + j);
        long mask = ((1L << maxLevel) - 1) & ~((1L << (minLevel - 1)) - 1);
        return inverse(result, wavelet, mode, mask, maxLevel == j);
    }

    private double[] inverse(MultiLevelMODWTResult result, Wavelet wavelet, BoundaryMode mode, long detailMask, boolean useApprox) {
        int levels = result.getLevels(), n = result.getSignalLength();
        double[] hs = scaled(wavelet.lowPassReconstruction()), gs = scaled(wavelet.highPassReconstruction());
        PinnedArena pin = PinnedArena.current();
        pin.reset();
        MemorySegment[] seg = pin.takeAll(8L * n * levels, 8L * n, 8L * n);
        MemorySegment w = seg[0], v = seg[1], x = seg[2];
        try (Arena a = Arena.ofConfined()) {
            for (int j = 1; j <= levels; j++)
                MemorySegment.copy(result.getDetailCoeffsAtLevel(j), 0, w, ValueLayout.JAVA_DOUBLE, 8L * (j - 1) * n, n);
            MemorySegment.copy(result.getApproximationCoeffs(), 0, v, ValueLayout.JAVA_DOUBLE, 0, n);
            MemorySegment hseg = a.allocateFrom(ValueLayout.JAVA_DOUBLE, hs), gseg = a.allocateFrom(ValueLayout.JAVA_DOUBLE, gs);
            check((int) VwNative.vw_modwt_inverse.invokeExact(ctx, w, (long) n, (long) n, v, (long) n, 1L, (long) n, hseg, gseg,
                    hs.length, levels, modeOf(mode), alignmentTable(a, wavelet, mode, levels), orderOf(mode), detailMask,
                    useApprox ? 1 : 0, x, (long) n, 0));
            return VwNative.copyOut(x, 0, n);
        } catch (Throwable t) {
            throw VwNative.rethrow(t);
        }
    }

    // ---- SWT adapter --------------------------------------------------------------------------------------------------
    /** VectorWaveSwtAdapter.denoise(signal, levels) (:532-536): universal threshold, soft. */
    public double[] denoise(double[] signal, int levels, Wavelet wavelet, BoundaryMode mode, int maxLevels) {
        return denoise(signal, levels, -1.0, true, wavelet, mode, maxLevels);
    }

    /** VectorWaveSwtAdapter.denoise(signal, levels, threshold, soft) (:546-562); threshold < 0 selects the universal one
     *  (median(|W_1|)/0.6745 * sqrt(2 ln n), :505-520,627-645).  One call: 8 B/sample each way over PCIe. */
    public double[] denoise(double[] signal, int levels, double threshold, boolean soft, Wavelet wavelet, BoundaryMode mode, int maxLevels) {
        Objects.requireNonNull(signal, "signal cannot be null");
        checkLevels(levels, maxLevels);
        int n = signal.length;
        double[] hs = scaled(wavelet.lowPassDecomposition()), gs = scaled(wavelet.highPassDecomposition());
        PinnedArena pin = PinnedArena.current();
        pin.reset();
        MemorySegment[] seg = pin.takeAll(8L * n, 8L * n);
        try (Arena a = Arena.ofConfined()) {
            MemorySegment.copy(signal, 0, seg[0], ValueLayout.JAVA_DOUBLE, 0, n);
            MemorySegment hseg = a.allocateFrom(ValueLayout.JAVA_DOUBLE, hs), gseg = a.allocateFrom(ValueLayout.JAVA_DOUBLE, gs);
            check((int) VwNative.vw_swt_denoise.invokeExact(ctx, seg[0], 1L, (long) n, (long) n, hseg, gseg, hs.length, levels,
                    modeOf(mode), alignmentTable(a, wavelet, mode, levels), orderOf(mode), threshold, soft ? 1 : 0, seg[1], (long) n,
                    MemorySegment.NULL, VwNative.FLAG_CHECK_FINITE));
            return VwNative.copyOut(seg[1], 0, n);
        } catch (Throwable t) {
            throw VwNative.rethrow(t);
        }
    }

    // ---- device-resident results --------------------------------------------------------------------------------------
    /** decompose with the coefficients kept in HBM: threshold / energy / reconstruct without crossing PCIe again. */
    public GpuResidentResult decomposeResident(double[] signal, int levels, Wavelet wavelet, BoundaryMode mode, int maxLevels) {
        Objects.requireNonNull(signal, "signal cannot be null");
        checkLevels(levels, maxLevels);
        int n = signal.length;
        double[] hs = scaled(wavelet.lowPassDecomposition()), gs = scaled(wavelet.highPassDecomposition());
        PinnedArena pin = PinnedArena.current();
        pin.reset();
        MemorySegment x = pin.takeAll(8L * n)[0];
        try (Arena a = Arena.ofConfined()) {
            MemorySegment.copy(signal, 0, x, ValueLayout.JAVA_DOUBLE, 0, n);
            MemorySegment hseg = a.allocateFrom(ValueLayout.JAVA_DOUBLE, hs), gseg = a.allocateFrom(ValueLayout.JAVA_DOUBLE, gs);
            MemorySegment handle = a.allocate(ValueLayout.ADDRESS);
            handle.set(ValueLayout.ADDRESS, 0, MemorySegment.NULL);
            check((int) VwNative.vw_modwt_decompose_h.invokeExact(ctx, x, 1L, (long) n, (long) n, hseg, gseg, hs.length, levels,
                    modeOf(mode), handle, VwNative.FLAG_CHECK_FINITE));
            return new GpuResidentResult(this, handle.get(ValueLayout.ADDRESS, 0), n, levels, wavelet, mode);
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
