package com.morphiqlabs.wavelet.modwt;

import com.morphiqlabs.wavelet.api.BoundaryMode;
import com.morphiqlabs.wavelet.api.Wavelet;
import com.morphiqlabs.wavelet.gpu.PinnedArena;
import com.morphiqlabs.wavelet.gpu.VwNative;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;

/**
 * A {@link MutableMultiLevelMODWTResult} whose coefficients stay in HBM (the C ABI's {@code vw_result}).  The reference's
 * result classes are opaque behind their interface (MultiLevelMODWTResultImpl.java:51-139,
 * MutableMultiLevelMODWTResult.java:30-123), so a device-backed implementation changes nothing for callers: a level is
 * copied to the host only when {@code getDetailCoeffsAtLevel} asks for it, thresholds and energies run where the data is,
 * and {@link #reconstruct()} moves 8 B/sample back instead of 8 * (levels + 2) each way.
 */
public final class GpuResidentResult implements MutableMultiLevelMODWTResult, AutoCloseable {
    private final GpuMODWT gpu;
    private MemorySegment handle;
    private final int n, levels;
    private final Wavelet wavelet;
    private final BoundaryMode mode;
    private final double[][] hostCopy;     // lazily fetched levels; index levels = approximation

    GpuResidentResult(GpuMODWT gpu, MemorySegment handle, int n, int levels, Wavelet wavelet, BoundaryMode mode) {
        this.gpu = gpu; this.handle = handle; this.n = n; this.levels = levels; this.wavelet = wavelet; this.mode = mode;
        this.hostCopy = new double[levels + 1][];
    }

    @Override public int getLevels() { return levels; }
    @Override public int getSignalLength() { return n; }

    private double[] fetch(int level) {     // level 0 = approximation
        int slot = level == 0 ? levels : level - 1;
        if (hostCopy[slot] == null) {
            PinnedArena pin = PinnedArena.current();
            pin.reset();
            MemorySegment dst = pin.takeAll(8L * n)[0];
            try {
                gpu.check((int) VwNative.vw_result_get_level.invokeExact(gpu.ctx, handle, level, dst, (long) n, 0));
            } catch (Throwable t) {
                throw VwNative.rethrow(t);
            }
            hostCopy[slot] = VwNative.copyOut(dst, 0, n);
        }
        return hostCopy[slot];
    }

    private void push(int level, double[] coeffs) {
        PinnedArena pin = PinnedArena.current();
        pin.reset();
        MemorySegment src = pin.takeAll(8L * n)[0];
        MemorySegment.copy(coeffs, 0, src, ValueLayout.JAVA_DOUBLE, 0, n);
        try {
            gpu.check((int) VwNative.vw_result_set_level.invokeExact(gpu.ctx, handle, level, src, (long) n, 0));
        } catch (Throwable t) {
            throw VwNative.rethrow(t);
        }
    }

    private void checkLevel(int level) {     // MultiLevelMODWTResultImpl.java:74-80
        if (level < 1 || level > levels) throw new IllegalArgumentException("Level must be between 1 and " + levels + ", got: " + level);
    }

    @Override public double[] getDetailCoeffsAtLevel(int level) { checkLevel(level); return fetch(level).clone(); }
    @Override public double[] getApproximationCoeffs() { return fetch(0).clone(); }
    /** The mutable views of the reference hand out the backing array; here they are host copies written back by
     *  {@link #setDetailCoeffs} / {@link #clearCaches()} -- callers that mutate in place call clearCaches() as the reference's
     *  own adapter does (VectorWaveSwtAdapter.java:489-500), which pushes modified levels back to the device. */
    @Override public double[] getMutableDetailCoeffs(int level) { checkLevel(level); return fetch(level); }
    @Override public double[] getMutableApproximationCoeffs() { return fetch(0); }

    @Override public void setDetailCoeffs(int level, double[] coeffs) {
        checkLevel(level);
        if (coeffs == null || coeffs.length != n) throw new IllegalArgumentException("coefficients must have length " + n);
        hostCopy[level - 1] = coeffs.clone();
        push(level, coeffs);
    }

    @Override public void setApproximationCoeffs(double[] coeffs) {
        if (coeffs == null || coeffs.length != n) throw new IllegalArgumentException("coefficients must have length " + n);
        hostCopy[levels] = coeffs.clone();
        push(0, coeffs);
    }

    @Override public void clearCaches() {
        for (int j = 1; j <= levels; j++) if (hostCopy[j - 1] != null) push(j, hostCopy[j - 1]);
        if (hostCopy[levels] != null) push(0, hostCopy[levels]);
    }

    /** MutableMultiLevelMODWTResult.applyThreshold (:83-118) where the coefficients are: no PCIe traffic. */
    @Override public void applyThreshold(int level, double threshold, boolean soft) {
        try (Arena a = Arena.ofConfined()) {
            if (level == 0) {                         // the reference thresholds the approximation for level 0
                double[] v = fetch(0);
                for (int i = 0; i < n; i++) {
                    double abs = Math.abs(v[i]);
                    v[i] = soft ? (abs > threshold ? Math.signum(v[i]) * (abs - threshold) : 0.0) : (abs <= threshold ? 0.0 : v[i]);
                }
                push(0, v);
                return;
            }
            checkLevel(level);
            MemorySegment thr = a.allocateFrom(ValueLayout.JAVA_DOUBLE, threshold);
            gpu.check((int) VwNative.vw_result_threshold.invokeExact(gpu.ctx, handle, level, thr, 0, soft ? 1 : 0));
            hostCopy[level - 1] = null;
        } catch (Throwable t) {
            throw VwNative.rethrow(t);
        }
    }

    /** VectorWaveSwtAdapter.applyUniversalThreshold (:505-520): returns the threshold it applied to every detail level. */
    public double applyUniversalThreshold(boolean soft) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment thr = a.allocate(ValueLayout.JAVA_DOUBLE);
            gpu.check((int) VwNative.vw_result_universal_threshold.invokeExact(gpu.ctx, handle, soft ? 1 : 0, thr));
            for (int j = 0; j < levels; j++) hostCopy[j] = null;
            return thr.get(ValueLayout.JAVA_DOUBLE, 0);
        } catch (Throwable t) {
            throw VwNative.rethrow(t);
        }
    }

    private double energy(int level) {
        try (Arena a = Arena.ofConfined()) {
            MemorySegment e = a.allocate(ValueLayout.JAVA_DOUBLE);
            gpu.check((int) VwNative.vw_result_energy.invokeExact(gpu.ctx, handle, level, e));
            return e.get(ValueLayout.JAVA_DOUBLE, 0);
        } catch (Throwable t) {
            throw VwNative.rethrow(t);
        }
    }

    @Override public double getDetailEnergyAtLevel(int level) { checkLevel(level); return energy(level); }
    @Override public double getApproximationEnergy() { return energy(0); }

    @Override public double getTotalEnergy() {                  // MultiLevelMODWTResultImpl.java:111-121: details, then approximation
        double total = 0.0;
        for (int j = 1; j <= levels; j++) total += energy(j);
        return total + energy(0);
    }

    @Override public double[] getRelativeEnergyDistribution() { // :124-139: [detail_1 .. detail_J, approximation] / total
        double[] e = new double[levels + 1];
        double total = 0.0;
        for (int j = 1; j <= levels; j++) { e[j - 1] = energy(j); total += e[j - 1]; }
        e[levels] = energy(0);
        total += e[levels];
        if (total > 0) for (int i = 0; i <= levels; i++) e[i] /= total;
        return e;
    }

    /** MultiLevelMODWTTransform.reconstruct / VectorWaveSwtAdapter.inverse from the resident coefficients. */
    public double[] reconstruct() {
        double[] hs = GpuMODWT.scaled(wavelet.lowPassReconstruction()), gs = GpuMODWT.scaled(wavelet.highPassReconstruction());
        PinnedArena pin = PinnedArena.current();
        pin.reset();
        MemorySegment x = pin.takeAll(8L * n)[0];
        try (Arena a = Arena.ofConfined()) {
            MemorySegment hseg = a.allocateFrom(ValueLayout.JAVA_DOUBLE, hs), gseg = a.allocateFrom(ValueLayout.JAVA_DOUBLE, gs);
            gpu.check((int) VwNative.vw_modwt_reconstruct_h.invokeExact(gpu.ctx, handle, hseg, gseg, hs.length, GpuMODWT.modeOf(mode),
                    GpuMODWT.alignmentTable(a, wavelet, mode, levels), GpuMODWT.orderOf(mode), levels >= 64 ? -1L : (1L << levels) - 1, 1,
                    x, (long) n, 0));
            return VwNative.copyOut(x, 0, n);
        } catch (Throwable t) {
            throw VwNative.rethrow(t);
        }
    }

    @Override public MultiLevelMODWTResult toImmutable() {
        MultiLevelMODWTResultImpl r = new MultiLevelMODWTResultImpl(n, levels);
        for (int j = 1; j <= levels; j++) r.setDetailCoeffsAtLevel(j, fetch(j));
        r.setApproximationCoeffs(fetch(0));
        return r;
    }

    @Override public MultiLevelMODWTResult copy() { return toImmutable(); }

    @Override public boolean isValid() {                       // finite everywhere: the engine rejected non-finite input already
        return !handle.equals(MemorySegment.NULL);
    }

    @Override public void close() {
        if (handle.equals(MemorySegment.NULL)) return;
        try {
            VwNative.vw_result_free.invokeExact(gpu.ctx, handle);
        } catch (Throwable t) {
            throw VwNative.rethrow(t);
        } finally {
            handle = MemorySegment.NULL;
        }
    }
}
