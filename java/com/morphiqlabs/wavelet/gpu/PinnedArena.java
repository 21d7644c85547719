package com.morphiqlabs.wavelet.gpu;

import java.lang.foreign.MemorySegment;

/**
 * A grow-only block of page-locked host memory per thread.  {@code cudaMallocHost} / {@code cudaFreeHost} cost
 * milliseconds; the engine answers a 1 x 4096 transform in ~25 us, so the staging segments of the {@code double[]} API
 * are carved from one pooled allocation instead of being allocated per call (VERDICT r1, engineering item 12).
 *
 * <p>Usage: {@code PinnedArena a = PinnedArena.current(); a.reset(); MemorySegment x = a.take(bytes); ...}.  Segments are
 * valid until the next {@code reset()} on the same thread; every engine call through the host-pointer API has finished
 * with them when it returns (the C ABI synchronises host-buffer calls).</p>
 */
public final class PinnedArena implements AutoCloseable {
    private static final ThreadLocal<PinnedArena> CURRENT = ThreadLocal.withInitial(PinnedArena::new);
    private static final long ALIGN = 256;

    private MemorySegment block = MemorySegment.NULL;
    private long capacity, used;

    public static PinnedArena current() {
        return CURRENT.get();
    }

    /** Forget every segment handed out since the last reset (the memory stays allocated). */
    public void reset() {
        used = 0;
    }

    /** {@code bytes} of pinned memory, 256-byte aligned.  Grows (re-allocates) only when the high-water mark rises; a
     *  growth invalidates earlier segments of this cycle, so callers take all their segments through {@link #takeAll}. */
    public MemorySegment[] takeAll(long... sizes) {
        long total = 0;
        for (long s : sizes) total += (s + ALIGN - 1) / ALIGN * ALIGN;
        if (used + total > capacity) {
            if (used != 0) throw new IllegalStateException("PinnedArena.takeAll after take in the same cycle");
            if (!block.equals(MemorySegment.NULL)) VwNative.freePinned(block);
            capacity = Math.max(total, 1L << 20);
            block = VwNative.allocPinned(capacity);
        }
        MemorySegment[] out = new MemorySegment[sizes.length];
        for (int i = 0; i < sizes.length; i++) {
            out[i] = block.asSlice(used, sizes[i]);
            used += (sizes[i] + ALIGN - 1) / ALIGN * ALIGN;
        }
        return out;
    }

    @Override
    public void close() {
        if (!block.equals(MemorySegment.NULL)) VwNative.freePinned(block);
        block = MemorySegment.NULL;
        capacity = used = 0;
    }
}
