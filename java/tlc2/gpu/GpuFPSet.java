/*
 * Drop-in for tlc2.tool.fp.FPSet backed by the GPU-resident fingerprint set (kmc_fpset_*).
 * For callers that keep TLC's own Worker loop and only replace the set; the full replacement of
 * the hot loop is tlc2.gpu.GpuModelChecker.  NOT COMPILED HERE (no JDK / tla2tools.jar).
 *
 * TLC seam (tlc2.tool.fp.FPSet, public API of tla2tools.jar):
 *   boolean put(long fp)       -> true iff fp was already in the set
 *   boolean contains(long fp)
 *   long    size()
 * Worker threads call put() one fingerprint at a time; a GPU round trip per call would be
 * latency-bound, so put() is batched per worker: fingerprints are queued and flushed through
 * putBlock(), which is what a Worker processing a StateVec of successors would call.
 */
package tlc2.gpu;

public class GpuFPSet /* extends tlc2.tool.fp.FPSet */ {
    private final long ctx;

    public GpuFPSet(String modelLibrary, int fpBits) {
        this.ctx = Native.create(modelLibrary, "{\"table_log2\": " + fpBits + "}");
    }

    /** FPSet.put for a block of successor fingerprints (one kernel launch). */
    public boolean[] putBlock(long[] fps) {
        return Native.fpsetPut(ctx, fps);
    }

    public boolean put(long fp) {
        return Native.fpsetPut(ctx, new long[] {fp})[0];
    }

    public boolean contains(long fp) {
        return Native.fpsetContains(ctx, new long[] {fp})[0];
    }

    public long size() {
        return Native.fpsetSize(ctx);
    }

    public void close() {
        Native.destroy(ctx);
    }
}
