/*
 * Thin Java driver with tlc2.TLC's argument surface: replaces ModelChecker.doInit/runTLC (the
 * Worker next-state loop, FPSet and StateQueue) by one kmc_run on the GPU.  NOT COMPILED HERE.
 *
 *   java -Djava.library.path=build -cp java tlc2.gpu.GpuModelChecker -config Kip320.cfg -deadlock Kip320
 *
 * The .tla/.cfg pair is lowered ahead of time by `python -m kafka_specification_b200.build`
 * (or by the tlc2-compatible CLI, which does both steps: python -m kafka_specification_b200.tlc2).
 */
package tlc2.gpu;

public final class GpuModelChecker {
    public static void main(String[] args) {
        String config = null, spec = null, modelLib = System.getProperty("kspec.model");
        boolean noDeadlock = false, cont = false;
        for (int i = 0; i < args.length; i++) {
            switch (args[i]) {
                case "-config": config = args[++i]; break;
                case "-deadlock": noDeadlock = true; break;
                case "-continue": cont = true; break;
                case "-workers": i++; break;           // accepted, unused: the GPU grid replaces workers
                default: spec = args[i];
            }
        }
        if (modelLib == null) {
            System.err.println("Error: -Dkspec.model=<libkmc_*.so> is required (AOT-lowered " + spec + " + " + config + ")");
            System.exit(150);
        }
        StringBuilder opts = new StringBuilder("{");
        opts.append("\"continue\": ").append(cont);
        if (noDeadlock) opts.append(", \"check_deadlock\": false");
        opts.append("}");
        long ctx = Native.create(modelLib, opts.toString());
        int rc = Native.run(ctx);
        long[] st = Native.stats(ctx);
        long[] v = Native.violation(ctx);
        int exit = 0;
        if (rc != 0) {
            System.out.println("Error: " + Native.strerror(ctx, rc));
            exit = 1;
        } else if (v != null) {
            System.out.println(v[0] == 2 ? "Error: Deadlock reached." : "Error: Invariant #" + v[1] + " is violated.");
            exit = v[0] == 2 ? 11 : 12;
        } else {
            System.out.println("Model checking completed. No error has been found.");
        }
        System.out.println(st[1] + " states generated, " + st[0] + " distinct states found, " + st[2] + " states left on queue.");
        if (st[8] == 1) System.out.println("The depth of the complete state graph search is " + st[3] + ".");
        Native.destroy(ctx);
        System.exit(exit);
    }
}
