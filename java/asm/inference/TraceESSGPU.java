package asm.inference;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import beast.base.core.Description;
import beast.base.core.Input;

import static java.lang.foreign.ValueLayout.*;

/**
 * Drop-in for TraceESS: converged(burnin, end) (TraceESS.java:38-69) with the gather + ACT + ESS of every
 * selected column (:53-62) done by one asm_ess_eval call.  The NaN-propagating Math.min (:63) and the
 * threshold (:68) stay here.
 *   <stoppingCriterion spec="asm.inference.TraceESSGPU" targetESS="100" traces="posterior,prior,likelihood"/>
 */
@Description("Stopping criterion based on ESS of selected items from trace, evaluated on a B200 GPU")
public class TraceESSGPU extends TraceESS {
    public Input<Integer> deviceInput = new Input<>("device", "CUDA device index (the first one when gpus > 1)", 0);
    public Input<Integer> gpusInput = new Input<>("gpus", "number of GPUs of this box to spread the evaluation over (devices device .. device+gpus-1)", 1);
    private GpuTraceStore store;
    private double[] ess;

    @Override
    public void setup(int nChains, TraceInfo traceInfo) {
        super.setup(nChains, traceInfo);
        store = new GpuTraceStore(nChains, deviceInput.get(), gpusInput.get());
    }

    @Override
    public boolean converged(int[] burnin, int end, boolean useMapping) {
        int[] map = traceInfo.getMap();                                           // :51
        int[] cols = new int[map.length];
        for (int j = 0; j < map.length; j++) cols[j] = useMapping ? map[j] : j;   // :54
        ess = new double[map.length];                                             // :52
        try (Arena a = Arena.ofConfined()) {
            store.syncTraces(traceInfo, end);
            MemorySegment b = a.allocateFrom(JAVA_INT, burnin), c = a.allocateFrom(JAVA_INT, cols);
            MemorySegment out = a.allocate(JAVA_DOUBLE, cols.length);
            int rc = (int) AsmGpuNative.asm_ess_eval.invokeExact(store.ctx, b, end, c, cols.length, out);
            AsmGpuNative.check(store.ctx, rc, "asm_ess_eval");
            for (int j = 0; j < cols.length; j++) ess[j] = out.getAtIndex(JAVA_DOUBLE, j);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
        double minESS = Double.POSITIVE_INFINITY;                                 // :50
        for (double e : ess) minESS = Math.min(e, minESS);                        // :63
        return minESS >= targetESS * nChains;                                     // :68
    }

    /** TraceESS.getLogMap (:191-198) reads the parent's private currentESSs; serve it from this class's copy. */
    @Override
    public java.util.Map<String, Double> getLogMap() {
        String[] traces = getTraceLabels();
        java.util.Map<String, Double> m = new java.util.HashMap<>(traces.length);
        for (int i = 0; i < traces.length; i++) m.put(traces[i], ess[i]);
        return m;
    }
}
