package asm.inference;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;

import beast.base.core.Description;
import beast.base.core.Input;

import static java.lang.foreign.ValueLayout.*;

/**
 * Drop-in for TreePSRF: same inputs (inherited), same converged2sided control flow (TreePSRF.java:107-197),
 * with the two row loops (:141-144, :154-156) -- i.e. every distancePlusOne/calcPSRF call -- replaced by ONE
 * asm_psrf_eval call that returns psrf1mean and psrf2mean.  The tree distance is Robinson-Foulds on clade
 * bit-vectors (BASELINE.json north_star) instead of BEASTLabs' RNNIMetric (TreeESS.java:179-180).
 * Register as a provider in version.xml next to asm.inference.TreePSRF and use
 *   <stoppingCriterion spec="asm.inference.TreePSRFGPU" cacheLimit="200" smoothing="0.9" targetESS="100"/>
 */
@Description("Gelman-Rubin like criterion for convergence based on trees alone, evaluated on a B200 GPU")
public class TreePSRFGPU extends TreePSRF {
    public Input<Integer> deviceInput = new Input<>("device", "CUDA device index (the first one when gpus > 1)", 0);
    public Input<Integer> gpusInput = new Input<>("gpus", "number of GPUs of this box to spread the evaluation over (devices device .. device+gpus-1)", 1);
    public Input<String> methodInput = new Input<>("method", "distance kernel: i8 (int8 tcgen05), fp4 (block-scaled FP4 tcgen05, same integers, fastest) or popc", "i8");

    private GpuTraceStore store;
    private int start0 = -1;
    private boolean prev = false;
    private double lower, upper;

    @Override
    public void initAndValidate() {
        super.initAndValidate();
        double b = bInput.get();
        upper = 1.0 + b;
        lower = 1.0 - b;
    }

    @Override
    public void setup(int nChains, TraceInfo traceInfo) {
        super.setup(nChains, traceInfo);           // keeps the 2-chain rule and grValues initialisation
        store = new GpuTraceStore(nChains, deviceInput.get(), gpusInput.get());
    }

    @Override
    public boolean converged(int[] burnin, int end) {
        if (!twoSidedInput.get()) return super.converged(burnin, end);   // one-sided path unchanged
        prev = converged2sidedGpu(burnin, end);
        return prev;
    }

    private boolean converged2sidedGpu(int[] burnin, int end) {
        try {
            if (end % delta != 0) return prev;                                   // :113-116
            int start = (int) (end * (1.0 - smoothingInput.get()));              // :117
            start = start - start % delta;                                       // :118
            if (end / delta > cacheLimit) {                                      // :120-124
                delta *= 2;
                return converged(burnin, end);
            }
            store.syncTrees(traceInfo, end);
            double psrf1mean, psrf2mean = 0;
            try (Arena a = Arena.ofConfined()) {
                MemorySegment b = a.allocateFrom(JAVA_INT, burnin);
                MemorySegment mean = a.allocate(JAVA_DOUBLE, 4);
                int method = "popc".equals(methodInput.get()) ? AsmGpuNative.ASM_RF_POPC
                        : "fp4".equals(methodInput.get()) ? AsmGpuNative.ASM_RF_FP4 : AsmGpuNative.ASM_RF_I8;
                int rc = (int) AsmGpuNative.asm_psrf_eval.invokeExact(store.ctx, b, start, end, delta, method,
                        MemorySegment.NULL, mean);
                AsmGpuNative.check(store.ctx, rc, "asm_psrf_eval");
                psrf1mean = mean.getAtIndex(JAVA_DOUBLE, 0 * 2 + 1);             // :141-145
                psrf2mean = mean.getAtIndex(JAVA_DOUBLE, 1 * 2 + 0);             // :154-157 (used only below)
            }
            getGrValues()[0] = psrf1mean;                                        // :146
            if (lower < psrf1mean && psrf1mean < upper) {                        // :150
                if (start0 < 0) start0 = start;                                  // :151-153
                getGrValues()[1] = psrf2mean;                                    // :158
                if (lower < psrf2mean && psrf2mean < upper) {                    // :161
                    if (!checkESSInput.get()) return true;                       // :162-164
                    if ((end - start0) <= targetESS) return false;               // :168-171
                    return true;                                                 // :176 (checkESS || ...)
                }
            } else {
                psrf2mean = 0;                                                   // :147
            }
            if (lower <= psrf1mean || psrf1mean >= upper || lower <= psrf2mean || psrf2mean >= upper) {
                start0 = -1;                                                     // :184-187
            }
        } catch (Throwable e) {                                                  // :188-194
            beast.base.core.Log.warning("Programmer error: Something is WRONG and needs fixing:");
            e.printStackTrace();
            start0 = -1;
            return false;
        }
        return false;
    }
}
