package asm.inference;

import java.lang.foreign.*;
import java.lang.invoke.MethodHandle;

import static java.lang.foreign.ValueLayout.*;

/**
 * Panama FFM (JDK 22+) binding of libasmgpu.so -- the C ABI declared in include/asmgpu.h.
 * One downcall handle per entry point the two GPU criteria use.  Not compiled in this repository
 * (the build image has no JDK); kept as the reference-side stub a maintainer adds to the ASM package.
 */
final class AsmGpuNative {
    static final Linker LINKER = Linker.nativeLinker();
    static final SymbolLookup LIB = SymbolLookup.libraryLookup(
            System.getProperty("asmgpu.library", "libasmgpu.so"), Arena.global());

    private static MethodHandle h(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(
                () -> new UnsatisfiedLinkError("libasmgpu.so does not export " + name)), fd);
    }

    static final MethodHandle asm_create = h("asm_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT));
    /** asm_create_multi(&ctx, devices[], n_devices, n_chains): one context over several GPUs of the box, driven from the
     *  one thread that calls the criteria (AutoStopMCMC.java:397-409); every other handle takes it unchanged. */
    static final MethodHandle asm_create_multi = h("asm_create_multi", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT));
    static final MethodHandle asm_destroy = h("asm_destroy", FunctionDescriptor.of(JAVA_INT, ADDRESS));
    static final MethodHandle asm_last_error = h("asm_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    static final MethodHandle asm_trees_append = h("asm_trees_append",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_LONG, JAVA_INT, ADDRESS));
    static final MethodHandle asm_trees_append_ids = h("asm_trees_append_ids",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_LONG, JAVA_INT, ADDRESS, JAVA_INT));
    static final MethodHandle asm_psrf_eval = h("asm_psrf_eval",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, ADDRESS));
    static final MethodHandle asm_rf_block = h("asm_rf_block",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_LONG, JAVA_LONG, JAVA_INT, JAVA_LONG, JAVA_LONG, JAVA_INT, ADDRESS));
    static final MethodHandle asm_traces_append = h("asm_traces_append",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_LONG, ADDRESS));
    static final MethodHandle asm_ess_eval = h("asm_ess_eval",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    static final MethodHandle asm_ess_batch = h("asm_ess_batch",
            FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_LONG, JAVA_LONG, ADDRESS, JAVA_INT, ADDRESS, ADDRESS));

    static final int ASM_OK = 0, ASM_RF_POPC = 0, ASM_RF_I8 = 1, ASM_RF_FP4 = 2;

    /** Throws with the library's message; the criteria catch it exactly where the reference catches Throwable. */
    static void check(MemorySegment ctx, int rc, String what) {
        if (rc == ASM_OK) return;
        String msg;
        try {
            MemorySegment p = (MemorySegment) asm_last_error.invokeExact(ctx);
            msg = p.reinterpret(1024).getString(0);
        } catch (Throwable t) {
            msg = "?";
        }
        throw new IllegalStateException(what + " failed (" + rc + "): " + msg);
    }

    private AsmGpuNative() {}
}
