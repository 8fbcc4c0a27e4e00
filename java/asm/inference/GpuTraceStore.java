package asm.inference;

import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.util.*;

import beast.base.evolution.tree.Node;
import beast.base.evolution.tree.Tree;

import static java.lang.foreign.ValueLayout.*;

/**
 * Keeps the device copy of a TraceInfo in step with the Java lists: every tree the runner appended to
 * traceInfo.trees[chain] (AutoStopMCMC.java:536) is encoded as a clade bit row and appended with
 * asm_trees_append; every log line appended to traceInfo.logLines[chain] (AutoStopMCMC.java:462) goes to
 * asm_traces_append.  Clade identity is CladeDifference.traverse's (CladeDifference.java:107-141): the sorted
 * leaf numbers below an internal node, root included; ids are handed out in first-seen order.
 * Lives in package asm.inference because TraceInfo.trees / logLines are package-private (TraceInfo.java:21-29).
 */
final class GpuTraceStore implements AutoCloseable {
    final MemorySegment ctx;
    private final Arena arena = Arena.ofShared();
    private final Map<BitSet, Integer> dictionary = new HashMap<>();
    private final int[] treesSent, linesSent;

    /** gpus == 1: one context on `device`; gpus > 1: devices device .. device+gpus-1 behind ONE context
     *  (asm_create_multi: tile pairs, traces and clade counts are dealt to the GPUs inside the library and combined
     *  with NCCL; results are bit-identical to one GPU). */
    GpuTraceStore(int nChains, int device, int gpus) {
        MemorySegment out = arena.allocate(ADDRESS);
        try {
            int rc;
            if (gpus <= 1) {
                rc = (int) AsmGpuNative.asm_create.invokeExact(out, device, nChains);
            } else {
                int[] devices = new int[gpus];
                for (int i = 0; i < gpus; i++) devices[i] = device + i;
                rc = (int) AsmGpuNative.asm_create_multi.invokeExact(out, arena.allocateFrom(JAVA_INT, devices), gpus, nChains);
            }
            if (rc != 0) throw new IllegalStateException("asm_create failed (" + rc + "): no usable B200; there is no CPU fallback");
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new IllegalStateException(t);
        }
        ctx = out.get(ADDRESS, 0);
        treesSent = new int[nChains];
        linesSent = new int[nChains];
    }

    /** words per row for the current dictionary, a multiple of 4 (256 bits) */
    private int words() {
        int w = (dictionary.size() + 63) / 64;
        return Math.max(4, (w + 3) / 4 * 4);
    }

    private BitSet clades(Node node, List<BitSet> out) {
        BitSet b = new BitSet();
        if (node.isLeaf()) {
            b.set(node.getNr());
        } else {
            b.or(clades(node.getLeft(), out));
            b.or(clades(node.getRight(), out));
            out.add(b);
        }
        return b;
    }

    void syncTrees(TraceInfo traceInfo, int end) throws Throwable {
        for (int chain = 0; chain < treesSent.length; chain++) {
            List<Tree> trees = traceInfo.trees[chain];
            int have = Math.min(end, trees.size());
            if (have <= treesSent[chain]) continue;
            int nNew = have - treesSent[chain];
            List<int[]> ids = new ArrayList<>();
            for (int i = treesSent[chain]; i < have; i++) {
                List<BitSet> cl = new ArrayList<>();
                clades(trees.get(i).getRoot(), cl);
                int[] row = new int[cl.size()];
                for (int k = 0; k < row.length; k++) {
                    Integer id = dictionary.get(cl.get(k));
                    if (id == null) {
                        id = dictionary.size();
                        dictionary.put(cl.get(k), id);
                    }
                    row[k] = id;
                }
                ids.add(row);
            }
            try (Arena a = Arena.ofConfined()) {
                if (dictionary.size() <= 0xFFFF) {
                    // the ids themselves go up (two bytes per clade); the library builds the bit rows on the GPU
                    int per = 1;
                    for (int[] row : ids) per = Math.max(per, row.length);
                    MemorySegment buf = a.allocate(JAVA_SHORT, (long) nNew * per);
                    for (int t = 0; t < nNew; t++) {
                        int[] row = ids.get(t);
                        for (int k = 0; k < per; k++)   // ASM_NO_CLADE in the slots a multifurcating tree leaves empty
                            buf.setAtIndex(JAVA_SHORT, (long) t * per + k, (short) (k < row.length ? row[k] : 0xFFFF));
                    }
                    int rc = (int) AsmGpuNative.asm_trees_append_ids.invokeExact(ctx, chain, (long) nNew, per, buf, dictionary.size());
                    AsmGpuNative.check(ctx, rc, "asm_trees_append_ids");
                } else {
                    int w = words();
                    MemorySegment bits = a.allocate(JAVA_LONG, (long) nNew * w);
                    for (int t = 0; t < nNew; t++)
                        for (int id : ids.get(t)) {
                            long off = ((long) t * w + (id >> 6)) * 8;
                            bits.set(JAVA_LONG, off, bits.get(JAVA_LONG, off) | (1L << (id & 63)));
                        }
                    int rc = (int) AsmGpuNative.asm_trees_append.invokeExact(ctx, chain, (long) nNew, w, bits);
                    AsmGpuNative.check(ctx, rc, "asm_trees_append");
                }
            }
            treesSent[chain] = have;
        }
    }

    void syncTraces(TraceInfo traceInfo, int end) throws Throwable {
        for (int chain = 0; chain < linesSent.length; chain++) {
            List<Double>[] cols = traceInfo.logLines[chain];
            int have = Math.min(end, cols[0].size());
            if (have <= linesSent[chain]) continue;
            int nNew = have - linesSent[chain], nCols = cols.length;
            try (Arena a = Arena.ofConfined()) {
                MemorySegment v = a.allocate(JAVA_DOUBLE, (long) nNew * nCols);
                for (int x = 0; x < nNew; x++)
                    for (int c = 0; c < nCols; c++)
                        v.setAtIndex(JAVA_DOUBLE, (long) x * nCols + c, cols[c].get(linesSent[chain] + x));
                int rc = (int) AsmGpuNative.asm_traces_append.invokeExact(ctx, chain, nCols, (long) nNew, v);
                AsmGpuNative.check(ctx, rc, "asm_traces_append");
            }
            linesSent[chain] = have;
        }
    }

    @Override
    public void close() {
        try {
            int rc = (int) AsmGpuNative.asm_destroy.invokeExact(ctx);
        } catch (Throwable ignored) {
        }
        arena.close();
    }
}
