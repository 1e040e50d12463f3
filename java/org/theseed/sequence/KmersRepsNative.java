package org.theseed.sequence;

import java.lang.foreign.Arena;
import java.lang.foreign.FunctionDescriptor;
import java.lang.foreign.Linker;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.SymbolLookup;
import java.lang.invoke.MethodHandle;
import java.nio.charset.StandardCharsets;
import java.util.List;

import static java.lang.foreign.ValueLayout.ADDRESS;
import static java.lang.foreign.ValueLayout.JAVA_BYTE;
import static java.lang.foreign.ValueLayout.JAVA_DOUBLE;
import static java.lang.foreign.ValueLayout.JAVA_INT;
import static java.lang.foreign.ValueLayout.JAVA_LONG;

/**
 * Panama FFM (java.lang.foreign, Java 22+, or 21 with --enable-preview) binding of libkmersreps_cuda.so,
 * one MethodHandle per entry point of include/kmersreps.h.  COMPILE-UNVERIFIED: the build container has no
 * JDK; the C side of every signature is exercised by the ctypes binding in kmers/reps_b200/binding.py.
 *
 * There is no CPU fallback: if the library or a GPU is missing, the static initialiser throws and every class
 * that needs a similarity fails with it (reference convention: RuntimeException, App.java:129).
 */
public final class KmersRepsNative {

    public static final int INFINITY = 9999;
    public static final int IDENT_IS_INFINITY = 1;
    public static final int NO_REP = -1;          // 0xFFFFFFFF
    public static final int ELEM_I32 = 0, ELEM_U16 = 1;
    public static final int LAYOUT_FULL = 0, LAYOUT_TRIANGLE = 1, LAYOUT_BLOCKTRI = 2;

    private static final Linker LINKER = Linker.nativeLinker();
    private static final SymbolLookup LIB =
            SymbolLookup.libraryLookup(System.getProperty("kmersreps.lib", "libkmersreps_cuda.so"), Arena.global());

    private static MethodHandle h(String name, FunctionDescriptor fd) {
        return LINKER.downcallHandle(LIB.find(name).orElseThrow(() -> new UnsatisfiedLinkError(name)), fd);
    }

    public static final MethodHandle kr_create = h("kr_create", FunctionDescriptor.of(JAVA_INT, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS));
    public static final MethodHandle kr_destroy = h("kr_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    public static final MethodHandle kr_last_error = h("kr_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    public static final MethodHandle kr_encode = h("kr_encode", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS));
    public static final MethodHandle kr_seqset_free = h("kr_seqset_free", FunctionDescriptor.ofVoid(ADDRESS));
    public static final MethodHandle kr_seqset_sizes = h("kr_seqset_sizes", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    public static final MethodHandle kr_seqset_keys = h("kr_seqset_keys", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS));
    public static final MethodHandle kr_key_decode = h("kr_key_decode", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
    public static final MethodHandle kr_seqset_gather = h("kr_seqset_gather", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS));
    public static final MethodHandle kr_sim_pairs = h("kr_sim_pairs", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS));
    public static final MethodHandle kr_distance = h("kr_distance", FunctionDescriptor.of(JAVA_DOUBLE, JAVA_INT, JAVA_INT, JAVA_INT));
    public static final MethodHandle kr_sim_matrix = h("kr_sim_matrix", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    public static final MethodHandle kr_db_create = h("kr_db_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    public static final MethodHandle kr_db_free = h("kr_db_free", FunctionDescriptor.ofVoid(ADDRESS));
    public static final MethodHandle kr_db_append = h("kr_db_append", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
    public static final MethodHandle kr_db_set_ranks = h("kr_db_set_ranks", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
    public static final MethodHandle kr_db_find_closest = h("kr_db_find_closest", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS));
    public static final MethodHandle kr_db_any_ge = h("kr_db_any_ge", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS));
    public static final MethodHandle kr_greedy_build = h("kr_greedy_build", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS));

    public static final MethodHandle kr_seqset_from_keys = h("kr_seqset_from_keys", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS));
    public static final MethodHandle kr_sim_matrix_sparse64 = h("kr_sim_matrix_sparse64", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
    public static final MethodHandle kr_sim_matrix_sparse = h("kr_sim_matrix_sparse", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
    // several GPUs behind one handle (include/kmersreps.h, last section): -Dkmersreps.devices=0,1,2,3
    public static final MethodHandle kr_multi_create = h("kr_multi_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS));
    public static final MethodHandle kr_multi_destroy = h("kr_multi_destroy", FunctionDescriptor.ofVoid(ADDRESS));
    public static final MethodHandle kr_multi_last_error = h("kr_multi_last_error", FunctionDescriptor.of(ADDRESS, ADDRESS));
    public static final MethodHandle kr_multi_ctx = h("kr_multi_ctx", FunctionDescriptor.of(ADDRESS, ADDRESS, JAVA_INT));
    public static final MethodHandle kr_multi_sim_matrix = h("kr_multi_sim_matrix", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS));
    public static final MethodHandle kr_multi_sim_matrix_sparse = h("kr_multi_sim_matrix_sparse", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS));
    public static final MethodHandle kr_multi_db_create = h("kr_multi_db_create", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS));
    public static final MethodHandle kr_multi_db_free = h("kr_multi_db_free", FunctionDescriptor.ofVoid(ADDRESS));
    public static final MethodHandle kr_multi_db_append = h("kr_multi_db_append", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
    public static final MethodHandle kr_multi_db_set_ranks = h("kr_multi_db_set_ranks", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_INT));
    public static final MethodHandle kr_multi_db_find_closest = h("kr_multi_db_find_closest", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, ADDRESS));
    public static final MethodHandle kr_multi_greedy_build = h("kr_multi_greedy_build", FunctionDescriptor.of(JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS, JAVA_INT, JAVA_INT, JAVA_INT, ADDRESS));

    /** One context per k-mer size (ProteinKmers.setKmerSize is process-global in the reference). */
    private static final java.util.Map<Integer, MemorySegment> CONTEXTS = new java.util.HashMap<>();

    public static synchronized MemorySegment context(int key) {
        MemorySegment m = multi(key);
        if (m != null) {                                      // single-object calls use the first GPU of the multi handle
            try { return (MemorySegment) kr_multi_ctx.invokeExact(m, 0); } catch (Throwable t) { throw new RuntimeException(t); }
        }
        return CONTEXTS.computeIfAbsent(key, k -> {
            try (Arena a = Arena.ofConfined()) {
                MemorySegment out = a.allocate(ADDRESS);
                int device = Integer.getInteger("kmersreps.device", 0);
                int rc = (int) kr_create.invokeExact(device, (int) k, IDENT_IS_INFINITY, out);
                if (rc != 0) throw new RuntimeException("kmersreps: " + lastError(MemorySegment.NULL));
                return out.get(ADDRESS, 0);
            } catch (RuntimeException e) {
                throw e;
            } catch (Throwable t) {
                throw new RuntimeException(t);
            }
        });
    }

    /** The GPUs named by -Dkmersreps.devices=0,1,... behind one kr_multi handle per k-mer size; null when the property is unset. */
    private static final java.util.Map<Integer, MemorySegment> MULTI = new java.util.HashMap<>();

    public static synchronized MemorySegment multi(int key) {
        String list = System.getProperty("kmersreps.devices");
        if (list == null || list.isBlank()) return null;
        return MULTI.computeIfAbsent(key, k -> {
            String[] parts = list.split(",");
            try (Arena a = Arena.ofConfined()) {
                MemorySegment devs = a.allocate(JAVA_INT, parts.length), out = a.allocate(ADDRESS);
                for (int i = 0; i < parts.length; i++) devs.setAtIndex(JAVA_INT, i, Integer.parseInt(parts[i].trim()));
                int rc = (int) kr_multi_create.invokeExact(devs, parts.length, (int) k, IDENT_IS_INFINITY, out);
                if (rc != 0) throw new RuntimeException("kmersreps: " + multiError(MemorySegment.NULL));
                return out.get(ADDRESS, 0);
            } catch (RuntimeException e) {
                throw e;
            } catch (Throwable t) {
                throw new RuntimeException(t);
            }
        });
    }

    static String multiError(MemorySegment m) {
        try {
            MemorySegment p = (MemorySegment) kr_multi_last_error.invokeExact(m);
            return p.reinterpret(512).getString(0, StandardCharsets.US_ASCII);
        } catch (Throwable t) {
            return t.toString();
        }
    }

    public static void checkMulti(MemorySegment m, int rc) {
        if (rc != 0) throw new RuntimeException("kmersreps: " + multiError(m));
    }

    /** flat residue buffer + offsets of a list of sequences, as every batch entry point takes them */
    public static final class Flat {
        public final MemorySegment residues, offsets;
        public final int count;
        Flat(Arena a, List<String> seqs) {
            long total = 0;
            for (String p : seqs) total += p.length();
            this.residues = a.allocate(Math.max(total, 1));
            this.offsets = a.allocate(JAVA_LONG, seqs.size() + 1L);
            long pos = 0;
            for (int i = 0; i < seqs.size(); i++) {
                byte[] b = seqs.get(i).getBytes(StandardCharsets.ISO_8859_1);
                MemorySegment.copy(b, 0, this.residues, JAVA_BYTE, pos, b.length);
                this.offsets.setAtIndex(JAVA_LONG, i, pos);
                pos += b.length;
            }
            this.offsets.setAtIndex(JAVA_LONG, seqs.size(), pos);
            this.count = seqs.size();
        }
    }

    public static Flat flat(Arena a, List<String> seqs) { return new Flat(a, seqs); }

    static String lastError(MemorySegment ctx) {
        try {
            MemorySegment p = (MemorySegment) kr_last_error.invokeExact(ctx);
            return p.reinterpret(512).getString(0, StandardCharsets.US_ASCII);
        } catch (Throwable t) {
            return t.toString();
        }
    }

    public static void check(MemorySegment ctx, int rc) {
        if (rc != 0) throw new RuntimeException("kmersreps: " + lastError(ctx));
    }

    /** A batch of proteins encoded on the GPU (kr_seqset). */
    public static final class SeqSet implements AutoCloseable {
        final MemorySegment ctx, handle;
        final int count;
        final int[] sizes;

        SeqSet(MemorySegment ctx, MemorySegment handle, int count, int[] sizes) {
            this.ctx = ctx; this.handle = handle; this.count = count; this.sizes = sizes;
        }
        public MemorySegment handle() { return handle; }
        public int sizeOf(int i) { return sizes[i]; }
        @Override public void close() {
            try { kr_seqset_free.invokeExact(handle); } catch (Throwable t) { throw new RuntimeException(t); }
        }
    }

    /** kr_encode: flat residue buffer + offsets, one launch for the whole list. */
    public static SeqSet encode(int K, List<String> proteins) {
        MemorySegment ctx = context(K);
        try (Arena a = Arena.ofConfined()) {
            Flat f = flat(a, proteins);
            MemorySegment res = f.residues, off = f.offsets;
            MemorySegment out = a.allocate(ADDRESS);
            check(ctx, (int) kr_encode.invokeExact(ctx, res, off, proteins.size(), out));
            MemorySegment set = out.get(ADDRESS, 0);
            MemorySegment sz = a.allocate(JAVA_INT, Math.max(proteins.size(), 1));
            check(ctx, (int) kr_seqset_sizes.invokeExact(set, sz));
            int[] sizes = new int[proteins.size()];
            for (int i = 0; i < sizes.length; i++) sizes[i] = sz.getAtIndex(JAVA_INT, i);
            return new SeqSet(ctx, set, proteins.size(), sizes);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    /** kr_seqset_keys: the sorted packed k-mers of one element (SequenceKmers.iterator / retainAll / removeAll). */
    public static long[] keys(SeqSet s, int i) {
        int n = s.sizes[i];
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(JAVA_LONG, Math.max(n, 1)), cnt = a.allocate(JAVA_INT);
            check(s.ctx, (int) kr_seqset_keys.invokeExact(s.handle, i, out, n, cnt));
            return out.asSlice(0, 8L * n).toArray(JAVA_LONG);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    /** kr_key_decode: a packed k-mer back to its String. */
    public static String decode(int K, long key) {
        MemorySegment ctx = context(K);
        try (Arena a = Arena.ofConfined()) {
            MemorySegment out = a.allocate(Math.max(K, 1));
            check(ctx, (int) kr_key_decode.invokeExact(ctx, key, out));
            return new String(out.toArray(JAVA_BYTE), 0, K, StandardCharsets.ISO_8859_1);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    /** kr_seqset_from_keys: objects whose k-mer sets were edited (retainAll / removeAll) become a device set again. */
    public static SeqSet fromKeys(int K, long[][] keyLists, List<String> sequences) {
        MemorySegment ctx = context(K);
        try (Arena a = Arena.ofConfined()) {
            long total = 0;
            for (long[] k : keyLists) total += k.length;
            MemorySegment keys = a.allocate(JAVA_LONG, Math.max(total, 1)), koff = a.allocate(JAVA_LONG, keyLists.length + 1L);
            long pos = 0;
            int[] sizes = new int[keyLists.length];
            for (int i = 0; i < keyLists.length; i++) {
                koff.setAtIndex(JAVA_LONG, i, pos);
                MemorySegment.copy(keyLists[i], 0, keys, JAVA_LONG, 8 * pos, keyLists[i].length);
                pos += keyLists[i].length;
                sizes[i] = keyLists[i].length;
            }
            koff.setAtIndex(JAVA_LONG, keyLists.length, pos);
            Flat f = flat(a, sequences);
            MemorySegment out = a.allocate(ADDRESS);
            check(ctx, (int) kr_seqset_from_keys.invokeExact(ctx, keys, koff, f.residues, f.offsets, keyLists.length, out));
            return new SeqSet(ctx, out.get(ADDRESS, 0), keyLists.length, sizes);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    /** kr_sim_pairs for one pair (SequenceKmers.similarity). */
    public static int similarity(SeqSet a, int ia, SeqSet b, int ib) {
        try (Arena ar = Arena.ofConfined()) {
            MemorySegment pa = ar.allocate(JAVA_INT), pb = ar.allocate(JAVA_INT), sim = ar.allocate(JAVA_INT);
            pa.set(JAVA_INT, 0, ia); pb.set(JAVA_INT, 0, ib);
            check(a.ctx, (int) kr_sim_pairs.invokeExact(a.ctx, a.handle, pa, b.handle, pb, 1L, sim, MemorySegment.NULL));
            return sim.get(JAVA_INT, 0);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    /** kr_sim_matrix, full row-major int32 (rows of a x columns of b; b == null: a against itself). */
    public static int[] simMatrix(SeqSet a, SeqSet b) {
        int cols = b == null ? a.count : b.count;
        try (Arena ar = Arena.ofConfined()) {
            MemorySegment desc = ar.allocate(JAVA_INT, 4);
            desc.setAtIndex(JAVA_INT, 0, ELEM_I32); desc.setAtIndex(JAVA_INT, 1, LAYOUT_FULL);
            desc.setAtIndex(JAVA_INT, 2, 0); desc.setAtIndex(JAVA_INT, 3, a.count);
            MemorySegment out = ar.allocate(JAVA_INT, Math.max(1L, (long) a.count * cols));
            check(a.ctx, (int) kr_sim_matrix.invokeExact(a.ctx, a.handle, b == null ? MemorySegment.NULL : b.handle, desc, out));
            return out.toArray(JAVA_INT);
        } catch (RuntimeException e) {
            throw e;
        } catch (Throwable t) {
            throw new RuntimeException(t);
        }
    }

    /**
     * kr_sim_matrix_sparse: the cells (row, col, sim) with sim >= minSim of a against b (b == null: the strict upper
     * triangle of a against itself), 3 ints per cell, unordered.  Every other cell is below minSim (0 for minSim = 1).
     */
    public static int[] simCells(SeqSet a, SeqSet b, int minSim) {
        long cap = Math.max(1024L, 16L * a.count);
        for (;;) {
            try (Arena ar = Arena.ofConfined()) {
                MemorySegment desc = ar.allocate(JAVA_INT, 4);
                desc.setAtIndex(JAVA_INT, 0, ELEM_I32); desc.setAtIndex(JAVA_INT, 1, b == null ? LAYOUT_TRIANGLE : LAYOUT_FULL);
                desc.setAtIndex(JAVA_INT, 2, 0); desc.setAtIndex(JAVA_INT, 3, a.count);
                // packed cells (8 bytes: row << 40 | col << 16 | sim) when both sets have at most 2^24 sequences, else 16-byte cells
                boolean packed = a.count <= (1 << 24) && (b == null || b.count <= (1 << 24));
                MemorySegment cells = ar.allocate(JAVA_LONG, (packed ? 1 : 2) * cap), cnt = ar.allocate(JAVA_LONG);
                MemorySegment bh = b == null ? MemorySegment.NULL : b.handle;
                int rc = packed ? (int) kr_sim_matrix_sparse64.invokeExact(a.ctx, a.handle, bh, desc, minSim, cells, cap, cnt)
                                : (int) kr_sim_matrix_sparse.invokeExact(a.ctx, a.handle, bh, desc, minSim, cells, cap, cnt);
                long n = cnt.get(JAVA_LONG, 0);
                if (rc != 0 && n > cap) { cap = n; continue; }      // the library told how many cells qualify: once more with room
                check(a.ctx, rc);
                int[] out = new int[(int) (3 * n)];
                for (long k = 0; k < n; k++) {
                    if (packed) {
                        long c = cells.getAtIndex(JAVA_LONG, k);
                        out[(int) (3 * k)] = (int) (c >>> 40); out[(int) (3 * k + 1)] = (int) ((c >>> 16) & 0xFFFFFF); out[(int) (3 * k + 2)] = (int) (c & 0xFFFF);
                    } else
                        for (int f = 0; f < 3; f++) out[(int) (3 * k + f)] = cells.getAtIndex(JAVA_INT, 4 * k + f);
                }
                return out;
            } catch (RuntimeException e) {
                throw e;
            } catch (Throwable t) {
                throw new RuntimeException(t);
            }
        }
    }

    public static double distance(int sim, int sizeA, int sizeB) {
        try { return (double) kr_distance.invokeExact(sim, sizeA, sizeB); } catch (Throwable t) { throw new RuntimeException(t); }
    }

    private KmersRepsNative() { }
}
