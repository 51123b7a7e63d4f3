// The affine traceback walk -- backtrace_affine (src/algn.c:1983-2097) with the median / medianwg / aligned pair /
// bitset / closest outputs -- as a device function of ONE thread, shared by the standalone traceback kernel
// (trace_kernels.cuh) and by the ring fill kernels, which walk the pairs they filled themselves (aff_ring_kernels.cuh).
// The direction band is read through a small decoder object, so the same walk serves every band format.
#pragma once
#include "cells.cuh"

namespace poyb200 {

// Writes a sequence right-to-left, exactly like the reference's seq_prepend (src/seq.c:147-153), into a row
// whose end is 4-byte aligned; bytes are gathered into 32-bit words before they are stored.
struct RevWriter {
    uint8_t *base;
    int pos;  // next byte goes to pos - 1
    uint32_t acc;
    int n;
    __device__ __forceinline__ void init(uint8_t *row, int cap) { base = row; pos = cap; acc = 0; n = 0; }
    __device__ __forceinline__ void put(int v) {
        pos--;
        n++;
        acc |= ((uint32_t) v & 0xffu) << ((pos & 3) * 8);
        if ((pos & 3) == 0) {
            __stcs(reinterpret_cast<uint32_t *>(base + pos), acc);  // streaming: written once, read by the host
            acc = 0;
        }
    }
    __device__ __forceinline__ void flush() {
        for (int k = pos; (k & 3) != 0; k++) base[k] = (uint8_t) (acc >> ((k & 3) * 8));
    }
};

// One flag per alignment column, written right-to-left like RevWriter: a right-aligned bit string, most significant bit
// of a byte first (numpy.unpackbits order), gathered into 32-bit words.  capbits is a multiple of 32.
struct RevBitWriter {
    uint32_t *base;
    int pos;
    uint32_t acc;
    __device__ __forceinline__ void init(uint8_t *row, int capbits) { base = reinterpret_cast<uint32_t *>(row); pos = capbits; acc = 0; }
    __device__ __forceinline__ void put(bool bit) {
        pos--;
        acc |= (uint32_t) bit << ((((pos >> 3) & 3) << 3) + 7 - (pos & 7));
        if ((pos & 31) == 0) {
            __stcs(base + (pos >> 5), acc);
            acc = 0;
        }
    }
    __device__ __forceinline__ void flush() {
        if (pos & 31) __stcs(base + (pos >> 5), acc);
    }
};

__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// Band of one byte per cell, addressed through the Task (dir_index, common.cuh): what the generic and the legacy stripe
// kernels write.  Written by an EARLIER kernel, so the read-only path (__ldg) is safe.
struct BandBytes {
    const uint8_t *dbase;
    uint32_t G, twoK, BL, recip;
    int dbase_d, tshift;
    bool dir6;  // TF_DIR6: BL = 4, five 6-bit codes per word, END_BLOCK implied
    uint64_t tile_bytes;
    __device__ __forceinline__ BandBytes(const Task &t, const uint8_t *base)
        : dbase(base), G(t.G), twoK(t.twoK), BL(t.BL), dbase_d(t.dbase), tshift(t.tshift), dir6((t.flags & TF_DIR6) != 0) {
        tile_bytes = (uint64_t) t.G * 8 * t.BL;
        // dd / twoK == (dd * recip) >> 16 for dd < 16384: twoK is 8..16 for the stripe kernels (stripes <= 512 diagonals) and
        // 0xFFFF for the generic ones, where the quotient is 0 for every dd the 16384-element cap allows
        recip = (65536u + t.twoK - 1) / t.twoK;
    }
    __device__ __forceinline__ void start(int, int) {}
    __device__ __forceinline__ void finish() {}
    __device__ __forceinline__ int fetch(int i, int j) const {
        const uint32_t dd = (uint32_t) ((j - i) - dbase_d), T = (uint32_t) (i + j - tshift);
        const uint32_t lane = (dd * recip) >> 16, m = (dd - lane * twoK) >> 1;
        if (dir6) {
            const uint64_t widx = ((((uint64_t) (T >> 3) * G + lane) << 3) + (T & 7)) * 4;
            if (widx >= 2 * tile_bytes) prefetch_l1(dbase + widx - 2 * tile_bytes);
            return (int) ((__ldg(reinterpret_cast<const uint32_t *>(dbase + widx)) >> (6 * m)) & 63u) | AB_ENDB;
        }
        const uint64_t idx = ((((uint64_t) (T >> 3) * G + lane) << 3) + (T & 7)) * BL + m;
        // the direction line needed two tiles (8 diagonal moves) further on is requested early, so that the common case --
        // a run of matches along one diagonal -- finds it in L1
        if (idx >= 2 * tile_bytes) prefetch_l1(dbase + idx - 2 * tile_bytes);
        return __ldg(dbase + idx);
    }
};

// One operand read backwards, element by element: the walk consumes indices in descending order, one step at a time, so
// the reader keeps the aligned 32-bit word that holds the current element and the word below it, and fetches the next word
// a whole word (up to four steps) before it is needed -- the element itself never waits for a global load.
struct SeqBack {
    uintptr_t acur, lo;  // aligned address of `cur`; lowest address that may be read
    uint32_t cur, nxt;
    const uint8_t *base;
    __device__ __forceinline__ void init(const uint8_t *pool, const uint8_t *seq, int i) {
        base = seq;
        lo = reinterpret_cast<uintptr_t>(pool) & ~(uintptr_t) 3;
        acur = reinterpret_cast<uintptr_t>(seq + i) & ~(uintptr_t) 3;
        cur = __ldg(reinterpret_cast<const uint32_t *>(acur));
        nxt = (acur >= lo + 4) ? __ldg(reinterpret_cast<const uint32_t *>(acur - 4)) : 0u;
    }
    __device__ __forceinline__ int get(int i) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(base + i);
        if ((a & ~(uintptr_t) 3) != acur) {  // one word down
            acur -= 4;
            cur = nxt;
            nxt = (acur >= lo + 4) ? __ldg(reinterpret_cast<const uint32_t *>(acur - 4)) : 0u;
        }
        return (int) ((cur >> (8 * (a & 3))) & 0xffu);
    }
};

// cm_get_median for 4-bit codes from a 16 x 16 byte table in shared memory (the walk of a ring kernel) or from the matrix
// in global memory.
struct MedianGlobal {
    const uint8_t *median;
    int lcm;
    __device__ __forceinline__ int get(int a, int b) const { return __ldg(median + (a << lcm) + b); }
};
struct MedianShared {
    uint32_t tab;  // shared address of 256 bytes, index (a << 4) | b
    __device__ __forceinline__ int get(int a, int b) const {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(tab + ((a << 4) | b)));
        return (int) v;
    }
};

// dcap = device row stride (multiple of 16).
template <class Band, class Med>
__device__ __forceinline__ void aff_walk_pair(const Task &t, const uint8_t *__restrict__ pool, Band band, const Med med_lut,
                                              const DevCM &cm, const OutPtrs &out) {
    SeqBack si, sj;
    const int dcap = (int) out.stride;
    const size_t row = (size_t) t.pair * out.stride;
    const bool w_clo = out.want & 16;
    const bool w_med = (out.want & 1) && !w_clo, w_wg = out.want & 2, w_al = out.want & 4, w_bits = out.want & 8;
    RevWriter med, wg, ri, rj;
    RevBitWriter bi, bj, bw;
    // resi belongs to the row sequence; rows may be the caller's operand b (algn.c:2606-2616)
    const bool rows_b = (t.flags & TF_ROWS_ARE_B) != 0;
    {
        const size_t brow = w_bits ? (size_t) t.pair * out.bstride : 0;
        bi.init((rows_b ? out.bits_b : out.bits_a) + brow, (int) out.bstride * 8);
        bj.init((rows_b ? out.bits_a : out.bits_b) + brow, (int) out.bstride * 8);
        bw.init(out.bits_wg + brow, (int) out.bstride * 8);
    }
    med.init(out.median + ((w_med || w_clo) ? row : 0), dcap);
    wg.init(out.medianwg + (w_wg ? row : 0), dcap);
    ri.init((rows_b ? out.al_b : out.al_a) + (w_al ? row : 0), dcap);
    rj.init((rows_b ? out.al_a : out.al_b) + (w_al ? row : 0), dcap);
    int i = t.lr - 1, j = t.lc - 1;
    si.init(pool, pool + t.off_r, i);
    sj.init(pool, pool + t.off_c, j);
    int ic = si.get(i), jc = sj.get(j);
    if (i > 0 && j > 0) band.start(i, j);
    int nmed = 0, nwg = 0, nres = 0, med_first = -1, nclo = 0;
    // The reference's mode machine (m_todo / vertical / horizontal / diagonal / align, :2003-2075) and its two tails
    // (`while (i)`, `while (j)`, :2076-2091) as ONE branch-free step: the 32 walkers of a warp are in different modes, and with
    // a branch per mode the warp issued every branch (ncu: ~250 warp instructions per step); here every walker runs the same
    // instructions and the mode only selects values.  Modes carry the AM_* codes of the direction byte, 4 = m_todo.
    constexpr int M_TODO = 4;
    int mode = M_TODO;
    while ((i | j) != 0) {
        const bool inside = (i != 0) & (j != 0);
        int b = 0;
        if (inside) {
            const int d = j - i;
            if (d < t.dlo) b = AFF_LEFT_EDGE_BYTE;
            else if (d > t.dhi) b = AFF_RIGHT_EDGE_BYTE;
            else b = band.fetch(i, j);
        }
        // m_todo reads the cell and the next iteration acts on the same cell (:2003-2013): resolved in place.  Outside the
        // matrix the tails move vertically while rows are left, then horizontally.
        const int eff = inside ? ((mode == M_TODO) ? ((b >> 2) & 3) : mode) : ((i != 0) ? AM_V : AM_H);
        const bool isH = eff == AM_H, isV = eff == AM_V, isA = eff == AM_A, isD = eff == AM_D;
        const int a_el = isH ? TMPGAP : ic, b_el = isV ? TMPGAP : jc;  // the column of resi / resj
        const int x = isV ? ic : jc;                                    // the element an indel column is built from
        const int p = med_lut.get(ic & 15, jc & 15);
        const int wgv = isA ? p : ((isD || (x & TMPGAP)) ? TMPGAP : (x | TMPGAP));
        const bool emit = isA || (!isD && !(x & TMPGAP));
        nres++;
        nwg++;
        if (w_al) { ri.put(a_el); rj.put(b_el); }
        if (w_wg) wg.put(wgv);
        if (w_bits) { bi.put(a_el != TMPGAP); bj.put(b_el != TMPGAP); bw.put(wgv != TMPGAP); }
        if (w_clo) {
            const int sel = rows_b ? closest_elem(cm, b_el, a_el) : closest_elem(cm, a_el, b_el);
            if (sel != TMPGAP) { nclo++; med.put(sel); }
        }
        if (emit) {
            nmed++;
            med_first = wgv;
            if (w_med) med.put(wgv);
        }
        const int nx = b & 3;
        const int after_align = (nx == AN_H) ? AM_H : (nx == AN_D) ? AM_D : (nx == AN_V) ? AM_V : AM_A;
        const int endbit = isV ? AB_ENDV : isH ? AB_ENDH : AB_ENDB;
        mode = isA ? after_align : ((b & endbit) ? M_TODO : eff);
        i -= !isH;
        j -= !isV;
        ic = si.get(i);
        jc = sj.get(j);
    }
    band.finish();
    // the leading column: (gap, gap), a gap in medianwg, and a gap in front of the median unless it starts with one
    nres++;
    nwg++;
    if (w_al) { ri.put(TMPGAP); rj.put(TMPGAP); }
    if (w_bits) { bi.put(false); bj.put(false); bw.put(false); }
    if (w_wg) wg.put(TMPGAP);
    if (med_first != TMPGAP) {  // :2093 (an empty median counts as "not a gap")
        nmed++;
        if (w_med) med.put(TMPGAP);
    }
    if (w_clo) { med.put(TMPGAP); nclo++; }  // `prepend res gap`, src/sequence.ml:980
    if (w_med || w_clo) med.flush();
    if (w_wg) wg.flush();
    if (w_al) { ri.flush(); rj.flush(); }
    if (w_bits) { bi.flush(); bj.flush(); bw.flush(); }
    int *ol = out.out_len + 4 * (size_t) t.pair;
    ol[0] = w_clo ? nclo : nmed; ol[1] = nwg; ol[2] = nres; ol[3] = nres;
}

// ---- linear walk ------------------------------------------------------------------------------------------------------

// The direction band of a linear pair in any of its formats (common.cuh dir_index / dir_fetch), with the addressing of
// the pair resolved once: 32-bit arithmetic, the division by the lane width as a multiply.
struct LinBand {
    const uint8_t *dbase;
    uint32_t G, twoK, BL, magic, flags;
    int dbase_d, tshift;
    __device__ __forceinline__ LinBand(const Task &t, const uint8_t *base)
        : dbase(base), G(t.G), twoK(t.twoK), BL(t.BL), flags(t.flags), dbase_d(t.dbase), tshift(t.tshift) {
        magic = (uint32_t) ((0x100000000ull + t.twoK - 1) / t.twoK);  // x / twoK == umulhi(x, magic) for x < 2^32 / twoK
    }
    // 2-bit resolved move (TF_DIR2: 0 ALIGN, 1 the preferred gap move, 2 the other) or the reference's mask of minima
    __device__ __forceinline__ int fetch(int i, int j) const {
        if (flags & TF_ROWMAJ) {
            const uint32_t lane = __umulhi((uint32_t) j, magic), m = (uint32_t) j - lane * twoK;
            const uint32_t w = ((((uint32_t) i >> 3) * G + lane) << 3) + ((uint32_t) i & 7);
            return (int) ((__ldg(reinterpret_cast<const uint32_t *>(dbase) + w) >> (2 * m)) & 3u);
        }
        const uint32_t dd = (uint32_t) ((j - i) - dbase_d), T = (uint32_t) (i + j - tshift);
        const uint32_t lane = __umulhi(dd, magic), m = (dd - lane * twoK) >> 1;
        const uint32_t chunk = (((T >> 3) * G + lane) << 3) + (T & 7);
        if (flags & TF_DIR2) return (int) ((__ldg(reinterpret_cast<const uint32_t *>(dbase) + chunk) >> (2 * m)) & 3u);
        return __ldg(dbase + (size_t) chunk * BL + m);
    }
};

// backtrack_2d, linear branch (src/algn.c:3606-3665), fused with algn_ancestor_2 (:4126-4147) and
// algn_get_median_2d_with_gaps (:4024-4035), one thread per pair.  Every walker runs the same instructions; the move only
// selects values (the walkers of a warp take different moves).
__device__ __forceinline__ void lin_walk_pair(const Task &t, const uint8_t *__restrict__ pool, const LinBand band, const DevCM &cm,
                                              const OutPtrs &out) {
    const int dcap = (int) out.stride, gap = cm.gap;
    const size_t row = (size_t) t.pair * out.stride;
    const bool w_clo = out.want & 16;
    const bool w_med = (out.want & 1) || w_clo, w_wg = out.want & 2, w_al = out.want & 4, w_bits = out.want & 8;
    const bool rows_b = (t.flags & TF_ROWS_ARE_B) != 0;
    const bool swaped = (t.flags & TF_SWAPED) != 0;
    const bool dir2 = (t.flags & TF_DIR2) != 0;
    RevWriter med, wg, r1, r2;
    RevBitWriter b1, b2, bw;
    {
        const size_t brow = w_bits ? (size_t) t.pair * out.bstride : 0;
        b1.init((rows_b ? out.bits_b : out.bits_a) + brow, (int) out.bstride * 8);
        b2.init((rows_b ? out.bits_a : out.bits_b) + brow, (int) out.bstride * 8);
        bw.init(out.bits_wg + brow, (int) out.bstride * 8);
    }
    med.init(out.median + (w_med ? row : 0), dcap);
    wg.init(out.medianwg + (w_wg ? row : 0), dcap);
    r1.init((rows_b ? out.al_b : out.al_a) + (w_al ? row : 0), dcap);
    r2.init((rows_b ? out.al_a : out.al_b) + (w_al ? row : 0), dcap);
    int i = t.lr - 1, j = t.lc - 1, n = 0, nmed = 0;
    SeqBack si, sj;
    si.init(pool, pool + t.off_r, i);
    sj.init(pool, pool + t.off_c, j);
    int ic = si.get(i), jc = sj.get(j);
    const int second = swaped ? D_INSERT : D_DELETE, third = swaped ? D_DELETE : D_INSERT;
    // `while (end >= beg)` over the row-major matrix: stops after the ALIGN step out of cell (0, 0)
    while ((i | j) >= 0) {
        const int m = band.fetch(i, j);
        int mv;
        if (dir2) mv = (m == 0) ? D_ALIGN : (m == 1) ? second : third;  // the stripe kernels resolved the tie already
        else mv = (m & D_ALIGN) ? D_ALIGN : (m & second) ? second : third;
        const bool isI = mv == D_INSERT, isD = mv == D_DELETE;
        const int x = isI ? gap : ic, y = isD ? gap : jc;  // elements of s1 / s2 in this column
        n++;
        if (w_al) { r1.put(x); r2.put(y); }
        const int ea = rows_b ? y : x, eb = rows_b ? x : y;  // caller's operand order
        const int mm = cm_median(cm, ea, eb);
        if (w_wg) wg.put(mm);
        if (w_bits) { b1.put(x != gap); b2.put(y != gap); bw.put(mm != gap); }
        if (w_clo) {
            const int sel = closest_elem(cm, ea, eb);
            if (sel != gap) { nmed++; med.put(sel); }
        } else if (mm != gap) {
            nmed++;
            if (w_med) med.put(mm);
        }
        i -= !isI;
        j -= !isD;
        if ((i | j) >= 0) { ic = si.get(i); jc = sj.get(j); }
    }
    nmed++;
    if (w_med) { med.put(gap); med.flush(); }
    if (w_wg) wg.flush();
    if (w_al) { r1.flush(); r2.flush(); }
    if (w_bits) { b1.flush(); b2.flush(); bw.flush(); }
    int *ol = out.out_len + 4 * (size_t) t.pair;
    ol[0] = nmed; ol[1] = n; ol[2] = n; ol[3] = n;
}

}  // namespace poyb200
