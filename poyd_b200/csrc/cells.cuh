// Per-cell recurrences, shared by the generic (state in memory) and stripe (state in registers) kernels.
#pragma once
#include "common.cuh"

namespace poyb200 {

// ---------------------------------------------------------------------------------------------------
// affine_3, src/algn.c:1693-2548
// ---------------------------------------------------------------------------------------------------

// Quantities that depend on the column j only (src/algn.c:2452-2458, 2500-2502).
struct AffCol {
    int hx;    // sj_horizontal_extension[j]
    int gopg;  // gap_open_prec[j] + gap_row[j]   (opening a horizontal gap at j)
    int gop;   // gap_open_prec[j]
    int lut;   // (sj[j] & 15)
    int t16;   // sj[j] & TMPGAP            (0 / 16)
    int ext;   // sj[j] != (sj[j] & 15)     (0 / 1)   "sj_base != sj_no_gap" :1947
};
// Quantities that depend on the row i only (src/algn.c:2478-2495).
struct AffRow {
    int vx;     // si_vertical_extension
    int gopge;  // si_gap_opening + si_gap_extension
    int gop;    // si_gap_opening
    int lut;    // (si[i] & 15)
    int t16, ext;
};

__device__ __forceinline__ AffCol aff_make_col(int cj, int pj, int j, int gap, int go, int g) {
    AffCol c;
    c.gop = (!(pj & gap) && (cj & gap)) ? 0 : go;  // HAS_GAP_OPENING :1730
    c.hx = ((pj & gap) && !(cj & gap)) ? c.gop + g : g;
    if (j == 1) c.hx = g;  // :2458
    c.gopg = c.gop + g;
    c.lut = cj & 15;
    c.t16 = cj & TMPGAP;
    c.ext = (cj != (cj & 15));
    return c;
}

__device__ __forceinline__ AffRow aff_make_row(int ci, int pi, int i, int gap, int go, int ge) {
    AffRow r;
    r.gop = (!(pi & gap) && (ci & gap)) ? 0 : go;
    r.vx = (i > 1 && (pi & gap) && !(ci & gap)) ? r.gop + ge : ge;  // :2483-2485
    r.gopge = r.gop + ge;
    r.lut = ci & 15;
    r.t16 = ci & TMPGAP;
    r.ext = (ci != (ci & 15));
    return r;
}

// One interior cell.  Inputs: left (ehl, cbl), up (evu, cbu), diagonal (cbd, evd, ehd, ebd), the row and
// column parameters, dcost = cost[(si&15) << lcm | (sj&15)].  Returns the direction byte (see common.cuh);
// BT = false gives the cost-only recurrence whose block-diagonal opening differs (:1848 vs :1872).
template <bool BT>
__device__ __forceinline__ int aff_cell(int ehl, int cbl, int evu, int cbu, int cbd, int evd, int ehd, int ebd,
                                        const AffRow &r, const AffCol &c, int dcost, int go, int &cb, int &ev,
                                        int &eh, int &eb) {
    int byte;
    // FILL_EXTEND_HORIZONTAL :1765-1787 -- extend wins only when strictly cheaper
    int x = ehl + c.hx, y = cbl + c.gopg;
    eh = min(x, y);
    byte = (x < y) ? 0 : AB_ENDH;
    // FILL_EXTEND_VERTICAL :1813-1830
    x = evu + r.vx;
    y = cbu + r.gopge;
    ev = min(x, y);
    byte |= (x < y) ? 0 : AB_ENDV;
    // FILL_EXTEND_BLOCK_DIAGONAL :1861-1882 (_NOBT :1837-1854: opening costs 2*go when both carry a gap;
    // its `flag2` needs !(sj & 16) and so is never true together with `flag`)
    const bool both = (r.t16 != 0) && (c.t16 != 0);
    const int dg = both ? 0 : HIGH_NUM;
    const int odg = BT ? dg : (both ? 2 * go : HIGH_NUM);
    x = ebd + dg;
    y = cbd + odg;
    eb = min(x, y);
    byte |= (x < y) ? 0 : AB_ENDB;
    // FILL_CLOSE_BLOCK_DIAGONAL :1923-1977
    const int a0 = cbd + dcost;
    const int a1 = evd + dcost + (r.ext ? c.gop : 0);
    const int a2 = ehd + dcost + (c.ext ? r.gop : 0);
    const int a3 = ebd + dcost + max(r.gop, c.gop);
    cb = min(min(a0, a1), min(a2, a3));
    if (BT) {
        // mask keeps every minimum; backtrace_affine reads it with priority H > D > V (:2049-2051)
        const int nxt = (a2 == cb) ? AN_H : (a3 == cb) ? AN_D : (a1 == cb) ? AN_V : AN_A;
        // ASSIGN_MINIMUM :2251-2280, read with priority H > A > V > D (:2006-2012)
        const int f = min(min(eh, ev), min(eb, cb));
        const int mode = (eh == f) ? AM_H : (cb == f) ? AM_A : (ev == f) ? AM_V : AM_D;
        byte |= (mode << 2) | nxt;
    }
    return byte;
}

// ---------------------------------------------------------------------------------------------------
// linear gaps, src/algn.c:375-431 (+ the edge rules folded into "neighbour = LIN_INF")
// ---------------------------------------------------------------------------------------------------
__device__ __forceinline__ int lin_cell(int ml, int mu, int md, int c_al, int c_ins, int c_del, int &mask) {
    const int t_al = md + c_al, t_ins = ml + c_ins, t_del = mu + c_del;
    const int v = min(t_al, min(t_ins, t_del));
    mask = ((t_al == v) ? D_ALIGN : 0) | ((t_ins == v) ? D_INSERT : 0) | ((t_del == v) ? D_DELETE : 0);
    return v;
}

// algn_fill_last_column :548-560
__device__ __forceinline__ int lin_last_column(int v, int mu, int tail_a, int &mask) {
    const int cst = tail_a + mu;
    if (cst < v) {
        mask = D_DELETE;
        return cst;
    }
    if (cst == v) mask |= D_DELETE;
    return v;
}

}  // namespace poyb200
