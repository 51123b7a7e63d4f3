// Host-side harness of poyd_b200/csrc/powell_core.h (tests only): the level-synchronous restatement of Powell's 3-D
// aligner run by ONE host thread, so that tests/test_powell.py can check the algorithm the CUDA kernel executes against
// the compiled reference (oracle/_ref) on a machine without a GPU.  Not part of the product: libpoyb200.so never runs this.
#include <stdlib.h>
#include <vector>

#include "../../poyd_b200/csrc/powell_core.h"

using namespace poyb200::powell;

static int collapse(const unsigned char *s, int n, std::vector<uint8_t> &out) {  // copySequence, src/ukkCommon.c:87-108
    out.clear();
    for (int i = 1; i < n; i++) {
        const int v = s[i];
        if (v & 1) out.push_back(1);
        else if (v & 2) out.push_back(2);
        else if (v & 4) out.push_back(4);
        else if (v & 8) out.push_back(8);
        else return -1;
    }
    return 0;
}

extern "C" int pw_host_align(const unsigned char *s1, int l1, const unsigned char *s2, int l2, const unsigned char *s3, int l3, int mm,
                             int go, int ge, int R, int Wd, unsigned char *r1, unsigned char *r2, unsigned char *r3, int *rlen, int *status,
                             long long *cells /* 5 statistics, may be NULL */) {
    std::vector<uint8_t> A, B, C;
    *status = PW_EINPUT;
    *rlen = 0;
    if (collapse(s1, l1, A) || collapse(s2, l2, B) || collapse(s3, l3, C)) return -1;
    Tables tb;
    make_tables(tb, mm, go, ge);
    Work w;
    memset(&w, 0, sizeof w);
    A.push_back(0); B.push_back(0); C.push_back(0);  // readable one past the end, never equal to a base
    w.A = A.data(); w.B = B.data(); w.C = C.data();
    w.Alen = (int) A.size() - 1; w.Blen = (int) B.size() - 1; w.Clen = (int) C.size() - 1;
    w.R = R; w.D = 2 * R + 1;
    w.cab = (w.Alen - w.Blen) / 2; w.cac = (w.Alen - w.Clen) / 2;
    if (Wd <= 0) { Wd = 16; while (Wd < 4 * tb.maxSingleStep + 2 * go + 8) Wd *= 2; }
    w.Wd = Wd;
    const size_t nx = (size_t) w.D * w.D * NS;
    std::vector<Entry> U(nx * Wd);
    memset(U.data(), 0xff, U.size() * sizeof(Entry));
    std::vector<int> top(nx, NEGBIG), prev(nx, NEGBIG), snap(nx, NEGBIG);
    w.maxlevels = 2 * (w.Alen + w.Blen + w.Clen) * (ge > mm ? ge : mm) + 6 * go + 16;
    std::vector<int> keycnt(2 * (w.maxlevels + 1) + 1 + 1024), list(4 * nx);
    w.keycap = 2 * (w.maxlevels + 1) + 1;
    w.U = U.data(); w.top = top.data(); w.prev = prev.data(); w.snap = snap.data(); w.keycnt = keycnt.data(); w.list = list.data();
    w.listcap = (int) (2 * nx);
    const int cap = w.Alen + w.Blen + w.Clen + 1;
    std::vector<uint8_t> ra(cap), rb(cap), rc(cap);
    w.resA = ra.data(); w.resB = rb.data(); w.resC = rc.data(); w.rescap = cap;
    std::vector<Task> stack(256);
    w.stack = stack.data(); w.stackcap = 256;
    w.nextOffset = 1;
    Engine e;
    memset(&e, 0, sizeof e);
    e.w = &w; e.tb = &tb;
    const int cost = e.run();
    *status = w.status;
    if (cells) { cells[0] = w.ncalc; cells[1] = w.st_sweeps; cells[2] = w.st_sweep_cells; cells[3] = w.st_levels; cells[4] = w.st_tops; }
    if (w.status) return -1;
    // printTraceBack :477-495: the rows, forward, behind one gap
    r1[0] = r2[0] = r3[0] = 16;
    for (int i = 0; i < w.nres; i++) {
        const int k = w.nres - 1 - i;
        r1[1 + i] = ra[k] == 0xff ? 16 : ra[k];
        r2[1 + i] = rb[k] == 0xff ? 16 : rb[k];
        r3[1 + i] = rc[k] == 0xff ? 16 : rc[k];
    }
    *rlen = w.nres + 1;
    return cost;
}

// make_tables() laid out like the globals of ukkCommon.c (:46-59), for the test that compares them with setup()'s.
extern "C" void pw_host_tables(int mm, int go, int ge, int *neighbours, int *contCost, int *secondCost, int *transCost /* 27 x 27 */,
                               int *numStates, int *maxSingleStep) {
    Tables tb;
    make_tables(tb, mm, go, ge);
    *numStates = NS;
    *maxSingleStep = tb.maxSingleStep;
    for (int s = 0; s < NS; s++) {
        neighbours[s] = tb.da[s] + 2 * tb.db[s] + 4 * tb.dc[s];
        contCost[s] = tb.cont[s];
        secondCost[s] = tb.second[s];
        for (int t = 0; t < NS; t++) transCost[s * 27 + t] = tb.trans[s][t];
    }
}
