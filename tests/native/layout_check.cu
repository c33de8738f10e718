// Host-only check of the shared-memory / workspace layouts the kernel and the host must agree on
// (altro_kernels.cuh: fixed_layout, make_layout, make_layout_big).  Built and run by tests/test_layout_host.py.
#include <algorithm>
#include <cstdio>
#include <utility>
#include <vector>

#include "../../altro_mpc_icra2021_b200/csrc/altro_kernels.cuh"

using altro::Layout;

struct Span { const char *name; long lo, hi; };

static int check(const char *what, std::vector<Span> v, long limit)
{
    int bad = 0;
    std::sort(v.begin(), v.end(), [](const Span &a, const Span &b) { return a.lo < b.lo; });
    for (size_t i = 0; i < v.size(); ++i) {
        if (v[i].lo < 0 || v[i].hi > limit) { printf("%s: %s [%ld,%ld) outside [0,%ld)\n", what, v[i].name, v[i].lo, v[i].hi, limit); ++bad; }
        if (i + 1 < v.size() && v[i].hi > v[i + 1].lo) { printf("%s: %s overlaps %s\n", what, v[i].name, v[i + 1].name); ++bad; }
    }
    return bad;
}

int main()
{
    int bad = 0, cases = 0;
    const int dims[][2] = {{6, 3}, {12, 12}, {6, 6}, {12, 3}, {12, 6}, {2, 2}, {30, 25}, {55, 2}, {64, 16}, {200, 25}};
    for (auto &d : dims)
        for (int N : {2, 15, 21, 80})
            for (int ncon : {0, 3, 7})
                for (int ref : {0, 1})
                    for (int spec : {0, 1, 3}) {
                        const int n = d[0], m = d[1], P = ncon * N * 4, EX = ncon * N * 9, ITAB = 4 * 40 + (n + n * n + m + m * m + 1) + 30;
                        const long NT = n + (long)n * n + m + (long)m * m;
                        Layout l = altro::make_layout(n, m, N, P, ncon, EX, ref, ITAB, spec);
                        const Layout f = altro::fixed_layout(n, m);
                        // the fixed part is a prefix of the full layout (compile-time offsets in the fixed-dimension kernels)
                        if (f.S != l.S || f.Qi != l.Qi || f.mu != l.mu || f.X != l.X || f.red != l.red) { printf("fixed prefix differs\n"); ++bad; }
                        std::vector<Span> v = {
                            {"Qd", l.Qd, l.Qd + n}, {"Qfd", l.Qfd, l.Qfd + n}, {"Rd", l.Rd, l.Rd + m}, {"sA", l.sA, l.sA + n * n},
                            {"sB", l.sB, l.sB + n * m}, {"sd", l.sd, l.sd + n}, {"S", l.S, l.S + n * n}, {"SA", l.SA, l.SA + n * n},
                            {"Qxx", l.Qxx, l.Qxx + n * n}, {"SB", l.SB, l.SB + n * m}, {"Qux", l.Qux, l.Qux + m * n},
                            {"T1", l.T1, l.T1 + m * n}, {"Quu", l.Quu, l.Quu + m * m}, {"L", l.L, l.L + m * m}, {"s", l.s, l.s + n},
                            {"Qx", l.Qx, l.Qx + n}, {"Qu", l.Qu, l.Qu + m}, {"t1", l.t1, l.t1 + m}, {"linv", l.linv, l.linv + m},
                            {"Qi", l.Qi, l.Qi + NT}, {"mu", l.mu, l.mu + altro::MAX_CON}, {"bc", l.bc, l.bc + 24}, {"red", l.red, l.red + 9},
                            {"X", l.X, l.X + N * n}, {"U", l.U, l.U + (N - 1) * m}, {"Xb", l.Xb, l.Xb + N * n}, {"Ub", l.Ub, l.Ub + (N - 1) * m},
                            {"K", l.K, l.K + (long)(N - 1) * m * n}, {"dv", l.dv, l.dv + (N - 1) * m}, {"lam", l.lam, l.lam + P},
                            {"ex", l.ex, l.ex + EX}, {"itm", l.itm, l.itm + N * (1 + ncon)},
                            {"cand", l.cand, l.cand + (long)spec * (N * n + (N - 1) * m + N * (1 + ncon))}};
                        if (ref) { v.push_back({"xr", l.xr, l.xr + N * n}); v.push_back({"ur", l.ur, l.ur + (N - 1) * m}); }
                        else if (l.xr != -1 || l.ur != -1) { printf("reference offsets set without ref_in_smem\n"); ++bad; }
                        bad += check("make_layout", v, l.cd);
                        const long tail = (long)l.cd * 8 + (long)std::max(ncon, 1) * sizeof(altro::ConDesc) + (long)ITAB * 4;
                        if ((l.cd & 1) || l.bytes < tail || l.bytes % 16 || l.big || l.ws_doubles) { printf("tail/alignment wrong\n"); ++bad; }
                        // workspace layout: shared-memory part and global part each disjoint
                        Layout b = altro::make_layout_big(n, m, N, P, ncon, 0);
                        std::vector<Span> sm = {
                            {"Qd", b.Qd, b.Qd + n}, {"Qfd", b.Qfd, b.Qfd + n}, {"Rd", b.Rd, b.Rd + m}, {"sd", b.sd, b.sd + n},
                            {"Quu", b.Quu, b.Quu + m * m}, {"L", b.L, b.L + m * m}, {"s", b.s, b.s + n}, {"Qx", b.Qx, b.Qx + n},
                            {"Qu", b.Qu, b.Qu + m}, {"t1", b.t1, b.t1 + m}, {"linv", b.linv, b.linv + m}, {"mu", b.mu, b.mu + altro::MAX_CON},
                            {"bc", b.bc, b.bc + 24}, {"red", b.red, b.red + 9}, {"X", b.X, b.X + N * n}, {"U", b.U, b.U + (N - 1) * m},
                            {"Xb", b.Xb, b.Xb + N * n}, {"Ub", b.Ub, b.Ub + (N - 1) * m}, {"dv", b.dv, b.dv + (N - 1) * m},
                            {"lam", b.lam, b.lam + P}, {"itm", b.itm, b.itm + N * (1 + ncon)}};
                        bad += check("make_layout_big/smem", sm, b.cd);
                        std::vector<Span> ws = {
                            {"S", b.S, b.S + n * n}, {"SA", b.SA, b.SA + n * n}, {"Qxx", b.Qxx, b.Qxx + n * n}, {"SB", b.SB, b.SB + n * m},
                            {"Qux", b.Qux, b.Qux + m * n}, {"T1", b.T1, b.T1 + m * n}, {"Qi", b.Qi, b.Qi + NT},
                            {"K", b.K, b.K + (long)(N - 1) * m * n}};
                        bad += check("make_layout_big/workspace", ws, b.ws_doubles);
                        if (!b.big || (b.cd & 1) || b.bytes % 16 || b.bytes < (long)b.cd * 8 + (long)std::max(ncon, 1) * (long)sizeof(altro::ConDesc)) { printf("big tail wrong\n"); ++bad; }
                        ++cases;
                    }
    if (sizeof(altro::ConDesc) % 16) { printf("ConDesc is not a multiple of 16 bytes\n"); ++bad; }
    printf("%d layouts checked, %d problems\n", cases, bad);
    return bad ? 1 : 0;
}
