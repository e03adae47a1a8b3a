// Host-side check of p2p_plan_numbers (csrc/p2p_exchange.cuh), the plan the peer-memory exchange derives on the device from
// the gathered histogram rows: reads "W me" and W rows of W + 1 bucket counts from stdin, prints the plan as one line of
// integers. Compiled and driven by tests/test_multirank_gloo.py::test_p2p_plan_matches_deque_plan (g++, no GPU involved).
#include <cstdio>
#include <vector>
#include "p2p_exchange.cuh"

int main() {
    int W, me;
    if (scanf("%d %d", &W, &me) != 2 || W < 1 || W > dprt::kP2PMaxWorld) return 2;
    std::vector<int32_t> cnt((size_t)W * 32, 0);
    for (int s = 0; s < W; s++) for (int b = 0; b <= W; b++) if (scanf("%d", &cnt[(size_t)s * 32 + b]) != 1) return 2;
    dprt::P2PPlanNumbers pn{};
    dprt::p2p_plan_numbers(cnt.data(), 32, W, me, &pn);
    for (int d = 0; d < W; d++) printf("%d ", pn.dstOffset[d]);
    printf("%d %d %d %d %d %d %d %d\n", pn.cL, pn.cR, pn.newNL, pn.newActive, pn.allLocal, pn.sent, pn.total, pn.maxArrivals);
    return 0;
}
