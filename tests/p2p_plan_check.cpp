// Host-side check of p2p_plan_from_rows (csrc/p2p_exchange.cuh), the plan the experimental peer-memory exchange derives
// on the device: reads "W me" and W rows of W + 2 offsets from stdin, prints the plan as one line of integers. Compiled
// and driven by tests/test_multirank_gloo.py::test_p2p_plan_matches_deque_plan (g++, no GPU involved).
#include <cstdio>
#include <vector>
#include "p2p_exchange.cuh"

int main() {
    int W, me;
    if (scanf("%d %d", &W, &me) != 2 || W < 1 || W > dprt::kP2PMaxWorld) return 2;
    std::vector<int32_t> rows((size_t)W * (W + 2));
    for (auto& v : rows) if (scanf("%d", &v) != 1) return 2;
    dprt::P2PPlan plan{};
    dprt::p2p_plan_from_rows(rows.data(), W + 2, W, me, &plan);
    for (int d = 0; d < W; d++) printf("%d %d %d ", plan.sendCnt[d], plan.recvCnt[d], plan.dstOffset[d]);
    printf("%d %d %d %d %d %d %d\n", plan.offL, plan.cL, plan.offR, plan.cR, plan.newNL, plan.newActive, plan.allLocal);
    return 0;
}
