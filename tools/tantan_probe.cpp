// Host probe: spb_tantan_mask on a synthetic target set of GenBank-phage shape (495 567 sequences), cold and warm, for the thread count
// given; SPB_TANTAN_SCALAR=1 selects the one-sequence-at-a-time path.  g++ -O2 -fopenmp -Iinclude tools/tantan_probe.cpp -Lspacepharer_b200/lib -lspb200
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <vector>
#include <chrono>
#include <random>
#include <omp.h>
#include "spb200.h"
int main(int argc, char **argv) {
    const uint32_t T = 495567; std::mt19937 rng(1);
    std::vector<uint64_t> off(T + 1, 0);
    for (uint32_t i = 0; i < T; i++) off[i + 1] = off[i] + 30 + rng() % 160;
    std::vector<uint8_t> res(off[T]);
    for (auto &r : res) r = rng() % 20;
    double prob[21 * 21], pb[21];
    for (int i = 0; i < 21; i++) pb[i] = 1.0 / 21;
    for (int i = 0; i < 441; i++) prob[i] = (i % 22 == 0 ? 0.02 : 0.0013);
    int th = atoi(argv[1]);
    if (argc > 2) {   // pre-create the OpenMP team with plain work
        double s = 0;
#pragma omp parallel num_threads(th) reduction(+:s)
        { for (int i = 0; i < 100000000; i++) s += i * 1e-9; }
        printf("team warmed %g\n", s);
    }
    for (int rep = 0; rep < 3; rep++) {
        std::vector<uint8_t> r2(res);
        uint64_t n = 0;
        auto t0 = std::chrono::steady_clock::now();
        spb_tantan_mask(prob, pb, r2.data(), off.data(), T, th, &n);
        printf("%.3f s masked %llu\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(), (unsigned long long)n);
    }
}
