// h2d_probe.cc — standalone (no Python) front end of orb_h2d_probe (orbslam_mapsave_b200/csrc/orb_probe.cu): the aggregate host->device
// rate of the box for 1, 2, 4 ... N GPUs streaming 640x480 frames at once, with and without concurrent device->host traffic, NUMA
// binding and write-combined staging.
// build: g++ -O2 -o tools/probes/h2d_probe tools/probes/h2d_probe.cc -Iinclude -Lorbslam_mapsave_b200 -lorb_b200 -Wl,-rpath,'$ORIGIN/../../orbslam_mapsave_b200'
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "orb_b200.h"

int main(int argc, char** argv) {
    const int maxDev = orb_device_count();
    if (maxDev <= 0) { fprintf(stderr, "no CUDA device\n"); return 1; }
    const size_t frame = 640 * 480, step = 4096 * frame, chunk = (argc > 1 ? atoi(argv[1]) : 128) * frame;
    printf("gpus  variant              total GB/s   min per-GPU GB/s   numa nodes\n");
    for (int n = 1; n <= maxDev; n *= 2) {
        std::vector<int> dev(n), nodes(n);
        for (int i = 0; i < n; i++) dev[i] = i;
        const struct { const char* name; int flags; } v[] = {{"h2d only", 0}, {"h2d + d2h", 2}, {"h2d + d2h, numa bound", 6}, {"h2d + d2h, write-combined", 3}};
        for (const auto& var : v) {
            std::vector<double> each(n);
            double tot = 0;
            if (orb_h2d_probe(n, dev.data(), step, chunk, 3, var.flags, each.data(), &tot, nodes.data()) != 0) { printf("%d %s: %s\n", n, var.name, orb_last_error()); continue; }
            double mn = each[0];
            for (double e : each) mn = e < mn ? e : mn;
            printf("%4d  %-26s %8.1f %12.1f        ", n, var.name, tot, mn);
            for (int i = 0; i < n; i++) printf("%d ", nodes[i]);
            printf("\n");
        }
    }
    return 0;
}
