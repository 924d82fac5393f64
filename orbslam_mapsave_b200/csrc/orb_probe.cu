// orb_probe.cu — platform probes behind the C ABI (measurement aids, not on the product path).
// orb_h2d_probe: the host->device rate the box can sustain when n GPUs stream frames at once — the floor under the end-to-end
// (host-buffer) numbers of bench.py.  One thread per GPU, plain cudaMemcpyAsync per chunk from pinned memory (no batched-copy API),
// all threads released together, wall clock from the common start to the last completion.
#include "orb_common.cuh"

#include <sched.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

static int gpu_numa_node(int dev, char* bdf_out) {
    char bdf[32] = "";
    if (cudaDeviceGetPCIBusId(bdf, sizeof(bdf), dev) != cudaSuccess) { cudaGetLastError(); return -2; }
    for (char* p = bdf; *p; ++p) if (*p >= 'A' && *p <= 'F') *p = (char)(*p - 'A' + 'a');
    if (bdf_out) strcpy(bdf_out, bdf);
    const std::string path = std::string("/sys/bus/pci/devices/") + bdf + "/numa_node";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return -3;                                            // sysfs entry not visible (container / VM)
    int node = -1;
    if (fscanf(f, "%d", &node) != 1) node = -4;
    fclose(f);
    return node;
}

static bool bind_thread_to_node(int node) {
    if (node < 0) return false;
    const std::string path = "/sys/devices/system/node/node" + std::to_string(node) + "/cpulist";
    FILE* f = fopen(path.c_str(), "r");
    if (!f) return false;
    char buf[4096] = "";
    const bool got = fgets(buf, sizeof(buf), f) != nullptr;
    fclose(f);
    if (!got) return false;
    cpu_set_t set;
    CPU_ZERO(&set);
    int n = 0;
    for (char* tok = strtok(buf, ",\n"); tok; tok = strtok(nullptr, ",\n")) {
        int a = 0, b = 0;
        const int k = sscanf(tok, "%d-%d", &a, &b);
        if (k == 1) b = a;
        if (k >= 1) for (int c = a; c <= b && c < CPU_SETSIZE; c++) { CPU_SET(c, &set); n++; }
    }
    return n > 0 && sched_setaffinity(0, sizeof(set), &set) == 0;
}

// flags: bit 0 = write-combined pinned memory, bit 1 = also stream 20 % of the volume device->host at the same time (the results),
// bit 2 = bind every copy thread (and therefore its first-touched pinned buffer) to the NUMA node of its GPU.
// gbs_each[n_dev]: per-GPU H2D GB/s; *gbs_total: aggregate; numa_nodes[n_dev]: sysfs numa_node of each GPU (-1 = the kernel reports
// none, -3 = sysfs entry not visible).
// orb_h2d_probe_at: the same with a common start time for SEVERAL PROCESSES (one per GPU, the layout of the e2e measurement): every
// process allocates and warms up on its own, then all of them wait for `start_unix_ns` (CLOCK_REALTIME) before the timed copies, so
// that the timed windows coincide.  *late_ms receives by how much this process missed the start (0 = on time; a late process measured
// against less contention and its figure must be discarded).  start_unix_ns = 0: start as soon as this process's threads are ready.
static int h2d_probe_impl(int n_dev, const int* devices, size_t bytes_per_step, size_t chunk_bytes, int steps, int flags,
                          long long start_unix_ns, double* gbs_each, double* gbs_total, int* numa_nodes, double* late_ms);
extern "C" int orb_h2d_probe(int n_dev, const int* devices, size_t bytes_per_step, size_t chunk_bytes, int steps, int flags,
                             double* gbs_each, double* gbs_total, int* numa_nodes) {
    return h2d_probe_impl(n_dev, devices, bytes_per_step, chunk_bytes, steps, flags, 0, gbs_each, gbs_total, numa_nodes, nullptr);
}
extern "C" int orb_h2d_probe_at(int n_dev, const int* devices, size_t bytes_per_step, size_t chunk_bytes, int steps, int flags,
                                long long start_unix_ns, double* gbs_each, double* gbs_total, int* numa_nodes, double* late_ms) {
    return h2d_probe_impl(n_dev, devices, bytes_per_step, chunk_bytes, steps, flags, start_unix_ns, gbs_each, gbs_total, numa_nodes, late_ms);
}
static int h2d_probe_impl(int n_dev, const int* devices, size_t bytes_per_step, size_t chunk_bytes, int steps, int flags,
                          long long start_unix_ns, double* gbs_each, double* gbs_total, int* numa_nodes, double* late_ms) {
    ORB_REQUIRE(n_dev > 0 && devices && bytes_per_step > 0 && chunk_bytes > 0 && steps > 0 && gbs_each && gbs_total, ORB_ERR_ARG, "bad arguments");
    ORB_REQUIRE(orb_device_count() >= n_dev, ORB_ERR_CUDA, "need %d CUDA devices (no CPU fallback)", n_dev);
    std::atomic<int> ready(0), failed(0);
    std::atomic<bool> go(false);
    std::vector<double> secs(n_dev, 0.0);
    std::vector<std::thread> th;
    for (int i = 0; i < n_dev; i++)
        th.emplace_back([&, i] {
            const int dev = devices[i];
            bool ok = cudaSetDevice(dev) == cudaSuccess;
            const int node = gpu_numa_node(dev, nullptr);
            if (numa_nodes) numa_nodes[i] = node;
            if (flags & 4) bind_thread_to_node(node);
            u8 *h = nullptr, *hout = nullptr, *d[2] = {nullptr, nullptr}, *dout = nullptr;
            cudaStream_t s = nullptr, s2 = nullptr;
            ok = ok && cudaHostAlloc(&h, bytes_per_step, (flags & 1) ? cudaHostAllocWriteCombined : cudaHostAllocDefault) == cudaSuccess;
            if (ok) memset(h, 1, bytes_per_step);                 // first touch on this thread
            ok = ok && cudaMalloc(&d[0], chunk_bytes) == cudaSuccess && cudaMalloc(&d[1], chunk_bytes) == cudaSuccess;
            ok = ok && cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) == cudaSuccess;
            const size_t outBytes = bytes_per_step / 5;
            if (flags & 2) {
                ok = ok && cudaHostAlloc(&hout, outBytes, cudaHostAllocDefault) == cudaSuccess && cudaMalloc(&dout, chunk_bytes) == cudaSuccess;
                ok = ok && cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking) == cudaSuccess;
                if (ok) memset(hout, 0, outBytes);
            }
            auto step = [&] {
                int k = 0;
                for (size_t o = 0; o < bytes_per_step && ok; o += chunk_bytes, k++) {
                    const size_t n = std::min(chunk_bytes, bytes_per_step - o);
                    ok = cudaMemcpyAsync(d[k & 1], h + o, n, cudaMemcpyHostToDevice, s) == cudaSuccess;
                    if ((flags & 2) && ok && o / 5 + n / 5 <= outBytes)
                        ok = cudaMemcpyAsync(hout + o / 5, dout, n / 5, cudaMemcpyDeviceToHost, s2) == cudaSuccess;
                }
                ok = ok && cudaStreamSynchronize(s) == cudaSuccess;
                if (flags & 2) ok = ok && cudaStreamSynchronize(s2) == cudaSuccess;
            };
            if (ok) step();                                        // warm-up
            ready.fetch_add(1);
            while (!go.load()) std::this_thread::yield();
            const auto t0 = std::chrono::steady_clock::now();
            for (int k = 0; k < steps && ok; k++) step();
            secs[i] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            if (!ok) { failed.fetch_add(1); cudaGetLastError(); }
            if (s) cudaStreamDestroy(s);
            if (s2) cudaStreamDestroy(s2);
            cudaFree(d[0]); cudaFree(d[1]); cudaFree(dout);
            if (h) cudaFreeHost(h);
            if (hout) cudaFreeHost(hout);
        });
    while (ready.load() < n_dev) std::this_thread::yield();
    if (late_ms) *late_ms = 0.0;
    if (start_unix_ns > 0) {
        auto now_ns = [] { return (long long)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::system_clock::now().time_since_epoch()).count(); };
        const long long t = now_ns();
        if (t > start_unix_ns) { if (late_ms) *late_ms = (double)(t - start_unix_ns) * 1e-6; }
        else while (now_ns() < start_unix_ns) std::this_thread::yield();
    }
    go.store(true);
    for (auto& t : th) t.join();
    ORB_REQUIRE(failed.load() == 0, ORB_ERR_CUDA, "a CUDA call failed in the copy probe");
    double worst = 0;
    for (int i = 0; i < n_dev; i++) {
        gbs_each[i] = (double)bytes_per_step * steps / secs[i] / 1e9;
        worst = std::max(worst, secs[i]);
    }
    *gbs_total = (double)bytes_per_step * steps * n_dev / worst / 1e9;
    return ORB_OK;
}
