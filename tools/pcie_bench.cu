// Staging roofline microbench (SURVEY.md section 8d "Staging: PCIe H2D ... measure pinned cudaMemcpyAsync on the box").
//
//   nvcc -O2 -std=c++17 -o tools/_build/pcie_bench tools/pcie_bench.cu -lpthread
//   tools/_build/pcie_bench [--mb 147.456] [--reps 20] [--gpus N]
//
// For 1 GPU alone and for all N GPUs at the same time (one host thread + one pinned buffer per GPU):
//   h2d      pinned host -> device, cudaMemcpyAsync of one batch-sized buffer (default 147.456 MB = 256 v2.4 segments)
//   d2h      device -> pinned host, 6.7 MB (256 x 6522 logits) per copy
//   bidir    both directions at once on two streams
//   pack     pageable -> pinned memcpy with T threads (the host gather of batch_context.rs:199-211): the host-DRAM
//            side of a true drop-in call with pageable slices
// Prints one JSON object.  Nothing here touches the engine; it bounds what any staging scheme on this box can do.
#include <cuda_runtime.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(2); } } while (0)

using clk = std::chrono::steady_clock;
static double now_s() { return std::chrono::duration<double>(clk::now().time_since_epoch()).count(); }

struct Barrier {
    std::atomic<int> count{0};
    std::atomic<int> gen{0};
    int n;
    explicit Barrier(int n_) : n(n_) {}
    void wait() {
        const int g = gen.load();
        if (count.fetch_add(1) + 1 == n) { count.store(0); gen.fetch_add(1); }
        else while (gen.load() == g) std::this_thread::yield();
    }
};

struct Result { double h2d = 0, d2h = 0, bi_h2d = 0, bi_d2h = 0; };

// every GPU in `devs` runs the same copies concurrently; returns the per-GPU rates (GB/s) and the wall-clock aggregate
static void run_set(const std::vector<int>& devs, size_t in_bytes, size_t out_bytes, int reps, std::vector<Result>& res, Result& agg) {
    const int n = (int)devs.size();
    res.assign(n, Result{});
    Barrier bar(n);
    std::vector<double> t_h2d(n), t_d2h(n), t_bi(n);
    std::vector<std::thread> th;
    for (int i = 0; i < n; ++i)
        th.emplace_back([&, i] {
            CK(cudaSetDevice(devs[i]));
            void *h_in, *h_out, *d_in, *d_out;
            CK(cudaHostAlloc(&h_in, in_bytes, cudaHostAllocDefault));
            CK(cudaHostAlloc(&h_out, out_bytes, cudaHostAllocDefault));
            memset(h_in, 1, in_bytes);
            CK(cudaMalloc(&d_in, in_bytes));
            CK(cudaMalloc(&d_out, out_bytes));
            cudaStream_t s0, s1;
            CK(cudaStreamCreateWithFlags(&s0, cudaStreamNonBlocking));
            CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
            for (int w = 0; w < 3; ++w) { CK(cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyHostToDevice, s0)); CK(cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, s1)); }
            CK(cudaDeviceSynchronize());
            bar.wait();
            double t0 = now_s();
            for (int r = 0; r < reps; ++r) CK(cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyHostToDevice, s0));
            CK(cudaStreamSynchronize(s0));
            t_h2d[i] = now_s() - t0;
            bar.wait();
            t0 = now_s();
            for (int r = 0; r < reps * 8; ++r) CK(cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, s1));
            CK(cudaStreamSynchronize(s1));
            t_d2h[i] = now_s() - t0;
            bar.wait();
            t0 = now_s();
            for (int r = 0; r < reps; ++r) {
                CK(cudaMemcpyAsync(d_in, h_in, in_bytes, cudaMemcpyHostToDevice, s0));
                CK(cudaMemcpyAsync(h_out, d_out, out_bytes, cudaMemcpyDeviceToHost, s1));
            }
            CK(cudaStreamSynchronize(s0));
            CK(cudaStreamSynchronize(s1));
            t_bi[i] = now_s() - t0;
            bar.wait();
            cudaFreeHost(h_in); cudaFreeHost(h_out); cudaFree(d_in); cudaFree(d_out);
            cudaStreamDestroy(s0); cudaStreamDestroy(s1);
        });
    for (auto& t : th) t.join();
    double mx_h = 0, mx_d = 0, mx_b = 0;
    for (int i = 0; i < n; ++i) {
        res[i].h2d = reps * (double)in_bytes / t_h2d[i] / 1e9;
        res[i].d2h = reps * 8 * (double)out_bytes / t_d2h[i] / 1e9;
        res[i].bi_h2d = reps * (double)in_bytes / t_bi[i] / 1e9;
        res[i].bi_d2h = reps * (double)out_bytes / t_bi[i] / 1e9;
        mx_h = std::max(mx_h, t_h2d[i]); mx_d = std::max(mx_d, t_d2h[i]); mx_b = std::max(mx_b, t_bi[i]);
    }
    agg.h2d = n * reps * (double)in_bytes / mx_h / 1e9;          // all GPUs' bytes over the slowest GPU's time
    agg.d2h = n * reps * 8 * (double)out_bytes / mx_d / 1e9;
    agg.bi_h2d = n * reps * (double)in_bytes / mx_b / 1e9;
    agg.bi_d2h = n * reps * (double)out_bytes / mx_b / 1e9;
}

static double pack_rate(size_t bytes, int threads, int reps) {
    std::vector<char> src(bytes, 1);
    void* dst;
    CK(cudaHostAlloc(&dst, bytes, cudaHostAllocDefault));
    memset(dst, 0, bytes);
    double best = 0;
    for (int r = 0; r < reps; ++r) {
        const double t0 = now_s();
        std::vector<std::thread> th;
        const size_t per = (bytes / threads + 4095) & ~(size_t)4095;
        for (int t = 0; t < threads; ++t)
            th.emplace_back([&, t] {
                const size_t lo = std::min(bytes, t * per), hi = std::min(bytes, lo + per);
                memcpy((char*)dst + lo, src.data() + lo, hi - lo);
            });
        for (auto& t : th) t.join();
        best = std::max(best, (double)bytes / (now_s() - t0) / 1e9);
    }
    cudaFreeHost(dst);
    return best;
}

int main(int argc, char** argv) {
    double mb = 147.456;
    int reps = 20, want = 0;
    for (int i = 1; i + 1 < argc; i += 2) {
        if (!strcmp(argv[i], "--mb")) mb = atof(argv[i + 1]);
        else if (!strcmp(argv[i], "--reps")) reps = atoi(argv[i + 1]);
        else if (!strcmp(argv[i], "--gpus")) want = atoi(argv[i + 1]);
    }
    int ndev = 0;
    CK(cudaGetDeviceCount(&ndev));
    if (want > 0 && want < ndev) ndev = want;
    const size_t in_bytes = (size_t)(mb * 1e6), out_bytes = (size_t)256 * 6522 * 4;
    printf("{\"in_bytes\": %zu, \"out_bytes\": %zu, \"reps\": %d, \"host_threads\": %u, \"sets\": [", in_bytes, out_bytes, reps, std::thread::hardware_concurrency());
    std::vector<std::vector<int>> sets;
    sets.push_back({0});
    for (int n = 2; n <= ndev; n *= 2) { std::vector<int> d; for (int i = 0; i < n; ++i) d.push_back(i); sets.push_back(d); }
    if (ndev > 1 && (ndev & (ndev - 1))) { std::vector<int> d; for (int i = 0; i < ndev; ++i) d.push_back(i); sets.push_back(d); }
    for (size_t s = 0; s < sets.size(); ++s) {
        std::vector<Result> res;
        Result agg;
        run_set(sets[s], in_bytes, out_bytes, reps, res, agg);
        double mn = 1e30, mxv = 0;
        for (auto& r : res) { mn = std::min(mn, r.h2d); mxv = std::max(mxv, r.h2d); }
        printf("%s{\"gpus\": %zu, \"h2d_gbs_aggregate\": %.2f, \"h2d_gbs_per_gpu_min\": %.2f, \"h2d_gbs_per_gpu_max\": %.2f, "
               "\"d2h_gbs_aggregate\": %.2f, \"bidir_h2d_gbs_aggregate\": %.2f, \"bidir_d2h_gbs_aggregate\": %.2f}",
               s ? ", " : "", sets[s].size(), agg.h2d, mn, mxv, agg.d2h, agg.bi_h2d, agg.bi_d2h);
        fflush(stdout);
    }
    printf("], \"pack_gbs\": {");
    const int tl[] = {1, 2, 4, 8, 16, 32};
    bool first = true;
    for (int t : tl) {
        if (t > (int)std::thread::hardware_concurrency()) break;
        printf("%s\"%d\": %.2f", first ? "" : ", ", t, pack_rate(in_bytes, t, 5));
        first = false;
    }
    printf("}}\n");
    return 0;
}
