// Host thread pool (see st_hostpool.cuh).  Plain C++: no device code here.
#include "st_hostpool.cuh"

#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

namespace {

struct Pool {
    std::mutex mu;                 // protects the fields below
    std::condition_variable cv_work, cv_done;
    std::vector<std::thread> workers;
    const std::function<void(int, int)> *fn = nullptr;
    int n_parts = 0, next_part = 0, pending = 0;
    uint64_t generation = 0;
    bool stop = false;
    std::mutex call_mu;            // one parallel_for at a time
    int n_threads = 1;

    Pool() {
        int hw = int(std::thread::hardware_concurrency());
#if defined(__linux__)
        cpu_set_t set;
        if (sched_getaffinity(0, sizeof(set), &set) == 0) hw = std::min(hw > 0 ? hw : 1 << 20, CPU_COUNT(&set));
#endif
        if (hw < 1) hw = 1;
        // one process per GPU (torchrun): share the cores between the local ranks
        if (const char *e = getenv("LOCAL_WORLD_SIZE")) {
            const int lw = atoi(e);
            if (lw > 1) hw = std::max(1, hw / lw);
        }
        n_threads = std::min(hw, 16);
        if (const char *e = getenv("SUCHTREE_B200_HOST_THREADS")) {
            int v = atoi(e);
            if (v >= 1 && v <= 256) n_threads = v;
        }
        for (int i = 1; i < n_threads; ++i) workers.emplace_back([this] { loop(); });
    }
    ~Pool() {
        {
            std::lock_guard<std::mutex> l(mu);
            stop = true;
        }
        cv_work.notify_all();
        for (auto &t : workers) t.join();
    }
    void run_parts(std::unique_lock<std::mutex> &l) {
        while (next_part < n_parts) {
            const int p = next_part++;
            const auto *f = fn;
            const int np = n_parts;
            l.unlock();
            (*f)(p, np);
            l.lock();
            if (--pending == 0) cv_done.notify_all();
        }
    }
    void loop() {
        std::unique_lock<std::mutex> l(mu);
        uint64_t seen = 0;
        for (;;) {
            cv_work.wait(l, [&] { return stop || generation != seen; });
            if (stop) return;
            seen = generation;
            run_parts(l);
        }
    }
    void parallel_for(int parts, const std::function<void(int, int)> &f) {
        if (parts <= 1 || n_threads <= 1) {
            for (int p = 0; p < parts; ++p) f(p, parts);
            return;
        }
        std::lock_guard<std::mutex> call(call_mu);
        std::unique_lock<std::mutex> l(mu);
        fn = &f;
        n_parts = parts;
        next_part = 0;
        pending = parts;
        ++generation;
        cv_work.notify_all();
        run_parts(l);  // the caller works too
        cv_done.wait(l, [&] { return pending == 0; });
        fn = nullptr;
    }
};

Pool &pool() {
    static Pool *p = new Pool();  // leaked on purpose: no static-destruction order issues at exit
    return *p;
}

}  // namespace

void st_parallel_for(int n_parts, const std::function<void(int, int)> &fn) { pool().parallel_for(n_parts, fn); }
int st_host_threads() { return pool().n_threads; }
