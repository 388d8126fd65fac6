// tests/emul only (compiled with -DFMB_EMULATE): run a "kernel launch" on host threads with a real barrier.
#ifdef FMB_EMULATE
#include <barrier>
#include <thread>
#include <vector>

#include "common.h"

namespace fmb {

void emulate_launch(long long tiles, int nt, size_t smem,
                    const std::function<void(long long, int, int, void *, const std::function<void()> &)> &body) {
    std::vector<unsigned char> sm(smem + 64);
    std::barrier<> bar(nt);
    std::function<void()> sync = [&bar]() { bar.arrive_and_wait(); };
    std::vector<std::thread> threads;
    threads.reserve(nt);
    for (int tid = 0; tid < nt; ++tid) {
        threads.emplace_back([&, tid]() {
            for (long long tile = 0; tile < tiles; ++tile) {
                body(tile, tid, nt, sm.data(), sync);
                bar.arrive_and_wait();          // the next tile reuses the shared buffer
            }
        });
    }
    for (auto &t : threads) t.join();
}

}  // namespace fmb
#endif
