// flye_b200 host mirror — processInParallel with the reference's signature (src/common/parallel.h:14-58):
// min(maxThreads, tasks) std::threads pulling indices from one atomic counter.
#pragma once
#include <algorithm>
#include <atomic>
#include <functional>
#include <thread>
#include <vector>

template <class T>
void processInParallel(const std::vector<T>& scheduledTasks, std::function<void(const T&)> updateFun, size_t maxThreads,
                       bool /*progressBar*/) {
    if (scheduledTasks.empty()) return;
    std::atomic<size_t> next(0);
    auto worker = [&]() {
        for (size_t i = next++; i < scheduledTasks.size(); i = next++) updateFun(scheduledTasks[i]);
    };
    std::vector<std::thread> pool(std::min(std::max<size_t>(maxThreads, 1), scheduledTasks.size()));
    for (auto& t : pool) t = std::thread(worker);
    for (auto& t : pool) t.join();
}
