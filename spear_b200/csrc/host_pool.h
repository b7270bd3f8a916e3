// host_pool.h -- a few persistent worker threads for the host passes that are per-system or per-effect independent
// (Context::update's phase 1 over 100k active effects, Context::renderBatch's draw records, BillboardSystem::add's
// conversion). The threads are created on first use and parked on a condition variable between jobs: creating and
// joining std::threads per call costs more (~20 us per thread) than the C5 frame's passes gain.
#pragma once

#include <stdint.h>
#include <stdlib.h>

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <exception>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

namespace spr {

class HostPool {
public:
	~HostPool()
	{
		{ std::lock_guard<std::mutex> lock(m); quit = true; }
		wake.notify_all();
		for (std::thread &t : threads) t.join();
	}

	// Threads a large host pass may use: SPR_HOST_THREADS (SPR_RENDER_THREADS, the older name, is still read), else
	// 8 on a box with >= 16 hardware threads, 4 with >= 8, 1 otherwise. 1 keeps everything on the calling thread.
	static uint32_t maxThreads()
	{
		static const uint32_t n = [] {
			const char *e = getenv("SPR_HOST_THREADS");
			if (!e) e = getenv("SPR_RENDER_THREADS");
			if (e) { int v = atoi(e); return (uint32_t)(v < 1 ? 1 : v > 64 ? 64 : v); }
			unsigned hw = std::thread::hardware_concurrency();
			return hw >= 16 ? 8u : hw >= 8 ? 4u : 1u;
		}();
		return n;
	}

	// fn(t) for t in [0, nt): t = 0 on the calling thread, the rest on the workers; returns when all are done and
	// rethrows the first exception any of them threw.
	void run(uint32_t nt, const std::function<void(uint32_t)> &fn)
	{
		if (nt <= 1) { fn(0); return; }
		std::lock_guard<std::mutex> serial(runLock); // one job at a time (systems of one context may be driven from several threads)
		{
			std::lock_guard<std::mutex> lock(m);
			while (threads.size() < nt - 1) { const uint32_t idx = (uint32_t)threads.size() + 1; threads.emplace_back([this, idx] { workerLoop(idx); }); }
			job = &fn; jobThreads = nt; failure = nullptr;
			remaining.store(nt - 1, std::memory_order_relaxed);
			generation++;
			generationSeen.store(generation, std::memory_order_release);
		}
		wake.notify_all();
		try { fn(0); } catch (...) { std::lock_guard<std::mutex> lock(m); if (!failure) failure = std::current_exception(); }
		// the workers' share is as long as the caller's: spin, then yield
		for (uint32_t spins = 0; remaining.load(std::memory_order_acquire) != 0; spins++) if (spins > 2000) std::this_thread::yield();
		std::exception_ptr f;
		{ std::lock_guard<std::mutex> lock(m); f = failure; job = nullptr; }
		if (f) std::rethrow_exception(f);
	}

private:
	void workerLoop(uint32_t idx)
	{
		uint64_t seen = 0;
		for (;;) {
			const std::function<void(uint32_t)> *fn = nullptr;
			// A frame loop hands out a job every few hundred microseconds: poll for the next one for a while before going to
			// sleep on the condition variable (a futex wake costs tens of microseconds, more under a sandboxed kernel)
			if (spinMicros()) {
				const auto until = std::chrono::steady_clock::now() + std::chrono::microseconds(spinMicros());
				while (generationSeen.load(std::memory_order_acquire) == seen && std::chrono::steady_clock::now() < until) { }
			}
			{
				std::unique_lock<std::mutex> lock(m);
				wake.wait(lock, [&] { return quit || generation != seen; });
				if (quit) return;
				seen = generation;
				if (idx < jobThreads) fn = job;
			}
			if (!fn) continue;
			try { (*fn)(idx); } catch (...) { std::lock_guard<std::mutex> lock(m); if (!failure) failure = std::current_exception(); }
			remaining.fetch_sub(1, std::memory_order_release);
		}
	}

	static uint32_t spinMicros()
	{
		static const uint32_t us = [] { const char *e = getenv("SPR_HOST_SPIN_US"); return e ? (uint32_t)atoi(e) : 0u; }();
		return us;
	}

	std::atomic<uint64_t> generationSeen { 0 };
	std::vector<std::thread> threads;
	std::mutex m, runLock;
	std::condition_variable wake;
	const std::function<void(uint32_t)> *job = nullptr;
	uint32_t jobThreads = 0;
	uint64_t generation = 0;
	std::atomic<uint32_t> remaining { 0 };
	std::exception_ptr failure;
	bool quit = false;
};

} // namespace spr
