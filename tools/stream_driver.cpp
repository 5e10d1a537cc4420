// Load generator for the REFERENCE entry points (april_api.h only: aas_create_session / aas_feed_pcm16 / aas_flush / aas_free),
// used by bench.py and the -m gpu tests: N native threads, one synchronous session each, exactly the usage
// april-docs/src/concepts.md:41-58 documents. Bench/test tooling, not part of libaprilasr.so; it calls nothing but the public ABI,
// so it drives the oracle library just as well.
//   asd_stream: BASELINE configs[3] — every thread feeds one `chunk` per tick (all threads released together, as 40 ms of audio
//               arrives for every live stream at once) and records (aas_feed_pcm16 returned - entered): the handler of a
//               synchronous session runs before the call returns, so this bounds "handler invoked - feed entered" from above.
//   asd_bulk:   every thread feeds its whole clip in `chunk`-sample calls and flushes; returns the wall time.
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

#include "../include/april_api.h"

// the four entry points under test, handed over as pointers (ctypes: cast(lib.aas_feed_pcm16, c_void_p)) so that the same driver
// binary runs against libaprilasr.so or the oracle library without being linked to either
struct AsdApi {
    AprilASRSession (*create_session)(AprilASRModel, AprilConfig);
    void (*feed_pcm16)(AprilASRSession, short*, size_t);
    void (*flush)(AprilASRSession);
    void (*free_session)(AprilASRSession);
};

namespace {
struct TickBarrier {
    std::mutex mu; std::condition_variable cv; int n, waiting = 0; unsigned gen = 0;
    explicit TickBarrier(int n_) : n(n_) {}
    void wait() {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned g = gen;
        if (++waiting == n) { waiting = 0; gen++; cv.notify_all(); }
        else cv.wait(lk, [&] { return gen != g; });
    }
};
struct Counters { unsigned long long partial = 0, final_ = 0, silence = 0, final_tokens = 0, errors = 0; };
void counting(void* ud, AprilResultType t, size_t count, const AprilToken*) {
    Counters* c = static_cast<Counters*>(ud);
    if (t == APRIL_RESULT_RECOGNITION_PARTIAL) c->partial++;
    else if (t == APRIL_RESULT_RECOGNITION_FINAL) { c->final_++; c->final_tokens += count; }
    else if (t == APRIL_RESULT_SILENCE) c->silence++;
    else c->errors++;
}
double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
}  // namespace

extern "C" {

// lat_ms: [ticks * n_sessions] (tick-major). counters: [5] summed over sessions {partial, final, silence, final tokens, errors}.
__attribute__((visibility("default")))
int asd_stream(const AsdApi* api, AprilASRModel model, int n_sessions, const short* const* pcm, size_t n_samples, int chunk, int warm_ticks, int ticks, int pace_us,
               double* lat_ms, unsigned long long* counters) {
    if (!api || !model || n_sessions <= 0 || !pcm || chunk <= 0 || ticks <= 0 || (size_t)chunk > n_samples) return -1;
    std::vector<Counters> cnt((size_t)n_sessions);
    std::vector<AprilASRSession> ss((size_t)n_sessions, nullptr);
    for (int i = 0; i < n_sessions; i++) {
        AprilConfig cfg; memset(&cfg, 0, sizeof cfg);
        cfg.handler = counting; cfg.userdata = &cnt[(size_t)i]; cfg.flags = APRIL_CONFIG_FLAG_ZERO_BIT;
        ss[(size_t)i] = api->create_session(model, cfg);
        if (!ss[(size_t)i]) return -2;
    }
    TickBarrier bar(n_sessions);
    std::vector<std::thread> th;
    for (int i = 0; i < n_sessions; i++)
        th.emplace_back([&, i] {
            const size_t span = n_samples - (size_t)chunk;
            for (int k = 0; k < warm_ticks + ticks; k++) {
                const size_t off = span ? ((size_t)k * (size_t)chunk) % span : 0;
                bar.wait();
                const double t0 = now_ms();
                api->feed_pcm16(ss[(size_t)i], const_cast<short*>(pcm[i] + off), (size_t)chunk);
                const double t1 = now_ms();
                if (k >= warm_ticks) lat_ms[(size_t)(k - warm_ticks) * (size_t)n_sessions + (size_t)i] = t1 - t0;
                if (pace_us > 0) std::this_thread::sleep_for(std::chrono::microseconds(pace_us));
            }
        });
    for (auto& t : th) t.join();
    for (auto s : ss) api->free_session(s);
    if (counters) { memset(counters, 0, 5 * sizeof *counters); for (auto& c : cnt) { counters[0] += c.partial; counters[1] += c.final_; counters[2] += c.silence; counters[3] += c.final_tokens; counters[4] += c.errors; } }
    return 0;
}

// Returns wall milliseconds (first feed entered -> last flush returned), or a negative error.
__attribute__((visibility("default")))
double asd_bulk(const AsdApi* api, AprilASRModel model, int n_sessions, const short* const* pcm, const size_t* n_samples, int chunk, unsigned long long* counters) {
    if (!api || !model || n_sessions <= 0 || !pcm || !n_samples) return -1.0;
    std::vector<Counters> cnt((size_t)n_sessions);
    std::vector<AprilASRSession> ss((size_t)n_sessions, nullptr);
    for (int i = 0; i < n_sessions; i++) {
        AprilConfig cfg; memset(&cfg, 0, sizeof cfg);
        cfg.handler = counting; cfg.userdata = &cnt[(size_t)i]; cfg.flags = APRIL_CONFIG_FLAG_ZERO_BIT;
        ss[(size_t)i] = api->create_session(model, cfg);
        if (!ss[(size_t)i]) return -2.0;
    }
    TickBarrier bar(n_sessions + 1);
    std::vector<std::thread> th;
    for (int i = 0; i < n_sessions; i++)
        th.emplace_back([&, i] {
            bar.wait();
            const size_t n = n_samples[i], step = chunk > 0 ? (size_t)chunk : (n ? n : 1);
            for (size_t off = 0; off < n; off += step) api->feed_pcm16(ss[(size_t)i], const_cast<short*>(pcm[i] + off), std::min(step, n - off));
            api->flush(ss[(size_t)i]);
        });
    bar.wait();
    const double t0 = now_ms();
    for (auto& t : th) t.join();
    const double wall = now_ms() - t0;
    for (auto s : ss) api->free_session(s);
    if (counters) { memset(counters, 0, 5 * sizeof *counters); for (auto& c : cnt) { counters[0] += c.partial; counters[1] += c.final_; counters[2] += c.silence; counters[3] += c.final_tokens; counters[4] += c.errors; } }
    return wall;
}

}  // extern "C"
