// C ABI of libaprilasr.so: the 12 entry points of april_api.h plus the additive batch API of
// april_b200.h. Session runtime semantics follow /root/reference/src/april_session.c (sync path
// :479-564, async path :479-493,541-545,567-585 with src/audio_provider.c's 3 s ring) — the
// forward path itself runs in the CUDA engine (engine.cu).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <mutex>
#include <set>
#include <thread>
#include <functional>
#include <vector>

#include "../../include/april_b200.h"
#include "engine.h"
#include "model_loader.h"
#include "session.h"
#include "plan.h"

using namespace april;

// ---------------------------------------------------------------------------------------------
// logging: same env var and levels as src/log.h:23-56, src/init.c:36-44
// ---------------------------------------------------------------------------------------------
enum { LEVEL_DEBUG = 0, LEVEL_INFO, LEVEL_WARNING, LEVEL_ERROR, LEVEL_NONE };
static int g_loglevel = LEVEL_WARNING;
static std::atomic<int> g_initialized{ 0 };
static thread_local int tl_device = -1;
static const char* kLevelNames[] = { "DEBUG", "INFO", "WARNING", "ERROR", "NONE" };
#define LOG_AT(level, fmt, ...) do { if ((level) >= g_loglevel) fprintf(stderr, "libapril: (%s:%d) [%s] " fmt "\n", __FILE__, __LINE__, kLevelNames[level], ##__VA_ARGS__); } while (0)
#define LOG_INFO(fmt, ...) LOG_AT(LEVEL_INFO, fmt, ##__VA_ARGS__)
#define LOG_WARNING(fmt, ...) LOG_AT(LEVEL_WARNING, fmt, ##__VA_ARGS__)
#define LOG_ERROR(fmt, ...) LOG_AT(LEVEL_ERROR, fmt, ##__VA_ARGS__)

static const size_t kRingSamples = 48000;      // MAX_AUDIO, src/audio_provider.c:31

struct AprilASRSession_i;
struct AprilASRModel_i;

struct SyncReq {                                 // one blocked aas_feed_pcm16 / aas_flush call of a synchronous session
    AprilASRSession_i* s = nullptr;
    PassItem item;
    // 0 = queued, 1 = "you lead the next pass", 2 = done. Every caller sleeps on ITS OWN word (C++20 atomic wait = one futex per
    // request): measured with one shared mutex + condition variable, 256 callers woken by a finished pass needed 1-5 ms to get
    // through the mutex one by one, which spread the next tick's arrivals over two passes.
    std::atomic<int> state{ 0 };
};

// One GPU of a model: the engine (weights replicated per device) plus the two things that build batched passes
// out of the reference's one-session-per-call API: the coalescer for synchronous callers and the batcher for async sessions.
struct Lane {
    AprilASRModel_i* model = nullptr;
    int index = 0;
    std::unique_ptr<Engine> engine;
    std::atomic<int> live_sessions{ 0 };
    // synchronous callers (april-docs/src/concepts.md:41-58: many sessions of one model driven from many threads, each call
    // blocking until its handler has run): calls that arrive while a pass is in flight, or within a short window of each
    // other, are merged into ONE device pass by whichever caller becomes the leader; every caller then replays its own
    // session's decisions (= invokes its handler) on its own thread before returning.
    std::mutex qmu;
    std::condition_variable qcv;
    std::vector<SyncReq*> queue;
    bool leader_active = false, leader_in_window = false;
    int live_sync = 0;                           // synchronous sessions placed on this lane
    int miss_streak = 0;                         // consecutive windows that caught nobody
    long last_pass_us = 0;                       // smoothed duration of this lane's coalesced passes
    // asynchronous sessions: one batcher thread per GPU (replaces one ProcThread per session, src/proc_thread.c:138-172)
    std::mutex amu;
    std::condition_variable acv;
    std::vector<AprilASRSession_i*> async_sessions;
    std::thread batcher;
    bool batcher_started = false, stop = false;
    std::mutex pass_mu;                          // held while the batcher touches sessions
};

struct AprilASRModel_i {
    std::string name, description, language;
    Params params;
    Vocab vocab;
    std::vector<std::unique_ptr<Lane>> lanes;    // one per GPU this model runs on
    std::mutex place_mu;
    long coalesce_us = 100;
};

struct AprilASRSession_i {
    AprilASRModel_i* model = nullptr;
    Lane* lane = nullptr;
    StreamFraming fr;
    std::unique_ptr<TokenShaper> shaper;
    AprilRecognitionResultHandler handler = nullptr;
    void* userdata = nullptr;
    bool sync = true;
    bool was_flushed = false;
    bool mismatch_reported = false;
    // async input ring (guarded by lane->amu)
    std::vector<int16_t> ring;
    size_t ring_count = 0;
    bool flush_requested = false;
    size_t flush_at = 0;                         // samples that were in the ring when the flush was requested: they go first
};

// ---------------------------------------------------------------------------------------------
static void replay_steps(AprilASRSession_i* s, PassItem& it, int t_from, int t_to) {
    const size_t stride = (size_t)s->lane->engine->stride_ms();
    for (int t = t_from; t < t_to; t++) {
        s->shaper->begin_step(stride);
        const int np = it.npass[(size_t)t];
        for (int p = 0; p < np; p++) {
            const StepDecision& d = it.decisions[(size_t)t * 3 + p];
            bool host_blank = s->shaper->on_joiner_pass(p == 0 ? 1.0f : 0.0f, d.max_idx, d.max_val, d.blank_val);
            if (host_blank != (d.is_blank != 0) && !s->mismatch_reported) {
                s->mismatch_reported = true;
                LOG_ERROR("internal: device/host greedy state diverged at step %lld pass %d", (long long)s->fr.steps_total, p);
            }
        }
    }
}
// what the pass decided for this session, in stream order (steps an early-steps listener has already replayed are skipped)
static void replay(AprilASRSession_i* s, PassItem& it, bool flush) {
    replay_steps(s, it, it.delivered, it.n_steps);
    it.delivered = it.n_steps;
    if (flush) s->shaper->on_flush_end();
}

// the batched entry points: sessions may live on different GPUs; each lane's pass runs concurrently, callbacks then run in index order
static void run_items(std::vector<AprilASRSession_i*>& ss, std::vector<PassItem>& items) {
    if (ss.empty()) return;
    bool one_lane = true;
    for (auto* s : ss) if (s->lane != ss[0]->lane) { one_lane = false; break; }
    if (one_lane) {
        // Big passes are computed in time chunks; the callbacks of a finished chunk run here, on the calling thread, while the device
        // works on the next one (the result replay is ~3 ms of a 64 x 30 s pass). Order: per session as always; across the sessions
        // of the call chunk by chunk, then session by session. The handlers run while the engine is locked: from THIS entry point
        // (aab_*_batch, an addition of this library) they must not call back into the library.
        const std::function<void(int, int)> early = [&](int, int t1) {
            for (size_t i = 0; i < ss.size(); i++) {
                const int upto = std::min(t1, items[i].n_steps);
                if (upto > items[i].delivered) { replay_steps(ss[i], items[i], items[i].delivered, upto); items[i].delivered = upto; }
            }
        };
        ss[0]->lane->engine->run_pass(items, &early);
    }
    else {
        std::vector<Lane*> lanes;
        for (auto* s : ss) if (std::find(lanes.begin(), lanes.end(), s->lane) == lanes.end()) lanes.push_back(s->lane);
        std::vector<std::vector<PassItem>> per(lanes.size());
        std::vector<std::vector<size_t>> idx(lanes.size());
        for (size_t i = 0; i < ss.size(); i++) {
            const size_t l = (size_t)(std::find(lanes.begin(), lanes.end(), ss[i]->lane) - lanes.begin());
            per[l].push_back(std::move(items[i])); idx[l].push_back(i);
        }
        std::vector<std::thread> th;
        for (size_t l = 0; l < lanes.size(); l++) th.emplace_back([&, l] { lanes[l]->engine->run_pass(per[l]); });
        for (auto& t : th) t.join();
        for (size_t l = 0; l < lanes.size(); l++) for (size_t k = 0; k < idx[l].size(); k++) items[idx[l][k]] = std::move(per[l][k]);
    }
    for (size_t i = 0; i < ss.size(); i++) replay(ss[i], items[i], items[i].flush);
}

// ---------------------------------------------------------------------------------------------
// synchronous call -> one request in the lane's queue; returns after this session's callbacks have run on this thread
enum { REQ_QUEUED = 0, REQ_LEAD = 1, REQ_DONE = 2 };

static void lead_one_pass(Lane& L) {
    std::unique_lock<std::mutex> lk(L.qmu);
    // Coalescing window: wait while callers keep arriving — until every live synchronous session of this GPU has a call queued,
    // or nobody new came for a quiet interval, or the cap is reached. The interval scales with what a pass costs (an eighth of
    // the last pass, at least `coalesce_us`): the threads released by one 40 ms audio tick arrive in bursts over a few hundred
    // microseconds, and one more pass for the stragglers costs far more than waiting for them; a lone caller among many idle
    // sessions pays one interval.
    const long win = L.model->coalesce_us;
    if (win > 0 && (int)L.queue.size() < L.live_sync && L.miss_streak < 4) {
        const long quiet = std::max(win, L.last_pass_us / 8), cap_us = std::max(10 * win, L.last_pass_us / 2);
        const auto cap = std::chrono::steady_clock::now() + std::chrono::microseconds(cap_us);
        L.leader_in_window = true;
        for (;;) {
            const size_t seen = L.queue.size();
            if ((int)seen >= L.live_sync) break;
            const auto now = std::chrono::steady_clock::now();
            if (now >= cap) break;
            L.qcv.wait_until(lk, std::min(cap, now + std::chrono::microseconds(quiet)));
            if (L.queue.size() == seen && std::chrono::steady_clock::now() >= now + std::chrono::microseconds(quiet)) break;   // a quiet interval
        }
        L.leader_in_window = false;
    }
    std::vector<SyncReq*> batch;
    batch.swap(L.queue);
    L.miss_streak = batch.size() > 1 ? 0 : std::min(L.miss_streak + 1, 1000);
    if (L.miss_streak >= 4 && (L.miss_streak & 63) == 0) L.miss_streak = 0;        // probe again now and then
    lk.unlock();
    std::vector<PassItem> items(batch.size());
    for (size_t i = 0; i < batch.size(); i++) items[i] = std::move(batch[i]->item);
    const auto p0 = std::chrono::steady_clock::now();
    L.engine->run_pass(items);
    const long pass_us = (long)std::chrono::duration_cast<std::chrono::microseconds>(std::chrono::steady_clock::now() - p0).count();
    for (size_t i = 0; i < batch.size(); i++) batch[i]->item = std::move(items[i]);
    lk.lock();
    L.last_pass_us = (3 * L.last_pass_us + pass_us) / 4;
    SyncReq* next = L.queue.empty() ? nullptr : L.queue.front();     // calls that queued up during the pass: the first of them leads the next one
    if (!next) L.leader_active = false;
    lk.unlock();
    if (next) { next->state.store(REQ_LEAD, std::memory_order_release); next->state.notify_one(); }
    for (auto* b : batch) { b->state.store(REQ_DONE, std::memory_order_release); b->state.notify_one(); }   // after this store the request may be gone
}

static void sync_submit(AprilASRSession_i* s, const int16_t* pcm, size_t count, bool flush) {
    Lane& L = *s->lane;
    SyncReq r;
    r.s = s;
    r.item.fr = &s->fr; r.item.cookie = s; r.item.pcm = pcm; r.item.count = count; r.item.flush = flush;
    {
        std::lock_guard<std::mutex> lk(L.qmu);
        L.queue.push_back(&r);
        if (!L.leader_active) { L.leader_active = true; r.state.store(REQ_LEAD, std::memory_order_relaxed); }
        else if (L.leader_in_window) L.qcv.notify_one();     // only a leader inside its window sleeps on the condition variable
    }
    for (;;) {
        const int st = r.state.load(std::memory_order_acquire);
        if (st == REQ_DONE) break;
        if (st == REQ_LEAD) { lead_one_pass(L); continue; }  // my own request is in the batch I run (it was queued before I became leader)
        r.state.wait(REQ_QUEUED, std::memory_order_acquire);
    }
    replay(s, r.item, flush);
}

// ---------------------------------------------------------------------------------------------
static void batcher_main(Lane* m) {
    std::vector<AprilASRSession_i*> ss;
    std::vector<PassItem> items;
    std::vector<std::vector<int16_t>> chunks;
    for (;;) {
        ss.clear(); items.clear(); chunks.clear();
        {
            std::unique_lock<std::mutex> lk(m->amu);
            m->acv.wait(lk, [&] {
                if (m->stop) return true;
                for (auto* s : m->async_sessions) if (s->ring_count || s->flush_requested) return true;
                return false;
            });
            if (m->stop) return;
        }
        std::unique_lock<std::mutex> pass_lk(m->pass_mu);     // lock order: pass_mu, then amu (aas_free does the same)
        {
            // Decided per session (the reference has one thread per session, src/april_session.c:567-585): audio that was in
            // the ring when a flush was requested goes first; a session whose flush is due is flushed in THIS pass whatever
            // the other sessions are doing; audio fed after the flush request waits for the pass after it.
            std::unique_lock<std::mutex> lk(m->amu);
            for (auto* s : m->async_sessions) {
                size_t take = s->ring_count;
                if (s->flush_requested) take = std::min(take, s->flush_at);
                if (take) {
                    ss.push_back(s);
                    chunks.emplace_back(s->ring.begin(), s->ring.begin() + (long)take);
                    if (take < s->ring_count) memmove(s->ring.data(), s->ring.data() + take, (s->ring_count - take) * sizeof(int16_t));
                    s->ring_count -= take;
                    if (s->flush_requested) s->flush_at -= take;
                    PassItem it; it.fr = &s->fr; it.cookie = s; it.pcm = chunks.back().data(); it.count = take;
                    items.push_back(std::move(it));
                    s->was_flushed = false;
                } else if (s->flush_requested) {
                    s->flush_requested = false;
                    if (!s->was_flushed) {
                        s->was_flushed = true;
                        ss.push_back(s);
                        chunks.emplace_back();
                        PassItem it; it.fr = &s->fr; it.cookie = s; it.flush = true;
                        items.push_back(std::move(it));
                    }
                }
            }
            for (size_t i = 0; i < items.size(); i++) if (!items[i].flush) items[i].pcm = chunks[i].data();   // chunks may have reallocated
        }
        if (ss.empty()) continue;
        m->engine->run_pass(items);
        for (size_t i = 0; i < ss.size(); i++) replay(ss[i], items[i], items[i].flush);
    }
}

static void ensure_batcher(Lane* m) {
    std::lock_guard<std::mutex> lk(m->amu);
    if (!m->batcher_started) { m->batcher_started = true; m->batcher = std::thread(batcher_main, m); }
}

// ---------------------------------------------------------------------------------------------
extern "C" {

APRIL_EXPORT void aam_api_init(int version) {
    if (version != APRIL_VERSION) LOG_WARNING("aam_api_init: client version %d, library version %d", version, APRIL_VERSION);
    const char* env = getenv("APRIL_LOG_LEVEL");
    if (env) for (int i = 0; i <= LEVEL_NONE; i++) if (strcmp(env, kLevelNames[i]) == 0) g_loglevel = i;
    g_initialized = 1;
}

APRIL_EXPORT int aab_device_count(void) { return (int)cuda_devices_sm100().size(); }

APRIL_EXPORT int aab_set_device(int device) {
    if (device < -1 || device >= (int)cuda_devices_sm100().size()) return -1;
    tl_device = device;                                      // -1: back to the default (all visible B200s)
    return 0;
}

APRIL_EXPORT const char* aab_engine_version(void) { return "april-b200 r2 (sm_100a; tcgen05 tier: pair GEMMs, wide-batch K-split scan; per-GPU lanes with sync-call coalescing)"; }

// CUDA ordinals this model should run on. A pinned device (aab_set_device on this thread, APRIL_B200_DEVICE, or LOCAL_RANK under
// torchrun: one process per GPU) gives one lane; otherwise every visible sm_100 device gets an engine and sessions are placed
// on the least-loaded one (the reference has no device notion: src/april_session.c:25-93).
static std::vector<int> pick_devices() {
    const std::vector<int> all = cuda_devices_sm100();       // index among B200s -> CUDA ordinal
    if (all.empty()) return all;
    int pin = tl_device;
    if (pin < 0) {
        const char* e = getenv("APRIL_B200_DEVICE");
        if (!e) e = getenv("LOCAL_RANK");
        if (e && *e) { const int v = atoi(e); if (v >= 0) pin = v % (int)all.size(); }
    }
    if (pin >= 0) return { all[(size_t)pin] };
    std::vector<int> out = all;
    if (const char* e = getenv("APRIL_B200_MAX_DEVICES")) { const int n = atoi(e); if (n > 0 && (size_t)n < out.size()) out.resize((size_t)n); }
    return out;
}

APRIL_EXPORT AprilASRModel aam_create_model(const char* model_path) {
    if (!g_initialized) { LOG_ERROR("aam: please make sure to call aam_api_init!"); return nullptr; }
    if (!model_path) return nullptr;
    LoadedModel lm;
    std::string err;
    if (!load_april_file(model_path, lm, err)) { LOG_ERROR("aam: failed to read model: %s", err.c_str()); return nullptr; }
    std::unique_ptr<AprilASRModel_i> m(new AprilASRModel_i());
    m->name = lm.name; m->description = lm.description; m->language = lm.language;
    m->params = lm.params;
    m->vocab.tokens = lm.params.tokens;
    m->vocab.blank_id = lm.params.blank_id;
    if (const char* e = getenv("APRIL_B200_COALESCE_US")) m->coalesce_us = atol(e);
    const std::vector<int> devs = pick_devices();
    if (devs.empty()) { LOG_ERROR("aam: no sm_100 (B200) device available; this library has no CPU fallback"); return nullptr; }
    std::vector<std::string> errs(devs.size());
    m->lanes.resize(devs.size());
    std::vector<std::thread> th;
    for (size_t i = 0; i < devs.size(); i++)
        th.emplace_back([&, i] {
            std::unique_ptr<Lane> L(new Lane());
            L->model = m.get(); L->index = (int)i;
            L->engine.reset(Engine::create(lm, devs[i], errs[i]));
            if (L->engine) m->lanes[i] = std::move(L);
        });
    for (auto& t : th) t.join();
    for (size_t i = 0; i < devs.size(); i++)
        if (!m->lanes[i]) { LOG_ERROR("aam: cannot create the B200 engine on cuda:%d: %s", devs[i], errs[i].c_str()); return nullptr; }
    LOG_INFO("aam: loaded model %s on %d GPU(s), first cuda:%d", m->name.c_str(), (int)devs.size(), devs[0]);
    return m.release();
}

APRIL_EXPORT const char* aam_get_name(AprilASRModel model) { return model->name.c_str(); }
APRIL_EXPORT const char* aam_get_description(AprilASRModel model) { return model->description.c_str(); }
APRIL_EXPORT const char* aam_get_language(AprilASRModel model) { return model->language.c_str(); }
APRIL_EXPORT size_t aam_get_sample_rate(AprilASRModel model) { return (size_t)model->params.sample_rate; }

APRIL_EXPORT void aam_free(AprilASRModel model) {
    if (!model) return;
    for (auto& L : model->lanes) {
        {
            std::lock_guard<std::mutex> lk(L->amu);
            L->stop = true;
        }
        L->acv.notify_all();
        if (L->batcher_started) L->batcher.join();
    }
    delete model;
}

APRIL_EXPORT AprilASRSession aas_create_session(AprilASRModel model, AprilConfig config) {
    if (!model) return nullptr;
    if (config.handler == nullptr) { LOG_ERROR("No handler provided! A handler is required, please provide a handler"); return nullptr; }
    std::unique_ptr<AprilASRSession_i> s(new AprilASRSession_i());
    s->model = model;
    s->sync = (config.flags & (APRIL_CONFIG_FLAG_ASYNC_RT_BIT | APRIL_CONFIG_FLAG_ASYNC_NO_RT_BIT)) == 0;
    s->handler = config.handler;
    s->userdata = config.userdata;
    s->shaper.reset(new TokenShaper(&model->vocab, config.handler, config.userdata));
    {
        std::lock_guard<std::mutex> lk(model->place_mu);     // least-loaded GPU; ties go to the lowest index
        Lane* best = model->lanes[0].get();
        for (auto& L : model->lanes) if (L->live_sessions.load() < best->live_sessions.load()) best = L.get();
        s->lane = best;
        best->live_sessions++;
    }
    Lane* L = s->lane;
    s->fr.slot = L->engine->alloc_slot();
    if (!s->sync) {
        s->ring.resize(kRingSamples);
        ensure_batcher(L);
        std::lock_guard<std::mutex> lk(L->amu);
        L->async_sessions.push_back(s.get());
    } else {
        std::lock_guard<std::mutex> lk(L->qmu);
        L->live_sync++;
    }
    return s.release();
}

APRIL_EXPORT float aas_realtime_get_speedup(AprilASRSession session) { (void)session; return 1.0f; }

APRIL_EXPORT void aas_free(AprilASRSession session) {
    if (!session) return;
    Lane* m = session->lane;
    if (!session->sync) {
        std::lock_guard<std::mutex> pl(m->pass_mu);       // wait for an in-flight batcher pass
        std::lock_guard<std::mutex> lk(m->amu);
        auto& v = m->async_sessions;
        for (size_t i = 0; i < v.size(); i++) if (v[i] == session) { v.erase(v.begin() + (long)i); break; }
    } else {
        std::lock_guard<std::mutex> lk(m->qmu);
        m->live_sync--;
        m->qcv.notify_all();                              // a leader waiting for "everybody" re-evaluates
    }
    m->engine->free_slot(session->fr.slot);
    m->live_sessions--;
    delete session;
}

APRIL_EXPORT void aas_feed_pcm16(AprilASRSession session, short* pcm16, size_t short_count) {
    if (!session) return;
    Lane* m = session->lane;
    if (session->sync) {
        session->was_flushed = false;                                   // april_session.c:510
        if (short_count == 0 || !pcm16) return;
        sync_submit(session, pcm16, short_count, false);
        return;
    }
    bool ok = true;
    {
        std::lock_guard<std::mutex> lk(m->amu);
        if (session->ring_count + short_count >= kRingSamples) ok = false;   // ap_push_audio, src/audio_provider.c:59-64 (>=): drop the whole chunk
        else { memcpy(session->ring.data() + session->ring_count, pcm16, short_count * sizeof(int16_t)); session->ring_count += short_count; }
    }
    m->acv.notify_one();
    if (!ok) session->handler(session->userdata, APRIL_RESULT_ERROR_CANT_KEEP_UP, 0, nullptr);   // april_session.c:485-492
}

APRIL_EXPORT void aas_flush(AprilASRSession session) {
    if (!session) return;
    Lane* m = session->lane;
    if (session->sync) {
        if (session->was_flushed) return;                               // april_session.c:548-550
        session->was_flushed = true;
        sync_submit(session, nullptr, 0, true);
        return;
    }
    {
        std::lock_guard<std::mutex> lk(m->amu);
        if (!session->flush_requested) { session->flush_requested = true; session->flush_at = session->ring_count; }
    }
    m->acv.notify_one();
}

// ---- additive batch API (april_b200.h) -----------------------------------------------------------
static int check_batch(AprilASRSession* sessions, size_t n) {
    if (!sessions || n == 0) return -1;
    std::set<AprilASRSession> seen;
    for (size_t i = 0; i < n; i++) {
        if (!sessions[i] || !sessions[i]->sync || sessions[i]->model != sessions[0]->model) return -1;
        if (!seen.insert(sessions[i]).second) return -1;
    }
    return 0;
}

APRIL_EXPORT int aab_feed_pcm16_batch(AprilASRSession* sessions, short** pcm16, const size_t* short_count, size_t n) {
    if (check_batch(sessions, n) != 0 || !short_count || !pcm16) return -1;
    for (size_t i = 0; i < n; i++) if (short_count[i] && !pcm16[i]) return -1;
    std::vector<AprilASRSession_i*> ss;
    std::vector<PassItem> items;
    for (size_t i = 0; i < n; i++) {
        sessions[i]->was_flushed = false;
        if (short_count[i] == 0) continue;
        ss.push_back(sessions[i]);
        PassItem it; it.fr = &sessions[i]->fr; it.cookie = sessions[i]; it.pcm = pcm16[i]; it.count = short_count[i];
        items.push_back(std::move(it));
    }
    run_items(ss, items);
    return 0;
}

APRIL_EXPORT int aab_flush_batch(AprilASRSession* sessions, size_t n) {
    if (check_batch(sessions, n) != 0) return -1;
    std::vector<AprilASRSession_i*> ss;
    std::vector<PassItem> items;
    for (size_t i = 0; i < n; i++) {
        if (sessions[i]->was_flushed) continue;
        sessions[i]->was_flushed = true;
        ss.push_back(sessions[i]);
        PassItem it; it.fr = &sessions[i]->fr; it.cookie = sessions[i]; it.flush = true;
        items.push_back(std::move(it));
    }
    run_items(ss, items);
    return 0;
}

APRIL_EXPORT int aab_session_device(AprilASRSession session) { return session ? session->lane->engine->device() : -1; }
APRIL_EXPORT int aab_pin_host_memory(void* ptr, size_t bytes) { return (ptr && bytes) ? host_pin(ptr, bytes) : -1; }
APRIL_EXPORT int aab_unpin_host_memory(void* ptr) { return ptr ? host_unpin(ptr) : -1; }

APRIL_EXPORT int aab_get_stats(AprilASRModel model, AprilB200Stats* out, int reset) {
    if (!model || !out) return -1;
    memset(out, 0, sizeof *out);
    for (auto& L : model->lanes) {                          // summed over the model's GPUs (times: the slowest lane)
        EngineStats s = L->engine->stats(reset != 0);
        out->encoder_steps += s.encoder_steps; out->joiner_passes += s.joiner_passes; out->decoder_runs += s.decoder_runs; out->frames += s.frames;
        out->kernel_launches += s.kernel_launches; out->h2d_bytes += s.h2d_bytes; out->d2h_bytes += s.d2h_bytes;
        out->gemm_flops += s.gemm_flops; out->gemm_launches += s.gemm_launches; out->passes += s.passes; out->pass_sessions += s.pass_sessions; out->lstm_gemm_flops += s.lstm_gemm_flops;
        out->lstm_gemm_ms = std::max(out->lstm_gemm_ms, s.lstm_gemm_ms); out->frontend_ms = std::max(out->frontend_ms, s.frontend_ms);
        out->device_ms = std::max(out->device_ms, s.device_ms); out->lstm_scan_ms = std::max(out->lstm_scan_ms, s.lstm_scan_ms); out->gemm_ms = std::max(out->gemm_ms, s.gemm_ms);
        out->compute_ms = std::max(out->compute_ms, s.compute_ms); out->greedy_ms = std::max(out->greedy_ms, s.greedy_ms);
    }
    return 0;
}

APRIL_EXPORT void aab_set_timing(AprilASRModel model, int enabled) { if (model) for (auto& L : model->lanes) L->engine->set_timing(enabled != 0); }

APRIL_EXPORT void aab_set_tap(AprilASRModel model, AprilB200Tap tap, void* userdata) {
    if (!model) return;
    for (auto& L : model->lanes) L->engine->set_tap(reinterpret_cast<TapFn>(tap), userdata);
}

// A ready-made handler for throughput measurements: userdata points at 4 uint64 counters
// {partial calls, final calls, silence calls, tokens in final results}.
APRIL_EXPORT void aab_counting_handler(void* userdata, AprilResultType type, size_t count, const AprilToken* tokens) {
    (void)tokens;
    uint64_t* c = static_cast<uint64_t*>(userdata);
    if (!c) return;
    if (type == APRIL_RESULT_RECOGNITION_PARTIAL) c[0]++;
    else if (type == APRIL_RESULT_RECOGNITION_FINAL) { c[1]++; c[3] += count; }
    else if (type == APRIL_RESULT_SILENCE) c[2]++;
}

APRIL_EXPORT int aab_debug_fbank(AprilASRModel model, const short* pcm16, size_t short_count, float* out_rows, int max_rows) {
    if (!model || !pcm16 || !out_rows) return -1;
    return model->lanes[0]->engine->debug_fbank(pcm16, short_count, out_rows, max_rows);
}

APRIL_EXPORT int aab_debug_gemm(AprilASRModel model, int tier, int M, int N, int K, const float* A, const float* W, const float* bias, float* out) {
    if (!model || !A || !W || !bias || !out) return -1;
    return model->lanes[0]->engine->debug_gemm(tier, M, N, K, A, W, bias, out);
}

APRIL_EXPORT int aab_debug_replay(const char* model_path, int n_steps, int n_records, const int* max_idx, const float* max_val, const float* blank_val,
                                  int flush, AprilRecognitionResultHandler handler, void* userdata) {
    if (!model_path || !handler || n_steps < 0 || n_records < 0 || (n_records && (!max_idx || !max_val || !blank_val))) return -1;
    LoadedModel lm;
    std::string err;
    if (!load_april_file(model_path, lm, err)) { LOG_ERROR("aab_debug_replay: failed to read model: %s", err.c_str()); return -1; }
    Vocab vocab;
    vocab.tokens = lm.params.tokens; vocab.blank_id = lm.params.blank_id;
    TokenShaper shaper(&vocab, handler, userdata);
    const size_t stride = (size_t)lm.params.segment_step * (size_t)lm.params.frame_shift_ms;   // fbank.c:359-361
    int r = 0;
    for (int t = 0; t < n_steps; t++) {
        shaper.begin_step(stride);
        for (int p = 0; p < 3; p++) {                                  // april_session.c:449-454
            if (r >= n_records) return -1;
            const bool blank = shaper.on_joiner_pass(p == 0 ? 1.0f : 0.0f, max_idx[r], max_val[r], blank_val[r]);
            r++;
            if (blank) break;
        }
    }
    if (flush) shaper.on_flush_end();
    return r;
}

APRIL_EXPORT int aab_debug_model_digest(const char* model_path, uint64_t* out16) {
    if (!model_path || !out16) return -1;
    LoadedModel lm;
    std::string err;
    if (!load_april_file(model_path, lm, err)) { LOG_INFO("aab_debug_model_digest: %s", err.c_str()); return -1; }
    model_digest(lm, out16);
    return 0;
}

APRIL_EXPORT int aab_debug_plan(const char* model_path, int n_calls, const size_t* counts, const int* flush, int* out_steps, int* out_frames) {
    if (!model_path || n_calls < 0 || (n_calls && (!counts || !flush || !out_steps || !out_frames))) return -1;
    LoadedModel lm;
    std::string err;
    if (!load_april_file(model_path, lm, err)) { LOG_ERROR("aab_debug_plan: failed to read model: %s", err.c_str()); return -1; }
    const Params& prm = lm.params;
    int N = prm.frame_length_ms * prm.sample_rate / 1000;              // engine.cu: same derivation (fbank.c:136-141)
    const int shift = prm.frame_shift_ms * prm.sample_rate / 1000;
    if (prm.round_pow2) { int p2 = 1; while (p2 < N) p2 <<= 1; N = p2; }
    PlanState st;
    for (int i = 0; i < n_calls; i++) {
        const PassPlan pl = plan_pass(st, counts[i], flush[i] != 0, prm.segment_size, prm.segment_step, N, shift, 6400);   // april_session.c:500,555-556
        out_steps[i] = pl.nsteps; out_frames[i] = pl.n_frames;
    }
    return 0;
}

}  // extern "C"
