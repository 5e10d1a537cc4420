// Pass planning: the integer logic of the reference's feature ring and flush, with no device code in it.
//   fbank_accept_waveform's frame count and leftover   /root/reference/src/fbank.c:198-226,299-300
//   fbank_pull_segments' availability counters          src/fbank.c:327-349
//   fbank_flush's padding loop                          src/fbank.c:308-325
//   _aas_flush: flush, 6400 zero samples, flush again   src/april_session.c:547-559
// Features are a pure function of the absolute sample position, so one pass can plan a whole feed (or flush) call: how many
// rows are carried in, how many pad rows precede / follow the frames computed from this call's samples, how many encoder
// steps run, which rows stay for the next call, and how many samples become the next call's leftover.
#pragma once
#include <cstddef>

namespace april {

struct PlanState {                 // per stream, across calls
    long long avail = 0;           // temp_segment_avail: feature rows in the ring not yet consumed
    long long avail_f = 0;         // temp_segment_avail_f: the signed counter fbank_flush tests
    size_t tail = 0;               // prev_leftover: samples kept from the previous call (from the next frame's start on)
};

struct PassPlan {
    int carry = 0;                 // rows carried in from earlier calls
    int pad_before = 0;            // pad rows emitted by the first flush loop (before this call's frames)
    int n_frames = 0;              // frames computed from [tail | new samples or flush zeros]
    int pad_after = 0;             // pad rows emitted by the second flush loop
    int n_new = 0;                 // pad_before + n_frames + pad_after
    int nsteps = 0;                // encoder steps this call runs
    int keep_first = 0;            // first row (of carry + new rows) that stays for the next call
    int n_keep = 0;                // rows that stay
    size_t n_buf = 0;              // samples staged for this call: tail + n_in
    size_t n_in = 0;               // new samples (flush: the zero samples)
};

inline PassPlan plan_pass(PlanState& st, size_t count, bool flush, int SEG, int STEP, int N, int shift, size_t flush_zeros) {
    PassPlan p;
    p.carry = (int)st.avail;
    long long avail = st.avail, avail_f = st.avail_f;
    int steps = 0;
    auto drain = [&]() { while (avail >= SEG) { steps++; avail -= STEP; avail_f -= STEP; } };
    auto flush_loop = [&](int& pads) {                                   // while (fbank_flush()) aas_infer();
        while (avail_f >= -(long long)(SEG * 3)) { while (avail < SEG) { pads++; avail++; } drain(); }
    };
    p.n_in = flush ? flush_zeros : count;
    if (flush) flush_loop(p.pad_before);
    p.n_buf = st.tail + p.n_in;
    p.n_frames = p.n_buf >= (size_t)N ? (int)((p.n_buf - (size_t)N) / (size_t)shift) + 1 : 0;
    avail += p.n_frames;
    if (p.n_frames > 0) avail_f = avail;                                 // fbank.c:299-300
    if (flush) flush_loop(p.pad_after); else drain();
    p.n_new = p.pad_before + p.n_frames + p.pad_after;
    p.nsteps = steps;
    p.keep_first = steps * STEP;
    p.n_keep = p.carry + p.n_new - steps * STEP;                         // == avail
    size_t consumed = (size_t)p.n_frames * (size_t)shift;                // the next frame starts here (fbank.c:198-224)
    if (consumed > p.n_buf) consumed = p.n_buf;
    st.avail = avail; st.avail_f = avail_f; st.tail = p.n_buf - consumed;
    return p;
}

}  // namespace april
