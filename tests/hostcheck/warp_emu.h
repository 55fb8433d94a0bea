// warp_emu.h — TEST INFRASTRUCTURE: runs a warp-cooperative kernel body (cgal_b200/csrc/closest_pair.cuh) on the CPU.
// The 32 lanes are ucontext fibers scheduled round-robin on ONE thread: a lane runs until it reaches a warp collective (ballot, shuffle,
// match, reduce, syncwarp), parks there, and is resumed once all 32 lanes have arrived.  Between collectives the lanes therefore run one
// after the other — a legal SIMT schedule, and a hostile one: code that silently relies on lock-step execution between two collectives
// (a missing __syncwarp) misbehaves here deterministically.
#pragma once
#include <ucontext.h>
#include <cstdint>
#include <cstdlib>
#include <functional>
#include <vector>

namespace b2 { struct WarpCtx; }

namespace b2emu {

struct Warp {
    ucontext_t sched;
    ucontext_t lane_ctx[32];
    std::vector<char> stacks[32];
    uint64_t xchg[2][32];
    unsigned arrived = 0, generation = 0;
    bool finished[32];
    int current = -1;
    std::function<void(unsigned)> body;
};

inline Warp*& active() { static thread_local Warp* w = nullptr; return w; }

inline void park(Warp* W, unsigned lane) { swapcontext(&W->lane_ctx[lane], &W->sched); }

}  // namespace b2emu

#include "../../cgal_b200/csrc/closest_pair.cuh"

namespace b2emu {

// every lane contributes v; returns all 32 values.  Double-buffered by collective parity: a lane can be at most one collective ahead.
inline void exchange(b2::WarpCtx& w, uint64_t v, uint64_t* all) {
    Warp* W = (Warp*)w.emu;
    uint64_t* buf = W->xchg[w.phase & 1u];
    buf[w.lane] = v;
    const unsigned gen = W->generation;
    ++W->arrived;
    while (W->generation == gen) park(W, w.lane);
    for (int i = 0; i < 32; ++i) all[i] = buf[i];
    ++w.phase;
}
inline void sync(b2::WarpCtx& w) { uint64_t all[32]; exchange(w, 0, all); }

inline void trampoline(unsigned lo, unsigned hi) {
    Warp* W = (Warp*)(((uint64_t)hi << 32) | lo);
    const unsigned lane = (unsigned)W->current;
    W->body(lane);
    W->finished[lane] = true;
    swapcontext(&W->lane_ctx[lane], &W->sched);
}

// runs body(lane) for 32 lanes to completion; returns false on a deadlock (some lanes finished while others wait at a collective)
inline bool run_warp(Warp& W, std::function<void(unsigned)> body) {
    W.body = std::move(body);
    W.arrived = 0; W.generation = 0;
    for (int l = 0; l < 32; ++l) {
        W.finished[l] = false;
        W.stacks[l].assign(1 << 20, 0);
        getcontext(&W.lane_ctx[l]);
        W.lane_ctx[l].uc_stack.ss_sp = W.stacks[l].data();
        W.lane_ctx[l].uc_stack.ss_size = W.stacks[l].size();
        W.lane_ctx[l].uc_link = &W.sched;
        const uint64_t p = (uint64_t)&W;
        makecontext(&W.lane_ctx[l], (void (*)())trampoline, 2, (unsigned)(p & 0xFFFFFFFFu), (unsigned)(p >> 32));
    }
    for (;;) {
        for (int l = 0; l < 32; ++l) {
            if (W.finished[l]) continue;
            W.current = l;
            swapcontext(&W.sched, &W.lane_ctx[l]);
        }
        unsigned alive = 0;
        for (int l = 0; l < 32; ++l) alive += W.finished[l] ? 0u : 1u;
        if (alive == 0) return true;
        if (W.arrived == 32) { W.arrived = 0; ++W.generation; }
        else if (W.arrived == alive) return false;   // every live lane waits at a collective some lanes will never reach: non-uniform control flow
    }
}

}  // namespace b2emu
