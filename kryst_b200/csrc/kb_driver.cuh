// kb_driver.cuh — host loop shared by the Krylov drivers: capture B iterations into a CUDA graph once,
// replay it, and poll the device control block's `done` flag once per replay.  Kernels issued after
// `done` is set exit immediately, so over-issued iterations change nothing.
#pragma once
#include <cstddef>
#include <vector>
#include "kb_objects.h"

struct KbGraphCache {
    cudaGraphExec_t exec = nullptr;
    uint64_t key = 0;
    int iters = 0;
    uint64_t launches = 0;
    void reset() {
        if (exec) cudaGraphExecDestroy(exec);
        exec = nullptr; key = 0; iters = 0; launches = 0;
    }
};

// launch_iter(): enqueue one iteration's kernels on c->stream, return kb_status.
// units_cap: upper bound on useful iterations (max_iters); the loop stops when done or after the cap.
// Slow mode (KB_FLAG_MONITOR): one unit per batch, no graph; after every unit the control block is read back
// and the host observer is called for every residual-history entry it has not seen yet, in order
// (pcg.rs:143-145,196-198: monitor(0, r0), monitor(i+1, res); fgmres.rs:286-289: monitor(total_iters, res)).
struct KbMonitor {
    kb_monitor_fn fn = nullptr; void* user = nullptr;
    const double* d_hist = nullptr; uint64_t cap = 0, seen = 0, index_offset = 0;
};
static inline int kb_monitor_deliver(kb_ctx_s* c, KbMonitor* m, const KbCtl* h_ctl) {
    const uint64_t have = h_ctl->hist_len < m->cap ? h_ctl->hist_len : m->cap;
    if (have <= m->seen) return KB_OK;
    std::vector<double> buf((size_t)(have - m->seen));
    KB_CUDA(cudaMemcpyAsync(buf.data(), m->d_hist + m->seen, buf.size() * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    KB_CUDA(cudaStreamSynchronize(c->stream));
    for (size_t k = 0; k < buf.size(); ++k) m->fn(m->seen + k + m->index_offset, buf[k], m->user);
    m->seen = have;
    return KB_OK;
}

template <class F>
static int kb_run_iterations(kb_ctx_s* c, KbGraphCache* gc, uint64_t key, int B, uint64_t units_cap, bool use_graph,
                             KbCtl* d_ctl, KbCtl* h_ctl, F&& launch_iter, KbMonitor* mon = nullptr) {
    if (mon && mon->fn) {
        for (uint64_t u = 0;; ++u) {
            if (cudaMemcpyAsync(h_ctl, d_ctl, offsetof(KbCtl, h), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
                cudaStreamSynchronize(c->stream) != cudaSuccess) { kb_set_error("device error during Krylov iterations: %s", cudaGetErrorString(cudaGetLastError())); return KB_SOLVE_ERROR; }
            KB_TRY(kb_monitor_deliver(c, mon, h_ctl));
            if (h_ctl->done || u >= units_cap) break;
            KB_TRY(launch_iter());
        }
        return KB_OK;
    }
    if (units_cap == 0) return KB_OK;
    if (B < 1) B = 1;
    if ((uint64_t)B > units_cap) B = (int)units_cap;
    if (use_graph && (!gc->exec || gc->key != key || gc->iters != B)) {
        gc->reset();
        cudaGraph_t g = nullptr;
        int st = KB_OK;
        c->capturing = true; c->captured_launches = 0;
        cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
        if (e == cudaSuccess) {
            for (int k = 0; k < B && st == KB_OK; ++k) st = launch_iter();
            cudaError_t e2 = cudaStreamEndCapture(c->stream, &g);
            if (e2 != cudaSuccess) e = e2;
        }
        c->capturing = false;
        if (st != KB_OK) { if (g) cudaGraphDestroy(g); return st; }
        if (e != cudaSuccess || !g) { kb_set_error("CUDA graph capture failed: %s", cudaGetErrorString(e)); return KB_SOLVE_ERROR; }
        e = cudaGraphInstantiate(&gc->exec, g, 0);
        cudaGraphDestroy(g);
        if (e != cudaSuccess) { gc->exec = nullptr; kb_set_error("CUDA graph instantiate failed: %s", cudaGetErrorString(e)); return KB_SOLVE_ERROR; }
        gc->key = key; gc->iters = B; gc->launches = c->captured_launches;
    }
    // Software-pipelined polling: replay k+1 is enqueued before the host looks at the `done` flag copied after
    // replay k, so the device never idles on the host round trip (the speculative replay is a string of no-ops
    // once `done` is set).
    uint64_t issued = 0;
    cudaEvent_t ev[2] = {nullptr, nullptr};
    for (int k = 0; k < 2; ++k) if (cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming) != cudaSuccess) { kb_set_error("event create failed"); return KB_SOLVE_ERROR; }
    int* flags = reinterpret_cast<int*>(reinterpret_cast<char*>(h_ctl) + sizeof(KbCtl) - 2 * sizeof(int));   // tail of the pinned mirror: unused by copies
    int st = KB_OK;
    int slot = 0;
    bool pending[2] = {false, false};
    auto enqueue = [&](int sl) -> int {
        if (use_graph) {
            cudaError_t e = cudaGraphLaunch(gc->exec, c->stream);
            if (e != cudaSuccess) { kb_set_error("CUDA graph launch failed: %s", cudaGetErrorString(e)); return KB_SOLVE_ERROR; }
            c->launches += gc->launches;
        } else {
            for (int k = 0; k < B; ++k) KB_TRY(launch_iter());
        }
        issued += (uint64_t)B;
        if (cudaMemcpyAsync(&flags[sl], &d_ctl->done, sizeof(int), cudaMemcpyDeviceToHost, c->stream) != cudaSuccess ||
            cudaEventRecord(ev[sl], c->stream) != cudaSuccess) { kb_set_error("poll enqueue failed"); return KB_SOLVE_ERROR; }
        pending[sl] = true;
        return KB_OK;
    };
    st = enqueue(slot);
    while (st == KB_OK) {
        const bool more = issued < units_cap + (uint64_t)B;
        if (more) { st = enqueue(slot ^ 1); if (st != KB_OK) break; }
        if (cudaEventSynchronize(ev[slot]) != cudaSuccess) { kb_set_error("device error during Krylov iterations: %s", cudaGetErrorString(cudaGetLastError())); st = KB_SOLVE_ERROR; break; }
        pending[slot] = false;
        if (flags[slot] || !more) break;
        slot ^= 1;
    }
    if (cudaStreamSynchronize(c->stream) != cudaSuccess && st == KB_OK) { kb_set_error("device error during Krylov iterations: %s", cudaGetErrorString(cudaGetLastError())); st = KB_SOLVE_ERROR; }
    for (int k = 0; k < 2; ++k) cudaEventDestroy(ev[k]);
    return st;
}

// iterations per replay so that one replay is ~4 ms of work
static inline int kb_batch_size(double bytes_per_iter, int kernels_per_iter) {
    double t = bytes_per_iter / 6.0e12 + 3.0e-6 * kernels_per_iter;
    double b = 4.0e-3 / t;
    return (int)(b < 4.0 ? 4.0 : (b > 64.0 ? 64.0 : b));
}
