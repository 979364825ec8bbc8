// nbx_graph.inl -- shared by nbx_multi.cu and nbx_slab.cu: run nsteps of a fixed launch sequence, replaying a CUDA graph
#pragma once
#include "nbx_internal.cuh"

namespace nbx {

// nsteps of a fixed launch sequence.  The buffer pointers a step sees rotate on the host with a fixed PERIOD (2: the acc /
// acc_old swap; 6 for list-less slabs, whose per-step compaction also rotates a third acceleration buffer and swaps the
// position buffers).  One period is captured once -- after two eager steps, which let every buffer and kernel attribute
// settle outside the capture -- and kept for later calls with the same dt; a call then is nsteps / period graph launches
// (eager steps first until the rotation is back where the capture started).
template <class F>
static int steps_graphed(nbx_ctx *c, int kind, double dt, int64_t nsteps, bool allow_graph, int period, F one_step)
{
    int64_t s = 0;
    const bool graphable = allow_graph && c->opt_graph && !c->timing && nsteps >= 2 + period && c->stream != nullptr &&
                           c->stream != cudaStreamLegacy && c->stream != cudaStreamPerThread;
    if (graphable) {
        if (!(c->mg_exec && c->mg_kind == kind && c->mg_dt == dt)) {
            graph_drop(c);
            for (int w = 0; w < 2; ++w, ++s) NBX_TRY(one_step());
            if (c->opt_cond_nodes && !c->cond_fail && !c->aux_stream &&
                cudaStreamCreateWithFlags(&c->aux_stream, cudaStreamNonBlocking) != cudaSuccess) { c->aux_stream = nullptr; cudaGetLastError(); }
            cudaGraph_t graph = nullptr;
            cudaError_t e = cudaSuccess;
            for (int attempt = 0; attempt < 2; ++attempt) { // second attempt: plain capture, should the IF nodes be refused
                cudaStream_t main_stream = c->stream;
                c->cond_capture = c->opt_cond_nodes && !c->cond_fail && c->aux_stream != nullptr;
                const bool with_nodes = c->cond_capture;
                double *acc0 = c->acc, *acc_old0 = c->acc_old;
                e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal);
                if (e != cudaSuccess) { c->cond_capture = false; break; }
                int rc = NBX_OK;
                for (int k = 0; k < period && rc == NBX_OK; ++k) rc = one_step();
                c->cond_capture = false;
                c->stream = main_stream;
                e = cudaStreamEndCapture(c->stream, &graph);
                if (rc == NBX_OK && e == cudaSuccess) e = cudaGraphInstantiate(&c->mg_exec, graph, 0);
                if (rc == NBX_OK && e == cudaSuccess) break;
                if (graph) { cudaGraphDestroy(graph); graph = nullptr; }
                c->mg_exec = nullptr;
                cudaGetLastError();
                c->acc = acc0; c->acc_old = acc_old0; c->forces_done = false; // nothing of the failed capture ran
                if (!with_nodes) { if (rc != NBX_OK) return rc; break; }
                c->cond_fail = true;
                e = cudaSuccess;
            }
            if (graph) cudaGraphDestroy(graph);
            if (e != cudaSuccess) { c->mg_exec = nullptr; return cuda_fail(c, e, "CUDA graph of the distributed velocity-Verlet step"); }
            if (c->mg_exec) { c->mg_kind = kind; c->mg_dt = dt; c->mg_acc0 = c->acc; c->mg_pos0 = c->pos; }
        }
        if (c->mg_exec) {
            auto aligned = [&]() { return c->acc == c->mg_acc0 && c->pos == c->mg_pos0; };
            for (int k = 0; k < period && !aligned() && s < nsteps; ++k, ++s) NBX_TRY(one_step()); // back to the captured rotation
            if (aligned())
                for (; s + period <= nsteps; s += period) NBX_CUDA(c, cudaGraphLaunch(c->mg_exec, c->stream));
        }
    }
    for (; s < nsteps; ++s) NBX_TRY(one_step());
    return NBX_OK;
}

} // namespace nbx
