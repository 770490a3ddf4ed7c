// multi.cu -- one context over several devices (twg_create_multi): the index split of SURVEY.md 8(e) behind the C ABI.
//
// The reference's callers hand one batch to one call (InoutFiltering.cpp:40-52: all tet centroids; LocalOperations.cpp:1046-1109:
// the faces of a candidate; VertexSmoother.cpp:216-241: all tets). A multi-device context keeps that call shape: the surface /
// winding handle is REPLICATED on every device, a host entry point splits its batch [0, n) into contiguous index ranges
// [k n / G, (k+1) n / G), and device k's range is staged and evaluated by device k's own host thread through device k's
// ordinary one-device context -- H2D, kernels and D2H of the G ranges run concurrently, results land directly in the
// caller's buffer (no collective: the ranges are disjoint slices of one host array). Small batches stay on device 0: waking
// G threads costs more than a few thousand queries.
#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include "common.cuh"

struct twg_worker {
    std::thread th;
    std::mutex m;
    std::condition_variable cv;
    std::function<int()> job;
    bool has_job = false, done = false, quit = false;
    int rc = 0;
};

namespace {

void worker_main(twg_worker* w, int device) {
    cudaSetDevice(device);
    std::unique_lock<std::mutex> lk(w->m);
    for (;;) {
        w->cv.wait(lk, [&] { return w->has_job || w->quit; });
        if (w->quit) return;
        std::function<int()> job = std::move(w->job);
        w->has_job = false;
        lk.unlock();
        const int rc = job();
        lk.lock();
        w->rc = rc;
        w->done = true;
        w->cv.notify_all();
    }
}

}  // namespace

// fn(k, child_k) on every device's own thread, concurrently; returns the first non-zero code (its message is copied to c->err)
int twg_multi_run(twg_ctx* c, const std::function<int(int, twg_ctx*)>& fn) {
    const int G = (int)c->children.size();
    for (int k = 0; k < G; ++k) {
        twg_worker* w = c->workers[k];
        twg_ctx* child = c->children[k];
        std::lock_guard<std::mutex> lk(w->m);
        w->job = [&fn, k, child]() { return fn(k, child); };
        w->done = false;
        w->has_job = true;
        w->cv.notify_all();
    }
    int rc = 0;
    for (int k = 0; k < G; ++k) {
        twg_worker* w = c->workers[k];
        std::unique_lock<std::mutex> lk(w->m);
        w->cv.wait(lk, [&] { return w->done; });
        if (w->rc != 0 && rc == 0) {
            rc = w->rc;
            snprintf(c->err, sizeof(c->err), "device %d: %s", c->children[k]->device, c->children[k]->err);
        }
    }
    return rc;
}

// result of an entry point forwarded to device 0 of a multi-device context
int twg_forward0(twg_ctx* c, int rc) {
    if (rc != 0) snprintf(c->err, sizeof(c->err), "device %d: %s", c->children[0]->device, c->children[0]->err);
    return rc;
}

void twg_multi_teardown(twg_ctx* c) {
    for (twg_worker* w : c->workers) {
        {
            std::lock_guard<std::mutex> lk(w->m);
            w->quit = true;
            w->cv.notify_all();
        }
        if (w->th.joinable()) w->th.join();
        delete w;
    }
    c->workers.clear();
    for (twg_ctx* k : c->children) twg_destroy(k);
    c->children.clear();
}

extern "C" int twg_create_multi(twg_ctx** out, const int* device_ids, int n_devices) {
    if (!out) return TWG_ERR_INVALID_ARG;
    *out = nullptr;
    if (!device_ids || n_devices < 1 || n_devices > 64) return TWG_ERR_INVALID_ARG;
    // ids may repeat: several worker contexts on one device (own streams, own replicas) -- how the index split is tested on a
    // one-GPU box; a production caller names each device once
    if (n_devices == 1) return twg_create(out, device_ids[0]);
    twg_ctx* c = new twg_ctx;
    c->device = device_ids[0];
    for (int k = 0; k < n_devices; ++k) {
        twg_ctx* child = nullptr;
        const int rc = twg_create(&child, device_ids[k]);
        if (rc != 0) {
            twg_multi_teardown(c);
            delete c;
            return rc;
        }
        child->parent = c;
        c->children.push_back(child);
    }
    c->opt = c->children[0]->opt;
    c->sm_count = c->children[0]->sm_count;
    for (int k = 0; k < n_devices; ++k) {
        twg_worker* w = new twg_worker;
        c->workers.push_back(w);
        w->th = std::thread(worker_main, w, device_ids[k]);
    }
    *out = c;
    return 0;
}
