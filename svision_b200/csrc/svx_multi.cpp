// Several GPUs behind ONE handle, in ONE process (include/svx.h, "multi-device").
//
// The reference runs Step 2 from a single `SVision` process (SVision:296-341); its replacement
// `Predict.run` makes one `classify` call per chromosome.  svx_multi gives that
// single caller every GPU of the box: one svx_handle and one host thread per device; the rows of a
// call are cut into chunks that the device threads take from a shared counter, so a GPU that runs
// slower (the B200s of one box differ by 10-20 % in sustained clock under the power cap) simply
// takes fewer chunks.  Sites are independent: a site's result does not depend on
// which device or chunk it lands in, and results are written straight to their place in the
// caller's arrays, so file order is preserved without any gather step.
//
// This file is host plumbing on top of the public C-ABI (svx_create / svx_classify); it has no
// arithmetic of its own.
#include <atomic>
#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/svx.h"

namespace svx {
int fail(int code, const std::string& msg);
}

struct svx_multi {
    std::vector<int> devices;
    std::vector<svx_handle*> handles;
    int64_t max_batch = 0;

    std::vector<std::thread> threads;
    std::mutex mu;
    std::condition_variable cv_job, cv_done;
    uint64_t job_id = 0;
    bool stop = false;
    int pending = 0;

    // the job in flight
    const int32_t* rows = nullptr;
    int64_t n = 0, chunk = 0, n_chunks = 0;
    int32_t* labels = nullptr;
    float* probs = nullptr;
    std::atomic<int64_t> next_chunk{0};

    std::vector<int> rc;
    std::vector<std::string> err;
    std::vector<int64_t> sites_done;          // per device, of the last call
};

namespace {

void worker(svx_multi* m, int idx, const svx_weights* weights, int precision) {
    // the handle is created by the thread that uses it (one CUDA context per device, selected per call)
    svx_handle* h = nullptr;
    int rc = svx_create(weights, m->devices[idx], m->max_batch, precision, &h);
    {
        std::lock_guard<std::mutex> lock(m->mu);
        m->handles[idx] = h;
        m->rc[idx] = rc;
        if (rc != SVX_OK) m->err[idx] = svx_last_error();
        --m->pending;
    }
    m->cv_done.notify_all();
    uint64_t seen = 0;
    for (;;) {
        {
            std::unique_lock<std::mutex> lock(m->mu);
            m->cv_job.wait(lock, [&] { return m->stop || m->job_id != seen; });
            if (m->stop) break;
            seen = m->job_id;
        }
        int job_rc = SVX_OK;
        std::string job_err;
        int64_t done = 0;
        if (h) {
            for (;;) {
                const int64_t c = m->next_chunk.fetch_add(1);
                if (c >= m->n_chunks) break;
                const int64_t off = c * m->chunk;
                const int64_t cnt = m->n - off < m->chunk ? m->n - off : m->chunk;
                job_rc = svx_classify(h, m->rows + off * SVX_ROW_FIELDS, cnt, m->labels + off,
                                      m->probs + off * SVX_NUM_CLASSES);
                if (job_rc != SVX_OK) { job_err = svx_last_error(); break; }
                done += cnt;
            }
        } else {
            job_rc = SVX_ERR_INVALID;
            job_err = "device handle missing";
        }
        {
            std::lock_guard<std::mutex> lock(m->mu);
            m->rc[idx] = job_rc;
            m->err[idx] = job_err;
            m->sites_done[idx] = done;
            --m->pending;
        }
        m->cv_done.notify_all();
    }
    if (h) svx_destroy(h);
}

void shutdown(svx_multi* m) {
    {
        std::lock_guard<std::mutex> lock(m->mu);
        m->stop = true;
    }
    m->cv_job.notify_all();
    for (std::thread& t : m->threads)
        if (t.joinable()) t.join();
}

}  // namespace

extern "C" {

int svx_multi_create(const svx_weights* weights, const int* devices, int ndev, int64_t max_batch,
                     int precision, svx_multi** out) {
    if (!out) return svx::fail(SVX_ERR_INVALID, "svx_multi_create: out is NULL");
    *out = nullptr;
    if (!weights || !devices || ndev < 1 || ndev > 64)
        return svx::fail(SVX_ERR_INVALID, "svx_multi_create: need weights and 1..64 devices");
    for (int i = 0; i < ndev; ++i)
        for (int j = 0; j < i; ++j)
            if (devices[i] == devices[j]) return svx::fail(SVX_ERR_INVALID, "svx_multi_create: duplicate device");
    svx_multi* m = new svx_multi();
    m->devices.assign(devices, devices + ndev);
    m->handles.assign(ndev, nullptr);
    m->rc.assign(ndev, SVX_OK);
    m->err.assign(ndev, std::string());
    m->sites_done.assign(ndev, 0);
    m->max_batch = max_batch;
    m->pending = ndev;
    // every device repacks and uploads its own copy of the weights: in parallel
    for (int i = 0; i < ndev; ++i) m->threads.emplace_back(worker, m, i, weights, precision);
    {
        std::unique_lock<std::mutex> lock(m->mu);
        m->cv_done.wait(lock, [&] { return m->pending == 0; });
    }
    for (int i = 0; i < ndev; ++i) {
        if (m->rc[i] != SVX_OK) {
            const int rc = m->rc[i];
            const std::string msg = "svx_multi_create: device " + std::to_string(m->devices[i]) + ": " + m->err[i];
            shutdown(m);
            delete m;
            return svx::fail(rc, msg);
        }
    }
    *out = m;
    return SVX_OK;
}

void svx_multi_destroy(svx_multi* m) {
    if (!m) return;
    shutdown(m);
    delete m;
}

int svx_multi_device_count(const svx_multi* m) { return m ? (int)m->devices.size() : 0; }

int svx_multi_classify(svx_multi* m, const int32_t* rows_host, int64_t n, int32_t* labels_host,
                       float* probs_host) {
    if (!m) return svx::fail(SVX_ERR_INVALID, "svx_multi_classify: NULL handle");
    if (n < 0 || (n > 0 && (!rows_host || !labels_host || !probs_host)))
        return svx::fail(SVX_ERR_INVALID, "svx_multi_classify: bad arguments");
    const int ndev = (int)m->devices.size();
    for (int i = 0; i < ndev; ++i) m->sites_done[i] = 0;
    if (n == 0) return SVX_OK;
    // Chunks: k equal chunks per device, k the fewest that fit the micro-batch (an equal split, one
    // pass each, when the call fits).  With devices of equal speed every device takes k chunks; a slower
    // one takes fewer, because the chunks are handed out through a shared counter.
    const int64_t rounds = (n + (int64_t)ndev * m->max_batch - 1) / ((int64_t)ndev * m->max_batch);
    const int64_t chunk = (n + ndev * rounds - 1) / (ndev * rounds);
    {
        std::lock_guard<std::mutex> lock(m->mu);
        m->rows = rows_host;
        m->labels = labels_host;
        m->probs = probs_host;
        m->n = n;
        m->chunk = chunk;
        m->n_chunks = (n + chunk - 1) / chunk;
        m->next_chunk.store(0);
        m->pending = ndev;
        ++m->job_id;
    }
    m->cv_job.notify_all();
    {
        std::unique_lock<std::mutex> lock(m->mu);
        m->cv_done.wait(lock, [&] { return m->pending == 0; });
    }
    for (int i = 0; i < ndev; ++i)
        if (m->rc[i] != SVX_OK)
            return svx::fail(m->rc[i], "svx_multi_classify: device " + std::to_string(m->devices[i]) + ": " + m->err[i]);
    return SVX_OK;
}

int svx_multi_last_split(const svx_multi* m, int64_t* sites_per_device) {
    if (!m || !sites_per_device) return svx::fail(SVX_ERR_INVALID, "svx_multi_last_split: bad arguments");
    for (size_t i = 0; i < m->devices.size(); ++i) sites_per_device[i] = m->sites_done[i];
    return SVX_OK;
}

}  // extern "C"
